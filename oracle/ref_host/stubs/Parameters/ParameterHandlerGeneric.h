// Stand-in for Parameters/ParameterHandlerGeneric.h (needs ROOT matrices and yaml-cpp): only the three getters
// Splines/SplineStructs.h's knot-capping helpers mention.  TEST INFRASTRUCTURE (oracle/ref_host).
#pragma once
#include <string>
#include "Parameters/ParameterStructs.h"
class ParameterHandlerGeneric {
 public:
  std::string GetParFancyName(int) const { return std::string(); }
  double GetParSplineKnotUpperBound(int) const { return M3::DefSplineKnotUpBound; }
  double GetParSplineKnotLowerBound(int) const { return M3::DefSplineKnotLowBound; }
};
