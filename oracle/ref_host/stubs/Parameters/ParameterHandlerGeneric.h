// Stand-in for Parameters/ParameterHandlerGeneric.h (the real one pulls in ROOT's matrix classes, yaml-cpp parsing
// and the whole ParameterHandlerBase): only the members the reference sources compiled through oracle/ref_host
// mention, declared with the reference's signatures (Parameters/ParameterHandlerGeneric.h:40-132,
// Parameters/ParameterHandlerBase.h:205).  Everything returns "nothing": the harness never routes a sample through
// its configuration-time set-up; it fills the event structures directly.  TEST INFRASTRUCTURE (oracle/ref_host).
#pragma once
#include <string>
#include <vector>
#include "Parameters/ParameterStructs.h"
#include "yaml-cpp/yaml.h"
class ParameterHandlerGeneric {
 public:
  std::string GetParFancyName(int) const { return std::string(); }
  double GetParSplineKnotUpperBound(int) const { return M3::DefSplineKnotUpBound; }
  double GetParSplineKnotLowerBound(int) const { return M3::DefSplineKnotLowBound; }
  const std::vector<SplineInterpolation> GetSplineInterpolationFromSampleName(const std::string&) { return {}; }
  const std::vector<int> GetGlobalSystIndexFromSampleName(const std::string&, const SystType) { return {}; }
  int GetNumParamsFromSampleName(const std::string&, const SystType) { return 0; }
  const std::vector<std::string> GetParsNamesFromSampleName(const std::string&, const SystType) { return {}; }
  const std::vector<int> GetParsIndexFromSampleName(const std::string&, const SystType) { return {}; }
  const std::vector<std::string> GetSplineParsNamesFromSampleName(const std::string&) { return {}; }
  const std::vector<std::vector<int>> GetSplineModeVecFromSampleName(const std::string&) { return {}; }
  const std::vector<NormParameter> GetNormParsFromSampleName(const std::string&) const { return {}; }
  const std::vector<FunctionalParameter> GetFunctionalParametersFromSampleName(const std::string&) const { return {}; }
  std::vector<const double*> GetOscParsFromSampleName(const std::string&) { return {}; }
  const double* RetPointer(const int) { return nullptr; }
  YAML::Node GetConfig() const { return YAML::Node(); }
};
