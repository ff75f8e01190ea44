// M3BMonolithFile.h -- the ROOT-free spline-monolith cache file, reference side.
//
// SMonolith::PrepareSplineFile (Splines/SplineMonolith.cpp:543-614) writes the flattened monolith into a ROOT file so
// that later runs skip the per-event TSpline3 flattening (LoadSplineFile, :454-540).  libm3b200 has no ROOT dependency;
// a MaCh3 maintainer adds ONE call next to PrepareSplineFile (SMonolith's members are private: from a member function,
// or a friend) and later runs load the flat file straight onto the device(s):
//
//     void SMonolith::PrepareSplineFile(std::string FileName) {
//       ...                                                       // the reference's ROOT output, unchanged
//       std::vector<int16_t> n_pts(nParams);
//       std::vector<double>  x_pts(size_t(nParams) * _max_knots, 0.0);
//       for (int i = 0; i < nParams; ++i) {                       // FastSplineInfo (Splines/SplineStructs.h:21-44)
//         n_pts[i] = SplineInfoArray[i].xPts.empty() ? 0 : SplineInfoArray[i].nPts;
//         for (size_t k = 0; k < SplineInfoArray[i].xPts.size(); ++k) x_pts[size_t(i) * _max_knots + k] = SplineInfoArray[i].xPts[k];
//       }
//       m3b200::MonolithArrays a = ...;                           // as for SampleHandlerB200::MoveToB200
//       m3b200::WriteMonolithFile(FileName + ".m3b", a, n_pts.data(), x_pts.data(), NEvents);
//     }
//
// and, instead of SMonolith(FileName) + MoveToGPU:   m3b_upload_from_file(handle, "SplineFile.root.m3b", 0);
// (or m3b_group_upload_from_file for a sample spread over several B200s).  Round trip tested in tests/test_monolith_file.py.
#pragma once
#include "SampleHandlerB200.h"

namespace m3b200 {

inline void WriteMonolithFile(const std::string& path, const MonolithArrays& a, const int16_t* n_pts, const double* x_pts_f64,
                              int64_t n_events) {
  const int rc = m3b_write_monolith_file(path.c_str(), a.n_params, a.max_knots, a.coeff_x, n_pts ? n_pts : a.n_pts, x_pts_f64, n_events,
                                         a.nParamPerEvent, a.paramNo_arr, a.nKnots_arr, a.total_knots, a.coeff_many,
                                         a.nParamPerEvent_tf1, a.paramNo_tf1, a.coeff_tf1);
  if (rc != M3B_OK) throw std::runtime_error(std::string("m3b200::WriteMonolithFile: ") + m3b_last_error(nullptr));
}

}  // namespace m3b200
