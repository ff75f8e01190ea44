// harness.cpp -- drives the REFERENCE's own binning code (struct SampleBinningInfo / BinInfo, header-only in
// /root/reference/Samples/SampleStructs.h) behind a C ABI, so that the oracle's restatement of
//   SampleBinningInfo::FindBin                      Samples/SampleStructs.h:577-613
//   InitialiseBinMigrationLookUp / Strides          :618-675
//   InitNonUniform + InitialiseGridMapping          :394-528
//   BinInfo::IsEventInside                          :207-219
// can be pinned against the reference itself (tests/test_reference_host.py, tests/golden/ref_host_binning.npz).
//
// TEST INFRASTRUCTURE (part of oracle/).  The reference header is compiled where it lies (-I/root/reference);
// ROOT, spdlog, yaml-cpp and NuOscillator are absent from this image, so oracle/ref_host/stubs/ provides empty
// stand-ins for their headers (none of their functionality is used by the code exercised here).  The ten lines
// of BinningHandler::FindGlobalBin (Samples/BinningHandler.cpp:257-291) are re-stated here on the reference's
// structures: FindBin per dimension, Strides, then for non-uniform binning BinGridMapping[mega] in order with
// Bins[b].IsEventInside().  (harness_path.cpp compiles Samples/BinningHandler.cpp itself and runs the REAL
// FindGlobalBin; tests/test_reference_host.py checks that both give the same bins.)
#include "Samples/SampleStructs.h"

#include <cstdint>
#include <vector>

namespace {
struct Holder {
  SampleBinningInfo info;
};
}  // namespace

#define REFH_API extern "C" __attribute__((visibility("default")))

// uniform: n_dim axes, nbins[d], edges concatenated
REFH_API void* refh_uniform(int n_dim, const int* nbins, const double* edges) {
  std::vector<std::vector<double>> e(n_dim);
  for (int d = 0; d < n_dim; ++d) { e[d].assign(edges, edges + nbins[d] + 1); edges += nbins[d] + 1; }
  Holder* h = new Holder();
  try { h->info.InitUniform(e); } catch (...) { delete h; return nullptr; }
  return h;
}

// non-uniform: n_boxes boxes of n_dim {lo,hi} pairs
REFH_API void* refh_nonuniform(int n_dim, int n_boxes, const double* extent) {
  std::vector<std::vector<std::vector<double>>> in(n_boxes, std::vector<std::vector<double>>(n_dim, std::vector<double>(2)));
  for (int b = 0; b < n_boxes; ++b)
    for (int d = 0; d < n_dim; ++d) { in[b][d][0] = extent[(size_t(b) * n_dim + d) * 2]; in[b][d][1] = extent[(size_t(b) * n_dim + d) * 2 + 1]; }
  Holder* h = new Holder();
  try { h->info.InitNonUniform(in); } catch (...) { delete h; return nullptr; }
  return h;
}

REFH_API void refh_destroy(void* p) { delete static_cast<Holder*>(p); }
REFH_API int refh_nbins(void* p) { return static_cast<Holder*>(p)->info.nBins; }
REFH_API int refh_axis_nbins(void* p, int d) { return static_cast<Holder*>(p)->info.AxisNBins[d]; }
REFH_API double refh_edge(void* p, int d, int i) { return static_cast<Holder*>(p)->info.BinEdges[d][i]; }
REFH_API int refh_stride(void* p, int d) { return static_cast<Holder*>(p)->info.Strides[d]; }
REFH_API int refh_grid_size(void* p, int mega) { return int(static_cast<Holder*>(p)->info.BinGridMapping[mega].size()); }
REFH_API int refh_grid_entry(void* p, int mega, int k) { return static_cast<Holder*>(p)->info.BinGridMapping[mega][k]; }

// SampleBinningInfo::FindBin(Dimension, Var, NomBin) for n values
REFH_API void refh_find_bin(void* p, int dim, int n, const double* var, const int* nom_bin, int* out) {
  const SampleBinningInfo& info = static_cast<Holder*>(p)->info;
  for (int i = 0; i < n; ++i) out[i] = info.FindBin(dim, var[i], nom_bin[i]);
}

// the event's bin within the sample (GlobalOffset 0), BinningHandler::FindGlobalBin's logic on the reference's structures;
// kin is dim-major [d*n + i]; nom_bin likewise (may hold -1)
REFH_API void refh_find_sample_bin(void* p, int n, const double* kin, const int* nom_bin, int* out) {
  const SampleBinningInfo& Binning = static_cast<Holder*>(p)->info;
  const int Dim = int(Binning.BinEdges.size());
  std::vector<const double*> KinVar(Dim);
  for (int i = 0; i < n; ++i) {
    int GlobalBin = 0;
    bool oob = false;
    for (int d = 0; d < Dim; ++d) {
      KinVar[d] = &kin[size_t(d) * n + i];
      const int Bin = Binning.FindBin(d, *KinVar[d], nom_bin[size_t(d) * n + i]);
      if (Bin < 0) { oob = true; break; }
      GlobalBin += Bin * Binning.Strides[d];
    }
    if (oob) { out[i] = M3::UnderOverFlowBin; continue; }
    if (Binning.Uniform) { out[i] = GlobalBin; continue; }
    out[i] = M3::UnderOverFlowBin;
    const auto& BinMapping = Binning.BinGridMapping[GlobalBin];
    for (size_t k = 0; k < BinMapping.size(); ++k) {
      const int BinNumber = BinMapping[k];
      if (Binning.Bins[BinNumber].IsEventInside(KinVar)) { out[i] = BinNumber; break; }
    }
  }
}
