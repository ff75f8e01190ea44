// harness_path.cpp -- drives the REFERENCE's own host implementation of the likelihood path behind a C ABI:
//   SMonolith (Splines/SplineMonolith.cpp, CPU build: no MaCh3_CUDA)
//     constructor -> ScanMasterSpline :318-440, PrepareForGPU :53-250   (flattening into the monolith arrays)
//     Evaluate :712-723 -> SplineBase::FindSplineSegment (Splines/SplineBase.cpp:44-111),
//                          CalcSplineWeights :727-789, CalcTotalEventWeight :792-832
//   SampleHandlerBase::GetTestStatLLH / GetPoissonLLH (Samples/SampleHandlerBase.cpp:17-193)
// Those three translation units are compiled WHERE THEY LIE under /root/reference (see Makefile); nothing of
// them is copied.  ROOT, spdlog and yaml-cpp are absent from this image: stubs/ holds compile-only stand-ins
// (root_fwd.h) and shadows of three reference headers that exist only to pull those libraries in
// (Manager/Manager.h, Manager/MaCh3Modes.h, Samples/HistogramUtils.h, Parameters/ParameterHandlerGeneric.h).
// The code exercised here never calls into them: responses are handed over as the reference's own reduced
// objects, TSpline3_red(X, Y, N, P) (Splines/SplineStructs.h:285) and TF1_red + SetSize/SetParameter (:215-234).
//
// TEST INFRASTRUCTURE (part of oracle/): pins the oracle's restatement of the host path to the reference itself
// (tests/test_reference_path.py, tests/golden/ref_host_path.npz).  Built with -fno-access-control so the private
// monolith arrays can be read back and handed, unchanged, to the oracle and to libm3b200's
// m3b_upload_spline_monolith.
#include "Splines/SplineMonolith.h"
#include "Samples/SampleHandlerBase.h"

#include <cstdint>
#include <cstring>
#include <vector>

TStyle* gStyle = nullptr;
TDirectory* gDirectory = nullptr;
TROOT* gROOT = nullptr;

#define REFP_API extern "C" __attribute__((visibility("default")))

namespace {
struct Mono {
  SMonolith* m = nullptr;
  std::vector<double> pars;
};

// the smallest concrete SampleHandlerBase: only the test statistic is used
struct Stat final : SampleHandlerBase {
  std::string GetSampleTitle(const int) const override { return ""; }
  std::string GetName() const override { return "ref_host"; }
  double GetSampleLikelihood(const int) const override { return 0; }
  void CleanMemoryBeforeFit() override {}
  void Reweight() override {}
  double GetLikelihood() const override { return 0; }
  void PrintRates(const bool) override {}
  int GetNOscChannels(const int) const override { return 0; }
  std::string GetKinVarName(const int, const int) const override { return ""; }
  TH1* GetDataHist(const int) override { return nullptr; }
  TH1* GetMCHist(const int) override { return nullptr; }
  TH1* GetW2Hist(const int) override { return nullptr; }
  int GetNDim(const int) const override { return 0; }
  std::string GetFlavourName(const int, const int) const override { return ""; }
  std::vector<double> ReturnKinematicParameterBinning(const int, const std::string&) const override { return {}; }
  TH1* Get1DVarHistByModeAndChannel(const int, const std::string&, int, int, int, TAxis*) override { return nullptr; }
  TH2* Get2DVarHistByModeAndChannel(const int, const std::string&, const std::string&, int, int, int, TAxis*, TAxis*) override { return nullptr; }
  TH1* Get1DVarHist(const int, const std::string&, const std::vector<KinematicCut>&, int, TAxis*, const std::vector<KinematicCut>&) override { return nullptr; }
  TH2* Get2DVarHist(const int, const std::string&, const std::string&, const std::vector<KinematicCut>&, int, TAxis*, TAxis*, const std::vector<KinematicCut>&) override { return nullptr; }
};
}  // namespace

// ---- SMonolith ------------------------------------------------------------------------------------------------
// type[P]: 0 = kTSpline3_red, 1 = kTF1_red.  npts[n_events*P]: knots of the event's response to parameter p
// (0 = the event has none: nullptr in MasterSpline; TF1: number of coefficients, 2).  vals: 5 doubles per knot, in
// event-major, parameter, knot order: {x, y, b, c, d}; for a TF1 coefficient k only column 1 is read.
REFP_API void* refp_mono_create(int n_events, int P, const int* type, const int* npts, const double* vals) {
  std::vector<std::vector<TResponseFunction_red*>> master(n_events, std::vector<TResponseFunction_red*>(P, nullptr));
  std::vector<RespFuncType> types(P);
  for (int p = 0; p < P; ++p) types[p] = type[p] ? kTF1_red : kTSpline3_red;
  const double* v = vals;
  std::vector<M3::float_t> X, Y, B, Cc, D;
  for (int e = 0; e < n_events; ++e)
    for (int p = 0; p < P; ++p) {
      const int n = npts[size_t(e) * P + p];
      if (n == 0) continue;
      X.resize(n); Y.resize(n); B.resize(n); Cc.resize(n); D.resize(n);
      for (int k = 0; k < n; ++k, v += 5) { X[k] = v[0]; Y[k] = v[1]; B[k] = v[2]; Cc[k] = v[3]; D[k] = v[4]; }
      if (type[p]) {
        // not TF1_red(n, array): that constructor (Splines/SplineStructs.h:166-171) writes through an unallocated Par
        TF1_red* f = new TF1_red();
        f->SetSize(M3::int_t(n));
        for (int k = 0; k < n; ++k) f->SetParameter(M3::int_t(k), Y[k]);
        master[e][p] = f;
      } else {
        std::vector<M3::float_t*> rows(n);
        std::vector<M3::float_t> bcd(size_t(n) * 3);
        for (int k = 0; k < n; ++k) { bcd[3 * k] = B[k]; bcd[3 * k + 1] = Cc[k]; bcd[3 * k + 2] = D[k]; rows[k] = &bcd[3 * k]; }
        master[e][p] = new TSpline3_red(X.data(), Y.data(), M3::int_t(n), rows.data());
      }
    }
  Mono* h = new Mono();
  try {
    h->m = new SMonolith(master, types, false);
  } catch (...) { delete h; return nullptr; }
  h->pars.assign(P, 0.0);
  std::vector<const double*> ptrs(P);
  for (int p = 0; p < P; ++p) ptrs[p] = &h->pars[p];
  h->m->setSplinePointers(ptrs);
  return h;
}
REFP_API void refp_mono_destroy(void* p) { Mono* h = static_cast<Mono*>(p); delete h->m; delete h; }

// sizes: {NEvents, nParams, _max_knots, NSplines_valid, NTF1_valid, nKnots, nTF1coeff}
REFP_API void refp_mono_sizes(void* p, int64_t* out) {
  SMonolith* m = static_cast<Mono*>(p)->m;
  out[0] = m->NEvents; out[1] = m->nParams; out[2] = m->_max_knots; out[3] = m->NSplines_valid;
  out[4] = m->NTF1_valid; out[5] = m->nKnots; out[6] = m->nTF1coeff;
}
// which: 0 coeff_x (f32) 1 coeff_many (f32) 2 nKnots_arr (u32) 3 paramNo_arr (i16) 4 cpu_nParamPerEvent (u32)
//        5 cpu_nParamPerEvent_tf1 (u32) 6 cpu_paramNo_TF1_arr (i16) 7 cpu_coeff_TF1_many (f32)
//        8 SplineInfoArray[].nPts (i16 out)   9 SplineInfoArray[].xPts as doubles, rows padded to _max_knots with 0
//        returns the element count; copies when out != nullptr
REFP_API int64_t refp_mono_array(void* p, int which, void* out) {
  SMonolith* m = static_cast<Mono*>(p)->m;
  auto give = [&](const void* src, size_t n, size_t sz) -> int64_t { if (out && n) std::memcpy(out, src, n * sz); return int64_t(n); };
  switch (which) {
    case 0: return give(m->cpu_spline_handler->coeff_x.data(), m->cpu_spline_handler->coeff_x.size(), 4);
    case 1: return give(m->cpu_spline_handler->coeff_many.data(), m->cpu_spline_handler->coeff_many.size(), 4);
    case 2: return give(m->cpu_spline_handler->nKnots_arr.data(), m->cpu_spline_handler->nKnots_arr.size(), 4);
    case 3: return give(m->cpu_spline_handler->paramNo_arr.data(), m->cpu_spline_handler->paramNo_arr.size(), 2);
    case 4: return give(m->cpu_nParamPerEvent.data(), m->cpu_nParamPerEvent.size(), 4);
    case 5: return give(m->cpu_nParamPerEvent_tf1.data(), m->cpu_nParamPerEvent_tf1.size(), 4);
    case 6: return give(m->cpu_paramNo_TF1_arr.data(), m->cpu_paramNo_TF1_arr.size(), 2);
    case 7: return give(m->cpu_coeff_TF1_many.data(), m->cpu_coeff_TF1_many.size(), 4);
    case 8: {
      const size_t n = m->SplineInfoArray.size();
      if (out) for (size_t i = 0; i < n; ++i) static_cast<int16_t*>(out)[i] = int16_t(m->SplineInfoArray[i].xPts.empty() ? 0 : m->SplineInfoArray[i].nPts);
      return int64_t(n);
    }
    case 9: {
      const size_t n = m->SplineInfoArray.size(), K = size_t(m->_max_knots);
      if (out) {
        double* o = static_cast<double*>(out);
        for (size_t i = 0; i < n * K; ++i) o[i] = 0.0;
        for (size_t i = 0; i < n; ++i)
          for (size_t k = 0; k < m->SplineInfoArray[i].xPts.size() && k < K; ++k) o[i * K + k] = double(m->SplineInfoArray[i].xPts[k]);
      }
      return int64_t(n * K);
    }
  }
  return -1;
}

// one step: copy the parameter values the pointers look at, SMonolith::Evaluate(), read back
REFP_API int refp_mono_evaluate(void* p, const double* pars, float* weights, int16_t* segments, float* param_values) {
  Mono* h = static_cast<Mono*>(p);
  SMonolith* m = h->m;
  for (size_t i = 0; i < h->pars.size(); ++i) h->pars[i] = pars[i];
  try { m->Evaluate(); } catch (...) { return 1; }
  if (weights) std::memcpy(weights, m->cpu_total_weights, size_t(m->NEvents) * sizeof(float));
  if (segments) std::memcpy(segments, m->SplineSegments, size_t(m->nParams) * sizeof(short));
  if (param_values) std::memcpy(param_values, m->ParamValues, size_t(m->nParams) * sizeof(float));
  return 0;
}

// ---- test statistics ------------------------------------------------------------------------------------------
// kind follows enum TestStatistic (Samples/SampleStructs.h:105-112).  Returns the number of bins the reference threw on.
REFP_API int refp_test_stat(int kind, int n, const double* data, const double* mc, const double* w2, double* out) {
  Stat s;
  s.SetTestStatistic(static_cast<TestStatistic>(kind));
  int thrown = 0;
  for (int i = 0; i < n; ++i) {
    try { out[i] = s.GetTestStatLLH(data[i], mc[i], w2[i]); } catch (...) { out[i] = -999.0; ++thrown; }
  }
  return thrown;
}
REFP_API void refp_poisson(int n, const double* data, const double* mc, double* out) {
  Stat s;
  for (int i = 0; i < n; ++i) out[i] = s.GetPoissonLLH(data[i], mc[i]);
}
REFP_API double refp_low_mc_bound() { return M3::_LOW_MC_BOUND_; }
