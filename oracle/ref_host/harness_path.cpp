// harness_path.cpp -- drives the REFERENCE's own host implementation of the likelihood path behind a C ABI:
//   SMonolith (Splines/SplineMonolith.cpp, CPU build: no MaCh3_CUDA)
//     constructor -> ScanMasterSpline :318-440, PrepareForGPU :53-250   (flattening into the monolith arrays)
//     Evaluate :712-723 -> SplineBase::FindSplineSegment (Splines/SplineBase.cpp:44-111),
//                          CalcSplineWeights :727-789, CalcTotalEventWeight :792-832
//   SampleHandlerBase::GetTestStatLLH / GetPoissonLLH (Samples/SampleHandlerBase.cpp:17-193)
// (the second half of this file adds SampleHandlerFD, BinningHandler and BinnedSplineHandler).  The reference's
// translation units are compiled WHERE THEY LIE under /root/reference (see Makefile); nothing of them is copied.
// ROOT, spdlog, yaml-cpp and NuOscillator are absent from this image: stubs/ holds compile-only stand-ins for their
// headers (root_fwd.h, yaml-cpp/yaml.h, spdlog/spdlog.h: permissive classes that abort if ever called) and ONE shadow
// of a reference header, Parameters/ParameterHandlerGeneric.h, whose real version needs ROOT's matrix classes; all
// other reference headers (Manager.h, YamlHelper.h, MaCh3Modes.h, HistogramUtils.h, ...) are the reference's own.
// The code exercised here never calls into the stand-ins: responses are handed over as the reference's own reduced
// objects, TSpline3_red(X, Y, N, P) (Splines/SplineStructs.h:285) and TF1_red + SetSize/SetParameter (:215-234).
//
// TEST INFRASTRUCTURE (part of oracle/): pins the oracle's restatement of the host path to the reference itself
// (tests/test_reference_path.py, tests/golden/ref_host_path.npz).  Built with -fno-access-control so the private
// monolith arrays can be read back and handed, unchanged, to the oracle and to libm3b200's
// m3b_upload_spline_monolith.
#include "Splines/SplineMonolith.h"
#include "Samples/SampleHandlerBase.h"
#ifdef MaCh3_CUDA
#include "Splines/gpuSplineUtils.cuh"
#endif

#include <algorithm>
#include <cstdint>
#include <cstdio>
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
// The same SMonolith, for workloads too large to hand over as one heap object per response (the bench's CPU
// baseline: millions of responses).  A one-event monolith goes through the reference's constructor, then its
// members are replaced by the monolith arrays as PrepareForGPU would have left them (the layout the tests above
// verify against the constructor's own output).  Evaluate() is the reference's, on the reference's layout.
REFP_API void* refp_mono_create_from_arrays(int P, int max_knots, const float* coeff_x, const int16_t* n_pts, const int8_t* type,
                                            int64_t n_events, const uint32_t* nParamPerEvent, const int16_t* paramNo_arr,
                                            const uint64_t* nKnots_arr, uint64_t total_knots, const float* coeff_many,
                                            const uint32_t* nParamPerEvent_tf1, const int16_t* paramNo_tf1, const float* coeff_tf1) {
  std::vector<std::vector<TResponseFunction_red*>> master(1, std::vector<TResponseFunction_red*>(P, nullptr));
  std::vector<RespFuncType> types(P);
  M3::float_t X[2] = {0, 1}, Y[2] = {1, 1}, z[3] = {0, 0, 0};
  M3::float_t* rows[2] = {z, z};
  for (int p = 0; p < P; ++p) {
    types[p] = type[p] ? kTF1_red : kTSpline3_red;
    if (type[p]) { TF1_red* f = new TF1_red(); f->SetSize(2); f->SetParameter(0, 0); f->SetParameter(1, 1); master[0][p] = f; }
    else master[0][p] = new TSpline3_red(X, Y, 2, rows);
  }
  Mono* h = new Mono();
  try { h->m = new SMonolith(master, types, false); } catch (...) { delete h; return nullptr; }
  SMonolith* m = h->m;
  uint64_t ns = 0, nl = 0;
  for (int64_t e = 0; e < n_events; ++e) { ns += nParamPerEvent[2 * e]; nl += nParamPerEvent_tf1[2 * e]; }
#ifdef MaCh3_CUDA
  // the one-event monolith already moved to the GPU and freed its host arrays (SplineMonolith.cpp:254-313): release
  // it the way the destructor does (:621-626) and start again from empty host structures
  m->gpu_spline_handler->CleanupGPU_SplineMonolith(m->cpu_total_weights);
  m->gpu_spline_handler->CleanupGPU_Segments(m->SplineSegments, m->ParamValues);
  delete m->gpu_spline_handler;
  m->gpu_spline_handler = nullptr;
  m->cpu_total_weights = nullptr; m->SplineSegments = nullptr; m->ParamValues = nullptr;
  m->cpu_spline_handler = new SplineMonoStruct();
#endif
  m->NEvents = unsigned(n_events); m->_max_knots = short(max_knots);
  m->NSplines_valid = unsigned(ns); m->NTF1_valid = unsigned(nl); m->nKnots = unsigned(total_knots); m->nTF1coeff = unsigned(nl * 2);
  m->cpu_spline_handler->coeff_x.assign(coeff_x, coeff_x + size_t(P) * max_knots);
  m->cpu_spline_handler->coeff_many.assign(coeff_many, coeff_many + total_knots * 4);
  m->cpu_spline_handler->nKnots_arr.resize(ns);
  for (uint64_t i = 0; i < ns; ++i) m->cpu_spline_handler->nKnots_arr[i] = unsigned(nKnots_arr[i]);
  m->cpu_spline_handler->paramNo_arr.assign(paramNo_arr, paramNo_arr + ns);
  m->cpu_nParamPerEvent.assign(nParamPerEvent, nParamPerEvent + 2 * n_events);
  m->cpu_nParamPerEvent_tf1.assign(nParamPerEvent_tf1, nParamPerEvent_tf1 + 2 * n_events);
  m->cpu_paramNo_TF1_arr.assign(paramNo_tf1, paramNo_tf1 + nl);
  m->cpu_coeff_TF1_many.assign(coeff_tf1, coeff_tf1 + nl * 2);
  for (int p = 0; p < P; ++p) {
    m->SplineInfoArray[p].nPts = M3::int_t(n_pts[p]);
    m->SplineInfoArray[p].xPts.assign(n_pts[p] > 0 ? size_t(n_pts[p]) : 0, M3::float_t(0));
    for (int k = 0; k < n_pts[p]; ++k) m->SplineInfoArray[p].xPts[k] = M3::float_t(coeff_x[size_t(p) * max_knots + k]);
    m->SplineInfoArray[p].CurrSegment = 0;
  }
#ifdef MaCh3_CUDA
  // what PrepareForGPU does before MoveToGPU in this build (:81-95): pinned segment / value arrays, then the move
  m->gpu_spline_handler->InitGPU_Segments(&m->SplineSegments);       // static-like: do not touch `this` (called on nullptr, :82)
  m->gpu_spline_handler->InitGPU_Vals(&m->ParamValues);
  for (int p = 0; p < P; ++p) { m->SplineSegments[p] = 0; m->ParamValues[p] = -999; }
  m->MoveToGPU();
#else
  delete[] m->cpu_total_weights; delete[] m->cpu_weights_spline_var; delete[] m->cpu_weights_tf1_var;
  m->cpu_total_weights = new float[size_t(n_events) + 1]();
  m->cpu_weights_spline_var = new float[ns + 1]();
  m->cpu_weights_tf1_var = new float[nl + 1]();
#endif
  h->pars.assign(P, 0.0);
  std::vector<const double*> ptrs(P);
  for (int p = 0; p < P; ++p) ptrs[p] = &h->pars[p];
  m->setSplinePointers(ptrs);
  return h;
}
#ifdef MULTITHREAD
#include <omp.h>
REFP_API int refp_num_threads() { return omp_get_max_threads(); }     // the reference's MULTITHREAD build
#else
REFP_API int refp_num_threads() { return 1; }
#endif
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
  if (!m->cpu_spline_handler) return -1;      // MaCh3_CUDA build: handed to the GPU class and freed (SplineMonolith.cpp:305-311)
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
  try { m->Evaluate(); m->SynchroniseMemTransfer(); } catch (...) { return 1; }     // the fence is a no-op in the CPU build
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

// ==================================================================================================================
// SampleHandlerFD (Samples/SampleHandlerFD.cpp), BinningHandler (Samples/BinningHandler.cpp) and
// BinnedSplineHandler (Splines/BinnedSplineHandler.cpp): the per-step part of the path run by THE REFERENCE'S CODE
//   Reweight :316-343 -> ResetHistograms :442, SplineHandler->Evaluate(), FillArray :352-386
//                        (ApplyShifts, IsEventSelected, CalcWeightTotal :568-596, BinningHandler::FindGlobalBin :257-291,
//                        the `<= 0` skip, the W2 freeze FirstTimeW2/UpdateW2)
//   GetLikelihood :1284-1300, GetSampleLikelihood :1262-1281, FindNominalBinAndEdges :858-887, SetSplinePointers
//   :1196-1259 (monolith arm, _LOW_MEMORY_STRUCTS_ build), BinnedSplineHandler::Evaluate/CalcSplineWeights :295-341.
// What a MaCh3 experiment supplies through YAML, ROOT files and its SampleHandlerFD subclass -- binning, events,
// which parameters and weights every event points at -- is filled into the reference's own structures directly
// (-fno-access-control); nothing of the configuration-time code is run.
// ==================================================================================================================
#include "Samples/SampleHandlerFD.h"
#include "Splines/BinnedSplineHandler.h"

// ---- link-time stand-ins for the four reference functions whose translation units need yaml-cpp / NuOscillator at
//      run time (Manager.cpp, MaCh3Modes.cpp, OscillationHandler.cpp).  None is reached by the code driven here,
//      except the Manager constructor, which SampleHandlerFD's constructor calls to read its YAML: here it reads nothing.
Manager::Manager(std::string const& filename) { FileName = filename; }
Manager::~Manager() {}
std::string MaCh3Modes::GetMaCh3ModeName(const int) const { m3stub::no_root("MaCh3Modes"); }
std::string MaCh3Modes::GetSplineSuffixFromMaCh3Mode(const int) { m3stub::no_root("MaCh3Modes"); }
void OscillationHandler::Evaluate() { m3stub::no_root("OscillationHandler"); }
const M3::float_t* OscillationHandler::GetNuOscillatorPointers(const int, const int, const int, const int, const FLOAT_T, const FLOAT_T) { m3stub::no_root("OscillationHandler"); }

namespace {
struct FD : SampleHandlerFD {
  std::vector<double> kin;          // [event][4]: what KinVar points at
  // functional ("shift") parameters of the linear family: parameter s adds (*valuePtr) * shift_coef[s][event] to column
  // shift_target[s] of the event's kinematic row; the reference's own ApplyShifts (Samples/SampleHandlerFD.cpp:545-564)
  // drives them through funcParsGrid / FunctionalShifter / ResetShifts
  std::vector<double> kin_nom, shift_pars;
  std::vector<int> shift_target;
  std::vector<std::vector<double>> shift_coef;
  std::vector<FuncParFuncType> shift_funcs;
  void ResetShifts(const int iEvent) override {
    for (int d = 0; d < 4; ++d) kin[size_t(iEvent) * 4 + size_t(d)] = kin_nom[size_t(iEvent) * 4 + size_t(d)];
  }
  std::vector<double> norm;         // what norm_pointers point at (ParameterHandler::_fPropVal)
  std::vector<M3::float_t> pool;    // what total_weight_pointers point at (oscillation weights, extra weights)
  Mono* mono = nullptr;             // spline parameters live with the spline handler
  std::vector<double>* binned_pars = nullptr;
  FD() : SampleHandlerFD("ref_host", nullptr, nullptr) {}
  void CleanMemoryBeforeFit() override {}
  void AddAdditionalWeightPointers() override {}
  void SetupSplines() override {}
  void Init() override {}
  int SetupExperimentMC() override { return 0; }
  void SetupFDMC() override {}
  void RegisterFunctionalParameters() override {}
  double ReturnKinematicParameter(std::string, int) override { return 0.0; }
  // the experiment's kinematic-variable look-up: variable v of the event = column v of its row in `kin` (columns past
  // the sample's binning dimensions are free for variables that are only cut on)
  double ReturnKinematicParameter(int var, int iEvent) override { return kin[size_t(iEvent) * 4 + size_t(var)]; }
  const double* GetPointerToKinematicParameter(std::string name, int iEvent) override { return &kin[size_t(iEvent) * 4 + size_t(name[1] - '0')]; }
  const double* GetPointerToKinematicParameter(double, int) override { return nullptr; }
};

#ifdef M3B_WITH_ADAPTER
}  // namespace
// The drop-in adapter (adapters/SampleHandlerB200.h) instantiated over the REAL SampleHandlerFD class hierarchy:
// the fitters' three virtuals then run on the B200 through libm3b200's C ABI.
#include "SampleHandlerB200.h"
namespace {
using FDType = m3b200::SampleHandlerB200<FD>;
#else
using FDType = FD;
#endif

struct Binned final : BinnedSplineHandler {
  std::vector<double> pars;
  int64_t n_coeff_keep = 0;          // length of xcoeff_arr / manycoeff_arr (in knots)
  static ParameterHandlerGeneric* fake_xsec() { static ParameterHandlerGeneric x; return &x; }
  // only stored by the constructor (Splines/BinnedSplineHandler.cpp:14-32), never dereferenced by Evaluate
  static MaCh3Modes* fake_modes() { static long long blob[64]; return reinterpret_cast<MaCh3Modes*>(blob); }
  Binned() : BinnedSplineHandler(fake_xsec(), fake_modes()) {}
  std::vector<std::string> GetTokensFromSplineName(std::string) override { return {}; }
};
}  // namespace

// samples: n_dim[s]; uniform[s]; uniform: nbins[4*s+d] + edges concatenated; non-uniform: nbins[4*s] boxes of n_dim {lo,hi}
REFP_API void* refp_fd_create(int n_samples, const int* n_dim, const int* uniform, const int* nbins, const double* edges,
                              int test_statistic, int update_w2) {
  FD* fd = nullptr;
  try {
    fd = new FDType();
    fd->nSamples = M3::int_t(n_samples);
    fd->SampleDetails.resize(n_samples);
    const double* ep = edges;
    for (int s = 0; s < n_samples; ++s) {
      fd->SampleDetails[s].nDimensions = n_dim[s];
      fd->SampleDetails[s].SampleTitle = "s" + std::to_string(s);
      for (int d = 0; d < n_dim[s]; ++d) fd->SampleDetails[s].VarStr.push_back("v" + std::to_string(d));
      SampleBinningInfo info;
      if (uniform[s]) {
        std::vector<std::vector<double>> e(n_dim[s]);
        for (int d = 0; d < n_dim[s]; ++d) { e[d].assign(ep, ep + nbins[4 * s + d] + 1); ep += nbins[4 * s + d] + 1; }
        info.InitUniform(e);
      } else {
        const int nb = nbins[4 * s];
        std::vector<std::vector<std::vector<double>>> in(nb, std::vector<std::vector<double>>(n_dim[s], std::vector<double>(2)));
        for (int b = 0; b < nb; ++b)
          for (int d = 0; d < n_dim[s]; ++d) { in[b][d][0] = ep[0]; in[b][d][1] = ep[1]; ep += 2; }
        info.InitNonUniform(in);
      }
      fd->Binning->SampleBinning.emplace_back(info);
    }
    fd->Binning->SetGlobalBinNumbers();
    const int nb = fd->Binning->GetNBins();
    fd->SampleHandlerFD_array.assign(nb, 0.0);
    fd->SampleHandlerFD_array_w2.assign(nb, 0.0);
    fd->SampleHandlerFD_data.assign(nb, 0.0);
    fd->Selection.resize(n_samples);
    fd->StoredSelection.resize(n_samples);
    fd->SetTestStatistic(static_cast<TestStatistic>(test_statistic));
    fd->UpdateW2 = update_w2 != 0;
    fd->FirstTimeW2 = true;
  } catch (...) { delete fd; return nullptr; }
  return fd;
}
REFP_API void refp_fd_destroy(void* p) { delete static_cast<FD*>(p); }
REFP_API int refp_fd_nbins(void* p) { return static_cast<FD*>(p)->Binning->GetNBins(); }
REFP_API int refp_float_t_bytes() { return int(sizeof(M3::float_t)); }

// The spline handler moves into the sample handler (std::unique_ptr<SplineBase> SplineHandler), as in
// SampleHandlerFD::SetupSplines of an experiment.
REFP_API void refp_fd_attach_monolith(void* p, void* mono) {
  FD* fd = static_cast<FD*>(p);
  fd->mono = static_cast<Mono*>(mono);
  fd->SplineHandler.reset(fd->mono->m);
  fd->mono->m = nullptr;                 // owned by the sample handler from here on
}

// BinnedSplineHandler with its monolith arrays handed over as they stand after FillSampleArray / TransferToMonolith.
REFP_API void* refp_fd_attach_binned(void* p, int P, int max_knots, const double* knot_x, const int16_t* n_pts,
                                     int64_t n_slots, const int* uniquesplinevec, const int* coeffindexvec,
                                     int64_t n_unique, const int* uniquecoeffindices, int64_t n_coeff,
                                     const double* manycoeff, const double* xcoeff) {
  FD* fd = static_cast<FD*>(p);
  Binned* b = new Binned();
  b->nParams = short(P);
  b->SplineSegments = new short int[P]();
  b->ParamValues = new float[P]();
  b->pars.assign(P, 0.0);
  b->SplineInfoArray.resize(P);
  for (int i = 0; i < P; ++i) {
    b->SplineInfoArray[i].nPts = M3::int_t(n_pts[i]);
    b->SplineInfoArray[i].xPts.assign(n_pts[i] > 0 ? size_t(n_pts[i]) : 0, M3::float_t(0));
    for (int k = 0; k < n_pts[i]; ++k) b->SplineInfoArray[i].xPts[k] = M3::float_t(knot_x[size_t(i) * max_knots + k]);
    b->SplineInfoArray[i].CurrSegment = 0;
    b->SplineInfoArray[i].splineParsPointer = &b->pars[i];
  }
  b->uniquesplinevec_Monolith.assign(uniquesplinevec, uniquesplinevec + n_slots);
  b->coeffindexvec.assign(coeffindexvec, coeffindexvec + n_slots);
  b->uniquecoeffindices.assign(uniquecoeffindices, uniquecoeffindices + n_unique);
  b->weightvec_Monolith.assign(size_t(n_slots), M3::float_t(1));        // flat slots stay 1 (.cpp:672)
  b->n_coeff_keep = n_coeff;
  b->xcoeff_arr = new M3::float_t[size_t(n_coeff)];
  b->manycoeff_arr = new M3::float_t[size_t(n_coeff) * 4];
  for (int64_t i = 0; i < n_coeff; ++i) b->xcoeff_arr[i] = M3::float_t(xcoeff[i]);
  for (int64_t i = 0; i < n_coeff * 4; ++i) b->manycoeff_arr[i] = M3::float_t(manycoeff[i]);
  fd->binned_pars = &b->pars;
  fd->SplineHandler.reset(b);
  return b;
}

// Events.  Weight pointers are pushed in the reference's order (SampleHandlerFD::Initialise :169-202):
//   w_before (oscillation weight; SetupOscParameters) -> spline weights (SetSplinePointers) -> w_after (extras).
// w_before / w_after: [n_events * n_*] indices into the M3::float_t pool, -1 = none.
// monolith: the reference's own SetSplinePointers() wires SMonolith::retPointer(event) (_LOW_MEMORY_STRUCTS_ only);
// binned: slot indices per event (CSR binned_start[n_events+1], binned_slot[]) -> &weightvec_Monolith[slot].
REFP_API int refp_fd_set_events(void* p, int n_events, const int* sample_id, const double* kin /* [n_events][4] */,
                                int n_norm_per_event, const int16_t* norm_idx, int n_norm_values,
                                int n_before, const int* w_before, int n_after, const int* w_after, int n_pool,
                                const int64_t* binned_start, const int* binned_slot) {
  FD* fd = static_cast<FD*>(p);
  try {
    fd->nEvents = unsigned(n_events);
    fd->kin.assign(kin, kin + size_t(n_events) * 4);
    fd->norm.assign(size_t(n_norm_values > 0 ? n_norm_values : 1), 1.0);
    fd->pool.assign(size_t(n_pool > 0 ? n_pool : 1), M3::float_t(1));
    fd->MCSamples = std::vector<EventInfo>(size_t(n_events));
    fd->funcParsGrid.assign(size_t(n_events), {});
    for (int e = 0; e < n_events; ++e) {
      EventInfo& ev = fd->MCSamples[e];
      ev.NominalSample = sample_id[e];
      for (int k = 0; k < n_norm_per_event; ++k) {
        const int16_t i = norm_idx[size_t(e) * n_norm_per_event + k];
        if (i >= 0) ev.norm_pointers.push_back(&fd->norm[i]);
      }
      for (int k = 0; k < n_before; ++k) {
        const int i = w_before[size_t(e) * n_before + k];
        if (i >= 0) ev.total_weight_pointers.push_back(&fd->pool[i]);
      }
    }
    fd->FindNominalBinAndEdges();                             // the reference's own (KinVar pointers, NomBin)
    if (fd->mono) {
      fd->SetSplinePointers();                                // the reference's own; throws in the double build
    } else if (binned_start) {
      Binned* b = static_cast<Binned*>(fd->SplineHandler.get());
      for (int e = 0; e < n_events; ++e)
        for (int64_t k = binned_start[e]; k < binned_start[e + 1]; ++k)
          fd->MCSamples[e].total_weight_pointers.push_back(&b->weightvec_Monolith[binned_slot[k]]);
    }
    for (int e = 0; e < n_events; ++e)
      for (int k = 0; k < n_after; ++k) {
        const int i = w_after[size_t(e) * n_after + k];
        if (i >= 0) fd->MCSamples[e].total_weight_pointers.push_back(&fd->pool[i]);
      }
  } catch (...) { return 1; }
  return 0;
}

// StoredSelection, as SampleHandlerFD::ReadConfig leaves it (Samples/SampleHandlerFD.cpp:140-165): per sample a list of
// KinematicCut {ParamToCutOnIt, LowerBound, UpperBound}.  FillArray[_MP] copies it into Selection at every fill and
// IsEventSelected (:281-294) applies it through ReturnKinematicParameter(ParamToCutOnIt, event).
REFP_API void refp_fd_set_selection(void* p, int n_cuts, const int* cut_sample, const int* cut_var, const double* lower, const double* upper) {
  FD* fd = static_cast<FD*>(p);
  for (auto& s : fd->StoredSelection) s.clear();
  for (int k = 0; k < n_cuts; ++k) {
    KinematicCut c;
    c.ParamToCutOnIt = cut_var[k]; c.LowerBound = lower[k]; c.UpperBound = upper[k];
    fd->StoredSelection[size_t(cut_sample[k])].emplace_back(c);
  }
  fd->Selection = fd->StoredSelection;
}
REFP_API void refp_fd_selected(void* p, unsigned char* out) {
  FD* fd = static_cast<FD*>(p);
  fd->Selection = fd->StoredSelection;
  for (unsigned e = 0; e < fd->nEvents; ++e) out[e] = fd->IsEventSelected(fd->MCSamples[e].NominalSample, int(e)) ? 1 : 0;
}

// Functional parameters: n_pars shifters; shifter s applies to the events with a finite coef[s*n_events + e] (NaN = the
// event is not in its funcParsGrid list) and adds (*valuePtr) * coef to kinematic column target[s].  Wired the way
// SampleHandlerFD::SetupFunctionalParameters leaves it (:470-540): funcParsMap[s] = {valuePtr, funcPtr}, funcParsGrid[e] =
// the shifters of event e in parameter order.  Call after refp_fd_set_events (the nominal kinematics are read now).
REFP_API void refp_fd_set_linear_shifts(void* p, int n_pars, const int* target, const double* coef) {
  FD* fd = static_cast<FD*>(p);
  const size_t E = fd->nEvents;
  fd->kin_nom = fd->kin;
  fd->shift_pars.assign(size_t(n_pars), 0.0);
  fd->shift_target.assign(target, target + n_pars);
  fd->shift_coef.assign(size_t(n_pars), std::vector<double>(E));
  fd->shift_funcs.resize(size_t(n_pars));
  fd->funcParsMap.resize(size_t(n_pars));
  for (int s = 0; s < n_pars; ++s) {
    std::copy(coef + size_t(s) * E, coef + size_t(s + 1) * E, fd->shift_coef[size_t(s)].begin());
    fd->shift_funcs[size_t(s)] = [fd, s](const double* par, std::size_t iEvent) {
      fd->kin[iEvent * 4 + size_t(fd->shift_target[size_t(s)])] += (*par) * fd->shift_coef[size_t(s)][iEvent];
    };
    fd->funcParsMap[size_t(s)].valuePtr = &fd->shift_pars[size_t(s)];
    fd->funcParsMap[size_t(s)].funcPtr = &fd->shift_funcs[size_t(s)];
  }
  fd->funcParsGrid.assign(E, {});
  for (size_t e = 0; e < E; ++e)
    for (int s = 0; s < n_pars; ++s)
      if (fd->shift_coef[size_t(s)][e] == fd->shift_coef[size_t(s)][e]) fd->funcParsGrid[e].push_back(&fd->funcParsMap[size_t(s)]);
}
REFP_API void refp_fd_set_shift_pars(void* p, const double* vals) {
  FD* fd = static_cast<FD*>(p);
  std::copy(vals, vals + fd->shift_pars.size(), fd->shift_pars.begin());
}
REFP_API void refp_fd_get_kin(void* p, double* out) {
  FD* fd = static_cast<FD*>(p);
  std::copy(fd->kin.begin(), fd->kin.end(), out);
}

REFP_API void refp_fd_set_data(void* p, const double* data) {
  FD* fd = static_cast<FD*>(p);
  std::copy(data, data + fd->SampleHandlerFD_data.size(), fd->SampleHandlerFD_data.begin());
}
// shifted kinematics (what functional parameters write): same storage, the nominal bins stay
REFP_API void refp_fd_set_kin(void* p, const double* kin) {
  FD* fd = static_cast<FD*>(p);
  std::copy(kin, kin + fd->kin.size(), fd->kin.begin());
}
REFP_API void refp_fd_set_test_statistic(void* p, int kind) { static_cast<FD*>(p)->SetTestStatistic(static_cast<TestStatistic>(kind)); }

// one step: the values every pointer looks at, then SampleHandlerFD::Reweight()
REFP_API int refp_fd_reweight(void* p, const double* spline_pars, const double* norm, const double* pool) {
  FD* fd = static_cast<FD*>(p);
  if (spline_pars && fd->mono) for (size_t i = 0; i < fd->mono->pars.size(); ++i) fd->mono->pars[i] = spline_pars[i];
  if (spline_pars && fd->binned_pars) for (size_t i = 0; i < fd->binned_pars->size(); ++i) (*fd->binned_pars)[i] = spline_pars[i];
  if (norm) std::copy(norm, norm + fd->norm.size(), fd->norm.begin());
  if (pool) for (size_t i = 0; i < fd->pool.size(); ++i) fd->pool[i] = M3::float_t(pool[i]);
  try { fd->Reweight(); } catch (...) { return 1; }
  return 0;
}
REFP_API double refp_fd_llh(void* p) { return static_cast<FD*>(p)->GetLikelihood(); }
REFP_API double refp_fd_sample_llh(void* p, int s) { return static_cast<FD*>(p)->GetSampleLikelihood(s); }
REFP_API void refp_fd_read(void* p, double* mc, double* w2) {
  FD* fd = static_cast<FD*>(p);
  std::copy(fd->SampleHandlerFD_array.begin(), fd->SampleHandlerFD_array.end(), mc);
  std::copy(fd->SampleHandlerFD_array_w2.begin(), fd->SampleHandlerFD_array_w2.end(), w2);
}
// per event: CalcWeightTotal and FindGlobalBin as FillArray calls them
REFP_API void refp_fd_events(void* p, double* weight, int* bin) {
  FD* fd = static_cast<FD*>(p);
  for (unsigned e = 0; e < fd->nEvents; ++e) {
    const EventInfo* ev = &fd->MCSamples[e];
    if (weight) weight[e] = double(fd->CalcWeightTotal(ev));
    if (bin) bin[e] = fd->Binning->FindGlobalBin(ev->NominalSample, ev->KinVar, ev->NomBin);
  }
}
REFP_API int64_t refp_fd_binned_weights(void* p, double* out) {
  Binned* b = dynamic_cast<Binned*>(static_cast<FD*>(p)->SplineHandler.get());
  if (!b) return -1;
  if (out) for (size_t i = 0; i < b->weightvec_Monolith.size(); ++i) out[i] = double(b->weightvec_Monolith[i]);
  return int64_t(b->weightvec_Monolith.size());
}
REFP_API void refp_fd_segments(void* p, int16_t* out) {
  SplineBase* s = static_cast<FD*>(p)->SplineHandler.get();
  if (s) std::memcpy(out, s->SplineSegments, size_t(s->nParams) * sizeof(short));
}

// ---- the adapter over the real class (only in the build that links libm3b200: libm3ref_path_lm_b200.so) ----------
REFP_API int refp_fd_move_to_b200_ex(void* p, int n_devices, const int* devices);
REFP_API int refp_fd_move_to_b200(void* p, int device) { return refp_fd_move_to_b200_ex(p, 1, &device); }
// devices: one entry = the whole sample on that B200; several = m3b_group_* (one process, one calling thread)
REFP_API int refp_fd_move_to_b200_ex(void* p, int n_devices, const int* devices) {
#ifdef M3B_WITH_ADAPTER
  FD* fd = static_cast<FD*>(p);
  FDType* b = static_cast<FDType*>(fd);
  m3b200::MonolithArrays a;
  std::vector<int16_t> n_pts;
#ifdef _LOW_MEMORY_STRUCTS_          // the event-by-event monolith is wired into SampleHandlerFD in this build only
  SMonolith* m = dynamic_cast<SMonolith*>(fd->SplineHandler.get());
  if (m) {
    n_pts.resize(size_t(m->nParams));
    for (int i = 0; i < m->nParams; ++i) {
      n_pts[i] = int16_t(m->SplineInfoArray[i].xPts.empty() ? 0 : m->SplineInfoArray[i].nPts);
      a.spline_par_pointers.push_back(m->SplineInfoArray[i].splineParsPointer);
    }
    a.n_params = m->nParams; a.max_knots = m->_max_knots;
    a.coeff_x = m->cpu_spline_handler->coeff_x.data(); a.n_pts = n_pts.data();
    a.nParamPerEvent = m->cpu_nParamPerEvent.data(); a.paramNo_arr = m->cpu_spline_handler->paramNo_arr.data();
    a.nKnots_arr = m->cpu_spline_handler->nKnots_arr.data(); a.total_knots = uint32_t(m->cpu_spline_handler->coeff_many.size() / 4);
    a.coeff_many = m->cpu_spline_handler->coeff_many.data(); a.nParamPerEvent_tf1 = m->cpu_nParamPerEvent_tf1.data();
    a.paramNo_tf1 = m->cpu_paramNo_TF1_arr.data(); a.coeff_tf1 = m->cpu_coeff_TF1_many.data();
    a.cpu_total_weights = m->cpu_total_weights;
  }
#endif
  m3b200::PointerBasesT<M3::float_t> pb;
  pb.norm_base = fd->norm.data(); pb.n_norm = int(fd->norm.size());
  pb.osc_base = fd->pool.data(); pb.n_osc = int64_t(fd->nEvents);       // pool = [osc per event | extra weights]
  pb.zero = &M3::Zero; pb.unity = &M3::Unity;
  pb.constant_weight_ranges.push_back({fd->pool.data() + fd->nEvents, fd->pool.data() + fd->pool.size()});   // the extras
  // the binned arm (Samples/SampleHandlerFD.cpp:1196-1242): BinnedSplineHandler's monolith arrays as they stand
  // (either M3::float_t build: the adapter's types follow it)
  m3b200::BinnedArraysT<M3::float_t> ba;
  std::vector<M3::float_t> knot_x; std::vector<int16_t> bn_pts;
  std::vector<int32_t> usv, civ, uci;
  if (Binned* bs = dynamic_cast<Binned*>(fd->SplineHandler.get())) {
    int K = 0;
    for (int i = 0; i < bs->nParams; ++i) K = std::max(K, int(bs->SplineInfoArray[i].nPts));
    knot_x.assign(size_t(bs->nParams) * K, M3::float_t(0)); bn_pts.resize(size_t(bs->nParams));
    for (int i = 0; i < bs->nParams; ++i) {
      bn_pts[i] = int16_t(bs->SplineInfoArray[i].nPts);
      for (int k = 0; k < bn_pts[i]; ++k) knot_x[size_t(i) * K + k] = bs->SplineInfoArray[i].xPts[k];
      ba.spline_par_pointers.push_back(bs->SplineInfoArray[i].splineParsPointer);
    }
    usv.assign(bs->uniquesplinevec_Monolith.begin(), bs->uniquesplinevec_Monolith.end());
    civ.assign(bs->coeffindexvec.begin(), bs->coeffindexvec.end());
    uci.assign(bs->uniquecoeffindices.begin(), bs->uniquecoeffindices.end());
    ba.n_params = bs->nParams; ba.max_knots = K; ba.knot_x = knot_x.data(); ba.n_pts = bn_pts.data();
    ba.n_slots = int64_t(bs->weightvec_Monolith.size()); ba.uniquesplinevec_Monolith = usv.data(); ba.coeffindexvec = civ.data();
    ba.n_unique = int64_t(uci.size()); ba.uniquecoeffindices = uci.data(); ba.n_coeff = bs->n_coeff_keep;
    ba.manycoeff_arr = bs->manycoeff_arr; ba.xcoeff_arr = bs->xcoeff_arr; ba.weightvec_Monolith = bs->weightvec_Monolith.data();
  }
  try { b->MoveToB200(a, pb, std::vector<int>(devices, devices + n_devices), ba); }
  catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
  return 0;
#else
  (void)p; (void)n_devices; (void)devices;
  return -1;
#endif
}
REFP_API int refp_fd_data_changed(void* p) {
#ifdef M3B_WITH_ADAPTER
  try { static_cast<FDType*>(static_cast<FD*>(p))->DataChanged(); } catch (...) { return 1; }
  return 0;
#else
  (void)p; return -1;
#endif
}
REFP_API int refp_fd_sync_host_arrays(void* p) {
#ifdef M3B_WITH_ADAPTER
  try { static_cast<FDType*>(static_cast<FD*>(p))->SyncHostArrays(); } catch (...) { return 1; }
  return 0;
#else
  (void)p; return -1;
#endif
}
