// mock_mach3.h -- TEST INFRASTRUCTURE.  A minimal stand-in for the parts of MaCh3 that
// adapters/SampleHandlerB200.h touches, with the reference's member names and pointer wiring
// (EventInfo: Samples/FarDetectorCoreInfoStruct.h:82-126; SampleHandlerFD members:
// Samples/SampleHandlerFD.h:214-217,335-398; BinningHandler accessors: Samples/BinningHandler.h:160-180),
// and an independent, scalar CPU implementation of the hot path written from the formulas in
// SURVEY.md Appendix A -- so the adapter can be compiled and run without ROOT, and its result checked
// against a CPU path that consumes the SAME pointer soup the real SampleHandlerFD would.
// Nothing here is part of the product.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <array>
#include <vector>

namespace M3 {
using float_t = float;
using int_t = short;
static const float_t Unity = 1.0f;
static const float_t Zero = 0.0f;
constexpr double _LOW_MC_BOUND_ = .00001;
}  // namespace M3

enum TestStatistic { kPoisson = 0, kBarlowBeeston = 1, kIceCube = 2, kPearson = 3, kDembinskiAbdelmotteleb = 4 };

struct KinematicCut {          // Samples/SampleStructs.h:149-157
  int ParamToCutOnIt = -999;
  double LowerBound = -999, UpperBound = -999;
};
struct FunctionalShifter {};

struct EventInfo {
  std::vector<const double*> norm_pointers;
  std::vector<const M3::float_t*> total_weight_pointers;
  std::vector<const double*> KinVar;
  std::vector<int> NomBin;
  int NominalSample = -1;
  bool isNC = false;
};

class BinningHandler {
 public:
  std::vector<std::vector<std::vector<double>>> edges;   // [sample][dim][edge]
  std::vector<int> offset;
  int total = 0;
  void Finalise() {
    offset.clear(); total = 0;
    for (auto& s : edges) { offset.push_back(total); int n = 1; for (auto& d : s) n *= int(d.size()) - 1; total += n; }
  }
  int GetNDim(const int s) const { return int(edges[s].size()); }
  std::vector<double> GetBinEdges(const int s, const int d) const { return edges[s].at(d); }
  // the mock only does uniform binning (the real class, with non-uniform samples, is exercised through
  // oracle/_ref/libm3ref_path_lm_b200.so: tests/test_adapter_gpu.py)
  bool IsUniform(const int) const { return true; }
  struct BinInfo { std::vector<std::array<double, 2>> Extent; };
  std::vector<BinInfo> GetNonUniformBins(const int) const { return {}; }
  int GetNBins() const { return total; }
  int GetSampleStartBin(const int s) const { return offset[s]; }
  int GetSampleEndBin(const int s) const { return s + 1 < int(offset.size()) ? offset[s + 1] : total; }
  int FindGlobalBin(const int s, const std::vector<const double*>& kin) const {   // BinningHandler.cpp:257-277
    int g = 0, stride = 1;
    for (size_t d = 0; d < edges[s].size(); ++d) {
      const auto& e = edges[s][d];
      const double x = *kin[d];
      if (x < e.front() || x >= e.back()) return -1;
      const int b = int(std::upper_bound(e.begin(), e.end(), x) - e.begin()) - 1;
      g += b * stride;
      stride *= int(e.size()) - 1;
    }
    return g + offset[s];
  }
};

class OscillationHandler {     // NuOscillator stand-in: Evaluate() refreshes the weight array in place
 public:
  std::vector<M3::float_t> weights;
  std::vector<std::vector<M3::float_t>> per_step;
  size_t next = 0;
  void Evaluate() { if (!per_step.empty()) { weights = per_step[next % per_step.size()]; ++next; } }
};

class SplineBase {
 public:
  virtual ~SplineBase() {}
  virtual void Evaluate() = 0;
  virtual void SynchroniseMemTransfer() const {}
};

// event-by-event monolith on the CPU, reference layout (Splines/SplineMonolith.cpp:727-830)
class SMonolith : public SplineBase {
 public:
  int nParams = 0, max_knots = 0;
  std::vector<float> coeff_x, coeff_many, coeff_tf1;
  std::vector<int16_t> n_pts, paramNo_arr, paramNo_tf1;
  std::vector<uint32_t> nKnots_arr, nParamPerEvent, nParamPerEvent_tf1;
  std::vector<const double*> pars;
  std::vector<int16_t> segments, curr;
  std::vector<float> vals, cpu_total_weights_v;
  float* cpu_total_weights = nullptr;
  void Finalise(size_t n_events) {
    segments.assign(nParams, 0); curr.assign(nParams, 0); vals.assign(nParams, 0.f);
    cpu_total_weights_v.assign(n_events, 1.f); cpu_total_weights = cpu_total_weights_v.data();
  }
  const float* retPointer(const int e) const { return &cpu_total_weights[e]; }
  void FindSplineSegment() {                               // Splines/SplineBase.cpp:44-109
    for (int i = 0; i < nParams; ++i) {
      const int n = n_pts[i];
      const float* x = coeff_x.data() + size_t(i) * max_knots;
      const float xvar = float(*pars[i]);
      vals[i] = xvar;
      if (n == 0) continue;
      int seg = 0, hi = n - 1;
      if (xvar <= x[0]) seg = 0;
      else if (xvar >= x[n - 1]) seg = hi;
      else if (x[curr[i] + 1] > xvar && xvar >= x[curr[i]]) seg = curr[i];
      else while (hi - seg > 1) { const int half = (seg + hi) / 2; if (xvar > x[half]) seg = half; else hi = half; }
      if (seg >= n - 1 && n > 1) seg = n - 2;
      curr[i] = segments[i] = int16_t(seg);
    }
  }
  void Evaluate() override {
    FindSplineSegment();
    const size_t E = cpu_total_weights_v.size();
    for (size_t e = 0; e < E; ++e) {
      float w = 1.0f;
      for (uint32_t k = 0; k < nParamPerEvent[2 * e]; ++k) {
        const uint32_t s = nParamPerEvent[2 * e + 1] + k;
        const int p = paramNo_arr[s];
        const int seg = segments[p];
        const float* c = coeff_many.data() + size_t(nKnots_arr[s]) * 4 + size_t(seg) * 4;
        const float dx = vals[p] - coeff_x[size_t(p) * max_knots + seg];
        w *= fmaf(dx, fmaf(dx, fmaf(dx, c[3], c[2]), c[1]), c[0]);
      }
      for (uint32_t k = 0; k < nParamPerEvent_tf1[2 * e]; ++k) {
        const uint32_t s = nParamPerEvent_tf1[2 * e + 1] + k;
        w *= fmaf(coeff_tf1[2 * s], vals[paramNo_tf1[s]], coeff_tf1[2 * s + 1]);
      }
      cpu_total_weights[e] = w;
    }
  }
};

class SampleHandlerBase {
 public:
  virtual ~SampleHandlerBase() {}
  virtual void Reweight() = 0;
  virtual double GetLikelihood() const = 0;
  virtual double GetSampleLikelihood(const int isample) const = 0;
  virtual M3::int_t GetNsamples() { return nSamples; }
  unsigned int GetNEvents() const { return nEvents; }
  double GetTestStatLLH(const double data, const double mc, const double w2) const {   // Poisson + Barlow-Beeston
    if (fTestStatistic == kBarlowBeeston) {
      double newmc = mc;
      if (mc < M3::_LOW_MC_BOUND_) { if (data > M3::_LOW_MC_BOUND_) newmc = M3::_LOW_MC_BOUND_; else if (data >= mc) return 0.; }
      const double f = std::sqrt(w2) / newmc, f2 = f * f, t = newmc * f2 - 1, t2 = t * t + 4 * data * f2;
      const double beta = (-t + std::sqrt(t2)) / 2.;
      double stat = mc * beta;
      if (data > 0) { newmc *= beta; stat = newmc - data + data * std::log(data / newmc); }
      return stat + (f > 0 ? (beta - 1) * (beta - 1) / (2 * f2) : 0.);
    }
    if (data == 0) return mc;
    if (mc < M3::_LOW_MC_BOUND_) {
      if (data > M3::_LOW_MC_BOUND_) return M3::_LOW_MC_BOUND_ - data + data * std::log(data / M3::_LOW_MC_BOUND_);
      else if (data >= mc) return 0.;
    }
    return mc - data + data * std::log(data / mc);
  }
 protected:
  TestStatistic fTestStatistic = kPoisson;
  M3::int_t nSamples = 1;
  unsigned int nEvents = 0;
};

class SampleHandlerFD : public SampleHandlerBase {
 public:
  const BinningHandler* GetBinningHandler() const { return Binning.get(); }
  void Reweight() override {                               // Samples/SampleHandlerFD.cpp:316-343 (serial FillArray)
    std::fill(SampleHandlerFD_array.begin(), SampleHandlerFD_array.end(), 0.);
    if (FirstTimeW2) std::fill(SampleHandlerFD_array_w2.begin(), SampleHandlerFD_array_w2.end(), 0.);
    if (Oscillator) Oscillator->Evaluate();
    if (SplineHandler) SplineHandler->Evaluate();
    Selection = StoredSelection;                           // :355
    for (unsigned int e = 0; e < GetNEvents(); ++e) {
      const EventInfo& ev = MCSamples[e];
      if (!IsEventSelected(ev.NominalSample, int(e))) continue;          // :361
      M3::float_t w = 1.0f;                                // CalcWeightTotal :568-594
      for (const double* p : ev.norm_pointers) w *= static_cast<M3::float_t>(*p);
      for (const M3::float_t* p : ev.total_weight_pointers) w *= *p;
      if (w <= 0.) continue;
      const int bin = Binning->FindGlobalBin(ev.NominalSample, ev.KinVar);
      if (bin > -1) { SampleHandlerFD_array[bin] += w; if (FirstTimeW2) SampleHandlerFD_array_w2[bin] += w * w; }
    }
    if (!UpdateW2) FirstTimeW2 = false;
  }
  double GetLikelihood() const override {
    double l = 0;
    for (int b = 0; b < Binning->GetNBins(); ++b) l += GetTestStatLLH(SampleHandlerFD_data[b], SampleHandlerFD_array[b], SampleHandlerFD_array_w2[b]);
    return l;
  }
  double GetSampleLikelihood(const int s) const override {
    double l = 0;
    for (int b = Binning->GetSampleStartBin(s); b < Binning->GetSampleEndBin(s); ++b)
      l += GetTestStatLLH(SampleHandlerFD_data[b], SampleHandlerFD_array[b], SampleHandlerFD_array_w2[b]);
    return l;
  }
  virtual double ReturnKinematicParameter(int, int) { return 0.0; }     // pure virtual in the reference (SampleHandlerFD.h:293)
  bool IsEventSelected(const int iSample, const int iEvent) {           // Samples/SampleHandlerFD.cpp:281-294
    if (size_t(iSample) >= Selection.size()) return true;
    for (const auto& Cut : Selection[size_t(iSample)]) {
      const double Val = ReturnKinematicParameter(Cut.ParamToCutOnIt, iEvent);
      if ((Val < Cut.LowerBound) || (Val >= Cut.UpperBound)) return false;
    }
    return true;
  }
 protected:
  std::vector<std::vector<KinematicCut>> StoredSelection, Selection;
  std::vector<std::vector<FunctionalShifter*>> funcParsGrid;
  std::unique_ptr<SplineBase> SplineHandler;
  std::shared_ptr<OscillationHandler> Oscillator;
  std::unique_ptr<BinningHandler> Binning;
  std::vector<double> SampleHandlerFD_array, SampleHandlerFD_array_w2, SampleHandlerFD_data;
  std::vector<EventInfo> MCSamples;
  bool FirstTimeW2 = true, UpdateW2 = false;
};
