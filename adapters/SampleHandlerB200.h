// SampleHandlerB200.h -- the fused B200 path behind MaCh3's SampleHandlerBase virtuals.
//
//   template <class FDBase> class SampleHandlerB200 : public FDBase
//
// FDBase is the experiment's concrete SampleHandlerFD subclass (the class that implements SetupSplines,
// SetupFDMC, ... -- Samples/SampleHandlerFD.h:189-304).  The adapter overrides exactly the three
// virtuals every fitter calls (Samples/SampleHandlerBase.h:37-90; Fitters/MR2T2.cpp:62-74):
//
//     void   Reweight()                       Samples/SampleHandlerFD.cpp:316-343
//     double GetLikelihood() const            :1284-1300     (returns -lnL, like the reference)
//     double GetSampleLikelihood(int) const   :1262-1281
//
// and turns the per-event pointer soup the reference builds in Initialise()
// (Samples/SampleHandlerFD.cpp:169-202; EventInfo, Samples/FarDetectorCoreInfoStruct.h:82-126) into the
// flat index tables of libm3b200 (include/m3b200.h) ONCE, in MoveToB200():
//
//     EventInfo::norm_pointers[j]          -> norm_idx = ptr - <base of ParameterHandlerBase::_fPropVal>
//     EventInfo::total_weight_pointers[k]  -> classified by address:
//           inside SMonolith::cpu_total_weights       the event's spline weight: produced on the device, dropped
//           inside BinnedSplineHandler::              slot index = ptr - weightvec_Monolith.data(): the binned arm
//             weightvec_Monolith                      (SetSplinePointers, Samples/SampleHandlerFD.cpp:1196-1242)
//           inside the oscillator's weight array      osc_idx = ptr - base          (SampleHandlerFD.cpp:1108-1122)
//           &M3::Zero / &M3::Unity                    static factor 0 / 1           (:1128-1131)
//           inside PointerBases::constant_weight_ranges   read now, folded into the event's static weight
//           anything else                             ERROR (a weight the device copy would never see change)
//     EventInfo::KinVar[d]                 -> kin[d][e] = *ptr  (bins are found on the device; functional
//                                             "shift" parameters, :545-564, and CalcWeightFunc overrides, :428, are
//                                             NOT supported by this adapter: it throws if funcParsGrid is non-empty)
//     EventInfo::NominalSample             -> sample_id
//     StoredSelection (KinematicCut lists, -> m3b_upload_selection: IsEventSelected (:281-294) runs on the device; the
//       Samples/SampleHandlerFD.h:364-371)    cut variables are read once through ReturnKinematicParameter(var, event)
//     BinningHandler (GetNDim/GetBinEdges/ -> m3b_upload_binning_ex         (uniform and non-uniform samples)
//       IsUniform/GetNonUniformBins)
//     SplineMonoStruct + SMonolith arrays  -> m3b_upload_spline_monolith    (MonolithArrays)
//     BinnedSplineHandler arrays           -> m3b_upload_binned_splines + m3b_upload_event_binned_splines (BinnedArrays)
//
// Devices: MoveToB200(..., {d}) puts the whole sample on one B200 (one fused launch per step);
// MoveToB200(..., {0,1,...,7}) spreads the events over several B200s of the box through m3b_group_* -- still ONE
// process and ONE calling thread, the fitters do not change.
//
// Per step Reweight() = [Oscillator->Evaluate()] + ONE m3b_step / m3b_group_step call: FindSplineSegment (host, the
// reference's own history-dependent rule) -> fused evaluate/product/fill/-lnL kernel.  The oscillation weights:
//     * the oscillator's array is REGISTERED with CUDA once (PointerBases::register_osc_array, default) and handed to
//       the library as it stands: the fill kernel streams it over PCIe (events in array order) or the library issues
//       one DMA copy (indexed events) -- no host-side copy, no per-step synchronisation;
//     * if the registration is refused, or register_osc_array is false, the array is copied into a pinned staging
//       buffer every step (the pre-round-2 behaviour: ~0.1 ms per MB on one host thread).
// Nothing but the -lnL scalar comes back; SampleHandlerFD_array / _array_w2 are refreshed lazily by SyncHostArrays()
// (call it before GetMCArray()/PrintRates()/plotting).  Before MoveToB200() every virtual forwards to the reference's
// own CPU implementation (FDBase::Reweight etc. -- e.g. for the Asimov pass of a set-up that moves later).
//
// The template keeps this header compilable against the real MaCh3 (needs ROOT; not available in this
// repository's build image) and against the mock in tests/adapters/mock_mach3.h, which the repository's
// own test builds and runs on the B200 (tests/test_adapter_gpu.py).
#pragma once
#include "m3b200.h"

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <type_traits>
#include <string>
#include <utility>
#include <vector>

namespace m3b200 {

// What the adapter needs from the experiment side that SampleHandlerFD keeps private or spreads over
// several objects.  All pointers are the reference's own arrays; the library copies on upload.
struct MonolithArrays {               // Splines/SplineCommon.h:30-50, Splines/SplineMonolith.h:96-137
  int n_params = 0, max_knots = 0;
  const float* coeff_x = nullptr;             // SplineMonoStruct::coeff_x           [n_params*max_knots]
  const int16_t* n_pts = nullptr;             // FastSplineInfo::nPts per parameter  [n_params]
  const uint32_t* nParamPerEvent = nullptr;   // SMonolith::cpu_nParamPerEvent       [2*n_events]
  const int16_t* paramNo_arr = nullptr;       // SplineMonoStruct::paramNo_arr
  const uint32_t* nKnots_arr = nullptr;       // SplineMonoStruct::nKnots_arr
  uint32_t total_knots = 0;
  const float* coeff_many = nullptr;          // SplineMonoStruct::coeff_many        [total_knots*4]
  const uint32_t* nParamPerEvent_tf1 = nullptr;
  const int16_t* paramNo_tf1 = nullptr;       // SMonolith::cpu_paramNo_TF1_arr
  const float* coeff_tf1 = nullptr;           // SMonolith::cpu_coeff_TF1_many       [n_tf1*2]
  std::vector<const double*> spline_par_pointers;   // what SMonolith::setSplinePointers receives
  const float* cpu_total_weights = nullptr;   // SMonolith::cpu_total_weights (address range only; never read)
};

// The other SplineBase implementation (Splines/BinnedSplineHandler.h:110-135): its monolith arrays as they stand after
// TransferToMonolith.  R = M3::float_t of the MaCh3 build: float with _LOW_MEMORY_STRUCTS_, double by default.
template <class R>
struct BinnedArraysT {
  int n_params = 0, max_knots = 0;
  const R* knot_x = nullptr;                  // FastSplineInfo::xPts per parameter, rows padded to max_knots
  const int16_t* n_pts = nullptr;             // FastSplineInfo::nPts                [n_params]
  int64_t n_slots = 0;                        // weightvec_Monolith.size()
  const int32_t* uniquesplinevec_Monolith = nullptr;   // [n_slots]
  const int32_t* coeffindexvec = nullptr;              // [n_slots]
  int64_t n_unique = 0;
  const int32_t* uniquecoeffindices = nullptr;         // [n_unique]
  int64_t n_coeff = 0;
  const R* manycoeff_arr = nullptr;           // [n_coeff*4]
  const R* xcoeff_arr = nullptr;              // [n_coeff]
  std::vector<const double*> spline_par_pointers;      // FastSplineInfo::splineParsPointer per parameter
  const R* weightvec_Monolith = nullptr;      // address range of the slots (never read)
};
using BinnedArrays = BinnedArraysT<float>;    // _LOW_MEMORY_STRUCTS_ build
using BinnedArraysD = BinnedArraysT<double>;  // default build

template <class R>
struct PointerBasesT {
  const double* norm_base = nullptr;  int n_norm = 0;        // ParameterHandlerBase::_fPropVal.data(), size
  const R* osc_base = nullptr;        int64_t n_osc = 0;     // the oscillator's weight array (may be null)
  const R* zero = nullptr;                                   // &M3::Zero
  const R* unity = nullptr;                                  // &M3::Unity
  // Address ranges of experiment-specific weights (AddAdditionalWeightPointers) that never change during a fit, e.g. a
  // per-event flux or POT weight array: {first, one-past-last}.  A weight pointer that is none of the above and lies
  // in none of these ranges is an ERROR (it might be rewritten every step, which the device copy would not see).
  std::vector<std::pair<const R*, const R*>> constant_weight_ranges;
  // page-lock the oscillator's weight array where it lies (cudaHostRegister) so the device reads it without a host copy
  // (float build; the double build's array is copied to the device every step)
  bool register_osc_array = true;
};
using PointerBases = PointerBasesT<float>;    // _LOW_MEMORY_STRUCTS_ build
using PointerBasesD = PointerBasesT<double>;  // default build

template <class FDBase>
class SampleHandlerB200 : public FDBase {
 public:
  using FDBase::FDBase;
  ~SampleHandlerB200() override {
    if (g_) m3b_group_destroy(g_);
    else if (h_) m3b_destroy(h_);
  }

  // Call once after the base class finished Initialise() (events, binning, splines, pointers wired).
  // Exactly one of mono.n_params / binned.n_params may be non-zero (SampleHandlerFD holds one SplineBase).
  // R is the build's M3::float_t (deduced from the arguments): float = the _LOW_MEMORY_STRUCTS_ build, the only one in
  // which SMonolith is wired into SampleHandlerFD (Samples/SampleHandlerFD.cpp:1244-1254); double = the default build,
  // BinnedSplineHandler only, weights and products in double like the reference's (m3b_upload_binned_splines_f64).
  template <class R>
  void MoveToB200(const MonolithArrays& mono, const PointerBasesT<R>& bases, const std::vector<int>& cuda_devices = {0},
                  const BinnedArraysT<R>& binned = BinnedArraysT<R>()) {
    constexpr bool kF64 = std::is_same<R, double>::value;
    static_assert(kF64 || std::is_same<R, float>::value, "M3::float_t is float or double");
    using WeightPtr = typename std::decay<decltype(this->MCSamples[0].total_weight_pointers[0])>::type;
    static_assert(std::is_same<WeightPtr, const R*>::value, "PointerBases / BinnedArrays of the other M3::float_t build");
    if (kF64 && mono.n_params > 0) throw std::runtime_error("SampleHandlerB200: the spline monolith exists only in the _LOW_MEMORY_STRUCTS_ build");
    if (cuda_devices.empty() || cuda_devices.size() > 8) throw std::runtime_error("SampleHandlerB200: 1..8 devices");
    if (mono.n_params > 0 && binned.n_params > 0) throw std::runtime_error("SampleHandlerB200: monolith OR binned splines, not both");
    if (binned.n_params > 0 && cuda_devices.size() > 1) throw std::runtime_error("SampleHandlerB200: the binned-spline arm runs on one device");
    m3b_config cfg{};
    cfg.device = cuda_devices[0];
    cfg.test_statistic = static_cast<int32_t>(this->fTestStatistic);   // enum TestStatistic == m3b_test_statistic
    cfg.update_w2 = this->UpdateW2 ? 1 : 0;
    if (cuda_devices.size() == 1) {
      check(m3b_create(&cfg, &h_), "m3b_create");
    } else {
      std::vector<int32_t> dev(cuda_devices.begin(), cuda_devices.end());
      if (m3b_group_create(&cfg, dev.data(), static_cast<int32_t>(dev.size()), &g_) != M3B_OK)
        throw std::runtime_error(std::string("SampleHandlerB200: m3b_group_create: ") + m3b_last_error(nullptr));
      h_ = m3b_group_member(g_, 0);
    }
    const int64_t E = static_cast<int64_t>(this->GetNEvents());
    // functional ("shift") parameters rewrite the kinematics through std::functions every step (ApplyShifts,
    // Samples/SampleHandlerFD.cpp:545-564): not something a constant device table can follow -- refuse, loudly
    for (const auto& shifts : this->funcParsGrid)
      if (!shifts.empty()) throw std::runtime_error("SampleHandlerB200: functional (shift) parameters are not supported by this adapter");
    norm_base_ = bases.norm_base; n_norm_ = bases.n_norm; osc_base_ = bases.osc_base; n_osc_ = bases.n_osc; f64_ = kF64;
    spline_ptrs_ = mono.n_params > 0 ? mono.spline_par_pointers : binned.spline_par_pointers;
    spline_vals_.assign(spline_ptrs_.size(), 0.0);

    // --- binning (BinningHandler, Samples/BinningHandler.cpp:257-291): uniform samples hand over their axis edges,
    //     non-uniform ones (Samples/SampleStructs.h:468-528) their boxes (BinInfo::Extent); the library rebuilds the
    //     mega-bin grid exactly as InitialiseGridMapping does
    const int nS = static_cast<int>(this->GetNsamples());
    std::vector<int32_t> ndim(nS), uniform(nS, 1), nbins(static_cast<size_t>(nS) * 4, 0);
    std::vector<double> edges;
    int max_dim = 0;
    for (int s = 0; s < nS; ++s) {
      ndim[s] = this->GetBinningHandler()->GetNDim(s);
      if (ndim[s] < 1 || ndim[s] > 4) throw std::runtime_error("SampleHandlerB200: 1..4 binning dimensions per sample");
      max_dim = ndim[s] > max_dim ? ndim[s] : max_dim;
      if (this->GetBinningHandler()->IsUniform(s)) {
        for (int d = 0; d < ndim[s]; ++d) {
          const std::vector<double> e = this->GetBinningHandler()->GetBinEdges(s, d);
          nbins[static_cast<size_t>(s) * 4 + d] = static_cast<int32_t>(e.size()) - 1;
          edges.insert(edges.end(), e.begin(), e.end());
        }
      } else {
        uniform[s] = 0;
        const auto boxes = this->GetBinningHandler()->GetNonUniformBins(s);
        nbins[static_cast<size_t>(s) * 4] = static_cast<int32_t>(boxes.size());
        for (const auto& b : boxes)
          for (int d = 0; d < ndim[s]; ++d) { edges.push_back(b.Extent[d][0]); edges.push_back(b.Extent[d][1]); }
      }
    }
    if (g_) gcheck(m3b_group_upload_binning_ex(g_, nS, ndim.data(), uniform.data(), nbins.data(), edges.data()), "m3b_group_upload_binning_ex");
    else check(m3b_upload_binning_ex(h_, nS, ndim.data(), uniform.data(), nbins.data(), edges.data()), "m3b_upload_binning_ex");
    n_bins_ = this->GetBinningHandler()->GetNBins();

    // --- the spline handler's arrays, as they are
    if (mono.n_params > 0) {
      if (g_) gcheck(m3b_group_upload_spline_monolith(g_, mono.n_params, mono.max_knots, mono.coeff_x, mono.n_pts, E, mono.nParamPerEvent,
                                                      mono.paramNo_arr, mono.nKnots_arr, mono.total_knots, mono.coeff_many,
                                                      mono.nParamPerEvent_tf1, mono.paramNo_tf1, mono.coeff_tf1), "m3b_group_upload_spline_monolith");
      else check(m3b_upload_spline_monolith(h_, mono.n_params, mono.max_knots, mono.coeff_x, mono.n_pts, E, mono.nParamPerEvent,
                                            mono.paramNo_arr, mono.nKnots_arr, mono.total_knots, mono.coeff_many,
                                            mono.nParamPerEvent_tf1, mono.paramNo_tf1, mono.coeff_tf1), "m3b_upload_spline_monolith");
    } else if (binned.n_params > 0) {
      check(upload_binned(binned), kF64 ? "m3b_upload_binned_splines_f64" : "m3b_upload_binned_splines");
    }

    // --- events: pointers -> indices
    size_t max_norm = 0;
    for (int64_t e = 0; e < E; ++e) max_norm = std::max(max_norm, this->MCSamples[e].norm_pointers.size());
    if (max_norm > 16) throw std::runtime_error("SampleHandlerB200: more than 16 norm pointers on one event");
    std::vector<int32_t> sample_id(E), osc_idx(E, -1);
    std::vector<double> kin(static_cast<size_t>(max_dim) * E, 0.0);
    std::vector<int16_t> norm_idx(static_cast<size_t>(max_norm) * E, -1);
    std::vector<R> static_w(E, R(1));
    std::vector<uint32_t> n_binned(binned.n_params > 0 ? E : 0, 0);
    std::vector<int32_t> binned_slot;
    bool any_osc = false, identity = true;
    for (int64_t e = 0; e < E; ++e) {
      const auto& ev = this->MCSamples[e];
      sample_id[e] = ev.NominalSample;
      for (size_t d = 0; d < ev.KinVar.size() && d < static_cast<size_t>(max_dim); ++d) kin[d * E + e] = *ev.KinVar[d];
      for (size_t j = 0; j < ev.norm_pointers.size(); ++j) {
        const std::ptrdiff_t off = ev.norm_pointers[j] - bases.norm_base;
        if (off < 0 || off >= bases.n_norm) throw std::runtime_error("SampleHandlerB200: norm pointer outside the parameter array");
        norm_idx[e * max_norm + j] = static_cast<int16_t>(off);
      }
      for (const auto* p : ev.total_weight_pointers) {
        if (is_monolith_weight(p, mono, E)) {
          if (monolith_index(p, mono) != e) throw std::runtime_error("SampleHandlerB200: event points at another event's spline weight");
        } else if (binned.weightvec_Monolith && p >= binned.weightvec_Monolith && p < binned.weightvec_Monolith + binned.n_slots) {
          binned_slot.push_back(static_cast<int32_t>(p - binned.weightvec_Monolith));      // pointer order kept (:1236-1242)
          ++n_binned[e];
        } else if (bases.osc_base && p >= bases.osc_base && p < bases.osc_base + bases.n_osc) {
          if (osc_idx[e] >= 0) throw std::runtime_error("SampleHandlerB200: two oscillation weights on one event");
          osc_idx[e] = static_cast<int32_t>(p - bases.osc_base);
          any_osc = true;
        } else if (p == bases.zero) {
          static_w[e] = R(0);                 // NC event with flavour change (SampleHandlerFD.cpp:1128-1131)
        } else if (p != bases.unity) {
          bool constant = false;
          for (const auto& r : bases.constant_weight_ranges) constant |= (p >= r.first && p < r.second);
          if (!constant)
            throw std::runtime_error("SampleHandlerB200: event " + std::to_string(e) + " has a weight pointer that is neither the "
                                     "oscillation array, the spline handler, M3::Zero/Unity nor inside PointerBases::constant_weight_ranges");
          static_w[e] *= *p;                  // experiment-specific constant weight, folded once
        }
      }
      identity &= (osc_idx[e] == e);
    }
    // events in the oscillator array's own order (one weight per event): no index table, the kernel streams the array
    osc_direct_ = any_osc && identity && bases.n_osc == E;
    n_osc_dev_ = any_osc ? bases.n_osc : 0;
    const int32_t* oi = (any_osc && !osc_direct_) ? osc_idx.data() : nullptr;       // -1 entries: no oscillation weight (reads 1.0)
    // (double build: the float entry point fixes the shapes, the double static weights follow)
    if (g_) gcheck(m3b_group_upload_events(g_, E, sample_id.data(), kin.data(), static_cast<int32_t>(max_norm), max_norm ? norm_idx.data() : nullptr,
                                           bases.n_norm, any_osc ? 1 : 0, oi, n_osc_dev_, float_or_null(static_w)), "m3b_group_upload_events");
    else check(m3b_upload_events(h_, E, sample_id.data(), kin.data(), static_cast<int32_t>(max_norm), max_norm ? norm_idx.data() : nullptr,
                                 bases.n_norm, any_osc ? 1 : 0, oi, n_osc_dev_, float_or_null(static_w)), "m3b_upload_events");
    if (binned.n_params > 0)
      check(m3b_upload_event_binned_splines(h_, E, n_binned.data(), binned_slot.data()), "m3b_upload_event_binned_splines");
    if (kF64) check(upload_static_f64(E, static_w), "m3b_upload_event_weights_f64");
    if (any_osc && !kF64) {
      // the oscillator's array, page-locked where it lies; else a pinned staging copy refreshed every step
      if (bases.register_osc_array &&
          m3b_register_host_buffer(h_, const_cast<void*>(osc_base_), sizeof(float) * static_cast<uint64_t>(bases.n_osc)) == M3B_OK) {
        osc_registered_ = true;
      } else {
        void* p = nullptr;
        if (g_) gcheck(m3b_group_alloc_host(g_, sizeof(float) * static_cast<size_t>(n_osc_dev_), &p), "m3b_group_alloc_host");
        else check(m3b_alloc_host(h_, sizeof(float) * static_cast<size_t>(n_osc_dev_), &p), "m3b_alloc_host");
        osc_stage_ = static_cast<float*>(p);
        std::fill(osc_stage_, osc_stage_ + n_osc_dev_, 1.0f);
      }
    }
    // --- selection cuts (Samples/SampleHandlerFD.cpp:281-294): one table row per distinct ParamToCutOnIt, filled
    //     through the experiment's own ReturnKinematicParameter; without functional shifts the values are constants
    {
      std::vector<int32_t> cut_sample, cut_var, distinct;
      std::vector<double> lo, hi;
      for (int s = 0; s < nS && s < static_cast<int>(this->StoredSelection.size()); ++s)
        for (const auto& c : this->StoredSelection[static_cast<size_t>(s)]) {
          size_t v = 0;
          while (v < distinct.size() && distinct[v] != c.ParamToCutOnIt) ++v;
          if (v == distinct.size()) distinct.push_back(c.ParamToCutOnIt);
          cut_sample.push_back(s); cut_var.push_back(static_cast<int32_t>(v)); lo.push_back(c.LowerBound); hi.push_back(c.UpperBound);
        }
      if (!cut_sample.empty()) {
        std::vector<double> values(distinct.size() * static_cast<size_t>(E));
        for (size_t v = 0; v < distinct.size(); ++v)
          for (int64_t e = 0; e < E; ++e) values[v * static_cast<size_t>(E) + e] = this->ReturnKinematicParameter(distinct[v], static_cast<int>(e));
        const int32_t nc = static_cast<int32_t>(cut_sample.size()), nv = static_cast<int32_t>(distinct.size());
        if (g_) gcheck(m3b_group_upload_selection(g_, nc, cut_sample.data(), cut_var.data(), lo.data(), hi.data(), nv, values.data()), "m3b_group_upload_selection");
        else check(m3b_upload_selection(h_, nc, cut_sample.data(), cut_var.data(), lo.data(), hi.data(), nv, values.data()), "m3b_upload_selection");
      }
    }
    if (g_) {
      gcheck(m3b_group_upload_data(g_, this->SampleHandlerFD_data.data(), n_bins_), "m3b_group_upload_data");
      gcheck(m3b_group_connect(g_, M3B_EXCHANGE_PEER), "m3b_group_connect");
    } else {
      check(m3b_upload_data(h_, this->SampleHandlerFD_data.data(), n_bins_), "m3b_upload_data");
    }
    ready_ = true;
  }
  // (pre-round-2 signature)
  template <class R>
  void MoveToB200(const MonolithArrays& mono, const PointerBasesT<R>& bases, int cuda_device) { MoveToB200(mono, bases, std::vector<int>{cuda_device}); }

  // SampleHandlerFD::AddData (Samples/SampleHandlerFD.cpp:955-1044) changed SampleHandlerFD_data
  void DataChanged() {
    if (g_) gcheck(m3b_group_upload_data(g_, this->SampleHandlerFD_data.data(), n_bins_), "m3b_group_upload_data");
    else check(m3b_upload_data(h_, this->SampleHandlerFD_data.data(), n_bins_), "m3b_upload_data");
  }

  // ---- the three virtuals the fitters call -------------------------------------------------------
  void Reweight() override {
    if (!ready_) { FDBase::Reweight(); return; }           // before MoveToB200: the reference's own CPU path
    if (this->Oscillator) this->Oscillator->Evaluate();    // NuOscillator stays where it is (input array)
    for (size_t p = 0; p < spline_ptrs_.size(); ++p) spline_vals_[p] = *spline_ptrs_[p];
    const float* osc = nullptr;
    if (n_osc_dev_ > 0 && f64_) {
      // default build: the oscillator's array is double; copied to the device (the call returns once the source is free)
      check(m3b_upload_osc_f64(h_, static_cast<const double*>(osc_base_), n_osc_dev_), "m3b_upload_osc_f64");
    } else if (n_osc_dev_ > 0) {
      const float* base = static_cast<const float*>(osc_base_);
      if (osc_registered_) {
        osc = base;                                         // read by the device where it lies; the caller's next
                                                            // Evaluate() comes after GetLikelihood(), which synchronises
      } else {
        if (g_) gcheck(m3b_group_synchronize(g_), "m3b_group_synchronize");
        else check(m3b_synchronize(h_), "m3b_synchronize"); // the previous step may still be reading the staging array
        std::copy(base, base + n_osc_, osc_stage_);
        osc = osc_stage_;
      }
    }
    const double* sp = spline_vals_.empty() ? nullptr : spline_vals_.data();
    if (g_) gcheck(m3b_group_step(g_, sp, norm_base_, osc), "m3b_group_step");
    else check(m3b_step(h_, sp, norm_base_, osc), "m3b_step");
    host_arrays_stale_ = true;
    if (!this->UpdateW2) this->FirstTimeW2 = false;        // Samples/SampleHandlerFD.cpp:342
  }

  double GetLikelihood() const override {
    if (!ready_) return FDBase::GetLikelihood();
    double total = 0;
    if (g_) gcheck(m3b_group_llh(g_, &total, nullptr), "m3b_group_llh");
    else check(m3b_llh(h_, &total, nullptr), "m3b_llh");
    return total;
  }

  double GetSampleLikelihood(const int isample) const override {
    if (!ready_) return FDBase::GetSampleLikelihood(isample);
    std::vector<double> per(static_cast<size_t>(const_cast<SampleHandlerB200*>(this)->GetNsamples()));
    double total = 0;
    if (g_) gcheck(m3b_group_llh(g_, &total, per.data()), "m3b_group_llh");
    else check(m3b_llh(h_, &total, per.data()), "m3b_llh");
    return per.at(static_cast<size_t>(isample));
  }

  // Lazy host mirrors of SampleHandlerFD_array / _array_w2 (Samples/SampleHandlerFD.h:337-341).
  void SyncHostArrays() {
    if (!ready_ || !host_arrays_stale_) return;
    if (g_) gcheck(m3b_group_read_hist(g_, this->SampleHandlerFD_array.data(), this->SampleHandlerFD_array_w2.data()), "m3b_group_read_hist");
    else check(m3b_read_hist(h_, this->SampleHandlerFD_array.data(), this->SampleHandlerFD_array_w2.data()), "m3b_read_hist");
    host_arrays_stale_ = false;
  }

  // ---- batched proposals (m3b_step_batch: ONE pass over the coefficient rows for up to 256 parameter sets) ----------
  // The caller moves the parameters to proposal j exactly as it would before Reweight() (cov->SetParProp, ThrowParameters,
  // ProposeStep ...) and calls CaptureProposal(): the values behind this sample's spline and normalisation pointers are
  // recorded, nothing is evaluated.  EvaluateCaptured() then returns -lnL of every captured proposal -- the same numbers,
  // in the same order, as Reweight() + GetLikelihood() called proposal by proposal (sequential semantics: cached spline
  // segments, frozen W2).  The oscillation weights are those of the last Reweight().  adapters/BatchFitters.h builds
  // RunLLHScan, the PredictiveThrower toy loop and DelayedMR2T2's stages on these three calls.
  void BeginBatch() { batch_sp_.clear(); batch_nm_.clear(); batch_n_ = 0; }
  void CaptureProposal() {
    for (const double* p : spline_ptrs_) batch_sp_.push_back(*p);
    batch_nm_.insert(batch_nm_.end(), norm_base_, norm_base_ + n_norm_);
    ++batch_n_;
  }
  int CapturedProposals() const { return batch_n_; }
  // llh[n]; per_sample (optional) [n * GetNsamples()]; mc (optional) [n * n_bins]: every proposal's MC histogram
  std::vector<double> EvaluateCaptured(std::vector<double>* per_sample = nullptr, std::vector<double>* mc = nullptr) {
    if (!ready_) throw std::runtime_error("SampleHandlerB200::EvaluateCaptured: call MoveToB200 first");
    if (g_) throw std::runtime_error("SampleHandlerB200::EvaluateCaptured: batched proposals run on one device (MoveToB200 with a single device)");
    std::vector<double> llh(static_cast<size_t>(batch_n_), 0.0);
    if (batch_n_ == 0) return llh;
    const int nS = static_cast<int>(this->GetNsamples());
    if (per_sample) per_sample->assign(static_cast<size_t>(batch_n_) * nS, 0.0);
    const double* sp = batch_sp_.empty() ? nullptr : batch_sp_.data();
    const double* nm = batch_nm_.empty() ? nullptr : batch_nm_.data();
    if (mc) {
      mc->assign(static_cast<size_t>(batch_n_) * n_bins_, 0.0);
      check(m3b_step_batch_hist(h_, batch_n_, sp, nm, nullptr, llh.data(), per_sample ? per_sample->data() : nullptr, mc->data()), "m3b_step_batch_hist");
    } else {
      check(m3b_step_batch(h_, batch_n_, sp, nm, nullptr, llh.data(), per_sample ? per_sample->data() : nullptr), "m3b_step_batch");
    }
    host_arrays_stale_ = true;
    BeginBatch();
    return llh;
  }

  m3b_handle* handle() const { return h_; }         // the (lead) device handle
  m3b_group* group() const { return g_; }           // non-null when the sample is spread over several devices
  bool OscillatorArrayRegistered() const { return osc_registered_; }

 private:
  void check(int rc, const char* what) const {
    if (rc != M3B_OK) throw std::runtime_error(std::string("SampleHandlerB200: ") + what + ": " + m3b_last_error(h_));
  }
  void gcheck(int rc, const char* what) const {
    if (rc != M3B_OK) throw std::runtime_error(std::string("SampleHandlerB200: ") + what + ": " + m3b_group_last_error(g_));
  }
  // the two M3::float_t builds differ only in these calls
  int upload_binned(const BinnedArraysT<float>& b) {
    return m3b_upload_binned_splines(h_, b.n_params, b.max_knots, b.knot_x, b.n_pts, b.n_slots, b.uniquesplinevec_Monolith, b.coeffindexvec,
                                     b.n_unique, b.uniquecoeffindices, b.n_coeff, b.manycoeff_arr, b.xcoeff_arr);
  }
  int upload_binned(const BinnedArraysT<double>& b) {
    return m3b_upload_binned_splines_f64(h_, b.n_params, b.max_knots, b.knot_x, b.n_pts, b.n_slots, b.uniquesplinevec_Monolith, b.coeffindexvec,
                                         b.n_unique, b.uniquecoeffindices, b.n_coeff, b.manycoeff_arr, b.xcoeff_arr);
  }
  static const float* float_or_null(const std::vector<float>& w) { return w.data(); }
  static const float* float_or_null(const std::vector<double>&) { return nullptr; }
  int upload_static_f64(int64_t, const std::vector<float>&) { return M3B_OK; }
  int upload_static_f64(int64_t n, const std::vector<double>& w) { return m3b_upload_event_weights_f64(h_, n, w.data()); }
  static bool is_monolith_weight(const float* p, const MonolithArrays& m, int64_t n) { return m.cpu_total_weights && p >= m.cpu_total_weights && p < m.cpu_total_weights + n; }
  static bool is_monolith_weight(const double*, const MonolithArrays&, int64_t) { return false; }
  static int64_t monolith_index(const float* p, const MonolithArrays& m) { return p - m.cpu_total_weights; }
  static int64_t monolith_index(const double*, const MonolithArrays&) { return -1; }
  m3b_handle* h_ = nullptr;
  m3b_group* g_ = nullptr;
  bool ready_ = false, host_arrays_stale_ = false, osc_direct_ = false, osc_registered_ = false;
  int n_bins_ = 0;
  int64_t n_osc_dev_ = 0;
  const double* norm_base_ = nullptr; int n_norm_ = 0;
  const void* osc_base_ = nullptr; int64_t n_osc_ = 0;      // the oscillator's array (const M3::float_t*)
  bool f64_ = false;                                         // default build (M3::float_t = double)
  std::vector<const double*> spline_ptrs_;
  std::vector<double> spline_vals_;
  std::vector<double> batch_sp_, batch_nm_;     // captured proposals, row-major
  int batch_n_ = 0;
  float* osc_stage_ = nullptr;             // pinned staging copy (only when the oscillator's array could not be registered)
};

}  // namespace m3b200
