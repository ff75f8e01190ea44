// adapter_test.cpp -- TEST: adapters/SampleHandlerB200.h compiled against tests/adapters/mock_mach3.h.
// An "experiment" class wires a synthetic workload the way SampleHandlerFD::Initialise does (per-event
// pointer vectors into the parameter array, the oscillator's weight array and the monolith's
// cpu_total_weights), then two instances run the same proposals: one through the mock's scalar CPU path,
// one through SampleHandlerB200 -> libm3b200 on the B200.  -lnL must agree to 1e-6 relative
// (north_star), histograms to 1e-9.    usage: adapter_test [n_events] [barlow|poisson] [n_members] [time]
// n_members > 1: the adapter spreads the sample over that many group members (devices 0..n-1 modulo the GPUs present;
// m3b_group_*: still one process, one calling thread).  "time": prints the adapter's real per-step cost
// (Oscillator->Evaluate() + Reweight() + GetLikelihood(), host wall clock) for the registered-array and the staging route.
#include "mock_mach3.h"
#include "SampleHandlerB200.h"
#include "BatchFitters.h"
#include "m3b_synth.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime_api.h>
#include <random>

struct Workload {
  m3s_config c{};
  std::vector<int8_t> type; std::vector<int16_t> n_pts; std::vector<float> coeff_x;
  std::vector<double> pars, norms;            // the doubles every pointer points into (ParameterHandler::_fPropVal)
};

class ExperimentFD : public SampleHandlerFD {
 public:
  // plain_osc: every event points at its own entry of the oscillator's array, in event order (no &M3::Zero / missing
  // pointers): the shape in which the adapter hands the array to the kernel as it stands (no index table, no copy)
  ExperimentFD(Workload& w, bool barlow, bool update_w2, bool plain_osc = false) : W(w) {
    const m3s_config& c = w.c;
    const int64_t E = c.n_events;
    nEvents = static_cast<unsigned>(E); nSamples = static_cast<M3::int_t>(c.n_samples);
    fTestStatistic = barlow ? kBarlowBeeston : kPoisson; UpdateW2 = update_w2;
    Binning = std::make_unique<BinningHandler>();
    for (int s = 0; s < c.n_samples; ++s) {
      Binning->edges.emplace_back();
      for (int d = 0; d < c.n_dims; ++d) {
        const int nb = d == 0 ? c.nbins_x : c.nbins_y;
        std::vector<double> e(nb + 1);
        m3s_bin_edges(&c, s, d, e.data());
        Binning->edges.back().push_back(e);
      }
    }
    Binning->Finalise();
    SampleHandlerFD_array.assign(Binning->GetNBins(), 0.); SampleHandlerFD_array_w2 = SampleHandlerFD_array; SampleHandlerFD_data = SampleHandlerFD_array;
    // monolith (reference arrays)
    auto mono = std::make_unique<SMonolith>();
    Mono = mono.get();
    mono->nParams = c.n_params; mono->max_knots = c.n_knots; mono->coeff_x = w.coeff_x; mono->n_pts = w.n_pts;
    std::vector<uint32_t> nc(E), nl(E); uint64_t tc = 0, tl = 0;
    m3s_count(&c, 0, E, nc.data(), nl.data(), &tc, &tl);
    mono->nParamPerEvent.resize(2 * E); mono->nParamPerEvent_tf1.resize(2 * E);
    mono->paramNo_arr.resize(tc); mono->paramNo_tf1.resize(tl); mono->coeff_many.resize(tc * c.n_knots * 4); mono->coeff_tf1.resize(tl * 2);
    std::vector<uint64_t> koff(tc);
    m3s_fill_splines(&c, 0, E, mono->nParamPerEvent.data(), mono->paramNo_arr.data(), koff.data(), mono->coeff_many.data(),
                     mono->nParamPerEvent_tf1.data(), mono->paramNo_tf1.data(), mono->coeff_tf1.data());
    mono->nKnots_arr.assign(koff.begin(), koff.end());
    for (int p = 0; p < c.n_params; ++p) mono->pars.push_back(&w.pars[p]);
    mono->Finalise(E);
    SplineHandler = std::move(mono);
    // oscillator
    Oscillator = std::make_shared<OscillationHandler>();
    for (int step = 0; step < 4; ++step) { std::vector<float> o(E); m3s_fill_osc(&c, 0, E, step, o.data()); Oscillator->per_step.push_back(o); }
    Oscillator->weights = Oscillator->per_step[0];
    // events: the pointer soup of SampleHandlerFD::Initialise
    sample_id.resize(E); kin.resize(size_t(c.n_dims) * E); norm_idx.resize(size_t(c.n_norm_per_event) * E); static_w.resize(E);
    m3s_fill_events(&c, 0, E, sample_id.data(), kin.data(), norm_idx.data(), static_w.data());
    MCSamples.resize(E);
    for (int64_t e = 0; e < E; ++e) {
      EventInfo& ev = MCSamples[e];
      ev.NominalSample = sample_id[e];
      for (int d = 0; d < c.n_dims; ++d) ev.KinVar.push_back(&kin[size_t(d) * E + e]);
      for (int j = 0; j < c.n_norm_per_event; ++j) if (norm_idx[e * c.n_norm_per_event + j] >= 0) ev.norm_pointers.push_back(&w.norms[norm_idx[e * c.n_norm_per_event + j]]);
      if (plain_osc) ev.total_weight_pointers.push_back(&Oscillator->weights[e]);
      else if (e % 97 == 13) ev.total_weight_pointers.push_back(&M3::Zero);           // NC flavour change (:1128-1131)
      else if (e % 89 != 7) ev.total_weight_pointers.push_back(&Oscillator->weights[e]);   // else: &M3::Unity is never pushed (:1116)
      ev.total_weight_pointers.push_back(Mono->retPointer(int(e)));               // :1248
      ev.total_weight_pointers.push_back(&static_w[e]);                           // AddAdditionalWeightPointers
    }
    // selection cuts (StoredSelection, SampleHandlerFD.cpp:140-165): sample 0 on its first binning variable, sample 1 on
    // a variable that is only cut on (ReturnKinematicParameter(7, e)), sample 2 none
    StoredSelection.resize(size_t(c.n_samples));
    cut_only.resize(E);
    for (int64_t e = 0; e < E; ++e) cut_only[e] = double((e * 2654435761u) % 1000) / 1000.0;
    { KinematicCut k; k.ParamToCutOnIt = 0; k.LowerBound = 0.2; k.UpperBound = 2.5; StoredSelection[0].push_back(k); }
    if (c.n_samples > 1) {
      KinematicCut k; k.ParamToCutOnIt = 7; k.LowerBound = 0.125; k.UpperBound = 0.875; StoredSelection[1].push_back(k);
      KinematicCut k2; k2.ParamToCutOnIt = 1; k2.LowerBound = 0.1; k2.UpperBound = 3.0; StoredSelection[1].push_back(k2);
    }
  }
  double ReturnKinematicParameter(int var, int e) override {
    return var == 7 ? cut_only[size_t(e)] : kin[size_t(var) * nEvents + size_t(e)];
  }
  void SetData(const std::vector<double>& d) { SampleHandlerFD_data = d; }
  void FreezeOscillator() { Oscillator->per_step.clear(); }
  const std::vector<double>& MC() const { return SampleHandlerFD_array; }
  const std::vector<double>& W2() const { return SampleHandlerFD_array_w2; }
  m3b200::MonolithArrays Arrays() const {
    m3b200::MonolithArrays a;
    a.n_params = Mono->nParams; a.max_knots = Mono->max_knots; a.coeff_x = Mono->coeff_x.data(); a.n_pts = Mono->n_pts.data();
    a.nParamPerEvent = Mono->nParamPerEvent.data(); a.paramNo_arr = Mono->paramNo_arr.data(); a.nKnots_arr = Mono->nKnots_arr.data();
    a.total_knots = uint32_t(Mono->coeff_many.size() / 4); a.coeff_many = Mono->coeff_many.data();
    a.nParamPerEvent_tf1 = Mono->nParamPerEvent_tf1.data(); a.paramNo_tf1 = Mono->paramNo_tf1.data(); a.coeff_tf1 = Mono->coeff_tf1.data();
    a.spline_par_pointers = Mono->pars; a.cpu_total_weights = Mono->cpu_total_weights;
    return a;
  }
  m3b200::PointerBases Bases() const {
    m3b200::PointerBases b;
    b.norm_base = W.norms.data(); b.n_norm = int(W.norms.size());
    b.osc_base = Oscillator->weights.data(); b.n_osc = int64_t(Oscillator->weights.size());
    b.zero = &M3::Zero; b.unity = &M3::Unity;
    b.constant_weight_ranges.push_back({static_w.data(), static_w.data() + static_w.size()});
    return b;
  }
 protected:
  Workload& W;
  SMonolith* Mono = nullptr;
  std::vector<int32_t> sample_id; std::vector<double> kin, cut_only; std::vector<int16_t> norm_idx; std::vector<float> static_w;
  template <class T> friend class m3b200::SampleHandlerB200;
};

// adapters/BatchFitters.h against the reference's sequential loops run on the mock's CPU path: FitterBase::RunLLHScan's
// inner loop (Fitters/FitterBase.cpp:742-798), PredictiveThrower's toy loop (Fitters/PredictiveThrower.cpp:507-563) and
// DelayedMR2T2::DoStep (Fitters/DelayedMR2T2.cpp:110-157) with the same pre-drawn proposals and accept/delay draws.
static bool batch_consumers(Workload& w, ExperimentFD& cpu, m3b200::SampleHandlerB200<ExperimentFD>& gpu) {
  using B200 = m3b200::SampleHandlerB200<ExperimentFD>;
  const m3s_config& c = w.c;
  bool ok = true;
  // a batch is evaluated with the oscillation weights of the last Reweight(): hold the oscillator still on both sides
  cpu.FreezeOscillator(); gpu.FreezeOscillator();
  std::vector<B200*> samples = {&gpu};
  // ---- RunLLHScan: parameter 3 over 100 points, central values from proposal 2
  m3s_proposal(&c, 2, w.pars.data(), w.norms.data());
  const std::vector<double> pars0 = w.pars;
  const int n_points = 100, ipar = 3;
  auto point = [&](int j) { w.pars = pars0; w.pars[ipar] = -2.9 + 5.8 * (j + 0.5) / n_points; };
  std::vector<std::vector<double>> split;
  const auto scan = m3b200::LLHScanBatched<B200>(samples, n_points, point, &split);
  double worst = 0, worst_split = 0;
  for (int j = 0; j < n_points; ++j) {
    point(j);
    cpu.Reweight();
    const double l = cpu.GetLikelihood();
    worst = std::max(worst, std::fabs(scan[0][j] - l) / std::max(1e-300, std::fabs(l)));
    double ls = 0;
    for (int s = 0; s < c.n_samples; ++s) ls += split[0][size_t(j) * c.n_samples + s];
    worst_split = std::max(worst_split, std::fabs(ls - scan[0][j]) / std::max(1e-300, std::fabs(l)));
  }
  printf("LLHScanBatched: %d points, worst rel diff to the sequential loop %.2e, per-sample sums %.1e  %s\n", n_points, worst, worst_split,
         worst <= 1e-6 && worst_split <= 1e-9 ? "OK" : "FAIL");
  ok &= worst <= 1e-6 && worst_split <= 1e-9;
  // ---- PredictiveThrower: 40 toys, each toy's MC histogram and -lnL
  const int n_toys = 40;
  auto toy = [&](int i) { m3s_proposal(&c, 100 + i, w.pars.data(), w.norms.data()); };
  std::vector<std::vector<double>> toy_llh;
  const auto mc = m3b200::ThrowToysBatched<B200>(samples, n_toys, toy, &toy_llh, 16);
  const size_t nb = cpu.MC().size();
  double worst_mc = 0, worst_l = 0;
  for (int i = 0; i < n_toys; ++i) {
    toy(i);
    cpu.Reweight();
    for (size_t b = 0; b < nb; ++b) worst_mc = std::max(worst_mc, std::fabs(mc[0][size_t(i) * nb + b] - cpu.MC()[b]) / std::max(1.0, std::fabs(cpu.MC()[b])));
    const double l = cpu.GetLikelihood();
    worst_l = std::max(worst_l, std::fabs(toy_llh[0][i] - l) / std::max(1e-300, std::fabs(l)));
  }
  printf("ThrowToysBatched: %d toys, worst histogram diff %.1e, worst -lnL rel diff %.2e  %s\n", n_toys, worst_mc, worst_l,
         worst_mc <= 1e-9 && worst_l <= 1e-6 ? "OK" : "FAIL");
  ok &= worst_mc <= 1e-9 && worst_l <= 1e-6;
  // ---- DelayedMR2T2::DoStep: 60 steps, 3 allowed rejections, against the reference's loop run stage by stage on the CPU
  //      instance with the SAME pre-drawn unit proposals, accept draws and delay draws
  {
    std::mt19937_64 rng(99);
    std::normal_distribution<double> gaus(0, 1);
    std::uniform_real_distribution<double> uni(0, 1);
    const int max_rej = 3; const double decay_rate = 0.3, large = 1234567890.0, delay_probability = 0.8;
    int agree = 0, accepted = 0, delayed = 0, redo = 0, n_steps = 60;
    std::vector<double> curr = pars0;
    w.pars = curr; cpu.Reweight();
    double logLCurr = cpu.GetLikelihood();
    double worst_stage = 0;
    for (int step = 0; step < n_steps; ++step) {
      std::vector<std::vector<double>> unit(max_rej + 1, std::vector<double>(curr.size()));
      std::vector<double> u_acc(max_rej + 1), u_del(max_rej + 1);
      for (int i = 0; i <= max_rej; ++i) { for (double& x : unit[i]) x = gaus(rng); u_acc[i] = uni(rng); u_del[i] = uni(rng); }
      // the proposal rule both runs share: centred on the previous stage's proposal (the AcceptStep "leapfrog",
      // :124-127), scale multiplied by decay_rate when the previous stage was evaluated and rejected (:152)
      std::vector<double> centre; double scale = 0; std::vector<std::vector<double>> props(max_rej + 1);
      auto propose = [&](int i, bool decay) {
        if (i == 0) { centre = curr; scale = 0.12; }
        if (decay) scale *= decay_rate;
        std::vector<double> p(curr.size());
        bool oob = false;
        for (size_t k = 0; k < p.size(); ++k) { p[k] = centre[k] + scale * unit[i][k]; oob |= std::fabs(p[k]) > 2.9; }
        centre = p; props[i] = p; w.pars = p;
        return oob;
      };
      // (1) the reference's loop, stage by stage, on the CPU instance
      int acc_ref = -1; double MinLL = large, ll_ref = 0; bool decay = false;
      std::vector<double> llh_ref(max_rej + 1, 0.0);
      for (int i = 0; i <= max_rej; ++i) {
        const bool oob = propose(i, decay);
        decay = false;
        double logLProp = large;
        if (!oob) { cpu.Reweight(); logLProp = cpu.GetLikelihood(); llh_ref[i] = logLProp; }
        if (oob || logLProp > MinLL) continue;
        double accProb;
        if (i == 0) accProb = std::min(1.0, std::exp(logLCurr - logLProp));
        else {
          const double num = std::max(0.0, std::exp(MinLL - logLProp) - 1.0), den = std::exp(MinLL - logLCurr) - 1.0;
          accProb = den <= 0.0 ? 1.0 : ((std::isinf(num) || std::isinf(den)) ? std::min(1.0, std::exp(logLCurr - logLProp)) : std::min(num / den, 1.0));
        }
        if (!(u_acc[i] > accProb)) { acc_ref = i; ll_ref = logLProp; break; }       // MCMCBase::IsStepAccepted
        if (u_del[i] > delay_probability) break;                                      // ProbabilisticDelay() == false: stop delaying
        decay = true; MinLL = logLProp;
      }
      const std::vector<std::vector<double>> props_ref = props;
      // (2) the batched form: all stages proposed first, one device pass, then the same decisions
      const auto r = m3b200::DelayedStagesBatched<B200>(samples, max_rej, logLCurr, propose, [](int) { return 0.0; },
                                                        [&](int i, double pacc) { return !(u_acc[i] > pacc); },
                                                        [&](int i) { return !(u_del[i] > delay_probability); });
      if (r.redo_from >= 0) { ++redo; }            // (the sequential re-proposal of the tail is the caller's: count, do not compare)
      else {
        agree += (r.accepted_stage == acc_ref);
        for (int i = 0; i < r.stages_used; ++i)
          if (llh_ref[i] != 0.0 && props[i] == props_ref[i])
            worst_stage = std::max(worst_stage, std::fabs(r.stage_llh[i] - llh_ref[i]) / std::max(1e-300, std::fabs(llh_ref[i])));
      }
      if (acc_ref >= 0) { curr = props_ref[acc_ref]; logLCurr = ll_ref; ++accepted; delayed += acc_ref > 0; }
    }
    const bool good = agree + redo == n_steps && worst_stage <= 1e-6 && accepted > 0;
    printf("DelayedStagesBatched: %d steps, %d accepted (%d after a delay), decisions agree with the stage-by-stage loop in %d, "
           "%d needed a sequential tail, worst stage -lnL rel diff %.2e  %s\n", n_steps, accepted, delayed, agree, redo, worst_stage, good ? "OK" : "FAIL");
    ok &= good;
  }
  return ok;
}

int main(int argc, char** argv) {
  const int64_t E = argc > 1 ? atoll(argv[1]) : 30011;
  const bool barlow = argc > 2 && !strcmp(argv[2], "barlow");
  Workload w;
  m3s_config& c = w.c;
  c.seed = 4242; c.n_events = E; c.n_params = 24; c.n_linear = 4; c.n_knots = 6; c.n_modes = 6; c.density = 0.6f;
  c.n_samples = 3; c.n_dims = 2; c.nbins_x = 20; c.nbins_y = 6; c.n_norm_params = 5; c.n_norm_per_event = 2; c.mode_block = 500;
  for (int s = 0; s <= c.n_samples; ++s) c.sample_start[s] = E * s / c.n_samples;
  w.type.resize(c.n_params); w.n_pts.resize(c.n_params); w.coeff_x.resize(size_t(c.n_params) * c.n_knots);
  m3s_param_layout(&c, w.type.data(), w.n_pts.data(), w.coeff_x.data());
  w.pars.assign(c.n_params, 0.); w.norms.assign(c.n_norm_params, 1.);

  const int n_members = argc > 3 ? atoi(argv[3]) : 1;
  const bool timing = argc > 4 && !strcmp(argv[4], "time");
  int n_gpus = 1;
  cudaGetDeviceCount(&n_gpus);
  std::vector<int> devices;
  for (int i = 0; i < n_members; ++i) devices.push_back(i % (n_gpus > 0 ? n_gpus : 1));
  ExperimentFD cpu(w, barlow, barlow);
  m3b200::SampleHandlerB200<ExperimentFD> gpu(w, barlow, barlow);
  gpu.MoveToB200(gpu.Arrays(), gpu.Bases(), devices);
  printf("adapter on %d member(s), %d GPU(s) present, oscillator array %s\n", n_members, n_gpus,
         gpu.OscillatorArrayRegistered() ? "registered in place" : "staged");

  SampleHandlerBase* handlers[2] = {&cpu, &gpu};       // the fitters only see the base-class virtuals
  m3s_proposal(&c, -1, w.pars.data(), w.norms.data());
  for (auto* h : handlers) h->Reweight();
  std::vector<double> data(cpu.MC().size());
  std::mt19937_64 rng(7);
  for (size_t b = 0; b < data.size(); ++b) data[b] = double(std::poisson_distribution<long>(cpu.MC()[b] + 1e-9)(rng));
  cpu.SetData(data); gpu.SetData(data); gpu.DataChanged();

  bool ok = true;
  for (int step : {-1, 0, 1, -2, 2, 3, -3, 4}) {
    m3s_proposal(&c, step, w.pars.data(), w.norms.data());
    double l[2];
    for (int i = 0; i < 2; ++i) { handlers[i]->Reweight(); l[i] = handlers[i]->GetLikelihood(); }
    gpu.SyncHostArrays();
    double dmax = 0, dw2 = 0;
    for (size_t b = 0; b < data.size(); ++b) {
      dmax = std::max(dmax, std::fabs(cpu.MC()[b] - gpu.MC()[b]) / std::max(1.0, std::fabs(cpu.MC()[b])));
      dw2 = std::max(dw2, std::fabs(cpu.W2()[b] - gpu.W2()[b]) / std::max(1.0, std::fabs(cpu.W2()[b])));
    }
    double ls = 0;
    for (int s = 0; s < c.n_samples; ++s) ls += gpu.GetSampleLikelihood(s);
    const double rel = std::fabs(l[0] - l[1]) / std::max(1e-300, std::fabs(l[0]));
    const bool good = rel <= 1e-6 && dmax <= 1e-9 && dw2 <= 1e-9 && std::fabs(ls - l[1]) <= 1e-9 * std::fabs(l[1]) + 1e-12;
    ok &= good;
    printf("step %2d: -lnL cpu %.9f  b200 %.9f  rel %.2e  hist %.1e  w2 %.1e  sum(per-sample) %.9f  %s\n", step, l[0], l[1], rel, dmax, dw2, ls,
           good ? "OK" : "FAIL");
  }
  if (argc > 4 && !strcmp(argv[4], "batch")) ok &= batch_consumers(w, cpu, gpu);
  if (timing) {
    // the adapter's real step cost, as a fitter pays it: Reweight() (incl. Oscillator->Evaluate()) + GetLikelihood()
    // (the mock oscillator is held still: its Evaluate() would copy the whole weight array on the host every step, which is
    //  the oscillator's cost, not the adapter's; the array the device reads is the same either way)
    for (int route = 0; route < 4; ++route) {
      const bool plain = route >= 2;
      m3b200::SampleHandlerB200<ExperimentFD> t(w, barlow, barlow, plain);
      m3b200::PointerBases pb = t.Bases();
      pb.register_osc_array = route % 2 == 0;
      t.MoveToB200(t.Arrays(), pb, devices);
      t.FreezeOscillator();
      t.SetData(data); t.DataChanged();
      for (int k = 0; k < 10; ++k) { m3s_proposal(&c, k, w.pars.data(), w.norms.data()); t.Reweight(); t.GetLikelihood(); }
      const auto t0 = std::chrono::steady_clock::now();
      const int K = 200;
      double l = 0;
      for (int k = 0; k < K; ++k) { m3s_proposal(&c, k % 16, w.pars.data(), w.norms.data()); t.Reweight(); l = t.GetLikelihood(); }
      const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / K;
      printf("adapter step cost (%s; %s): %.1f us per Reweight+GetLikelihood, %lld events, -lnL %.6f\n",
             plain ? "one oscillation weight per event in event order: streamed by the kernel" : "indexed oscillation weights: one H2D copy of the array per step",
             t.OscillatorArrayRegistered() ? "oscillator array registered, read in place" : "oscillator array copied to a pinned staging buffer every step",
             us, (long long)E, l);
    }
  }
  printf(ok ? "ADAPTER OK\n" : "ADAPTER FAILED\n");
  return ok ? 0 : 1;
}
