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
  ExperimentFD(Workload& w, bool barlow, bool update_w2) : W(w) {
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
      if (e % 97 == 13) ev.total_weight_pointers.push_back(&M3::Zero);           // NC flavour change (:1128-1131)
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
  if (timing) {
    // the adapter's real step cost, as a fitter pays it: Reweight() (incl. Oscillator->Evaluate()) + GetLikelihood()
    for (int route = 0; route < 2; ++route) {
      m3b200::SampleHandlerB200<ExperimentFD> t(w, barlow, barlow);
      m3b200::PointerBases pb = t.Bases();
      pb.register_osc_array = route == 0;
      t.MoveToB200(t.Arrays(), pb, devices);
      t.SetData(data); t.DataChanged();
      for (int k = 0; k < 10; ++k) { m3s_proposal(&c, k, w.pars.data(), w.norms.data()); t.Reweight(); t.GetLikelihood(); }
      const auto t0 = std::chrono::steady_clock::now();
      const int K = 200;
      double l = 0;
      for (int k = 0; k < K; ++k) { m3s_proposal(&c, k % 16, w.pars.data(), w.norms.data()); t.Reweight(); l = t.GetLikelihood(); }
      const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / K;
      printf("adapter step cost (%s): %.1f us per Reweight+GetLikelihood, %lld events, -lnL %.6f\n",
             t.OscillatorArrayRegistered() ? "oscillator array registered, read in place" : "oscillator array copied to a pinned staging buffer every step",
             us, (long long)E, l);
    }
  }
  printf(ok ? "ADAPTER OK\n" : "ADAPTER FAILED\n");
  return ok ? 0 : 1;
}
