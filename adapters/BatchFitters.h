// BatchFitters.h -- the reference's multi-evaluation fitter loops on the batched device step.
//
// Every loop below exists in the reference as "move the parameters, samples[i]->Reweight(), samples[i]->GetLikelihood()"
// repeated N times; one repetition costs a full pass over the spline coefficients.  m3b_step_batch evaluates up to 256
// parameter sets in ONE pass (the coefficient rows are read once and reused for every set), so N evaluations cost about
// N/8 single steps.  The templates are written against the call surface of the reference's classes and compile against
// the real MaCh3 as well as against tests/adapters/mock_mach3.h:
//
//   Sample   m3b200::SampleHandlerB200<...>: BeginBatch / CaptureProposal / EvaluateCaptured, GetNsamples
//   the parameter side is reached only through the caller's callbacks, which do exactly what the reference loop does
//   between two Reweight() calls (cov->SetParProp, ThrowParameters, ProposeStep ...), so PCA, fixed parameters, prior
//   throws etc. keep the reference's own code.
//
//   LLHScanBatched        FitterBase::RunLLHScan's inner loop          Fitters/FitterBase.cpp:742-798
//   ThrowToysBatched      PredictiveThrower's toy loop                   Fitters/PredictiveThrower.cpp:507-563
//   DelayedStagesBatched  DelayedMR2T2::DoStep's stage loop              Fitters/DelayedMR2T2.cpp:110-157
//   SwarmValuesBatched    PSO::CalcChi for a list of positions           Fitters/PSO.cpp:287-298 (uncertainty_check's 5000-point
//                         scans, :300-335; a synchronous swarm update)   scans are the exact use; see the note at the function)
//   Chi2Batch             LikelihoodFit::CalcChi2 for a list of points   Fitters/LikelihoodFit.cpp:39-137 (finite-difference
//                         gradients / Hesse grids)
//
// Results are identical to the sequential loops (same -lnL to the last bits the single-step path gives: the batched kernel
// forms every per-event weight in the reference's order; tests/test_batch_fitters_gpu.py, tests/adapters/batch_test.cpp).
#pragma once
#include "SampleHandlerB200.h"

#include <cmath>
#include <functional>
#include <limits>

namespace m3b200 {

// ---- FitterBase::RunLLHScan, one scanned parameter ---------------------------------------------------------------
// set_point(j): put the scanned parameter on point j (cov->SetParProp(i, hScan->GetBinCenter(j+1)) or the PCA form,
// Fitters/FitterBase.cpp:745-751).  Returns sample_llh[ivs][j] = samples[ivs]->GetLikelihood() at point j; with
// `split` != nullptr also split[ivs][j * nsamples + is] = GetSampleLikelihood(is) (PlotLLHScanBySample, :773-783).
// The penalty terms (systematics[ivc]->GetLikelihood()) stay with the caller: they need no device.
template <class Sample>
std::vector<std::vector<double>> LLHScanBatched(const std::vector<Sample*>& samples, int n_points,
                                                const std::function<void(int)>& set_point,
                                                std::vector<std::vector<double>>* split = nullptr) {
  for (Sample* s : samples) s->BeginBatch();
  for (int j = 0; j < n_points; ++j) {
    set_point(j);
    for (Sample* s : samples) s->CaptureProposal();
  }
  std::vector<std::vector<double>> out(samples.size());
  if (split) split->assign(samples.size(), {});
  for (size_t ivs = 0; ivs < samples.size(); ++ivs)
    out[ivs] = samples[ivs]->EvaluateCaptured(split ? &(*split)[ivs] : nullptr, nullptr);
  return out;
}

// ---- PredictiveThrower's toy loop ---------------------------------------------------------------------------------
// set_toy(i): what the reference does per toy before samples[iPDF]->Reweight() (draw a posterior step or
// ThrowParameters, SetParamters(); Fitters/PredictiveThrower.cpp:514-545).  Returns mc[ivs][i * n_bins + b] -- each
// toy's MC prediction, what WriteToy saves (:549) -- and, through llh, -lnL of every toy against the loaded data.
template <class Sample>
std::vector<std::vector<double>> ThrowToysBatched(const std::vector<Sample*>& samples, int n_toys,
                                                  const std::function<void(int)>& set_toy,
                                                  std::vector<std::vector<double>>* llh = nullptr, int chunk = 256) {
  std::vector<std::vector<double>> mc(samples.size());
  if (llh) llh->assign(samples.size(), {});
  for (int i0 = 0; i0 < n_toys; i0 += chunk) {
    const int n = std::min(chunk, n_toys - i0);
    for (Sample* s : samples) s->BeginBatch();
    for (int i = 0; i < n; ++i) {
      set_toy(i0 + i);
      for (Sample* s : samples) s->CaptureProposal();
    }
    for (size_t ivs = 0; ivs < samples.size(); ++ivs) {
      std::vector<double> m;
      const std::vector<double> l = samples[ivs]->EvaluateCaptured(nullptr, &m);
      mc[ivs].insert(mc[ivs].end(), m.begin(), m.end());
      if (llh) (*llh)[ivs].insert((*llh)[ivs].end(), l.begin(), l.end());
    }
  }
  return mc;
}

// ---- DelayedMR2T2::DoStep ------------------------------------------------------------------------------------------
// The reference proposes stage i, evaluates it, and only then decides whether stage i+1 is needed
// (Fitters/DelayedMR2T2.cpp:110-157).  Here all stages are proposed first and evaluated in one batch; the decision logic
// then replays the reference's loop on the pre-computed values.  What the caller's hooks must do:
//   propose(i, decay)   ScaleSystematics(decay ? decay_rate : unchanged) as of the END of stage i-1, then ProposeStep's
//                       proposal part + the AcceptStep "leapfrog" (:120-127); returns out_of_bounds of the proposal.
//                       `decay` is what the reference would have applied after stage i-1 (:152) IF that stage was
//                       evaluated and rejected; a stage that `continue`s (:129-131: out of bounds, or logLProp >
//                       MinLogLikelihood) does NOT decay.  Out-of-bounds is known when proposing; the second condition
//                       only after evaluation, so the batch is built on the prediction "logLProp <= MinLogLikelihood" and
//                       the result says from which stage on that prediction failed (redo_from): the caller then re-proposes
//                       those stages sequentially, exactly like the reference (rare: it needs a stage WORSE than the
//                       previous rejected one).
//   prior_llh(i)        sum of systematics[s]->GetLikelihood() for stage i's proposal (host; captured by the caller when
//                       proposing)
//   accept(i, p), delay(i)  IsStepAccepted(accProb) and "go on delaying" (= !ProbabilisticDelay()'s break, :148-150) for
//                       stage i -- the reference's own random draws.
// RNG order: all proposals are drawn before the first accept/delay draw, whereas the reference interleaves them stage by
// stage: the chain is statistically equivalent to, not draw-for-draw identical with, the reference's.
struct DelayedResult {
  int accepted_stage = -1;      // -1: the step is rejected
  bool accepted_delayed = false;
  int stages_used = 0;          // stages the reference loop would have evaluated
  int redo_from = -1;           // >= 0: the pre-proposed stages from here on assumed a decay the reference would not apply
  double logLProp = 0;          // of the last stage looked at
  std::vector<double> stage_llh;
};
template <class Sample>
DelayedResult DelayedStagesBatched(const std::vector<Sample*>& samples, int max_rejections, double logLCurr,
                                   const std::function<bool(int, bool)>& propose, const std::function<double(int)>& prior_llh,
                                   const std::function<bool(int, double)>& accept, const std::function<bool(int)>& delay,
                                   bool delay_on_oob_only = false, double large_logl = 1234567890.0) {
  const int n = max_rejections + 1;
  std::vector<char> oob(n, 0);
  for (Sample* s : samples) s->BeginBatch();
  bool decay = false;
  for (int i = 0; i < n; ++i) {
    oob[i] = propose(i, decay) ? 1 : 0;
    decay = !oob[i];                                    // predicted: an in-bounds stage is evaluated and rejected
    for (Sample* s : samples) s->CaptureProposal();
  }
  DelayedResult r;
  r.stage_llh.assign(n, 0.0);
  for (Sample* s : samples) {
    const std::vector<double> l = s->EvaluateCaptured();
    for (int i = 0; i < n; ++i) r.stage_llh[i] += l[i];
  }
  // replay of the reference loop (:116-154) on the evaluated stages
  double MinLogLikelihood = large_logl;
  bool is_delayed = false;
  for (int i = 0; i < n; ++i) {
    r.stages_used = i + 1;
    const double logLProp = oob[i] ? large_logl : r.stage_llh[i] + prior_llh(i);     // ProposeStep: out of bounds -> _LARGE_LOGL_
    r.logLProp = logLProp;
    if (oob[i] || logLProp > MinLogLikelihood) {
      if (!oob[i] && i + 1 < n) { r.redo_from = i + 1; break; }     // the next stage was proposed with a decay that did not happen
      continue;
    }
    double accProb;
    if (i == 0) {
      accProb = std::min(1.0, std::exp(logLCurr - logLProp));                      // MR2T2::AcceptanceProbability
    } else {
      const double num = std::max(0.0, std::exp(MinLogLikelihood - logLProp) - 1.0);   // DelayedMR2T2::AcceptanceProbability :79-94
      const double den = std::exp(MinLogLikelihood - logLCurr) - 1.0;
      if (den <= 0.0) accProb = 1.0;
      else if (std::isinf(num) || std::isinf(den)) accProb = std::min(1.0, std::exp(logLCurr - logLProp));
      else accProb = std::min(num / den, 1.0);
      is_delayed = true;
    }
    if (accept(i, accProb)) { r.accepted_stage = i; r.accepted_delayed = is_delayed; break; }
    if (delay_on_oob_only) break;                       // :141-145 (not out of bounds here)
    if (!delay(i)) break;                               // :148-150
    MinLogLikelihood = logLProp;                        // :153
  }
  return r;
}

// ---- PSO::CalcChi / LikelihoodFit::CalcChi2 for a list of positions ----------------------------------------------------
// set_position(k): the body of CalcChi up to the Reweight (SetParameters of every covariance object, Fitters/PSO.cpp:
// 287-292; LikelihoodFit.cpp:47-66).  Returns the SAMPLE part of the value per position; the caller adds its penalty
// terms.  Exact for PSO::uncertainty_check's scans (:300-335) and for finite-difference grids.  For PSO::swarmIterate
// (:338-373) batching the particles of one iteration makes the swarm SYNCHRONOUS (every particle sees the best position of
// the previous iteration; the reference updates the global best particle by particle): a standard PSO variant, not the
// reference's algorithm -- use it knowingly.
template <class Sample>
std::vector<double> SwarmValuesBatched(const std::vector<Sample*>& samples, int n_positions, const std::function<void(int)>& set_position) {
  std::vector<double> v(static_cast<size_t>(n_positions), 0.0);
  for (int k0 = 0; k0 < n_positions; k0 += 256) {
    const int n = std::min(256, n_positions - k0);
    for (Sample* s : samples) s->BeginBatch();
    for (int k = 0; k < n; ++k) { set_position(k0 + k); for (Sample* s : samples) s->CaptureProposal(); }
    for (Sample* s : samples) {
      const std::vector<double> l = s->EvaluateCaptured();
      for (int k = 0; k < n; ++k) v[k0 + k] += l[k];
    }
  }
  return v;
}
template <class Sample>
std::vector<double> Chi2Batch(const std::vector<Sample*>& samples, int n_points, const std::function<void(int)>& set_point) {
  std::vector<double> v = SwarmValuesBatched(samples, n_points, set_point);
  for (double& x : v) x *= 2.0;                            // LikelihoodFit::CalcChi2 returns 2 * (-lnL), :134
  return v;
}

}  // namespace m3b200
