"""Batch consumers of the hot path: fitter-side loops of the reference that evaluate many
parameter sets against the same events, rewritten on top of m3b_step_batch so that one pass over the coefficient
rows serves the whole loop.

    RunLLHScan     FitterBase::RunLLHScan (Fitters/FitterBase.cpp:620-798): for every scanned parameter, n_points
                   values at the bin centres of [lower, upper], all other parameters at their central values,
                   samples[i]->Reweight(); samples[i]->GetLikelihood() (+ GetSampleLikelihood when split by sample).
                   Returns what the reference writes into hScanSam / hScanSamSplit: 2 x (-lnL).

    ProduceToys    the sample side of PredictiveThrower::ProduceToys (Fitters/PredictiveThrower.cpp:507-563): for every
                   thrown parameter set samples[i]->Reweight() and the MC histogram of every sample (what WriteToy
                   stores); here all throws go through one m3b_step_batch_hist call.

    EvaluateDelayedStages
                   DelayedMR2T2::DoStep (Fitters/DelayedMR2T2.cpp:110-157) proposes up to max_rejections+1 stage
                   points per step, each needing samples[i]->Reweight() + GetLikelihood() before the accept test.
                   The stage proposals do not depend on the earlier stages' likelihoods (only on the random draws
                   and the decaying step scale), so a fitter can draw them up front and evaluate them speculatively
                   in one batch; it then walks the stages in order and stops at the first accepted one.

Only the sample-likelihood part is computed here; the systematic (prior) terms of the scan come from the
ParameterHandler and are out of scope.
"""
from __future__ import annotations

import numpy as np


def scan_points(lower: float, upper: float, n_points: int) -> np.ndarray:
    """TH1D(n_points, lower, upper)->GetBinCenter(j+1), j = 0..n_points-1 (Fitters/FitterBase.cpp:738)."""
    width = (upper - lower) / n_points
    return lower + (np.arange(n_points) + 0.5) * width


def RunLLHScan(sample, central_spline_pars, central_norm_pars=None, spline_ranges=None, norm_ranges=None, n_points=100,
               by_sample=False):
    """`sample`: mach3_b200.handlers.SampleHandlerFD.  spline_ranges / norm_ranges: {parameter index: (lower, upper)}.
    Returns {("spline"|"norm", index): {"x": points, "llh2": 2*(-lnL)[n_points], "llh2_by_sample": [n_points, n_samples]}}.
    The handle is left at the central values (one extra Reweight), like the reference resets the parameter."""
    h = sample.handle
    c_sp = np.ascontiguousarray(central_spline_pars, np.float64)
    c_nm = None if central_norm_pars is None else np.ascontiguousarray(central_norm_pars, np.float64)
    out = {}
    for kind, ranges in (("spline", spline_ranges or {}), ("norm", norm_ranges or {})):
        for idx, (lo, hi) in ranges.items():
            x = scan_points(lo, hi, n_points)
            sps = np.tile(c_sp, (n_points, 1))
            nms = None if c_nm is None else np.tile(c_nm, (n_points, 1))
            if kind == "spline":
                sps[:, idx] = x
            else:
                nms[:, idx] = x
            tot, per = h.step_batch(sps, nms, per_sample=True)
            out[(kind, idx)] = {"x": x, "llh2": 2.0 * tot, "llh2_by_sample": 2.0 * per if by_sample else None}
    h.step(c_sp, c_nm)
    h.llh()
    return out


def ProduceToys(sample, spline_par_throws, norm_par_throws=None, chunk=256):
    """`sample`: mach3_b200.handlers.SampleHandlerFD.  spline_par_throws[n_toys, n_params] / norm_par_throws[n_toys, n_norm]:
    the parameter sets PredictiveThrower draws from the posterior chain (or from the prior).  Returns
    (mc[n_toys, n_bins], llh[n_toys]): each toy's MC prediction in global-bin order (slice it per sample with the
    binning's sample offsets) and the sample -lnL of that throw against the loaded data.  Sequential semantics: the
    result equals looping Reweight() over the throws."""
    h = sample.handle
    sp = np.ascontiguousarray(spline_par_throws, np.float64)
    nm = None if norm_par_throws is None else np.ascontiguousarray(norm_par_throws, np.float64)
    n = sp.shape[0]
    mcs, llhs = [], []
    for i0 in range(0, n, chunk):
        tot, mc = h.step_batch_hist(sp[i0:i0 + chunk], None if nm is None else nm[i0:i0 + chunk])
        mcs.append(mc); llhs.append(tot)
    return np.concatenate(mcs, 0), np.concatenate(llhs)


def EvaluateDelayedStages(sample, stage_spline_pars, stage_norm_pars=None):
    """`sample`: mach3_b200.handlers.SampleHandlerFD.  stage_spline_pars[n_stages, n_params] /
    stage_norm_pars[n_stages, n_norm]: the stage proposals of one delayed-rejection step, in stage order.  Returns
    -lnL[n_stages] (sample part of logLProp for every stage), evaluated with the reference's sequential semantics in
    stage order.

    This is a speculative evaluator, NOT a re-implementation of DelayedMR2T2::DoStep (Fitters/DelayedMR2T2.cpp:110-157);
    the caller that pre-draws the stages must reproduce the reference's rules itself:
      * stage i+1 is proposed around stage i's proposal (the AcceptStep "leapfrog", :124-127), with the step scale
        multiplied by DecayRate only when stage i was actually evaluated and rejected (:152-153) -- a stage that is out
        of bounds or has logLProp > MinLogLikelihood `continue`s WITHOUT the decay (:129-131);
      * the reference interleaves its random draws (ProposeStep, IsStepAccepted, ProbabilisticDelay) stage by stage;
        drawing all proposals first changes the order in which the RNG stream is consumed, so a chain built on this
        helper is statistically equivalent to, but not step-for-step identical with, the reference's chain;
      * stages after the accepted one are evaluated too, so the cached spline segment (Splines/SplineBase.cpp:76) may
        sit elsewhere afterwards -- it only matters for a later proposal exactly on a knot.
    The C++ form with these rules built in is adapters/BatchFitters.h (DelayedStages)."""
    sp = np.ascontiguousarray(stage_spline_pars, np.float64)
    nm = None if stage_norm_pars is None else np.ascontiguousarray(stage_norm_pars, np.float64)
    return sample.handle.step_batch(sp, nm)
