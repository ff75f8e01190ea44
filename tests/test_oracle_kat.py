"""Known-answer tests of the CPU oracle.  The reference ships no golden
vectors for this path, so these are derived from the formulas in the cited reference lines."""
import math

import numpy as np
import pytest

from mach3_b200 import synth
from oracle import binding as O


def _tiny_monolith(x, coeffs_per_event, lin=None):
    """One cubic parameter (p0) with knots x and, optionally, one TF1 parameter (p1)."""
    K = len(x)
    E = len(coeffs_per_event)
    P = 2 if lin is not None else 1
    cx = np.full(P * K, -999, np.float32)
    cx[:K] = x
    npts = np.array([K] + ([0] if lin is not None else []), np.int16)
    spl = dict(n_events=E,
               nParamPerEvent=np.array([[1, e] for e in range(E)], np.uint32).ravel(),
               paramNo_arr=np.zeros(E, np.int16), nKnots_arr=(np.arange(E) * K).astype(np.uint64),
               coeff_many=np.asarray(coeffs_per_event, np.float32).ravel(),
               nParamPerEvent_tf1=np.array([[1 if lin is not None else 0, e if lin is not None else 0] for e in range(E)], np.uint32).ravel(),
               paramNo_tf1=np.ones(E if lin is not None else 0, np.int16),
               coeff_tf1=np.asarray(lin if lin is not None else [], np.float32).ravel())
    return O.SMonolith(P, K, cx, npts, spl)


DOC = [[1, -1, 0, 0.5], [0.5, 0.5, 1.5, -0.5], [2, 0, 0, 0]]   # x=[0,1,2], y=[1,.5,2] natural spline (§8c iv)


def test_doc_example_natural_spline_generator_convention():
    # the generator's own coefficient builder reproduces the hand-computed rows
    m = _tiny_monolith([0, 1, 2], [DOC])
    for xv, seg, want in [(0.5, 0, 1 - 0.5 + 0.5 * 0.125), (1.5, 1, 0.5 + 0.25 + 1.5 * 0.25 - 0.5 * 0.125)]:
        m.set_params([xv]); m.Evaluate()
        assert m.segments[0] == seg
        assert m.total_weights[0] == pytest.approx(want, rel=1e-6)
    # segment ends: seg0 at dx=1 -> 0.5 ; seg1 at dx=1 -> 2
    m.set_params([2.0]); m.Evaluate()
    assert m.segments[0] == 1 and m.total_weights[0] == pytest.approx(2.0, rel=1e-6)


def test_segment_rules_edges_and_clamp():
    # (ii) xvar <= x0 -> 0 ; xvar >= x_last -> K-2   (Splines/SplineBase.cpp:69-74,97)
    x = [-3, -1.5, 0, 1.5, 3]
    m = _tiny_monolith(x, [[[1, 0, 0, 0]] * 5])
    for xv, seg in [(-5, 0), (-3, 0), (3, 3), (9, 3), (-2.9, 0), (2.9, 3), (0.1, 2), (-0.1, 1)]:
        m.set_params([xv]); m.FindSplineSegment()
        assert m.segments[0] == seg, (xv, seg)
        assert m.param_values[0] == np.float32(xv)


def test_segment_history_dependence_at_exact_knot():
    # (iii) cached-segment test x[prev] <= xvar < x[prev+1] (:76) vs binary search x[seg] < xvar <= x[seg+1] (:84-94)
    x = [-3, -1.5, 0, 1.5, 3]
    m = _tiny_monolith(x, [[[1, 0, 0, 0]] * 5])
    m.set_curr_segment(0, 0)
    m.set_params([0.0]); m.FindSplineSegment()
    assert m.segments[0] == 1            # binary search puts an exact knot hit in the segment BELOW
    m.set_curr_segment(0, 2)
    m.FindSplineSegment()
    assert m.segments[0] == 2            # cached segment [x2, x3) accepts it
    m.FindSplineSegment()
    assert m.segments[0] == 2            # and keeps it


def test_exact_knot_hit_value():
    # (i) with segment k and dx == 0 the fmaf chain returns Y_k bit-exactly
    x = [0, 1, 2]
    m = _tiny_monolith(x, [DOC])
    m.set_curr_segment(0, 1)
    m.set_params([1.0]); m.Evaluate()
    assert m.segments[0] == 1 and m.total_weights[0] == np.float32(0.5)
    m.set_curr_segment(0, 0)             # binary search side: segment 0 at dx=1 -> 1-1+0+0.5
    m.set_params([1.0]); m.FindSplineSegment()
    assert m.segments[0] == 0
    m.CalcSplineWeights(); m.CalcTotalEventWeight()
    assert m.total_weights[0] == pytest.approx(0.5, rel=1e-6)


def test_tf1_and_product_order():
    # a*x+b (Splines/SplineMonolith.cpp:780); total = cubic product then TF1 product (:799-828)
    m = _tiny_monolith([0, 1, 2], [DOC, DOC], lin=[[0.1, 1.0], [-0.05, 1.0]])
    m.set_params([0.5, 2.0]); m.Evaluate()
    cub = np.float32(1 - 0.5 + 0.5 * 0.125)
    assert m.tf1_weights[0] == np.float32(np.float32(0.1) * np.float32(2.0) + np.float32(1.0))
    assert m.total_weights[0] == pytest.approx(float(cub) * 1.2, rel=1e-6)
    assert m.total_weights[1] == pytest.approx(float(cub) * 0.9, rel=1e-6)


def test_poisson_llh_kats():
    # (v) Samples/SampleHandlerBase.cpp:17-31
    P = O.kPoisson
    assert O.test_stat_llh(P, 0, 3.25, 0) == 3.25
    assert O.test_stat_llh(P, 0, 0, 0) == 0
    assert O.test_stat_llh(P, 7.5, 7.5, 0) == 0
    d = 4.0
    assert O.test_stat_llh(P, d, 1e-7, 0) == pytest.approx(1e-5 - d + d * math.log(d / 1e-5), rel=1e-14)
    assert O.test_stat_llh(P, 5e-6, 1e-6, 0) == 0.0       # data <= bound, data >= mc
    assert O.test_stat_llh(P, 3.0, 2.0, 0) == pytest.approx(2 - 3 + 3 * math.log(1.5), rel=1e-14)


def test_barlow_beeston_kats():
    # (vi) w2 = 0 -> f = 0, t = -1, beta = 1, no penalty -> Poisson ; data = 0 -> mc*beta + penalty
    B, P = O.kBarlowBeeston, O.kPoisson
    for d, mc in [(3.0, 2.0), (10.0, 12.5), (1.0, 0.3)]:
        assert O.test_stat_llh(B, d, mc, 0.0) == pytest.approx(O.test_stat_llh(P, d, mc, 0.0), rel=1e-14)
    mc, w2 = 5.0, 0.8
    f2 = w2 / mc ** 2
    t = mc * f2 - 1
    beta = (-t + math.sqrt(t * t)) / 2
    assert O.test_stat_llh(B, 0.0, mc, w2) == pytest.approx(mc * beta + (beta - 1) ** 2 / (2 * f2), rel=1e-13)
    d = 7.0
    beta = (-t + math.sqrt(t * t + 4 * d * f2)) / 2
    want = mc * beta - d + d * math.log(d / (mc * beta)) + (beta - 1) ** 2 / (2 * f2)
    assert O.test_stat_llh(B, d, mc, w2) == pytest.approx(want, rel=1e-13)


def test_other_test_statistics():
    assert O.test_stat_llh(O.kPearson, 0, 4.0, 0) == 2.0
    assert O.test_stat_llh(O.kPearson, 6.0, 4.0, 0) == pytest.approx(4 / 8)
    assert O.test_stat_llh(O.kDembinskiAbdelmotteleb, 3.0, 2.0, 0) == O.test_stat_llh(O.kPoisson, 3.0, 2.0, 0)
    assert O.test_stat_llh(O.kIceCube, 3.0, 2.0, 0) == O.test_stat_llh(O.kPoisson, 3.0, 2.0, 0)
    # large effective MC statistics -> every MC-stat-aware statistic tends to Poisson
    p = O.test_stat_llh(O.kPoisson, 30.0, 25.0, 0)
    assert O.test_stat_llh(O.kBarlowBeeston, 30.0, 25.0, 1e-9) == pytest.approx(p, rel=1e-8)
    assert O.test_stat_llh(O.kDembinskiAbdelmotteleb, 30.0, 25.0, 1e-6) == pytest.approx(p, rel=1e-6)
    # IceCube is -ln L itself, not a ratio to the saturated model: tends to -ln Poisson(d; mc)
    full = 25.0 - 30.0 * math.log(25.0) + math.lgamma(31.0)
    assert O.test_stat_llh(O.kIceCube, 30.0, 25.0, 1e-3) == pytest.approx(full, rel=1e-4)


def test_find_bin_kats():
    # (viii) Samples/SampleStructs.h:577-613
    edges = [[np.array([0.0, 1.0, 2.5, 4.0, 10.0])]]
    sh = O.SampleHandlerFD(1, edges)
    for nom in (-1, 0, 1, 2, 3):
        assert sh.find_bin(0, 0, 0.0, nom) == 0          # == lower edge -> that bin
        assert sh.find_bin(0, 0, 1.0, nom) == 1
        assert sh.find_bin(0, 0, 9.999, nom) == 3
        assert sh.find_bin(0, 0, 10.0, nom) == -1        # == top edge -> overflow
        assert sh.find_bin(0, 0, -1e-9, nom) == -1
        assert sh.find_bin(0, 0, 3.0, nom) == 2


@pytest.fixture(scope="module")
def small():
    w = synth.CFG1.scaled(6000)
    mono, sh, d = O.build_from_workload(w)
    return w, mono, sh, d


def test_asimov_is_exactly_zero(small):
    # (vii) data := MC at nominal -> -lnL == 0 for Poisson (Fitters/MaCh3Factory.h:143-152)
    w, mono, sh, d = small
    sp, nm = synth.proposal(w, -1)
    mono.set_params(sp); sh.norm_vals[:] = nm
    sh.Reweight()
    sh.AddData(sh.mc.copy())
    assert sh.GetLikelihood() == 0.0
    assert sh.mc.sum() > 0


def test_zero_weight_events_contribute_nothing_and_w2_freeze(small):
    # (ix) w <= 0 skipped (Samples/SampleHandlerFD.cpp:432) ; (x) W2 frozen after the first Reweight (:342)
    w, mono, sh, d = small
    tw = sh.event_weights()
    assert (tw == 0).sum() > 0                  # the generator plants exact-zero osc weights
    bins = sh.event_bins()
    ref = np.zeros(sh.n_bins)
    ok = (tw > 0) & (bins >= 0)
    np.add.at(ref, bins[ok], tw[ok].astype(np.float64))
    np.testing.assert_allclose(sh.mc, ref, rtol=1e-12)
    w2_before = sh.w2.copy()
    sp, nm = synth.proposal(w, 5)
    mono.set_params(sp); sh.norm_vals[:] = nm
    sh.Reweight()
    np.testing.assert_array_equal(sh.w2, w2_before)
    assert not np.allclose(sh.mc, ref)


def test_per_sample_llh_sums_to_total():
    w = synth.SPARSE.scaled(5000)
    mono, sh, d = O.build_from_workload(w, update_w2=True)
    sp, nm = synth.proposal(w, -1)
    mono.set_params(sp); sh.norm_vals[:] = nm
    sh.Reweight()
    rng = np.random.default_rng(3)
    sh.AddData(rng.poisson(sh.mc).astype(np.float64))
    sp, nm = synth.proposal(w, 2)
    mono.set_params(sp); sh.norm_vals[:] = nm
    sh.Reweight()
    tot = sh.GetLikelihood()
    parts = [sh.GetSampleLikelihood(i) for i in range(w.n_samples)]
    assert tot == pytest.approx(sum(parts), rel=1e-12)
    assert tot > 0


def test_generator_is_chunk_invariant():
    w = synth.SPARSE.scaled(3000)
    full = synth.make_splines(w)
    a = synth.make_splines(w, 0, 1280)
    b = synth.make_splines(w, 1280, 3000)
    np.testing.assert_array_equal(np.concatenate([a["coeff_many"], b["coeff_many"]]), full["coeff_many"])
    np.testing.assert_array_equal(np.concatenate([a["paramNo_arr"], b["paramNo_arr"]]), full["paramNo_arr"])
    np.testing.assert_array_equal(np.concatenate([a["coeff_tf1"], b["coeff_tf1"]]), full["coeff_tf1"])
    cnt = full["nParamPerEvent"][0::2]
    assert cnt.min() < cnt.max()            # sparse: ragged responses per event
