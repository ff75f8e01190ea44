"""Pins the HOST side of the path and the test statistics to the REFERENCE'S OWN CODE, run in this container:
Splines/SplineMonolith.cpp (SMonolith: ScanMasterSpline, PrepareForGPU, CPU Evaluate = FindSplineSegment +
CalcSplineWeights + CalcTotalEventWeight), Splines/SplineBase.cpp (FindSplineSegment) and
Samples/SampleHandlerBase.cpp (GetTestStatLLH, GetPoissonLLH) were compiled from /root/reference where they lie
(oracle/ref_host: stand-ins only for the absent ROOT/spdlog/yaml-cpp headers) and driven on the seeded inputs of
tests/refpath_cases.py; the outputs are committed as tests/golden/ref_host_path.npz
(generator: tests/golden/make_ref_host_path.py).
  * CPU: the oracle, fed the monolith arrays THE REFERENCE BUILT, reproduces its segments, float parameter values
    and per-event weights bit for bit over 40 proposals per case (history-dependent segment cache included), and
    its five test statistics;
  * live: where oracle/_ref/libm3ref_path.so exists the reference is re-run and must reproduce the vectors;
  * GPU: libm3b200, fed the same arrays through m3b_upload_spline_monolith, gives the reference's segments and
    per-event weights, and its device test statistics the reference's per-bin values."""
import hashlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refpath_cases as RC  # noqa: E402
from oracle import binding as O
from oracle import ref_path_binding as RP

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_host_path.npz")
ARR = ("coeff_x", "coeff_many", "nKnots_arr", "paramNo_arr", "nParamPerEvent", "nParamPerEvent_tf1", "paramNo_tf1",
       "coeff_tf1", "n_pts", "x_pts_f64")
BOUNDARY_STEPS = (5, 9, 13, 17, 21, 25, 33)       # proposals placed on / one ulp off a knot (refpath_cases.make_case)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _arrays(g, name):
    a = {k: g[f"{name}/arr/{k}"] for k in ARR}
    a["n_events"] = int(g[f"{name}/sizes"][0])
    return a


def _oracle_monolith(g, name, double_knots=True):
    a = _arrays(g, name)
    P, K = int(g[f"{name}/sizes"][1]), int(g[f"{name}/sizes"][2])
    m = O.SMonolith(P, K, a["coeff_x"], a["n_pts"], a)
    if double_knots:
        m.set_knots_f64(a["x_pts_f64"])
    return m, a


def _digest(c):
    h = hashlib.sha256()
    for k in ("type", "npts", "vals", "pars"):
        h.update(np.ascontiguousarray(c[k]).tobytes())
    return np.frombuffer(h.digest(), np.uint8)


@pytest.mark.parametrize("name", RC.CASES)
def test_inputs_are_the_ones_the_vectors_were_made_from(gold, name):
    np.testing.assert_array_equal(_digest(RC.make_case(name)), gold[f"{name}/input_sha256"])


@pytest.fixture
def serial_oracle():
    O.set_multithread(False)                       # the serial build: strict left-to-right float product
    yield
    O.set_multithread(True)


@pytest.mark.parametrize("name", RC.CASES)
def test_oracle_host_path_matches_reference_vectors(gold, name, serial_oracle):
    m, _ = _oracle_monolith(gold, name)
    pars = RC.make_case(name)["pars"]
    for t in range(pars.shape[0]):
        m.set_params(pars[t])
        m.Evaluate()
        np.testing.assert_array_equal(m.segments, gold[f"{name}/segments"][t], err_msg=f"segments, step {t}")
        np.testing.assert_array_equal(m.param_values, gold[f"{name}/param_values"][t])
        np.testing.assert_array_equal(m.total_weights.view(np.uint32), gold[f"{name}/weights"][t].view(np.uint32),
                                      err_msg=f"weights, step {t}")


@pytest.mark.parametrize("name", RC.CASES)
def test_float_knots_differ_from_double_knots_only_at_knot_boundaries(gold, name):
    """coeff_x carries the knots as floats; FindSplineSegment compares against FastSplineInfo::xPts, doubles when the
    monolith was built in-process.  With float knots (a monolith reloaded from file, or _LOW_MEMORY_STRUCTS_) the
    segments can differ only where a proposal sits within a float ulp of a knot that is not exact in float."""
    m, a = _oracle_monolith(gold, name, double_knots=False)
    P, K = int(gold[f"{name}/sizes"][1]), int(gold[f"{name}/sizes"][2])
    x = a["x_pts_f64"].reshape(P, K)
    exact = np.array([np.array_equal(x[p], x[p].astype(np.float32).astype(np.float64)) for p in range(P)])
    pars = RC.make_case(name)["pars"]
    for t in range(pars.shape[0]):
        m.set_params(pars[t])
        m.FindSplineSegment()
        diff = m.segments != gold[f"{name}/segments"][t]
        assert not diff[exact].any()
        if t not in BOUNDARY_STEPS:
            assert not diff.any(), (t, np.nonzero(diff)[0])


def test_oracle_test_statistics_match_reference(gold):
    d, mc, w2 = gold["stat/data"], gold["stat/mc"], gold["stat/w2"]
    d2, mc2, w22 = RC.stat_inputs()
    np.testing.assert_array_equal(d, d2); np.testing.assert_array_equal(mc, mc2); np.testing.assert_array_equal(w2, w22)
    for kind in range(5):
        got = np.array([O.test_stat_llh(kind, d[i], mc[i], w2[i]) for i in range(d.size)])
        np.testing.assert_allclose(got, gold[f"stat/llh{kind}"], rtol=1e-14, atol=1e-300, err_msg=f"test statistic {kind}")
    np.testing.assert_array_equal(gold["stat/low_mc_bound"], [1e-5])
    # kPoisson is GetPoissonLLH except that it guards data == 0 differently: both were recorded
    assert np.isfinite(gold["stat/poisson"]).all()


@pytest.mark.skipif(not RP.available(), reason="oracle/_ref/libm3ref_path.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name", RC.CASES)
def test_reference_rerun_reproduces_the_vectors(gold, name):
    c = RC.make_case(name)
    m = RP.RefSMonolith(c["type"], c["npts"], c["vals"])
    try:
        got = m.arrays()
        for k in ARR:
            np.testing.assert_array_equal(got[k], gold[f"{name}/arr/{k}"], err_msg=k)
        for t in range(c["pars"].shape[0]):
            w, s, v = m.evaluate(c["pars"][t])
            np.testing.assert_array_equal(s, gold[f"{name}/segments"][t])
            np.testing.assert_array_equal(v, gold[f"{name}/param_values"][t])
            np.testing.assert_array_equal(w.view(np.uint32), gold[f"{name}/weights"][t].view(np.uint32))
    finally:
        m.close()


@pytest.mark.skipif(not RP.available(), reason="oracle/_ref/libm3ref_path.so not built (needs /root/reference at build time)")
def test_reference_rerun_reproduces_the_test_statistics(gold):
    for kind in range(5):
        v, thrown = RP.test_stat(kind, gold["stat/data"], gold["stat/mc"], gold["stat/w2"])
        assert thrown == 0
        np.testing.assert_array_equal(v, gold[f"stat/llh{kind}"])


# ---- device -----------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", RC.CASES)
def test_device_segments_and_weights_match_reference_host(gold, name):
    """m3b_upload_spline_monolith consumes the arrays SMonolith::PrepareForGPU built (unused tail entries, padded
    coeff_x rows, parameters without any spline included); per step the segments equal the reference's and the
    per-event spline weight equals SMonolith::Evaluate's cpu_total_weights: same fmaf chain, same product order."""
    from mach3_b200 import lib as L
    a = _arrays(gold, name)
    E, P, K = (int(v) for v in gold[f"{name}/sizes"][:3])
    h = L.Handle(flags=L.FLAG_KEEP_EVENT_WEIGHTS)
    h.upload_binning([[np.array([0.0, 1.0])]])
    h.upload_spline_monolith(P, K, a["coeff_x"], a["n_pts"], a)
    h.set_spline_knots_f64(a["x_pts_f64"])
    h.upload_events(np.zeros(E, np.int32), np.full(E, 0.5))
    pars = RC.make_case(name)["pars"]
    worst = 0.0
    for t in range(pars.shape[0]):
        seg, val = h.find_segments(pars[t])
        np.testing.assert_array_equal(seg, gold[f"{name}/segments"][t], err_msg=f"segments, step {t}")
        np.testing.assert_array_equal(val, gold[f"{name}/param_values"][t])
        h.step(pars[t])
        h.llh()
        sw, tw = h.read_event_weights()
        ref = gold[f"{name}/weights"][t]
        worst = max(worst, float(np.max(np.abs(sw - ref) / np.maximum(np.abs(ref), 1e-30))))
        np.testing.assert_array_equal(sw.view(np.uint32), ref.view(np.uint32), err_msg=f"weights, step {t}")
    assert worst <= 1e-5            # north_star's bound; the assertion above is the stronger, bit-exact one
    h.close()


@pytest.mark.gpu
def test_device_test_statistics_match_reference(gold):
    """Every (data, mc, w2) triple gets a one-bin sample of its own: the device's per-sample -lnL is the reference's
    GetTestStatLLH of that bin.  The histogram is written straight into the device buffer (the multi-GPU entry point,
    m3b_hist_device_ptr + m3b_llh_from_hist), so the statistics are tested on exactly the reference's inputs."""
    import torch
    from mach3_b200 import lib as L
    from mach3_b200.sharding import _DevArray
    d, mc, w2 = gold["stat/data"], gold["stat/mc"], gold["stat/w2"]
    NS = 64
    for kind in range(5):
        h = L.Handle(test_statistic=kind, update_w2=True, flags=L.FLAG_NO_FUSED_LLH)
        h.upload_binning([[np.array([0.0, 1.0])] for _ in range(NS)])
        h.upload_events(np.arange(NS, dtype=np.int32), np.full(NS, 0.5))
        for b in range(d.size // NS):
            sl = slice(b * NS, (b + 1) * NS)
            h.upload_data(d[sl])
            h.step(np.zeros(0), mode="fill")
            h.synchronize()
            ptr, nb, live = h.hist_device_ptr()
            assert nb == NS and live == 1
            hist = torch.as_tensor(_DevArray(ptr, 2 * nb), device="cuda:0")
            hist.copy_(torch.from_numpy(np.concatenate([mc[sl], w2[sl]])))
            torch.cuda.synchronize()
            h.llh_from_hist()
            tot, per = h.llh(per_sample=True)
            ref = gold[f"stat/llh{kind}"][sl]
            # device log / lgamma are a few ulp off glibc's; north_star's bound on -lnL is 1e-6 relative
            np.testing.assert_allclose(per, ref, rtol=1e-10, atol=1e-13, err_msg=f"test statistic {kind}, batch {b}")
            assert tot == pytest.approx(ref.sum(), rel=1e-10)
        h.close()


# =====================================================================================================================
# The full per-step path against the reference's own SampleHandlerFD::Reweight / GetLikelihood
# (tests/golden/ref_host_fd.npz, generator tests/golden/make_ref_host_fd.py; both reference builds)
# =====================================================================================================================
GOLD_FD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_host_fd.npz")
BARLOW_BEESTON = 1


@pytest.fixture(scope="module")
def gold_fd():
    return np.load(GOLD_FD)


@pytest.mark.parametrize("update_w2", [False, True])
def test_oracle_full_path_matches_reference_sample_handler(gold, gold_fd, update_w2, serial_oracle):
    """Monolith + SampleHandlerFD, _LOW_MEMORY_STRUCTS_ build (the only one that wires SMonolith in): per-event
    CalcWeightTotal, FindGlobalBin (three binnings, events outside / on edges, shifted kinematics with stale nominal
    bins), the `<= 0` skip, the W2 freeze and Barlow-Beeston -lnL, step by step."""
    tag = f"mono_w2{int(update_w2)}"
    f = RC.fd_case()
    E = f["sample_id"].size
    mono, a = _oracle_monolith(gold, "mixed", double_knots=False)       # float build: FastSplineInfo::xPts are floats
    sh = O.SampleHandlerFD(E, RC.fd_edges(), BARLOW_BEESTON, update_w2)
    assert sh.n_bins == int(gold_fd[f"{tag}/n_bins"][0])
    norm, osc = np.ones(RC.N_NORM), np.ones(E, np.float32)
    sh.set_events(f["sample_id"], f["kin"].reshape(-1).copy(), f["norm_idx"].reshape(-1), RC.NPE, norm, osc, mono, f["static_w"])
    for t in range(RC.FD_STEPS):
        if t == 8:
            sh._keep[1][:] = f["kin_shift"].reshape(-1)           # the KinVar pointers look into this array
        mono.set_params(f["pars"][t]); sh.norm_vals[:] = f["norm"][t]; sh.osc_w[:] = f["osc"][t]
        sh.Reweight()
        if t == 0:
            sh.AddData(gold_fd[f"{tag}/data"])
        np.testing.assert_array_equal(mono.segments, gold_fd[f"{tag}/segments"][t])
        np.testing.assert_array_equal(sh.event_bins(), gold_fd[f"{tag}/event_bin"][t], err_msg=f"bins, step {t}")
        np.testing.assert_array_equal(sh.event_weights().view(np.uint32), gold_fd[f"{tag}/event_w"][t].view(np.uint32),
                                      err_msg=f"event weights, step {t}")
        np.testing.assert_array_equal(sh.mc, gold_fd[f"{tag}/mc"][t], err_msg=f"mc, step {t}")          # same summation order
        np.testing.assert_array_equal(sh.w2, gold_fd[f"{tag}/w2"][t], err_msg=f"w2, step {t}")
        assert sh.GetLikelihood() == pytest.approx(float(gold_fd[f"{tag}/llh"][t]), rel=1e-14)
        got = np.array([sh.GetSampleLikelihood(s) for s in range(3)])
        np.testing.assert_allclose(got, gold_fd[f"{tag}/sample_llh"][t], rtol=1e-14)


@pytest.mark.parametrize("build", ["float", "double"])
def test_oracle_binned_path_matches_reference_sample_handler(gold_fd, build, serial_oracle):
    """BinnedSplineHandler::Evaluate (segments, weightvec_Monolith incl. the clamp at 0) + SampleHandlerFD in both
    reference builds: M3::float_t = float and the default double."""
    from mach3_b200.synth import binned as B
    tag = f"binned_{build}"
    w = RC.binned_workload()
    f64 = build == "double"
    b, sh, d = O.build_binned_from_workload(w, update_w2=True, test_statistic=BARLOW_BEESTON, f64=f64)
    for i, step in enumerate(RC.BINNED_STEPS):
        sp, nm = B.proposal(w, step)
        b.set_params(sp); sh.norm_vals[:] = nm; sh.osc_w[:] = B.make_osc(w, max(step, 0), f64=f64)
        sh.Reweight()
        if i == 0:
            sh.AddData(gold_fd[f"{tag}/data"])
        np.testing.assert_array_equal(b.segments, gold_fd[f"{tag}/segments"][i])
        np.testing.assert_array_equal(b.weights, gold_fd[f"{tag}/slot_w"][i], err_msg=f"weightvec_Monolith, step {step}")
        np.testing.assert_array_equal(sh.event_weights(), gold_fd[f"{tag}/event_w"][i], err_msg=f"event weights, step {step}")
        np.testing.assert_array_equal(sh.mc, gold_fd[f"{tag}/mc"][i])
        np.testing.assert_array_equal(sh.w2, gold_fd[f"{tag}/w2"][i])
        assert sh.GetLikelihood() == pytest.approx(float(gold_fd[f"{tag}/llh"][i]), rel=1e-14)


@pytest.mark.skipif(not RP.available(), reason="oracle/_ref/libm3ref_path*.so not built (needs /root/reference at build time)")
def test_reference_rerun_reproduces_the_sample_handler_vectors(gold_fd):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_ref_host_fd as G
    out = {}
    G.run_monolith(out, False); G.run_monolith(out, True); G.run_binned(out, "float"); G.run_binned(out, "double")
    assert set(out) == set(gold_fd.files)
    for k, v in out.items():
        np.testing.assert_array_equal(np.asarray(v), gold_fd[k], err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("update_w2", [False, True])
def test_device_full_path_matches_reference_sample_handler(gold, gold_fd, update_w2):
    """libm3b200's fused step against the reference's SampleHandlerFD::Reweight + GetLikelihood: bins and per-event
    weights bit for bit, histograms to 1e-12 (f64 atomics reorder the sum), -lnL to 1e-10."""
    from mach3_b200 import handlers
    tag = f"mono_w2{int(update_w2)}"
    f = RC.fd_case()
    E = f["sample_id"].size
    a = _arrays(gold, "mixed")
    P, K = int(gold["mixed/sizes"][1]), int(gold["mixed/sizes"][2])
    sh = handlers.SampleHandlerFD(RC.fd_edges(), BARLOW_BEESTON, update_w2, keep_event_weights=True, keep_kinematics=True)
    sh.SetupSplines(P, K, a["coeff_x"], a["n_pts"], a)               # float build: the float knots of coeff_x
    pars, norm, osc = np.zeros(P), np.ones(RC.N_NORM), np.ones(E, np.float32)
    sh.SetupEvents(f["sample_id"], f["kin"].reshape(-1), f["norm_idx"].reshape(-1), RC.NPE, norm, osc, None, f["static_w"])
    sh.SetSplinePointers(pars)
    assert sh.n_bins == int(gold_fd[f"{tag}/n_bins"][0])
    for t in range(RC.FD_STEPS):
        if t == 8:
            sh.handle.update_kinematics(f["kin_shift"].reshape(-1))
        pars[:] = f["pars"][t]; norm[:] = f["norm"][t]; osc[:] = f["osc"][t]
        sh.OscillatorEvaluated()
        sh.Reweight()
        if t == 0:
            sh.GetLikelihood()
            sh.AddData(gold_fd[f"{tag}/data"])
        llh = sh.GetLikelihood()
        np.testing.assert_array_equal(sh.handle.read_event_bins(), gold_fd[f"{tag}/event_bin"][t], err_msg=f"bins, step {t}")
        sw, tw = sh.handle.read_event_weights()
        ref_w = gold_fd[f"{tag}/event_w"][t]
        live = ref_w > 0                                        # skipped events: the device leaves w <= 0 as computed
        np.testing.assert_array_equal(tw[live].view(np.uint32), ref_w[live].view(np.uint32), err_msg=f"event weights, step {t}")
        mc, w2 = sh.handle.read_hist()
        np.testing.assert_allclose(mc, gold_fd[f"{tag}/mc"][t], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(w2, gold_fd[f"{tag}/w2"][t], rtol=1e-12, atol=1e-13)
        if t > 0:
            assert llh == pytest.approx(float(gold_fd[f"{tag}/llh"][t]), rel=1e-10)
            per = sh.handle.llh(per_sample=True)[1]
            np.testing.assert_allclose(per, gold_fd[f"{tag}/sample_llh"][t], rtol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("build", ["float", "double"])
def test_device_binned_path_matches_reference_sample_handler(gold_fd, build):
    from mach3_b200 import handlers
    from mach3_b200.synth import binned as B
    tag = f"binned_{build}"
    w = RC.binned_workload()
    f64 = build == "double"
    sh, d = handlers.build_binned_from_workload(w, update_w2=True, test_statistic=BARLOW_BEESTON, keep_event_weights=True, f64=f64)
    for i, step in enumerate(RC.BINNED_STEPS):
        sp, nm = B.proposal(w, step)
        d["pars"][:] = sp; d["norm"][:] = nm; d["osc"][:] = B.make_osc(w, max(step, 0), f64=f64)
        sh.OscillatorEvaluated()
        sh.Reweight()
        if i == 0:
            sh.GetLikelihood()
            sh.AddData(gold_fd[f"{tag}/data"])
        llh = sh.GetLikelihood()
        np.testing.assert_array_equal(np.asarray(sh.SplineHandler.weightvec_Monolith), gold_fd[f"{tag}/slot_w"][i],
                                      err_msg=f"weightvec_Monolith, step {step}")
        mc, w2 = sh.handle.read_hist()
        np.testing.assert_allclose(mc, gold_fd[f"{tag}/mc"][i], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(w2, gold_fd[f"{tag}/w2"][i], rtol=1e-12, atol=1e-13)
        if i > 0:
            assert llh == pytest.approx(float(gold_fd[f"{tag}/llh"][i]), rel=1e-10)


@pytest.mark.skipif(not RP.available(), reason="oracle/_ref/libm3ref_path*.so not built (needs /root/reference at build time)")
def test_array_fed_reference_monolith_equals_the_constructed_one(gold):
    """bench.py's CPU baseline hands the reference's SMonolith its arrays directly (millions of responses would not
    fit as heap objects): same Evaluate() results as the monolith the reference's constructor builds."""
    c = RC.make_case("mixed")
    a = _arrays(gold, "mixed")
    P, K = int(gold["mixed/sizes"][1]), int(gold["mixed/sizes"][2])
    spl = dict(a); spl["nKnots_arr"] = a["nKnots_arr"].astype(np.uint64)
    m0 = RP.RefSMonolith(c["type"], c["npts"], c["vals"], build="float")
    m1 = RP.RefSMonolith.from_arrays(P, K, a["coeff_x"], a["n_pts"], c["type"].astype(np.int8), spl, build="float")
    try:
        for t in range(0, 40, 3):
            w0, s0, v0 = m0.evaluate(c["pars"][t])
            w1, s1, v1 = m1.evaluate(c["pars"][t])
            np.testing.assert_array_equal(s0, s1)
            np.testing.assert_array_equal(w0.view(np.uint32), w1.view(np.uint32))
    finally:
        m0.close(); m1.close()


@pytest.mark.skipif(not RP.available_mt(), reason="oracle/_ref/libm3ref_path_lm_mt.so not built")
def test_reference_release_build_agrees_with_the_oracle_multithread_path():
    """The library bench.py times as cpu_baseline kind "reference" (release flags, MULTITHREAD) against the oracle's
    MULTITHREAD arm on a cfg2 sample: reassociated float products and sums, so 1e-6 on -lnL (north_star's bound)."""
    from mach3_b200 import synth
    w = synth.CFG2.scaled(20_000)
    typ, npts, cx = synth.param_layout(w)
    spl, ev = synth.make_splines(w), synth.make_events(w)
    mono = RP.RefSMonolith.from_arrays(w.n_params, w.n_knots, cx, npts, typ, spl, build="float_mt")
    fd = RP.RefSampleHandlerFD(synth.bin_edges(w), w.test_statistic, False, build="float_mt")
    fd.attach_monolith(mono)
    E = w.n_events
    idx = np.arange(E, dtype=np.int32)
    fd.set_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, w_before=idx, w_after=E + idx, n_pool=2 * E)
    pool = np.concatenate([synth.make_osc(w, 0), ev["static_w"]]).astype(np.float64)
    O.set_multithread(True)
    omono, osh, od = O.build_from_workload(w)
    data = None
    for step in (-1, 0, 1):
        sp, nm = synth.proposal(w, step)
        fd.reweight(sp, nm, pool)
        omono.set_params(sp); osh.norm_vals[:] = nm
        osh.Reweight()
        if data is None:
            data = np.random.default_rng(1).poisson(osh.mc).astype(np.float64)
            fd.set_data(data); osh.AddData(data)
        np.testing.assert_allclose(fd.hist()[0], osh.mc, rtol=1e-6, atol=1e-9)
        assert fd.llh() == pytest.approx(osh.GetLikelihood(), rel=1e-6, abs=1e-9)
    fd.close()


# =====================================================================================================================
# BASELINE config 1 at full size (the reference's own CPU-runnable case) against the reference's own code
# =====================================================================================================================
GOLD_CFG1 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_host_baseline.npz")
CFG1_STEPS = (-1, 0, 1, 2, 3, -2, -3, 4)
BASELINE_SHAPES = ("cfg1", "cfg2s", "cfg3s")


def _shape(name):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_ref_host_baseline as G
    return G, G.workloads()[name]


@pytest.mark.parametrize("name", BASELINE_SHAPES)
def test_oracle_baseline_shapes_match_reference(name, serial_oracle):
    """cfg1 at full size, cfg2 / cfg3 shapes on samples: the oracle against the reference's own Reweight +
    GetLikelihood, histogram bit for bit."""
    from mach3_b200 import synth
    g = np.load(GOLD_CFG1)
    _, w = _shape(name)
    mono, sh, d = O.build_from_workload(w)
    for i, step in enumerate(CFG1_STEPS):
        sp, nm = synth.proposal(w, step)
        mono.set_params(sp); sh.norm_vals[:] = nm
        sh.Reweight()
        if i == 0:
            sh.AddData(g[f"{name}/data"])
        np.testing.assert_array_equal(sh.mc, g[f"{name}/mc"][i], err_msg=f"step {step}")
        assert sh.GetLikelihood() == pytest.approx(float(g[f"{name}/llh"][i]), rel=1e-14)


@pytest.mark.skipif(not RP.available(), reason="oracle/_ref/libm3ref_path_lm.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name", BASELINE_SHAPES)
def test_reference_rerun_reproduces_baseline_shapes(name):
    G, w = _shape(name)
    out, g = G.run(w), np.load(GOLD_CFG1)
    for k in ("data", "mc", "llh"):
        np.testing.assert_array_equal(out[k], g[f"{name}/{k}"], err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", BASELINE_SHAPES)
def test_device_baseline_shapes_match_reference(name):
    """The fused step against the reference's own Reweight + GetLikelihood on BASELINE's shapes (config 1 at full
    size).  Bars of north_star: -lnL within 1e-6 relative (asserted: 1e-10), histogram 1e-12."""
    from mach3_b200 import handlers, synth
    g = np.load(GOLD_CFG1)
    _, w = _shape(name)
    sh, d = handlers.build_from_workload(w)
    for i, step in enumerate(CFG1_STEPS):
        sp, nm = synth.proposal(w, step)
        d["pars"][:] = sp; d["norm"][:] = nm
        sh.Reweight()
        if i == 0:
            sh.GetLikelihood(); sh.AddData(g[f"{name}/data"])
        llh = sh.GetLikelihood()
        np.testing.assert_allclose(sh.handle.read_hist()[0], g[f"{name}/mc"][i], rtol=1e-12, atol=1e-12, err_msg=f"step {step}")
        if i > 0:
            assert llh == pytest.approx(float(g[f"{name}/llh"][i]), rel=1e-10)


@pytest.mark.gpu
@pytest.mark.skipif(not RP.available_mt(), reason="oracle/_ref/libm3ref_path_lm_mt.so not built (needs /root/reference at build time)")
def test_full_size_cfg2_against_the_reference_itself():
    """BASELINE config 2 at FULL size (1M events x 50 responses, 900 bins), live: the reference's own multithreaded
    SampleHandlerFD::Reweight + GetLikelihood (release build of its sources, on the box's host cores) next to the fused
    B200 step, same proposals.  north_star's bar on -lnL: 1e-6 relative (the MULTITHREAD build reassociates float
    products and double sums, so not bit-exact)."""
    from mach3_b200 import handlers, synth
    w = synth.CFG2
    typ, npts, cx = synth.param_layout(w)
    spl, ev = synth.make_splines(w), synth.make_events(w)
    mono = RP.RefSMonolith.from_arrays(w.n_params, w.n_knots, cx, npts, typ, spl, build="float_mt")
    fd = RP.RefSampleHandlerFD(synth.bin_edges(w), w.test_statistic, False, build="float_mt")
    fd.attach_monolith(mono)
    E = w.n_events
    idx = np.arange(E, dtype=np.int32)
    fd.set_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, w_before=idx, w_after=E + idx, n_pool=2 * E)
    del spl
    pool = np.concatenate([synth.make_osc(w, 0), ev["static_w"]]).astype(np.float64)
    gsh, gd = handlers.build_from_workload(w)
    worst = 0.0
    for i, step in enumerate((-1, 0, 1, 2, 3)):
        sp, nm = synth.proposal(w, step)
        fd.reweight(sp, nm, pool if i == 0 else None)
        gd["pars"][:] = sp; gd["norm"][:] = nm
        gsh.Reweight()
        if i == 0:
            data = np.random.default_rng(w.seed).poisson(fd.hist()[0]).astype(np.float64)
            fd.set_data(data); gsh.GetLikelihood(); gsh.AddData(data)
        r, g = fd.llh(), gsh.GetLikelihood()
        np.testing.assert_allclose(gsh.handle.read_hist()[0], fd.hist()[0], rtol=1e-6, atol=1e-9)
        if i > 0:
            worst = max(worst, abs(g - r) / abs(r))
            assert abs(g - r) <= 1e-6 * abs(r), (step, g, r)
    print(f"full-size cfg2: worst relative -lnL difference to the reference {worst:.2e}")
    fd.close()
