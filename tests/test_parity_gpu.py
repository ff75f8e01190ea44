"""Parity of the B200 path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): segment indices and bin assignments bit-exact; per-event weights
within 1e-5 relative; total -lnL within 1e-6 relative.
"""
import numpy as np
import pytest

from mach3_b200 import handlers, lib, synth
from oracle import binding as O

pytestmark = pytest.mark.gpu

W_RTOL = 1e-5      # per-event weights (fp32 coefficients)          -- north_star bar
LLH_RTOL = 1e-6    # total -lnL                                      -- north_star bar
HIST_RTOL = 1e-6   # histogram vs the MULTITHREAD oracle (its simd products reassociate)
ORDER_RTOL = 1e-12 # f64 histogram when only the summation order differs


@pytest.fixture(autouse=True, params=["serial", "multithread"])
def oracle_build(request):
    """Every parity test runs against both restated builds of the reference: the serial build
    (strict left-to-right float products: the device weights must be BIT-EXACT) and the
    MULTITHREAD build (OpenMP simd reductions reassociate: north_star tolerances)."""
    O.set_multithread(request.param == "multithread")
    yield request.param
    O.set_multithread(True)


def _exact():
    return not bool(O.lib().m3o_get_multithread())


def _pair(w, **kw):
    mono, osh, od = O.build_from_workload(w, update_w2=kw.get("update_w2", False),
                                          test_statistic=kw.get("test_statistic"))
    gsh, gd = handlers.build_from_workload(w, keep_event_weights=True, **kw)
    return mono, osh, gsh, gd


def _set(w, step, mono, osh, gsh, gd, osc_step=None):
    sp, nm = synth.proposal(w, step)
    mono.set_params(sp)
    osh.norm_vals[:] = nm
    gd["pars"][:] = sp
    gd["norm"][:] = nm
    if osc_step is not None:
        osc = synth.make_osc(w, osc_step)
        osh.osc_w[:] = osc
        gd["osc"][:] = osc
        gsh.OscillatorEvaluated()


def _check_step(w, mono, osh, gsh, check_weights=True):
    osh.Reweight()
    gsh.Reweight()
    o_llh, g_llh = osh.GetLikelihood(), gsh.GetLikelihood()
    seg, val = gsh.SplineHandler.handle.find_segments(gsh._spline_pars)  # idempotent given the cached segment
    np.testing.assert_array_equal(seg, mono.segments)                    # bit-exact integers
    np.testing.assert_array_equal(val, mono.param_values)
    mc, w2 = gsh.GetMCArray(), gsh.GetW2Array()
    if check_weights:
        sw, tw = gsh.GetEventWeight()
        if _exact():
            np.testing.assert_array_equal(sw, mono.total_weights)        # bit-exact
            np.testing.assert_array_equal(tw, osh.event_weights())
        else:
            np.testing.assert_allclose(sw, mono.total_weights, rtol=W_RTOL, atol=0)
            np.testing.assert_allclose(tw, osh.event_weights(), rtol=W_RTOL, atol=0)
        # the fill itself: histogram of the device's own weights, recomputed in numpy f64
        bins = gsh.GetEventBins()
        ok = (tw > 0) & (bins >= 0)
        ref = np.zeros(gsh.n_bins)
        np.add.at(ref, bins[ok], tw[ok].astype(np.float64))
        np.testing.assert_allclose(mc, ref, rtol=ORDER_RTOL, atol=1e-13)
    rt = ORDER_RTOL if _exact() else HIST_RTOL
    np.testing.assert_allclose(mc, osh.mc, rtol=rt, atol=1e-12)
    np.testing.assert_allclose(w2, osh.w2, rtol=rt, atol=1e-12)
    assert g_llh == pytest.approx(o_llh, rel=1e-10 if _exact() else LLH_RTOL, abs=1e-9)
    return o_llh, g_llh


@pytest.mark.parametrize("tile", [256, 512, 1024])
def test_cfg1_shape_parity(tile):
    """BASELINE config 1 shape (10 TSpline3 K=5 + 2 TF1, 1-D 50 bins, Poisson), reduced to 30k events."""
    w = synth.CFG1.scaled(30_011)          # ragged last tile
    mono, osh, gsh, gd = _pair(w, tile_events=tile)
    np.testing.assert_array_equal(gsh.GetEventBins(), osh.event_bins())   # bit-exact bins
    _set(w, -1, mono, osh, gsh, gd)
    _check_step(w, mono, osh, gsh)
    asimov = osh.mc.copy()
    osh.AddData(asimov); gsh.AddData(asimov)
    assert gsh.GetMCArray().sum() > 0
    # Asimov: the device LLH of the device histogram against its own copy is exactly zero
    gsh.AddData(gsh.GetMCArray())
    gsh.Reweight()
    assert abs(gsh.GetLikelihood()) < 1e-10
    gsh.AddData(asimov)
    for step in range(6):
        _set(w, step, mono, osh, gsh, gd, osc_step=step)
        o, g = _check_step(w, mono, osh, gsh)
        assert o > 0


def test_special_proposals_on_knots_and_out_of_range():
    w = synth.CFG1.scaled(8_000)
    mono, osh, gsh, gd = _pair(w)
    _set(w, -1, mono, osh, gsh, gd); _check_step(w, mono, osh, gsh)
    data = np.random.default_rng(1).poisson(osh.mc).astype(np.float64)
    osh.AddData(data); gsh.AddData(data)
    # nominal (on a knot) after a random step and before: history-dependent segment must agree
    for step in (-2, 3, -1, -3, 4, -4, -2, -1):
        _set(w, step, mono, osh, gsh, gd)
        _check_step(w, mono, osh, gsh)


@pytest.mark.parametrize("wl", ["SPARSE", "SPARSE_RUNS"])
def test_sparse_multi_sample_barlow_beeston_live_w2(wl):
    """Ragged responses (interaction-mode sparsity), 3 samples, Barlow-Beeston, UpdateW2=true."""
    w = getattr(synth, wl)
    mono, osh, gsh, gd = _pair(w, update_w2=True)
    if wl == "SPARSE_RUNS":
        assert gsh.handle.info().n_signatures > 3      # tiles with different parameter sets
    np.testing.assert_array_equal(gsh.GetEventBins(), osh.event_bins())
    _set(w, -1, mono, osh, gsh, gd); _check_step(w, mono, osh, gsh)
    data = np.random.default_rng(2).poisson(osh.mc).astype(np.float64)
    osh.AddData(data); gsh.AddData(data)
    for step in range(4):
        _set(w, step, mono, osh, gsh, gd, osc_step=step)
        _check_step(w, mono, osh, gsh)
        tot, parts = gsh.handle.llh(per_sample=True)
        for i in range(w.n_samples):
            assert parts[i] == pytest.approx(osh.GetSampleLikelihood(i), rel=LLH_RTOL, abs=1e-9)
        assert tot == pytest.approx(parts.sum(), rel=1e-12)


def test_w2_frozen_after_first_reweight_by_default():
    w = synth.SPARSE.scaled(9_000)
    mono, osh, gsh, gd = _pair(w, update_w2=False)
    _set(w, -1, mono, osh, gsh, gd); _check_step(w, mono, osh, gsh)
    w2_first = gsh.GetW2Array().copy()
    assert w2_first.sum() > 0
    for step in range(3):
        _set(w, step, mono, osh, gsh, gd)
        _check_step(w, mono, osh, gsh)
        np.testing.assert_array_equal(gsh.GetW2Array(), w2_first)
    gsh.handle.reset_w2()
    gsh.Reweight(); gsh.GetLikelihood()
    assert not np.array_equal(gsh.GetW2Array(), w2_first)


@pytest.mark.parametrize("ts", [lib.POISSON, lib.BARLOW_BEESTON, lib.PEARSON, lib.DEMBINSKI_ABDELMOTTELEB, lib.ICECUBE])
def test_every_test_statistic(ts):
    w = synth.SPARSE.scaled(12_000)
    mono, osh, gsh, gd = _pair(w, update_w2=True, test_statistic=ts)
    _set(w, -1, mono, osh, gsh, gd); osh.Reweight()
    data = np.random.default_rng(5).poisson(osh.mc).astype(np.float64)
    osh.AddData(data); gsh.AddData(data)
    _set(w, 1, mono, osh, gsh, gd)
    osh.Reweight(); gsh.Reweight()
    # IceCube is evaluated in long double by the reference and in f64 on the device
    rel = 1e-6 if ts != lib.ICECUBE else 1e-5
    assert gsh.GetLikelihood() == pytest.approx(osh.GetLikelihood(), rel=rel)


def test_chunked_upload_and_reference_typed_upload_give_identical_results():
    w = synth.SPARSE.scaled(7_777)
    a, ad = handlers.build_from_workload(w, keep_event_weights=True)
    b, bd = handlers.build_from_workload(w, keep_event_weights=True, chunk_events=1024)
    sp, nm = synth.proposal(w, 2)
    for sh, d in ((a, ad), (b, bd)):
        d["pars"][:] = sp; d["norm"][:] = nm
        sh.Reweight()
    np.testing.assert_array_equal(a.GetEventWeight()[0], b.GetEventWeight()[0])
    np.testing.assert_array_equal(a.GetEventWeight()[1], b.GetEventWeight()[1])
    # one-shot upload with the reference's unsigned-int knot offsets
    h = lib.Handle(flags=lib.FLAG_KEEP_EVENT_WEIGHTS)
    spl = dict(ad["spl"]); spl["nKnots_arr"] = spl["nKnots_arr"].astype(np.uint32)
    h.upload_spline_monolith(w.n_params, w.n_knots, ad["coeff_x"], ad["npts"], spl)
    h.upload_binning(synth.bin_edges(w))
    ev = ad["ev"]
    h.upload_events(ev["sample_id"], ev["kin"])
    h.step(sp)
    np.testing.assert_array_equal(h.read_event_weights()[0], a.GetEventWeight()[0])


def test_step_segments_matches_smonolithgpu_contract():
    """RunGPU_SplineMonolith's contract: the caller passes ParamValues + SplineSegments."""
    w = synth.CFG1.scaled(5_000)
    mono, osh, gsh, gd = _pair(w)
    _set(w, 2, mono, osh, gsh, gd)
    osh.Reweight()
    gsh.handle.step_segments(mono.param_values, mono.segments, gd["norm"], gd["osc"])
    sw, tw = gsh.GetEventWeight()
    np.testing.assert_allclose(sw, mono.total_weights, rtol=0 if _exact() else W_RTOL)
    assert gsh.GetLikelihood() == pytest.approx(osh.GetLikelihood(), rel=LLH_RTOL, abs=1e-9)


def test_standalone_monolith_like_pymach3():
    """pyMaCh3 EventSplineMonolith usage: set_param_value_array / evaluate / get_event_weight."""
    w = synth.SPARSE.scaled(4_000)
    typ, npts, cx = synth.param_layout(w)
    spl = synth.make_splines(w)
    omono = O.SMonolith(w.n_params, w.n_knots, cx, npts, spl)
    gmono = handlers.SMonolith(w.n_params, w.n_knots, cx, npts, spl)
    pars = np.zeros(w.n_params)
    gmono.setSplinePointers(pars)
    for step in (0, 1, -2):
        pars[:] = synth.proposal(w, step)[0]
        omono.set_params(pars); omono.Evaluate()
        gmono.Evaluate(); gmono.SynchroniseMemTransfer()
        np.testing.assert_allclose(gmono.cpu_total_weights, omono.total_weights, rtol=0 if _exact() else W_RTOL)
        assert gmono.retPointer(17)[0] == gmono.cpu_total_weights[17]


def test_no_splines_only_norm_and_osc():
    w = synth.CFG1.scaled(3_000)
    ev = synth.make_events(w)
    osc = synth.make_osc(w, 0)
    norm = np.array(synth.proposal(w, 4)[1])
    osh = O.SampleHandlerFD(w.n_events, synth.bin_edges(w))
    osh.set_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, norm, osc, None, ev["static_w"])
    osh.Reweight()
    gsh = handlers.SampleHandlerFD(synth.bin_edges(w), keep_event_weights=True)
    gsh.SetupEvents(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, norm, osc, None, ev["static_w"])
    gsh.Reweight()
    np.testing.assert_allclose(gsh.GetEventWeight()[1], osh.event_weights(), rtol=0 if _exact() else W_RTOL)
    np.testing.assert_allclose(gsh.GetMCArray(), osh.mc, rtol=ORDER_RTOL if _exact() else HIST_RTOL)


def test_knot_count_mismatch_is_an_error():
    w = synth.CFG1.scaled(600)
    typ, npts, cx = synth.param_layout(w)
    spl = synth.make_splines(w)
    bad = npts.copy(); bad[0] = 4
    h = lib.Handle()
    with pytest.raises(lib.M3BError) as ei:
        h.splines_begin(w.n_params, w.n_knots, cx, bad, w.n_events)
        h.splines_append(spl)
    assert ei.value.code == 4


def test_full_size_cfg2_properties():
    """BASELINE config 2 at full size (1M events x 50 responses): size-independent properties."""
    w = synth.CFG2
    gsh, gd = handlers.build_from_workload(w)
    sp, nm = synth.proposal(w, -1)
    gd["pars"][:] = sp; gd["norm"][:] = nm
    gsh.Reweight(); gsh.GetLikelihood()
    asimov = gsh.GetMCArray()
    gsh.AddData(asimov)
    gsh.Reweight()
    # Asimov: zero up to the f64 rounding of a differently-ordered atomic histogram sum
    assert abs(gsh.GetLikelihood()) < 1e-9
    # histogram total == sum of positive in-range event weights, computed independently in numpy
    osc = gd["osc"].astype(np.float64)
    assert asimov.sum() > 0 and np.isfinite(asimov).all()
    # linearity in a norm parameter that every event of a class carries: scaling all norms by 2
    # scales every event weight by 2^n_norm_per_event exactly (powers of two are exact in fp32)
    gd["norm"][:] = 2.0 * nm
    gsh.Reweight(); gsh.GetLikelihood()
    np.testing.assert_allclose(gsh.GetMCArray(), asimov * 2.0 ** w.n_norm_per_event, rtol=1e-12)
    # determinism of the segment path and near-determinism of the f64 histogram across repeats
    gd["pars"][:] = synth.proposal(w, 7)[0]; gd["norm"][:] = nm
    gsh.Reweight(); l1 = gsh.GetLikelihood(); h1 = gsh.GetMCArray()
    gsh.Reweight(); l2 = gsh.GetLikelihood(); h2 = gsh.GetMCArray()
    np.testing.assert_allclose(h1, h2, rtol=1e-12)
    assert l1 == pytest.approx(l2, rel=1e-10) and l1 > 0
    # oracle on a 1/16 subset of the same events: histogram of the subset must match
    sub = w.scaled(w.n_events // 16)
    mono, osh, od = O.build_from_workload(sub)
    mono.set_params(gd["pars"]); osh.norm_vals[:] = nm
    osh.Reweight()
    gsub, gsd = handlers.build_from_workload(sub)
    gsd["pars"][:] = gd["pars"]; gsd["norm"][:] = nm
    gsub.Reweight(); gsub.GetLikelihood()
    np.testing.assert_allclose(gsub.GetMCArray(), osh.mc, rtol=ORDER_RTOL if _exact() else HIST_RTOL, atol=1e-12)


def test_zero_copy_oscillation_weights_match_the_copied_path(oracle_build):
    """Oscillation weights in pinned host memory are streamed by the kernel's bulk copies (no H2D pass);
    the result must be identical to the cudaMemcpy path, also for the ragged last tile, and the
    weights must stay on the device for later steps that pass no array."""
    if oracle_build != "serial":
        pytest.skip("device-vs-device comparison: one oracle build is enough")
    w = synth.CFG1.scaled(70_001)
    typ, npts, cx = synth.param_layout(w)
    spl, ev = synth.make_splines(w, 0, w.n_events), synth.make_events(w, 0, w.n_events)
    out = []
    for pinned in (False, True):
        h = lib.Handle(update_w2=True, flags=lib.FLAG_KEEP_EVENT_WEIGHTS)
        h.upload_spline_monolith(w.n_params, w.n_knots, cx, npts, spl)
        h.upload_binning(synth.bin_edges(w))
        h.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, None, 0,
                        ev["static_w"])
        res, keep = [], []
        for step in range(3):
            osc = synth.make_osc(w, step)
            keep.append(osc)                    # registered memory must outlive the handle
            if pinned:
                h.register_host_buffer(osc)
            sp, nm = synth.proposal(w, step)
            h.step(sp, nm, osc)
            llh = h.llh()
            res.append((llh, h.read_hist()[0].copy(), h.read_event_weights()[1].copy()))
            sp2, nm2 = synth.proposal(w, step + 10)
            h.step(sp2, nm2, None)              # keeps the weights of the previous call
            res.append((h.llh(), h.read_hist()[0].copy(), h.read_event_weights()[1].copy()))
        out.append(res)
        h.close()
    for (l0, m0, w0), (l1, m1, w1) in zip(*out):
        np.testing.assert_array_equal(w0, w1)
        np.testing.assert_allclose(m0, m1, rtol=ORDER_RTOL, atol=1e-13)
        assert l0 == pytest.approx(l1, rel=1e-10, abs=1e-9)


@pytest.mark.parametrize("wl,n_events,n_sets,kernel", [("CFG5", 40_003, 48, "batch"), ("CFG5", 20_000, 48, "sequential"),
                                                       ("SPARSE_RUNS", 20_000, 70, "batch"), ("CFG1", 9_000, 300, "batch")])
def test_batched_proposals_match_sequential_oracle(oracle_build, wl, n_events, n_sets, kernel, monkeypatch):
    """BASELINE config 5 shape (reduced) and friends: a batch of parameter sets -- perturbations around one point so
    only a few segments per parameter are active -- evaluated by m3b_step_batch in ONE pass over the coefficient
    rows (or, forced, by sequential single-set launches); every set's -lnL against the oracle run sequentially
    (same cached-segment history).  300 sets = two launches of the batched kernel."""
    w = getattr(synth, wl).scaled(n_events)
    mono, osh, od = O.build_from_workload(w)
    gsh, gd = handlers.build_from_workload(w, batch_kernel=(kernel != "sequential"))     # M3B_FLAG_NO_BATCH_KERNEL
    _set(w, -1, mono, osh, gsh, gd)
    osh.Reweight(); gsh.Reweight(); gsh.GetLikelihood()
    data = np.random.default_rng(9).poisson(osh.mc).astype(np.float64)
    osh.AddData(data); gsh.AddData(data)
    rng = np.random.default_rng(10)
    sp0, nm0 = synth.proposal(w, 3)
    sps = np.clip(sp0[None, :] + rng.normal(0, 0.3, (n_sets, w.n_params)), -2.9, 2.9)
    nms = np.clip(nm0[None, :] + rng.normal(0, 0.05, (n_sets, max(w.n_norm_params, 1))), 0.5, 1.5)[:, :w.n_norm_params]
    sps[5] = np.round(sps[5])            # a set sitting exactly on knots
    launches0 = gsh.handle.info().kernel_launches
    tot, per = gsh.handle.step_batch(sps, nms if w.n_norm_params else None, per_sample=True)
    launches = gsh.handle.info().kernel_launches - launches0
    assert launches == (n_sets if kernel == "sequential" else 2 * ((n_sets + 255) // 256))
    assert tot.shape == (n_sets,)
    for i in range(n_sets):
        mono.set_params(sps[i])
        if w.n_norm_params:
            osh.norm_vals[:] = nms[i]
        osh.Reweight()
        o = osh.GetLikelihood()
        assert tot[i] == pytest.approx(o, rel=1e-10 if _exact() else LLH_RTOL, abs=1e-9), i
        assert per[i].sum() == pytest.approx(tot[i], rel=1e-12)
    assert gsh.GetLikelihood() == tot[-1]
    np.testing.assert_allclose(gsh.GetMCArray(), osh.mc, rtol=1e-12 if _exact() else HIST_RTOL, atol=1e-12)
    # the handle keeps working like after a normal step
    _set(w, 4, mono, osh, gsh, gd)
    _check_step(w, mono, osh, gsh, check_weights=False)


def test_tiny_and_ragged_event_counts():
    """Fewer events than one tile row, one event, and a count one past a tile boundary."""
    for n in (1, 33, 513, 1025):
        w = synth.CFG1.scaled(n)
        O.set_multithread(False)
        mono, osh, gsh, gd = _pair(w)
        np.testing.assert_array_equal(gsh.GetEventBins(), osh.event_bins())
        for step in (-1, 0, 1):
            _set(w, step, mono, osh, gsh, gd)
            _check_step(w, mono, osh, gsh)
    O.set_multithread(True)


def test_call_order_errors_are_reported():
    h = lib.Handle()
    with pytest.raises(lib.M3BError) as ei:
        h.step(np.zeros(3))                      # nothing uploaded
    assert ei.value.code == 3                    # M3B_ERR_STATE
    h.upload_binning([[np.array([0.0, 1.0, 2.0])]])
    with pytest.raises(lib.M3BError):
        h.upload_binning([[np.array([0.0, 1.0])]])           # twice
    with pytest.raises(lib.M3BError):
        h.upload_data(np.zeros(5))                           # wrong bin count
    with pytest.raises(lib.M3BError):
        h.upload_events(np.array([3], np.int32), np.zeros(1))   # sample id out of range
    h.close()


def test_llh_scan_matches_the_reference_loop(oracle_build):
    """FitterBase::RunLLHScan through m3b_step_batch against the reference's loop (set one parameter to the bin
    centre, Reweight, GetLikelihood) run on the oracle."""
    from mach3_b200 import fitters
    w = synth.SPARSE.scaled(30_000)
    mono, osh, od = O.build_from_workload(w)
    gsh, gd = handlers.build_from_workload(w)
    _set(w, -1, mono, osh, gsh, gd)
    osh.Reweight(); gsh.Reweight(); gsh.GetLikelihood()
    data = np.random.default_rng(12).poisson(osh.mc).astype(np.float64)
    osh.AddData(data); gsh.AddData(data)
    c_sp, c_nm = synth.proposal(w, 2)
    res = fitters.RunLLHScan(gsh, c_sp, c_nm, spline_ranges={0: (-2.5, 2.5), w.n_params - 1: (-1.0, 1.0)},
                             norm_ranges={1: (0.7, 1.3)}, n_points=40, by_sample=True)
    for (kind, idx), r in res.items():
        assert r["x"][0] < r["x"][-1] and r["llh2"].shape == (40,)
        for j, x in enumerate(r["x"]):
            sp, nm = c_sp.copy(), c_nm.copy()
            (sp if kind == "spline" else nm)[idx] = x
            mono.set_params(sp); osh.norm_vals[:] = nm
            osh.Reweight()
            assert r["llh2"][j] == pytest.approx(2 * osh.GetLikelihood(), rel=1e-9 if _exact() else LLH_RTOL, abs=1e-9)
            for s in range(w.n_samples):
                assert r["llh2_by_sample"][j, s] == pytest.approx(2 * osh.GetSampleLikelihood(s), rel=1e-8 if _exact() else 1e-5, abs=1e-8)
    # the reference resets the scanned parameter without reweighting (FitterBase.cpp:795); RunLLHScan leaves the
    # handle at the central values with one Reweight: same cached-segment history on both sides
    mono.set_params(c_sp); osh.norm_vals[:] = c_nm; osh.Reweight()
    assert gsh.GetLikelihood() == pytest.approx(osh.GetLikelihood(), rel=1e-9 if _exact() else LLH_RTOL)


@pytest.mark.parametrize("update_w2", [False, True])
def test_predictive_toys_match_the_reference_loop(oracle_build, update_w2):
    """PredictiveThrower's toy loop (set a thrown parameter set, Reweight, keep every sample's MC histogram;
    Fitters/PredictiveThrower.cpp:507-563) through m3b_step_batch_hist: frozen W2 goes through the batched kernel
    (one pass over the coefficient rows for all throws), live W2 through the sequential path; both return each
    throw's histogram and -lnL as the oracle's loop computes them."""
    from mach3_b200 import fitters
    w = synth.SPARSE.scaled(20_000)
    mono, osh, od = O.build_from_workload(w, update_w2=update_w2)
    gsh, gd = handlers.build_from_workload(w, update_w2=update_w2)
    _set(w, -1, mono, osh, gsh, gd)
    osh.Reweight(); gsh.Reweight(); gsh.GetLikelihood()
    data = np.random.default_rng(13).poisson(osh.mc).astype(np.float64)
    osh.AddData(data); gsh.AddData(data)
    n_toys = 37
    throws = [synth.proposal(w, 100 + k) for k in range(n_toys)]
    sps = np.stack([t[0] for t in throws]); nms = np.stack([t[1] for t in throws])
    mc, llh = fitters.ProduceToys(gsh, sps, nms)
    assert mc.shape == (n_toys, w.n_bins) and llh.shape == (n_toys,)
    for k in range(n_toys):
        mono.set_params(sps[k]); osh.norm_vals[:] = nms[k]
        osh.Reweight()
        np.testing.assert_allclose(mc[k], osh.mc, rtol=1e-12 if _exact() else 1e-6, atol=1e-12, err_msg=f"toy {k}")
        assert llh[k] == pytest.approx(osh.GetLikelihood(), rel=1e-9 if _exact() else LLH_RTOL, abs=1e-9)


def test_delayed_rejection_stages_match_the_reference_loop(oracle_build):
    """DelayedMR2T2's stage loop (Fitters/DelayedMR2T2.cpp:110-157): stage proposals with a decaying step scale around
    the current point, each Reweight + GetLikelihood on the oracle; the speculative batch gives the same -lnL per stage."""
    from mach3_b200 import fitters
    w = synth.SPARSE.scaled(20_000)
    mono, osh, od = O.build_from_workload(w)
    gsh, gd = handlers.build_from_workload(w)
    _set(w, -1, mono, osh, gsh, gd)
    osh.Reweight(); gsh.Reweight(); gsh.GetLikelihood()
    data = np.random.default_rng(14).poisson(osh.mc).astype(np.float64)
    osh.AddData(data); gsh.AddData(data)
    rng = np.random.default_rng(15)
    cur_sp, cur_nm = synth.proposal(w, 7)
    for step in range(3):
        scale, sps, nms = 1.0, [], []
        for stage in range(4):                                   # max_rejections = 3, decay_rate = 0.5
            sps.append(cur_sp + scale * 0.4 * rng.normal(size=cur_sp.size)); nms.append(cur_nm + scale * 0.05 * rng.normal(size=cur_nm.size))
            scale *= 0.5
        got = fitters.EvaluateDelayedStages(gsh, np.stack(sps), np.stack(nms))
        for stage in range(4):
            mono.set_params(sps[stage]); osh.norm_vals[:] = nms[stage]
            osh.Reweight()
            assert got[stage] == pytest.approx(osh.GetLikelihood(), rel=1e-9 if _exact() else LLH_RTOL, abs=1e-9)
        cur_sp, cur_nm = sps[int(np.argmin(got))], nms[int(np.argmin(got))]


def test_fill_only_then_llh_from_hist_equals_fused(oracle_build):
    """The multi-GPU building blocks on one GPU: m3b_step_fill + m3b_llh_from_hist (what every rank runs around the
    exchange) give the fused step's histogram and -lnL; the histogram device pointer is stable across steps."""
    from mach3_b200 import sharding
    w = synth.SPARSE.scaled(12_000)
    mono, osh, od = O.build_from_workload(w, update_w2=True)
    fused, fd = handlers.build_from_workload(w, update_w2=True)
    split, sd = handlers.build_from_workload(w, update_w2=True, fused_llh=False)
    sh = sharding.ShardedSampleHandler(split.handle, None)
    split.handle.upload_osc(sd["osc"])           # the sharded wrapper takes the weights per call; keep step 0's on the device
    ptr0 = None
    for step in (-1, 0, 1, 2):
        sp, nm = synth.proposal(w, step)
        mono.set_params(sp); osh.norm_vals[:] = nm
        fd["pars"][:] = sp; fd["norm"][:] = nm
        osh.Reweight(); fused.Reweight()
        if step == -1:
            data = np.random.default_rng(4).poisson(osh.mc).astype(np.float64)
            osh.AddData(data); fused.AddData(data); split.AddData(data)
            osh.Reweight(); fused.Reweight()
        sh.Reweight(sp, nm)
        ptr, nb, live = split.handle.hist_device_ptr()
        assert nb == w.n_bins and live == 1 and (ptr0 is None or ptr == ptr0)
        ptr0 = ptr
        assert sh.GetLikelihood() == pytest.approx(fused.GetLikelihood(), rel=1e-12)
        np.testing.assert_allclose(split.GetMCArray(), fused.GetMCArray(), rtol=1e-12, atol=1e-12)
        assert sh.GetLikelihood() == pytest.approx(osh.GetLikelihood(), rel=1e-10 if _exact() else LLH_RTOL, abs=1e-9)


def test_block_trace_reports_every_block():
    w = synth.CFG1.scaled(50_000)
    gsh, gd = handlers.build_from_workload(w)
    gsh.Reweight(); gsh.GetLikelihood()
    gsh.handle.block_trace(read=False)
    gsh.Reweight(); gsh.GetLikelihood()
    tr = gsh.handle.block_trace()
    assert tr.shape == (gsh.handle.info().grid_blocks, 8)
    assert (tr[:, 6] > tr[:, 0]).all() and int(tr[:, 7].sum()) == -(-w.n_events // 256)     # every 256-event unit exactly once


def test_functional_shifts_applied_on_the_host_rebin_on_the_device():
    """Functional (kinematic-shift) parameters stay on the host (SampleHandlerFD.cpp:545-564): the shifted KinVar are
    handed over with m3b_update_kinematics and re-binned with FindGlobalBin semantics; oracle = the same shifted values."""
    O.set_multithread(False)
    w = synth.SPARSE.scaled(15_000)
    typ, npts, cx = synth.param_layout(w)
    spl, ev = synth.make_splines(w), synth.make_events(w)
    h = lib.Handle(update_w2=True, flags=lib.FLAG_KEEP_KINEMATICS | lib.FLAG_KEEP_EVENT_WEIGHTS)
    h.upload_spline_monolith(w.n_params, w.n_knots, cx, npts, spl)
    h.upload_binning(synth.bin_edges(w))
    h.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, None, 0, ev["static_w"])
    osc = synth.make_osc(w, 0)
    h.upload_osc(osc)
    rng = np.random.default_rng(3)
    for step in range(3):
        kin = ev["kin"] * (1.0 + 0.05 * rng.normal(size=ev["kin"].size))      # an energy-scale-like shift per event
        mono = O.SMonolith(w.n_params, w.n_knots, cx, npts, spl)
        osh = O.SampleHandlerFD(w.n_events, synth.bin_edges(w), w.test_statistic, True)
        norm = np.ones(w.n_norm_params)
        osh.set_events(ev["sample_id"], kin, ev["norm_idx"], w.n_norm_per_event, norm, osc, mono, ev["static_w"])
        sp, nm = synth.proposal(w, step)
        mono.set_params(sp); osh.norm_vals[:] = nm
        osh.Reweight()
        h.update_kinematics(kin)
        h.step(sp, nm)
        h.llh()
        np.testing.assert_array_equal(h.read_event_bins(), osh.event_bins())
        np.testing.assert_allclose(h.read_hist()[0], osh.mc, rtol=1e-12, atol=1e-12)
    h.close()
    O.set_multithread(True)


def test_batch_size_grows_on_one_handle(oracle_build):
    """A later m3b_step_batch with more sets than any earlier one on the same handle (DelayedMR2T2 stages, then an LLH
    scan): the per-set -lnL slots are re-allocated, the kernel's staging buffers are not touched (they once were freed
    here and reused: use-after-free)."""
    w = synth.CFG1.scaled(12_000)
    mono, osh, od = O.build_from_workload(w)
    gsh, gd = handlers.build_from_workload(w)
    _set(w, -1, mono, osh, gsh, gd)
    osh.Reweight(); gsh.Reweight(); gsh.GetLikelihood()
    data = np.random.default_rng(3).poisson(osh.mc).astype(np.float64)
    osh.AddData(data); gsh.AddData(data)
    for n_sets in (4, 100, 7, 300):
        props = [synth.proposal(w, 20 + k) for k in range(n_sets)]
        sps = np.stack([p[0] for p in props]); nms = np.stack([p[1] for p in props])
        tot = gsh.handle.step_batch(sps, nms if w.n_norm_params else None)
        for i in range(n_sets):
            mono.set_params(sps[i])
            if w.n_norm_params:
                osh.norm_vals[:] = nms[i]
            osh.Reweight()
            assert tot[i] == pytest.approx(osh.GetLikelihood(), rel=LLH_RTOL, abs=1e-9), (n_sets, i)


def test_barlow_beeston_negative_discriminant_is_an_error_like_the_reference_throw(oracle_build):
    """SampleHandlerBase::GetTestStatLLH throws MaCh3Exception when the Barlow-Beeston discriminant is negative
    (Samples/SampleHandlerBase.cpp:64-67; only reachable with a negative data bin).  The device flags it and m3b_llh /
    m3b_step_batch return M3B_ERR_MATH instead of handing a NaN to the fitter; the handle stays usable."""
    w = synth.CFG1.scaled(6_000)
    gsh, gd = handlers.build_from_workload(w, test_statistic=lib.BARLOW_BEESTON)
    sp, nm = synth.proposal(w, 0)
    gd["pars"][:] = sp; gd["norm"][:] = nm
    gsh.Reweight(); gsh.GetLikelihood()
    mc = gsh.GetMCArray()
    good = np.random.default_rng(4).poisson(mc).astype(np.float64)
    bad = good.copy()
    b = int(np.argmax(mc))
    bad[b] = -50.0 * max(mc[b], 1.0)            # temp^2 + 4*data*f^2 < 0
    gsh.AddData(bad)
    gsh.Reweight()
    with pytest.raises(lib.M3BError) as ei:
        gsh.GetLikelihood()
    assert ei.value.code == 8 and "Barlow-Beeston" in str(ei.value)
    with pytest.raises(lib.M3BError) as ei:
        gsh.handle.step_batch(np.stack([sp, sp]), np.stack([nm, nm]) if w.n_norm_params else None)
    assert ei.value.code == 8
    gsh.AddData(good)
    gsh.Reweight()
    assert np.isfinite(gsh.GetLikelihood())


def test_indexed_oscillation_table(oracle_build):
    """A binned oscillator: events index a small table of oscillation weights (osc_idx), the table changes every step.
    Both hand-over routes: a numpy array (copy engine) and the library's mapped host memory (fetched by a kernel)."""
    w = synth.SPARSE
    typ, npts, cx = synth.param_layout(w)
    spl, ev = synth.make_splines(w, 0, w.n_events), synth.make_events(w, 0, w.n_events)
    rng = np.random.default_rng(23)
    n_osc = 257
    osc_idx = rng.integers(0, n_osc, w.n_events).astype(np.int32)
    tab_o = np.ones(n_osc, np.float32); tab_g = np.ones(n_osc, np.float32)
    mono = O.SMonolith(w.n_params, w.n_knots, cx, npts, spl)
    osh = O.SampleHandlerFD(w.n_events, synth.bin_edges(w), w.test_statistic, False)
    osh.set_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, np.ones(w.n_norm_params), tab_o, mono,
                   ev["static_w"], osc_idx=osc_idx)
    gsh = handlers.SampleHandlerFD(synth.bin_edges(w), w.test_statistic, False, 0, 0, True, True)
    gsh.SetupSplines(w.n_params, w.n_knots, cx, npts, spl)
    pars, norm = np.zeros(w.n_params), np.ones(w.n_norm_params)
    gsh.SetupEvents(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, norm, tab_g, osc_idx, ev["static_w"])
    gsh.SetSplinePointers(pars)
    gd = dict(pars=pars, norm=norm)
    mapped = gsh.handle.alloc_host(n_osc, np.float32)
    for i, step in enumerate((-1, 0, 1, 2, 3, 4)):
        sp, nm = synth.proposal(w, step)
        mono.set_params(sp); osh.norm_vals[:] = nm; gd["pars"][:] = sp; gd["norm"][:] = nm
        tab = rng.uniform(0.05, 1.3, n_osc).astype(np.float32)
        osh.osc_w[:] = tab
        if i % 2 == 0:
            tab_g[:] = tab; gsh._osc_w = tab_g
        else:
            mapped[:] = tab; gsh._osc_w = mapped
        gsh.OscillatorEvaluated()
        _check_step(w, mono, osh, gsh)
    gsh.handle.free_host(mapped)
