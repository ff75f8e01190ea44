"""BASELINE config 4 shape: BinnedSplineHandler evaluation + N weight pointers per event + Barlow-Beeston
with live W2, B200 path (C ABI) against the CPU oracle.  Bars: segments and bins bit-exact, binned-spline
weights and per-event weights bit-exact against the serial oracle build (same fmaf nesting, same product
order), -lnL <= 1e-6 relative (north_star)."""
import numpy as np
import pytest

from mach3_b200 import handlers, lib
from mach3_b200.synth import binned as B
from oracle import binding as O

pytestmark = pytest.mark.gpu
LLH_RTOL = 1e-6


@pytest.fixture(autouse=True, params=["serial", "multithread"])
def oracle_build(request):
    O.set_multithread(request.param == "multithread")
    yield request.param
    O.set_multithread(True)


def _set(w, step, bsh, osh, gsh, gd, osc_step=None):
    sp, nm = B.proposal(w, step)
    bsh.set_params(sp); osh.norm_vals[:] = nm
    gd["pars"][:] = sp; gd["norm"][:] = nm
    if osc_step is not None:
        osc = B.make_osc(w, osc_step)
        osh.osc_w[:] = osc; gd["osc"][:] = osc
        gsh.OscillatorEvaluated()


@pytest.mark.parametrize("ts", [lib.BARLOW_BEESTON, lib.POISSON])
def test_binned_spline_path_matches_oracle(oracle_build, ts):
    w = B.CFG4_SMALL
    exact = oracle_build == "serial"
    bsh, osh, od = O.build_binned_from_workload(w, update_w2=True, test_statistic=ts)
    gsh, gd = handlers.build_binned_from_workload(w, update_w2=True, test_statistic=ts, keep_event_weights=True)
    np.testing.assert_array_equal(gsh.GetEventBins(), osh.event_bins())
    data = None
    for i, step in enumerate((-1, 0, 1, -2, 2, -3, -4, 3)):
        _set(w, step, bsh, osh, gsh, gd, osc_step=i)
        osh.Reweight(); gsh.Reweight()
        if data is None:
            data = np.random.default_rng(3).poisson(osh.mc).astype(np.float64)
            osh.AddData(data); gsh.AddData(data)
            osh.Reweight(); gsh.Reweight()      # W2 is live: same state on both sides
        o, g = osh.GetLikelihood(), gsh.GetLikelihood()
        seg, _ = gsh.SplineHandler.FindSplineSegment()
        np.testing.assert_array_equal(seg, bsh.segments)
        wv = gsh.SplineHandler.weightvec_Monolith
        np.testing.assert_array_equal(wv, bsh.weights)            # same fmaf nesting, same clamp: bit-exact
        assert (wv >= 0).all()
        sw, tw = gsh.GetEventWeight()
        if exact:
            np.testing.assert_array_equal(tw, osh.event_weights())
        else:
            np.testing.assert_allclose(tw, osh.event_weights(), rtol=1e-5, atol=0)
        mc, w2 = gsh.GetMCArray(), gsh.GetW2Array()
        np.testing.assert_allclose(mc, osh.mc, rtol=1e-12 if exact else 1e-6, atol=1e-12)
        np.testing.assert_allclose(w2, osh.w2, rtol=1e-12 if exact else 1e-6, atol=1e-12)
        assert g == pytest.approx(o, rel=1e-10 if exact else LLH_RTOL, abs=1e-9)
    assert (bsh.weights == 0).any(), "the workload should exercise the negative-weight clamp"


def _scrambled(w, mode):
    """The same physics in a layout the upload's walking-order heuristics were not written for: the result must not care."""
    spl, ev = B.make_binned_splines(w), B.make_binned_events(w)
    rng = np.random.default_rng(17)
    npe = ev["n_per_event"].astype(np.int64)
    start = np.concatenate(([0], np.cumsum(npe)))
    si = ev["spline_index"].copy()
    if mode == "pointer_order":            # every event's pointers shuffled (the product follows the given order) and
        keep = np.ones(si.size, bool)      # every seventh event has no pointer at all
        for e in range(w.n_events):
            a, b = start[e], start[e + 1]
            if e % 7 == 3:
                keep[a:b] = False; npe[e] = 0
            else:
                si[a:b] = rng.permutation(si[a:b])
        ev = dict(ev, n_per_event=npe.astype(np.uint32), spline_index=si[keep])
    else:                                  # the slots themselves in random order: no systematic owns a run of slots
        perm = rng.permutation(w.n_slots).astype(np.int32)
        usv = np.empty_like(spl["uniquesplinevec_Monolith"]); usv[perm] = spl["uniquesplinevec_Monolith"]
        civ = np.empty_like(spl["coeffindexvec"]); civ[perm] = spl["coeffindexvec"]
        spl = dict(spl, uniquesplinevec_Monolith=usv, coeffindexvec=civ, uniquecoeffindices=perm[spl["uniquecoeffindices"]])
        ev = dict(ev, spline_index=perm[si])
    return spl, ev


@pytest.mark.parametrize("mode", ["pointer_order", "slot_order"])
def test_binned_walking_order_does_not_depend_on_the_layout(oracle_build, mode):
    if oracle_build != "serial":
        pytest.skip("bit-exact comparison: serial oracle build")
    w = B.CFG4_SMALL
    spl, ev = _scrambled(w, mode)
    bsh, osh, od = O.build_binned_from_workload(w, update_w2=True, spl=spl, ev=ev)
    gsh, gd = handlers.build_binned_from_workload(w, update_w2=True, keep_event_weights=True, spl=spl, ev=ev)
    for i, step in enumerate((-1, 0, 1, 2)):
        _set(w, step, bsh, osh, gsh, gd, osc_step=i)
        osh.Reweight(); gsh.Reweight()
        np.testing.assert_array_equal(gsh.SplineHandler.weightvec_Monolith, bsh.weights)
        sw, tw = gsh.GetEventWeight()
        np.testing.assert_array_equal(tw, osh.event_weights())
        np.testing.assert_allclose(gsh.GetMCArray(), osh.mc, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(gsh.GetW2Array(), osh.w2, rtol=1e-12, atol=1e-12)


def test_binned_rejects_inconsistent_input():
    w = B.CFG4_SMALL
    spl = B.make_binned_splines(w)
    h = lib.Handle()
    bad = dict(spl); bad["uniquecoeffindices"] = np.array([w.n_slots + 5], np.int32)
    with pytest.raises(lib.M3BError):
        h.upload_binned_splines(bad)
    h.close()
    h = lib.Handle()
    h.upload_binned_splines(spl)
    with pytest.raises(lib.M3BError):       # events first
        h.upload_event_binned_splines(np.zeros(4, np.uint32), np.zeros(0, np.int32))
    h.close()


def test_binned_spline_path_default_double_build(oracle_build):
    """The reference's DEFAULT build (M3::float_t = double): double coefficients / weights / osc / static weights,
    std::fma, the parameter read un-narrowed by the evaluation but narrowed to float by FindSplineSegment."""
    w = B.CFG4_SMALL
    exact = oracle_build == "serial"
    bsh, osh, od = O.build_binned_from_workload(w, update_w2=True, f64=True)
    gsh, gd = handlers.build_binned_from_workload(w, update_w2=True, keep_event_weights=True, f64=True)
    np.testing.assert_array_equal(gsh.GetEventBins(), osh.event_bins())
    data = None
    for i, step in enumerate((-1, 0, 1, -2, 2, -3, -4, 3)):
        sp, nm = B.proposal(w, step)
        sp = sp + 1e-9 * (i + 1)            # not representable in float: the eval must use the double value
        bsh.set_params(sp); osh.norm_vals[:] = nm
        gd["pars"][:] = sp; gd["norm"][:] = nm
        osc = B.make_osc(w, i, f64=True)
        osh.osc_w[:] = osc; gd["osc"][:] = osc; gsh.OscillatorEvaluated()
        osh.Reweight(); gsh.Reweight()
        if data is None:
            data = np.random.default_rng(3).poisson(osh.mc).astype(np.float64)
            osh.AddData(data); gsh.AddData(data)
            osh.Reweight(); gsh.Reweight()
        o, g = osh.GetLikelihood(), gsh.GetLikelihood()
        seg, _ = gsh.SplineHandler.FindSplineSegment()
        np.testing.assert_array_equal(seg, bsh.segments)
        wv = gsh.SplineHandler.weightvec_Monolith
        assert wv.dtype == np.float64
        np.testing.assert_array_equal(wv, bsh.weights)            # same fma nesting in double: bit-exact
        sw, tw = gsh.GetEventWeight()
        if exact:
            np.testing.assert_array_equal(tw, osh.event_weights())
        else:
            np.testing.assert_allclose(tw, osh.event_weights(), rtol=1e-13, atol=0)
        np.testing.assert_allclose(gsh.GetMCArray(), osh.mc, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(gsh.GetW2Array(), osh.w2, rtol=1e-12, atol=1e-12)
        assert g == pytest.approx(o, rel=1e-10, abs=1e-9)
    # and the double build really differs from the float build at the 1e-8 level (so the test can tell them apart)
    fsh, fd = handlers.build_binned_from_workload(w, update_w2=True)
    fd["pars"][:] = gd["pars"]; fd["norm"][:] = gd["norm"]
    fsh.Reweight(); fsh.GetLikelihood()
    assert not np.array_equal(fsh.SplineHandler.weightvec_Monolith.astype(np.float64), gsh.SplineHandler.weightvec_Monolith)


def test_full_size_cfg4_properties(oracle_build):
    """BASELINE config 4 at its full size (2 M events, 100 M spline slots, 20 M non-flat): size-independent properties of
    the device path -- the histogram IS the f64 sum of the device's own per-event weights over its own bins, and the
    per-event weight of a random sample of events IS the float product, in the reference's order, of the norm values,
    the oscillation weight, the event's binned-spline weights as read back from the device, and the static weight."""
    if oracle_build != "serial":
        pytest.skip("no oracle involved: run once")
    w = B.CFG4
    gsh, gd = handlers.build_binned_from_workload(w, update_w2=True, keep_event_weights=True)
    ev = gd["ev"]
    start = np.concatenate(([0], np.cumsum(ev["n_per_event"].astype(np.int64))))
    norm_idx = ev["norm_idx"].reshape(w.n_events, w.n_norm_per_event)
    rng = np.random.default_rng(5)
    sample = rng.choice(w.n_events, 3000, replace=False)
    bins = gsh.GetEventBins()
    for step in (0, 1):
        sp, nm = B.proposal(w, step)
        gd["pars"][:] = sp; gd["norm"][:] = nm
        osc = B.make_osc(w, step); gd["osc"][:] = osc; gsh.OscillatorEvaluated()
        gsh.Reweight()
        llh = gsh.GetLikelihood()
        assert np.isfinite(llh)
        sw, tw = gsh.GetEventWeight()
        mc, w2 = gsh.GetMCArray(), gsh.GetW2Array()
        ok = (tw > 0) & (bins >= 0)
        np.testing.assert_allclose(mc, np.bincount(bins[ok], tw[ok].astype(np.float64), minlength=gsh.n_bins), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(w2, np.bincount(bins[ok], (tw[ok] * tw[ok]).astype(np.float64), minlength=gsh.n_bins), rtol=1e-12, atol=1e-12)
        wv = gsh.SplineHandler.weightvec_Monolith            # 1.0 for flat slots
        assert (wv >= 0).all()
        normf = nm.astype(np.float32)
        for e in sample:
            x = np.float32(1)
            for j in range(w.n_norm_per_event):
                x = np.float32(x * normf[norm_idx[e, j]])
            x = np.float32(x * osc[e])
            s = np.float32(1)
            for k in ev["spline_index"][start[e]:start[e + 1]]:
                x = np.float32(x * wv[k]); s = np.float32(s * wv[k])
            x = np.float32(x * ev["static_w"][e])
            assert tw[e] == x and sw[e] == s, (e, tw[e], x)
