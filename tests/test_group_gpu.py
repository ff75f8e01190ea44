"""m3b_group_*: one sample handler over several devices from ONE process and ONE calling thread -- the form in which
the reference's single-process fitters (Fitters/MR2T2.cpp:62-74: Reweight() then GetLikelihood() per sample handler)
reach more than one GPU.  Checked against the CPU oracle run on the whole workload.

On a one-GPU box the group lists device 0 twice (two members with their own streams and shards: same code path, the
"peer" loads stay on the device); with >= 2 GPUs the members sit on distinct devices and the NCCL arm runs too."""
import numpy as np
import pytest

from mach3_b200 import handlers, lib, synth
from oracle import binding as O

pytestmark = pytest.mark.gpu


def _devices(n):
    import torch
    have = torch.cuda.device_count()
    return [i % have for i in range(n)], have


def _oracle(w, update_w2, test_statistic):
    O.set_multithread(False)
    mono, osh, od = O.build_from_workload(w, update_w2=update_w2, test_statistic=test_statistic)
    return mono, osh, od


def _group_whole(w, devices, update_w2, test_statistic, od):
    g = lib.Group(devices, test_statistic=test_statistic, update_w2=update_w2)
    spl = dict(od["spl"]); spl["nKnots_arr"] = np.asarray(spl["nKnots_arr"], np.uint32)
    g.upload_binning(synth.bin_edges(w))
    g.upload_spline_monolith(w.n_params, w.n_knots, od["coeff_x"], od["npts"], spl)
    ev = od["ev"]
    g.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, None, 0, ev["static_w"])
    return g


@pytest.mark.parametrize("n_members,wl,n_events,update_w2", [(2, "SPARSE_RUNS", 30_011, True), (3, "CFG1", 9_000, False),
                                                             (4, "CFG3", 20_000, False), (8, "CFG1", 1_500, False)])
def test_group_step_matches_oracle(n_members, wl, n_events, update_w2):
    w = getattr(synth, wl).scaled(n_events)
    ts = lib.BARLOW_BEESTON if update_w2 else w.test_statistic
    mono, osh, od = _oracle(w, update_w2, ts)
    devices, have = _devices(n_members)
    g = _group_whole(w, devices, update_w2, ts, od)
    g.connect("peer")
    osc = g.alloc_host(w.n_events, np.float32)
    osc[:] = synth.make_osc(w, 0)
    g.upload_osc(osc)
    sp, nm = synth.proposal(w, -1)
    mono.set_params(sp); osh.norm_vals[:] = nm
    osh.Reweight()
    g.step(sp, nm if w.n_norm_params else None); g.llh()          # the first Reweight freezes W2 (UpdateW2 = false)
    data = np.random.default_rng(5).poisson(osh.mc).astype(np.float64)
    osh.AddData(data); g.upload_data(data)
    for step in range(5):
        sp, nm = synth.proposal(w, step)
        mono.set_params(sp); osh.norm_vals[:] = nm
        host_osc = step >= 2                      # from step 2 on the oscillation weights come from host memory every step
        if host_osc:
            osc[:] = synth.make_osc(w, step); osh.osc_w[:] = osc
        osh.Reweight()
        g.step(sp, nm if w.n_norm_params else None, osc if host_osc else None)
        tot, per = g.llh(per_sample=True)
        mc, w2 = g.read_hist()
        np.testing.assert_allclose(mc, osh.mc, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(w2, osh.w2, rtol=1e-12, atol=1e-12)
        assert tot == pytest.approx(osh.GetLikelihood(), rel=1e-10, abs=1e-9)
        np.testing.assert_allclose(per, [osh.GetSampleLikelihood(s) for s in range(w.n_samples)], rtol=1e-10, atol=1e-9)
    g.close()
    O.set_multithread(True)


def test_group_equals_single_handle_and_member_uploads():
    """The per-member upload route (m3b_group_member + the ordinary calls on the member's shard) and the whole-workload
    route give the same histograms as ONE handle holding everything; selection cuts are sliced with the events."""
    w = synth.SPARSE.scaled(25_000)
    mono, osh, od = _oracle(w, False, w.test_statistic)
    devices, _ = _devices(3)
    gw = _group_whole(w, devices, False, w.test_statistic, od)
    gm = lib.Group(devices, test_statistic=w.test_statistic, update_w2=False)
    typ, npts, cx = synth.param_layout(w)
    for i in range(gm.n):
        e0, e1 = gm.shard(w.n_events, i)
        m = gm.member(i)
        m.upload_binning(synth.bin_edges(w))
        m.splines_begin(w.n_params, w.n_knots, cx, npts, e1 - e0)
        m.splines_append(synth.make_splines(w, e0, e1))
        m.splines_end()
        ev = synth.make_events(w, e0, e1)
        m.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, None, 0, ev["static_w"])
    one, od1 = handlers.build_from_workload(w)
    cutv = np.stack([od["ev"]["kin"].reshape(-1, w.n_events)[0], od["ev"]["kin"].reshape(-1, w.n_events)[1]])
    cuts = [(0, 0, 0.2, 2.2), (1, 1, 0.3, 2.9), (2, 0, 0.1, 1.5)]
    osh.SetSelection(cuts, cutv)
    for g in (gw, gm):
        g.connect("peer")
        g.upload_osc(synth.make_osc(w, 0))
        g.upload_selection(cuts, cutv, n_events=w.n_events)
    one.SetSelection(cuts, cutv)
    sp, nm = synth.proposal(w, -1)
    mono.set_params(sp); osh.norm_vals[:] = nm; osh.Reweight()
    od1["pars"][:] = sp; od1["norm"][:] = nm; one.Reweight(); one.GetLikelihood()
    for g in (gw, gm):
        g.step(sp, nm); g.llh()
    data = np.random.default_rng(6).poisson(osh.mc + 0.3).astype(np.float64)
    osh.AddData(data); gw.upload_data(data); gm.upload_data(data); one.AddData(data)
    for step in range(3):
        sp, nm = synth.proposal(w, step)
        mono.set_params(sp); osh.norm_vals[:] = nm; osh.Reweight()
        od1["pars"][:] = sp; od1["norm"][:] = nm; one.Reweight()
        l1 = one.GetLikelihood()
        np.testing.assert_array_equal(one.handle.read_event_selected(), osh.event_selected(), err_msg="single handle: selection mask")
        np.testing.assert_allclose(one.GetMCArray(), osh.mc, rtol=1e-12, atol=1e-12, err_msg="single handle vs oracle")
        assert l1 == pytest.approx(osh.GetLikelihood(), rel=1e-10, abs=1e-9)
        for name, g in (("whole-workload uploads", gw), ("per-member uploads", gm)):
            g.step(sp, nm)
            lg = g.llh()
            mask = np.concatenate([g.member(i).read_event_selected() for i in range(g.n) if g.member(i).n_events])
            np.testing.assert_array_equal(mask, osh.event_selected(), err_msg=f"{name}: selection mask, step {step}")
            np.testing.assert_allclose(g.read_hist()[0], osh.mc, rtol=1e-12, atol=1e-12, err_msg=f"{name}: histogram, step {step}")
            assert lg == pytest.approx(l1, rel=1e-11, abs=1e-10), name
    gw.close(); gm.close()
    O.set_multithread(True)


def test_group_nccl_exchange_on_distinct_devices():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("the NCCL arm needs two distinct devices")
    w = synth.CFG3.scaled(60_000)
    mono, osh, od = _oracle(w, True, lib.BARLOW_BEESTON)
    n = min(4, torch.cuda.device_count())
    res = {}
    for ex in ("peer", "nccl"):
        g = _group_whole(w, list(range(n)), True, lib.BARLOW_BEESTON, od)
        g.connect(ex)
        g.upload_osc(synth.make_osc(w, 0))
        sp, nm = synth.proposal(w, -1)
        mono.set_params(sp); osh.norm_vals[:] = nm
        O.lib().m3o_set_first_time_w2(osh.h, 1)
        osh.Reweight()
        data = np.random.default_rng(7).poisson(osh.mc).astype(np.float64)
        osh.AddData(data); g.upload_data(data)
        out = []
        for step in range(4):
            sp, nm = synth.proposal(w, step)
            mono.set_params(sp); osh.norm_vals[:] = nm; osh.Reweight()
            g.step(sp, nm)
            tot = g.llh()
            assert tot == pytest.approx(osh.GetLikelihood(), rel=1e-10, abs=1e-9), (ex, step)
            out.append(tot)
        res[ex] = out
        g.close()
    np.testing.assert_allclose(res["peer"], res["nccl"], rtol=1e-12)
    O.set_multithread(True)


def test_group_argument_errors():
    devices, have = _devices(2)
    g = lib.Group(devices)
    with pytest.raises(lib.M3BError):
        g.connect("peer")                    # nothing uploaded
    with pytest.raises(lib.M3BError):
        g.step(np.zeros(3), None)            # not connected
    g.close()
    with pytest.raises(lib.M3BError):
        lib.Group([have + 3])                # no such device
