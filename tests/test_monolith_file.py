"""The ROOT-free spline-monolith cache file (reference: SMonolith::PrepareSplineFile / LoadSplineFile,
Splines/SplineMonolith.cpp:454-614): m3b_write_monolith_file -> m3b_upload_from_file must give the device exactly what
m3b_upload_spline_monolith gives it from the in-memory arrays."""
import os

import numpy as np
import pytest

from mach3_b200 import lib, synth


def _arrays(w):
    typ, npts, cx = synth.param_layout(w)
    spl = synth.make_splines(w)
    spl["nKnots_arr"] = spl["nKnots_arr"].astype(np.uint32)
    return typ, npts, cx, spl


@pytest.mark.parametrize("wl,n", [("SPARSE_RUNS", 7_013), ("CFG1", 3_000)])
def test_file_round_trip_on_the_host(tmp_path, wl, n):
    """Writer and header reader need no GPU: sizes and every section come back bit for bit (read with numpy from the
    documented layout: 64-byte header, 48-byte section records)."""
    w = getattr(synth, wl).scaled(n)
    typ, npts, cx, spl = _arrays(w)
    path = str(tmp_path / "mono.m3b")
    xp = np.arange(w.n_params * w.n_knots, dtype=np.float64)
    lib.write_monolith_file(path, w.n_params, w.n_knots, cx, npts, spl, x_pts_f64=xp)
    info = lib.monolith_file_info(path)
    assert info == dict(n_events=n, n_params=w.n_params, max_knots=w.n_knots, total_knots=spl["coeff_many"].size // 4)
    raw = open(path, "rb").read()
    assert raw[:8] == b"M3BMONO1" and len(raw) % 64 == 0
    n_sec = int(np.frombuffer(raw, np.uint32, 1, 12)[0])
    want = {"coeff_x": cx, "coeff_many": spl["coeff_many"], "nKnots_arr": spl["nKnots_arr"], "paramNo_arr": spl["paramNo_arr"],
            "cpu_nParamPerEvent": spl["nParamPerEvent"], "cpu_nParamPerEvent_tf1": spl["nParamPerEvent_tf1"],
            "cpu_coeff_TF1_many": spl["coeff_tf1"], "cpu_paramNo_TF1_arr": spl["paramNo_tf1"], "nPts": npts, "xPts_f64": xp}
    seen = set()
    for i in range(n_sec):
        rec = raw[64 + 48 * i: 64 + 48 * (i + 1)]
        name = rec[:24].split(b"\0")[0].decode()
        eb = int(np.frombuffer(rec, np.uint32, 1, 24)[0])
        count, off = (int(v) for v in np.frombuffer(rec, np.uint64, 2, 32))
        assert off % 64 == 0
        a = want[name]
        assert eb == a.dtype.itemsize and count == a.size, name
        np.testing.assert_array_equal(np.frombuffer(raw, a.dtype, count, off), a.reshape(-1), err_msg=name)
        seen.add(name)
    assert seen == set(want)
    with pytest.raises(lib.M3BError):
        lib.monolith_file_info(str(tmp_path / "nope.m3b"))
    open(tmp_path / "junk.m3b", "wb").write(b"not a monolith" * 10)
    with pytest.raises(lib.M3BError):
        lib.monolith_file_info(str(tmp_path / "junk.m3b"))


@pytest.mark.gpu
@pytest.mark.parametrize("wl,n,chunk", [("SPARSE_RUNS", 30_011, 4096), ("CFG2", 20_000, 0), ("CFG1", 5_000, 1024)])
def test_upload_from_file_equals_upload_from_memory(tmp_path, wl, n, chunk):
    w = getattr(synth, wl).scaled(n)
    typ, npts, cx, spl = _arrays(w)
    path = str(tmp_path / "mono.m3b")
    lib.write_monolith_file(path, w.n_params, w.n_knots, cx, npts, spl)
    ev = synth.make_events(w)
    out = []
    for from_file in (False, True):
        h = lib.Handle(test_statistic=w.test_statistic, flags=lib.FLAG_KEEP_EVENT_WEIGHTS)
        if from_file:
            h.upload_from_file(path, chunk)
        else:
            h.upload_spline_monolith(w.n_params, w.n_knots, cx, npts, spl)
        h.upload_binning(synth.bin_edges(w))
        h.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, None, 0, ev["static_w"])
        h.upload_osc(synth.make_osc(w, 0))
        res = []
        for step in (-1, 0, 1, 2):
            sp, nm = synth.proposal(w, step)
            h.step(sp, nm)
            res.append((h.llh(), h.read_event_weights()[0].copy(), h.read_hist()[0].copy()))
        out.append(res)
        h.close()
    for (l0, w0, m0), (l1, w1, m1) in zip(*out):
        np.testing.assert_array_equal(w0.view(np.uint32), w1.view(np.uint32))
        np.testing.assert_allclose(m0, m1, rtol=1e-12, atol=1e-13)
        assert l0 == pytest.approx(l1, rel=1e-11, abs=1e-10)


@pytest.mark.gpu
def test_group_upload_from_file(tmp_path):
    import torch
    w = synth.SPARSE.scaled(20_000)
    typ, npts, cx, spl = _arrays(w)
    path = str(tmp_path / "mono.m3b")
    lib.write_monolith_file(path, w.n_params, w.n_knots, cx, npts, spl)
    ev = synth.make_events(w)
    one = lib.Handle(test_statistic=w.test_statistic)
    one.upload_spline_monolith(w.n_params, w.n_knots, cx, npts, spl)
    g = lib.Group([i % torch.cuda.device_count() for i in range(3)], test_statistic=w.test_statistic)
    g.upload_binning(synth.bin_edges(w))
    g.upload_from_file(path, 2048)
    for x in (one, g):
        if x is one:
            x.upload_binning(synth.bin_edges(w))
        x.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, None, 0, ev["static_w"])
    g.connect("peer")
    one.upload_osc(synth.make_osc(w, 0)); g.upload_osc(synth.make_osc(w, 0))
    for step in (-1, 0, 1):
        sp, nm = synth.proposal(w, step)
        one.step(sp, nm); g.step(sp, nm)
        assert g.llh() == pytest.approx(one.llh(), rel=1e-11, abs=1e-10)
        np.testing.assert_allclose(g.read_hist()[0], one.read_hist()[0], rtol=1e-12, atol=1e-13)
    g.close(); one.close()
