"""Non-uniform ("mega-bin") binning, SURVEY §8a row a11 / §8f rank 3: boxes behind a 10-per-dimension grid
(Samples/SampleStructs.h:394-528), first listed box with lo < x <= hi wins (BinningHandler.cpp:278-290).
CPU part: known answers against the oracle.  GPU part: device bin ids bit-exact against the oracle, and the
whole step (weights -> boxes -> -lnL) in parity."""
import numpy as np
import pytest

from mach3_b200 import synth
from oracle import binding as O


def _boxes(rng, nx=7, x_max=3.0, y_max=np.pi):
    """Irregular tiling of [0,x_max] x [0,y_max]: nx columns of random widths, each cut into a random number of rows."""
    xs = np.concatenate([[0.0], np.sort(rng.uniform(0.1, x_max - 0.1, nx - 1)), [x_max]])
    out = []
    for i in range(nx):
        ny = int(rng.integers(1, 6))
        ys = np.concatenate([[0.0], np.sort(rng.uniform(0.05, y_max - 0.05, ny - 1)), [y_max]])
        for j in range(ny):
            out.append([[xs[i], xs[i + 1]], [ys[j], ys[j + 1]]])
    return np.array(out, np.float64)


def _edges(w, rng):
    """sample 0 uniform (the workload's own), the rest non-uniform"""
    ed = synth.bin_edges(w)
    return [ed[0]] + [_boxes(rng) for _ in range(1, len(ed))]


def test_nonuniform_known_answers():
    boxes = np.array([[[0, 1], [0, 1]], [[1, 3], [0, 0.5]], [[1, 3], [0.5, 1]], [[0, 1], [1, 2]], [[1, 3], [1, 2]]], np.float64)
    sh = O.SampleHandlerFD(7, [boxes])
    #            inside box 0   on box0's upper x edge (lo<x<=hi: still box 0)   x=1+eps -> box 1   y on 0.5 edge -> box 1
    kin_x = np.array([0.5, 1.0, 1.0000001, 2.0, 0.0, 3.0, 2.0])
    kin_y = np.array([0.5, 0.5, 0.25, 0.5, 0.5, 1.5, 2.0])
    sh.set_events(np.zeros(7, np.int32), np.concatenate([kin_x, kin_y]), None, 0, None, None, None, None)
    bins = sh.event_bins()
    # x == 0.0 is the lowest lower edge: accepted by the mega grid ([lo,hi)) but in no box ((lo,hi]) -> -1;
    # x == 3.0 / y == 2.0 are the top edges: rejected by the mega grid (x >= last edge) -> -1
    assert bins.tolist() == [0, 0, 1, 1, -1, -1, -1]
    assert sh.n_bins == 5


@pytest.mark.gpu
def test_nonuniform_bins_and_step_parity_on_gpu():
    from mach3_b200 import handlers, lib
    O.set_multithread(False)
    w = synth.SPARSE.scaled(25_003)
    rng = np.random.default_rng(77)
    edges = _edges(w, rng)
    typ, npts, cx = synth.param_layout(w)
    spl, ev = synth.make_splines(w), synth.make_events(w)
    # a few events exactly on box edges
    kin = ev["kin"].reshape(2, -1).copy()
    b1 = edges[1]
    kin[0, :40] = np.resize(b1[:, 0, 1], 40); kin[1, :40] = np.resize(b1[:, 1, 1], 40)
    ev["sample_id"][:40] = 1
    osc = synth.make_osc(w, 0)
    mono = O.SMonolith(w.n_params, w.n_knots, cx, npts, spl)
    osh = O.SampleHandlerFD(w.n_events, edges, lib.BARLOW_BEESTON, True)
    norm = np.ones(w.n_norm_params)
    osh.set_events(ev["sample_id"], kin.reshape(-1), ev["norm_idx"], w.n_norm_per_event, norm, osc, mono, ev["static_w"])
    gsh = handlers.SampleHandlerFD(edges, lib.BARLOW_BEESTON, True, keep_event_weights=True)
    gsh.SetupSplines(w.n_params, w.n_knots, cx, npts, spl)
    pars, gnorm = np.zeros(w.n_params), np.ones(w.n_norm_params)
    gsh.SetupEvents(ev["sample_id"], kin.reshape(-1), ev["norm_idx"], w.n_norm_per_event, gnorm, osc.copy(), None, ev["static_w"])
    gsh.SetSplinePointers(pars)
    assert gsh.n_bins == osh.n_bins
    ob, gb = osh.event_bins(), gsh.GetEventBins()
    np.testing.assert_array_equal(gb, ob)                       # bit-exact bin ids, both arms
    assert (ob[ev["sample_id"] > 0] >= 0).mean() > 0.9
    data = None
    for step in (-1, 0, 1, 2):
        sp, nm = synth.proposal(w, step)
        mono.set_params(sp); osh.norm_vals[:] = nm; pars[:] = sp; gnorm[:] = nm
        osh.Reweight(); gsh.Reweight()
        if data is None:
            data = np.random.default_rng(5).poisson(osh.mc).astype(np.float64)
            osh.AddData(data); gsh.AddData(data)
            osh.Reweight(); gsh.Reweight()          # the fused -lnL is formed inside Reweight: redo it with the data in place
        o, g = osh.GetLikelihood(), gsh.GetLikelihood()
        np.testing.assert_allclose(gsh.GetMCArray(), osh.mc, rtol=1e-12, atol=1e-12)
        assert g == pytest.approx(o, rel=1e-10, abs=1e-9)
        for i in range(w.n_samples):
            assert gsh.GetSampleLikelihood(i) == pytest.approx(osh.GetSampleLikelihood(i), rel=1e-9, abs=1e-9)
    O.set_multithread(True)
