"""Synthetic BinnedSplineHandler workload (BASELINE config 4), in the reference's own
array layout (Splines/BinnedSplineHandler.h:110-135):

    weightvec_Monolith[n_slots]        slot index = syst * n_grid + grid_bin   (the 7-D index space
                                       [sample][osc][syst][mode][v1][v2][v3] flattened, one sample/osc channel)
    uniquesplinevec_Monolith[slot]     spline parameter (= syst) of the slot
    coeffindexvec[slot]                first knot of the slot's spline in manycoeff_arr / xcoeff_arr
    uniquecoeffindices[]               the non-flat slots
    manycoeff_arr[n_coeff*4]           AoS {y,b,c,d} per knot (float, _LOW_MEMORY_STRUCTS_ build)
    xcoeff_arr[n_coeff]                knot x per knot per spline

Events point at one slot per systematic that applies to them (SampleHandlerFD::SetSplinePointers,
Samples/SampleHandlerFD.cpp:1196-1242): `spline_index` concatenated, `n_per_event` counts, in syst order.
Workload generation only (numpy); shared by the oracle, the tests and the bench.
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass(frozen=True)
class BinnedWorkload:
    name: str
    seed: int
    n_events: int
    n_systs: int           # spline parameters
    n_grid: int            # spline bins per systematic (Etrue x Erec x mode)
    n_knots: int
    fill: float            # fraction of non-flat slots
    mean_per_event: int    # systematics applying to an event (pointers per event), at most n_systs
    nbins_x: int
    nbins_y: int
    n_norm_params: int = 5
    n_norm_per_event: int = 2
    test_statistic: int = 1   # Barlow-Beeston (live W2)

    @property
    def n_slots(self):
        return self.n_systs * self.n_grid

    @property
    def n_bins(self):
        return self.nbins_x * self.nbins_y

    def scaled(self, n_events=None, n_grid=None):
        return dataclasses.replace(self, n_events=self.n_events if n_events is None else n_events,
                                   n_grid=self.n_grid if n_grid is None else n_grid)


CFG4 = BinnedWorkload("cfg4: 200 binned-spline systematics x 500k spline bins (20% non-flat, K=7), 2M events, "
                      "100x50 bins, Barlow-Beeston", 4004, 2_000_000, 200, 500_000, 7, 0.2, 40, 100, 50)
CFG4_SMALL = BinnedWorkload("cfg4-small: 24 systematics x 700 spline bins, 30k events, 20x8 bins, Barlow-Beeston",
                            4005, 30_011, 24, 700, 6, 0.3, 9, 20, 8)


def _natural_spline_coeffs(x, y):
    """{y,b,c,d} of the natural cubic spline through (x, y[..., K]) for every row, in double; the last
    knot's row is zero (never used: the segment is clamped to K-2, Splines/SplineBase.cpp:97)."""
    K = x.size
    h = np.diff(x)
    A = np.zeros((K, K))
    A[0, 0] = A[-1, -1] = 1.0
    for i in range(1, K - 1):
        A[i, i - 1], A[i, i], A[i, i + 1] = h[i - 1], 2 * (h[i - 1] + h[i]), h[i]
    Ainv = np.linalg.inv(A)
    rhs = np.zeros(y.shape)
    rhs[..., 1:-1] = 6 * ((y[..., 2:] - y[..., 1:-1]) / h[1:] - (y[..., 1:-1] - y[..., :-2]) / h[:-1])
    M = rhs @ Ainv.T
    out = np.zeros(y.shape + (4,))
    out[..., :-1, 0] = y[..., :-1]
    out[..., :-1, 1] = (y[..., 1:] - y[..., :-1]) / h - h * (2 * M[..., :-1] + M[..., 1:]) / 6
    out[..., :-1, 2] = M[..., :-1] / 2
    out[..., :-1, 3] = (M[..., 1:] - M[..., :-1]) / (6 * h)
    return out


def make_binned_splines(w: BinnedWorkload, f64=False):
    rng = np.random.default_rng(w.seed)
    P, G, K = w.n_systs, w.n_grid, w.n_knots
    x = np.linspace(-3.0, 3.0, K)
    ft = np.float64 if f64 else np.float32                         # M3::float_t of the build
    knot_x = np.tile(x.astype(ft), P)                               # SplineInfoArray[p].xPts
    n_pts = np.full(P, K, np.int16)
    active = rng.random(P * G) < w.fill
    uniquecoeffindices = np.nonzero(active)[0].astype(np.int32)
    n_act = uniquecoeffindices.size
    uniquesplinevec = np.repeat(np.arange(P, dtype=np.int32), G)
    coeffindexvec = np.zeros(P * G, np.int32)
    coeffindexvec[uniquecoeffindices] = np.arange(n_act, dtype=np.int32) * K
    many = np.zeros((n_act, K, 4), ft)
    CH = 1 << 20
    for c0 in range(0, n_act, CH):
        n = min(CH, n_act - c0)
        a = rng.uniform(-0.25, 0.25, (n, 1))
        b = rng.uniform(-0.12, 0.03, (n, 1))
        y = 1 + a * x + b * x * x            # dips below zero at the edges for some splines: exercises the clamp
        many[c0:c0 + n] = _natural_spline_coeffs(x, y).astype(ft)
    xcoeff = np.tile(x.astype(ft), n_act)
    return dict(n_params=P, max_knots=K, knot_x=knot_x, n_pts=n_pts, n_slots=P * G, uniquesplinevec_Monolith=uniquesplinevec,
                coeffindexvec=coeffindexvec, uniquecoeffindices=uniquecoeffindices, manycoeff_arr=many.reshape(-1),
                xcoeff_arr=xcoeff)


def make_binned_events(w: BinnedWorkload, f64=False):
    rng = np.random.default_rng(w.seed + 1)
    E, P, G = w.n_events, w.n_systs, w.n_grid
    grid_bin = rng.integers(0, G, E)
    lo, hi = max(1, w.mean_per_event // 2), min(P, w.mean_per_event * 3 // 2)
    n_per = rng.integers(lo, hi + 1, E).astype(np.uint32)
    # a random subset of n_per systematics per event, ascending (syst order)
    keys = rng.random((E, P)) if E * P <= 50_000_000 else None
    idx = []
    if keys is not None:
        order = np.argsort(keys, axis=1)
        for e in range(E):
            idx.append(np.sort(order[e, :n_per[e]]) * G + grid_bin[e])
        spline_index = np.concatenate(idx).astype(np.int32) if idx else np.zeros(0, np.int32)
    else:  # large workloads: a contiguous run of systematics starting at a random one (cheap to generate)
        start = rng.integers(0, P, E)
        npi = n_per.astype(np.int64)
        tot = int(npi.sum())
        ev = np.repeat(np.arange(E, dtype=np.int64), npi)
        j = np.arange(tot, dtype=np.int64) - np.repeat(np.cumsum(npi) - npi, npi)
        key = np.sort(ev * P + (start[ev] + j) % P)              # ascending systematic inside each event
        spline_index = ((key % P) * G + grid_bin[key // P]).astype(np.int32)
    kin = np.empty((2, E))
    kin[0] = rng.gamma(3.0, 0.3, E)
    kin[1] = rng.uniform(0, np.pi, E)
    norm_idx = rng.integers(0, w.n_norm_params, (E, w.n_norm_per_event)).astype(np.int16)
    static_w = rng.uniform(0.5, 1.5, E).astype(np.float64 if f64 else np.float32)
    return dict(sample_id=np.zeros(E, np.int32), kin=kin.reshape(-1), norm_idx=norm_idx.reshape(-1), static_w=static_w,
                n_per_event=n_per, spline_index=spline_index)


def bin_edges(w: BinnedWorkload):
    return [[np.linspace(0.0, 3.0, w.nbins_x + 1), np.linspace(0.0, np.pi, w.nbins_y + 1)]]


def make_osc(w: BinnedWorkload, step=0, f64=False):
    rng = np.random.default_rng(w.seed + 100 + step)
    o = rng.random(w.n_events).astype(np.float64 if f64 else np.float32)
    o[rng.integers(0, w.n_events, max(1, w.n_events // 997))] = 0.0
    return o


def proposal(w: BinnedWorkload, step: int):
    """step >= 0: N(0,1) clipped to (-2.9, 2.9); -1 nominal (on a knot for odd K), -2 every parameter on a knot,
    -3 below the first knot, -4 above the last."""
    P = w.n_systs
    rng = np.random.default_rng(w.seed + 1000 + max(step, 0))
    if step >= 0:
        sp = np.clip(rng.normal(0, 1, P), -2.9, 2.9)
    elif step == -1:
        sp = np.zeros(P)
    elif step == -2:
        x = np.linspace(-3.0, 3.0, w.n_knots)
        sp = x[rng.integers(1, w.n_knots - 1, P)].astype(np.float32).astype(np.float64)
    elif step == -3:
        sp = np.full(P, -3.5)
    else:
        sp = np.full(P, 3.5)
    nm = np.clip(rng.normal(1, 0.1, w.n_norm_params), 0.5, 1.5) if step >= 0 else np.ones(w.n_norm_params)
    return sp, nm
