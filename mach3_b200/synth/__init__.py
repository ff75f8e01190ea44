"""Deterministic synthetic workloads -- numpy front-end of libm3bsynth.so.

The arrays returned by :func:`make_splines` are exactly what the reference's ``SMonolith`` holds
after ``PrepareForGPU`` (Splines/SplineMonolith.cpp:53-250): ``coeff_x``, AoS ``coeff_many``,
``nKnots_arr`` (first-knot offset), ``paramNo_arr``, ``{count,start}`` pairs, TF1 arrays.
:func:`make_events` returns what ``SampleHandlerFD::Initialise`` wires per event
(Samples/SampleHandlerFD.cpp:169-202).  Host only; used by tests, bench and smoke to build inputs
for BOTH the oracle and the B200 library.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libm3bsynth.so")
MAX_SAMPLES = 16


class _Cfg(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_events", C.c_int64), ("n_params", C.c_int32), ("n_linear", C.c_int32),
        ("n_knots", C.c_int32), ("n_modes", C.c_int32), ("density", C.c_float), ("n_samples", C.c_int32),
        ("n_dims", C.c_int32), ("nbins_x", C.c_int32), ("nbins_y", C.c_int32), ("n_norm_params", C.c_int32),
        ("n_norm_per_event", C.c_int32), ("sample_start", C.c_int64 * (MAX_SAMPLES + 1)),
        ("mode_block", C.c_int32), ("reserved", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(_LIB_PATH)
        for name in ("m3s_param_layout", "m3s_count", "m3s_fill_splines", "m3s_fill_events", "m3s_fill_osc",
                     "m3s_bin_edges", "m3s_proposal"):
            getattr(_lib, name).restype = None
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class Workload:
    """Shape of one synthetic configuration (BASELINE.json ``configs``)."""
    name: str
    seed: int
    n_events: int
    n_cubic: int
    n_linear: int
    n_knots: int
    sample_fracs: tuple = (1.0,)
    n_dims: int = 2
    nbins_x: int = 60
    nbins_y: int = 15
    n_modes: int = 1
    density: float = 1.0
    n_norm_params: int = 5
    n_norm_per_event: int = 2
    test_statistic: int = 0     # kPoisson
    mode_block: int = 1         # events come in runs of this many sharing one interaction mode

    @property
    def n_params(self):
        return self.n_cubic + self.n_linear

    @property
    def n_samples(self):
        return len(self.sample_fracs)

    @property
    def bins_per_sample(self):
        return self.nbins_x * (self.nbins_y if self.n_dims > 1 else 1)

    @property
    def n_bins(self):
        return self.bins_per_sample * self.n_samples

    @property
    def bytes_per_event(self):
        """Algorithmic bytes per event per step: active {y,b,c,d} per cubic,
        {a,b} per linear, one osc weight, one bin id.  Dense workloads only."""
        return 16 * self.n_cubic + 8 * self.n_linear + 4 + 4

    def scaled(self, n_events, name=None):
        import dataclasses
        return dataclasses.replace(self, n_events=int(n_events), name=name or f"{self.name}[E={n_events}]")

    def cfg(self) -> _Cfg:
        c = _Cfg()
        c.seed, c.n_events = self.seed, self.n_events
        c.n_params, c.n_linear, c.n_knots = self.n_params, self.n_linear, self.n_knots
        c.n_modes, c.density = self.n_modes, self.density
        c.n_samples, c.n_dims = self.n_samples, self.n_dims
        c.nbins_x, c.nbins_y = self.nbins_x, self.nbins_y
        c.n_norm_params, c.n_norm_per_event = self.n_norm_params, self.n_norm_per_event
        c.mode_block = self.mode_block
        acc = 0.0
        for s in range(self.n_samples):
            c.sample_start[s] = int(round(acc * self.n_events))
            acc += self.sample_fracs[s]
        for s in range(self.n_samples, MAX_SAMPLES + 1):
            c.sample_start[s] = self.n_events
        return c


# BASELINE.json configs
CFG1 = Workload("cfg1: 100k ev x (10 TSpline3 K=5 + 2 TF1), 1D 50 bins, Poisson", 1001, 100_000, 10, 2, 5,
                n_dims=1, nbins_x=50, nbins_y=1)
CFG2 = Workload("cfg2: T2K-FD-like 1M ev x (40 TSpline3 K=7 + 10 TF1), 60x15 bins, Poisson", 2002, 1_000_000,
                40, 10, 7, n_dims=2, nbins_x=60, nbins_y=15)
CFG3 = Workload("cfg3: DUNE-FD-scale 20M ev x (48 TSpline3 K=7 + 12 TF1), 4 samples x 80x20 bins, Poisson",
                3003, 20_000_000, 48, 12, 7, sample_fracs=(0.4, 0.3, 0.2, 0.1), n_dims=2, nbins_x=80, nbins_y=20)
CFG5 = Workload("cfg5: 5M ev x (48 TSpline3 K=7 + 12 TF1), 900 bins, 256 proposals", 5005, 5_000_000, 48, 12, 7,
                n_dims=2, nbins_x=60, nbins_y=15)
# a small sparse case with interaction-mode structure, several samples and Barlow-Beeston
SPARSE = Workload("sparse: 40k ev, 24 params (20+4), 6 modes, density 0.6, 3 samples", 777, 40_000, 20, 4, 6,
                  sample_fracs=(0.5, 0.3, 0.2), n_dims=2, nbins_x=20, nbins_y=8, n_modes=6, density=0.6,
                  test_statistic=1)


# the same, with events grouped in runs of one mode (as after sorting by interaction mode):
# tiles then carry different parameter signatures, and runs do not align with tile boundaries
SPARSE_RUNS = Workload("sparse-runs: 40k ev, 24 params, 6 modes in runs of 700, density 0.5, 3 samples", 778, 40_000,
                       20, 4, 6, sample_fracs=(0.5, 0.3, 0.2), n_dims=2, nbins_x=20, nbins_y=8, n_modes=6,
                       density=0.5, test_statistic=1, mode_block=700)


def param_layout(w: Workload):
    """-> type[P] int8 (0 TSpline3 / 1 TF1), n_pts[P] int16, coeff_x[P*K] float32."""
    P, K = w.n_params, w.n_knots
    typ = np.zeros(P, np.int8)
    npts = np.zeros(P, np.int16)
    cx = np.zeros(P * K, np.float32)
    cfg = w.cfg()
    lib().m3s_param_layout(C.byref(cfg), _p(typ), _p(npts), _p(cx))
    return typ, npts, cx


def count_responses(w: Workload, e0=0, e1=None):
    """-> (TSpline3 responses, TF1 responses) of events [e0,e1)."""
    e1 = w.n_events if e1 is None else e1
    cfg = w.cfg()
    tc, tl = C.c_uint64(0), C.c_uint64(0)
    lib().m3s_count(C.byref(cfg), C.c_int64(e0), C.c_int64(e1), None, None, C.byref(tc), C.byref(tl))
    return tc.value, tl.value


def make_splines(w: Workload, e0=0, e1=None, into=None):
    """Reference monolith arrays for events [e0,e1) (offsets relative to the chunk).  `into`: a dict of
    preallocated arrays of the same names (e.g. pinned staging buffers), at least as long as needed; the
    result then holds views of their leading parts."""
    e1 = w.n_events if e1 is None else e1
    n = e1 - e0
    cfg = w.cfg()
    tc, tl = count_responses(w, e0, e1)
    sizes = dict(nParamPerEvent=2 * n, paramNo_arr=tc, nKnots_arr=tc, coeff_many=tc * w.n_knots * 4,
                 nParamPerEvent_tf1=2 * n, paramNo_tf1=tl, coeff_tf1=tl * 2)
    if into is not None:
        out = {k: into[k][:sz] for k, sz in sizes.items()}
        assert all(out[k].size == sz for k, sz in sizes.items()), "staging buffers too small"
    else:
        out = dict(
            nParamPerEvent=np.zeros(2 * n, np.uint32), paramNo_arr=np.zeros(tc, np.int16),
            nKnots_arr=np.zeros(tc, np.uint64), coeff_many=np.zeros(tc * w.n_knots * 4, np.float32),
            nParamPerEvent_tf1=np.zeros(2 * n, np.uint32), paramNo_tf1=np.zeros(tl, np.int16),
            coeff_tf1=np.zeros(tl * 2, np.float32))
    lib().m3s_fill_splines(C.byref(cfg), C.c_int64(e0), C.c_int64(e1), _p(out["nParamPerEvent"]),
                           _p(out["paramNo_arr"]), _p(out["nKnots_arr"]), _p(out["coeff_many"]),
                           _p(out["nParamPerEvent_tf1"]), _p(out["paramNo_tf1"]), _p(out["coeff_tf1"]))
    out["n_events"] = n
    return out


def make_events(w: Workload, e0=0, e1=None):
    """sample_id[n] i32, kin[n_dims*n] f64 (dim-major), norm_idx[n*npe] i16, static_w[n] f32."""
    e1 = w.n_events if e1 is None else e1
    n = e1 - e0
    cfg = w.cfg()
    out = dict(sample_id=np.zeros(n, np.int32), kin=np.zeros(w.n_dims * n, np.float64),
               norm_idx=np.zeros(n * w.n_norm_per_event, np.int16), static_w=np.zeros(n, np.float32))
    lib().m3s_fill_events(C.byref(cfg), C.c_int64(e0), C.c_int64(e1), _p(out["sample_id"]), _p(out["kin"]),
                          _p(out["norm_idx"]), _p(out["static_w"]))
    return out


def make_osc(w: Workload, step=0, e0=0, e1=None, out=None):
    e1 = w.n_events if e1 is None else e1
    if out is None:
        out = np.zeros(e1 - e0, np.float32)
    cfg = w.cfg()
    lib().m3s_fill_osc(C.byref(cfg), C.c_int64(e0), C.c_int64(e1), C.c_int64(step), _p(out))
    return out


def bin_edges(w: Workload):
    """-> list over samples of list over dims of float64 edge arrays."""
    cfg = w.cfg()
    res = []
    for s in range(w.n_samples):
        dims = []
        for d in range(w.n_dims):
            nb = w.nbins_x if d == 0 else w.nbins_y
            e = np.zeros(nb + 1, np.float64)
            lib().m3s_bin_edges(C.byref(cfg), C.c_int(s), C.c_int(d), _p(e))
            dims.append(e)
        res.append(dims)
    return res


def proposal(w: Workload, step: int):
    """-> (spline parameter values f64[P], norm parameter values f64[N]).  step<0: special cases."""
    sp = np.zeros(w.n_params, np.float64)
    nm = np.zeros(max(w.n_norm_params, 1), np.float64)
    cfg = w.cfg()
    lib().m3s_proposal(C.byref(cfg), C.c_int64(step), _p(sp), _p(nm))
    return sp, nm[:w.n_norm_params]
