"""ctypes front-end of the CPU oracle (oracle/m3_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this.  Class and method names follow the reference (SMonolith, SampleHandlerFD, Evaluate,
Reweight, GetLikelihood ...).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libm3oracle.so")
MAX_DIM = 4
kPoisson, kBarlowBeeston, kIceCube, kPearson, kDembinskiAbdelmotteleb = range(5)

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} missing: run make -C oracle (or __graft_entry__.build())")
        L = C.CDLL(_LIB_PATH)
        L.m3o_monolith_create.restype = C.c_void_p
        L.m3o_sample_create.restype = C.c_void_p
        L.m3o_sample_create_ex.restype = C.c_void_p
        L.m3o_binned_create.restype = C.c_void_p
        L.m3o_binnedd_create.restype = C.c_void_p
        L.m3o_binnedd_weights.restype = C.c_void_p
        L.m3o_binnedd_segments.restype = C.c_void_p
        L.m3o_binned_weights.restype = C.c_void_p
        L.m3o_binned_segments.restype = C.c_void_p
        for n in ("m3o_segments", "m3o_param_values", "m3o_total_weights", "m3o_spline_weights", "m3o_tf1_weights",
                  "m3o_mc_array", "m3o_w2_array", "m3o_data_array"):
            getattr(L, n).restype = C.c_void_p
        L.m3o_n_splines_valid.restype = C.c_uint64
        L.m3o_n_tf1_valid.restype = C.c_uint64
        L.m3o_get_likelihood.restype = C.c_double
        L.m3o_get_sample_likelihood.restype = C.c_double
        L.m3o_test_stat_llh.restype = C.c_double
        L.m3o_test_stat_llh.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
        L.m3o_find_bin.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]
        L.m3o_bin_edge.restype = C.c_double
        L.m3o_bin_edge.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.m3o_grid_entry.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.m3o_grid_size.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.m3o_axis_nbins.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _view(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


class SMonolith:
    """Oracle mirror of the reference's CPU ``SMonolith`` (Splines/SplineMonolith.cpp)."""

    def __init__(self, n_params, max_knots, coeff_x, n_pts, spl):
        self._keep = [np.ascontiguousarray(coeff_x, np.float32), np.ascontiguousarray(n_pts, np.int16),
                      np.ascontiguousarray(spl["nParamPerEvent"], np.uint32),
                      np.ascontiguousarray(spl["paramNo_arr"], np.int16),
                      np.ascontiguousarray(spl["nKnots_arr"], np.uint64),
                      np.ascontiguousarray(spl["coeff_many"], np.float32),
                      np.ascontiguousarray(spl["nParamPerEvent_tf1"], np.uint32),
                      np.ascontiguousarray(spl["paramNo_tf1"], np.int16),
                      np.ascontiguousarray(spl["coeff_tf1"], np.float32)]
        k = self._keep
        self.n_events = int(spl["n_events"])
        self.n_params = int(n_params)
        self.h = C.c_void_p(lib().m3o_monolith_create(
            C.c_int(n_params), C.c_int(max_knots), _p(k[0]), _p(k[1]), C.c_uint(self.n_events),
            _p(k[2]), _p(k[3]), _p(k[4]), _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8])))
        self.pars = np.zeros(n_params, np.float64)      # the doubles the splineParsPointers point at
        lib().m3o_set_spline_pointers(self.h, _p(self.pars))

    def set_params(self, values):
        self.pars[:] = values

    def FindSplineSegment(self):
        lib().m3o_find_spline_segment(self.h)

    def CalcSplineWeights(self):
        lib().m3o_calc_spline_weights(self.h)

    def CalcTotalEventWeight(self):
        lib().m3o_calc_total_event_weight(self.h)

    def Evaluate(self):
        lib().m3o_evaluate(self.h)

    def set_knots_f64(self, x_pts):
        """FastSplineInfo::xPts in double (the default build with splines built in-process); None: float knots."""
        self._knots_d = None if x_pts is None else np.ascontiguousarray(x_pts, np.float64)
        lib().m3o_set_knots_f64(self.h, None if x_pts is None else _p(self._knots_d))

    def set_curr_segment(self, p, seg):
        lib().m3o_set_curr_segment(self.h, C.c_int(p), C.c_int(seg))

    @property
    def segments(self):
        return _view(lib().m3o_segments(self.h), self.n_params, np.int16)

    @property
    def param_values(self):
        return _view(lib().m3o_param_values(self.h), self.n_params, np.float32)

    @property
    def total_weights(self):
        return _view(lib().m3o_total_weights(self.h), self.n_events, np.float32)

    @property
    def spline_weights(self):
        return _view(lib().m3o_spline_weights(self.h), lib().m3o_n_splines_valid(self.h), np.float32)

    @property
    def tf1_weights(self):
        return _view(lib().m3o_tf1_weights(self.h), lib().m3o_n_tf1_valid(self.h), np.float32)

    def __del__(self):
        try:
            lib().m3o_monolith_destroy(self.h)
        except Exception:
            pass


class BinnedSplineHandler:
    """Oracle mirror of the reference's ``BinnedSplineHandler`` evaluation (Splines/BinnedSplineHandler.cpp:295-341),
    _LOW_MEMORY_STRUCTS_ (float) build.  `spl` = dict from mach3_b200.synth.binned.make_binned_splines."""

    def __init__(self, spl):
        k = dict(knot_x=np.ascontiguousarray(spl["knot_x"], np.float32), n_pts=np.ascontiguousarray(spl["n_pts"], np.int16),
                 usv=np.ascontiguousarray(spl["uniquesplinevec_Monolith"], np.int32),
                 civ=np.ascontiguousarray(spl["coeffindexvec"], np.int32),
                 uci=np.ascontiguousarray(spl["uniquecoeffindices"], np.int32),
                 many=np.ascontiguousarray(spl["manycoeff_arr"], np.float32), xc=np.ascontiguousarray(spl["xcoeff_arr"], np.float32))
        self._keep = k
        self.n_params = int(spl["n_params"])
        self.n_slots = int(spl["n_slots"])
        self.h = C.c_void_p(lib().m3o_binned_create(C.c_int(self.n_params), C.c_int(int(spl["max_knots"])), _p(k["knot_x"]),
                                                    _p(k["n_pts"]), C.c_int64(self.n_slots), _p(k["usv"]), _p(k["civ"]),
                                                    C.c_int64(k["uci"].size), _p(k["uci"]), _p(k["many"]), _p(k["xc"])))
        self.pars = np.zeros(self.n_params, np.float64)
        lib().m3o_binned_set_pointers(self.h, _p(self.pars))

    def set_params(self, values):
        self.pars[:] = values

    def Evaluate(self):
        lib().m3o_binned_evaluate(self.h)

    @property
    def weights(self):
        return _view(lib().m3o_binned_weights(self.h), self.n_slots, np.float32)

    @property
    def segments(self):
        return _view(lib().m3o_binned_segments(self.h), self.n_params, np.int16)

    def __del__(self):
        try:
            lib().m3o_binned_destroy(self.h)
        except Exception:
            pass


class BinnedSplineHandlerD:
    """The same in the reference's DEFAULT build (M3::float_t = double)."""

    def __init__(self, spl):
        k = dict(knot_x=np.ascontiguousarray(spl["knot_x"], np.float64), n_pts=np.ascontiguousarray(spl["n_pts"], np.int16),
                 usv=np.ascontiguousarray(spl["uniquesplinevec_Monolith"], np.int32),
                 civ=np.ascontiguousarray(spl["coeffindexvec"], np.int32),
                 uci=np.ascontiguousarray(spl["uniquecoeffindices"], np.int32),
                 many=np.ascontiguousarray(spl["manycoeff_arr"], np.float64), xc=np.ascontiguousarray(spl["xcoeff_arr"], np.float64))
        self._keep = k
        self.n_params = int(spl["n_params"])
        self.n_slots = int(spl["n_slots"])
        self.h = C.c_void_p(lib().m3o_binnedd_create(C.c_int(self.n_params), C.c_int(int(spl["max_knots"])), _p(k["knot_x"]),
                                                     _p(k["n_pts"]), C.c_int64(self.n_slots), _p(k["usv"]), _p(k["civ"]),
                                                     C.c_int64(k["uci"].size), _p(k["uci"]), _p(k["many"]), _p(k["xc"])))
        self.pars = np.zeros(self.n_params, np.float64)
        lib().m3o_binnedd_set_pointers(self.h, _p(self.pars))

    def set_params(self, values):
        self.pars[:] = values

    def Evaluate(self):
        lib().m3o_binnedd_evaluate(self.h)

    @property
    def weights(self):
        return _view(lib().m3o_binnedd_weights(self.h), self.n_slots, np.float64)

    @property
    def segments(self):
        return _view(lib().m3o_binnedd_segments(self.h), self.n_params, np.int16)

    def __del__(self):
        try:
            lib().m3o_binnedd_destroy(self.h)
        except Exception:
            pass


class SampleHandlerFD:
    """Oracle mirror of the reference's ``SampleHandlerFD`` reweight / likelihood path."""

    def __init__(self, n_events, edges, test_statistic=kPoisson, update_w2=False):
        """edges: list over samples; uniform sample = list over dims of edge arrays, non-uniform sample = float
        array [n_boxes, n_dim, 2] (BinInfo::Extent)."""
        n_samples = len(edges)
        ndim, uniform, nbins, flat = np.zeros(n_samples, np.int32), np.ones(n_samples, np.int32), np.zeros(n_samples * MAX_DIM, np.int32), []
        for s, dims in enumerate(edges):
            if isinstance(dims, np.ndarray) and dims.ndim == 3:
                uniform[s], ndim[s], nbins[s * MAX_DIM] = 0, dims.shape[1], dims.shape[0]
                flat.append(np.asarray(dims, np.float64).reshape(-1))
                continue
            ndim[s] = len(dims)
            for d, e in enumerate(dims):
                nbins[s * MAX_DIM + d] = len(e) - 1
                flat.append(np.asarray(e, np.float64))
        flat = np.concatenate(flat)
        self.n_events = int(n_events)
        self.n_samples = n_samples
        self.h = C.c_void_p(lib().m3o_sample_create_ex(C.c_uint(n_events), C.c_int(n_samples), _p(ndim), _p(uniform), _p(nbins),
                                                       _p(flat), C.c_int(test_statistic), C.c_int(int(update_w2))))
        self.n_bins = lib().m3o_total_bins(self.h)
        self._keep = []

    def set_events(self, sample_id, kin, norm_idx, n_norm_per_event, norm_vals, osc_w, spline: SMonolith | None,
                   static_w, osc_idx=None):
        """norm_vals / osc_w are the live arrays the per-event pointers will point INTO."""
        self.norm_vals = None if norm_vals is None else np.ascontiguousarray(norm_vals, np.float64)
        self.osc_w = None if osc_w is None else np.ascontiguousarray(osc_w, np.float32)
        k = [np.ascontiguousarray(sample_id, np.int32), np.ascontiguousarray(kin, np.float64),
             None if norm_idx is None else np.ascontiguousarray(norm_idx, np.int16),
             None if static_w is None else np.ascontiguousarray(static_w, np.float32),
             None if osc_idx is None else np.ascontiguousarray(osc_idx, np.int32), spline]
        self._keep = k
        lib().m3o_sample_set_events(self.h, _p(k[0]), _p(k[1]), C.c_int(n_norm_per_event if k[2] is not None else 0),
                                    _p(k[2]), _p(self.norm_vals), _p(self.osc_w), _p(k[4]),
                                    spline.h if spline is not None else None, _p(k[3]))

    def set_events_binned(self, sample_id, kin, norm_idx, n_norm_per_event, norm_vals, osc_w, binned: BinnedSplineHandler,
                          n_per_event, spline_index, static_w):
        """Events whose spline weights are pointers into a BinnedSplineHandler's weightvec_Monolith."""
        self.norm_vals = None if norm_vals is None else np.ascontiguousarray(norm_vals, np.float64)
        self.osc_w = None if osc_w is None else np.ascontiguousarray(osc_w, np.float32)
        k = [np.ascontiguousarray(sample_id, np.int32), np.ascontiguousarray(kin, np.float64),
             None if norm_idx is None else np.ascontiguousarray(norm_idx, np.int16),
             None if static_w is None else np.ascontiguousarray(static_w, np.float32),
             np.ascontiguousarray(n_per_event, np.uint32), np.ascontiguousarray(spline_index, np.int32), binned]
        self._keep = k
        lib().m3o_sample_set_events_binned(self.h, _p(k[0]), _p(k[1]), C.c_int(n_norm_per_event if k[2] is not None else 0),
                                           _p(k[2]), _p(self.norm_vals), _p(self.osc_w), None, binned.h, _p(k[4]), _p(k[5]),
                                           _p(k[3]))

    def set_events_binned_d(self, sample_id, kin, norm_idx, n_norm_per_event, norm_vals, osc_w, binned: BinnedSplineHandlerD,
                            n_per_event, spline_index, static_w):
        """Default build: osc / static weights are double arrays."""
        self.norm_vals = None if norm_vals is None else np.ascontiguousarray(norm_vals, np.float64)
        self.osc_w = None if osc_w is None else np.ascontiguousarray(osc_w, np.float64)
        k = [np.ascontiguousarray(sample_id, np.int32), np.ascontiguousarray(kin, np.float64),
             None if norm_idx is None else np.ascontiguousarray(norm_idx, np.int16),
             None if static_w is None else np.ascontiguousarray(static_w, np.float64),
             np.ascontiguousarray(n_per_event, np.uint32), np.ascontiguousarray(spline_index, np.int32), binned]
        self._keep = k
        self._double_build = True
        lib().m3o_sample_set_events_binned_d(self.h, _p(k[0]), _p(k[1]), C.c_int(n_norm_per_event if k[2] is not None else 0),
                                             _p(k[2]), _p(self.norm_vals), _p(self.osc_w), binned.h, _p(k[4]), _p(k[5]), _p(k[3]))

    def Reweight(self):
        if getattr(self, "_double_build", False):
            lib().m3o_reweight_d(self.h)
        else:
            lib().m3o_reweight(self.h)

    def FillOnly(self):
        lib().m3o_fill_only(self.h)

    def SetSelection(self, cuts, values):
        """StoredSelection: cuts = [(sample, var, lower, upper), ...]; values[n_vars, n_events] (kept alive, may be
        rewritten in place: shifted cut variables) = ReturnKinematicParameter(var, event)."""
        cs = np.array([c[0] for c in cuts], np.int32); cv = np.array([c[1] for c in cuts], np.int32)
        lo = np.array([c[2] for c in cuts], np.float64); hi = np.array([c[3] for c in cuts], np.float64)
        self.cut_values = np.ascontiguousarray(np.asarray(values, np.float64).reshape(-1, self.n_events))
        assert cv.size == 0 or (cv.min() >= 0 and cv.max() < self.cut_values.shape[0])
        lib().m3o_sample_set_selection(self.h, C.c_int(len(cuts)), _p(cs), _p(cv), _p(lo), _p(hi), _p(self.cut_values))

    def SetLinearShifts(self, n_per_event, shift_par, target, coef, values):
        """funcParsGrid for linear functional parameters (SampleHandlerFD::ApplyShifts): per event, in order, entries
        {shift_par, target, coef}; `values` is the live parameter array the FunctionalShifters' valuePtr look into.
        Targets < (kinematic rows) shift the live kinematic array handed to set_events, the rest the cut-variable table."""
        npe = np.ascontiguousarray(n_per_event, np.uint32); sp = np.ascontiguousarray(shift_par, np.int32)
        tg = np.ascontiguousarray(target, np.int32); cf = np.ascontiguousarray(coef, np.float64)
        self.shift_values = np.ascontiguousarray(values, np.float64)
        kin = self._keep[1]
        n_rows = kin.size // self.n_events
        cv = getattr(self, "cut_values", None)
        self._keep_shift = (npe, sp, tg, cf)
        lib().m3o_sample_set_linear_shifts(self.h, _p(npe), _p(sp), _p(tg), _p(cf), _p(self.shift_values), _p(kin), C.c_int(n_rows),
                                           None if cv is None else _p(cv), C.c_int(0 if cv is None else cv.shape[0]))

    def event_selected(self):
        out = np.zeros(self.n_events, np.uint8)
        lib().m3o_event_selected(self.h, _p(out))
        return out.astype(bool)

    def GetLikelihood(self):
        return lib().m3o_get_likelihood(self.h)

    def GetSampleLikelihood(self, i):
        return lib().m3o_get_sample_likelihood(self.h, C.c_int(i))

    def AddData(self, data):
        d = np.ascontiguousarray(data, np.float64)
        assert d.size == self.n_bins
        lib().m3o_add_data(self.h, _p(d))

    def SetTestStatistic(self, t):
        lib().m3o_set_test_statistic(self.h, C.c_int(t))

    @property
    def mc(self):
        return _view(lib().m3o_mc_array(self.h), self.n_bins, np.float64)

    @property
    def w2(self):
        return _view(lib().m3o_w2_array(self.h), self.n_bins, np.float64)

    @property
    def data(self):
        return _view(lib().m3o_data_array(self.h), self.n_bins, np.float64)

    def event_bins(self):
        out = np.zeros(self.n_events, np.int32)
        lib().m3o_event_bins(self.h, _p(out))
        return out

    def event_weights(self):
        if getattr(self, "_double_build", False):
            out = np.zeros(self.n_events, np.float64)
            lib().m3o_event_weights_d(self.h, _p(out))
            return out
        out = np.zeros(self.n_events, np.float32)
        lib().m3o_event_weights(self.h, _p(out))
        return out

    def find_bin(self, sample, dim, var, nom_bin):
        return lib().m3o_find_bin(self.h, sample, dim, float(var), int(nom_bin))

    def axis_edges(self, sample, dim):
        n = lib().m3o_axis_nbins(self.h, sample, dim)
        return np.array([lib().m3o_bin_edge(self.h, sample, dim, i) for i in range(n + 1)])

    def grid_mapping(self, sample, mega):
        n = lib().m3o_grid_size(self.h, sample, mega)
        return [lib().m3o_grid_entry(self.h, sample, mega, k) for k in range(n)]

    def __del__(self):
        try:
            lib().m3o_sample_destroy(self.h)
        except Exception:
            pass


def test_stat_llh(test_statistic, data, mc, w2):
    return lib().m3o_test_stat_llh(int(test_statistic), float(data), float(mc), float(w2))


def set_multithread(on: bool):
    """True: the reference's MULTITHREAD build (OpenMP, reassociating simd reductions);
    False: its serial build (strict left-to-right float products, FillArray)."""
    lib().m3o_set_multithread(C.c_int(int(on)))


def num_threads():
    return lib().m3o_num_threads()


def build_from_workload(w, e0=0, e1=None, with_osc=True, update_w2=False, test_statistic=None):
    """Convenience: oracle SMonolith + SampleHandlerFD wired on a synthetic workload."""
    from mach3_b200 import synth
    e1 = w.n_events if e1 is None else e1
    typ, npts, cx = synth.param_layout(w)
    spl = synth.make_splines(w, e0, e1)
    ev = synth.make_events(w, e0, e1)
    mono = SMonolith(w.n_params, w.n_knots, cx, npts, spl)
    sh = SampleHandlerFD(e1 - e0, synth.bin_edges(w), w.test_statistic if test_statistic is None else test_statistic,
                         update_w2)
    osc = synth.make_osc(w, 0, e0, e1) if with_osc else None
    norm_vals = np.ones(max(w.n_norm_params, 1), np.float64)
    sh.set_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, norm_vals, osc, mono, ev["static_w"])
    return mono, sh, dict(typ=typ, npts=npts, coeff_x=cx, spl=spl, ev=ev)


def build_binned_from_workload(w, update_w2=True, test_statistic=None, with_osc=True, f64=False, spl=None, ev=None):
    """Oracle BinnedSplineHandler + SampleHandlerFD wired on a synthetic binned-spline workload
    (f64: the reference's default build).  spl / ev: use these arrays instead of generating them."""
    from mach3_b200.synth import binned as B
    spl = B.make_binned_splines(w, f64=f64) if spl is None else spl
    ev = B.make_binned_events(w, f64=f64) if ev is None else ev
    sh = SampleHandlerFD(w.n_events, B.bin_edges(w), w.test_statistic if test_statistic is None else test_statistic, update_w2)
    osc = B.make_osc(w, 0, f64=f64) if with_osc else None
    norm_vals = np.ones(w.n_norm_params, np.float64)
    if f64:
        bsh = BinnedSplineHandlerD(spl)
        sh.set_events_binned_d(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, norm_vals, osc, bsh,
                               ev["n_per_event"], ev["spline_index"], ev["static_w"])
    else:
        bsh = BinnedSplineHandler(spl)
        sh.set_events_binned(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, norm_vals, osc, bsh,
                             ev["n_per_event"], ev["spline_index"], ev["static_w"])
    return bsh, sh, dict(spl=spl, ev=ev)
