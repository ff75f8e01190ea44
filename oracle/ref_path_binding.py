"""ctypes front-end of oracle/_ref/libm3ref_path.so: the REFERENCE's own host implementation of the likelihood path
-- SMonolith (CPU build: ScanMasterSpline, PrepareForGPU, Evaluate = SplineBase::FindSplineSegment +
CalcSplineWeights + CalcTotalEventWeight) and SampleHandlerBase::GetTestStatLLH -- compiled from
/root/reference/Splines/SplineMonolith.cpp, Splines/SplineBase.cpp and Samples/SampleHandlerBase.cpp by
oracle/ref_host/Makefile (harness: oracle/ref_host/harness_path.cpp).  TEST INFRASTRUCTURE ONLY: it pins the
oracle's restatement (and, through the golden vectors, the CUDA path) to the reference itself."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libm3ref_path.so")          # default build: M3::float_t = double
LIB_PATH_LM = os.path.join(_HERE, "_ref", "libm3ref_path_lm.so")    # -D_LOW_MEMORY_STRUCTS_: M3::float_t = float
LIB_PATH_LM_MT = os.path.join(_HERE, "_ref", "libm3ref_path_lm_mt.so")   # ... with the release flags + MULTITHREAD
LIB_PATH_LM_B200 = os.path.join(_HERE, "_ref", "libm3ref_path_lm_b200.so")   # ... with adapters/SampleHandlerB200.h over the real class
LIB_PATH_LM_CUDA = os.path.join(_HERE, "_ref", "libm3ref_path_lm_cuda.so")   # ... MaCh3_CUDA build, adapters/SMonolithGPU_m3b200.cu
LIB_PATH_B200 = os.path.join(_HERE, "_ref", "libm3ref_path_b200.so")   # default (double) build with the adapter over the real class
_PATHS = {"float_cuda": LIB_PATH_LM_CUDA, "double": LIB_PATH, "float": LIB_PATH_LM, "float_mt": LIB_PATH_LM_MT, "float_b200": LIB_PATH_LM_B200,
          "double_b200": LIB_PATH_B200}
_LIBS = {}


def available():
    return os.path.exists(LIB_PATH) and os.path.exists(LIB_PATH_LM)


def lib(build="double"):
    """build: "double" (the reference's default), "float" (_LOW_MEMORY_STRUCTS_), or "float_mt" (the float build with
    the reference's release flags and MULTITHREAD: the one `bench.py --impl reference` times)."""
    if build not in _LIBS:
        # "float_refcuda_P<n>": the whole incumbent -- MaCh3_CUDA build with the reference's own kernels for n parameters
        path = (os.path.join(_HERE, "_ref", f"libm3ref_path_lm_refcuda_{build.split('_')[-1]}.so")
                if build.startswith("float_refcuda_") else _PATHS[build])
        L = C.CDLL(path)
        assert L.refp_float_t_bytes() == (8 if build.startswith("double") else 4)
        L.refp_mono_create_from_arrays.restype = C.c_void_p
        L.refp_mono_create_from_arrays.argtypes = ([C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 3
                                                   + [C.c_uint64] + [C.c_void_p] * 4)
        L.refp_fd_create.restype = C.c_void_p
        L.refp_fd_create.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int]
        L.refp_fd_destroy.argtypes = [C.c_void_p]
        L.refp_fd_nbins.argtypes = [C.c_void_p]
        L.refp_fd_attach_monolith.argtypes = [C.c_void_p, C.c_void_p]
        L.refp_fd_attach_binned.restype = C.c_void_p
        L.refp_fd_attach_binned.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                            C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.refp_fd_set_events.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                         C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.refp_fd_set_data.argtypes = [C.c_void_p, C.c_void_p]
        L.refp_fd_set_kin.argtypes = [C.c_void_p, C.c_void_p]
        L.refp_fd_set_test_statistic.argtypes = [C.c_void_p, C.c_int]
        L.refp_fd_reweight.argtypes = [C.c_void_p] * 4
        L.refp_fd_llh.restype = C.c_double
        L.refp_fd_llh.argtypes = [C.c_void_p]
        L.refp_fd_sample_llh.restype = C.c_double
        L.refp_fd_sample_llh.argtypes = [C.c_void_p, C.c_int]
        L.refp_fd_read.argtypes = [C.c_void_p] * 3
        L.refp_fd_events.argtypes = [C.c_void_p] * 3
        L.refp_fd_binned_weights.restype = C.c_int64
        L.refp_fd_binned_weights.argtypes = [C.c_void_p, C.c_void_p]
        L.refp_fd_segments.argtypes = [C.c_void_p, C.c_void_p]
        L.refp_fd_move_to_b200.argtypes = [C.c_void_p, C.c_int]
        L.refp_fd_data_changed.argtypes = [C.c_void_p]
        L.refp_fd_sync_host_arrays.argtypes = [C.c_void_p]
        L.refp_mono_create.restype = C.c_void_p
        L.refp_mono_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refp_mono_destroy.argtypes = [C.c_void_p]
        L.refp_mono_sizes.argtypes = [C.c_void_p, C.c_void_p]
        L.refp_mono_array.restype = C.c_int64
        L.refp_mono_array.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.refp_mono_evaluate.argtypes = [C.c_void_p] * 5
        L.refp_test_stat.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4
        L.refp_poisson.argtypes = [C.c_int] + [C.c_void_p] * 3
        L.refp_low_mc_bound.restype = C.c_double
        _LIBS[build] = L
    return _LIBS[build]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


_ARRAYS = (("coeff_x", np.float32), ("coeff_many", np.float32), ("nKnots_arr", np.uint32), ("paramNo_arr", np.int16),
           ("nParamPerEvent", np.uint32), ("nParamPerEvent_tf1", np.uint32), ("paramNo_tf1", np.int16),
           ("coeff_tf1", np.float32), ("n_pts", np.int16), ("x_pts_f64", np.float64))


class RefSMonolith:
    """The reference's SMonolith built from per-event response functions.

    type[P]: 0 TSpline3_red, 1 TF1_red.  npts[n_events, P]: knots (0: the event has no response to the parameter).
    vals[total_knots, 5] = {x, y, b, c, d} per knot in (event, parameter, knot) order; TF1: column 1 = coefficient."""

    def __init__(self, type_, npts, vals, build="double"):
        L = self.L = lib(build)
        self.type = np.ascontiguousarray(type_, np.int32)
        npts = np.ascontiguousarray(npts, np.int32)
        vals = np.ascontiguousarray(vals, np.float64)
        assert vals.shape == (int(npts.sum()), 5)
        self.n_events, self.n_params = npts.shape
        self.h = L.refp_mono_create(self.n_events, self.n_params, _p(self.type), _p(npts), _p(vals))
        if not self.h:
            raise RuntimeError("the reference threw while building the monolith (MaCh3Exception)")
        s = np.zeros(7, np.int64)
        L.refp_mono_sizes(self.h, _p(s))
        (self.NEvents, self.nParams, self.max_knots, self.NSplines_valid, self.NTF1_valid, self.nKnots,
         self.nTF1coeff) = (int(v) for v in s)

    @classmethod
    def from_arrays(cls, n_params, max_knots, coeff_x, n_pts, type_, spl, build="float"):
        """SMonolith over monolith arrays that already have the reference's layout (mach3_b200.synth.make_splines):
        for workloads too large to hand over as one object per response."""
        self = cls.__new__(cls)
        L = self.L = lib(build)
        k = [np.ascontiguousarray(coeff_x, np.float32), np.ascontiguousarray(n_pts, np.int16), np.ascontiguousarray(type_, np.int8),
             np.ascontiguousarray(spl["nParamPerEvent"], np.uint32), np.ascontiguousarray(spl["paramNo_arr"], np.int16),
             np.ascontiguousarray(spl["nKnots_arr"], np.uint64), np.ascontiguousarray(spl["coeff_many"], np.float32),
             np.ascontiguousarray(spl["nParamPerEvent_tf1"], np.uint32), np.ascontiguousarray(spl["paramNo_tf1"], np.int16),
             np.ascontiguousarray(spl["coeff_tf1"], np.float32)]
        n_events = k[3].size // 2
        self.h = L.refp_mono_create_from_arrays(int(n_params), int(max_knots), _p(k[0]), _p(k[1]), _p(k[2]), n_events, _p(k[3]),
                                                _p(k[4]), _p(k[5]), k[6].size // 4, _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]))
        if not self.h:
            raise RuntimeError("the reference threw while building the monolith")
        s = np.zeros(7, np.int64)
        L.refp_mono_sizes(self.h, _p(s))
        (self.NEvents, self.nParams, self.max_knots, self.NSplines_valid, self.NTF1_valid, self.nKnots,
         self.nTF1coeff) = (int(v) for v in s)
        self.n_events, self.n_params = self.NEvents, self.nParams
        return self

    def arrays(self):
        """The monolith arrays exactly as SMonolith::PrepareForGPU left them (the arguments of
        SMonolithGPU::CopyToGPU_SplineMonolith, and of m3b_upload_spline_monolith)."""
        L = self.L
        out = {}
        for which, (name, dt) in enumerate(_ARRAYS):
            n = L.refp_mono_array(self.h, which, None)
            a = np.zeros(n, dt)
            if n:
                L.refp_mono_array(self.h, which, _p(a))
            out[name] = a
        out["n_events"] = self.NEvents
        return out

    def evaluate(self, pars):
        pars = np.ascontiguousarray(pars, np.float64)
        w = np.zeros(self.NEvents, np.float32)
        seg = np.zeros(self.nParams, np.int16)
        val = np.zeros(self.nParams, np.float32)
        if self.L.refp_mono_evaluate(self.h, _p(pars), _p(w), _p(seg), _p(val)):
            raise RuntimeError("the reference threw in SMonolith::Evaluate")
        return w, seg, val

    def close(self):
        if self.h:
            self.L.refp_mono_destroy(self.h)
            self.h = None


class RefSampleHandlerFD:
    """The reference's SampleHandlerFD (+ BinningHandler) filled directly with a binning, events and their pointers;
    Reweight() / GetLikelihood() are the reference's own.  edges: list over samples -- a list over dims of edge arrays
    (uniform) or an array [n_boxes, n_dim, 2] (non-uniform).  build "float" is the only one in which SMonolith can be
    attached (Samples/SampleHandlerFD.cpp:1244-1254)."""

    def __init__(self, edges, test_statistic=0, update_w2=False, build="double"):
        self.L = lib(build)
        self.build = build
        ns = len(edges)
        ndim, uniform, nbins, flat = np.zeros(ns, np.int32), np.ones(ns, np.int32), np.zeros(ns * 4, np.int32), []
        for s_, dims in enumerate(edges):
            if isinstance(dims, np.ndarray) and dims.ndim == 3:
                uniform[s_], ndim[s_], nbins[4 * s_] = 0, dims.shape[1], dims.shape[0]
                flat.append(np.asarray(dims, np.float64).reshape(-1))
                continue
            ndim[s_] = len(dims)
            for d, e in enumerate(dims):
                nbins[4 * s_ + d] = len(e) - 1
                flat.append(np.asarray(e, np.float64))
        flat = np.ascontiguousarray(np.concatenate(flat))
        self.ndim = ndim
        self.h = self.L.refp_fd_create(ns, _p(ndim), _p(uniform), _p(nbins), _p(flat), int(test_statistic), int(update_w2))
        if not self.h:
            raise RuntimeError("the reference rejected this binning (MaCh3Exception)")
        self.n_samples = ns
        self.n_bins = self.L.refp_fd_nbins(self.h)
        self.mono = None

    def attach_monolith(self, mono: RefSMonolith):
        assert self.build.startswith("float") and mono.L is self.L
        self.L.refp_fd_attach_monolith(self.h, mono.h)
        self.mono = mono          # its SMonolith now belongs to the sample handler

    def attach_binned(self, spl):
        """spl: the dict of mach3_b200.synth.binned.make_binned_splines (the reference's monolith arrays)."""
        k = [np.ascontiguousarray(spl["knot_x"], np.float64), np.ascontiguousarray(spl["n_pts"], np.int16),
             np.ascontiguousarray(spl["uniquesplinevec_Monolith"], np.int32), np.ascontiguousarray(spl["coeffindexvec"], np.int32),
             np.ascontiguousarray(spl["uniquecoeffindices"], np.int32), np.ascontiguousarray(spl["manycoeff_arr"], np.float64),
             np.ascontiguousarray(spl["xcoeff_arr"], np.float64)]
        self.n_slots = k[2].size
        self.n_params = int(spl["n_params"])
        self.L.refp_fd_attach_binned(self.h, self.n_params, int(spl["max_knots"]), _p(k[0]), _p(k[1]), k[2].size, _p(k[2]),
                                     _p(k[3]), k[4].size, _p(k[4]), k[6].size, _p(k[5]), _p(k[6]))

    def set_events(self, sample_id, kin, norm_idx=None, n_norm_per_event=0, n_norm_values=0, w_before=None, w_after=None,
                   n_pool=0, binned_n_per_event=None, binned_slot=None):
        """kin: [max_dim, n_events] (the layout of m3b_upload_events); w_before / w_after: [n_events, k] pool indices."""
        sid = np.ascontiguousarray(sample_id, np.int32)
        E = self.n_events = sid.size
        kin = np.asarray(kin, np.float64).reshape(-1, E)
        k4 = np.zeros((E, 4), np.float64)
        k4[:, :kin.shape[0]] = kin.T
        ni = None if norm_idx is None else np.ascontiguousarray(norm_idx, np.int16)
        wb = None if w_before is None else np.ascontiguousarray(np.asarray(w_before, np.int32).reshape(E, -1))
        wa = None if w_after is None else np.ascontiguousarray(np.asarray(w_after, np.int32).reshape(E, -1))
        bs = bi = None
        if binned_n_per_event is not None:
            bs = np.zeros(E + 1, np.int64)
            bs[1:] = np.cumsum(np.asarray(binned_n_per_event, np.int64))
            bi = np.ascontiguousarray(binned_slot, np.int32)
        self._keep = (sid, k4, ni, wb, wa, bs, bi)
        rc = self.L.refp_fd_set_events(self.h, E, _p(sid), _p(k4), int(n_norm_per_event if ni is not None else 0),
                                       None if ni is None else _p(ni), int(n_norm_values),
                                       0 if wb is None else wb.shape[1], None if wb is None else _p(wb),
                                       0 if wa is None else wa.shape[1], None if wa is None else _p(wa), int(n_pool),
                                       None if bs is None else _p(bs), None if bi is None else _p(bi))
        if rc:
            raise RuntimeError("the reference threw while wiring the events")

    def set_kin(self, kin):
        kin = np.asarray(kin, np.float64).reshape(-1, self.n_events)
        k4 = np.zeros((self.n_events, 4), np.float64)
        k4[:, :kin.shape[0]] = kin.T
        self.L.refp_fd_set_kin(self.h, _p(k4))

    def set_selection(self, cuts):
        """StoredSelection: cuts = [(sample, var, lower, upper), ...]; var = column of the event's kinematic row
        (FD::ReturnKinematicParameter in oracle/ref_host/harness_path.cpp)."""
        cs = np.array([c[0] for c in cuts], np.int32); cv = np.array([c[1] for c in cuts], np.int32)
        lo = np.array([c[2] for c in cuts], np.float64); hi = np.array([c[3] for c in cuts], np.float64)
        self.L.refp_fd_set_selection(self.h, len(cuts), _p(cs), _p(cv), _p(lo), _p(hi))

    def set_linear_shifts(self, target, coef):
        """target[n_pars]: kinematic column each functional parameter shifts; coef[n_pars, n_events]: its per-event
        coefficient, NaN = the event is not in the parameter's funcParsGrid list."""
        tg = np.ascontiguousarray(target, np.int32)
        cf = np.ascontiguousarray(np.asarray(coef, np.float64).reshape(tg.size, self.n_events))
        self.n_shift_pars = tg.size
        self.L.refp_fd_set_linear_shifts(self.h, int(tg.size), _p(tg), _p(cf))

    def set_shift_pars(self, vals):
        v = np.ascontiguousarray(vals, np.float64)
        assert v.size == self.n_shift_pars
        self.L.refp_fd_set_shift_pars(self.h, _p(v))

    def kin(self):
        """The live kinematic rows [n_events, 4] (after the last Reweight: shifted)."""
        out = np.zeros((self.n_events, 4), np.float64)
        self.L.refp_fd_get_kin(self.h, _p(out))
        return out

    def selected(self):
        out = np.zeros(self.n_events, np.uint8)
        self.L.refp_fd_selected(self.h, _p(out))
        return out.astype(bool)

    def set_data(self, data):
        d = np.ascontiguousarray(data, np.float64)
        assert d.size == self.n_bins
        self.L.refp_fd_set_data(self.h, _p(d))

    def set_test_statistic(self, kind):
        self.L.refp_fd_set_test_statistic(self.h, int(kind))

    def reweight(self, spline_pars=None, norm=None, pool=None):
        a = [None if v is None else np.ascontiguousarray(v, np.float64) for v in (spline_pars, norm, pool)]
        if self.L.refp_fd_reweight(self.h, *[None if v is None else _p(v) for v in a]):
            raise RuntimeError("the reference threw in SampleHandlerFD::Reweight")

    def llh(self):
        return self.L.refp_fd_llh(self.h)

    def sample_llh(self):
        return np.array([self.L.refp_fd_sample_llh(self.h, s_) for s_ in range(self.n_samples)])

    def hist(self):
        mc, w2 = np.zeros(self.n_bins), np.zeros(self.n_bins)
        self.L.refp_fd_read(self.h, _p(mc), _p(w2))
        return mc, w2

    def events(self):
        w, b = np.zeros(self.n_events), np.zeros(self.n_events, np.int32)
        self.L.refp_fd_events(self.h, _p(w), _p(b))
        return w, b

    # -- only in builds "float_b200" / "double_b200": the object is an m3b200::SampleHandlerB200<FD> over the reference's class
    def move_to_b200(self, device=0):
        """device: one ordinal, or a list of ordinals (the sample is then spread over them through m3b_group_*)."""
        dev = np.ascontiguousarray([device] if np.isscalar(device) else list(device), np.int32)
        if self.L.refp_fd_move_to_b200_ex(self.h, int(dev.size), _p(dev)):
            raise RuntimeError("SampleHandlerB200::MoveToB200 failed")

    def data_changed(self):
        if self.L.refp_fd_data_changed(self.h):
            raise RuntimeError("SampleHandlerB200::DataChanged failed")

    def sync_host_arrays(self):
        if self.L.refp_fd_sync_host_arrays(self.h):
            raise RuntimeError("SampleHandlerB200::SyncHostArrays failed")

    def binned_weights(self):
        out = np.zeros(self.n_slots)
        self.L.refp_fd_binned_weights(self.h, _p(out))
        return out

    def segments(self, n_params):
        out = np.zeros(n_params, np.int16)
        self.L.refp_fd_segments(self.h, _p(out))
        return out

    def close(self):
        if self.h:
            self.L.refp_fd_destroy(self.h)
            self.h = None


def test_stat(kind, data, mc, w2):
    data, mc, w2 = (np.ascontiguousarray(a, np.float64) for a in (data, mc, w2))
    out = np.zeros(data.size)
    thrown = lib().refp_test_stat(int(kind), data.size, _p(data), _p(mc), _p(w2), _p(out))
    return out, thrown


def poisson(data, mc):
    data, mc = (np.ascontiguousarray(a, np.float64) for a in (data, mc))
    out = np.zeros(data.size)
    lib().refp_poisson(data.size, _p(data), _p(mc), _p(out))
    return out


def num_threads(build="float_mt"):
    return lib(build).refp_num_threads()


def available_refcuda(n_params):
    return os.path.exists(os.path.join(_HERE, "_ref", f"libm3ref_path_lm_refcuda_P{n_params}.so"))


def available_cuda():
    return os.path.exists(LIB_PATH_LM_CUDA)


def available_b200(build="float_b200"):
    return os.path.exists(_PATHS[build])


def available_mt():
    return os.path.exists(LIB_PATH_LM_MT)


def low_mc_bound():
    return lib().refp_low_mc_bound()
