"""ctypes binding of libm3b200.so (include/m3b200.h).  Fails loudly when the CUDA library is
missing or cannot create a device context: there is no CPU fallback in the product path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# M3B_LIB: an experiments build of the library (mach3_b200/build.py with M3B_BUILD_EXPERIMENTS=1: legacy kernel variants and
# M3B_* tuning knobs read from the environment) for A/B measurements; the product library reads no environment
LIB_PATH = os.environ.get("M3B_LIB") or os.path.join(_HERE, "libm3b200.so")

OK = 0
POISSON, BARLOW_BEESTON, ICECUBE, PEARSON, DEMBINSKI_ABDELMOTTELEB = range(5)
FLAG_KEEP_EVENT_WEIGHTS = 1
FLAG_KEEP_KINEMATICS = 2
FLAG_NO_FUSED_LLH = 4
FLAG_NO_BATCH_KERNEL = 8
FLAG_BATCH_KERNEL_V1 = 16
FLAG_NO_SPIN_LLH = 32

#: every symbol include/m3b200.h declares
EXPORTS = (
    "m3b_create", "m3b_destroy", "m3b_last_error", "m3b_abi_version", "m3b_set_stream",
    "m3b_splines_begin", "m3b_splines_append", "m3b_splines_end", "m3b_upload_spline_monolith",
    "m3b_upload_binned_splines", "m3b_upload_event_binned_splines", "m3b_read_binned_weights",
    "m3b_upload_binned_splines_f64", "m3b_upload_event_weights_f64", "m3b_upload_osc_f64", "m3b_read_binned_weights_f64",
    "m3b_read_event_weights_f64",
    "m3b_upload_binning", "m3b_upload_binning_ex", "m3b_upload_events", "m3b_update_kinematics", "m3b_upload_selection", "m3b_update_selection_values", "m3b_read_event_selected",
    "m3b_upload_linear_shifts", "m3b_set_shift_pars", "m3b_upload_data", "m3b_upload_osc", "m3b_register_host_buffer", "m3b_alloc_host", "m3b_free_host",
    "m3b_set_test_statistic", "m3b_set_flags", "m3b_reset_w2",
    "m3b_step", "m3b_step_segments", "m3b_step_batch", "m3b_step_batch_hist", "m3b_llh", "m3b_eval_weights", "m3b_find_segments", "m3b_set_spline_knots_f64", "m3b_synchronize",
    "m3b_read_hist", "m3b_read_event_weights", "m3b_read_event_bins",
    "m3b_step_fill", "m3b_hist_device_ptr", "m3b_llh_from_hist", "m3b_peer_export", "m3b_peer_import", "m3b_step_peer",
    "m3b_get_info", "m3b_set_timing", "m3b_kernel_time", "m3b_block_trace",
    "m3b_group_create", "m3b_group_destroy", "m3b_group_last_error", "m3b_group_size", "m3b_group_member", "m3b_group_shard",
    "m3b_group_upload_spline_monolith", "m3b_group_upload_binning_ex", "m3b_group_upload_events", "m3b_group_upload_selection",
    "m3b_group_upload_data", "m3b_group_upload_osc", "m3b_group_connect", "m3b_group_alloc_host", "m3b_group_step",
    "m3b_group_llh", "m3b_group_read_hist", "m3b_group_synchronize",
    "m3b_write_monolith_file", "m3b_monolith_file_info", "m3b_upload_from_file", "m3b_group_upload_from_file",
    "m3b_group_upload_linear_shifts", "m3b_group_set_shift_pars",
)
EXCHANGE_PEER, EXCHANGE_NCCL = 0, 1


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("test_statistic", C.c_int32), ("update_w2", C.c_int32),
                ("tile_events", C.c_int32), ("flags", C.c_int32), ("reserved", C.c_int32 * 11)]


class Info(C.Structure):
    _fields_ = [("n_events", C.c_int64), ("n_tiles", C.c_int64), ("n_params", C.c_int32), ("n_bins", C.c_int32),
                ("n_samples", C.c_int32), ("n_signatures", C.c_int32), ("tile_events", C.c_int32),
                ("grid_blocks", C.c_int32), ("smem_bytes", C.c_int32), ("hist_in_smem", C.c_int32),
                ("device_bytes", C.c_uint64), ("active_bytes_per_step", C.c_uint64), ("steps", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("kernel_variant", C.c_int32), ("tma_stages", C.c_int32)]


class M3BError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libm3b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libm3b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                               "mach3_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.m3b_last_error.restype = C.c_char_p
        L.m3b_last_error.argtypes = [C.c_void_p]
        L.m3b_destroy.restype = None
        L.m3b_destroy.argtypes = [C.c_void_p]
        L.m3b_group_last_error.restype = C.c_char_p
        L.m3b_group_last_error.argtypes = [C.c_void_p]
        L.m3b_group_destroy.restype = None
        L.m3b_group_destroy.argtypes = [C.c_void_p]
        L.m3b_group_member.restype = C.c_void_p
        L.m3b_group_member.argtypes = [C.c_void_p, C.c_int32]
        L.m3b_group_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.m3b_group_llh.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def bind_to_gpu_cpus(device=0):
    """Pin the calling process to the CPUs NVML reports as local to `device` so that pinned host
    buffers are first-touched on the GPU's socket (zero-copy osc weights cross PCIe once, not UPI +
    PCIe).  Host-side placement only; returns a short description or None when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(device)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or allowed
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} cpus local to gpu {device}"
    except Exception as e:       # noqa: BLE001 - best effort
        return None


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def addr(a):
    """Integer address of a C-contiguous array (for Handle.step_addr: callers that, like a C++ host, already hold
    pointers and must not pay ctypes' per-call pointer extraction).  The array must stay alive."""
    assert a.flags.c_contiguous
    return a.ctypes.data


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype)


def write_monolith_file(path, n_params, max_knots, coeff_x, n_pts, spl, x_pts_f64=None):
    """The ROOT-free spline-monolith cache file (m3b_write_monolith_file); needs no GPU."""
    L = load()
    cx, npt = _c(coeff_x, np.float32), _c(n_pts, np.int16)
    xp = _c(x_pts_f64, np.float64)
    a = [_c(spl["nParamPerEvent"], np.uint32), _c(spl["paramNo_arr"], np.int16), _c(spl["nKnots_arr"], np.uint32),
         _c(spl["coeff_many"], np.float32), _c(spl["nParamPerEvent_tf1"], np.uint32),
         _c(spl["paramNo_tf1"], np.int16), _c(spl["coeff_tf1"], np.float32)]
    rc = L.m3b_write_monolith_file(os.fsencode(path), C.c_int32(n_params), C.c_int32(max_knots), _p(cx), _p(npt), _p(xp),
                                   C.c_int64(a[0].size // 2), _p(a[0]), _p(a[1]), _p(a[2]), C.c_uint32(a[3].size // 4), _p(a[3]),
                                   _p(a[4]), _p(a[5]), _p(a[6]))
    if rc != OK:
        raise M3BError(rc, (L.m3b_last_error(None) or b"").decode())


def monolith_file_info(path):
    """-> dict(n_events, n_params, max_knots, total_knots) of an M3BMONO1 file; needs no GPU."""
    L = load()
    ne, npar, mk, tk = C.c_int64(0), C.c_int32(0), C.c_int32(0), C.c_uint64(0)
    rc = L.m3b_monolith_file_info(os.fsencode(path), C.byref(ne), C.byref(npar), C.byref(mk), C.byref(tk))
    if rc != OK:
        raise M3BError(rc, (L.m3b_last_error(None) or b"").decode())
    return dict(n_events=ne.value, n_params=npar.value, max_knots=mk.value, total_knots=tk.value)


class Handle:
    """One device context = one sample handler + its spline monolith."""

    @classmethod
    def borrowed(cls, ptr, owner):
        """A Handle object over an m3b_handle* that belongs to something else (a Group member): never destroyed here."""
        self = cls.__new__(cls)
        self.L = load()
        self.h = C.c_void_p(ptr)
        self._owner = owner
        self._borrowed = True
        self.n_samples = self.n_bins = self.n_events = self.n_params = 0
        self._tot = C.c_double(0)
        self._tot_ref = C.byref(self._tot)
        return self

    def __init__(self, device=0, test_statistic=POISSON, update_w2=False, tile_events=0, flags=0):
        self.L = load()
        cfg = Config(device=device, test_statistic=test_statistic, update_w2=int(update_w2),
                     tile_events=tile_events, flags=flags)
        self.h = C.c_void_p()
        rc = self.L.m3b_create(C.byref(cfg), C.byref(self.h))
        if rc != OK:
            raise M3BError(rc, (self.L.m3b_last_error(None) or b"").decode())
        self.n_samples = 0
        self.n_bins = 0
        self.n_events = 0
        self.n_params = 0
        self._tot = C.c_double(0)
        self._tot_ref = C.byref(self._tot)
        self.L.m3b_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.m3b_llh.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]

    def _ck(self, rc):
        if rc != OK:
            raise M3BError(rc, (self.L.m3b_last_error(self.h) or b"").decode())

    def close(self):
        if self.h and not getattr(self, "_borrowed", False):
            self.L.m3b_destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.L.m3b_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    # ---- splines
    def splines_begin(self, n_params, max_knots, coeff_x, n_pts, n_events_total):
        cx, npt = _c(coeff_x, np.float32), _c(n_pts, np.int16)
        self.n_params = int(n_params)
        self._ck(self.L.m3b_splines_begin(self.h, C.c_int32(n_params), C.c_int32(max_knots), _p(cx), _p(npt),
                                          C.c_int64(n_events_total)))

    def splines_append(self, spl):
        a = [_c(spl["nParamPerEvent"], np.uint32), _c(spl["paramNo_arr"], np.int16), _c(spl["nKnots_arr"], np.uint64),
             _c(spl["coeff_many"], np.float32), _c(spl["nParamPerEvent_tf1"], np.uint32),
             _c(spl["paramNo_tf1"], np.int16), _c(spl["coeff_tf1"], np.float32)]
        n = a[0].size // 2
        self._ck(self.L.m3b_splines_append(self.h, C.c_int64(n), _p(a[0]), _p(a[1]), _p(a[2]),
                                           C.c_uint64(a[3].size // 4), _p(a[3]), _p(a[4]), _p(a[5]), _p(a[6])))

    def splines_end(self):
        self._ck(self.L.m3b_splines_end(self.h))

    def upload_spline_monolith(self, n_params, max_knots, coeff_x, n_pts, spl):
        """One-shot upload with the reference's own types (unsigned int knot offsets)."""
        cx, npt = _c(coeff_x, np.float32), _c(n_pts, np.int16)
        a = [_c(spl["nParamPerEvent"], np.uint32), _c(spl["paramNo_arr"], np.int16), _c(spl["nKnots_arr"], np.uint32),
             _c(spl["coeff_many"], np.float32), _c(spl["nParamPerEvent_tf1"], np.uint32),
             _c(spl["paramNo_tf1"], np.int16), _c(spl["coeff_tf1"], np.float32)]
        self.n_params = int(n_params)
        self._ck(self.L.m3b_upload_spline_monolith(
            self.h, C.c_int32(n_params), C.c_int32(max_knots), _p(cx), _p(npt), C.c_int64(a[0].size // 2),
            _p(a[0]), _p(a[1]), _p(a[2]), C.c_uint32(a[3].size // 4), _p(a[3]), _p(a[4]), _p(a[5]), _p(a[6])))

    def upload_from_file(self, path, chunk_events=0):
        """SMonolith::LoadSplineFile + MoveToGPU from the ROOT-free cache file, streamed in chunks."""
        self._ck(self.L.m3b_upload_from_file(self.h, os.fsencode(path), C.c_int64(chunk_events)))
        self.n_params = monolith_file_info(path)["n_params"]

    # ---- binning / events / data
    def upload_binned_splines(self, spl, f64=False):
        """`spl`: the reference's BinnedSplineHandler arrays (see mach3_b200.synth.binned.make_binned_splines).
        f64=True: the reference's default build (M3::float_t = double)."""
        ft = np.float64 if f64 else np.float32
        a = [_c(spl["knot_x"], ft), _c(spl["n_pts"], np.int16), _c(spl["uniquesplinevec_Monolith"], np.int32),
             _c(spl["coeffindexvec"], np.int32), _c(spl["uniquecoeffindices"], np.int32), _c(spl["manycoeff_arr"], ft),
             _c(spl["xcoeff_arr"], ft)]
        fn = self.L.m3b_upload_binned_splines_f64 if f64 else self.L.m3b_upload_binned_splines
        self._ck(fn(self.h, C.c_int32(int(spl["n_params"])), C.c_int32(int(spl["max_knots"])),
                    _p(a[0]), _p(a[1]), C.c_int64(int(spl["n_slots"])), _p(a[2]), _p(a[3]),
                    C.c_int64(a[4].size), _p(a[4]), C.c_int64(a[6].size), _p(a[5]), _p(a[6])))
        self.n_params = int(spl["n_params"])
        self.n_slots = int(spl["n_slots"])
        self.f64 = bool(f64)

    def upload_event_weights_f64(self, static_w):
        sw = _c(static_w, np.float64)
        self._ck(self.L.m3b_upload_event_weights_f64(self.h, C.c_int64(sw.size), _p(sw)))

    def upload_osc_f64(self, osc_w):
        o = _c(osc_w, np.float64)
        self._keep_osc = o
        self._ck(self.L.m3b_upload_osc_f64(self.h, _p(o), C.c_int64(o.size)))

    def upload_event_binned_splines(self, n_per_event, spline_index):
        n, si = _c(n_per_event, np.uint32), _c(spline_index, np.int32)
        self._ck(self.L.m3b_upload_event_binned_splines(self.h, C.c_int64(n.size), _p(n), _p(si)))

    def read_binned_weights(self):
        if getattr(self, "f64", False):
            out = np.zeros(self.n_slots, np.float64)
            self._ck(self.L.m3b_read_binned_weights_f64(self.h, _p(out)))
            return out
        out = np.zeros(self.n_slots, np.float32)
        self._ck(self.L.m3b_read_binned_weights(self.h, _p(out)))
        return out

    def upload_binning(self, edges):
        """edges: list over samples; a uniform sample is a list over dims of edge arrays, a NON-UNIFORM sample
        (Samples/SampleStructs.h:468-528) is a float array [n_boxes, n_dim, 2] of {lo, hi} box extents."""
        ns = len(edges)
        ndim, uniform, nbins, flat, total = np.zeros(ns, np.int32), np.ones(ns, np.int32), np.zeros(ns * 4, np.int32), [], 0
        for s, dims in enumerate(edges):
            if isinstance(dims, np.ndarray) and dims.ndim == 3:
                uniform[s], ndim[s], nbins[s * 4] = 0, dims.shape[1], dims.shape[0]
                flat.append(np.asarray(dims, np.float64).reshape(-1))
                total += dims.shape[0]
                continue
            ndim[s] = len(dims)
            for d, e in enumerate(dims):
                nbins[s * 4 + d] = len(e) - 1
                flat.append(np.asarray(e, np.float64))
            total += int(np.prod([len(e) - 1 for e in dims]))
        flat = np.ascontiguousarray(np.concatenate(flat))
        if uniform.all():
            self._ck(self.L.m3b_upload_binning(self.h, C.c_int32(ns), _p(ndim), _p(nbins), _p(flat)))
        else:
            self._ck(self.L.m3b_upload_binning_ex(self.h, C.c_int32(ns), _p(ndim), _p(uniform), _p(nbins), _p(flat)))
        self.n_samples = ns
        self.n_bins = total

    def upload_events(self, sample_id, kin, norm_idx=None, n_norm_per_event=0, n_norm_values=0, use_osc=False,
                      osc_idx=None, n_osc_values=0, static_w=None):
        sid, k = _c(sample_id, np.int32), _c(kin, np.float64)
        ni, oi, sw = _c(norm_idx, np.int16), _c(osc_idx, np.int32), _c(static_w, np.float32)
        self.n_events = sid.size
        self._ck(self.L.m3b_upload_events(self.h, C.c_int64(sid.size), _p(sid), _p(k), C.c_int32(n_norm_per_event),
                                          _p(ni), C.c_int32(n_norm_values), C.c_int32(int(use_osc)), _p(oi),
                                          C.c_int64(n_osc_values), _p(sw)))

    def update_kinematics(self, kin):
        """Shifted kinematic variables (functional parameters applied on the host): re-bins on the device."""
        k = _c(kin, np.float64)
        self._keep_kin = k
        self._ck(self.L.m3b_update_kinematics(self.h, _p(k)))

    def upload_selection(self, cuts, values=None):
        """SampleHandlerFD::IsEventSelected: cuts = [(sample, var, lower, upper), ...] in StoredSelection order;
        values[n_vars, n_events] = the cut variables (var >= 0 indexes its rows; var = -1-d: binning variable d)."""
        n = len(cuts)
        cs = np.array([c[0] for c in cuts], np.int32); cv = np.array([c[1] for c in cuts], np.int32)
        lo = np.array([c[2] for c in cuts], np.float64); hi = np.array([c[3] for c in cuts], np.float64)
        v = None if values is None else np.ascontiguousarray(np.asarray(values, np.float64).reshape(-1, self.n_events))
        self._ck(self.L.m3b_upload_selection(self.h, C.c_int32(n), _p(cs), _p(cv), _p(lo), _p(hi),
                                             C.c_int32(0 if v is None else v.shape[0]), _p(v)))

    def update_selection_values(self, values):
        v = np.ascontiguousarray(np.asarray(values, np.float64).reshape(-1, self.n_events))
        self._keep_sel = v
        self._ck(self.L.m3b_update_selection_values(self.h, _p(v)))

    def read_event_selected(self):
        out = np.zeros(self.n_events, np.uint8)
        self._ck(self.L.m3b_read_event_selected(self.h, _p(out)))
        return out.astype(bool)

    def upload_linear_shifts(self, n_shift_pars, n_per_event, shift_par, target, coef):
        """SampleHandlerFD::ApplyShifts for linear functional parameters: per event (funcParsGrid order) entries
        {shift_par, target, coef}: x_target += value[shift_par] * coef."""
        npe, sp, tg, cf = _c(n_per_event, np.uint32), _c(shift_par, np.int32), _c(target, np.int32), _c(coef, np.float64)
        self._ck(self.L.m3b_upload_linear_shifts(self.h, C.c_int32(n_shift_pars), C.c_int64(npe.size), _p(npe), _p(sp), _p(tg), _p(cf)))

    def set_shift_pars(self, values):
        v = _c(values, np.float64)
        self._keep_shift = v
        self._ck(self.L.m3b_set_shift_pars(self.h, _p(v)))

    def upload_data(self, data):
        d = _c(data, np.float64)
        self._ck(self.L.m3b_upload_data(self.h, _p(d), C.c_int32(d.size)))

    def upload_osc(self, osc_w):
        o = _c(osc_w, np.float32)
        self._ck(self.L.m3b_upload_osc(self.h, _p(o), C.c_int64(o.size)))

    def register_host_buffer(self, arr):
        self._ck(self.L.m3b_register_host_buffer(self.h, _p(arr), C.c_uint64(arr.nbytes)))

    def alloc_host(self, n, dtype=np.float32):
        """numpy array over pinned + mapped host memory owned by the handle (valid until close())."""
        dt = np.dtype(dtype)
        p = C.c_void_p()
        self._ck(self.L.m3b_alloc_host(self.h, C.c_uint64(int(n) * dt.itemsize), C.byref(p)))
        buf = (C.c_char * (int(n) * dt.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=dt, count=int(n))

    def free_host(self, arr):
        """Give an alloc_host array back (the array must not be used afterwards)."""
        self._ck(self.L.m3b_free_host(self.h, C.c_void_p(arr.ctypes.data)))

    def set_test_statistic(self, ts):
        self._ck(self.L.m3b_set_test_statistic(self.h, C.c_int32(ts)))

    def set_flags(self, set_mask=0, clear_mask=0):
        self._ck(self.L.m3b_set_flags(self.h, C.c_int32(set_mask), C.c_int32(clear_mask)))

    def reset_w2(self):
        self._ck(self.L.m3b_reset_w2(self.h))

    # ---- step
    def step(self, spline_pars, norm_pars=None, osc_w=None, mode="fused"):
        """Asynchronous.  Arrays must stay alive until the next synchronising call."""
        sp = _c(spline_pars, np.float64)
        nm = _c(norm_pars, np.float64)
        if osc_w is not None:
            assert osc_w.dtype == np.float32 and osc_w.flags.c_contiguous
        self._keep = (sp, nm, osc_w)
        fn = {"fused": self.L.m3b_step, "fill": self.L.m3b_step_fill, "peer": self.L.m3b_step_peer}[mode]
        self._ck(fn(self.h, _p(sp), _p(nm), _p(osc_w)))

    def step_batch(self, spline_pars, norm_pars=None, osc_w=None, per_sample=False):
        """n_sets proposals (rows of spline_pars / norm_pars) in one call; returns -lnL per set."""
        sp = _c(spline_pars, np.float64)
        n_sets = sp.shape[0] if sp is not None else np.asarray(norm_pars).shape[0]
        nm = _c(norm_pars, np.float64)
        tot = np.zeros(n_sets, np.float64)
        ps = np.zeros((n_sets, max(self.n_samples, 1)), np.float64) if per_sample else None
        self._ck(self.L.m3b_step_batch(self.h, C.c_int32(n_sets), _p(sp), _p(nm), _p(osc_w), _p(tot), _p(ps)))
        return (tot, ps[:, :self.n_samples]) if per_sample else tot

    def step_batch_hist(self, spline_pars, norm_pars=None, osc_w=None):
        """Like step_batch, and every set's MC histogram comes back: (-lnL[n_sets], mc[n_sets, n_bins])."""
        sp = _c(spline_pars, np.float64)
        n_sets = sp.shape[0] if sp is not None else np.asarray(norm_pars).shape[0]
        nm = _c(norm_pars, np.float64)
        tot = np.zeros(n_sets, np.float64)
        mc = np.zeros((n_sets, self.n_bins), np.float64)
        self._ck(self.L.m3b_step_batch_hist(self.h, C.c_int32(n_sets), _p(sp), _p(nm), _p(osc_w), _p(tot), None, _p(mc)))
        return tot, mc

    def step_addr(self, spline_pars_addr, norm_pars_addr=0, osc_w_addr=0):
        """m3b_step on raw addresses (float64 spline pars, float64 norm pars, float32 osc weights or 0)."""
        rc = self.L.m3b_step(self.h, spline_pars_addr or None, norm_pars_addr or None, osc_w_addr or None)
        if rc != OK:
            self._ck(rc)

    def llh_fast(self):
        """m3b_llh total only, no allocation."""
        rc = self.L.m3b_llh(self.h, self._tot_ref, None)
        if rc != OK:
            self._ck(rc)
        return self._tot.value

    def step_segments(self, param_values, segments, norm_pars=None, osc_w=None):
        pv, sg, nm = _c(param_values, np.float32), _c(segments, np.int16), _c(norm_pars, np.float64)
        self._keep = (pv, sg, nm, osc_w)
        self._ck(self.L.m3b_step_segments(self.h, _p(pv), _p(sg), _p(nm), _p(osc_w)))

    def eval_weights(self, param_values, segments, host_out):
        """SMonolithGPU::RunGPU_SplineMonolith: asynchronous; host_out valid after synchronize()."""
        pv, sg = _c(param_values, np.float32), _c(segments, np.int16)
        assert host_out.dtype == np.float32 and host_out.flags.c_contiguous
        self._keep = (pv, sg, host_out)
        self._ck(self.L.m3b_eval_weights(self.h, _p(pv), _p(sg), _p(host_out)))
        if self.n_events == 0:
            self.n_events = host_out.size

    def llh(self, per_sample=False):
        tot = C.c_double(0)
        ps = np.zeros(max(self.n_samples, 1), np.float64)
        self._ck(self.L.m3b_llh(self.h, C.byref(tot), _p(ps)))
        return (tot.value, ps[:self.n_samples]) if per_sample else tot.value

    def find_segments(self, spline_pars):
        sp = _c(spline_pars, np.float64)
        seg = np.zeros(self.n_params, np.int16)
        val = np.zeros(self.n_params, np.float32)
        self._ck(self.L.m3b_find_segments(self.h, _p(sp), _p(seg), _p(val)))
        return seg, val

    def set_spline_knots_f64(self, x_pts):
        """FastSplineInfo::xPts as doubles ([n_params, max_knots]); None restores the float knots of coeff_x."""
        self._ck(self.L.m3b_set_spline_knots_f64(self.h, None if x_pts is None else _p(_c(x_pts, np.float64))))

    def synchronize(self):
        self._ck(self.L.m3b_synchronize(self.h))

    # ---- read-back
    def read_hist(self):
        mc, w2 = np.zeros(self.n_bins, np.float64), np.zeros(self.n_bins, np.float64)
        self._ck(self.L.m3b_read_hist(self.h, _p(mc), _p(w2)))
        return mc, w2

    def read_event_weights(self):
        if getattr(self, "f64", False):
            sw, tw = np.zeros(self.n_events, np.float64), np.zeros(self.n_events, np.float64)
            self._ck(self.L.m3b_read_event_weights_f64(self.h, _p(sw), _p(tw)))
            return sw, tw
        sw, tw = np.zeros(self.n_events, np.float32), np.zeros(self.n_events, np.float32)
        self._ck(self.L.m3b_read_event_weights(self.h, _p(sw), _p(tw)))
        return sw, tw

    def read_event_bins(self):
        b = np.zeros(self.n_events, np.int32)
        self._ck(self.L.m3b_read_event_bins(self.h, _p(b)))
        return b

    # ---- multi-GPU
    def hist_device_ptr(self):
        ptr, nb, live = C.c_void_p(), C.c_int32(), C.c_int32()
        self._ck(self.L.m3b_hist_device_ptr(self.h, C.byref(ptr), C.byref(nb), C.byref(live)))
        return ptr.value, nb.value, bool(live.value)

    def llh_from_hist(self):
        self._ck(self.L.m3b_llh_from_hist(self.h))

    def peer_export(self, rank, world):
        buf = (C.c_ubyte * 64)()
        self._ck(self.L.m3b_peer_export(self.h, C.c_int32(rank), C.c_int32(world), buf))
        return bytes(buf)

    def peer_import(self, peer_rank, handle_bytes):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle_bytes)
        self._ck(self.L.m3b_peer_import(self.h, C.c_int32(peer_rank), buf))

    def set_timing(self, on=True):
        self._ck(self.L.m3b_set_timing(self.h, C.c_int32(int(on))))

    def kernel_time(self):
        """-> (summed fill-kernel ms, launches) since the last call; synchronises."""
        ms, n = C.c_double(0), C.c_int64(0)
        self._ck(self.L.m3b_kernel_time(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def block_trace(self, read=True):
        """Per-block timeline of the last fill launch: array [grid, 8] of globaltimer ns (see m3b200.h)."""
        g = C.c_int32(0)
        self._ck(self.L.m3b_block_trace(self.h, None, C.byref(g)))
        if not read:
            return None
        out = np.zeros((max(g.value, 1), 8), np.uint64)
        self._ck(self.L.m3b_block_trace(self.h, _p(out), C.byref(g)))
        return out[:g.value]

    def info(self) -> Info:
        i = Info()
        self._ck(self.L.m3b_get_info(self.h, C.byref(i)))
        return i


class Group:
    """ONE sample handler over several devices, driven by one process / one calling thread (m3b_group_*): the form in
    which the reference's single-process fitters (Fitters/MR2T2.cpp:62-74) reach more than one GPU."""

    def __init__(self, devices, test_statistic=POISSON, update_w2=False, tile_events=0, flags=0):
        self.L = load()
        cfg = Config(device=0, test_statistic=test_statistic, update_w2=int(update_w2), tile_events=tile_events, flags=flags)
        dev = np.ascontiguousarray(devices, np.int32)
        self.g = C.c_void_p()
        rc = self.L.m3b_group_create(C.byref(cfg), _p(dev), C.c_int32(dev.size), C.byref(self.g))
        if rc != OK:
            raise M3BError(rc, (self.L.m3b_last_error(None) or b"").decode())
        self.n = int(dev.size)
        self.devices = [int(d) for d in dev]
        self._members = [Handle.borrowed(self.L.m3b_group_member(self.g, i), self) for i in range(self.n)]
        self.exchange = None
        self.n_bins = self.n_samples = 0
        self._tot = C.c_double(0)
        self._tot_ref = C.byref(self._tot)

    def _ck(self, rc):
        if rc != OK:
            raise M3BError(rc, (self.L.m3b_group_last_error(self.g) or b"").decode())

    def close(self):
        if self.g:
            self.L.m3b_group_destroy(self.g)
            self.g = C.c_void_p()
            for m in self._members:
                m.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def member(self, i) -> Handle:
        return self._members[i]

    def shard(self, n_events, i):
        e0, e1 = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.m3b_group_shard(self.g, C.c_int64(n_events), C.c_int32(i), C.byref(e0), C.byref(e1)))
        return e0.value, e1.value

    # ---- whole-workload uploads (the reference's arrays, sliced by the library)
    def upload_spline_monolith(self, n_params, max_knots, coeff_x, n_pts, spl):
        cx, npt = _c(coeff_x, np.float32), _c(n_pts, np.int16)
        a = [_c(spl["nParamPerEvent"], np.uint32), _c(spl["paramNo_arr"], np.int16), _c(spl["nKnots_arr"], np.uint32),
             _c(spl["coeff_many"], np.float32), _c(spl["nParamPerEvent_tf1"], np.uint32),
             _c(spl["paramNo_tf1"], np.int16), _c(spl["coeff_tf1"], np.float32)]
        self._ck(self.L.m3b_group_upload_spline_monolith(
            self.g, C.c_int32(n_params), C.c_int32(max_knots), _p(cx), _p(npt), C.c_int64(a[0].size // 2),
            _p(a[0]), _p(a[1]), _p(a[2]), C.c_uint32(a[3].size // 4), _p(a[3]), _p(a[4]), _p(a[5]), _p(a[6])))
        for m in self._members:
            m.n_params = int(n_params)

    def upload_from_file(self, path, chunk_events=0):
        self._ck(self.L.m3b_group_upload_from_file(self.g, os.fsencode(path), C.c_int64(chunk_events)))
        for m in self._members:
            m.n_params = monolith_file_info(path)["n_params"]

    def upload_binning(self, edges):
        for m in self._members:
            m.upload_binning(edges)
        self.n_bins, self.n_samples = self._members[0].n_bins, self._members[0].n_samples

    def upload_events(self, sample_id, kin, norm_idx=None, n_norm_per_event=0, n_norm_values=0, use_osc=False,
                      osc_idx=None, n_osc_values=0, static_w=None):
        sid, k = _c(sample_id, np.int32), _c(kin, np.float64)
        ni, oi, sw = _c(norm_idx, np.int16), _c(osc_idx, np.int32), _c(static_w, np.float32)
        self._ck(self.L.m3b_group_upload_events(self.g, C.c_int64(sid.size), _p(sid), _p(k), C.c_int32(n_norm_per_event),
                                                _p(ni), C.c_int32(n_norm_values), C.c_int32(int(use_osc)), _p(oi),
                                                C.c_int64(n_osc_values), _p(sw)))
        for i, m in enumerate(self._members):
            e0, e1 = self.shard(sid.size, i)
            m.n_events = e1 - e0

    def upload_selection(self, cuts, values=None, n_events=None):
        cs = np.array([c[0] for c in cuts], np.int32); cv = np.array([c[1] for c in cuts], np.int32)
        lo = np.array([c[2] for c in cuts], np.float64); hi = np.array([c[3] for c in cuts], np.float64)
        n_events = n_events or sum(m.n_events for m in self._members)
        v = None if values is None else np.ascontiguousarray(np.asarray(values, np.float64).reshape(-1, n_events))
        self._ck(self.L.m3b_group_upload_selection(self.g, C.c_int32(len(cuts)), _p(cs), _p(cv), _p(lo), _p(hi),
                                                   C.c_int32(0 if v is None else v.shape[0]), _p(v)))

    def upload_linear_shifts(self, n_shift_pars, n_per_event, shift_par, target, coef):
        npe, sp, tg, cf = _c(n_per_event, np.uint32), _c(shift_par, np.int32), _c(target, np.int32), _c(coef, np.float64)
        self._ck(self.L.m3b_group_upload_linear_shifts(self.g, C.c_int32(n_shift_pars), C.c_int64(npe.size), _p(npe), _p(sp), _p(tg), _p(cf)))

    def set_shift_pars(self, values):
        v = _c(values, np.float64)
        self._ck(self.L.m3b_group_set_shift_pars(self.g, _p(v)))

    def upload_data(self, data):
        d = _c(data, np.float64)
        self._ck(self.L.m3b_group_upload_data(self.g, _p(d), C.c_int32(d.size)))

    def upload_osc(self, osc_w):
        o = _c(osc_w, np.float32)
        self._ck(self.L.m3b_group_upload_osc(self.g, _p(o), C.c_int64(o.size)))

    def connect(self, exchange="peer"):
        if not self.n_bins:
            self.n_bins, self.n_samples = self._members[0].n_bins, self._members[0].n_samples
        self._ck(self.L.m3b_group_connect(self.g, C.c_int32(EXCHANGE_NCCL if exchange == "nccl" else EXCHANGE_PEER)))
        self.exchange = exchange

    def alloc_host(self, n, dtype=np.float32):
        dt = np.dtype(dtype)
        p = C.c_void_p()
        self._ck(self.L.m3b_group_alloc_host(self.g, C.c_uint64(int(n) * dt.itemsize), C.byref(p)))
        buf = (C.c_char * (int(n) * dt.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=dt, count=int(n))

    def step(self, spline_pars, norm_pars=None, osc_w=None):
        sp, nm = _c(spline_pars, np.float64), _c(norm_pars, np.float64)
        if osc_w is not None:
            assert osc_w.dtype == np.float32 and osc_w.flags.c_contiguous
        self._keep = (sp, nm, osc_w)
        rc = self.L.m3b_group_step(self.g, None if sp is None else sp.ctypes.data, None if nm is None or nm.size == 0 else nm.ctypes.data,
                                   None if osc_w is None else osc_w.ctypes.data)
        if rc != OK:
            self._ck(rc)

    def llh(self, per_sample=False):
        if not per_sample:
            rc = self.L.m3b_group_llh(self.g, self._tot_ref, None)
            if rc != OK:
                self._ck(rc)
            return self._tot.value
        ps = np.zeros(max(self.n_samples, 1), np.float64)
        self._ck(self.L.m3b_group_llh(self.g, self._tot_ref, ps.ctypes.data))
        return self._tot.value, ps[:self.n_samples]

    def read_hist(self):
        mc, w2 = np.zeros(self.n_bins, np.float64), np.zeros(self.n_bins, np.float64)
        self._ck(self.L.m3b_group_read_hist(self.g, _p(mc), _p(w2)))
        return mc, w2

    def synchronize(self):
        self._ck(self.L.m3b_group_synchronize(self.g))
