"""ctypes front-end of oracle/_ref/libm3ref_gpu_P<N>.so: the REFERENCE's own CUDA spline kernels
(Splines/gpuSplineUtils.cu) built from /root/reference by oracle/ref_gpu/Makefile.
TEST INFRASTRUCTURE ONLY (second oracle for the per-event spline weights + incumbent-GPU timing)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path(n_params):
    return os.path.join(_HERE, "_ref", f"libm3ref_gpu_P{n_params}.so")


def available(n_params):
    return os.path.exists(lib_path(n_params))


def adapter_path():
    """adapters/SMonolithGPU_m3b200.cu (the drop-in class on top of libm3b200) + the same harness."""
    return os.path.join(_HERE, "_ref", "libm3adapter_gpu.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefSMonolithGPU:
    """The reference's SMonolithGPU driven like SMonolith does (MoveToGPU / Evaluate)."""

    def __init__(self, n_params, max_knots, coeff_x, spl, adapter=False):
        L = C.CDLL(adapter_path() if adapter else lib_path(n_params))
        L.m3ref_create.restype = C.c_void_p
        L.m3ref_total_weights.restype = C.c_void_p
        L.m3ref_time_ms.restype = C.c_double
        L.m3ref_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.m3ref_time_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.m3ref_total_weights.argtypes = [C.c_void_p]
        L.m3ref_destroy.argtypes = [C.c_void_p]
        assert L.m3ref_compiled_nparams() == (-1 if adapter else n_params)
        self.L = L
        self.n_events = int(spl["n_events"])
        self.n_params = n_params
        a = [np.ascontiguousarray(coeff_x, np.float32), np.ascontiguousarray(spl["nParamPerEvent"], np.uint32),
             np.ascontiguousarray(spl["paramNo_arr"], np.int16), np.ascontiguousarray(spl["nKnots_arr"], np.uint32),
             np.ascontiguousarray(spl["coeff_many"], np.float32), np.ascontiguousarray(spl["nParamPerEvent_tf1"], np.uint32),
             np.ascontiguousarray(spl["paramNo_tf1"], np.int16), np.ascontiguousarray(spl["coeff_tf1"], np.float32)]
        self._keep = a
        self.h = L.m3ref_create(C.c_int(n_params), C.c_int(max_knots), _p(a[0]), C.c_uint(self.n_events), _p(a[1]),
                                _p(a[2]), _p(a[3]), C.c_uint(a[4].size // 4), _p(a[4]), _p(a[5]), _p(a[6]), _p(a[7]))
        if not self.h:
            raise RuntimeError("reference SMonolithGPU could not be created (nParams != NSplines_GPU or CUDA error)")

    def run(self, param_values, segments):
        pv = np.ascontiguousarray(param_values, np.float32)
        sg = np.ascontiguousarray(segments, np.int16)
        if self.L.m3ref_run(self.h, _p(pv), _p(sg)) != 0:
            raise RuntimeError("reference GPU kernels failed")
        ptr = self.L.m3ref_total_weights(self.h)
        buf = (C.c_char * (4 * self.n_events)).from_address(ptr)
        return np.frombuffer(buf, np.float32, self.n_events).copy()

    def time_ms(self, param_values, segments, laps=20):
        pv = np.ascontiguousarray(param_values, np.float32)
        sg = np.ascontiguousarray(segments, np.int16)
        return self.L.m3ref_time_ms(self.h, _p(pv), _p(sg), C.c_int(laps))

    def close(self):
        if self.h:
            self.L.m3ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
