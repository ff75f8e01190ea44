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
LIB_PATH = os.path.join(_HERE, "_ref", "libm3ref_path.so")
_L = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _L
    if _L is None:
        L = C.CDLL(LIB_PATH)
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
        _L = L
    return _L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


_ARRAYS = (("coeff_x", np.float32), ("coeff_many", np.float32), ("nKnots_arr", np.uint32), ("paramNo_arr", np.int16),
           ("nParamPerEvent", np.uint32), ("nParamPerEvent_tf1", np.uint32), ("paramNo_tf1", np.int16),
           ("coeff_tf1", np.float32), ("n_pts", np.int16), ("x_pts_f64", np.float64))


class RefSMonolith:
    """The reference's SMonolith built from per-event response functions.

    type[P]: 0 TSpline3_red, 1 TF1_red.  npts[n_events, P]: knots (0: the event has no response to the parameter).
    vals[total_knots, 5] = {x, y, b, c, d} per knot in (event, parameter, knot) order; TF1: column 1 = coefficient."""

    def __init__(self, type_, npts, vals):
        L = lib()
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

    def arrays(self):
        """The monolith arrays exactly as SMonolith::PrepareForGPU left them (the arguments of
        SMonolithGPU::CopyToGPU_SplineMonolith, and of m3b_upload_spline_monolith)."""
        L = lib()
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
        if lib().refp_mono_evaluate(self.h, _p(pars), _p(w), _p(seg), _p(val)):
            raise RuntimeError("the reference threw in SMonolith::Evaluate")
        return w, seg, val

    def close(self):
        if self.h:
            lib().refp_mono_destroy(self.h)
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


def low_mc_bound():
    return lib().refp_low_mc_bound()
