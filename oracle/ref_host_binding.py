"""ctypes front-end of oracle/_ref/libm3ref_host.so: the REFERENCE's own binning code (struct SampleBinningInfo,
Samples/SampleStructs.h) compiled from /root/reference by oracle/ref_host/Makefile.  TEST INFRASTRUCTURE ONLY:
pins the oracle's FindBin / non-uniform binning restatement to the reference itself."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libm3ref_host.so")


def available():
    return os.path.exists(LIB_PATH)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefBinning:
    """One SampleBinningInfo of the reference: uniform (list over dims of edge arrays) or non-uniform (array
    [n_boxes, n_dim, 2])."""

    def __init__(self, spec):
        L = C.CDLL(LIB_PATH)
        L.refh_uniform.restype = C.c_void_p
        L.refh_nonuniform.restype = C.c_void_p
        L.refh_edge.restype = C.c_double
        for f in ("refh_destroy", "refh_nbins"):
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("refh_axis_nbins", "refh_stride", "refh_grid_size"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        L.refh_edge.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.refh_grid_entry.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.refh_find_bin.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refh_find_sample_bin.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L = L
        if isinstance(spec, np.ndarray) and spec.ndim == 3:
            ext = np.ascontiguousarray(spec, np.float64)
            self.n_dim = ext.shape[1]
            self.h = L.refh_nonuniform(C.c_int(self.n_dim), C.c_int(ext.shape[0]), _p(ext))
        else:
            self.n_dim = len(spec)
            nb = np.array([len(e) - 1 for e in spec], np.int32)
            ed = np.ascontiguousarray(np.concatenate([np.asarray(e, np.float64) for e in spec]))
            self.h = L.refh_uniform(C.c_int(self.n_dim), _p(nb), _p(ed))
        if not self.h:
            raise RuntimeError("the reference rejected this binning (MaCh3Exception)")

    @property
    def n_bins(self):
        return self.L.refh_nbins(self.h)

    def axis_edges(self, d):
        n = self.L.refh_axis_nbins(self.h, d)
        return np.array([self.L.refh_edge(self.h, d, i) for i in range(n + 1)])

    def grid_mapping(self, mega):
        return [self.L.refh_grid_entry(self.h, mega, k) for k in range(self.L.refh_grid_size(self.h, mega))]

    def find_bin(self, dim, var, nom_bin):
        var = np.ascontiguousarray(var, np.float64); nom = np.ascontiguousarray(nom_bin, np.int32)
        out = np.zeros(var.size, np.int32)
        self.L.refh_find_bin(self.h, C.c_int(dim), C.c_int(var.size), _p(var), _p(nom), _p(out))
        return out

    def find_sample_bin(self, kin, nom_bin):
        """kin, nom_bin: [n_dim, n]"""
        kin = np.ascontiguousarray(kin, np.float64); nom = np.ascontiguousarray(nom_bin, np.int32)
        n = kin.shape[1]
        out = np.zeros(n, np.int32)
        self.L.refh_find_sample_bin(self.h, C.c_int(n), _p(kin), _p(nom), _p(out))
        return out

    def close(self):
        if self.h:
            self.L.refh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
