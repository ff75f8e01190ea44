"""Generates tests/golden/ref_host_path.npz: outputs of the REFERENCE's own host code for the likelihood path --
SMonolith (constructor flattening + Evaluate = FindSplineSegment + CalcSplineWeights + CalcTotalEventWeight) and
SampleHandlerBase::GetTestStatLLH -- compiled from /root/reference by oracle/ref_host/Makefile into
oracle/_ref/libm3ref_path.so and run HERE (this container) on the seeded inputs of tests/refpath_cases.py.

    python tests/golden/make_ref_host_path.py

Stored per case: the monolith arrays the reference built, and per proposal the segments, float parameter values and
per-event weights it computed; for the test statistics the per-bin values of all five statistics."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refpath_cases as RC                      # noqa: E402
from oracle import ref_path_binding as RP      # noqa: E402


def digest(c):
    h = hashlib.sha256()
    for k in ("type", "npts", "vals", "pars"):
        h.update(np.ascontiguousarray(c[k]).tobytes())
    return np.frombuffer(h.digest(), np.uint8)


def main():
    out = {}
    for name in RC.CASES:
        c = RC.make_case(name)
        m = RP.RefSMonolith(c["type"], c["npts"], c["vals"])
        for k, v in m.arrays().items():
            out[f"{name}/arr/{k}"] = np.asarray(v)
        out[f"{name}/sizes"] = np.array([m.NEvents, m.nParams, m.max_knots, m.NSplines_valid, m.NTF1_valid, m.nKnots, m.nTF1coeff], np.int64)
        W, S, V = [], [], []
        for t in range(c["pars"].shape[0]):
            w, s, v = m.evaluate(c["pars"][t])
            W.append(w); S.append(s); V.append(v)
        out[f"{name}/weights"] = np.stack(W)
        out[f"{name}/segments"] = np.stack(S)
        out[f"{name}/param_values"] = np.stack(V)
        out[f"{name}/input_sha256"] = digest(c)
        m.close()
        print(name, "events", m.NEvents, "splines", m.NSplines_valid, "tf1", m.NTF1_valid, "knots", m.nKnots, "max_knots", m.max_knots)
    d, mc, w2 = RC.stat_inputs()
    out["stat/data"], out["stat/mc"], out["stat/w2"] = d, mc, w2
    for kind in range(5):
        v, thrown = RP.test_stat(kind, d, mc, w2)
        assert thrown == 0, (kind, thrown)
        out[f"stat/llh{kind}"] = v
    out["stat/poisson"] = RP.poisson(d, mc)
    out["stat/low_mc_bound"] = np.array([RP.low_mc_bound()])
    path = os.path.join(ROOT, "tests", "golden", "ref_host_path.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
