"""Generates tests/golden/ref_host_baseline.npz: BASELINE config 1 (and the shapes of configs 2 and 3 on samples)
-- config 1 (the reference's own CPU-runnable case: 100k events x
(10 TSpline3, 5 knots + 2 linear), 1-D 50 bins, Poisson) run at FULL size through the REFERENCE's own
SampleHandlerFD::Reweight + GetLikelihood over its SMonolith (oracle/_ref/libm3ref_path_lm.so: the reference
sources compiled from /root/reference, serial float build).  Stored: the data histogram, and per proposal the MC
histogram and -lnL.      python tests/golden/make_ref_host_baseline.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mach3_b200 import synth                   # noqa: E402
from oracle import ref_path_binding as RP      # noqa: E402

STEPS = (-1, 0, 1, 2, 3, -2, -3, 4)


def workloads():
    """cfg1 at full size; cfg2's shape (50 responses, 7 knots, 60x15 bins) on 30k events; cfg3's (60 responses,
    4 samples x 80x20 bins) on 40k events."""
    return {"cfg1": synth.CFG1, "cfg2s": synth.CFG2.scaled(30_000), "cfg3s": synth.CFG3.scaled(40_000)}


def run(w, build="float"):
    typ, npts, cx = synth.param_layout(w)
    spl, ev = synth.make_splines(w), synth.make_events(w)
    mono = RP.RefSMonolith.from_arrays(w.n_params, w.n_knots, cx, npts, typ, spl, build=build)
    fd = RP.RefSampleHandlerFD(synth.bin_edges(w), w.test_statistic, False, build=build)
    fd.attach_monolith(mono)
    E = w.n_events
    idx = np.arange(E, dtype=np.int32)
    fd.set_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, w_before=idx, w_after=E + idx, n_pool=2 * E)
    pool = np.concatenate([synth.make_osc(w, 0), ev["static_w"]]).astype(np.float64)
    out = {"mc": [], "llh": []}
    for i, step in enumerate(STEPS):
        sp, nm = synth.proposal(w, step)
        fd.reweight(sp, nm, pool if i == 0 else None)
        if i == 0:
            out["data"] = np.random.default_rng(w.seed).poisson(fd.hist()[0]).astype(np.float64)
            fd.set_data(out["data"])
        out["mc"].append(fd.hist()[0]); out["llh"].append(fd.llh())
    fd.close()
    return {k: np.asarray(v) for k, v in out.items()}


if __name__ == "__main__":
    out = {}
    for name, w in workloads().items():
        r = run(w)
        for k, v in r.items():
            out[f"{name}/{k}"] = v
        print(name, "-lnL", r["llh"][:4])
    path = os.path.join(ROOT, "tests", "golden", "ref_host_baseline.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
