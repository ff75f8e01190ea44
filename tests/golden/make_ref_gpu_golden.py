"""Generates tests/golden/ref_gpu_weights.npz: per-event spline weights computed by the REFERENCE's
own CUDA kernels (Splines/gpuSplineUtils.cu built into oracle/_ref) on a B200, for small seeded
synthetic monoliths.  Run on the GPU box:

    gpurun -- 'python tests/golden/make_ref_gpu_golden.py gpurun_out/ref_gpu_weights.npz'

then copy the file to tests/golden/.  The CPU test tests/test_golden.py checks the oracle (and the
GPU test the B200 kernel) against these vectors; a checksum of the generated inputs guards against
generator drift."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mach3_b200 import synth          # noqa: E402
from oracle import binding as O       # noqa: E402
from oracle import ref_gpu_binding as R  # noqa: E402

CASES = [("CFG1", 3000, [-1, 0, 1, 2, -2, -3, -4]), ("SPARSE", 3000, [-1, 0, 1, 2, -2]),
         ("SPARSE_RUNS", 3000, [-1, 0, 3])]


def inputs_digest(spl, cx):
    h = hashlib.sha256()
    for k in ("nParamPerEvent", "paramNo_arr", "nKnots_arr", "coeff_many", "nParamPerEvent_tf1", "paramNo_tf1", "coeff_tf1"):
        h.update(np.ascontiguousarray(spl[k]).tobytes())
    h.update(np.ascontiguousarray(cx).tobytes())
    return h.hexdigest()


def main(out):
    res = {}
    for name, n, steps in CASES:
        w = getattr(synth, name).scaled(n)
        typ, npts, cx = synth.param_layout(w)
        spl = synth.make_splines(w)
        ref = R.RefSMonolithGPU(w.n_params, w.n_knots, cx, spl)
        mono = O.SMonolith(w.n_params, w.n_knots, cx, npts, spl)     # only for FindSplineSegment's history
        res[f"{name}.n_events"] = np.int64(n)
        res[f"{name}.digest"] = np.frombuffer(inputs_digest(spl, cx).encode(), np.uint8)
        res[f"{name}.steps"] = np.array(steps, np.int64)
        for i, st in enumerate(steps):
            sp, _ = synth.proposal(w, st)
            mono.set_params(sp); mono.FindSplineSegment()
            res[f"{name}.pars.{i}"] = sp.copy()
            res[f"{name}.segments.{i}"] = mono.segments.copy()
            res[f"{name}.values.{i}"] = mono.param_values.copy()
            res[f"{name}.weights.{i}"] = ref.run(mono.param_values, mono.segments)
        ref.close()
    np.savez_compressed(out, **res)
    print("wrote", out, {k: v.shape for k, v in res.items() if "weights" in k})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_gpu_weights.npz"))
