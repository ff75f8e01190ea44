"""The incumbent: MaCh3's own CUDA spline kernels (Splines/gpuSplineUtils.cu built into oracle/_ref, driven like
SMonolith::Evaluate + SynchroniseMemTransfer) on BASELINE config 2, next to the drop-in adapter (same call sequence on
libm3b200) and the fused step.  The reference's GPU path stops at per-event spline weights (n_events x 4 B copied to the
host every step); fill and likelihood then run on the CPU.    python scripts/incumbent_gpu.py [n_events]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mach3_b200 import handlers, lib, synth          # noqa: E402
from oracle import binding as O                        # noqa: E402
from oracle import ref_gpu_binding as R                # noqa: E402

w = synth.CFG2 if len(sys.argv) < 2 else synth.CFG2.scaled(int(sys.argv[1]))
typ, npts, cx = synth.param_layout(w)
t0 = time.time()
spl = synth.make_splines(w)
print(f"monolith in the reference layout: {spl['coeff_many'].nbytes / 1e9:.2f} GB of knots, built in {time.time() - t0:.1f} s")
omono = O.SMonolith(w.n_params, w.n_knots, cx, npts, spl)          # only for FindSplineSegment on the host
pars = synth.proposal(w, 3)[0]
omono.set_params(pars); omono.FindSplineSegment()
vals, segs = omono.param_values.copy(), omono.segments.copy()
out = {}
for name, adapter in (("reference gpuSplineUtils.cu", False), ("adapter SMonolithGPU_m3b200.cu", True)):
    if not adapter and not R.available(w.n_params):
        print("reference kernels not built for", w.n_params, "parameters"); continue
    g = R.RefSMonolithGPU(w.n_params, w.n_knots, cx, spl, adapter=adapter)
    wts = g.run(vals, segs)
    ms = g.time_ms(vals, segs, laps=50)
    out[name] = (ms, wts)
    print(f"{name:32s} Evaluate + SynchroniseMemTransfer: {ms * 1e3:8.1f} us per step  ({w.n_events / ms / 1e6:.2f} G events/s, weights only)")
    g.close()
if len(out) == 2:
    a, b = out["reference gpuSplineUtils.cu"][1], out["adapter SMonolithGPU_m3b200.cu"][1]
    print("weights bit-identical:", bool(np.array_equal(a, b)))
del spl
