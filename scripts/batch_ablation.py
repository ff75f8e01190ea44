#!/usr/bin/env python
"""Timing ablations of fill_batch2_kernel (experiments library only; the results of the ablated launches are wrong on
purpose).  M3B_LIB=mach3_b200/libm3b200_exp.so python scripts/batch_ablation.py [events]"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                              # noqa: E402
from mach3_b200 import lib, synth                         # noqa: E402

n_events = int(sys.argv[1]) if len(sys.argv) > 1 else 1_200_000
w = synth.CFG5.scaled(n_events)
h = lib.Handle(device=0, test_statistic=w.test_statistic, update_w2=False)
bench.upload_monolith(h, w, 0, w.n_events)
h.upload_binning(synth.bin_edges(w))
ev = synth.make_events(w, 0, w.n_events)
h.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, None, 0, ev["static_w"])
h.upload_osc(synth.make_osc(w, 0, 0, w.n_events))
sp, nm = synth.proposal(w, -1)
h.step(sp, nm); h.llh()
h.upload_data(np.random.default_rng(w.seed).poisson(h.read_hist()[0]).astype(np.float64))
rng = np.random.default_rng(w.seed + 7)
sp0, nm0 = synth.proposal(w, 1)
sps = np.clip(sp0[None, :] + rng.normal(0, 0.3, (256, w.n_params)), -2.9, 2.9)
nms = np.clip(nm0[None, :] + rng.normal(0, 0.05, (256, w.n_norm_params)), 0.5, 1.5)
scan = np.tile(sp0, (256, 1)); scan[:, 0] = np.linspace(-2.9, 2.9, 256)
h.set_timing(True)
for name, pars in (("sigma 0.3 throws", sps), ("LLH scan (one segment per slot)", scan)):
    for dbg, what in ((0, "full kernel"), (1, "no atomics"), (2, "no epilogue"), (4, "no spline arithmetic"), (8, "no re-pack / second barrier"),
                      (6, "no arithmetic, no epilogue: ring + tables + barriers only"), (14, "ring + tables only")):
        os.environ["M3B_BATCH_DBG"] = str(dbg)
        ms = []
        for rep in range(3):
            h.kernel_time()
            try:
                h.step_batch(pars, nms)
            except Exception as e:                        # ablated launches may trip the math check: timing is still valid
                pass
            t, n = h.kernel_time()
            ms.append(t / max(n, 1))
        print(f"{name:34s} dbg {dbg:2d} {what:58s} kernel_ms {min(ms):8.3f}", flush=True)
