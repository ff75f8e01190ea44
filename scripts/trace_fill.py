"""Per-block timeline of one fill launch on cfg2 (debug aid): ramp-up, steady state, tail."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mach3_b200 import lib, synth

w = synth.CFG2
if len(sys.argv) > 1:
    w = w.scaled(int(sys.argv[1]))
h = lib.Handle(tile_events=int(os.environ.get("M3B_TILE", "0")))
typ, npts, cx = synth.param_layout(w)
h.splines_begin(w.n_params, w.n_knots, cx, npts, w.n_events)
C = 131072
for c0 in range(0, w.n_events, C):
    h.splines_append(synth.make_splines(w, c0, min(w.n_events, c0 + C)))
h.splines_end()
h.upload_binning(synth.bin_edges(w))
ev = synth.make_events(w, 0, w.n_events)
h.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, None, 0, ev["static_w"])
h.upload_osc(synth.make_osc(w, 0, 0, w.n_events))
for k in range(5):
    sp, nm = synth.proposal(w, k); h.step(sp, nm); h.llh()
h.block_trace(read=False)
h.set_timing(True); h.kernel_time()
sp, nm = synth.proposal(w, 7); h.step(sp, nm); h.llh()
ms, n = h.kernel_time()
tr = h.block_trace().astype(np.float64)
t0 = tr[:, 0].min()
names = ["start", "tables", "first_stage", "producer_done", "consumers_done", "flushed", "end"]
print(f"kernel {ms/n*1e3:.1f} us (events); grid {tr.shape[0]}")
for i, nme in enumerate(names):
    col = (tr[:, i] - t0) / 1e3
    print(f"{nme:15s} min {col.min():8.2f}  median {np.median(col):8.2f}  max {col.max():8.2f} us")
units = tr[:, 7]
print("units per block: min %d median %d max %d sum %d" % (units.min(), np.median(units), units.max(), units.sum()))
busy = (tr[:, 4] - tr[:, 2]) / 1e3
print("consume span per block: min %.2f median %.2f max %.2f us; per unit median %.3f us" % (busy.min(), np.median(busy), busy.max(), np.median(busy / np.maximum(units, 1))))
