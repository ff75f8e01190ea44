"""The incumbent: MaCh3's own CUDA spline kernels (Splines/gpuSplineUtils.cu built into oracle/_ref, driven like
SMonolith::Evaluate + SynchroniseMemTransfer) on BASELINE config 2, next to the drop-in adapter (same call sequence on
libm3b200) and the fused step.  The reference's GPU path stops at per-event spline weights (n_events x 4 B copied to the
host every step); fill and likelihood then run on the CPU.    python tests/perf/incumbent_gpu.py [n_events]
(lives under tests/ because it drives oracle/: test infrastructure, never the product)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
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


# ---- the whole incumbent step: the reference's MaCh3_CUDA build (its own SMonolith + gpuSplineUtils.cu kernels, weights
#      copied back every step) + its SampleHandlerFD::FillArray_MP and GetLikelihood on the host cores
#      (oracle/_ref/libm3ref_path_lm_refcuda_P<n>.so: the reference's sources compiled where they lie)
from oracle import ref_path_binding as RP              # noqa: E402
if RP.available_refcuda(w.n_params):
    build = f"float_refcuda_P{w.n_params}"
    spl = synth.make_splines(w)
    ev = synth.make_events(w)
    mono = RP.RefSMonolith.from_arrays(w.n_params, w.n_knots, cx, npts, typ, spl, build=build)
    fd = RP.RefSampleHandlerFD(synth.bin_edges(w), w.test_statistic, False, build=build)
    fd.attach_monolith(mono)
    E = w.n_events
    idx = np.arange(E, dtype=np.int32)
    fd.set_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, w_before=idx, w_after=E + idx, n_pool=2 * E)
    pool = np.concatenate([synth.make_osc(w, 0), ev["static_w"]]).astype(np.float64)
    sp, nm = synth.proposal(w, -1)
    fd.reweight(sp, nm, pool)
    fd.set_data(np.random.default_rng(w.seed).poisson(fd.hist()[0]).astype(np.float64))
    ts = []
    for k in range(23):
        sp, nm = synth.proposal(w, k)
        t0 = time.perf_counter()
        fd.reweight(sp, nm, None)
        llh_inc = fd.llh()
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.median(ts[3:]))
    print(f"{'incumbent, whole step (ref. CUDA build)':32s} Reweight + GetLikelihood: {ms * 1e3:8.1f} us per step  "
          f"({E / ms / 1e3:.3f} M events/s, {RP.num_threads(build)} host threads for the fill), -lnL {llh_inc:.6f}")
    # the fused path on the same proposals
    gsh, gd = handlers.build_from_workload(w)
    sp, nm = synth.proposal(w, -1)
    gd["pars"][:] = sp; gd["norm"][:] = nm
    gsh.Reweight(); gsh.GetLikelihood()
    gsh.AddData(np.random.default_rng(w.seed).poisson(gsh.handle.read_hist()[0]).astype(np.float64))
    ts = []
    for k in range(23):
        sp, nm = synth.proposal(w, k)
        gd["pars"][:] = sp; gd["norm"][:] = nm
        t0 = time.perf_counter()
        gsh.Reweight()
        llh_b = gsh.GetLikelihood()
        ts.append(time.perf_counter() - t0)
    msb = 1e3 * float(np.median(ts[3:]))
    print(f"{'libm3b200 fused step':32s} Reweight + GetLikelihood: {msb * 1e3:8.1f} us per step  ({E / msb / 1e6:.3f} G events/s, host wall clock), "
          f"-lnL {llh_b:.6f}  (rel. diff to the incumbent {abs(llh_b - llh_inc) / abs(llh_inc):.2e})")
else:
    print("whole-incumbent library not built for", w.n_params, "parameters")
