#!/usr/bin/env python
"""bench.py -- one reweight + fill + likelihood step of the MaCh3 hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched through torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (the reference's own CPU path, host cores)

Workload at EVERY N (so the driver's 1/2/4/8 points are one strong-scaling curve): BASELINE.json configs[2],
"DUNE-FD-scale" 20M events x (48 TSpline3 K=7 + 12 TF1), 4 samples x 80x20 bins, Poisson -- the configuration the
north_star target is quoted on; it fits one B200 (94 GB resident).  N>1: events sharded contiguously over the ranks,
partial histograms exchanged every step.  Synthetic data (mach3_b200.synth), fresh proposal every step so the active
spline segments change.

Prints ONE JSON line (rank 0):
  value    events/s of the HOST-SYNCHRONISED step -- m3b_step + m3b_llh every step, the way MR2T2 drives it
           (Fitters/MR2T2.cpp:62-74: step k+1 cannot be proposed before step k's -lnL is on the host) -- with all inputs
           resident in HBM (only the <2 KB per-step parameter table crosses PCIe);
  e2e      the same with the oscillation-weight array taken from (pinned) HOST memory inside every step;
  extra.queued   K steps enqueued without a host sync (LLH scans, batches: consecutive launches overlap ramp and tail);
  roofline the fill kernel's isolated launch duration (the library's CUDA events around every launch) against HBM;
  N=1 only: extra.cfg2 / extra.cfg4 / extra.cfg5 = the other BASELINE configs as complete sub-records (clocks, roofline,
           cpu_baseline each), extra.incumbent_gpu = the reference's own MaCh3_CUDA build on the same B200;
  N=1 only: extra.cfg2_e2e_binned_osc / N>1: extra.e2e_binned_osc = the e2e step with a binned oscillator (a 4096-entry
           table of oscillation weights per step instead of one weight per event);
  N>1 only: parity = an untimed check of the sharded -lnL / histogram against the CPU oracle (both exchanges) --
           the run FAILS (rc != 0) above 1e-6; single_process = the same N-GPU step driven by ONE host thread through
           m3b_group_* (the drop-in boundary for the reference's single-process fitters).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CHUNK = 131072          # events generated / uploaded per chunk (multiple of every tile size)
METRIC = "reweighted events/s per host-synchronised MCMC step (reweight+fill+LLH)"
DTYPE = "f32 weights / f64 histogram+LLH"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--events", type=int, default=0, help="override the total event count (debug)")
    ap.add_argument("--exchange", default=os.environ.get("M3B_EXCHANGE", "auto"), choices=["auto", "nccl", "peer"],
                    help="N>1 histogram exchange: the library's own peer-memory pull fused with the likelihood (peer), "
                         "NCCL all-reduce + likelihood launch (nccl), or peer with NCCL as fallback (auto)")
    ap.add_argument("--tile", type=int, default=int(os.environ.get("M3B_TILE", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-events", type=int, default=1_000_000,
                    help="events of the workload the CPU reference is timed on (bounded sample; the full 20M-event monolith "
                         "needs 109 GB of host memory and overflows the reference's 32-bit knot offsets)")
    ap.add_argument("--extras", default="auto", choices=["auto", "all", "none"],
                    help="N=1 sub-records cfg2/cfg4/cfg5/incumbent_gpu, N>1 parity + single-process record (auto: on for the default workload)")
    return ap.parse_args()


def pick_workload(args):
    from mach3_b200 import synth
    name = args.workload
    if name == "auto":
        name = "cfg3"
    w = {"cfg1": synth.CFG1, "cfg2": synth.CFG2, "cfg3": synth.CFG3}[name]
    if args.events:
        w = w.scaled(args.events, name=w.name + f" [events overridden to {args.events}]")
    return w


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md: sample nvidia-smi DURING the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(self.index)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.25)          # let the first samples land before the timed region starts
        except Exception:
            self.p = None
        return self

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline.  kind "reference": the reference's OWN CPU implementation of the path --
# Splines/SplineMonolith.cpp, SplineBase.cpp, BinnedSplineHandler.cpp, Samples/SampleHandlerFD.cpp, SampleHandlerBase.cpp,
# BinningHandler.cpp compiled from /root/reference with its release flags + MULTITHREAD + _LOW_MEMORY_STRUCTS_ into
# oracle/_ref/libm3ref_path_lm_mt.so (oracle/ref_host/Makefile; built in the container that has the reference, it
# travels to the GPU box).  kind "port": the oracle's restatement, when that library is absent.
# ---------------------------------------------------------------------------------------------
def cpu_path(w, n_sample, steps, warmup, budget_s=None):
    """Times SampleHandlerFD::Reweight (FindSplineSegment + CalcSplineWeights + CalcTotalEventWeight
    + FillArray_MP) + GetLikelihood on a bounded sample of workload `w`, DragRace style
    (Fitters/FitterBase.cpp:461-520).  Returns (events/s, ms/step, laps, cores, llh, sample workload, kind)."""
    from mach3_b200 import synth
    from oracle import ref_path_binding as RP     # the checker's reference build, here as the timed CPU baseline
    ws = w if n_sample >= w.n_events else w.scaled(n_sample)
    use_ref = RP.available_mt()
    if use_ref:
        try:
            RP.lib("float_mt")
        except OSError as e:                      # the prebuilt library does not load on this box: time the port
            print(f"bench.py: {e}; falling back to the oracle port for the CPU baseline", file=sys.stderr)
            use_ref = False
    if use_ref:
        kind = "reference"
        typ, npts, cx = synth.param_layout(ws)
        spl, ev = synth.make_splines(ws), synth.make_events(ws)
        mono = RP.RefSMonolith.from_arrays(ws.n_params, ws.n_knots, cx, npts, typ, spl, build="float_mt")
        del spl
        fd = RP.RefSampleHandlerFD(synth.bin_edges(ws), ws.test_statistic, False, build="float_mt")
        fd.attach_monolith(mono)
        E = ws.n_events
        idx = np.arange(E, dtype=np.int32)
        fd.set_events(ev["sample_id"], ev["kin"], ev["norm_idx"] if ws.n_norm_per_event else None, ws.n_norm_per_event,
                      ws.n_norm_params, w_before=idx, w_after=E + idx, n_pool=2 * E)
        pool = np.concatenate([synth.make_osc(ws, 0), ev["static_w"]]).astype(np.float64)
        cores = RP.num_threads("float_mt")

        def step(k, first=False):
            sp, nm = synth.proposal(ws, k)
            t0 = time.perf_counter()
            fd.reweight(sp, nm, pool if first else None)      # the weights the pointers look at change only once here
            llh = fd.llh()
            return time.perf_counter() - t0, llh
        step(-1, first=True)
        fd.set_data(np.random.default_rng(ws.seed).poisson(fd.hist()[0]).astype(np.float64))
    else:
        kind = "port"
        from oracle import binding as O
        O.set_multithread(True)
        mono, sh, d = O.build_from_workload(ws)
        cores = O.num_threads()

        def step(k, first=False):
            sp, nm = synth.proposal(ws, k)
            mono.set_params(sp); sh.norm_vals[:] = nm
            t0 = time.perf_counter()
            sh.Reweight()
            llh = sh.GetLikelihood()
            return time.perf_counter() - t0, llh
        step(-1)
        sh.AddData(np.random.default_rng(ws.seed).poisson(sh.mc).astype(np.float64))
    for k in range(max(warmup, 1)):
        step(k)
    t_all, laps, llh = 0.0, 0, 0.0
    for k in range(steps):
        dt, llh = step(warmup + k)
        t_all += dt
        laps += 1
        if budget_s is not None and t_all > budget_s and laps >= 3:
            break
    ms = 1e3 * t_all / laps
    return ws.n_events / (ms * 1e-3), ms, laps, cores, llh, ws, kind


def cpu_baseline_record(w, n_sample, steps, warmup, budget_s):
    evs, ms, laps, cores, _, ws, kind = cpu_path(w, n_sample, steps, warmup, budget_s=budget_s)
    whole = ws.n_events == w.n_events
    return {"value": evs, "unit": "events/s", "cores": cores, "kind": kind, "ms_per_step": ms,
            "sample": (f"the whole {w.n_events}-event workload" if whole else f"{ws.n_events} events of the {w.n_events}-event workload")
                      + f", {laps} DragRace laps of Reweight+GetLikelihood after {max(warmup, 1)} warm-up, OpenMP {cores} threads"
                      + (" (the reference's own sources, oracle/_ref/libm3ref_path_lm_mt.so)" if kind == "reference" else " (oracle port)")}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses all host threads (before libgomp initialises)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from mach3_b200 import build
    build.build_synth(); build.build_oracle()
    w = pick_workload(args)
    evs, ms, laps, cores, llh, ws, kind = cpu_path(w, args.cpu_sample_events, args.steps, args.warmup, budget_s=150.0)
    what = ("the reference's own SampleHandlerFD::Reweight + GetLikelihood over its SMonolith (compiled from the reference "
            "sources: release flags -O3 -flto, MULTITHREAD, _LOW_MEMORY_STRUCTS_, no -march)" if kind == "reference" else
            "oracle port of the reference's MULTITHREAD CPU path (flags -O3 -fopenmp -flto, no -march)")
    whole = ws.n_events == w.n_events
    sample = (f"the whole {w.n_events}-event workload" if whole else
              f"{ws.n_events} events of the {w.n_events}-event workload (the full monolith needs 109 GB of host memory and "
              f"overflows the reference's unsigned-int knot offsets, Splines/SplineCommon.h:30-50)") + \
             f", {laps} timed steps, OpenMP {cores} threads, {what}"
    line = {"impl": "reference", "metric": METRIC, "value": evs,
            "unit": "events/s", "n_gpus": args.gpus, "steps": laps, "warmup": max(args.warmup, 1), "ms_per_step": ms,
            "llh_evals_per_s": 1e3 / ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": {"workload": w.name, "sample": sample},
            "cpu_baseline": {"value": evs, "unit": "events/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": evs, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# B200 arm: the event-by-event monolith path (configs 1-3)
# ---------------------------------------------------------------------------------------------
class _ChunkBuffers:
    """Two sets of pinned staging arrays for the monolith upload: the generator (host, OpenMP) fills one while the
    library copies the other to the device (true DMA, no pageable bounce) and re-tiles it."""

    def __init__(self, h, w, n_max):
        from mach3_b200 import synth
        self.sets = []
        tc, tl = synth.count_responses(w, 0, min(n_max, w.n_events))
        # dense upper bounds: every event has at most n_cubic / n_linear responses
        tc, tl = max(tc, n_max * w.n_cubic), max(tl, n_max * w.n_linear)
        for _ in range(2):
            self.sets.append(dict(
                nParamPerEvent=h.alloc_host(2 * n_max, np.uint32), paramNo_arr=h.alloc_host(max(tc, 1), np.int16),
                nKnots_arr=h.alloc_host(max(tc, 1), np.uint64), coeff_many=h.alloc_host(max(tc, 1) * w.n_knots * 4, np.float32),
                nParamPerEvent_tf1=h.alloc_host(2 * n_max, np.uint32), paramNo_tf1=h.alloc_host(max(tl, 1), np.int16),
                coeff_tf1=h.alloc_host(max(tl, 1) * 2, np.float32)))

    def free(self, h):
        for s in self.sets:
            for a in s.values():
                h.free_host(a)
        self.sets = []


def upload_monolith(h, w, e0, e1):
    """Generates events [e0, e1) of workload `w` in the reference's monolith layout chunk by chunk and appends them;
    generation of chunk k+1 overlaps the upload of chunk k."""
    from mach3_b200 import synth
    typ, npts, cx = synth.param_layout(w)
    h.splines_begin(w.n_params, w.n_knots, cx, npts, e1 - e0)
    starts = list(range(e0, e1, CHUNK))
    if starts:
        bufs = _ChunkBuffers(h, w, min(CHUNK, e1 - e0))
        with ThreadPoolExecutor(1) as pool:
            fut = pool.submit(synth.make_splines, w, starts[0], min(e1, starts[0] + CHUNK), bufs.sets[0])
            for i, c0 in enumerate(starts):
                spl = fut.result()
                if i + 1 < len(starts):
                    fut = pool.submit(synth.make_splines, w, starts[i + 1], min(e1, starts[i + 1] + CHUNK), bufs.sets[(i + 1) % 2])
                h.splines_append(spl)            # synchronous: the buffer is free again when it returns
        bufs.free(h)
    h.splines_end()


def build_handle(w, e0, e1, local, flags, tile, stream_ptr, n_osc_bufs=4, osc_table=0):
    """osc_table > 0: a binned oscillator -- the events index a table of that many oscillation weights (global event
    number hashed, so every shard sees the same table) and the per-step host input is the table."""
    from mach3_b200 import lib, synth
    h = lib.Handle(device=local, test_statistic=w.test_statistic, update_w2=False, tile_events=tile, flags=flags)
    if stream_ptr is not None:
        h.set_stream(stream_ptr)
    t0 = time.perf_counter()
    upload_monolith(h, w, e0, e1)
    h.upload_binning(synth.bin_edges(w))
    ev = synth.make_events(w, e0, e1)
    if osc_table:
        osc_idx = ((np.arange(e0, e1, dtype=np.int64) * 2654435761) % osc_table).astype(np.int32)
        h.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, osc_idx, osc_table,
                        ev["static_w"])
        del ev, osc_idx
        rng = np.random.default_rng(11)
        osc_bufs = []
        for k in range(n_osc_bufs):
            b = h.alloc_host(osc_table, np.float32)
            b[:] = rng.uniform(0.2, 1.0, osc_table).astype(np.float32)
            osc_bufs.append(b)
        h.upload_osc(osc_bufs[0])
        return h, osc_bufs, time.perf_counter() - t0
    h.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, None, 0,
                    ev["static_w"])
    del ev
    # the caller's persistent oscillation-weight arrays live in pinned + mapped host memory from the library
    # (m3b_alloc_host; registering malloc'ed numpy memory gave less than half the PCIe rate on this pool)
    osc_bufs = []
    for k in range(n_osc_bufs):
        b = h.alloc_host(e1 - e0, np.float32)
        synth.make_osc(w, k, e0, e1, out=b)
        osc_bufs.append(b)
    h.upload_osc(osc_bufs[0])
    return h, osc_bufs, time.perf_counter() - t0


def measure_monolith(args, w, world, rank, local, dist, torch, W, K, want_cpu_baseline, cpu_sample):
    """Builds this rank's shard of workload `w` and measures the step three ways (host-synchronised with resident
    inputs = value; queued; host-synchronised with host inputs = e2e) plus the fill kernel's own duration.
    Returns the rank-0 record (None on other ranks) and keeps nothing alive."""
    from mach3_b200 import lib, sharding, synth
    E = w.n_events
    e0, e1 = sharding.shard_range(E, world, rank)
    n_local = e1 - e0
    flags = lib.FLAG_NO_FUSED_LLH if world > 1 else 0
    stream = torch.cuda.current_stream()
    h, osc_bufs, t_setup = build_handle(w, e0, e1, local, flags, args.tile, stream.cuda_stream)
    n_osc_bufs = len(osc_bufs)
    sh = sharding.ShardedSampleHandler(h, dist, args.exchange, device=f"cuda:{local}") if world > 1 else None

    props = {k: synth.proposal(w, k) for k in range(-1, 3 * (W + K) + 16)}
    props = {k: (np.ascontiguousarray(sp, np.float64), np.ascontiguousarray(nm, np.float64)) for k, (sp, nm) in props.items()}
    prop_addr = {k: (lib.addr(sp), lib.addr(nm) if nm.size else 0) for k, (sp, nm) in props.items()}
    osc_addr = {id(b): lib.addr(b) for b in osc_bufs}

    def step(k, osc=None):
        if world == 1:
            # raw addresses, like the C++ host the library is made for (no per-call ctypes pointer extraction)
            h.step_addr(prop_addr[k][0], prop_addr[k][1], 0 if osc is None else osc_addr[id(osc)])
        else:
            sp, nm = props[k]
            sh.Reweight(sp, nm, osc)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # Asimov data at nominal, Poisson-fluctuated with a seed every rank shares
    step(-1); h.llh()
    mc, _ = h.read_hist()
    data = np.random.default_rng(w.seed).poisson(mc).astype(np.float64)
    h.upload_data(data)

    def timed(fn, n):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        t0 = time.perf_counter()
        out = fn(n)
        ev1.record(stream)
        barrier()
        t_host = time.perf_counter() - t0
        return max(ev0.elapsed_time(ev1), 1e3 * t_host), out

    # ---------------- value: host-synchronised steps, inputs resident in HBM --------------------
    for k in range(W):
        step(k); h.llh_fast()
    llh_w = h.llh()
    clocks = ClockSampler(local).start()
    launches0 = h.info().kernel_launches

    def sync_steps(n):
        v = 0.0
        for k in range(W, W + n):
            step(k)
            v = h.llh_fast()
        return v
    ms_sync, llh_sync = timed(sync_steps, K)
    launches = h.info().kernel_launches - launches0

    # ---------------- queued: K steps enqueued without a host sync ------------------------------
    def queued_steps(n):
        for k in range(W, W + n):
            step(k)
        return None
    ms_queued, _ = timed(queued_steps, K)
    llh_queued = h.llh()
    assert llh_queued == llh_sync or abs(llh_queued - llh_sync) <= 1e-9 * abs(llh_sync), (llh_sync, llh_queued)

    # ---------------- the fill kernel alone (library's events around every launch) --------------
    h.set_timing(True)
    h.kernel_time()
    for k in range(W, W + K):
        step(k)
    h.llh()
    kern_ms, kern_n = h.kernel_time()
    h.set_timing(False)
    info_value = h.info()          # launch configuration of the value / roofline passes (the e2e pass adds staging slots)

    # ---------------- e2e: host buffers in, scalar out, every step ------------------------------
    for k in range(min(W, 5)):
        step(W + K + k, osc_bufs[k % n_osc_bufs]); h.llh()

    def e2e_steps(n):
        v = 0.0
        for k in range(n):
            step(W + K + 5 + k, osc_bufs[k % n_osc_bufs])
            v = h.llh_fast()
        return v
    ms_e2e, llh_e2e = timed(e2e_steps, K)
    clk = clocks.stop()

    info = h.info()
    kern_avg = kern_ms / max(kern_n, 1)
    if dist is not None:
        t = torch.tensor([ms_sync, ms_queued, ms_e2e, kern_avg], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_sync, ms_queued, ms_e2e, kern_avg = t.tolist()
    exchange = "none" if world == 1 else sh.exchange
    h.close()
    del h, sh, osc_bufs
    if rank != 0:
        return None

    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ms_step = ms_sync / K
    alg_bytes_local = n_local * w.bytes_per_event        # SURVEY §8d per-event figure x events of one launch
    achieved = alg_bytes_local / (kern_avg * 1e-3) / 1e9
    step_bytes = 12 * w.n_params + 4 * w.n_norm_params
    traffic, traffic_src = committed_traffic(w, world, info_value)
    rec = {
        "metric": METRIC,
        "value": E / (ms_step * 1e-3), "unit": "events/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "llh_evals_per_s": 1e3 / ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": w.name, "events": E, "events_per_gpu": n_local, "responses_per_event": w.n_params,
                   "bins": w.n_bins, "tile_events": info_value.tile_events, "grid_blocks": info_value.grid_blocks,
                   "smem_bytes": info_value.smem_bytes, "tma_stages": info_value.tma_stages, "exchange": exchange,
                   "step": "m3b_step (async) + m3b_llh (blocks) every step; inputs resident in HBM",
                   "l2": "inputs larger than L2: %.0f MB of coefficient rows stream per step per GPU, fresh "
                         "proposal (different segments) every step" % (info.active_bytes_per_step / 1e6),
                   "device_bytes": info.device_bytes, "setup_s": round(t_setup, 2)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth, read+write)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                     "unit": "GB/s", "frac": achieved / peak, "frac_of_8TBs_nominal": achieved / 8000.0,
                     "kernel": "m3b::fill_tma_kernel", "kernel_ms": kern_avg, "algorithmic_bytes_per_launch": alg_bytes_local,
                     "loaded_bytes_per_launch": info.active_bytes_per_step, "traffic": traffic, "traffic_source": traffic_src,
                     "note": "kernel_ms = isolated launch duration (library's CUDA events around every launch); a read-only "
                             "stream can exceed the copy (read+write) peak",
                     "frac_of_step": kern_avg / ms_step},
        "e2e": {"value": E / (ms_e2e / K * 1e-3), "unit": "events/s", "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": int(4 * n_local + step_bytes), "d2h_bytes_per_step": int(8 * (1 + w.n_samples)),
                "smem_bytes": info.smem_bytes, "tma_stages": info.tma_stages,
                "api": "m3b_step(host pars, host norms, host osc weights in pinned memory) + m3b_llh(); the osc weights "
                       "are streamed over PCIe by the fill kernel's own bulk copies (no separate H2D pass)"},
        "gpu_launches": int(launches), "clocks": clk,
        "extra": {"queued": {"value": E / (ms_queued / K * 1e-3), "unit": "events/s", "ms_per_step": ms_queued / K,
                             "what": "K steps enqueued back to back, ONE host synchronisation at the end (LLH scans, "
                                     "m3b_step_batch's sequential path): consecutive fused launches overlap ramp and tail "
                                     "through programmatic dependent launch",
                             "achieved_gbs": alg_bytes_local / (ms_queued / K * 1e-3) / 1e9}},
        "llh": {"last_value_step": llh_sync, "last_e2e_step": llh_e2e, "after_warmup": llh_w},
    }
    if want_cpu_baseline:
        rec["cpu_baseline"] = cpu_baseline_record(w, cpu_sample, 100, 2, budget_s=15.0)
    return rec


def committed_traffic(w, world, info):
    """DRAM bytes per launch from the committed `ncu --set full` capture whose launch configuration is the one this run
    used (profiles/r02_ncu_*.json carry tile_events / tma_stages / events); None when no capture matches."""
    pdir = os.path.join(ROOT, "profiles")
    try:
        names = sorted(f for f in os.listdir(pdir) if f.startswith("r02_ncu_full_fill_tma") and f.endswith(".json"))
    except OSError:
        return None, None
    for f in names:
        try:
            pj = json.load(open(os.path.join(pdir, f)))
            c = pj.get("config", {})
            if c.get("events_per_gpu") != info.n_events or c.get("tile_events") != info.tile_events or \
                    c.get("tma_stages") != info.tma_stages or c.get("responses_per_event") != w.n_params:
                continue
            return float(pj["dram_bytes_per_launch"]), f"profiles/{f} (dram__bytes_read.sum + dram__bytes_write.sum per launch, same launch configuration)"
        except Exception:
            continue
    return None, None


# ---------------------------------------------------------------------------------------------
# N>1: untimed parity of the sharded path against the CPU oracle (the checker), both exchanges
# ---------------------------------------------------------------------------------------------
def sharded_parity(args, world, rank, local, dist, torch, n_events=160_001):
    sys.path.insert(0, os.path.join(ROOT, "tests", "multigpu"))
    import parity_ranks
    return parity_ranks.run_parity(dist, rank, world, local, n_events, verbose=False)


# ---------------------------------------------------------------------------------------------
# N>1: the same N-GPU step from ONE process / ONE host thread (m3b_group_*), rank 0 after the ranks have left
# ---------------------------------------------------------------------------------------------
def measure_single_process(args, w, n_dev, torch, W, K, expect_llh):
    from mach3_b200 import lib, synth
    t0 = time.perf_counter()
    g = lib.Group(list(range(n_dev)), test_statistic=w.test_statistic, update_w2=False, tile_events=args.tile)
    osc_bufs = []
    for i in range(n_dev):
        e0, e1 = g.shard(w.n_events, i)
        m = g.member(i)
        upload_monolith(m, w, e0, e1)
        m.upload_binning(synth.bin_edges(w))
        ev = synth.make_events(w, e0, e1)
        m.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, None, 0, ev["static_w"])
        del ev
    g.connect()
    # per-step oscillation weights: one pinned host array over ALL events (what the oscillator owns), sharded by the group
    for k in range(4):
        b = g.alloc_host(w.n_events, np.float32)
        synth.make_osc(w, k, 0, w.n_events, out=b)
        osc_bufs.append(b)
    g.upload_osc(osc_bufs[0])
    t_setup = time.perf_counter() - t0
    props = {k: synth.proposal(w, k) for k in range(-1, 3 * (W + K) + 16)}
    props = {k: (np.ascontiguousarray(sp, np.float64), np.ascontiguousarray(nm, np.float64)) for k, (sp, nm) in props.items()}
    g.step(*props[-1]); g.llh()
    mc, _ = g.read_hist()
    g.upload_data(np.random.default_rng(w.seed).poisson(mc).astype(np.float64))
    for k in range(W):
        g.step(*props[k]); g.llh()
    t1 = time.perf_counter()
    for k in range(W, W + K):
        g.step(*props[k]); llh = g.llh()
    ms_sync = 1e3 * (time.perf_counter() - t1) / K
    for k in range(min(W, 5)):
        g.step(*props[W + K + k], osc_w=osc_bufs[k % 4]); g.llh()
    t1 = time.perf_counter()
    for k in range(K):
        g.step(*props[W + K + 5 + k], osc_w=osc_bufs[k % 4]); llh_e = g.llh()
    ms_e2e = 1e3 * (time.perf_counter() - t1) / K
    rec = {"value": w.n_events / (ms_sync * 1e-3), "unit": "events/s", "ms_per_step": ms_sync, "n_gpus": n_dev,
           "e2e": {"value": w.n_events / (ms_e2e * 1e-3), "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(4 * w.n_events),
                   "d2h_bytes_per_step": int(8 * (1 + w.n_samples))},
           "timing": "host wall clock around K x (m3b_group_step + m3b_group_llh), one process, one calling thread",
           "exchange": g.exchange, "setup_s": round(t_setup, 1), "llh_last": llh,
           "llh_rel_diff_to_multi_process": (abs(llh - expect_llh) / abs(expect_llh)) if expect_llh else None}
    g.close()
    return rec


# ---------------------------------------------------------------------------------------------
# main, B200 arm
# ---------------------------------------------------------------------------------------------
def main_b200(args):
    # torchrun exports OMP_NUM_THREADS=1; the synthetic-workload generator (host, OpenMP) would then build each
    # rank's shard on one core.  Give every rank its share of the host cores BEFORE the generator library loads.
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if world_env > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world_env))
    import torch
    from mach3_b200 import lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun for --gpus > 1 (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    numa = lib.bind_to_gpu_cpus(local)     # pinned host buffers (osc weights, -lnL mirror) land on the GPU's socket
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w = pick_workload(args)
    W, K = max(args.warmup, 3), args.steps
    extras = args.extras == "all" or (args.extras == "auto" and args.workload == "auto" and not args.events)
    line = measure_monolith(args, w, world, rank, local, dist, torch, W, K,
                            want_cpu_baseline=(world == 1 and not args.no_cpu_baseline), cpu_sample=args.cpu_sample_events)
    if rank == 0:
        line["config"]["host_cpu_affinity"] = numa
    rc = 0
    if world > 1 and extras:
        par = sharded_parity(args, world, rank, local, dist, torch)
        if rank == 0:
            line["parity"] = par
            if not par["ok"]:
                rc = 3
        try:      # every rank takes part; a failure on any rank must not leave the others in a collective
            bo = measure_sharded_binned_osc(args, w, world, rank, local, dist, torch, W, K)
        except Exception as e:                                   # noqa: BLE001
            bo = {"error": f"{type(e).__name__}: {e}"}
        if rank == 0:
            line["extra"]["e2e_binned_osc"] = bo
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return 0
        # this process no longer shares the GPUs: the host threads may use every core again
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))
        except Exception:
            pass
    if world > 1 and extras:
        try:
            line["single_process"] = measure_single_process(args, w, world, torch, W, K, line["llh"]["last_value_step"])
        except Exception as e:                                   # noqa: BLE001 -- the headline line must still be printed
            line["single_process"] = {"error": f"{type(e).__name__}: {e}"}
    if world == 1 and extras:
        ex = line["extra"]
        for name, fn in (("cfg2", lambda: measure_monolith(args, synth.CFG2, 1, 0, local, None, torch, W, max(K, 100),
                                                           want_cpu_baseline=not args.no_cpu_baseline, cpu_sample=synth.CFG2.n_events)),
                         ("cfg2_e2e_binned_osc", lambda: measure_binned_osc(args, local, W, max(K, 100))),
                         ("cfg4", lambda: measure_cfg4(args, local, not args.no_cpu_baseline)),
                         ("cfg5", lambda: measure_cfg5(args, local, not args.no_cpu_baseline)),
                         ("incumbent_gpu", lambda: measure_incumbent(local))):
            try:
                ex[name] = fn()
            except Exception as e:                               # noqa: BLE001
                ex[name] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line), flush=True)
    return rc


def measure_sharded_binned_osc(args, w, world, rank, local, dist, torch, W, K):
    """N > 1: the e2e step with a binned oscillator (a 4096-entry table per step instead of one weight per event), same
    sharding and exchange as the headline line; max over ranks.  Every rank calls it."""
    from mach3_b200 import lib, sharding, synth
    e0, e1 = sharding.shard_range(w.n_events, world, rank)
    stream = torch.cuda.current_stream()
    h, tabs, t_setup = build_handle(w, e0, e1, local, lib.FLAG_NO_FUSED_LLH, args.tile, stream.cuda_stream, osc_table=4096)
    sh = sharding.ShardedSampleHandler(h, dist, args.exchange, device=f"cuda:{local}")
    props = [synth.proposal(w, k) for k in range(-1, W + K)]
    props = [(np.ascontiguousarray(sp, np.float64), np.ascontiguousarray(nm, np.float64)) for sp, nm in props]
    sh.Reweight(*props[0], tabs[0]); h.llh()
    mc, _ = h.read_hist()
    h.upload_data(np.random.default_rng(w.seed).poisson(mc).astype(np.float64))
    for k in range(W):
        sh.Reweight(*props[1 + k], tabs[k % 4]); h.llh_fast()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t0 = time.perf_counter()
    for k in range(K):
        sh.Reweight(*props[1 + W + k], tabs[k % 4]); llh = h.llh_fast()
    ev1.record(stream)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ms = max(ev0.elapsed_time(ev1), 1e3 * (time.perf_counter() - t0))
    t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    exchange = sh.exchange
    h.close()
    del h, sh, tabs
    return {"value": w.n_events / (ms * 1e-3), "unit": "events/s", "ms_per_step": ms, "osc_table_entries": 4096, "exchange": exchange,
            "h2d_bytes_per_step": int(4 * 4096 + 12 * w.n_params + 4 * w.n_norm_params), "d2h_bytes_per_step": int(8 * (1 + w.n_samples)),
            "llh_last": float(llh), "setup_s": round(t_setup, 1),
            "what": "host-synchronised step, oscillation weights as a 4096-entry table in host memory every step (binned oscillator)"}


def measure_binned_osc(args, local, W, K):
    """e2e with a BINNED oscillator (NuOscillator's binned mode: events index a small table of oscillation weights,
    SampleHandlerFD's osc pointers then point into that table): config 2, the per-step host input is the table, not one
    weight per event.  Host-synchronised m3b_step(host pars, host norms, host table) + m3b_llh per step."""
    from mach3_b200 import lib, synth
    w = synth.CFG2
    n_osc = 4096
    h = lib.Handle(device=local, test_statistic=w.test_statistic, update_w2=False, tile_events=args.tile)
    upload_monolith(h, w, 0, w.n_events)
    h.upload_binning(synth.bin_edges(w))
    ev = synth.make_events(w, 0, w.n_events)
    osc_idx = ((np.arange(w.n_events, dtype=np.int64) * 2654435761) % n_osc).astype(np.int32)
    h.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, osc_idx, n_osc,
                    ev["static_w"])
    del ev
    rng = np.random.default_rng(11)
    tabs = []
    for _ in range(4):
        t = h.alloc_host(n_osc, np.float32)
        t[:] = rng.uniform(0.2, 1.0, n_osc).astype(np.float32)
        tabs.append(t)
    props = [synth.proposal(w, k) for k in range(W + K + 1)]
    h.step(*props[0], tabs[0]); h.llh()
    h.upload_data(np.random.default_rng(w.seed).poisson(h.read_hist()[0]).astype(np.float64))
    for k in range(W):
        h.step(*props[k], tabs[k % 4]); h.llh_fast()
    h.synchronize()
    t0 = time.perf_counter()
    for k in range(K):
        h.step(*props[W + k], tabs[k % 4]); llh = h.llh_fast()
    dt = (time.perf_counter() - t0) / K
    for t in tabs:
        h.free_host(t)
    h.close()
    return {"workload": w.name, "osc_table_entries": n_osc, "ms_per_step": 1e3 * dt, "value": w.n_events / dt, "unit": "events/s",
            "h2d_bytes_per_step": int(4 * n_osc + 12 * w.n_params + 4 * w.n_norm_params), "d2h_bytes_per_step": 16, "llh_last": float(llh),
            "what": "host-synchronised step with the oscillation weights as a 4096-entry table in host memory (binned "
                    "oscillator): the per-step host input shrinks from 4 B per event to the table"}


# ---------------------------------------------------------------------------------------------
# BASELINE config 4: BinnedSplineHandler workload
# ---------------------------------------------------------------------------------------------
def cpu_binned(w, steps, warmup, budget_s):
    """The reference's own BinnedSplineHandler::Evaluate + SampleHandlerFD::FillArray_MP + GetLikelihood (release build,
    MULTITHREAD, float) on workload `w` (a bounded sample of config 4)."""
    from mach3_b200.synth import binned as B
    from oracle import ref_path_binding as RP
    if not RP.available_mt():
        return None
    spl, ev = B.make_binned_splines(w), B.make_binned_events(w)
    E = w.n_events
    fd = RP.RefSampleHandlerFD(B.bin_edges(w), w.test_statistic, True, build="float_mt")
    fd.attach_binned(spl)
    idx = np.arange(E, dtype=np.int32)
    fd.set_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, w_before=idx,
                  w_after=E + idx, n_pool=2 * E, binned_n_per_event=ev["n_per_event"], binned_slot=ev["spline_index"])
    pool = np.concatenate([B.make_osc(w, 0), ev["static_w"]]).astype(np.float64)
    sp, nm = B.proposal(w, -1)
    fd.reweight(sp, nm, pool)
    fd.set_data(np.random.default_rng(w.seed).poisson(fd.hist()[0]).astype(np.float64))
    t_all, laps = 0.0, 0
    for k in range(warmup + steps):
        sp, nm = B.proposal(w, k)
        t0 = time.perf_counter()
        fd.reweight(sp, nm, None); fd.llh()
        dt = time.perf_counter() - t0
        if k >= warmup:
            t_all += dt; laps += 1
            if t_all > budget_s and laps >= 3:
                break
    cores = RP.num_threads("float_mt")
    fd.close()
    ms = 1e3 * t_all / laps
    return {"value": E / (ms * 1e-3), "unit": "events/s", "cores": cores, "kind": "reference", "ms_per_step": ms,
            "binned_spline_evals_per_s": int(spl["uniquecoeffindices"].size) / (ms * 1e-3),
            "sample": f"{E} events x {w.n_systs} systematics x {w.n_grid} spline bins ({int(spl['uniquecoeffindices'].size)} non-flat "
                      f"binned splines): a 1/10 sample of config 4, {laps} laps of the reference's BinnedSplineHandler::Evaluate + "
                      f"FillArray_MP + GetLikelihood (oracle/_ref/libm3ref_path_lm_mt.so), OpenMP {cores} threads"}


def measure_cfg4(args, local=0, want_cpu=True):
    import torch
    from mach3_b200 import handlers
    from mach3_b200.synth import binned as B
    w = B.CFG4 if not args.events or args.workload != "cfg4" else B.CFG4.scaled(n_events=args.events, n_grid=max(1000, int(B.CFG4.n_grid * args.events / B.CFG4.n_events)))
    t0 = time.perf_counter()
    sh, d = handlers.build_binned_from_workload(w, update_w2=True, device=local)
    h = sh.handle
    t_setup = time.perf_counter() - t0
    W, K = max(args.warmup, 3), max(args.steps, 20)
    props = {k: B.proposal(w, k) for k in range(-1, 2 * (W + K) + 2)}

    def step(k):
        d["pars"][:], d["norm"][:] = props[k]
        sh.Reweight()

    step(-1); sh.GetLikelihood()
    sh.AddData(np.random.default_rng(w.seed).poisson(sh.GetMCArray()).astype(np.float64))
    for k in range(W):
        step(k); sh.GetLikelihood()
    clocks = ClockSampler(local).start()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for k in range(W, W + K):
        step(k); llh = sh.GetLikelihood()
    t_sync = time.perf_counter() - t1
    t1 = time.perf_counter()
    for k in range(W, W + K):
        step(k)
    sh.GetLikelihood()
    t_queued = time.perf_counter() - t1
    h.set_timing(True); h.kernel_time()
    for k in range(W, W + K):
        step(k)
    sh.GetLikelihood()
    kern_ms, kern_n = h.kernel_time()
    h.set_timing(False)
    clk = clocks.stop()
    n_act = int(d["spl"]["uniquecoeffindices"].size)
    n_ptr = int(d["ev"]["spline_index"].size)
    mask = np.zeros(w.n_slots, bool); mask[d["spl"]["uniquecoeffindices"]] = True
    n_nonflat = int(mask[d["ev"]["spline_index"]].sum())              # pointers at flat splines (exactly 1.0) are dropped at upload
    del mask
    # eval: {y,b,c,d}+x read, weight write per non-flat spline; fill: event table + (index + gathered weight) per non-flat pointer
    alg = n_act * (16 + 4 + 4) + w.n_events * 8 + n_nonflat * 4 * 2
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kms = kern_ms / max(kern_n, 1)
    info = h.info()
    traffic, traffic_src = None, None
    try:      # DRAM bytes of the two launches from the committed ncu captures of this same workload
        pj = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_full_binned_fill_cfg4.json")))
        if pj.get("config", {}).get("events") == w.n_events:
            traffic = float(pj["dram_bytes_per_launch"]) + float(pj["eval_kernel"]["dram_bytes_per_launch"])
            traffic_src = "profiles/r02_ncu_full_binned_fill_cfg4.json (dram__bytes_read.sum + dram__bytes_write.sum: fill launch + eval launch)"
    except Exception:
        pass
    rec = {"metric": "reweighted events/s per host-synchronised MCMC step (binned-spline eval + fill + Barlow-Beeston LLH)",
           "value": w.n_events / (t_sync / K), "unit": "events/s", "n_gpus": 1, "steps": K, "warmup": W,
           "ms_per_step": 1e3 * t_sync / K, "binned_spline_evals_per_s": n_act / (t_sync / K), "higher_is_better": True,
           "dtype": DTYPE, "data": "synthetic",
           "config": {"workload": w.name, "events": w.n_events, "active_binned_splines": n_act, "weight_pointers": n_ptr, "non_flat_weight_pointers": n_nonflat,
                      "slots": w.n_slots, "bins": w.n_bins, "setup_s": round(t_setup, 1),
                      "l2": "coefficient rows (%.0f MB/step) stream from HBM; the compact weight array (%.0f MB) is gathered through L2"
                            % (n_act * 20 / 1e6, n_act * 4 / 1e6)},
           "roofline": {"bound": "hbm", "achieved": alg / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (kms * 1e-3) / 1e9 / peak, "kernel": "m3b::binned_eval_kernel + m3b::binned_fill_kernel",
                        "kernel_ms": kms, "algorithmic_bytes_per_launch": alg, "traffic": traffic, "traffic_source": traffic_src,
                        "note": "kernel_ms = eval + fill launches of one step, from the library's CUDA events"},
           "e2e": {"value": w.n_events / (t_sync / K), "unit": "events/s", "ms_per_step": 1e3 * t_sync / K,
                   "h2d_bytes_per_step": 12 * w.n_systs + 4 * w.n_norm_params, "d2h_bytes_per_step": 16,
                   "note": "config 4 has no per-step host array: the step's only inputs are the parameter values"},
           "extra": {"queued": {"ms_per_step": 1e3 * t_queued / K}},
           "gpu_launches": int(2 * K), "clocks": clk, "llh": {"last": llh}, "device_bytes": info.device_bytes}
    h.close()
    del sh, d
    if want_cpu:
        rec["cpu_baseline"] = cpu_binned(B.CFG4.scaled(n_events=w.n_events // 10, n_grid=w.n_grid // 10), 20, 2, 10.0)
    return rec


# ---------------------------------------------------------------------------------------------
# BASELINE config 5: batched proposals
# ---------------------------------------------------------------------------------------------
def measure_cfg5(args, local=0, want_cpu=True):
    import torch
    from mach3_b200 import lib, synth
    w = synth.CFG5 if not args.events or args.workload != "cfg5" else synth.CFG5.scaled(args.events)
    n_sets = 256
    h = lib.Handle(device=local, test_statistic=w.test_statistic, update_w2=False, tile_events=args.tile)
    t0 = time.perf_counter()
    upload_monolith(h, w, 0, w.n_events)
    h.upload_binning(synth.bin_edges(w))
    ev = synth.make_events(w, 0, w.n_events)
    h.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, None, 0, ev["static_w"])
    del ev
    h.upload_osc(synth.make_osc(w, 0, 0, w.n_events))
    t_setup = time.perf_counter() - t0
    sp, nm = synth.proposal(w, -1)
    h.step(sp, nm); h.llh()
    h.upload_data(np.random.default_rng(w.seed).poisson(h.read_hist()[0]).astype(np.float64))
    # single-set reference point on the same handle
    h.set_timing(True); h.kernel_time()
    for k in range(10):
        sp, nm = synth.proposal(w, k); h.step(sp, nm)
    h.llh()
    ms1, n1 = h.kernel_time()
    rng = np.random.default_rng(w.seed + 7)

    def batch(k):
        sp0, nm0 = synth.proposal(w, k)
        sps = np.clip(sp0[None, :] + rng.normal(0, 0.3, (n_sets, w.n_params)), -2.9, 2.9)
        nms = np.clip(nm0[None, :] + rng.normal(0, 0.05, (n_sets, w.n_norm_params)), 0.5, 1.5)
        return sps, nms

    W, K = 2, 4
    bs = [batch(k) for k in range(W + K)]
    for k in range(W):
        h.step_batch(*bs[k])
    h.kernel_time()
    clocks = ClockSampler(local).start()
    t1 = time.perf_counter()
    for k in range(W, W + K):
        tot = h.step_batch(*bs[k])
    t_b = (time.perf_counter() - t1) / K
    msb, nb = h.kernel_time()
    kms = msb / max(nb, 1)
    # the first-generation kernel on the same handle and batches (A/B)
    h.set_flags(set_mask=lib.FLAG_BATCH_KERNEL_V1)
    h.step_batch(*bs[0]); h.kernel_time()
    for k in range(W, W + 2):
        h.step_batch(*bs[k])
    ms_v1, n_v1 = h.kernel_time()
    h.set_flags(clear_mask=lib.FLAG_BATCH_KERNEL_V1)
    # LLH scan of one parameter (FitterBase::RunLLHScan): 256 points, every other parameter fixed -> one segment per slot
    sp0, nm0 = synth.proposal(w, 1)
    sc_sp = np.tile(sp0, (n_sets, 1)); sc_sp[:, 0] = np.linspace(-2.9, 2.9, n_sets)
    sc_nm = np.tile(nm0, (n_sets, 1))
    h.step_batch(sc_sp, sc_nm); h.kernel_time()
    t2 = time.perf_counter()
    for k in range(K):
        h.step_batch(sc_sp, sc_nm)
    t_scan = (time.perf_counter() - t2) / K
    ms_scan, n_scan = h.kernel_time()
    clk = clocks.stop()
    fp_instr = w.n_events * n_sets * (4 * (w.n_params - w.n_linear) + 2 * w.n_linear)     # FMA/MUL issue slots
    sm_ghz = (clk.get("sm_max_mhz") or 1965.0) * 1e-3
    peak_issue = 148 * 128 * sm_ghz * 1e9
    rec = {"metric": "LLH evaluations/s over batched proposals (reweight+fill+LLH per parameter set)", "value": n_sets / t_b,
           "unit": "LLH evals/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": 1e3 * t_b, "us_per_llh": 1e6 * t_b / n_sets,
           "event_sets_per_s": w.n_events * n_sets / t_b, "higher_is_better": True, "dtype": DTYPE, "data": "synthetic",
           "config": {"workload": w.name, "events": w.n_events, "sets_per_batch": n_sets, "bins": w.n_bins, "setup_s": round(t_setup, 1),
                      "proposals": "256 perturbations N(theta0, 0.3^2) around one point: 1-3 active segments per parameter",
                      "single_set_kernel_ms": ms1 / max(n1, 1), "amortisation_vs_single_set": (ms1 / max(n1, 1)) * n_sets / kms,
                      "first_generation_kernel_ms": ms_v1 / max(n_v1, 1),
                      "llh_scan_256_points": {"ms": 1e3 * t_scan, "us_per_llh": 1e6 * t_scan / n_sets, "kernel_ms": ms_scan / max(n_scan, 1),
                                              "amortisation_vs_single_set": (ms1 / max(n1, 1)) * n_sets / (ms_scan / max(n_scan, 1))}},
           "roofline": {"bound": "fp32 issue (CUDA cores; no contraction, so no tensor cores)", "achieved": fp_instr / (kms * 1e-3) / 1e12,
                        "peak": peak_issue / 1e12, "peak_source": "148 SMs x 128 FP32 lanes x max SM clock", "unit": "T FP32 instr/s",
                        "frac": fp_instr / (kms * 1e-3) / peak_issue, "kernel": "m3b::fill_batch2_kernel", "kernel_ms": kms, "traffic": None,
                        "algorithmic_instr_per_launch": fp_instr},
           "e2e": {"value": n_sets / t_b, "unit": "LLH evals/s", "h2d_bytes_per_step": int(n_sets * (8 * w.n_params + 8 * w.n_norm_params)),
                   "d2h_bytes_per_step": int(n_sets * 8 * (1 + w.n_samples))},
           "gpu_launches": int(2 * K), "clocks": clk, "llh": {"first": float(tot[0]), "last": float(tot[-1])}}
    h.close()
    if want_cpu:
        evs, ms, laps, cores, _, ws, kind = cpu_path(synth.CFG5, 500_000, 20, 2, budget_s=8.0)
        rec["cpu_baseline"] = {"value": 1e3 / (ms * w.n_events / ws.n_events), "unit": "LLH evals/s", "cores": cores, "kind": kind,
                               "sample": f"{ws.n_events} events of the workload, {laps} sequential Reweight+GetLikelihood laps (the reference has no "
                                         f"batched path: one parameter set per step), scaled to the full {w.n_events} events; OpenMP {cores} threads"}
    return rec


# ---------------------------------------------------------------------------------------------
# the incumbent GPU path: the reference's own MaCh3_CUDA build (SMonolith + Splines/gpuSplineUtils.cu kernels compiled
# from the reference's sources, weights copied back every step, FillArray_MP + GetLikelihood on the host cores)
# ---------------------------------------------------------------------------------------------
class _QuietStdout:
    """The reference's CUDA code printf()s its allocations: keep the C-level stdout of this process clean (bench.py's
    contract is ONE JSON line on stdout)."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)
        return self

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved); os.close(self.null)
        return False


def measure_incumbent(local=0):
    with _QuietStdout():
        return _measure_incumbent(local)


def _measure_incumbent(local=0):
    from mach3_b200 import synth
    from oracle import ref_path_binding as RP
    w = synth.CFG2
    if not RP.available_refcuda(w.n_params):
        return {"unavailable": f"oracle/_ref/libm3ref_path_lm_refcuda_P{w.n_params}.so not built (needs /root/reference at build time)"}
    build = f"float_refcuda_P{w.n_params}"
    typ, npts, cx = synth.param_layout(w)
    spl, ev = synth.make_splines(w), synth.make_events(w)
    mono = RP.RefSMonolith.from_arrays(w.n_params, w.n_knots, cx, npts, typ, spl, build=build)
    del spl
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
    for k in range(25):
        sp, nm = synth.proposal(w, k)
        t0 = time.perf_counter()
        fd.reweight(sp, nm, None)
        llh = fd.llh()
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(ts[5:]))
    cores = RP.num_threads(build)
    fd.close()
    return {"value": E / (ms * 1e-3), "unit": "events/s", "ms_per_step": ms, "workload": w.name, "host_threads": cores, "llh_last": llh,
            "what": "the reference's MaCh3_CUDA build on this B200: its SMonolith + gpuSplineUtils.cu kernels (per-event spline weights, "
                    "4 B/event copied back), then SampleHandlerFD::FillArray_MP + GetLikelihood on the host cores; compiled from the "
                    "reference's own sources (oracle/ref_host/Makefile); 20 timed steps after 5 warm-up; compare with extra.cfg2"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    elif a.workload in ("cfg4", "cfg5"):
        fn = measure_cfg4 if a.workload == "cfg4" else measure_cfg5
        print(json.dumps(fn(a, int(os.environ.get("LOCAL_RANK", "0")), not a.no_cpu_baseline)), flush=True)
    else:
        sys.exit(main_b200(a))
