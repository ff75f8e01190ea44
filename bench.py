#!/usr/bin/env python
"""bench.py -- one reweight + fill + likelihood step of the MaCh3 hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched through torchrun)
    python bench.py --impl reference ...                      (the reference's CPU path, host cores)

Workload (BASELINE.json): N=1 -> configs[1] "T2K-FD-like" 1M events x (40 TSpline3 K=7 + 10 TF1),
60x15 bins, Poisson.  N>1 -> configs[2] "DUNE-FD-scale" 20M events x (48+12), 4 samples x 80x20 bins,
events sharded contiguously over the ranks, partial histograms exchanged every step.
Synthetic data (mach3_b200.synth), fresh proposal every step so the active spline segments change.

Prints ONE JSON line (rank 0).  `value` = events/s with all inputs resident in HBM (only the
<2 KB per-step parameter table crosses PCIe); `e2e` = the same through the C ABI with the
oscillation-weight array copied from host memory inside every step and the -lnL read back.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CHUNK = 131072          # events generated / uploaded per chunk (multiple of every tile size)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--events", type=int, default=0, help="override the total event count (debug)")
    ap.add_argument("--exchange", default=os.environ.get("M3B_EXCHANGE", "auto"), choices=["auto", "nccl", "peer"],
                    help="N>1 histogram exchange: the library's own peer-memory pull fused with the likelihood (peer), "
                         "NCCL all-reduce + likelihood launch (nccl), or peer with NCCL as fallback (auto)")
    ap.add_argument("--tile", type=int, default=int(os.environ.get("M3B_TILE", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-events", type=int, default=200_000)
    return ap.parse_args()


def pick_workload(args):
    from mach3_b200 import synth
    name = args.workload
    if name == "auto":
        name = "cfg2" if args.gpus == 1 else "cfg3"
    w = {"cfg1": synth.CFG1, "cfg2": synth.CFG2, "cfg3": synth.CFG3}[name]
    if args.events:
        w = w.scaled(args.events, name=w.name + f" [events overridden to {args.events}]")
    return w


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
                                       "-lms", "100", "-i", str(self.index)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

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
# Splines/SplineMonolith.cpp, SplineBase.cpp, Samples/SampleHandlerFD.cpp, SampleHandlerBase.cpp, BinningHandler.cpp
# compiled from /root/reference with its release flags + MULTITHREAD + _LOW_MEMORY_STRUCTS_ into
# oracle/_ref/libm3ref_path_lm_mt.so (oracle/ref_host/Makefile; built in the container that has the reference, it
# travels to the GPU box).  kind "port": the oracle's restatement, when that library is absent.
# ---------------------------------------------------------------------------------------------
def cpu_path(w, n_sample, steps, warmup, budget_s=None):
    """Times SampleHandlerFD::Reweight (FindSplineSegment + CalcSplineWeights + CalcTotalEventWeight
    + FillArray_MP) + GetLikelihood on a bounded sample of workload `w`, DragRace style
    (Fitters/FitterBase.cpp:461-520).  Returns (events/s, ms/step, laps, cores, llh, sample workload, kind)."""
    from mach3_b200 import synth
    from oracle import ref_path_binding as RP     # the checker's reference build, here as the timed CPU baseline
    ws = w.scaled(min(n_sample, w.n_events))
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
    evs, ms, laps, cores, llh, ws, kind = cpu_path(w, args.cpu_sample_events, args.steps, min(args.warmup, 3), budget_s=120.0)
    what = ("the reference's own SampleHandlerFD::Reweight + GetLikelihood over its SMonolith (compiled from the reference "
            "sources: release flags -O3 -flto, MULTITHREAD, _LOW_MEMORY_STRUCTS_, no -march)" if kind == "reference" else
            "oracle port of the reference's MULTITHREAD CPU path (flags -O3 -fopenmp -flto, no -march)")
    sample = f"{ws.n_events} events of the {w.n_events}-event workload, {laps} timed steps, OpenMP {cores} threads, {what}"
    line = {"impl": "reference", "metric": "reweighted events/s per MCMC step (reweight+fill+LLH)", "value": evs,
            "unit": "events/s", "n_gpus": args.gpus, "steps": laps, "warmup": min(args.warmup, 3), "ms_per_step": ms,
            "llh_evals_per_s": 1e3 / ms, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "f32 weights / f64 histogram+LLH", "data": "synthetic",
            "config": {"workload": w.name, "sample": sample},
            "cpu_baseline": {"value": evs, "unit": "events/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": evs, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# B200 arm
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
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
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
    E = w.n_events
    # contiguous, tile-aligned event shards
    from mach3_b200 import sharding
    e0, e1 = sharding.shard_range(E, world, rank)
    n_local = e1 - e0

    flags = lib.FLAG_NO_FUSED_LLH if world > 1 else 0
    h = lib.Handle(device=local, test_statistic=w.test_statistic, update_w2=False, tile_events=args.tile, flags=flags)
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)
    t_setup = time.perf_counter()
    typ, npts, cx = synth.param_layout(w)
    h.splines_begin(w.n_params, w.n_knots, cx, npts, n_local)
    for c0 in range(e0, e1, CHUNK):
        h.splines_append(synth.make_splines(w, c0, min(e1, c0 + CHUNK)))
    h.splines_end()
    h.upload_binning(synth.bin_edges(w))
    ev = synth.make_events(w, e0, e1)
    h.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, None, 0,
                    ev["static_w"])
    del ev
    n_osc_bufs = 4
    # the caller's persistent oscillation-weight arrays live in pinned + mapped host memory from the library
    # (m3b_alloc_host; registering malloc'ed numpy memory gave less than half the PCIe rate on this pool)
    osc_bufs = []
    for k in range(n_osc_bufs):
        b = h.alloc_host(n_local, np.float32)
        b[:] = synth.make_osc(w, k, e0, e1)
        osc_bufs.append(b)
    h.upload_osc(osc_bufs[0])
    t_setup = time.perf_counter() - t_setup

    sh = sharding.ShardedSampleHandler(h, dist, args.exchange, device=f"cuda:{local}") if world > 1 else None

    def step(k, osc=None):
        sp, nm = props[k]
        if world == 1:
            # raw addresses, like the C++ host the library is made for (no per-call ctypes pointer extraction)
            h.step_addr(prop_addr[k][0], prop_addr[k][1], 0 if osc is None else osc_addr[id(osc)])
        else:
            sh.Reweight(sp, nm, osc)

    W, K = args.warmup, args.steps
    props = {k: synth.proposal(w, k) for k in range(-1, 2 * (W + K) + 16)}
    props = {k: (np.ascontiguousarray(sp, np.float64), np.ascontiguousarray(nm, np.float64)) for k, (sp, nm) in props.items()}
    prop_addr = {k: (lib.addr(sp), lib.addr(nm) if nm.size else 0) for k, (sp, nm) in props.items()}
    osc_addr = {id(b): lib.addr(b) for b in osc_bufs}

    # Asimov data at nominal, Poisson-fluctuated with a seed every rank shares
    step(-1); h.llh()
    mc, _ = h.read_hist()
    data = np.random.default_rng(w.seed).poisson(mc).astype(np.float64)
    h.upload_data(data)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM --------------------------------------------
    # K queued steps between two events on the launching stream (consecutive fused launches may overlap their
    # ramp/tail through programmatic dependent launch), then the same K steps again with the library's own events
    # around every launch (which serialises them): `value` comes from the first pass, the kernel's isolated launch
    # duration -- what the roofline uses -- from the second.
    for k in range(W):
        step(k)
    llh_w = h.llh()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = h.info().kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for k in range(W, W + K):
        step(k)
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = h.info().kernel_launches - launches0
    llh_last = h.llh()
    h.set_timing(True)
    h.kernel_time()
    for k in range(W, W + K):
        step(k)
    llh_last2 = h.llh()
    kern_ms, kern_n = h.kernel_time()
    h.set_timing(False)
    assert llh_last2 == llh_last or abs(llh_last2 - llh_last) <= 1e-9 * abs(llh_last), (llh_last, llh_last2)

    # ---------------- e2e: host buffers in, scalar out, every step -----------------------------
    for k in range(min(W, 5)):
        step(W + K + k, osc_bufs[k % n_osc_bufs]); h.llh()
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(stream)
    t_host = time.perf_counter()
    for k in range(K):
        step(W + K + 5 + k, osc_bufs[k % n_osc_bufs])
        llh_e2e = h.llh_fast()
    ev3.record(stream)
    barrier()
    t_host = time.perf_counter() - t_host
    ms_e2e = max(ev2.elapsed_time(ev3), 1e3 * t_host)
    clk = clocks.stop()

    info = h.info()
    if dist is not None:
        t = torch.tensor([ms_total, ms_e2e, kern_ms / max(kern_n, 1)], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e, kern_avg = t.tolist()
    else:
        kern_avg = kern_ms / max(kern_n, 1)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        ms_step = ms_total / K
        alg_bytes_local = n_local * w.bytes_per_event        # SURVEY §8d per-event figure x events of one launch
        achieved = alg_bytes_local / (kern_avg * 1e-3) / 1e9
        step_bytes = 12 * w.n_params + 4 * w.n_norm_params
        kname = "m3b::fill_tma_kernel" if info.kernel_variant < 0 else f"m3b::fill_kernel<{info.tile_events},variant {info.kernel_variant}>"
        # DRAM traffic per launch from the committed `ncu --set full` capture of this same command
        traffic, traffic_src = None, None
        prof = os.path.join(ROOT, "profiles", "r01_ncu_full_fill_tma_cfg2.json")
        if world == 1 and w is synth.CFG2 and info.kernel_variant < 0 and os.path.exists(prof):
            try:
                pj = json.load(open(prof))
                rd = [float(x) for x in pj["dram__bytes_read.sum"]["per_launch"]]
                wr = [float(x) for x in pj["dram__bytes_write.sum"]["per_launch"]]
                scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
                traffic = (sum(rd) / len(rd)) * scale[pj["dram__bytes_read.sum"]["unit"]] + \
                          (sum(wr) / len(wr)) * scale[pj["dram__bytes_write.sum"]["unit"]]
                traffic_src = "profiles/r01_ncu_full_fill_tma_cfg2.json (dram__bytes_read.sum + dram__bytes_write.sum, mean of 3 launches)"
            except Exception:
                traffic = None
        line = {
            "metric": "reweighted events/s per MCMC step (reweight+fill+LLH)",
            "value": E / (ms_step * 1e-3), "unit": "events/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "llh_evals_per_s": 1e3 / ms_step, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "f32 weights / f64 histogram+LLH", "data": "synthetic",
            "config": {"workload": w.name, "events": E, "events_per_gpu": n_local, "responses_per_event": w.n_params,
                       "bins": w.n_bins, "tile_events": info.tile_events, "grid_blocks": info.grid_blocks,
                       "smem_bytes": info.smem_bytes, "tma_stages": info.tma_stages, "exchange": ("none" if world == 1 else sh.exchange),
                       "l2": "inputs larger than L2: %.0f MB of coefficient rows stream per step per GPU, fresh "
                             "proposal (different segments) every step" % (info.active_bytes_per_step / 1e6),
                       "device_bytes": info.device_bytes, "setup_s": round(t_setup, 2), "host_cpu_affinity": numa},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                         "unit": "GB/s", "frac": achieved / peak, "frac_of_8TBs_nominal": achieved / 8000.0,
                         "kernel": kname, "kernel_ms": kern_avg, "algorithmic_bytes_per_launch": alg_bytes_local,
                         "loaded_bytes_per_launch": info.active_bytes_per_step, "traffic": traffic, "traffic_source": traffic_src,
                         "note": "kernel_ms = isolated launch duration (library's CUDA events around every launch, which "
                                 "serialises the launches); `value`/ms_per_step come from K queued steps whose ramp/tail "
                                 "overlap through programmatic dependent launch",
                         "achieved_queued": alg_bytes_local / (ms_step * 1e-3) / 1e9,
                         "frac_queued": alg_bytes_local / (ms_step * 1e-3) / 1e9 / peak},
            "e2e": {"value": E / (ms_e2e / K * 1e-3), "unit": "events/s", "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": int(4 * n_local + step_bytes), "d2h_bytes_per_step": int(8 * (1 + w.n_samples)),
                    "api": "m3b_step(host pars, host norms, host osc weights in pinned memory) + m3b_llh(); the osc weights "
                           "are streamed over PCIe by the fill kernel's own bulk copies (no separate H2D pass)"},
            "gpu_launches": int(launches), "clocks": clk,
            "llh": {"last_value_step": llh_last, "last_e2e_step": llh_e2e, "after_warmup": llh_w},
        }
        if world > 1:
            # the driver's N=1 run is cfg2 (BASELINE configs[1]); the same-workload single-GPU point of THIS strong-scaling
            # curve was measured once and committed (python bench.py --workload cfg3 on one B200: all events resident)
            try:
                one = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_cfg3_1gpu.json")))
                if one["config"]["events"] == E and one["config"]["responses_per_event"] == w.n_params:
                    line["same_workload_on_one_gpu"] = {"value": one["value"], "ms_per_step": one["ms_per_step"], "unit": "events/s",
                                                        "source": "profiles/r01_bench_cfg3_1gpu.json (committed measurement, not this run)"}
            except Exception:
                pass
        if world == 1 and not args.no_cpu_baseline:
            evs, ms, laps, cores, _, ws, kind = cpu_path(w, args.cpu_sample_events, 100, 2, budget_s=15.0)
            line["cpu_baseline"] = {"value": evs, "unit": "events/s", "cores": cores, "kind": kind, "ms_per_step": ms,
                                    "sample": f"{ws.n_events} events of the same workload, {laps} DragRace laps of "
                                              f"Reweight+GetLikelihood, OpenMP {cores} threads"
                                              + (" (the reference's own sources, oracle/_ref/libm3ref_path_lm_mt.so)"
                                                 if kind == "reference" else " (oracle port)")}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    h.close()


def main_cfg4(args):
    """BASELINE config 4 (BinnedSplineHandler workload), single GPU, optional bench line:
    python bench.py --workload cfg4 [--events N]   (not the default; the headline stays cfg2)."""
    import torch
    from mach3_b200 import handlers, lib
    from mach3_b200.synth import binned as B
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; no CPU fallback")
    w = B.CFG4 if not args.events else B.CFG4.scaled(n_events=args.events, n_grid=max(1000, int(B.CFG4.n_grid * args.events / B.CFG4.n_events)))
    t0 = time.perf_counter()
    sh, d = handlers.build_binned_from_workload(w, update_w2=True)
    h = sh.handle
    t_setup = time.perf_counter() - t0
    W, K = args.warmup, args.steps
    props = {k: B.proposal(w, k) for k in range(-1, W + K + 2)}

    def step(k):
        d["pars"][:], d["norm"][:] = props[k]
        sh.Reweight()

    step(-1); sh.GetLikelihood()
    sh.AddData(np.random.default_rng(w.seed).poisson(sh.GetMCArray()).astype(np.float64))
    for k in range(W):
        step(k)
    sh.GetLikelihood()
    h.set_timing(True); h.kernel_time()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for k in range(W, W + K):
        step(k)
    llh = sh.GetLikelihood()
    t_async = time.perf_counter() - t1
    kern_ms, kern_n = h.kernel_time()
    t1 = time.perf_counter()
    for k in range(K):
        step(W + (k % K)); llh_e = sh.GetLikelihood()
    t_e2e = time.perf_counter() - t1
    n_act = int(d["spl"]["uniquecoeffindices"].size)
    n_ptr = int(d["ev"]["spline_index"].size)
    mask = np.zeros(w.n_slots, bool); mask[d["spl"]["uniquecoeffindices"]] = True
    n_nonflat = int(mask[d["ev"]["spline_index"]].sum())              # pointers at flat splines (exactly 1.0) are dropped at upload
    del mask
    # eval: {y,b,c,d}+x read, weight write per non-flat spline; fill: event table + (index + gathered weight) per non-flat pointer
    alg = n_act * (16 + 4 + 4) + w.n_events * 8 + n_nonflat * 4 * 2
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kms = kern_ms / max(kern_n, 1)
    line = {"metric": "reweighted events/s per MCMC step (binned-spline eval + fill + Barlow-Beeston LLH)",
            "value": w.n_events / (t_async / K), "unit": "events/s", "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_async / K, "binned_spline_evals_per_s": n_act / (t_async / K), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 weights / f64 histogram+LLH", "data": "synthetic",
            "config": {"workload": w.name, "events": w.n_events, "active_binned_splines": n_act, "weight_pointers": n_ptr, "non_flat_weight_pointers": n_nonflat,
                       "slots": w.n_slots, "bins": w.n_bins, "setup_s": round(t_setup, 1),
                       "l2": "coefficient rows (%.0f MB/step) stream from HBM; the compact weight array (%.0f MB) is gathered through L2"
                             % (n_act * 20 / 1e6, n_act * 4 / 1e6)},
            "roofline": {"bound": "hbm", "achieved": alg / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (kms * 1e-3) / 1e9 / peak, "kernel": "m3b::binned_eval_kernel + m3b::binned_fill_kernel",
                         "kernel_ms": kms, "algorithmic_bytes_per_launch": alg, "traffic": None},
            "e2e": {"value": w.n_events / (t_e2e / K), "unit": "events/s", "ms_per_step": 1e3 * t_e2e / K,
                    "h2d_bytes_per_step": 12 * w.n_systs + 4 * w.n_norm_params, "d2h_bytes_per_step": 16},
            "gpu_launches": int(2 * K), "llh": {"last": llh, "last_e2e": llh_e}}
    print(json.dumps(line), flush=True)


def main_cfg5(args):
    """BASELINE config 5 (batched proposals), single GPU, optional bench line:
    python bench.py --workload cfg5 [--events N] [--steps batches]   -- 256 parameter sets per batch."""
    import torch
    from mach3_b200 import lib, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; no CPU fallback")
    w = synth.CFG5 if not args.events else synth.CFG5.scaled(args.events)
    n_sets = 256
    h = lib.Handle(test_statistic=w.test_statistic, update_w2=False, tile_events=args.tile)
    t0 = time.perf_counter()
    typ, npts, cx = synth.param_layout(w)
    h.splines_begin(w.n_params, w.n_knots, cx, npts, w.n_events)
    for c0 in range(0, w.n_events, CHUNK):
        h.splines_append(synth.make_splines(w, c0, min(w.n_events, c0 + CHUNK)))
    h.splines_end()
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

    W, K = max(1, min(args.warmup, 2)), max(1, min(args.steps, 5))
    bs = [batch(k) for k in range(W + K)]
    for k in range(W):
        h.step_batch(*bs[k])
    h.kernel_time()
    t1 = time.perf_counter()
    for k in range(W, W + K):
        tot = h.step_batch(*bs[k])
    t_b = (time.perf_counter() - t1) / K
    msb, nb = h.kernel_time()
    kms = msb / max(nb, 1)
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
    fp_instr = w.n_events * n_sets * (4 * (w.n_params - w.n_linear) + 2 * w.n_linear)     # FMA/MUL issue slots
    peak_issue = 148 * 128 * 1.965e9
    line = {"metric": "LLH evaluations/s over batched proposals (reweight+fill+LLH per parameter set)", "value": n_sets / t_b,
            "unit": "LLH evals/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": 1e3 * t_b, "us_per_llh": 1e6 * t_b / n_sets,
            "event_sets_per_s": w.n_events * n_sets / t_b, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 weights / f64 histogram+LLH", "data": "synthetic",
            "config": {"workload": w.name, "events": w.n_events, "sets_per_batch": n_sets, "bins": w.n_bins, "setup_s": round(t_setup, 1),
                       "single_set_kernel_ms": ms1 / max(n1, 1), "amortisation_vs_single_set": (ms1 / max(n1, 1)) * n_sets / kms,
                       "llh_scan_256_points": {"ms": 1e3 * t_scan, "us_per_llh": 1e6 * t_scan / n_sets, "kernel_ms": ms_scan / max(n_scan, 1),
                                               "amortisation_vs_single_set": (ms1 / max(n1, 1)) * n_sets / (ms_scan / max(n_scan, 1))}},
            "roofline": {"bound": "fp32 issue (CUDA cores; no contraction, so no tensor cores)", "achieved": fp_instr / (kms * 1e-3) / 1e12,
                         "peak": peak_issue / 1e12, "unit": "T FP32 instr/s", "frac": fp_instr / (kms * 1e-3) / peak_issue,
                         "kernel": "m3b::fill_batch_kernel", "kernel_ms": kms, "traffic": None},
            "e2e": {"value": n_sets / t_b, "unit": "LLH evals/s", "h2d_bytes_per_step": int(n_sets * (8 * w.n_params + 8 * w.n_norm_params)),
                    "d2h_bytes_per_step": int(n_sets * 8 * (1 + w.n_samples))},
            "gpu_launches": int(2 * K), "llh": {"first": float(tot[0]), "last": float(tot[-1])}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.workload == "cfg4":
        main_cfg4(a)
    elif a.workload == "cfg5":
        main_cfg5(a)
    elif a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
