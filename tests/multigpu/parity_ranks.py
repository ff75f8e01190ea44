"""Launched by torchrun (one process per GPU): event-sharded fill + histogram exchange + -lnL on the
B200s against the single-process CPU oracle on the same seeded workload.  Both exchanges
(NCCL all-reduce on the library's buffer; the library's own peer-memory pull) must give every rank
the same -lnL, equal to the oracle's within the north_star tolerance (1e-6 relative).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu/parity_ranks.py

run_parity() is also what `bench.py --gpus N` calls, untimed, after its timed region (TEST INFRASTRUCTURE: it drives
the oracle as the checker).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LLH_RTOL = 1e-6
HIST_RTOL = 1e-10


def run_parity(dist, rank, world, local, n_events=150_001, verbose=True, exchanges=("nccl", "peer")):
    """Every rank calls this inside an initialised NCCL process group.  Returns (on every rank) a dict
    {ok, n_events, steps, exchanges: {name: {rel_llh, max_rel_hist, ranks_identical}}}: the worst case over the steps of
    |-lnL_gpu - -lnL_oracle| / |-lnL_oracle|, of the histogram difference (mc and w2, relative, bins above 1e-10), and
    whether all ranks hold the same -lnL (bitwise for the peer pull, which sums in rank order on every rank; to 1e-12
    for NCCL)."""
    import torch
    from mach3_b200 import lib, sharding, synth
    from oracle import binding as O              # the checker
    w = synth.CFG3.scaled(n_events)
    e0, e1 = sharding.shard_range(w.n_events, world, rank)
    typ, npts, cx = synth.param_layout(w)
    steps = [-1, 0, 1, 2, 3]

    # oracle, whole workload, rank 0 only
    ref, data = None, None
    if rank == 0:
        O.set_multithread(False)
        mono, osh, _ = O.build_from_workload(w, update_w2=True, test_statistic=lib.BARLOW_BEESTON)
        sp, nm = synth.proposal(w, -1)
        mono.set_params(sp); osh.norm_vals[:] = nm
        osh.Reweight()
        data = np.random.default_rng(11).poisson(osh.mc).astype(np.float64)
        osh.AddData(data)
        ref = []
        for s in steps:
            sp, nm = synth.proposal(w, s)
            mono.set_params(sp); osh.norm_vals[:] = nm
            osh.Reweight()
            ref.append((osh.GetLikelihood(), osh.mc.copy(), osh.w2.copy()))
        O.set_multithread(True)
    box = [data]
    dist.broadcast_object_list(box, src=0)
    data = box[0]

    def rel_hist(a, b):
        big = np.abs(b) > 1e-10
        d = np.zeros_like(a)
        d[big] = np.abs(a[big] - b[big]) / np.abs(b[big])
        d[~big] = np.abs(a[~big] - b[~big])
        return float(d.max()) if d.size else 0.0

    out = {"ok": True, "n_events": int(n_events), "steps": len(steps), "tolerance_rel_llh": LLH_RTOL,
           "checker": "oracle/m3_oracle.c, serial build, whole workload in one process, Barlow-Beeston with live W2", "exchanges": {}}
    for exchange in exchanges:
        h = lib.Handle(device=local, test_statistic=lib.BARLOW_BEESTON, update_w2=True, flags=lib.FLAG_NO_FUSED_LLH)
        h.set_stream(torch.cuda.current_stream().cuda_stream)
        h.splines_begin(w.n_params, w.n_knots, cx, npts, e1 - e0)
        if e1 > e0:
            h.splines_append(synth.make_splines(w, e0, e1))
        h.splines_end()
        h.upload_binning(synth.bin_edges(w))
        ev = synth.make_events(w, e0, e1)
        h.upload_events(ev["sample_id"], ev["kin"], ev["norm_idx"], w.n_norm_per_event, w.n_norm_params, True, None, 0,
                        ev["static_w"])
        h.upload_osc(synth.make_osc(w, 0, e0, e1))
        h.upload_data(data)
        sh = sharding.ShardedSampleHandler(h, dist, exchange, device=f"cuda:{local}")
        worst_llh, worst_hist, identical = 0.0, 0.0, True
        for i, s in enumerate(steps):
            sp, nm = synth.proposal(w, s)
            sh.Reweight(sp, nm)
            llh = sh.GetLikelihood()
            mc, w2 = h.read_hist()
            t = torch.tensor([llh], dtype=torch.float64, device=f"cuda:{local}")
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            vals = [float(v.item()) for v in allv]
            if rank == 0:
                r_llh, r_mc, r_w2 = ref[i]
                same = all(v == vals[0] for v in vals) if exchange == "peer" else all(abs(v - vals[0]) <= 1e-12 * abs(vals[0]) for v in vals)
                rl = abs(llh - r_llh) / max(abs(r_llh), 1e-300)
                rh = max(rel_hist(mc, r_mc), rel_hist(w2, r_w2))
                worst_llh, worst_hist, identical = max(worst_llh, rl), max(worst_hist, rh), identical and same
                if verbose:
                    print(f"[{exchange}] step {s:2d}: -lnL gpu {llh:.9f} oracle {r_llh:.9f} rel {rl:.2e} hist {rh:.1e} "
                          f"ranks agree {same}", flush=True)
        dist.barrier()
        h.close()
        good = worst_llh <= LLH_RTOL and worst_hist <= HIST_RTOL and identical
        out["exchanges"][sh.exchange if exchange == "auto" else exchange] = {
            "rel_llh": worst_llh, "max_rel_hist": worst_hist, "ranks_identical": bool(identical), "ok": bool(good)}
        out["ok"] = bool(out["ok"] and good)
    box = [out]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = run_parity(dist, rank, world, local, int(os.environ.get("M3B_PARITY_EVENTS", "150001")))
    dist.destroy_process_group()
    if rank == 0:
        print(res, flush=True)
    if not res["ok"]:
        raise SystemExit(1)
    if rank == 0:
        print("MULTI-GPU PARITY OK", flush=True)


if __name__ == "__main__":
    main()
