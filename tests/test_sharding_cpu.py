"""Multi-GPU host logic on CPU: event sharding + ONE all-reduce of partial histograms (gloo,
world_size 2 and 3) reproduce the single-process histogram and -lnL.  The partial histograms come
from the oracle (the checker); the product's kernels are exercised by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mach3_b200 import sharding, synth  # noqa: E402


def test_shard_ranges_tile_the_event_list():
    for n in (0, 1, 1023, 1024, 1025, 100_000, 20_000_000):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                e0, e1 = sharding.shard_range(n, world, r)
                assert e0 == prev and e0 <= e1 <= n
                assert e0 % sharding.TILE_ALIGN == 0 or e0 == n
                prev = e1
            assert prev == n
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_events, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      OMP_NUM_THREADS="2")
    import torch.distributed as dist
    from oracle import binding as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = synth.SPARSE.scaled(n_events)
        e0, e1 = sharding.shard_range(w.n_events, world, rank, align=256)
        out = []
        data = None
        for step in (-1, 0, 1, 2):
            sp, nm = synth.proposal(w, step)
            mc = np.zeros(w.n_bins); w2 = np.zeros(w.n_bins)
            if e1 > e0:
                mono, sh, d = O.build_from_workload(w, e0, e1, update_w2=True)
                mono.set_params(sp); sh.norm_vals[:] = nm
                sh.Reweight()
                mc[:] = sh.mc; w2[:] = sh.w2
            sharding.allreduce_partial_histograms(dist, mc, w2)
            if data is None:
                data = np.random.default_rng(5).poisson(mc).astype(np.float64)
            llh = sum(O.test_stat_llh(w.test_statistic, data[b], mc[b], w2[b]) for b in range(w.n_bins))
            out.append((mc.copy(), w2.copy(), llh))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_fill_allreduce_llh_equals_single_process(world):
    from oracle import binding as O
    n_events = 6_007
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_events, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process reference
    w = synth.SPARSE.scaled(n_events)
    mono, sh, d = O.build_from_workload(w, update_w2=True)
    data = None
    for i, step in enumerate((-1, 0, 1, 2)):
        sp, nm = synth.proposal(w, step)
        mono.set_params(sp); sh.norm_vals[:] = nm
        sh.Reweight()
        if data is None:
            data = np.random.default_rng(5).poisson(sh.mc).astype(np.float64)
            sh.AddData(data)
        ref = sh.GetLikelihood()
        for r in range(world):
            mc, w2, llh = res[r][i]
            np.testing.assert_allclose(mc, sh.mc, rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(w2, sh.w2, rtol=1e-12, atol=1e-12)
            assert llh == pytest.approx(ref, rel=1e-10, abs=1e-9)
            # every rank holds bit-identical totals after the all-reduce
            np.testing.assert_array_equal(mc, res[0][i][0])
