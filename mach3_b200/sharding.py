"""Event sharding of one sample handler over the GPUs of a node.

The reference runs on one GPU (Manager/gpuUtils.cu:71).  Events are independent given the
parameters; the only coupling of the path is the histogram sum in front of the (non-linear)
likelihood (Samples/SampleHandlerFD.cpp:1284-1300).  So: contiguous, tile-aligned event shards, one
process per GPU, every rank fills a partial histogram, ONE exchange of the partials, every rank
reduces the same -lnL.

    shard_range           which events rank r owns
    ShardedSampleHandler  Reweight()/GetLikelihood() over a lib.Handle per rank; the exchange is
                          either torch.distributed all_reduce (NCCL over NVLink) on the library's
                          own histogram buffer, or the library's peer-memory push (m3b_step_peer)
"""
from __future__ import annotations

import numpy as np

TILE_ALIGN = 1024      # largest tile row of the device layout: shards never split a tile


def shard_range(n_events: int, world: int, rank: int, align: int = TILE_ALIGN):
    """Events [e0, e1) of `rank`: equal contiguous shards rounded up to `align`; the last ranks may
    get fewer (or no) events.  The union over ranks is exactly [0, n_events), in rank order."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = -(-n_events // world)
    per = -(-per // align) * align
    e0 = min(n_events, rank * per)
    e1 = min(n_events, (rank + 1) * per)
    return e0, e1


class _DevArray:
    """torch view of a raw device pointer (plumbing for the all-reduce; no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}


class ShardedSampleHandler:
    """One rank's part of an event-sharded sample handler.

    `handle` is a lib.Handle created with FLAG_NO_FUSED_LLH that holds this rank's shard (splines,
    events, full binning, full data histogram).  `dist` is an initialised torch.distributed module
    (or None for world 1).  With the NCCL exchange the handle is put on torch's current CUDA stream of
    `device` (the stream the all-reduce is issued on): fill -> all-reduce -> likelihood are then stream-ordered;
    callers must keep that stream current while they call Reweight()."""

    def __init__(self, handle, dist=None, exchange="nccl", device=None):
        self.h = handle
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.exchange = exchange if self.world > 1 else "none"
        self._hist = None
        if self.world > 1:
            if exchange in ("peer", "auto"):
                # the library's own exchange: CUDA-IPC peer mappings of every rank's partial histogram.
                # "auto" falls back to NCCL (on all ranks together) if the mappings cannot be made.
                ok = 1
                try:
                    mine = handle.peer_export(self.rank, self.world)
                except Exception:
                    if exchange == "peer":
                        raise
                    mine, ok = None, 0
                allh = [None] * self.world
                dist.all_gather_object(allh, mine)
                if all(x is not None for x in allh):
                    try:
                        for r in range(self.world):
                            handle.peer_import(r, allh[r])
                    except Exception:
                        if exchange == "peer":
                            raise
                        ok = 0
                else:
                    ok = 0
                oks = [None] * self.world
                dist.all_gather_object(oks, ok)
                self.exchange = exchange = "peer" if all(oks) else "nccl"
            if exchange == "peer":
                pass
            elif exchange == "nccl":
                import torch
                # the all-reduce runs on torch's current stream; the library's fill and likelihood launches must be
                # ordered with it, so the handle is moved onto that stream (it owns a private non-blocking one by default)
                handle.set_stream(torch.cuda.current_stream(device).cuda_stream)
                ptr, nb, _ = handle.hist_device_ptr()
                self._hist = torch.as_tensor(_DevArray(ptr, 2 * nb), device=device)
            else:
                raise ValueError(exchange)

    def Reweight(self, spline_pars, norm_pars=None, osc_w=None):
        h = self.h
        if self.world == 1:
            h.step(spline_pars, norm_pars, osc_w, mode="fill")
            h.llh_from_hist()
        elif self.exchange == "peer":
            h.step(spline_pars, norm_pars, osc_w, mode="peer")
        else:
            h.step(spline_pars, norm_pars, osc_w, mode="fill")
            _, nb, live = h.hist_device_ptr()
            self.dist.all_reduce(self._hist[: (2 * nb if live else nb)])
            h.llh_from_hist()

    def GetLikelihood(self):
        return self.h.llh()


def allreduce_partial_histograms(dist, mc: np.ndarray, w2: np.ndarray | None = None):
    """Host-array form of the exchange (any backend, e.g. gloo): sums the ranks' partial histograms in
    place.  mc‖w2 travel as ONE message, like the device path."""
    import torch
    buf = np.concatenate([mc, w2]) if w2 is not None else mc.copy()
    t = torch.from_numpy(buf)
    dist.all_reduce(t)
    n = mc.size
    mc[:] = buf[:n]
    if w2 is not None:
        w2[:] = buf[n:]
    return mc, w2
