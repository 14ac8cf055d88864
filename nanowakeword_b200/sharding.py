"""Multi-GPU plumbing: independent windows / streams shard across ranks, one process per GPU.

The reference has no distributed path at all (no torch.distributed, NCCL or MPI);
per-window scoring has no cross-window dependency (per-stream state is private,
reference nanointerpreter.py:150-154, 181-182), so the path shards by a contiguous block
partition of the window axis with replicated weights and NO data-path collective.  The only
exchanges are the optional ingest scatter (when one rank owns the audio) and the gather of
scores (4 bytes per window) — both through ``torch.distributed`` (NCCL on GPUs, gloo in the
CPU tests).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np


def partition(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition of ``n`` units: rank r owns [start, start + count)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def scatter_windows(pcm_root, n_total: int, clip_samples: int, rank: int, world: int, device, src: int = 0):
    """Rank ``src`` holds ``pcm_root`` (n_total, clip_samples) int16; every rank receives its
    block.  Point-to-point sends of exactly the owned rows (no padding), as the partition is
    contiguous."""
    import torch
    import torch.distributed as dist
    start, count = partition(n_total, world, rank)
    local = torch.empty((count, clip_samples), dtype=torch.int16, device=device)
    if world == 1:
        local.copy_(pcm_root[start:start + count])
        return local
    # ONE grouped NCCL call (ncclGroupStart / End around every send): the per-destination transfers run concurrently
    # over NVSwitch instead of being serialised as independent point-to-point operations.
    ops = []
    if rank == src:
        for r in range(world):
            s, c = partition(n_total, world, r)
            if r == src:
                local.copy_(pcm_root[s:s + c])
            elif c:
                # NCCL has no int16: ship the rows as raw bytes
                ops.append(dist.P2POp(dist.isend, pcm_root[s:s + c].contiguous().view(torch.uint8), r))
    elif count:
        ops.append(dist.P2POp(dist.irecv, local.view(torch.uint8), src))
    for q in (dist.batch_isend_irecv(ops) if ops else []):
        q.wait()
    return local


def gather_scores(local_scores, n_total: int, rank: int, world: int, dst: int = 0):
    """Collect per-rank score blocks on ``dst`` in window order (padded all_gather, since
    NCCL collectives need equal counts).  Returns the full (n_total,) tensor on ``dst``,
    None elsewhere."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_scores
    width = -(-n_total // world)
    buf = torch.zeros(width, dtype=local_scores.dtype, device=local_scores.device)
    buf[:local_scores.numel()] = local_scores
    out = torch.empty(world * width, dtype=local_scores.dtype, device=local_scores.device)
    dist.all_gather_into_tensor(out, buf)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        _, c = partition(n_total, world, r)
        parts.append(out[r * width:r * width + c])
    return torch.cat(parts)


class ShardedScorer:
    """Score a global batch across ranks: scatter (optional) -> local engine -> gather."""

    def __init__(self, score_local: Callable, clip_samples: int, rank: int, world: int, device):
        self.score_local = score_local          # (count, clip) int16 tensor on `device` -> (count,) float32 tensor
        self.clip_samples = clip_samples
        self.rank, self.world, self.device = rank, world, device

    def score_from_root(self, pcm_root, n_total: int):
        local = scatter_windows(pcm_root, n_total, self.clip_samples, self.rank, self.world, self.device)
        scores = self.score_local(local)
        return gather_scores(scores, n_total, self.rank, self.world)

    def score_from_root_pipelined(self, pcm_root, n_total: int, n_chunks: int = 2):
        """Same result as :meth:`score_from_root`, with the ingest overlapped with compute: every rank's block is
        cut into ``n_chunks`` pieces; all transfers are queued on a side CUDA stream (rank ``0`` sends piece k of
        every block as one grouped NCCL call) and the engine scores piece k as soon as its event fires, while
        piece k+1 is still on NVLink.  On CPU tensors (gloo tests) it degrades to the sequential path."""
        import torch
        import torch.distributed as dist
        if self.world == 1 or torch.device(self.device).type != "cuda":
            return self.score_from_root(pcm_root, n_total)
        start, count = partition(n_total, self.world, self.rank)
        n_chunks = max(1, min(n_chunks, count if count else 1))
        local = torch.empty((count, self.clip_samples), dtype=torch.int16, device=self.device)
        bounds = [partition(count, n_chunks, k) for k in range(n_chunks)]           # (offset, length) inside the block
        comm = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        comm.wait_stream(main)
        events = []
        with torch.cuda.stream(comm):
            for k in range(n_chunks):
                ops = []
                if self.rank == 0:
                    for r in range(self.world):
                        s_r, c_r = partition(n_total, self.world, r)
                        o, l = partition(c_r, n_chunks, k)
                        if l == 0:
                            continue
                        piece = pcm_root[s_r + o:s_r + o + l]
                        if r == 0:
                            local[o:o + l].copy_(piece, non_blocking=True)
                        else:
                            ops.append(dist.P2POp(dist.isend, piece.view(torch.uint8), r))
                else:
                    o, l = bounds[k]
                    if l:
                        ops.append(dist.P2POp(dist.irecv, local[o:o + l].view(torch.uint8), 0))
                for req in (dist.batch_isend_irecv(ops) if ops else []):
                    req.wait()
                ev = torch.cuda.Event()
                ev.record(comm)
                events.append(ev)
        scores = torch.empty(count, dtype=torch.float32, device=self.device)
        for k, (o, l) in enumerate(bounds):
            main.wait_event(events[k])
            if l:
                scores[o:o + l] = self.score_local(local[o:o + l])
        main.wait_stream(comm)
        return gather_scores(scores, n_total, self.rank, self.world)

    def score_resident(self, local_pcm, n_total: Optional[int] = None):
        """Each rank already holds its block (the replicas / per-GPU ingest case)."""
        scores = self.score_local(local_pcm)
        if n_total is None:
            n_total = local_pcm.shape[0] * self.world
        return gather_scores(scores, n_total, self.rank, self.world)
