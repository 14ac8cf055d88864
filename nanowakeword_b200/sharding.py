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
    if rank == src:
        reqs = []
        for r in range(world):
            s, c = partition(n_total, world, r)
            if r == src:
                local.copy_(pcm_root[s:s + c])
            elif c:
                # NCCL has no int16: ship the rows as raw bytes
                reqs.append(dist.isend(pcm_root[s:s + c].contiguous().view(torch.uint8), dst=r))
        for q in reqs:
            q.wait()
    elif count:
        dist.recv(local.view(torch.uint8), src=src)
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

    def score_resident(self, local_pcm, n_total: Optional[int] = None):
        """Each rank already holds its block (the replicas / per-GPU ingest case)."""
        scores = self.score_local(local_pcm)
        if n_total is None:
            n_total = local_pcm.shape[0] * self.world
        return gather_scores(scores, n_total, self.rank, self.world)
