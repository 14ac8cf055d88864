"""world_size-2 gloo test of the N>1 host path: partition, ingest scatter, score gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nanowakeword_b200.sharding import ShardedScorer, partition


def test_partition_covers_everything():
    for n in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [partition(n, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == n
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_score(pcm):           # stand-in for the engine: any per-window map
    return (pcm.float().abs().mean(dim=1) / 32768.0).contiguous()


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        clip = 160
        full = torch.from_numpy(np.random.default_rng(3).integers(-32768, 32768, (n_total, clip), dtype=np.int16))
        scorer = ShardedScorer(_fake_score, clip, rank, world, torch.device("cpu"))
        got = scorer.score_from_root(full if rank == 0 else None, n_total)
        start, count = partition(n_total, world, rank)
        got2 = scorer.score_resident(full[start:start + count], n_total)
        got3 = scorer.score_from_root_pipelined(full if rank == 0 else None, n_total, n_chunks=4)
        if rank == 0:
            ref = _fake_score(full)
            q.put((torch.equal(got, ref), torch.equal(got2, ref) and torch.equal(got3, ref)))
        else:
            assert got is None and got2 is None and got3 is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [64, 37])
def test_scatter_score_gather_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == (True, True)
