"""N>1 host logic on CPU with the gloo backend (world_size 2): environment sharding covers the batch
exactly once and the benchmark reductions (max of timings, sum of env-steps) agree on every rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fluidgym_b200.distributed import max_over_ranks, shard_envs, sum_over_ranks


def test_shards_partition_the_batch():
    for total in (1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                seen += list(shard_envs(total, world, r))
            assert seen == list(range(total))
            sizes = [len(shard_envs(total, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_envs(4, 2, 2)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = shard_envs(10, world, rank)
    ms = 5.0 + 3.0 * rank                     # rank-dependent "device time"
    steps = len(shard) * 25
    res = (max_over_ranks(ms), sum_over_ranks(steps), list(shard))
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        out.put(gathered)
    dist.destroy_process_group()


def test_gloo_world_size_2_reductions():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (m0, s0, sh0), (m1, s1, sh1) = gathered
    assert m0 == m1 == 8.0
    assert s0 == s1 == 250.0
    assert sh0 + sh1 == list(range(10))
