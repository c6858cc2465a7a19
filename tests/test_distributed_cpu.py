"""N>1 host logic on CPU: frame sharding and the sum-reduce of accumulation buffers with the gloo backend,
world_size 2 (the GPU path uses the same functions with NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spectral_raytracer_b200.distributed import frame_shard, reduce_sum_


def test_frame_shard_partitions_exactly():
    for first, n, world in [(0, 1024, 8), (5, 17, 4), (0, 3, 8), (100, 64, 1), (0, 0, 2)]:
        seen = []
        for r in range(world):
            s, c = frame_shard(first, n, r, world)
            seen += list(range(s, s + c))
        assert seen == list(range(first, first + n))
        counts = [frame_shard(first, n, r, world)[1] for r in range(world)]
        assert max(counts) - min(counts) <= 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank "renders" its shard: buffer = sum over its frames of a frame-dependent pattern
        first, count = frame_shard(0, 11, rank, world)
        buf = torch.zeros(6 * 4 * 8, dtype=torch.float32)
        for f in range(first, first + count):
            buf += torch.arange(buf.numel(), dtype=torch.float32) * (f + 1)
        total = reduce_sum_(buf, count, dst=0)
        if rank == 0:
            want = torch.arange(buf.numel(), dtype=torch.float32) * sum(range(1, 12))
            out.put((total, bool(torch.allclose(buf, want))))
    finally:
        dist.destroy_process_group()


def test_gloo_reduce_of_sharded_accumulation_buffers():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    total, ok = out.get(timeout=10)
    assert total == 11 and ok
