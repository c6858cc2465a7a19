"""N>1 host logic on CPU: frame sharding and the sum-reduce of accumulation buffers with the gloo backend,
world_size 2 (the GPU path uses the same functions with NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spectral_raytracer_b200.distributed import frame_shard, reduce_sum_, render_sharded


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


class _CpuContext:
    """Stand-in for a render context on CPU: 'rendering' frame f adds a frame-dependent pattern."""

    def __init__(self, n):
        self.buf = torch.zeros(n, dtype=torch.float32)
        self.frames_accumulated = 0

    def accum_tensor(self):
        return self.buf

    def render_frames(self, first, count):
        for f in range(first, first + count):
            self.buf += torch.arange(self.buf.numel(), dtype=torch.float32) * (f + 1)
        self.frames_accumulated += count

    def clear(self):
        self.buf.zero_()
        self.frames_accumulated = 0


def _worker_rounds(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = _CpuContext(6 * 4 * 8)
        if rank == 0:
            ctx.render_frames(100, 3)  # the root already holds frames (a resumed render)
        t1 = render_sharded(ctx, 0, 11, frames_per_call=4)
        after1 = (ctx.frames_accumulated, float(ctx.buf.abs().sum()))
        t2 = render_sharded(ctx, 11, 5, frames_per_call=4)  # a second round must not add the old shards again
        if rank == 0:
            frames = list(range(100, 103)) + list(range(0, 16))
            want = torch.arange(ctx.buf.numel(), dtype=torch.float32) * sum(f + 1 for f in frames)
            out.put(("root", t1, t2, ctx.frames_accumulated, bool(torch.allclose(ctx.buf, want))))
        else:
            out.put(("other", after1, ctx.frames_accumulated, float(ctx.buf.abs().sum())))
    finally:
        dist.destroy_process_group()


def test_gloo_render_sharded_twice_counts_every_frame_once():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker_rounds, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict((r[0], r[1:]) for r in (out.get(timeout=10), out.get(timeout=10)))
    t1, t2, frames, ok = got["root"]
    assert (t1, t2, frames) == (14, 19, 19) and ok
    after1, frames_other, abs_sum = got["other"]
    assert after1 == (0, 0.0) and frames_other == 0 and abs_sum == 0.0   # non-root contexts end every round empty
