"""world_size-2 gloo test (CPU) of the image-sharding host logic: shards are a partition of the batch, results come
back in input order, and sharded == unsharded bit for bit for a per-image-independent forward."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from candle_birefnet_b200.shard import forward_sharded, shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 5, 16, 17, 64):
        for world in (1, 2, 3, 4, 8):
            b = shard_bounds(n, world)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _fake_forward(x: np.ndarray) -> np.ndarray:
    # per-image independent, deterministic stand-in for forward_logits ([b,3,H,W] -> [b,1,H,W])
    return (x[:, :1] * 2.0 + x[:, 1:2] - x[:, 2:3] * 0.5).astype(np.float32)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 3, 8, 12)).astype(np.float32)
    local = forward_sharded(_fake_forward, x, rank, world, gather=False)
    lo, hi = shard_bounds(5, world)[rank]
    ok_local = np.array_equal(local, _fake_forward(x)[lo:hi])
    full = forward_sharded(_fake_forward, x, rank, world, gather=True)
    dist.barrier()
    if rank == 0:
        q.put((ok_local, bool(np.array_equal(full, _fake_forward(x)))))
    else:
        q.put((ok_local, full is None))
    dist.destroy_process_group()


def test_sharded_equals_unsharded_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(a and b for a, b in res), res
