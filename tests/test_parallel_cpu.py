"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: slice sharding, label-shard gather, gradient all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cswin_unet_b200 import parallel
from cswin_unet_b200.engine import shard_slices


def test_shard_slices_partition():
    for n in (0, 1, 7, 150, 151):
        for world in (1, 2, 3, 8):
            parts = [list(shard_slices(n, world, r)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. label gather: every rank "segments" its shard (label = slice index mod 251), gathers the volume
        rng = shard_slices(n_total, world, rank)
        local = torch.stack([torch.full((4, 5), d % 251, dtype=torch.uint8) for d in rng]) if len(rng) else torch.zeros((0, 4, 5), dtype=torch.uint8)
        full = parallel.gather_label_shards(local, n_total)
        assert full.shape == (n_total, 4, 5)
        assert torch.equal(full[:, 0, 0], (torch.arange(n_total) % 251).to(torch.uint8))
        # 2. gradient all-reduce: mean over ranks, bucketed, grads of different sizes
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(s)) for s in ((3, 4), (5,), (1000,), (2, 2, 2))]
        for i, p in enumerate(params):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        n = parallel.allreduce_gradients(params, bucket_bytes=64)
        assert n >= 2
        for i, p in enumerate(params):
            want = (i + 1) * sum(range(1, world + 1)) / world
            assert torch.allclose(p.grad, torch.full_like(p, want))
        # 3. overlapped pooled reducer: slices of ONE flat buffer handed out in backward order, reduced in place in buckets
        from cswin_unet_b200 import autograd as ag
        pool = ag.ZeroPool(4096, "cpu")
        red = parallel.PoolGradReducer(pool, bucket_bytes=4 * 300)
        for step in range(2):                                   # two steps: the pool and the reducer are re-armed every step
            pool.reset()
            red.begin()
            slices = []
            for i, shape in enumerate(((10, 10), (7,), (333,), (64, 3), (5,))):
                v = pool.take(shape)
                v += float(rank + 1) * (i + 1 + step)           # "wgrad kernel" accumulates into the zeroed slice
                slices.append(v)
                pool.commit()
            red.finish()
            assert red.n_coll >= 2 and pool.on_commit is None
            for i, v in enumerate(slices):
                want = (i + 1 + step) * sum(range(1, world + 1)) / world
                assert red.in_pool(v) and torch.allclose(v, torch.full_like(v, want)), (step, i)
            assert not red.in_pool(torch.zeros(3))
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 150])
def test_gloo_world2_gather_and_allreduce(tmp_path, n_total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}.npy") for r in range(world))
