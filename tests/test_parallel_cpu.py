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
                before = pool.off
                v = pool.take(shape)
                v += float(rank + 1) * (i + 1 + step)           # "wgrad kernel" accumulates into the zeroed slice
                slices.append(v)
                pool.commit()
                # a bucket launched at this commit must not contain the committing node's own range —
                # autograd may still read it (AccumulateGrad clone of a permuted view, dtype cast) after the commit
                assert all(b <= before for (_, b) in red.launched), (red.launched, before)
            red.finish()
            assert red.n_coll >= 2 and pool.on_commit is None
            for i, v in enumerate(slices):
                want = (i + 1 + step) * sum(range(1, world + 1)) / world
                assert red.in_pool(v) and torch.allclose(v, torch.full_like(v, want)), (step, i)
            assert not red.in_pool(torch.zeros(3))
        # 4. global-batch Dice (reference nn.DataParallel semantics, trainer.py:55-57): the per-rank gradient of
        #    seg_loss(..., global_dice_group=True) w.r.t. the local logits == world x the gradient of the single-process loss on
        #    the concatenated batch (the gradient all-reduce AVERAGES over ranks)
        from cswin_unet_b200.train import seg_loss
        g = torch.Generator().manual_seed(7)
        full_logits = torch.randn(2 * world, 4, 6, 5, generator=g, dtype=torch.float64)
        full_labels = torch.randint(0, 4, (2 * world, 6, 5), generator=g)
        fl = full_logits.clone().requires_grad_(True)
        want_loss = seg_loss(fl.float(), full_labels, 4)
        (g_full,) = torch.autograd.grad(want_loss, fl)
        mine = full_logits[2 * rank:2 * rank + 2].clone().requires_grad_(True)
        loss = seg_loss(mine.float(), full_labels[2 * rank:2 * rank + 2], 4, global_dice_group=True)
        (g_mine,) = torch.autograd.grad(loss, mine)
        assert torch.allclose(g_mine, world * g_full[2 * rank:2 * rank + 2], rtol=1e-4, atol=1e-7), (g_mine - world * g_full[2 * rank:2 * rank + 2]).abs().max()
        # loss value: CE is the local mean, Dice the global one; averaged over ranks it is the single-process loss
        lv = loss.detach().double().clone()
        dist.all_reduce(lv)
        assert abs(float(lv) / world - float(want_loss)) < 1e-5
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 150])
def test_gloo_world2_gather_and_allreduce(tmp_path, n_total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}.npy") for r in range(world))
