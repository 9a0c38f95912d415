"""Multi-GPU plumbing (torch.distributed; one process per GPU, NCCL on GPUs, gloo in the CPU tests).

Inference shards slices with no collective on the forward path; the only exchange is the gather of the uint8 label
shards at the end of a volume.  Training (data parallel) has exactly one exchange per step: the gradient all-reduce
(reference: single-process nn.DataParallel, trainer.py:37-38).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from .engine import shard_slices  # noqa: F401  (re-export)

Tensor = torch.Tensor


def gather_label_shards(local: Tensor, n_total: int, group=None) -> Tensor:
    """All-gather contiguous slice shards (n_local, H, W) uint8 -> (n_total, H, W) on every rank (shard_slices order)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        assert local.shape[0] == n_total
        return local
    world = dist.get_world_size(group)
    per = (n_total + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = [parts[r][: len(shard_slices(n_total, world, r))] for r in range(world)]
    return torch.cat(out, dim=0)


def allreduce_gradients(params: Iterable[Tensor], group=None, bucket_bytes: int = 32 << 20, average: bool = True) -> int:
    """Bucketed gradient all-reduce (sum, then / world).  Gradients are packed in REVERSE parameter order (the order the
    backward produces them) into flat ~32 MB buckets with one multi-tensor copy per bucket, reduced with one NCCL
    all-reduce per bucket, and unpacked with one multi-tensor copy.  Returns the number of collectives issued."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    grads = [p.grad for p in reversed(list(params)) if p.grad is not None]
    n_coll, i = 0, 0
    while i < len(grads):
        bucket: List[Tensor] = []
        size = 0
        dtype = grads[i].dtype
        while i < len(grads) and grads[i].dtype == dtype and (not bucket or size + grads[i].numel() * grads[i].element_size() <= bucket_bytes):
            bucket.append(grads[i]); size += grads[i].numel() * grads[i].element_size(); i += 1
        flat = torch.empty(sum(g.numel() for g in bucket), dtype=dtype, device=bucket[0].device)
        views, off = [], 0
        for g in bucket:
            views.append(flat[off: off + g.numel()].view(g.shape)); off += g.numel()
        torch._foreach_copy_(views, bucket)
        if average and flat.is_cuda and dist.get_backend(group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)       # averaged inside the collective: no div_ launch
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                flat.div_(world)
        torch._foreach_copy_(bucket, views)
        n_coll += 1
    return n_coll


class PoolGradReducer:
    """Gradient all-reduce overlapped with the backward, without packing: the weight-gradient kernels accumulate into slices of
    ONE flat fp32 buffer (autograd.ZeroPool) handed out in the order the backward reaches them, so a finished prefix of that
    buffer IS a gradient bucket.  Whenever ~bucket_bytes more have been committed (ZeroPool.commit, called after the producing
    kernels are enqueued) the range is all-reduced in place on a side stream (NCCL over NVLink / NVSwitch) while the backward
    keeps running on the main stream; `finish()` reduces the tail and joins the streams.  Capturable into the step's CUDA graph.
    Gradients that do not live in the pool (a few produced by torch ops on the tape) are handled by `allreduce_gradients`."""

    def __init__(self, pool, group=None, bucket_bytes: int = 16 << 20, world: Optional[int] = None):
        self.pool, self.group, self.bucket = pool, group, bucket_bytes // 4
        # world = 1 (no process group needed): no collective at all, but `after` still fires per bucket on the side stream — the
        # single-GPU train step uses it to overlap the optimizer update of finished buckets with the rest of the backward
        self.world = world if world is not None else dist.get_world_size(group)
        self.comm = torch.cuda.Stream() if pool.buf.is_cuda else None     # (CPU / gloo: same logic, no streams — used by the tests)
        # NCCL averages inside the collective (no per-bucket div_ launch); gloo has no AVG
        self.avg = self.world > 1 and pool.buf.is_cuda and dist.get_backend(group) == "nccl"
        self.after = None                  # callable(a, b): runs on the comm stream right after the range [a, b) has been reduced
        self.start = 0
        self.prev = 0
        self.n_coll = 0
        self.launched = []                 # [(a, b)] ranges reduced this step (tests / timelines)

    def begin(self) -> None:
        self.start = 0
        self.prev = 0
        self.n_coll = 0
        self.launched = []
        self.pool.on_commit = self._on_commit

    def _on_commit(self, off: int) -> None:
        # A bucket never includes the range of the node that is committing right now: that node has not returned yet, so
        # autograd may still read its pool slices on the main stream (AccumulateGrad clones a gradient that reaches its
        # parameter through a permute / reshape of the weight — Merge_Block, CARAFE.encoder, stem — and `.to(param dtype)`
        # runs after the commit).  Everything below the PREVIOUS commit's offset has had all of its consumers enqueued: the
        # engine runs the view-backward and AccumulateGrad nodes of a finished node before the next kernel-bearing node.
        if self.prev - self.start >= self.bucket:
            self._launch(self.start, self.prev)
            self.start = self.prev
        self.prev = off

    def _launch(self, a: int, b: int) -> None:
        seg = self.pool.buf[a:b]
        if self.comm is None:
            if self.world > 1:
                dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.group)
                seg.div_(self.world)
            if self.after is not None:
                self.after(a, b)
        else:
            self.comm.wait_stream(torch.cuda.current_stream())     # the kernels that filled / read [a, b) are already enqueued there
            with torch.cuda.stream(self.comm):
                if self.world > 1:
                    if self.avg:
                        dist.all_reduce(seg, op=dist.ReduceOp.AVG, group=self.group)
                    else:
                        dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.group)
                        seg.div_(self.world)
                if self.after is not None:
                    self.after(a, b)
        if self.world > 1:
            self.n_coll += 1
        self.launched.append((a, b))

    def reduce_outside_pool(self, grads) -> None:
        """Gradients that autograd materialised OUTSIDE the pool (a permuted clone of a re-packed conv weight's gradient): one
        grouped all-reduce on the comm stream, in place, no packing copies; `finish()` joins."""
        grads = [g for g in grads if g is not None]
        if self.world <= 1 or not grads:
            return
        if self.comm is None:
            for g in grads:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
                g.div_(self.world)
            self.n_coll += 1
            return
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            op = dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM
            with dist._coalescing_manager(group=self.group, device=grads[0].device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, op=op, group=self.group)
            if not self.avg:
                torch._foreach_div_(grads, float(self.world))
        self.n_coll += 1

    def finish(self) -> None:
        """After backward() has returned: every consumer of every pool slice is enqueued, reduce the tail and join."""
        self.pool.on_commit = None
        if self.pool.off > self.start:
            self._launch(self.start, self.pool.off)
            self.start = self.prev = self.pool.off
        if self.comm is not None:
            torch.cuda.current_stream().wait_stream(self.comm)

    def in_pool(self, t: Tensor) -> bool:
        lo = self.pool.buf.data_ptr()
        return lo <= t.data_ptr() < lo + self.pool.buf.numel() * 4
