"""Data-parallel training step of the reference's baseline trainer (trainer.py:42-60) on the native kernels.

    step = TrainStep(model, lr=0.05)            # SGD(momentum .9, weight decay 1e-4), loss 0.4 CE + 0.6 Dice
    loss = step(images, labels)                 # forward + backward (native kernels) + gradient all-reduce + SGD update

One process per GPU; with torch.distributed initialised the only exchange is the bucketed NCCL all-reduce of the
23.57 M gradients (`parallel.allreduce_gradients`) — the reference uses single-process nn.DataParallel
(trainer.py:37-38).  Loss (cswin_seg_loss_fwd/bwd) and optimizer (cswin_sgd_momentum_step) are native too (SURVEY 8f
rank 3); the Dice term is formed per rank (a ratio of sums over the LOCAL batch), as SURVEY 8e notes.
The reference's logging `.item()` calls (>= 11 host syncs per step) are not reproduced: the step returns a device tensor.
"""
from __future__ import annotations

from typing import Optional

import os

import torch
import torch.nn.functional as F

from . import autograd as ag
from . import modules
from . import ops
from . import parallel

Tensor = torch.Tensor
NATIVE_LOSS = os.environ.get("CSWIN_TORCH_LOSS") != "1"


def seg_loss(logits: Tensor, target: Tensor, n_classes: int, global_dice_group=None) -> Tensor:
    """0.4 * CrossEntropy + 0.6 * DiceLoss(softmax=True)   (trainer.py:55-57, utils.py:9-45).

    `global_dice_group` (a process group, or True for the default one): form Dice over the GLOBAL batch like the reference's
    nn.DataParallel does (the Dice sums are all-reduced; see autograd.SegLossFn).  Default None = per-rank Dice."""
    logits = logits.float()
    if (logits.is_cuda and NATIVE_LOSS and ops.seg_loss_supported(n_classes) and logits.shape[1] == n_classes
            and target.dtype in (torch.uint8, torch.int32, torch.int64)):
        return ag.SegLossFn.apply(logits.contiguous(), target.contiguous(), 0.4, 0.6, global_dice_group)   # native: one pass each way
    ce = F.cross_entropy(logits, target.long())
    prob = torch.softmax(logits, dim=1)
    # (F.one_hot validates its input with a host sync, which would break CUDA-graph capture of the step)
    onehot = (target.long().unsqueeze(1) == torch.arange(n_classes, device=target.device).view(1, -1, 1, 1)).to(prob.dtype)
    dims = (0, 2, 3)
    inter = (prob * onehot).sum(dims)
    zsum = (prob * prob).sum(dims)
    ysum = onehot.sum(dims)
    world = 1
    if global_dice_group is not None and torch.distributed.is_initialized():
        g = None if global_dice_group is True else global_dice_group
        world = torch.distributed.get_world_size(g)
        if world > 1:                                         # other ranks' sums enter as constants (their gradient lives there)
            packed = torch.stack([inter, zsum, ysum]).detach().clone()
            torch.distributed.all_reduce(packed, group=g)
            inter = inter + (packed[0] - inter.detach())
            zsum = zsum + (packed[1] - zsum.detach())
            ysum = ysum + (packed[2] - ysum.detach())
    dice = (1.0 - (2 * inter + 1e-5) / (zsum + ysum + 1e-5)).mean()
    if world > 1:                                             # value unchanged, gradient x world (the all-reduce averages)
        dice = dice.detach() + world * (dice - dice.detach())
    return 0.4 * ce + 0.6 * dice


class TrainStep:
    """forward + backward + gradient all-reduce + SGD.  With `graph=True` (default) the whole step is captured once into a
    CUDA graph (static input buffers; ~2000 launches per step would otherwise be bound by Python launch overhead) and
    replayed; the NCCL all-reduce stays outside the graph (eager, bucketed) when more than one rank trains."""

    def __init__(self, model, lr: float = 0.05, momentum: float = 0.9, weight_decay: float = 1e-4,
                 compute_dtype: torch.dtype = torch.bfloat16, group=None, graph: bool = True, warmup: int = 3,
                 batched_drop_path: bool = True, global_dice: bool = False, distributed: Optional[bool] = None):
        """global_dice: Dice over the global batch as the reference's nn.DataParallel computes it (all-reduce of the Dice sums);
        default = per-rank Dice.  distributed: None = follow torch.distributed; False = never all-reduce (used by the
        N-rank gradient-parity check to compute per-rank gradients inside an initialised process group)."""
        self.model = model.train()
        self.model.compute_dtype = compute_dtype
        self.n_classes = model.num_classes
        self.group = group
        self.momentum, self.weight_decay = momentum, weight_decay
        # torch.optim.SGD is kept for zero_grad() / the non-fp32-parameter fallback; the step itself is ONE native launch over
        # every parameter (cswin_sgd_momentum_step: weight decay + momentum + update + bf16 shadow refresh)
        self.opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=momentum, weight_decay=weight_decay)
        self._params = [p for p in model.parameters() if p.requires_grad]
        self._native_sgd = os.environ.get("CSWIN_TORCH_SGD") != "1" and all(
            p.dtype == torch.float32 and p.is_cuda and p.is_contiguous() for p in self._params)
        dev = self._params[0].device
        self.lr_dev = torch.tensor([lr], dtype=torch.float32, device=dev)
        self.momentum_buffers = [torch.zeros_like(p) for p in self._params] if self._native_sgd else []
        n_chunks = sum((p.numel() + ops.SGD_CHUNK - 1) // ops.SGD_CHUNK for p in self._params)
        self._tbl_host = torch.empty((n_chunks, 5), dtype=torch.int64).pin_memory() if self._native_sgd else None
        self._tbl_dev = torch.empty((n_chunks, 5), dtype=torch.int64, device=dev) if self._native_sgd else None
        self._tbl_key, self._tbl_n = None, 0
        self._versions = None
        self.use_graph = graph
        self.warmup = warmup
        self._graphs = None
        self.native_launches_per_step = 0
        self._static = None
        self._seen = 0
        self._make_shadow(compute_dtype)
        # every DropPath mask of a step from one Bernoulli draw (modules.DropPathPlan); False = timm's per-call RNG consumption
        self._drop_plan = modules.DropPathPlan() if (batched_drop_path and os.environ.get("CSWIN_DROPPATH_PER_CALL") != "1") else None
        self._pool = None if os.environ.get("CSWIN_NO_POOL") == "1" else ag.ZeroPool(
            sum(p.numel() + 4 for p in model.parameters()), next(model.parameters()).device)
        self._distributed = torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size(group) > 1
        if distributed is not None:
            self._distributed = self._distributed and distributed
        self._dice_group = (group if group is not None else True) if (global_dice and self._distributed) else None
        # all-reduce overlapped with the backward on the pooled gradient buffer (CSWIN_DDP_OVERLAP=0: bucketed all-reduce
        # after the backward, between two CUDA graphs)
        self._reducer = parallel.PoolGradReducer(self._pool, group, bucket_bytes=int(float(os.environ.get("CSWIN_DDP_BUCKET_MB", "16")) * (1 << 20))) if (
            self._distributed and self._pool is not None and os.environ.get("CSWIN_DDP_OVERLAP", "1") != "0") else None
        # Optional (CSWIN_BUCKET_SGD=1): optimizer update per gradient bucket, on the reducer's side stream right after the bucket's
        # all-reduce, so that neither the ~95 us fused SGD pass nor (N > 1) the wait for the last all-reduce sits at the end of the
        # step.  Safe: a bucket holds gradients whose layers' backward kernels are already enqueued (PoolGradReducer), and nothing
        # later in the backward reads those layers' parameters.  MEASURED slower on B200 — 1 GPU 6.12 vs 6.04 ms, 2 GPUs 6.34 vs
        # 6.25 ms (profiles/r02_ddp_suite_2gpu.log): the update kernels then share HBM / SMs with the latency-bound backward
        # kernels on the critical path instead of running alone at 5.5 TB/s — so the default stays ONE launch after the backward.
        env = os.environ.get("CSWIN_BUCKET_SGD", "0")
        self._bucket_sgd = (self._native_sgd and self._pool is not None and next(model.parameters()).is_cuda and env == "1"
                            and (self._reducer is not None or not self._distributed))
        if self._bucket_sgd and self._reducer is None:
            self._reducer = parallel.PoolGradReducer(self._pool, None, world=1)
        self._bk = None                      # bucket-SGD tables, built from the first backward's gradient layout

    def _make_shadow(self, dtype: torch.dtype) -> None:
        ps = [p for p in self.model.parameters() if p.dtype != dtype]
        self._shadow_src = ps
        n = sum(p.numel() + 8 for p in ps)                       # every view 16-byte aligned
        flat = torch.empty(n, dtype=dtype, device=ps[0].device) if ps else None
        self._shadow_views, off = [], 0
        for p in ps:
            self._shadow_views.append(flat[off:off + p.numel()].view(p.shape))
            off += (p.numel() + 7) // 8 * 8
        self._shadow_of = {id(p): v for p, v in zip(ps, self._shadow_views)}

    @property
    def lr(self) -> float:
        return self.opt.param_groups[0]["lr"]

    @lr.setter
    def lr(self, value: float) -> None:
        """Learning-rate schedule hook (trainer.py:63-66): takes effect at the next step, also under CUDA-graph replay."""
        for g in self.opt.param_groups:
            g["lr"] = float(value)
        self.lr_dev.fill_(float(value))

    def _refresh_shadow(self) -> None:
        if self._shadow_src:
            torch._foreach_copy_(self._shadow_views, [p.detach() for p in self._shadow_src])

    def _optimizer_step(self) -> None:
        if getattr(self, "_sgd_in_backward", False):           # already done, bucket by bucket, inside the backward
            return
        if not self._native_sgd:
            self.opt.step()
            return
        ps = [(p, m) for p, m in zip(self._params, self.momentum_buffers) if p.grad is not None]
        key = tuple(p.grad.data_ptr() for p, _ in ps)
        if key != self._tbl_key:                               # gradient buffers moved (always static under graph replay)
            tbl = ops.sgd_chunk_table([p.detach() for p, _ in ps], [p.grad.contiguous() for p, _ in ps], [m for _, m in ps],
                                      [self._shadow_of.get(id(p)) for p, _ in ps])
            self._tbl_n = tbl.shape[0]
            self._tbl_host[:self._tbl_n].copy_(tbl)
            self._tbl_key = key
        self._tbl_dev.copy_(self._tbl_host, non_blocking=True)
        ops.sgd_momentum_step(self._tbl_dev[:self._tbl_n], self.lr_dev, self.momentum, self.weight_decay)
        modules.bump_param_epoch()                             # raw-pointer write: invalidate every derived-weight cache

    # ---- optimizer update per reduced bucket (see __init__) ------------------------------------------------------------------
    def _build_bucket_tables(self) -> None:
        """After a backward: split the parameters into those whose gradient lives in the pool (sorted by pool offset; their chunk
        rows go to one device table that `_sgd_range` slices) and the rest (one small table)."""
        import numpy as np
        base = self._pool.buf.data_ptr()
        inp, rest = [], []
        for p, m in zip(self._params, self.momentum_buffers):
            if p.grad is None:
                continue
            g = p.grad
            (inp if (self._reducer.in_pool(g) and g.is_contiguous()) else rest).append((g.data_ptr(), p, g, m))
        inp.sort(key=lambda t: t[0])

        def table(rows):
            if not rows:
                return None, None
            t = ops.sgd_chunk_table([r[1].detach() for r in rows], [r[2] for r in rows], [r[3] for r in rows],
                                    [self._shadow_of.get(id(r[1])) for r in rows])
            return t.to(self._pool.buf.device), (t[:, 1].numpy() - base) // 4
        tp, offs = table(inp)
        n_rest = sum((r[1].numel() + ops.SGD_CHUNK - 1) // ops.SGD_CHUNK for r in rest)
        self._bk = {"pool": tp, "offs": np.asarray(offs) if offs is not None else None,
                    "rest_params": [(r[1], r[3]) for r in rest],
                    "rest_host": torch.empty((max(n_rest, 1), 5), dtype=torch.int64).pin_memory(),
                    "rest_dev": torch.empty((max(n_rest, 1), 5), dtype=torch.int64, device=self._pool.buf.device), "n_rest": n_rest,
                    # pooled gradients keep their addresses from step to step (same backward order); the few gradients autograd
                    # clones outside the pool are fresh allocations every step and get their (tiny) table rebuilt below
                    "key": tuple(t[0] for t in inp)}

    def _sgd_rest(self) -> None:
        """Update of the parameters whose gradient lives outside the pool (rebuilt table: their gradient tensors move)."""
        bk = self._bk
        if not bk["rest_params"]:
            return
        ps = [(p, m) for p, m in bk["rest_params"] if p.grad is not None]
        t = ops.sgd_chunk_table([p.detach() for p, _ in ps], [p.grad.contiguous() for p, _ in ps], [m for _, m in ps],
                                [self._shadow_of.get(id(p)) for p, _ in ps])
        n = t.shape[0]
        assert n <= bk["rest_host"].shape[0]
        bk["rest_host"][:n].copy_(t)
        bk["rest_dev"].copy_(bk["rest_host"], non_blocking=True)
        ops.sgd_momentum_step(bk["rest_dev"][:n], self.lr_dev, self.momentum, self.weight_decay)

    def _sgd_range(self, a: int, b: int) -> None:
        """SGD update of every parameter whose (pooled) gradient lies in pool range [a, b): runs on the reducer's side stream."""
        import numpy as np
        bk = self._bk
        if bk["pool"] is None:
            return
        lo, hi = int(np.searchsorted(bk["offs"], a, "left")), int(np.searchsorted(bk["offs"], b, "left"))
        if hi > lo:
            ops.sgd_momentum_step(bk["pool"][lo:hi], self.lr_dev, self.momentum, self.weight_decay)

    def _fwd_bwd(self, images: Tensor, labels: Tensor) -> Tensor:
        if self._shadow_src:
            if not self._native_sgd:                             # one multi-tensor fp32 -> compute-dtype copy per step
                self._refresh_shadow()                           # (the native SGD kernel refreshes the shadows itself)
            ag.SHADOW = self._shadow_of
        if self._pool is not None:
            self._pool.reset()
            ag.POOL = self._pool
        self._sgd_in_backward = False
        if self._reducer is not None:
            self._reducer.begin()
            self._reducer.after = None
            if self._bucket_sgd and self._bk is not None and not getattr(self, "_grads_only", False):
                self._reducer.after = self._sgd_range
                self._sgd_in_backward = True
        if self._drop_plan is not None:
            self._drop_plan.begin(images.shape[0], images.device)
            modules.DROP_PATH_PLAN = self._drop_plan
        try:
            logits = self.model(images)
            if self._drop_plan is not None:
                self._drop_plan.end()
            modules.DROP_PATH_PLAN = None
            loss = seg_loss(logits, labels, self.n_classes, self._dice_group)
            loss.backward()
            if self._reducer is not None:                       # gradients outside the pool, then the tail bucket, then join
                rest = [p for p in self.model.parameters() if p.grad is not None and not self._reducer.in_pool(p.grad)]
                self._reducer.reduce_outside_pool([p.grad for p in rest])
                if self._sgd_in_backward:
                    key = tuple(sorted(p.grad.data_ptr() for p in self._params if p.grad is not None and self._reducer.in_pool(p.grad)
                                       and p.grad.is_contiguous()))
                    if key != self._bk["key"]:
                        raise RuntimeError("TrainStep: the pooled gradient layout changed between steps while the optimizer update runs "
                                           "per bucket inside the backward (set CSWIN_BUCKET_SGD=0 for models whose backward order varies)")
                    cs = self._reducer.comm                      # on the side stream, after the grouped all-reduce of `rest`
                    cs.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(cs):
                        self._sgd_rest()
                self._reducer.finish()
                if self._sgd_in_backward:
                    modules.bump_param_epoch()
                elif self._bucket_sgd and self._bk is None:
                    self._build_bucket_tables()                 # layout known now: later steps update per bucket
        finally:
            modules.DROP_PATH_PLAN = None
            ag.SHADOW = {}
            ag.POOL = None
            if self._pool is not None:
                self._pool.on_commit = None
        return loss.detach()

    def gradients(self, images: Tensor, labels: Tensor) -> Tensor:
        """Forward + backward (+ the gradient all-reduce when distributed) WITHOUT the optimizer step: returns the loss and
        leaves the (averaged) gradients in `p.grad`.  Eager; used by the N-rank gradient-parity check."""
        self._check_external_writes()
        self.opt.zero_grad(set_to_none=True)
        self._grads_only = True
        try:
            loss = self._fwd_bwd(images, labels)
        finally:
            self._grads_only = False
        if self._reducer is None and self._distributed:
            parallel.allreduce_gradients(self.model.parameters(), self.group)
        return loss

    def _eager(self, images: Tensor, labels: Tensor) -> Tensor:
        self.opt.zero_grad(set_to_none=True)
        loss = self._fwd_bwd(images, labels)
        if self._reducer is None and self._distributed:
            parallel.allreduce_gradients(self.model.parameters(), self.group)
        self._optimizer_step()
        return loss

    def _capture(self, images: Tensor, labels: Tensor) -> None:
        self._static = (images.clone(), labels.clone())
        sx, sy = self._static
        self.opt.zero_grad(set_to_none=True)
        from ._lib import launch_count
        n0 = launch_count()
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):                         # forward + backward (+ optimizer when single-rank)
            self._loss = self._fwd_bwd(sx, sy)
            if not self._distributed or self._reducer is not None:
                self._optimizer_step()                     # (overlapped DDP: the NCCL all-reduces are part of this graph)
        g2 = None
        if self._distributed and self._reducer is None:    # all-reduce eagerly between the two graphs
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, pool=g1.pool()):
                self._optimizer_step()
        self._graphs = (g1, g2)
        self.native_launches_per_step = launch_count() - n0        # launches of this library recorded into the graph(s)

    def _check_external_writes(self) -> None:
        """The native step writes parameters through raw pointers (no version bump): a changed `_version` means someone else
        (load_state_dict, TPGM projection, ...) wrote them since the last step, so the bf16 shadows are refreshed first."""
        if not (self._native_sgd and self._shadow_src):
            return
        vers = [p._version for p in self._shadow_src]
        if vers != self._versions:
            self._refresh_shadow()
            self._versions = vers

    def close(self) -> None:
        """Drop the captured CUDA graphs (and with them the captured NCCL all-reduces).  Call before
        torch.distributed.destroy_process_group(): tearing a communicator down while graphs that contain its collectives
        are alive can block forever."""
        self._graphs = None
        self._static = None
        self._loss = None
        if self._reducer is not None:
            self._pool.on_commit = None
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def __call__(self, images: Tensor, labels: Tensor) -> Tensor:
        self._check_external_writes()
        if not self.use_graph:
            return self._eager(images, labels)
        if self._graphs is None:
            if self._seen < self.warmup:                   # eager warm-up steps (allocator, lazy state, momentum buffers),
                self._seen += 1                            # on a side stream as CUDA-graph capture of autograd requires
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    loss = self._eager(images, labels)
                torch.cuda.current_stream().wait_stream(side)
                return loss
            self._capture(images, labels)
        sx, sy = self._static
        sx.copy_(images, non_blocking=True)
        sy.copy_(labels, non_blocking=True)
        g1, g2 = self._graphs
        g1.replay()
        if g2 is not None:
            parallel.allreduce_gradients(self.model.parameters(), self.group)
            g2.replay()
        modules.bump_param_epoch()                         # the replayed optimizer kernel wrote the parameters (no Python ran)
        return self._loss
