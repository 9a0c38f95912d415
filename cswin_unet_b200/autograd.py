"""torch.autograd.Functions that route the backward of the hot path through the native backward kernels
(include/cswin_b200.h: cswin_lepe_attention_bwd, cswin_act_bwd, cswin_linear_wgrad, cswin_layernorm_bwd,
cswin_col2im_tokens, cswin_carafe_reassemble_bwd; the data gradient of a Linear is a forward Linear on W^T).

The reference has no backward source — it relies on torch.autograd through its eager ops (trainer.py:59).  Here
autograd is only the tape: every node's forward and backward is one of our kernels.  Parameter gradients are
accumulated in fp32 by the kernels and returned in the parameter's dtype (fp32 master weights + bf16 compute work).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops

Tensor = torch.Tensor


def needs_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(t is not None and torch.is_tensor(t) and t.requires_grad for t in ts)


import os as _os
_DGRAD_TRANSPOSE = _os.environ.get("CSWIN_DGRAD_TRANSPOSE") == "1"      # A/B switch: materialise W^T for the data gradient


# SHADOW[id(param)] = compute-dtype copy of the parameter, refreshed once per step by ONE multi-tensor copy
# (train.TrainStep) instead of a cast kernel per weight per use.  Empty outside a TrainStep.
SHADOW: dict = {}


class ZeroPool:
    """Bump allocator over ONE flat fp32 buffer that train.TrainStep zeroes once per step: the weight-gradient kernels
    accumulate into slices of it instead of ~460 separately zero-filled tensors."""

    def __init__(self, numel: int, device):
        self.buf = torch.zeros(numel, dtype=torch.float32, device=device)
        self.off = 0

        self.on_commit = None           # callable(offset): everything below `offset` has had its producing kernels launched

    def reset(self):
        self.buf.zero_()
        self.off = 0

    def commit(self):
        if self.on_commit is not None:
            self.on_commit(self.off)

    def take(self, shape) -> Optional[Tensor]:
        n = 1
        for s in shape:
            n *= s
        n4 = (n + 3) // 4 * 4
        if self.off + n4 > self.buf.numel():
            return None
        v = self.buf[self.off:self.off + n].view(shape)
        self.off += n4
        return v


POOL: Optional[ZeroPool] = None


def _commit() -> None:
    """Called at the end of every backward that filled pooled gradient slices (see parallel.PoolGradReducer)."""
    if POOL is not None:
        POOL.commit()


def _zeros(shape, device) -> Tensor:
    v = POOL.take(shape) if POOL is not None else None
    return v if v is not None else torch.zeros(shape, dtype=torch.float32, device=device)


def _tma_rows(t: Tensor) -> Tensor:
    """Gradient rows laid out for the tcgen05 kernels: unit column stride and, for bf16, a row pitch that is a multiple of
    16 bytes (N = 36 or 9 output channels would otherwise fall back to the SIMT dgrad / wgrad kernels)."""
    n = t.shape[-1]
    if t.dtype == torch.bfloat16 and n % 8 != 0:
        ld = (n + 7) // 8 * 8
        if not (t.dim() >= 2 and t.stride(-1) == 1 and t.stride(-2) == ld
                and _uniform_rows(t, ld)):
            buf = torch.empty(t.shape[:-1] + (ld,), dtype=t.dtype, device=t.device)[..., :n]
            buf.copy_(t)
            return buf
        return t
    return t.contiguous()


def _uniform_rows(t: Tensor, ld: int) -> bool:
    expect = ld
    for size, stride in zip(reversed(t.shape[:-1]), reversed(t.stride()[:-1])):
        if size != 1 and stride != expect:
            return False
        expect *= size
    return True


def _c(p: Optional[Tensor], dt: torch.dtype) -> Optional[Tensor]:
    if p is None:
        return None
    sh = SHADOW.get(id(p))
    if sh is not None and sh.dtype == dt:
        return sh
    p = p.detach()
    return p if (p.dtype == dt and p.is_contiguous()) else p.to(dt).contiguous()


PARAM_GRAD_SIDE_STREAM = _os.environ.get("CSWIN_PARAM_GRAD_SIDE_STREAM", "1") != "0"
_SIDE_STREAMS: dict = {}


def _side_stream(device) -> "torch.cuda.Stream":
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    s = _SIDE_STREAMS.get(key)
    if s is None:
        s = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return s


# Off by default: measured 6.03 vs 6.05 ms per train step — the wgrad CTA (4-stage ring, up to 193 KB of shared memory) and the
# dgrad Linear's CTAs do not fit on one SM together, so the two kernels take turns instead of overlapping, and the fork / join
# breaks the programmatic-dependent-launch chain of the main stream.  A 2-stage wgrad ring (97 KB, -DCSWIN_WGRAD_STAGES=2) does
# overlap, but is itself 0.33 ms per step slower: 6.07 ms with the fork vs 6.37 ms without, i.e. no net gain either.
WGRAD_SIDE_STREAM = _os.environ.get("CSWIN_WGRAD_SIDE_STREAM", "0") == "1"


class _Fork:
    """Runs the weight-gradient kernel(s) of a Linear on the second stream while the data-gradient GEMM runs on the main one: both
    need only dZ, neither fills the GPU for long (<= 1 wave each), and the optimizer / all-reduce are the only consumers of dW.
    `with _Fork(t) as f:` enqueues on the side stream after everything already enqueued on the main stream; `f.join()` makes the
    main stream wait for it (before ZeroPool.commit: an all-reduce may start right after)."""

    def __init__(self, like: Tensor, enabled: bool = True):
        self.on = enabled and WGRAD_SIDE_STREAM and like.is_cuda and like.dtype == torch.bfloat16
        self.side = self.cur = None

    def __enter__(self):
        if self.on:
            self.cur = torch.cuda.current_stream()
            self.side = _side_stream(self.cur.device)
            self.side.wait_stream(self.cur)
            self._ctx = torch.cuda.stream(self.side)
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.on:
            self._ctx.__exit__(*exc)
        return False

    def join(self) -> None:
        if self.on:
            self.cur.wait_stream(self.side)


class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        dt = x.dtype
        g, b = _c(gamma, dt), _c(beta, dt)
        y, mean, rstd = ops.layernorm(x, g, b, eps, stats=True)
        ctx.save_for_backward(x, g, mean, rstd)
        ctx.pd = gamma.dtype
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, g, mean, rstd = ctx.saved_tensors
        Cn = x.shape[-1]
        dx, dg, db = ops.layernorm_bwd(x, dy.contiguous(), g, mean, rstd, _zeros((Cn,), x.device), _zeros((Cn,), x.device))
        _commit()
        return dx, dg.to(ctx.pd), db.to(ctx.pd), None


class LayerNormForkFn(Function):
    """x -> (LayerNorm(x), x): the pre-norm residual pattern `x + f(LN(x))` (cswin_unet.py:178-179).  Handing out the residual
    operand here makes this node the only consumer of x, so the two gradients of x (through the LayerNorm and around it)
    meet inside the LayerNorm backward kernel's store instead of in a separate accumulation kernel of the autograd engine."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        dt = x.dtype
        g, b = _c(gamma, dt), _c(beta, dt)
        y, mean, rstd = ops.layernorm(x, g, b, eps, stats=True)
        ctx.save_for_backward(x, g, mean, rstd)
        ctx.pd = gamma.dtype
        return y, x.view_as(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, dres):
        x, g, mean, rstd = ctx.saved_tensors
        Cn = x.shape[-1]
        if dy is None:                                      # only the residual path was used
            return dres, None, None, None
        dx, dg, db = ops.layernorm_bwd(x, dy.contiguous(), g, mean, rstd, _zeros((Cn,), x.device), _zeros((Cn,), x.device),
                                       dx_add=None if dres is None else dres.contiguous())
        _commit()
        return dx, dg.to(ctx.pd), db.to(ctx.pd), None


class LinearFn(Function):
    """out = residual + sample_scale[row // rps] * ([a | a2] @ w.T + bias)   (no activation: see GeluFn)."""

    @staticmethod
    def forward(ctx, a, w, bias, a2, residual, sample_scale, rps):
        dt = a.dtype
        wc, bc = _c(w, dt), _c(bias, dt)
        out = ops.linear(a, wc, bc, a2=a2, residual=residual, sample_scale=sample_scale, rows_per_sample=rps)
        ctx.save_for_backward(a, a2, wc, sample_scale)
        ctx.rps, ctx.has_bias, ctx.has_res = rps, bias is not None, residual is not None
        ctx.wd = w.dtype
        ctx.bd = bias.dtype if bias is not None else None
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        a, a2, wc, ss = ctx.saved_tensors
        dout = _tma_rows(dout)
        dz = dout if ss is None else ops.act_bwd(dout, None, ss, ctx.rps, act=0)
        K1 = a.shape[-1]
        N, K = wc.shape
        need = ctx.needs_input_grad
        dw = db = dwf = dbf = None
        fk = None
        if need[1] or need[2]:                                     # dW on the second stream, next to dA below
            dwf = _zeros((N, K), dz.device)
            dbf = _zeros((N,), dz.device) if ctx.has_bias else None
            with _Fork(dz) as fk:
                ops.linear_wgrad(dz, a, dwf[:, :K1], dbf)
                if a2 is not None:
                    ops.linear_wgrad(dz, a2, dwf[:, K1:], None)
        # dA = dZ @ W: a forward Linear that reads W (N, K) in place as (K', N')
        if _DGRAD_TRANSPOSE:                                       # A/B switch: materialise W^T instead
            wt = wc.t().contiguous()
            da = ops.linear(dz, wt[:K1]) if need[0] else None
            da2 = ops.linear(dz, wt[K1:]) if (a2 is not None and need[3]) else None
        else:
            da = ops.linear(dz, wc[:, :K1], w_kn=True) if need[0] else None
            da2 = ops.linear(dz, wc[:, K1:], w_kn=True) if (a2 is not None and need[3]) else None
        if fk is not None:
            fk.join()
            dw = dwf.to(ctx.wd)
            db = dbf.to(ctx.bd) if ctx.has_bias else None
            _commit()
        return da, dw, db, da2, (dout if ctx.has_res else None), None, None


class MlpFn(Function):
    """out = residual + sample_scale * fc2(GELU(fc1(u)))  (Mlp, cswin_unet.py:12-28 + the residual / DropPath of :179) with the
    activation work inside the GEMM epilogues: fc1 writes z and GELU(z) from one accumulator read (aux_out), and the backward's
    dZ = dH o GELU'(z) is the epilogue of the fc2 data-gradient GEMM (act 2) — no separate activation passes either way."""

    @staticmethod
    def forward(ctx, u, w1, b1, w2, b2, residual, sample_scale, rps):
        dt = u.dtype
        w1c, b1c, w2c, b2c = _c(w1, dt), _c(b1, dt), _c(w2, dt), _c(b2, dt)
        z = torch.empty(u.shape[:-1] + (w1c.shape[0],), dtype=dt, device=u.device)
        h = ops.linear(u, w1c, b1c, act=1, aux_out=z)
        out = ops.linear(h, w2c, b2c, residual=residual, sample_scale=sample_scale, rows_per_sample=rps)
        ctx.save_for_backward(u, z, h, w1c, w2c, sample_scale)
        ctx.rps, ctx.has_res = rps, residual is not None
        ctx.dts = (w1.dtype, b1.dtype, w2.dtype, b2.dtype)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        u, z, h, w1c, w2c, ss = ctx.saved_tensors
        dout = _tma_rows(dout)
        dz2 = dout if ss is None else ops.act_bwd(dout, None, ss, ctx.rps, act=0)
        dw2f, db2f = _zeros(tuple(w2c.shape), dout.device), _zeros((w2c.shape[0],), dout.device)
        with _Fork(dz2) as f2:                                           # the weight gradients run on the second stream,
            ops.linear_wgrad(dz2, h, dw2f, db2f)                         # next to the data-gradient GEMMs
        dz = ops.linear(dz2, w2c, w_kn=True, act=2, residual=z)          # dH o GELU'(z)
        dw1f, db1f = _zeros(tuple(w1c.shape), dout.device), _zeros((w1c.shape[0],), dout.device)
        with _Fork(dz) as f1:
            ops.linear_wgrad(dz, u, dw1f, db1f)
        du = ops.linear(dz, w1c, w_kn=True) if ctx.needs_input_grad[0] else None
        f2.join(); f1.join()
        _commit()
        d = ctx.dts
        return du, dw1f.to(d[0]), db1f.to(d[1]), dw2f.to(d[2]), db2f.to(d[3]), (dout if ctx.has_res else None), None, None


def mlp(u, w1, b1, w2, b2, residual=None, sample_scale=None, rps=0):
    """MlpFn when the tcgen05 training epilogues apply (bf16, aligned rows), else the composed Linear / GELU / Linear nodes."""
    if b1 is not None and b2 is not None and ops.train_epilogues_supported(u, w1.shape[0]) and w2.shape[0] % 8 == 0 and FUSE_TRAIN_GELU:
        return MlpFn.apply(u, w1, b1, w2, b2, residual, sample_scale, rps)
    hid = GeluFn.apply(linear(u, w1, b1))
    return linear(hid, w2, b2, residual=residual, sample_scale=sample_scale, rps=rps)


FUSE_TRAIN_GELU = _os.environ.get("CSWIN_FUSE_TRAIN_GELU", "1") != "0"


class GeluFn(Function):
    @staticmethod
    def forward(ctx, z):
        ctx.save_for_backward(z)
        return ops.act_fwd(z, act=1)

    @staticmethod
    @once_differentiable
    def backward(ctx, dh):
        (z,) = ctx.saved_tensors
        return ops.act_bwd(dh.contiguous(), z, None, 0, act=1)


class LepeAttentionFn(Function):
    """(B, L, 3C) qkv -> (B, L, C): one or two stripe branches (cswin_unet.py:172-176 + :82-109)."""

    @staticmethod
    def forward(ctx, qkv, cw0, cb0, cw1, cb1, meta):
        B, L, C3 = qkv.shape
        Cn = C3 // 3
        dt = qkv.dtype
        out = torch.empty((B, L, Cn), dtype=dt, device=qkv.device)
        ws = [(_c(cw0, dt), _c(cb0, dt))] + ([(_c(cw1, dt), _c(cb1, dt))] if cw1 is not None else [])
        # per-row log-sum-exp of the scaled scores, (B, L, heads_b) fp32 per branch: saved for the backward kernel
        lses = [torch.empty((B, L, h), dtype=torch.float32, device=qkv.device) for h in meta["heads"]]
        ops.lepe_attention_fwd(LepeAttentionFn._descs(qkv, out, ws, meta, Cn, lambda i, sl: dict(lse=lses[i])), B,
                               meta["reso"], meta["scale"], dt)
        ctx.save_for_backward(qkv, *[t for pair in ws for t in pair], *lses)
        ctx.meta, ctx.pd = meta, cw0.dtype
        return out

    @staticmethod
    def _descs(qkv, out, ws, meta, Cn, extra=None):
        q, k, v = qkv[..., :Cn], qkv[..., Cn:2 * Cn], qkv[..., 2 * Cn:]
        nb = len(ws)
        h = Cn // nb
        descs = []
        for i, (cw, cb) in enumerate(ws):
            sl = slice(i * h, (i + 1) * h)
            d = dict(q=q[..., sl], k=k[..., sl], v=v[..., sl], out=out[..., sl], conv_w=cw, conv_b=cb,
                     heads=meta["heads"][i], H_sp=meta["win"][i][0], W_sp=meta["win"][i][1])
            if extra is not None:
                d.update(extra(i, sl))
            descs.append(d)
        return descs

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        meta = ctx.meta
        nbr = len(meta["heads"])
        qkv, *rest = ctx.saved_tensors
        wflat, lses = rest[:2 * nbr], rest[2 * nbr:]
        ws = [(wflat[2 * i], wflat[2 * i + 1]) for i in range(nbr)]
        B, L, C3 = qkv.shape
        Cn = C3 // 3
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        dq, dk, dv = dqkv[..., :Cn], dqkv[..., Cn:2 * Cn], dqkv[..., 2 * Cn:]
        h = Cn // len(ws)
        gw = [_zeros((h, 9), qkv.device) for _ in ws]
        gb = [_zeros((h,), qkv.device) for _ in ws]

        def extra(i, sl):
            return dict(dout=dout[..., sl], dq=dq[..., sl], dk=dk[..., sl], dv=dv[..., sl], dconv_w=gw[i], dconv_b=gb[i],
                        lse=lses[i])
        descs = LepeAttentionFn._descs(qkv, dout, ws, meta, Cn, extra)
        # d get_v.{weight,bias} need only dout and v: they run as their own streaming kernel on a second stream, next to the
        # attention backward kernel (which is latency-bound and leaves most issue slots idle)
        side = None
        if PARAM_GRAD_SIDE_STREAM and qkv.dtype == torch.bfloat16:
            cur = torch.cuda.current_stream()
            side = _side_stream(qkv.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                if not ops.lepe_param_grad(descs, B, meta["reso"], qkv.dtype):
                    side = None
        ops.lepe_attention_bwd(descs, B, meta["reso"], meta["scale"], qkv.dtype, param_grads=side is None)
        if side is not None:
            cur.wait_stream(side)
        _commit()
        g0 = (gw[0].view(h, 1, 3, 3).to(ctx.pd), gb[0].to(ctx.pd))
        g1 = (gw[1].view(h, 1, 3, 3).to(ctx.pd), gb[1].to(ctx.pd)) if len(ws) == 2 else (None, None)
        return dqkv, g0[0], g0[1], g1[0], g1[1], None


class Im2colTokensFn(Function):
    @staticmethod
    def forward(ctx, x, H, W, KH, KW, stride, pad):
        ctx.geom = (x.shape[0], H, W, x.shape[-1], KH, KW, stride, pad)
        return ops.im2col_tokens(x, H, W, KH, KW, stride, pad)

    @staticmethod
    @once_differentiable
    def backward(ctx, dcol):
        B, H, W, Cn, KH, KW, stride, pad = ctx.geom
        return ops.col2im_tokens(dcol, B, H, W, Cn, KH, KW, stride, pad), None, None, None, None, None, None


class CarafeReassembleFn(Function):
    """enc (B*H*W, 9 up^2), z (B*H*W, C), bias (C) -> (B, up^2 H W, C)."""

    @staticmethod
    def forward(ctx, enc, z, bias, B, H, W, up):
        dt = z.dtype
        y = ops.carafe_reassemble(enc, z, _c(bias, dt), B, H, W, up)
        ctx.save_for_backward(enc, z)
        ctx.geom, ctx.bd = (B, H, W, up), bias.dtype
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        enc, z = ctx.saved_tensors
        B, H, W, up = ctx.geom
        denc, dz, dbias = ops.carafe_reassemble_bwd(enc, z, dy, B, H, W, up)
        return denc, dz, dbias.to(ctx.bd), None, None, None, None


class CarafeHeadFn(Function):
    """Folded segmentation head on the tape: enc (M, 144), z (M, zcols >= classes), folded bias (classes) -> fp32 NCHW logits
    (cswin_unet.py:536-544 with CARAFE4.out and `output` folded into z's producer)."""

    @staticmethod
    def forward(ctx, enc, z, bias, B, H, W, up, n_classes):
        logits, _ = ops.carafe_head(enc, z, bias.to(z.dtype), B, H, W, up, want_logits=True, logits_dtype=torch.float32,
                                    n_classes=n_classes)
        ctx.save_for_backward(enc, z)
        ctx.geom, ctx.bd = (B, H, W, up, n_classes), bias.dtype
        return logits

    @staticmethod
    @once_differentiable
    def backward(ctx, dlogits):
        enc, z = ctx.saved_tensors
        B, H, W, up, nc = ctx.geom
        denc, dz, dbias = ops.carafe_head_bwd(enc, z, dlogits, B, H, W, up, nc)
        return denc, dz, dbias.to(ctx.bd), None, None, None, None, None


class SegLossFn(Function):
    """w_ce * CE + w_dice * Dice(softmax) on fp32 NCHW logits (trainer.py:55-57, utils.py:9-45): one native pass each way.

    `group` (a torch.distributed group, or True for the default group) selects the reference's GLOBAL-batch Dice: the
    reference's nn.DataParallel gathers the logits and forms Dice — a ratio of sums — over the whole batch (trainer.py:37-38,
    :55-57), so the 3 x classes Dice sums are all-reduced between the two passes (one tiny collective, capturable) and the
    Dice part of the local gradient is scaled by the world size (the gradient all-reduce AVERAGES over ranks, while the global
    Dice gradient is the SUM of the per-rank contributions).  CE is a mean over local pixels, which averages correctly."""

    @staticmethod
    def forward(ctx, logits, labels, w_ce, w_dice, group=None):
        sums = ops.seg_loss_fwd(logits, labels)
        nc = logits.shape[1]
        npix = labels.numel()
        world = 1
        if group is not None and torch.distributed.is_initialized():
            g = None if group is True else group
            world = torch.distributed.get_world_size(g)
            if world > 1:
                torch.distributed.all_reduce(sums[1:], op=torch.distributed.ReduceOp.SUM, group=g)
        inter, z, y = sums[1:1 + nc], sums[1 + nc:1 + 2 * nc], sums[1 + 2 * nc:1 + 3 * nc]
        loss = w_ce * sums[0] / npix + w_dice * (1.0 - (2 * inter + 1e-5) / (z + y + 1e-5)).mean()
        ctx.save_for_backward(logits, labels, sums)
        ctx.w = (w_ce, w_dice * world)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        logits, labels, sums = ctx.saved_tensors
        return ops.seg_loss_bwd(logits, labels, sums, gout, *ctx.w), None, None, None, None


def linear(a, w, bias=None, a2=None, residual=None, sample_scale=None, rps=0):
    return LinearFn.apply(a, w, bias, a2, residual, sample_scale, rps)
