"""Drop-in nn.Modules for the CSWin-UNet hot path, backed by libcswin_b200.so.

Same class names, constructor signatures / defaults, public attributes, sub-module names and
state_dict keys as the reference classes in /root/reference/networks/cswin_unet.py
(`LePEAttention` :31-109, `CSWinBlock` :112-181, `Mlp` :12-28, `Merge_Block` :205-220,
`CARAFE` :222-269, `CARAFE4` :272-319, free functions `img2windows` :184-191 / `windows2img`
:194-202), so `load_state_dict(strict=True)` works both ways and `install()` (install.py) can
rebind them inside the reference's own module.  Parameters are ordinary nn.Parameters held by
nn.Linear / nn.Conv2d / nn.LayerNorm containers that are never *called*: every forward goes
through the C ABI.  Compute dtype = dtype of the incoming activation (float32 -> exact SIMT
path, bfloat16 -> tcgen05 path); parameters of another dtype are cast through a cache keyed on
`param._version` (re-derived every call in training mode, where in-place `.data` writes are
common and do not bump the version).
"""
from __future__ import annotations

import math
import os
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import autograd as ag
from . import ops

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------
class DropPathPlan:
    """All stochastic-depth masks of one training step from TWO launches instead of three per DropPath call (152 -> 2 at
    cswin_tiny: bernoulli_, div_, cast per call are ~2 us each and sit on the step's critical path).  The first step under
    a plan records the keep probabilities in call order (every module still draws for itself); later steps draw one
    (calls, B) Bernoulli matrix up front and hand out its rows.  Statistically identical to timm's DropPath; the RNG
    stream is consumed in one draw per step instead of one per call, so masks differ from a per-call run with the same
    seed (plain module use, without a plan, keeps timm's per-call consumption)."""

    def __init__(self):
        self.keeps: List[Tuple[float, bool]] = []
        self.recording = True
        self.rows: Optional[Tensor] = None
        self._probs = self._inv = None
        self.i = 0

    def begin(self, batch: int, device) -> None:
        self.i = 0
        self.rows = None
        if self.recording or not self.keeps:
            return
        if self._probs is None or self._probs.device != torch.device(device):
            self._probs = torch.tensor([k for k, _ in self.keeps], dtype=torch.float32, device=device).view(-1, 1)
            self._inv = torch.tensor([(1.0 / k if (k > 0.0 and sc) else 1.0) for k, sc in self.keeps], dtype=torch.float32,
                                     device=device).view(-1, 1)
        self.rows = torch.bernoulli(self._probs.expand(-1, batch)) * self._inv

    def take(self, keep: float, scale_by_keep: bool, batch: int) -> Optional[Tensor]:
        if self.recording:
            self.keeps.append((keep, scale_by_keep))
            return None
        if self.rows is None or self.i >= len(self.keeps) or self.keeps[self.i] != (keep, scale_by_keep) or self.rows.shape[1] != batch:
            self.i += 1
            return None                                     # call sequence changed: this call draws for itself
        row = self.rows[self.i]
        self.i += 1
        return row

    def end(self) -> None:
        self.recording = False


DROP_PATH_PLAN: Optional[DropPathPlan] = None            # set by train.TrainStep around the forward


class DropPath(nn.Module):
    """Stochastic depth with timm's semantics and RNG consumption (one bernoulli_ of shape (B,1,..,1) in x.dtype)."""

    def __init__(self, drop_prob: float = 0., scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def sample_scale(self, x: Tensor) -> Optional[Tensor]:
        """fp32 (B,) per-sample factor m_b/(1-p), or None when inactive (eval or p == 0)."""
        if self.drop_prob == 0. or not self.training:
            return None
        keep = 1.0 - self.drop_prob
        if DROP_PATH_PLAN is not None:
            row = DROP_PATH_PLAN.take(keep, self.scale_by_keep, x.shape[0])
            if row is not None:
                return row
        m = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            m.div_(keep)
        return m.reshape(-1).float()

    def forward(self, x: Tensor) -> Tensor:      # only used if someone calls the module directly
        s = self.sample_scale(x)
        return x if s is None else x * s.to(x.dtype).view((-1,) + (1,) * (x.ndim - 1))

    def extra_repr(self):
        return f"drop_prob={round(self.drop_prob, 3):0.3f}"


# fc1 + GELU + fc2 + residual as one tcgen05 kernel (needs FOLD_LN) for dim <= FUSE_MLP_MAX_DIM.  Measured per block on B200
# (batch 24): dim 64 fused 71.3 us vs 76.8 us composed; dim 128 45.5 vs 44.3; dim 256 44.6 vs 38.8 -> default 64.
FUSE_MLP = os.environ.get("CSWIN_FUSE_MLP", "1") != "0"
FUSE_MLP_MAX_DIM = int(os.environ.get("CSWIN_FUSE_MLP_MAX_DIM", "64"))
# ... and for at most this many rows: the fused kernel shortens the latency chain of a small launch (batch 24: 75 k rows at stage 1),
# in the throughput regime the two composed Linears (16 epilogue warps per SM) are faster (batch 96, 301 k rows: 30.8 k -> 31.3 k slices/s)
FUSE_MLP_MAX_ROWS = int(os.environ.get("CSWIN_FUSE_MLP_MAX_ROWS", "150000"))
FOLD_LN = os.environ.get("CSWIN_FOLD_LN", "1") != "0"     # LayerNorm folded into the tcgen05 Linear epilogue (bf16 inference)
# Merge_Block / CARAFE.encoder convolutions as implicit GEMMs (cswin_conv_tokens_fwd: strided TMA boxes of the token image, no
# column matrix) where the channel count allows (C % 64 == 0); CSWIN_IMPLICIT_CONV=0 restores im2col + Linear (A/B switch)
IMPLICIT_CONV = os.environ.get("CSWIN_IMPLICIT_CONV", "1") != "0"
# [LN1 -> qkv -> both LePE attention branches] as one kernel (csrc/qkv_attn_tc.cu) where the block shape allows it.  Bit-identical
# to the composed path, one launch and the (B, L, 3C) qkv round trip fewer per block — and still SLOWER in the forward on B200
# (17,433 vs 18,984 slices/s, profiles/r02_sweep_streams.log): with 24 images there are only 192 (image, branch, head pair, window)
# work units at stage 3, each CTA runs GEMM -> epilogue -> two heads of attention back to back (14.3 us per CTA, in-kernel trace
# profiles/r02_trace_qa_s3.log) where the two separate launches spread the same work over 444 + 384 shorter CTAs (5.7 + 8.5 us).
# Default off; CSWIN_FUSE_QKV_ATTN=1 enables it.
FUSE_QKV_ATTN = os.environ.get("CSWIN_FUSE_QKV_ATTN", "0") == "1"


# Parameters can be written behind torch's back: the native fused SGD (train.TrainStep -> cswin_sgd_momentum_step) updates them
# through raw pointers and never bumps `param._version`.  Every such writer calls `bump_param_epoch()`; the epoch is part of the
# cache signature below (and of SliceEngine's captured-graph signature), so derived tensors are re-derived after training.
_PARAM_EPOCH = [0]


def bump_param_epoch() -> int:
    _PARAM_EPOCH[0] += 1
    return _PARAM_EPOCH[0]


def param_epoch() -> int:
    return _PARAM_EPOCH[0]


class _Derived:
    """Cache of tensors derived from parameters (dtype casts, re-packed conv weights)."""

    def __init__(self):
        self._d: Dict[str, Tuple[tuple, Tensor]] = {}

    def get(self, key: str, params, dtype: torch.dtype, fn: Optional[Callable] = None, fresh: bool = False) -> Tensor:
        if isinstance(params, Tensor):
            params = (params,)
        if fn is None and params[0].dtype == dtype and params[0].is_contiguous():
            return params[0].detach()                    # read the parameter in place: can never go stale
        sig = tuple((p._version, p.data_ptr(), p.device) for p in params) + (dtype, _PARAM_EPOCH[0])
        hit = self._d.get(key)
        if not fresh and hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            src = [p.detach() for p in params]
            t = fn(*src) if fn is not None else src[0]
            t = t.to(dtype).contiguous()
        self._d[key] = (sig, t)
        return t

    def clear(self):
        self._d.clear()


class _Native(nn.Module):
    """Base: owns a derived-weight cache that is not part of state_dict and is dropped on deepcopy / pickling."""

    def __init__(self):
        super().__init__()
        object.__setattr__(self, "_derived", _Derived())

    def _w(self, key: str, params, dtype, fn=None) -> Tensor:
        return self._derived.get(key, params, dtype, fn, fresh=self.training)

    def train(self, mode: bool = True):
        # `.data` writes (TPGM: universal_train.py:380-388, 451-453) and raw-pointer optimizers do not bump `_version`; they
        # happen in training mode, so a train <-> eval switch drops the derived tensors and invalidates captured engines
        if mode != self.training:
            self._derived.clear()
            bump_param_epoch()
        return super().train(mode)

    def __getstate__(self):
        st = self.__dict__.copy()
        st.pop("_derived", None)
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        object.__setattr__(self, "_derived", _Derived())

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k != "_derived":
                new.__dict__[k] = copy.deepcopy(v, memo)
        object.__setattr__(new, "_derived", _Derived())
        return new

    def _replicate_for_data_parallel(self):
        rep = super()._replicate_for_data_parallel()
        object.__setattr__(rep, "_derived", _Derived())
        return rep


def _no_autograd(*ts: Tensor):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts):
        raise NotImplementedError(
            "cswin_unet_b200: this op has no native backward kernel yet; run it under torch.no_grad() "
            "(there is deliberately no autograd fallback through library ops)")


def img2windows(img: Tensor, H_sp: int, W_sp: int) -> Tensor:
    """(B,C,H,W) -> (B*nWin, H_sp*W_sp, C); window id b*nH*nW + ih*nW + iw, token id r*W_sp + c.
    API-compat helper (cswin_unet.py:184-191); the kernels never materialise this layout."""
    B, Cn, H, W = img.shape
    t = img.reshape(B, Cn, H // H_sp, H_sp, W // W_sp, W_sp)
    return t.permute(0, 2, 4, 3, 5, 1).reshape(-1, H_sp * W_sp, Cn)


def windows2img(img_splits_hw: Tensor, H_sp: int, W_sp: int, H: int, W: int) -> Tensor:
    """Inverse of img2windows, returning (B,H,W,C) (cswin_unet.py:194-202)."""
    nwin = (H // H_sp) * (W // W_sp)
    B = img_splits_hw.shape[0] // nwin
    t = img_splits_hw.reshape(B, H // H_sp, W // W_sp, H_sp, W_sp, -1)
    return t.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


# ------------------------------------------------------------------------------------------
# LePEAttention
# ------------------------------------------------------------------------------------------
class LePEAttention(_Native):
    """Cross-shaped-window attention branch with locally-enhanced positional encoding (cswin_unet.py:31-109)."""

    def __init__(self, dim, resolution, idx, split_size, dim_out=None, num_heads=9, attn_drop=0., proj_drop=0.,
                 qk_scale=None):
        super().__init__()
        self.dim = dim
        self.dim_out = dim_out or dim
        self.resolution = resolution
        self.split_size = split_size
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        if idx == -1:
            H_sp, W_sp = resolution, resolution
        elif idx == 0:
            H_sp, W_sp = resolution, split_size
        elif idx == 1:
            W_sp, H_sp = resolution, split_size
        else:
            raise ValueError(f"LePEAttention: idx must be -1, 0 or 1 (got {idx})")   # reference prints and exit(0)s
        self.idx = idx
        self.H_sp = H_sp
        self.W_sp = W_sp
        self.get_v = nn.Conv2d(dim, dim, kernel_size=3, stride=1, padding=1, groups=dim)   # parameter container
        self.attn_drop = nn.Dropout(attn_drop)

    def _check(self, L: int):
        H = W = self.resolution
        assert L == H * W, "flatten img_tokens has wrong size"
        if H % self.H_sp or W % self.W_sp:
            raise RuntimeError(f"LePEAttention: resolution {H} is not divisible by the stripe window "
                               f"{self.H_sp}x{self.W_sp}")
        if self.training and self.attn_drop.p > 0:
            raise NotImplementedError("LePEAttention: attn_drop > 0 is not supported by the fused kernel "
                                      "(the reference never sets it: vision_transformer.py:23-35)")

    def branch_desc(self, q: Tensor, k: Tensor, v: Tensor, out: Tensor, lse: Optional[Tensor] = None) -> dict:
        dt = q.dtype
        return dict(q=q, k=k, v=v, out=out, conv_w=self._w("cw", self.get_v.weight, dt),
                    conv_b=self._w("cb", self.get_v.bias, dt), heads=self.num_heads, H_sp=self.H_sp, W_sp=self.W_sp,
                    lse=lse)

    def forward(self, qkv: Tensor) -> Tensor:
        """qkv: (3, B, L, C) view of any strides (the reference passes a strided slice of the qkv Linear output)."""
        q, k, v = qkv[0], qkv[1], qkv[2]
        B, L, Cn = q.shape
        self._check(L)
        if ag.needs_grad(qkv, self.get_v.weight, self.get_v.bias):
            packed = torch.cat([q, k, v], dim=-1)                       # (B, L, 3 C_b); torch.cat is the tape's job here
            meta = dict(reso=self.resolution, scale=float(self.scale), heads=[self.num_heads], win=[(self.H_sp, self.W_sp)])
            return ag.LepeAttentionFn.apply(packed, self.get_v.weight, self.get_v.bias, None, None, meta)
        out = torch.empty((B, L, Cn), dtype=q.dtype, device=q.device)
        ops.lepe_attention_fwd([self.branch_desc(q, k, v, out)], B, self.resolution, float(self.scale), q.dtype)
        return out


# ------------------------------------------------------------------------------------------
# CSWinBlock
# ------------------------------------------------------------------------------------------
class Mlp(nn.Module):
    """Parameter container with the reference's names (cswin_unet.py:12-28); CSWinBlock runs it fused."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x: Tensor) -> Tensor:
        if ag.needs_grad(x, self.fc1.weight, self.fc2.weight):
            return ag.linear(ag.GeluFn.apply(ag.linear(x, self.fc1.weight, self.fc1.bias)), self.fc2.weight, self.fc2.bias)
        dt = x.dtype
        h = ops.linear(x, self.fc1.weight.detach().to(dt), self.fc1.bias.detach().to(dt), act=1)
        return ops.linear(h, self.fc2.weight.detach().to(dt), self.fc2.bias.detach().to(dt))


class CSWinBlock(_Native):
    """Pre-norm CSWin transformer block (cswin_unet.py:112-181) as 5 fused launches:
    [LN1+qkv] -> [both LePE stripe-attention branches, written straight into the concat layout] ->
    [proj + residual + DropPath] -> [LN2 + fc1 + GELU] -> [fc2 + residual + DropPath]."""

    def __init__(self, dim, reso, num_heads, split_size, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0.,
                 attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, last_stage=False):
        super().__init__()
        if norm_layer is not nn.LayerNorm or act_layer is not nn.GELU:
            raise NotImplementedError("CSWinBlock: the fused kernels implement nn.LayerNorm and nn.GELU only")
        self.dim = dim
        self.num_heads = num_heads
        self.patches_resolution = reso
        self.split_size = split_size
        self.mlp_ratio = mlp_ratio
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.norm1 = norm_layer(dim)
        if self.patches_resolution == split_size:
            last_stage = True
        self.branch_num = 1 if last_stage else 2
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(drop)
        if last_stage:
            self.attns = nn.ModuleList([
                LePEAttention(dim, resolution=self.patches_resolution, idx=-1, split_size=split_size,
                              num_heads=num_heads, dim_out=dim, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
                for _ in range(self.branch_num)])
        else:
            self.attns = nn.ModuleList([
                LePEAttention(dim // 2, resolution=self.patches_resolution, idx=i, split_size=split_size,
                              num_heads=num_heads // 2, dim_out=dim // 2, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)
                for i in range(self.branch_num)])
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, out_features=dim, act_layer=act_layer, drop=drop)
        self.norm2 = norm_layer(dim)
        if drop > 0:
            raise NotImplementedError("CSWinBlock: drop > 0 (proj/MLP dropout) is not fused; the reference uses 0")

    def _sample_scale(self, x: Tensor) -> Optional[Tensor]:
        return self.drop_path.sample_scale(x) if isinstance(self.drop_path, DropPath) else None

    def forward(self, x: Tensor) -> Tensor:
        H = W = self.patches_resolution
        B, L, Cn = x.shape
        assert L == H * W, "flatten img_tokens has wrong size"
        if ag.needs_grad(x, *self.parameters()):
            return self._forward_train(x)
        dt = x.dtype
        w = self._w
        fold = dt == torch.bfloat16 and FOLD_LN
        if fold:
            # LayerNorm folded into the tcgen05 Linear: the GEMM reads the RAW rows; mean / rstd come from the (sum, sum^2)
            # side channel the producing Linear's epilogue wrote (or one row_stats pass for a stage's first block)
            st = getattr(x, "_cswin_stats", None)
            if st is None or st.shape[0] != B * L:
                st = ops.row_stats(x)
            wq, csq, bq = self._folded("qkv", self.qkv, self.norm1)
            if FUSE_QKV_ATTN and x.is_contiguous() and self._qkv_attn_ok():
                for a in self.attns:
                    a._check(L)
                brs = [dict(conv_w=a._w("cw", a.get_v.weight, dt), conv_b=a._w("cb", a.get_v.bias, dt), heads=a.num_heads,
                            H_sp=a.H_sp, W_sp=a.W_sp) for a in self.attns]
                att = ops.qkv_lepe_attention(x, wq, bq, (st, csq, self.norm1.eps), brs, H, float(self.attns[0].scale))
                return self._after_attention(x, att, dt, True)
            qkv = ops.linear(x, wq, None, ln_fold=(st, csq, self.norm1.eps), bias_f32=bq)
        else:
            qkv = ops.linear(x, w("qkv.w", self.qkv.weight, dt),
                             None if self.qkv.bias is None else w("qkv.b", self.qkv.bias, dt),
                             ln=(w("n1.w", self.norm1.weight, dt), w("n1.b", self.norm1.bias, dt), self.norm1.eps))
        att = torch.empty((B, L, Cn), dtype=dt, device=x.device)
        q, k, v = qkv[..., :Cn], qkv[..., Cn:2 * Cn], qkv[..., 2 * Cn:]
        if self.branch_num == 2:
            h = Cn // 2
            descs = []
            for i, a in enumerate(self.attns):
                a._check(L)
                sl = slice(i * h, (i + 1) * h)
                descs.append(a.branch_desc(q[..., sl], k[..., sl], v[..., sl], att[..., sl]))
            scale = float(self.attns[0].scale)
        else:
            self.attns[0]._check(L)
            descs = [self.attns[0].branch_desc(q, k, v, att)]
            scale = float(self.attns[0].scale)
        ops.lepe_attention_fwd(descs, B, H, scale, dt)
        return self._after_attention(x, att, dt, fold)

    def _qkv_attn_ok(self) -> bool:
        ok = self.__dict__.get("_qkv_attn_ok_cache")
        if ok is None:
            ok = ops.qkv_attention_supported(self.dim, self.patches_resolution, [(a.num_heads, a.H_sp, a.W_sp) for a in self.attns])
            self.__dict__["_qkv_attn_ok_cache"] = ok
        return ok

    def _after_attention(self, x: Tensor, att: Tensor, dt, fold: bool) -> Tensor:
        """proj + residual, then the MLP half of the block (cswin_unet.py:177-179)."""
        B, L, Cn = x.shape
        w = self._w
        if fold:
            x1, st1 = ops.linear(att, w("proj.w", self.proj.weight, dt), w("proj.b", self.proj.bias, dt), residual=x,
                                 sample_scale=self._sample_scale(x), rows_per_sample=L, want_stats=True)
            w1, cs1, b1 = self._folded("fc1", self.mlp.fc1, self.norm2)
            if (FUSE_MLP and Cn <= FUSE_MLP_MAX_DIM and B * L <= FUSE_MLP_MAX_ROWS and self._sample_scale(x) is None
                    and ops.mlp_supported(Cn, self.mlp.fc1.out_features)):
                # fc1 + GELU + fc2 + residual in one launch; the hidden activation stays in TMEM / shared memory
                y, st2 = ops.mlp_fused(x1, w1, cs1, b1, w("fc2.w", self.mlp.fc2.weight, dt),
                                       w("fc2.b32", self.mlp.fc2.bias, torch.float32), st1, self.norm2.eps)
                y._cswin_stats = st2
                return y
            hid = ops.linear(x1, w1, None, ln_fold=(st1, cs1, self.norm2.eps), bias_f32=b1, act=1)
            y, st2 = ops.linear(hid, w("fc2.w", self.mlp.fc2.weight, dt), w("fc2.b", self.mlp.fc2.bias, dt), residual=x1,
                                sample_scale=self._sample_scale(x), rows_per_sample=L, want_stats=True)
            y._cswin_stats = st2                        # rides along to the next block of the stage
            return y
        x1 = ops.linear(att, w("proj.w", self.proj.weight, dt), w("proj.b", self.proj.bias, dt), residual=x,
                        sample_scale=self._sample_scale(x), rows_per_sample=L)
        hid = ops.linear(x1, w("fc1.w", self.mlp.fc1.weight, dt), w("fc1.b", self.mlp.fc1.bias, dt),
                         ln=(w("n2.w", self.norm2.weight, dt), w("n2.b", self.norm2.bias, dt), self.norm2.eps), act=1)
        return ops.linear(hid, w("fc2.w", self.mlp.fc2.weight, dt), w("fc2.b", self.mlp.fc2.bias, dt), residual=x1,
                          sample_scale=self._sample_scale(x), rows_per_sample=L)


    def _folded(self, key: str, lin: nn.Linear, norm: nn.LayerNorm):
        """(bf16 W o gamma, fp32 column sums of that bf16 matrix, fp32 b + W beta) for LN -> Linear (see cswin_b200.h)."""
        ps = (lin.weight, norm.weight, norm.bias) + (() if lin.bias is None else (lin.bias,))
        wf = self._w(key + ".fw", ps, torch.bfloat16, lambda W, g, b, *r: W.float() * g.float()[None, :])
        cs = self._w(key + ".fcs", ps, torch.float32, lambda W, g, b, *r: (W.float() * g.float()[None, :]).bfloat16().float().sum(1))
        bf = self._w(key + ".fb", ps, torch.float32,
                     lambda W, g, b, *r: W.float() @ b.float() + (r[0].float() if r else 0.0))
        return wf, cs, bf

    def _forward_train(self, x: Tensor) -> Tensor:
        """Same graph on the autograd tape: every node is a native forward / backward kernel pair (autograd.py)."""
        H = self.patches_resolution
        B, L, Cn = x.shape
        for a in self.attns:
            a._check(L)
        u, xr = ag.LayerNormForkFn.apply(x, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        qkv = ag.linear(u, self.qkv.weight, self.qkv.bias)
        meta = dict(reso=H, scale=float(self.attns[0].scale), heads=[a.num_heads for a in self.attns],
                    win=[(a.H_sp, a.W_sp) for a in self.attns])
        a0 = self.attns[0]
        a1 = self.attns[1] if self.branch_num == 2 else None
        att = ag.LepeAttentionFn.apply(qkv, a0.get_v.weight, a0.get_v.bias, a1.get_v.weight if a1 else None,
                                       a1.get_v.bias if a1 else None, meta)
        x1 = ag.linear(att, self.proj.weight, self.proj.bias, residual=xr, sample_scale=self._sample_scale(x), rps=L)
        u2, x1r = ag.LayerNormForkFn.apply(x1, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        return ag.mlp(u2, self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias, residual=x1r,
                      sample_scale=self._sample_scale(x), rps=L)


# All blocks of a stage as ONE persistent dataflow launch (csrc/stage_tc.cu) instead of 5 launches per block: same tiles, same
# arithmetic as the composed path, ordered by per-row-tile completion counters instead of kernel boundaries.  CSWIN_STAGE_EXEC
# lists the block dims that use it (e.g. "256" or "64,128,256,512"); anything outside the kernel's envelope, training, fp32 and
# DropPath-active calls run the composed path.  MEASURED on B200 (profiles/r02_stage_kernel_*.log): bit-identical to the composed
# path, but SLOWER at batch 24 (stage 3, 9 blocks: 639 us vs 309 us): a block is a chain of 5 dependent ops, and the chain is
# latency-bound either way — per link the dataflow version pays deps-visible -> load 0.6 us -> mainloop 1.4 us (fc2 7 us with the
# 2-stage ring that fits next to a second CTA) -> epilogue + bulk-store completion 3.2-4.4 us -> red.release 0.8-1.4 us -> flag
# poll 1.3-1.9 us = 8-13 us, more than the 5-8 us a PDL-chained launch costs.  Default: off.
STAGE_EXEC_DIMS = tuple(int(t) for t in os.environ.get("CSWIN_STAGE_EXEC", "").split(",") if t.strip() not in ("", "0"))


def _stage_block_desc(blk: "CSWinBlock") -> dict:
    dt = torch.bfloat16
    w = blk._w
    wq, csq, bq = blk._folded("qkv", blk.qkv, blk.norm1)
    w1, cs1, b1 = blk._folded("fc1", blk.mlp.fc1, blk.norm2)
    b32 = lambda t: t.to(dt).float()                          # the composed path adds the bf16-rounded bias in fp32
    return dict(w_qkv=wq, cs_qkv=csq, b_qkv=bq, w_proj=w("proj.w", blk.proj.weight, dt),
                b_proj=w("proj.b32", blk.proj.bias, torch.float32, b32), w_fc1=w1, cs_fc1=cs1, b_fc1=b1,
                w_fc2=w("fc2.w", blk.mlp.fc2.weight, dt), b_fc2=w("fc2.b32r", blk.mlp.fc2.bias, torch.float32, b32),
                lepe_w=[a._w("cw", a.get_v.weight, dt).contiguous() for a in blk.attns],
                lepe_b=[a._w("cb", a.get_v.bias, dt).contiguous() for a in blk.attns],
                eps1=float(blk.norm1.eps), eps2=float(blk.norm2.eps))


def run_stage(blocks, x: Tensor) -> Tensor:
    """`for blk in blocks: x = blk(x)` (cswin_unet.py:462-478, :505-533).  bf16 inference on an eligible stage runs all blocks
    in one persistent launch and OVERWRITES x (always a fresh activation: stem / Merge_Block / skip-Linear output)."""
    blocks = list(blocks)
    b0 = blocks[0] if blocks else None
    ok = (b0 is not None and x.is_cuda and x.dtype == torch.bfloat16 and FOLD_LN and b0.dim in STAGE_EXEC_DIMS
          and not ag.needs_grad(x, *(p for b in blocks for p in b.parameters()))
          and all((not b.training or not isinstance(b.drop_path, DropPath) or b.drop_path.drop_prob == 0.) for b in blocks)
          and all(b.dim == b0.dim and b.patches_resolution == b0.patches_resolution and b.branch_num == b0.branch_num
                  and b.mlp.fc1.out_features == b0.mlp.fc1.out_features and b.qkv.weight.shape[0] == 3 * b0.dim
                  and [(a.num_heads, a.H_sp, a.W_sp) for a in b.attns] == [(a.num_heads, a.H_sp, a.W_sp) for a in b0.attns]
                  and float(b.attns[0].scale) == float(b0.attns[0].scale) for b in blocks))
    plan = None
    if ok:
        B, L, Cn = x.shape
        branches = [(a.num_heads, a.H_sp, a.W_sp) for a in b0.attns]
        key = ("_stage_plan", B)
        plan = b0.__dict__.get(key, False)
        if plan is False:
            plan = ops.stage_plan(B, b0.patches_resolution, Cn, b0.mlp.fc1.out_features, branches) if L == b0.patches_resolution ** 2 else None
            b0.__dict__[key] = plan
    if plan is None:
        for blk in blocks:
            x = blk(x)
        return x
    for a in b0.attns:
        a._check(L)
    st = getattr(x, "_cswin_stats", None)
    if st is None or st.shape[0] != B * L:
        st = ops.row_stats(x)
    if not x.is_contiguous():
        x = x.contiguous()
    return ops.stage_forward(x, st, [_stage_block_desc(b) for b in blocks], b0.patches_resolution, b0.mlp.fc1.out_features,
                             branches, float(b0.attns[0].scale), plan)


# ------------------------------------------------------------------------------------------
# Merge_Block / CARAFE / CARAFE4
# ------------------------------------------------------------------------------------------
def _side(L: int) -> int:
    r = int(round(math.sqrt(L)))
    assert r * r == L, "token count is not a square image"
    return r


class Merge_Block(_Native):
    """Stage down-sampling (cswin_unet.py:205-220): 3x3 stride-2 conv on the token image + LayerNorm, as a
    channel-innermost im2col gather -> Linear (weights re-packed to (2C, ky, kx, C)) -> LayerNorm."""

    def __init__(self, dim, dim_out, norm_layer=nn.LayerNorm):
        super().__init__()
        if norm_layer is not nn.LayerNorm:
            raise NotImplementedError("Merge_Block: only nn.LayerNorm is implemented")
        self.conv = nn.Conv2d(dim, dim_out, 3, 2, 1)
        self.norm = norm_layer(dim_out)

    def forward(self, x: Tensor) -> Tensor:
        B, L, Cn = x.shape
        H = W = _side(L)
        if ag.needs_grad(x, *self.parameters()):
            wk = self.conv.weight.permute(0, 2, 3, 1).reshape(self.conv.weight.shape[0], -1)     # re-pack on the tape
            y = ag.linear(ag.Im2colTokensFn.apply(x, H, W, 3, 3, 2, 1), wk, self.conv.bias)
            y = ag.LayerNormFn.apply(y, self.norm.weight, self.norm.bias, self.norm.eps)
            return y.view(B, ((H + 2 - 3) // 2 + 1) ** 2, -1)
        dt = x.dtype
        wk = self._w("conv.w", self.conv.weight, dt, lambda t: t.permute(0, 2, 3, 1).reshape(t.shape[0], -1))
        y = ops.conv_tokens(x, H, W, wk, self._w("conv.b", self.conv.bias, dt), 3, 3, 2, 1) if IMPLICIT_CONV and Cn % 64 == 0 else None
        if y is None:
            col = ops.im2col_tokens(x, H, W, 3, 3, 2, 1)
            y = ops.linear(col, wk, self._w("conv.b", self.conv.bias, dt))
        y = ops.layernorm_with_row_stats(y, self._w("n.w", self.norm.weight, dt), self._w("n.b", self.norm.bias, dt), self.norm.eps)
        Ho = (H + 2 - 3) // 2 + 1
        v = y.view(B, Ho * Ho, -1)
        if hasattr(y, "_cswin_stats"):
            v._cswin_stats = y._cswin_stats                  # rides along to the first block of the next stage
        return v


class CARAFE(_Native):
    """Content-aware re-assembly up-sampling (cswin_unet.py:222-269).  The 1x1 `out` conv is applied at LOW
    resolution (it commutes with the convex re-assembly; see include/cswin_b200.h) and the softmax / pixel
    shuffles / 3x3 gather run in one kernel."""

    def __init__(self, dim, dim_out, kernel_size=3, up_factor=2):
        super().__init__()
        if kernel_size != 3:
            raise NotImplementedError("CARAFE: only kernel_size=3 is implemented (the reference never uses another)")
        self.kernel_size = kernel_size
        self.up_factor = up_factor
        self.down = nn.Conv2d(dim, dim // 4, 1)
        self.encoder = nn.Conv2d(dim // 4, self.up_factor ** 2 * self.kernel_size ** 2, self.kernel_size, 1,
                                 self.kernel_size // 2)
        self.out = nn.Conv2d(dim, dim_out, 1)

    def _kernel_logits(self, x: Tensor, H: int, W: int, z_weight: Optional[Tensor] = None, z_key: str = "out"):
        """Kernel logits of the re-assembly (down -> 3x3 encoder).  With `z_weight` ((Nz, C): the 1x1 map that is applied at low
        resolution, `out.weight` or the folded head) ONE Linear computes [down | z] from x — both read the same rows — and the
        method returns (logits, z) with z a column view of that Linear's output."""
        dt = x.dtype
        nd0 = self.down.weight.shape[0]
        # fewer than 64 compressed channels (C/4 = 16 / 32 at 56^2 / 28^2): `down` is given zero rows up to 64 outputs and the encoder
        # zero input channels to match, so that the 3x3 encoder runs as an implicit GEMM straight from the token image (K blocks of
        # one tap x 64 channels) — the zero channels add exact zeros; cheaper than writing and re-reading a column matrix
        pad_to = 64 if (IMPLICIT_CONV and dt == torch.bfloat16 and nd0 < 64) else nd0
        F = torch.nn.functional
        if pad_to == nd0:
            wd = self._w("down.w", self.down.weight, dt, lambda t: t.reshape(t.shape[0], -1))
            bd = self._w("down.b", self.down.bias, dt)
            we = self._w("enc.w", self.encoder.weight, dt, lambda t: t.permute(0, 2, 3, 1).reshape(t.shape[0], -1))
        else:
            wd = self._w("down.w.p64", self.down.weight, dt, lambda t: F.pad(t.reshape(t.shape[0], -1), (0, 0, 0, pad_to - nd0)))
            bd = self._w("down.b.p64", self.down.bias, dt, lambda t: F.pad(t, (0, pad_to - nd0)))
            we = self._w("enc.w.p64", self.encoder.weight, dt,
                         lambda t: F.pad(t.permute(0, 2, 3, 1), (0, pad_to - nd0)).reshape(t.shape[0], -1))
        be = self._w("enc.b", self.encoder.bias, dt)
        z = None
        if z_weight is None:
            d = ops.linear(x, wd, bd)                                                   # (B, L, C/4)
        else:
            nd, nz = wd.shape[0], z_weight.shape[0]
            pk = "" if pad_to == nd0 else ".p64"
            wcat = self._w("downz.w." + z_key + pk, (self.down.weight, z_weight), dt,
                           lambda a, b: torch.cat([F.pad(a.reshape(a.shape[0], -1).to(dt), (0, 0, 0, pad_to - nd0)), b.reshape(b.shape[0], -1).to(dt)], 0))
            bcat = self._w("downz.b." + z_key + pk, (self.down.bias, z_weight), dt,
                           lambda a, b: torch.cat([F.pad(a.to(dt), (0, pad_to - nd0)), torch.zeros(b.shape[0], dtype=dt, device=a.device)]))
            y = ops.linear(x, wcat, bcat)                                               # (B, L, C/4 + Nz): [down(x) | z(x)], no bias on z yet
            d, z = y[..., :nd], y[..., nd:]
        enc = ops.conv_tokens(d, H, W, we, be, 3, 3, 1, 1) if IMPLICIT_CONV and d.shape[-1] % 64 == 0 else None
        if enc is None:
            col = ops.im2col_tokens(d, H, W, 3, 3, 1, 1)
            enc = ops.linear(col, we, be)                                               # (B*L, 9 s^2)
        return enc if z_weight is None else (enc, z)

    def _kernel_logits_tape(self, x: Tensor, H: int, W: int) -> Tensor:
        """_kernel_logits on the autograd tape (training)."""
        d = ag.linear(x, self.down.weight.reshape(self.down.weight.shape[0], -1), self.down.bias)
        col = ag.Im2colTokensFn.apply(d, H, W, 3, 3, 1, 1)
        return ag.linear(col, self.encoder.weight.permute(0, 2, 3, 1).reshape(self.encoder.weight.shape[0], -1), self.encoder.bias)

    def forward(self, x: Tensor) -> Tensor:
        B, L, Cn = x.shape
        H = W = _side(L)
        if ag.needs_grad(x, *self.parameters()):
            rs = lambda t: t.reshape(t.shape[0], -1)
            enc = self._kernel_logits_tape(x, H, W)
            z = ag.linear(x, rs(self.out.weight), None)
            return ag.CarafeReassembleFn.apply(enc, z.view(B * L, -1), self.out.bias, B, H, W, self.up_factor)
        dt = x.dtype
        enc, z = self._kernel_logits(x, H, W, self.out.weight)                          # z = out(x) without its bias, at LOW resolution
        return ops.carafe_reassemble(enc, z.reshape(B * L, -1) if z.is_contiguous() else z.flatten(0, 1), self._w("out.b", self.out.bias, dt),
                                     B, H, W, self.up_factor)


class CARAFE4(CARAFE):
    """x4 variant (cswin_unet.py:272-319): identical code in the reference, only `up_factor` differs."""

    def __init__(self, dim, dim_out, kernel_size=3, up_factor=4):
        super().__init__(dim, dim_out, kernel_size, up_factor)
