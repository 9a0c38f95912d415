"""Tensor-level wrappers over the C ABI (include/cswin_b200.h): torch is used for device memory and
streams only.  CUDA tensors only — there is no CPU path; CPU tensors raise.

Every function enqueues on `torch.cuda.current_stream()` and allocates outputs / workspaces with
`torch.empty`, so calls are CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (BF16, F32, LepeBranch, LepeBranchGrad, LinearArgs, MlpArgs, QkvAttnArgs, StageArgs, StageBlock, StagePlan,
                   check, lib)

Tensor = torch.Tensor


def _dtype_code(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"cswin_unet_b200 computes in float32 or bfloat16, got {t.dtype}")


def _need_cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cswin_unet_b200 has no CPU path: tensors must live on a CUDA device "
                               "(the reference's CPU path is only restated in oracle/ for testing)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _rows(t: Tensor) -> Tuple[Tensor, int, int]:
    """View a (..., K) tensor with uniform row pitch as (M, K) rows: returns (t, M, ld). Copies only if it must."""
    if t.stride(-1) != 1:
        t = t.contiguous()
    K = t.shape[-1]
    if t.dim() == 1:
        return t, 1, K
    if t.dim() == 2:
        return t, t.shape[0], t.stride(0)
    lead = t.shape[:-1]
    ld = t.stride(-2)
    # rows are uniformly spaced iff every outer stride equals the product of inner extents times ld
    expect = ld
    ok = True
    for size, stride in zip(reversed(lead), reversed(t.stride()[:-1])):
        if size != 1 and stride != expect:
            ok = False
            break
        expect *= size
    if not ok:
        t = t.contiguous()
        ld = K
    M = 1
    for s in lead:
        M *= s
    return t, M, ld


# ----------------------------------------------------------------------------------------------
# LePE attention
# ----------------------------------------------------------------------------------------------
def _branch(q: Tensor, k: Tensor, v: Tensor, out: Tensor, conv_w: Tensor, conv_b: Tensor, heads: int,
            H_sp: int, W_sp: int, lse: Optional[Tensor]) -> LepeBranch:
    br = LepeBranch()
    br.q, br.k, br.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    br.q_bs, br.q_ts = q.stride(0), q.stride(1)
    br.k_bs, br.k_ts = k.stride(0), k.stride(1)
    br.v_bs, br.v_ts = v.stride(0), v.stride(1)
    br.out, br.o_bs, br.o_ts = out.data_ptr(), out.stride(0), out.stride(1)
    br.conv_w, br.conv_b = conv_w.data_ptr(), conv_b.data_ptr()
    br.lse = _ptr(lse)
    br.C_b, br.heads, br.H_sp, br.W_sp = q.shape[-1], heads, H_sp, W_sp
    return br


def _unit_channel_stride(t: Tensor) -> Tensor:
    return t if t.stride(-1) == 1 else t.contiguous()


def lepe_attention_fwd(branches: Sequence[dict], B: int, reso: int, scale: float, dtype: torch.dtype) -> None:
    """branches: [{q,k,v:(B,L,C_b) views, out:(B,L,C_b) view, conv_w:(C_b,1,3,3), conv_b:(C_b), heads, H_sp, W_sp, lse}]."""
    arr = (LepeBranch * len(branches))()
    keep = []
    for i, b in enumerate(branches):
        q, k, v = (_unit_channel_stride(b[n]) for n in ("q", "k", "v"))
        out = b["out"]
        assert out.stride(-1) == 1
        cw, cb = b["conv_w"].contiguous(), b["conv_b"].contiguous()
        _need_cuda(q, k, v, out, cw, cb)
        keep += [q, k, v, cw, cb]
        arr[i] = _branch(q, k, v, out, cw, cb, b["heads"], b["H_sp"], b["W_sp"], b.get("lse"))
    code = F32 if dtype == torch.float32 else BF16
    check(lib().cswin_lepe_attention_fwd(arr, len(branches), B, reso, C.c_float(scale), code, _stream()),
          "cswin_lepe_attention_fwd")


# ----------------------------------------------------------------------------------------------
# LayerNorm / Linear
# ----------------------------------------------------------------------------------------------
def layernorm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5, out: Optional[Tensor] = None,
              stats: bool = False):
    _need_cuda(x, gamma, beta)
    x2, M, ldx = _rows(x)
    Cn = x.shape[-1]
    y = torch.empty(x.shape, dtype=x.dtype, device=x.device) if out is None else out
    y2, _, ldy = _rows(y)
    assert y2.data_ptr() == y.data_ptr()
    mean = rstd = None
    if stats:
        mean = torch.empty(M, dtype=torch.float32, device=x.device)
        rstd = torch.empty(M, dtype=torch.float32, device=x.device)
    check(lib().cswin_layernorm_fwd(x2.data_ptr(), ldx, gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), ldy, M, Cn,
                                    C.c_float(eps), _ptr(mean), _ptr(rstd), _dtype_code(x), _stream()),
          "cswin_layernorm_fwd")
    return (y, mean, rstd) if stats else y


def layernorm_with_row_stats(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5) -> Tensor:
    """bf16 LayerNorm whose output carries `_cswin_stats` = (M, 1, 2) row (sum, sum^2), for the next block's folded LayerNorm.
    Falls back to the plain kernel (the block then runs one row_stats pass) when the shape is outside the fast path."""
    Cn = x.shape[-1]
    x2, M, ldx = _rows(x)
    if (x.dtype != torch.bfloat16 or Cn not in (64, 128, 256, 512) or ldx % 8 or x2.data_ptr() % 16
            or gamma.data_ptr() % 16 or beta.data_ptr() % 16):
        return layernorm(x, gamma, beta, eps)
    _need_cuda(x, gamma, beta)
    y = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    st = torch.empty((M, 1, 2), dtype=torch.float32, device=x.device)
    check(lib().cswin_layernorm_stats_fwd(x2.data_ptr(), ldx, gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), Cn, M, Cn,
                                          C.c_float(eps), st.data_ptr(), _dtype_code(x), _stream()), "cswin_layernorm_stats_fwd")
    y._cswin_stats = st
    return y


def linear(a: Tensor, w: Tensor, bias: Optional[Tensor] = None, *, a2: Optional[Tensor] = None,
           ln: Optional[Tuple[Tensor, Tensor, float]] = None, act: int = 0, residual: Optional[Tensor] = None,
           sample_scale: Optional[Tensor] = None, rows_per_sample: int = 0, out: Optional[Tensor] = None,
           n_out: Optional[int] = None, w_kn: bool = False, ln_fold=None, bias_f32: Optional[Tensor] = None,
           want_stats: bool = False, aux_out: Optional[Tensor] = None):
    """out = residual + sample_scale[row // rows_per_sample] * act(LN?([a | a2]) @ w[:n_out].T + bias).

    bf16 only: `ln_fold = (stats (M, parts, 2) fp32, colsum (N) fp32, eps)` applies LayerNorm algebraically in the
    epilogue (w must already be W o gamma, bias_f32 = b + W beta — see include/cswin_b200.h); `want_stats=True`
    returns (out, stats) where stats holds the per-row partial (sum, sum^2) of `out` for the next folded Linear."""
    _need_cuda(a, w, bias, a2, residual, sample_scale)
    if ln is not None and a.dtype == torch.bfloat16:
        # bf16 / tcgen05 path: the operand is normalised in fp32 and rounded to bf16 ONCE by the LayerNorm kernel,
        # then streamed by TMA (the fp32 SIMT kernel fuses the normalisation into its operand load instead)
        a = layernorm(a, ln[0], ln[1], ln[2])
        ln = None
    a_, M, lda = _rows(a)
    K1 = a.shape[-1]
    args = LinearArgs()
    args.a, args.lda, args.K1 = a_.data_ptr(), lda, K1
    keep = [a_]
    K2 = 0
    if a2 is not None:
        a2_, M2, lda2 = _rows(a2)
        assert M2 == M
        K2 = a2.shape[-1]
        args.a2, args.lda2, args.K2 = a2_.data_ptr(), lda2, K2
        keep.append(a2_)
    if w_kn:                                              # w is (K, N): out = a @ w  (dgrad reads nn.Linear.weight in place)
        assert w.dim() == 2 and w.stride(1) == 1 and w.shape[0] == K1 + K2, (w.shape, K1, K2)
        N = w.shape[1] if n_out is None else n_out
        args.w_layout = 1
    else:
        assert w.dim() == 2 and w.stride(1) == 1 and w.shape[1] == K1 + K2, (w.shape, K1, K2)
        N = w.shape[0] if n_out is None else n_out
    args.w, args.ldw = w.data_ptr(), w.stride(0)
    args.bias = _ptr(bias)
    if ln is not None:
        args.ln_gamma, args.ln_beta, args.ln_eps = ln[0].data_ptr(), ln[1].data_ptr(), ln[2]
    if out is None:
        out = torch.empty(a.shape[:-1] + (N,), dtype=a.dtype, device=a.device)
    o_, Mo, ldo = _rows(out)
    assert o_.data_ptr() == out.data_ptr() and Mo == M
    if residual is not None:
        r_, Mr, ldr = _rows(residual)
        assert Mr == M and residual.shape[-1] == N
        args.residual, args.ldr = r_.data_ptr(), ldr
        keep.append(r_)
    if sample_scale is not None:
        assert sample_scale.dtype == torch.float32 and rows_per_sample > 0
        args.sample_scale, args.rows_per_sample = sample_scale.data_ptr(), rows_per_sample
    args.out, args.ldo, args.M, args.N, args.act = out.data_ptr(), ldo, M, N, act
    if bias_f32 is not None:
        assert bias_f32.dtype == torch.float32 and bias_f32.numel() >= N
        args.bias_f32 = bias_f32.data_ptr()
    if ln_fold is not None:
        st, cs, eps = ln_fold
        assert st.dtype == torch.float32 and st.is_contiguous() and st.shape[0] == M and st.shape[2] == 2
        assert cs.dtype == torch.float32 and cs.numel() >= N and a2 is None
        args.ln_stats, args.ln_stats_parts, args.ln_C = st.data_ptr(), st.shape[1], K1
        args.ln_colsum, args.ln_eps = cs.data_ptr(), eps
    if aux_out is not None:                               # training: act = 1 and the pre-activation goes here as well
        x_, Mx, ldx = _rows(aux_out)
        assert x_.data_ptr() == aux_out.data_ptr() and Mx == M and aux_out.shape[-1] == N and aux_out.dtype == a.dtype
        args.aux_out, args.ld_aux = aux_out.data_ptr(), ldx
    stats = None
    if want_stats:
        parts = lib().cswin_linear_stats_parts(M, N, K1 + K2, act)
        stats = torch.empty((M, parts, 2), dtype=torch.float32, device=a.device)
        args.stats_out = stats.data_ptr()
    check(lib().cswin_linear_fwd(C.byref(args), _dtype_code(a), _stream()), "cswin_linear_fwd")
    return (out, stats) if want_stats else out


def train_epilogues_supported(a: Tensor, n_out: int) -> bool:
    """The training epilogues of cswin_linear_fwd (aux_out, act 2) need the tcgen05 fast path: bf16, 16-byte aligned rows."""
    return (a.dtype == torch.bfloat16 and a.is_cuda and a.shape[-1] % 8 == 0 and n_out % 8 == 0 and a.stride(-1) == 1
            and a.data_ptr() % 16 == 0)


def mlp_supported(C_: int, hidden: int) -> bool:
    return lib().cswin_mlp_stats_parts(C_, hidden) > 0


def mlp_fused(x: Tensor, w1f: Tensor, cs1: Tensor, b1f: Tensor, w2: Tensor, b2: Tensor, stats: Tensor, eps: float,
              want_stats: bool = True):
    """out = x + GELU(LN(x) W1^T + b1) W2^T + b2 in one tcgen05 launch (LayerNorm folded: w1f = W1 o gamma, cs1 its fp32 row
    sums, b1f = b1 + W1 beta; `stats` = (M, parts, 2) row sums of x).  Returns (out, stats of out | None)."""
    _need_cuda(x, w1f, cs1, b1f, w2, b2, stats)
    assert x.dtype == torch.bfloat16 and w1f.dtype == torch.bfloat16 and w2.dtype == torch.bfloat16
    assert cs1.dtype == b1f.dtype == b2.dtype == stats.dtype == torch.float32 and stats.is_contiguous()
    x_, M, ldx = _rows(x)
    Cn, hid = x.shape[-1], w1f.shape[0]
    assert w1f.shape == (hid, Cn) and w2.shape == (Cn, hid) and w1f.stride(1) == 1 and w2.stride(1) == 1
    assert stats.shape[0] == M and stats.shape[2] == 2 and cs1.numel() == hid and b1f.numel() == hid and b2.numel() == Cn
    out = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    a = MlpArgs()
    a.x, a.ldx, a.w1, a.ldw1 = x_.data_ptr(), ldx, w1f.data_ptr(), w1f.stride(0)
    a.ln_colsum, a.b1, a.w2, a.ldw2, a.b2 = cs1.data_ptr(), b1f.data_ptr(), w2.data_ptr(), w2.stride(0), b2.data_ptr()
    a.ln_stats, a.ln_stats_parts, a.ln_eps = stats.data_ptr(), stats.shape[1], eps
    a.out, a.ldo, a.M, a.C, a.hidden = out.data_ptr(), Cn, M, Cn, hid
    st = None
    if want_stats:
        st = torch.empty((M, lib().cswin_mlp_stats_parts(Cn, hid), 2), dtype=torch.float32, device=x.device)
        a.stats_out = st.data_ptr()
    check(lib().cswin_mlp_fwd(C.byref(a), _dtype_code(x), _stream()), "cswin_mlp_fwd")
    return out, st


def qkv_attention_supported(Cn: int, reso: int, branches: Sequence[Tuple[int, int, int]]) -> bool:
    """branches: [(heads, H_sp, W_sp)] — is the fused [LayerNorm -> qkv -> LePE attention] kernel available for this block?"""
    n = len(branches)
    arr = lambda i: (C.c_int32 * 2)(*([b[i] for b in branches] + [0] * (2 - n)))
    return bool(lib().cswin_qkv_lepe_attention_supported(Cn, reso, n, arr(0), arr(1), arr(2)))


def qkv_lepe_attention(x: Tensor, w: Tensor, bias_f32: Optional[Tensor], ln_fold, branches: Sequence[dict], reso: int,
                       scale: float, out: Optional[Tensor] = None) -> Tensor:
    """att = cat_i LePEAttention_i(qkv(LN(x)) slices) in one tcgen05 launch (cswin_qkv_lepe_attention_fwd).
    x (B, L, C) bf16; w (3C, C) bf16 (gamma-folded when `ln_fold = (stats (M, parts, 2), colsum (3C) fp32, eps)`); bias_f32 (3C) fp32;
    branches: [{conv_w (C_b,1,3,3) bf16, conv_b (C_b) bf16, heads, H_sp, W_sp}]."""
    _need_cuda(x, w, bias_f32)
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.dim() == 3 and x.stride(2) == 1 and w.stride(1) == 1
    B, L, Cn = x.shape
    assert L == reso * reso and w.shape == (3 * Cn, Cn)
    if out is None:
        out = torch.empty((B, L, Cn), dtype=x.dtype, device=x.device)
    a = QkvAttnArgs()
    a.x, a.x_bs, a.x_ts = x.data_ptr(), x.stride(0), x.stride(1)
    a.w, a.ldw = w.data_ptr(), w.stride(0)
    if bias_f32 is not None:
        assert bias_f32.dtype == torch.float32 and bias_f32.numel() == 3 * Cn
        a.bias_f32 = bias_f32.data_ptr()
    if ln_fold is not None:
        st, cs, eps = ln_fold
        assert st.dtype == torch.float32 and st.is_contiguous() and st.shape[0] == B * L and st.shape[2] == 2
        assert cs.dtype == torch.float32 and cs.numel() == 3 * Cn and x.is_contiguous()
        a.ln_stats, a.ln_colsum, a.ln_stats_parts, a.ln_eps = st.data_ptr(), cs.data_ptr(), st.shape[1], eps
    a.out, a.o_bs, a.o_ts = out.data_ptr(), out.stride(0), out.stride(1)
    a.B, a.reso, a.C, a.n_branches, a.scale = B, reso, Cn, len(branches), scale
    keep = []
    for i, b in enumerate(branches):
        cw, cb = b["conv_w"].contiguous(), b["conv_b"].contiguous()
        _need_cuda(cw, cb)
        assert cw.dtype == torch.bfloat16 and cb.dtype == torch.bfloat16
        keep += [cw, cb]
        a.br[i].conv_w, a.br[i].conv_b = cw.data_ptr(), cb.data_ptr()
        a.br[i].heads, a.br[i].H_sp, a.br[i].W_sp = b["heads"], b["H_sp"], b["W_sp"]
    check(lib().cswin_qkv_lepe_attention_fwd(C.byref(a), BF16, _stream()), "cswin_qkv_lepe_attention_fwd")
    return out


def stage_plan(B: int, reso: int, Cn: int, hidden: int, branches: Sequence[Tuple[int, int, int]]):
    """branches: [(heads, H_sp, W_sp)].  Returns the StagePlan of the persistent stage kernel, or None outside its envelope."""
    n = len(branches)
    arr = lambda i: (C.c_int32 * 2)(*([b[i] for b in branches] + [0] * (2 - n)))
    plan = StagePlan()
    rc = lib().cswin_stage_plan(B, reso, Cn, hidden, n, arr(0), arr(1), arr(2), C.byref(plan))
    return plan if rc == 0 else None


_STAGE_WS = {}


def _stage_workspace(x: Tensor, hidden: int, plan) -> dict:
    """qkv / att / x1 / hid / statistics / ctrl buffers of the stage kernel, cached per (device, stream, shape): the launches of
    one stream are ordered, so consecutive stages of the same shape share them; ctrl is zeroed once here and left zeroed by the
    kernel."""
    B, L, Cn = x.shape
    key = (x.device, torch.cuda.current_stream().cuda_stream, B, L, Cn, hidden)
    ws = _STAGE_WS.get(key)
    if ws is None:
        M = B * L
        e = lambda *shape, dt=torch.bfloat16: torch.empty(shape, dtype=dt, device=x.device)
        ws = dict(qkv=e(M, 3 * Cn), att=e(M, Cn), x1=e(M, Cn), hid=e(M, hidden),
                  stats_x=e(M, plan.parts_x, 2, dt=torch.float32), stats_x1=e(M, plan.parts_x1, 2, dt=torch.float32),
                  ctrl=torch.zeros(plan.ctrl_ints, dtype=torch.int32, device=x.device))
        _STAGE_WS[key] = ws
    return ws


def stage_forward(x: Tensor, stats_in: Tensor, blocks: Sequence[dict], reso: int, hidden: int,
                  branches: Sequence[Tuple[int, int, int]], scale: float, plan=None) -> Tensor:
    """All CSWinBlocks of one stage in one persistent launch per <= plan.max_blocks blocks (cswin_stage_fwd).  x (B, L, C) bf16
    contiguous is OVERWRITTEN with the stage's output and returned.  blocks: [{w_qkv, cs_qkv, b_qkv, w_proj, b_proj, w_fc1, cs_fc1,
    b_fc1, w_fc2, b_fc2, lepe_w: [..], lepe_b: [..], eps1, eps2}] (see include/cswin_b200.h); stats_in (M, parts, 2) fp32."""
    _need_cuda(x, stats_in)
    assert x.dtype == torch.bfloat16 and x.dim() == 3 and x.is_contiguous()
    B, L, Cn = x.shape
    assert L == reso * reso and stats_in.dtype == torch.float32 and stats_in.is_contiguous() and stats_in.shape[0] == B * L
    plan = plan or stage_plan(B, reso, Cn, hidden, branches)
    if plan is None:
        raise _lib.CswinError("stage_forward: shape outside the persistent stage kernel's envelope")
    ws = _stage_workspace(x, hidden, plan)
    n = len(branches)
    st, parts = stats_in, stats_in.shape[1]
    for lo in range(0, len(blocks), plan.max_blocks):
        chunk = blocks[lo:lo + plan.max_blocks]
        arr = (StageBlock * len(chunk))()
        for i, b in enumerate(chunk):
            for k in ("w_qkv", "w_proj", "w_fc1", "w_fc2"):
                assert b[k].dtype == torch.bfloat16 and b[k].is_contiguous() and b[k].is_cuda
            for k in ("cs_qkv", "b_qkv", "b_proj", "cs_fc1", "b_fc1", "b_fc2"):
                assert b[k].dtype == torch.float32 and b[k].is_contiguous() and b[k].is_cuda
            assert b["w_qkv"].shape == (3 * Cn, Cn) and b["w_proj"].shape == (Cn, Cn) and b["w_fc1"].shape == (hidden, Cn)
            assert b["w_fc2"].shape == (Cn, hidden)
            d = arr[i]
            for k in ("w_qkv", "cs_qkv", "b_qkv", "w_proj", "b_proj", "w_fc1", "cs_fc1", "b_fc1", "w_fc2", "b_fc2"):
                setattr(d, k, b[k].data_ptr())
            for j in range(n):
                cw, cb = b["lepe_w"][j], b["lepe_b"][j]
                assert cw.dtype == torch.bfloat16 and cb.dtype == torch.bfloat16 and cw.is_contiguous() and cb.is_contiguous()
                d.lepe_w[j], d.lepe_b[j] = cw.data_ptr(), cb.data_ptr()
            d.eps1, d.eps2 = b["eps1"], b["eps2"]
        a = StageArgs()
        a.x, a.stats_in, a.stats_in_parts, a.n_blocks = x.data_ptr(), st.data_ptr(), parts, len(chunk)
        a.qkv, a.att, a.x1, a.hid = (ws[k].data_ptr() for k in ("qkv", "att", "x1", "hid"))
        a.stats_x, a.stats_x1, a.ctrl, a.ctrl_ints = ws["stats_x"].data_ptr(), ws["stats_x1"].data_ptr(), ws["ctrl"].data_ptr(), ws["ctrl"].numel()
        a.blocks = arr
        a.B, a.reso, a.C, a.hidden, a.n_branches, a.scale = B, reso, Cn, hidden, n, scale
        for j, (h, hs, wsp) in enumerate(branches):
            a.heads[j], a.H_sp[j], a.W_sp[j] = h, hs, wsp
        check(lib().cswin_stage_fwd(C.byref(a), BF16, _stream()), "cswin_stage_fwd")
        st, parts = ws["stats_x"], plan.parts_x                 # the next chunk continues from this one's row statistics
    return x


def conv_tokens(x: Tensor, H: int, W: int, w: Tensor, bias: Optional[Tensor], KH: int, KW: int, stride: int, pad: int) -> Optional[Tensor]:
    """Convolution over the token image x (B, H*W, C) as an implicit GEMM on the tcgen05 Linear (cswin_conv_tokens_fwd): w (N, KH*KW*C)
    in (ky, kx, c) order.  Returns (B*OH*OW, N) bf16, or None outside the kernel's envelope (the caller composes im2col + linear)."""
    _need_cuda(x, w, bias)
    if x.dtype != torch.bfloat16 or x.dim() != 3 or x.stride(2) != 1:
        return None
    B, L, Cn = x.shape
    assert L == H * W and w.dim() == 2 and w.stride(1) == 1 and w.shape[1] == KH * KW * Cn and w.dtype == x.dtype
    OH, OW = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    N = w.shape[0]
    out = torch.empty((B * OH * OW, N), dtype=x.dtype, device=x.device)
    handled = C.c_int32(0)
    check(lib().cswin_conv_tokens_fwd(x.data_ptr(), x.stride(0), x.stride(1), w.data_ptr(), w.stride(0), _ptr(bias), out.data_ptr(),
                                      out.stride(0), B, H, W, Cn, N, KH, KW, stride, pad, BF16, _stream(), C.byref(handled)),
          "cswin_conv_tokens_fwd")
    return out if handled.value else None


def stem_fused(x: Tensor, w_packed: Tensor, bias: Tensor, gamma: Tensor, beta: Tensor, eps: float) -> Optional[Tensor]:
    """Conv2d(3, 64, 7, 4, 2) + token layout + LayerNorm(64) in one tcgen05 launch (cswin_stem_fwd).  x (B, 3, H, W) fp32 / bf16,
    w_packed (64, 192) bf16, bias / gamma / beta (64) fp32.  Returns (B, Ho Wo, 64) bf16 carrying `_cswin_stats`, or None when the
    shape is outside the kernel's envelope (the caller composes im2col + Linear + LayerNorm)."""
    _need_cuda(x, w_packed, bias, gamma, beta)
    assert x.dim() == 4 and x.shape[1] == 3 and x.dtype in (torch.float32, torch.bfloat16)
    assert w_packed.dtype == torch.bfloat16 and tuple(w_packed.shape) == (64, 192) and w_packed.is_contiguous()
    assert bias.dtype == gamma.dtype == beta.dtype == torch.float32 and bias.numel() == gamma.numel() == beta.numel() == 64
    x = x.contiguous()
    B, _, H, W = x.shape
    Ho, Wo = (H + 4 - 7) // 4 + 1, (W + 4 - 7) // 4 + 1
    out = torch.empty((B, Ho * Wo, 64), dtype=torch.bfloat16, device=x.device)
    st = torch.empty((B * Ho * Wo, 1, 2), dtype=torch.float32, device=x.device)
    handled = C.c_int32(0)
    check(lib().cswin_stem_fwd(x.data_ptr(), int(x.dtype == torch.float32), w_packed.data_ptr(), bias.data_ptr(), gamma.data_ptr(),
                               beta.data_ptr(), C.c_float(eps), out.data_ptr(), st.data_ptr(), B, H, W, BF16, _stream(),
                               C.byref(handled)), "cswin_stem_fwd")
    if not handled.value:
        return None
    out._cswin_stats = st
    return out


SGD_CHUNK = 65536


def sgd_chunk_table(params, grads, momenta, shadows) -> Tensor:
    """Host-side int64 (n_chunks, 5) table == cswin_sgd_chunk_t[]: one row per <= SGD_CHUNK-element piece of a parameter."""
    rows = []
    for p, g, m, sh in zip(params, grads, momenta, shadows):
        assert p.dtype == g.dtype == m.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous() and m.is_contiguous()
        assert g.numel() == p.numel() == m.numel() and (sh is None or (sh.dtype == torch.bfloat16 and sh.is_contiguous()))
        n = p.numel()
        for off in range(0, n, SGD_CHUNK):
            rows.append((p.data_ptr() + 4 * off, g.data_ptr() + 4 * off, m.data_ptr() + 4 * off,
                         0 if sh is None else sh.data_ptr() + 2 * off, min(SGD_CHUNK, n - off)))
    return torch.tensor(rows, dtype=torch.int64).view(-1, 5)


def sgd_momentum_step(table_dev: Tensor, lr_dev: Tensor, momentum: float, weight_decay: float) -> None:
    """One launch: m = momentum m + (g + wd p); p -= lr m; shadow = bf16(p) for every chunk of table_dev (device int64 (n,5))."""
    _need_cuda(table_dev, lr_dev)
    assert table_dev.dtype == torch.int64 and table_dev.is_contiguous() and lr_dev.dtype == torch.float32
    check(lib().cswin_sgd_momentum_step(table_dev.data_ptr(), table_dev.shape[0], lr_dev.data_ptr(), momentum, weight_decay,
                                        _stream()), "cswin_sgd_momentum_step")


def row_stats(x: Tensor) -> Tensor:
    """(M, 1, 2) fp32 per-row (sum, sum^2) of a (..., C) activation — seeds the folded-LayerNorm chain."""
    _need_cuda(x)
    x_, M, ldx = _rows(x)
    st = torch.empty((M, 1, 2), dtype=torch.float32, device=x.device)
    check(lib().cswin_row_stats(x_.data_ptr(), ldx, M, x.shape[-1], st.data_ptr(), _dtype_code(x), _stream()), "cswin_row_stats")
    return st


# ----------------------------------------------------------------------------------------------
# conv gathers / CARAFE reassembly
# ----------------------------------------------------------------------------------------------
def im2col_tokens(x: Tensor, H: int, W: int, KH: int, KW: int, stride: int, pad: int) -> Tensor:
    """x: (B, H*W, C) token-major -> (B*Ho*Wo, KH*KW*C), column index (ky*KW+kx)*C + c."""
    _need_cuda(x)
    x = _unit_channel_stride(x)
    B, L, Cn = x.shape
    assert L == H * W
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    col = torch.empty((B * Ho * Wo, KH * KW * Cn), dtype=x.dtype, device=x.device)
    check(lib().cswin_im2col_tokens(x.data_ptr(), x.stride(0), x.stride(1), col.data_ptr(), col.stride(0), B, H, W, Cn,
                                    KH, KW, stride, pad, _dtype_code(x), _stream()), "cswin_im2col_tokens")
    return col


def im2col_nchw(x: Tensor, KH: int, KW: int, stride: int, pad: int, ldcol: int, dtype: torch.dtype) -> Tensor:
    """x: (B, C, H, W) fp32 or bf16 -> (B*Ho*Wo, ldcol) of `dtype`, column (c*KH+ky)*KW+kx, zero padded to ldcol."""
    _need_cuda(x)
    x = x.contiguous()
    B, Cn, H, W = x.shape
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    col = torch.empty((B * Ho * Wo, ldcol), dtype=dtype, device=x.device)
    code = F32 if dtype == torch.float32 else BF16
    check(lib().cswin_im2col_nchw(x.data_ptr(), int(x.dtype == torch.float32), col.data_ptr(), ldcol, B, Cn, H, W, KH,
                                  KW, stride, pad, code, _stream()), "cswin_im2col_nchw")
    return col


def carafe_reassemble(enc: Tensor, z: Tensor, bias: Tensor, B: int, H: int, W: int, up: int, *, nchw_out: bool = False,
                      out_dtype: Optional[torch.dtype] = None) -> Tensor:
    """enc: (B*H*W, 9 up^2) logits, z: (B*H*W, C) -> (B, up^2 H W, C) token-major, or (B, C, up H, up W) if nchw_out."""
    _need_cuda(enc, z, bias)
    Cn = z.shape[-1]
    od = out_dtype or z.dtype
    if nchw_out:
        y = torch.empty((B, Cn, H * up, W * up), dtype=od, device=z.device)
        ldy = 0
    else:
        y = torch.empty((B, H * up * W * up, Cn), dtype=od, device=z.device)
        ldy = Cn
    check(lib().cswin_carafe_reassemble_fwd(enc.data_ptr(), enc.stride(0), z.data_ptr(), z.stride(0), bias.data_ptr(),
                                            y.data_ptr(), ldy, int(nchw_out), int(od == torch.float32), B, H, W, Cn, up,
                                            _dtype_code(z), _stream()), "cswin_carafe_reassemble_fwd")
    return y


def carafe_head(enc: Tensor, z: Tensor, bias: Tensor, B: int, H: int, W: int, up: int, *, want_logits: bool = True,
                want_labels: bool = False, logits_dtype: Optional[torch.dtype] = None, n_classes: Optional[int] = None,
                out_logits: Optional[Tensor] = None, out_labels: Optional[Tensor] = None):
    """Folded segmentation head: returns (logits (B,C,up H,up W) or None, labels uint8 (B,up H,up W) or None).
    z may be padded beyond n_classes columns (rows of 16 bf16 let the kernel use 16-byte loads).  `out_logits` / `out_labels`:
    contiguous destination tensors of those shapes (e.g. a batch slice of a larger buffer) instead of fresh allocations."""
    _need_cuda(enc, z, bias, out_logits, out_labels)
    Cn = n_classes or z.shape[-1]
    ld = logits_dtype or z.dtype
    logits = labels = None
    if want_logits:
        logits = out_logits if out_logits is not None else torch.empty((B, Cn, H * up, W * up), dtype=ld, device=z.device)
        assert logits.shape == (B, Cn, H * up, W * up) and logits.dtype == ld and logits.is_contiguous()
    if want_labels:
        labels = out_labels if out_labels is not None else torch.empty((B, H * up, W * up), dtype=torch.uint8, device=z.device)
        assert labels.shape == (B, H * up, W * up) and labels.dtype == torch.uint8 and labels.is_contiguous()
    check(lib().cswin_carafe_head_fwd(enc.data_ptr(), enc.stride(0), z.data_ptr(), z.stride(0), bias.data_ptr(),
                                      _ptr(logits), int(ld == torch.float32), _ptr(labels), B, H, W, Cn, up,
                                      _dtype_code(z), _stream()), "cswin_carafe_head_fwd")
    return logits, labels


# ----------------------------------------------------------------------------------------------
# backward ops (training step)
# ----------------------------------------------------------------------------------------------
def seg_loss_supported(n_classes: int) -> bool:
    return n_classes in (2, 3, 4, 9)


def _label_bytes(labels: Tensor) -> int:
    return {torch.uint8: 1, torch.int32: 4, torch.int64: 8}[labels.dtype]


def seg_loss_fwd(logits: Tensor, labels: Tensor) -> Tensor:
    """logits fp32 NCHW contiguous, labels (B, H, W) uint8/int32/int64 -> sums (1 + 3C) fp32 (see cswin_seg_loss_fwd)."""
    _need_cuda(logits, labels)
    assert logits.dtype == torch.float32 and logits.is_contiguous() and labels.is_contiguous()
    B, Cn = logits.shape[:2]
    HW = logits[0, 0].numel()
    assert labels.numel() == B * HW
    sums = torch.zeros(1 + 3 * Cn, dtype=torch.float32, device=logits.device)
    check(lib().cswin_seg_loss_fwd(logits.data_ptr(), labels.data_ptr(), _label_bytes(labels), sums.data_ptr(), B, Cn, HW, _stream()),
          "cswin_seg_loss_fwd")
    return sums


def seg_loss_bwd(logits: Tensor, labels: Tensor, sums: Tensor, grad_out: Tensor, w_ce: float, w_dice: float) -> Tensor:
    _need_cuda(logits, labels, sums, grad_out)
    B, Cn = logits.shape[:2]
    HW = logits[0, 0].numel()
    go = grad_out.reshape(1).float().contiguous()
    dl = torch.empty_like(logits)
    check(lib().cswin_seg_loss_bwd(logits.data_ptr(), labels.data_ptr(), _label_bytes(labels), sums.data_ptr(), go.data_ptr(),
                                   dl.data_ptr(), w_ce, w_dice, B, Cn, HW, _stream()), "cswin_seg_loss_bwd")
    return dl


def carafe_head_bwd_supported(n_classes: int, up: int) -> bool:
    return up == 4 and n_classes in (2, 3, 4, 9)


def carafe_head_bwd(enc: Tensor, z: Tensor, dlogits: Tensor, B: int, H: int, W: int, up: int, n_classes: int):
    """Backward of carafe_head: dlogits fp32 NCHW -> (d enc (M, 144) [row pitch padded to 16 B], d z like z, d bias fp32 (C))."""
    _need_cuda(enc, z, dlogits)
    dlogits = dlogits.float().contiguous()
    M = B * H * W
    assert dlogits.shape == (B, n_classes, H * up, W * up) and enc.shape[0] == M and z.shape[0] == M and enc.stride(1) == 1
    ne = enc.shape[-1]
    denc = torch.empty((M, (ne + 7) // 8 * 8), dtype=enc.dtype, device=enc.device)[:, :ne]
    dz = torch.empty((M, z.shape[-1]), dtype=z.dtype, device=z.device)
    dbias = torch.zeros(n_classes, dtype=torch.float32, device=z.device)
    kws = torch.empty(M * ne, dtype=torch.float32, device=z.device)
    check(lib().cswin_carafe_head_bwd(enc.data_ptr(), enc.stride(0), z.data_ptr(), z.stride(0), dlogits.data_ptr(),
                                      denc.data_ptr(), denc.stride(0), dz.data_ptr(), dz.stride(0), z.shape[-1],
                                      dbias.data_ptr(), kws.data_ptr(), B, H, W, n_classes, up, _dtype_code(z), _stream()),
          "cswin_carafe_head_bwd")
    return denc, dz, dbias
def _lepe_grad_array(branches: Sequence[dict], param_grads: bool):
    arr = (LepeBranchGrad * len(branches))()
    keep = []
    for i, b in enumerate(branches):
        q, k, v = (_unit_channel_stride(b[n]) for n in ("q", "k", "v"))
        cw, cb = b["conv_w"].contiguous(), b["conv_b"].contiguous()
        dout = _unit_channel_stride(b["dout"])
        _need_cuda(q, k, v, cw, cb, dout, b["dq"], b["dk"], b["dv"], b["dconv_w"], b["dconv_b"])
        keep += [q, k, v, cw, cb, dout]
        g = LepeBranchGrad()
        g.fwd = _branch(q, k, v, dout, cw, cb, b["heads"], b["H_sp"], b["W_sp"], b.get("lse"))
        g.dout, g.do_bs, g.do_ts = dout.data_ptr(), dout.stride(0), dout.stride(1)
        for n in ("dq", "dk", "dv"):
            t = b[n]
            assert t.stride(-1) == 1
            setattr(g, n, t.data_ptr()); setattr(g, n + "_bs", t.stride(0)); setattr(g, n + "_ts", t.stride(1))
        assert b["dconv_w"].dtype == torch.float32 and b["dconv_b"].dtype == torch.float32
        if param_grads:
            g.dconv_w, g.dconv_b = b["dconv_w"].data_ptr(), b["dconv_b"].data_ptr()
        else:
            g.dconv_w, g.dconv_b = None, None
        arr[i] = g
    return arr, keep


def lepe_attention_bwd(branches: Sequence[dict], B: int, reso: int, scale: float, dtype: torch.dtype,
                       param_grads: bool = True) -> None:
    """branches: forward description + dout, dq, dk, dv (views with unit channel stride), dconv_w (C_b,9) / dconv_b (C_b) fp32.
    param_grads=False: dconv_w / dconv_b are left alone (the caller ran `lepe_param_grad`)."""
    arr, keep = _lepe_grad_array(branches, param_grads)
    code = F32 if dtype == torch.float32 else BF16
    check(lib().cswin_lepe_attention_bwd(arr, len(branches), B, reso, C.c_float(scale), code, _stream()),
          "cswin_lepe_attention_bwd")


def lepe_param_grad(branches: Sequence[dict], B: int, reso: int, dtype: torch.dtype) -> bool:
    """dconv_w / dconv_b += gradients of get_v from dout and v alone, on the current stream.  False = configuration outside the
    kernel's envelope, nothing launched (then call lepe_attention_bwd with param_grads=True)."""
    if dtype != torch.bfloat16:
        return False
    arr, keep = _lepe_grad_array(branches, True)
    handled = C.c_int32(0)
    check(lib().cswin_lepe_param_grad(arr, len(branches), B, reso, BF16, _stream(), C.byref(handled)), "cswin_lepe_param_grad")
    return bool(handled.value)


def act_fwd(z: Tensor, act: int = 1) -> Tensor:
    _need_cuda(z)
    z2, M, ldz = _rows(z)
    out = torch.empty(z.shape, dtype=z.dtype, device=z.device)
    check(lib().cswin_act_fwd(z2.data_ptr(), ldz, out.data_ptr(), z.shape[-1], M, z.shape[-1], act, _dtype_code(z), _stream()),
          "cswin_act_fwd")
    return out


def act_bwd(dout: Tensor, z: Optional[Tensor], sample_scale: Optional[Tensor], rows_per_sample: int, act: int) -> Tensor:
    _need_cuda(dout, z, sample_scale)
    d2, M, ldd = _rows(dout)
    N = dout.shape[-1]
    z2, ldz = (None, 0)
    if z is not None:
        z2, _, ldz = _rows(z)
    dz = torch.empty(dout.shape, dtype=dout.dtype, device=dout.device)
    check(lib().cswin_act_bwd(d2.data_ptr(), ldd, _ptr(z2), ldz, _ptr(sample_scale), rows_per_sample, dz.data_ptr(), N, M, N,
                              act, _dtype_code(dout), _stream()), "cswin_act_bwd")
    return dz


def linear_wgrad(dz: Tensor, a: Tensor, dw: Tensor, db: Optional[Tensor]) -> None:
    """dw (N, K) view of an fp32 buffer (unit column stride) += dz^T a ; db (N) fp32 += column sums of dz."""
    _need_cuda(dz, a, dw, db)
    d2, M, ldz = _rows(dz)
    a2, Ma, lda = _rows(a)
    assert Ma == M and dw.dtype == torch.float32 and dw.stride(1) == 1 and dz.dtype == a.dtype
    check(lib().cswin_linear_wgrad(d2.data_ptr(), ldz, a2.data_ptr(), lda, dw.data_ptr(), dw.stride(0), _ptr(db), M,
                                   dz.shape[-1], a.shape[-1], _dtype_code(dz), _stream()), "cswin_linear_wgrad")


def layernorm_bwd(x: Tensor, dy: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor, dg: Optional[Tensor] = None,
                  db: Optional[Tensor] = None, dx_add: Optional[Tensor] = None):
    """dx = LayerNorm'(dy) (+ dx_add: the gradient that reached x through the residual connection around the LayerNorm)."""
    _need_cuda(x, dy, gamma, mean, rstd, dx_add)
    x2, M, ldx = _rows(x)
    d2, _, ldy = _rows(dy)
    add2, lda = None, 0
    if dx_add is not None:
        assert dx_add.shape == x.shape and dx_add.dtype == x.dtype
        add2, _, lda = _rows(dx_add)
    Cn = x.shape[-1]
    dx = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    if dg is None:                                        # accumulated into: the caller may pass pre-zeroed slices
        dg = torch.zeros(Cn, dtype=torch.float32, device=x.device)
    if db is None:
        db = torch.zeros(Cn, dtype=torch.float32, device=x.device)
    check(lib().cswin_layernorm_bwd(x2.data_ptr(), ldx, d2.data_ptr(), ldy, gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                    dx.data_ptr(), Cn, _ptr(add2), lda, dg.data_ptr(), db.data_ptr(), M, Cn, _dtype_code(x),
                                    _stream()),
          "cswin_layernorm_bwd")
    return dx, dg, db


def col2im_tokens(dcol: Tensor, B: int, H: int, W: int, Cn: int, KH: int, KW: int, stride: int, pad: int) -> Tensor:
    _need_cuda(dcol)
    dcol = dcol.contiguous()
    dx = torch.empty((B, H * W, Cn), dtype=dcol.dtype, device=dcol.device)
    check(lib().cswin_col2im_tokens(dcol.data_ptr(), dcol.stride(0), dx.data_ptr(), dx.stride(0), dx.stride(1), B, H, W, Cn,
                                    KH, KW, stride, pad, _dtype_code(dcol), _stream()), "cswin_col2im_tokens")
    return dx


def carafe_reassemble_bwd(enc: Tensor, z: Tensor, dy: Tensor, B: int, H: int, W: int, up: int, nchw: bool = False):
    """dy: (B, up^2 H W, C) token-major, or (B, C, up H, up W) if nchw.  Returns (d enc, d z, d bias fp32)."""
    _need_cuda(enc, z, dy)
    Cn = z.shape[-1]
    dy = dy.contiguous()
    Ho, Wo = H * up, W * up
    if nchw:
        sb, sy, sx, sc = Cn * Ho * Wo, Wo, 1, Ho * Wo
    else:
        sb, sy, sx, sc = Ho * Wo * Cn, Wo * Cn, Cn, 1
    ne = enc.shape[-1]                                    # row pitch padded to 16 bytes: the encoder Linear's dgrad / wgrad stay on tcgen05
    denc = torch.empty(enc.shape[:-1] + ((ne + 7) // 8 * 8,), dtype=enc.dtype, device=enc.device)[..., :ne]
    dz = torch.empty_like(z)
    dbias = torch.zeros(Cn, dtype=torch.float32, device=z.device)
    kws = torch.empty(B * H * W * up * up * 9, dtype=torch.float32, device=z.device)
    check(lib().cswin_carafe_reassemble_bwd(enc.data_ptr(), enc.stride(0), z.data_ptr(), z.stride(0), dy.data_ptr(),
                                            int(dy.dtype == torch.float32), sb, sy, sx, sc, denc.data_ptr(), denc.stride(0),
                                            dz.data_ptr(), dz.stride(0), dbias.data_ptr(), kws.data_ptr(), B, H, W, Cn, up,
                                            _dtype_code(z), _stream()), "cswin_carafe_reassemble_bwd")
    return denc, dz, dbias


def zoom_cubic(x: Tensor, out_hw: Tuple[int, int], out: Optional[Tensor] = None, work: Optional[Tensor] = None) -> Tensor:
    """scipy.ndimage.zoom(slice, (oh / H, ow / W), order=3) for every slice of x (n, H, W) float32 (utils.py:69).
    `out`: (n, C, oh, ow) float32 contiguous per channel plane — every channel receives the same plane (the model's 1 -> 3
    channel repeat) — default (n, 1, oh, ow).  `work`: n*H*W float64 scratch."""
    _need_cuda(x, out, work)
    assert x.dtype == torch.float32 and x.dim() == 3 and x.is_contiguous()
    n, H, W = x.shape
    oh, ow = out_hw
    if out is None:
        out = torch.empty((n, 1, oh, ow), dtype=torch.float32, device=x.device)
    assert out.dtype == torch.float32 and out.dim() == 4 and out.shape[0] >= n and tuple(out.shape[2:]) == (oh, ow)
    assert out.stride(3) == 1 and out.stride(2) == ow
    if work is None:
        work = torch.empty(n * H * W, dtype=torch.float64, device=x.device)
    assert work.dtype == torch.float64 and work.numel() >= n * H * W and work.is_contiguous()
    check(lib().cswin_zoom_cubic_fwd(x.data_ptr(), n, H, W, work.data_ptr(), out.data_ptr(), out.stride(0), out.stride(1),
                                     out.shape[1], oh, ow, _stream()), "cswin_zoom_cubic_fwd")
    return out


def zoom_nearest_u8(lab: Tensor, out_hw: Tuple[int, int], out: Optional[Tensor] = None) -> Tensor:
    """scipy.ndimage.zoom(label_map, (oh / H, ow / W), order=0) for every (H, W) uint8 map of lab (n, H, W) (utils.py:77)."""
    _need_cuda(lab, out)
    assert lab.dtype == torch.uint8 and lab.dim() == 3 and lab.is_contiguous()
    n, H, W = lab.shape
    oh, ow = out_hw
    if out is None:
        out = torch.empty((n, oh, ow), dtype=torch.uint8, device=lab.device)
    assert out.dtype == torch.uint8 and out.is_contiguous() and out.shape[0] >= n and tuple(out.shape[1:]) == (oh, ow)
    check(lib().cswin_zoom_nearest_u8(lab.data_ptr(), n, H, W, out.data_ptr(), oh, ow, _stream()), "cswin_zoom_nearest_u8")
    return out
