"""ORACLE — test infrastructure, not product code.

A CPU restatement (numpy loops for small cases, vectorised torch fp32/fp64 for full sizes) of
the CSWin-UNet hot path of BoloniniD/CSWin-UNet.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import this package; the
product (`cswin_unet_b200`) never does and has no CPU path of its own.

Parity pin: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against outputs of the *unmodified* reference module run in the build
container (`tests/golden/make_golden.py` imports /root/reference/networks/cswin_unet.py through a
3-symbol timm shim and commits the vectors under tests/golden/); `tests/test_oracle_golden.py`
checks every function below against them.

Everything is written functionally over a flat `state_dict`-style mapping (same keys as the
reference's `CSWinTransformer.state_dict()`), token-major (B, L, C) throughout, with explicit
index math instead of the reference's view/permute/contiguous chains:

  lepe_attention_loops   /root/reference/networks/cswin_unet.py:82-109 (+59-80, 184-202)   numpy loops
  lepe_attention         same, vectorised (window gather by index table, LePE as masked image shifts)
  cswin_block            :160-181 (+ Mlp :12-28)
  merge_block            :211-220
  carafe                 :232-269 and :282-319 (up_factor 2 / 4)
  cswin_unet_forward     :462-554 (stem :338-342, skips :509-527, head :536-544)
  dice_loss / seg_loss   /root/reference/utils.py:9-45, /root/reference/trainer.py:55-57
  dice_hd95_percase      /root/reference/utils.py:48-58 (medpy binary.dc / binary.hd95 restated)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration (the 6 numbers of configs/cswin_tiny_224_lite.yaml + config.py:57-66)
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class OracleConfig:
    img_size: int = 224
    in_chans: int = 3
    num_classes: int = 9
    embed_dim: int = 64
    depth: Tuple[int, ...] = (1, 2, 9, 1)
    split_size: Tuple[int, ...] = (1, 2, 7, 7)
    num_heads: Tuple[int, ...] = (2, 4, 8, 16)
    mlp_ratio: float = 4.0
    ln_eps: float = 1e-5


def stripe_shape(reso: int, idx: int, split: int) -> Tuple[int, int]:
    """(H_sp, W_sp) — cswin_unet.py:43-53. idx 0: vertical stripe, 1: horizontal, -1: whole image."""
    if idx == -1:
        return reso, reso
    if idx == 0:
        return reso, split
    if idx == 1:
        return split, reso
    raise ValueError(f"idx must be -1, 0 or 1, got {idx}")


# --------------------------------------------------------------------------------------
# LePE attention — explicit loops (small cases only)
# --------------------------------------------------------------------------------------
def lepe_attention_loops(q: np.ndarray, k: np.ndarray, v: np.ndarray, conv_w: np.ndarray,
                         conv_b: np.ndarray, reso: int, idx: int, split: int, heads: int,
                         scale: Optional[float] = None) -> np.ndarray:
    """SURVEY Appendix A, literally.  q,k,v: (B, L, C_b) float; conv_w: (C_b,1,3,3); returns (B,L,C_b) f64."""
    q = np.asarray(q, np.float64); k = np.asarray(k, np.float64); v = np.asarray(v, np.float64)
    w = np.asarray(conv_w, np.float64).reshape(-1, 3, 3); beta = np.asarray(conv_b, np.float64)
    B, L, C = q.shape
    H = W = reso
    assert L == H * W, "flatten img_tokens has wrong size"
    hs, ws = stripe_shape(reso, idx, split)
    if H % hs or W % ws:
        raise ValueError("resolution not divisible by stripe shape")
    d = C // heads
    sc = float(scale) if scale is not None else d ** -0.5
    out = np.zeros((B, L, C), np.float64)
    for b in range(B):
        for ih in range(H // hs):
            for iw in range(W // ws):
                tok = [(ih * hs + r) * W + (iw * ws + c) for r in range(hs) for c in range(ws)]
                for g in range(heads):
                    ch = slice(g * d, (g + 1) * d)
                    Q = q[b, tok, ch]; K = k[b, tok, ch]; V = v[b, tok, ch]
                    S = sc * (Q @ K.T)
                    S = S - S.max(axis=1, keepdims=True)
                    P = np.exp(S); P /= P.sum(axis=1, keepdims=True)
                    O = P @ V
                    for n, t in enumerate(tok):
                        r, c = divmod(n, ws)
                        acc = beta[ch].copy()
                        for dr in (-1, 0, 1):
                            for dc in (-1, 0, 1):
                                rr, cc = r + dr, c + dc
                                if 0 <= rr < hs and 0 <= cc < ws:     # WINDOW-local zero padding
                                    acc += w[ch, dr + 1, dc + 1] * V[rr * ws + cc]
                        out[b, t, ch] = O[n] + acc
    return out


# --------------------------------------------------------------------------------------
# LePE attention — vectorised
# --------------------------------------------------------------------------------------
def _window_token_table(reso: int, hs: int, ws: int) -> Tensor:
    """(nWin, N) int64: token ids of each window in the reference's order (img2windows :184-191)."""
    y = torch.arange(reso).view(reso // hs, hs, 1, 1)
    x = torch.arange(reso).view(1, 1, reso // ws, ws)
    tok = (y * reso + x).permute(0, 2, 1, 3)            # (nH, nW, hs, ws)
    return tok.reshape(-1, hs * ws)


def lepe_attention(q: Tensor, k: Tensor, v: Tensor, conv_w: Tensor, conv_b: Tensor, reso: int,
                   idx: int, split: int, heads: int, scale: Optional[float] = None,
                   return_lse: bool = False):
    """softmax(scale q k^T) v + LePE(v) per (window, head); q,k,v (B,L,C_b) any strides -> (B,L,C_b)."""
    B, L, C = q.shape
    if L != reso * reso:
        raise AssertionError("flatten img_tokens has wrong size")
    hs, ws = stripe_shape(reso, idx, split)
    if reso % hs or reso % ws:
        raise ValueError("resolution not divisible by stripe shape")
    d = C // heads
    sc = float(scale) if scale is not None else d ** -0.5
    table = _window_token_table(reso, hs, ws)                       # (nWin, N)
    nwin, N = table.shape

    def gather(t: Tensor) -> Tensor:                                 # (B, nWin, heads, N, d)
        return t[:, table.reshape(-1)].reshape(B, nwin, N, heads, d).permute(0, 1, 3, 2, 4)

    Q, K, V = gather(q), gather(k), gather(v)
    S = sc * (Q @ K.transpose(-1, -2))
    lse = torch.logsumexp(S, dim=-1)
    P = torch.exp(S - lse.unsqueeze(-1))
    O = (P @ V).permute(0, 1, 3, 2, 4).reshape(B, nwin * N, C)        # windows order
    out = torch.empty_like(O)
    out[:, table.reshape(-1)] = O                                    # scatter back (windows2img :194-202)

    # LePE: depthwise 3x3 cross-correlation with zero padding at the WINDOW border (:67-80),
    # written as nine masked image-level shifts.
    vi = v.reshape(B, reso, reso, C)
    yy = torch.arange(reso).view(reso, 1)
    xx = torch.arange(reso).view(1, reso)
    lepe = conv_b.view(1, 1, 1, C).expand(B, reso, reso, C).clone()
    wk = conv_w.reshape(C, 3, 3)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            ny, nx = yy + dy, xx + dx
            ok = (ny >= 0) & (ny < reso) & (nx >= 0) & (nx < reso)
            ok = ok & (torch.div(ny.clamp(0, reso - 1), hs, rounding_mode="floor") == torch.div(yy, hs, rounding_mode="floor"))
            ok = ok & (torch.div(nx.clamp(0, reso - 1), ws, rounding_mode="floor") == torch.div(xx, ws, rounding_mode="floor"))
            shifted = torch.roll(vi, shifts=(-dy, -dx), dims=(1, 2))
            lepe = lepe + shifted * ok.view(1, reso, reso, 1).to(vi.dtype) * wk[:, dy + 1, dx + 1].view(1, 1, 1, C)
    out = out + lepe.reshape(B, L, C)
    if return_lse:
        # (B, L, heads) in image token order
        lse_img = torch.empty(B, L, heads, dtype=lse.dtype)
        lse_img[:, table.reshape(-1)] = lse.permute(0, 1, 3, 2).reshape(B, nwin * N, heads)
        return out, lse_img
    return out


# --------------------------------------------------------------------------------------
# block / merge / carafe
# --------------------------------------------------------------------------------------
def _ln(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _gelu(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def cswin_block(sd: Mapping[str, Tensor], prefix: str, x: Tensor, reso: int, heads: int, split: int,
                last_stage: bool = False, eps: float = 1e-5, qk_scale: Optional[float] = None,
                sample_scale: Optional[Tensor] = None) -> Tensor:
    """CSWinBlock.forward (cswin_unet.py:160-181). `sample_scale` (B,) = DropPath m_b/(1-p) or None."""
    p = lambda n: sd[prefix + n]
    B, L, C = x.shape
    if L != reso * reso:
        raise AssertionError("flatten img_tokens has wrong size")
    one_branch = last_stage or reso == split
    u = _ln(x, p("norm1.weight"), p("norm1.bias"), eps)
    z = u @ p("qkv.weight").T
    if (prefix + "qkv.bias") in sd:
        z = z + p("qkv.bias")
    q, k, v = z[..., :C], z[..., C:2 * C], z[..., 2 * C:]
    if one_branch:
        a = lepe_attention(q, k, v, p("attns.0.get_v.weight"), p("attns.0.get_v.bias"), reso, -1, split,
                           heads, qk_scale)
    else:
        h = C // 2
        a0 = lepe_attention(q[..., :h], k[..., :h], v[..., :h], p("attns.0.get_v.weight"),
                            p("attns.0.get_v.bias"), reso, 0, split, heads // 2, qk_scale)
        a1 = lepe_attention(q[..., h:], k[..., h:], v[..., h:], p("attns.1.get_v.weight"),
                            p("attns.1.get_v.bias"), reso, 1, split, heads // 2, qk_scale)
        a = torch.cat([a0, a1], dim=-1)
    y = a @ p("proj.weight").T + p("proj.bias")
    if sample_scale is not None:
        y = y * sample_scale.view(B, 1, 1)
    x1 = x + y
    hdn = _gelu(_ln(x1, p("norm2.weight"), p("norm2.bias"), eps) @ p("mlp.fc1.weight").T + p("mlp.fc1.bias"))
    y2 = hdn @ p("mlp.fc2.weight").T + p("mlp.fc2.bias")
    if sample_scale is not None:
        y2 = y2 * sample_scale.view(B, 1, 1)
    return x1 + y2


def _tokens_to_image(x: Tensor) -> Tuple[Tensor, int]:
    B, L, C = x.shape
    r = int(round(math.sqrt(L)))
    assert r * r == L
    return x.reshape(B, r, r, C).permute(0, 3, 1, 2), r


def merge_block(sd: Mapping[str, Tensor], prefix: str, x: Tensor, eps: float = 1e-5) -> Tensor:
    """Merge_Block.forward (:211-220): 3x3 stride-2 pad-1 conv on the token image, then LayerNorm."""
    img, _ = _tokens_to_image(x)
    y = F.conv2d(img, sd[prefix + "conv.weight"], sd[prefix + "conv.bias"], stride=2, padding=1)
    y = y.flatten(2).transpose(1, 2)
    return _ln(y, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], eps)


def carafe(sd: Mapping[str, Tensor], prefix: str, x: Tensor, up: int) -> Tensor:
    """CARAFE (:232-269, up=2) / CARAFE4 (:282-319, up=4) in the restated form of SURVEY Appendix A:

        E = encoder(down(X))                    channel t*s^2 + a*s + e  (tap t, sub-pixel (a,e))
        kappa[t] = softmax_t E[t*s^2+a*s+e, y, x]
        Y[c, s*y+a, s*x+e] = sum_t kappa[t] * Xpad[c, y+dy-1, x+dx-1]          (image-level zero pad)
        result = out_conv(Y), flattened row-major to (B, s^2 L, C_out)
    """
    img, r = _tokens_to_image(x)
    B, C = img.shape[:2]
    s = up
    e = F.conv2d(img, sd[prefix + "down.weight"], sd[prefix + "down.bias"])
    e = F.conv2d(e, sd[prefix + "encoder.weight"], sd[prefix + "encoder.bias"], padding=1)   # (B, 9 s^2, r, r)
    kap = torch.softmax(e.reshape(B, 9, s, s, r, r), dim=1)                                   # (B,9,a,e,y,x)
    xp = F.pad(img, (1, 1, 1, 1))
    y = torch.zeros(B, C, r, s, r, s, dtype=img.dtype)
    for t in range(9):
        dy, dx = divmod(t, 3)
        nb = xp[:, :, dy:dy + r, dx:dx + r]                                                     # (B,C,y,x)
        kt = kap[:, t]                                                                          # (B,a,e,y,x)
        y = y + nb.view(B, C, r, 1, r, 1) * kt.permute(0, 3, 1, 4, 2).reshape(B, 1, r, s, r, s)
    y = y.reshape(B, C, r * s, r * s)
    o = F.conv2d(y, sd[prefix + "out.weight"], sd[prefix + "out.bias"])
    return o.flatten(2).transpose(1, 2)


# --------------------------------------------------------------------------------------
# whole model
# --------------------------------------------------------------------------------------
def stage_plan(cfg: OracleConfig) -> List[Tuple[str, int, int, int, int, bool]]:
    """[(stage name, dim, reso, heads, split, last_stage)] for encoder stages 1..4."""
    r = cfg.img_size // 4
    plan = []
    for i in range(4):
        plan.append((f"stage{i + 1}", cfg.embed_dim << i, r >> i, cfg.num_heads[i],
                     cfg.split_size[i if i < 3 else -1], i == 3))
    return plan


def cswin_unet_forward(sd: Mapping[str, Tensor], x: Tensor, cfg: OracleConfig = OracleConfig(),
                       taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """CSWinTransformer.forward (:546-554), eval mode (DropPath = identity). x: (B,3,H,W) -> (B,classes,H,W)."""
    eps = cfg.ln_eps
    plan = stage_plan(cfg)
    t = F.conv2d(x, sd["stage1_conv_embed.0.weight"], sd["stage1_conv_embed.0.bias"], stride=4, padding=2)
    t = t.flatten(2).transpose(1, 2)
    t = _ln(t, sd["stage1_conv_embed.2.weight"], sd["stage1_conv_embed.2.bias"], eps)
    skips = []
    for si, (name, dim, reso, heads, split, last) in enumerate(plan):
        for bi in range(cfg.depth[si]):
            t = cswin_block(sd, f"{name}.{bi}.", t, reso, heads, split, last, eps)
        if taps is not None:
            taps[name] = t
        if si < 3:
            skips.append(t)
            t = merge_block(sd, f"merge{si + 1}.", t, eps)
    t = _ln(t, sd["norm.weight"], sd["norm.bias"], eps)
    for si in (3, 2, 1, 0):
        name, dim, reso, heads, split, last = plan[si]
        for bi in range(cfg.depth[si]):
            t = cswin_block(sd, f"stage_up{si + 1}.{bi}.", t, reso, heads, split, last, eps)
        if taps is not None:
            taps[f"stage_up{si + 1}"] = t
        if si > 0:
            t = carafe(sd, f"upsample{si + 1}.", t, 2)
            t = torch.cat([skips[si - 1], t], dim=-1)
            t = t @ sd[f"concat_linear{si + 1}.weight"].T + sd[f"concat_linear{si + 1}.bias"]
    t = _ln(t, sd["norm_up.weight"], sd["norm_up.bias"], eps)
    t = carafe(sd, "upsample1.", t, 4)                     # (B, 16 L, 64)
    B = t.shape[0]
    side = cfg.img_size
    img = t.reshape(B, side, side, -1).permute(0, 3, 1, 2)
    return F.conv2d(img, sd["output.weight"])


def state_dict_shapes(cfg: OracleConfig = OracleConfig()) -> Dict[str, Tuple[int, ...]]:
    """Key -> shape of the reference's CSWinTransformer.state_dict() (463 tensors at T224) in its order."""
    shapes: Dict[str, Tuple[int, ...]] = {}
    E = cfg.embed_dim

    def block(prefix: str, C: int, one_branch: bool):
        shapes[prefix + "qkv.weight"] = (3 * C, C); shapes[prefix + "qkv.bias"] = (3 * C,)
        shapes[prefix + "norm1.weight"] = (C,); shapes[prefix + "norm1.bias"] = (C,)
        shapes[prefix + "proj.weight"] = (C, C); shapes[prefix + "proj.bias"] = (C,)
        for i in range(1 if one_branch else 2):
            cb = C if one_branch else C // 2
            shapes[prefix + f"attns.{i}.get_v.weight"] = (cb, 1, 3, 3)
            shapes[prefix + f"attns.{i}.get_v.bias"] = (cb,)
        hid = int(C * cfg.mlp_ratio)
        shapes[prefix + "mlp.fc1.weight"] = (hid, C); shapes[prefix + "mlp.fc1.bias"] = (hid,)
        shapes[prefix + "mlp.fc2.weight"] = (C, hid); shapes[prefix + "mlp.fc2.bias"] = (C,)
        shapes[prefix + "norm2.weight"] = (C,); shapes[prefix + "norm2.bias"] = (C,)

    def carafe_shapes(prefix: str, C: int, Cout: int, s: int):
        shapes[prefix + "down.weight"] = (C // 4, C, 1, 1); shapes[prefix + "down.bias"] = (C // 4,)
        shapes[prefix + "encoder.weight"] = (9 * s * s, C // 4, 3, 3); shapes[prefix + "encoder.bias"] = (9 * s * s,)
        shapes[prefix + "out.weight"] = (Cout, C, 1, 1); shapes[prefix + "out.bias"] = (Cout,)

    shapes["stage1_conv_embed.0.weight"] = (E, cfg.in_chans, 7, 7); shapes["stage1_conv_embed.0.bias"] = (E,)
    shapes["stage1_conv_embed.2.weight"] = (E,); shapes["stage1_conv_embed.2.bias"] = (E,)
    plan = stage_plan(cfg)
    for si, (name, dim, reso, heads, split, last) in enumerate(plan):
        for bi in range(cfg.depth[si]):
            block(f"{name}.{bi}.", dim, last or reso == split)
        if si < 3:
            shapes[f"merge{si + 1}.conv.weight"] = (2 * dim, dim, 3, 3); shapes[f"merge{si + 1}.conv.bias"] = (2 * dim,)
            shapes[f"merge{si + 1}.norm.weight"] = (2 * dim,); shapes[f"merge{si + 1}.norm.bias"] = (2 * dim,)
    shapes["norm.weight"] = (8 * E,); shapes["norm.bias"] = (8 * E,)
    for si in (3, 2, 1, 0):
        name, dim, reso, heads, split, last = plan[si]
        for bi in range(cfg.depth[si]):
            block(f"stage_up{si + 1}.{bi}.", dim, last or reso == split)
        if si > 0:
            carafe_shapes(f"upsample{si + 1}.", dim, dim // 2, 2)
            shapes[f"concat_linear{si + 1}.weight"] = (dim // 2, dim); shapes[f"concat_linear{si + 1}.bias"] = (dim // 2,)
    carafe_shapes("upsample1.", E, 64, 4)
    shapes["norm_up.weight"] = (E,); shapes["norm_up.bias"] = (E,)
    shapes["output.weight"] = (cfg.num_classes, E, 1, 1)
    return shapes


# --------------------------------------------------------------------------------------
# losses and metrics (trainer.py:55-57, utils.py:9-58)
# --------------------------------------------------------------------------------------
def dice_loss(logits: Tensor, target: Tensor, n_classes: int) -> Tensor:
    """DiceLoss(softmax=True) of utils.py:32-45: mean over classes of 1 - (2 I + s)/(Z + Y + s), sums over the batch."""
    prob = torch.softmax(logits, dim=1)
    smooth = 1e-5
    loss = logits.new_zeros(())
    for c in range(n_classes):
        t = (target == c).to(prob.dtype)
        s = prob[:, c]
        inter = (s * t).sum(); ysum = (t * t).sum(); zsum = (s * s).sum()
        loss = loss + (1.0 - (2 * inter + smooth) / (zsum + ysum + smooth))
    return loss / n_classes


def seg_loss(logits: Tensor, target: Tensor, n_classes: int) -> Tensor:
    """0.4 CE + 0.6 Dice (trainer.py:55-57)."""
    return 0.4 * F.cross_entropy(logits, target.long()) + 0.6 * dice_loss(logits, target, n_classes)


def _surface_distances(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    from scipy.ndimage import binary_erosion, distance_transform_edt, generate_binary_structure
    fp = generate_binary_structure(a.ndim, 1)
    a = a.astype(bool); b = b.astype(bool)
    a_border = a ^ binary_erosion(a, structure=fp, iterations=1)
    b_border = b ^ binary_erosion(b, structure=fp, iterations=1)
    dt = distance_transform_edt(~b_border)
    return dt[a_border]


def dice_hd95_percase(pred: np.ndarray, gt: np.ndarray) -> Tuple[float, float]:
    """calculate_metric_percase (utils.py:48-58) with medpy's binary.dc / binary.hd95 restated
    (medpy is not installed: dc = 2|A&B|/(|A|+|B|); hd95 = 95th percentile of the two directed
    surface-distance sets, 1-connectivity borders, unit voxel spacing)."""
    pred = (np.asarray(pred) > 0); gt = (np.asarray(gt) > 0)
    if pred.sum() > 0 and gt.sum() > 0:
        inter = np.count_nonzero(pred & gt)
        dc = 2.0 * inter / float(np.count_nonzero(pred) + np.count_nonzero(gt))
        hd = np.percentile(np.hstack((_surface_distances(pred, gt), _surface_distances(gt, pred))), 95)
        return float(dc), float(hd)
    if pred.sum() > 0 and gt.sum() == 0:
        return 1.0, 0.0
    return 0.0, 0.0


def volume_metrics(pred: np.ndarray, label: np.ndarray, classes: int) -> List[Tuple[float, float]]:
    """Per-class (dice, hd95) over a 3-D prediction / label volume (utils.py:88-90)."""
    return [dice_hd95_percase(pred == c, label == c) for c in range(1, classes)]
