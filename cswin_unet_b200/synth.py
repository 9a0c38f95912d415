"""Deterministic synthetic weights and inputs (no datasets / checkpoints exist offline).

Everything is generated with numpy's PCG64 (`np.random.default_rng`), whose stream is
stable across numpy versions, and is keyed on the *name* of the tensor, so the values do
not depend on module construction order or on torch's RNG.  The same functions feed the
reference (when the golden fixtures are generated), the oracle, the CUDA path, the tests
and bench.py.

Weight scales are chosen so that the network is numerically "alive": Linear / conv
weights are N(0, gain/fan_in) (so q.k scores have O(1) spread and softmax is far from
uniform), LayerNorm gains are 1 + 0.1 N(0,1).  The reference's own init
(trunc-normal 0.02, /root/reference/networks/cswin_unet.py:444-451) gives nearly flat
attention, which would hide softmax / masking bugs.
"""
from __future__ import annotations

import zlib
from typing import Dict, Mapping, Sequence, Tuple

import numpy as np


def _rng(name: str, seed: int) -> np.random.Generator:
    return np.random.default_rng([zlib.crc32(name.encode()) & 0xFFFFFFFF, seed & 0xFFFFFFFF])


def _is_norm_key(key: str, shape: Sequence[int]) -> bool:
    if len(shape) != 1:
        return False
    parts = key.split(".")
    owner = parts[-2] if len(parts) >= 2 else ""
    # norm1 / norm2 / norm / norm_up / merge*.norm / stage1_conv_embed.2
    return owner.startswith("norm") or (owner == "2" and "conv_embed" in key)


def synth_tensor(key: str, shape: Sequence[int], seed: int = 0) -> np.ndarray:
    """One float32 tensor for state_dict entry `key` of the given shape."""
    shape = tuple(int(s) for s in shape)
    g = _rng(key, seed)
    leaf = key.split(".")[-1]
    if _is_norm_key(key, shape):
        if leaf == "weight":
            return (1.0 + 0.1 * g.standard_normal(shape)).astype(np.float32)
        return (0.05 * g.standard_normal(shape)).astype(np.float32)
    if leaf == "bias":
        return (0.02 * g.standard_normal(shape)).astype(np.float32)
    # weight of a Linear (out,in) / Conv2d (out, in/groups, kh, kw)
    fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])
    gain = 1.0
    if ".get_v." in key:           # depthwise 3x3 LePE conv: keep it comparable to attn output
        gain = 0.5
    return (gain * g.standard_normal(shape) / np.sqrt(max(fan_in, 1))).astype(np.float32)


def synth_state_dict(shapes: Mapping[str, Sequence[int]], seed: int = 0) -> Dict[str, np.ndarray]:
    return {k: synth_tensor(k, s, seed) for k, s in shapes.items()}


def synth_image_batch(batch: int, chans: int = 3, size: int = 224, seed: int = 0,
                      kind: str = "randn") -> np.ndarray:
    """(B, chans, size, size) float32. kind: 'randn'  or 'ct' ([0,1] blobs)."""
    g = _rng(f"image/{kind}/{batch}x{chans}x{size}", seed)
    if kind == "randn":
        return g.standard_normal((batch, chans, size, size)).astype(np.float32)
    if kind == "ct":
        yy, xx = np.mgrid[0:size, 0:size].astype(np.float32) / size
        out = np.zeros((batch, chans, size, size), np.float32)
        for b in range(batch):
            img = np.zeros((size, size), np.float32)
            for _ in range(6):
                cy, cx, r, a = g.uniform(0.2, 0.8), g.uniform(0.2, 0.8), g.uniform(0.05, 0.25), g.uniform(0.2, 1.0)
                img += a * np.exp(-(((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r)))
            img = img / max(float(img.max()), 1e-6)
            img += 0.02 * g.standard_normal((size, size)).astype(np.float32)
            out[b, :] = np.clip(img, 0.0, 1.0)
        return out
    raise ValueError(f"unknown synthetic image kind {kind!r}")


def synth_qkv(batch: int, reso: int, dim: int, seed: int = 0) -> np.ndarray:
    """The (B, L, 3, C) buffer a qkv Linear would have produced.

    `LePEAttention` sees it as `base.permute(2,0,1,3)[..., off:off+C_b]`, i.e. a
    non-contiguous (3,B,L,C_b) view with strides (C, 3LC, 3C, 1).
    """
    g = _rng(f"qkv/{batch}/{reso}/{dim}", seed)
    return g.standard_normal((batch, reso * reso, 3, dim)).astype(np.float32)


def synth_labels(batch: int, size: int = 224, classes: int = 9, seed: int = 0) -> np.ndarray:
    """(B, size, size) int64 blob label maps with `classes` classes (0 = background)."""
    g = _rng(f"labels/{batch}/{size}/{classes}", seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32) / size
    lab = np.zeros((batch, size, size), np.int64)
    for b in range(batch):
        for c in range(1, classes):
            cy, cx, r = g.uniform(0.15, 0.85), g.uniform(0.15, 0.85), g.uniform(0.04, 0.14)
            lab[b][((yy - cy) ** 2 + (xx - cx) ** 2) < r * r] = c
    return lab


# the seven LePEAttention configurations of cswin_tiny_224_lite:
# (C_b, reso, idx, split, heads_b)
LEPE_CONFIGS_T224: Tuple[Tuple[int, int, int, int, int], ...] = (
    (32, 56, 0, 1, 1), (32, 56, 1, 1, 1),
    (64, 28, 0, 2, 2), (64, 28, 1, 2, 2),
    (128, 14, 0, 7, 4), (128, 14, 1, 7, 4),
    (512, 7, -1, 7, 16),
)
