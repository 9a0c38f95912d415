#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py [--ref /root/reference]

The reference's `networks/cswin_unet.py` is imported as-is; the only thing injected is a
3-symbol `timm.models.layers` shim (timm is not installed) and empty
`medpy` / `SimpleITK` modules so that `utils.py` (DiceLoss) imports.  Inputs and weights come
from `cswin_unet_b200.synth` (numpy PCG64 keyed on tensor names), so the tests regenerate them
bit-identically and only OUTPUTS are stored.  Large outputs are stored as every `stride`-th token
row plus float64 checksums of the full tensor.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from cswin_unet_b200 import synth  # noqa: E402


def install_shims():
    class DropPath(torch.nn.Module):
        def __init__(self, drop_prob=0., scale_by_keep=True):
            super().__init__()
            self.drop_prob = drop_prob
            self.scale_by_keep = scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0. or not self.training:
                return x
            keep = 1 - self.drop_prob
            m = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                m.div_(keep)
            return x * m

    layers = types.ModuleType("timm.models.layers")
    layers.DropPath = DropPath
    layers.to_2tuple = lambda x: (x, x)
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm = types.ModuleType("timm"); models = types.ModuleType("timm.models")
    timm.models = models; models.layers = layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    medpy = types.ModuleType("medpy"); medpy.metric = types.ModuleType("medpy.metric")
    sys.modules.update({"medpy": medpy, "medpy.metric": medpy.metric, "SimpleITK": types.ModuleType("SimpleITK")})


def pack(t: torch.Tensor, stride_if_big: int = 5, big: int = 60_000):
    """-> dict with 'rows' (sampled token rows), 'stride', 'shape', 'sum', 'abssum' (float64)."""
    a = t.detach().double().cpu().numpy()
    flat = a.reshape(-1, a.shape[-1])
    stride = stride_if_big if a.size > big else 1
    return {"rows": flat[::stride].astype(np.float32), "stride": np.int64(stride),
            "shape": np.array(a.shape, np.int64), "sum": np.float64(a.sum()), "abssum": np.float64(np.abs(a).sum())}


def save(name: str, **groups):
    flat = {}
    for g, d in groups.items():
        if isinstance(d, dict):
            for k, v in d.items():
                flat[f"{g}.{k}"] = v
        else:
            flat[g] = d
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **flat)
    print(f"wrote {path}  {os.path.getsize(path) / 1024:.0f} KiB")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def load_synth(module: torch.nn.Module, prefix: str, seed: int):
    sd = {k: T(synth.synth_tensor(prefix + k, tuple(v.shape), seed)) for k, v in module.state_dict().items()}
    module.load_state_dict(sd, strict=True)
    return sd


# extra LePE configs beyond T224: small / 512^2-like / odd head dims
LEPE_EXTRA = ((32, 16, 0, 2, 1), (64, 16, 1, 2, 2), (64, 8, -1, 8, 2), (128, 16, 0, 8, 4), (48, 12, 1, 3, 3))
# BASELINE configs[4] (512^2, split [1,2,8,8]): stripe windows of 256 tokens (32x8, 8x32, 16x16), ragged 192 (24x8) and 196 (14x14)
LEPE_WIDE = ((128, 32, 0, 8, 4), (128, 32, 1, 8, 4), (512, 16, -1, 8, 16), (64, 24, 0, 8, 2), (64, 14, -1, 14, 2))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", default="", help="comma-separated groups to (re)write: lepe_t224, lepe_extra, lepe_wide, rest; default all")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    install_shims()
    sys.path.insert(0, args.ref)
    torch.set_grad_enabled(True)
    torch.manual_seed(0)
    import networks.cswin_unet as ref          # the unmodified reference

    # ---------------- LePEAttention forward + backward ----------------
    for tag, cfgs, B in (("t224", synth.LEPE_CONFIGS_T224, 1), ("extra", LEPE_EXTRA, 2), ("wide", LEPE_WIDE, 1)):
        if only and f"lepe_{tag}" not in only:
            continue
        out = {}
        for (cb, reso, idx, split, heads) in cfgs:
            m = ref.LePEAttention(cb, resolution=reso, idx=idx, split_size=split, num_heads=heads).double()
            load_synth(m, f"lepe/{cb}/{reso}/{idx}/", 1)
            full_c = cb if idx == -1 else 2 * cb
            base = T(synth.synth_qkv(B, reso, full_c, seed=0)).double().requires_grad_(True)
            off = cb if idx == 1 else 0
            qkv = base.permute(2, 0, 1, 3)[..., off:off + cb]          # the real strided view
            y = m(qkv)
            key = f"c{cb}_r{reso}_i{idx}_s{split}_h{heads}"
            out[key] = pack(y)
            # backward with a fixed synthetic upstream gradient
            gup = T(synth.synth_tensor(f"lepe_grad/{key}", tuple(y.shape), 2)).double()
            gb, gw, gbias = torch.autograd.grad(y, [base, m.get_v.weight, m.get_v.bias], gup)
            out[key + "_dqkv"] = pack(gb[..., off:off + cb].permute(2, 0, 1, 3).reshape(3 * B, reso * reso, cb))
            out[key + "_dw"] = pack(gw.reshape(cb, 9))
            out[key + "_db"] = pack(gbias.reshape(1, cb))
        save(f"lepe_{tag}", **out)

    if not only or "model_512" in only:
        # ---------------- whole model, BASELINE configs[4]: 512^2 input, 3 classes, split [1,2,8,8] ----------------
        m512 = ref.CSWinTransformer(img_size=512, patch_size=4, in_chans=3, num_classes=3, embed_dim=64,
                                    depth=[1, 2, 9, 1], split_size=[1, 2, 8, 8], num_heads=[2, 4, 8, 16],
                                    mlp_ratio=4., qkv_bias=True, qk_scale=None, drop_rate=0., drop_path_rate=0.2).eval()
        shapes512 = {k: tuple(v.shape) for k, v in m512.state_dict().items()}
        m512.load_state_dict({k: T(v) for k, v in synth.synth_state_dict(shapes512, seed=1234).items()}, strict=True)
        x = T(synth.synth_image_batch(1, 3, 512, seed=0, kind="ct"))
        with torch.no_grad():
            logits64 = m512.double()(x.double())
        out = {"logits_ct": pack(logits64.permute(0, 2, 3, 1), stride_if_big=37),       # rows = pixels, 3 classes
               "argmax_ct": logits64.argmax(1).to(torch.uint8).numpy(),
               "key_shapes": np.array([k + ":" + ",".join(map(str, v)) for k, v in shapes512.items()])}
        save("model_512", **out)
    if only and "rest" not in only:
        return
    # ---------------- CSWinBlock forward (one per stage) ----------------
    out = {}
    for (dim, reso, heads, split, last) in ((64, 56, 2, 1, False), (128, 28, 4, 2, False), (256, 14, 8, 7, False), (512, 7, 16, 7, True)):
        m = ref.CSWinBlock(dim=dim, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).double().eval()
        load_synth(m, f"block/{dim}/", 3)
        x = T(synth.synth_tensor(f"block_in/{dim}", (2, reso * reso, dim), 4)).double()
        out[f"d{dim}"] = pack(m(x))
    save("block", **out)

    # ---------------- Merge_Block, CARAFE, CARAFE4 ----------------
    out = {}
    for (dim, reso) in ((64, 56), (128, 28), (256, 14)):
        m = ref.Merge_Block(dim, dim * 2).double().eval()
        load_synth(m, f"merge/{dim}/", 5)
        x = T(synth.synth_tensor(f"merge_in/{dim}", (2, reso * reso, dim), 6)).double()
        out[f"merge_d{dim}"] = pack(m(x))
    for (cls, dim, dout, reso, up) in ((ref.CARAFE, 512, 256, 7, 2), (ref.CARAFE, 128, 64, 28, 2), (ref.CARAFE4, 64, 64, 14, 4)):
        m = cls(dim, dout).double().eval()
        load_synth(m, f"carafe/{dim}/{up}/", 7)
        x = T(synth.synth_tensor(f"carafe_in/{dim}/{up}", (2, reso * reso, dim), 8)).double()
        out[f"carafe_d{dim}_u{up}"] = pack(m(x))
    save("merge_carafe", **out)

    # ---------------- whole model, T224 ----------------
    model = ref.CSWinTransformer(img_size=224, patch_size=4, in_chans=3, num_classes=9, embed_dim=64,
                                 depth=[1, 2, 9, 1], split_size=[1, 2, 7, 7], num_heads=[2, 4, 8, 16],
                                 mlp_ratio=4., qkv_bias=True, qk_scale=None, drop_rate=0., drop_path_rate=0.2).eval()
    ref_sd = model.state_dict()
    keys = np.array(list(ref_sd.keys()))
    shapes = {k: tuple(v.shape) for k, v in ref_sd.items()}
    sd = {k: T(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}
    model.load_state_dict(sd, strict=True)
    out = {"keys": keys, "key_shapes": np.array([",".join(map(str, shapes[k])) for k in keys]),
           "n_params": np.int64(sum(int(np.prod(s)) for s in shapes.values()))}
    for kind in ("randn", "ct"):
        x = T(synth.synth_image_batch(2, 3, 224, seed=0, kind=kind))
        with torch.no_grad():
            logits32 = model(x)
            logits64 = model.double()(x.double())
            model.float()
        out[f"logits_{kind}"] = pack(logits64.permute(0, 2, 3, 1), stride_if_big=7)      # rows = pixels, 9 classes
        out[f"fp32_vs_fp64_maxabs_{kind}"] = np.float64((logits32.double() - logits64).abs().max())
        out[f"argmax_{kind}"] = logits64.argmax(1).to(torch.uint8).numpy()
        srt = logits64.sort(dim=1, descending=True).values
        out[f"margin_{kind}"] = (srt[:, 0] - srt[:, 1]).float().numpy().astype(np.float16)
    # stage taps (encoder skips) for localising a mismatch
    x = T(synth.synth_image_batch(2, 3, 224, seed=0, kind="randn")).double()
    model.double()
    with torch.no_grad():
        feats = model.forward_features(x)
        out["tap_x1"] = pack(model.x1); out["tap_x2"] = pack(model.x2); out["tap_x3"] = pack(model.x3)
        out["tap_bottleneck"] = pack(feats)
        up = model.forward_up_features(feats)
        out["tap_up"] = pack(up)
    model.float()
    save("model_t224", **out)

    # ---------------- whole model with the reference's init DISTRIBUTIONS (bf16 tolerance is quoted on these) -----
    sd = {k: T(v) for k, v in synth.synth_state_dict(shapes, seed=1234, mode="refinit").items()}
    model.load_state_dict(sd, strict=True)
    out = {}
    for kind in ("randn", "ct"):
        x = T(synth.synth_image_batch(2, 3, 224, seed=0, kind=kind))
        with torch.no_grad():
            logits64 = model.double()(x.double())
            model.float()
            logits32 = model(x)
            import copy
            lbf = copy.deepcopy(model).bfloat16()(x.bfloat16()).double()      # a COPY: .bfloat16() rounds weights in place
        out[f"logits_{kind}"] = pack(logits64.permute(0, 2, 3, 1), stride_if_big=7)
        out[f"argmax_{kind}"] = logits64.argmax(1).to(torch.uint8).numpy()
        srt = logits64.sort(dim=1, descending=True).values
        out[f"margin_{kind}"] = (srt[:, 0] - srt[:, 1]).float().numpy().astype(np.float16)
        out[f"fp32_vs_fp64_maxabs_{kind}"] = np.float64((logits32.double() - logits64).abs().max())
        # yard-stick: the reference's OWN bf16 path (CPU) against its fp64 path
        out[f"refbf16_maxabs_{kind}"] = np.float64((lbf - logits64).abs().max())
        out[f"refbf16_argmax_agree_{kind}"] = np.float64((lbf.argmax(1) == logits64.argmax(1)).double().mean())
    save("model_refinit", **out)

    # ---------------- loss (trainer.py:55-57 with utils.DiceLoss) ----------------
    import utils as ref_utils
    logits = T(synth.synth_tensor("loss/logits", (2, 9, 32, 32), 9)).double().requires_grad_(True)
    labels = T(synth.synth_labels(2, 32, 9, seed=9))
    ce = torch.nn.CrossEntropyLoss()(logits, labels.long())
    dl = ref_utils.DiceLoss(9)(logits, labels, softmax=True)
    loss = 0.4 * ce + 0.6 * dl
    (g,) = torch.autograd.grad(loss, logits)
    save("loss", loss=np.float64(loss.item()), ce=np.float64(ce.item()), dice=np.float64(dl.item()),
         grad=g.detach().numpy().astype(np.float32))


if __name__ == "__main__":
    main()
