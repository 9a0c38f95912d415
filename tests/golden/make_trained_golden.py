#!/usr/bin/env python
"""Golden fixtures with DECISIVE logits: the unmodified reference, briefly trained.

    python tests/golden/make_trained_golden.py --config t224|512 [--steps 300]

Runs only in the build container (imports /root/reference through baseline/ref_loader.py).  At random init the logits are
nearly flat (11 % of the pixels have a top-2 margin < 0.01), so per-pixel argmax agreement >= 99.9 % / Dice / HD95 <= 1e-3
cannot be decided for ANY reduced-precision path — the reference's own bf16 autocast reaches 99.4 %.  This script trains the
reference with its own loop (trainer.py:42-63: SGD lr .05 / momentum .9 / wd 1e-4, poly decay, 0.4 CE + 0.6 Dice with the
reference's own DiceLoss, utils.py:9-45) on the seeded blob task of `synth.synth_seg_batch`.

To keep the fixture small (the full model is 94 MB) only the sub-network on the full-resolution path is trained — stem,
stage1, concat_linear2, stage_up1, norm_up, upsample1 (CARAFE4), output: ~146 k parameters, stored here — while every other
parameter keeps its `synth` reference-init value (regenerated from its name by the tests).  The deep path is evaluated under
no_grad during training (it is frozen), which is only a cheaper way to obtain the same kind of weights; every golden OUTPUT
below comes from the plain `model(x)` call of the unmodified reference in fp32.

Stored: trained tensors; logits of the reference on held-out slices (fp32, sampled rows + checksums); full arg-max maps; the
reference's label volume for a synthetic volume run through the test_single_volume loop (utils.py:61-80) and its per-class
Dice / HD95 against the task's ground truth; the reference's own bf16-autocast agreement as the yard-stick.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from baseline import ref_loader  # noqa: E402
from cswin_unet_b200 import synth  # noqa: E402
from oracle import cswin_oracle as O  # noqa: E402  (dice / hd95 restatement for the stored metrics)

CONFIGS = {
    "t224": dict(img_size=224, num_classes=9, split_size=[1, 2, 7, 7], batch=8, vol=(12, 256)),
    "512": dict(img_size=512, num_classes=3, split_size=[1, 2, 8, 8], batch=2, vol=(4, 512)),
}
TRAINED_PREFIXES = ("stage1_conv_embed.", "stage1.", "concat_linear2.", "stage_up1.", "norm_up.", "upsample1.", "output.")


def is_trained(key: str) -> bool:
    return key.startswith(TRAINED_PREFIXES)


def reference_dice_loss(n_classes):
    """The reference's own DiceLoss (utils.py:9-45); utils.py imports medpy / SimpleITK at module top, which are stubbed."""
    for name in ("medpy", "medpy.metric", "SimpleITK"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["medpy"].metric = sys.modules["medpy.metric"]
    sys.path.insert(0, ref_loader.REF_SOURCE)
    import utils as ref_utils
    return ref_utils.DiceLoss(n_classes)


def pack(t: torch.Tensor, stride: int):
    a = t.detach().double().cpu().numpy()
    flat = a.reshape(-1, a.shape[-1])
    return {"rows": flat[::stride].astype(np.float32), "stride": np.int64(stride), "shape": np.array(a.shape, np.int64),
            "sum": np.float64(a.sum()), "abssum": np.float64(np.abs(a).sum())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="t224", choices=sorted(CONFIGS))
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--lr", type=float, default=0.05)
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    S, NC, B = cfg["img_size"], cfg["num_classes"], cfg["batch"]
    torch.manual_seed(1234)
    torch.set_num_threads(os.cpu_count() or 1)

    m = ref_loader.build_reference_model(img_size=S, num_classes=NC, split_size=cfg["split_size"]).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234, mode="refinit").items()}, strict=True)
    trained = [k for k in shapes if is_trained(k)]
    for k, p in m.named_parameters():
        p.requires_grad_(is_trained(k))
    n_tr = sum(int(np.prod(shapes[k])) for k in trained)
    print(f"[{args.config}] training {len(trained)} tensors / {n_tr} parameters of {sum(int(np.prod(s)) for s in shapes.values())}")

    ce = torch.nn.CrossEntropyLoss()
    dice = reference_dice_loss(NC)
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=args.lr, momentum=0.9, weight_decay=1e-4)

    def train_forward(x):
        """Same function as CSWinTransformer.forward (cswin_unet.py:462-554), with the frozen deep path under no_grad."""
        x = m.stage1_conv_embed(x)
        x = m.pos_drop(x)
        for blk in m.stage1:
            x = blk(x)
        x1 = x
        with torch.no_grad():
            y = m.merge1(x1)
            for blk in m.stage2:
                y = blk(y)
            x2 = y
            y = m.merge2(y)
            for blk in m.stage3:
                y = blk(y)
            x3 = y
            y = m.merge3(y)
            for blk in m.stage4:
                y = blk(y)
            y = m.norm(y)
            for blk in m.stage_up4:
                y = blk(y)
            y = m.concat_linear4(torch.cat([x3, m.upsample4(y)], -1))
            for blk in m.stage_up3:
                y = blk(y)
            y = m.concat_linear3(torch.cat([x2, m.upsample3(y)], -1))
            for blk in m.stage_up2:
                y = blk(y)
            y = m.upsample2(y)
        y = m.concat_linear2(torch.cat([x1, y], -1))
        for blk in m.stage_up1:
            y = blk(y)
        y = m.norm_up(y)
        return m.up_x4(y)

    # sanity: the piecewise forward IS the reference forward
    with torch.no_grad():
        x0 = torch.from_numpy(synth.synth_seg_batch(1, S, NC, seed=999)[0]).repeat(1, 3, 1, 1)
        assert torch.equal(train_forward(x0), m(x0)), "piecewise training forward differs from the reference forward"

    t0 = time.time()
    for it in range(args.steps):
        xs, ys = synth.synth_seg_batch(B, S, NC, seed=it)
        x = torch.from_numpy(xs).repeat(1, 3, 1, 1)               # CSwinUnet.forward's 1 -> 3 channel repeat (vision_transformer.py:40-41)
        y = torch.from_numpy(ys)
        out = train_forward(x)
        loss = 0.4 * ce(out, y) + 0.6 * dice(out, y, softmax=True)   # trainer.py:55-57
        opt.zero_grad()
        loss.backward()
        opt.step()
        lr_ = args.lr * (1.0 - it / args.steps) ** 0.9              # trainer.py:61-63
        for g in opt.param_groups:
            g["lr"] = lr_
        if it % 20 == 0 or it == args.steps - 1:
            acc = (out.argmax(1) == y).float().mean().item()
            print(f"  step {it:4d} loss {loss.item():.4f} pixel acc {acc:.4f}  ({time.time() - t0:.0f} s)", flush=True)

    # ---------------- golden outputs of the UNMODIFIED reference on the trained weights ----------------
    m.eval()
    out = {}
    for k in trained:
        out["w." + k] = m.state_dict()[k].detach().numpy().astype(np.float32)
    n_test = 4 if args.config == "t224" else 2
    xs, ys = synth.synth_seg_batch(n_test, S, NC, seed=10_000)
    x = torch.from_numpy(xs).repeat(1, 3, 1, 1)
    with torch.no_grad():
        logits = m(x)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            logits_bf16 = m(x).float()
    top2 = logits.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])
    out["argmax"] = logits.argmax(1).numpy().astype(np.uint8)
    out["margin_q"] = np.quantile(margin.numpy().ravel(), [0.0001, 0.001, 0.01, 0.1, 0.5]).astype(np.float32)
    out["pixel_acc_vs_labels"] = np.float32((logits.argmax(1).numpy() == ys).mean())
    out["ref_bf16_autocast_maxabs"] = np.float32((logits_bf16 - logits).abs().max().item())
    out["ref_bf16_autocast_agree"] = np.float32((logits_bf16.argmax(1) == logits.argmax(1)).float().mean().item())
    out["logit_absmax"] = np.float32(logits.abs().max().item())
    # logits in NHWC row form (rows = pixels, width = classes) so that golden_util.compare works; every 3rd pixel stored
    lg = logits.permute(0, 2, 3, 1).reshape(-1, NC)
    for k, v in pack(lg, 3).items():
        out["logits." + k] = v
    print(f"[{args.config}] held-out: pixel acc vs labels {out['pixel_acc_vs_labels']:.4f}, |logit| max {out['logit_absmax']:.2f}, "
          f"margin quantiles (1e-4, 1e-3, 1e-2, .1, .5) {out['margin_q']}, reference bf16 autocast: max-abs "
          f"{out['ref_bf16_autocast_maxabs']:.3e}, agreement {out['ref_bf16_autocast_agree']:.5f}")

    # volume through the reference's test_single_volume loop (utils.py:61-80): zoom order 3 in, argmax(softmax), zoom order 0 out
    from scipy.ndimage import zoom
    D, VS = cfg["vol"]
    vol, gt = synth.synth_seg_volume(D, VS, NC, seed=77)
    pred = np.zeros((D, VS, VS), np.uint8)
    with torch.no_grad():
        for d in range(D):
            sl = vol[d] if VS == S else zoom(vol[d], (S / VS, S / VS), order=3)
            inp = torch.from_numpy(np.ascontiguousarray(sl))[None, None].float().repeat(1, 3, 1, 1)
            o = torch.argmax(torch.softmax(m(inp), dim=1), dim=1)[0].numpy()
            pred[d] = o if VS == S else zoom(o, (VS / S, VS / S), order=0)
    out["vol_pred"] = pred
    mets = np.array([O.dice_hd95_percase(pred == c, gt == c) for c in range(1, NC)], np.float64)
    out["vol_metrics"] = mets
    # yard-stick: the reference's OWN bf16 autocast through the same loop, and how far its metrics move
    pred16 = np.zeros((D, VS, VS), np.uint8)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        for d in range(D):
            sl = vol[d] if VS == S else zoom(vol[d], (S / VS, S / VS), order=3)
            inp = torch.from_numpy(np.ascontiguousarray(sl))[None, None].float().repeat(1, 3, 1, 1)
            o = torch.argmax(torch.softmax(m(inp).float(), dim=1), dim=1)[0].numpy()
            pred16[d] = o if VS == S else zoom(o, (VS / S, VS / S), order=0)
    mets16 = np.array([O.dice_hd95_percase(pred16 == c, gt == c) for c in range(1, NC)], np.float64)
    out["ref_bf16_vol_agree"] = np.float64((pred16 == pred).mean())
    out["ref_bf16_vol_ddice"] = np.float64(np.abs(mets16[:, 0] - mets[:, 0]).max())
    out["ref_bf16_vol_dhd95"] = np.float64(np.abs(mets16[:, 1] - mets[:, 1]).max())
    print(f"[{args.config}] reference bf16 autocast through the same loop: label agreement {out['ref_bf16_vol_agree']:.6f}, worst |dDice| "
          f"{out['ref_bf16_vol_ddice']:.2e}, worst |dHD95| {out['ref_bf16_vol_dhd95']:.2e}")
    print(f"[{args.config}] volume {D}x{VS}^2: per-class (Dice, HD95) of the reference vs ground truth:\n{np.round(mets, 4)}")
    out["meta"] = np.array([args.steps, B, S, NC, D, VS], np.int64)
    path = os.path.join(HERE, f"trained_{args.config}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
