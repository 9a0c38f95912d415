#!/usr/bin/env python
"""Per-stage device time of the fused LePE attention forward / backward kernels (torch.profiler / CUPTI), plus the
in-kernel %globaltimer phase trace of the backward kernel.  Usage: python tools/bench_attn_bwd.py [B] [512]"""
import collections
import os
import sys

import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw  # noqa: E402
from cswin_unet_b200 import _lib, autograd as ag  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
STAGES = [(64, 56, 2, 1, False), (128, 28, 4, 2, False), (256, 14, 8, 7, False), (512, 7, 16, 7, True)]
if len(sys.argv) > 2 and sys.argv[2] == "512":          # 512^2 configuration, split [1,2,8,8]
    STAGES = [(64, 128, 2, 1, False), (128, 64, 4, 2, False), (256, 32, 8, 8, False), (512, 16, 16, 8, True)]
NAMES = ["entry", "prologue", "S/dP ready", "P,dS published", "lepe dv done", "(unused)", "dQ/dK/dV ready", "exit"]

for si, (C, reso, heads, split, last) in enumerate(STAGES):
    blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).cuda().train()
    L = reso * reso
    qkv = torch.randn(B, L, 3 * C, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    g = torch.randn(B, L, C, device="cuda", dtype=torch.bfloat16)

    # call the autograd Function the block uses
    def fb():
        qkv.grad = None
        a = blk.attns
        nb = len(a)
        meta = dict(reso=reso, scale=float(a[0].scale), heads=[m.num_heads for m in a], win=[(m.H_sp, m.W_sp) for m in a])
        cw0, cb0 = a[0].get_v.weight, a[0].get_v.bias
        cw1, cb1 = (a[1].get_v.weight, a[1].get_v.bias) if nb == 2 else (None, None)
        y = ag.LepeAttentionFn.apply(qkv, cw0, cb0, cw1, cb1, meta)
        y.backward(g)

    for _ in range(3):
        fb()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            fb()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and "lepe_" in e.name:
            n = "param_grad" if "param_grad" in e.name else "bwd" if "bwd" in e.name else "fwd"
            agg[n][0] += 1
            agg[n][1] += e.device_time
    msg = ", ".join(f"{n} {t / c:7.2f} us" for n, (c, t) in sorted(agg.items()))
    print(f"stage {si + 1} B={B} C={C} reso={reso}: {msg}")
    buf = torch.zeros(1024 * 16, dtype=torch.int64, device="cuda")
    a = blk.attns
    meta = dict(reso=reso, scale=float(a[0].scale), heads=[m.num_heads for m in a], win=[(m.H_sp, m.W_sp) for m in a])
    y = ag.LepeAttentionFn.apply(qkv, a[0].get_v.weight, a[0].get_v.bias, *((a[1].get_v.weight, a[1].get_v.bias) if len(a) == 2 else (None, None)), meta)
    torch.cuda.synchronize()
    _lib.lib().cswin_debug_set_trace(buf.data_ptr())
    y.backward(g)
    torch.cuda.synchronize()
    _lib.lib().cswin_debug_set_trace(None)
    t = buf.cpu().numpy().reshape(1024, 16)[:, :8].astype(np.float64)
    live = (t[:, 0] > 0) & (t[:, 7] >= t[:, 0])
    t = t[live]
    if len(t):
        t0 = t[:, 0].min()
        print("   " + " | ".join(f"{n} +{np.median(t[:, i] - t0) / 1e3:.2f}" for i, n in enumerate(NAMES)))
        print(f"   CTA lifetime median {np.median(t[:, 7] - t[:, 0]) / 1e3:.2f} us, first entry -> last exit {(t[:, 7].max() - t0) / 1e3:.2f} us ({len(t)} CTAs traced)")
