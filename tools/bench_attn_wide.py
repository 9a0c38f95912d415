#!/usr/bin/env python
"""512^2 configuration (split [1,2,8,8]): fused LePE attention forward per stage, bf16, batch B — CUDA-graph timing.
Run twice (CSWIN_ATTN_FWD_SIMT=1 for the general SIMT kernel) to compare.  Usage: python tools/bench_attn_wide.py [B]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
DEV = "cuda"
for (C, reso, heads, split, last) in ((64, 128, 2, 1, False), (128, 64, 4, 2, False), (256, 32, 8, 8, False), (512, 16, 16, 8, True)):
    blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).to(DEV).eval()
    L = reso * reso
    qkv = torch.randn(B, L, 3 * C, device=DEV, dtype=torch.bfloat16)
    out = torch.empty(B, L, C, device=DEV, dtype=torch.bfloat16)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    if blk.branch_num == 2:
        h = C // 2
        descs = [a.branch_desc(q[..., i * h:(i + 1) * h], k[..., i * h:(i + 1) * h], v[..., i * h:(i + 1) * h], out[..., i * h:(i + 1) * h]) for i, a in enumerate(blk.attns)]
    else:
        descs = [blk.attns[0].branch_desc(q, k, v, out)]
    fn = lambda: ops.lepe_attention_fwd(descs, B, reso, float(blk.attns[0].scale), torch.bfloat16)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        t0 = cw.tc_launch_count()
        for _ in range(3): fn()
        tc = cw.tc_launch_count() > t0
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10): fn()
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(20): g.replay()
        e1.record(s); s.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    N = blk.attns[0].H_sp * blk.attns[0].W_sp
    byts = 4 * B * L * C * 2
    print(f"512^2 stage C={C:3d} reso={reso:3d} window N={N:3d} B={B}: {us:8.2f} us  {'tcgen05' if tc else 'SIMT   '}  {byts / us / 1e3:7.1f} GB/s algorithmic")
