#!/usr/bin/env python
"""Fused stem kernel (csrc/stem_tc.cu) vs im2col + Linear + LayerNorm, CUDA-graph replay.  usage: bench_stem.py [B] [S]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import model as mm, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
S = int(sys.argv[2]) if len(sys.argv) > 2 else 224
m = cw.cswin_tiny_224(num_classes=9).eval().cuda()
x = torch.from_numpy(synth.synth_image_batch(B, 3, S, seed=0, kind="ct")).cuda()
for fused in (True, False):
    mm.FUSE_STEM = fused
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        m._stem(x, torch.bfloat16); m._stem(x, torch.bfloat16); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10): y = m._stem(x, torch.bfloat16)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"stem B={B} S={S} fused={fused}: {e0.elapsed_time(e1) / 100 * 1e3:.2f} us per call")
