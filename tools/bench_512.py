#!/usr/bin/env python
"""BASELINE configs[4]: cswin_tiny at 512^2 (3 classes, split [1,2,8,8]) — bf16 forward (CUDA-graph replay) and train step
(TrainStep, graph) on one GPU, batch B.  Usage: python tools/bench_512.py [B]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
m = cw.cswin_tiny_224(num_classes=3, img_size=512, split_size=[1, 2, 8, 8]).eval()
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
m = m.to(dev); m.compute_dtype = torch.bfloat16
x = torch.from_numpy(synth.synth_image_batch(B, 3, 512, seed=0, kind="ct")).to(dev)
y = torch.from_numpy(synth.synth_labels(B, 512, 3, seed=0)).to(dev)
with torch.no_grad():
    t0, n0 = cw.tc_launch_count(), cw.launch_count()
    m(x)
    print(f"forward: {cw.launch_count() - n0} native launches, {cw.tc_launch_count() - t0} on tcgen05")
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        m(x); m(x)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = m(x)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
print(f"512^2 forward  B={B}: {ms:.3f} ms per batch, {B / ms * 1e3:.0f} slices/s, {B / ms * 56.60:.1f} TFLOP/s (56.60 GFLOP per slice)")
import copy
step = cw.TrainStep(copy.deepcopy(m).train(), lr=0.05)
for _ in range(5): step(x, y)
torch.cuda.synchronize()
e0.record()
for _ in range(10): step(x, y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"512^2 train step B={B}: {ms:.3f} ms per step, {B / ms * 1e3:.0f} slices/s, {B / ms * 186.23:.1f} TFLOP/s (186.23 GFLOP per slice)")
