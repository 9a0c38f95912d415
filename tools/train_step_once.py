#!/usr/bin/env python
"""Runs a few train steps (batch B, bf16 compute) — target for `ncu --metrics gpu__time_duration.sum`."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m = cw.cswin_tiny_224(num_classes=9).train()
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
m = m.cuda()
step = cw.TrainStep(m, lr=0.05, graph=(os.environ.get('NOGRAPH') is None))
x = torch.from_numpy(synth.synth_image_batch(B, 3, 224, seed=0, kind="ct")).cuda()
y = torch.from_numpy(synth.synth_labels(B, 224, 9, seed=0)).cuda()
for i in range(n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loss = step(x, y)
    torch.cuda.synchronize()
    print(f"step {i}: {1e3*(time.perf_counter()-t0):.1f} ms loss {float(loss):.4f}")
