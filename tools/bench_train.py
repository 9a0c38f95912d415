#!/usr/bin/env python
"""Train step (batch B, bf16 compute, CUDA graph) timing for A/B runs of env knobs / library variants.  usage: bench_train.py [B] [steps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
m = cw.cswin_tiny_224(num_classes=9).train()
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
m = m.cuda()
step = cw.TrainStep(m, lr=0.05)
x = torch.from_numpy(synth.synth_image_batch(B, 3, 224, seed=0, kind="ct")).cuda()
y = torch.from_numpy(synth.synth_labels(B, 224, 9, seed=0)).cuda()
for _ in range(6): loss = step(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): loss = step(x, y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"train step B={B}: {ms:.3f} ms, {B / ms * 1e3:.0f} slices/s, loss {float(loss):.4f}  [{os.environ.get('CSWIN_LIB_PATH', 'default lib')}, side={os.environ.get('CSWIN_WGRAD_SIDE_STREAM', '0')}, cap={os.environ.get('CSWIN_GEMM_SMEM_CAP_KB', '-')}]")
step.close()
