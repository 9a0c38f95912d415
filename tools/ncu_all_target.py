#!/usr/bin/env python
"""Target for `ncu --profile-from-start off`: after warm-up, ONE pass over every kernel family of the library inside a
cudaProfilerStart / Stop bracket — an eager bf16 train step at batch B (forward, backward, loss, SGD), the bf16 eval forward
(folded-LayerNorm Linears, fused MLP, folded head with arg-max), a resampled volume (spline prefilter / zoom kernels) and the
512^2 configuration's wide attention kernels (forward + backward).  Usage: ncu_all_target.py [B]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = torch.device("cuda", 0)


def model(nc=9, img=224, split=(1, 2, 7, 7)):
    m = cw.cswin_tiny_224(num_classes=nc, img_size=img, split_size=list(split))
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
    return m.to(dev)


mt = model().train()
step = cw.TrainStep(mt, lr=0.05, graph=False)
x = torch.from_numpy(synth.synth_image_batch(B, 3, 224, seed=0, kind="ct")).to(dev)
y = torch.from_numpy(synth.synth_labels(B, 224, 9, seed=0)).to(dev)
me = model().eval()
me.compute_dtype = torch.bfloat16
eng = cw.SliceEngine(me, batch=B, compute_dtype=torch.bfloat16)
vol = np.random.default_rng(0).random((B, 512, 512), dtype=np.float32)
blk = cw.CSWinBlock(dim=256, reso=32, num_heads=8, split_size=8, qkv_bias=True).to(dev).train()      # 512^2 stage 3: 256-token windows
xw = torch.randn(2, 32 * 32, 256, device=dev, dtype=torch.bfloat16, requires_grad=True)


def everything():
    step(x, y)
    with torch.no_grad():
        me.predict_labels(x)
    cw.predict_volume(eng, vol, resample="gpu")
    blk(xw).sum().backward()


for _ in range(3):
    everything()
torch.cuda.synchronize()
torch.cuda.profiler.start()
everything()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
