#!/usr/bin/env python
"""Second target for `ncu --profile-from-start off` (see ncu_all_target.py): the kernels an eager train step does not launch —
the bf16 EVAL forward (folded-LayerNorm Linear variants, fused MLP, folded head with arg-max), the resampled-volume path (spline
prefilter / zoom kernels), the 512^2 configuration's wide attention kernels (forward + backward), the fused qkv+attention kernel
and the persistent stage kernel.  Usage: ncu_rest_target.py [B]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import modules, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = torch.device("cuda", 0)
me = cw.cswin_tiny_224(num_classes=9)
shapes = {k: tuple(v.shape) for k, v in me.state_dict().items()}
me.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
me = me.to(dev).eval()
me.compute_dtype = torch.bfloat16
x = torch.from_numpy(synth.synth_image_batch(B, 3, 224, seed=0, kind="ct")).to(dev)
eng = cw.SliceEngine(me, batch=8, compute_dtype=torch.bfloat16, inflight=1)
vol = np.random.default_rng(0).random((8, 512, 512), dtype=np.float32)
blk = cw.CSWinBlock(dim=256, reso=32, num_heads=8, split_size=8, qkv_bias=True).to(dev).train()      # 512^2 stage 3: 256-token windows
xw = torch.randn(2, 32 * 32, 256, device=dev, dtype=torch.bfloat16, requires_grad=True)
blocks3 = [cw.CSWinBlock(dim=256, reso=14, num_heads=8, split_size=7, qkv_bias=True).to(dev).eval() for _ in range(2)]
x3 = torch.randn(B, 196, 256, device=dev, dtype=torch.bfloat16)


def everything():
    with torch.no_grad():
        me.predict_labels(x)
        modules.STAGE_EXEC_DIMS = (256,)
        modules.run_stage(blocks3, x3.clone())
        modules.STAGE_EXEC_DIMS = ()
        modules.FUSE_QKV_ATTN = True
        blocks3[0](x3)
        modules.FUSE_QKV_ATTN = False
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
