#!/usr/bin/env python
"""Persistent stage kernel (csrc/stage_tc.cu) vs the composed path (5 launches per block), CUDA-graph replay, CUDA events.
usage: bench_stage.py [B] [stage ...]   (stages 1-4 of cswin_tiny_224; default 3)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import modules, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
stages = [int(s) for s in sys.argv[2:]] or [3]
SHAPES = {1: (64, 56, 2, 1, False, 1), 2: (128, 28, 4, 2, False, 2), 3: (256, 14, 8, 7, False, 9), 4: (512, 7, 16, 7, True, 1)}
DEV = "cuda"


def timed(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        fn(); fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, out


for st in stages:
    C, reso, heads, split, last, n = SHAPES[st]
    blocks = []
    for i in range(n):
        blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).eval()
        sd = {k: torch.from_numpy(synth.synth_tensor(f"bs/{C}/{i}/" + k, tuple(v.shape), 31)) for k, v in blk.state_dict().items()}
        blk.load_state_dict(sd, strict=True)
        blocks.append(blk.to(DEV))
    x = torch.from_numpy(synth.synth_tensor(f"bs_in/{C}", (B, reso * reso, C), 32)).bfloat16().to(DEV)
    res = {}
    for name, dims in (("composed", ()), ("stage", (C,))):
        modules.STAGE_EXEC_DIMS = dims
        us, out = timed(lambda: modules.run_stage(blocks, x.clone()))
        res[name] = (us, out.float())
    d = (res["composed"][1] - res["stage"][1]).abs().max().item()
    flop = 2.0 * B * reso * reso * C * C * 12 * n
    print(f"stage {st} (C={C}, {n} blocks, B={B}): composed {res['composed'][0]:8.1f} us, stage kernel {res['stage'][0]:8.1f} us "
          f"({res['composed'][0] / res['stage'][0]:.2f}x, {flop / res['stage'][0] * 1e-6:.0f} TFLOP/s Linear), max-abs diff {d:.2e}")
