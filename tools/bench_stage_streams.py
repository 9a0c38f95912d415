#!/usr/bin/env python
"""Experiment: the stage kernel as K independent chains (K sub-batches on K streams, grid/K CTAs each) vs one chain.
usage: bench_stage_streams.py [B] [K]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
os.environ["CSWIN_STAGE_CTAS"] = str(296 // K)
import cswin_unet_b200 as cw
from cswin_unet_b200 import modules, synth
C, reso, heads, split, last, n = 256, 14, 8, 7, False, 9
DEV = "cuda"
blocks = []
for i in range(n):
    blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).eval()
    blk.load_state_dict({k: torch.from_numpy(synth.synth_tensor(f"bs/{C}/{i}/" + k, tuple(v.shape), 31)) for k, v in blk.state_dict().items()})
    blocks.append(blk.to(DEV))
x = torch.from_numpy(synth.synth_tensor(f"bs_in/{C}", (B, reso * reso, C), 32)).bfloat16().to(DEV)
modules.STAGE_EXEC_DIMS = (C,)
side = [torch.cuda.Stream() for _ in range(K)]
def fn():
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(cur)
    outs = []
    for i, s in enumerate(side):
        s.wait_event(ev)
        with torch.cuda.stream(s):
            outs.append(modules.run_stage(blocks, x[B * i // K: B * (i + 1) // K].clone()))
    for s in side: cur.wait_stream(s)
    return outs
s0 = torch.cuda.Stream(); s0.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s0), torch.no_grad():
    fn(); fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s0):
        out = fn()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): g.replay()
e1.record(); torch.cuda.synchronize()
print(f"stage 3, B={B}, {K} chains x {296 // K} CTAs: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per 9 blocks")
