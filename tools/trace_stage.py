#!/usr/bin/env python
"""Per-CTA cycle accounting of the persistent stage kernel (needs a library built with `make EXTRA=-DCSWIN_STAGE_PROFILE`).
usage: trace_stage.py [B] [stage] [n_blocks]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import modules, synth, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
st = int(sys.argv[2]) if len(sys.argv) > 2 else 3
SHAPES = {1: (64, 56, 2, 1, False, 1), 2: (128, 28, 4, 2, False, 2), 3: (256, 14, 8, 7, False, 9), 4: (512, 7, 16, 7, True, 1)}
C, reso, heads, split, last, n = SHAPES[st]
if len(sys.argv) > 3: n = int(sys.argv[3])
DEV = "cuda"
blocks = []
for i in range(n):
    blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).eval()
    blk.load_state_dict({k: torch.from_numpy(synth.synth_tensor(f"bs/{C}/{i}/" + k, tuple(v.shape), 31)) for k, v in blk.state_dict().items()})
    blocks.append(blk.to(DEV))
x = torch.from_numpy(synth.synth_tensor(f"bs_in/{C}", (B, reso * reso, C), 32)).bfloat16().to(DEV)
modules.STAGE_EXEC_DIMS = (C,)
with torch.no_grad():
    for _ in range(3): modules.run_stage(blocks, x.clone())
    torch.cuda.synchronize()
    buf = torch.zeros(1024 * 16, dtype=torch.int64, device=DEV)
    xin = x.clone()
    _lib.lib().cswin_debug_set_trace(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); modules.run_stage(blocks, xin); e1.record(); torch.cuda.synchronize()
    _lib.lib().cswin_debug_set_trace(None)
t = buf.cpu().numpy().reshape(1024, 16)
t = t[t[:, 0] > 0]
us = lambda c: c / 1965.0
print(f"stage {st} B={B} blocks={n}: {e0.elapsed_time(e1) * 1e3:.1f} us (events, eager incl. row_stats), {len(t)} CTAs traced; cycles -> us at 1965 MHz; median [min..max] per CTA")
def line(name, v):
    print(f"  {name:34s} {us(np.median(v)):8.1f} [{us(v.min()):8.1f} .. {us(v.max()):8.1f}]")
line("producer: total", t[:, 0]); line("producer: wait dependency flags", t[:, 1]); line("producer: wait ring slot empty", t[:, 2]); line("producer: wait tile queue + atomic", t[:, 3])
line("mma: total", t[:, 4]); line("mma: wait operands (full)", t[:, 5]); line("mma: wait accumulator free", t[:, 6]); line("mma: wait P (softmax)", t[:, 7]); line("mma: wait tile queue", t[:, 14])
line("epilogue: total", t[:, 8]); line("epilogue: wait tile queue", t[:, 9]); line("epilogue: wait accumulator", t[:, 10]); line("epilogue: own dependency poll", t[:, 11])
line("epilogue: store complete + barrier", t[:, 12]); line("epilogue: fence + release", t[:, 13] & 0xffffffff); line("epilogue: attention tiles total", t[:, 13] >> 32)
print(f"  tiles per CTA: median {np.median(t[:, 15]):.0f} [{t[:, 15].min()} .. {t[:, 15].max()}]")
