#!/usr/bin/env python
"""Per-TILE timeline of the persistent stage kernel (library built with EXTRA=-DCSWIN_STAGE_PROFILE; CSWIN_STAGE_TILE_TRACE=1).
usage: trace_stage_tiles.py [B] [n_blocks]"""
import os, sys
os.environ["CSWIN_STAGE_TILE_TRACE"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import modules, synth, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
C, reso, heads, split, last = 256, 14, 8, 7, False
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
    buf = torch.zeros(200000 * 10, dtype=torch.int64, device=DEV)
    xin = x.clone()
    _lib.lib().cswin_debug_set_trace(buf.data_ptr())
    modules.run_stage(blocks, xin); torch.cuda.synchronize()
    _lib.lib().cswin_debug_set_trace(None)
t = buf.cpu().numpy().reshape(-1, 10)
nt = int((t[:, 0] > 0).sum())
t = t[:nt].astype(np.int64)
t0 = t[:, 0].min()
info = t[:, 9]
j, op, m, nn = info & 0xff, (info >> 8) & 0xf, (info >> 12) & 0xfffff, info >> 32
S = (t[:, :8] - t0) / 1e3                       # us
names = ["qkv", "att", "proj", "fc1", "fc2"]
print(f"{nt} tiles, {S[:, 7].max():.1f} us first grab -> last release; stamps: 0 grabbed, 1 deps ok, 2 loads issued, 3 first operands landed, 4 MMAs issued, 5 acc ready seen, 6 outputs complete, 7 released")
print("median per op [us]:   deps wait (1-0)   load lat (3-1)   mainloop (4-3)   epilogue (6-5)   release (7-6)   grab->release (7-0)   active (7-1)")
for o in range(5):
    k = op == o
    f = lambda a, b: np.median(S[k, a] - S[k, b])
    print(f"  {names[o]:5s} n={k.sum():5d}     {f(1,0):8.2f}         {f(3,1):8.2f}         {f(4,3):8.2f}         {f(6,5):8.2f}        {f(7,6):8.2f}        {f(7,0):8.2f}          {f(7,1):8.2f}")
# signalling latency: deps-ok time of a GEMM tile minus the release time of the LAST tile it depends on (same block, previous op, same row tile)
idx = {}
for i in range(nt):
    idx.setdefault((int(j[i]), int(op[i]), int(m[i])), []).append(i)
for o, po in ((3, 2), (4, 3)):
    lat = []
    for i in np.nonzero(op == o)[0]:
        d = idx.get((int(j[i]), po, int(m[i])))
        if d: lat.append(S[i, 1] - max(S[k2, 7] for k2 in d))
    lat = np.array(lat)
    print(f"  signal latency {names[po]} release -> {names[o]} deps-ok: median {np.median(lat):.2f} us, p90 {np.percentile(lat, 90):.2f}, share of tiles that were already waiting (lat < 3 us): {(lat < 3).mean():.2f}")
# per block wall time
for jj in range(n):
    k = j == jj
    print(f"  block {jj}: first grab {S[k, 0].min():8.1f}  last release {S[k, 7].max():8.1f}; per op first deps-ok / last release: " +
          "  ".join(f"{names[o]} {S[k & (op == o), 1].min():.1f}/{S[k & (op == o), 7].max():.1f}" for o in range(5)))
# timeline of row tile 5 of block 1
jj = min(1, n - 1)
print(f"  chain of row tile 5 in block {jj} (grab, deps ok, landed, mma, acc seen, complete, released; CTA):")
for o in range(5):
    for i in idx.get((jj, o, 5 if o != 1 else 3), [])[:3]:
        print(f"    {names[o]:5s} n={int(nn[i]):3d}  " + " ".join(f"{v:8.2f}" for v in S[i, [0, 1, 3, 4, 5, 6, 7]]) + f"   cta {int(t[i, 8])}")
