import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import _lib
stage, B = int(sys.argv[1]), int(sys.argv[2])
C, reso, heads, split, last = [(64, 56, 2, 1, False), (128, 28, 4, 2, False), (256, 14, 8, 7, False), (512, 7, 16, 7, True)][stage - 1]
blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).cuda().train()
x = torch.randn(B, reso * reso, C, device="cuda", dtype=torch.bfloat16, requires_grad=True)
for _ in range(3):
    y = blk(x); y.sum().backward()
torch.cuda.synchronize()
buf = torch.zeros(1024 * 16, dtype=torch.int64, device="cuda")
y = blk(x)
_lib.lib().cswin_debug_set_trace(buf.data_ptr())
y.sum().backward(); torch.cuda.synchronize()
_lib.lib().cswin_debug_set_trace(None)
t = buf.cpu().numpy().reshape(1024, 16)[:, :8].astype(np.float64)
# the buffer is shared by all traced kernels of the backward; the LAST writer per CTA slot wins: attention bwd runs after
# fc/proj dgrad... so report whatever is there for rows with monotone stamps
live = (t[:, 0] > 0) & (t[:, 7] >= t[:, 0])
t = t[live]; t0 = t[:, 0].min()
names = ["entry", "prologue", "S/dP ready", "P,dS published", "lepe dv done", "dw/db done", "dQ/dK/dV ready", "exit"]
print(f"stage {stage} B={B}: {live.sum()} CTA slots")
for i, n in enumerate(names):
    d = np.median(t[:, i + 1] - t[:, i]) / 1e3 if i < 7 else 0.0
    print(f"  {n:18s} start {np.median(t[:, i] - t0)/1e3:8.2f} us   dur {d:7.2f}")
print(f"lifetime median {np.median(t[:,7]-t[:,0])/1e3:.2f} us")
