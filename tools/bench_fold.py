"""Per-stage CSWinBlock forward time (bf16, batch 24, CUDA graph of 8 chained blocks) with the LayerNorm fold on / off."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cswin_unet_b200 as cw
from cswin_unet_b200 import modules as M_

DEV = "cuda"
B = int(os.environ.get("B", 24))
for (dim, reso, heads, split, last) in ((64, 56, 2, 1, False), (128, 28, 4, 2, False), (256, 14, 8, 7, False), (512, 7, 16, 7, True)):
    blk = cw.CSWinBlock(dim=dim, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).to(DEV).eval()
    x = torch.randn(B, reso * reso, dim, device=DEV).bfloat16()
    res = {}
    for fold in (0, 1, 2, 0, 1, 2):
        M_.FOLD_LN = fold > 0
        M_.FUSE_MLP = fold > 1
        M_.FUSE_MLP_MAX_DIM = 512
        s = torch.cuda.Stream()
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(3):
                y = x
                for _ in range(8):
                    y = blk(y)
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                y = x
                for _ in range(8):
                    y = blk(y)
            for _ in range(5):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(50):
                g.replay()
            e1.record(s)
            s.synchronize()
        res.setdefault(fold, []).append(e0.elapsed_time(e1) / 50 / 8 * 1e3)
    print(f"dim {dim:4d} reso {reso:3d}: unfused {min(res[0]):7.2f} us/block   folded LN {min(res[1]):7.2f}   folded LN + fused MLP {min(res[2]):7.2f}")
