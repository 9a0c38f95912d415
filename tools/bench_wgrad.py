"""Time cswin_linear_wgrad (tcgen05) per layer shape of the T224 / batch-24 train step: CUDA graph of 10 calls, 30 replays."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cswin_unet_b200 import ops
DEV = "cuda"
shapes = [(75264, 192, 64, 2), (75264, 64, 64, 2), (75264, 256, 64, 2), (75264, 64, 256, 2),
          (18816, 384, 128, 4), (18816, 128, 128, 4), (18816, 512, 128, 4), (18816, 128, 512, 4),
          (4704, 768, 256, 18), (4704, 256, 256, 18), (4704, 1024, 256, 18), (4704, 256, 1024, 18),
          (1176, 1536, 512, 2), (1176, 512, 512, 2), (1176, 2048, 512, 2), (1176, 512, 2048, 2)]
tot = 0.0
for M, N, K, cnt in shapes:
    dz = torch.randn(M, N, device=DEV).bfloat16(); a = torch.randn(M, K, device=DEV).bfloat16()
    dw = torch.zeros(N, K, device=DEV); db = torch.zeros(N, device=DEV)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): ops.linear_wgrad(dz, a, dw, db)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10): ops.linear_wgrad(dz, a, dw, db)
        for _ in range(3): g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(30): g.replay()
        e1.record(s); s.synchronize()
    us = e0.elapsed_time(e1) / 300 * 1e3
    tot += us * cnt
    print(f"M={M:6d} N={N:5d} K={K:5d}: {us:7.2f} us  x{cnt:3d} = {us*cnt:8.1f}   ({2*M*N*K/us/1e6:7.1f} TFLOP/s)")
print(f"sum over block Linears: {tot:.0f} us")
