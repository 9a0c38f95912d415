#!/usr/bin/env python
"""Tiny driver for `ncu --set full`: runs a few launches of one op at one shape.
usage: ncu_target.py attn <stage 1-4> <B> | attn512 <stage 1-4> <B> | block_infer <stage> <B> | block_train <stage> <B> | conv <B> <H> <C> <N> <stride> | linear <M> <N> <K> [act] [res]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import ops

DEV = "cuda"
kind = sys.argv[1]
if kind in ("attn", "attn512"):
    stage, B = int(sys.argv[2]), int(sys.argv[3])
    T224 = [(64, 56, 2, 1, False), (128, 28, 4, 2, False), (256, 14, 8, 7, False), (512, 7, 16, 7, True)]
    T512 = [(64, 128, 2, 1, False), (128, 64, 4, 2, False), (256, 32, 8, 8, False), (512, 16, 16, 8, True)]      # split [1,2,8,8]
    C, reso, heads, split, last = (T224 if kind == "attn" else T512)[stage - 1]
    blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).to(DEV).eval()
    L = reso * reso
    qkv = torch.randn(B, L, 3 * C, device=DEV, dtype=torch.bfloat16)
    out = torch.empty(B, L, C, device=DEV, dtype=torch.bfloat16)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    if blk.branch_num == 2:
        h = C // 2
        descs = [a.branch_desc(q[..., i * h:(i + 1) * h], k[..., i * h:(i + 1) * h], v[..., i * h:(i + 1) * h], out[..., i * h:(i + 1) * h]) for i, a in enumerate(blk.attns)]
    else:
        descs = [blk.attns[0].branch_desc(q, k, v, out)]
    fn = lambda: ops.lepe_attention_fwd(descs, B, reso, float(blk.attns[0].scale), torch.bfloat16)
elif kind == "block_infer":          # inference forward of one CSWinBlock (bf16): folded-LN Linears, attention, (fused) MLP
    stage, B = int(sys.argv[2]), int(sys.argv[3])
    C, reso, heads, split, last = [(64, 56, 2, 1, False), (128, 28, 4, 2, False), (256, 14, 8, 7, False), (512, 7, 16, 7, True)][stage - 1]
    blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).to(DEV).eval()
    x = torch.randn(B, reso * reso, C, device=DEV, dtype=torch.bfloat16)
    def fn():
        with torch.no_grad():
            blk(blk(x))
elif kind == "block_train":          # forward + backward of one CSWinBlock (bf16): attention fwd/bwd, Linear fwd/dgrad/wgrad ...
    stage, B = int(sys.argv[2]), int(sys.argv[3])
    C, reso, heads, split, last = [(64, 56, 2, 1, False), (128, 28, 4, 2, False), (256, 14, 8, 7, False), (512, 7, 16, 7, True)][stage - 1]
    blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).to(DEV).train()
    x = torch.randn(B, reso * reso, C, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    def fn():
        y = blk(x); y.sum().backward()
elif kind == "conv":                 # implicit-GEMM conv over a token image (Merge_Block: stride 2; CARAFE.encoder: stride 1): conv B H C N stride
    B, H, C, N, stride = [int(v) for v in sys.argv[2:7]]
    x = torch.randn(B, H * H, C, device=DEV, dtype=torch.bfloat16)
    w = torch.randn(N, 9 * C, device=DEV, dtype=torch.bfloat16) / (9 * C) ** 0.5
    bias = torch.randn(N, device=DEV, dtype=torch.bfloat16)
    fn = lambda: ops.conv_tokens(x, H, H, w, bias, 3, 3, stride, 1)
else:
    M, N, K = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    act = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    res = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=DEV, dtype=torch.bfloat16) / K ** 0.5
    bias = torch.randn(N, device=DEV, dtype=torch.bfloat16)
    r = torch.randn(M, N, device=DEV, dtype=torch.bfloat16) if res else None
    o = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    fn = lambda: ops.linear(a, w, bias, act=act, residual=r, out=o)
for _ in range(6):
    fn()
torch.cuda.synchronize()
print("ok")
