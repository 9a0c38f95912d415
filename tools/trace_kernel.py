#!/usr/bin/env python
"""In-kernel phase timeline (%globaltimer stamps, ns) of the tcgen05 kernels for one op/shape.
usage: trace_kernel.py attn <stage> <B> | linear <M> <N> <K> [act] [res]"""
import os, sys, subprocess
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import ops, _lib
DEV = "cuda"
kind = sys.argv[1]
if kind == "attn":
    stage, B = int(sys.argv[2]), int(sys.argv[3])
    C, reso, heads, split, last = [(64, 56, 2, 1, False), (128, 28, 4, 2, False), (256, 14, 8, 7, False), (512, 7, 16, 7, True)][stage - 1]
    blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).to(DEV).eval()
    L = reso * reso
    qkv = torch.randn(B, L, 3 * C, device=DEV, dtype=torch.bfloat16); out = torch.empty(B, L, C, device=DEV, dtype=torch.bfloat16)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    h = C // 2
    descs = ([a.branch_desc(q[..., i * h:(i + 1) * h], k[..., i * h:(i + 1) * h], v[..., i * h:(i + 1) * h], out[..., i * h:(i + 1) * h]) for i, a in enumerate(blk.attns)]
             if blk.branch_num == 2 else [blk.attns[0].branch_desc(q, k, v, out)])
    fn = lambda: ops.lepe_attention_fwd(descs, B, reso, float(blk.attns[0].scale), torch.bfloat16)
    names = ["entry", "prologue", "qkv_landed", "S_ready", "P_published", "lepe_done", "O_ready", "exit"]
elif kind == "qa":
    stage, B = int(sys.argv[2]), int(sys.argv[3])
    C, reso, heads, split, last = [(64, 56, 2, 1, False), (128, 28, 4, 2, False), (256, 14, 8, 7, False)][stage - 1]
    blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).to(DEV).eval()
    x = torch.randn(B, reso * reso, C, device=DEV, dtype=torch.bfloat16)
    st = ops.row_stats(x)
    wq, csq, bq = blk._folded("qkv", blk.qkv, blk.norm1)
    brs = [dict(conv_w=a.get_v.weight.detach().bfloat16(), conv_b=a.get_v.bias.detach().bfloat16(), heads=a.num_heads, H_sp=a.H_sp, W_sp=a.W_sp) for a in blk.attns]
    out = torch.empty_like(x)
    fn = lambda: ops.qkv_lepe_attention(x, wq, bq, (st, csq, 1e-5), brs, reso, float(blk.attns[0].scale), out=out)
    names = ["entry", "prologue", "acc_ready", "tiles_published", "pv_issued", "-", "-", "exit"]
elif kind == "mlp":
    M, C = int(sys.argv[2]), int(sys.argv[3])
    x = torch.randn(M, C, device=DEV, dtype=torch.bfloat16)
    w1 = (torch.randn(4 * C, C, device=DEV) / C ** 0.5).bfloat16(); w2 = (torch.randn(C, 4 * C, device=DEV) / (4 * C) ** 0.5).bfloat16()
    cs, b1, b2 = w1.float().sum(1), torch.randn(4 * C, device=DEV) * 0.1, torch.randn(C, device=DEV) * 0.1
    st = ops.row_stats(x)
    fn = lambda: ops.mlp_fused(x, w1, cs, b1, w2, b2, st, 1e-5)
    names = ["entry", "prologue", "tma_issued", "x_landed", "mma_issued", "acc2_ready", "exchanged", "reduced",
             "cyc:mma_wait_ring", "cyc:mma_wait_acc1free", "cyc:mma_wait_H", "cyc:mma_total",
             "cyc:epi_wait_acc1", "cyc:issue_mma2", "cyc:commits", "cyc:issue_mma1"]
else:
    M, N, K = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    act = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    res = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16); w = torch.randn(N, K, device=DEV, dtype=torch.bfloat16) / K ** 0.5
    bias = torch.randn(N, device=DEV, dtype=torch.bfloat16); r = torch.randn(M, N, device=DEV, dtype=torch.bfloat16) if res else None
    o = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    fn = lambda: ops.linear(a, w, bias, act=act, residual=r, out=o)
    names = ["entry", "prologue", "tma_issued", "first_landed", "mma_issued", "acc_ready", "epi_done", "exit"]
for _ in range(5): fn()
torch.cuda.synchronize()
buf = torch.zeros(1024 * 16, dtype=torch.int64, device=DEV)
_lib.lib().cswin_debug_set_trace(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); e1.record(); torch.cuda.synchronize()
_lib.lib().cswin_debug_set_trace(None)
NS = len(names)
t = buf.cpu().numpy().reshape(1024, 16)[:, :max(NS, 8)].astype(np.float64)
live = t[:, 0] > 0
t = t[live]
t0 = t[:, 0].min()
if NS > 8:
    print(f"{kind} {sys.argv[2:]}: {live.sum()} CTAs traced; kernel (events) {e0.elapsed_time(e1)*1e3:.1f} us; first entry -> last exit {(t[:,7].max()-t0)/1e3:.2f} us")
    for i in range(8, NS):
        print(f"  {names[i]:24s} median {np.median(t[:, i]):9.0f} cycles  (needs -DCSWIN_MLP_PROFILE)")
    order = np.argsort(np.median(t[:, :8] - t0, axis=0))
    for i in order:
        print(f"  {names[i]:14s} median {np.median(t[:, i] - t0)/1e3:8.2f} us   (per-CTA since entry {np.median(t[:, i] - t[:, 0])/1e3:7.2f})")
    sys.exit(0)
print(f"{kind} {sys.argv[2:]}: {live.sum()} CTAs traced; kernel (events) {e0.elapsed_time(e1)*1e3:.1f} us; first entry -> last exit {(t[:,7].max()-t0)/1e3:.2f} us")
print("phase          median-start(us)  median-dur-to-next(us)   [relative to first CTA entry]")
for i, n in enumerate(names):
    d = np.median(t[:, i + 1] - t[:, i]) / 1e3 if i < 7 else 0.0
    print(f"  {n:14s} {np.median(t[:, i] - t0)/1e3:10.2f} {d:14.2f}")
if os.environ.get("GEMM_PROFILE"):
    tt = buf.cpu().numpy().reshape(1024, 16)[live].astype(np.float64)
    lab = ["tmem_ld issue", "tmem_ld done", "math+st.shared+syncwarp", "lds+residual issued", "residual arrived", "stores done"]
    if os.environ.get("GEMM_PROFILE") == "tma":          # TMA-store fast path stamps
        lab = ["unit start", "tmem_ld + residual loads done", "staging box free (wait_read)", "math + st.shared", "fence.proxy.async + syncwarp", "TMA store issued"]
    for i in range(9, 14):
        print(f"  cyc {lab[i-8]:28s} +{np.median(tt[:, i] - tt[:, i-1]):7.0f}")
print(f"per-CTA lifetime median {np.median(t[:,7]-t[:,0])/1e3:.2f} us, max {np.max(t[:,7]-t[:,0])/1e3:.2f}; entry spread {np.ptp(t[:,0])/1e3:.2f} us")
