#!/usr/bin/env python
"""Op-level timings on one GPU (CUDA events, L2 flushed between iterations): the LePEAttention op sweep of
BASELINE.json configs[1] and the Linear shapes of cswin_tiny_224_lite.  Prints one JSON line per case with the
achieved fraction of the measured HBM / bf16 peak.  Usage: python tools/bench_ops.py [attn|linear|all] [B ...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cswin_unet_b200 as cw  # noqa: E402
from cswin_unet_b200 import ops  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
DEV = "cuda"
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    _flush.zero_()


def time_op(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    if os.environ.get("BENCH_NOFLUSH"):          # L2-warm, 50 back-to-back launches captured in a CUDA graph
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            for _ in range(50):
                fn()
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        t = a.elapsed_time(b) * 1e-3 / 50
        return t, t
    ts = []
    for _ in range(iters):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3, ts[0] * 1e-3


def attn_cases(B, dtype):
    stages = [(64, 56, 2, 1, False), (128, 28, 4, 2, False), (256, 14, 8, 7, False), (512, 7, 16, 7, True)]
    for (C, reso, heads, split, last) in stages:
        blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).to(DEV).eval()
        L = reso * reso
        qkv = torch.randn(B, L, 3 * C, device=DEV, dtype=dtype)
        out = torch.empty(B, L, C, device=DEV, dtype=dtype)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        if blk.branch_num == 2:
            h = C // 2
            descs = [a.branch_desc(q[..., i * h:(i + 1) * h], k[..., i * h:(i + 1) * h], v[..., i * h:(i + 1) * h],
                                   out[..., i * h:(i + 1) * h]) for i, a in enumerate(blk.attns)]
        else:
            descs = [blk.attns[0].branch_desc(q, k, v, out)]
        scale = float(blk.attns[0].scale)
        fn = lambda: ops.lepe_attention_fwd(descs, B, reso, scale, dtype)
        med, best = time_op(fn)
        es = 2 if dtype == torch.bfloat16 else 4
        byts = 4 * B * L * C * es
        N = blk.attns[0].H_sp * blk.attns[0].W_sp
        flops = 4 * B * L * N * C + 18 * B * L * C
        print(json.dumps({"op": "lepe_attention_fwd", "dtype": str(dtype).split(".")[-1], "B": B, "C": C, "reso": reso, "N": N,
                          "us_median": round(med * 1e6, 2), "us_best": round(best * 1e6, 2), "bytes": byts,
                          "GBps": round(byts / med / 1e9, 1), "hbm_frac": round(byts / med / 1e9 / PEAKS["hbm_gbs"], 4),
                          "TFLOPs": round(flops / med / 1e12, 2)}))


def linear_cases(B, dtype):
    shapes = []
    for (C, L) in ((64, 3136), (128, 784), (256, 196), (512, 49)):
        M = B * L
        shapes += [("qkv", M, 3 * C, C, 0, 0), ("proj+res", M, C, C, 0, 1), ("fc1+gelu", M, 4 * C, C, 1, 0), ("fc2+res", M, C, 4 * C, 0, 1)]
    for (name, M, N, K, act, res) in shapes:
        a = torch.randn(M, K, device=DEV, dtype=dtype)
        w = torch.randn(N, K, device=DEV, dtype=dtype) / K ** 0.5
        bias = torch.randn(N, device=DEV, dtype=dtype)
        r = torch.randn(M, N, device=DEV, dtype=dtype) if res else None
        out = torch.empty(M, N, device=DEV, dtype=dtype)
        fn = lambda: ops.linear(a, w, bias, act=act, residual=r, out=out)
        med, best = time_op(fn)
        es = 2 if dtype == torch.bfloat16 else 4
        byts = es * (M * K + N * K + M * N * (2 if res else 1))
        flops = 2.0 * M * N * K
        print(json.dumps({"op": "linear_fwd:" + name, "dtype": str(dtype).split(".")[-1], "M": M, "N": N, "K": K,
                          "us_median": round(med * 1e6, 2), "us_best": round(best * 1e6, 2),
                          "GBps": round(byts / med / 1e9, 1), "hbm_frac": round(byts / med / 1e9 / PEAKS["hbm_gbs"], 4),
                          "TFLOPs": round(flops / med / 1e12, 2), "tensor_frac": round(flops / med / 1e12 / PEAKS["bf16_tflops"], 4)}))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    Bs = [int(b) for b in sys.argv[2:]] or [24, 192]
    t0 = cw.tc_launch_count()
    for B in Bs:
        if what in ("attn", "all"):
            attn_cases(B, torch.bfloat16)
        if what in ("linear", "all"):
            linear_cases(B, torch.bfloat16)
    print(json.dumps({"tc_launches": cw.tc_launch_count() - t0, "launches": cw.launch_count()}))
