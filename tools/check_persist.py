"""Persistent tcgen05 Linear (gemm_tc.cu, kPersist): parity against fp64 on multi-wave shapes + timing against the one-tile form.
Run twice (CSWIN_GEMM_PERSIST=0 / 1) for the A/B; prints one line per shape."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import ops

DEV = "cuda"
torch.manual_seed(0)
CASES = [  # M, N, K, act, res, stats, fold
    (18816, 1024, 256, 1, 0, 0, 1), (18816, 768, 256, 0, 0, 0, 1), (18816, 256, 1024, 0, 1, 1, 0), (18816, 256, 256, 0, 1, 1, 0),
    (75264, 512, 128, 1, 0, 0, 1), (75264, 128, 512, 0, 1, 1, 0), (75264, 384, 128, 0, 0, 0, 1), (301056, 192, 64, 0, 0, 0, 1),
    (301056, 64, 64, 0, 1, 1, 0), (38457, 96, 152, 0, 0, 0, 0), (38457, 72, 64, 1, 1, 0, 0),
    (301056, 16, 64, 0, 0, 0, 0), (75264, 32, 128, 0, 0, 0, 0), (301056, 144, 144, 0, 0, 0, 0), (75264, 36, 288, 0, 0, 0, 0),
]
def run(M, N, K, act, res, stats, fold, check=True):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).bfloat16().to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(DEV)
    b = (0.1 * torch.randn(N, generator=g)).float().to(DEV)
    r = torch.randn(M, N, generator=g).bfloat16().to(DEV) if res else None
    kw = {}
    if fold:
        st = ops.row_stats(a)
        cs = w.float().sum(1)
        kw = dict(ln_fold=(st, cs, 1e-5), bias_f32=b)
    else:
        kw = dict(bias_f32=b) if False else {}
    bias = None if fold else b.bfloat16()
    def call():
        return ops.linear(a, w, bias, act=act, residual=r, want_stats=bool(stats), **kw)
    out = call()
    y, st_out = (out if stats else (out, None))
    err = serr = 0.0
    if check:
        A = a.double()
        if fold:
            mu = A.mean(1, keepdim=True); var = A.var(1, unbiased=False, keepdim=True)
            ref = ((A - mu) / (var + 1e-5).sqrt()) @ w.double().T + b.double()
        else:
            ref = A @ w.double().T + bias.double()
        if act: ref = torch.nn.functional.gelu(ref)
        if res: ref = ref + r.double()
        err = (y.double() - ref).abs().max().item()
        if stats:
            s = st_out.double().sum(1)
            serr = (s[:, 0] - y.double().sum(1)).abs().max().item()
    for _ in range(3): call()
    torch.cuda.synchronize()
    n = 20
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_):
        for _ in range(n): call()
    g_.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    g_.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"M={M:7d} N={N:5d} K={K:5d} act={act} res={res} stats={stats} fold={fold}: {us:8.2f} us  {2.0*M*N*K/us*1e-6:7.1f} TFLOP/s  err {err:.3e} stat-err {serr:.2e}", flush=True)
    return err
print("CSWIN_GEMM_PERSIST =", os.environ.get("CSWIN_GEMM_PERSIST", "(default)"))
bad = 0
for c in CASES:
    e = run(*c)
    if not (e <= 6e-2): bad += 1
print("bad", bad)
sys.exit(1 if bad else 0)
