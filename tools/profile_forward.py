#!/usr/bin/env python
"""Per-call device time of every native op in one forward (batch B, bf16): each recorded call is re-launched 20x from its
own CUDA graph (L2-warm, as inside the step) and timed with CUDA events.  Prints calls sorted by time + family totals."""
import collections, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import ops, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = "cuda"
model = cw.cswin_tiny_224(num_classes=9).eval()
shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
model = model.to(dev); model.compute_dtype = torch.bfloat16
x = torch.from_numpy(synth.synth_image_batch(B, 3, 224, seed=0, kind="ct")).to(dev)
names = ["qkv_lepe_attention", "mlp_fused", "row_stats", "layernorm_with_row_stats", "lepe_attention_fwd", "linear", "layernorm", "im2col_tokens", "im2col_nchw", "carafe_reassemble", "carafe_head"]
calls = []
origs = {n: getattr(ops, n) for n in names}
def mk(n):
    def rec(*a, **k):
        calls.append((n, a, k)); return origs[n](*a, **k)
    return rec
with torch.no_grad():
    model(x)                      # warm caches
    for n in names: setattr(ops, n, mk(n))
    # layernorm is called inside linear(ln=...) on the bf16 path: record it separately, not twice
    model(x)
    for n in names: setattr(ops, n, origs[n])
torch.cuda.synchronize()

def desc(n, a, k):
    if n == "linear":
        m = a[0].numel() // a[0].shape[-1]
        return f"M={m} N={k.get('n_out') or a[1].shape[0]} K={a[1].shape[1]} act={k.get('act',0)} res={int(k.get('residual') is not None)} ln={int(k.get('ln') is not None)} a2={int(k.get('a2') is not None)}"
    if n == "lepe_attention_fwd":
        d = a[0][0]; return f"B={a[1]} reso={a[2]} C_b={d['q'].shape[-1]} N={d['H_sp']*d['W_sp']} branches={len(a[0])}"
    if n == "qkv_lepe_attention":
        return f"B={a[0].shape[0]} reso={a[5]} C={a[0].shape[-1]} N={a[4][0]['H_sp']*a[4][0]['W_sp']} branches={len(a[4])}"
    if n == "mlp_fused":
        return f"M={a[0].numel()//a[0].shape[-1]} C={a[0].shape[-1]}"
    if n in ("row_stats", "layernorm_with_row_stats"):
        return f"M={a[0].numel()//a[0].shape[-1]} C={a[0].shape[-1]}"
    if n == "layernorm":
        return f"M={a[0].numel()//a[0].shape[-1]} C={a[0].shape[-1]}"
    return " ".join(str(tuple(t.shape)) for t in a if torch.is_tensor(t))

rows = []
with torch.no_grad():
    for (n, a, k) in calls:
        if n == "linear" and k.get("ln") is not None:
            k = dict(k); a = (origs["layernorm"](a[0], *k.pop("ln")),) + tuple(a[1:]); k["ln"] = None    # LN timed as its own call
        fn = lambda: origs[n](*a, **k)
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s): fn()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        rows.append((e0.elapsed_time(e1) * 1e3 / 20, n, desc(n, a, k)))
tot = collections.defaultdict(lambda: [0, 0.0])
agg = collections.defaultdict(lambda: [0, 0.0])
for t, n, d in rows:
    tot[n][0] += 1; tot[n][1] += t; agg[(n, d)][0] += 1; agg[(n, d)][1] += t
print(f"batch {B}: {len(rows)} calls, sum {sum(r[0] for r in rows):.1f} us")
for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {n:22s} {c:4d} calls {t:9.1f} us")
print("by shape:")
for (n, d), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {t:8.1f} us = {c:3d} x {t/c:7.2f}  {n:20s} {d}")
