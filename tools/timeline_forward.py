#!/usr/bin/env python
"""In-situ kernel timeline of ONE graph-replayed forward (batch B, bf16) from CUPTI (torch.profiler): per kernel start / duration,
the gap to the previous kernel's end (negative = overlapped through programmatic dependent launch), and per-kernel-name totals.
Answers: is the forward the sum of kernel bodies, or of launch gaps?   usage: timeline_forward.py [B] [--train]"""
import collections, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 24
dev = "cuda"
model = cw.cswin_tiny_224(num_classes=9).eval()
shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
model = model.to(dev); model.compute_dtype = torch.bfloat16
x = torch.from_numpy(synth.synth_image_batch(B, 3, 224, seed=0, kind="ct")).to(dev)
with torch.no_grad():
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): model(x)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y = model(x)
for _ in range(5): g.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.name and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
evs.sort(key=lambda e: e.time_range.start)
n = len(evs) // 3
evs = evs[2 * n:]                                   # last replay
t0 = evs[0].time_range.start
end_prev = None
rows = []
for e in evs:
    st, en = e.time_range.start - t0, e.time_range.end - t0
    gap = (st - end_prev) if end_prev is not None else 0.0
    rows.append((st, en - st, gap, e.name))
    end_prev = en if end_prev is None else max(end_prev, en)
total = rows[-1][0] + rows[-1][1]
busy = 0.0; cur_end = 0.0
for st, du, gap, nm in rows:
    s0 = max(st, cur_end); e0 = st + du
    if e0 > s0: busy += e0 - s0
    cur_end = max(cur_end, e0)
print(f"batch {B}: {len(rows)} kernels in one replay; first start -> last end {total:.1f} us; union of kernel intervals {busy:.1f} us; "
      f"sum of durations {sum(r[1] for r in rows):.1f} us; sum of positive gaps {sum(max(r[2], 0) for r in rows):.1f} us; "
      f"sum of overlaps {-sum(min(r[2], 0) for r in rows):.1f} us")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for st, du, gap, nm in rows:
    k = nm.split("(")[0][-60:]
    agg[k][0] += 1; agg[k][1] += du; agg[k][2] += gap
for k, (c, du, gap) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {du:8.1f} us dur = {c:3d} x {du / c:6.2f}   mean gap before {gap / c:+6.2f} us   {k}")
if "--all" in sys.argv:
    for st, du, gap, nm in rows:
        print(f"{st:9.2f} {du:7.2f} {gap:+7.2f}  {nm[:90]}")
