#!/usr/bin/env python
"""Per-kernel device time of one eager train step (torch.profiler / CUPTI; no ncu)."""
import os, sys, collections
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
m = cw.cswin_tiny_224(num_classes=9).train()
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
m = m.cuda()
step = cw.TrainStep(m, lr=0.05, graph=False)
x = torch.from_numpy(synth.synth_image_batch(B, 3, 224, seed=0, kind="ct")).cuda()
y = torch.from_numpy(synth.synth_labels(B, 224, 9, seed=0)).cuda()
for _ in range(3): step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(x, y); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name.replace("(anonymous namespace)::", "").replace("cswin::", "").split("(")[0][-60:]
        agg[n][0] += 1; agg[n][1] += e.device_time
tot = sum(v[1] for v in agg.values()); cnt = sum(v[0] for v in agg.values())
print(f"batch {B}: {cnt} kernels, {tot/1e3:.2f} ms device time")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f"{t:9.1f} us {c:5d} x {t/c:8.2f}  {n}")

if len(sys.argv) > 2 and sys.argv[2] == "ops":
    # which torch-level (non-native) ops still launch kernels: aten op x input shapes, by device time
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof2:
        step(x, y); torch.cuda.synchronize()
    rows = [e for e in prof2.key_averages(group_by_input_shape=True) if e.key.startswith("aten::") and e.self_device_time_total > 0]
    rows.sort(key=lambda e: -e.self_device_time_total)
    print("torch-level ops with device time:")
    for e in rows[:40]:
        print(f"{e.self_device_time_total:9.1f} us {e.count:5d} x  {e.key:28s} {str(e.input_shapes)[:110]}")
