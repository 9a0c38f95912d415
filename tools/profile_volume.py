#!/usr/bin/env python
"""Where the wall time of predict_volume(resample='gpu') goes: pinned result allocation, host staging, device work."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import synth, ops
D, S = 150, 512
vol = np.random.default_rng(0).random((D, S, S), dtype=np.float32)
m = cw.cswin_tiny_224(num_classes=9).eval()
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
m = m.cuda()
eng = cw.SliceEngine(m, batch=24, compute_dtype=torch.bfloat16)
for _ in range(2): cw.predict_volume(eng, vol, resample="gpu")
torch.cuda.synchronize()
t = time.perf_counter(); out = torch.empty((D, S, S), dtype=torch.uint8).pin_memory(); t_pin = time.perf_counter() - t
t = time.perf_counter(); st = torch.empty((24, S, S), dtype=torch.float32); st.copy_(torch.from_numpy(vol[:24])); t_stage = time.perf_counter() - t
x = torch.empty((24, S, S), dtype=torch.float32, device="cuda"); pin = torch.empty((24, S, S), dtype=torch.float32).pin_memory()
work = torch.empty(24 * S * S, dtype=torch.float64, device="cuda")
e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
slot = eng.slots[0]
torch.cuda.synchronize()
e[0].record(); x.copy_(pin, non_blocking=True); e[1].record()
ops.zoom_cubic(x, (224, 224), out=slot["x"], work=work); e[2].record()
slot["graph"].replay(); e[3].record()
lab = ops.zoom_nearest_u8(slot["y"].contiguous(), (S, S)); e[4].record()
out[:24].copy_(lab, non_blocking=True); e[5].record()
torch.cuda.synchronize()
names = ["H2D 25 MB", "zoom_cubic (prefilter x2 + interpolation)", "forward graph", "zoom_nearest", "D2H 6.3 MB"]
print(f"pinned result buffer allocation ({D * S * S / 1e6:.0f} MB): {t_pin * 1e3:.2f} ms; host staging copy of one batch (25 MB, 1 thread): {t_stage * 1e3:.2f} ms")
for i, n in enumerate(names): print(f"  {n:44s} {e[i].elapsed_time(e[i + 1]):7.3f} ms per batch of 24")
t = time.perf_counter(); cw.predict_volume(eng, vol, resample="gpu"); torch.cuda.synchronize(); print(f"predict_volume total: {(time.perf_counter() - t) * 1e3:.1f} ms")
