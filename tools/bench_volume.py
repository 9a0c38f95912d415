#!/usr/bin/env python
"""BASELINE configs[3]: Synapse-shaped synthetic CT volume (D slices of 512x512 -> 224x224 -> 512x512), the slice loop of
test_single_volume (utils.py:61-90) on one GPU: host-side scipy resampling (the reference's way) vs both zooms on the device.
Usage: python tools/bench_volume.py [D]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import synth
D = int(sys.argv[1]) if len(sys.argv) > 1 else 150
S = 512
g = np.random.default_rng(0)
yy, xx = np.mgrid[0:S, 0:S].astype(np.float32) / S
vol = np.zeros((D, S, S), np.float32)
for d in range(D):                                                    # smooth blobs drifting through the volume, values in [0, 1]
    for k in range(6):
        cy, cx, r = 0.2 + 0.6 * g.random(), 0.2 + 0.6 * g.random(), 0.05 + 0.15 * g.random()
        vol[d] += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r)).astype(np.float32) * g.random()
vol /= max(vol.max(), 1e-6)
m = cw.cswin_tiny_224(num_classes=9).eval()
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
m = m.cuda()
eng = cw.SliceEngine(m, batch=24, compute_dtype=torch.bfloat16)
res = {}
for mode in ("gpu", "scipy"):
    cw.predict_volume(eng, vol[:24], resample=mode)                  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lab, _ = cw.predict_volume(eng, vol, resample=mode)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res[mode] = lab
    print(f"volume {D} x {S}x{S}, resample={mode:5s}: {dt * 1e3:9.1f} ms  = {D / dt:9.1f} slices/s end to end (host volume in, host label volume out)")
print(f"label agreement gpu vs scipy resampling: {(res['gpu'] == res['scipy']).mean():.6f}")
