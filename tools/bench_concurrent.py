#!/usr/bin/env python
"""Throughput of K batch-B forwards IN FLIGHT (K CUDA graphs of the same model on K streams, replayed round-robin) vs one.
usage: bench_concurrent.py [B] [K ...]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
Ks = [int(k) for k in sys.argv[2:]] or [1, 2, 3]
dev = torch.device("cuda", 0)
m = cw.cswin_tiny_224(num_classes=9).eval()
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
m = m.to(dev); m.compute_dtype = torch.bfloat16
xs = [torch.from_numpy(synth.synth_image_batch(B, 3, 224, seed=i, kind="ct")).to(dev) for i in range(4)]
for K in Ks:
    streams = [torch.cuda.Stream() for _ in range(K)]
    graphs, outs = [], []
    with torch.no_grad():
        for i, s in enumerate(streams):
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                m.predict_labels(xs[i]); m.predict_labels(xs[i])
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    outs.append(m.predict_labels(xs[i]))
            graphs.append(g)
    torch.cuda.synchronize()
    steps = 60
    def run():
        for it in range(steps):
            with torch.cuda.stream(streams[it % K]):
                graphs[it % K].replay()
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record(cur)
    for s in streams: s.wait_event(e0)
    run()
    for s in streams: cur.wait_stream(s)
    e1.record(cur); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"B={B}, {K} forwards in flight: {ms:.3f} ms per batch, {B / ms * 1e3:.0f} slices/s")
