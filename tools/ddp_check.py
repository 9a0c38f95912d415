#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/ddp_check.py : every rank trains on DIFFERENT data for a few steps; with a correct gradient
all-reduce all replicas stay bit-identical-ish (same averaged gradients -> same update).  Prints the worst parameter spread
across ranks and the loss per step; run with CSWIN_DDP_OVERLAP=1 and =0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import faulthandler, time
if os.environ.get("CSWIN_HANG_DUMP"):
    faulthandler.dump_traceback_later(float(os.environ["CSWIN_HANG_DUMP"]), exit=False)
_t0 = time.time()
import cswin_unet_b200 as cw
from cswin_unet_b200 import synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
torch.manual_seed(0)
m = cw.cswin_tiny_224(num_classes=9, drop_path_rate=0.0)
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
m = m.to(dev)
step = cw.TrainStep(m, lr=0.05, warmup=2)
x = torch.from_numpy(synth.synth_image_batch(4, 3, 224, seed=10 + rank, kind="ct")).to(dev)
y = torch.from_numpy(synth.synth_labels(4, 224, 9, seed=10 + rank)).to(dev)
losses = []
for i in range(6):
    losses.append(float(step(x, y)))
    if rank == 0:
        print(f"step {i} done at {time.time() - _t0:.1f} s", flush=True)
worst, wname = 0.0, ""
for k, p in m.named_parameters():
    hi, lo = p.detach().clone(), p.detach().clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    d = float((hi - lo).abs().max() / (p.detach().abs().max() + 1e-12))
    if d > worst:
        worst, wname = d, k
if rank == 0:
    print(f"overlap={os.environ.get('CSWIN_DDP_OVERLAP', '1')} world={world} reducer={'pool' if step._reducer is not None else 'bucketed'} "
          f"losses={[round(l, 4) for l in losses]} worst relative parameter spread across ranks={worst:.3e} ({wname})")
    assert worst < 1e-6, "replicas diverged: some gradient was not all-reduced"
step.close()
del step
import gc; gc.collect()
print(f'rank {rank}: graphs released, destroying the process group', flush=True)
dist.destroy_process_group()
print(f'rank {rank}: clean exit', flush=True)
