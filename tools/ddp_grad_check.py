#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/ddp_grad_check.py [--dtype fp32|bf16] [--global-dice] [--batch B]

N-rank gradient parity (replaces the reference's nn.DataParallel reduce, trainer.py:37-38):

  (A) every rank runs forward + backward on ITS batch with the overlapped pool all-reduce (parallel.PoolGradReducer) ->
      averaged gradients;
  (B) every rank then computes, without any collective, the gradient of EVERY rank's batch on its own GPU and averages them
      locally — the single-process value of the same per-rank loss definition.

(A) must equal (B) per parameter (relative L2; fp32 compute: <= 1e-4, bf16 compute: <= 3e-2 — the bf16 weight-gradient kernels
accumulate with fp32 atomics, so two runs of the same batch already differ at the 1e-3 level).  With --global-dice (Dice over
the global batch like nn.DataParallel) (B) is one forward + backward of the CONCATENATED batch.  Exit code 0 = parity."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import cswin_unet_b200 as cw
from cswin_unet_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
ap.add_argument("--global-dice", action="store_true")
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--classes", type=int, default=9)
args = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
dt = torch.float32 if args.dtype == "fp32" else torch.bfloat16


def make_model():
    m = cw.cswin_tiny_224(num_classes=args.classes, drop_path_rate=0.0)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
    return m.to(dev)


def batch_of(r):
    x = torch.from_numpy(synth.synth_image_batch(args.batch, 3, 224, seed=10 + r, kind="ct")).to(dev)
    y = torch.from_numpy(synth.synth_labels(args.batch, 224, args.classes, seed=10 + r)).to(dev)
    return x, y


m = make_model()
# (A) data parallel: overlapped pool reducer, eager (same code path the captured step records)
step = cw.TrainStep(m, lr=0.05, graph=False, compute_dtype=dt, global_dice=args.global_dice)
assert step._reducer is not None, "expected the overlapped PoolGradReducer"
loss_a = float(step.gradients(*batch_of(rank)))
ga = {k: p.grad.detach().clone().float() for k, p in m.named_parameters()}
n_coll = step._reducer.n_coll
# (B) single process, no collective
solo = cw.TrainStep(m, lr=0.05, graph=False, compute_dtype=dt, distributed=False)
gb = {k: torch.zeros_like(v) for k, v in ga.items()}
if args.global_dice:
    xs, ys = zip(*[batch_of(r) for r in range(world)])
    solo.gradients(torch.cat(xs), torch.cat(ys))
    for k, p in m.named_parameters():
        gb[k] += p.grad.detach().float()
else:
    for r in range(world):
        solo.gradients(*batch_of(r))
        for k, p in m.named_parameters():
            gb[k] += p.grad.detach().float() / world
worst, wname, tot_num, tot_den = 0.0, "", 0.0, 0.0
for k in ga:
    num = float((ga[k] - gb[k]).norm())
    den = float(gb[k].norm())
    tot_num += num * num; tot_den += den * den
    if den > 1e-12 and num / den > worst:
        worst, wname = num / den, k
overall = (tot_num / max(tot_den, 1e-30)) ** 0.5
tol = 1e-4 if args.dtype == "fp32" else 3e-2
w = torch.tensor([worst], device=dev)
dist.all_reduce(w, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"ddp_grad_check world={world} dtype={args.dtype} global_dice={args.global_dice} batch/rank={args.batch} "
          f"collectives={n_coll} loss={loss_a:.5f}: overall relative L2 {overall:.3e}, worst parameter {float(w):.3e} ({wname}), "
          f"tolerance {tol:.0e} -> {'OK' if float(w) <= tol else 'FAIL'}", flush=True)
ok = float(w) <= tol
step.close(); solo.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
