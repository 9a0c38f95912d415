import os, time, torch, torch.distributed as dist
t0 = time.time()
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
x = torch.ones(1 << 22, device="cuda")
for _ in range(3):
    dist.all_reduce(x)
torch.cuda.synchronize()
if rank == 0:
    print(f"nccl all_reduce ok, world={dist.get_world_size()} value={x[0].item()} in {time.time() - t0:.1f} s", flush=True)
dist.destroy_process_group()
