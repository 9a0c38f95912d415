#!/bin/bash
# train-step A/B at N ranks: gradient bucket size x NCCL CTA cap x per-bucket SGD
N=${1:-2}
out=gpurun_out/r02_ddp_sweep_${N}gpu.log; : > $out
port() { echo $((29600 + RANDOM % 300)); }
for mb in 8 16 32; do for ctas in 0 4 8; do for bs in 0; do
  if [ $ctas = 0 ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$ctas; fi
  r=$(CSWIN_DDP_BUCKET_MB=$mb CSWIN_BUCKET_SGD=$bs timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) bench.py --gpus $N --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train ms', round(d['train_step'].get('ms_per_step', -1), 4), d['train_step'].get('error'))")
  echo "N=$N bucket_mb=$mb nccl_max_ctas=$ctas bucket_sgd=$bs : $r" | tee -a $out
done; done; done
