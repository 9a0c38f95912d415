#!/bin/bash
# 2-rank checks on real GPUs: gradient parity (fp32 / bf16 / global Dice), replica identity, train-step A/B of the bucket-wise SGD
N=${1:-2}
out=gpurun_out/r02_ddp_suite_${N}gpu.log; : > $out
port() { echo $((29600 + RANDOM % 300)); }
run() { timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) "$@" 2>&1 | grep -E "ddp_grad_check|overlap=|OK|FAIL|Error|error" | tail -3 | tee -a $out; }
run tools/ddp_grad_check.py --dtype fp32
run tools/ddp_grad_check.py --dtype bf16
run tools/ddp_grad_check.py --dtype fp32 --global-dice
run tools/ddp_check.py
for v in 1 0; do
  r=$(CSWIN_BUCKET_SGD=$v timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) bench.py --gpus $N --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'fwd', round(d['value']), 'e2e', round(d['e2e']['value']), 'train ms', d['train_step'].get('ms_per_step'), d['train_step'].get('error'))")
  echo "bucket_sgd=$v : $r" | tee -a $out
done
