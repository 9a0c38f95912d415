#!/bin/bash
# A/B sweep of the forward: sub-batch streams x fused qkv+attention kernel x GEMM shared-memory cap (co-residency of two launch chains)
out=gpurun_out/r02_sweep_streams.log; : > $out
for fuse in 0 1; do for st in 1 2 3 4; do for cap in 0 100; do
  v=$(CSWIN_STREAMS=$st CSWIN_FUSE_QKV_ATTN=$fuse CSWIN_GEMM_SMEM_CAP_KB=$cap python bench.py --steps 30 --warmup 5 --no-train --no-cpu-baseline --no-reference-cuda --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))")
  echo "fuse_qkv_attn=$fuse streams=$st gemm_smem_cap_kb=$cap : $v" | tee -a $out
done; done; done
