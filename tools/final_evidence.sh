#!/bin/bash
# Round-end evidence on one B200: GPU test suite, the bench line, the ncu launch list of the bench command and ncu --set full
# captures of the main kernels (each only after the plain command exited 0).  Outputs under gpurun_out/.
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
python bench.py --steps 50 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
cap() { name=$1; kre=$2; cnt=$3; shift 3; python tools/ncu_target.py "$@" > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:$kre" -c $cnt -o $O/$name python tools/ncu_target.py "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; }
cap full_attn_fwd_s3 lepe_attn_fwd 2 attn 3 24
cap full_attn_fwd_s1 lepe_attn_fwd 2 attn 1 24
cap full_attn_wide_s3 lepe_attn_fwd 2 attn512 3 4
cap full_train_block3 "lepe_attn_bwd|lepe_param_grad|linear_tc_kernel|linear_wgrad" 24 block_train 3 24
