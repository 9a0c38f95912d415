#!/bin/bash
# Round-end evidence on one B200: GPU test suite, the bench line (native + both reference arms), the ncu launch list of the bench
# command, ncu metrics for every kernel and ncu --set full captures of the main kernels at the headline batch (each only after the
# plain command exited 0).  Outputs under gpurun_out/.
set -u
O=gpurun_out
T=${1:-r02}
python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${T}_pytest_gpu_final.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference_final.json 2> $O/${T}_bench_reference_final.err; echo "bench reference rc=$?"
python bench.py --steps 50 --warmup 5 > $O/${T}_bench_final.json 2> $O/${T}_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference-cuda --steps 5 --warmup 3 > $O/${T}_bench_reference_cuda_final.json 2> $O/${T}_bench_reference_cuda_final.err; echo "bench reference-cuda rc=$?"
SHORT="--steps 4 --warmup 3 --no-train --no-extras --no-cpu-baseline --no-reference-cuda"
python bench.py $SHORT > /dev/null 2>&1 && \
CSWIN_BENCH_PROFILER=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/${T}_ncu_launches_bench.csv \
    python bench.py $SHORT > $O/${T}_ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
cap() { name=$1; kre=$2; cnt=$3; shift 3; python tools/ncu_target.py "$@" > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:$kre" -s 3 -c $cnt -o $O/$name python tools/ncu_target.py "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; \
  ncu -i $O/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null; }
cap ${T}_full_attn_fwd_s3_b96 lepe_attn_fwd 1 attn 3 96
cap ${T}_full_attn_fwd_s2_b96 lepe_attn_fwd 1 attn 2 96
cap ${T}_full_attn_fwd_s1_b96 lepe_attn_fwd 1 attn 1 96
cap ${T}_full_attn_fwd_s4_b96 lepe_attn_fwd 1 attn 4 96
cap ${T}_full_linear_fc1_s3_b96 linear_tc 1 linear 18816 1024 256 1 0
cap ${T}_full_linear_qkv_s3_b96 linear_tc 1 linear 18816 768 256 0 0
cap ${T}_full_linear_fc2_s3_b96 linear_tc 1 linear 18816 256 1024 0 1
cap ${T}_full_linear_proj_s3_b96 linear_tc 1 linear 18816 256 256 0 1
cap ${T}_full_linear_persist_s1_b96 linear_tc 1 linear 301056 64 64 0 1
cap ${T}_full_conv_merge1_b96 linear_tc 1 conv 96 56 64 128 2
