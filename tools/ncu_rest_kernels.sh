#!/bin/bash
# ncu metrics for the kernels tools/ncu_all_kernels.sh (train step) does not reach, plus ncu --set full captures of the dominant
# Linear shape (fc1 of stage 3) and the stage-3 attention for the DRAM-traffic figures.  Each only after the plain run exited 0.
set -u
O=gpurun_out
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__registers_per_thread
python tools/ncu_rest_target.py 24 > $O/ncu_rest_plain.log 2>&1 && \
timeout 700 ncu --metrics $MET --clock-control none --profile-from-start off --csv --log-file $O/ncu_rest.csv \
    python tools/ncu_rest_target.py 24 > $O/ncu_rest.log 2>&1
echo "ncu_rest rc=$?"; tail -2 $O/ncu_rest_plain.log
cap() { name=$1; kre=$2; cnt=$3; shift 3; python tools/ncu_target.py "$@" > /dev/null 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$kre" -s 2 -c $cnt -o $O/$name python tools/ncu_target.py "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; }
cap r02_full_linear_fc1_s3 linear_tc 2 linear 4704 1024 256 1 0
cap r02_full_linear_fc2_s3 linear_tc 2 linear 4704 256 1024 0 1
cap r02_full_attn_fwd_s3 lepe_attn_fwd 2 attn 3 24
for r in r02_full_linear_fc1_s3 r02_full_linear_fc2_s3 r02_full_attn_fwd_s3; do ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null; python tools/ncu_summary.py $O/$r.raw.csv > $O/$r.summary.txt 2>&1; done
