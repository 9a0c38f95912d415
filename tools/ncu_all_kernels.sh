#!/bin/bash
# ncu metrics for EVERY kernel of the library in one pass (tools/ncu_all_target.py), only after the plain command exited 0.
# Output: gpurun_out/ncu_all.csv (read with tools/ncu_all_summary.py).
set -u
O=gpurun_out
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__registers_per_thread
python tools/ncu_all_target.py ${1:-24} > $O/ncu_all_plain.log 2>&1 && \
ncu --metrics $MET --clock-control none --profile-from-start off --csv --log-file $O/ncu_all.csv \
    python tools/ncu_all_target.py ${1:-24} > $O/ncu_all.log 2>&1
echo "ncu_all rc=$?"; tail -3 $O/ncu_all_plain.log; tail -2 $O/ncu_all.log
