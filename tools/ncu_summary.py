#!/usr/bin/env python
"""Summarise an `ncu --set full` report (ncu -i rep --page raw --csv > raw.csv): one line per launch.
usage: ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
M = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
     ("smsp__inst_executed.sum", "warp_inst"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_confl"),
     ("lts__t_sector_hit_rate.pct", "l2hit%"), ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem%")]
cols = [(hdr.index(m), lab) for m, lab in M if m in hdr]
ki = hdr.index("Kernel Name")
print("units: " + ", ".join(f"{lab}={units[i]}" for i, lab in cols))
print(f"{'kernel':58s} " + " ".join(f"{lab:>10s}" for _, lab in cols))
for d in data:
    name = d[ki].replace("cswin::<unnamed>::", "").replace("void ", "").split("(")[0][:58]
    vals = []
    for i, _ in cols:
        try:
            v = float(d[i].replace(",", ""))
            vals.append(f"{v:10.2f}" if v < 1e5 else f"{v:10.3g}")
        except ValueError:
            vals.append(f"{d[i][:10]:>10s}")
    print(f"{name:58s} " + " ".join(vals))
