#!/usr/bin/env python
"""One line per `ncu --set full` capture (ncu -i X.ncu-rep --page raw --csv > X.raw.csv): time, DRAM bytes and achieved GB/s, tensor-pipe
%, issue-slot utilisation, SM / L2 / DRAM throughput %, occupancy, registers, grid.  usage: ncu_full_table.py a.raw.csv [b.raw.csv ...]"""
import csv, os, re, sys
M = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
     ("lts__t_sector_hit_rate.pct", "l2hit%"), ("smsp__inst_executed.sum", "warp_inst"),
     ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
print(f"{'capture':34s} {'kernel':44s} " + " ".join(f"{l:>10s}" for _, l in M) + f" {'dramGB/s':>9s}")
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        print(f"{os.path.basename(f):34s} (empty)"); continue
    hdr, units = rows[0], rows[1]
    for d in rows[2:]:
        name = re.sub(r"cswin::|<unnamed>::|\(anonymous namespace\)::|void ", "", d[hdr.index("Kernel Name")]).split("(")[0][:44]
        vals = {}
        for m, lab in M:
            if m not in hdr:
                vals[lab] = float("nan"); continue
            i = hdr.index(m)
            try:
                vals[lab] = float(d[i].replace(",", "")) * SCALE.get(units[i], 1.0)
            except ValueError:
                vals[lab] = float("nan")
        gbs = (vals["dram_rd_MB"] + vals["dram_wr_MB"]) * 1e6 / (vals["time_us"] * 1e-6) / 1e9
        cap = os.path.basename(f).replace(".raw.csv", "")[:34]
        print(f"{cap:34s} {name:44s} " + " ".join(f"{vals[l]:10.3g}" if vals[l] >= 1e5 else f"{vals[l]:10.2f}" for _, l in M) + f" {gbs:9.0f}")
