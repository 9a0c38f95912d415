#!/usr/bin/env python
"""Per-kernel table from `ncu --metrics ... --csv --log-file x.csv` (long format: one row per launch x metric).
One line per kernel NAME: launches, total / mean time, DRAM bytes per launch, achieved DRAM GB/s, tensor-pipe %, warps-active %,
SM throughput %, DRAM throughput %, L2 hit %, registers, grid range.  usage: ncu_all_summary.py x.csv [min_launches]"""
import collections, csv, re, sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ix = {n: i for i, n in enumerate(hdr)}
per = collections.OrderedDict()           # launch id -> {metric: value}
names = {}
for r in rd:
    if len(r) < len(hdr):
        continue
    lid = r[ix["ID"]]
    names[lid] = r[ix["Kernel Name"]]
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[ix["Metric Unit"]]
    m = r[ix["Metric Name"]]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)              # -> us
    if m.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)   # -> bytes
    per.setdefault(lid, {})[m] = v
    per[lid]["grid"] = r[ix["Grid Size"]]


def short(n):
    n = re.sub(r"\(anonymous namespace\)::|cswin::|void ", "", n)
    n = re.sub(r"\(.*", "", n)
    return n[:72]


agg = collections.OrderedDict()
for lid, m in per.items():
    a = agg.setdefault(short(names[lid]), collections.defaultdict(list))
    for k, v in m.items():
        a[k].append(v)
tot = sum(sum(a["gpu__time_duration.sum"]) for a in agg.values())
mean = lambda xs: sum(xs) / len(xs) if xs else float("nan")
print(f"{len(per)} launches, {len(agg)} kernels, {tot:.1f} us summed (cold-cache, serialised under ncu: compare shares)")
print(f"{'kernel':72s} {'n':>5s} {'sum us':>9s} {'share':>6s} {'mean us':>8s} {'dramMB':>8s} {'GB/s':>7s} {'tens%':>6s} {'warp%':>6s} {'sm%':>6s} {'dram%':>6s} {'l2hit':>6s} {'regs':>5s} grid")
for n, a in sorted(agg.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
    t = a["gpu__time_duration.sum"]
    by = [r + w for r, w in zip(a["dram__bytes_read.sum"], a["dram__bytes_write.sum"])]
    gbs = sum(by) / max(sum(t), 1e-9) * 1e-3
    g = a["grid"]
    print(f"{n:72s} {len(t):5d} {sum(t):9.1f} {100 * sum(t) / tot:5.1f}% {mean(t):8.2f} {mean(by) / 1e6:8.3f} {gbs:7.0f} "
          f"{mean(a['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']):6.1f} "
          f"{mean(a['sm__warps_active.avg.pct_of_peak_sustained_active']):6.1f} "
          f"{mean(a['sm__throughput.avg.pct_of_peak_sustained_elapsed']):6.1f} "
          f"{mean(a['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']):6.1f} "
          f"{mean(a['lts__t_sector_hit_rate.pct']):6.1f} {int(mean(a['launch__registers_per_thread'])):5d} {g[0]}..{g[-1]}")
