#!/usr/bin/env python
"""Top stall locations from `ncu -i X.ncu-rep --page source --csv` (SASS view). usage: ncu_top_stalls.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]; end = his[1] - 1 if len(his) > 1 else len(rows)          # first profiled launch only
hdr = rows[hi]; data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
si = hdr.index("# Samples"); src = hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[:N]
for i in sorted(idx):
    r = data[i]
    st = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[si]):6d} {100*int(r[si])/max(tot,1):5.1f}%  {r[src].strip()[:70]:70s} {st}")
