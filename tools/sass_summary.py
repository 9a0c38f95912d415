#!/usr/bin/env python
"""Per-kernel SASS evidence from the shipped library: counts of the Blackwell-native mnemonics (UTC*MMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy) and of the legacy tensor path (HMMA = mma.sync).
usage: sass_summary.py [path/to/libcswin_b200.so]   (runs `cuobjdump -sass`; no GPU needed)"""
import collections, os, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cswin_unet_b200", "libcswin_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True, check=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU", "BAR.SYNC", "REDG", "RED.", "ATOMG"]
cur, rows = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); rows[cur] = collections.Counter(); continue
    if cur is None: continue
    mm = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if mm:
        op = mm.group(1)
        rows[cur]["instr"] += 1
        for p in pats:
            if op.startswith(p): rows[cur][p] += 1
def demangle(n):
    try: return subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()
    except Exception: return n
print(f"{len(rows)} kernels in {os.path.basename(so)} (sm_100a SASS); tensor-core kernels first")
print(f"{'kernel':70s} {'instr':>6s} " + " ".join(f"{p:>8s}" for p in pats))
def short(n):
    d = demangle(n)
    d = re.sub(r"\(anonymous namespace\)::|cswin::|void ", "", d)
    return re.sub(r"\(.*", "", d)[:70]
for n, c in sorted(rows.items(), key=lambda kv: (-(kv[1]["UTCHMMA"] + kv[1]["UTCQMMA"]), short(kv[0]))):
    print(f"{short(n):70s} {c['instr']:6d} " + " ".join(f"{c[p]:8d}" for p in pats))
tot_h = sum(c["HMMA"] for c in rows.values())
print(f"legacy HMMA (mma.sync) instructions in the whole library: {tot_h}")
