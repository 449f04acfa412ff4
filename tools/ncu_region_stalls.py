#!/usr/bin/env python
"""Per-region stall-reason totals from an ncu SASS source page (same region arguments as tools/ncu_regions.py).
    python tools/ncu_region_stalls.py <src.csv> <cubin> <kernel substring> <file> name:lo-hi [...]"""
import csv, re, subprocess, sys, collections
src_csv, cubin, kname, fname = sys.argv[1:5]
regions = []
for s in sys.argv[5:]:
    n, r = s.split(":"); lo, hi = r.split("-"); regions.append((n, int(lo), int(hi)))
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
addr_line = {}; cur = None; inside = False
for ln in dis:
    if ln.startswith("\t.text.") or ln.strip().startswith(".section\t.text."):
        inside = kname in ln
    m = re.match(r'\s*//## File ".*?([^/"]+)", line (\d+)', ln)
    if m: cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m and inside: addr_line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
h = rows[1]; ia = h.index('Address')
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
base = None; reg = "pre"
agg = collections.OrderedDict()
for r in rows[2:]:
    a = int(r[ia], 16)
    if base is None: base = a
    key = addr_line.get(a - base)
    if key and key[0] == fname:
        for n, lo, hi in regions:
            if lo <= key[1] <= hi: reg = n; break
    d = agg.setdefault(reg, collections.defaultdict(float))
    for i, c in stall_cols:
        d[c] += float(r[i] or 0)
tot = sum(sum(d.values()) for d in agg.values())
for k, d in agg.items():
    s = sum(d.values())
    if s == 0: continue
    top = sorted(d.items(), key=lambda kv: -kv[1])[:7]
    print(f"{k:12s} {100*s/tot:5.1f}%  " + " ".join(f"{c[6:]}:{100*v/s:.0f}%" for c, v in top))
