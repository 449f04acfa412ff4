#!/usr/bin/env python
"""Split an ncu SASS source page into program regions by source-line ranges of the kernel's own file (inlined helper lines
inherit the region of the surrounding code in address order).
    python tools/ncu_regions.py <src.csv> <cubin> <kernel substring> <file> <n_units> name:lo-hi [name:lo-hi ...]"""
import csv, re, subprocess, sys, collections
src_csv, cubin, kname, fname, nunits = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
regions = []
for s in sys.argv[6:]:
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
h = rows[1]; ia = h.index('Address'); ie = h.index('Instructions Executed'); ins = h.index('# Samples'); iso = h.index('Source')
base = None; reg = "pre"
agg = collections.OrderedDict()
ops = collections.defaultdict(lambda: collections.defaultdict(float))
for r in rows[2:]:
    a = int(r[ia], 16)
    if base is None: base = a
    key = addr_line.get(a - base)
    if key and key[0] == fname:
        for n, lo, hi in regions:
            if lo <= key[1] <= hi: reg = n; break
    v = agg.setdefault(reg, [0.0, 0.0, 0])
    e = float(r[ie] or 0)
    v[0] += e; v[1] += float(r[ins] or 0); v[2] += 1
    op = r[iso].split()[0] if not r[iso].startswith('@') else r[iso].split()[1]
    ops[reg][op.split('.')[0]] += e
te = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total warp-instr {te:.0f} ({te/nunits:.1f}/unit), samples {ts:.0f}")
for k, v in agg.items():
    top = sorted(ops[k].items(), key=lambda kv: -kv[1])[:8]
    print(f"{k:12s} sass={v[2]:5d} instr/unit={v[0]/nunits:8.1f} exec%={100*v[0]/te:5.1f} samples%={100*v[1]/ts:5.1f}  " +
          " ".join(f"{o}:{c/nunits:.0f}" for o, c in top))
