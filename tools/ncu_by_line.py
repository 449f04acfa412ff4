#!/usr/bin/env python
"""Aggregate an ncu SASS source page by CUDA source line, using nvdisasm -g line markers of the matching cubin.
    python tools/ncu_by_line.py <src.csv from `ncu --page source --csv`> <cubin> <kernel substring> [n_units]"""
import csv, re, subprocess, sys, collections
src_csv, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
nunits = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# restrict to the kernel's function
addr_line = {}
cur = None; inside = False
for ln in dis:
    if ln.startswith("\t.text.") or ln.strip().startswith(".section\t.text."):
        inside = kname in ln
    m = re.match(r'\s*//## File ".*?([^/"]+)", line (\d+)( inlined at ".*?([^/"]+)", line (\d+))?', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m and inside:
        addr_line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
h = rows[1]; ia = h.index('Address'); ie = h.index('Instructions Executed'); ins = h.index('# Samples')
base = None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
for r in rows[2:]:
    a = int(r[ia], 16)
    if base is None: base = a
    key = addr_line.get(a - base)
    agg[key][0] += float(r[ie] or 0); agg[key][1] += float(r[ins] or 0); agg[key][2] += 1
te = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total warp-instr {te:.0f} ({te/nunits:.1f}/unit), samples {ts:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{str(k):28s} sass={v[2]:5d} instr/unit={v[0]/nunits:8.1f} exec%={100*v[0]/te:5.1f} samples%={100*v[1]/ts:5.1f}")
