#!/usr/bin/env python
"""Stall samples per CUDA source line from an .ncu-rep captured with --import-source on (needs -lineinfo):
python tools/ncu_lines.py report.ncu-rep [top]   -> file:line, samples, share, instructions executed, dominant stall reasons"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try:
            n = int(d["# Samples"])
        except (KeyError, ValueError):
            continue
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
        lines.append((n, fname, int(r[0]), r[1].strip()[:70], d.get("Instructions Executed", "?"), stalls))
tot = sum(x[0] for x in lines) or 1
print(f"{tot} samples")
for n, f, ln, src, ie, st in sorted(lines, key=lambda x: -x[0])[:top]:
    s = " ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(f"{100 * n / tot:5.1f}%  {f}:{ln:<4d} inst {ie:>10s}  {s}\n        {src}")
