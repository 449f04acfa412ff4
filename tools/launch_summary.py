#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): launches, total ms, share.
python tools/launch_summary.py profiles/r02_launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) != len(H):
        continue
    d = dict(zip(H, r))
    v = float(d["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1.0)
    name = d["Kernel Name"]
    name = name.replace("<unnamed>::", "").split("(")[0][:70]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
print(f"{sum(v[0] for v in agg.values())} launches, {tot:.2f} ms in total (per-launch durations are cold-cache and serialised)")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f} % |")
