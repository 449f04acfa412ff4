#!/usr/bin/env python
"""Summarise a gpurun_out/<tag>/ directory (ncu launch list + one `--set full` report + the bench line) into
profiles/<tag>_summary.md -- the tracked evidence behind the numbers in DESIGN.md.

    python tools/ncu_summary.py gpurun_out/r01a profiles/r01a
"""
import collections
import csv
import json
import os
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
]


def launch_shares(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(r[ki].split("(")[0][-48:], [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total ms (cold, serialised) | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 5e-4:
            continue
        out.append(f"| `{k}` | {v[0]} | {v[1]:.3f} | {v[1] / tot:.3f} |")
    return "\n".join(out)


def full_metrics(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        out.append(f"\n**`{name}`**\n")
        out.append("| metric | value | unit |\n|---|---|---|")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                out.append(f"| {m} | {r[i]} | {units[i]} |")
    return "\n".join(out)


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    os.makedirs(os.path.dirname(dst) or ".", exist_ok=True)
    parts = [f"# Profile summary `{os.path.basename(src)}`\n"]
    b = os.path.join(src, "bench.json")
    if os.path.exists(b) and os.path.getsize(b):
        line = json.loads(open(b).read().strip().splitlines()[-1])
        parts.append("## bench.py line (not under a profiler)\n\n```json\n" + json.dumps(line, indent=1) + "\n```\n")
    l = os.path.join(src, "launches.csv")
    if os.path.exists(l):
        parts.append("## ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`) of "
                     "`python bench.py --steps 2 --warmup 1 --no-cpu`\n\n" + launch_shares(l) + "\n")
    rep = os.path.join(src, "prof.ncu-rep")
    if os.path.exists(rep):
        parts.append("## `ncu --set full --clock-control none`, one launch per kernel at 1024 filters (= one front-end sub-batch of the 4096-filter step; tools/gpu_check.sh)\n" + full_metrics(rep) + "\n")
    open(dst + "_summary.md", "w").write("\n".join(parts))
    print("wrote", dst + "_summary.md")
