#!/usr/bin/env python
"""Summarise the ncu capture of the matching-front kernels (tools/gpu_match.sh) into profiles/<tag>_match_summary.md.
    python tools/ncu_match_summary.py gpurun_out/m4 profiles/r01v"""
import csv
import io
import os
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]

src, dst = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", os.path.join(src, "prof_match.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = [f"# Matching-front kernels, ncu `--set full` (`{src}`, 1024 filters x 256 tracks x 30 observations, 300 detections x 64-d)\n"]
seen = set()
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    if name in seen:
        continue
    seen.add(name)
    out.append(f"## `{name}`\n\n| metric | value | unit |\n|---|---|---|")
    for m in METRICS:
        if m in d:
            out.append(f"| `{m}` | {d[m]} | {units[hdr.index(m)]} |")
    out.append("")
probe = os.path.join(src, "tracks_probe.txt")
if os.path.exists(probe):
    out.append("## Timing at the benchmark scale (`tools/tracks_probe.py 4096`, CUDA events, not under the profiler)\n\n```")
    out.append(open(probe).read().rstrip())
    out.append("```")
test = os.path.join(src, "pytest_match.log")
if os.path.exists(test):
    out.append("\n## `pytest tests/test_gpu_matching.py`\n\n```\n" + "\n".join(open(test).read().rstrip().splitlines()[-3:]) + "\n```")
open(dst + "_match_summary.md", "w").write("\n".join(out) + "\n")
print("wrote", dst + "_match_summary.md")
