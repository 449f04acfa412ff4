#!/usr/bin/env python
"""Digest of `ncu --page raw --csv` exports: the metrics the profiles/ summaries quote, one block per kernel.
python tools/ncu_digest.py gpurun_out/r02m_*.csv"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"), ("launch__shared_mem_per_block_static", "static smem/CTA"),
    ("launch__occupancy_limit_registers", "occ limit regs (CTAs)"), ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe (inst) %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe cycles active %"),
    ("sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_active", "tensor fp64 (DMMA) %"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor pipe inst %"),
    ("sm__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp issue active %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "shared wavefronts %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio throttle / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg throttle / issue"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no instruction / issue"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch / issue"),
    ("smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "stall tex / issue"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping / issue"),
    ("smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "stall drain / issue"),
    ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "stall imc miss / issue"),
    ("smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "selected / issue"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / inst"),
]


def digest(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    d = {n: (v, u) for n, u, v in zip(names, units, vals)}
    out = [f"### {d.get('Kernel Name', ('?',))[0][:90]}  ({path})"]
    for k, label in KEYS:
        if k in d and d[k][0] != "":
            out.append(f"- {label}: {d[k][0]} {d[k][1]}  (`{k}`)")
    return "\n".join(out), d


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(digest(p)[0])
        print()
