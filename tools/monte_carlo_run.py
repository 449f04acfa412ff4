#!/usr/bin/env python
"""Monte-Carlo run over IMU-noise levels x sequence seeds, sharded across the GPUs of one box: every rank
owns a contiguous slice of the filter indices and keeps its filters resident for the whole sequence; the only collective is
the final gather of the per-filter absolute trajectory errors and end positions.

    python tools/monte_carlo_run.py --filters 96 --frames 40
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/monte_carlo_run.py --filters 768
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from monocular_visual_inertial_msckf_b200 import synth
from monocular_visual_inertial_msckf_b200.monte_carlo import run_monte_carlo
from monocular_visual_inertial_msckf_b200.sharding import gather_filter_results, shard_range

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--filters", type=int, default=96, help="filters in the whole job")
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--landmarks", type=int, default=48)
    ap.add_argument("--window", type=int, default=8)
    ap.add_argument("--matcher", choices=("host", "device"), default="host",
                    help="device: the descriptor matcher / geometric tests of add_camera_measurements run on the GPU")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    lo, hi = shard_range(args.filters, rank, world)
    t0 = time.time()
    res = run_monte_carlo(list(range(lo, hi)), n_frames=args.frames, n_landmarks=args.landmarks, window=args.window,
                          device=f"cuda:{local}", matcher=args.matcher)
    torch.cuda.synchronize()
    wall = time.time() - t0
    stats = torch.from_numpy(np.concatenate([res.ate_rmse[:, None], res.positions[:, -1], res.gt_positions[:, -1]], axis=1)).to(f"cuda:{local}")
    allstats = gather_filter_results(stats, args.filters, dst=0) if world > 1 else stats
    if rank == 0:
        a = allstats.cpu().numpy()
        levels = list(synth.NOISE_LEVELS.keys())
        out = dict(filters=args.filters, frames=args.frames, n_gpus=world, matcher=args.matcher, wall_s_rank0=wall, ate_rmse_m={})
        for k, name in enumerate(levels):
            sel = a[k::len(levels), 0]
            out["ate_rmse_m"][name] = dict(mean=float(sel.mean()), p50=float(np.median(sel)), max=float(sel.max()), n=int(len(sel)))
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
