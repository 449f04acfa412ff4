#!/usr/bin/env python
"""Timing probe of the BASELINE.json stress shape (60-pose window, d = 375, 1024 features x 60 observations) on a few
filters: which kernels serve it and how long one update takes.  python tools/stress_probe.py [filters]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import make_case
from monocular_visual_inertial_msckf_b200 import BatchedMSCKF, UpdateParams, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
F, M, N = 1024, 60, 60
base = make_case(F, M, N, seed=7)
flt = BatchedMSCKF(B, F, M, N, UpdateParams.from_case(base, _lib.COMPRESS_GRAM), device="cuda")
flt.load_cases([base] * B)
P0 = flt.P.clone(); R0 = flt.cam_R.clone(); t0 = flt.cam_t.clone(); s0 = flt.imu_state.clone()
def restore():
    flt.P.copy_(P0); flt.cam_R.copy_(R0); flt.cam_t.copy_(t0); flt.imu_state.copy_(s0)
for stage in ("front", "compress", "apply_update"):
    best = 1e9
    for _ in range(3):
        restore()
        if stage != "front": flt.front()
        if stage == "apply_update": flt.compress()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); getattr(flt, stage)(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{stage:14s} {best:9.3f} ms for {B} filters x {F} features x {M} obs, window {N}  -> {best / B:.3f} ms/filter")
print("accepted", int(flt.n_accepted.sum()), "rejected", int(flt.n_rejected.sum()))
