#!/usr/bin/env python
"""Timing of the IMU propagation alone: B filters, 30-pose window, n samples per call (CUDA events, best of reps).
    python tools/prop_probe.py [filters] [samples] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import load_base_case  # noqa: E402
from monocular_visual_inertial_msckf_b200 import BatchedMSCKF, UpdateParams, _lib, pack_cases  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 9
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
base = load_base_case(30)
F, M, N = len(base["feat_len"]), base["obs_uv"].shape[1], int(base["n_cams"])
flt = BatchedMSCKF(B, F, M, N, UpdateParams.from_case(base, _lib.COMPRESS_GRAM, _lib.GATE_EXACT), device="cuda")
h = pack_cases([base], F, M, N, flt.ldp)
for k, v in h.items():
    getattr(flt, k).copy_(torch.from_numpy(v).to("cuda").expand_as(getattr(flt, k)))
g = torch.Generator(device="cuda").manual_seed(3)
t0 = flt.imu_time.clone()
smp = torch.zeros((B, ns, 7), dtype=torch.float64, device="cuda")
smp[:, :, 0] = t0[:, None] + 0.005 * torch.arange(1, ns + 1, device="cuda", dtype=torch.float64)[None, :]
smp[:, :, 1:4] = 0.05 * torch.randn((B, ns, 3), dtype=torch.float64, device="cuda", generator=g)
smp[:, :, 4:7] = torch.tensor([0.0, 0.0, 9.81], dtype=torch.float64, device="cuda") + 0.2 * torch.randn((B, ns, 3), dtype=torch.float64, device="cuda", generator=g)
snap = flt.snapshot()
best = 1e9
for _ in range(reps):
    flt.restore(snap)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); flt.propagate(smp); e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
d = 15 + 6 * N
byt = B * (3 * 15 * d * 8 + ns * 56)
print(f"{os.environ.get('MSCKF_B200_LIB', 'in-tree').split('/')[-1]:24s} {B} filters x {ns} samples: {best:.3f} ms  "
      f"{byt / best / 1e6:.0f} GB/s (read 15 d + write 2 x 15 d doubles per filter)  checksum {float(flt.P.abs().sum()):.12e} {float(flt.imu_state.abs().sum()):.12e}")
