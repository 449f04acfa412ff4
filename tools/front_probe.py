#!/usr/bin/env python
"""Timing probe of the per-feature front end (geometry / assembly / factorisation kernels) on the benchmark shape:
B filters x 256 features x 30 observations, 30-pose window, every filter the steady-state snapshot of the golden fixture
with its own pixel noise.  python tools/front_probe.py [filters] [repeats] [gate: exact|bounds] [front|update]
("update" runs the whole measurement update, so the compression and correction kernels launch too)
Short enough to sit under `ncu -k regex:front_asm -c 1`."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import load_base_case  # noqa: E402
from monocular_visual_inertial_msckf_b200 import BatchedMSCKF, UpdateParams, _lib, pack_cases  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
gate = _lib.GATE_BOUNDS if (len(sys.argv) > 3 and sys.argv[3] == "bounds") else _lib.GATE_EXACT
base = load_base_case(30)
F, M, N = len(base["feat_len"]), base["obs_uv"].shape[1], int(base["n_cams"])
flt = BatchedMSCKF(B, F, M, N, UpdateParams.from_case(base, _lib.COMPRESS_GRAM, gate), device="cuda")
h = pack_cases([base], F, M, N, flt.ldp)
for k, v in h.items():
    getattr(flt, k).copy_(torch.from_numpy(v).to("cuda").expand_as(getattr(flt, k)))
g = torch.Generator(device="cuda").manual_seed(7)
flt.obs_uv.add_(0.1 * torch.randn(flt.obs_uv.shape, dtype=torch.float64, device="cuda", generator=g))
m0, rho0 = flt.idp_m.clone(), flt.idp_rho.clone()
whole = len(sys.argv) > 4 and sys.argv[4] == "update"
flt.lib.msckf_set_stage_timing(1)
best = None
for _ in range(reps):
    flt.idp_m.copy_(m0)
    flt.idp_rho.copy_(rho0)
    flt.update() if whole else flt.front()
    t = [C.c_double(0) for _ in range(3)]
    n = flt.lib.msckf_front_stage_ms(*[C.byref(x) for x in t])
    v = [x.value for x in t]
    if best is None or sum(v) < sum(best):
        best = v
code = flt.feat_status & 0xF
print(f"{B} filters x {F} x {M}, window {N}: geo {best[0]:.3f} ms  asm {best[1]:.3f} ms  chol {best[2]:.3f} ms  "
      f"(x{4096 // B if B <= 4096 else 1} per 4096 filters: {best[0] * 4096 / B:.2f} / {best[1] * 4096 / B:.2f} / {best[2] * 4096 / B:.2f}); "
      f"accepted {int((code == 1).sum())} gated {int((code == 2).sum())} sub-batches {n}")
