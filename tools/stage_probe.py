#!/usr/bin/env python
"""Per-stage timing of ONE measurement update on the benchmark shape (B filters x 256 features x 30 observations, 30-pose
window; every filter the steady-state snapshot of the golden fixture with its own pixel noise): front (geometry / assembly /
factorisation from the library's own events), compression, correction, each timed with CUDA events; best of `reps`.
    python tools/stage_probe.py [filters] [reps]        (MSCKF_B200_LIB=<variant .so> selects another build)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import load_base_case  # noqa: E402
from monocular_visual_inertial_msckf_b200 import BatchedMSCKF, UpdateParams, _lib, pack_cases  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
base = load_base_case(30)
F, M, N = len(base["feat_len"]), base["obs_uv"].shape[1], int(base["n_cams"])
flt = BatchedMSCKF(B, F, M, N, UpdateParams.from_case(base, _lib.COMPRESS_GRAM, _lib.GATE_EXACT), device="cuda")
h = pack_cases([base], F, M, N, flt.ldp)
for k, v in h.items():
    getattr(flt, k).copy_(torch.from_numpy(v).to("cuda").expand_as(getattr(flt, k)))
g = torch.Generator(device="cuda").manual_seed(7)
flt.obs_uv.add_(0.1 * torch.randn(flt.obs_uv.shape, dtype=torch.float64, device="cuda", generator=g))
snap = flt.snapshot()
flt.lib.msckf_set_stage_timing(1)


def ev():
    return torch.cuda.Event(enable_timing=True)


best = None
for _ in range(reps):
    flt.restore(snap)
    e = [ev() for _ in range(4)]
    torch.cuda.synchronize()
    e[0].record(); flt.front(); e[1].record(); flt.compress(); e[2].record(); flt.apply_update(); e[3].record()
    torch.cuda.synchronize()
    t = [C.c_double(0) for _ in range(3)]
    flt.lib.msckf_front_stage_ms(*[C.byref(x) for x in t])
    v = [x.value for x in t] + [e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])]
    if best is None or sum(v) < sum(best):
        best = v
chk = float(flt.P.double().abs().sum().item())
print(f"{os.environ.get('MSCKF_B200_LIB', 'in-tree').split('/')[-1]:28s} {B} filters: geo {best[0]:.3f} asm {best[1]:.3f} chol {best[2]:.3f} "
      f"gram {best[3]:.3f} update {best[4]:.3f}  total {sum(best):.3f} ms   checksum(P) {chk:.12e}")
