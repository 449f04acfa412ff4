#!/usr/bin/env python
"""Timing of the window / track surgery kernels at the benchmark's scale: 4096 filters x 256 tracks x
30 observations, 30-pose window.  HBM-bound integer / byte work: reported as ms and GB/s of the bytes each call touches.
    python tools/tracks_probe.py [filters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from monocular_visual_inertial_msckf_b200 import BatchedMSCKF, UpdateParams, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
base = bench.load_base_case(30)
F, M, N = len(base["feat_len"]), base["obs_uv"].shape[1] + 1, int(base["n_cams"]) + 1
flt = BatchedMSCKF(B, F, M, N, UpdateParams.from_case(base, _lib.COMPRESS_GRAM), device="cuda")
flt.load_cases([base] * B)
snap = {n: getattr(flt, n).clone() for n in _lib.TRACK_FIELDS + ("P", "cam_R", "cam_t", "n_cams", "imu_state")}
obs_bytes = 52.0 * float(flt.feat_len.sum().item())

def restore():
    for n, t in snap.items():
        getattr(flt, n).copy_(t)

def timed(fn, prep=lambda: None, reps=3):
    best = 1e9
    for _ in range(reps):
        restore(); prep()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

d = 15 + 6 * (N - 1)
action = torch.full((B, F), 2, dtype=torch.uint8, device="cuda")
new_uv = torch.rand((B, F, 2), dtype=torch.float64, device="cuda") * 400
new_conf = torch.rand((B, F), dtype=torch.float64, device="cuda")
mask = torch.zeros((B, N), dtype=torch.uint8, device="cuda"); mask[:, 0] = 1; mask[:, 3] = 1
R_I_C, t_I_C = np.array([[0.0, 0, 1], [-1, 0, 0], [0, -1, 0]]), np.zeros(3)
rows = [
    ("augment (P grows by one pose)", lambda: flt.augment(R_I_C, t_I_C), None, 18.0 * 8 * (d + 6) * B),
    ("append_observations", lambda: flt.append_observations(action, new_uv, new_conf), None, (25.0 + 52.0 + 16.0) * B * F),
    ("select_poorest_cameras", lambda: flt.select_poorest_cameras(2), None, 4.0 * float(flt.feat_len.sum().item()) + 4.0 * B * F),
    ("drop_observations (2 cameras)", lambda: flt.drop_observations(mask), None, 2.0 * obs_bytes),
    ("remove_cameras (2 cameras)", lambda: flt.remove_cameras(mask), None, 2.0 * 8 * d * d * B),
    ("compact_tracks", lambda: flt.compact_tracks(False), None, 2.0 * obs_bytes),
    ("unobserved_cameras", lambda: flt.unobserved_cameras(), None, 4.0 * float(flt.feat_len.sum().item())),
]
# matching front: proposals = every track re-detected near its last observation; 300 detections x 64-d
F2, D = 300, 64
gen = torch.Generator(device="cuda").manual_seed(1)
desc_t = torch.nn.functional.normalize(torch.randn((B, F, D), device="cuda", generator=gen), dim=2).contiguous()
desc_n = torch.zeros((B, F2, D), device="cuda")
desc_n[:, :F] = desc_t
desc_n[:, F:] = torch.nn.functional.normalize(torch.randn((B, F2 - F, D), device="cuda", generator=gen), dim=2)
desc_n = torch.nn.functional.normalize(desc_n + 0.03 * torch.randn((B, F2, D), device="cuda", generator=gen), dim=2).contiguous()
kp_n = torch.rand((B, F2, 2), dtype=torch.float64, device="cuda") * 400
sc_n = torch.rand((B, F2), dtype=torch.float64, device="cuda")
n_trk = torch.full((B,), F, dtype=torch.int32, device="cuda")
last = (flt.feat_len - 1).clamp(min=0).long()
uv_last = torch.gather(flt.obs_uv, 2, last[:, :, None, None].expand(B, F, 1, 2))[:, :, 0]
prop_uv = (uv_last + 0.5 * torch.randn((B, F, 2), dtype=torch.float64, device="cuda")).contiguous()
act_v = [None]
def prep_verify():
    act_v[0] = torch.full((B, F), 2, dtype=torch.uint8, device="cuda")
    flt.imu_state[:, 9:12] += 0.05     # the platform moved 5 cm since the last frame (at rest every walk would end in the homography branch)
    flt.augment(R_I_C, t_I_C)          # the frame's own pose enters the window before add_camera_measurements (MSCKF.py:150-153)
match_flop = 2.0 * 2.0 * F * F2 * D * B
rows += [
    ("match_descriptors (256x300x64)", lambda: flt.match_descriptors(desc_t, desc_n, 0.82, n_trk, None, kp_n, sc_n), None,
     4.0 * (F + F2) * D * B + 29.0 * F * B),
    ("verify_matches (epipolar walk)", lambda: flt.verify_matches(act_v[0], prop_uv, 5.0, 5.0), prep_verify,
     20.0 * float(flt.feat_len.sum().item()) + 18.0 * B * F),
]
print(f"{B} filters x {F} tracks x {M - 1} observations, window {N - 1} (capacity {N}); HBM peak 6550 GB/s")
for name, fn, prep, nbytes in rows:
    ms = timed(fn, prep or (lambda: None))
    if name.startswith("match_descriptors"):
        print(f"{name:34s} {ms:8.3f} ms   {match_flop / 1e9:7.2f} GFLOP fp32        {match_flop / (ms * 1e-3) / 1e12:8.2f} TFLOP/s")
    print(f"{name:34s} {ms:8.3f} ms   {nbytes / 1e9:7.2f} GB touched   {nbytes / (ms * 1e-3) / 1e9:8.1f} GB/s")
