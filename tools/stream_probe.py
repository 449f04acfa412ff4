#!/usr/bin/env python
"""Device-resident frame loop on device-built streams (streams.DeviceStreams + BatchedMSCKF.frame_step_detections): window,
track counts, errors and time per frame.  python tools/stream_probe.py [filters] [frames] [p_drop]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from monocular_visual_inertial_msckf_b200 import BatchedMSCKF, MetricsAccumulator, UpdateParams, _lib  # noqa: E402
from monocular_visual_inertial_msckf_b200.streams import DeviceStreams  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 40
p_drop = float(sys.argv[3]) if len(sys.argv) > 3 else 0.03
p_out = float(sys.argv[4]) if len(sys.argv) > 4 else 0.02
mode = sys.argv[5] if len(sys.argv) > 5 else "vio"
min_par = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
t0 = time.time()
st = DeviceStreams(range(B), nf, n_landmarks=256, p_drop=p_drop, p_outlier=p_out)
torch.cuda.synchronize()
print(f"streams for {B} filters x {nf} frames built in {time.time() - t0:.2f} s; detections per frame "
      f"{float(st.count.float().mean()):.1f} (min {int(st.count.min())})")
window = 30
par = UpdateParams(K=st.K, width=st.width, height=st.height, sigma_image=0.1, min_lost=2, min_tracked=3, use_parallax=True,
                   min_parallax=min_par, compress_mode=_lib.COMPRESS_GRAM, gate_mode=_lib.GATE_EXACT)
F, M, N = 256 + 32, 32, window + 1
flt = BatchedMSCKF(B, F, M, N, par)
flt.P.zero_()
flt.n_cams.zero_()
flt.feat_len.zero_()
flt.Qc_diag.copy_(st.Qc_diag)
flt.count_feature_updates = True
acc = MetricsAccumulator(B)
for f in range(nf):
    smp, desc, kp, sc, cnt = st.frame_inputs(f)
    if mode == "imu":
        cnt = torch.zeros_like(cnt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fu0 = int(flt.feature_updates.item()) if flt.feature_updates is not None else 0
    torch.cuda.synchronize()
    e0.record()
    flt.frame_step_detections(smp, st.R_I_C, st.t_I_C, desc, kp, sc, cnt, window, min_cosine_similarity=0.82,
                              epipolar_rejection_threshold=1e9, homography_rejection_threshold=1e9)
    e1.record()
    torch.cuda.synchronize()
    acc.step(st.gt_pose(f), flt)
    err = (flt.imu_state[:, 9:12] - st.gt_pose(f)[:, 9:]).norm(dim=1)
    fu = int(flt.feature_updates.item()) - fu0
    if f in (nf - 1,) or (mode == "vio" and len(sys.argv) <= 4 and (f < 4 or f % 5 == 0 or f >= nf - 4)):
        print(f"frame {f:3d}: {e0.elapsed_time(e1):8.2f} ms  cams {int(flt.n_cams.min())}-{int(flt.n_cams.max())}  tracks "
              f"{int((flt.feat_len > 0).sum(1).min())}-{int((flt.feat_len > 0).sum(1).max())}  len max {int(flt.feat_len.max())}  "
              f"feature-updates {fu / B:7.1f}/filter  accepted {float(flt.n_accepted.float().mean()):6.1f}  |pos err| mean {float(err.mean()):.4f} max {float(err.max()):.4f}")
lv = np.array([["low", "mid", "high"].index(x) for x in st.noise_level])
errn = err.cpu().numpy()
print("final |pos err| by IMU-noise level low/mid/high:", [float(errn[lv == k].mean()) for k in range(3)], "rejected/filter", float(flt.n_rejected.float().mean()))
if mode != "imu":
    flt.check_window_overflow()
print("track overflow", int(flt._overflow.item()) if flt._overflow is not None else 0)
s = acc.summary().cpu().numpy()
print("summary mean over filters:", dict(zip(MetricsAccumulator.SUMMARY, np.round(s.mean(0), 6))))
