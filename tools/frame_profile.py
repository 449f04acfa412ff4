#!/usr/bin/env python
"""Where a camera frame of the resident loop spends its time: every BatchedMSCKF call of frame_step_detections timed with
CUDA events (synchronised per call, so the sum slightly exceeds the un-instrumented frame).  python tools/frame_profile.py [filters] [split]
("split": every update as its three C-ABI stages front / compress / apply_update, timed separately)"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from monocular_visual_inertial_msckf_b200 import BatchedMSCKF, UpdateParams, _lib  # noqa: E402
from monocular_visual_inertial_msckf_b200.streams import DeviceStreams  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NW = 36
st = DeviceStreams(range(B), NW + 4, n_landmarks=256, p_drop=0.03, p_outlier=0.02)
par = UpdateParams(K=st.K, sigma_image=0.1, min_lost=2, min_tracked=3, min_parallax=0.0, compress_mode=_lib.COMPRESS_GRAM)
flt = BatchedMSCKF(B, 288, 32, 31, par)
flt.P.zero_(); flt.n_cams.zero_(); flt.feat_len.zero_(); flt.Qc_diag.copy_(st.Qc_diag)
times = collections.OrderedDict()
names = ["propagate", "augment", "track_descriptors", "match_descriptors", "verify_matches", "append_descriptors", "append_observations",
         "update", "compact_tracks", "unobserved_cameras", "drop_observations", "remove_cameras", "select_poorest_cameras", "score_filter"]
active = [False]


def wrap(name):
    orig = getattr(flt, name)

    def f(*a, **k):
        if not active[0]:
            return orig(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        r = orig(*a, **k)
        e1.record()
        torch.cuda.synchronize()
        times.setdefault(name, []).append(e0.elapsed_time(e1))
        return r
    setattr(flt, name, f)


if len(sys.argv) > 2 and sys.argv[2] == "split":
    names += ["front", "compress", "apply_update"]
    names.remove("update")
    flt.lib.msckf_set_stage_timing(1)

    def split_update():
        import ctypes as C
        flt.front()
        t = [C.c_double(0) for _ in range(3)]
        flt.lib.msckf_front_stage_ms(*[C.byref(x) for x in t])
        if active[0]:
            times.setdefault("  (geo, asm, chol)", []).extend(x.value for x in t)
        flt.compress()
        flt.apply_update()
        if flt.count_feature_updates:
            pass
    flt.update = split_update
for n in names:
    wrap(n)
for f in range(NW + 4):
    active[0] = f >= NW
    times.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    flt.frame_step_detections(*st.frame_inputs(f)[:1], st.R_I_C, st.t_I_C, *st.frame_inputs(f)[1:], 30, epipolar_rejection_threshold=1e9,
                              homography_rejection_threshold=1e9)
    e1.record()
    torch.cuda.synchronize()
    if active[0]:
        tot = sum(sum(v) for v in times.values())
        print(f"frame {f}: {e0.elapsed_time(e1):.2f} ms wall (instrumented), calls {tot:.2f} ms, window now {int(flt.n_cams.max())}")
        for k, v in times.items():
            print(f"    {k:24s} " + "  ".join(f"{x:7.3f}" for x in v))
