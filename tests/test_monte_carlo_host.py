"""CPU: the host side of the Monte-Carlo driver (monte_carlo.TrackAssociator, the stand-in for the reference's descriptor
matcher on synthetic streams) against what the reference's add_camera_measurements did in a whole run
(tests/golden/small_trace.npz): same actions, same new observations, same order of the new tracks, frame after frame."""
import os

import numpy as np

from oracle import msckf_oracle as o
from monocular_visual_inertial_msckf_b200 import synth
from monocular_visual_inertial_msckf_b200.monte_carlo import TrackAssociator


def test_association_matches_the_reference_matcher(golden_dir):
    tr = o.load_trace(os.path.join(golden_dir, "small_trace.npz"))
    z = np.load(os.path.join(golden_dir, "small_trace.npz"))
    kw = {k[4:]: z[k].item() for k in z.files if k.startswith("seq.")} if any(k.startswith("seq.") for k in z.files) else \
        dict(seed=42, n_frames=60, n_landmarks=40, noise_level="mid", p_drop=0.06, p_outlier=0.03, shuffle=True)
    seq = synth.make_sequence(**kw)
    apps, frames = tr["appends"], tr["frames"]
    assert len(apps) == len(seq.frames) - 1          # the reference drops the very first image (IMU not initialised yet)
    landmark_of = {}                                   # reference feature id -> landmark id
    assoc = TrackAssociator()
    n_new = n_lost = n_app = 0
    for i, a in enumerate(apps):
        fr = seq.frames[i + 1]
        prev_ids = frames[i - 1]["feat_ids"] if i > 0 else np.zeros(0, np.int64)
        assoc.slots = [landmark_of[int(k)] for k in prev_ids]
        action, uv, conf = assoc.frame(fr.landmark_ids, fr.keypoints, fr.scores)
        assert len(action) == len(a["action"]), i
        assert (action == a["action"]).all(), i
        assert np.array_equal(uv, a["new_uv"]) and np.array_equal(conf, a["new_conf"]), i
        for j, fid in enumerate(a["feat_ids"]):
            if a["action"][j] == 3:
                landmark_of[int(fid)] = assoc.slots[j]
        n_new += int((action == 3).sum()); n_lost += int((action == 1).sum()); n_app += int((action == 2).sum())
    assert n_new >= 40 and n_lost > 20 and n_app > 1000


def test_apply_map_keeps_order():
    a = TrackAssociator()
    a.slots = [7, 3, 9, 4]
    a.apply_map(np.array([0, -1, 1, 2, -1, -1]))
    assert a.slots == [7, 9, 4]
