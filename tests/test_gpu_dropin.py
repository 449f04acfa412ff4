"""GPU: the drop-in filter class (reference API: MSCKF.process_imu / get_valid_features / update on Feature / Camera / IMU
objects) against the golden vectors captured from the unmodified reference.  These read like the reference's own calls:
build the objects the reference holds, call the method, compare the objects' state."""
import os

import numpy as np
import pytest

from helpers import rel_err
from oracle import msckf_oracle as o

pytestmark = pytest.mark.gpu

from monocular_visual_inertial_msckf_b200 import (MSCKF, Camera, Feature, IMUMeasurement, InverseDepthPoint,  # noqa: E402
                                                  Isometry3D, Line, MSCKFParameters, _lib)


def filter_from_case(c, mode):
    """The reference's objects exactly as they were when the golden case was captured (tools/make_golden.py)."""
    p = MSCKFParameters(K=np.array(c["K"], float), width=int(c["width"]), height=int(c["height"]),
                        sigma_image=float(c["sigma"]), W_gravity=np.array(c["gravity"], float),
                        use_parallax=bool(c["use_parallax"]), min_parallax=float(c["min_parallax"]),
                        min_number_of_frames_to_be_lost=int(c["min_lost"]),
                        min_number_of_frames_to_be_tracked=int(c["min_tracked"]))
    f = MSCKF(p)
    f.compress_mode = mode
    imu = f.state.imu
    imu.T_W_Ii = Isometry3D(np.array(c["imu_R"], float), np.array(c["imu_t"], float))
    imu.v_W_Ii = np.array(c["imu_v"], float)
    imu.gyroscope_bias = np.array(c["imu_bg"], float)
    imu.accelerometer_bias = np.array(c["imu_ba"], float)
    f.state.covariance = np.array(c["P"], float)
    n = int(c["n_cams"])
    ids = [1000 + 7 * s for s in range(n)]          # camera keys are imu ids, not slots (MSCKF.py:255)
    for s in range(n):
        f.state.cameras[ids[s]] = Camera(p.K, p.width, p.height, Isometry3D(np.array(c["cam_R"][s], float), np.array(c["cam_t"][s], float)))
    feats = {}
    for j in range(len(c["feat_len"])):
        m = int(c["feat_len"][j])
        ft = Feature()
        for i in range(m):
            s = int(c["obs_slot"][j, i])
            ft.keypoints.append(np.array(c["obs_uv"][j, i], float))
            ft.scores.append(float(c["obs_conf"][j, i]))
            ft.camera_indices.append(ids[s])
            ft.lines.append(Line(f.state.cameras[ids[s]].T_W_Ci.t, np.array(c["obs_dir"][j, i], float), float(c["obs_conf"][j, i])))
        a = int(c["idp_anchor_slot"][j])
        idp = InverseDepthPoint(f.state.cameras[ids[a]].T_W_Ci if a >= 0 else Isometry3D(np.eye(3), np.array(c["idp_base"][j], float)))
        idp.m = np.array(c["idp_m"][j], float)
        idp.rho = float(c["idp_rho"][j])
        ft.inverse_depth_point = idp
        ft.lost_for_n_frames = int(c["lost_cnt"][j])
        ft.tracked_for_n_frames = int(c["tracked_cnt"][j])
        feats[int(c["feat_ids"][j])] = ft
    return f, feats, ids


@pytest.mark.parametrize("mode", [pytest.param(_lib.COMPRESS_QR, id="qr"), pytest.param(_lib.COMPRESS_GRAM, id="gram")])
@pytest.mark.parametrize("name,stride", [("small_updates.npz", 4), ("baseline_256x30_updates.npz", 1)])
def test_get_valid_features_and_update_like_the_reference(golden_dir, name, stride, mode):
    cases, _ = o.load_cases(os.path.join(golden_dir, name))
    n_updated = 0
    for c in cases[::stride]:
        f, feats, ids = filter_from_case(c, mode)
        keys = list(feats.keys())
        valid, lost = f.get_valid_features(feats)
        assert [k in valid for k in keys] == list(c["out_valid"]) and [k in lost for k in keys] == list(c["out_lost"])
        assert rel_err(np.stack([feats[k].inverse_depth_point.m for k in keys]), c["out_idp_m"]) < 1e-9
        assert rel_err(np.array([feats[k].inverse_depth_point.rho for k in keys]), c["out_idp_rho"]) < 1e-9
        P_before = f.state.covariance.copy()
        f.update(valid)
        assert f.number_of_residuals_discarded_for_gasting_test == int(c["out_n_rejected"])
        if not c["out_updated"]:
            assert np.array_equal(f.state.covariance, P_before)
            continue
        n_updated += 1
        assert rel_err(f.state.covariance, c["out_P"]) < 1e-9
        assert rel_err(f.state.imu.T_W_Ii.R, c["out_imu_R"]) < 1e-9 and rel_err(f.state.imu.T_W_Ii.t, c["out_imu_t"]) < 1e-9
        assert rel_err(f.state.imu.v_W_Ii, c["out_imu_v"]) < 1e-9
        assert rel_err(f.state.imu.gyroscope_bias, c["out_imu_bg"]) < 1e-9
        assert rel_err(f.state.imu.accelerometer_bias, c["out_imu_ba"]) < 1e-9
        cams = list(f.state.cameras.values())
        assert rel_err(np.stack([cam.T_W_Ci.R for cam in cams]), c["out_cam_R"]) < 1e-9
        assert rel_err(np.stack([cam.T_W_Ci.t for cam in cams]), c["out_cam_t"]) < 1e-9
        # aliasing survives the correction: Line.base / InverseDepthPoint.base still ARE the camera position arrays
        for ft in feats.values():
            for ln, cid in zip(ft.lines, ft.camera_indices):
                assert ln.base is f.state.cameras[cid].T_W_Ci.t
            assert all(cam.T_W_Ci_null is cam.T_W_Ci for cam in cams)
    assert n_updated >= 3


def test_process_imu_like_the_reference(golden_dir):
    cases, _ = o.load_cases(os.path.join(golden_dir, "small_propagate.npz"))
    for c in cases[::5]:
        p = MSCKFParameters(W_gravity=np.array(c["gravity"], float))
        f = MSCKF(p)
        f.continuous_noise_covariance = np.diag(np.asarray(c["Qc_diag"], float))
        imu = f.state.imu
        imu.T_W_Ii = Isometry3D(np.array(c["imu_R"], float), np.array(c["imu_t"], float))
        imu.v_W_Ii = np.array(c["imu_v"], float)
        imu.gyroscope_bias = np.array(c["imu_bg"], float)
        imu.accelerometer_bias = np.array(c["imu_ba"], float)
        imu.T_W_Ii_null = Isometry3D(np.array(c["null_R"], float), np.array(c["null_t"], float))
        imu.v_W_Ii_null = np.array(c["null_v"], float)
        imu.timestamp = float(c["timestamp"])
        d = c["P"].shape[0]
        f.state.covariance = np.array(c["P"], float)
        for s in range((d - 15) // 6):
            f.state.cameras[s] = Camera(p.K, p.width, p.height, Isometry3D())
        s = np.asarray(c["sample"], float)
        f.process_imu(IMUMeasurement(float(s[0]), s[1:4].copy(), s[4:7].copy()))
        assert rel_err(f.state.covariance, c["out_P"]) < 1e-12
        assert rel_err(imu.T_W_Ii.R, c["out_imu_R"]) < 1e-13 and rel_err(imu.T_W_Ii.t, c["out_imu_t"]) < 1e-13
        assert rel_err(imu.v_W_Ii, c["out_imu_v"]) < 1e-13
        assert imu.T_W_Ii_null is imu.T_W_Ii and imu.v_W_Ii_null is imu.v_W_Ii and imu.timestamp == float(s[0])


def test_window_surgery_like_the_reference(golden_dir):
    """state_augmentation / prune_poorest_camera_states / remove_cameras of the drop-in class on the reference's object
    model, against the oracle's restatement of the same reference steps."""
    cases, _ = o.load_cases(os.path.join(golden_dir, "mid_48x10_updates.npz"))
    c = cases[0]
    f, feats, ids = filter_from_case(c, _lib.COMPRESS_GRAM)
    f.features = feats
    n = int(c["n_cams"])
    # --- augmentation: a new pose under the key imu.id, P grows by one 6x6 block (MSCKF.py:250-265) ---
    f.state.imu.T_W_I = Isometry3D(np.eye(3), np.zeros(3))
    key = f.state.imu.id                      # the camera is filed under the IMU's running id (MSCKF.py:164, 255)
    P0 = f.state.covariance.copy()
    T_I_C = f.state.imu.T_W_I.inv() * f.T_W_C
    f.state_augmentation()
    assert list(f.state.cameras.keys())[-1] == key and len(f.state.cameras) == n + 1
    assert rel_err(f.state.covariance, o.augment_covariance(P0, T_I_C.R, T_I_C.t, f.state.imu.T_W_Ii.R)) < 1e-12
    new_cam = f.state.cameras[key].T_W_Ci
    ref_pose = f.state.imu.T_W_Ii * T_I_C
    assert rel_err(new_cam.R, ref_pose.R) < 1e-13 and rel_err(new_cam.t, ref_pose.t) < 1e-13
    # --- prune: selection as the reference's dict + stable sort, one last update, removal ---
    tracks = {k: list(ft.camera_indices) for k, ft in f.features.items()}
    m_ref, t_ref = o.select_poorest_cameras(c)
    chosen = [ids[s] for s in np.nonzero(m_ref)[0]]
    P1 = f.state.covariance.copy()
    f.prune_poorest_camera_states()
    assert all(k not in f.state.cameras for k in chosen) and len(f.state.cameras) == n + 1 - len(chosen)
    assert f.state.covariance.shape[0] == P1.shape[0] - 6 * len(chosen)
    for k, ft in f.features.items():
        assert ft.camera_indices == [cid for cid in tracks[k] if cid not in chosen]
        assert len(ft.keypoints) == len(ft.camera_indices) == len(ft.lines)
    assert all(len(tracks[k]) > 0 and all(cid in chosen for cid in tracks[k]) for k in tracks if k not in f.features)
    # --- remove_cameras alone: covariance rows / columns of that slot disappear (MSCKF.py:753-757) ---
    P2 = f.state.covariance.copy()
    first = list(f.state.cameras.keys())[0]
    f.remove_cameras({first: f.state.cameras[first]})
    assert rel_err(f.state.covariance, o.remove_camera_covariance(P2, 0)) == 0.0


def test_whole_reference_run_through_the_drop_in_class(golden_dir):
    """Sequence level, through the reference's OWN method names on the reference's object model: the operation trace of a
    whole reference run (540 process_imu, 59 state_augmentation, 85 get_valid_features + update, 79 remove_cameras) is
    replayed call by call on the drop-in class.  Camera keys come from imu.id as it advances with every process_imu
    (MSCKF.py:164, 255) -- nothing is set by hand -- the covariance is uploaded ONCE and stays resident on the device, and
    after every camera frame the state equals the reference's to 1e-9."""
    from dropin_replay import replay_trace_through_drop_in
    f, worst = replay_trace_through_drop_in(lambda p: MSCKF(p), golden_dir)
    print("drop-in class over the whole reference run: worst relative error", worst)
