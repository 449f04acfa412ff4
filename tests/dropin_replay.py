"""Shared by the CPU and the GPU drop-in tests: replays the operation trace of a whole reference run
(tests/golden/small_trace.npz) through the reference's own method names on the drop-in class."""
import os

import numpy as np

from helpers import rel_err
from oracle import msckf_oracle as o
from monocular_visual_inertial_msckf_b200 import (Feature, IMUMeasurement, InverseDepthPoint, Isometry3D, Line,
                                                  MSCKFParameters, _lib)


def replay_trace_through_drop_in(make_filter, golden_dir):
    """make_filter(MSCKFParameters) -> drop-in filter object.  Returns (filter, worst relative state error)."""
    tr = o.load_trace(os.path.join(golden_dir, "small_trace.npz"))
    ini, cases = tr["init"], tr["cases"]
    c0 = cases[0]
    p = MSCKFParameters(K=np.array(c0["K"], float), width=int(c0["width"]), height=int(c0["height"]), sigma_image=float(c0["sigma"]),
                        W_gravity=np.array(ini["gravity"], float), use_parallax=bool(c0["use_parallax"]),
                        min_parallax=float(c0["min_parallax"]), min_number_of_frames_to_be_lost=int(c0["min_lost"]),
                        min_number_of_frames_to_be_tracked=int(c0["min_tracked"]), max_number_of_camera_states=int(tr["max_cams"]),
                        T_W_C=Isometry3D(np.array(ini["R_I_C"], float), np.array(ini["t_I_C"], float)))
    f = make_filter(p)
    f.gate_mode = _lib.GATE_EXACT
    f.continuous_noise_covariance = np.diag(np.asarray(ini["Qc_diag"], float))
    imu = f.state.imu
    imu.T_W_I = Isometry3D(np.eye(3), np.zeros(3))
    imu.T_W_Ii = Isometry3D(np.array(ini["imu_R"], float), np.array(ini["imu_t"], float))
    imu.v_W_Ii = np.array(ini["imu_v"], float)
    imu.gyroscope_bias, imu.accelerometer_bias = np.array(ini["imu_bg"], float), np.array(ini["imu_ba"], float)
    imu.T_W_Ii_null = Isometry3D(np.array(ini["null_R"], float), np.array(ini["null_t"], float))
    imu.v_W_Ii_null = np.array(ini["null_v"], float)
    imu.timestamp = float(ini["timestamp"])
    f.state.covariance = np.array(ini["P"], float)
    boundaries = {int(fr["op_index"]): i for i, fr in enumerate(tr["frames"])}
    worst, n_imu, keys_seen = 0.0, 0, []
    for n, (kind, idx) in enumerate(tr["ops"]):
        if kind == 0:
            s = np.asarray(tr["samples"][idx], float)
            f.process_imu(IMUMeasurement(float(s[0]), s[1:4].copy(), s[4:7].copy()))
            n_imu += 1
            assert imu.id == n_imu
        elif kind == 1:
            f.state_augmentation()
            assert list(f.state.cameras.keys())[-1] == imu.id
            keys_seen.append(imu.id)
        elif kind == 2:
            c = cases[idx]
            cam_keys = list(f.state.cameras.keys())
            assert len(cam_keys) == int(c["n_cams"])
            feats = {}
            for j in range(len(c["feat_len"])):
                ft = Feature()
                for i in range(int(c["feat_len"][j])):
                    cam = f.state.cameras[cam_keys[int(c["obs_slot"][j, i])]]
                    ft.keypoints.append(np.array(c["obs_uv"][j, i], float))
                    ft.scores.append(float(c["obs_conf"][j, i]))
                    ft.camera_indices.append(cam_keys[int(c["obs_slot"][j, i])])
                    ft.lines.append(Line(cam.T_W_Ci.t, np.array(c["obs_dir"][j, i], float), float(c["obs_conf"][j, i])))
                a = int(c["idp_anchor_slot"][j])
                idp = InverseDepthPoint(f.state.cameras[cam_keys[a]].T_W_Ci if a >= 0 else
                                        Isometry3D(np.eye(3), np.array(c["idp_base"][j], float)))
                idp.m, idp.rho = np.array(c["idp_m"][j], float), float(c["idp_rho"][j])
                ft.inverse_depth_point = idp
                ft.lost_for_n_frames, ft.tracked_for_n_frames = int(c["lost_cnt"][j]), int(c["tracked_cnt"][j])
                feats[int(c["feat_ids"][j])] = ft
            valid, lost = f.get_valid_features(feats)
            assert [k in valid for k in feats] == list(c["out_valid"])
            if len(valid) > 0:
                f.update(valid)
        elif kind == 3:
            cam_keys = list(f.state.cameras.keys())
            f.features = {}
            f.remove_cameras({cam_keys[s]: f.state.cameras[cam_keys[s]] for s in np.nonzero(tr["masks"][idx])[0]})
        if (n + 1) in boundaries:
            fr = tr["frames"][boundaries[n + 1]]
            cams = list(f.state.cameras.values())
            errs = [rel_err(f.state.covariance, fr["P"]), rel_err(imu.T_W_Ii.R, fr["imu_R"]), rel_err(imu.T_W_Ii.t, fr["imu_t"]),
                    rel_err(imu.v_W_Ii, fr["imu_v"]), rel_err(imu.gyroscope_bias, fr["imu_bg"]),
                    rel_err(imu.accelerometer_bias, fr["imu_ba"])]
            assert f.state.covariance.shape[0] == 15 + 6 * len(cams) == fr["P"].shape[0]
            if cams:
                errs += [rel_err(np.stack([c_.T_W_Ci.R for c_ in cams]), fr["cam_R"]),
                         rel_err(np.stack([c_.T_W_Ci.t for c_ in cams]), fr["cam_t"])]
            worst = max(worst, max(errs))
            assert worst < 1e-9, (boundaries[n + 1], errs)
    assert len(set(keys_seen)) == len(keys_seen) == 59 and n_imu == 540
    assert f._b200_res.p_uploads == 1          # resident across all 763 calls
    assert len(f.estimated_world_points) > 100
    return f, worst
