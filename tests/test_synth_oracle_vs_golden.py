"""CPU: the synthesis / metrics oracle (oracle/synth_oracle.py) against fixtures produced by executing the reference's own
code (tools/make_golden_synth.py): IMU synthesis of both dataset generators, landmark projection, the error block of main.py."""
import os

import numpy as np

from helpers import rel_err
from oracle import synth_oracle as so


def test_quaternion_helpers_match_scipy():
    from scipy.spatial.transform import Rotation, Slerp
    rng = np.random.default_rng(0)
    for _ in range(200):
        R0 = Rotation.from_rotvec(rng.uniform(-3, 3, 3)).as_matrix()
        R1 = Rotation.from_rotvec(rng.uniform(-3, 3, 3)).as_matrix()
        q = Rotation.from_matrix(R0).as_quat()
        assert np.allclose(so.quat_from_matrix(R0), q, atol=1e-14)
        t = rng.uniform(0, 1)
        ref = Slerp([0, 1], Rotation.from_matrix([R0, R1]))(t).as_matrix()
        assert np.max(np.abs(so.slerp_matrix(R0, R1, t) - ref)) < 1e-13
        assert np.max(np.abs(so.euler_XYZ(R0) - Rotation.from_matrix(R0).as_euler("XYZ"))) < 1e-12


def test_imu_synthesis_photorealistic_generator(golden_dir):
    z = np.load(os.path.join(golden_dir, "synth_photoreal.npz"))
    sig = z["sigmas"]                                         # acc, gyro, acc walk, gyro walk
    noise = z["normals"] * np.array([sig[0], sig[2], sig[1], sig[3]])[None, :, None]
    poses, times, imu_gt, imu_noisy = so.synth_imu(z["key_poses"], z["key_times"], 10, z["gravity"], noise, accumulate_bias=False,
                                                   sign_fix=True)
    assert len(imu_gt) == len(z["imu_gt"]) == 9 * (len(z["key_poses"]) - 1) + 1
    assert rel_err(times, z["imu_gt"][:, 0]) < 1e-14
    assert rel_err(imu_gt[:, :3], z["imu_gt"][:, 1:4]) < 1e-9 and rel_err(imu_gt[:, 3:], z["imu_gt"][:, 4:7]) < 1e-9
    assert rel_err(imu_noisy[:, :3], z["imu_noisy"][:, 1:4]) < 1e-9 and rel_err(imu_noisy[:, 3:], z["imu_noisy"][:, 4:7]) < 1e-9
    tr = z["transforms"]                                      # timestamp, T00 T01 T02 T03, T10 ...
    R = tr[:, 1:].reshape(-1, 3, 4)[:, :, :3].reshape(-1, 9)
    p = tr[:, 1:].reshape(-1, 3, 4)[:, :, 3]
    assert rel_err(poses[:, :9], R) < 1e-12 and rel_err(poses[:, 9:], p) < 1e-12


def test_imu_synthesis_synthetic_generator(golden_dir):
    z = np.load(os.path.join(golden_dir, "synth_generator.npz"))
    P = np.vstack([z["poses"], z["last_pose"][None]])         # the generator differentiates len(frames) + 1 poses
    T = len(z["poses"])
    d = z["imu_draws"]                                        # gyro white, acc white, gyro walk, acc walk
    noise = np.stack([d[:, 1], d[:, 3], d[:, 0], d[:, 2]], 1)
    _, _, imu_gt, imu_noisy = so.synth_imu(P, np.arange(T + 1) * float(z["delta_t"]), 1, z["gravity"], noise, accumulate_bias=True,
                                           sign_fix=False)
    # the reference's `angular_velocity += noise` acts in place on the arrays its "gt" measurement already holds
    # (synthetic_generator.py:165-181), so its gt stream IS the noisy stream; the clean values are noisy - noise - bias
    assert np.array_equal(z["imu_gt"], z["imu_noisy"])
    assert rel_err(imu_noisy[1:], z["imu_noisy"]) < 1e-9
    bias_g, bias_a = np.cumsum(d[:, 2], 0), np.cumsum(d[:, 3], 0)
    clean = z["imu_noisy"] - np.hstack([d[:, 0] + bias_g, d[:, 1] + bias_a])
    assert rel_err(imu_gt[1:], clean) < 1e-9


def test_landmark_projection(golden_dir):
    z = np.load(os.path.join(golden_dir, "synth_generator.npz"))
    n_vis = 0
    for i in range(0, len(z["poses"]), 7):
        vis, kp, kpn = so.project_landmarks(z["poses"][i], z["R_W_C"], z["t_W_C"], z["K"], int(z["width"]), int(z["height"]),
                                            z["landmarks"], z["noise_uv"][i])
        assert np.array_equal(vis, z["vis"][i])
        assert rel_err(kp[vis], z["kp_gt"][i][vis]) < 1e-12 and rel_err(kpn[vis], z["kp_noisy"][i][vis]) < 1e-12
        n_vis += int(vis.sum())
    assert n_vis > 300 and abs(float(z["score"][z["vis"]].mean()) - 1 / (1 + 2 * float(z["sigma_px"]) ** 2)) < 1e-15


def test_metrics_stream(golden_dir):
    z = np.load(os.path.join(golden_dir, "metrics_main.npz"))
    out = so.metrics_stream(z["gt_R"], z["gt_p"], z["est_R"], z["est_p"], z["P"])
    for k in ("rte", "roe", "ate", "aoe", "bounds", "rel_sq_pos", "rel_sq_ori"):
        assert rel_err(out[k], z[k]) < 1e-9, k
    assert np.max(np.abs(np.round(out["summary"], 8) - z["summary"])) <= 1e-8
    # both special branches are in the fixture: zero relative motion (:349-350) and the 10x smoothing of a spike (:352-358)
    assert (z["rel_sq_pos"][:30] == 0).all() and z["rel_sq_pos"][200] > 100 * z["rte"][200] ** 2
