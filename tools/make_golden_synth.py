#!/usr/bin/env python
"""Golden fixtures for SURVEY.md section 8f row 4 (data synthesis + metrics), produced by running the UNMODIFIED reference
code in this container (CPU; the GPU box has no /root/reference):

  tests/golden/synth_generator.npz   MeasurementsGenerator.generate() of dataset/tools/dataset_generators/synthetic_generator.py
                                     (:130-227): poses, IMU from finite differences (:145-186), landmark projection (:189-223).
                                     The file ends in a script that writes under the author's home directory, so only its
                                     definitions (everything above the "Synthetic dataset generation" banner) are executed.
  tests/golden/synth_photoreal.npz   PhotorealisticGenerator.process_data() of photorealistic_generator.py (:47-228) in a scratch
                                     ./data tree: ground-truth camera poses -> interpolated body poses + IMU (:145-216).
  tests/golden/metrics_main.npz      the error block of the reference driver main.py (:312-391, :436-437 + the summary :440-445), executed
                                     line for line (compile() of those source lines) over a stream of estimated / ground-truth
                                     poses and covariances.

Random draws are recorded (white_gaussian_noise is wrapped; np.random.normal draws are replayed from the same seed), so the
device kernels can be given exactly the noise the reference used."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = ref_shim.REF_ROOT


def fixture_generator():
    ref_shim.load_reference()
    path = os.path.join(REF, "dataset/tools/dataset_generators/synthetic_generator.py")
    # The file cannot be imported here: it ends in a script that writes under the author's home directory, and its
    # parameter dataclasses use ndarray / dataclass-instance defaults, which Python >= 3.11 rejects.  The three classes that
    # hold the algorithm (TrajectorySegment, MeasurementsGenerator, WorldPointsGenerator) are compiled from the reference's
    # source unchanged (ast); the six plain data holders they fill are re-declared below with the same field names.
    import ast
    from dataclasses import dataclass, field
    from typing import List
    tree = ast.parse(open(path).read(), path)
    keep = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ("TrajectorySegment", "MeasurementsGenerator", "WorldPointsGenerator")]
    assert len(keep) == 3
    import scipy.interpolate
    import scipy.spatial.transform
    from src.msckf.Camera import Camera
    from src.utils import geometry as G
    ns = dict(vars(G))
    ns.update(np=np, List=List, Camera=Camera, CubicSpline=scipy.interpolate.CubicSpline, Slerp=scipy.spatial.transform.Slerp,
              __name__="synthetic_generator_definitions")

    @dataclass
    class CameraParameters:
        K: object = None
        T_W_C: object = None
        width: int = 640
        height: int = 480
        sigma_noise: float = 0

    @dataclass
    class CameraMeasurement:
        positions: list = field(default_factory=list)
        descriptors: list = field(default_factory=list)
        scores: list = field(default_factory=list)
        indices: list = field(default_factory=list)

    @dataclass
    class IMUParameters:
        sigma_noise_angular_velocity: float = 0
        sigma_noise_linear_acceleration: float = 0
        sigma_noise_bias_gyroscope: float = 0
        sigma_noise_bias_accelerometer: float = 0
        T_W_I: object = None

    @dataclass
    class IMUMeasurement:
        angular_velocity: object = None
        linear_acceleration: object = None

    @dataclass
    class WorldPoints:
        positions: list = field(default_factory=list)
        indices: list = field(default_factory=list)
        descriptors: list = field(default_factory=list)

    @dataclass
    class Frame:
        timestamp: float = 0
        pose: object = None
        IMU_measurement_gt: object = None
        IMU_measurement_noisy: object = None
        camera_measurement_gt: object = None
        camera_measurement_noisy: object = None

    ns.update(CameraParameters=CameraParameters, CameraMeasurement=CameraMeasurement, IMUParameters=IMUParameters,
              IMUMeasurement=IMUMeasurement, WorldPoints=WorldPoints, Frame=Frame)
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    draws = []
    orig = ns["white_gaussian_noise"]

    def recorder(sigma):
        noise, cov = orig(sigma)
        draws.append(np.array(noise, float))
        return noise, cov
    ns["white_gaussian_noise"] = recorder

    np.random.seed(7)
    Iso = ns["Isometry3D"]
    cam = ns["CameraParameters"](K=np.array([[320.0, 0, 320], [0, 320.0, 240], [0, 0, 1]]),
                                 T_W_C=Iso(np.array([[0.0, 0, 1], [-1, 0, 0], [0, -1, 0]]), np.array([0.0, 0, 0])),
                                 width=640, height=480, sigma_noise=0.05)
    imu = ns["IMUParameters"](T_W_I=Iso(np.eye(3), np.zeros(3)), sigma_noise_linear_acceleration=1e-3,
                              sigma_noise_angular_velocity=1e-4, sigma_noise_bias_accelerometer=1e-4,
                              sigma_noise_bias_gyroscope=1e-5)
    wpg = ns["WorldPointsGenerator"]()
    wpg.add_random_world_points(80, 12, 12, 5, np.array([-6.0, -4, -1]))
    rate = 40.0
    segs = [ns["TrajectorySegment"](np.array([[0.0, 0, 0], [1.4142, 0.5858, 0.2], [2, 2, 0.3]]), np.array([[0.0, 0, 0], [0.05, -0.03, np.pi / 2]])),
            ns["TrajectorySegment"](np.array([[2.0, 2, 0.3], [1.4142, 3.4142, 0.1], [0, 4, 0]]), np.array([[0.05, -0.03, np.pi / 2], [0.0, 0.02, np.pi]]))]
    gen = ns["MeasurementsGenerator"](segs, imu, cam, wpg.world_points, rate)
    gen.generate()
    frames = gen.frames
    T = len(frames)
    # the generator computed len(poses) = T + 1 camera measurements but keeps T frames (synthetic_generator.py:225-229)
    poses = np.array([np.concatenate([f.pose.R.flatten(), f.pose.t]) for f in frames])
    imu_gt = np.array([np.concatenate([f.IMU_measurement_gt.angular_velocity, f.IMU_measurement_gt.linear_acceleration]) for f in frames])
    imu_noisy = np.array([np.concatenate([f.IMU_measurement_noisy.angular_velocity, f.IMU_measurement_noisy.linear_acceleration]) for f in frames])
    # draws: per IMU step 4 vectors of 3 (gyro white, acc white, gyro walk, acc walk: synthetic_generator.py:168-181), then per
    # pose one vector of 2 per visible landmark (:212)
    n_imu = T            # len(poses) - 1 IMU steps; frames = len(imu_measurements_noisy)
    imu_draws = np.array(draws[:4 * n_imu]).reshape(n_imu, 4, 3)
    cam_draws = draws[4 * n_imu:]
    L = len(wpg.world_points.positions)
    vis = np.zeros((T, L), bool)
    kp_gt = np.zeros((T, L, 2))
    kp_noisy = np.zeros((T, L, 2))
    noise_uv = np.zeros((T, L, 2))
    score = np.zeros((T, L))
    k = 0
    for i, f in enumerate(frames):
        ids = f.camera_measurement_gt.indices
        for j, lid in enumerate(ids):
            vis[i, lid] = True
            # (the reference's `Im_p += noise` acts in place on the array it had already appended to the gt list, so its "gt"
            #  positions carry the noise too: the clean projection is noisy - noise)
            kp_noisy[i, lid] = f.camera_measurement_noisy.positions[j]
            noise_uv[i, lid] = cam_draws[k]
            kp_gt[i, lid] = kp_noisy[i, lid] - cam_draws[k]
            score[i, lid] = f.camera_measurement_noisy.scores[j]
            k += 1
    # pose T (the last one) also consumed draws; they are not part of any frame
    np.savez_compressed(os.path.join(GOLDEN, "synth_generator.npz"), poses=poses, imu_gt=imu_gt, imu_noisy=imu_noisy,
                        imu_draws=imu_draws, delta_t=1.0 / rate, K=cam.K, R_W_C=cam.T_W_C.R, t_W_C=cam.T_W_C.t, width=640, height=480,
                        sigma_px=cam.sigma_noise, landmarks=np.array(wpg.world_points.positions), vis=vis, kp_gt=kp_gt,
                        kp_noisy=kp_noisy, noise_uv=noise_uv, score=score, gravity=np.array([0.0, 0, -9.81]),
                        last_pose=np.concatenate([segs[-1].poses[-1].R.flatten(), segs[-1].poses[-1].t]))
    print(f"synth_generator.npz: {T} frames, {L} landmarks, {int(vis.sum())} projections, {k} of {len(cam_draws)} camera draws used")


def fixture_photoreal():
    import pandas as pd
    ref_shim.load_reference()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            seq = "data/tartanair/golden"
            os.makedirs(seq + "/cam")
            K = 14
            for i in range(K):
                open(f"{seq}/cam/{i:06d}.png", "w").close()
            R_W_C = np.array([[0.0, 0, 1], [-1, 0, 0], [0, -1, 0]])
            pd.DataFrame([{f"R{r}{c}": R_W_C[r, c] for r in range(3) for c in range(3)}]).to_csv("data/tartanair/camera_info.csv", index=False)
            from scipy.spatial.transform import Rotation
            rng = np.random.default_rng(11)
            t = np.arange(K) * 0.05
            pos = np.stack([0.6 * np.sin(1.3 * t) + 0.2 * t, 0.4 * (1 - np.cos(0.9 * t)), 0.1 * np.sin(2.1 * t)], 1) + 1.5
            eul = np.stack([0.2 * np.sin(0.8 * t), 0.1 * np.sin(1.1 * t + 0.3), 0.6 * t], 1)
            quat = Rotation.from_euler("xyz", eul).as_quat()      # x y z w
            quat[5] = -quat[5]                                    # same rotation, opposite sign: the (q, -q) branch of :180-182
            # the generator names the columns itself (header row consumed): first row is a dummy header line
            with open(f"{seq}/trajectory.csv", "w") as f:
                f.write("x y z qx qy qz qw\n")
                for p, q in zip(pos, quat):
                    f.write(" ".join(f"{v:.17g}" for v in np.concatenate([p, q])) + "\n")
            sys.path.insert(0, REF)
            import importlib
            mod = importlib.import_module("dataset.tools.dataset_generators.photorealistic_generator")
            sig = dict(accelerometer_noise_density=5e-3, gyroscope_noise_density=5e-4, accelerometer_random_walk=5e-4,
                       gyroscope_random_walk=5e-5)
            gen = mod.PhotorealisticGenerator("tartanair", "golden", **sig)
            np.random.seed(123)
            gen.process_data()
            tr = pd.read_csv(f"{seq}/transforms_gt.csv").to_numpy()
            imu_gt = pd.read_csv(f"{seq}/imu_gt.csv").to_numpy()
            imu_noisy = pd.read_csv(f"{seq}/imu.csv").to_numpy()
        finally:
            os.chdir(cwd)
    n = len(imu_gt) - 1
    np.random.seed(123)
    normals = np.array([[np.random.normal(0, 1, 3) for _ in range(4)] for _ in range(n)])    # acc white, acc walk, gyro white, gyro walk
    # the key poses the generator differentiated: 3 identity rows at 0, 0.05, 0.1 then the trajectory rows re-based (:76-113)
    T_W_C = np.eye(4)
    T_W_C[:3, :3] = R_W_C
    def T_of(p, q):
        T = np.eye(4)
        T[:3, :3] = Rotation.from_quat(q).as_matrix()
        T[:3, 3] = p
        return T
    T0 = T_of(pos[0], quat[0]) @ np.linalg.inv(T_W_C)
    keys = [np.eye(4)] * 3 + [np.linalg.inv(T0) @ (T_of(pos[i], quat[i]) @ np.linalg.inv(T_W_C)) for i in range(1, K)]
    key_times = [0.0, 0.05, 0.1] + [i * 0.05 + 0.1 for i in range(1, K)]
    key_poses = np.array([np.concatenate([T[:3, :3].flatten(), T[:3, 3]]) for T in keys])
    np.savez_compressed(os.path.join(GOLDEN, "synth_photoreal.npz"), key_poses=key_poses, key_times=np.array(key_times),
                        transforms=tr, imu_gt=imu_gt, imu_noisy=imu_noisy, normals=normals, gravity=np.array([0.0, 0, -9.81]),
                        sigmas=np.array([sig["accelerometer_noise_density"], sig["gyroscope_noise_density"],
                                         sig["accelerometer_random_walk"], sig["gyroscope_random_walk"]]))
    print(f"synth_photoreal.npz: {len(keys)} key poses -> {len(imu_gt)} IMU rows")


def fixture_metrics():
    """main.py:312-391 executed from the reference's own source text over a synthetic stream of poses."""
    ns_ref = ref_shim.load_reference()
    import pandas as pd
    from collections import deque
    lines = open(os.path.join(REF, "main.py")).read().split("\n")
    # 1-based inclusive line ranges of the reference driver: ground truth + errors + stores; the update of the previous poses
    block = "\n".join(lines[311:391]) + "\n" + "\n".join(lines[435:437]) + "\n"
    import textwrap
    code = compile(textwrap.dedent(block), "main.py:312-391,436-437", "exec")
    summary = compile(textwrap.dedent("\n".join(lines[439:445])), "main.py:440-445", "exec")
    from src.utils import geometry as G
    rng = np.random.default_rng(5)
    n = 400
    t = np.arange(n) * 0.005
    from scipy.spatial.transform import Rotation
    gt_p = np.stack([np.sin(0.7 * t) * 2, 1 - np.cos(0.5 * t), 0.2 * t], 1)
    gt_p[:30] = 0.0                                         # stationary start: relative motion exactly zero (the 0-branches of :349-350)
    gt_R = Rotation.from_euler("xyz", np.stack([0.1 * np.sin(t), 0.05 * t, 0.8 * t], 1)).as_matrix()
    gt_R[:30] = np.eye(3)
    drift = np.cumsum(1e-4 * rng.standard_normal((n, 3)), 0)
    est_p = gt_p + drift
    est_p[200] += 0.05                                      # a spike: exercises the 10x smoothing branch (:352-358)
    rot_err = Rotation.from_rotvec(np.cumsum(2e-5 * rng.standard_normal((n, 3)), 0)).as_matrix()
    est_R = np.einsum("nij,njk->nik", gt_R, rot_err)
    P = np.zeros((n, 21, 21))
    for i in range(n):
        A = rng.standard_normal((21, 21)) * 1e-3
        P[i] = A @ A.T + np.eye(21) * (1e-6 + 1e-8 * i)
    cols = ["T00", "T01", "T02", "T03", "T10", "T11", "T12", "T13", "T20", "T21", "T22", "T23"]
    rows = []
    for i in range(n):
        M = np.hstack([gt_R[i], gt_p[i][:, None]])
        rows.append(dict(zip(cols, M.flatten())))
    gt_trajectory = pd.DataFrame(rows)

    class _S:
        pass
    msckf = _S()
    msckf.state = _S()
    msckf.state.imu = _S()
    g = dict(vars(G))
    g.update(np=np, gt_trajectory=gt_trajectory, msckf=msckf, Isometry3D=G.Isometry3D, R2euler=G.R2euler,
             T_W_I0_gt=G.Isometry3D(np.eye(3), np.zeros(3)), T_W_I0_est=G.Isometry3D(np.eye(3), np.zeros(3)),
             relative_position_error_deque=deque(maxlen=10), relative_orientation_error_deque=deque(maxlen=10))
    for name in ("gt_positions", "gt_rotations", "estimated_positions", "estimated_rotations", "rte_array", "roe_array", "ate_x",
                 "ate_x_bounds", "ate_y", "ate_y_bounds", "ate_z", "ate_z_bounds", "aoe_roll", "aoe_roll_bounds", "aoe_pitch",
                 "aoe_pitch_bounds", "aoe_yaw", "aoe_yaw_bounds", "relative_RMSE_position", "relative_RMSE_orientation"):
        g[name] = []
    for i in range(n):
        msckf.state.imu.T_W_Ii = G.Isometry3D(est_R[i].copy(), est_p[i].copy())
        msckf.state.covariance = P[i]
        g["i"] = i
        exec(code, g)
    exec(summary, g)
    out = dict(gt_R=gt_R, gt_p=gt_p, est_R=est_R, est_p=est_p, P=P,
               rte=np.array(g["rte_array"]), roe=np.array(g["roe_array"]),
               ate=np.stack([g["ate_x"], g["ate_y"], g["ate_z"]], 1), aoe=np.stack([g["aoe_roll"], g["aoe_pitch"], g["aoe_yaw"]], 1),
               bounds=np.stack([g["ate_x_bounds"], g["ate_y_bounds"], g["ate_z_bounds"], g["aoe_roll_bounds"], g["aoe_pitch_bounds"],
                                g["aoe_yaw_bounds"]], 1),
               rel_sq_pos=np.array(g["relative_RMSE_position"]), rel_sq_ori=np.array(g["relative_RMSE_orientation"]),
               summary=np.array([g["rmse_position"], g["rmse_orientation"], g["mean_rte"], g["std_rte"], g["mean_roe"], g["std_roe"]]))
    np.savez_compressed(os.path.join(GOLDEN, "metrics_main.npz"), **out)
    print("metrics_main.npz:", n, "steps; summary", out["summary"])


if __name__ == "__main__":
    which = sys.argv[1:] or ["generator", "photoreal", "metrics"]
    if "generator" in which:
        fixture_generator()
    if "photoreal" in which:
        fixture_photoreal()
    if "metrics" in which:
        fixture_metrics()
