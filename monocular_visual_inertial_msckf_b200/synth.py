"""Synthetic visual-inertial streams (IMU samples + per-frame keypoint measurements) for tests and benchmarks.

This is this repo's own generator (numpy only, no dependency on the reference).  It produces streams in the shape the
reference's driver feeds its filter (reference main.py:261-298: one IMUMeasurement(t, omega, a) per sample, one
CameraMeasurement(keypoints, descriptors, scores) per camera frame), following the conventions of the reference's
data tools so that the reference filter and this repo's filter both run on it unchanged:

  * accelerometer reading a = R_prev^T (dv/dt + W_gravity), W_gravity = (0, 0, -9.81)
    (reference dataset/tools/dataset_generators/synthetic_generator.py:157), matching IMU.integrate's
    "R a - W_gravity" (reference src/msckf/IMU.py:92);
  * gyroscope reading = body rate between consecutive poses; white noise + random-walk bias per sample
    (reference photorealistic_generator.py:189-198 / synthetic_generator.py:168-181);
  * camera extrinsics T_W_C = [[0,0,1],[-1,0,0],[0,-1,0]] (camera looks along body +x), pinhole K,
    keypoint noise N(0, sigma_px), scores in (0.7, 1], unit-norm random descriptors
    (reference synthetic_generator.py:342-343,365-368).

"TartanAir-shaped" (BASELINE.json configs[1..3]): K = fx 320, 640x480, camera 20 Hz, 9 IMU samples per frame.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

NOISE_LEVELS = {
    # (sigma_acc, sigma_gyro, sigma_acc_walk, sigma_gyro_walk) -- reference main.py:139-160
    "low": (1e-3, 1e-4, 1e-4, 1e-5),
    "mid": (5e-3, 5e-4, 5e-4, 5e-5),
    "high": (1e-2, 1e-3, 1e-3, 1e-4),
}

R_BODY_CAM = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
W_GRAVITY = np.array([0.0, 0.0, -9.81])


def _skew(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def _exp_so3(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3) + _skew(w)
    a = w / th
    A = _skew(a)
    return np.eye(3) + np.sin(th) * A + (1.0 - np.cos(th)) * (A @ A)


def _log_so3(R):
    c = np.clip((np.trace(R) - 1.0) * 0.5, -1.0, 1.0)
    th = np.arccos(c)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if th < 1e-9:
        return 0.5 * v
    return th / (2.0 * np.sin(th)) * v


def _euler_zyx(yaw, pitch, roll):
    cy, sy = np.cos(yaw), np.sin(yaw)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cr, sr = np.cos(roll), np.sin(roll)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    Ry = np.array([[cp, 0, sp], [0, 1.0, 0], [-sp, 0, cp]])
    Rx = np.array([[1.0, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return Rz @ Ry @ Rx


@dataclass
class Frame:
    imu_index: int                 # this frame is delivered right after IMU sample `imu_index`
    keypoints: np.ndarray          # (n, 2) float64 pixels
    descriptors: np.ndarray        # (n, k) float64 unit norm
    scores: np.ndarray             # (n,) float64
    landmark_ids: np.ndarray       # (n,) int, ground-truth association (not given to the filters)


@dataclass
class Sequence:
    K: np.ndarray
    width: int
    height: int
    noise_level: str
    imu: np.ndarray                               # (T, 7): timestamp, omega(3), acc(3)
    frames: List[Frame] = field(default_factory=list)
    gt_R: Optional[np.ndarray] = None             # (T+1, 3, 3) body poses at IMU rate
    gt_p: Optional[np.ndarray] = None             # (T+1, 3)
    landmarks: Optional[np.ndarray] = None        # (L, 3)


def make_sequence(seed: int = 42, n_frames: int = 40, n_landmarks: int = 64, imu_per_frame: int = 9,
                  frame_dt: float = 0.05, noise_level: str = "mid", fx: float = 320.0, width: int = 640,
                  height: int = 480, sigma_px: float = 0.1, p_drop: float = 0.0, p_outlier: float = 0.0,
                  outlier_px: float = 60.0, descriptor_dim: int = 32, speed: float = 1.0,
                  depth_range=(6.0, 14.0), shuffle: bool = False) -> Sequence:
    """One independent visual-inertial stream; `seed` follows SURVEY.md section 8d (default_rng(42 + filter index))."""
    rng = np.random.default_rng(seed)
    K = np.array([[fx, 0.0, width / 2.0], [0.0, fx, height / 2.0], [0.0, 0.0, 1.0]])
    dt = frame_dt / imu_per_frame
    T = n_frames * imu_per_frame
    sa, sg, sba, sbg = NOISE_LEVELS[noise_level]

    # ---- smooth bounded 6-DoF trajectory: Lissajous position + small attitude oscillation, ramped in from rest ----
    amp = rng.uniform([0.5, 0.7, 0.3], [0.9, 1.1, 0.5])
    wp = rng.uniform(0.5, 0.9, 3) * speed
    php = rng.uniform(0, 2 * np.pi, 3)
    ang_amp = rng.uniform([0.10, 0.05, 0.03], [0.18, 0.09, 0.06])
    wa = rng.uniform(0.3, 0.7, 3) * speed
    pha = rng.uniform(0, 2 * np.pi, 3)
    t_rest = frame_dt * 1.0
    t_ramp = 1.0

    def ramp(t):
        x = np.clip((t - t_rest) / t_ramp, 0.0, 1.0)
        return x * x * x * (x * (6 * x - 15) + 10)

    ts = np.arange(T + 1) * dt
    gt_p = np.zeros((T + 1, 3))
    gt_R = np.zeros((T + 1, 3, 3))
    for k, t in enumerate(ts):
        s = ramp(t)
        gt_p[k] = s * amp * (np.sin(wp * t + php) - np.sin(php))
        e = s * ang_amp * (np.sin(wa * t + pha) - np.sin(pha))
        gt_R[k] = _euler_zyx(e[0], e[1], e[2])

    # ---- IMU by finite differences (exactly what the reference's Euler integrator inverts) ----
    imu = np.zeros((T, 7))
    bg = np.zeros(3)
    ba = np.zeros(3)
    v_prev = np.zeros(3)
    for k in range(1, T + 1):
        v = (gt_p[k] - gt_p[k - 1]) / dt
        acc = gt_R[k - 1].T @ ((v - v_prev) / dt + W_GRAVITY)
        omega = _log_so3(gt_R[k - 1].T @ gt_R[k]) / dt
        bg = bg + sbg * rng.standard_normal(3)
        ba = ba + sba * rng.standard_normal(3)
        omega = omega + sg * rng.standard_normal(3) + bg
        acc = acc + sa * rng.standard_normal(3) + ba
        imu[k - 1, 0] = ts[k]
        imu[k - 1, 1:4] = omega
        imu[k - 1, 4:7] = acc
        v_prev = v

    # ---- landmarks that stay inside the image for every camera frame ----
    frame_pose_idx = [(f + 1) * imu_per_frame for f in range(n_frames)]      # frame f arrives after that many samples
    frame_pose_idx[0] = imu_per_frame
    cam_R = np.stack([gt_R[i] @ R_BODY_CAM for i in frame_pose_idx])
    cam_t = np.stack([gt_p[i] for i in frame_pose_idx])

    def visible_everywhere(P):
        ok = np.ones(len(P), bool)
        for R, t in zip(cam_R, cam_t):
            c = (P - t) @ R
            z = c[:, 2]
            u = fx * c[:, 0] / z + width / 2.0
            v = fx * c[:, 1] / z + height / 2.0
            ok &= (z > 0.5) & (u > 8) & (u < width - 8) & (v > 8) & (v < height - 8)
        return ok

    pts = np.zeros((0, 3))
    while len(pts) < n_landmarks:
        n = 4 * n_landmarks
        depth = rng.uniform(depth_range[0], depth_range[1], n)
        yy = depth * np.tan(rng.uniform(-0.70, 0.70, n))
        zz = depth * np.tan(rng.uniform(-0.55, 0.55, n))
        cand = np.stack([depth, yy, zz], axis=1)
        cand = cand[visible_everywhere(cand)]
        pts = np.concatenate([pts, cand])[:n_landmarks]
    desc = rng.standard_normal((n_landmarks, descriptor_dim))
    desc /= np.linalg.norm(desc, axis=1, keepdims=True)

    seq = Sequence(K=K, width=width, height=height, noise_level=noise_level, imu=imu, gt_R=gt_R, gt_p=gt_p,
                   landmarks=pts)
    for f in range(n_frames):
        R, t = cam_R[f], cam_t[f]
        c = (pts - t) @ R
        uv = np.stack([fx * c[:, 0] / c[:, 2] + width / 2.0, fx * c[:, 1] / c[:, 2] + height / 2.0], axis=1)
        uv = uv + sigma_px * rng.standard_normal(uv.shape)
        if p_outlier > 0:
            bad = rng.random(n_landmarks) < p_outlier
            uv[bad] += outlier_px * rng.standard_normal((int(bad.sum()), 2))
        keep = np.ones(n_landmarks, bool)
        if p_drop > 0 and f > 0:
            keep = rng.random(n_landmarks) >= p_drop
        keep &= (uv[:, 0] >= 0) & (uv[:, 0] < width) & (uv[:, 1] >= 0) & (uv[:, 1] < height)
        ids = np.nonzero(keep)[0]
        if shuffle:
            ids = rng.permutation(ids)
        scores = rng.uniform(0.7, 1.0, n_landmarks)
        seq.frames.append(Frame(imu_index=frame_pose_idx[f] - 1, keypoints=uv[ids].copy(),
                                descriptors=desc[ids].copy(), scores=scores[ids].copy(), landmark_ids=ids))
    return seq
