"""Drop-in host side of the accelerated path: the reference's filter-class API over the sm_100a kernels.

The reference (ValerioSpagnoli/Monocular-Visual-Inertial-MSCKF) has no plugin/FFI layer; its boundary for this path is the
method set of `class MSCKF` in src/msckf/MSCKF.py.  `B200UpdateMixin` re-implements exactly the hot methods with the
reference's names, arguments, return values and side effects:

    process_imu(imu_measurement)              MSCKF.py:160-248   (+ IMU.integrate, IMU.py:78-100)
    get_valid_features(features)              MSCKF.py:458-495   -> (valid_features, lost_features), re-triangulates in place
    update(features)                          MSCKF.py:570-609   (+ correct, 611-661): mutates state.covariance, the IMU
                                                                  state, every camera pose, the gating discard counter

It is duck-typed: it reads/writes only attributes the reference's own objects have (state.imu.T_W_Ii.R/.t, v_W_Ii,
gyroscope_bias, accelerometer_bias, T_W_Ii_null, v_W_Ii_null, timestamp, W_gravity; state.cameras {id: cam.T_W_Ci};
state.covariance; Feature.keypoints/lines/camera_indices/inverse_depth_point/...), so a maintainer of the reference
mixes it into the reference class (`class MSCKF(B200UpdateMixin, reference.MSCKF)`), while this
package's own `MSCKF` below combines it with minimal mirrors of the reference's data holders so that the path runs (and
is tested) where the reference is not installed.  Aliasing is preserved the way the reference produces it: positions are
corrected in place (`t[:] = ...`), so `Line.base` / `InverseDepthPoint.base` arrays that alias a camera position follow
the correction (MSCKF.py:305,410,637-661).

There is no CPU fallback: every method ends in a C-ABI call into csrc/libmsckf_b200.so.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import _lib
from .batched import BatchedMSCKF, UpdateParams


# ---------------------------------------------------------------------------------------------------------------------
# minimal mirrors of the reference's data holders (same attribute names; only what the hot path touches)
# ---------------------------------------------------------------------------------------------------------------------
class Isometry3D:
    """reference src/utils/geometry.py:23-51 (R 3x3, t 3)"""

    def __init__(self, R=None, t=None):
        self.R = np.eye(3) if R is None else np.asarray(R, float)
        self.t = (np.zeros(3) if t is None else np.asarray(t, float)).reshape((1, 3))[0]

    def __mul__(self, other):
        return Isometry3D(self.R @ other.R, self.R @ other.t + self.t)

    def inv(self):
        return Isometry3D(self.R.T, -self.R.T @ self.t)

    def transform(self, p, rotation_only=False):
        return self.R @ p if rotation_only else self.R @ p + self.t


@dataclass
class Line:
    """reference src/utils/geometry.py:5-21; `base` aliases the observing camera's live position array"""
    base: np.ndarray = field(default_factory=lambda: np.zeros(3))
    direction: np.ndarray = field(default_factory=lambda: np.zeros(3))
    confidence: float = 1.0


class InverseDepthPoint:
    """reference src/utils/geometry.py:53-71; `base` aliases the anchor camera's position array"""

    def __init__(self, camera_pose=None, direction=None):
        camera_pose = camera_pose if camera_pose is not None else Isometry3D()
        direction = np.array([0.0, 0.0, 1.0]) if direction is None else np.asarray(direction, float)
        self.base = camera_pose.t
        self.m = direction / max(np.linalg.norm(direction), 1e-300)
        self.rho = 0.1


@dataclass
class Feature:
    """reference src/msckf/FeatureExtractor.py:20-31"""
    keypoints: List[np.ndarray] = field(default_factory=list)
    descriptors: List[np.ndarray] = field(default_factory=list)
    scores: List[float] = field(default_factory=list)
    camera_indices: List[int] = field(default_factory=list)
    lines: List[Line] = field(default_factory=list)
    inverse_depth_point: InverseDepthPoint = field(default_factory=InverseDepthPoint)
    tracked_for_n_frames: int = 0
    lost_for_n_frames: int = 0


class Camera:
    """reference src/msckf/Camera.py:6-11: the null pose IS the pose object"""

    def __init__(self, K, width, height, T_W_Ci):
        self.K, self.width, self.height = K, width, height
        self.T_W_Ci = T_W_Ci
        self.T_W_Ci_null = T_W_Ci


@dataclass
class IMUMeasurement:
    """reference src/msckf/IMU.py:6-11"""
    timestamp: float = 0.0
    angular_velocity: np.ndarray = field(default_factory=lambda: np.zeros(3))
    linear_acceleration: np.ndarray = field(default_factory=lambda: np.zeros(3))


class IMU:
    """reference src/msckf/IMU.py:13-41 (state fields only)"""

    def __init__(self, W_gravity=None, id=0, timestamp=0.0):
        self.id, self.timestamp = id, timestamp
        self.T_W_I = Isometry3D()
        self.T_W_Ii = Isometry3D()
        self.v_W_Ii = np.zeros(3)
        self.gyroscope_bias = np.zeros(3)
        self.accelerometer_bias = np.zeros(3)
        self.W_gravity = np.array([0.0, 0.0, -9.81]) if W_gravity is None else np.asarray(W_gravity, float)
        self.T_W_Ii_null = Isometry3D()
        self.v_W_Ii_null = np.zeros(3)
        self.is_initialized = True


@dataclass
class MSCKFState:
    """reference src/msckf/MSCKF.py:72-77"""
    imu: IMU = None
    cameras: Dict[int, Camera] = field(default_factory=dict)
    covariance: np.ndarray = field(default_factory=lambda: np.zeros((15, 15)))


@dataclass
class MSCKFParameters:
    """the MSCKFParameters fields this path reads (reference src/msckf/MSCKF.py:14-46, same names and defaults)"""
    T_W_C: Isometry3D = field(default_factory=lambda: Isometry3D(np.array([[0.0, 0, 1], [-1, 0, 0], [0, -1, 0]]), np.zeros(3)))
    K: np.ndarray = field(default_factory=lambda: np.array([[180.0, 0, 320], [0, 180.0, 240], [0, 0, 1]]))
    width: int = 640
    height: int = 480
    sigma_image: float = 0.2
    accelerometer_noise_density: float = 0.001
    accelerometer_random_walk: float = 0.00001
    gyroscope_noise_density: float = 0.0001
    gyroscope_random_walk: float = 0.000001
    W_gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.81]))
    use_parallax: bool = True
    min_parallax: float = 20
    min_number_of_frames_to_be_lost: int = 1
    min_number_of_frames_to_be_tracked: int = 5
    max_number_of_camera_states: int = 30


# ---------------------------------------------------------------------------------------------------------------------
# the accelerated methods
# ---------------------------------------------------------------------------------------------------------------------
def _round_up(x, m):
    return ((max(int(x), 1) + m - 1) // m) * m


class B200UpdateMixin:
    """process_imu / get_valid_features / update of the reference's MSCKF, executed by the CUDA library."""

    compress_mode = _lib.COMPRESS_GRAM   # or _lib.COMPRESS_QR: explicit Householder TSQR like np.linalg.qr
    gate_mode = _lib.GATE_BOUNDS         # the reference's gating_test only returns the decision; GATE_EXACT also works
    device = "cuda"

    # -- device context (capacities grow in steps so that consecutive frames reuse the buffers) --------------------
    def _b200_context(self, F, M, N) -> BatchedMSCKF:
        ctx = getattr(self, "_b200_ctx", None)
        if ctx is None or ctx.F < F or ctx.M < M or ctx.N < N:
            params = UpdateParams(K=np.asarray(self.K, float), width=int(self.width), height=int(self.height),
                                  sigma_image=float(self.sigma_image), gravity=np.asarray(self.state.imu.W_gravity, float),
                                  min_lost=int(self.min_number_of_frames_to_be_lost),
                                  min_tracked=int(self.min_number_of_frames_to_be_tracked),
                                  use_parallax=bool(self.use_parallax), min_parallax=float(self.min_parallax),
                                  compress_mode=self.compress_mode, gate_mode=self.gate_mode)
            Fc, Mc, Nc = _round_up(F, 64), _round_up(M, 8), max(int(N), 1)   # N exact: the warp-per-feature kernel needs N <= 31
            if ctx is not None:
                Fc, Mc, Nc = max(Fc, ctx.F), max(Mc, ctx.M), max(Nc, ctx.N)
            ctx = BatchedMSCKF(1, Fc, Mc, Nc, params, device=self.device)
            self._b200_ctx = ctx
        return ctx

    def _upload_state(self, ctx):
        imu, P = self.state.imu, np.asarray(self.state.covariance, float)
        cams = list(self.state.cameras.values())
        n, d = len(cams), P.shape[0]
        assert d == 15 + 6 * n, "covariance size does not match the number of camera states"
        h = np.zeros((ctx.ldp, ctx.ldp))
        h[:d, :d] = P
        ctx.P[0].copy_(torch.from_numpy(h))
        st = np.concatenate([np.asarray(imu.T_W_Ii.R, float).flatten(), imu.T_W_Ii.t, imu.v_W_Ii, imu.gyroscope_bias,
                             imu.accelerometer_bias]).astype(float)
        ctx.imu_state[0].copy_(torch.from_numpy(st))
        R = np.zeros((ctx.N, 9))
        t = np.zeros((ctx.N, 3))
        for s, cam in enumerate(cams):
            R[s] = np.asarray(cam.T_W_Ci.R, float).flatten()
            t[s] = cam.T_W_Ci.t
        ctx.cam_R[0].copy_(torch.from_numpy(R))
        ctx.cam_t[0].copy_(torch.from_numpy(t))
        ctx.n_cams.fill_(n)
        return n, d

    def _pack_features(self, ctx, features):
        cam_ids = list(self.state.cameras.keys())
        slot_of = {cid: s for s, cid in enumerate(cam_ids)}
        cam_t = [self.state.cameras[c].T_W_Ci.t for c in cam_ids]
        F, M = ctx.F, ctx.M
        h = dict(feat_len=np.zeros(F, np.int32), obs_uv=np.zeros((F, M, 2)), obs_dir=np.zeros((F, M, 3)),
                 obs_conf=np.zeros((F, M)), obs_slot=np.zeros((F, M), np.int32), idp_base=np.zeros((F, 3)),
                 idp_anchor_slot=-np.ones(F, np.int32), idp_m=np.zeros((F, 3)), idp_rho=np.zeros(F),
                 lost_cnt=np.zeros(F, np.int32), tracked_cnt=np.zeros(F, np.int32))
        for j, ft in enumerate(features.values()):
            n = len(ft.keypoints)
            h["feat_len"][j] = n
            if n:
                h["obs_uv"][j, :n] = np.asarray(ft.keypoints, float)
                h["obs_dir"][j, :n] = np.stack([ln.direction for ln in ft.lines])
                h["obs_conf"][j, :n] = [ln.confidence for ln in ft.lines]
                h["obs_slot"][j, :n] = [slot_of[c] for c in ft.camera_indices]   # position in the window (MSCKF.py:539)
            idp = ft.inverse_depth_point
            h["idp_base"][j] = idp.base
            for s, tarr in enumerate(cam_t):          # InverseDepthPoint.base aliases the anchor camera's live position
                if idp.base is tarr:
                    h["idp_anchor_slot"][j] = s
                    break
            h["idp_m"][j] = idp.m
            h["idp_rho"][j] = idp.rho
            h["lost_cnt"][j] = ft.lost_for_n_frames
            h["tracked_cnt"][j] = ft.tracked_for_n_frames
        for k, v in h.items():
            getattr(ctx, k)[0].copy_(torch.from_numpy(v))

    def _download_idp(self, ctx, features):
        m = ctx.idp_m[0].cpu().numpy()
        rho = ctx.idp_rho[0].cpu().numpy()
        for j, ft in enumerate(features.values()):
            ft.inverse_depth_point.m = m[j].copy()
            ft.inverse_depth_point.rho = float(rho[j])

    # -- reference API ------------------------------------------------------------------------------------------------
    def get_valid_features(self, features: Dict[int, "Feature"]) -> Tuple[Dict[int, "Feature"], Dict[int, "Feature"]]:
        """MSCKF.get_valid_features (MSCKF.py:458-495): selection by lost / parallax, re-triangulation of the selected
        tracks (their inverse_depth_point is updated in place); returns (valid_features, lost_features)."""
        valid, lost = {}, {}
        if len(features) == 0:
            return valid, lost
        M = max([len(ft.keypoints) for ft in features.values()] + [1])
        ctx = self._b200_context(len(features), M, len(self.state.cameras))
        self._upload_state(ctx)
        self._pack_features(ctx, features)
        ctx.front()
        st = ctx.feat_status[0].cpu().numpy()
        self._download_idp(ctx, features)
        for j, (idx, ft) in enumerate(features.items()):
            code = st[j] & 0xF
            if code in (_lib.FEAT_ACCEPTED, _lib.FEAT_GATED_OUT):
                valid[idx] = ft
            if st[j] & _lib.FEAT_LOST_FLAG:
                lost[idx] = ft
        return valid, lost

    def update(self, features: Dict[int, "Feature"]) -> None:
        """MSCKF.update + MSCKF.correct (MSCKF.py:570-661) on `features` (the valid_features dict the reference passes).
        Silent return when nothing survives the gate (MSCKF.py:584-585)."""
        if len(features) == 0:
            return
        M = max([len(ft.keypoints) for ft in features.values()] + [1])
        ctx = self._b200_context(len(features), M, len(self.state.cameras))
        n, d = self._upload_state(ctx)
        self._pack_features(ctx, features)
        ctx.update()
        n_rej = int(ctx.n_rejected[0])
        self.number_of_residuals_discarded_for_gasting_test = getattr(self, "number_of_residuals_discarded_for_gasting_test", 0) + n_rej
        self.last_gate_status = ctx.feat_status[0, :len(features)].cpu().numpy() & 0xF
        self.last_gamma = ctx.gamma[0, :len(features)].cpu().numpy()
        self.last_delta_x = ctx.delta_x[0, :d].cpu().numpy()
        self._download_idp(ctx, features)     # the fused call re-triangulates (same values get_valid_features produced)
        if int(ctx.n_accepted[0]) == 0:
            return
        # ---- results back into the reference's objects (in place where the reference is in place) ----
        self.state.covariance = ctx.P[0, :d, :d].cpu().numpy()
        st = ctx.imu_state[0].cpu().numpy()
        imu = self.state.imu
        imu.T_W_Ii.R = st[0:9].reshape(3, 3).copy()
        imu.T_W_Ii.t[:] = st[9:12]
        imu.v_W_Ii[:] = st[12:15]
        imu.gyroscope_bias[:] = st[15:18]
        imu.accelerometer_bias[:] = st[18:21]
        cam_R = ctx.cam_R[0, :n].cpu().numpy()
        cam_t = ctx.cam_t[0, :n].cpu().numpy()
        for s, cam in enumerate(self.state.cameras.values()):
            cam.T_W_Ci.R = cam_R[s].reshape(3, 3).copy()
            cam.T_W_Ci.t[:] = cam_t[s]

    def process_imu(self, imu_measurement) -> None:
        """MSCKF.process_imu (MSCKF.py:160-248): integrates the IMU state and propagates the covariance."""
        imu = self.state.imu
        ctx = self._b200_context(1, 2, max(len(self.state.cameras), 1))
        n, d = self._upload_state(ctx)
        null = np.concatenate([np.asarray(imu.T_W_Ii_null.R, float).flatten(), imu.T_W_Ii_null.t, imu.v_W_Ii_null])
        ctx.imu_null[0].copy_(torch.from_numpy(null.astype(float)))
        ctx.imu_time.fill_(float(imu.timestamp))
        q = np.diag(np.asarray(self.continuous_noise_covariance, float))[[0, 3, 6, 9]].copy()
        ctx.Qc_diag[0].copy_(torch.from_numpy(q))
        sample = np.concatenate([[imu_measurement.timestamp], imu_measurement.angular_velocity,
                                 imu_measurement.linear_acceleration]).astype(float)
        ctx.propagate(torch.from_numpy(sample).reshape(1, 1, 7).to(ctx.device), use_null_state=True)
        st = ctx.imu_state[0].cpu().numpy()
        # IMU.integrate creates NEW pose / velocity objects; the null state then aliases them (MSCKF.py:247-248)
        imu.T_W_Ii = Isometry3D(st[0:9].reshape(3, 3).copy(), st[9:12].copy())
        imu.v_W_Ii = st[12:15].copy()
        imu.T_W_Ii_null = imu.T_W_Ii
        imu.v_W_Ii_null = imu.v_W_Ii
        imu.timestamp = float(imu_measurement.timestamp)
        self.state.covariance = ctx.P[0, :d, :d].cpu().numpy()


    # -- window surgery: same names / arguments / side effects as the reference ---------
    def state_augmentation(self) -> None:
        """MSCKF.state_augmentation (MSCKF.py:250-265): the current camera pose enters the window under the key
        `state.imu.id`, P <- [I; J] P [I; J]^T on the device (msckf_augment_batched)."""
        imu = self.state.imu
        T_I_C = imu.T_W_I.inv() * self.T_W_C
        n = len(self.state.cameras)
        ctx = self._b200_context(1, 2, n + 1)
        self._upload_state(ctx)
        ctx.augment(T_I_C.R, T_I_C.t)
        d = 15 + 6 * (n + 1)
        self.state.covariance = ctx.P[0, :d, :d].cpu().numpy()
        R = ctx.cam_R[0, n].cpu().numpy().reshape(3, 3).copy()
        t = ctx.cam_t[0, n].cpu().numpy().copy()
        self.state.cameras[imu.id] = Camera(self.K, self.width, self.height, Isometry3D(R, t))

    def remove_cameras(self, cameras: Dict[int, "Camera"]) -> None:
        """MSCKF.remove_cameras (MSCKF.py:751-770): covariance rows / columns and poses of `cameras` leave the window
        (msckf_remove_cameras_batched), their observations leave every track, empty tracks are deleted."""
        if len(cameras) == 0:
            return
        ids = list(self.state.cameras.keys())
        n = len(ids)
        ctx = self._b200_context(1, 2, n)
        self._upload_state(ctx)
        mask = torch.zeros((1, ctx.N), dtype=torch.uint8)
        for k in cameras.keys():
            mask[0, ids.index(k)] = 1
        ctx.remove_cameras(mask.to(ctx.device))
        d = 15 + 6 * (n - int(mask.sum()))
        self.state.covariance = ctx.P[0, :d, :d].cpu().numpy()
        for k in cameras.keys():
            del self.state.cameras[k]
        emptied = []
        for index, ft in self.features.items():
            for k in cameras.keys():
                if k in ft.camera_indices:
                    i = ft.camera_indices.index(k)
                    for lst in (ft.keypoints, ft.descriptors, ft.scores, ft.camera_indices, ft.lines):
                        if i < len(lst):
                            del lst[i]
            if len(ft.camera_indices) == 0:
                emptied.append(index)
        for index in emptied:
            del self.features[index]

    def prune_poorest_camera_states(self) -> None:
        """MSCKF.prune_poorest_camera_states (MSCKF.py:710-737): the two cameras observed by the fewest tracks get one
        last update with the tracks that see them and are removed."""
        per_camera: Dict[int, int] = {}
        for ft in self.features.values():
            for k in ft.camera_indices:
                per_camera[k] = per_camera.get(k, 0) + 1
        chosen = [k for k, _ in sorted(per_camera.items(), key=lambda kv: kv[1])][:2]
        cams = {k: self.state.cameras[k] for k in chosen}
        touched = {i: ft for i, ft in self.features.items() if any(k in cams for k in ft.camera_indices)}
        valid, _ = self.get_valid_features(touched)
        if len(valid) > 0:
            self.update(valid)
        self.remove_cameras(cams)


class MSCKF(B200UpdateMixin):
    """Stand-alone filter object with the reference's constructor signature and state layout for this path
    (reference MSCKF.__init__, MSCKF.py:80-132; track management / feature extraction are not part of this package)."""

    def __init__(self, parameters: MSCKFParameters, rr=None, device="cuda"):
        self.rr = rr
        self.device = device
        self.state = MSCKFState()
        self.state.imu = IMU(W_gravity=parameters.W_gravity)
        self.K, self.T_W_C = parameters.K, parameters.T_W_C
        self.width, self.height, self.sigma_image = parameters.width, parameters.height, parameters.sigma_image
        q = np.eye(12)
        q[:3, :3] *= parameters.gyroscope_noise_density ** 2
        q[3:6, 3:6] *= parameters.gyroscope_random_walk ** 2
        q[6:9, 6:9] *= parameters.accelerometer_noise_density ** 2
        q[9:12, 9:12] *= parameters.accelerometer_random_walk ** 2
        self.continuous_noise_covariance = q
        self.features: Dict[int, Feature] = {}
        self.use_parallax, self.min_parallax = parameters.use_parallax, parameters.min_parallax
        self.min_number_of_frames_to_be_lost = max(parameters.min_number_of_frames_to_be_lost, 1)         # MSCKF.py:119
        self.min_number_of_frames_to_be_tracked = max(parameters.min_number_of_frames_to_be_tracked, 2)   # MSCKF.py:120
        self.max_number_of_camera_states = parameters.max_number_of_camera_states
        self.number_of_residuals_discarded_for_gasting_test = 0

    def process_features(self):
        """MSCKF.process_features (MSCKF.py:450-456)"""
        valid_features, lost_features = self.get_valid_features(self.features)
        if len(valid_features) > 0:
            self.update(valid_features)
            for index in lost_features:
                self.features.pop(index, None)
