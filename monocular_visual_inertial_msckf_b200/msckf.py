"""Drop-in host side of the accelerated path: the reference's filter-class API over the sm_100a kernels.

The reference (ValerioSpagnoli/Monocular-Visual-Inertial-MSCKF) has no plugin/FFI layer; its boundary for this path is the
method set of `class MSCKF` in src/msckf/MSCKF.py.  `B200UpdateMixin` re-implements exactly the hot methods with the
reference's names, arguments, return values and side effects:

    process_imu(imu_measurement)              MSCKF.py:160-248   (+ IMU.integrate, IMU.py:78-100; imu.id += 1, :164)
    get_valid_features(features)              MSCKF.py:458-495   -> (valid_features, lost_features), re-triangulates in place
    update(features)                          MSCKF.py:570-609   (+ correct, 611-661): mutates state.covariance, the IMU
                                                                  state, every camera pose, the gating discard counter
    state_augmentation()                      MSCKF.py:250-265   new camera under the key state.imu.id
    remove_cameras(cameras)                   MSCKF.py:751-778   window, covariance, tracks, last_camera_measurement
    prune_poorest_camera_states()             MSCKF.py:710-737

The covariance stays RESIDENT on the device between calls: `state.covariance` becomes a property of the filter's state
object that downloads it only when somebody reads it (the reference's driver reads two 3x3 blocks per frame,
main.py:334-335) and uploads it only after somebody assigned a new matrix.  The small state (IMU pose / velocity / biases,
null state, window poses) travels with every call in one copy each way, because the reference's own Python code reads
and writes those objects between the accelerated calls.

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

import sys
import weakref

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
    """reference src/msckf/Camera.py:6-52: the null pose IS the pose object; the pinhole helpers are the ones the
    reference's add_camera_measurements / get_valid_features call on a window camera"""

    def __init__(self, K, width, height, T_W_Ci):
        self.K, self.width, self.height = K, width, height
        self.T_W_Ci = T_W_Ci
        self.T_W_Ci_null = T_W_Ci

    def project_point(self, Ci_p):
        """Camera.py:13-28: (inside the image and in front of the camera?, pixel)"""
        if Ci_p[2] <= 0:
            return False, None
        h = self.K @ Ci_p
        px = h[:2] / h[2]
        if px[0] < 0 or px[0] >= self.width or px[1] < 0 or px[1] >= self.height:
            return False, None
        return True, px

    def inverse_project_point(self, Im_p):
        """Camera.py:30-37: K^-1 [u, v, 1]"""
        return np.linalg.inv(self.K) @ np.append(Im_p, 1)

    def Ci2W(self, Ci_n, is_versor=False):
        """Camera.py:39-45"""
        return self.T_W_Ci.transform(Ci_n, rotation_only=bool(is_versor))

    def W2Ci(self, W_n, is_versor=False):
        """Camera.py:47-52"""
        return self.T_W_Ci.inv().transform(W_n, rotation_only=bool(is_versor))


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


class _Resident:
    """What the drop-in keeps on the device for ONE filter between calls."""

    def __init__(self):
        self.ctx = None          # BatchedMSCKF(1, F, M, N)
        self.dev_valid = False   # ctx.P[0, :d, :d] is the filter's current covariance
        self.d = 15              # its dimension
        self.n_dev = -1          # value last written to ctx.n_cams
        self.p_uploads = 0       # host -> device copies of the covariance so far (1 for a whole run under the binding)


_STATE_CLASSES = {}


def _resident_state_class(base):
    """Subclass of the filter's state class whose `covariance` is a property backed by the device copy."""
    cls = _STATE_CLASSES.get(base)
    if cls is None:
        def _get(st):
            d = st.__dict__
            if d.get("_b200_cov") is None:
                owner = d["_b200_owner"]()
                d["_b200_cov"] = owner._b200_download_covariance()
            return d["_b200_cov"]

        def _set(st, value):
            d = st.__dict__
            d["_b200_cov"] = value
            owner = d["_b200_owner"]()
            if owner is not None:
                owner._b200_res.dev_valid = False      # the host matrix is authoritative until the next device call

        cls = type(base.__name__, (base,), {"covariance": property(_get, _set), "_b200_resident": True})
        _STATE_CLASSES[base] = cls
    return cls


class B200UpdateMixin:
    """process_imu / get_valid_features / update / window surgery of the reference's MSCKF, executed by the CUDA library."""

    compress_mode = _lib.COMPRESS_GRAM   # or _lib.COMPRESS_QR: explicit Householder TSQR like np.linalg.qr
    gate_mode = _lib.GATE_BOUNDS         # the reference's gating_test only returns the decision; GATE_EXACT also works
    device = "cuda"

    # -- device context (capacities grow in steps so that consecutive frames reuse the buffers) --------------------
    def _b200_make_context(self, F, M, N, params):
        return BatchedMSCKF(1, F, M, N, params, device=self.device)

    @property
    def _b200_res(self) -> _Resident:
        r = self.__dict__.get("_b200_resident")
        if r is None:
            r = self.__dict__["_b200_resident"] = _Resident()
        return r

    def _b200_context(self, F, M, N):
        r = self._b200_res
        ctx = r.ctx
        if ctx is None or ctx.F < F or ctx.M < M or ctx.N < N:
            params = UpdateParams(K=np.asarray(self.K, float), width=int(self.width), height=int(self.height),
                                  sigma_image=float(self.sigma_image), gravity=np.asarray(self.state.imu.W_gravity, float),
                                  min_lost=int(self.min_number_of_frames_to_be_lost),
                                  min_tracked=int(self.min_number_of_frames_to_be_tracked),
                                  use_parallax=bool(self.use_parallax), min_parallax=float(self.min_parallax),
                                  compress_mode=self.compress_mode, gate_mode=self.gate_mode)
            # the window reaches max_number_of_camera_states + 1 poses before it is pruned (MSCKF.py:157-158)
            Fc, Mc = _round_up(F, 64), _round_up(M, 8)
            Nc = max(int(N), int(getattr(self, "max_number_of_camera_states", 0)) + 1, 1)
            if ctx is not None:
                Fc, Mc, Nc = max(Fc, ctx.F), max(Mc, ctx.M), max(Nc, ctx.N)
            new = self._b200_make_context(Fc, Mc, Nc, params)
            if ctx is not None and r.dev_valid:          # the resident covariance moves device to device
                new.P[0, :r.d, :r.d].copy_(ctx.P[0, :r.d, :r.d])
            r.ctx, r.n_dev = new, -1
            ctx = new
        return ctx

    # -- state in / out ---------------------------------------------------------------------------------------------
    def _b200_bind_state(self):
        st = self.state
        d = st.__dict__
        owner = d.get("_b200_owner")
        if owner is not None and owner() is self:
            return
        cov = d.pop("covariance", None)
        if not getattr(type(st), "_b200_resident", False):
            st.__class__ = _resident_state_class(type(st))
        d["_b200_owner"] = weakref.ref(self)
        d["_b200_cov"] = cov if cov is not None else d.get("_b200_cov")
        self._b200_res.dev_valid = False

    def _b200_download_covariance(self):
        r = self._b200_res
        P = r.ctx.P[0, :r.d, :r.d].cpu().numpy()
        P.flags.writeable = False      # in-place edits would silently miss the device copy: assign a new matrix instead
        return P

    def _b200_sync_in(self, ctx, sample=None):
        """Host objects -> device: the small state in one copy; the covariance only if the host holds a newer one."""
        self._b200_bind_state()
        r = self._b200_res
        imu, cams = self.state.imu, list(self.state.cameras.values())
        n = len(cams)
        h = ctx.small_host_views()
        s = h["imu_state"][0]
        s[0:9] = np.asarray(imu.T_W_Ii.R, float).reshape(9)
        s[9:12], s[12:15], s[15:18], s[18:21] = imu.T_W_Ii.t, imu.v_W_Ii, imu.gyroscope_bias, imu.accelerometer_bias
        z = h["imu_null"][0]
        z[0:9] = np.asarray(imu.T_W_Ii_null.R, float).reshape(9)
        z[9:12], z[12:15] = imu.T_W_Ii_null.t, imu.v_W_Ii_null
        h["imu_time"][0] = float(imu.timestamp)
        h["Qc_diag"][0] = np.diag(np.asarray(self.continuous_noise_covariance, float))[[0, 3, 6, 9]]
        R, t = h["cam_R"][0], h["cam_t"][0]
        for k, cam in enumerate(cams):
            R[k] = np.asarray(cam.T_W_Ci.R, float).reshape(9)
            t[k] = cam.T_W_Ci.t
        if sample is not None:
            h["samples_stage"][0, 0] = sample
        ctx.push_small()
        if r.n_dev != n:
            ctx.n_cams.fill_(n)
            r.n_dev = n
        if not r.dev_valid:
            P = np.asarray(self.state.__dict__["_b200_cov"], float)
            assert P.shape[0] == 15 + 6 * n, "covariance size does not match the number of camera states"
            hp = np.zeros((ctx.ldp, ctx.ldp))
            hp[:P.shape[0], :P.shape[0]] = P
            ctx.P[0].copy_(torch.from_numpy(hp))
            r.dev_valid, r.d = True, P.shape[0]
            r.p_uploads += 1
        assert r.d == 15 + 6 * n, "covariance size does not match the number of camera states"
        # device copy authoritative from here on: readers get a fresh download (see _resident_state_class)
        self.state.__dict__["_b200_cov"] = None
        return n, r.d

    def _b200_iso(self, R, t):
        """A pose object of the class the host filter uses (the reference's Isometry3D under the binding)."""
        return type(self.state.imu.T_W_Ii)(R, t)

    def _b200_camera_class(self):
        for cam in self.state.cameras.values():
            return type(cam)
        for k in type(self).__mro__:                       # the reference's MSCKF module star-imports its Camera
            mod = sys.modules.get(k.__module__)
            c = getattr(mod, "Camera", None)
            if c is not None and mod is not sys.modules[__name__]:
                return c
        return Camera

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
        tracks (their inverse_depth_point is updated in place, the triangulated points are appended to the debug lists
        estimated_world_points / currently_processed_world_points, :489-490); returns (valid_features, lost_features)."""
        valid, lost = {}, {}
        if len(features) == 0:
            return valid, lost
        M = max([len(ft.keypoints) for ft in features.values()] + [1])
        ctx = self._b200_context(len(features), M, len(self.state.cameras))
        self._b200_sync_in(ctx)
        self._pack_features(ctx, features)
        ctx.front()
        st = ctx.feat_status[0].cpu().numpy()
        self._download_idp(ctx, features)
        est = self.__dict__.setdefault("estimated_world_points", [])
        cur = self.__dict__.setdefault("currently_processed_world_points", [])
        for j, (idx, ft) in enumerate(features.items()):
            code = st[j] & 0xF
            if code in (_lib.FEAT_ACCEPTED, _lib.FEAT_GATED_OUT):
                valid[idx] = ft
                if st[j] & _lib.FEAT_TRIANGULATED:
                    # W_p_hat = t_0 + depth R_0 K^-1 [u, v, 1]: the point (m, rho) describe from the first observing camera
                    cam0 = self.state.cameras[ft.camera_indices[0]].T_W_Ci
                    m, rho = ft.inverse_depth_point.m, ft.inverse_depth_point.rho
                    W_p_hat = cam0.t + m / (rho * float(np.asarray(cam0.R, float)[:, 2] @ m))
                    est.append(W_p_hat)
                    cur.append(W_p_hat)
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
        n, d = self._b200_sync_in(ctx)
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
        ctx.pull_small()
        h = ctx.small_host_views()
        st = h["imu_state"][0]
        imu = self.state.imu
        imu.T_W_Ii.R = st[0:9].reshape(3, 3).copy()
        imu.T_W_Ii.t[:] = st[9:12]
        imu.v_W_Ii[:] = st[12:15]
        imu.gyroscope_bias[:] = st[15:18]
        imu.accelerometer_bias[:] = st[18:21]
        cam_R, cam_t = h["cam_R"][0], h["cam_t"][0]
        for s, cam in enumerate(self.state.cameras.values()):
            cam.T_W_Ci.R = cam_R[s].reshape(3, 3).copy()
            cam.T_W_Ci.t[:] = cam_t[s]

    def process_imu(self, imu_measurement) -> None:
        """MSCKF.process_imu (MSCKF.py:160-248): integrates the IMU state, propagates the covariance (which stays on the
        device), advances imu.id (:164) -- the key the next state_augmentation files the camera under."""
        imu = self.state.imu
        ctx = self._b200_context(1, 2, max(len(self.state.cameras), 1))
        sample = np.concatenate([[imu_measurement.timestamp], imu_measurement.angular_velocity,
                                 imu_measurement.linear_acceleration]).astype(float)
        self._b200_sync_in(ctx, sample=sample)
        ctx.propagate(ctx.samples_stage[:, :1], use_null_state=True)
        ctx.pull_small()
        st = ctx.small_host_views()["imu_state"][0]
        # IMU.integrate creates NEW pose / velocity objects; the null state then aliases them (MSCKF.py:247-248)
        imu.T_W_Ii = self._b200_iso(st[0:9].reshape(3, 3).copy(), st[9:12].copy())
        imu.v_W_Ii = st[12:15].copy()
        imu.T_W_Ii_null = imu.T_W_Ii
        imu.v_W_Ii_null = imu.v_W_Ii
        imu.timestamp = imu_measurement.timestamp
        if hasattr(imu, "id"):
            imu.id += 1

    # -- window surgery: same names / arguments / side effects as the reference ---------
    def state_augmentation(self) -> None:
        """MSCKF.state_augmentation (MSCKF.py:250-265): the current camera pose enters the window under the key
        `state.imu.id`, P <- [I; J] P [I; J]^T on the device (msckf_augment_batched)."""
        imu = self.state.imu
        T_I_C = imu.T_W_I.inv() * self.T_W_C
        n = len(self.state.cameras)
        ctx = self._b200_context(1, 2, n + 1)
        self._b200_sync_in(ctx)
        ctx.augment(T_I_C.R, T_I_C.t)
        r = self._b200_res
        r.d, r.n_dev = 15 + 6 * (n + 1), n + 1
        ctx.pull_small()
        h = ctx.small_host_views()
        pose = self._b200_iso(h["cam_R"][0, n].reshape(3, 3).copy(), h["cam_t"][0, n].copy())
        self.state.cameras[imu.id] = self._b200_camera_class()(self.K, self.width, self.height, pose)

    def remove_cameras(self, cameras: Dict[int, "Camera"]) -> None:
        """MSCKF.remove_cameras (MSCKF.py:751-778): covariance rows / columns and poses of `cameras` leave the window
        (msckf_remove_cameras_batched), their observations leave every track, empty tracks are deleted together with
        their rows of last_camera_measurement (the matcher's view of the tracks, :772-778)."""
        if len(cameras) == 0:
            return
        ids = list(self.state.cameras.keys())
        n = len(ids)
        ctx = self._b200_context(1, 2, n)
        self._b200_sync_in(ctx)
        mask = torch.zeros((1, ctx.N), dtype=torch.uint8)
        for k in cameras.keys():
            mask[0, ids.index(k)] = 1
        ctx.remove_cameras(mask.to(ctx.device))
        r = self._b200_res
        r.n_dev = n - int(mask.sum())
        r.d = 15 + 6 * r.n_dev
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
            self._b200_forget_track(index)

    def _b200_forget_track(self, index_feature):
        """The deleted track leaves last_camera_measurement as well (MSCKF.py:742-747, 774-778)."""
        lcm = getattr(self, "last_camera_measurement", None)
        if lcm is None:
            return
        hit = np.where(np.asarray(lcm.features_indices) == index_feature)[0]
        if len(hit) > 0:
            lcm.descriptors = np.delete(lcm.descriptors, hit[0], axis=0)
            lcm.features_indices = np.delete(lcm.features_indices, hit[0], axis=0)

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
    (reference MSCKF.__init__, MSCKF.py:80-132; feature extraction / matching are not part of this class)."""

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
        self.last_camera_measurement = None
        self.use_parallax, self.min_parallax = parameters.use_parallax, parameters.min_parallax
        self.min_number_of_frames_to_be_lost = max(parameters.min_number_of_frames_to_be_lost, 1)         # MSCKF.py:119
        self.min_number_of_frames_to_be_tracked = max(parameters.min_number_of_frames_to_be_tracked, 2)   # MSCKF.py:120
        self.max_number_of_camera_states = parameters.max_number_of_camera_states
        self.number_of_residuals_discarded_for_gasting_test = 0
        self.estimated_world_points = []
        self.currently_processed_world_points = []

    def process_features(self):
        """MSCKF.process_features (MSCKF.py:450-456)"""
        self.currently_processed_world_points = []
        valid_features, lost_features = self.get_valid_features(self.features)
        if len(valid_features) > 0:
            self.update(valid_features)
            self.remove_features(lost_features)

    def remove_features(self, features):
        """MSCKF.remove_features (MSCKF.py:739-749): the lost tracks go, then every camera no track observes any more."""
        for index_feature in list(features.keys()):
            del self.features[index_feature]
            self._b200_forget_track(index_feature)
        self.remove_cameras(self.get_cameras_without_features())

    def get_cameras_without_features(self):
        """MSCKF.get_cameras_without_features (MSCKF.py:780-791)"""
        seen = set()
        for ft in self.features.values():
            seen.update(ft.camera_indices)
        return {k: cam for k, cam in self.state.cameras.items() if k not in seen}
