"""Device-resident batch of independent MSCKF filters: the host-side owner of the HBM state the CUDA kernels work on.

Mirrors, for B filters at once, the calls the reference makes on one `MSCKF` object (reference src/msckf/MSCKF.py):
    propagate()       <-> process_imu            (MSCKF.py:160-248)   n samples per launch
    augment()         <-> state_augmentation     (MSCKF.py:250-265)
    update()          <-> get_valid_features + update + correct   (MSCKF.py:458-495, 570-661)
    remove_cameras()  <-> remove_cameras (covariance/pose part)   (MSCKF.py:751-758)
PyTorch is used for device memory and streams only; all arithmetic happens in csrc/*.cu through include/msckf_b200.h.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib


def _chi2_table(n):
    from scipy.stats import chi2
    with np.errstate(all="ignore"):
        t = np.array([chi2.ppf(0.95, k) for k in range(n)], float)
    t[0] = np.nan     # chi2.ppf(0.95, 0) is NaN: a feature with an empty null space is rejected (MSCKF.py:566-568)
    return t


@dataclass
class UpdateParams:
    """The MSCKFParameters fields the hot path reads (reference MSCKF.py:14-46) plus the compression mode."""
    K: np.ndarray = field(default_factory=lambda: np.array([[320.0, 0, 320], [0, 320.0, 240], [0, 0, 1]]))
    width: int = 640
    height: int = 480
    sigma_image: float = 0.1
    gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.81]))
    min_lost: int = 2
    min_tracked: int = 4
    use_parallax: bool = True
    min_parallax: float = 45.0
    compress_mode: int = _lib.COMPRESS_QR
    gate_mode: int = _lib.GATE_EXACT

    def to_c(self):
        p = _lib.Params()
        K = np.asarray(self.K, float)
        Kinv = np.linalg.inv(K)
        for i in range(9):
            p.K[i] = float(K.flat[i])
            p.Kinv[i] = float(Kinv.flat[i])
        p.width, p.height = int(self.width), int(self.height)
        p.sigma_image = float(self.sigma_image)
        for i in range(3):
            p.gravity[i] = float(self.gravity[i])
        p.min_lost = max(int(self.min_lost), 1)            # clamps of MSCKF.__init__, MSCKF.py:119-120
        p.min_tracked = max(int(self.min_tracked), 2)
        p.use_parallax = int(bool(self.use_parallax))
        p.min_parallax_deg = float(self.min_parallax)
        p.compress_mode = int(self.compress_mode)
        p.gate_mode = int(self.gate_mode)
        return p

    @staticmethod
    def from_case(c, compress_mode=_lib.COMPRESS_QR, gate_mode=_lib.GATE_EXACT):
        return UpdateParams(K=np.array(c["K"], float), width=int(c["width"]), height=int(c["height"]),
                            sigma_image=float(c["sigma"]), gravity=np.array(c["gravity"], float),
                            min_lost=int(c["min_lost"]), min_tracked=int(c["min_tracked"]),
                            use_parallax=bool(c["use_parallax"]), min_parallax=float(c["min_parallax"]),
                            compress_mode=compress_mode, gate_mode=gate_mode)


SAMPLES_STAGE = 16      # IMU samples the small-state slab can stage per filter (the reference delivers ~10 per frame)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class BatchedMSCKF:
    """B filters x F feature slots x M observations x N window poses, all state resident in HBM (FP64)."""

    def __init__(self, B, F, M, N, params: UpdateParams, device="cuda", ldp=None):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedMSCKF needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load_library()
        self.B, self.F, self.M, self.N = int(B), int(F), int(M), int(N)
        self.ldp = int(ldp) if ldp is not None else 15 + 6 * self.N
        self.params = params
        self.device = torch.device(device)
        self.dims = _lib.Dims(self.B, self.F, self.M, self.N, self.ldp)
        self.cpar = params.to_c()
        if self.M > self.lib.msckf_max_obs_per_feature() or self.N > self.lib.msckf_max_cameras(params.compress_mode):
            raise ValueError(f"capacity exceeded: M={self.M} (max {self.lib.msckf_max_obs_per_feature()}), "
                             f"N={self.N} (max {self.lib.msckf_max_cameras(params.compress_mode)} in this mode)")
        dv, f64, i32 = self.device, torch.float64, torch.int32
        z = lambda *s, dtype=f64: torch.zeros(*s, dtype=dtype, device=dv)  # noqa: E731
        # filter state.  Everything except P lives in ONE slab (imu_state, imu_null, imu_time, Qc_diag, cam_R, cam_t and a
        # staging area for IMU samples are views of it, each padded to 256 B), so that a host-side owner of a single
        # filter moves the whole small state with one copy in each direction (push_small / pull_small).
        self.P = z(B, self.ldp, self.ldp)
        layout, off = {}, 0
        for name, shape in (("imu_state", (B, 21)), ("imu_null", (B, 15)), ("imu_time", (B,)), ("Qc_diag", (B, 4)),
                            ("cam_R", (B, N, 9)), ("cam_t", (B, N, 3)), ("samples_stage", (B, SAMPLES_STAGE, 7))):
            n = int(np.prod(shape))
            layout[name] = (off, n, shape)
            off += (n + 31) // 32 * 32
        self.small_layout = layout
        self.small_slab = z(off)
        for name, (o, n, shape) in layout.items():
            setattr(self, name, self.small_slab[o:o + n].view(shape))
        self.small_host = None                   # pinned host mirror of the slab, created on first use
        self.imu_state[:, [0, 4, 8]] = 1.0
        self.imu_null[:, [0, 4, 8]] = 1.0
        self.n_cams = z(B, dtype=i32)
        # tracks
        self.feat_len = z(B, F, dtype=i32)
        self.obs_uv = z(B, F, M, 2)
        self.obs_dir = z(B, F, M, 3)
        self.obs_conf = z(B, F, M)
        self.obs_slot = z(B, F, M, dtype=i32)
        self.idp_base = z(B, F, 3)
        self.idp_anchor_slot = torch.full((B, F), -1, dtype=i32, device=dv)
        self.idp_m = z(B, F, 3)
        self.idp_rho = z(B, F)
        self.lost_cnt = z(B, F, dtype=i32)
        self.tracked_cnt = z(B, F, dtype=i32)
        # outputs
        self.feat_status = z(B, F, dtype=torch.uint8)
        self.gamma = z(B, F)
        self.dof = z(B, F, dtype=i32)
        self.delta_x = z(B, self.ldp)
        self.n_accepted = z(B, dtype=i32)
        self.n_rejected = z(B, dtype=i32)
        self.T_H = None
        self.r_n = None
        self.chi2_len = 2 * self.M + 3
        self.chi2 = torch.from_numpy(_chi2_table(self.chi2_len)).to(dv)
        nbytes = self.lib.msckf_update_workspace_bytes(C.byref(self.dims), params.compress_mode)
        if nbytes == 0:
            raise ValueError("unsupported dimensions")
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=dv)
        self._remove_ws = None
        self._overflow = None
        self.window_overflow = None
        self._alt_tracks = None
        self.obs_desc = None                     # [B, F, R, D] float32: Feature.descriptors, one row per camera of the window
        self._alt_desc = None
        self.lcm_desc = None                     # [B, F, D] float32: last_camera_measurement.descriptors (the matcher's view)
        self.desc_sum = None                     # [B, F, D + 2] float64: running (sum_i w_i d_i | sum_i w_i | pad) of every track's list
        self.cam_row = None                      # [B, N] int32: descriptor row of the camera in window slot s (the camera ring)
        self.desc_rows = max(self.M, self.N)     # R: descriptor rows per track (one per camera of the window)
        self._alt_lcm = None
        self._alt_sum = None
        self.desc_dim = 0
        self.feature_updates = None              # device int64: features that went through rows 3-8 (set count_feature_updates)
        self.count_feature_updates = False
        self._samples_dev = None

    # ------------------------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def small_host_views(self):
        """numpy views (name -> array of the field's shape) of the pinned host mirror of the small-state slab."""
        if self.small_host is None:
            self.small_host = torch.zeros(self.small_slab.numel(), dtype=torch.float64).pin_memory()
            self.small_host.copy_(self.small_slab)
            a = self.small_host.numpy()
            self._small_np = {name: a[o:o + n].reshape(shape) for name, (o, n, shape) in self.small_layout.items()}
        return self._small_np

    def push_small(self):
        """host mirror -> device: IMU state, null state, time, noise densities, window poses, staged samples (one H2D)."""
        self.small_slab.copy_(self.small_host, non_blocking=True)

    def pull_small(self):
        """device -> host mirror (one D2H) and wait for it."""
        self.small_host.copy_(self.small_slab, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()

    def _front_tail(self):
        return [_ptr(self.cam_R), _ptr(self.cam_t), _ptr(self.n_cams), _ptr(self.feat_len), _ptr(self.obs_uv),
                _ptr(self.obs_dir), _ptr(self.obs_conf), _ptr(self.obs_slot), _ptr(self.idp_base),
                _ptr(self.idp_anchor_slot), _ptr(self.idp_m), _ptr(self.idp_rho), _ptr(self.lost_cnt),
                _ptr(self.tracked_cnt), _ptr(self.feat_status), _ptr(self.gamma), _ptr(self.dof)]

    def update(self):
        """get_valid_features + update + correct for every filter (one C-ABI call, three kernels)."""
        rc = self.lib.msckf_update_batched(C.byref(self.dims), C.byref(self.cpar), _ptr(self.chi2), self.chi2_len,
                                           _ptr(self.P), _ptr(self.imu_state), *self._front_tail(),
                                           _ptr(self.delta_x), _ptr(self.n_accepted), _ptr(self.n_rejected),
                                           _ptr(self.workspace), self.workspace.numel(), self._stream())
        _lib.check(rc, "msckf_update_batched")
        if self.count_feature_updates:
            if self.feature_updates is None:
                self.feature_updates = torch.zeros((), dtype=torch.int64, device=self.device)
            code = self.feat_status & 0xF
            self.feature_updates += ((code == _lib.FEAT_ACCEPTED) | (code == _lib.FEAT_GATED_OUT)).sum()

    def front(self):
        rc = self.lib.msckf_front_batched(C.byref(self.dims), C.byref(self.cpar), _ptr(self.chi2), self.chi2_len,
                                          _ptr(self.P), *self._front_tail(), _ptr(self.workspace),
                                          self.workspace.numel(), self._stream())
        _lib.check(rc, "msckf_front_batched")

    def compress(self):
        c = 6 * self.N
        if self.T_H is None:
            self.T_H = torch.zeros(self.B, c, c, dtype=torch.float64, device=self.device)
            self.r_n = torch.zeros(self.B, c, dtype=torch.float64, device=self.device)
        rc = self.lib.msckf_compress_batched(C.byref(self.dims), C.byref(self.cpar), _ptr(self.n_cams),
                                             _ptr(self.feat_len), _ptr(self.obs_slot), _ptr(self.feat_status),
                                             _ptr(self.T_H), _ptr(self.r_n), _ptr(self.n_accepted),
                                             _ptr(self.n_rejected), _ptr(self.workspace), self.workspace.numel(),
                                             self._stream())
        _lib.check(rc, "msckf_compress_batched")
        return self.T_H, self.r_n

    def apply_update(self):
        rc = self.lib.msckf_apply_update_batched(C.byref(self.dims), C.byref(self.cpar), _ptr(self.P),
                                                 _ptr(self.imu_state), _ptr(self.cam_R), _ptr(self.cam_t),
                                                 _ptr(self.n_cams), _ptr(self.T_H), _ptr(self.r_n),
                                                 _ptr(self.n_accepted), _ptr(self.delta_x), _ptr(self.workspace),
                                                 self.workspace.numel(), self._stream())
        _lib.check(rc, "msckf_apply_update_batched")

    def propagate(self, samples, use_null_state=False):
        """samples: [B, n, 7] device tensor (timestamp, omega, acc).  With use_null_state=False the null state is the
        state on entry -- the aliasing the reference exhibits after its first process_imu call (MSCKF.py:247-248)."""
        assert samples.shape[0] == self.B and samples.shape[2] == 7 and samples.dtype == torch.float64
        samples = samples.contiguous()
        g = (C.c_double * 3)(*[float(x) for x in self.params.gravity])
        rc = self.lib.msckf_propagate_batched(C.byref(self.dims), int(samples.shape[1]), _ptr(samples),
                                              _ptr(self.imu_time), _ptr(self.Qc_diag), g, _ptr(self.imu_state),
                                              _ptr(self.imu_null) if use_null_state else C.c_void_p(0),
                                              _ptr(self.P), _ptr(self.n_cams), self._stream())
        _lib.check(rc, "msckf_propagate_batched")

    def step_host(self, host_in, host_samples=None, host_out=None, chunk_filters=592, overlap_previous=False,
                  use_null_state=False):
        """One camera-frame step of every filter driven from HOST buffers (msckf_filter_step_host): the IMU samples
        since the last frame are propagated, then get_valid_features + update + correct run, chunk by chunk, with the
        uploads / downloads of neighbouring chunks overlapping the kernels.

        host_in:  dict name -> host tensor for any of _lib.HOST_INPUTS (the track arrays, [B, ...] in the device
                  layout; pinned memory makes the copies asynchronous); missing names stay as resident on the device.
        host_samples: [B, n, 7] host tensor of IMU samples (timestamp, omega, acc) or None.
        host_out: dict name -> host tensor for any of _lib.HOST_OUTPUTS.
        Results are valid once the current stream is synchronised."""
        io = _lib.HostStep()
        shapes = dict(feat_len=self.feat_len, obs_uv=self.obs_uv, obs_dir=self.obs_dir, obs_conf=self.obs_conf,
                      obs_slot=self.obs_slot, idp_base=self.idp_base, idp_anchor_slot=self.idp_anchor_slot,
                      idp_m=self.idp_m, idp_rho=self.idp_rho, lost_cnt=self.lost_cnt, tracked_cnt=self.tracked_cnt,
                      imu_state=self.imu_state, cam_R=self.cam_R, cam_t=self.cam_t, idp_m_out=self.idp_m,
                      idp_rho_out=self.idp_rho, feat_status=self.feat_status, gamma=self.gamma, delta_x=self.delta_x,
                      n_accepted=self.n_accepted, n_rejected=self.n_rejected)
        for group, names in ((host_in or {}, _lib.HOST_INPUTS), (host_out or {}, _lib.HOST_OUTPUTS)):
            for k, t in group.items():
                if k not in names:
                    raise KeyError(f"{k} is not a host {'input' if names is _lib.HOST_INPUTS else 'output'} of step_host")
                d = shapes[k]
                if t.device.type != "cpu" or t.dtype != d.dtype or tuple(t.shape) != tuple(d.shape) or not t.is_contiguous():
                    raise ValueError(f"{k}: expected a contiguous host tensor of shape {tuple(d.shape)}, dtype {d.dtype}")
                setattr(io, k, t.data_ptr())
        n = 0
        if host_samples is not None:
            if (host_samples.device.type != "cpu" or host_samples.dtype != torch.float64 or host_samples.dim() != 3 or
                    host_samples.shape[0] != self.B or host_samples.shape[2] != 7 or not host_samples.is_contiguous()):
                raise ValueError("host_samples: expected a contiguous [B, n, 7] float64 host tensor")
            n = int(host_samples.shape[1])
            if self._samples_dev is None or self._samples_dev.shape[1] != n:     # device staging, [B][n][7] exactly
                self._samples_dev = torch.empty(self.B, n, 7, dtype=torch.float64, device=self.device)
            io.imu_samples = host_samples.data_ptr()
        io.n_samples = n
        io.use_null_state = int(bool(use_null_state))
        for i in range(3):
            io.gravity[i] = float(self.params.gravity[i])
        dev = _lib.DeviceBuffers()
        for k, _ in _lib.DeviceBuffers._fields_:
            t = self._samples_dev if k == "imu_samples" else getattr(self, k)
            setattr(dev, k, t.data_ptr() if t is not None else None)
        rc = self.lib.msckf_filter_step_host(C.byref(self.dims), C.byref(self.cpar), _ptr(self.chi2), self.chi2_len,
                                             C.byref(dev), C.byref(io), int(chunk_filters), int(bool(overlap_previous)),
                                             _ptr(self.workspace), self.workspace.numel(), self._stream())
        _lib.check(rc, "msckf_filter_step_host")

    def augment(self, R_I_C, t_I_C, active=None):
        """state_augmentation (MSCKF.py:250-265) for every (active) filter.  A filter whose window is already full is left
        untouched and counted in self.window_overflow (device int32; the reference's dict has no capacity, so size N for
        max_number_of_camera_states + 1 + the frames that may pass without a valid feature); check_window_overflow() raises."""
        T = (C.c_double * 12)(*[float(x) for x in np.asarray(R_I_C, float).flatten()],
                              *[float(x) for x in np.asarray(t_I_C, float).flatten()])
        if self.window_overflow is None:
            self.window_overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        if self.cam_row is not None:             # the new camera's descriptor row (lowest free one), before n_cams grows
            rc = self.lib.msckf_descriptor_rows_assign_batched(C.byref(self.dims), self.desc_rows, _ptr(self.n_cams), _ptr(active),
                                                               _ptr(self.cam_row), self._stream())
            _lib.check(rc, "msckf_descriptor_rows_assign_batched")
        rc = self.lib.msckf_augment_batched(C.byref(self.dims), T, _ptr(self.imu_state), _ptr(self.cam_R),
                                            _ptr(self.cam_t), _ptr(self.P), _ptr(self.n_cams), _ptr(active),
                                            _ptr(self.window_overflow), self._stream())
        _lib.check(rc, "msckf_augment_batched")

    def check_window_overflow(self):
        """Raises if any state_augmentation found its filter's window full since the last check (synchronises)."""
        if self.window_overflow is not None:
            n = int(self.window_overflow.item())
            if n:
                self.window_overflow.zero_()
                raise OverflowError(f"{n} state augmentation(s) found the camera window full (N = {self.N}): the new pose was "
                                    "NOT appended; allocate a larger N")

    def remove_cameras(self, mask):
        """mask: [B, N] uint8 device tensor, 1 = delete that window slot."""
        if self._remove_ws is None:
            n = self.lib.msckf_remove_workspace_bytes(C.byref(self.dims))
            self._remove_ws = torch.empty(n, dtype=torch.uint8, device=self.device)
        if self.cam_row is not None:             # the descriptor rows of the deleted cameras become free again
            rc = self.lib.msckf_descriptor_rows_remove_batched(C.byref(self.dims), _ptr(self.n_cams), _ptr(mask), _ptr(self.cam_row),
                                                               self._stream())
            _lib.check(rc, "msckf_descriptor_rows_remove_batched")
        rc = self.lib.msckf_remove_cameras_batched(C.byref(self.dims), _ptr(mask), _ptr(self.cam_R),
                                                   _ptr(self.cam_t), _ptr(self.P), _ptr(self.n_cams),
                                                   _ptr(self._remove_ws), self._remove_ws.numel(), self._stream())
        _lib.check(rc, "msckf_remove_cameras_batched")

    def append_observations(self, action, new_uv, new_conf):
        """Track bookkeeping of MSCKF.add_camera_measurements after the matcher (MSCKF.py:400-438): action [B, F] uint8
        (0 untouched, 1 lost this frame, 2 append, 3 new track), new_uv [B, F, 2], new_conf [B, F] on the device.
        Returns the device int counting tracks that were full."""
        if self._overflow is None:
            self._overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        kinv = (C.c_double * 9)(*[float(self.cpar.Kinv[i]) for i in range(9)])
        rc = self.lib.msckf_append_observations_batched(
            C.byref(self.dims), kinv, _ptr(self.n_cams), _ptr(self.cam_R), _ptr(self.cam_t), _ptr(action), _ptr(new_uv),
            _ptr(new_conf), _ptr(self.feat_len), _ptr(self.obs_uv), _ptr(self.obs_dir), _ptr(self.obs_conf),
            _ptr(self.obs_slot), _ptr(self.idp_base), _ptr(self.idp_anchor_slot), _ptr(self.idp_m), _ptr(self.idp_rho),
            _ptr(self.lost_cnt), _ptr(self.tracked_cnt), _ptr(self._overflow), self._stream())
        _lib.check(rc, "msckf_append_observations_batched")
        return self._overflow

    def match_descriptors(self, desc_tracks, desc_new, min_cosine_similarity=0.82, n_tracks=None, n_new=None, kp_new=None,
                          score_new=None):
        """FeatureExtractor.match / XFeat.match (FeatureExtractor.py:62-84, xfeat.py:327-348) for every filter:
        desc_tracks [B, F, D] float32 (rows = track slots), desc_new [B, F2, D] float32 (the frame's detections), optional
        per-filter counts n_tracks / n_new [B] int32.  Returns (match12 [B, F] int32, match21 [B, F2] int32), -1 = no
        match.  With kp_new [B, F2, 2] and score_new [B, F2] (float64) the proposal list of MSCKF.py:317-318 / :400-434
        is produced too: returns (match12, match21, action, new_uv, new_conf, n_tracks_out) ready for verify_matches and
        append_observations (new tracks in the free slots from n_tracks[b] on)."""
        B, F1, D = desc_tracks.shape
        F2 = desc_new.shape[1]
        assert B == self.B and F1 == self.F and desc_new.shape[0] == B and desc_new.shape[2] == D
        assert desc_tracks.dtype == torch.float32 and desc_new.dtype == torch.float32
        assert desc_tracks.is_contiguous() and desc_new.is_contiguous()
        m12 = torch.empty((B, F1), dtype=torch.int32, device=self.device)
        m21 = torch.empty((B, F2), dtype=torch.int32, device=self.device)
        prop = kp_new is not None
        if prop:
            action = torch.empty((B, F1), dtype=torch.uint8, device=self.device)
            new_uv = torch.empty((B, F1, 2), dtype=torch.float64, device=self.device)
            new_conf = torch.empty((B, F1), dtype=torch.float64, device=self.device)
            n_out = torch.empty(B, dtype=torch.int32, device=self.device)
            if self._overflow is None:
                self._overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
            assert kp_new.dtype == torch.float64 and score_new.dtype == torch.float64
        rc = self.lib.msckf_match_descriptors_batched(
            B, F1, F2, D, float(min_cosine_similarity), _ptr(n_tracks) if n_tracks is not None else None,
            _ptr(n_new) if n_new is not None else None, _ptr(desc_tracks), _ptr(desc_new), _ptr(m12), _ptr(m21),
            _ptr(kp_new) if prop else None, _ptr(score_new) if prop else None, _ptr(action) if prop else None,
            _ptr(new_uv) if prop else None, _ptr(new_conf) if prop else None, _ptr(n_out) if prop else None,
            _ptr(self._overflow) if prop else None, self._stream())
        _lib.check(rc, "msckf_match_descriptors_batched")
        return (m12, m21, action, new_uv, new_conf, n_out) if prop else (m12, m21)

    def verify_matches(self, action, new_uv, epipolar_rejection_threshold=5.0, homography_rejection_threshold=5.0,
                       counters=None):
        """The epipolar / homography walk of MSCKF.add_camera_measurements (MSCKF.py:333-401) over the resident tracks:
        proposals (action == 2, new_uv) that fail against any past observation are demoted to action 1 IN PLACE.
        Returns verdict [B, F] uint8 (0 none, 1 passed, 2 homography, 3 epipolar); counters = optional pair of [B] int32
        device tensors incremented like the reference's two discard counters."""
        verdict = torch.empty((self.B, self.F), dtype=torch.uint8, device=self.device)
        k = (C.c_double * 9)(*[float(self.cpar.K[i]) for i in range(9)])
        kinv = (C.c_double * 9)(*[float(self.cpar.Kinv[i]) for i in range(9)])
        nh, ne = counters if counters is not None else (None, None)
        rc = self.lib.msckf_verify_matches_batched(
            C.byref(self.dims), k, kinv, float(epipolar_rejection_threshold), float(homography_rejection_threshold),
            _ptr(self.n_cams), _ptr(self.cam_R), _ptr(self.cam_t), _ptr(self.feat_len), _ptr(self.obs_uv), _ptr(self.obs_slot),
            _ptr(new_uv), _ptr(action), _ptr(verdict), _ptr(nh) if nh is not None else None,
            _ptr(ne) if ne is not None else None, self._stream())
        _lib.check(rc, "msckf_verify_matches_batched")
        return verdict

    def select_poorest_cameras(self, max_select=2, prune_above=-1):
        """MSCKF.prune_poorest_camera_states' selection (MSCKF.py:712-733) for every filter: returns
        (remove_mask [B, N] uint8, feat_touch [B, F] uint8) on the device.  Filters whose window holds at most
        `prune_above` cameras select nothing (MSCKF.py:157-158)."""
        mask = torch.empty((self.B, self.N), dtype=torch.uint8, device=self.device)
        touch = torch.empty((self.B, self.F), dtype=torch.uint8, device=self.device)
        rc = self.lib.msckf_select_poorest_cameras_batched(C.byref(self.dims), int(max_select), int(prune_above), _ptr(self.n_cams),
                                                           _ptr(self.feat_len), _ptr(self.obs_slot), _ptr(mask), _ptr(touch),
                                                           self._stream())
        _lib.check(rc, "msckf_select_poorest_cameras_batched")
        return mask, touch

    def drop_observations(self, mask):
        """Track part of MSCKF.remove_cameras (MSCKF.py:760-770); call before remove_cameras(mask).  The per-observation
        descriptors (if this batch keeps them) lose the same rows (:764-765)."""
        if self.obs_desc is not None:
            rc = self.lib.msckf_drop_descriptors_batched(C.byref(self.dims), self.desc_dim, _ptr(mask), _ptr(self.feat_len),
                                                         _ptr(self.obs_slot), _ptr(self.obs_desc), _ptr(self.obs_conf),
                                                         _ptr(self.desc_sum), self.desc_rows, _ptr(self.cam_row), self._stream())
            _lib.check(rc, "msckf_drop_descriptors_batched")
        rc = self.lib.msckf_drop_observations_batched(C.byref(self.dims), _ptr(mask), _ptr(self.n_cams), _ptr(self.cam_t),
                                                      _ptr(self.feat_len), _ptr(self.obs_uv), _ptr(self.obs_dir),
                                                      _ptr(self.obs_conf), _ptr(self.obs_slot), _ptr(self.idp_base),
                                                      _ptr(self.idp_anchor_slot), self._stream())
        _lib.check(rc, "msckf_drop_observations_batched")

    def _track_struct(self, tensors):
        st = _lib.TrackArrays()
        for n in _lib.TRACK_FIELDS:
            setattr(st, n, tensors[n].data_ptr())
        return st

    def compact_tracks(self, drop_lost=False, return_map=False):
        """Stable compaction of the track slots (MSCKF.remove_features / the track deletions of remove_cameras), in place:
        the survivors move, in order, to the front; a filter in which nothing moves costs one scan of its lengths.
        Returns active [B] uint8 (drop_lost: filters that deleted their lost tracks); with return_map also dest [B, F]
        int32: the new slot of every old slot, -1 for a deleted track."""
        cur = {n: getattr(self, n) for n in _lib.TRACK_FIELDS}
        if self._alt_tracks is None:
            self._alt_tracks = {n: torch.zeros_like(t) for n, t in cur.items()}      # scratch: stages the tracks that move
        active = torch.zeros(self.B, dtype=torch.uint8, device=self.device)
        want_map = return_map or self.obs_desc is not None
        dest = torch.empty((self.B, self.F), dtype=torch.int32, device=self.device) if want_map else None
        a, b = self._track_struct(cur), self._track_struct(self._alt_tracks)
        rc = self.lib.msckf_compact_tracks_inplace_batched(C.byref(self.dims), int(bool(drop_lost)), _ptr(self.feat_status),
                                                           C.byref(a), C.byref(b), _ptr(active), _ptr(dest), self._stream())
        _lib.check(rc, "msckf_compact_tracks_inplace_batched")
        if self.obs_desc is not None:            # the descriptor lists move with their tracks ...
            if self._alt_desc is None:
                self._alt_desc = torch.zeros_like(self.obs_desc)
                self._alt_lcm = torch.zeros_like(self.lcm_desc)
                self._alt_sum = torch.zeros_like(self.desc_sum)
            dr = _lib.Dims(self.B, self.F, self.desc_rows, self.N, self.ldp)
            rows = (self.feat_len > 0).to(torch.int32) * self.desc_rows          # a live track moves all its camera rows
            rc = self.lib.msckf_compact_descriptors_inplace_batched(C.byref(dr), self.desc_dim, _ptr(dest), _ptr(rows),
                                                                    _ptr(self.obs_desc), _ptr(self._alt_desc), 0, self._stream())
            _lib.check(rc, "msckf_compact_descriptors_inplace_batched")
            # ... and so do their rows of last_camera_measurement (MSCKF.py:742-747, 774-778): one row per track
            d1 = _lib.Dims(self.B, self.F, 1, self.N, self.ldp)
            rc = self.lib.msckf_compact_descriptors_inplace_batched(C.byref(d1), self.desc_dim, _ptr(dest), _ptr(self.feat_len),
                                                                    _ptr(self.lcm_desc), _ptr(self._alt_lcm), 1, self._stream())
            _lib.check(rc, "msckf_compact_descriptors_inplace_batched")
            # ... and so do the running sums (rows of D + 2 doubles = 2 (D + 2) floats)
            rc = self.lib.msckf_compact_descriptors_inplace_batched(C.byref(d1), 2 * (self.desc_dim + 2), _ptr(dest), _ptr(self.feat_len),
                                                                    _ptr(self.desc_sum), _ptr(self._alt_sum), 1, self._stream())
            _lib.check(rc, "msckf_compact_descriptors_inplace_batched")
        return (active, dest) if return_map else active

    def unobserved_cameras(self, active=None):
        """MSCKF.get_cameras_without_features (MSCKF.py:739-749): [B, N] uint8 mask of the window slots no track observes."""
        mask = torch.empty((self.B, self.N), dtype=torch.uint8, device=self.device)
        rc = self.lib.msckf_unobserved_cameras_batched(C.byref(self.dims), _ptr(self.n_cams), _ptr(self.feat_len),
                                                       _ptr(self.obs_slot), _ptr(active), _ptr(mask), self._stream())
        _lib.check(rc, "msckf_unobserved_cameras_batched")
        return mask

    def process_features(self, return_map=False):
        """MSCKF.process_features (MSCKF.py:450-456) on resident tracks: one update with every track; filters that had a
        valid feature then delete their lost tracks (remove_features) and the cameras nothing observes any more."""
        self.update()
        active = self.compact_tracks(drop_lost=True, return_map=return_map)
        dest = None
        if return_map:
            active, dest = active
        mask = self.unobserved_cameras(active)
        self.drop_observations(mask)          # renumbers the slots of the surviving observations
        self.remove_cameras(mask)
        return dest

    def frame_step(self, samples, R_I_C, t_I_C, action, new_uv, new_conf, max_cameras, use_null_state=False, return_maps=False):
        """One camera frame of every filter with everything resident (MSCKF.imu_callback x n + feature_callback,
        MSCKF.py:135-158): the IMU samples since the last frame, state_augmentation, the track bookkeeping of
        add_camera_measurements (the matcher's decisions come from the caller), process_features, and the pruning of the
        two poorest cameras in the filters whose window exceeds max_cameras.  With return_maps: the two slot maps
        ([B, F] int32, -1 = deleted) of the track deletions, for callers that keep per-track records of their own."""
        self.propagate(samples, use_null_state=use_null_state)
        self.augment(R_I_C, t_I_C)
        self.append_observations(action, new_uv, new_conf)
        m1 = self.process_features(return_map=return_maps)
        m2 = self.prune_poorest_camera_states(prune_above=max_cameras, return_map=return_maps)
        return (m1, m2) if return_maps else None

    # ------------------------------------------------------------------------------------------------------------
    def _ensure_descriptors(self, D):
        if self.obs_desc is None or self.desc_dim != D:
            self.desc_dim = int(D)
            self.obs_desc = torch.zeros((self.B, self.F, self.desc_rows, D), dtype=torch.float32, device=self.device)
            # cameras already in the window get the rows 0 .. n-1 in slot order
            self.cam_row = torch.arange(self.N, dtype=torch.int32, device=self.device).repeat(self.B, 1).contiguous()
            self.lcm_desc = torch.zeros((self.B, self.F, D), dtype=torch.float32, device=self.device)
            self.desc_sum = torch.zeros((self.B, self.F, D + 2), dtype=torch.float64, device=self.device)
            self._alt_desc = self._alt_lcm = self._alt_sum = None

    def score_filter(self, desc_new, kp_new, score_new, n_new=None):
        """The score filter of MSCKF.add_camera_measurements (MSCKF.py:279-284) on the device: detections with
        score >= 0.5 mean(score) survive, in order.  Returns (desc, kp, score, n) compacted copies."""
        B, F2, D = desc_new.shape
        do, ko, so = torch.zeros_like(desc_new), torch.zeros_like(kp_new), torch.zeros_like(score_new)
        no = torch.empty(B, dtype=torch.int32, device=self.device)
        rc = self.lib.msckf_score_filter_batched(B, F2, D, _ptr(n_new), _ptr(score_new), _ptr(kp_new), _ptr(desc_new), _ptr(so),
                                                 _ptr(ko), _ptr(do), _ptr(no), self._stream())
        _lib.check(rc, "msckf_score_filter_batched")
        return do, ko, so, no

    def track_descriptors(self, out=None, keep_old=None, from_lists=False):
        """[B, F, D] float32: the score-weighted mean descriptor of every track over its current descriptor list
        (np.average(..., weights=scores), MSCKF.py:436-441).  keep_old [B] uint8: filters whose rows of `out` stay.
        By default from the running sums the append / drop kernels maintain (D + 2 doubles per track); from_lists=True
        re-reads every list (the two agree to FP64 summation order, i.e. below the float32 rounding of the rows)."""
        if out is None:
            out = torch.empty((self.B, self.F, self.desc_dim), dtype=torch.float32, device=self.device)
        if from_lists:
            rc = self.lib.msckf_track_descriptors_batched(C.byref(self.dims), self.desc_dim, _ptr(self.feat_len), _ptr(self.obs_conf),
                                                          _ptr(self.obs_desc), _ptr(keep_old), _ptr(out), None, self.desc_rows,
                                                          _ptr(self.obs_slot), _ptr(self.cam_row), self._stream())
            _lib.check(rc, "msckf_track_descriptors_batched")
        else:
            rc = self.lib.msckf_track_descriptors_from_sums_batched(C.byref(self.dims), self.desc_dim, _ptr(self.feat_len),
                                                                    _ptr(self.desc_sum), _ptr(keep_old), _ptr(out), self._stream())
            _lib.check(rc, "msckf_track_descriptors_from_sums_batched")
        return out

    def rebuild_descriptor_sums(self):
        """The running sums recomputed from the lists (after the lists were written directly, e.g. loaded from a checkpoint)."""
        rc = self.lib.msckf_track_descriptors_batched(C.byref(self.dims), self.desc_dim, _ptr(self.feat_len), _ptr(self.obs_conf),
                                                      _ptr(self.obs_desc), None, None, _ptr(self.desc_sum), self.desc_rows,
                                                      _ptr(self.obs_slot), _ptr(self.cam_row), self._stream())
        _lib.check(rc, "msckf_track_descriptors_batched")

    def append_descriptors(self, action, match12, match21, n_tracks, n_new, desc_new, new_conf):
        """feature.descriptors.append(...) / feature.scores.append(...) for the accepted matches and the new tracks
        (MSCKF.py:407-408, 425-427); call with the final actions and before append_observations.  new_conf [B, F] = the
        score of the detection every track slot receives (what append_observations stores as obs_conf)."""
        rc = self.lib.msckf_append_descriptors_batched(C.byref(self.dims), self.desc_dim, int(desc_new.shape[1]), _ptr(action),
                                                       _ptr(match12), _ptr(match21), _ptr(n_tracks), _ptr(n_new), _ptr(desc_new),
                                                       _ptr(self.feat_len), _ptr(self.obs_desc), _ptr(new_conf), _ptr(self.desc_sum),
                                                       self.desc_rows, _ptr(self.n_cams), _ptr(self.cam_row), self._stream())
        _lib.check(rc, "msckf_append_descriptors_batched")

    def frame_step_detections(self, samples, R_I_C, t_I_C, desc_new, kp_new, score_new, n_new, max_cameras,
                              min_cosine_similarity=0.82, epipolar_rejection_threshold=5.0,
                              homography_rejection_threshold=5.0, use_null_state=False, counters=None, score_filter=False):
        """One camera frame of every filter from the frame's DETECTIONS (MSCKF.imu_callback x n + feature_callback,
        MSCKF.py:135-158, with add_camera_measurements :267-448 on the device): propagation, state_augmentation, the
        score filter (:279-284, score_filter=True; otherwise the caller filtered), the tracks' score-weighted mean
        descriptors (:436-441) over their per-observation lists, the mutual-NN matcher (:317), the epipolar / homography
        walk (:333-401), the track bookkeeping with the descriptor lists (:400-438), process_features and the pruning --
        the descriptor rows of a pruned camera leave their tracks (:764-765) before the next mean.
        desc_new [B, F2, D] float32, kp_new [B, F2, 2] / score_new [B, F2] float64, n_new [B] int32.
        Returns (action, verdict)."""
        B, F2, D = desc_new.shape
        self._ensure_descriptors(D)
        self.propagate(samples, use_null_state=use_null_state)
        self.augment(R_I_C, t_I_C)
        if score_filter:
            desc_new, kp_new, score_new, n_new = self.score_filter(desc_new, kp_new, score_new, n_new)
        n_tracks = (self.feat_len > 0).sum(1).to(torch.int32)            # the slots in use are 0 .. n-1 (tracks stay compacted)
        # the matcher sees last_camera_measurement: the means built at the end of the PREVIOUS frame's bookkeeping
        m12, m21, action, new_uv, new_conf, _ = self.match_descriptors(self.lcm_desc, desc_new, min_cosine_similarity, n_tracks,
                                                                      n_new, kp_new, score_new)
        # `if len(keypoints) == 0: return` / `if len(matched.keypoints) == 0: return` (MSCKF.py:286, 320): such a filter
        # skips the frame's bookkeeping and keeps its old last_camera_measurement
        skip = (n_new <= 0) | ((n_tracks > 0) & ((m12 >= 0).sum(1) == 0))
        action.masked_fill_(skip[:, None], 0)
        verdict = self.verify_matches(action, new_uv, epipolar_rejection_threshold, homography_rejection_threshold, counters)
        self.append_descriptors(action, m12, m21, n_tracks, n_new, desc_new, new_conf)
        self.append_observations(action, new_uv, new_conf)
        self.track_descriptors(out=self.lcm_desc, keep_old=skip.to(torch.uint8))      # MSCKF.py:436-441
        self.process_features()
        self.prune_poorest_camera_states(prune_above=max_cameras)
        return action, verdict

    def capture_frame_graph(self, samples, R_I_C, t_I_C, desc_new, kp_new, score_new, n_new, max_cameras, **kwargs):
        """frame_step_detections captured ONCE into a CUDA graph (a frame is ~60 kernel launches; for one filter or a few
        dozen the launches, not the kernels, set the frame time).  The arguments give the shapes; the returned FrameGraph
        copies each frame's inputs into its static buffers and replays the graph.  The filter state is left untouched: the
        one eager warm-up frame that sizes every lazily allocated buffer runs between snapshot() and restore()."""
        return FrameGraph(self, (samples, desc_new, kp_new, score_new, n_new), R_I_C, t_I_C, max_cameras, kwargs)

    # ------------------------------------------------------------------------------------------------------------
    def snapshot(self):
        """Every state tensor of the batch (filter state, window, tracks, descriptor lists) cloned: restore() brings the
        batch back to this point (benchmarks and A/B comparisons re-run the same frames from it)."""
        snap = dict(P=self.P.clone(), small_slab=self.small_slab.clone(), n_cams=self.n_cams.clone(),
                    tracks={n: getattr(self, n).clone() for n in _lib.TRACK_FIELDS},
                    obs_desc=None if self.obs_desc is None else self.obs_desc.clone(),
                    lcm_desc=None if self.lcm_desc is None else self.lcm_desc.clone(),
                    desc_sum=None if self.desc_sum is None else self.desc_sum.clone(),
                    cam_row=None if self.cam_row is None else self.cam_row.clone())
        return snap

    def restore(self, snap):
        self.P.copy_(snap["P"])
        self.small_slab.copy_(snap["small_slab"])
        self.n_cams.copy_(snap["n_cams"])
        for n in _lib.TRACK_FIELDS:
            getattr(self, n).copy_(snap["tracks"][n])
        if snap["obs_desc"] is not None:
            self.obs_desc.copy_(snap["obs_desc"])
            self.lcm_desc.copy_(snap["lcm_desc"])
            self.desc_sum.copy_(snap["desc_sum"])
            self.cam_row.copy_(snap["cam_row"])

    def prune_poorest_camera_states(self, max_select=2, prune_above=-1, return_map=False):
        """MSCKF.prune_poorest_camera_states (MSCKF.py:710-737) for every filter, everything resident: select the two
        cameras with the fewest observations, one last update with the tracks that observe them, then delete those
        cameras from the tracks, the window and the covariance.  Returns the (remove_mask, feat_touch) that were used."""
        mask, touch = self.select_poorest_cameras(max_select, prune_above)
        full_len = self.feat_len.clone()
        self.feat_len.mul_(touch.to(torch.int32))         # features_to_process: the other tracks sit this update out
        self.update()
        self.feat_len.copy_(full_len)
        self.drop_observations(mask)
        self.remove_cameras(mask)
        out = self.compact_tracks(drop_lost=False, return_map=return_map)   # tracks left without observations are deleted (MSCKF.py:768-772)
        return out[1] if return_map else (mask, touch)

    # ------------------------------------------------------------------------------------------------------------
    def load_cases(self, cases):
        """Fill filter b from packed case dict cases[b] (the layout of tests/golden and tools/make_golden.py)."""
        assert len(cases) == self.B
        h = pack_cases(cases, self.F, self.M, self.N, self.ldp)
        for k, v in h.items():
            getattr(self, k).copy_(torch.from_numpy(v))

    def case_outputs(self, b, case):
        """Outputs of filter b in the golden `out_*` vocabulary (numpy, trimmed to the case's true sizes)."""
        F = len(case["feat_len"])
        n = int(case["n_cams"])
        d = 15 + 6 * n
        st = self.feat_status[b, :F].cpu().numpy()
        code = st & 0xF
        imu = self.imu_state[b].cpu().numpy()
        return dict(
            out_valid=(code == _lib.FEAT_ACCEPTED) | (code == _lib.FEAT_GATED_OUT),
            out_lost=(st & _lib.FEAT_LOST_FLAG) != 0,
            out_gate=code == _lib.FEAT_ACCEPTED,
            out_gamma=self.gamma[b, :F].cpu().numpy(), out_dof=self.dof[b, :F].cpu().numpy(),
            out_idp_m=self.idp_m[b, :F].cpu().numpy(), out_idp_rho=self.idp_rho[b, :F].cpu().numpy(),
            out_delta_x=self.delta_x[b, :d].cpu().numpy(), out_P=self.P[b, :d, :d].cpu().numpy(),
            out_cam_R=self.cam_R[b, :n].cpu().numpy().reshape(n, 3, 3), out_cam_t=self.cam_t[b, :n].cpu().numpy(),
            out_imu_R=imu[:9].reshape(3, 3), out_imu_t=imu[9:12], out_imu_v=imu[12:15], out_imu_bg=imu[15:18],
            out_imu_ba=imu[18:21], out_n_rejected=int(self.n_rejected[b]), out_n_accepted=int(self.n_accepted[b]),
            out_updated=int(self.n_accepted[b]) > 0)


def pack_cases(cases, F, M, N, ldp):
    """numpy [B, ...] arrays in the C-ABI layout from a list of single-filter case dicts (padding with empty slots)."""
    B = len(cases)
    h = dict(
        P=np.zeros((B, ldp, ldp)), imu_state=np.zeros((B, 21)), cam_R=np.zeros((B, N, 9)), cam_t=np.zeros((B, N, 3)),
        n_cams=np.zeros(B, np.int32), feat_len=np.zeros((B, F), np.int32), obs_uv=np.zeros((B, F, M, 2)),
        obs_dir=np.zeros((B, F, M, 3)), obs_conf=np.zeros((B, F, M)), obs_slot=np.zeros((B, F, M), np.int32),
        idp_base=np.zeros((B, F, 3)), idp_anchor_slot=-np.ones((B, F), np.int32), idp_m=np.zeros((B, F, 3)),
        idp_rho=np.zeros((B, F)), lost_cnt=np.zeros((B, F), np.int32), tracked_cnt=np.zeros((B, F), np.int32))
    for b, c in enumerate(cases):
        n = int(c["n_cams"])
        d = 15 + 6 * n
        f = len(c["feat_len"])
        m = c["obs_uv"].shape[1] if f else 0
        assert n <= N and f <= F and m <= M and d <= ldp, (n, f, m, d)
        h["P"][b, :d, :d] = c["P"]
        h["imu_state"][b] = np.concatenate([np.asarray(c["imu_R"], float).flatten(), c["imu_t"], c["imu_v"],
                                            c["imu_bg"], c["imu_ba"]])
        h["cam_R"][b, :n] = np.asarray(c["cam_R"], float).reshape(n, 9)
        h["cam_t"][b, :n] = c["cam_t"]
        h["n_cams"][b] = n
        if f:
            h["feat_len"][b, :f] = c["feat_len"]
            h["obs_uv"][b, :f, :m] = c["obs_uv"]
            h["obs_dir"][b, :f, :m] = c["obs_dir"]
            h["obs_conf"][b, :f, :m] = c["obs_conf"]
            h["obs_slot"][b, :f, :m] = np.maximum(c["obs_slot"], 0)
            h["idp_base"][b, :f] = c["idp_base"]
            h["idp_anchor_slot"][b, :f] = c["idp_anchor_slot"]
            h["idp_m"][b, :f] = c["idp_m"]
            h["idp_rho"][b, :f] = c["idp_rho"]
            h["lost_cnt"][b, :f] = c["lost_cnt"]
            h["tracked_cnt"][b, :f] = c["tracked_cnt"]
    return h


class FrameGraph:
    """One camera frame of a BatchedMSCKF as a replayable CUDA graph (see BatchedMSCKF.capture_frame_graph)."""

    def __init__(self, flt, inputs, R_I_C, t_I_C, max_cameras, kwargs):
        self.flt = flt
        self.static = [x.clone() for x in inputs]
        self.R_I_C, self.t_I_C = np.array(R_I_C, float), np.array(t_I_C, float)     # host values, baked into the graph
        snap = flt.snapshot()
        counting, flt.count_feature_updates = flt.count_feature_updates, False
        side = torch.cuda.Stream(device=flt.device)
        side.wait_stream(torch.cuda.current_stream(flt.device))
        with torch.cuda.stream(side):                 # warm-up: lazy allocations, kernel attributes
            self._run(max_cameras, kwargs)
        torch.cuda.current_stream(flt.device).wait_stream(side)
        flt.restore(snap)
        flt.count_feature_updates = counting
        torch.cuda.synchronize(flt.device)
        self.graph = torch.cuda.CUDAGraph()
        n0 = flt.lib.msckf_launch_count()
        with torch.cuda.graph(self.graph):
            self.outputs = self._run(max_cameras, kwargs)
        self.launches_per_replay = int(flt.lib.msckf_launch_count() - n0)     # this library's kernels inside one replay
        flt.restore(snap)                             # (capture does not execute; restore() is for symmetry with a failed capture)

    def _run(self, max_cameras, kwargs):
        smp, desc, kp, sc, cnt = self.static
        return self.flt.frame_step_detections(smp, self.R_I_C, self.t_I_C, desc, kp, sc, cnt, max_cameras, **kwargs)

    def step(self, samples, desc_new, kp_new, score_new, n_new):
        """Copy one frame's inputs (device or pinned host tensors) into the static buffers and replay.  Returns the
        (action, verdict) tensors of the captured call (overwritten by the next replay)."""
        for dst, src in zip(self.static, (samples, desc_new, kp_new, score_new, n_new)):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.outputs
