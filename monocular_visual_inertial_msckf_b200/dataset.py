"""Data synthesis and metrics on the device: host wrappers of the kernels in csrc/synth.cu.

Mirrors, batched over B independent streams, what the reference's data tools and driver do on the host:
    synth_imu()            <-> PhotorealisticGenerator.process_data, dataset/tools/dataset_generators/photorealistic_generator.py:145-216
                               (interp=10) and MeasurementsGenerator.generate, synthetic_generator.py:145-186 (interp=1)
    project_landmarks()    <-> MeasurementsGenerator.generate's camera loop, synthetic_generator.py:189-223
    MetricsAccumulator     <-> the error block of main.py:312-391 and its summary :440-445
PyTorch is used for device memory and streams only."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise ValueError("expected contiguous CUDA tensors (there is no CPU fallback)")


def synth_imu(key_poses, key_times, interp=10, gravity=(0.0, 0.0, -9.81), noise=None, sigmas=None, seed=0,
              accumulate_bias=False, sign_fix=True, want_poses=True):
    """key_poses [B, K, 12] (R row-major, t), key_times [B, K] float64 on the device.  Returns dict(poses [B, T+1, 12],
    times [B, T+1], imu_gt [B, T+1, 7], imu_noisy [B, T+1, 7]); rows are (timestamp, w, a), row 0 the reference's zero row.
    noise [B, T, 12] already scaled (acc white, acc walk, gyro white, gyro walk), or sigmas [B, 4] (acc, gyro, acc walk,
    gyro walk) for draws made on the device, or neither (noisy = gt)."""
    lib = _lib.load_library()
    _need_cuda(key_poses, key_times, noise, sigmas)
    B, K = int(key_poses.shape[0]), int(key_poses.shape[1])
    assert key_poses.dtype == torch.float64 and tuple(key_poses.shape) == (B, K, 12) and tuple(key_times.shape) == (B, K)
    sub = interp - 1 if interp > 1 else 1
    T = sub * (K - 1)
    dev = key_poses.device
    if noise is not None:
        assert tuple(noise.shape) == (B, T, 12) and noise.dtype == torch.float64
    if sigmas is not None:
        assert tuple(sigmas.shape) == (B, 4) and sigmas.dtype == torch.float64
    poses = torch.empty((B, T + 1, 12), dtype=torch.float64, device=dev) if want_poses else None
    times = torch.empty((B, T + 1), dtype=torch.float64, device=dev) if want_poses else None
    imu_gt = torch.empty((B, T + 1, 7), dtype=torch.float64, device=dev)
    imu_noisy = torch.empty((B, T + 1, 7), dtype=torch.float64, device=dev)
    g = (C.c_double * 3)(*[float(x) for x in gravity])
    rc = lib.msckf_synth_imu_batched(B, K, int(interp), _ptr(key_poses), _ptr(key_times), g, _ptr(noise), _ptr(sigmas),
                                     C.c_ulonglong(int(seed)), int(bool(accumulate_bias)), int(bool(sign_fix)), _ptr(poses),
                                     _ptr(times), _ptr(imu_gt), _ptr(imu_noisy), _stream(dev))
    _lib.check(rc, "msckf_synth_imu_batched")
    return dict(poses=poses, times=times, imu_gt=imu_gt, imu_noisy=imu_noisy)


def project_landmarks(body_poses, R_I_C, t_I_C, K, width, height, landmarks, noise_uv=None, sigma_px=0.0, seed=0,
                      want_scores=False):
    """body_poses [B, n_frames, 12], landmarks [B, L, 3] float64 on the device.  Returns (kp [B, n_frames, L, 2],
    ids [B, n_frames, L] int32, count [B, n_frames] int32[, scores]): per frame the visible landmarks in id order."""
    lib = _lib.load_library()
    _need_cuda(body_poses, landmarks, noise_uv)
    B, nf = int(body_poses.shape[0]), int(body_poses.shape[1])
    L = int(landmarks.shape[1])
    assert tuple(body_poses.shape) == (B, nf, 12) and tuple(landmarks.shape) == (B, L, 3)
    dev = body_poses.device
    if noise_uv is not None:
        assert tuple(noise_uv.shape) == (B, nf, L, 2) and noise_uv.dtype == torch.float64
    kp = torch.zeros((B, nf, L, 2), dtype=torch.float64, device=dev)
    ids = torch.full((B, nf, L), -1, dtype=torch.int32, device=dev)
    cnt = torch.zeros((B, nf), dtype=torch.int32, device=dev)
    sc = torch.zeros((B, nf, L), dtype=torch.float64, device=dev) if want_scores else None
    T = (C.c_double * 12)(*[float(x) for x in np.asarray(R_I_C, float).flatten()], *[float(x) for x in np.asarray(t_I_C, float).flatten()])
    Kc = (C.c_double * 9)(*[float(x) for x in np.asarray(K, float).flatten()])
    rc = lib.msckf_project_landmarks_batched(B, L, nf, _ptr(body_poses), T, Kc, int(width), int(height), _ptr(landmarks),
                                             _ptr(noise_uv), float(sigma_px), C.c_ulonglong(int(seed)), _ptr(kp), _ptr(ids),
                                             _ptr(cnt), _ptr(sc), _stream(dev))
    _lib.check(rc, "msckf_project_landmarks_batched")
    return (kp, ids, cnt, sc) if want_scores else (kp, ids, cnt)


class MetricsAccumulator:
    """The reference driver's error block (main.py:312-391) for B filters, one step() per loop iteration of that driver."""

    RECORD = ("rte", "roe", "ate_x", "ate_y", "ate_z", "aoe_roll", "aoe_pitch", "aoe_yaw", "x_bound", "y_bound", "z_bound",
              "roll_bound", "pitch_bound", "yaw_bound")
    SUMMARY = ("rmse_position", "rmse_orientation", "mean_rte", "std_rte", "mean_roe", "std_roe", "ate_rmse", "n")

    def __init__(self, B, device="cuda"):
        self.lib = _lib.load_library()
        self.B = int(B)
        self.device = torch.device(device)
        self.state = torch.zeros((self.B, self.lib.msckf_metrics_state_doubles()), dtype=torch.float64, device=self.device)

    def step(self, gt_pose, flt, record=False):
        """gt_pose [B, 12] device tensor (R row-major, t) of this step; flt: the BatchedMSCKF holding the estimates."""
        _need_cuda(gt_pose)
        assert tuple(gt_pose.shape) == (self.B, 12) and flt.B == self.B
        rec = torch.empty((self.B, 14), dtype=torch.float64, device=self.device) if record else None
        rc = self.lib.msckf_metrics_step_batched(C.byref(flt.dims), _ptr(gt_pose), _ptr(flt.imu_state), _ptr(flt.P), _ptr(self.state),
                                                 _ptr(rec), _stream(self.device))
        _lib.check(rc, "msckf_metrics_step_batched")
        return rec

    def summary(self):
        """[B, 8] device tensor: SUMMARY columns (what the final gather of a Monte-Carlo run carries)."""
        out = torch.empty((self.B, 8), dtype=torch.float64, device=self.device)
        rc = self.lib.msckf_metrics_finalize_batched(self.B, _ptr(self.state), _ptr(out), _stream(self.device))
        _lib.check(rc, "msckf_metrics_finalize_batched")
        return out
