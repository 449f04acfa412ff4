"""Synthetic visual-inertial streams built ON THE DEVICE for thousands of independent filters (BASELINE.json configs[2..3]:
filter g uses seed 42 + g for its trajectory and IMU-noise level g mod 3).

The per-filter trajectory parameters (a few numbers) come from a host generator; everything heavy runs on the device:
poses (torch element-wise, plumbing), IMU samples (msckf_synth_imu_batched: finite differences + white noise + random-walk
bias, the reference's generator arithmetic, csrc/synth.cu), landmark projection with pixel noise and id-ordered lists
(msckf_project_landmarks_batched), outliers / drop-outs / descriptors (torch indexing, set-up only).  The conventions are
those of synth.make_sequence (this repo's host generator the golden fixtures were made with): TartanAir-shaped camera
(640 x 480, fx 320), 20 Hz frames, 9 IMU samples per frame, camera looking along body +x."""
import numpy as np
import torch

from .dataset import project_landmarks, synth_imu
from .synth import NOISE_LEVELS, R_BODY_CAM, W_GRAVITY


def _euler_zyx(e):
    """[..., 3] yaw, pitch, roll -> [..., 3, 3] = Rz Ry Rx"""
    cy, sy = torch.cos(e[..., 0]), torch.sin(e[..., 0])
    cp, sp = torch.cos(e[..., 1]), torch.sin(e[..., 1])
    cr, sr = torch.cos(e[..., 2]), torch.sin(e[..., 2])
    R = torch.stack([cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                     sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                     -sp, cp * sr, cp * cr], dim=-1)
    return R


class DeviceStreams:
    def __init__(self, filter_indices, n_frames, n_landmarks=256, imu_per_frame=9, frame_dt=0.05, device="cuda", fx=320.0,
                 width=640, height=480, sigma_px=0.1, p_drop=0.0, p_outlier=0.02, outlier_px=60.0, desc_dim=32, desc_sigma=0.0,
                 seed=42, speed=1.0, depth_range=(6.0, 14.0)):
        dev = torch.device(device)
        idx = np.asarray(list(filter_indices), np.int64)
        B, L = len(idx), int(n_landmarks)
        self.B, self.L, self.n_frames, self.imu_per_frame, self.device = B, L, int(n_frames), int(imu_per_frame), dev
        self.K = np.array([[fx, 0.0, width / 2.0], [0.0, fx, height / 2.0], [0.0, 0.0, 1.0]])
        self.width, self.height, self.sigma_px = int(width), int(height), float(sigma_px)
        self.R_I_C, self.t_I_C = R_BODY_CAM.copy(), np.zeros(3)
        levels = list(NOISE_LEVELS.keys())
        self.noise_level = [levels[int(g) % len(levels)] for g in idx]
        dt = frame_dt / imu_per_frame
        T = self.n_frames * imu_per_frame
        # ---- per-filter trajectory parameters (host, a few numbers per filter) ----
        par = np.zeros((B, 6, 3))
        sig = np.zeros((B, 4))
        for b, g in enumerate(idx):
            rng = np.random.default_rng(seed + int(g))
            par[b, 0] = rng.uniform([0.5, 0.7, 0.3], [0.9, 1.1, 0.5])              # position amplitudes
            par[b, 1] = rng.uniform(0.5, 0.9, 3) * speed                            # position rates
            par[b, 2] = rng.uniform(0, 2 * np.pi, 3)
            par[b, 3] = rng.uniform([0.10, 0.05, 0.03], [0.18, 0.09, 0.06])        # attitude amplitudes
            par[b, 4] = rng.uniform(0.3, 0.7, 3) * speed
            par[b, 5] = rng.uniform(0, 2 * np.pi, 3)
            sa, sg, sba, sbg = NOISE_LEVELS[self.noise_level[b]]
            sig[b] = [sa, sg, sba, sbg]
        par_d = torch.from_numpy(par).to(dev)
        self.sigmas = torch.from_numpy(sig).to(dev)
        # continuous_noise_covariance diagonal blocks (MSCKF.py:99-103): gyro, gyro walk, acc, acc walk
        self.Qc_diag = torch.from_numpy(np.stack([sig[:, 1] ** 2, sig[:, 3] ** 2, sig[:, 0] ** 2, sig[:, 2] ** 2], 1)).to(dev)
        # ---- poses at IMU rate: Lissajous position + small attitude oscillation, ramped in from rest ----
        ts = torch.arange(T + 1, dtype=torch.float64, device=dev) * dt
        x = ((ts - frame_dt) / 1.0).clamp(0.0, 1.0)
        ramp = x * x * x * (x * (6 * x - 15) + 10)
        tt = ts[None, :, None]
        pos = ramp[None, :, None] * par_d[:, None, 0] * (torch.sin(par_d[:, None, 1] * tt + par_d[:, None, 2]) - torch.sin(par_d[:, None, 2]))
        eul = ramp[None, :, None] * par_d[:, None, 3] * (torch.sin(par_d[:, None, 4] * tt + par_d[:, None, 5]) - torch.sin(par_d[:, None, 5]))
        self.times = ts[None].expand(B, -1).contiguous()
        self.poses = torch.cat([_euler_zyx(eul), pos], dim=-1).contiguous()       # [B, T + 1, 12]
        # ---- IMU: finite differences + white noise + accumulated random-walk bias (csrc/synth.cu) ----
        out = synth_imu(self.poses, self.times, interp=1, gravity=W_GRAVITY, sigmas=self.sigmas, seed=seed, accumulate_bias=True,
                        sign_fix=True, want_poses=False)
        self.imu = out["imu_noisy"]                                               # [B, T + 1, 7]; row k = sample k (row 0 unused)
        # ---- frames: frame f is delivered after IMU sample (f + 1) * imu_per_frame ----
        fidx = (torch.arange(self.n_frames, device=dev) + 1) * imu_per_frame
        self.frame_sample = fidx
        self.frame_pose = self.poses[:, fidx].contiguous()                        # [B, n_frames, 12] body poses
        # ---- landmarks that stay inside the image for every frame (first L of 6 L candidates per filter) ----
        gen = torch.Generator(device=dev).manual_seed(1000 + seed + int(idx[0]))
        C = 6 * L
        Rbc = torch.from_numpy(R_BODY_CAM).to(dev)
        lm = torch.empty((B, L, 3), dtype=torch.float64, device=dev)
        ch = max(1, min(B, 64))
        for b0 in range(0, B, ch):
            nb = min(ch, B - b0)
            depth = depth_range[0] + (depth_range[1] - depth_range[0]) * torch.rand((nb, C), dtype=torch.float64, device=dev, generator=gen)
            a1 = -0.62 + 1.24 * torch.rand((nb, C), dtype=torch.float64, device=dev, generator=gen)
            a2 = -0.48 + 0.96 * torch.rand((nb, C), dtype=torch.float64, device=dev, generator=gen)
            cand = torch.stack([depth, depth * torch.tan(a1), depth * torch.tan(a2)], dim=-1)          # body frame at rest (x ahead)
            Rb = self.frame_pose[b0:b0 + nb, :, :9].reshape(nb, -1, 3, 3)
            tb = self.frame_pose[b0:b0 + nb, :, 9:]
            Rc = Rb @ Rbc                                                                               # camera rotations
            ok = torch.ones((nb, C), dtype=torch.bool, device=dev)
            for f0 in range(0, self.n_frames, 8):                                                       # frames in groups: memory
                d = cand[:, None, :, :] - tb[:, f0:f0 + 8, None, :]                                     # [nb, f, C, 3]
                c = torch.einsum("bfij,bfci->bfcj", Rc[:, f0:f0 + 8], d)
                u = fx * c[..., 0] / c[..., 2] + width / 2.0
                v = fx * c[..., 1] / c[..., 2] + height / 2.0
                ok &= ((c[..., 2] > 0.5) & (u > 8) & (u < width - 8) & (v > 8) & (v < height - 8)).all(dim=1)
            order = torch.argsort((~ok).to(torch.int8), dim=1, stable=True)[:, :L]
            lm[b0:b0 + nb] = torch.gather(cand, 1, order[..., None].expand(-1, -1, 3))
        self.landmarks = lm
        # ---- detections: projection + pixel noise on the device, then outliers / drop-outs ----
        kp, ids, cnt = project_landmarks(self.frame_pose, self.R_I_C, self.t_I_C, self.K, width, height, lm, sigma_px=sigma_px,
                                         seed=seed + 7)
        keep = torch.arange(L, device=dev)[None, None, :] < cnt[..., None]
        if p_outlier > 0:
            bad = torch.rand((B, self.n_frames, L), device=dev, generator=gen) < p_outlier
            kp = kp + (bad & keep)[..., None] * outlier_px * torch.randn((B, self.n_frames, L, 2), dtype=torch.float64, device=dev, generator=gen)
        inside = (kp[..., 0] >= 0) & (kp[..., 0] < width) & (kp[..., 1] >= 0) & (kp[..., 1] < height)
        keep &= inside
        if p_drop > 0:
            drop = torch.rand((B, self.n_frames, L), device=dev, generator=gen) < p_drop
            drop[:, 0] = False
            keep &= ~drop
        order = torch.argsort((~keep).to(torch.int8), dim=2, stable=True)
        self.kp = torch.gather(kp, 2, order[..., None].expand(-1, -1, -1, 2)).contiguous()
        self.ids = torch.gather(ids, 2, order).contiguous()
        self.count = keep.sum(dim=2).to(torch.int32).contiguous()
        self.scores = (0.7 + 0.3 * torch.rand((B, self.n_frames, L), dtype=torch.float64, device=dev, generator=gen)).contiguous()
        d = torch.randn((B, L, desc_dim), device=dev, generator=gen)
        self.desc = torch.nn.functional.normalize(d, dim=2).contiguous()          # one unit descriptor per landmark (float32)
        self.desc_sigma = float(desc_sigma)
        self._gen = gen

    def frame_inputs(self, f):
        """Frame f as the filters receive it: IMU samples since the previous frame [B, n, 7], the frame's detections
        (descriptors [B, L, D] float32, keypoints [B, L, 2], scores [B, L], count [B] int32)."""
        n = self.imu_per_frame
        k1 = int(self.frame_sample[f])
        samples = self.imu[:, k1 - n + 1:k1 + 1].contiguous()
        ids = self.ids[:, f].clamp(min=0).long()
        desc = torch.gather(self.desc, 1, ids[..., None].expand(-1, -1, self.desc.shape[2]))
        if self.desc_sigma > 0:
            desc = torch.nn.functional.normalize(desc + self.desc_sigma * torch.randn(desc.shape, device=self.device, generator=self._gen), dim=2)
        return samples, desc.contiguous(), self.kp[:, f].contiguous(), self.scores[:, f].contiguous(), self.count[:, f].contiguous()

    def gt_pose(self, f):
        """[B, 12] ground-truth body pose at frame f (for the metrics kernels)."""
        return self.frame_pose[:, f].contiguous()
