"""Monte-Carlo runs of whole visual-inertial sequences with every filter resident on the device.

The data-parallel unit of the path is the filter instance: B independent filters, one synthetic stream
each (`synth.make_sequence(seed=42 + index, noise_level=index mod 3)`), advance frame by frame through
`BatchedMSCKF.frame_step`.  What stays on the host is only what the scope contract leaves there -- the association of the
frame's detections with the existing tracks (the reference's descriptor matcher, MSCKF.py:268-331; with the synthetic
streams the ground-truth landmark ids give the same association) -- and the bookkeeping of which landmark each track slot
follows, kept aligned with the device through the slot maps `frame_step` returns.

Per filter the run yields the trajectory at camera rate and its absolute trajectory error against the ground truth
(main.py:334-391 plots the same quantity); ranks gather them with one collective at the end (`sharding.gather_filter_results`).
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

from . import synth
from .batched import BatchedMSCKF, UpdateParams


class TrackAssociator:
    """Host side of add_camera_measurements for ONE filter: which landmark every track slot follows (dict order of the
    reference's `features`), turned into the action / new_uv / new_conf arrays of msckf_append_observations_batched."""

    def __init__(self):
        self.slots: List[int] = []          # landmark id per track slot, in the reference's dict order

    def frame(self, landmark_ids, keypoints, scores):
        """-> (action [n], new_uv [n, 2], new_conf [n]) for the n track slots after this frame.  Matched tracks get their
        detection (action 2), tracks whose landmark was not detected are lost this frame (1), detections of unknown
        landmarks open new tracks at the end, in detection order (3) -- MSCKF.py:333-434."""
        landmark_ids = np.asarray(landmark_ids)
        scores = np.asarray(scores, float)
        keep = scores >= 0.5 * scores.mean() if len(scores) else np.zeros(0, bool)      # MSCKF.py:281-284
        det = {int(l): i for i, l in enumerate(landmark_ids) if keep[i]}
        had_tracks = len(self.slots) > 0
        any_match = any(l in det for l in self.slots)
        n_old = len(self.slots)
        if had_tracks and not any_match:                      # no match at all: the reference returns early (MSCKF.py:316)
            return np.zeros(n_old, np.uint8), np.zeros((n_old, 2)), np.zeros(n_old)
        known = set(self.slots)
        new = [int(l) for i, l in enumerate(landmark_ids) if keep[i] and int(l) not in known]
        self.slots = self.slots + new
        n = len(self.slots)
        action = np.zeros(n, np.uint8)
        uv = np.zeros((n, 2))
        conf = np.zeros(n)
        for j, l in enumerate(self.slots):
            i = det.get(l)
            if i is None:
                action[j] = 1
            else:
                action[j] = 2 if j < n_old else 3
                uv[j] = keypoints[i]
                conf[j] = scores[i]
        return action, uv, conf

    def apply_map(self, dest):
        """dest[j] = new slot of old slot j, -1 = deleted (msckf_compact_tracks_batched)."""
        out = [None] * int((np.asarray(dest[:len(self.slots)]) >= 0).sum())
        for j, l in enumerate(self.slots):
            d = int(dest[j])
            if d >= 0:
                out[d] = l
        self.slots = out


@dataclass
class MonteCarloResult:
    positions: np.ndarray                   # [B, n_frames, 3] estimated IMU positions after every camera frame
    gt_positions: np.ndarray                # [B, n_frames, 3]
    ate_rmse: np.ndarray                    # [B] root-mean-square absolute translation error (m)
    n_tracks: np.ndarray                    # [B, n_frames] live tracks after every frame
    n_cams: np.ndarray                      # [B, n_frames]
    seeds: List[int] = field(default_factory=list)


def run_monte_carlo(filter_indices, n_frames=30, n_landmarks=48, window=8, device="cuda", imu_per_frame=9,
                    min_parallax=2.0, p_drop=0.05, p_outlier=0.02, matcher="host", min_cosine_similarity=0.95,
                    epipolar_rejection_threshold=1e9, homography_rejection_threshold=1e9, window_slack=2):
    """Run the filters `filter_indices` (global indices: seed 42 + index, IMU-noise level index mod 3) for n_frames camera
    frames, all of them resident on `device`.  matcher = "host": the association of detections with tracks comes from the
    streams' ground-truth landmark ids (TrackAssociator); "device": the host only uploads each frame's detections
    (descriptors, keypoints, scores) and the mutual-NN matcher, the epipolar / homography tests and the bookkeeping run
    on the device (BatchedMSCKF.frame_step_detections, MSCKF.add_camera_measurements :284-448).
    The window capacity is window + 1 (the frame's own pose before pruning) + window_slack: a camera added on a frame that
    produced no valid feature stays in the reference's window too (it only prunes cameras some track observes), so a run
    of such frames needs the slack; an augmentation that finds the window full is an error (check_window_overflow)."""
    assert matcher in ("host", "device")
    levels = list(synth.NOISE_LEVELS.keys())
    seqs = [synth.make_sequence(seed=42 + int(i), n_frames=n_frames, n_landmarks=n_landmarks, imu_per_frame=imu_per_frame,
                                noise_level=levels[int(i) % len(levels)], p_drop=p_drop, p_outlier=p_outlier, shuffle=True)
            for i in filter_indices]
    B = len(seqs)
    s0 = seqs[0]
    par = UpdateParams(K=s0.K, width=s0.width, height=s0.height, sigma_image=0.1, min_lost=2, min_tracked=3, use_parallax=True,
                       min_parallax=min_parallax)
    from . import _lib
    par.compress_mode = _lib.COMPRESS_GRAM
    F, M, N = n_landmarks + 8, window + 2 + window_slack, window + 1 + window_slack
    flt = BatchedMSCKF(B, F, M, N, par, device=device)
    # initial state as the reference's IMU.initialize leaves it for a stream that starts at rest, level (IMU.py:43-60)
    flt.P.zero_()
    flt.imu_state.zero_()
    flt.imu_state[:, [0, 4, 8]] = 1.0
    flt.imu_time.zero_()
    flt.n_cams.zero_()
    flt.feat_len.zero_()
    qc = np.zeros((B, 4))
    for b, i in enumerate(filter_indices):
        sa, sg, sba, sbg = synth.NOISE_LEVELS[levels[int(i) % len(levels)]]
        qc[b] = [sg ** 2, sbg ** 2, sa ** 2, sba ** 2]          # continuous_noise_covariance diagonal blocks (MSCKF.py:99-103)
    flt.Qc_diag.copy_(torch.from_numpy(qc))
    R_I_C, t_I_C = synth.R_BODY_CAM, np.zeros(3)
    assoc = [TrackAssociator() for _ in range(B)]
    pos = np.zeros((B, n_frames, 3))
    gtp = np.zeros((B, n_frames, 3))
    ntr = np.zeros((B, n_frames), np.int32)
    ncm = np.zeros((B, n_frames), np.int32)
    prev = [0] * B
    for f in range(n_frames):
        n_smp = max(sq.frames[f].imu_index + 1 - prev[b] for b, sq in enumerate(seqs))
        samples = np.zeros((B, n_smp, 7))
        action = np.zeros((B, F), np.uint8)
        new_uv = np.zeros((B, F, 2))
        new_conf = np.zeros((B, F))
        for b, sq in enumerate(seqs):
            fr = sq.frames[f]
            rows = sq.imu[prev[b]:fr.imu_index + 1]
            assert len(rows) == n_smp, "every stream delivers the same number of IMU samples per frame"
            samples[b] = rows
            prev[b] = fr.imu_index + 1
            a, uv, cf = assoc[b].frame(fr.landmark_ids, fr.keypoints, fr.scores)
            assert len(a) <= F, "track capacity exceeded"
            action[b, :len(a)], new_uv[b, :len(a)], new_conf[b, :len(a)] = a, uv, cf
            gtp[b, f] = sq.gt_p[fr.imu_index + 1]
        if matcher == "device":
            F2 = max(len(sq.frames[f].keypoints) for sq in seqs)
            D = seqs[0].frames[f].descriptors.shape[1]
            desc = np.zeros((B, F2, D), np.float32)
            kp = np.zeros((B, F2, 2))
            sc = np.zeros((B, F2))
            n_new = np.zeros(B, np.int32)
            for b, sq in enumerate(seqs):
                fr = sq.frames[f]
                n = len(fr.scores)                        # RAW detections: the score filter of MSCKF.py:279-284 runs on the device
                desc[b, :n], kp[b, :n], sc[b, :n], n_new[b] = fr.descriptors, fr.keypoints, fr.scores, n
            dv = flt.device
            flt.frame_step_detections(torch.from_numpy(samples).to(dv), R_I_C, t_I_C, torch.from_numpy(desc).to(dv),
                                      torch.from_numpy(kp).to(dv), torch.from_numpy(sc).to(dv), torch.from_numpy(n_new).to(dv),
                                      window, min_cosine_similarity=min_cosine_similarity,
                                      epipolar_rejection_threshold=epipolar_rejection_threshold,
                                      homography_rejection_threshold=homography_rejection_threshold, score_filter=True)
            pos[:, f] = flt.imu_state[:, 9:12].cpu().numpy()
            ntr[:, f] = (flt.feat_len > 0).sum(dim=1).cpu().numpy()
            ncm[:, f] = flt.n_cams.cpu().numpy()
            continue
        m1, m2 = flt.frame_step(torch.from_numpy(samples).to(flt.device), R_I_C, t_I_C, torch.from_numpy(action).to(flt.device),
                                torch.from_numpy(new_uv).to(flt.device), torch.from_numpy(new_conf).to(flt.device), window,
                                use_null_state=False, return_maps=True)
        m1, m2 = m1.cpu().numpy(), m2.cpu().numpy()
        for b in range(B):
            assoc[b].apply_map(m1[b])
            assoc[b].apply_map(m2[b])
        pos[:, f] = flt.imu_state[:, 9:12].cpu().numpy()
        ntr[:, f] = (flt.feat_len > 0).sum(dim=1).cpu().numpy()
        ncm[:, f] = flt.n_cams.cpu().numpy()
        for b in range(B):
            assert ntr[b, f] == len(assoc[b].slots), "host track records out of step with the device"
    flt.check_window_overflow()
    ate = np.sqrt(np.mean(np.sum((pos - gtp) ** 2, axis=2), axis=1))
    return MonteCarloResult(pos, gtp, ate, ntr, ncm, [42 + int(i) for i in filter_indices])
