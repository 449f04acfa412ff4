#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference (imported from
/root/reference through tools/ref_shim.py) on this repo's synthetic streams.

Run HERE (CPU container; the GPU box has no /root/reference):   python tools/make_golden.py [--only NAME]

For every reference `get_valid_features` + `update` pair (the unit the C-ABI call msckf_update_batched replaces,
reference src/msckf/MSCKF.py:450-456, 733-735) an "update case" is captured: the packed inputs exactly as the
reference's objects held them on entry, and everything the call produced (validity/lost flags, re-triangulated
inverse-depth parameters, per-feature gamma/dof/gate decision, delta_x, P+, corrected IMU state and camera poses).
For `process_imu` (MSCKF.py:160-248) "propagate cases" hold (state, null state, P, sample) -> (state, P).
Trajectory fixtures hold the per-frame state of a whole run.

The captures are taken by wrapping bound methods on the reference instance; the reference source is not modified.
"""
import argparse
import os
import sys
import time

import numpy as np
from scipy.stats import chi2

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402
from monocular_visual_inertial_msckf_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pack_update_inputs(ns, f, features):
    """Pack the reference's objects into the SoA layout of include/msckf_b200.h (one filter)."""
    cams = list(f.state.cameras.items())
    cam_ids = [k for k, _ in cams]
    N = len(cams)
    F = len(features)
    M = max([len(ft.keypoints) for ft in features.values()] + [1])
    c = dict(
        K=np.array(f.K, float), Kinv=np.linalg.inv(f.K), width=f.width, height=f.height, sigma=f.sigma_image,
        gravity=np.array(f.state.imu.W_gravity, float), min_lost=f.min_number_of_frames_to_be_lost,
        min_tracked=f.min_number_of_frames_to_be_tracked, use_parallax=int(f.use_parallax),
        min_parallax=float(f.min_parallax), n_cams=N,
        cam_R=np.stack([cam.T_W_Ci.R for _, cam in cams]).astype(float) if N else np.zeros((0, 3, 3)),
        cam_t=np.stack([cam.T_W_Ci.t for _, cam in cams]).astype(float) if N else np.zeros((0, 3)),
        imu_R=np.array(f.state.imu.T_W_Ii.R, float), imu_t=np.array(f.state.imu.T_W_Ii.t, float),
        imu_v=np.array(f.state.imu.v_W_Ii, float), imu_bg=np.array(f.state.imu.gyroscope_bias, float),
        imu_ba=np.array(f.state.imu.accelerometer_bias, float),
        P=np.array(f.state.covariance, float),
        feat_ids=np.array(list(features.keys()), np.int64),
        feat_len=np.zeros(F, np.int32), obs_uv=np.zeros((F, M, 2)), obs_dir=np.zeros((F, M, 3)),
        obs_conf=np.zeros((F, M)), obs_slot=-np.ones((F, M), np.int32),
        idp_base=np.zeros((F, 3)), idp_anchor_slot=-np.ones(F, np.int32), idp_m=np.zeros((F, 3)),
        idp_rho=np.zeros(F), lost_cnt=np.zeros(F, np.int32), tracked_cnt=np.zeros(F, np.int32),
    )
    for j, ft in enumerate(features.values()):
        n = len(ft.keypoints)
        c["feat_len"][j] = n
        for i in range(n):
            c["obs_uv"][j, i] = ft.keypoints[i]
            c["obs_dir"][j, i] = ft.lines[i].direction
            c["obs_conf"][j, i] = ft.lines[i].confidence
            slot = cam_ids.index(ft.camera_indices[i])
            c["obs_slot"][j, i] = slot
            # Line.base is the camera's live position array (MSCKF.py:305,410)
            assert ft.lines[i].base is f.state.cameras[ft.camera_indices[i]].T_W_Ci.t
        idp = ft.inverse_depth_point
        c["idp_base"][j] = idp.base
        for s, (_, cam) in enumerate(cams):
            if idp.base is cam.T_W_Ci.t:
                c["idp_anchor_slot"][j] = s
        c["idp_m"][j] = idp.m
        c["idp_rho"][j] = idp.rho
        c["lost_cnt"][j] = ft.lost_for_n_frames
        c["tracked_cnt"][j] = ft.tracked_for_n_frames
    return c


class Capture:
    def __init__(self, ns, f, keep_update=lambda i: True, keep_propagate=lambda i: False):
        self.ns, self.f = ns, f
        self.update_cases = []
        self.propagate_cases = []
        self.keep_update, self.keep_propagate = keep_update, keep_propagate
        self.n_update_calls = 0
        self.n_propagate_calls = 0
        self._cur = None
        self._orig_gvf = f.get_valid_features
        self._orig_update = f.update
        self._orig_gate = f.gating_test
        self._orig_correct = f.correct
        self._orig_process_imu = f.process_imu
        f.get_valid_features = self.get_valid_features
        f.update = self.update
        f.gating_test = self.gating_test
        f.correct = self.correct
        f.process_imu = self.process_imu

    # -- update path ---------------------------------------------------------------------------------------------
    def get_valid_features(self, features):
        f = self.f
        idx = self.n_update_calls
        self.n_update_calls += 1
        keep = self.keep_update(idx)
        if keep:
            c = pack_update_inputs(self.ns, f, features)
            c["call_index"] = idx
        valid, lost = self._orig_gvf(features)
        if keep:
            keys = list(features.keys())
            F = len(keys)
            c["out_valid"] = np.array([k in valid for k in keys], bool)
            c["out_lost"] = np.array([k in lost for k in keys], bool)
            c["out_idp_m"] = np.stack([features[k].inverse_depth_point.m for k in keys]) if F else np.zeros((0, 3))
            c["out_idp_rho"] = np.array([features[k].inverse_depth_point.rho for k in keys], float)
            c["out_gamma"] = np.full(F, np.nan)
            c["out_dof"] = np.zeros(F, np.int32)
            c["out_gate"] = np.zeros(F, bool)
            c["out_delta_x"] = np.zeros(f.state.covariance.shape[0])
            c["out_updated"] = False
            self._cur = c
            self._valid_keys = [k for k in keys if k in valid]
            self._gate_i = 0
            self._keys = keys
            self._n_rej0 = f.number_of_residuals_discarded_for_gasting_test
            if len(valid) == 0:
                self._finish()
        return valid, lost

    def gating_test(self, r, H):
        ok = self._orig_gate(r, H)
        c = self._cur
        if c is not None:
            f = self.f
            S = H @ f.state.covariance @ H.T + f.sigma_image ** 2 * np.eye(H.shape[0])
            gamma = float((r.T @ np.linalg.inv(S) @ r).flatten()[0]) if H.shape[0] > 0 else 0.0
            j = self._keys.index(self._valid_keys[self._gate_i])
            c["out_gamma"][j] = gamma
            c["out_dof"][j] = r.shape[0]
            c["out_gate"][j] = bool(ok)
            self._gate_i += 1
        return ok

    def correct(self, K, T_H, R_n, delta_x):
        if self._cur is not None:
            self._cur["out_delta_x"] = np.array(delta_x, float).flatten()
            self._cur["out_updated"] = True
        return self._orig_correct(K, T_H, R_n, delta_x)

    def update(self, features):
        out = self._orig_update(features)
        if self._cur is not None:
            self._finish()
        return out

    def _finish(self):
        f, c = self.f, self._cur
        cams = list(f.state.cameras.values())
        N = len(cams)
        c["out_P"] = np.array(f.state.covariance, float)
        c["out_cam_R"] = np.stack([cam.T_W_Ci.R for cam in cams]).astype(float) if N else np.zeros((0, 3, 3))
        c["out_cam_t"] = np.stack([cam.T_W_Ci.t for cam in cams]).astype(float) if N else np.zeros((0, 3))
        c["out_imu_R"] = np.array(f.state.imu.T_W_Ii.R, float)
        c["out_imu_t"] = np.array(f.state.imu.T_W_Ii.t, float)
        c["out_imu_v"] = np.array(f.state.imu.v_W_Ii, float)
        c["out_imu_bg"] = np.array(f.state.imu.gyroscope_bias, float)
        c["out_imu_ba"] = np.array(f.state.imu.accelerometer_bias, float)
        c["out_n_rejected"] = f.number_of_residuals_discarded_for_gasting_test - self._n_rej0
        self.update_cases.append(c)
        self._cur = None

    # -- propagation path ----------------------------------------------------------------------------------------
    def process_imu(self, m):
        f = self.f
        idx = self.n_propagate_calls
        self.n_propagate_calls += 1
        keep = self.keep_propagate(idx)
        if keep:
            imu = f.state.imu
            c = dict(call_index=idx, sample=np.concatenate([[m.timestamp], m.angular_velocity, m.linear_acceleration]),
                     timestamp=imu.timestamp, gravity=np.array(imu.W_gravity, float),
                     Qc_diag=np.diag(f.continuous_noise_covariance).copy(),
                     imu_R=np.array(imu.T_W_Ii.R, float), imu_t=np.array(imu.T_W_Ii.t, float),
                     imu_v=np.array(imu.v_W_Ii, float), imu_bg=np.array(imu.gyroscope_bias, float),
                     imu_ba=np.array(imu.accelerometer_bias, float),
                     null_R=np.array(imu.T_W_Ii_null.R, float), null_t=np.array(imu.T_W_Ii_null.t, float),
                     null_v=np.array(imu.v_W_Ii_null, float), P=np.array(f.state.covariance, float))
        out = self._orig_process_imu(m)
        if keep:
            imu = f.state.imu
            c.update(out_imu_R=np.array(imu.T_W_Ii.R, float), out_imu_t=np.array(imu.T_W_Ii.t, float),
                     out_imu_v=np.array(imu.v_W_Ii, float), out_P=np.array(f.state.covariance, float))
            self.propagate_cases.append(c)
        return out


def pack_tracks(f):
    """All tracks of the filter in the SoA layout (dict order), slots = position in the insertion-ordered camera dict."""
    cams = list(f.state.cameras.items())
    cam_ids = [k for k, _ in cams]
    feats = f.features
    F = len(feats)
    M = max([len(ft.keypoints) for ft in feats.values()] + [1])
    c = dict(n_cams=len(cams), feat_ids=np.array(list(feats.keys()), np.int64), feat_len=np.zeros(F, np.int32),
             obs_uv=np.zeros((F, M, 2)), obs_dir=np.zeros((F, M, 3)), obs_conf=np.zeros((F, M)),
             obs_slot=-np.ones((F, M), np.int32), idp_base=np.zeros((F, 3)), idp_anchor_slot=-np.ones(F, np.int32),
             cam_t=np.stack([cam.T_W_Ci.t for _, cam in cams]).astype(float) if cams else np.zeros((0, 3)),
             cam_R=np.stack([cam.T_W_Ci.R for _, cam in cams]).astype(float) if cams else np.zeros((0, 3, 3)),
             idp_m=np.zeros((F, 3)), idp_rho=np.zeros(F), lost_cnt=np.zeros(F, np.int32), tracked_cnt=np.zeros(F, np.int32))
    for j, ft in enumerate(feats.values()):
        n = len(ft.keypoints)
        c["feat_len"][j] = n
        for i in range(n):
            c["obs_uv"][j, i] = ft.keypoints[i]
            c["obs_dir"][j, i] = ft.lines[i].direction
            c["obs_conf"][j, i] = ft.lines[i].confidence
            c["obs_slot"][j, i] = cam_ids.index(ft.camera_indices[i])
        c["idp_base"][j] = ft.inverse_depth_point.base
        for s, (_, cam) in enumerate(cams):
            if ft.inverse_depth_point.base is cam.T_W_Ci.t:
                c["idp_anchor_slot"][j] = s
        c["idp_m"][j] = ft.inverse_depth_point.m
        c["idp_rho"][j] = ft.inverse_depth_point.rho
        c["lost_cnt"][j] = ft.lost_for_n_frames
        c["tracked_cnt"][j] = ft.tracked_for_n_frames
    return c


class SurgeryCapture:
    """prune_poorest_camera_states (selection, MSCKF.py:710-727) and remove_cameras (track part, MSCKF.py:751-770)."""

    def __init__(self, f):
        self.f = f
        self.select_cases, self.drop_cases = [], []
        self._orig_prune, self._orig_remove = f.prune_poorest_camera_states, f.remove_cameras
        self._orig_gvf = f.get_valid_features
        self._in_prune = None
        f.prune_poorest_camera_states = self.prune
        f.remove_cameras = self.remove
        f.get_valid_features = self.gvf
        self.append_cases = []
        self._orig_add = f.add_camera_measurements
        f.add_camera_measurements = self.add

    def add(self, image, extracted_features=None):
        """add_camera_measurements (MSCKF.py:268-448): the matching is out of scope; what it did to the tracks is
        recovered by comparing them before / after: action 0 untouched, 1 lost this frame, 2 observation appended,
        3 new track (appended to the dict)."""
        f = self.f
        before = pack_tracks(f)
        out = self._orig_add(image, extracted_features)
        after = pack_tracks(f)
        Fb, Fa = len(before["feat_ids"]), len(after["feat_ids"])
        assert list(after["feat_ids"][:Fb]) == list(before["feat_ids"])          # dict order: old tracks, then new ones
        c = {k: before[k] for k in ("n_cams", "cam_R", "cam_t")}
        c["K"] = np.array(f.K, float)
        c["Kinv"] = np.linalg.inv(f.K)
        c["call_index"] = len(self.append_cases)
        c["feat_ids"] = after["feat_ids"]
        Ma = after["obs_slot"].shape[1]
        for k in ("feat_len", "obs_uv", "obs_dir", "obs_conf", "obs_slot", "idp_base", "idp_anchor_slot", "idp_m", "idp_rho",
                  "lost_cnt", "tracked_cnt"):
            a = after[k]
            pre = np.zeros_like(a) if k != "obs_slot" else -np.ones_like(a)
            if k == "idp_anchor_slot":
                pre = -np.ones_like(a)
            b = before[k]
            if b.ndim >= 2 and k.startswith("obs_"):
                pre[:Fb, :b.shape[1]] = b
            else:
                pre[:Fb] = b
            c[k] = pre
            c["out_" + k] = a
        action = np.zeros(Fa, np.uint8)
        new_uv = np.zeros((Fa, 2))
        new_conf = np.zeros(Fa)
        for j in range(Fa):
            n1 = int(after["feat_len"][j])
            if j >= Fb:
                action[j] = 3
            elif n1 > int(before["feat_len"][j]):
                action[j] = 2
            elif int(after["lost_cnt"][j]) > int(before["lost_cnt"][j]):
                action[j] = 1
            if action[j] >= 2:
                new_uv[j] = after["obs_uv"][j, n1 - 1]
                new_conf[j] = after["obs_conf"][j, n1 - 1]
        c["action"], c["new_uv"], c["new_conf"] = action, new_uv, new_conf
        self.append_cases.append(c)
        return out

    def prune(self):
        c = pack_tracks(self.f)
        c["call_index"] = len(self.select_cases)
        self._in_prune = c
        out = self._orig_prune()
        self._in_prune = None
        self.select_cases.append(c)
        return out

    def gvf(self, features):
        c = self._in_prune
        if c is not None and "out_feat_touch" not in c:
            c["out_feat_touch"] = np.array([k in features for k in c["feat_ids"]], bool)
        return self._orig_gvf(features)

    def remove(self, cameras):
        f = self.f
        cam_ids = list(f.state.cameras.keys())
        if self._in_prune is not None:
            m = np.zeros(len(cam_ids), bool)
            for k in cameras.keys():
                m[cam_ids.index(k)] = True
            self._in_prune["out_remove_mask"] = m
            self._in_prune.setdefault("out_feat_touch", np.zeros(len(self._in_prune["feat_ids"]), bool))
        c = pack_tracks(f)
        c["call_index"] = len(self.drop_cases)
        c["remove_mask"] = np.array([k in cameras for k in cam_ids], bool)
        out = self._orig_remove(cameras)
        after = pack_tracks(f)
        F, M = c["obs_slot"].shape
        for k in ("feat_len", "obs_uv", "obs_dir", "obs_conf", "obs_slot", "idp_base", "idp_anchor_slot"):
            c["out_" + k] = np.zeros_like(c[k]) if k != "obs_slot" else -np.ones_like(c[k])
        pos = {int(fid): j for j, fid in enumerate(c["feat_ids"])}
        for j2, fid in enumerate(after["feat_ids"]):
            j = pos[int(fid)]
            n = int(after["feat_len"][j2])
            c["out_feat_len"][j] = n
            for k in ("obs_uv", "obs_dir", "obs_conf", "obs_slot"):
                c["out_" + k][j, :n] = after[k][j2, :n]
            c["out_idp_base"][j] = after["idp_base"][j2]
            c["out_idp_anchor_slot"][j] = after["idp_anchor_slot"][j2]
        c["out_n_cams"] = after["n_cams"]
        self.drop_cases.append(c)
        return out


class TraceCapture:
    """Operation trace of a whole reference run for the chained (sequence-level) parity test: every process_imu sample,
    every state_augmentation, every get_valid_features + update pair (tracks only -- the filter state is carried by
    the replay) and every remove_cameras mask, in call order, plus the state at the first traced call and after every
    camera frame."""

    def __init__(self, ns, f):
        self.ns, self.f = ns, f
        self.ops = []            # (kind, payload index)
        self.samples, self.masks, self.cases, self.frames = [], [], [], []
        self.init = None
        self._o_imu, self._o_aug = f.process_imu, f.state_augmentation
        self._o_gvf, self._o_upd, self._o_rm = f.get_valid_features, f.update, f.remove_cameras
        f.process_imu, f.state_augmentation = self.imu, self.aug
        f.get_valid_features, f.update, f.remove_cameras = self.gvf, self.upd, self.rm
        self._pending = None
        self.appends = []
        self._o_add = f.add_camera_measurements
        f.add_camera_measurements = self.add

    def add(self, image, extracted_features=None):
        """what add_camera_measurements did to the tracks (the matching itself is out of scope): kind-4 op"""
        f = self.f
        before = pack_tracks(f)
        out = self._o_add(image, extracted_features)
        after = pack_tracks(f)
        Fb, Fa = len(before["feat_ids"]), len(after["feat_ids"])
        assert list(after["feat_ids"][:Fb]) == list(before["feat_ids"])
        action = np.zeros(Fa, np.uint8)
        new_uv, new_conf = np.zeros((Fa, 2)), np.zeros(Fa)
        for j in range(Fa):
            n1 = int(after["feat_len"][j])
            if j >= Fb:
                action[j] = 3
            elif n1 > int(before["feat_len"][j]):
                action[j] = 2
            elif int(after["lost_cnt"][j]) > int(before["lost_cnt"][j]):
                action[j] = 1
            if action[j] >= 2:
                new_uv[j] = after["obs_uv"][j, n1 - 1]
                new_conf[j] = after["obs_conf"][j, n1 - 1]
        self.ops.append((4, len(self.appends)))
        self.appends.append(dict(action=action, new_uv=new_uv, new_conf=new_conf, feat_ids=after["feat_ids"]))
        return out

    def _state(self):
        f = self.f
        imu = f.state.imu
        cams = list(f.state.cameras.values())
        return dict(P=np.array(f.state.covariance, float), imu_R=np.array(imu.T_W_Ii.R, float), imu_t=np.array(imu.T_W_Ii.t, float),
                    imu_v=np.array(imu.v_W_Ii, float), imu_bg=np.array(imu.gyroscope_bias, float),
                    imu_ba=np.array(imu.accelerometer_bias, float), timestamp=float(imu.timestamp),
                    null_R=np.array(imu.T_W_Ii_null.R, float), null_t=np.array(imu.T_W_Ii_null.t, float),
                    null_v=np.array(imu.v_W_Ii_null, float),
                    cam_R=np.stack([c.T_W_Ci.R for c in cams]).astype(float) if cams else np.zeros((0, 3, 3)),
                    cam_t=np.stack([c.T_W_Ci.t for c in cams]).astype(float) if cams else np.zeros((0, 3)))

    def imu(self, m):
        if self.init is None:
            f = self.f
            self.init = self._state()
            T_I_C = f.state.imu.T_W_I.inv() * f.T_W_C
            self.init.update(Qc_diag=np.diag(f.continuous_noise_covariance).copy(), gravity=np.array(f.state.imu.W_gravity, float),
                             R_I_C=np.array(T_I_C.R, float), t_I_C=np.array(T_I_C.t, float))
        self.ops.append((0, len(self.samples)))
        self.samples.append(np.concatenate([[m.timestamp], m.angular_velocity, m.linear_acceleration]))
        return self._o_imu(m)

    def aug(self):
        self.ops.append((1, 0))
        return self._o_aug()

    def gvf(self, features):
        c = pack_update_inputs(self.ns, self.f, features)
        for k in ("P", "cam_R", "cam_t", "imu_R", "imu_t", "imu_v", "imu_bg", "imu_ba"):
            c.pop(k)                                  # carried by the replay
        self._pending = c
        valid, lost = self._o_gvf(features)
        c["out_valid"] = np.array([k in valid for k in features.keys()], bool)
        if len(valid) == 0:                           # the reference does not call update(): selection only
            c["n_valid"] = 0
            self.ops.append((2, len(self.cases)))
            self.cases.append(c)
            self._pending = None
        return valid, lost

    def upd(self, features):
        c = self._pending
        out = self._o_upd(features)
        if c is not None:
            c["n_valid"] = len(features)
            self.ops.append((2, len(self.cases)))
            self.cases.append(c)
            self._pending = None
        return out

    def rm(self, cameras):
        ids = list(self.f.state.cameras.keys())
        self.ops.append((3, len(self.masks)))
        self.masks.append(np.array([k in cameras for k in ids], bool))
        return self._o_rm(cameras)

    def frame_done(self):
        st = self._state()
        trk = pack_tracks(self.f)
        self.frames.append(dict(op_index=len(self.ops), feat_ids=trk["feat_ids"], feat_len=trk["feat_len"],
                                obs_slot=trk["obs_slot"], lost_cnt=trk["lost_cnt"], tracked_cnt=trk["tracked_cnt"],
                                idp_anchor_slot=trk["idp_anchor_slot"], **st))


def noisy_detections(seq, seed=2024, desc_sigma=0.03, score_range=(0.1, 1.0)):
    """Per-frame descriptor noise and widely spread scores on a synthetic stream: the tracks' score-weighted mean
    descriptors then MOVE from frame to frame and the score filter of MSCKF.py:279-282 drops detections."""
    rng = np.random.default_rng(seed)
    for fr in seq.frames:
        d = fr.descriptors + desc_sigma * rng.standard_normal(fr.descriptors.shape)
        fr.descriptors = d / np.linalg.norm(d, axis=1, keepdims=True)
        fr.scores = rng.uniform(score_range[0], score_range[1], len(fr.scores))
    return seq


def fixture_trace(ns, desc_noise=False):
    """Whole-run operation trace of the small configuration (window 8, 40 landmarks, 60 frames).  desc_noise: the variant
    with per-frame descriptor noise / spread scores (small_desc_trace.npz), which also records every frame's raw
    detections and the matcher's view of the tracks afterwards (last_camera_measurement: the score-weighted mean
    descriptor of every track, MSCKF.py:436-441, after the deletions of :742-747 / :772-778)."""
    kw = dict(seed=42, n_frames=60, n_landmarks=40, noise_level="mid", p_drop=0.06, p_outlier=0.03, shuffle=True)
    seq = synth.make_sequence(**kw)
    if desc_noise:
        noisy_detections(seq)
    fp = dict(min_parallax=2.0, max_number_of_camera_states=8, epipolar_rejection_threshold=1e9,
              homography_rejection_threshold=1e9, min_number_of_frames_to_be_tracked=3)
    params, pd = make_params(ns, seq, **fp)
    f = ns.MSCKF(parameters=params, rr=None)
    cap = TraceCapture(ns, f)
    image = np.zeros((seq.height, seq.width, 3), np.uint8)
    frame_at = {fr.imu_index: fr for fr in seq.frames}
    extra_frames = []
    for i in range(len(seq.imu)):
        row = seq.imu[i]
        f.imu_callback(ns.IMUMeasurement(row[0], row[1:4].copy(), row[4:7].copy()))
        fr = frame_at.get(i)
        if fr is not None:
            f.feature_callback(image, ns.CameraMeasurement(keypoints=fr.keypoints.copy(), descriptors=fr.descriptors.copy(),
                                                           scores=fr.scores.copy()))
            if cap.init is not None:
                cap.frame_done()
                if desc_noise:
                    lcm = f.last_camera_measurement
                    extra_frames.append(dict(kp=fr.keypoints, desc=fr.descriptors, scores=fr.scores,
                                             lcm_desc=np.array(lcm.descriptors, float), lcm_idx=np.array(lcm.features_indices, np.int64),
                                             n_hom=f.number_of_features_discarder_for_homography_test,
                                             n_epi=f.number_of_features_discarded_for_epipolar_test))
    flat = {"ops": np.array(cap.ops, np.int32), "samples": np.array(cap.samples), "n_cases": np.asarray(len(cap.cases)),
            "n_masks": np.asarray(len(cap.masks)), "n_frames": np.asarray(len(cap.frames))}
    for k, v in cap.init.items():
        flat["init." + k] = np.asarray(v)
    for i, m in enumerate(cap.masks):
        flat[f"m{i}"] = m
    flat["n_appends"] = np.asarray(len(cap.appends))
    for i, a in enumerate(cap.appends):
        for k, v in a.items():
            flat[f"a{i}.{k}"] = np.asarray(v)
    flat["max_number_of_camera_states"] = np.asarray(int(f.max_number_of_camera_states))
    for i, c in enumerate(cap.cases):
        for k, v in c.items():
            flat[f"c{i}.{k}"] = np.asarray(v)
    for i, fr in enumerate(cap.frames):
        for k, v in fr.items():
            flat[f"f{i}.{k}"] = np.asarray(v)
    for i, x in enumerate(extra_frames):
        for k, v in x.items():
            flat[f"x{i}.{k}"] = np.asarray(v)
    flat["min_cosine_similarity"] = np.asarray(float(f.min_cosine_similarity))
    path = os.path.join(GOLDEN, "small_desc_trace.npz" if desc_noise else "small_trace.npz")
    np.savez_compressed(path, **flat)
    kinds = np.bincount(np.array(cap.ops)[:, 0], minlength=5)
    print(f"wrote {path}: {len(cap.ops)} ops (imu {kinds[0]}, augment {kinds[1]}, update {kinds[2]}, remove {kinds[3]}), "
          f"{len(cap.frames)} frames, {os.path.getsize(path) / 1e6:.2f} MB")


def fixture_prune(ns):
    """Window / track surgery cases from the `small` run: every prune selection and every remove_cameras call."""
    kw = dict(seed=42, n_frames=60, n_landmarks=40, noise_level="mid", p_drop=0.06, p_outlier=0.03, shuffle=True)
    seq = synth.make_sequence(**kw)
    fp = dict(min_parallax=2.0, max_number_of_camera_states=8, epipolar_rejection_threshold=1e9,
              homography_rejection_threshold=1e9, min_number_of_frames_to_be_tracked=3)
    params, pd = make_params(ns, seq, **fp)
    f = ns.MSCKF(parameters=params, rr=None)
    cap = SurgeryCapture(f)
    image = np.zeros((seq.height, seq.width, 3), np.uint8)
    frame_at = {fr.imu_index: fr for fr in seq.frames}
    for i in range(len(seq.imu)):
        row = seq.imu[i]
        f.imu_callback(ns.IMUMeasurement(row[0], row[1:4].copy(), row[4:7].copy()))
        fr = frame_at.get(i)
        if fr is not None:
            f.feature_callback(image, ns.CameraMeasurement(keypoints=fr.keypoints.copy(), descriptors=fr.descriptors.copy(),
                                                           scores=fr.scores.copy()))
    extra = seq_spec(**kw)
    save_cases(os.path.join(GOLDEN, "small_prune_select.npz"), cap.select_cases, extra)
    save_cases(os.path.join(GOLDEN, "small_prune_drop.npz"), cap.drop_cases, extra)
    save_cases(os.path.join(GOLDEN, "small_track_append.npz"), cap.append_cases[::2], extra)


class MatchCapture:
    """add_camera_measurements with the geometric tests ENABLED (MSCKF.py:317-404): what the matcher was given and
    returned (FeatureExtractor.match), which proposals survived the epipolar / homography walk, and the increments of the
    reference's two discard counters."""

    def __init__(self, f):
        self.f = f
        self.cases = []
        self._orig_add = f.add_camera_measurements
        self._orig_match = f.feature_extractor.match
        f.add_camera_measurements = self.add
        f.feature_extractor.match = self.match
        self._m = None

    def match(self, in1, in2, min_cossim=0.82):
        out = self._orig_match(in1, in2, min_cossim)
        matched, not_matched, (i1, i2, n1, n2) = out
        self._m = dict(desc_tracks=np.array(in1.descriptors), desc_new=np.array(in2.descriptors),
                       track_ids=np.array(in1.features_indices, np.int64), min_cossim=float(min_cossim),
                       idxs1=np.asarray(i1, np.int64), idxs2=np.asarray(i2, np.int64),
                       kp_new=np.array(in2.keypoints, float), sc_new=np.array(in2.scores, float))
        return out

    def add(self, image, extracted_features=None):
        f = self.f
        before = pack_tracks(f)
        nh0, ne0 = f.number_of_features_discarder_for_homography_test, f.number_of_features_discarded_for_epipolar_test
        self._m = None
        out = self._orig_add(image, extracted_features)
        m = self._m
        if m is None:
            return out
        after = pack_tracks(f)
        Fb = len(before["feat_ids"])
        c = {k: before[k] for k in ("n_cams", "cam_R", "cam_t", "feat_ids", "feat_len", "obs_uv", "obs_slot", "lost_cnt")}
        c["K"] = np.array(f.K, float)
        c["Kinv"] = np.linalg.inv(f.K)
        c["epipolar_rejection_threshold"] = float(f.epipolar_rejection_threshold)
        c["homography_rejection_threshold"] = float(f.homography_rejection_threshold)
        c["call_index"] = len(self.cases)
        assert list(m["track_ids"]) == list(before["feat_ids"])          # matcher rows are the tracks in dict order
        action = np.zeros(Fb, np.uint8)
        new_uv = np.zeros((Fb, 2))
        new_conf = np.zeros(Fb)
        action[:] = 1                                                     # not matched: lost this frame (:437)
        action[m["idxs1"]] = 2
        new_uv[m["idxs1"]] = m["kp_new"][m["idxs2"]]
        new_conf[m["idxs1"]] = m["sc_new"][m["idxs2"]]
        passed = after["feat_len"][:Fb] > before["feat_len"]
        assert not (passed & (action != 2)).any()
        c.update(action=action, new_uv=new_uv, new_conf=new_conf, out_passed=passed,
                 out_lost_cnt=after["lost_cnt"][:Fb],
                 out_n_homography=f.number_of_features_discarder_for_homography_test - nh0,
                 out_n_epipolar=f.number_of_features_discarded_for_epipolar_test - ne0)
        for k in ("desc_tracks", "desc_new", "min_cossim", "idxs1", "idxs2"):
            c["match_" + k] = m[k]
        self.cases.append(c)
        return out


def fixture_match(ns):
    """Geometric verification + matcher cases: the `small` stream with the reference's tests switched on (the other
    fixtures run with both thresholds at 1e9), a slow start so that the homography branch (baseline < 1 cm) is taken, and
    direct calls of the reference's XFeat.match on float32 64-dimensional descriptor sets."""
    kw = dict(seed=42, n_frames=40, n_landmarks=40, noise_level="mid", p_drop=0.06, p_outlier=0.08, shuffle=True)
    seq = synth.make_sequence(**kw)
    fp = dict(min_parallax=2.0, max_number_of_camera_states=8, epipolar_rejection_threshold=0.002,
              homography_rejection_threshold=0.15, min_number_of_frames_to_be_tracked=3)
    params, pd = make_params(ns, seq, **fp)
    f = ns.MSCKF(parameters=params, rr=None)
    cap = MatchCapture(f)
    image = np.zeros((seq.height, seq.width, 3), np.uint8)
    frame_at = {fr.imu_index: fr for fr in seq.frames}
    for i in range(len(seq.imu)):
        row = seq.imu[i]
        f.imu_callback(ns.IMUMeasurement(row[0], row[1:4].copy(), row[4:7].copy()))
        fr = frame_at.get(i)
        if fr is not None:
            f.feature_callback(image, ns.CameraMeasurement(keypoints=fr.keypoints.copy(), descriptors=fr.descriptors.copy(),
                                                           scores=fr.scores.copy()))
    nh = sum(int(c["out_n_homography"]) for c in cap.cases)
    ne = sum(int(c["out_n_epipolar"]) for c in cap.cases)
    npass = sum(int(c["out_passed"].sum()) for c in cap.cases)
    print(f"match fixture: {len(cap.cases)} frames, passed {npass}, homography rejections {nh}, epipolar rejections {ne}")
    extra = seq_spec(**kw)
    # the reference's matcher on XFeat-like inputs: float32, 64 dimensions, noisy re-detections, distractors, permutation
    rng = np.random.default_rng(7)
    xf = f.feature_extractor.xfeat
    import torch
    for t, (n1, n2, noise, thr) in enumerate([(40, 48, 0.05, 0.82), (256, 300, 0.6, 0.82), (300, 256, 0.45, 0.9),
                                              (64, 64, 0.9, 0.82), (17, 5, 1.2, 0.0), (1, 9, 0.02, 0.82)]):
        base = rng.standard_normal((max(n1, n2), 64))
        base /= np.linalg.norm(base, axis=1, keepdims=True)
        d1 = base[:n1].astype(np.float32)
        perm = rng.permutation(max(n1, n2))[:n2]
        d2 = base[perm] + noise * rng.standard_normal((n2, 64)) / 8.0
        d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(np.float32)
        i0, i1 = xf.match(torch.from_numpy(d1), torch.from_numpy(d2), min_cossim=thr)
        extra[f"xm{t}.d1"], extra[f"xm{t}.d2"], extra[f"xm{t}.thr"] = d1, d2, np.asarray(thr)
        extra[f"xm{t}.idx0"], extra[f"xm{t}.idx1"] = i0.numpy().astype(np.int64), i1.numpy().astype(np.int64)
        print(f"  xfeat.match set {t}: {n1} x {n2}, {len(i0)} matches")
    extra["n_xm"] = np.asarray(6)
    save_cases(os.path.join(GOLDEN, "small_match_verify.npz"), cap.cases, extra)


def fixture_multi(ns, n_filters=64):
    """BASELINE.json configs[2] as a parity fixture: 64 DISTINCT filters (IMU noise low / mid / high x 22 stream seeds), each
    run through the unmodified reference on its own stream (30 landmarks, window 6, 16 frames, drop-outs, outliers, moving
    descriptors, spread scores).  Stored per filter: the state the reference held when its first process_imu began, every IMU
    sample, every frame's raw detections, and after every frame the IMU state, the window poses, the track lengths and the
    covariance's diagonal (the full covariance after the last frame)."""
    levels = list(synth.NOISE_LEVELS.keys())
    flat = {"n_filters": np.asarray(n_filters)}
    t0 = time.time()
    for i in range(n_filters):
        kw = dict(seed=42 + i, n_frames=16, n_landmarks=30, noise_level=levels[i % 3], p_drop=0.06, p_outlier=0.03, shuffle=True)
        seq = noisy_detections(synth.make_sequence(**kw), seed=7000 + i)
        fp = dict(min_parallax=2.0, max_number_of_camera_states=6, epipolar_rejection_threshold=1e9,
                  homography_rejection_threshold=1e9, min_number_of_frames_to_be_tracked=3)
        params, pd = make_params(ns, seq, **fp)
        f = ns.MSCKF(parameters=params, rr=None)
        cap = TraceCapture(ns, f)
        image = np.zeros((seq.height, seq.width, 3), np.uint8)
        frame_at = {fr.imu_index: fr for fr in seq.frames}
        k = 0
        for j in range(len(seq.imu)):
            row = seq.imu[j]
            f.imu_callback(ns.IMUMeasurement(row[0], row[1:4].copy(), row[4:7].copy()))
            fr = frame_at.get(j)
            if fr is not None:
                f.feature_callback(image, ns.CameraMeasurement(keypoints=fr.keypoints.copy(), descriptors=fr.descriptors.copy(),
                                                               scores=fr.scores.copy()))
                if cap.init is not None:
                    st = cap._state()
                    trk = pack_tracks(f)
                    pre = f"t{i}.f{k}."
                    flat[pre + "kp"], flat[pre + "desc"], flat[pre + "scores"] = fr.keypoints, fr.descriptors.astype(np.float32), fr.scores
                    flat[pre + "n_samples"] = np.asarray(len(cap.samples))
                    flat[pre + "imu"] = np.concatenate([st["imu_R"].flatten(), st["imu_t"], st["imu_v"], st["imu_bg"], st["imu_ba"]])
                    flat[pre + "cam_t"] = st["cam_t"]
                    flat[pre + "feat_len"] = trk["feat_len"]
                    flat[pre + "P_diag"] = np.diag(st["P"]).copy()
                    flat[pre + "n_gate"] = np.asarray(f.number_of_residuals_discarded_for_gasting_test)
                    k += 1
        for key, v in cap.init.items():
            flat[f"t{i}.init.{key}"] = np.asarray(v)
        flat[f"t{i}.samples"] = np.array(cap.samples)
        flat[f"t{i}.n_frames"] = np.asarray(k)
        flat[f"t{i}.P_last"] = cap._state()["P"]
        flat[f"t{i}.sigma"] = np.asarray(float(f.sigma_image))
        if i % 8 == 0:
            print(f"  filter {i}: {k} frames, {len(f.features)} tracks, {len(f.state.cameras)} cameras, {time.time() - t0:.0f} s", flush=True)
    flat["par.min_parallax"], flat["par.max_cams"], flat["par.min_tracked"], flat["par.min_lost"] = (np.asarray(2.0), np.asarray(6),
                                                                                                   np.asarray(3), np.asarray(2))
    flat["par.min_cosine_similarity"] = np.asarray(float(f.min_cosine_similarity))
    flat["K"] = np.asarray(seq.K, float)
    path = os.path.join(GOLDEN, "multi_64.npz")
    np.savez_compressed(path, **flat)
    print(f"wrote {path}: {n_filters} filters, {os.path.getsize(path) / 1e6:.2f} MB")


def make_params(ns, seq, **kw):
    sa, sg, sba, sbg = synth.NOISE_LEVELS[seq.noise_level]
    p = dict(K=seq.K, width=seq.width, height=seq.height, sigma_image=0.1, only_imu=False,
             accelerometer_noise_density=sa, gyroscope_noise_density=sg, accelerometer_random_walk=sba,
             gyroscope_random_walk=sbg, W_gravity=np.array([0, 0, -9.81]), number_of_extracted_features=300,
             min_cosine_similarity=0.95, use_parallax=True, min_parallax=45, epipolar_rejection_threshold=0.005,
             homography_rejection_threshold=5, min_number_of_frames_to_be_tracked=4,
             min_number_of_frames_to_be_lost=2, max_number_of_camera_states=30)
    p.update(kw)
    return ns.MSCKFParameters(**p), p


def run_reference(ns, seq, params, keep_update, keep_propagate, verbose=False):
    f = ns.MSCKF(parameters=params, rr=None)
    cap = Capture(ns, f, keep_update, keep_propagate)
    image = np.zeros((seq.height, seq.width, 3), np.uint8)
    frame_at = {fr.imu_index: fr for fr in seq.frames}
    traj = []
    t0 = time.time()
    for i in range(len(seq.imu)):
        row = seq.imu[i]
        f.imu_callback(ns.IMUMeasurement(row[0], row[1:4].copy(), row[4:7].copy()))
        fr = frame_at.get(i)
        if fr is not None:
            f.feature_callback(image, ns.CameraMeasurement(keypoints=fr.keypoints.copy(),
                                                           descriptors=fr.descriptors.copy(),
                                                           scores=fr.scores.copy()))
            imu = f.state.imu
            traj.append(np.concatenate([[row[0]], imu.T_W_Ii.R.flatten(), imu.T_W_Ii.t, imu.v_W_Ii,
                                        imu.gyroscope_bias, imu.accelerometer_bias,
                                        [len(f.state.cameras), len(f.features), f.state.covariance.shape[0],
                                         f.number_of_residuals_discarded_for_gasting_test,
                                         np.trace(f.state.covariance)]]))
            if verbose:
                print(f"  frame {len(traj):3d} cams {len(f.state.cameras):2d} feats {len(f.features):4d} "
                      f"d {f.state.covariance.shape[0]:3d} updates {cap.n_update_calls:3d} "
                      f"rej {f.number_of_residuals_discarded_for_gasting_test:4d} {time.time() - t0:6.1f}s",
                      flush=True)
    return f, cap, np.array(traj)


def save_cases(path, cases, extra=None):
    flat = {}
    for i, c in enumerate(cases):
        for k, v in c.items():
            flat[f"c{i}.{k}"] = np.asarray(v)
    flat["n_cases"] = np.asarray(len(cases))
    if extra:
        flat.update(extra)
    np.savez_compressed(path, **flat)
    print(f"wrote {path}: {len(cases)} cases, {os.path.getsize(path) / 1e6:.2f} MB")


def seq_spec(**kw):
    """The make_sequence kwargs are stored beside the fixture so tests can regenerate the identical stream."""
    return {f"seq.{k}": np.asarray(v) for k, v in kw.items()}


def fixture_small(ns):
    """Natural-like workload: parallax/lost-driven updates, dropout, outliers, window 8; every update case kept."""
    kw = dict(seed=42, n_frames=60, n_landmarks=40, noise_level="mid", p_drop=0.06, p_outlier=0.03, shuffle=True)
    seq = synth.make_sequence(**kw)
    fp = dict(min_parallax=2.0, max_number_of_camera_states=8, epipolar_rejection_threshold=1e9,
              homography_rejection_threshold=1e9, min_number_of_frames_to_be_tracked=3)
    params, pd = make_params(ns, seq, **fp)
    f, cap, traj = run_reference(ns, seq, params, lambda i: True, lambda i: i < 40 or (200 <= i < 230), verbose=True)
    extra = seq_spec(**kw)
    extra.update({f"par.{k}": np.asarray(v) for k, v in fp.items()})
    extra["traj"] = traj
    save_cases(os.path.join(GOLDEN, "small_updates.npz"), cap.update_cases, extra)
    save_cases(os.path.join(GOLDEN, "small_propagate.npz"), cap.propagate_cases, extra)


def fixture_baseline(ns, n_landmarks=256, window=30, n_frames=36, name="baseline_256x30"):
    """BASELINE.json configs[1] shape: every track updated every frame (min_parallax = 0), F = 256, M -> window."""
    kw = dict(seed=42, n_frames=n_frames, n_landmarks=n_landmarks, noise_level="mid", p_outlier=0.02)
    seq = synth.make_sequence(**kw)
    fp = dict(min_parallax=0.0, max_number_of_camera_states=window, epipolar_rejection_threshold=1e9,
              homography_rejection_threshold=1e9, min_number_of_frames_to_be_tracked=2)
    params, pd = make_params(ns, seq, **fp)
    # steady state: keep the last three update calls (d alternates 195 / 201) and 9 consecutive propagations
    f, cap, traj = run_reference(ns, seq, params, lambda i: i >= KEEP_FROM[0], lambda i: i >= KEEP_FROM[1],
                                 verbose=True)
    extra = seq_spec(**kw)
    extra.update({f"par.{k}": np.asarray(v) for k, v in fp.items()})
    extra["traj"] = traj
    save_cases(os.path.join(GOLDEN, f"{name}_updates.npz"), cap.update_cases[-3:], extra)
    save_cases(os.path.join(GOLDEN, f"{name}_propagate.npz"), cap.propagate_cases[-9:][:3 if n_landmarks > 100 else 9], extra)


KEEP_FROM = [0, 0]

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    ns = ref_shim.load_reference()
    np.random.seed(42)
    if args.only in (None, "small"):
        fixture_small(ns)
    if args.only in (None, "prune"):
        fixture_prune(ns)
    if args.only in (None, "trace"):
        fixture_trace(ns)
    if args.only in (None, "desc"):
        fixture_trace(ns, desc_noise=True)
    if args.only in (None, "multi"):
        fixture_multi(ns)
    if args.only in (None, "match"):
        fixture_match(ns)
    if args.only in (None, "mid"):
        # mid-size steady-state fixture (cheap to regenerate): F = 48, window 10
        KEEP_FROM[:] = [18, 14 * 9]
        fixture_baseline(ns, n_landmarks=48, window=10, n_frames=18, name="mid_48x10")
    if args.only in (None, "baseline"):
        KEEP_FROM[:] = [30, 34 * 9]
        fixture_baseline(ns, n_landmarks=256, window=30, n_frames=38, name="baseline_256x30")
