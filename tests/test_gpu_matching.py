"""GPU parity of the matching front of add_camera_measurements through the C ABI:
msckf_match_descriptors_batched (XFeat.match + the proposal list) and msckf_verify_matches_batched (epipolar / homography
walk), against the reference's own decisions (tests/golden/small_match_verify.npz) and the oracle."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
torch = pytest.importorskip("torch")
from oracle import msckf_oracle as o                                   # noqa: E402  (checker only)
from monocular_visual_inertial_msckf_b200.batched import BatchedMSCKF, UpdateParams   # noqa: E402

pytestmark = pytest.mark.gpu


def _cases(golden_dir):
    return o.load_cases(os.path.join(golden_dir, "small_match_verify.npz"))


def _verify_batch(cases, pad_F=3, pad_M=1):
    F = max(len(c["feat_ids"]) for c in cases) + pad_F
    M = max(c["obs_slot"].shape[1] for c in cases) + pad_M
    N = max(int(c["n_cams"]) for c in cases)
    flt = BatchedMSCKF(len(cases), F, M, N, UpdateParams(K=np.array(cases[0]["K"], float)))
    action = torch.zeros((flt.B, F), dtype=torch.uint8, device="cuda")
    new_uv = torch.zeros((flt.B, F, 2), dtype=torch.float64, device="cuda")
    for b, c in enumerate(cases):
        f, m, n = c["obs_slot"].shape[0], c["obs_slot"].shape[1], int(c["n_cams"])
        flt.n_cams[b] = n
        flt.feat_len[b].zero_()
        flt.feat_len[b, :f] = torch.from_numpy(c["feat_len"])
        flt.obs_slot[b, :f, :m] = torch.from_numpy(np.maximum(c["obs_slot"], 0))
        flt.obs_uv[b, :f, :m] = torch.from_numpy(c["obs_uv"])
        flt.cam_t[b, :n] = torch.from_numpy(c["cam_t"])
        flt.cam_R[b, :n] = torch.from_numpy(np.asarray(c["cam_R"], float).reshape(n, 9))
        flt.lost_cnt[b, :f] = torch.from_numpy(c["lost_cnt"])
        action[b, :f] = torch.from_numpy(c["action"])
        new_uv[b, :f] = torch.from_numpy(c["new_uv"])
    return flt, action, new_uv


def test_verify_matches_vs_golden(golden_dir):
    """Every frame of a reference run with the geometric tests switched on (38 frames, 69 homography and 197 epipolar
    rejections): the device keeps exactly the proposals the reference kept, demotes the others to `lost this frame`, and
    its two counters move like the reference's; the verdict per track equals the oracle's (which names the failing test)."""
    cases, _ = _cases(golden_dir)
    flt, action, new_uv = _verify_batch(cases)
    nh = torch.zeros(flt.B, dtype=torch.int32, device="cuda")
    ne = torch.zeros(flt.B, dtype=torch.int32, device="cuda")
    c0 = cases[0]
    verdict = flt.verify_matches(action, new_uv, c0["epipolar_rejection_threshold"], c0["homography_rejection_threshold"],
                                 counters=(nh, ne))
    torch.cuda.synchronize()
    verdict, act, nh, ne = verdict.cpu().numpy(), action.cpu().numpy(), nh.cpu().numpy(), ne.cpu().numpy()
    n_rej = 0
    for b, c in enumerate(cases):
        f = len(c["feat_ids"])
        ref = o.verify_matches(c)
        assert np.array_equal(verdict[b, :f] == 1, c["out_passed"]), c["call_index"]
        assert np.array_equal(verdict[b, :f], ref["verdict"]), c["call_index"]
        assert np.array_equal(act[b, :f], ref["action"]) and (act[b, f:] == 0).all() and (verdict[b, f:] == 0).all()
        assert nh[b] == c["out_n_homography"] and ne[b] == c["out_n_epipolar"], c["call_index"]
        n_rej += int(nh[b] + ne[b])
    assert n_rej > 200
    # the bookkeeping kernel then reproduces the reference's lost counters (a rejected proposal is a lost frame)
    zeros = torch.zeros((flt.B, flt.F), dtype=torch.float64, device="cuda")
    flt.append_observations(action, new_uv, zeros)
    torch.cuda.synchronize()
    lost = flt.lost_cnt.cpu().numpy()
    for b, c in enumerate(cases):
        f = len(c["feat_ids"])
        assert np.array_equal(lost[b, :f], c["out_lost_cnt"]), c["call_index"]


def test_verify_matches_long_tracks_and_edges():
    """Tracks longer than one warp (40 observations, 48-pose window): the first failing observation decides, wherever it
    sits; empty tracks and filters without cameras pass every proposal; thresholds at +-inf."""
    rng = np.random.default_rng(3)
    B, F, M, N = 6, 9, 40, 36
    K = np.array([[320.0, 0, 320], [0, 320, 240], [0, 0, 1]])
    cases = []
    for b in range(B):
        cam_R = np.stack([np.linalg.qr(np.eye(3) + 0.02 * rng.standard_normal((3, 3)))[0] for _ in range(N)])
        cam_R *= np.sign(np.linalg.det(cam_R))[:, None, None]
        cam_t = 0.3 * rng.standard_normal((N, 3))
        cam_t[5] = cam_t[N - 1] + 1e-3                      # one camera within 1 cm of the newest: homography branch
        c = dict(K=K, Kinv=np.linalg.inv(K), n_cams=N, cam_R=cam_R, cam_t=cam_t, feat_len=rng.integers(0, M + 1, F).astype(np.int32),
                 obs_uv=rng.uniform(0, 480, (F, M, 2)), obs_slot=rng.integers(0, N - 1, (F, M)).astype(np.int32),
                 action=np.full(F, 2, np.uint8), new_uv=rng.uniform(0, 480, (F, 2)),
                 epipolar_rejection_threshold=0.35, homography_rejection_threshold=5.0, feat_ids=np.arange(F))
        c["feat_len"][0] = M
        c["feat_len"][1] = 0
        c["action"][2] = 1
        c["feat_len"][3] = 35
        c["obs_slot"][3] = 5                                # a track seen only from that camera: homography tests alone
        cases.append(c)
    flt = BatchedMSCKF(B, F, M, N, UpdateParams(K=K))
    action = torch.zeros((B, F), dtype=torch.uint8, device="cuda")
    new_uv = torch.zeros((B, F, 2), dtype=torch.float64, device="cuda")
    for b, c in enumerate(cases):
        flt.n_cams[b] = N
        flt.feat_len[b] = torch.from_numpy(c["feat_len"])
        flt.obs_slot[b] = torch.from_numpy(c["obs_slot"])
        flt.obs_uv[b] = torch.from_numpy(c["obs_uv"])
        flt.cam_t[b] = torch.from_numpy(c["cam_t"])
        flt.cam_R[b] = torch.from_numpy(c["cam_R"].reshape(N, 9))
        action[b] = torch.from_numpy(c["action"])
        new_uv[b] = torch.from_numpy(c["new_uv"])
    a0 = action.clone()
    verdict = flt.verify_matches(action, new_uv, 0.35, 5.0).cpu().numpy()
    seen = set()
    for b, c in enumerate(cases):
        ref = o.verify_matches(c)
        far = np.abs(ref["score"] - np.where(ref["verdict"] == 2, 5.0, 0.35)) > 1e-9
        assert np.array_equal(verdict[b][far], ref["verdict"][far])
        seen.update(int(v) for v in verdict[b])
        assert verdict[b, 1] == 1 and verdict[b, 2] == 0
    assert seen == {0, 1, 2, 3}
    act = a0.clone()
    v_inf = flt.verify_matches(act, new_uv, float("inf"), float("inf")).cpu().numpy()
    assert ((v_inf == 1) == (a0.cpu().numpy() == 2)).all() and torch.equal(act, a0)
    flt.n_cams.zero_()
    act = a0.clone()
    v_none = flt.verify_matches(act, new_uv, -1e9, -1e9).cpu().numpy()
    assert ((v_none == 1) == (a0.cpu().numpy() == 2)).all()


def _run_matcher(d1_list, d2_list, thr, F1, F2, D, kp=None, sc=None):
    B = len(d1_list)
    flt = BatchedMSCKF(B, F1, 2, 2, UpdateParams())
    dt = torch.zeros((B, F1, D), dtype=torch.float32, device="cuda")
    dn = torch.zeros((B, F2, D), dtype=torch.float32, device="cuda")
    n1 = torch.zeros(B, dtype=torch.int32, device="cuda")
    n2 = torch.zeros(B, dtype=torch.int32, device="cuda")
    for b, (d1, d2) in enumerate(zip(d1_list, d2_list)):
        dt[b, :len(d1)] = torch.from_numpy(np.asarray(d1, np.float32))
        dn[b, :len(d2)] = torch.from_numpy(np.asarray(d2, np.float32))
        n1[b], n2[b] = len(d1), len(d2)
    out = flt.match_descriptors(dt, dn, thr, n1, n2, kp, sc)
    torch.cuda.synchronize()
    return flt, [x.cpu().numpy() for x in out]


def test_matcher_vs_reference_xfeat_sets(golden_dir):
    """Direct XFeat.match calls of the reference on float32 64-dimensional descriptor sets (full, partial, almost empty
    match sets, min_cossim = 0, a single track): identical index pairs."""
    _, extra = _cases(golden_dir)
    for t in range(int(extra["n_xm"])):
        d1, d2, thr = extra[f"xm{t}.d1"], extra[f"xm{t}.d2"], float(extra[f"xm{t}.thr"])
        _, (m12, m21) = _run_matcher([d1], [d2], thr, len(d1) + 3, len(d2) + 2, 64)
        i1 = np.nonzero(m12[0] >= 0)[0]
        assert np.array_equal(i1, extra[f"xm{t}.idx0"]) and np.array_equal(m12[0][i1], extra[f"xm{t}.idx1"]), t
        o12, o21 = o.match_descriptors(d1, d2, thr)
        assert np.array_equal(m12[0, :len(d1)], o12) and np.array_equal(m21[0, :len(d2)], o21)
        assert (m12[0, len(d1):] == -1).all() and (m21[0, len(d2):] == -1).all()


def test_matcher_proposals_vs_reference_frames(golden_dir):
    """The matcher calls of the traced run, one filter per frame (32-dimensional synthetic descriptors): index pairs equal
    to the reference's, and the proposal list (action / new_uv / new_conf, new tracks in ascending detection order in the
    free slots) equal to what the reference derived from them; chained with verify_matches it gives the reference's
    surviving proposals."""
    cases, _ = _cases(golden_dir)
    F1 = max(len(c["feat_ids"]) for c in cases) + 30
    F2 = max(len(c["match_desc_new"]) for c in cases) + 1
    B = len(cases)
    kp = torch.zeros((B, F2, 2), dtype=torch.float64, device="cuda")
    sc = torch.zeros((B, F2), dtype=torch.float64, device="cuda")
    rng = np.random.default_rng(5)
    kps, scs = [], []
    for b, c in enumerate(cases):
        n2 = len(c["match_desc_new"])
        # the detections' keypoints / scores: those of the matched ones are in the fixture (new_uv / new_conf by track)
        k_, s_ = rng.uniform(0, 400, (n2, 2)), rng.uniform(0.5, 1, n2)
        k_[c["match_idxs2"]] = c["new_uv"][c["match_idxs1"]]
        s_[c["match_idxs2"]] = c["new_conf"][c["match_idxs1"]]
        kp[b, :n2], sc[b, :n2] = torch.from_numpy(k_), torch.from_numpy(s_)
        kps.append(k_), scs.append(s_)
    flt, (m12, m21, action, new_uv, new_conf, n_out) = _run_matcher(
        [c["match_desc_tracks"] for c in cases], [c["match_desc_new"] for c in cases], float(cases[0]["match_min_cossim"]),
        F1, F2, 32, kp, sc)
    assert int(flt._overflow.item()) == 0
    for b, c in enumerate(cases):
        f, n2 = len(c["feat_ids"]), len(c["match_desc_new"])
        i1 = np.nonzero(m12[b] >= 0)[0]
        assert np.array_equal(i1, c["match_idxs1"]) and np.array_equal(m12[b][i1], c["match_idxs2"]), c["call_index"]
        assert np.array_equal(action[b, :f], c["action"])
        assert np.array_equal(new_uv[b, :f][c["action"] == 2], c["new_uv"][c["action"] == 2])
        assert np.array_equal(new_conf[b, :f][c["action"] == 2], c["new_conf"][c["action"] == 2])
        un = np.setdiff1d(np.arange(n2), c["match_idxs2"])               # FeatureExtractor.py:73
        assert n_out[b] == f + len(un)
        assert (action[b, f:f + len(un)] == 3).all() and (action[b, f + len(un):] == 0).all()
        assert np.array_equal(new_uv[b, f:f + len(un)], kps[b][un]) and np.array_equal(new_conf[b, f:f + len(un)], scs[b][un])
    # chained: proposals from the device matcher -> device verification == the reference's surviving proposals
    vflt, act_ref, uv_ref = _verify_batch(cases, pad_F=F1 - max(len(c["feat_ids"]) for c in cases), pad_M=1)
    act_dev = torch.from_numpy(np.where(action == 3, 0, action).astype(np.uint8)).cuda()
    uv_dev = torch.from_numpy(new_uv).cuda()
    v = vflt.verify_matches(act_dev, uv_dev, cases[0]["epipolar_rejection_threshold"],
                            cases[0]["homography_rejection_threshold"]).cpu().numpy()
    for b, c in enumerate(cases):
        assert np.array_equal(v[b, :len(c["feat_ids"])] == 1, c["out_passed"]), c["call_index"]


def test_matcher_capacity_and_overflow():
    """Benchmark-shape sets (256 tracks x 300 detections x 64) against the oracle on noisy data away from ties; more new
    detections than free slots are counted, not written out of bounds; unsupported descriptor lengths are refused."""
    rng = np.random.default_rng(12)
    d1s, d2s = [], []
    for b in range(5):
        base = rng.standard_normal((300, 64))
        base /= np.linalg.norm(base, axis=1, keepdims=True)
        d2 = base[rng.permutation(300)] + 0.7 * rng.standard_normal((300, 64)) / 8
        d1s.append(base[:256].astype(np.float32))
        d2s.append((d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(np.float32))
    kp = torch.rand((5, 300, 2), dtype=torch.float64, device="cuda")
    sc = torch.rand((5, 300), dtype=torch.float64, device="cuda")
    flt, (m12, m21, action, new_uv, new_conf, n_out) = _run_matcher(d1s, d2s, 0.82, 256, 300, 64, kp, sc)
    n_over = 0
    for b in range(5):
        o12, o21 = o.match_descriptors(d1s[b], d2s[b], 0.82)
        cs = d1s[b] @ d2s[b].T
        top2 = np.sort(cs, axis=1)[:, -2:]
        clear = (top2[:, 1] - top2[:, 0] > 1e-5) & (np.abs(top2[:, 1] - 0.82) > 1e-5)
        col2 = np.sort(cs, axis=0)[-2:, :]
        assert clear.all() and (col2[1] - col2[0] > 1e-5).all()          # the data keeps clear of ties / the threshold
        assert np.array_equal(m12[b], o12) and np.array_equal(m21[b], o21)
        assert 0 < (o12 >= 0).sum() < 256
        n_over += int((o21 < 0).sum())                                   # all 256 slots are tracks: no room for new ones
        assert n_out[b] == 256 and (action[b] != 3).all()
    assert int(flt._overflow.item()) == n_over
    with pytest.raises(RuntimeError):
        _run_matcher([np.zeros((4, 48), np.float32)], [np.zeros((4, 48), np.float32)], 0.82, 4, 4, 48)
    with pytest.raises(RuntimeError):
        _run_matcher([np.zeros((4, 64), np.float32)], [np.zeros((4, 64), np.float32)], 0.82, 600, 600, 64)


def test_whole_run_from_detections(golden_dir):
    """The camera-frame loop of a whole reference run driven by nothing but the IMU samples and each frame's DETECTIONS
    (keypoints, descriptors, scores of the synthetic stream the reference was run on): matcher, geometric verification,
    track bookkeeping, updates and pruning all on the device (BatchedMSCKF.frame_step_detections).  After every camera
    frame the proposals equal the reference matcher's, and the filter state and the tracks equal the reference's."""
    from monocular_visual_inertial_msckf_b200 import synth, _lib
    from helpers import rel_err
    tr = o.load_trace(os.path.join(golden_dir, "small_trace.npz"))
    seq = synth.make_sequence(seed=42, n_frames=60, n_landmarks=40, noise_level="mid", p_drop=0.06, p_outlier=0.03, shuffle=True)
    ini, cases, apps = tr["init"], tr["cases"], tr["appends"]
    assert len(apps) == len(seq.frames) - 1
    F = max(len(a["action"]) for a in apps) + 1
    M = max(c["obs_uv"].shape[1] for c in cases) + 1
    N = tr["max_cams"] + 1
    F2 = max(len(fr.keypoints) for fr in seq.frames) + 2
    D = seq.frames[0].descriptors.shape[1]
    par = UpdateParams.from_case(cases[0], _lib.COMPRESS_GRAM)
    par.gravity = np.array(ini["gravity"], float)
    B = 2
    flt = BatchedMSCKF(B, F, M, N, par)
    flt.P.zero_()
    flt.P[:, :15, :15] = torch.from_numpy(np.asarray(ini["P"], float))
    flt.imu_state[:] = torch.from_numpy(np.concatenate([np.asarray(ini["imu_R"], float).flatten(), ini["imu_t"], ini["imu_v"],
                                                        ini["imu_bg"], ini["imu_ba"]]))
    flt.imu_null[:] = torch.from_numpy(np.concatenate([np.asarray(ini["null_R"], float).flatten(), ini["null_t"], ini["null_v"]]))
    flt.imu_time[:] = float(ini["timestamp"])
    flt.Qc_diag[:] = torch.from_numpy(np.asarray(ini["Qc_diag"], float)[[0, 3, 6, 9]].copy())
    flt.n_cams.zero_()
    flt.feat_len.zero_()
    ops = tr["ops"]
    first, worst, start = True, 0.0, 0
    for i, fr in enumerate(tr["frames"]):
        seg = ops[start:int(fr["op_index"])]
        start = int(fr["op_index"])
        smp = np.array([tr["samples"][idx] for kind, idx in seg if kind == 0], float)
        a = apps[[idx for kind, idx in seg if kind == 4][0]]
        det = seq.frames[i + 1]
        keep = det.scores >= 0.5 * det.scores.mean()                     # the score filter of MSCKF.py:279-282 (host side)
        n2 = int(keep.sum())
        desc = torch.zeros((B, F2, D), dtype=torch.float32, device="cuda")
        kp = torch.zeros((B, F2, 2), dtype=torch.float64, device="cuda")
        sc = torch.zeros((B, F2), dtype=torch.float64, device="cuda")
        desc[:, :n2] = torch.from_numpy(det.descriptors[keep].astype(np.float32)).cuda()
        kp[:, :n2] = torch.from_numpy(det.keypoints[keep]).cuda()
        sc[:, :n2] = torch.from_numpy(det.scores[keep]).cuda()
        n_new = torch.full((B,), n2, dtype=torch.int32, device="cuda")
        samples = torch.from_numpy(smp).reshape(1, -1, 7).repeat(B, 1, 1).cuda()
        action, verdict = flt.frame_step_detections(samples, ini["R_I_C"], ini["t_I_C"], desc, kp, sc, n_new, tr["max_cams"],
                                                    min_cosine_similarity=0.95, epipolar_rejection_threshold=1e9,
                                                    homography_rejection_threshold=1e9, use_null_state=first)
        first = False
        torch.cuda.synchronize()
        fa = len(a["action"])
        act = action.cpu().numpy()
        assert (act[:, :fa] == a["action"]).all() and (act[:, fa:] == 0).all(), i
        d = fr["P"].shape[0]
        nc = (d - 15) // 6
        nf = len(fr["feat_len"])
        for b in range(B):
            assert int(flt.n_cams[b]) == nc, (i, int(flt.n_cams[b]), nc)
            ln = flt.feat_len[b].cpu().numpy()
            assert (ln[:nf] == fr["feat_len"]).all() and (ln[nf:] == 0).all(), i
            sl = flt.obs_slot[b].cpu().numpy()
            for j in range(nf):
                assert (sl[j, :ln[j]] == fr["obs_slot"][j, :ln[j]]).all(), (i, j)
            assert (flt.lost_cnt[b, :nf].cpu().numpy() == fr["lost_cnt"]).all(), i
            assert (flt.tracked_cnt[b, :nf].cpu().numpy() == fr["tracked_cnt"]).all(), i
            st = flt.imu_state[b].cpu().numpy()
            errs = [rel_err(flt.P[b, :d, :d].cpu().numpy(), fr["P"]), rel_err(st[:9].reshape(3, 3), fr["imu_R"]),
                    rel_err(st[9:12], fr["imu_t"]), rel_err(st[12:15], fr["imu_v"]),
                    rel_err(flt.cam_t[b, :nc].cpu().numpy(), fr["cam_t"])]
            worst = max(worst, max(errs))
            assert worst < 1e-9, (i, errs)
    assert len(tr["frames"]) >= 50
    print("frame loop from detections, worst relative error over", len(tr["frames"]), "frames:", worst)


def test_monte_carlo_device_matcher_equals_host_association():
    """Monte-Carlo driver with the association done on the device from the detections (matcher = "device") against the
    same sequences associated on the host from the ground-truth landmark ids: same tracks, same windows, same
    trajectories, for six filters (three IMU-noise levels) over 24 frames."""
    from monocular_visual_inertial_msckf_b200.monte_carlo import run_monte_carlo
    host = run_monte_carlo(list(range(6)), n_frames=24, n_landmarks=40, window=8)
    dev = run_monte_carlo(list(range(6)), n_frames=24, n_landmarks=40, window=8, matcher="device")
    assert np.array_equal(host.n_tracks, dev.n_tracks) and np.array_equal(host.n_cams, dev.n_cams)
    assert np.max(np.abs(host.positions - dev.positions)) < 1e-9
    assert np.isfinite(dev.ate_rmse).all()


def test_descriptor_kernels_vs_oracle():
    """Score filter (MSCKF.py:279-284), per-observation descriptor lists with the deletion of a pruned camera's rows
    (:764-765), the move of whole tracks on compaction, and the score-weighted mean the matcher sees (:439), on random
    tracks against the oracle's restatements."""
    import ctypes as C
    from monocular_visual_inertial_msckf_b200 import _lib
    rng = np.random.default_rng(3)
    B, F, M, N, D, F2 = 3, 20, 8, 8, 32, 26
    flt = BatchedMSCKF(B, F, M, N, UpdateParams())
    flt._ensure_descriptors(D)
    # --- score filter ---
    sc = rng.uniform(0.05, 1.0, (B, F2))
    kp = rng.uniform(0, 400, (B, F2, 2))
    de = rng.standard_normal((B, F2, D)).astype(np.float32)
    n_in = np.array([F2, 17, 0], np.int32)
    do, ko, so, no = flt.score_filter(torch.from_numpy(de).cuda(), torch.from_numpy(kp).cuda(), torch.from_numpy(sc).cuda(),
                                      torch.from_numpy(n_in).cuda())
    for b in range(B):
        keep = o.score_filter(sc[b, :n_in[b]])
        n = int(keep.sum())
        assert int(no[b]) == n
        assert np.array_equal(so[b, :n].cpu().numpy(), sc[b, :n_in[b]][keep]) and np.array_equal(ko[b, :n].cpu().numpy(), kp[b, :n_in[b]][keep])
        assert np.array_equal(do[b, :n].cpu().numpy(), de[b, :n_in[b]][keep])
    # --- lists: random tracks, mean, drop two cameras, mean again ---
    lens = rng.integers(0, M + 1, (B, F)).astype(np.int32)
    slots = np.zeros((B, F, M), np.int32)
    for b in range(B):
        for j in range(F):
            slots[b, j, :lens[b, j]] = np.sort(rng.choice(N, lens[b, j], replace=False))
    conf = rng.uniform(0.1, 1.0, (B, F, M))
    od = rng.standard_normal((B, F, M, D)).astype(np.float32)
    flt.feat_len.copy_(torch.from_numpy(lens))
    flt.obs_slot.copy_(torch.from_numpy(slots))
    flt.obs_conf.copy_(torch.from_numpy(conf))
    # the descriptor of observation i sits in the row of its camera (the camera ring; here a scrambled slot -> row map)
    R = flt.desc_rows
    ring = np.stack([rng.permutation(R)[:N] for _ in range(B)]).astype(np.int32)
    od_ring = np.zeros((B, F, R, D), np.float32)
    for b in range(B):
        for j in range(F):
            for i in range(lens[b, j]):
                od_ring[b, j, ring[b, slots[b, j, i]]] = od[b, j, i]
    flt.cam_row.copy_(torch.from_numpy(ring))
    flt.obs_desc.copy_(torch.from_numpy(od_ring))
    flt.n_cams.fill_(N)
    flt.rebuild_descriptor_sums()          # the rows were written directly: the running sums start from them

    def check_means(lens_now, conf_now, od_now):
        # both routes to the matcher's rows: from the running sums (default) and from the lists
        for got in (flt.track_descriptors().cpu().numpy(), flt.track_descriptors(from_lists=True).cpu().numpy()):
            for b in range(B):
                for j in range(F):
                    n = int(lens_now[b, j])
                    want = o.track_mean_descriptor(od_now[b, j, :n], conf_now[b, j, :n]).astype(np.float32) if n else np.zeros(D, np.float32)
                    assert np.max(np.abs(got[b, j] - want)) <= 2e-7 * max(1.0, np.max(np.abs(want))), (b, j)

    check_means(lens, conf, od)
    mask = np.zeros((B, N), np.uint8)
    mask[:, [1, 4]] = 1
    mask[2, 6] = 1
    before = flt.obs_desc.clone()
    flt.drop_observations(torch.from_numpy(mask).cuda())
    flt.remove_cameras(torch.from_numpy(mask).cuda())        # (compacts the slot -> row map like the window)
    assert torch.equal(before, flt.obs_desc)                 # no descriptor moves when cameras leave
    lens2, conf2, od2 = lens.copy(), conf.copy(), od.copy()
    for b in range(B):
        for j in range(F):
            keep = [i for i in range(lens[b, j]) if not mask[b, slots[b, j, i]]]
            lens2[b, j] = len(keep)
            conf2[b, j, :len(keep)] = conf[b, j, keep]
            od2[b, j, :len(keep)] = od[b, j, keep]
    assert np.array_equal(flt.feat_len.cpu().numpy(), lens2)
    check_means(lens2, conf2, od2)
    # --- compaction: emptied tracks leave, the lists follow their tracks ---
    active, dest = flt.compact_tracks(drop_lost=False, return_map=True)
    dest = dest.cpu().numpy()
    lens3, conf3, od3 = np.zeros_like(lens2), np.zeros_like(conf2), np.zeros_like(od2)
    for b in range(B):
        for j in range(F):
            if dest[b, j] >= 0:
                lens3[b, dest[b, j]], conf3[b, dest[b, j]], od3[b, dest[b, j]] = lens2[b, j], conf2[b, j], od2[b, j]
    assert np.array_equal(flt.feat_len.cpu().numpy(), lens3)
    check_means(lens3, conf3, od3)


def test_whole_run_from_detections_with_moving_descriptors(golden_dir):
    """A whole reference run whose detections carry per-frame descriptor noise and widely spread scores
    (tests/golden/small_desc_trace.npz): the score filter drops detections, the tracks' score-weighted mean descriptors
    move every frame and lose the rows of every pruned camera.  The device loop gets the RAW detections
    (frame_step_detections(score_filter=True)); after every frame its proposals equal what the reference's matcher led to,
    the matcher's view of the tracks equals the reference's last_camera_measurement.descriptors (the means of MSCKF.py:436-441
    built after the frame's appends, minus the rows of the tracks deleted by :742-747 / :774-778; a pruned camera's
    descriptors leave a track's list at once, :764-765, and its mean at the next rebuild), and the state equals the reference's."""
    from monocular_visual_inertial_msckf_b200 import _lib
    from helpers import rel_err
    path = os.path.join(golden_dir, "small_desc_trace.npz")
    tr = o.load_trace(path)
    z = np.load(path)
    ini, cases, apps = tr["init"], tr["cases"], tr["appends"]
    nfr = len(tr["frames"])
    F = max(len(a["action"]) for a in apps) + 1
    M = max(c["obs_uv"].shape[1] for c in cases) + 1
    N = tr["max_cams"] + 1
    F2 = max(len(z[f"x{i}.scores"]) for i in range(nfr)) + 2
    D = z["x0.desc"].shape[1]
    par = UpdateParams.from_case(cases[0], _lib.COMPRESS_GRAM)
    par.gravity = np.array(ini["gravity"], float)
    B = 2
    flt = BatchedMSCKF(B, F, M, N, par)
    flt.P.zero_()
    flt.P[:, :15, :15] = torch.from_numpy(np.asarray(ini["P"], float))
    flt.imu_state[:] = torch.from_numpy(np.concatenate([np.asarray(ini["imu_R"], float).flatten(), ini["imu_t"], ini["imu_v"],
                                                        ini["imu_bg"], ini["imu_ba"]]))
    flt.imu_null[:] = torch.from_numpy(np.concatenate([np.asarray(ini["null_R"], float).flatten(), ini["null_t"], ini["null_v"]]))
    flt.imu_time[:] = float(ini["timestamp"])
    flt.Qc_diag[:] = torch.from_numpy(np.asarray(ini["Qc_diag"], float)[[0, 3, 6, 9]].copy())
    flt.n_cams.zero_()
    flt.feat_len.zero_()
    thr = float(z["min_cosine_similarity"])
    ops = tr["ops"]
    first, worst, worst_desc, start, n_filtered, n_moved = True, 0.0, 0.0, 0, 0, 0
    for i, fr in enumerate(tr["frames"]):
        seg = ops[start:int(fr["op_index"])]
        start = int(fr["op_index"])
        smp = np.array([tr["samples"][idx] for kind, idx in seg if kind == 0], float)
        a = apps[[idx for kind, idx in seg if kind == 4][0]]
        kp_raw, de_raw, sc_raw = z[f"x{i}.kp"], z[f"x{i}.desc"], z[f"x{i}.scores"]
        n2 = len(sc_raw)
        n_filtered += n2 - int(o.score_filter(sc_raw).sum())
        desc = torch.zeros((B, F2, D), dtype=torch.float32, device="cuda")
        kp = torch.zeros((B, F2, 2), dtype=torch.float64, device="cuda")
        sc = torch.zeros((B, F2), dtype=torch.float64, device="cuda")
        desc[:, :n2] = torch.from_numpy(de_raw.astype(np.float32)).cuda()
        kp[:, :n2] = torch.from_numpy(kp_raw).cuda()
        sc[:, :n2] = torch.from_numpy(sc_raw).cuda()
        n_new = torch.full((B,), n2, dtype=torch.int32, device="cuda")
        samples = torch.from_numpy(smp).reshape(1, -1, 7).repeat(B, 1, 1).cuda()
        action, verdict = flt.frame_step_detections(samples, ini["R_I_C"], ini["t_I_C"], desc, kp, sc, n_new, tr["max_cams"],
                                                    min_cosine_similarity=thr, epipolar_rejection_threshold=1e9,
                                                    homography_rejection_threshold=1e9, use_null_state=first, score_filter=True)
        first = False
        torch.cuda.synchronize()
        fa = len(a["action"])
        act = action.cpu().numpy()
        assert (act[:, :fa] == a["action"]).all() and (act[:, fa:] == 0).all(), i
        d = fr["P"].shape[0]
        nc = (d - 15) // 6
        nf = len(fr["feat_len"])
        lcm_desc, lcm_idx = z[f"x{i}.lcm_desc"], z[f"x{i}.lcm_idx"]
        assert np.array_equal(lcm_idx, fr["feat_ids"])            # the matcher's view lists exactly the live tracks, in order
        got = flt.lcm_desc.cpu().numpy()
        for b in range(B):
            assert int(flt.n_cams[b]) == nc, (i, int(flt.n_cams[b]), nc)
            ln = flt.feat_len[b].cpu().numpy()
            assert (ln[:nf] == fr["feat_len"]).all() and (ln[nf:] == 0).all(), i
            e = np.max(np.abs(got[b, :nf] - lcm_desc))
            worst_desc = max(worst_desc, e)
            assert e < 5e-7, (i, e)                               # float32 descriptors against the reference's float64 means
            assert (flt.lost_cnt[b, :nf].cpu().numpy() == fr["lost_cnt"]).all(), i
            st = flt.imu_state[b].cpu().numpy()
            errs = [rel_err(flt.P[b, :d, :d].cpu().numpy(), fr["P"]), rel_err(st[9:12], fr["imu_t"]),
                    rel_err(flt.cam_t[b, :nc].cpu().numpy(), fr["cam_t"])]
            worst = max(worst, max(errs))
            assert worst < 1e-9, (i, errs)
        n_moved += int(np.sum(np.abs(np.linalg.norm(lcm_desc, axis=1) - 1.0) > 1e-4))
    assert nfr >= 50 and n_filtered > 100 and n_moved > 500       # the filter fired and the means were not unit vectors
    print(f"frame loop from raw detections with moving descriptors: state {worst:.2e}, mean descriptors {worst_desc:.2e}, "
          f"{n_filtered} detections filtered by score")


def test_sixty_four_distinct_filters_vs_their_reference_runs(golden_dir):
    """BASELINE.json configs[2] as a parity test: 64 DISTINCT filters (IMU noise low / mid / high x 22 stream seeds, each with
    its own trajectory, landmarks, drop-outs, outliers, descriptor noise and scores) run as ONE batch through the resident
    frame loop from their raw detections; after every camera frame every filter's IMU state, window poses, track lengths and
    covariance diagonal -- and after the last frame its whole covariance -- equal what the unmodified reference produced when
    it ran that filter's stream alone (tests/golden/multi_64.npz, tools/make_golden.py --only multi)."""
    from monocular_visual_inertial_msckf_b200 import _lib
    from helpers import rel_err
    z = np.load(os.path.join(golden_dir, "multi_64.npz"))
    B = int(z["n_filters"])
    nfr = int(z["t0.n_frames"])
    assert all(int(z[f"t{i}.n_frames"]) == nfr for i in range(B))
    F, M, N, F2, D = 64, 8, int(z["par.max_cams"]) + 1, 32, z["t0.f0.desc"].shape[1]
    par = UpdateParams(K=z["K"], width=640, height=480, sigma_image=float(z["t0.sigma"]), gravity=np.array(z["t0.init.gravity"], float),
                       min_lost=int(z["par.min_lost"]), min_tracked=int(z["par.min_tracked"]), use_parallax=True,
                       min_parallax=float(z["par.min_parallax"]), compress_mode=_lib.COMPRESS_GRAM)
    flt = BatchedMSCKF(B, F, M, N, par)
    flt.P.zero_()
    flt.n_cams.zero_()
    flt.feat_len.zero_()
    for i in range(B):
        g = lambda k: np.asarray(z[f"t{i}.init.{k}"], float)  # noqa: E731
        flt.P[i, :15, :15] = torch.from_numpy(g("P"))
        flt.imu_state[i] = torch.from_numpy(np.concatenate([g("imu_R").flatten(), g("imu_t"), g("imu_v"), g("imu_bg"), g("imu_ba")]))
        flt.imu_null[i] = torch.from_numpy(np.concatenate([g("null_R").flatten(), g("null_t"), g("null_v")]))
        flt.imu_time[i] = float(z[f"t{i}.init.timestamp"])
        flt.Qc_diag[i] = torch.from_numpy(g("Qc_diag")[[0, 3, 6, 9]].copy())
    R_I_C, t_I_C = np.asarray(z["t0.init.R_I_C"], float), np.asarray(z["t0.init.t_I_C"], float)
    thr = float(z["par.min_cosine_similarity"])
    prev = [0] * B
    worst, levels_seen, n_updates_like = 0.0, set(), 0
    for k in range(nfr):
        ns = [int(z[f"t{i}.f{k}.n_samples"]) for i in range(B)]
        n = ns[0] - prev[0]
        assert all(ns[i] - prev[i] == n for i in range(B)), "every stream delivers the same number of IMU samples per frame"
        smp = np.stack([z[f"t{i}.samples"][prev[i]:ns[i]] for i in range(B)])
        prev = ns
        desc = torch.zeros((B, F2, D), dtype=torch.float32, device="cuda")
        kp = torch.zeros((B, F2, 2), dtype=torch.float64, device="cuda")
        sc = torch.zeros((B, F2), dtype=torch.float64, device="cuda")
        n_new = torch.zeros(B, dtype=torch.int32, device="cuda")
        for i in range(B):
            m = len(z[f"t{i}.f{k}.scores"])
            desc[i, :m] = torch.from_numpy(z[f"t{i}.f{k}.desc"]).cuda()
            kp[i, :m] = torch.from_numpy(z[f"t{i}.f{k}.kp"]).cuda()
            sc[i, :m] = torch.from_numpy(z[f"t{i}.f{k}.scores"]).cuda()
            n_new[i] = m
        flt.frame_step_detections(torch.from_numpy(smp).cuda(), R_I_C, t_I_C, desc, kp, sc, n_new, int(z["par.max_cams"]),
                                  min_cosine_similarity=thr, epipolar_rejection_threshold=1e9, homography_rejection_threshold=1e9,
                                  use_null_state=(k == 0), score_filter=True)
        torch.cuda.synchronize()
        st = flt.imu_state.cpu().numpy()
        nc = flt.n_cams.cpu().numpy()
        ln = flt.feat_len.cpu().numpy()
        ct = flt.cam_t.cpu().numpy()
        Pd = torch.diagonal(flt.P, dim1=1, dim2=2).cpu().numpy()
        for i in range(B):
            pre = f"t{i}.f{k}."
            ref_ct, ref_len = z[pre + "cam_t"], z[pre + "feat_len"]
            assert nc[i] == len(ref_ct), (k, i, nc[i], len(ref_ct))
            assert (ln[i, :len(ref_len)] == ref_len).all() and (ln[i, len(ref_len):] == 0).all(), (k, i)
            d = 15 + 6 * nc[i]
            errs = [rel_err(st[i], z[pre + "imu"]), rel_err(ct[i, :nc[i]], ref_ct) if nc[i] else 0.0, rel_err(Pd[i, :d], z[pre + "P_diag"])]
            worst = max(worst, max(errs))
            assert worst < 1e-9, (k, i, errs)
    assert int(flt._overflow.item()) == 0
    flt.check_window_overflow()
    for i in range(B):
        Pl = z[f"t{i}.P_last"]
        assert rel_err(flt.P[i, :Pl.shape[0], :Pl.shape[0]].cpu().numpy(), Pl) < 1e-9, i
    gates = [int(z[f"t{i}.f{nfr - 1}.n_gate"]) for i in range(B)]
    assert sum(gates) > 50 and len(set(tuple(z[f"t{i}.f{nfr - 1}.feat_len"]) for i in range(B))) > 32      # distinct, with rejections
    print(f"64 distinct filters x {nfr} frames as one batch vs their reference runs: worst relative error {worst:.2e}, "
          f"{sum(gates)} gate rejections in the reference runs")


@pytest.mark.gpu
def test_frame_graph_replay_equals_eager_frames():
    """BatchedMSCKF.capture_frame_graph: 24 camera frames of 6 heterogeneous filters (device-built streams, window cycling
    through the pruning) replayed from ONE captured CUDA graph leave exactly the bits the eager frame loop leaves -- state,
    covariance, window, tracks, descriptor lists -- and the capture itself does not touch the filter state."""
    from monocular_visual_inertial_msckf_b200 import _lib
    from monocular_visual_inertial_msckf_b200.streams import DeviceStreams
    B, L, WINDOW, NF = 6, 64, 8, 24
    st = DeviceStreams(range(B), NF + 1, n_landmarks=L, p_drop=0.05, p_outlier=0.03, desc_dim=32, seed=11)
    par = UpdateParams(K=st.K, width=st.width, height=st.height, sigma_image=0.1, min_lost=2, min_tracked=3, use_parallax=True,
                       min_parallax=0.0, compress_mode=_lib.COMPRESS_GRAM)

    def fresh():
        f = BatchedMSCKF(B, L + 16, WINDOW + 4, WINDOW + 2, par)
        f.P.zero_(); f.n_cams.zero_(); f.feat_len.zero_()
        f.Qc_diag.copy_(st.Qc_diag)
        return f

    kw = dict(min_cosine_similarity=0.82, epipolar_rejection_threshold=1e9, homography_rejection_threshold=1e9)
    eager, graphed = fresh(), fresh()
    before = graphed.snapshot()
    fg = graphed.capture_frame_graph(*st.frame_inputs(0)[:1], st.R_I_C, st.t_I_C, *st.frame_inputs(0)[1:], WINDOW, **kw)
    after = graphed.snapshot()
    assert torch.equal(before["P"], after["P"]) and torch.equal(before["small_slab"], after["small_slab"])
    assert all(torch.equal(before["tracks"][k], after["tracks"][k]) for k in before["tracks"])
    for f in range(NF):
        smp, desc, kp, sc, cnt = st.frame_inputs(f)
        eager.frame_step_detections(smp, st.R_I_C, st.t_I_C, desc, kp, sc, cnt, WINDOW, **kw)
        fg.step(smp, desc, kp, sc, cnt)
    torch.cuda.synchronize()
    eager.check_window_overflow()
    graphed.check_window_overflow()
    assert int(eager.n_cams.max()) >= WINDOW                 # the window filled up and was pruned on the way
    assert torch.equal(eager.n_cams, graphed.n_cams) and torch.equal(eager.feat_len, graphed.feat_len)
    assert torch.equal(eager.P, graphed.P) and torch.equal(eager.imu_state, graphed.imu_state)
    assert torch.equal(eager.obs_uv, graphed.obs_uv) and torch.equal(eager.obs_slot, graphed.obs_slot)
    assert torch.equal(eager.obs_desc, graphed.obs_desc) and torch.equal(eager.lcm_desc, graphed.lcm_desc)
    # the running descriptor sums (appended to / subtracted from over 24 frames of matches, prunings and track deletions)
    # still give the rows a fresh pass over the lists gives
    a, b = eager.track_descriptors(), eager.track_descriptors(from_lists=True)
    assert float((a - b).abs().max()) <= 1e-6 and float(b.abs().max()) > 0.01
