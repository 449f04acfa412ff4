"""GPU parity tests: the CUDA path (through the C ABI, via the ctypes binding) against the golden vectors captured
from the unmodified reference and against the NumPy oracle on the same inputs.

Tolerances (BASELINE.json north_star): state and covariance within 1e-9 relative (relative to the scale of the
quantity: max|a-b| / max|b|) in FP64; gating decisions identical away from the threshold (|gamma/chi2 - 1| > 1e-6).
"""
import os

import numpy as np
import pytest
import torch

from helpers import compare_update_outputs, rel_err
from oracle import msckf_oracle as o

pytestmark = pytest.mark.gpu

from monocular_visual_inertial_msckf_b200 import BatchedMSCKF, UpdateParams, _lib  # noqa: E402

CHI = o.chi2_table(512)


def _load(golden_dir, name):
    return o.load_cases(os.path.join(golden_dir, name))


def _batch_from_cases(cases, mode=_lib.COMPRESS_QR, pad_F=0, pad_M=0, pad_N=0):
    F = max(len(c["feat_len"]) for c in cases) + pad_F
    M = max(c["obs_uv"].shape[1] for c in cases) + pad_M
    N = max(int(c["n_cams"]) for c in cases) + pad_N
    flt = BatchedMSCKF(len(cases), F, M, N, UpdateParams.from_case(cases[0], mode))
    flt.load_cases(cases)
    return flt


def _summarise(all_errs):
    keys = sorted({k for e in all_errs for k in e})
    return {k: max(e.get(k, 0) for e in all_errs) for k in keys}


MODES = [pytest.param(_lib.COMPRESS_QR, id="qr"), pytest.param(_lib.COMPRESS_GRAM, id="gram")]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", ["small_updates.npz", "mid_48x10_updates.npz", "baseline_256x30_updates.npz"])
def test_update_vs_golden(golden_dir, name, mode):
    cases, _ = _load(golden_dir, name)
    flt = _batch_from_cases(cases, mode)
    flt.update()
    torch.cuda.synchronize()
    errs = []
    for b, c in enumerate(cases):
        out = flt.case_outputs(b, c)
        errs.append(compare_update_outputs(out, c, CHI, label=f"{name}#{c['call_index']}"))
    print(name, "gram" if mode else "qr", _summarise(errs))


@pytest.mark.parametrize("pad_M", [pytest.param(2, id="split-assembly+factorisation"),
                                   pytest.param(30, id="cta-per-feature")])
def test_front_stage_vs_golden(golden_dir, pad_M):
    """rows 3-8 alone: selection, triangulation, gamma/dof/gate per feature, state untouched.  The observation capacity
    picks the kernel: <= 32 observations per track -> warp-per-feature kernels (assembly + factorisation pair), more ->
    CTA-per-feature kernel."""
    cases, _ = _load(golden_dir, "small_updates.npz")
    flt = _batch_from_cases(cases, pad_F=3, pad_M=pad_M, pad_N=1)
    assert (flt.M <= 32) == (pad_M == 2)
    P0 = flt.P.clone()
    flt.front()
    torch.cuda.synchronize()
    assert torch.equal(P0, flt.P)
    n_checked = 0
    for b, c in enumerate(cases):
        out = flt.case_outputs(b, c)
        v = c["out_valid"]
        assert (out["out_valid"] == v).all() and (out["out_lost"] == c["out_lost"]).all()
        assert (out["out_dof"][v] == c["out_dof"][v]).all()
        assert rel_err(out["out_idp_m"], c["out_idp_m"]) < 1e-9 and rel_err(out["out_idp_rho"], c["out_idp_rho"]) < 1e-9
        if v.any():
            g = c["out_gamma"][v]
            assert np.max(np.abs(out["out_gamma"][v] - g) / np.maximum(np.abs(g), 1e-3)) < 1e-7
            n_checked += int(v.sum())
        F = len(v)
        assert (flt.feat_status[b, F:] == 0).all()           # padded slots are skipped
    assert n_checked > 1000


@pytest.mark.parametrize("mode", MODES)
def test_compress_invariants_vs_oracle(golden_dir, mode):
    """row 9: T_H^T T_H = sum_j H_o^T H_o and T_H^T r_n = sum_j H_o^T r_o (the basis-invariant content of the QR);
    in information form the compression stage returns exactly those two sums (G, g)."""
    cases, _ = _load(golden_dir, "mid_48x10_updates.npz")
    flt = _batch_from_cases(cases, mode, pad_N=1 if mode else 0)
    flt.front()
    T_H, r_n = flt.compress()
    torch.cuda.synchronize()
    for b, c in enumerate(cases):
        ref = o.update_case(c, return_intermediates=True)
        n = int(c["n_cams"])
        cc = 6 * n
        acc = [i for i, j in enumerate(np.nonzero(ref["out_valid"])[0]) if ref["out_gate"][j]]
        G = sum(ref["inter"]["HtH"][i] for i in acc)[15:, 15:]
        g = sum(ref["inter"]["Htr"][i] for i in acc)[15:]
        T = T_H[b, :cc, :cc].cpu().numpy()
        rn = r_n[b, :cc].cpu().numpy()
        if mode == _lib.COMPRESS_QR:
            assert np.allclose(np.tril(T, -1), 0.0)
            assert rel_err(T.T @ T, G) < 1e-11
            assert rel_err(T.T @ rn, g) < 1e-10
        else:
            assert np.array_equal(T, T.T)
            assert rel_err(T, G) < 1e-12
            assert rel_err(rn, g) < 1e-11
        assert int(flt.n_accepted[b]) == len(acc)
        assert (T_H[b, cc:, :] == 0).all() and (T_H[b, :, cc:] == 0).all()


@pytest.mark.parametrize("mode", MODES)
def test_update_matches_oracle_on_perturbed_inputs(golden_dir, mode):
    """Same seeded perturbation handed to the oracle and to the GPU (inputs the reference run never saw)."""
    cases, _ = _load(golden_dir, "mid_48x10_updates.npz")
    rng = np.random.default_rng(7)
    pert = []
    for c in cases:
        c = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in c.items()}
        c["obs_uv"] = c["obs_uv"] + 0.3 * rng.standard_normal(c["obs_uv"].shape)
        bad = rng.random(c["obs_uv"].shape[:2]) < 0.03
        c["obs_uv"][bad] += 80.0 * rng.standard_normal((int(bad.sum()), 2))
        c["cam_t"] = c["cam_t"] + 1e-3 * rng.standard_normal(c["cam_t"].shape)
        ref = o.update_case(c)
        c.update(ref)
        c["out_n_rejected"] = ref["out_n_rejected"]
        pert.append(c)
    flt = _batch_from_cases(pert, mode)
    flt.update()
    torch.cuda.synchronize()
    n_rej = 0
    for b, c in enumerate(pert):
        compare_update_outputs(flt.case_outputs(b, c), c, CHI, label=f"perturbed#{b}")
        n_rej += c["out_n_rejected"]
    assert n_rej > 0


@pytest.mark.parametrize("mode", MODES)
def test_edge_cases(golden_dir, mode):
    cases, _ = _load(golden_dir, "mid_48x10_updates.npz")
    base = cases[0]

    def clone():
        return {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in base.items()}

    # (a) every track cut to one observation: dof 0 -> all rejected, state untouched (MSCKF.py:566-585)
    a = clone()
    a["feat_len"][:] = 1
    a["lost_cnt"][:] = 5
    a["tracked_cnt"][:] = 10
    # (b) nothing selected: no lost tracks and parallax threshold unreachable
    b_ = clone()
    b_["lost_cnt"][:] = 0
    b_["min_parallax"] = 1e9
    # (c) lost before tracked long enough: discarded, flagged lost, no update
    c_ = clone()
    c_["lost_cnt"][:] = 5
    c_["tracked_cnt"][:] = 1
    # (d) zero baseline: every window pose identical and two identical rays per track (parallel rays, rank-deficient
    #     H_f -> the null space grows to 2M - 2, SURVEY.md section 8a edge cases)
    d_ = clone()
    d_["feat_len"][:] = 2
    d_["cam_R"][:] = d_["cam_R"][0]
    d_["cam_t"][:] = d_["cam_t"][0]
    d_["obs_uv"][:, 1] = d_["obs_uv"][:, 0] + np.array([1.5, -0.7])
    d_["obs_dir"][:, 1] = d_["obs_dir"][:, 0]
    d_["lost_cnt"][:] = 5
    d_["tracked_cnt"][:] = 10
    # (e) empty filter (no features at all)
    e_ = clone()
    e_["feat_len"][:] = 0
    batch = []
    for c in (a, b_, c_, d_, e_):
        # the per-case params differ (min_parallax): run them as separate single-filter batches
        ref = o.update_case(c) if c["feat_len"].max() > 0 else None
        flt = _batch_from_cases([c], mode)
        flt.update()
        torch.cuda.synchronize()
        out = flt.case_outputs(0, c)
        batch.append((c, ref, out))
    (a, ra, oa), (b_, rb, ob), (c_, rc, oc), (d_, rd, od), (e_, _, oe) = batch
    assert oa["out_valid"].all() and not oa["out_gate"].any() and oa["out_n_rejected"] == len(a["feat_len"])
    assert rel_err(oa["out_P"], a["P"]) == 0.0 and (oa["out_delta_x"] == 0).all()
    assert not ob["out_valid"].any() and not ob["out_updated"] and rel_err(ob["out_P"], b_["P"]) == 0.0
    assert not oc["out_valid"].any() and oc["out_lost"].all() and rel_err(oc["out_P"], c_["P"]) == 0.0
    assert (od["out_dof"] == rd["out_dof"]).all() and (rd["out_dof"][rd["out_valid"]] == 2).all()
    # information form on a rank-deficient, zero-baseline problem: delta_x (1e-4 of the state) carries ~1e-9 of ITS OWN
    # scale, i.e. ~1e-13 of the state, which is what the 1e-9 bar is about
    compare_update_outputs(od, rd, CHI, tol_dx=1e-9 if mode == _lib.COMPRESS_QR else 1e-8, label="rank-deficient")
    assert not oe["out_valid"].any() and not oe["out_updated"]


@pytest.mark.parametrize("name", ["small_propagate.npz", "baseline_256x30_propagate.npz"])
def test_propagate_vs_golden(golden_dir, name):
    cases, _ = _load(golden_dir, name)
    B = len(cases)
    N = max((c["P"].shape[0] - 15) // 6 for c in cases)
    N = max(N, 1)
    flt = BatchedMSCKF(B, 1, 2, N, UpdateParams(gravity=np.array(cases[0]["gravity"], float)))
    samples = torch.zeros(B, 1, 7, dtype=torch.float64, device="cuda")
    for b, c in enumerate(cases):
        d = c["P"].shape[0]
        flt.P[b, :d, :d] = torch.from_numpy(c["P"])
        flt.n_cams[b] = (d - 15) // 6
        flt.imu_state[b] = torch.from_numpy(np.concatenate([c["imu_R"].flatten(), c["imu_t"], c["imu_v"], c["imu_bg"], c["imu_ba"]]))
        flt.imu_null[b] = torch.from_numpy(np.concatenate([c["null_R"].flatten(), c["null_t"], c["null_v"]]))
        flt.imu_time[b] = float(c["timestamp"])
        flt.Qc_diag[b] = torch.from_numpy(np.asarray(c["Qc_diag"])[[0, 3, 6, 9]].copy())
        samples[b, 0] = torch.from_numpy(np.asarray(c["sample"], float))
    flt.propagate(samples, use_null_state=True)
    torch.cuda.synchronize()
    worst = 0.0
    for b, c in enumerate(cases):
        d = c["P"].shape[0]
        st = flt.imu_state[b].cpu().numpy()
        e = max(rel_err(st[:9].reshape(3, 3), c["out_imu_R"]), rel_err(st[9:12], c["out_imu_t"]),
                rel_err(st[12:15], c["out_imu_v"]), rel_err(flt.P[b, :d, :d].cpu().numpy(), c["out_P"]))
        worst = max(worst, e)
        assert e < 1e-12, (name, c["call_index"], e)
        assert np.array_equal(flt.imu_null[b].cpu().numpy(), st[:15])      # null state <- state (MSCKF.py:247-248)
    print(name, "worst rel err", worst)


def test_propagate_fused_samples_match_sequential_oracle(golden_dir):
    """n samples in one launch == n reference process_imu calls (null state aliasing included)."""
    cases, _ = _load(golden_dir, "mid_48x10_propagate.npz")
    c0 = cases[0]
    d = c0["P"].shape[0]
    n = len(cases)
    assert all(cases[i + 1]["call_index"] == cases[i]["call_index"] + 1 for i in range(n - 1))
    flt = BatchedMSCKF(2, 1, 2, (d - 15) // 6, UpdateParams(gravity=np.array(c0["gravity"], float)))
    samples = torch.zeros(2, n, 7, dtype=torch.float64, device="cuda")
    for b in range(2):
        flt.P[b, :d, :d] = torch.from_numpy(c0["P"])
        flt.n_cams[b] = (d - 15) // 6
        flt.imu_state[b] = torch.from_numpy(np.concatenate([c0["imu_R"].flatten(), c0["imu_t"], c0["imu_v"], c0["imu_bg"], c0["imu_ba"]]))
        flt.imu_null[b] = torch.from_numpy(np.concatenate([c0["null_R"].flatten(), c0["null_t"], c0["null_v"]]))
        flt.imu_time[b] = float(c0["timestamp"])
        flt.Qc_diag[b] = torch.from_numpy(np.asarray(c0["Qc_diag"])[[0, 3, 6, 9]].copy())
        for i in range(n):
            samples[b, i] = torch.from_numpy(np.asarray(cases[i]["sample"], float))
    flt.propagate(samples, use_null_state=True)
    torch.cuda.synchronize()
    last = cases[-1]          # consecutive reference calls: the golden output of the last one is the chained result
    for b in range(2):
        st = flt.imu_state[b].cpu().numpy()
        assert rel_err(flt.P[b, :d, :d].cpu().numpy(), last["out_P"]) < 1e-12
        assert rel_err(st[:9].reshape(3, 3), last["out_imu_R"]) < 1e-13
        assert rel_err(st[9:12], last["out_imu_t"]) < 1e-13 and rel_err(st[12:15], last["out_imu_v"]) < 1e-13


def test_window_surgery_vs_oracle(golden_dir):
    """next rows: state_augmentation and remove_cameras on device against the oracle restatement."""
    cases, _ = _load(golden_dir, "mid_48x10_updates.npz")
    c = cases[0]
    n = int(c["n_cams"])
    flt = _batch_from_cases([c, c], pad_N=2)
    R_I_C = np.array([[0.0, 0, 1], [-1, 0, 0], [0, -1, 0]])
    t_I_C = np.array([0.05, -0.02, 0.01])
    flt.augment(R_I_C, t_I_C)
    torch.cuda.synchronize()
    P_ref = o.augment_covariance(np.array(c["P"], float), R_I_C, t_I_C, np.array(c["imu_R"], float))
    d1 = P_ref.shape[0]
    assert int(flt.n_cams[0]) == n + 1
    assert rel_err(flt.P[0, :d1, :d1].cpu().numpy(), P_ref) < 1e-13
    Rc = np.array(c["imu_R"]) @ R_I_C
    tc = np.array(c["imu_R"]) @ t_I_C + np.array(c["imu_t"])
    assert rel_err(flt.cam_R[0, n].cpu().numpy().reshape(3, 3), Rc) < 1e-15
    assert rel_err(flt.cam_t[0, n].cpu().numpy(), tc) < 1e-15
    mask = torch.zeros(2, flt.N, dtype=torch.uint8, device="cuda")
    mask[:, 1] = 1
    mask[:, 4] = 1
    flt.remove_cameras(mask)
    torch.cuda.synchronize()
    P2 = o.remove_camera_covariance(o.remove_camera_covariance(P_ref, 4), 1)
    d2 = P2.shape[0]
    assert int(flt.n_cams[1]) == n - 1
    assert np.array_equal(flt.P[1, :d2, :d2].cpu().numpy(), flt.P[0, :d2, :d2].cpu().numpy())
    assert rel_err(flt.P[0, :d2, :d2].cpu().numpy(), P2) < 1e-13
    keep = [s for s in range(n + 1) if s not in (1, 4)]
    cam_t_ref = np.concatenate([np.asarray(c["cam_t"]), tc[None]])[keep]
    assert rel_err(flt.cam_t[0, :len(keep)].cpu().numpy(), cam_t_ref) < 1e-15


def _load_tracks(flt, cases):
    """tracks + window positions of packed surgery cases (tools/make_golden.py::pack_tracks) into a BatchedMSCKF"""
    for b, c in enumerate(cases):
        f, m, n = c["obs_slot"].shape[0], c["obs_slot"].shape[1], int(c["n_cams"])
        flt.n_cams[b] = n
        flt.feat_len[b].zero_()
        flt.feat_len[b, :f] = torch.from_numpy(c["feat_len"])
        flt.obs_slot[b, :f, :m] = torch.from_numpy(np.maximum(c["obs_slot"], 0))
        flt.obs_uv[b, :f, :m] = torch.from_numpy(c["obs_uv"])
        flt.obs_dir[b, :f, :m] = torch.from_numpy(c["obs_dir"])
        flt.obs_conf[b, :f, :m] = torch.from_numpy(c["obs_conf"])
        flt.idp_base[b, :f] = torch.from_numpy(c["idp_base"])
        flt.idp_anchor_slot[b, :f] = torch.from_numpy(c["idp_anchor_slot"])
        flt.cam_t[b, :n] = torch.from_numpy(c["cam_t"])


def _surgery_batch(cases):
    F = max(c["obs_slot"].shape[0] for c in cases) + 3
    M = max(c["obs_slot"].shape[1] for c in cases) + 1
    N = max(int(c["n_cams"]) for c in cases) + 1
    flt = BatchedMSCKF(len(cases), F, M, N, UpdateParams())
    _load_tracks(flt, cases)
    return flt


def test_prune_selection_vs_golden(golden_dir):
    """next rows: the camera / track selection of prune_poorest_camera_states on device, bit-exact against every
    selection the reference made in the small run (ties included)."""
    cases, _ = _load(golden_dir, "small_prune_select.npz")
    flt = _surgery_batch(cases)
    mask, touch = flt.select_poorest_cameras(2)
    torch.cuda.synchronize()
    mask, touch = mask.cpu().numpy().astype(bool), touch.cpu().numpy().astype(bool)
    for b, c in enumerate(cases):
        n, f = int(c["n_cams"]), len(c["feat_len"])
        assert (mask[b, :n] == c["out_remove_mask"]).all() and not mask[b, n:].any(), c["call_index"]
        assert (touch[b, :f] == c["out_feat_touch"]).all() and not touch[b, f:].any(), c["call_index"]


def test_drop_observations_vs_golden(golden_dir):
    """next rows: the track part of remove_cameras on device (compaction in place, slot renumbering, frozen anchors),
    bit-exact against every remove_cameras call of the small run; then the window / covariance part on the same mask."""
    cases, _ = _load(golden_dir, "small_prune_drop.npz")
    flt = _surgery_batch(cases)
    mask = torch.zeros((flt.B, flt.N), dtype=torch.uint8, device="cuda")
    for b, c in enumerate(cases):
        mask[b, :int(c["n_cams"])] = torch.from_numpy(c["remove_mask"].astype(np.uint8))
    flt.drop_observations(mask)
    flt.remove_cameras(mask)
    torch.cuda.synchronize()
    ln, sl = flt.feat_len.cpu().numpy(), flt.obs_slot.cpu().numpy()
    uv, dr, cf = flt.obs_uv.cpu().numpy(), flt.obs_dir.cpu().numpy(), flt.obs_conf.cpu().numpy()
    an, bs = flt.idp_anchor_slot.cpu().numpy(), flt.idp_base.cpu().numpy()
    n_checked = 0
    for b, c in enumerate(cases):
        f = len(c["feat_len"])
        assert (ln[b, :f] == c["out_feat_len"]).all() and (ln[b, f:] == 0).all(), c["call_index"]
        assert int(flt.n_cams[b]) == int(c["out_n_cams"])
        for j in range(f):
            n = int(c["out_feat_len"][j])
            if n == 0:
                continue
            assert (sl[b, j, :n] == c["out_obs_slot"][j, :n]).all()
            assert np.array_equal(uv[b, j, :n], c["out_obs_uv"][j, :n]) and np.array_equal(dr[b, j, :n], c["out_obs_dir"][j, :n])
            assert np.array_equal(cf[b, j, :n], c["out_obs_conf"][j, :n])
            assert an[b, j] == c["out_idp_anchor_slot"][j] and np.array_equal(bs[b, j], c["out_idp_base"][j])
            n_checked += n
    assert n_checked > 5000


def test_track_append_vs_golden(golden_dir):
    """next rows: the track bookkeeping of add_camera_measurements on device (append / lost / new track) against what
    the reference did to its tracks in 30 frames of the small run."""
    cases, _ = _load(golden_dir, "small_track_append.npz")
    F = max(len(c["feat_ids"]) for c in cases) + 2
    M = max(c["obs_slot"].shape[1] for c in cases)
    N = max(int(c["n_cams"]) for c in cases)
    flt = BatchedMSCKF(len(cases), F, M, N, UpdateParams(K=np.array(cases[0]["K"], float)))
    action = torch.zeros((flt.B, F), dtype=torch.uint8, device="cuda")
    new_uv = torch.zeros((flt.B, F, 2), dtype=torch.float64, device="cuda")
    new_conf = torch.zeros((flt.B, F), dtype=torch.float64, device="cuda")
    _load_tracks(flt, cases)
    for b, c in enumerate(cases):
        f, n = len(c["feat_ids"]), int(c["n_cams"])
        flt.cam_R[b, :n] = torch.from_numpy(np.asarray(c["cam_R"], float).reshape(n, 9))
        flt.idp_m[b, :f] = torch.from_numpy(c["idp_m"])
        flt.idp_rho[b, :f] = torch.from_numpy(c["idp_rho"])
        flt.lost_cnt[b, :f] = torch.from_numpy(c["lost_cnt"])
        flt.tracked_cnt[b, :f] = torch.from_numpy(c["tracked_cnt"])
        action[b, :f] = torch.from_numpy(c["action"])
        new_uv[b, :f] = torch.from_numpy(c["new_uv"])
        new_conf[b, :f] = torch.from_numpy(c["new_conf"])
    overflow = flt.append_observations(action, new_uv, new_conf)
    torch.cuda.synchronize()
    assert int(overflow.item()) == 0
    got = {k: getattr(flt, k).cpu().numpy() for k in ("feat_len", "obs_uv", "obs_dir", "obs_conf", "obs_slot", "idp_base",
                                                      "idp_anchor_slot", "idp_m", "idp_rho", "lost_cnt", "tracked_cnt")}
    n_app = 0
    for b, c in enumerate(cases):
        f = len(c["feat_ids"])
        for k in ("feat_len", "idp_anchor_slot", "lost_cnt", "tracked_cnt", "idp_rho"):
            assert np.array_equal(got[k][b, :f], c["out_" + k]), (c["call_index"], k)
        assert (got["feat_len"][b, f:] == 0).all()
        for j in range(f):
            n = int(c["out_feat_len"][j])
            assert (got["obs_slot"][b, j, :n] == c["out_obs_slot"][j, :n]).all()
            assert np.array_equal(got["obs_uv"][b, j, :n], c["out_obs_uv"][j, :n])
            assert np.array_equal(got["obs_conf"][b, j, :n], c["out_obs_conf"][j, :n])
            assert rel_err(got["obs_dir"][b, j, :n], c["out_obs_dir"][j, :n]) < 1e-14
            assert np.array_equal(got["idp_base"][b, j], c["out_idp_base"][j])
            assert rel_err(got["idp_m"][b, j], c["out_idp_m"][j]) < 1e-14
            n_app += int(c["action"][j] >= 2)
    assert n_app > 1000
    # a full track is reported, not overrun
    full = BatchedMSCKF(1, 2, 2, 2, UpdateParams())
    full.n_cams.fill_(2)
    full.feat_len[0, 0] = 2
    act = torch.tensor([[2, 0]], dtype=torch.uint8, device="cuda")
    ov = full.append_observations(act, torch.zeros((1, 2, 2), dtype=torch.float64, device="cuda"),
                                  torch.zeros((1, 2), dtype=torch.float64, device="cuda"))
    torch.cuda.synchronize()
    assert int(ov.item()) == 1 and int(full.feat_len[0, 0]) == 2


@pytest.mark.parametrize("mode", MODES)
def test_whole_run_trace_chained_on_device(golden_dir, mode):
    """Sequence-level parity: a whole reference run replayed call by call with the filter state RESIDENT on the device
    (540 process_imu, 59 state_augmentation, 85 get_valid_features + update + correct, 79 remove_cameras; only the tracks
    of each update come from the host) reproduces the reference's state after every camera frame to 1e-9."""
    tr = o.load_trace(os.path.join(golden_dir, "small_trace.npz"))
    ini, cases = tr["init"], tr["cases"]
    F = max(len(c["feat_len"]) for c in cases)
    M = max(c["obs_uv"].shape[1] for c in cases)
    N = max(int(c["n_cams"]) for c in cases)
    par = UpdateParams.from_case(cases[0], mode)
    par.gravity = np.array(ini["gravity"], float)
    flt = BatchedMSCKF(1, F, M, N, par)
    flt.P.zero_()
    flt.P[0, :15, :15] = torch.from_numpy(np.asarray(ini["P"], float))
    flt.imu_state[0] = torch.from_numpy(np.concatenate([np.asarray(ini["imu_R"], float).flatten(), ini["imu_t"], ini["imu_v"],
                                                        ini["imu_bg"], ini["imu_ba"]]))
    flt.imu_null[0] = torch.from_numpy(np.concatenate([np.asarray(ini["null_R"], float).flatten(), ini["null_t"], ini["null_v"]]))
    flt.imu_time[0] = float(ini["timestamp"])
    flt.Qc_diag[0] = torch.from_numpy(np.asarray(ini["Qc_diag"], float)[[0, 3, 6, 9]].copy())
    flt.n_cams.zero_()
    boundaries = {int(fr["op_index"]): i for i, fr in enumerate(tr["frames"])}
    first, worst, n_updates = True, 0.0, 0
    for n, (kind, idx) in enumerate(tr["ops"]):
        if kind == 0:
            smp = torch.from_numpy(np.asarray(tr["samples"][idx], float)).reshape(1, 1, 7).cuda()
            flt.propagate(smp, use_null_state=first)     # afterwards the null state IS the state (aliasing, MSCKF.py:247-248)
            first = False
        elif kind == 1:
            flt.augment(ini["R_I_C"], ini["t_I_C"])
        elif kind == 2:
            c = cases[idx]
            assert int(flt.n_cams[0]) == int(c["n_cams"])
            f, m = len(c["feat_len"]), c["obs_uv"].shape[1]
            flt.feat_len.zero_()
            if f:
                flt.feat_len[0, :f] = torch.from_numpy(c["feat_len"])
                flt.obs_uv[0, :f, :m] = torch.from_numpy(c["obs_uv"])
                flt.obs_dir[0, :f, :m] = torch.from_numpy(c["obs_dir"])
                flt.obs_conf[0, :f, :m] = torch.from_numpy(c["obs_conf"])
                flt.obs_slot[0, :f, :m] = torch.from_numpy(np.maximum(c["obs_slot"], 0))
                flt.idp_base[0, :f] = torch.from_numpy(c["idp_base"])
                flt.idp_anchor_slot[0, :f] = torch.from_numpy(c["idp_anchor_slot"])
                flt.idp_m[0, :f] = torch.from_numpy(c["idp_m"])
                flt.idp_rho[0, :f] = torch.from_numpy(c["idp_rho"])
                flt.lost_cnt[0, :f] = torch.from_numpy(c["lost_cnt"])
                flt.tracked_cnt[0, :f] = torch.from_numpy(c["tracked_cnt"])
            flt.update()
            code = flt.feat_status[0, :f].cpu().numpy() & 0xF
            assert (((code == 1) | (code == 2)) == c["out_valid"]).all()
            n_updates += int(flt.n_accepted[0]) > 0
        elif kind == 3:
            mask = torch.zeros((1, flt.N), dtype=torch.uint8, device="cuda")
            mk = tr["masks"][idx]
            mask[0, :len(mk)] = torch.from_numpy(mk.astype(np.uint8))
            flt.remove_cameras(mask)
        if (n + 1) in boundaries:
            fr = tr["frames"][boundaries[n + 1]]
            d = fr["P"].shape[0]
            nc = (d - 15) // 6
            assert int(flt.n_cams[0]) == nc
            st = flt.imu_state[0].cpu().numpy()
            errs = [rel_err(flt.P[0, :d, :d].cpu().numpy(), fr["P"]), rel_err(st[:9].reshape(3, 3), fr["imu_R"]),
                    rel_err(st[9:12], fr["imu_t"]), rel_err(st[12:15], fr["imu_v"]), rel_err(st[15:18], fr["imu_bg"]),
                    rel_err(st[18:21], fr["imu_ba"])]
            if nc:
                errs += [rel_err(flt.cam_R[0, :nc].cpu().numpy().reshape(nc, 3, 3), fr["cam_R"]),
                         rel_err(flt.cam_t[0, :nc].cpu().numpy(), fr["cam_t"])]
            worst = max(worst, max(errs))
            assert worst < 1e-9, (boundaries[n + 1], errs)
    assert n_updates >= 40
    print("chained device replay, worst relative error over", len(tr["frames"]), "frames:", worst)


def test_whole_run_resident_frame_loop(golden_dir):
    """The whole camera-frame loop on the device (BatchedMSCKF.frame_step: propagation, state augmentation, track
    bookkeeping, process_features with deletion of the lost tracks and of the unobserved cameras, pruning of the two
    poorest cameras with its last update) driven only by the IMU samples and the matcher's decisions of a reference run:
    after every camera frame the filter state AND the tracks (order, lengths, slots, counters, anchors) equal the
    reference's."""
    tr = o.load_trace(os.path.join(golden_dir, "small_trace.npz"))
    ini, cases, apps = tr["init"], tr["cases"], tr["appends"]
    F = max(len(a["action"]) for a in apps) + 1
    M = max(c["obs_uv"].shape[1] for c in cases) + 1
    N = tr["max_cams"] + 1
    par = UpdateParams.from_case(cases[0], _lib.COMPRESS_GRAM)
    par.gravity = np.array(ini["gravity"], float)
    B = 2                                                   # two identical filters: the batch path, not a special case
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
        app = [idx for kind, idx in seg if kind == 4]
        assert sum(1 for kind, _ in seg if kind == 1) == 1 and len(app) == 1
        a = apps[app[0]]
        fa = len(a["action"])
        action = torch.zeros((B, F), dtype=torch.uint8, device="cuda")
        new_uv = torch.zeros((B, F, 2), dtype=torch.float64, device="cuda")
        new_conf = torch.zeros((B, F), dtype=torch.float64, device="cuda")
        action[:, :fa] = torch.from_numpy(a["action"]).cuda()
        new_uv[:, :fa] = torch.from_numpy(a["new_uv"]).cuda()
        new_conf[:, :fa] = torch.from_numpy(a["new_conf"]).cuda()
        samples = torch.from_numpy(smp).reshape(1, -1, 7).repeat(B, 1, 1).cuda()
        flt.frame_step(samples, ini["R_I_C"], ini["t_I_C"], action, new_uv, new_conf, tr["max_cams"], use_null_state=first)
        first = False
        torch.cuda.synchronize()
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
            assert (flt.idp_anchor_slot[b, :nf].cpu().numpy() == fr["idp_anchor_slot"]).all(), i
            st = flt.imu_state[b].cpu().numpy()
            errs = [rel_err(flt.P[b, :d, :d].cpu().numpy(), fr["P"]), rel_err(st[:9].reshape(3, 3), fr["imu_R"]),
                    rel_err(st[9:12], fr["imu_t"]), rel_err(st[12:15], fr["imu_v"]), rel_err(st[15:18], fr["imu_bg"]),
                    rel_err(st[18:21], fr["imu_ba"]), rel_err(flt.cam_R[b, :nc].cpu().numpy().reshape(nc, 3, 3), fr["cam_R"]),
                    rel_err(flt.cam_t[b, :nc].cpu().numpy(), fr["cam_t"])]
            worst = max(worst, max(errs))
            assert worst < 1e-9, (i, errs)
    print("resident frame loop, worst relative error over", len(tr["frames"]), "frames:", worst)


def test_monte_carlo_sequences_resident():
    """Six independent synthetic sequences (two seeds per IMU-noise level) run for 30 camera frames with everything
    resident (monte_carlo.run_monte_carlo): the host's track records stay in step with the device through every
    deletion (asserted inside), the windows respect their maximum, and the trajectories follow the ground truth."""
    from monocular_visual_inertial_msckf_b200.monte_carlo import run_monte_carlo
    res = run_monte_carlo(list(range(6)), n_frames=30, n_landmarks=40, window=8)
    assert np.isfinite(res.positions).all()
    assert (res.n_cams <= 9).all() and (res.n_cams[:, -1] >= 7).all()
    assert (res.n_tracks[:, 5:] >= 20).all()
    travelled = np.linalg.norm(np.diff(res.gt_positions, axis=1), axis=2).sum(axis=1)
    print("ATE rmse [m]:", np.round(res.ate_rmse, 4), "path length [m]:", np.round(travelled, 2))
    assert (res.ate_rmse < 0.05 * np.maximum(travelled, 1.0)).all()


def test_prune_flow_vs_oracle(golden_dir):
    """prune_poorest_camera_states end to end on resident state (select -> update with the touched tracks -> drop ->
    remove) against the oracle's composition of the same reference steps."""
    cases, _ = _load(golden_dir, "mid_48x10_updates.npz")
    cases = [c for c in cases if int(c["n_cams"]) >= 4][:2]
    flt = _batch_from_cases(cases, _lib.COMPRESS_GRAM)
    mask, touch = flt.prune_poorest_camera_states()
    torch.cuda.synchronize()
    mask, touch = mask.cpu().numpy().astype(bool), touch.cpu().numpy().astype(bool)
    for b, c in enumerate(cases):
        n, f = int(c["n_cams"]), len(c["feat_len"])
        m_ref, t_ref = o.select_poorest_cameras(c)
        assert (mask[b, :n] == m_ref).all() and (touch[b, :f] == t_ref).all()
        sub = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in c.items()}
        sub["feat_len"] = np.where(t_ref, c["feat_len"], 0).astype(np.int32)
        ref = o.update_case(sub)
        P = ref["out_P"]
        for s in sorted(np.nonzero(m_ref)[0], reverse=True):
            P = o.remove_camera_covariance(P, int(s))
        keep = ~m_ref
        d = P.shape[0]
        assert int(flt.n_cams[b]) == int(keep.sum())
        assert rel_err(flt.P[b, :d, :d].cpu().numpy(), P) < 1e-9
        assert rel_err(flt.cam_t[b, :keep.sum()].cpu().numpy(), ref["out_cam_t"][keep]) < 1e-9
        after = dict(c)
        after["cam_t"] = ref["out_cam_t"]
        tr = o.drop_observations(after, m_ref)
        keep = np.nonzero(tr["feat_len"] > 0)[0]                 # tracks left empty are deleted, the others keep their order
        ln = flt.feat_len[b].cpu().numpy()
        assert (ln[:len(keep)] == tr["feat_len"][keep]).all() and (ln[len(keep):] == 0).all()
        sl = flt.obs_slot[b].cpu().numpy()
        for d, j in enumerate(keep):
            assert (sl[d, :tr["feat_len"][j]] == tr["obs_slot"][j, :tr["feat_len"][j]]).all()


@pytest.mark.parametrize("mode", MODES)
def test_full_size_properties(golden_dir, mode):
    """BASELINE shape (256 features x 30 poses) replicated over a batch with per-filter pixel noise: size-independent
    properties -- P+ symmetric, trace reduced, invariance to the order of the features, determinism."""
    cases, _ = _load(golden_dir, "baseline_256x30_updates.npz")
    base = cases[0]
    rng = np.random.default_rng(11)
    B = 24
    batch = []
    for b in range(B):
        c = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in base.items()}
        if b >= 2:
            c["obs_uv"] = c["obs_uv"] + 0.1 * rng.standard_normal(c["obs_uv"].shape)
        batch.append(c)
    # filter 1 = filter 0 with its features permuted
    perm = rng.permutation(len(base["feat_len"]))
    for k in ("feat_len", "obs_uv", "obs_dir", "obs_conf", "obs_slot", "idp_base", "idp_anchor_slot", "idp_m", "idp_rho",
              "lost_cnt", "tracked_cnt"):
        batch[1][k] = batch[1][k][perm]
    flt = _batch_from_cases(batch, mode)
    flt.update()
    torch.cuda.synchronize()
    P = flt.P.cpu().numpy()
    d = base["P"].shape[0]
    assert np.array_equal(P, np.transpose(P, (0, 2, 1)))                       # exactly symmetric after (P+P^T)/2
    for b in range(B):
        assert np.trace(P[b, :d, :d]) < np.trace(base["P"])                     # information was added
        assert np.linalg.eigvalsh(P[b, :d, :d]).min() > -1e-12 * np.abs(P[b]).max()
    compare_update_outputs(flt.case_outputs(0, base), base, CHI, label="baseline#0 in batch")
    assert rel_err(P[1], P[0]) < 1e-11                                          # feature order only changes rounding
    assert rel_err(flt.delta_x[1].cpu().numpy(), flt.delta_x[0].cpu().numpy()) < 1e-9
    assert (flt.feat_status[1].cpu().numpy() == flt.feat_status[0].cpu().numpy()[perm]).all()
    # determinism: a second run on identical inputs is bit-identical
    flt2 = _batch_from_cases(batch, mode)
    flt2.update()
    torch.cuda.synchronize()
    assert torch.equal(flt2.P, flt.P) and torch.equal(flt2.delta_x, flt.delta_x)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("F,M,N", [(24, 12, 36), (16, 40, 40), (40, 33, 34), (48, 60, 60)])
def test_large_windows_vs_oracle(mode, F, M, N):
    """Windows / tracks beyond the warp-per-feature kernels' budget (N > 32 or M > 32): the wide front end (CTA-per-feature
    geometry, P-stationary assembly over several CTAs per filter, shared-memory factorisation), multi-pass information
    accumulation (6N > 192), larger update factorisations -- against the oracle on synthetic
    cases (the golden fixtures stop at N = 31).  (48, 60, 60) is the per-feature / per-filter shape of BASELINE.json's
    stress config (60-pose window, state 375, 60 observations per track)."""
    from helpers import make_case
    if mode == _lib.COMPRESS_QR and N > _lib.load_library().msckf_max_cameras(_lib.COMPRESS_QR):
        pytest.skip("explicit-QR kernel keeps R in one SM's shared memory: N <= %d" % _lib.load_library().msckf_max_cameras(0))
    cases = [make_case(F, M, N, seed=s) for s in (1, 2)]
    refs = [o.update_case(c) for c in cases]
    flt = _batch_from_cases(cases, mode)
    flt.update()
    torch.cuda.synchronize()
    n_rej = 0
    for b, (c, ref) in enumerate(zip(cases, refs)):
        compare_update_outputs(flt.case_outputs(b, c), ref, CHI, label=f"synthetic F{F} M{M} N{N} #{b}")
        n_rej += ref["out_n_rejected"]
        assert ref["out_updated"]


def test_stress_shape_full_size_vs_oracle():
    """BASELINE.json configs[4] at its FULL per-filter size -- 1024 features x 60 observations, 60-pose window (state 375) --
    on sampled filters against the oracle: gate decisions, gamma, delta_x, covariance and poses (the reference's own
    update() would need a 115 GB matrix here; the oracle's only algebraic change is R_n = sigma^2 Q^T Q).  Ragged track
    lengths (33..60) ride along in the second filter: geometry by the CTA-per-feature kernel, P-stationary assembly over
    four CTAs per filter, shared-memory factorisation of up to 16 tile columns, multi-pass information accumulation."""
    from helpers import make_case
    cases = [make_case(1024, 60, 60, seed=7), make_case(1024, 60, 60, seed=8)]
    rng = np.random.default_rng(5)
    c = cases[1]
    for j in range(0, 1024, 3):
        c["feat_len"][j] = int(rng.integers(33, 61))
        c["tracked_cnt"][j] = c["feat_len"][j]
    refs = [o.update_case(c) for c in cases]
    flt = _batch_from_cases(cases, _lib.COMPRESS_GRAM)
    assert (flt.F, flt.M, flt.N) == (1024, 60, 60)
    flt.update()
    torch.cuda.synchronize()
    for b, (c, ref) in enumerate(zip(cases, refs)):
        errs = compare_update_outputs(flt.case_outputs(b, c), ref, CHI, label=f"stress 1024x60x60 #{b}")
        assert ref["out_updated"] and ref["out_n_rejected"] > 10
        print(f"stress shape filter {b}: {int(ref['out_gate'].sum())} accepted, {ref['out_n_rejected']} gated out, "
              f"P {errs['out_P']:.1e}, delta_x {errs['delta_x']:.1e}, gamma {errs.get('gamma', 0):.1e}")


def _ragged_reversed_gapped(F, M, N, seeds):
    """Synthetic cases with RAGGED track lengths 1..M (every tile count of the factorisation kernels), tracks stored
    newest-first (slots decreasing with the observation index: the transposed-block path of the assembly) and tracks with
    GAPS (a lost track that was matched again later skips window slots: MSCKF.py:400-434 resets lost_for_n_frames)."""
    from helpers import make_case
    cases = []
    for seed in seeds:
        c = make_case(F, M, N, seed=seed)
        rng = np.random.default_rng(100 + seed)
        for j in range(F):
            if j % 3 == 1:                                  # ragged: keep the first L observations
                c["feat_len"][j] = int(rng.integers(1, M + 1))
                c["tracked_cnt"][j] = c["feat_len"][j]
            L = int(c["feat_len"][j])
            if j % 5 == 3 and L >= 6:                       # gaps: drop two observations from the middle of the track
                for cut in sorted(rng.choice(np.arange(1, L - 1), size=2, replace=False))[::-1]:
                    for k in ("obs_uv", "obs_dir", "obs_conf", "obs_slot"):
                        c[k][j, cut:L - 1] = c[k][j, cut + 1:L].copy()
                    L -= 1
                c["feat_len"][j] = L
                c["tracked_cnt"][j] = L
            if j % 4 == 2:                                  # newest-first storage order
                for k in ("obs_uv", "obs_dir", "obs_conf", "obs_slot"):
                    c[k][j, :L] = c[k][j, :L][::-1].copy()
                c["idp_anchor_slot"][j] = c["obs_slot"][j, 0]
                c["idp_base"][j] = c["cam_t"][c["obs_slot"][j, 0]]
                c["idp_m"][j] = c["obs_dir"][j, 0] / np.linalg.norm(c["obs_dir"][j, 0])
        cases.append(c)
    return cases


@pytest.mark.parametrize("F,M,N", [(40, 32, 32), (24, 31, 31), (24, 29, 32), (48, 30, 30)])
def test_warp_front_end_capacity_edges(F, M, N):
    """The warp-per-feature front ends at the edges of their tile budget (M = 32: no padding rows in the last tile row;
    N = 32: the staged covariance block fills the shared memory), with ragged, reversed and gapped tracks in one batch --
    against the oracle."""
    cases = _ragged_reversed_gapped(F, M, N, (3, 4))
    refs = [o.update_case(c) for c in cases]
    flt = _batch_from_cases(cases, _lib.COMPRESS_GRAM)
    assert flt.M <= 32 and flt.N <= 32
    flt.update()
    torch.cuda.synchronize()
    for b, (c, ref) in enumerate(zip(cases, refs)):
        compare_update_outputs(flt.case_outputs(b, c), ref, CHI, label=f"edge F{F} M{M} N{N} #{b}")
        assert ref["out_updated"]


@pytest.mark.parametrize("F,M,N", [(40, 48, 50), (32, 64, 64), (36, 33, 61)])
def test_wide_front_end_ragged_reversed_gapped(F, M, N):
    """The same irregular tracks through the WIDE front end (tracks of 33..64 observations / windows of 33..64 poses:
    CTA-per-feature geometry, P-stationary assembly over several CTAs per filter, shared-memory factorisation of 1..17 tile
    columns), including its capacity corner (64, 64) -- against the oracle."""
    cases = _ragged_reversed_gapped(F, M, N, (5, 6))
    refs = [o.update_case(c) for c in cases]
    flt = _batch_from_cases(cases, _lib.COMPRESS_GRAM)
    flt.update()
    torch.cuda.synchronize()
    for b, (c, ref) in enumerate(zip(cases, refs)):
        compare_update_outputs(flt.case_outputs(b, c), ref, CHI, label=f"wide F{F} M{M} N{N} #{b}")
        assert ref["out_updated"]


@pytest.mark.parametrize("name", ["small_updates.npz", "baseline_256x30_updates.npz"])
def test_gate_bounds_mode_gives_identical_decisions_and_state(golden_dir, name):
    """MSCKF_GATE_BOUNDS: features whose upper bound passes / lower bound fails skip the factorisation; the decisions,
    the update and the covariance are the reference's; the reported gamma of such a feature is the deciding bound."""
    cases, _ = _load(golden_dir, name)
    F = max(len(c["feat_len"]) for c in cases)
    M = max(c["obs_uv"].shape[1] for c in cases)
    N = max(int(c["n_cams"]) for c in cases)
    par = UpdateParams.from_case(cases[0], _lib.COMPRESS_GRAM, _lib.GATE_BOUNDS)
    flt = BatchedMSCKF(len(cases), F, M, N, par)
    flt.load_cases(cases)
    flt.update()
    torch.cuda.synchronize()
    n_bound = n_full = 0
    for b, c in enumerate(cases):
        out = flt.case_outputs(b, c)
        st = flt.feat_status[b, :len(c["feat_len"])].cpu().numpy()
        flagged = (st & _lib.FEAT_GAMMA_BOUND) != 0
        v = np.asarray(c["out_valid"], bool)
        assert not flagged[~v].any()
        crit = CHI[c["out_dof"]]
        with np.errstate(invalid="ignore", divide="ignore"):
            decided = ~(np.abs(c["out_gamma"] / crit - 1.0) <= 1e-6)
        assert (out["out_gate"][v & decided] == c["out_gate"][v & decided]).all()
        acc_b, rej_b = flagged & out["out_gate"], flagged & ~out["out_gate"]
        assert (out["out_gamma"][acc_b] >= c["out_gamma"][acc_b] * (1 - 1e-12)).all()      # upper bound decided "accept"
        assert (out["out_gamma"][rej_b] <= c["out_gamma"][rej_b] * (1 + 1e-12)).all()      # lower bound decided "reject"
        exact = v & ~flagged
        if exact.any():
            g = c["out_gamma"][exact]
            assert np.max(np.abs(out["out_gamma"][exact] - g) / np.maximum(np.abs(g), 1e-3)) < 1e-6
        n_bound += int(flagged.sum())
        n_full += int(exact.sum())
        if (out["out_gate"][v] == c["out_gate"][v]).all():
            assert out["out_n_rejected"] == c["out_n_rejected"]
            for k in ("out_P", "out_imu_t", "out_imu_v", "out_cam_t", "out_cam_R"):
                assert rel_err(out[k], c[k]) < 1e-9, (k, rel_err(out[k], c[k]))
    print(name, "decided by bounds:", n_bound, "full test:", n_full)
    assert n_bound > 0


@pytest.mark.parametrize("chunk", [2, 3, 0])
def test_step_host_pipeline_equals_resident_calls(golden_dir, chunk):
    """msckf_filter_step_host (host buffers in, chunked upload / kernels / download, results out) gives bit-identical
    results to propagate() + update() on resident inputs, for ragged chunking, and matches the golden update."""
    cases, _ = _load(golden_dir, "mid_48x10_updates.npz")
    cases = (cases * 3)[:7]
    B = len(cases)
    mode = _lib.COMPRESS_GRAM
    ref = _batch_from_cases(cases, mode)
    flt = _batch_from_cases(cases, mode)
    n = 4
    rng = np.random.default_rng(5)
    samples = torch.zeros(B, n, 7, dtype=torch.float64)
    samples[:, :, 0] = torch.arange(1, n + 1, dtype=torch.float64) * 0.005
    samples[:, :, 1:4] = torch.from_numpy(0.02 * rng.standard_normal((B, n, 3)))
    samples[:, :, 4:7] = torch.from_numpy(np.array([0.0, 0.0, 9.81]) + 0.1 * rng.standard_normal((B, n, 3)))
    for f in (ref, flt):
        f.Qc_diag[:] = torch.tensor([1e-8, 1e-10, 1e-6, 1e-8], dtype=torch.float64)
    # resident path
    ref.propagate(samples.cuda())
    ref.update()
    # host path: every track array comes from (pinned) host memory; the device copies are scrambled first
    host_in = {k: getattr(flt, k).cpu().pin_memory() for k in _lib.HOST_INPUTS}
    for k in _lib.HOST_INPUTS:
        getattr(flt, k).fill_(0)
    host_out = dict(imu_state=torch.empty(B, 21, dtype=torch.float64).pin_memory(),
                    cam_R=torch.empty(B, flt.N, 9, dtype=torch.float64).pin_memory(),
                    cam_t=torch.empty(B, flt.N, 3, dtype=torch.float64).pin_memory(),
                    idp_m_out=torch.empty(B, flt.F, 3, dtype=torch.float64).pin_memory(),
                    idp_rho_out=torch.empty(B, flt.F, dtype=torch.float64).pin_memory(),
                    feat_status=torch.empty(B, flt.F, dtype=torch.uint8).pin_memory(),
                    gamma=torch.empty(B, flt.F, dtype=torch.float64).pin_memory(),
                    delta_x=torch.empty(B, flt.ldp, dtype=torch.float64).pin_memory(),
                    n_accepted=torch.empty(B, dtype=torch.int32).pin_memory(),
                    n_rejected=torch.empty(B, dtype=torch.int32).pin_memory())
    for rep in range(2):      # second call exercises the cross-call event ordering (overlap_previous)
        if rep == 1:
            fresh = _batch_from_cases(cases, mode)
            for k in ("P", "imu_state", "cam_R", "cam_t", "imu_time", "imu_null"):
                getattr(flt, k).copy_(getattr(fresh, k))
        flt.step_host(host_in, samples.pin_memory(), host_out, chunk_filters=chunk, overlap_previous=(rep == 1))
        torch.cuda.synchronize()
        for k, dev_name in (("imu_state", "imu_state"), ("cam_R", "cam_R"), ("cam_t", "cam_t"), ("idp_m_out", "idp_m"),
                            ("idp_rho_out", "idp_rho"), ("feat_status", "feat_status"), ("delta_x", "delta_x"),
                            ("n_accepted", "n_accepted"), ("n_rejected", "n_rejected")):
            assert torch.equal(host_out[k], getattr(ref, dev_name).cpu()), (k, rep)
        g_host, g_ref = host_out["gamma"].numpy(), ref.gamma.cpu().numpy()
        assert np.array_equal(np.isnan(g_host), np.isnan(g_ref)) and np.array_equal(g_host[~np.isnan(g_host)], g_ref[~np.isnan(g_ref)])
        assert torch.equal(flt.P, ref.P)
    with pytest.raises(KeyError):
        flt.step_host(dict(P=torch.zeros(1)))
    with pytest.raises(ValueError):
        flt.step_host(dict(obs_uv=torch.zeros(1, dtype=torch.float64)))
