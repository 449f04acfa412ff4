"""Shared comparison helpers for the parity tests."""
import numpy as np


def rel_err(a, b):
    """max |a - b| / max |b|  (the "relative" of BASELINE.json's 1e-9 bar: relative to the quantity's scale)."""
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


# "away from the threshold" band for gating decisions: |gamma / chi2 - 1| > 1e-6
GATE_BAND = 1e-6


def compare_update_outputs(out, ref, chi2_table, tol_state=1e-9, tol_gamma=1e-6, tol_dx=1e-9, label=""):
    """Checks one filter's update outputs against golden/oracle outputs; returns a dict of the errors seen."""
    errs = {}
    assert (out["out_valid"] == ref["out_valid"]).all(), f"{label}: valid mask differs"
    assert (out["out_lost"] == ref["out_lost"]).all(), f"{label}: lost mask differs"
    v = np.asarray(ref["out_valid"], bool)
    assert (out["out_dof"][v] == ref["out_dof"][v]).all(), f"{label}: dof differs"
    errs["idp_m"] = rel_err(out["out_idp_m"], ref["out_idp_m"])
    errs["idp_rho"] = rel_err(out["out_idp_rho"], ref["out_idp_rho"])
    assert errs["idp_m"] < tol_state and errs["idp_rho"] < tol_state, f"{label}: triangulation {errs}"
    if v.any():
        g_ref = ref["out_gamma"][v]
        g_out = out["out_gamma"][v]
        crit = chi2_table[ref["out_dof"][v]]
        with np.errstate(invalid="ignore", divide="ignore"):
            margin = np.abs(g_ref / crit - 1.0)
        decided = ~(margin <= GATE_BAND)          # NaN critical value (dof 0) counts as decided: always rejected
        assert (out["out_gate"][v][decided] == ref["out_gate"][v][decided]).all(), f"{label}: gate decision differs"
        errs["gamma"] = float(np.max(np.abs(g_out - g_ref) / np.maximum(np.abs(g_ref), 1e-3)))
        assert errs["gamma"] < tol_gamma, f"{label}: gamma {errs['gamma']}"
        errs["near_threshold"] = int((~decided).sum())
        same_gates = (out["out_gate"][v] == ref["out_gate"][v]).all()
    else:
        same_gates = True
    if same_gates:
        assert out["out_n_rejected"] == ref["out_n_rejected"], f"{label}: n_rejected"
        assert bool(out["out_updated"]) == bool(ref["out_updated"]), f"{label}: updated flag"
        errs["delta_x"] = rel_err(out["out_delta_x"], ref["out_delta_x"]) if ref["out_updated"] else 0.0
        for k in ("out_P", "out_cam_R", "out_cam_t", "out_imu_R", "out_imu_t", "out_imu_v", "out_imu_bg", "out_imu_ba"):
            errs[k] = rel_err(out[k], ref[k])
            assert errs[k] < tol_state, f"{label}: {k} rel err {errs[k]}"
        assert errs["delta_x"] < tol_dx, f"{label}: delta_x rel err {errs['delta_x']}"
    return errs


def make_case(F, M, N, seed=0, outlier_frac=0.05):
    """A self-consistent packed update case of arbitrary shape (F tracks x M observations, N-pose window) for sizes the
    golden fixtures do not hold: poses on a smooth arc, landmarks 6-14 m ahead, every track observed in M consecutive
    window slots, a random SPD covariance with realistic scales.  The oracle supplies the expected outputs."""
    rng = np.random.default_rng(seed)
    K = np.array([[320.0, 0, 320], [0, 320.0, 240], [0, 0, 1]])
    Kinv = np.linalg.inv(K)
    Rbc = np.array([[0.0, 0, 1], [-1, 0, 0], [0, -1, 0]])
    cam_R = np.zeros((N, 3, 3))
    cam_t = np.zeros((N, 3))
    for s in range(N):
        a = 0.01 * s
        Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
        cam_R[s] = Rz @ Rbc
        cam_t[s] = [0.05 * s, 0.6 * np.sin(0.15 * s), 0.1 * np.cos(0.1 * s)]
    d = 15 + 6 * N
    Q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    P = (Q * 10.0 ** rng.uniform(-7, -4, d)) @ Q.T
    P = (P + P.T) / 2
    c = dict(K=K, Kinv=Kinv, width=640, height=480, sigma=0.1, gravity=np.array([0.0, 0.0, -9.81]), min_lost=2,
             min_tracked=3, use_parallax=1, min_parallax=1e9, n_cams=N, cam_R=cam_R, cam_t=cam_t,
             imu_R=cam_R[-1] @ Rbc.T, imu_t=cam_t[-1].copy(), imu_v=np.array([0.3, 0.1, 0.0]),
             imu_bg=1e-3 * rng.standard_normal(3), imu_ba=1e-2 * rng.standard_normal(3), P=P,
             feat_ids=np.arange(F, dtype=np.int64), feat_len=np.full(F, M, np.int32), obs_uv=np.zeros((F, M, 2)),
             obs_dir=np.zeros((F, M, 3)), obs_conf=rng.uniform(0.7, 1.0, (F, M)), obs_slot=np.zeros((F, M), np.int32),
             idp_base=np.zeros((F, 3)), idp_anchor_slot=np.zeros(F, np.int32), idp_m=np.zeros((F, 3)),
             idp_rho=np.full(F, 0.1), lost_cnt=np.full(F, 5, np.int32), tracked_cnt=np.full(F, M, np.int32), call_index=0)
    for j in range(F):
        depth = rng.uniform(6, 14)
        p = np.array([depth, depth * np.tan(rng.uniform(-0.3, 0.3)), depth * np.tan(rng.uniform(-0.25, 0.25))])
        s0 = rng.integers(0, N - M + 1)
        bad = rng.random() < outlier_frac
        for i in range(M):
            s = s0 + i
            pc = cam_R[s].T @ (p - cam_t[s])
            uv = (K @ pc)[:2] / pc[2] + 0.1 * rng.standard_normal(2)
            if bad and i % 3 == 0:
                uv += 80.0 * rng.standard_normal(2)
            c["obs_uv"][j, i] = uv
            c["obs_dir"][j, i] = cam_R[s] @ (Kinv @ np.array([uv[0], uv[1], 1.0]))
            c["obs_slot"][j, i] = s
        c["idp_anchor_slot"][j] = s0
        c["idp_base"][j] = cam_t[s0]
        c["idp_m"][j] = c["obs_dir"][j, 0] / np.linalg.norm(c["obs_dir"][j, 0])
    return c
