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
