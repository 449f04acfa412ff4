"""NumPy prototype of the GPU formulation (projector-form gating, H~ = blockdiag(A) - Q1 W stacking, Householder-TSQR
vs Gram/Cholesky compression) used to check its numerics against the golden fixtures before writing the kernels.
Developer tool; not imported by the product."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import msckf_oracle as o


def feature_blocks(case, j, m, rho):
    r, Hx, Hf = o.residual_and_jacobians(case, j, m, rho)
    M = int(case["feat_len"][j])
    A = np.zeros((2 * M, 6))
    for i in range(M):
        s = int(case["obs_slot"][j, i])
        A[2 * i:2 * i + 2] = Hx[2 * i:2 * i + 2, 15 + 6 * s:21 + 6 * s]
    return r.flatten(), A, Hf, case["obs_slot"][j, :M].astype(int)


def range_basis(Hf):
    """column-pivoted Householder-like orthonormal basis of range(Hf) with SciPy's rank rule."""
    n2 = Hf.shape[0]
    sv = np.linalg.svd(Hf, compute_uv=False)
    tol = np.finfo(float).eps * max(3, n2) * sv[0]
    rank = int((sv > tol).sum())
    import scipy.linalg as sl
    Q, R, piv = sl.qr(Hf, mode="economic", pivoting=True)
    return Q[:, :rank], rank


def semidef_chol_upper(G, tol=1e-14):
    n = G.shape[0]
    R = np.zeros((n, n))
    Gw = G.copy()
    d0 = np.diag(G).copy()
    for k in range(n):
        p = Gw[k, k]
        if p <= tol * max(d0[k], 1e-300):
            continue
        R[k, k] = np.sqrt(p)
        R[k, k + 1:] = Gw[k, k + 1:] / R[k, k]
        Gw[k + 1:, k + 1:] -= np.outer(R[k, k + 1:], R[k, k + 1:])
    return R


def semidef_chol_lower_panel(Pp, c):
    """first c columns of the Cholesky factor of the PSD matrix Pp (n x n), zero column on a non-positive pivot."""
    n = Pp.shape[0]
    L = np.zeros((n, c))
    W = Pp.copy()
    for k in range(c):
        p = W[k, k]
        if p <= 0:
            continue
        L[k:, k] = W[k:, k] / np.sqrt(p)
        W[k + 1:, k + 1:] -= np.outer(L[k + 1:, k], L[k + 1:, k])
    return L


def info_update(P, G, g, sigma, mode):
    d = P.shape[0]; c = d - 15
    s2 = sigma ** 2
    if mode == "white":
        # order [C; I]: U~ = P[:, C] L^-T
        perm = np.r_[15:d, 0:15]
        Pp = P[np.ix_(perm, perm)]
        Lp = semidef_chol_lower_panel(Pp, c)          # (d x c): rows 0:c = L, rows c: = U~_I
        L = Lp[:c]; UI = Lp[c:]
        Gt = L.T @ G @ L
        Gt = (Gt + Gt.T) / 2
        B = s2 * np.eye(c) + Gt
        CB = np.linalg.cholesky(B)
        Wp = np.linalg.solve(CB, Lp.T).T               # U~ C_B^-T  (d x c) in permuted row order
        Pn_p = s2 * Wp @ Wp.T
        Pn_p[c:, c:] += Pp[c:, c:] - UI @ UI.T
        inv = np.argsort(perm)
        Pn = Pn_p[np.ix_(inv, inv)]
        y = np.linalg.solve(CB, L.T @ g)
        dx = (Wp @ y)[inv]
        return (Pn + Pn.T) / 2, dx
    Pcc = P[15:, 15:]; Pc = P[:, 15:]
    A = s2 * np.eye(c) + G @ Pcc
    import scipy.linalg as sl
    lu = sl.lu_factor(A)
    X = sl.lu_solve(lu, G)
    mg = sl.lu_solve(lu, g)
    dx = Pc @ mg
    if mode == "lu_direct":
        Z = (X + X.T) / 2
    else:
        Mi = sl.lu_solve(lu, np.eye(c))
        E = X @ Pcc + s2 * Mi
        Z = X + X.T - X @ E.T
        Z = (Z + Z.T) / 2
    Pn = P - Pc @ Z @ Pc.T
    return (Pn + Pn.T) / 2, dx


def run(case, mode="gram"):
    valid, lost, m_new, rho_new = o.select_and_triangulate(case)
    P = np.array(case["P"], float)
    d = P.shape[0]
    c = d - 15
    Pcc = P[15:, 15:]
    sigma = float(case["sigma"])
    F = len(valid)
    chi = o.chi2_table(2 * case["obs_uv"].shape[1] + 3)
    gamma = np.full(F, np.nan); dof = np.zeros(F, int); gate = np.zeros(F, bool)
    G = np.zeros((c, c)); g = np.zeros(c)
    rows = []; rhs = []
    for j in range(F):
        if not valid[j]:
            continue
        r, A, Hf, slots = feature_blocks(case, j, m_new[j], rho_new[j])
        M = len(slots)
        Q1, rank = range_basis(Hf)
        dof[j] = 2 * M - rank
        # S_x from 6x6 blocks of Pcc
        Sx = np.zeros((2 * M, 2 * M))
        for i in range(M):
            for k in range(M):
                Sx[2*i:2*i+2, 2*k:2*k+2] = A[2*i:2*i+2] @ Pcc[6*slots[i]:6*slots[i]+6, 6*slots[k]:6*slots[k]+6] @ A[2*k:2*k+2].T
        Y = Sx @ Q1
        Z = Q1.T @ Y
        Sp = Sx - Q1 @ Y.T - Y @ Q1.T + Q1 @ Z @ Q1.T + sigma**2 * np.eye(2 * M)
        rt = r - Q1 @ (Q1.T @ r)
        if dof[j] > 0:
            L = np.linalg.cholesky(Sp)
            y = np.linalg.solve(L, rt)
            gamma[j] = y @ y
        else:
            gamma[j] = 0.0
        gate[j] = gamma[j] <= chi[dof[j]]
        if not gate[j]:
            continue
        W = np.zeros((Q1.shape[1], c))
        Hs = np.zeros((2 * M, c))
        for i in range(M):
            W[:, 6*slots[i]:6*slots[i]+6] = Q1[2*i:2*i+2].T @ A[2*i:2*i+2]
            Hs[2*i:2*i+2, 6*slots[i]:6*slots[i]+6] = A[2*i:2*i+2]
            G[6*slots[i]:6*slots[i]+6, 6*slots[i]:6*slots[i]+6] += A[2*i:2*i+2].T @ A[2*i:2*i+2]
            g[6*slots[i]:6*slots[i]+6] += A[2*i:2*i+2].T @ r[2*i:2*i+2]
        w = Q1.T @ r
        G -= W.T @ W
        g -= W.T @ w
        rows.append(Hs - Q1 @ W); rhs.append(rt)
    out = dict(out_valid=valid, out_lost=lost, out_idp_m=m_new, out_idp_rho=rho_new, out_gamma=gamma, out_dof=dof, out_gate=gate)
    if not gate.any():
        return out
    if mode in ("white", "lu", "lu_direct"):
        out["out_P"], out["out_delta_x"] = info_update(P, G, g, sigma, mode)
        return out
    if mode == "qr":
        Hs = np.vstack(rows); rs = np.concatenate(rhs)
        if Hs.shape[0] < c:
            Hs = np.vstack([Hs, np.zeros((c - Hs.shape[0], c))]); rs = np.concatenate([rs, np.zeros(c - len(rs))])
        Q, R = np.linalg.qr(Hs)
        rn = Q.T @ rs
    else:
        R = semidef_chol_upper(G)
        rn = np.zeros(c)
        for k in range(c):
            if R[k, k] > 0:
                rn[k] = (g[k] - R[:k, k] @ rn[:k]) / R[k, k]
    Pc = P[:, 15:]
    S = R @ Pcc @ R.T + sigma**2 * np.eye(c)
    L = np.linalg.cholesky(S)
    PHt = Pc @ R.T
    Kg = np.linalg.solve(L.T, np.linalg.solve(L, PHt.T)).T
    dx = Kg @ rn
    KT = Kg @ R
    IKH = np.eye(d); IKH[:, 15:] -= KT
    Pn = IKH @ P @ IKH.T + sigma**2 * Kg @ Kg.T
    out["out_P"] = (Pn + Pn.T) / 2
    out["out_delta_x"] = dx
    return out


if __name__ == "__main__":
    cases, extra = o.load_cases(sys.argv[1])
    mode = sys.argv[2] if len(sys.argv) > 2 else "gram"
    worst = {}
    nmis = 0; ngate = 0
    for c in cases:
        out = run(c, mode)
        assert (out["out_valid"] == c["out_valid"]).all()
        v = c["out_valid"]
        assert (out["out_dof"][v] == c["out_dof"][v]).all(), (out["out_dof"][v], c["out_dof"][v])
        mis = out["out_gate"][v] != c["out_gate"][v]
        chi = o.chi2_table(200)
        ngate += v.sum()
        if mis.any():
            idx = np.nonzero(v)[0][mis]
            print("gate mismatch", c["call_index"], c["out_gamma"][idx] / chi[c["out_dof"][idx]], out["out_gamma"][idx] / chi[c["out_dof"][idx]])
            nmis += mis.sum()
            continue
        eg = np.nanmax(np.abs(out["out_gamma"][v] - c["out_gamma"][v]) / np.maximum(np.abs(c["out_gamma"][v]), 1e-9)) if v.any() else 0
        worst["gamma"] = max(worst.get("gamma", 0), eg)
        if c["out_updated"]:
            for k in ("out_delta_x", "out_P"):
                e = np.max(np.abs(out[k] - c[k])) / np.max(np.abs(c[k]))
                worst[k] = max(worst.get(k, 0), e)
            # state-relative error of dx
            e = np.max(np.abs(out["out_delta_x"] - c["out_delta_x"])[12:15]) / max(np.max(np.abs(c["imu_t"])), 1e-3)
            worst["dx_vs_state"] = max(worst.get("dx_vs_state", 0), e)
    print(mode, len(cases), "cases; gate calls", ngate, "mismatch", nmis, worst)
