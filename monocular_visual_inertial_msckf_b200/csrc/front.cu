// Per-feature front end of the MSCKF update (rows 3-8 of SURVEY.md section 8a), one CTA per (filter, feature):
//   selection + closed-form line-intersection triangulation   reference src/msckf/MSCKF.py:458-495,
//                                                              src/utils/geometry.py:274-303
//   residual + observability-projected Jacobian blocks         MSCKF.py:497-548, src/msckf/Camera.py:54-67
//   left-nullspace projection of H_f                           MSCKF.py:554-559 (scipy null_space)
//   chi-square Mahalanobis gate                                MSCKF.py:561-568
//
// Formulation (B200-first, not the reference's dense one).  With Q1 an orthonormal basis of range(H_f) and
// PI = I - Q1 Q1^T, every downstream quantity of the reference only depends on the null-space basis A through A A^T = PI:
//   gamma = (PI r)^T (PI S_x PI + sigma^2 I)^-1 (PI r),   S_x = H_x P H_x^T assembled from 2x6 . 6x6 . 6x2 blocks of P,
//   dof   = 2M - rank(H_f)   (rank by SciPy's rule s > eps max(3, 2M) s_max on the singular values),
// so no explicit null-space basis and no dense 2M x d Jacobian is ever formed.  The Cholesky is done as a square-root-free
// LDL^T on the packed lower triangle in shared memory with the projected residual appended as an extra row, which yields
// gamma without a separate triangular solve.  The factors (A_i, Q1, PI r) go to HBM for the compression stage.
#include "msckf_common.cuh"
#include "front_geom.cuh"
#include "front_tiles.cuh"

namespace {

constexpr int FRONT_THREADS = 128;

// what the geometry-only mode hands to the assembly / factorisation kernels of the split front end (long tracks)
struct GeoOut {
    double2* Kbuf;            // tile buffer of the sub-batch, kt_d2 double2 per feature slot
    int kt_d2;
    unsigned char* slot_map;  // [sub][F][64]
    int* pend_list;
    int* pend_count;
    int b0;                   // first filter of the sub-batch
};

// cyclic Jacobi eigen-decomposition of a symmetric 3x3 (a = [xx, xy, xz, yy, yz, zz]); V columns = eigenvectors
__device__ void sym3_eig(const double* a, double* lam, double* V) {
    double A[3][3] = {{a[0], a[1], a[2]}, {a[1], a[3], a[4]}, {a[2], a[4], a[5]}};
    double Q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        const double dia = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-300 || off <= 1e-18 * dia) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double qkp = Q[k][p], qkq = Q[k][q];
                    Q[k][p] = c * qkp - s * qkq;
                    Q[k][q] = s * qkp + c * qkq;
                }
            }
    }
    for (int k = 0; k < 3; ++k) {
        lam[k] = A[k][k];
        for (int r = 0; r < 3; ++r) V[r * 3 + k] = Q[r][k];
    }
}

// singular values of a 3x3 matrix G (row-major, destroyed) by one-sided Jacobi
__device__ void svals3(double* G, double* sv) {
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double al = 0, be = 0, ga = 0;
                for (int k = 0; k < 3; ++k) {
                    al += G[k * 3 + p] * G[k * 3 + p];
                    be += G[k * 3 + q] * G[k * 3 + q];
                    ga += G[k * 3 + p] * G[k * 3 + q];
                }
                if (ga == 0.0 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
                rotated = true;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < 3; ++k) {
                    const double gp = G[k * 3 + p], gq = G[k * 3 + q];
                    G[k * 3 + p] = c * gp - s * gq;
                    G[k * 3 + q] = s * gp + c * gq;
                }
            }
        if (!rotated) break;
    }
    for (int c = 0; c < 3; ++c) sv[c] = sqrt(G[c] * G[c] + G[3 + c] * G[3 + c] + G[6 + c] * G[6 + c]);
}

// GEO_ONLY: stop after the factors -- write the appended / padding rows of the padded K, the slot map and the pending list
// (what front_geo_kernel does for short tracks) and leave K's measurement blocks and its factorisation to the P-stationary
// assembly kernel and front_chol_smem_kernel.
template <bool GEO_ONLY>
__global__ void __launch_bounds__(FRONT_THREADS) front_kernel(const FrontArgs a, const GeoOut go) {
    const int j = blockIdx.x, b = blockIdx.y + (GEO_ONLY ? go.b0 : 0);
    const int tid = threadIdx.x, nth = blockDim.x;
    const int F = a.dims.F, Mcap = a.dims.M, Ncap = a.dims.N, ldp = a.dims.ldp;
    const size_t fj = (size_t)b * F + j;
    const int M = a.feat_len[fj];
    const int n2 = 2 * M;
    const int n2cap = 2 * Mcap;

    extern __shared__ double smem[];
    double* sA = smem;                      // [n2cap][6]
    double* sQ = sA + n2cap * 6;            // [n2cap][3]   H_f, then Q1
    double* sr = sQ + n2cap * 3;            // [n2cap]      raw residual
    double* srt = sr + n2cap;               // [n2cap]      projected residual
    double* sU = srt + n2cap;               // [n2cap][3]
    double* sS = sU + n2cap * 3;            // packed lower (n2cap+1)(n2cap+2)/2, aliased by the triangulation terms
    const int nS = GEO_ONLY ? Mcap * 9 : max(((n2cap + 1) * (n2cap + 2)) / 2, Mcap * 9);
    double* sm = sS + nS;                   // [64] small scratch
    int* sslot = (int*)(sm + 64);           // [Mcap]
    unsigned char* s_trow = (unsigned char*)(sslot + Mcap);   // [tri(n2cap + 1)] packed index -> row lookup
    __shared__ int s_flag[4];               // 0: selection code, 1: rank, 2: MSCKF_FEAT_TRIANGULATED or 0
    __shared__ double s_red[FRONT_THREADS / 32];

    double* facA = a.fac_A + fj * (size_t)Mcap * 12;
    double* facQ = a.fac_Q + fj * (size_t)Mcap * 6;
    double* facR = a.fac_rt + fj * (size_t)Mcap * 2;

    if (M <= 0 || M > Mcap) {
        if (tid == 0) {
            a.feat_status[fj] = MSCKF_FEAT_SKIPPED;
            a.gamma[fj] = nan("");
            a.dof[fj] = 0;
            a.fac_rank[fj] = 0;
        }
        return;
    }

    const double* uv = a.obs_uv + fj * (size_t)Mcap * 2;
    const double* dir = a.obs_dir + fj * (size_t)Mcap * 3;
    const double* conf = a.obs_conf + fj * (size_t)Mcap;
    const int* slot = a.obs_slot + fj * (size_t)Mcap;
    const double* camR = a.cam_R + (size_t)b * Ncap * 9;
    const double* camt = a.cam_t + (size_t)b * Ncap * 3;
    const double* P = a.P + (size_t)b * ldp * ldp;

    // ---- selection (MSCKF.py:463-479) -------------------------------------------------------------------------------
    if (tid == 0) {
        const bool lost = a.lost_cnt[fj] >= a.par.min_lost;
        int code;
        if (lost && a.tracked_cnt[fj] < a.par.min_tracked) {
            code = MSCKF_FEAT_LOST_DISCARD | MSCKF_FEAT_LOST_FLAG;
        } else {
            bool enough = false;
            if (a.par.use_parallax && M > 1) {
                const double* d1 = dir;
                const double* d2 = dir + 3 * (M - 1);
                const double n1 = sqrt(d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2]);
                const double n2_ = sqrt(d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2]);
                double dp = (d1[0] / n1) * (d2[0] / n2_) + (d1[1] / n1) * (d2[1] / n2_) + (d1[2] / n1) * (d2[2] / n2_);
                dp = fmin(1.0, fmax(-1.0, dp));
                const double parallax = acos(dp) * (180.0 / 3.14159265358979323846);
                enough = parallax > a.par.min_parallax_deg;
            }
            code = (lost || enough) ? (MSCKF_FEAT_ACCEPTED | (lost ? MSCKF_FEAT_LOST_FLAG : 0)) : MSCKF_FEAT_SKIPPED;
        }
        s_flag[0] = code;
        s_flag[2] = 0;
    }
    for (int i = tid; i < M; i += nth) sslot[i] = slot[i];
    if (!GEO_ONLY)      // (the packed-index lookup serves the assembly of S below)
        for (int e = tid; e < tri(n2 + 1); e += nth) s_trow[e] = (unsigned char)tri_row(e);
    __syncthreads();
    const int code = s_flag[0];
    if ((code & 0xF) != MSCKF_FEAT_ACCEPTED) {
        if (tid == 0) {
            a.feat_status[fj] = (unsigned char)code;
            a.gamma[fj] = nan("");
            a.dof[fj] = 0;
            a.fac_rank[fj] = 0;
        }
        return;
    }

    // ---- triangulation: X = sum c (I - d d^T), y = sum c (I - d d^T) t_slot  (geometry.py:286-299) ------------------
    double* sT = sS;   // [M][9]: xx xy xz yy yz zz y0 y1 y2
    for (int i = tid; i < M; i += nth) {
        const double dx = dir[3 * i], dy = dir[3 * i + 1], dz = dir[3 * i + 2];
        const double nn = sqrt(dx * dx + dy * dy + dz * dz);
        const double ux = dx / nn, uy = dy / nn, uz = dz / nn;
        const double c = conf[i];
        const double* bt = camt + 3 * sslot[i];
        const double pxx = c * (1.0 - ux * ux), pxy = c * (0.0 - ux * uy), pxz = c * (0.0 - ux * uz);
        const double pyy = c * (1.0 - uy * uy), pyz = c * (0.0 - uy * uz), pzz = c * (1.0 - uz * uz);
        double* t = sT + 9 * i;
        t[0] = pxx; t[1] = pxy; t[2] = pxz; t[3] = pyy; t[4] = pyz; t[5] = pzz;
        t[6] = pxx * bt[0] + pxy * bt[1] + pxz * bt[2];
        t[7] = pxy * bt[0] + pyy * bt[1] + pyz * bt[2];
        t[8] = pxz * bt[0] + pyz * bt[1] + pzz * bt[2];
    }
    __syncthreads();
    if (tid < 9) {   // sequential accumulation in observation order, like the reference's loop
        double acc = 0.0;
        for (int i = 0; i < M; ++i) acc += sT[9 * i + tid];
        sm[tid] = acc;
    }
    __syncthreads();
    if (tid == 0) {
        double p[3] = {0, 0, 0};
        // comfortably non-singular X: LDL^T solve; otherwise pinv with its singular-value cut-off (as the warp kernels do)
        const bool solved = fgeom::sym3_solve_spd(sm, sm + 6, p);
        double lam[3] = {0, 0, 0}, V[9];
        if (!solved) sym3_eig(sm, lam, V);
        const double lmax = fmax(fabs(lam[0]), fmax(fabs(lam[1]), fabs(lam[2])));
        for (int k = 0; k < 3 && !solved; ++k) {
            if (fabs(lam[k]) > 1e-15 * lmax) {   // np.linalg.pinv default rcond
                const double coef = (V[0 + k] * sm[6] + V[3 + k] * sm[7] + V[6 + k] * sm[8]) / lam[k];
                p[0] += coef * V[0 + k];
                p[1] += coef * V[3 + k];
                p[2] += coef * V[6 + k];
            }
        }
        // first observing camera (MSCKF.py:481-488)
        const double* R0 = camR + 9 * sslot[0];
        const double* t0 = camt + 3 * sslot[0];
        const double q0 = p[0] - t0[0], q1 = p[1] - t0[1], q2 = p[2] - t0[2];
        const double cx = R0[0] * q0 + R0[3] * q1 + R0[6] * q2;
        const double cy = R0[1] * q0 + R0[4] * q1 + R0[7] * q2;
        const double cz = R0[2] * q0 + R0[5] * q1 + R0[8] * q2;
        double rho = a.idp_rho[fj];
        double m0 = a.idp_m[fj * 3], m1 = a.idp_m[fj * 3 + 1], m2 = a.idp_m[fj * 3 + 2];
        if (cz > 0.0) {
            const double* K = a.par.K;
            const double iw = K[6] * cx + K[7] * cy + K[8] * cz;
            const double iu = (K[0] * cx + K[1] * cy + K[2] * cz) / iw;
            const double iv = (K[3] * cx + K[4] * cy + K[5] * cz) / iw;
            if (!(iu < 0.0 || iu >= (double)a.par.width || iv < 0.0 || iv >= (double)a.par.height)) {
                const double* Ki = a.par.Kinv;
                const double vx = Ki[0] * iu + Ki[1] * iv + Ki[2];
                const double vy = Ki[3] * iu + Ki[4] * iv + Ki[5];
                const double vz = Ki[6] * iu + Ki[7] * iv + Ki[8];
                const double wx = R0[0] * vx + R0[1] * vy + R0[2] * vz;
                const double wy = R0[3] * vx + R0[4] * vy + R0[5] * vz;
                const double wz = R0[6] * vx + R0[7] * vy + R0[8] * vz;
                const double wn = sqrt(wx * wx + wy * wy + wz * wz);
                rho = 1.0 / cz;                      // InverseDepthPoint.update, geometry.py:61-71
                m0 = wx / wn; m1 = wy / wn; m2 = wz / wn;
                a.idp_rho[fj] = rho;
                a.idp_m[fj * 3] = m0; a.idp_m[fj * 3 + 1] = m1; a.idp_m[fj * 3 + 2] = m2;
                s_flag[2] = MSCKF_FEAT_TRIANGULATED;
            }
        }
        const int as = a.idp_anchor_slot[fj];
        const double* base = (as >= 0) ? (camt + 3 * as) : (a.idp_base + fj * 3);
        sm[16] = rho; sm[17] = m0; sm[18] = m1; sm[19] = m2;
        sm[20] = base[0]; sm[21] = base[1]; sm[22] = base[2];
    }
    __syncthreads();

    // ---- residual and Jacobian blocks per observation (MSCKF.py:505-544, Camera.py:54-67) ---------------------------
    {
        const double rho = sm[16], m0 = sm[17], m1 = sm[18], m2 = sm[19];
        const double b0 = sm[20], b1 = sm[21], b2 = sm[22];
        const double g0 = a.par.gravity[0], g1 = a.par.gravity[1], g2 = a.par.gravity[2];
        const double* Ki = a.par.Kinv;
        for (int i = tid; i < M; i += nth) {
            const double* R = camR + 9 * sslot[i];
            const double* t = camt + 3 * sslot[i];
            const double q0 = rho * (b0 - t[0]) + m0, q1 = rho * (b1 - t[1]) + m1, q2 = rho * (b2 - t[2]) + m2;
            const double f0 = R[0] * q0 + R[3] * q1 + R[6] * q2;
            const double f1 = R[1] * q0 + R[4] * q1 + R[7] * q2;
            const double f2 = R[2] * q0 + R[5] * q1 + R[8] * q2;
            // W_f - t_null with W_f = R f + t (camera null pose == pose, Camera.py:11)
            const double w0 = (R[0] * f0 + R[1] * f1 + R[2] * f2 + t[0]) - t[0];
            const double w1 = (R[3] * f0 + R[4] * f1 + R[5] * f2 + t[1]) - t[1];
            const double w2 = (R[6] * f0 + R[7] * f1 + R[8] * f2 + t[2]) - t[2];
            const double u = uv[2 * i], v = uv[2 * i + 1];
            const double zx = Ki[0] * u + Ki[1] * v + Ki[2];
            const double zy = Ki[3] * u + Ki[4] * v + Ki[5];
            const double zw = Ki[6] * u + Ki[7] * v + Ki[8];
            sr[2 * i] = zx / zw - f0 / f2;
            sr[2 * i + 1] = zy / zw - f1 / f2;
            const double J00 = 1.0 / f2, J02 = -f0 / (f2 * f2), J11 = 1.0 / f2, J12 = -f1 / (f2 * f2);
            // H_x = [J skew(f), -J R^T];  H_f = J R^T
            double Hx[2][6];
            Hx[0][0] = J02 * (-f1);            // J00*0 + 0*f2 + J02*(-f1)
            Hx[0][1] = J00 * (-f2) + J02 * f0;
            Hx[0][2] = J00 * f1;
            Hx[1][0] = J11 * f2 + J12 * (-f1);
            Hx[1][1] = J12 * f0;
            Hx[1][2] = J11 * (-f0);
            double Hf[2][3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                Hf[0][k] = J00 * R[3 * k + 0] + J02 * R[3 * k + 2];
                Hf[1][k] = J11 * R[3 * k + 1] + J12 * R[3 * k + 2];
                Hx[0][3 + k] = -Hf[0][k];
                Hx[1][3 + k] = -Hf[1][k];
            }
            double u6[6];
            u6[0] = R[0] * g0 + R[3] * g1 + R[6] * g2;
            u6[1] = R[1] * g0 + R[4] * g1 + R[7] * g2;
            u6[2] = R[2] * g0 + R[5] * g1 + R[8] * g2;
            u6[3] = w1 * g2 - w2 * g1;
            u6[4] = w2 * g0 - w0 * g2;
            u6[5] = w0 * g1 - w1 * g0;
            double den = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) den += u6[k] * u6[k];
            if (den > 1e-6) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    double au = 0.0;
#pragma unroll
                    for (int k = 0; k < 6; ++k) au += Hx[r][k] * u6[k];
#pragma unroll
                    for (int k = 0; k < 6; ++k) Hx[r][k] = Hx[r][k] - (au * u6[k]) / den;
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int k = 0; k < 6; ++k) sA[(2 * i + r) * 6 + k] = Hx[r][k];
#pragma unroll
                for (int k = 0; k < 3; ++k) sQ[(2 * i + r) * 3 + k] = Hf[r][k];
            }
        }
    }
    __syncthreads();

    // ---- orthonormal basis of range(H_f): column-pivoted Gram-Schmidt with re-orthogonalisation (warp 0) ------------
    if (tid < 32) {
        const int lane = tid;
        double R3[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < 3; ++k) {
            double nrm[3] = {0, 0, 0};
            for (int r = lane; r < n2; r += 32)
                for (int c = k; c < 3; ++c) nrm[c] += sQ[r * 3 + c] * sQ[r * 3 + c];
            int piv = k;
            double best = -1.0;
            for (int c = k; c < 3; ++c) {
                nrm[c] = warp_sum(nrm[c]);
                if (nrm[c] > best) { best = nrm[c]; piv = c; }
            }
            if (piv != k) {
                for (int r = lane; r < n2; r += 32) {
                    const double tmp = sQ[r * 3 + k];
                    sQ[r * 3 + k] = sQ[r * 3 + piv];
                    sQ[r * 3 + piv] = tmp;
                }
                for (int r = 0; r < k; ++r) {
                    const double tmp = R3[r * 3 + k];
                    R3[r * 3 + k] = R3[r * 3 + piv];
                    R3[r * 3 + piv] = tmp;
                }
            }
            __syncwarp();
            const double nk = sqrt(best);
            R3[k * 3 + k] = nk;
            if (!(nk > 0.0)) break;
            const double ink = 1.0 / nk;
            for (int r = lane; r < n2; r += 32) sQ[r * 3 + k] *= ink;
            __syncwarp();
            for (int c = k + 1; c < 3; ++c) {
                for (int pass = 0; pass < 2; ++pass) {
                    double h = 0.0;
                    for (int r = lane; r < n2; r += 32) h += sQ[r * 3 + k] * sQ[r * 3 + c];
                    h = warp_sum(h);
                    for (int r = lane; r < n2; r += 32) sQ[r * 3 + c] -= h * sQ[r * 3 + k];
                    R3[k * 3 + c] += h;
                    __syncwarp();
                }
            }
        }
        if (lane == 0) {
            int rank = 0;
            // column pivoting bounds s_min >= |r33| / 4 and s_max <= sqrt(6) |r11|, so |r33| > 1e-10 |r11| already proves rank 3
            // (SciPy's threshold is ~1e-14 s_max); only doubtful factors pay for the Jacobi SVD
            if (n2 >= 3 && fabs(R3[8]) > 1e-10 * fabs(R3[0])) {
                rank = 3;
            } else {
                double G[9], sv[3];
                for (int k = 0; k < 9; ++k) G[k] = R3[k];
                svals3(G, sv);
                const double smax = fmax(sv[0], fmax(sv[1], sv[2]));
                const double tol = 2.220446049250313e-16 * (double)(n2 > 3 ? n2 : 3) * smax;   // scipy null_space rcond
                for (int k = 0; k < 3; ++k) rank += (sv[k] > tol) ? 1 : 0;
                if (rank > n2) rank = n2;
            }
            s_flag[1] = rank;
        }
    }
    __syncthreads();
    const int rank = s_flag[1];
    const int dofv = n2 - rank;
    for (int e = tid; e < n2 * 3; e += nth)
        if ((e % 3) >= rank) sQ[e] = 0.0;
    __syncthreads();

    // ---- w = Q1^T r, rt = r - Q1 w ----------------------------------------------------------------------------------
    if (tid < 3) {
        double acc = 0.0;
        for (int r = 0; r < n2; ++r) acc += sQ[r * 3 + tid] * sr[r];
        sm[24 + tid] = acc;
    }
    __syncthreads();
    for (int r = tid; r < n2; r += nth)
        srt[r] = sr[r] - (sQ[r * 3] * sm[24] + sQ[r * 3 + 1] * sm[25] + sQ[r * 3 + 2] * sm[26]);

    // ---- factors for the compression stage --------------------------------------------------------------------------
    for (int e = tid; e < n2 * 6; e += nth) facA[e] = sA[e];
    for (int e = tid; e < n2 * 3; e += nth) facQ[e] = sQ[e];
    __syncthreads();
    for (int e = tid; e < n2; e += nth) facR[e] = srt[e];

    if (dofv <= 0) {   // empty null space: gamma = 0 against chi2.ppf(0.95, 0) = NaN -> rejected (MSCKF.py:566-568)
        if (tid == 0) {
            const double crit = a.chi2_table[0];
            const bool pass = 0.0 <= crit;
            a.feat_status[fj] = (unsigned char)((pass ? MSCKF_FEAT_ACCEPTED : MSCKF_FEAT_GATED_OUT) | (code & MSCKF_FEAT_LOST_FLAG) | s_flag[2]);
            a.gamma[fj] = 0.0;
            a.dof[fj] = 0;
            a.fac_rank[fj] = rank;
        }
        return;
    }

    if (GEO_ONLY) {
        const int fl = (int)blockIdx.y * F + j;
        double2* Kt = go.Kbuf + (size_t)fl * go.kt_d2;
        for (int ob = tid; ob < M; ob += nth)
            ftile::write_appended_obs(Kt, M, ob, srt[2 * ob], srt[2 * ob + 1], sQ[(2 * ob) * 3], sQ[(2 * ob) * 3 + 1], sQ[(2 * ob) * 3 + 2],
                                      sQ[(2 * ob + 1) * 3], sQ[(2 * ob + 1) * 3 + 1], sQ[(2 * ob + 1) * 3 + 2]);
        ftile::write_padding_rows(Kt, M, tid, nth);
        unsigned char* mp = go.slot_map + (size_t)fl * 64;
        for (int sl = tid; sl < 64; sl += nth) mp[sl] = 0xFF;
        __syncthreads();
        for (int ob = tid; ob < M; ob += nth)
            if (sslot[ob] >= 0 && sslot[ob] < 64) mp[sslot[ob]] = (unsigned char)ob;
        if (tid == 0) {
            a.feat_status[fj] = (unsigned char)(MSCKF_FEAT_PENDING | (code & MSCKF_FEAT_LOST_FLAG) | s_flag[2]);
            a.dof[fj] = dofv;
            a.fac_rank[fj] = rank;
            go.pend_list[atomicAdd(go.pend_count, 1)] = fl;
        }
        return;
    }

    // ---- S_x = H_x P H_x^T from 6x6 blocks of P (lower triangle, packed) --------------------------------------------
    const int npairs = tri(M);
    for (int p = tid; p < npairs; p += nth) {
        const int i = s_trow[p];
        const int k = p - tri(i);
        const double* Pb = P + (size_t)(15 + 6 * sslot[i]) * ldp + 15 + 6 * sslot[k];
        double T0[6] = {0, 0, 0, 0, 0, 0}, T1[6] = {0, 0, 0, 0, 0, 0};
        const double* Ai = sA + (2 * i) * 6;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const double a0 = Ai[r], a1 = Ai[6 + r];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const double pv = __ldg(Pb + (size_t)r * ldp + c);
                T0[c] += a0 * pv;
                T1[c] += a1 * pv;
            }
        }
        const double* Ak = sA + (2 * k) * 6;
        double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            s00 += T0[c] * Ak[c];
            s01 += T0[c] * Ak[6 + c];
            s10 += T1[c] * Ak[c];
            s11 += T1[c] * Ak[6 + c];
        }
        const int r0 = 2 * i, c0 = 2 * k;
        sS[tri(r0) + c0] = s00;
        sS[tri(r0 + 1) + c0] = s10;
        sS[tri(r0 + 1) + c0 + 1] = s11;
        if (i != k) sS[tri(r0) + c0 + 1] = s01;
    }
    __syncthreads();

    // ---- U = Y - Q1 Z / 2 with Y = S_x Q1, Z = Q1^T Y;  S <- S - Q1 U^T - U Q1^T + sigma^2 I ------------------------
    for (int r = tid; r < n2; r += nth) {
        double y0 = 0, y1 = 0, y2 = 0;
        for (int c = 0; c < n2; ++c) {
            const double s = (c <= r) ? sS[tri(r) + c] : sS[tri(c) + r];
            y0 += s * sQ[c * 3];
            y1 += s * sQ[c * 3 + 1];
            y2 += s * sQ[c * 3 + 2];
        }
        sU[r * 3] = y0; sU[r * 3 + 1] = y1; sU[r * 3 + 2] = y2;
    }
    __syncthreads();
    if (tid < 9) {
        const int c1 = tid / 3, c2 = tid % 3;
        double acc = 0.0;
        for (int r = 0; r < n2; ++r) acc += sQ[r * 3 + c1] * sU[r * 3 + c2];
        sm[32 + tid] = acc;
    }
    __syncthreads();
    for (int r = tid; r < n2; r += nth) {
        const double q0 = sQ[r * 3], q1 = sQ[r * 3 + 1], q2 = sQ[r * 3 + 2];
#pragma unroll
        for (int c = 0; c < 3; ++c)
            sU[r * 3 + c] -= 0.5 * (q0 * sm[32 + c] + q1 * sm[35 + c] + q2 * sm[38 + c]);   // Z is symmetric up to rounding
    }
    __syncthreads();
    {
        const double s2 = a.par.sigma_image * a.par.sigma_image;
        const int nel = tri(n2);
        for (int e = tid; e < nel; e += nth) {
            const int r = s_trow[e];
            const int c = e - tri(r);
            double v = sS[e];
            v -= sQ[r * 3] * sU[c * 3] + sQ[r * 3 + 1] * sU[c * 3 + 1] + sQ[r * 3 + 2] * sU[c * 3 + 2];
            v -= sU[r * 3] * sQ[c * 3] + sU[r * 3 + 1] * sQ[c * 3 + 1] + sU[r * 3 + 2] * sQ[c * 3 + 2];
            if (r == c) v += s2;
            sS[e] = v;
        }
        for (int c = tid; c < n2; c += nth) sS[tri(n2) + c] = srt[c];   // appended row: the projected residual
    }
    __syncthreads();

    // ---- square-root-free Cholesky (LDL^T) of the (n2+1)-row packed triangle; gamma = sum_k y_k^2 / d_k ------------
    for (int k = 0; k < n2; ++k) {
        const double ipiv = 1.0 / sS[tri(k) + k];
        const int nt = n2 - k;   // trailing rows k+1 .. n2 (n2 = appended row)
        const int nel = tri(nt);
        for (int e = tid; e < nel; e += nth) {
            const int ri = s_trow[e];
            const int ci = e - tri(ri);
            const int r = k + 1 + ri, c = k + 1 + ci;
            if (r == n2 && c == n2) continue;
            sS[tri(r) + c] -= sS[tri(r) + k] * sS[tri(c) + k] * ipiv;
        }
        __syncthreads();
    }
    double part = 0.0;
    for (int k = tid; k < n2; k += nth) {
        const double y = sS[tri(n2) + k];
        part += y * y / sS[tri(k) + k];
    }
    part = warp_sum(part);
    if ((tid & 31) == 0) s_red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
        double g = 0.0;
        for (int w = 0; w < (nth + 31) / 32; ++w) g += s_red[w];
        const double crit = (dofv < a.chi2_len) ? a.chi2_table[dofv] : nan("");
        const bool pass = g <= crit;
        a.feat_status[fj] = (unsigned char)((pass ? MSCKF_FEAT_ACCEPTED : MSCKF_FEAT_GATED_OUT) | (code & MSCKF_FEAT_LOST_FLAG) | s_flag[2]);
        a.gamma[fj] = g;
        a.dof[fj] = dofv;
        a.fac_rank[fj] = rank;
    }
}

}  // namespace

static size_t front_smem_bytes(int Mcap) {
    const int n2 = 2 * Mcap;
    const size_t ntri = ((size_t)(n2 + 1) * (n2 + 2)) / 2;
    const size_t nS = ntri > (size_t)Mcap * 9 ? ntri : (size_t)Mcap * 9;   // the triangulation terms alias the S region
    const size_t doubles = (size_t)n2 * 6 + n2 * 3 + n2 + n2 + n2 * 3 + nS + 64;
    return doubles * sizeof(double) + (size_t)Mcap * sizeof(int) + ntri + 16;
}

cudaError_t launch_front(const FrontArgs& a, cudaStream_t s) {
    const size_t smem = front_smem_bytes(a.dims.M);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(front_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    dim3 grid(a.dims.F, a.dims.B);
    front_kernel<false><<<grid, FRONT_THREADS, smem, s>>>(a, GeoOut{});
    msckf_count_launch(1);
    return cudaGetLastError();
}

// geometry-only mode for the filters [b0, b0 + nb) (the split front end of tracks longer than 32 observations)
cudaError_t launch_front_geo_wide(const FrontArgs& a, double2* Kbuf, int kt_d2, unsigned char* slot_map, int* pend_list,
                                  int* pend_count, int b0, int nb, cudaStream_t s) {
    const size_t smem = front_smem_bytes(a.dims.M);      // (sized for the full kernel: an upper bound)
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(front_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const int n2 = 2 * a.dims.M;
    const size_t geo_smem = ((size_t)n2 * 6 + n2 * 3 + n2 + n2 + n2 * 3 + (size_t)a.dims.M * 9 + 64) * sizeof(double) +
                            (size_t)a.dims.M * sizeof(int) + (size_t)(n2 + 1) * (n2 + 2) / 2 + 16;
    dim3 grid(a.dims.F, nb);
    front_kernel<true><<<grid, FRONT_THREADS, geo_smem, s>>>(a, GeoOut{Kbuf, kt_d2, slot_map, pend_list, pend_count, b0});
    msckf_count_launch(1);
    return cudaGetLastError();
}
