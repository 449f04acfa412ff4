// Per-feature front end, warp-per-feature version (rows 3-8 of SURVEY.md section 8a) for tracks of up to 32 observations
// and windows whose camera covariance block fits one SM's shared memory (N <= 31):
//   selection + closed-form triangulation     reference src/msckf/MSCKF.py:458-495, src/utils/geometry.py:274-303
//   residual + Jacobian blocks                MSCKF.py:497-548, src/msckf/Camera.py:54-67
//   left-null-space projection of H_f         MSCKF.py:554-559
//   chi-square Mahalanobis gate               MSCKF.py:561-568
//
// One 256-thread CTA per (filter, chunk of features).  The CTA stages the camera block of P as 6x6 blocks (lower block
// triangle, 143 KB at N = 31) and the window poses in shared memory ONCE and its 8 warps then walk the chunk, one feature
// per warp at a time, lane i = observation i:
//   * geometry (triangulation terms, residual, 2x6 / 2x3 Jacobian blocks, Gram-Schmidt basis Q1 of range(H_f)) is
//     lane-local arithmetic plus warp-shuffle reductions;
//   * K = H_x P H_x^T + sigma^2 I is assembled from the staged 6x6 blocks directly into the 8x8 accumulator-fragment
//     layout of the FP64 tensor instruction (36 lower tiles = 72 registers per lane);
//   * gamma uses the basis-free identity  A (A^T K A)^-1 A^T = K^-1 - K^-1 Q (Q^T K^-1 Q)^-1 Q^T K^-1  (A = null-space
//     basis, Q = Q1): the rows [r~; Q1^T] are appended to K as a ninth tile row, one register-resident blocked Cholesky
//     (DMMA.8x8x4 for the panel solves and trailing updates, shuffles inside the 8x8 diagonal tiles) turns them into
//     y = L^-1 r~ and V = L^-1 Q1, and  gamma = y.y - (V^T y)^T (V^T V)^-1 (V^T y).
//     Feeding both operands of the tensor instruction with the k-index permutation kappa = 2 t + step makes the
//     accumulator layout double as the A and the B layout, so the factorisation needs no data movement between lanes
//     outside the diagonal tiles.
// No explicit null-space basis, no dense 2M x d Jacobian, no projected 2M x 2M matrix is ever formed.
#include "msckf_common.cuh"
#include "dense.cuh"

namespace {

constexpr int FW_THREADS = 256;
constexpr int FW_WARPS = FW_THREADS / 32;
constexpr int TC = 8;                         // 8x8 tile columns -> up to 64 measurement rows (32 observations)
constexpr int NTILE = TC * (TC + 1) / 2;
constexpr int FW_MAX_OBS = 4 * TC;

struct WarpScratch {
    double A[2 * FW_MAX_OBS][6];     // observability-projected Jacobian rows
    double aug[2 * FW_MAX_OBS][4];   // r~, Q1[:, 0..2] per measurement row
    double2 rowbuf[TC][32];          // one tile row of K in accumulator-fragment layout (staging for the registers)
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// 1/sqrt(x) for x > 0 (0 for x <= 0): hardware seed (MUFU.RSQ64H, ~2^-22) + two Newton steps, no slow-path branch
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    return x > 0.0 ? y : 0.0;
}

__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// cyclic Jacobi eigen-decomposition of a symmetric 3x3 (a = [xx, xy, xz, yy, yz, zz]); V columns = eigenvectors
__device__ void sym3_eig_w(const double* a, double* lam, double* V) {
    double A[3][3] = {{a[0], a[1], a[2]}, {a[1], a[3], a[4]}, {a[2], a[4], a[5]}};
    double Q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        const double dia = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-300 || off <= 1e-18 * dia) break;
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double qkp = Q[k][p], qkq = Q[k][q];
                    Q[k][p] = c * qkp - s * qkq;
                    Q[k][q] = s * qkp + c * qkq;
                }
            }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lam[k] = A[k][k];
#pragma unroll
        for (int r = 0; r < 3; ++r) V[r * 3 + k] = Q[r][k];
    }
}

// p = X^-1 y for a comfortably non-singular symmetric positive definite 3x3 X (a = [xx, xy, xz, yy, yz, zz]) by LDL^T.
// Returns false when det(X) <= 1e-9 tr(X)^3 (lambda_min / lambda_max could be below ~1e-9): only then can the 1e-15
// singular-value cut-off of np.linalg.pinv (geometry.py:299) matter, and the caller takes the eigen-decomposition path.
__device__ __forceinline__ bool sym3_solve_spd(const double* a, const double* y, double* p) {
    const double d0 = a[0];
    if (!(d0 > 0.0)) return false;
    const double i0 = 1.0 / d0;
    const double l10 = a[1] * i0, l20 = a[2] * i0;
    const double d1 = a[3] - l10 * a[1];
    if (!(d1 > 0.0)) return false;
    const double i1 = 1.0 / d1;
    const double e21 = a[4] - l20 * a[1];
    const double l21 = e21 * i1;
    const double d2 = a[5] - l20 * a[2] - l21 * e21;
    const double tr = a[0] + a[3] + a[5];
    if (!(d2 > 0.0) || !(d0 * d1 * d2 > 1e-9 * tr * tr * tr)) return false;
    const double z0 = y[0], z1 = y[1] - l10 * z0, z2 = y[2] - l20 * z0 - l21 * z1;
    p[2] = z2 / d2;
    p[1] = z1 * i1 - l21 * p[2];
    p[0] = z0 * i0 - l10 * p[1] - l20 * p[2];
    return true;
}

// singular values of a 3x3 matrix G (row-major, destroyed) by one-sided Jacobi
__device__ void svals3_w(double* G, double* sv) {
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = p + 1; q < 3; ++q) {
                double al = 0, be = 0, ga = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    al += G[k * 3 + p] * G[k * 3 + p];
                    be += G[k * 3 + q] * G[k * 3 + q];
                    ga += G[k * 3 + p] * G[k * 3 + q];
                }
                if (ga == 0.0 || fabs(ga) <= 1e-16 * sqrt(al * be)) continue;
                rotated = true;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double gp = G[k * 3 + p], gq = G[k * 3 + q];
                    G[k * 3 + p] = c * gp - s * gq;
                    G[k * 3 + q] = s * gp + c * gq;
                }
            }
        if (!rotated) break;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) sv[c] = sqrt(G[c] * G[c] + G[3 + c] * G[3 + c] + G[6 + c] * G[6 + c]);
}

// files one finished tile row (tiles (I, 0..I)) into the statically indexed lower-triangular register array
template <int I>
__device__ __forceinline__ void file_row(double (&S)[NTILE][2], const double2 (*rowbuf)[32], int lane) {
#pragma unroll
    for (int J = 0; J <= I; ++J) {
        const double2 v = rowbuf[J][lane];
        S[I * (I + 1) / 2 + J][0] = v.x;
        S[I * (I + 1) / 2 + J][1] = v.y;
    }
}

struct ObsRegs {
    double u, v, dx, dy, dz, cf;
    int slot;
};

__device__ __forceinline__ void load_obs(ObsRegs& o, const FrontArgs& a, size_t fj, int M, int lane) {
    const int Mcap = a.dims.M;
    if (lane < M) {
        const double2 uv = *reinterpret_cast<const double2*>(a.obs_uv + (fj * Mcap + lane) * 2);
        o.u = uv.x; o.v = uv.y;
        const double* d = a.obs_dir + (fj * Mcap + lane) * 3;
        o.dx = d[0]; o.dy = d[1]; o.dz = d[2];
        o.cf = a.obs_conf[fj * Mcap + lane];
        o.slot = a.obs_slot[fj * Mcap + lane];
    } else {
        o.u = o.v = 0.0; o.dx = o.dy = 0.0; o.dz = 1.0; o.cf = 0.0; o.slot = 0;
    }
}

__global__ void __launch_bounds__(FW_THREADS, 1) front_warp_kernel(const FrontArgs a, const int feats_per_cta) {
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int F = a.dims.F, Mcap = a.dims.M, Ncap = a.dims.N, ldp = a.dims.ldp;
    const int nc = min(a.n_cams[b], Ncap);

    extern __shared__ double smem[];
    double* sP = smem;                                              // [nblk(Ncap)][36] lower block triangle of P_CC
    double* sR = sP + (size_t)(Ncap * (Ncap + 1) / 2) * 36;         // [Ncap][9]
    double* sT = sR + (size_t)Ncap * 9;                             // [Ncap][3]
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(sT + (size_t)Ncap * 3) + warp;

    const int f_begin = blockIdx.x * feats_per_cta;
    const int f_end = min(F, f_begin + feats_per_cta);
    if (f_begin >= f_end) return;

    // ---- stage the window once per CTA ----
    {
        const double* P = a.P + (size_t)b * ldp * ldp;
        const int c = 6 * nc;
        for (int e = tid; e < c * c; e += FW_THREADS) {
            const int row = e / c, col = e - row * c;
            const int ba = row / 6, bb = col / 6;
            if (bb <= ba) sP[(size_t)(ba * (ba + 1) / 2 + bb) * 36 + (row - 6 * ba) * 6 + (col - 6 * bb)] = P[(size_t)(15 + row) * ldp + 15 + col];
        }
        const double* camR = a.cam_R + (size_t)b * Ncap * 9;
        const double* camt = a.cam_t + (size_t)b * Ncap * 3;
        for (int e = tid; e < nc * 9; e += FW_THREADS) sR[e] = camR[e];
        for (int e = tid; e < nc * 3; e += FW_THREADS) sT[e] = camt[e];
    }
    __syncthreads();

    const double s2 = a.par.sigma_image * a.par.sigma_image;
    const double g0 = a.par.gravity[0], g1 = a.par.gravity[1], g2 = a.par.gravity[2];
    const bool cos_ok = a.par.min_parallax_deg > 0.0 && a.par.min_parallax_deg < 180.0;
    const double cos_thr = cos_ok ? cos(a.par.min_parallax_deg * (3.14159265358979323846 / 180.0)) : 0.0;

    ObsRegs nxt;
    int Mn = 0;
    {
        const int j0 = f_begin + warp;
        if (j0 < f_end) {
            Mn = a.feat_len[(size_t)b * F + j0];
            load_obs(nxt, a, (size_t)b * F + j0, min(max(Mn, 0), FW_MAX_OBS), lane);
        }
    }

    for (int j = f_begin + warp; j < f_end; j += FW_WARPS) {
        const size_t fj = (size_t)b * F + j;
        const ObsRegs o = nxt;
        const int M = Mn;
        {   // prefetch the next feature of this warp
            const int jn = j + FW_WARPS;
            if (jn < f_end) {
                Mn = a.feat_len[(size_t)b * F + jn];
                load_obs(nxt, a, (size_t)b * F + jn, min(max(Mn, 0), FW_MAX_OBS), lane);
            }
        }
        if (M <= 0 || M > Mcap || M > FW_MAX_OBS) {
            if (lane == 0) {
                a.feat_status[fj] = MSCKF_FEAT_SKIPPED;
                a.gamma[fj] = nan("");
                a.dof[fj] = 0;
                a.fac_rank[fj] = 0;
            }
            continue;
        }
        const int n2 = 2 * M;
        const bool live = lane < M;

        // ---- selection (MSCKF.py:463-479) ----
        int code;
        {
            const bool lost = a.lost_cnt[fj] >= a.par.min_lost;
            if (lost && a.tracked_cnt[fj] < a.par.min_tracked) {
                code = MSCKF_FEAT_LOST_DISCARD | MSCKF_FEAT_LOST_FLAG;
            } else {
                bool enough = false;
                if (a.par.use_parallax && M > 1) {
                    const double ax = shfl_d(o.dx, 0), ay = shfl_d(o.dy, 0), az = shfl_d(o.dz, 0);
                    const double bx = shfl_d(o.dx, M - 1), by = shfl_d(o.dy, M - 1), bz = shfl_d(o.dz, M - 1);
                    // acos(dp) in degrees > min_parallax  <=>  dp < cos(min_parallax): decided on the cosine unless it sits
                    // within 1e-9 of the threshold, where the reference's own expression is evaluated (geometry.py:237-256)
                    const double na = ax * ax + ay * ay + az * az, nb = bx * bx + by * by + bz * bz;
                    const double dq = (ax * bx + ay * by + az * bz) * fast_rsqrt(na * nb);
                    if (cos_ok && na * nb > 0.0 && fabs(dq - cos_thr) > 1e-9) {
                        enough = dq < cos_thr;
                    } else {
                        const double n1 = sqrt(na), n2_ = sqrt(nb);
                        double dp = (ax / n1) * (bx / n2_) + (ay / n1) * (by / n2_) + (az / n1) * (bz / n2_);
                        dp = fmin(1.0, fmax(-1.0, dp));
                        enough = acos(dp) * (180.0 / 3.14159265358979323846) > a.par.min_parallax_deg;
                    }
                }
                code = (lost || enough) ? (MSCKF_FEAT_ACCEPTED | (lost ? MSCKF_FEAT_LOST_FLAG : 0)) : MSCKF_FEAT_SKIPPED;
            }
        }
        if ((code & 0xF) != MSCKF_FEAT_ACCEPTED) {
            if (lane == 0) {
                a.feat_status[fj] = (unsigned char)code;
                a.gamma[fj] = nan("");
                a.dof[fj] = 0;
                a.fac_rank[fj] = 0;
            }
            continue;
        }

        const double* R = sR + 9 * o.slot;
        const double* tt = sT + 3 * o.slot;

        // ---- triangulation: X = sum c (I - d d^T), y = sum c (I - d d^T) t_slot, p = pinv(X) y (geometry.py:286-299) ----
        double rho, m0, m1, m2, b0, b1, b2;
        {
            double term[9];
            {
                const double nn = sqrt(o.dx * o.dx + o.dy * o.dy + o.dz * o.dz);
                const double ux = o.dx / nn, uy = o.dy / nn, uz = o.dz / nn;
                const double c = live ? o.cf : 0.0;
                term[0] = c * (1.0 - ux * ux); term[1] = c * (0.0 - ux * uy); term[2] = c * (0.0 - ux * uz);
                term[3] = c * (1.0 - uy * uy); term[4] = c * (0.0 - uy * uz); term[5] = c * (1.0 - uz * uz);
                term[6] = term[0] * tt[0] + term[1] * tt[1] + term[2] * tt[2];
                term[7] = term[1] * tt[0] + term[3] * tt[1] + term[4] * tt[2];
                term[8] = term[2] * tt[0] + term[4] * tt[1] + term[5] * tt[2];
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) term[k] = warp_sum_all(live ? term[k] : 0.0);
            double p[3] = {0, 0, 0};
            if (!sym3_solve_spd(term, term + 6, p)) {   // (nearly) singular X: pinv with its singular-value cut-off
                double lam[3], V[9];
                sym3_eig_w(term, lam, V);
                const double lmax = fmax(fabs(lam[0]), fmax(fabs(lam[1]), fabs(lam[2])));
                p[0] = p[1] = p[2] = 0.0;
#pragma unroll 1
                for (int k = 0; k < 3; ++k) {
                    if (fabs(lam[k]) > 1e-15 * lmax) {   // np.linalg.pinv default rcond
                        const double coef = (V[0 + k] * term[6] + V[3 + k] * term[7] + V[6 + k] * term[8]) / lam[k];
                        p[0] += coef * V[0 + k];
                        p[1] += coef * V[3 + k];
                        p[2] += coef * V[6 + k];
                    }
                }
            }
            // first observing camera (MSCKF.py:481-488)
            const int s0 = __shfl_sync(0xffffffffu, o.slot, 0);
            const double* R0 = sR + 9 * s0;
            const double* t0 = sT + 3 * s0;
            const double q0 = p[0] - t0[0], q1 = p[1] - t0[1], q2 = p[2] - t0[2];
            const double cx = R0[0] * q0 + R0[3] * q1 + R0[6] * q2;
            const double cy = R0[1] * q0 + R0[4] * q1 + R0[7] * q2;
            const double cz = R0[2] * q0 + R0[5] * q1 + R0[8] * q2;
            rho = a.idp_rho[fj];
            m0 = a.idp_m[fj * 3]; m1 = a.idp_m[fj * 3 + 1]; m2 = a.idp_m[fj * 3 + 2];
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
                    if (lane == 0) {
                        a.idp_rho[fj] = rho;
                        a.idp_m[fj * 3] = m0; a.idp_m[fj * 3 + 1] = m1; a.idp_m[fj * 3 + 2] = m2;
                    }
                }
            }
            const int as = a.idp_anchor_slot[fj];
            const double* base = (as >= 0) ? (sT + 3 * as) : (a.idp_base + fj * 3);
            b0 = base[0]; b1 = base[1]; b2 = base[2];
        }

        // ---- residual and Jacobian blocks of this lane's observation (MSCKF.py:505-544, Camera.py:54-67) ----
        double Hx[2][6], Hf[2][3], res[2];
        {
            const double q0 = rho * (b0 - tt[0]) + m0, q1 = rho * (b1 - tt[1]) + m1, q2 = rho * (b2 - tt[2]) + m2;
            const double f0 = R[0] * q0 + R[3] * q1 + R[6] * q2;
            const double f1 = R[1] * q0 + R[4] * q1 + R[7] * q2;
            const double f2 = R[2] * q0 + R[5] * q1 + R[8] * q2;
            const double w0 = (R[0] * f0 + R[1] * f1 + R[2] * f2 + tt[0]) - tt[0];
            const double w1 = (R[3] * f0 + R[4] * f1 + R[5] * f2 + tt[1]) - tt[1];
            const double w2 = (R[6] * f0 + R[7] * f1 + R[8] * f2 + tt[2]) - tt[2];
            const double* Ki = a.par.Kinv;
            const double zx = Ki[0] * o.u + Ki[1] * o.v + Ki[2];
            const double zy = Ki[3] * o.u + Ki[4] * o.v + Ki[5];
            const double zw = Ki[6] * o.u + Ki[7] * o.v + Ki[8];
            res[0] = zx / zw - f0 / f2;
            res[1] = zy / zw - f1 / f2;
            const double J00 = 1.0 / f2, J02 = -f0 / (f2 * f2), J11 = 1.0 / f2, J12 = -f1 / (f2 * f2);
            Hx[0][0] = J02 * (-f1);
            Hx[0][1] = J00 * (-f2) + J02 * f0;
            Hx[0][2] = J00 * f1;
            Hx[1][0] = J11 * f2 + J12 * (-f1);
            Hx[1][1] = J12 * f0;
            Hx[1][2] = J11 * (-f0);
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
            if (!live) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    res[r] = 0.0;
#pragma unroll
                    for (int k = 0; k < 6; ++k) Hx[r][k] = 0.0;
#pragma unroll
                    for (int k = 0; k < 3; ++k) Hf[r][k] = 0.0;
                }
            }
        }

        // ---- orthonormal basis of range(H_f): column-pivoted Gram-Schmidt with re-orthogonalisation ----
        int rank;
        {
            double R3[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            bool stop = false;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (stop) break;
                double nrm[3] = {0, 0, 0};
                int piv = k;
                double best = -1.0;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (c < k) continue;
                    nrm[c] = warp_sum_all(Hf[0][c] * Hf[0][c] + Hf[1][c] * Hf[1][c]);
                    if (nrm[c] > best) { best = nrm[c]; piv = c; }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (c > k && c == piv) {
#pragma unroll
                        for (int r = 0; r < 2; ++r) { const double tmp = Hf[r][k]; Hf[r][k] = Hf[r][c]; Hf[r][c] = tmp; }
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            if (r < k) { const double tmp = R3[r * 3 + k]; R3[r * 3 + k] = R3[r * 3 + c]; R3[r * 3 + c] = tmp; }
                        }
                    }
                }
                const double nk = sqrt(best);
                R3[k * 3 + k] = nk;
                if (!(nk > 0.0)) { stop = true; break; }
                const double ink = 1.0 / nk;
                Hf[0][k] *= ink; Hf[1][k] *= ink;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (c <= k) continue;
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        const double h = warp_sum_all(Hf[0][k] * Hf[0][c] + Hf[1][k] * Hf[1][c]);
                        Hf[0][c] -= h * Hf[0][k];
                        Hf[1][c] -= h * Hf[1][k];
                        R3[k * 3 + c] += h;
                    }
                }
            }
            // rank by SciPy's rule s > eps max(3, 2M) s_max on the singular values of the 3x3 factor.  Column pivoting
            // bounds s_min >= |r33| / 4 and s_max <= sqrt(6) |r11|, so |r33| > 1e-10 |r11| already proves rank 3 (the
            // threshold is ~1e-14 s_max); only doubtful factors pay for the Jacobi SVD.
            if (n2 >= 3 && fabs(R3[8]) > 1e-10 * fabs(R3[0])) {
                rank = 3;
            } else {
                double G3[9], sv[3];
#pragma unroll
                for (int k = 0; k < 9; ++k) G3[k] = R3[k];
                svals3_w(G3, sv);
                const double smax = fmax(sv[0], fmax(sv[1], sv[2]));
                const double tol = 2.220446049250313e-16 * (double)(n2 > 3 ? n2 : 3) * smax;   // scipy null_space rcond
                rank = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) rank += (sv[k] > tol) ? 1 : 0;
                if (rank > n2) rank = n2;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (c >= rank) { Hf[0][c] = 0.0; Hf[1][c] = 0.0; }
        }
        const int dofv = n2 - rank;

        // ---- r~ = r - Q1 Q1^T r ----
        double rt0, rt1;
        {
            double w[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) w[c] = warp_sum_all(Hf[0][c] * res[0] + Hf[1][c] * res[1]);
            rt0 = res[0] - (Hf[0][0] * w[0] + Hf[0][1] * w[1] + Hf[0][2] * w[2]);
            rt1 = res[1] - (Hf[1][0] * w[0] + Hf[1][1] * w[1] + Hf[1][2] * w[2]);
        }

        // ---- factors for the compression stage + the warp's shared scratch ----
        __syncwarp();
        if (live) {
            double* fA = a.fac_A + (fj * Mcap + lane) * 12;
            double* fQ = a.fac_Q + (fj * Mcap + lane) * 6;
            double* fR = a.fac_rt + (fj * Mcap + lane) * 2;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int k = 0; k < 6; ++k) { fA[r * 6 + k] = Hx[r][k]; ws->A[2 * lane + r][k] = Hx[r][k]; }
#pragma unroll
                for (int k = 0; k < 3; ++k) { fQ[r * 3 + k] = Hf[r][k]; ws->aug[2 * lane + r][1 + k] = Hf[r][k]; }
            }
            fR[0] = rt0; fR[1] = rt1;
            ws->aug[2 * lane][0] = rt0;
            ws->aug[2 * lane + 1][0] = rt1;
        }
        __syncwarp();

        if (dofv <= 0) {   // empty null space: gamma = 0 against chi2.ppf(0.95, 0) = NaN -> rejected (MSCKF.py:566-568)
            if (lane == 0) {
                const double crit = a.chi2_table[0];
                const bool pass = 0.0 <= crit;
                a.feat_status[fj] = (unsigned char)((pass ? MSCKF_FEAT_ACCEPTED : MSCKF_FEAT_GATED_OUT) | (code & MSCKF_FEAT_LOST_FLAG));
                a.gamma[fj] = 0.0;
                a.dof[fj] = 0;
                a.fac_rank[fj] = rank;
            }
            continue;
        }

        // ---- exact early-out (MSCKF_GATE_BOUNDS): |r~|^2 / (s^2 + tr S_x) <= gamma <= |r~|^2 / s^2 ----
        if (a.par.gate_mode == MSCKF_GATE_BOUNDS) {
            double trs = 0.0;
            if (live) {
                const double* Pd = sP + (size_t)(o.slot * (o.slot + 1) / 2 + o.slot) * 36;
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    double v0 = 0.0, v1 = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) { v0 += Pd[r * 6 + c] * Hx[0][c]; v1 += Pd[r * 6 + c] * Hx[1][c]; }
                    trs += Hx[0][r] * v0 + Hx[1][r] * v1;
                }
            }
            trs = warp_sum_all(trs);
            const double rr = warp_sum_all(rt0 * rt0 + rt1 * rt1);
            const double crit = (dofv < a.chi2_len) ? a.chi2_table[dofv] : nan("");
            const double gub = rr / s2, glb = rr / (s2 + trs);
            const bool acc = gub <= crit, rej = glb > crit;
            if (acc || rej) {
                if (lane == 0) {
                    a.feat_status[fj] = (unsigned char)((acc ? MSCKF_FEAT_ACCEPTED : MSCKF_FEAT_GATED_OUT) |
                                                        (code & MSCKF_FEAT_LOST_FLAG) | MSCKF_FEAT_GAMMA_BOUND);
                    a.gamma[fj] = acc ? gub : glb;
                    a.dof[fj] = dofv;
                    a.fac_rank[fj] = rank;
                }
                continue;
            }
        }

        // ---- K = H_x P H_x^T + sigma^2 I in accumulator-fragment layout; appended tile row [r~; Q1^T] ----
        // (runtime loop over tile rows + a switch that files the finished row into the statically indexed register
        //  array: the fully unrolled form of this kernel was 400 KB of SASS and stalled 70% of the time on instruction
        //  fetch, profiles/r01d)
        const int Tn = (n2 + 7) >> 3;
        double S[NTILE][2];
        double Ga[TC][2];
        // Each lane computes one whole 2x2 block  A_i P[s_i, s_k] A_k^T  (so every 6x6 block of P is read from shared
        // memory by exactly one lane): row observation i = 4 I + (g >> 1), column observation k = 4 J + t, and the low bit
        // of g picks the tile J = 2 Jp + h of a PAIR of tiles handled per pass.  One exchange with lane ^ 4 then puts row
        // 2 i + (g & 1) of both tiles into the accumulator-fragment layout.
        const int h = g & 1;
        unsigned sk_lo = 0, sk_hi = 0;      // slots of this lane's column observations k = 4 J + t, one byte per J
#pragma unroll
        for (int J = 0; J < TC; ++J) {
            const int s_k = __shfl_sync(0xffffffffu, o.slot, (4 * J + t) & 31) & 0xff;
            if (J < 4) sk_lo |= (unsigned)s_k << (8 * J); else sk_hi |= (unsigned)s_k << (8 * (J - 4));
        }
#pragma unroll 1
        for (int I = 0; I < Tn; ++I) {
            const int i = 4 * I + (g >> 1);
            const bool iv = i < M;
            const int si = __shfl_sync(0xffffffffu, o.slot, i & 31);
            double ai[2][6];
            {
                const double2* Ai = reinterpret_cast<const double2*>(&ws->A[2 * (iv ? i : 0)][0]);
#pragma unroll
                for (int c2 = 0; c2 < 6; ++c2) {
                    const double2 x = Ai[c2];
                    ai[c2 / 3][2 * (c2 % 3)] = iv ? x.x : 0.0;
                    ai[c2 / 3][2 * (c2 % 3) + 1] = iv ? x.y : 0.0;
                }
            }
#pragma unroll 1
            for (int Jp = 0; 2 * Jp <= I; ++Jp) {
                const int J = 2 * Jp + h;
                const int k = 4 * J + t;
                double o00 = 0.0, o01 = 0.0, o10 = 0.0, o11 = 0.0;
                if (iv && J <= I && k < M && !(I == J && k > i)) {
                    const int sk = (int)(((J < 4 ? sk_lo : sk_hi) >> (8 * (J & 3))) & 0xffu);
                    double v0[6], v1[6];
#pragma unroll
                    for (int c = 0; c < 6; ++c) { v0[c] = 0.0; v1[c] = 0.0; }
                    if (si >= sk) {
                        const double2* Pb = reinterpret_cast<const double2*>(sP + (size_t)(si * (si + 1) / 2 + sk) * 36);
#pragma unroll
                        for (int r = 0; r < 6; ++r) {
#pragma unroll
                            for (int c2 = 0; c2 < 3; ++c2) {
                                const double2 pv = Pb[r * 3 + c2];
                                v0[2 * c2] += ai[0][r] * pv.x;
                                v0[2 * c2 + 1] += ai[0][r] * pv.y;
                                v1[2 * c2] += ai[1][r] * pv.x;
                                v1[2 * c2 + 1] += ai[1][r] * pv.y;
                            }
                        }
                    } else {   // transposed block (slots not increasing with the observation index)
                        const double2* Pb = reinterpret_cast<const double2*>(sP + (size_t)(sk * (sk + 1) / 2 + si) * 36);
#pragma unroll
                        for (int c = 0; c < 6; ++c) {
#pragma unroll
                            for (int r2 = 0; r2 < 3; ++r2) {
                                const double2 pv = Pb[c * 3 + r2];
                                v0[c] += ai[0][2 * r2] * pv.x + ai[0][2 * r2 + 1] * pv.y;
                                v1[c] += ai[1][2 * r2] * pv.x + ai[1][2 * r2 + 1] * pv.y;
                            }
                        }
                    }
                    const double2* Ak = reinterpret_cast<const double2*>(&ws->A[2 * k][0]);
#pragma unroll
                    for (int c2 = 0; c2 < 3; ++c2) {
                        const double2 x0 = Ak[c2], x1 = Ak[3 + c2];
                        o00 += v0[2 * c2] * x0.x + v0[2 * c2 + 1] * x0.y;
                        o01 += v0[2 * c2] * x1.x + v0[2 * c2 + 1] * x1.y;
                        o10 += v1[2 * c2] * x0.x + v1[2 * c2 + 1] * x0.y;
                        o11 += v1[2 * c2] * x1.x + v1[2 * c2 + 1] * x1.y;
                    }
                }
                // even lanes (h = 0) computed tile 2 Jp and keep its row 2 i; odd lanes computed tile 2 Jp + 1 and keep its
                // row 2 i + 1; the other row of each block goes to the partner lane
                const double r0 = __shfl_xor_sync(0xffffffffu, h ? o00 : o10, 4);
                const double r1 = __shfl_xor_sync(0xffffffffu, h ? o01 : o11, 4);
                double e0 = h ? r0 : o00, e1 = h ? r1 : o01;       // fragment of tile 2 Jp
                double f0 = h ? o10 : r0, f1 = h ? o11 : r1;       // fragment of tile 2 Jp + 1
                {   // sigma^2 on the diagonal, identity on padded rows
                    const int row = 8 * I + g;
                    const double dg = (row < n2) ? s2 : 1.0;
                    if (2 * Jp == I) {
                        if (g == 2 * t) e0 += dg;
                        if (g == 2 * t + 1) e1 += dg;
                    }
                    if (2 * Jp + 1 == I) {
                        if (g == 2 * t) f0 += dg;
                        if (g == 2 * t + 1) f1 += dg;
                    }
                }
                ws->rowbuf[2 * Jp][lane] = make_double2(e0, e1);
                if (2 * Jp + 1 <= I) ws->rowbuf[2 * Jp + 1][lane] = make_double2(f0, f1);
            }
            // the finished tile row goes from the staging buffer into the statically indexed register array
            switch (I) {
                case 0: file_row<0>(S, ws->rowbuf, lane); break;
                case 1: file_row<1>(S, ws->rowbuf, lane); break;
                case 2: file_row<2>(S, ws->rowbuf, lane); break;
                case 3: file_row<3>(S, ws->rowbuf, lane); break;
                case 4: file_row<4>(S, ws->rowbuf, lane); break;
                case 5: file_row<5>(S, ws->rowbuf, lane); break;
                case 6: file_row<6>(S, ws->rowbuf, lane); break;
                default: file_row<7>(S, ws->rowbuf, lane); break;
            }
        }
#pragma unroll
        for (int J = 0; J < TC; ++J) {
            double e0 = 0.0, e1 = 0.0;
            if (J < Tn && g < 4) {
                const int col = 8 * J + 2 * t;
                if (col < n2) e0 = ws->aug[col][g];
                if (col + 1 < n2) e1 = ws->aug[col + 1][g];
            }
            Ga[J][0] = e0;
            Ga[J][1] = e1;
        }

        // ---- blocked Cholesky of K with the appended rows riding along.  Every pass factors block column 0 of the
        //      remaining matrix and then shifts the trailing tiles up-left by one, so the loop body is one fixed piece of
        //      code (statically indexed registers) that stays in the instruction cache. ----
        // gamma = y.y - b^T (V V^T)^-1 b with y = row 0 of the appended tile row, V = rows 1..3, b = V y
        double gr0 = 0.0, gr1 = 0.0;   // accumulator tile of [y; V] [y; V]^T (rows/cols 0..3 are meaningful)
#pragma unroll 1
        for (int bk = 0; bk < Tn; ++bk) {
            const int rem = Tn - bk;
            double D0 = S[0][0], D1 = S[0][1];
            double x0 = (g == 2 * t) ? 1.0 : 0.0, x1 = (g == 2 * t + 1) ? 1.0 : 0.0;
            // in-tile factorisation: two columns (k = 2 tk, 2 tk + 1) per pass so the register parity is static.  Column
            // entries of finished rows (g <= k) are zeroed at the source lane, which makes every rank-1 update below
            // unconditional: finished rows and finished columns of X = L^-T receive exact zeros.
#pragma unroll 1
            for (int tk = 0; tk < 4; ++tk) {
                {
                    const int k = 2 * tk;
                    const double piv = shfl_d(D0, k * 4 + tk);
                    const double rs = fast_rsqrt(piv);
                    const double colv = (g > k) ? D0 * rs : 0.0;      // column k of L below the diagonal
                    const double xsel = x0 * rs;                       // column k of X L^-T
                    if (t == tk) x0 = xsel;
                    const double lg = shfl_d(colv, g * 4 + tk);
                    const double xg = shfl_d(xsel, g * 4 + tk);
                    const double lc0 = shfl_d(colv, (2 * t) * 4 + tk);
                    const double lc1 = shfl_d(colv, (2 * t + 1) * 4 + tk);
                    D0 -= lg * lc0; x0 -= xg * lc0;
                    D1 -= lg * lc1; x1 -= xg * lc1;
                }
                {
                    const int k = 2 * tk + 1;
                    const double piv = shfl_d(D1, k * 4 + tk);
                    const double rs = fast_rsqrt(piv);
                    const double colv = (g > k) ? D1 * rs : 0.0;
                    const double xsel = x1 * rs;
                    if (t == tk) x1 = xsel;
                    const double lg = shfl_d(colv, g * 4 + tk);
                    const double xg = shfl_d(xsel, g * 4 + tk);
                    const double lc0 = shfl_d(colv, (2 * t) * 4 + tk);
                    const double lc1 = shfl_d(colv, (2 * t + 1) * 4 + tk);
                    D0 -= lg * lc0; x0 -= xg * lc0;
                    D1 -= lg * lc1; x1 -= xg * lc1;
                }
            }
            // B fragments of the panel solve: B[slot t][n = g] = (L^-T)[2 t + step][g]
            double bf0, bf1;
            {
                const int src0 = (2 * t) * 4 + (g >> 1), src1 = (2 * t + 1) * 4 + (g >> 1);
                const double v00 = shfl_d(x0, src0), v01 = shfl_d(x1, src0);
                const double v10 = shfl_d(x0, src1), v11 = shfl_d(x1, src1);
                bf0 = (g & 1) ? v01 : v00;
                bf1 = (g & 1) ? v11 : v10;
            }
            // panel: L[I][0] = S[I][0] L_00^-T   (and the appended row)
#pragma unroll
            for (int I = 1; I < TC; ++I) {
                if (I < rem) {
                    const int idx = I * (I + 1) / 2;
                    double n0 = 0.0, n1 = 0.0;
                    dense::dmma(n0, n1, S[idx][0], bf0);
                    dense::dmma(n0, n1, S[idx][1], bf1);
                    S[idx][0] = n0;
                    S[idx][1] = n1;
                }
            }
            {
                double n0 = 0.0, n1 = 0.0;
                dense::dmma(n0, n1, Ga[0][0], bf0);
                dense::dmma(n0, n1, Ga[0][1], bf1);
                Ga[0][0] = n0;
                Ga[0][1] = n1;
            }
            // trailing update: S[I][J] -= L[I][0] L[J][0]^T
#pragma unroll
            for (int I = 1; I < TC; ++I) {
                if (I < rem) {
                    const double na0 = -S[I * (I + 1) / 2][0], na1 = -S[I * (I + 1) / 2][1];
#pragma unroll
                    for (int J = 1; J <= I; ++J) {
                        const int idx = I * (I + 1) / 2 + J;
                        dense::dmma(S[idx][0], S[idx][1], na0, S[J * (J + 1) / 2][0]);
                        dense::dmma(S[idx][0], S[idx][1], na1, S[J * (J + 1) / 2][1]);
                    }
                }
            }
            {
                const double na0 = -Ga[0][0], na1 = -Ga[0][1];
#pragma unroll
                for (int J = 1; J < TC; ++J) {
                    if (J < rem) {
                        dense::dmma(Ga[J][0], Ga[J][1], na0, S[J * (J + 1) / 2][0]);
                        dense::dmma(Ga[J][0], Ga[J][1], na1, S[J * (J + 1) / 2][1]);
                    }
                }
            }
            // the finished tile of the appended rows (y and V entries of this block column) enters their Gram matrix
            dense::dmma(gr0, gr1, Ga[0][0], Ga[0][0]);
            dense::dmma(gr0, gr1, Ga[0][1], Ga[0][1]);
            // shift the trailing matrix up-left by one tile
#pragma unroll
            for (int I = 1; I < TC; ++I) {
#pragma unroll
                for (int J = 1; J <= I; ++J) {
                    S[(I - 1) * I / 2 + (J - 1)][0] = S[I * (I + 1) / 2 + J][0];
                    S[(I - 1) * I / 2 + (J - 1)][1] = S[I * (I + 1) / 2 + J][1];
                }
                Ga[I - 1][0] = Ga[I][0];
                Ga[I - 1][1] = Ga[I][1];
            }
        }
        // Gram tile element (m, n) sits in lane 4 m + (n >> 1), register n & 1:  yy = (0,0), b_q = (q,0), M_qp = (q,p)
        const double yy = shfl_d(gr0, 0);
        double bq[3], Mq[6];
        bq[0] = shfl_d(gr0, 4); bq[1] = shfl_d(gr0, 8); bq[2] = shfl_d(gr0, 12);
        Mq[0] = shfl_d(gr1, 4);                                    // (1,1)
        Mq[1] = shfl_d(gr0, 5);                                    // (1,2)
        Mq[2] = shfl_d(gr1, 5);                                    // (1,3)
        Mq[3] = shfl_d(gr0, 9);                                    // (2,2)
        Mq[4] = shfl_d(gr1, 9);                                    // (2,3)
        Mq[5] = shfl_d(gr1, 13);                                   // (3,3)
        double gam = yy;
        {   // Cholesky solve on the leading rank x rank block of V V^T
            double l00 = 0, l10 = 0, l11 = 0, l20 = 0, l21 = 0, l22 = 0, z0 = 0, z1 = 0, z2 = 0;
            if (rank >= 1 && Mq[0] > 0.0) {
                l00 = sqrt(Mq[0]);
                z0 = bq[0] / l00;
                if (rank >= 2) {
                    l10 = Mq[1] / l00;
                    const double d1 = Mq[3] - l10 * l10;
                    if (d1 > 0.0) {
                        l11 = sqrt(d1);
                        z1 = (bq[1] - l10 * z0) / l11;
                        if (rank >= 3) {
                            l20 = Mq[2] / l00;
                            l21 = (Mq[4] - l20 * l10) / l11;
                            const double d2 = Mq[5] - l20 * l20 - l21 * l21;
                            if (d2 > 0.0) {
                                l22 = sqrt(d2);
                                z2 = (bq[2] - l20 * z0 - l21 * z1) / l22;
                            }
                        }
                    }
                }
            }
            gam = yy - (z0 * z0 + z1 * z1 + z2 * z2);
        }
        if (lane == 0) {
            const double crit = (dofv < a.chi2_len) ? a.chi2_table[dofv] : nan("");
            const bool pass = gam <= crit;
            a.feat_status[fj] = (unsigned char)((pass ? MSCKF_FEAT_ACCEPTED : MSCKF_FEAT_GATED_OUT) | (code & MSCKF_FEAT_LOST_FLAG));
            a.gamma[fj] = gam;
            a.dof[fj] = dofv;
            a.fac_rank[fj] = rank;
        }
    }
}

size_t front_warp_smem(const msckf_dims& d) {
    return ((size_t)(d.N * (d.N + 1) / 2) * 36 + (size_t)d.N * 12) * sizeof(double) + FW_WARPS * sizeof(WarpScratch);
}

}  // namespace

bool front_warp_supported(const msckf_dims& d) { return d.M <= FW_MAX_OBS && front_warp_smem(d) <= 227 * 1024; }

cudaError_t launch_front_warp(const FrontArgs& a, cudaStream_t s) {
    const size_t smem = front_warp_smem(a.dims);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(front_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    // one CTA per filter when the batch fills the machine; otherwise split the features so that ~2 waves of CTAs exist
    int chunks = 1;
    if (a.dims.B < 2 * 148) chunks = (2 * 148 + a.dims.B - 1) / a.dims.B;
    int fpc = (a.dims.F + chunks - 1) / chunks;
    if (fpc < 2 * FW_WARPS) fpc = 2 * FW_WARPS;
    chunks = (a.dims.F + fpc - 1) / fpc;
    dim3 grid(chunks, a.dims.B);
    front_warp_kernel<<<grid, FW_THREADS, smem, s>>>(a, fpc);
    msckf_count_launch(1);
    return cudaGetLastError();
}
