// Lane-per-observation geometry of one feature track for the warp-per-feature front-end kernels
// (front_split.cu: assembly + factorisation as two kernels):
//   selection + closed-form triangulation     reference src/msckf/MSCKF.py:458-495, src/utils/geometry.py:274-303
//   residual + Jacobian blocks                MSCKF.py:497-548, src/msckf/Camera.py:54-67
//   orthonormal basis Q1 of range(H_f), rank by SciPy's null_space rule, r~ = (I - Q1 Q1^T) r   MSCKF.py:554-559
#pragma once
#include "msckf_common.cuh"

namespace fgeom {

constexpr int FW_MAX_OBS = 32;   // one lane per observation

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// 1/sqrt(x) for x > 0 (0 for x <= 0): hardware seed (MUFU.RSQ64H, ~2^-22) + one third-order step, no slow-path branch
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // one cubically convergent step: e = 1 - x y^2, y <- y (1 + e/2 + 3 e^2/8); seed error 2^-22 -> ~2^-64
    const double e = fma(-(x * y), y, 1.0);
    y = fma(y * e, fma(0.375, e, 0.5), y);
    return x > 0.0 ? y : 0.0;
}

// 1/x for finite non-zero x: hardware seed (MUFU.RCP64H) + two Newton steps, no slow-path branch (x = 0 gives NaN where
// the IEEE division gives inf; every caller's downstream comparison treats both as "not valid")
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

// same without the x <= 0 guard: a non-positive (or NaN) pivot of a matrix that is positive definite by construction gives
// NaN, which propagates to gamma and fails the gate instead of silently zeroing a column
__device__ __forceinline__ double fast_rsqrt_spd(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}

__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// All-reduce of EIGHT values at once by recursive halving: every exchange step halves the number of values a lane still
// carries (4 + 2 + 1 + 1 + 1 = 9 double shuffles), then each total is fetched from the lane that owns it (8 shuffles):
// 17 instead of 40 trips through the shuffle / shared-memory queue, which is what the geometry waits for when the pair
// loops of the other warps keep that queue full.
__device__ __forceinline__ void warp_allreduce8(double (&v)[8], const int lane) {
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
    double a[4], b[2];
#pragma unroll
    for (int q = 0; q < 4; ++q) a[q] = (h16 ? v[4 + q] : v[q]) + __shfl_xor_sync(0xffffffffu, h16 ? v[q] : v[4 + q], 16);
#pragma unroll
    for (int q = 0; q < 2; ++q) b[q] = (h8 ? a[2 + q] : a[q]) + __shfl_xor_sync(0xffffffffu, h8 ? a[q] : a[2 + q], 8);
    double c = (h4 ? b[1] : b[0]) + __shfl_xor_sync(0xffffffffu, h4 ? b[0] : b[1], 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __shfl_sync(0xffffffffu, c, 4 * i);
}

// cyclic Jacobi eigen-decomposition of a symmetric 3x3 (a = [xx, xy, xz, yy, yz, zz]); V columns = eigenvectors
__device__ inline void sym3_eig_w(const double* a, double* lam, double* V) {
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
    const double i0 = fast_rcp(d0);
    const double l10 = a[1] * i0, l20 = a[2] * i0;
    const double d1 = a[3] - l10 * a[1];
    if (!(d1 > 0.0)) return false;
    const double i1 = fast_rcp(d1);
    const double e21 = a[4] - l20 * a[1];
    const double l21 = e21 * i1;
    const double d2 = a[5] - l20 * a[2] - l21 * e21;
    const double tr = a[0] + a[3] + a[5];
    if (!(d2 > 0.0) || !(d0 * d1 * d2 > 1e-9 * tr * tr * tr)) return false;
    const double z0 = y[0], z1 = y[1] - l10 * z0, z2 = y[2] - l20 * z0 - l21 * z1;
    p[2] = z2 * fast_rcp(d2);
    p[1] = z1 * i1 - l21 * p[2];
    p[0] = z0 * i0 - l10 * p[1] - l20 * p[2];
    return true;
}

// singular values of a 3x3 matrix G (row-major, destroyed) by one-sided Jacobi
__device__ inline void svals3_w(double* G, double* sv) {
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

struct ObsRegs {
    double u, v, dx, dy, dz, cf;
    int slot;
    // the track's own scalars, fetched with the observations so that their latency is hidden behind the previous
    // feature: fd = idp_rho (lane 0), idp_m (lanes 1-3), idp_base (lanes 4-6); fi = lost_cnt, tracked_cnt, idp_anchor_slot
    double fd;
    int fi;
};

// All loads are unconditional (clamped indices, no register written on two paths): the prefetch of the next feature then
// costs no scoreboard wait until the values are used.  Lanes beyond the track length get their neutral values from
// mask_obs() when the feature is processed.
__device__ __forceinline__ void load_obs(ObsRegs& o, const FrontArgs& a, size_t fj, int M, int lane) {
    const int Mcap = a.dims.M;
    const double* pd = (lane == 0 || lane >= 7) ? a.idp_rho + fj : (lane < 4 ? a.idp_m + fj * 3 + (lane - 1) : a.idp_base + fj * 3 + (lane - 4));
    const int* pi = (lane == 1) ? a.tracked_cnt + fj : (lane == 2 ? a.idp_anchor_slot + fj : a.lost_cnt + fj);
    o.fd = *pd;
    o.fi = *pi;
    const size_t e = fj * Mcap + (lane < M ? lane : 0);
    const double2 uv = *reinterpret_cast<const double2*>(a.obs_uv + e * 2);
    o.u = uv.x; o.v = uv.y;
    const double* d = a.obs_dir + e * 3;
    o.dx = d[0]; o.dy = d[1]; o.dz = d[2];
    o.cf = a.obs_conf[e];
    o.slot = a.obs_slot[e];
}

__device__ __forceinline__ void mask_obs(ObsRegs& o, int M, int lane) {
    if (lane >= M) { o.u = o.v = 0.0; o.dx = o.dy = 0.0; o.dz = 1.0; o.cf = 0.0; o.slot = 0; }
}

struct GeoConsts {
    double s2, g0, g1, g2, cos_thr;
    bool cos_ok;
};

__device__ __forceinline__ GeoConsts make_geo_consts(const FrontArgs& a) {
    GeoConsts c;
    c.s2 = a.par.sigma_image * a.par.sigma_image;
    c.g0 = a.par.gravity[0]; c.g1 = a.par.gravity[1]; c.g2 = a.par.gravity[2];
    c.cos_ok = a.par.min_parallax_deg >= 0.0 && a.par.min_parallax_deg < 180.0;   // (0 degrees: cos = 1, same guard band)
    c.cos_thr = c.cos_ok ? cos(a.par.min_parallax_deg * (3.14159265358979323846 / 180.0)) : 0.0;
    return c;
}

// Rows 3-7 for the feature fj of M observations held one per lane in `o` (sR / sT: the staged window poses).
// Returns the selection code (MSCKF_FEAT_*); when it is MSCKF_FEAT_ACCEPTED (| LOST_FLAG | TRIANGULATED) the outputs hold the
// observability-projected 2x6 block Hx, the lane's two rows of Q1 (zero columns beyond `rank`) and r~.  Re-triangulated
// (rho, m) are written back by lane 0 (MSCKF.py:485-488).
__device__ __forceinline__ int feature_geometry(const FrontArgs& a, const GeoConsts& gc, size_t fj, int M, int lane,
                                                const ObsRegs& o, const double* sR, const double* sT,
                                                double (&Hx)[2][6], double (&Hf)[2][3], double& rt0, double& rt1,
                                                int& rank) {
    const double g0 = gc.g0, g1 = gc.g1, g2 = gc.g2, cos_thr = gc.cos_thr;
    const bool cos_ok = gc.cos_ok;
    const int n2 = 2 * M;
    const bool live = lane < M;
    rank = 0; rt0 = 0.0; rt1 = 0.0;
    // ---- selection (MSCKF.py:463-479) ----
    int code;
    {
        const bool lost = __shfl_sync(0xffffffffu, o.fi, 0) >= a.par.min_lost;
        if (lost && __shfl_sync(0xffffffffu, o.fi, 1) < a.par.min_tracked) {
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
    if ((code & 0xF) != MSCKF_FEAT_ACCEPTED) return code;

    const double* R = sR + 9 * o.slot;
    const double* tt = sT + 3 * o.slot;

    // ---- triangulation: X = sum c (I - d d^T), y = sum c (I - d d^T) t_slot, p = pinv(X) y (geometry.py:286-299) ----
    double rho, m0, m1, m2, b0, b1, b2;
    {
        double term[9];
        {
            const double inn = fast_rsqrt(o.dx * o.dx + o.dy * o.dy + o.dz * o.dz);
            const double ux = o.dx * inn, uy = o.dy * inn, uz = o.dz * inn;
            const double c = live ? o.cf : 0.0;
            term[0] = c * (1.0 - ux * ux); term[1] = c * (0.0 - ux * uy); term[2] = c * (0.0 - ux * uz);
            term[3] = c * (1.0 - uy * uy); term[4] = c * (0.0 - uy * uz); term[5] = c * (1.0 - uz * uz);
            term[6] = term[0] * tt[0] + term[1] * tt[1] + term[2] * tt[2];
            term[7] = term[1] * tt[0] + term[3] * tt[1] + term[4] * tt[2];
            term[8] = term[2] * tt[0] + term[4] * tt[1] + term[5] * tt[2];
        }
        {
            double t8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t8[k] = live ? term[k] : 0.0;
            const double t9 = warp_sum_all(live ? term[8] : 0.0);
            warp_allreduce8(t8, lane);
#pragma unroll
            for (int k = 0; k < 8; ++k) term[k] = t8[k];
            term[8] = t9;
        }
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
        rho = shfl_d(o.fd, 0);
        m0 = shfl_d(o.fd, 1); m1 = shfl_d(o.fd, 2); m2 = shfl_d(o.fd, 3);
        if (cz > 0.0) {
            const double* K = a.par.K;
            const double iw = K[6] * cx + K[7] * cy + K[8] * cz;
            const double iiw = fast_rcp(iw);
            const double iu = (K[0] * cx + K[1] * cy + K[2] * cz) * iiw;
            const double iv = (K[3] * cx + K[4] * cy + K[5] * cz) * iiw;
            if (!(iu < 0.0 || iu >= (double)a.par.width || iv < 0.0 || iv >= (double)a.par.height)) {
                const double* Ki = a.par.Kinv;
                const double vx = Ki[0] * iu + Ki[1] * iv + Ki[2];
                const double vy = Ki[3] * iu + Ki[4] * iv + Ki[5];
                const double vz = Ki[6] * iu + Ki[7] * iv + Ki[8];
                const double wx = R0[0] * vx + R0[1] * vy + R0[2] * vz;
                const double wy = R0[3] * vx + R0[4] * vy + R0[5] * vz;
                const double wz = R0[6] * vx + R0[7] * vy + R0[8] * vz;
                rho = 1.0 / cz;                      // InverseDepthPoint.update, geometry.py:61-71
                const double iwn = fast_rsqrt(wx * wx + wy * wy + wz * wz);
                m0 = wx * iwn; m1 = wy * iwn; m2 = wz * iwn;
                code |= MSCKF_FEAT_TRIANGULATED;
                if (lane == 0) {
                    a.idp_rho[fj] = rho;
                    a.idp_m[fj * 3] = m0; a.idp_m[fj * 3 + 1] = m1; a.idp_m[fj * 3 + 2] = m2;
                }
            }
        }
        const int as = __shfl_sync(0xffffffffu, o.fi, 2);
        const double e0 = shfl_d(o.fd, 4), e1 = shfl_d(o.fd, 5), e2 = shfl_d(o.fd, 6);
        if (as >= 0) { b0 = sT[3 * as]; b1 = sT[3 * as + 1]; b2 = sT[3 * as + 2]; }
        else { b0 = e0; b1 = e1; b2 = e2; }
    }


    double res[2];
    // ---- residual and Jacobian blocks of this lane's observation (MSCKF.py:505-544, Camera.py:54-67) ----
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
        // (one reciprocal per denominator: the quotients differ from the reference's divisions by an ulp, far inside 1e-9)
        const double izw = fast_rcp(zw), if2 = fast_rcp(f2);
        res[0] = zx * izw - f0 * if2;
        res[1] = zy * izw - f1 * if2;
        const double J00 = if2, J02 = -f0 * (if2 * if2), J11 = if2, J12 = -f1 * (if2 * if2);
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
            const double iden = fast_rcp(den);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                double au = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) au += Hx[r][k] * u6[k];
#pragma unroll
                for (int k = 0; k < 6; ++k) Hx[r][k] = Hx[r][k] - (au * u6[k]) * iden;
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

    // ---- orthonormal basis of range(H_f) and r~ = (I - Q1 Q1^T) r ----
    // Fast path (well-conditioned H_f, the normal case): Cholesky-QR applied twice.  Two all-reduce rounds (the Gram matrix
    // of H_f; then the Gram matrix of the first-pass Q together with Q^T r) replace the nine sequential reductions of the
    // pivoted Gram-Schmidt below.  It is taken only when every Cholesky pivot of H_f^T H_f exceeds 1e-10 of its largest
    // diagonal entry (sigma_min / sigma_max > ~1e-5: rank 3 by SciPy's rule with nine orders of magnitude to spare, and the
    // second pass restores orthogonality to rounding); anything else goes through the Gram-Schmidt + SVD path.
    bool fast_done = false;
    if (n2 >= 3) {
        double G[8];
        G[0] = Hf[0][0] * Hf[0][0] + Hf[1][0] * Hf[1][0];
        G[1] = Hf[0][0] * Hf[0][1] + Hf[1][0] * Hf[1][1];
        G[2] = Hf[0][0] * Hf[0][2] + Hf[1][0] * Hf[1][2];
        G[3] = Hf[0][1] * Hf[0][1] + Hf[1][1] * Hf[1][1];
        G[4] = Hf[0][1] * Hf[0][2] + Hf[1][1] * Hf[1][2];
        G[5] = Hf[0][2] * Hf[0][2] + Hf[1][2] * Hf[1][2];
        G[6] = 0.0; G[7] = 0.0;
        warp_allreduce8(G, lane);
        const double gmax = fmax(G[0], fmax(G[3], G[5]));
        // G = R^T R, R upper triangular
        const double d0 = G[0];
        const double i0 = fast_rsqrt(d0);
        const double r01 = G[1] * i0, r02 = G[2] * i0;
        const double d1 = G[3] - r01 * r01;
        const double i1 = fast_rsqrt(d1);
        const double r12 = (G[4] - r01 * r02) * i1;
        const double d2 = G[5] - r02 * r02 - r12 * r12;
        const double i2 = fast_rsqrt(d2);
        if (d0 > 1e-10 * gmax && d1 > 1e-10 * gmax && d2 > 1e-10 * gmax && gmax < 1e300) {
            // first-pass Q = H_f R^-1 (row by row)
            double Q[2][3];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                Q[r][0] = Hf[r][0] * i0;
                Q[r][1] = (Hf[r][1] - Q[r][0] * r01) * i1;
                Q[r][2] = (Hf[r][2] - Q[r][0] * r02 - Q[r][1] * r12) * i2;
            }
            double S[8];
            S[0] = Q[0][0] * Q[0][0] + Q[1][0] * Q[1][0];
            S[1] = Q[0][0] * Q[0][1] + Q[1][0] * Q[1][1];
            S[2] = Q[0][0] * Q[0][2] + Q[1][0] * Q[1][2];
            S[3] = Q[0][1] * Q[0][1] + Q[1][1] * Q[1][1];
            S[4] = Q[0][1] * Q[0][2] + Q[1][1] * Q[1][2];
            S[5] = Q[0][2] * Q[0][2] + Q[1][2] * Q[1][2];
            S[6] = Q[0][0] * res[0] + Q[1][0] * res[1];
            S[7] = Q[0][1] * res[0] + Q[1][1] * res[1];
            const double b2 = warp_sum_all(Q[0][2] * res[0] + Q[1][2] * res[1]);
            warp_allreduce8(S, lane);
            // second pass: S = R2^T R2 (close to the identity), Q <- Q R2^-1, w = Q^T r = R2^-T (Q1^T r)
            const double j0 = fast_rsqrt(S[0]);
            const double s01 = S[1] * j0, s02 = S[2] * j0;
            const double j1 = fast_rsqrt(S[3] - s01 * s01);
            const double s12 = (S[4] - s01 * s02) * j1;
            const double j2 = fast_rsqrt(S[5] - s02 * s02 - s12 * s12);
            const double w0 = S[6] * j0;
            const double w1 = (S[7] - s01 * w0) * j1;
            const double w2 = (b2 - s02 * w0 - s12 * w1) * j2;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                Hf[r][0] = Q[r][0] * j0;
                Hf[r][1] = (Q[r][1] - Hf[r][0] * s01) * j1;
                Hf[r][2] = (Q[r][2] - Hf[r][0] * s02 - Hf[r][1] * s12) * j2;
            }
            rank = 3;
            rt0 = res[0] - (Hf[0][0] * w0 + Hf[0][1] * w1 + Hf[0][2] * w2);
            rt1 = res[1] - (Hf[1][0] * w0 + Hf[1][1] * w1 + Hf[1][2] * w2);
            fast_done = true;
        }
    }
    if (fast_done) return code;

    // ---- general path: column-pivoted Gram-Schmidt with re-orthogonalisation ----
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
            const double ink = fast_rsqrt(best);
            const double nk = best * ink;
            R3[k * 3 + k] = nk;
            if (!(nk > 0.0)) { stop = true; break; }
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

    // ---- r~ = r - Q1 Q1^T r ----
    {
        double w[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) w[c] = warp_sum_all(Hf[0][c] * res[0] + Hf[1][c] * res[1]);
        rt0 = res[0] - (Hf[0][0] * w[0] + Hf[0][1] * w[1] + Hf[0][2] * w[2]);
        rt1 = res[1] - (Hf[1][0] * w[0] + Hf[1][1] * w[1] + Hf[1][2] * w[2]);
    }
    return code;
}

}  // namespace fgeom
