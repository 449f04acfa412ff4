// Row 11 of SURVEY.md section 8a: MSCKF.process_imu (reference src/msckf/MSCKF.py:160-248) + IMU.integrate
// (src/msckf/IMU.py:78-100), n_samples per launch, one warp per filter.
//
// The 15x15 IMU block and the composed transition matrix stay in shared memory across all the samples between two camera
// frames; HBM is touched once: read P[0:15, :] (15 d doubles), write P[0:15, :] and its transpose P[:, 0:15].  The camera
// block of P is not touched: the reference's global (P + P^T)/2 (MSCKF.py:244) is the identity on it because every
// producer of P leaves it exactly symmetric.
#include "msckf_common.cuh"
#include "dense.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// One WARP per filter, no CTA barrier anywhere.  The round-2 profile of the previous kernel (a CTA per filter, nine serial
// 15 x 15 recurrences out of shared memory behind CTA barriers, dense powers of F) showed 4.6 % of the HBM roofline at 15 %
// of the warps active: it was a chain of dependent 15-term DFMA sums (0.64 ms per 4096 filters x 9 samples; this one 0.22).  Here
//   * the transition matrix of a sample is never built by dense powers of F: F dt has five non-zero 3 x 3 blocks
//     (MSCKF.py:179-192), so Phi = I + F + F^2/2 + F^3/6 is eight 3 x 3 blocks in closed form (below), which lane k
//     computes for sample k in registers, all samples of a chunk at once, together with the observability fix-up;
//   * the 15 x 15 recurrences run on the FP64 tensor pipe as 16 x 16 products of one warp (2 x 2 tiles of DMMA.8x8x4,
//     four dependent instructions per tile instead of fifteen dependent DFMAs per element), three per sample:
//     T = Phi P_II, Phi_tot <- Phi Phi_tot, P_II <- (T + dt Phi N) Phi^T with N = G Qc G^T = diag(s_g^2, s_bg^2, s_a^2,
//     s_ba^2, 0) (x) I_3 (MSCKF.py:237; R R^T = I), symmetrised like the reference's (P + P^T)/2;
//   * the composed Phi_tot is applied to the 15 x 6N strip straight from / to global memory in 8-column tiles, and the
//     strip's transpose is written in the same pass.
// A filter's working set is 12 KB of shared memory, so a dozen filters are in flight per SM.
namespace v3 {

constexpr int WPC = 4;            // warps (= filters) per CTA
constexpr int LD = 20;            // leading dimension of the 16 x 16 matrices (A- and B-fragment loads conflict-free)
constexpr int MAT = 16 * LD;
constexpr int CPH = 74;           // compact Phi record: 8 blocks of 9 + dt (+ pad)
constexpr int SIN = 14;           // per-sample record of phase A: dt, a[3], dR[9] (+ pad)
constexpr int CHUNK = 16;         // samples per chunk

using dense::dmma;

__device__ __forceinline__ void mul33(double* C, const double* A, const double* B) {   // C = A B (row-major 3 x 3)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}
__device__ __forceinline__ void skew9(double* S, const double* v) {
    S[0] = 0.0; S[1] = -v[2]; S[2] = v[1];
    S[3] = v[2]; S[4] = 0.0; S[5] = -v[0];
    S[6] = -v[1]; S[7] = v[0]; S[8] = 0.0;
}

// C (2 x 2 tiles, accumulator layout) += A B with A[r][k] = MA[r * LD + k] and B[k][n] = TB ? MB[n * LD + k] : MB[k * LD + n]
template <bool TB>
__device__ __forceinline__ void mm16(double (&C)[2][2][2], const double* MA, const double* MB, int g, int t) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const double a0 = MA[g * LD + 4 * ks + t], a1 = MA[(8 + g) * LD + 4 * ks + t];
        const double b0 = TB ? MB[g * LD + 4 * ks + t] : MB[(4 * ks + t) * LD + g];
        const double b1 = TB ? MB[(8 + g) * LD + 4 * ks + t] : MB[(4 * ks + t) * LD + 8 + g];
        dmma(C[0][0][0], C[0][0][1], a0, b0);
        dmma(C[0][1][0], C[0][1][1], a0, b1);
        dmma(C[1][0][0], C[1][0][1], a1, b0);
        dmma(C[1][1][0], C[1][1][1], a1, b1);
    }
}

__global__ void __launch_bounds__(WPC * 32, 3) propagate_kernel_v3(msckf_dims dims, int n_samples, const double* __restrict__ samples,
                                                                 double* imu_time, const double* __restrict__ Qc_diag, double g0,
                                                                 double g1, double g2, double* imu_state, double* imu_null,
                                                                 double* Pall, const int* __restrict__ n_cams, int chunk_cap) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * WPC + warp;
    if (b >= dims.B) return;                              // (whole warps; nothing below synchronises across warps)
    const int g = lane >> 2, t = lane & 3;
    const int ldp = dims.ldp;
    const int c = 6 * n_cams[b];
    double* P = Pall + (size_t)b * ldp * ldp;
    const double gv[3] = {g0, g1, g2};

    extern __shared__ double sm[];
    const int per_warp = 4 * MAT + chunk_cap * (SIN + CPH);
    double* sPhi = sm + (size_t)warp * per_warp;
    double* sP = sPhi + MAT;
    double* sTot = sP + MAT;
    double* sT = sTot + MAT;
    double* sIn = sT + MAT;                               // [chunk_cap][SIN]
    double* sC = sIn + chunk_cap * SIN;                   // [chunk_cap][CPH]

    // the 15 x 6N strip is needed last: ask for it now, so that phase E finds it in the L2 (a warp has at most sixteen
    // loads of its own in flight there, which leaves the HBM latency exposed)
    {
        const int lines = (c * 8 + 127) / 128 + 1;        // 128-byte lines per strip row (the strip starts 120 bytes into its row)
        for (int e = lane; e < 15 * lines; e += 32) {
            const int row = e / lines, ln = e - row * lines;
            const double* ptr = P + (size_t)row * ldp + 15 + 16 * ln;
            if (16 * ln < c) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
        }
    }
    // P_II (padded to 16 x 16 with zeros), Phi_tot = I
    for (int e = lane; e < 256; e += 32) {
        const int i = e >> 4, j = e & 15;
        sP[i * LD + j] = (i < 15 && j < 15) ? P[(size_t)i * ldp + j] : 0.0;
        sTot[i * LD + j] = (i == j) ? 1.0 : 0.0;
    }
    // running state, held redundantly by every lane
    // (the null state differs from the state only before the first sample of the launch: MSCKF.py:247-248 makes it the
    //  current state after every sample)
    double R[9], p[3], v[3];
    const double* st_in = imu_state + (size_t)b * 21;
    const double* null_in = imu_null ? imu_null + (size_t)b * 15 : st_in;
    {
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = st_in[e];
#pragma unroll
        for (int e = 0; e < 3; ++e) { p[e] = st_in[9 + e]; v[e] = st_in[12 + e]; }
    }
    double tcur = imu_time[b];
    const double qn0 = Qc_diag[(size_t)b * 4], qn1 = Qc_diag[(size_t)b * 4 + 1], qn2 = Qc_diag[(size_t)b * 4 + 2], qn3 = Qc_diag[(size_t)b * 4 + 3];
    __syncwarp();

    for (int k0 = 0; k0 < n_samples; k0 += chunk_cap) {
        const int nk = min(chunk_cap, n_samples - k0);
        // ---- A: lane k, sample k: dt, bias-corrected w / a, rotation increment (IMU.py:78-100) ----
        double dt = 0.0, w[3] = {0.0, 0.0, 0.0}, a[3] = {0.0, 0.0, 0.0};
        if (lane < nk) {
            const double* z = samples + ((size_t)b * n_samples + k0 + lane) * 7;
            const double tprev = (lane == 0) ? tcur : z[-7];
            dt = z[0] - tprev;
#pragma unroll
            for (int k = 0; k < 3; ++k) { w[k] = z[1 + k] - st_in[15 + k]; a[k] = z[4 + k] - st_in[18 + k]; }
            const double wn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
            const double theta = wn * dt;
            double dR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            if (theta > 0.0) {
                const double ax[3] = {w[0] / wn, w[1] / wn, w[2] / wn};
                double K[9], K2[9];
                skew9(K, ax);
                mul33(K2, K, K);
                const double sn_ = sin(theta), cs = 1.0 - cos(theta);
#pragma unroll
                for (int k = 0; k < 9; ++k) dR[k] = dR[k] + sn_ * K[k] + cs * K2[k];
            }
            double* q = sIn + lane * SIN;
            q[0] = dt;
#pragma unroll
            for (int k = 0; k < 3; ++k) q[1 + k] = a[k];
#pragma unroll
            for (int k = 0; k < 9; ++k) q[4 + k] = dR[k];
        }
        __syncwarp();
        // ---- B: the state recurrence, every lane the same; lane k keeps sample k's states ----
        double mR[9], mp[3], mv[3], mRn[9], mtn[3], mvn[3];   // post-integration state and null state of MY sample
#pragma unroll
        for (int e = 0; e < 9; ++e) mR[e] = mRn[e] = 0.0;
#pragma unroll
        for (int e = 0; e < 3; ++e) mp[e] = mv[e] = mtn[e] = mvn[e] = 0.0;
        if (k0 == 0 && lane == 0) {                       // the launch's first sample sees the null state it was given
#pragma unroll
            for (int e = 0; e < 9; ++e) mRn[e] = null_in[e];
#pragma unroll
            for (int e = 0; e < 3; ++e) { mtn[e] = null_in[9 + e]; mvn[e] = null_in[12 + e]; }
        }
        for (int k = 0; k < nk; ++k) {
            const double* q = sIn + k * SIN;
            const double dtk = q[0];
            const bool mine = lane == k;
            if (mine && (k0 + k) > 0) {                   // every later one the state before its own integration
#pragma unroll
                for (int e = 0; e < 9; ++e) mRn[e] = R[e];
#pragma unroll
                for (int e = 0; e < 3; ++e) { mtn[e] = p[e]; mvn[e] = v[e]; }
            }
            double Rn[9], aw[3];
            mul33(Rn, R, q + 4);
#pragma unroll
            for (int r = 0; r < 3; ++r) aw[r] = (R[r * 3] * q[1] + R[r * 3 + 1] * q[2] + R[r * 3 + 2] * q[3]) - gv[r];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                p[r] = p[r] + v[r] * dtk + 0.5 * aw[r] * (dtk * dtk);
                v[r] = v[r] + aw[r] * dtk;
            }
#pragma unroll
            for (int e = 0; e < 9; ++e) R[e] = Rn[e];
            if (mine) {
#pragma unroll
                for (int e = 0; e < 9; ++e) mR[e] = R[e];
#pragma unroll
                for (int e = 0; e < 3; ++e) { mp[e] = p[e]; mv[e] = v[e]; }
            }
        }
        tcur = samples[((size_t)b * n_samples + k0 + nk - 1) * 7];
        // ---- C: lane k, sample k: the eight non-trivial blocks of Phi in closed form.  With A = -[w]x dt, B = -R [a]x dt,
        //      C = -R dt (R the post-integration rotation, MSCKF.py:186-188) the series I + F + F^2/2 + F^3/6 is
        //        (th,th) = R Rn^T (the fix-up, MSCKF.py:221)   (th,bg) = -dt S          S = I + A/2 + A^2/6
        //        (v,th)  = B S  (+ fix-up)   (v,bg) = -dt W    (v,ba) = C               W = B/2 + B A/6
        //        (p,th)  = dt W (+ fix-up)   (p,bg) = -dt^2 B/6   (p,v) = dt I   (p,ba) = dt C/2
        if (lane < nk) {
            double A[9], A2[9], S[9], Bm[9], BA[9], Wm[9], sk[9], tmp[9];
            skew9(sk, w);
#pragma unroll
            for (int e = 0; e < 9; ++e) A[e] = (-sk[e]) * dt;
            mul33(A2, A, A);
#pragma unroll
            for (int e = 0; e < 9; ++e) S[e] = ((e % 4 == 0 ? 1.0 : 0.0) + 0.5 * A[e]) + (1.0 / 6.0) * A2[e];
            skew9(sk, a);
            mul33(tmp, mR, sk);
#pragma unroll
            for (int e = 0; e < 9; ++e) Bm[e] = (-tmp[e]) * dt;
            mul33(BA, Bm, A);
#pragma unroll
            for (int e = 0; e < 9; ++e) Wm[e] = 0.5 * Bm[e] + (1.0 / 6.0) * BA[e];
            double* q = sC + lane * CPH;
            // (th,th) = R Rn^T
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
                    q[r * 3 + cc] = mR[r * 3] * mRn[cc * 3] + mR[r * 3 + 1] * mRn[cc * 3 + 1] + mR[r * 3 + 2] * mRn[cc * 3 + 2];
#pragma unroll
            for (int e = 0; e < 9; ++e) q[9 + e] = (-dt) * S[e];                       // (th,bg)
            double Av[9], Ap[9];
            mul33(Av, Bm, S);                                                          // (v,th) before the fix-up
#pragma unroll
            for (int e = 0; e < 9; ++e) Ap[e] = dt * Wm[e];                            // (p,th) before the fix-up
            // observability fix-up (MSCKF.py:221-233); null state = state before this integration
            double u[3], sv[3];
#pragma unroll
            for (int e = 0; e < 3; ++e) u[e] = mRn[e * 3] * gv[0] + mRn[e * 3 + 1] * gv[1] + mRn[e * 3 + 2] * gv[2];
            const double uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
#pragma unroll
            for (int e = 0; e < 3; ++e) sv[e] = u[e] / uu;
            const double d1[3] = {mvn[0] - mv[0], mvn[1] - mv[1], mvn[2] - mv[2]};
            const double d2[3] = {dt * mvn[0] + mtn[0] - mp[0], dt * mvn[1] + mtn[1] - mp[1], dt * mvn[2] + mtn[2] - mp[2]};
            const double w1[3] = {d1[1] * gv[2] - d1[2] * gv[1], d1[2] * gv[0] - d1[0] * gv[2], d1[0] * gv[1] - d1[1] * gv[0]};
            const double w2[3] = {d2[1] * gv[2] - d2[2] * gv[1], d2[2] * gv[0] - d2[0] * gv[2], d2[0] * gv[1] - d2[1] * gv[0]};
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double ev = (Av[r * 3] * u[0] + Av[r * 3 + 1] * u[1] + Av[r * 3 + 2] * u[2]) - w1[r];
                const double ep = (Ap[r * 3] * u[0] + Ap[r * 3 + 1] * u[1] + Ap[r * 3 + 2] * u[2]) - w2[r];
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    q[18 + r * 3 + cc] = Av[r * 3 + cc] - ev * sv[cc];                 // (v,th)
                    q[45 + r * 3 + cc] = Ap[r * 3 + cc] - ep * sv[cc];                 // (p,th)
                }
            }
#pragma unroll
            for (int e = 0; e < 9; ++e) {
                q[27 + e] = (-dt) * Wm[e];                                             // (v,bg)
                q[36 + e] = (-mR[e]) * dt;                                             // (v,ba)
                q[54 + e] = (-(dt * dt) * (1.0 / 6.0)) * Bm[e];                        // (p,bg)
                q[63 + e] = (-(0.5 * dt * dt)) * mR[e];                                // (p,ba)
            }
            q[72] = dt;
        }
        __syncwarp();
        // ---- D: the 16 x 16 recurrences, sample by sample, on the tensor pipe ----
        for (int k = 0; k < nk; ++k) {
            const double* q = sC + k * CPH;
            const double dtk = q[72];
            // expand Phi (identity + blocks) into shared memory
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = lane + 32 * i;
                const int r = e >> 4, cc = e & 15;
                const int br = r / 3, bc = cc / 3, rr = r - 3 * br, cr = cc - 3 * bc;
                double val = (r == cc && r < 15) ? 1.0 : 0.0;
                int off = -1;
                if (br == 0 && bc == 0) off = 0;
                else if (br == 0 && bc == 1) off = 9;
                else if (br == 2 && bc == 0) off = 18;
                else if (br == 2 && bc == 1) off = 27;
                else if (br == 2 && bc == 3) off = 36;
                else if (br == 4 && bc == 0) off = 45;
                else if (br == 4 && bc == 1) off = 54;
                else if (br == 4 && bc == 3) off = 63;
                if (off >= 0) val = q[off + rr * 3 + cr];
                if (br == 4 && bc == 2) val = (rr == cr) ? dtk : 0.0;
                if (r == 15 || cc == 15) val = 0.0;
                sPhi[r * LD + cc] = val;
            }
            __syncwarp();
            double cT[2][2][2], cU[2][2][2];
#pragma unroll
            for (int x = 0; x < 8; ++x) { cT[x >> 2][(x >> 1) & 1][x & 1] = 0.0; cU[x >> 2][(x >> 1) & 1][x & 1] = 0.0; }
            mm16<false>(cT, sPhi, sP, g, t);              // T = Phi P_II
            mm16<false>(cU, sPhi, sTot, g, t);            // Phi_tot <- Phi Phi_tot
            __syncwarp();
            // A2 = T + dt Phi N (N = G Qc G^T, diagonal) -> sT;  Phi_tot -> sTot
#pragma unroll
            for (int I = 0; I < 2; ++I)
#pragma unroll
                for (int J = 0; J < 2; ++J)
#pragma unroll
                    for (int x = 0; x < 2; ++x) {
                        const int r = 8 * I + g, cc = 8 * J + 2 * t + x;
                        const int bc = cc / 3;
                        const double nn = bc == 0 ? qn0 : (bc == 1 ? qn1 : (bc == 2 ? qn2 : (bc == 3 ? qn3 : 0.0)));
                        sT[r * LD + cc] = cT[I][J][x] + (dtk * nn) * sPhi[r * LD + cc];
                        sTot[r * LD + cc] = cU[I][J][x];
                    }
            __syncwarp();
            double cP[2][2][2];
#pragma unroll
            for (int x = 0; x < 8; ++x) cP[x >> 2][(x >> 1) & 1][x & 1] = 0.0;
            mm16<true>(cP, sT, sPhi, g, t);               // (T + dt Phi N) Phi^T = Phi P Phi^T + Q
            __syncwarp();
#pragma unroll
            for (int I = 0; I < 2; ++I)
#pragma unroll
                for (int J = 0; J < 2; ++J)
#pragma unroll
                    for (int x = 0; x < 2; ++x) sT[(8 * I + g) * LD + 8 * J + 2 * t + x] = cP[I][J][x];
            __syncwarp();
#pragma unroll
            for (int I = 0; I < 2; ++I)
#pragma unroll
                for (int J = 0; J < 2; ++J)
#pragma unroll
                    for (int x = 0; x < 2; ++x) {
                        const int r = 8 * I + g, cc = 8 * J + 2 * t + x;
                        sP[r * LD + cc] = (cP[I][J][x] + sT[cc * LD + r]) / 2;     // (P + P^T) / 2, MSCKF.py:244
                    }
            __syncwarp();
        }
    }

    // ---- E: P_IC <- Phi_tot P_IC in 8-column tiles straight from / to global memory; the transpose in the same pass ----
    {
        double at[2][4];                                  // A fragments of Phi_tot: rows 8 I + g, k = 4 ks + t
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) { at[0][ks] = sTot[g * LD + 4 * ks + t]; at[1][ks] = sTot[(8 + g) * LD + 4 * ks + t]; }
        const int ntile = (c + 7) >> 3;
        auto load_pair = [&](double (&bx)[2][4], int n0) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const int row = 4 * ks + t, col = 8 * (n0 + h) + g;
                    bx[h][ks] = (row < 15 && col < c) ? P[(size_t)row * ldp + 15 + col] : 0.0;
                }
        };
        // the loads of the next tile pair are issued BEFORE the stores of this one (the compiler cannot move them there: the
        // strip is read and written through the same pointer), so only the first pair's latency is exposed
        auto fold_pair = [&](const double (&bx)[2][4], int n0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    dmma(o[0][0], o[0][1], at[0][ks], bx[h][ks]);
                    dmma(o[1][0], o[1][1], at[1][ks], bx[h][ks]);
                }
#pragma unroll
                for (int I = 0; I < 2; ++I)
#pragma unroll
                    for (int x = 0; x < 2; ++x) {
                        const int row = 8 * I + g, col = 8 * (n0 + h) + 2 * t + x;
                        if (row < 15 && col < c) {
                            P[(size_t)row * ldp + 15 + col] = o[I][x];
                            P[(size_t)(15 + col) * ldp + row] = o[I][x];
                        }
                    }
            }
        };
        double ba[2][4], bb[2][4];
        load_pair(ba, 0);
        for (int n0 = 0; n0 < ntile; n0 += 4) {
            if (n0 + 2 < ntile) load_pair(bb, n0 + 2);
            fold_pair(ba, n0);
            if (n0 + 4 < ntile) load_pair(ba, n0 + 4);
            if (n0 + 2 < ntile) fold_pair(bb, n0 + 2);
        }
    }
    __syncwarp();
    for (int e = lane; e < 225; e += 32) {
        const int i = e / 15, j = e - 15 * i;
        P[(size_t)i * ldp + j] = sP[i * LD + j];
    }
    if (lane == 0) {
        double* st = imu_state + (size_t)b * 21;
#pragma unroll
        for (int e = 0; e < 9; ++e) st[e] = R[e];
#pragma unroll
        for (int e = 0; e < 3; ++e) { st[9 + e] = p[e]; st[12 + e] = v[e]; }
        if (imu_null) {
            double* sn = imu_null + (size_t)b * 15;
#pragma unroll
            for (int e = 0; e < 9; ++e) sn[e] = R[e];                      // null <- current (MSCKF.py:247-248)
#pragma unroll
            for (int e = 0; e < 3; ++e) { sn[9 + e] = p[e]; sn[12 + e] = v[e]; }
        }
        imu_time[b] = tcur;
    }
}

}  // namespace v3

}  // namespace

cudaError_t launch_propagate(const msckf_dims& d, int n_samples, const double* samples, double* imu_time,
                             const double* Qc_diag, const double* g, double* imu_state, double* imu_null, double* P,
                             const int* n_cams, cudaStream_t s) {
    const int cap = n_samples < v3::CHUNK ? (n_samples > 0 ? n_samples : 1) : v3::CHUNK;
    const size_t smem = (size_t)v3::WPC * (4 * v3::MAT + cap * (v3::SIN + v3::CPH)) * sizeof(double);
    static size_t configured3 = 0;
    if (smem > 48 * 1024 && smem > configured3) {
        cudaError_t e = cudaFuncSetAttribute(v3::propagate_kernel_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured3 = smem;
    }
    v3::propagate_kernel_v3<<<(d.B + v3::WPC - 1) / v3::WPC, v3::WPC * 32, smem, s>>>(
        d, n_samples, samples, imu_time, Qc_diag, g[0], g[1], g[2], imu_state, imu_null, P, n_cams, cap);
    msckf_count_launch(1);
    return cudaGetLastError();
}
