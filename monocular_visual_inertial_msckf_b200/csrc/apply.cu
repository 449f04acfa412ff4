// Rows 9b-10 of SURVEY.md section 8a: S = T_H P T_H^T + sigma^2 I, K = P T_H^T S^-1, delta_x = K r_n
// (reference src/msckf/MSCKF.py:604-607) and MSCKF.correct (611-661): Joseph-form covariance update, symmetrisation,
// Rodrigues + polar re-orthogonalisation of the IMU and every camera pose, additive corrections.
//
// One persistent CTA per filter (grid = a few CTAs per SM, looping over the batch).  The measurement matrix only
// touches the 6N camera columns (its 15 IMU columns are identically zero, MSCKF.py:538-540), so with C = camera index
// range, R = T_H (upper triangular, c x c):
//     Bf  = R P[C,:]                    (c x d)
//     S   = Bf[:,C] R^T + sigma^2 I     (c x c, lower)      -> blocked Cholesky S = L L^T, inverses of the diagonal blocks
//     Kt  = S^-1 Bf = K^T               (c x d)             -> two right-looking block triangular solves (in place)
//     dx  = Kt^T r_n
//     M1  = Kt^T R = K T_H              (d x c)
//     Y1  = P - M1 P[C,:]               (d x d)   = (I - K T_H) P
//     P+  = Y1 - Y1[:,C] M1^T + sigma^2 Kt^T Kt,   then (P+ + P+^T)/2
// All of it is FP64 GEMM-shaped work on <= 400 x 400 operands that live in L2 (per-CTA scratch in the caller's workspace);
// the CTA-level GEMM stages 64x16 / 16x64 tiles through shared memory and keeps a 4x4 accumulator tile per thread.
#include "msckf_common.cuh"

namespace {

constexpr int AT = 256;          // threads per CTA
constexpr int BM = 64, BN = 64, BK = 16;
constexpr int NB = 16;           // Cholesky / TRSM block size

enum KRange { KR_FULL = 0, KR_GE_I = 1, KR_GE_J = 2, KR_LE_J = 3 };

struct GemmSmem {
    double As[BK][BM + 4];
    double Bs[BK][BN + 4];
};

// Cout(i,j) = beta * Cin(i,j) + alpha * sum_k A(i,k) B(k,j);   A(i,k) = A[i*ars + k*acs], B(k,j) = B[k*brs + j*bcs],
// C row-major with leading dimension ldc.  lower_only: tiles strictly above the diagonal are skipped.
// krange lets callers skip the structurally-zero part of a triangular operand.
__device__ void cta_gemm(GemmSmem& sm, int M, int N, int K, const double* __restrict__ A, int ars, int acs,
                         const double* __restrict__ B, int brs, int bcs, const double* Cin, double* Cout, int ldc,
                         double alpha, double beta, int krange, bool lower_only) {
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    for (int i0 = 0; i0 < M; i0 += BM) {
        for (int j0 = 0; j0 < N; j0 += BN) {
            if (lower_only && j0 > i0 + BM - 1) continue;
            int klo = 0, khi = K;
            if (krange == KR_GE_I) klo = (i0 / BK) * BK;
            else if (krange == KR_GE_J) klo = (j0 / BK) * BK;
            else if (krange == KR_LE_J) khi = min(K, j0 + BN);
            double acc[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
            for (int k0 = klo; k0 < khi; k0 += BK) {
                // ---- stage tiles (coalesced along whichever index has unit stride) ----
#pragma unroll
                for (int it = 0; it < (BM * BK) / AT; ++it) {
                    const int e = tid + it * AT;
                    int i, k;
                    if (acs == 1) { k = e % BK; i = e / BK; } else { i = e % BM; k = e / BM; }
                    const int gi = i0 + i, gk = k0 + k;
                    sm.As[k][i] = (gi < M && gk < khi) ? A[(size_t)gi * ars + (size_t)gk * acs] : 0.0;
                }
#pragma unroll
                for (int it = 0; it < (BN * BK) / AT; ++it) {
                    const int e = tid + it * AT;
                    int j, k;
                    if (brs == 1) { k = e % BK; j = e / BK; } else { j = e % BN; k = e / BN; }
                    const int gj = j0 + j, gk = k0 + k;
                    sm.Bs[k][j] = (gj < N && gk < khi) ? B[(size_t)gk * brs + (size_t)gj * bcs] : 0.0;
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < BK; ++k) {
                    const double2 a01 = *reinterpret_cast<const double2*>(&sm.As[k][ty * 4]);
                    const double2 a23 = *reinterpret_cast<const double2*>(&sm.As[k][ty * 4 + 2]);
                    const double2 b01 = *reinterpret_cast<const double2*>(&sm.Bs[k][tx * 4]);
                    const double2 b23 = *reinterpret_cast<const double2*>(&sm.Bs[k][tx * 4 + 2]);
                    const double av[4] = {a01.x, a01.y, a23.x, a23.y};
                    const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[r][c] += av[r] * bv[c];
                }
                __syncthreads();
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int gi = i0 + ty * 4 + r;
                if (gi >= M) continue;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int gj = j0 + tx * 4 + c;
                    if (gj >= N) continue;
                    double v = alpha * acc[r][c];
                    if (beta != 0.0) v += beta * Cin[(size_t)gi * ldc + gj];
                    Cout[(size_t)gi * ldc + gj] = v;
                }
            }
        }
    }
    __syncthreads();
}

// In-place blocked Cholesky of the lower triangle of S (n x n, leading dimension lds); also writes the inverse of every
// NB x NB diagonal block of L to Linv[blk][NB*NB] (row-major lower triangular) for the triangular solves.
__device__ void cta_cholesky(GemmSmem& gs, double (*D)[NB + 1], double (*Di)[NB + 1], double* S, int lds, int n,
                             double* Linv) {
    const int tid = threadIdx.x;
    for (int kb = 0; kb < n; kb += NB) {
        const int nb = min(NB, n - kb);
        // diagonal block -> shared
        for (int e = tid; e < NB * NB; e += AT) {
            const int r = e / NB, c = e % NB;
            D[r][c] = (r < nb && c <= r) ? S[(size_t)(kb + r) * lds + kb + c] : ((r == c) ? 1.0 : 0.0);
        }
        __syncthreads();
        if (tid < 32) {   // unblocked right-looking Cholesky by one warp
            for (int k = 0; k < nb; ++k) {
                const double dk = sqrt(D[k][k]);
                __syncwarp();
                if (tid == 0) D[k][k] = dk;
                for (int r = k + 1 + tid; r < nb; r += 32) D[r][k] /= dk;
                __syncwarp();
                for (int e = tid; e < (nb - k - 1) * (nb - k - 1); e += 32) {
                    const int r = k + 1 + e / (nb - k - 1), c = k + 1 + e % (nb - k - 1);
                    if (c <= r) D[r][c] -= D[r][k] * D[c][k];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // inverse of the diagonal block: thread t solves L x = e_t
        if (tid < NB) {
            const int t = tid;
            double x[NB];
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                double v = (r == t) ? 1.0 : 0.0;
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c < r) v -= D[r][c] * x[c];
                x[r] = v / D[r][r];
            }
#pragma unroll
            for (int r = 0; r < NB; ++r) Di[r][t] = (r >= t) ? x[r] : 0.0;
        }
        __syncthreads();
        for (int e = tid; e < NB * NB; e += AT) {
            const int r = e / NB, c = e % NB;
            Linv[(size_t)(kb / NB) * NB * NB + e] = Di[r][c];
            if (r < nb && c <= r) S[(size_t)(kb + r) * lds + kb + c] = D[r][c];
        }
        // panel below: S[i, kb:kb+nb] <- S[i, kb:kb+nb] L_kk^-T     (row i times Linv^T)
        for (int i = kb + nb + tid; i < n; i += AT) {
            double x[NB], y[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) x[c] = (c < nb) ? S[(size_t)i * lds + kb + c] : 0.0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < NB; ++k)
                    if (k <= c) v += x[k] * Di[c][k];
                y[c] = v;
            }
#pragma unroll
            for (int c = 0; c < NB; ++c)
                if (c < nb) S[(size_t)i * lds + kb + c] = y[c];
        }
        __syncthreads();
        // trailing update (lower only): S[i,j] -= sum_k L[i,kb+k] L[j,kb+k]
        const int n2 = n - kb - nb;
        if (n2 > 0) {
            double* S22 = S + (size_t)(kb + nb) * lds + kb + nb;
            const double* Lp = S + (size_t)(kb + nb) * lds + kb;
            cta_gemm(gs, n2, n2, nb, Lp, lds, 1, Lp, 1, lds, S22, S22, lds, -1.0, 1.0, KR_FULL, true);
        }
    }
    __syncthreads();
}

// X <- L^-1 X  (forward, right-looking) then X <- L^-T X (backward), X is n x m with leading dimension ldx.
__device__ void cta_chol_solve(GemmSmem& gs, double (*Di)[NB + 1], const double* L, int ldl, const double* Linv, int n,
                               double* X, int ldx, int m) {
    const int tid = threadIdx.x;
    // forward
    for (int kb = 0; kb < n; kb += NB) {
        const int nb = min(NB, n - kb);
        for (int e = tid; e < NB * NB; e += AT) Di[e / NB][e % NB] = Linv[(size_t)(kb / NB) * NB * NB + e];
        __syncthreads();
        for (int j = tid; j < m; j += AT) {
            double x[NB], y[NB];
#pragma unroll
            for (int r = 0; r < NB; ++r) x[r] = (r < nb) ? X[(size_t)(kb + r) * ldx + j] : 0.0;
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < NB; ++k)
                    if (k <= r) v += Di[r][k] * x[k];
                y[r] = v;
            }
#pragma unroll
            for (int r = 0; r < NB; ++r)
                if (r < nb) X[(size_t)(kb + r) * ldx + j] = y[r];
        }
        __syncthreads();
        const int n2 = n - kb - nb;
        if (n2 > 0)   // X[kb+nb:, :] -= L[kb+nb:, kb:kb+nb] X[kb:kb+nb, :]
            cta_gemm(gs, n2, m, nb, L + (size_t)(kb + nb) * ldl + kb, ldl, 1, X + (size_t)kb * ldx, ldx, 1,
                     X + (size_t)(kb + nb) * ldx, X + (size_t)(kb + nb) * ldx, ldx, -1.0, 1.0, KR_FULL, false);
    }
    // backward
    const int nblk = (n + NB - 1) / NB;
    for (int bi = nblk - 1; bi >= 0; --bi) {
        const int kb = bi * NB;
        const int nb = min(NB, n - kb);
        for (int e = tid; e < NB * NB; e += AT) Di[e / NB][e % NB] = Linv[(size_t)bi * NB * NB + e];
        __syncthreads();
        for (int j = tid; j < m; j += AT) {
            double x[NB], y[NB];
#pragma unroll
            for (int r = 0; r < NB; ++r) x[r] = (r < nb) ? X[(size_t)(kb + r) * ldx + j] : 0.0;
#pragma unroll
            for (int r = 0; r < NB; ++r) {   // y = Linv^T x
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < NB; ++k)
                    if (k >= r) v += Di[k][r] * x[k];
                y[r] = v;
            }
#pragma unroll
            for (int r = 0; r < NB; ++r)
                if (r < nb) X[(size_t)(kb + r) * ldx + j] = y[r];
        }
        __syncthreads();
        if (kb > 0)   // X[0:kb, :] -= L[kb:kb+nb, 0:kb]^T X[kb:kb+nb, :]
            cta_gemm(gs, kb, m, nb, L + (size_t)kb * ldl, 1, ldl, X + (size_t)kb * ldx, ldx, 1, X, X, ldx, -1.0, 1.0,
                     KR_FULL, false);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(AT) apply_kernel(const ApplyArgs a) {
    __shared__ GemmSmem gs;
    __shared__ double D[NB][NB + 1];
    __shared__ double Di[NB][NB + 1];
    const int tid = threadIdx.x;
    const int ldp = a.dims.ldp, Ncap = a.dims.N, ccap = 6 * Ncap;
    double* scratch = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;
    double* Bf = scratch;                                   // [ccap][ldp]
    double* Sm = Bf + (size_t)ccap * ldp;                   // [ccap][ccap]
    double* M1 = Sm + (size_t)ccap * ccap;                  // [ldp][ccap]
    double* Y1 = M1 + (size_t)ldp * ccap;                   // [ldp][ldp]
    double* Linv = Y1 + (size_t)ldp * ldp;                  // [ceil(ccap/NB)][NB*NB]
    double* dx = Linv + (size_t)((ccap + NB - 1) / NB) * NB * NB;   // [ldp]
    const double s2 = a.sigma * a.sigma;

    for (int b = blockIdx.x; b < a.dims.B; b += gridDim.x) {
        double* P = a.P + (size_t)b * ldp * ldp;
        double* dxo = a.delta_x + (size_t)b * ldp;
        if (a.n_accepted[b] <= 0) {   // nothing survived gating: the reference returns before touching the state
            for (int i = tid; i < ldp; i += AT) dxo[i] = 0.0;
            continue;
        }
        const int nc = a.n_cams[b];
        const int c = 6 * nc, d = 15 + c;
        const double* R = a.T_H + (size_t)b * ccap * ccap;
        const double* rn = a.r_n + (size_t)b * ccap;
        const double* PC = P + (size_t)15 * ldp;            // rows of the camera block

        // Bf = R P[C,:]
        cta_gemm(gs, c, d, c, R, ccap, 1, PC, ldp, 1, nullptr, Bf, ldp, 1.0, 0.0, KR_GE_I, false);
        // S = Bf[:,C] R^T (+ sigma^2 I), lower triangle
        cta_gemm(gs, c, c, c, Bf + 15, ldp, 1, R, 1, ccap, nullptr, Sm, ccap, 1.0, 0.0, KR_GE_J, true);
        for (int i = tid; i < c; i += AT) Sm[(size_t)i * ccap + i] += s2;
        __syncthreads();
        cta_cholesky(gs, D, Di, Sm, ccap, c, Linv);
        // Kt = S^-1 Bf (in place)
        cta_chol_solve(gs, Di, Sm, ccap, Linv, c, Bf, ldp, d);
        // dx = Kt^T r_n
        for (int i = tid; i < ldp; i += AT) {
            double acc = 0.0;
            if (i < d)
                for (int k = 0; k < c; ++k) acc += Bf[(size_t)k * ldp + i] * rn[k];
            dx[i] = acc;
            dxo[i] = acc;
        }
        // M1 = Kt^T R   (d x c)
        cta_gemm(gs, d, c, c, Bf, 1, ldp, R, ccap, 1, nullptr, M1, ccap, 1.0, 0.0, KR_LE_J, false);
        // Y1 = P - M1 P[C,:]
        cta_gemm(gs, d, d, c, M1, ccap, 1, PC, ldp, 1, P, Y1, ldp, -1.0, 1.0, KR_FULL, false);
        // P = Y1 - Y1[:,C] M1^T
        cta_gemm(gs, d, d, c, Y1 + 15, ldp, 1, M1, 1, ccap, Y1, P, ldp, -1.0, 1.0, KR_FULL, false);
        // P += sigma^2 Kt^T Kt
        cta_gemm(gs, d, d, c, Bf, 1, ldp, Bf, ldp, 1, P, P, ldp, s2, 1.0, KR_FULL, false);
        // symmetrise
        for (int e = tid; e < d * d; e += AT) {
            const int i = e / d, j = e - i * d;
            if (j < i) {
                const double v = 0.5 * (P[(size_t)i * ldp + j] + P[(size_t)j * ldp + i]);
                P[(size_t)i * ldp + j] = v;
                P[(size_t)j * ldp + i] = v;
            }
        }
        // state correction (MSCKF.py:616-661)
        if (tid == 0) {
            double* st = a.imu_state + (size_t)b * MSCKF_IMU_STATE_DOUBLES;
            apply_rotation_correction(st, dx);
            for (int k = 0; k < 3; ++k) {
                st[9 + k] += dx[12 + k];    // position
                st[12 + k] += dx[6 + k];    // velocity
                st[15 + k] += dx[3 + k];    // gyroscope bias
                st[18 + k] += dx[9 + k];    // accelerometer bias
            }
        }
        for (int i = tid; i < nc; i += AT) {
            double* Rc = a.cam_R + ((size_t)b * Ncap + i) * 9;
            double* tc = a.cam_t + ((size_t)b * Ncap + i) * 3;
            apply_rotation_correction(Rc, dx + 15 + 6 * i);
            for (int k = 0; k < 3; ++k) tc[k] += dx[15 + 6 * i + 3 + k];
        }
        __syncthreads();
    }
}

}  // namespace

size_t apply_scratch_doubles(const msckf_dims& d) {
    const size_t ccap = 6 * (size_t)d.N, ldp = d.ldp;
    return ccap * ldp + ccap * ccap + ldp * ccap + ldp * ldp + ((ccap + NB - 1) / NB) * NB * NB + ldp + 64;
}

int apply_num_ctas(const msckf_dims& d) {
    const int cap = 148 * 3;
    return d.B < cap ? d.B : cap;
}

cudaError_t launch_apply(const ApplyArgs& a, int n_ctas, cudaStream_t s) {
    apply_kernel<<<n_ctas, AT, 0, s>>>(a);
    msckf_count_launch(1);
    return cudaGetLastError();
}
