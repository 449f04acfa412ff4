// CTA-level FP64 dense linear algebra on the sm_100a tensor pipe (DMMA: mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4), used by
// the per-filter update kernels.  Operands live in global memory (L2-resident per-filter matrices of <= 400 x 400
// doubles); a 256-thread CTA walks 64 x 64 output tiles, stages 64 x 16 operand panels through shared memory with a
// register prefetch of the next panel (double-buffered), and every warp owns a 16 x 32 sub-tile = 2 x 4 DMMA tiles.
//
// Why DMMA and not DFMA: both paths peak at ~37 TFLOP/s on B200 (tools/probes/dmma_probe.cu, profiles/), but one DMMA
// replaces 8 warp-wide DFMAs, so the issue slots and the shared-memory bandwidth go to operand movement instead.
#pragma once
#include <cuda_runtime.h>

namespace dense {

constexpr int NT = 256;          // threads per CTA
constexpr int TM = 64, TN = 64, TK = 16;
constexpr int LDX = 20;          // leading dimension of a [64][16] (k-contiguous) panel: 20 mod 16 = 4 -> conflict-free
constexpr int LDK = 72;          // leading dimension of a [16][64] (k-major) panel: 72 mod 16 = 8 -> conflict-free
constexpr int PANEL = 1280;      // doubles per staged operand panel = max(64 * 20, 16 * 72)
constexpr int NB = 16;           // block size of the Cholesky / triangular-solve routines

struct Smem {
    double A[2][PANEL];
    double B[2][PANEL];
    double D[NB][NB + 1];
    double Di[NB][NB + 1];
};

enum Mask { MASK_NONE = 0, MASK_K_LE_X = 1, MASK_K_GE_X = 2 };

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Loads this thread's 4 elements of a 64 (x) by 16 (k) operand panel.  KC: element (x, k) at p[x * ld + k]; otherwise at
// p[k * ld + x].  Elements outside [0, X) x [klo, khi) or rejected by the triangular mask read as zero.
template <bool KC>
__device__ __forceinline__ void panel_load(double (&r)[4], const double* __restrict__ p, int ld, int x0, int X, int k0,
                                           int klo, int khi, int mask) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int e = tid + it * NT;
        int x, k;
        if (KC) { k = e & 15; x = e >> 4; } else { x = e & 63; k = e >> 6; }
        const int gx = x0 + x, gk = k0 + k;
        bool ok = gx < X && gk >= klo && gk < khi;
        if (mask == MASK_K_LE_X) ok = ok && gk <= gx;
        else if (mask == MASK_K_GE_X) ok = ok && gk >= gx;
        r[it] = ok ? (KC ? p[(size_t)gx * ld + gk] : p[(size_t)gk * ld + gx]) : 0.0;
    }
}

template <bool KC>
__device__ __forceinline__ void panel_store(double* s, const double (&r)[4]) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int e = tid + it * NT;
        if (KC) s[(e >> 4) * LDX + (e & 15)] = r[it];
        else s[(e >> 6) * LDK + (e & 63)] = r[it];
    }
}

template <bool KC>
__device__ __forceinline__ double panel_at(const double* s, int x, int k) {
    return KC ? s[x * LDX + k] : s[k * LDK + x];
}

// For every (i, j) of the M x N product (restricted to tiles on or below the diagonal when lower_only):
//     epi(i, j, sum_k A(i, k) B(k, j))
// AKC: A(i, k) = A[i * lda + k], else A[k * lda + i].   BKC: B(k, j) = B[j * ldb + k], else B[k * ldb + j].
// maskA applies to (x = i, k), maskB to (x = j, k); masked-out elements are zero and the k range of every tile is cut
// to the part where both operands can be non-zero.
template <bool AKC, bool BKC, class Epi>
__device__ __noinline__ void gemm(Smem& sm, int M, int N, int K, const double* __restrict__ A, int lda, const double* __restrict__ B,
                     int ldb, int maskA, int maskB, bool lower_only, Epi epi) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wr = warp >> 1, wc = warp & 1;
    for (int i0 = 0; i0 < M; i0 += TM) {
        for (int j0 = 0; j0 < N; j0 += TN) {
            if (lower_only && j0 > i0 + TM - 1) continue;
            int klo = 0, khi = K;
            if (maskA == MASK_K_LE_X) khi = min(khi, i0 + TM);
            else if (maskA == MASK_K_GE_X) klo = max(klo, i0);
            if (maskB == MASK_K_LE_X) khi = min(khi, j0 + TN);
            else if (maskB == MASK_K_GE_X) klo = max(klo, j0);
            const int kbeg = (klo / TK) * TK;
            double acc[2][4][2];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
            if (kbeg < khi) {
                double ra[4], rb[4];
                panel_load<AKC>(ra, A, lda, i0, M, kbeg, klo, khi, maskA);
                panel_load<BKC>(rb, B, ldb, j0, N, kbeg, klo, khi, maskB);
                panel_store<AKC>(sm.A[0], ra);
                panel_store<BKC>(sm.B[0], rb);
                __syncthreads();
                int buf = 0;
                for (int k0 = kbeg; k0 < khi; k0 += TK) {
                    const bool more = k0 + TK < khi;
                    if (more) {
                        panel_load<AKC>(ra, A, lda, i0, M, k0 + TK, klo, khi, maskA);
                        panel_load<BKC>(rb, B, ldb, j0, N, k0 + TK, klo, khi, maskB);
                    }
                    const double* sa = sm.A[buf];
                    const double* sb = sm.B[buf];
#pragma unroll
                    for (int kk = 0; kk < TK; kk += 4) {
                        double af[2], bf[4];
#pragma unroll
                        for (int a = 0; a < 2; ++a) af[a] = panel_at<AKC>(sa, wr * 16 + a * 8 + g, kk + t);
#pragma unroll
                        for (int b = 0; b < 4; ++b) bf[b] = panel_at<BKC>(sb, wc * 32 + b * 8 + g, kk + t);
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
                    }
                    if (more) {
                        panel_store<AKC>(sm.A[buf ^ 1], ra);
                        panel_store<BKC>(sm.B[buf ^ 1], rb);
                    }
                    __syncthreads();
                    buf ^= 1;
                }
            }
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int gi = i0 + wr * 16 + a * 8 + g;
                if (gi >= M) continue;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int gj = j0 + wc * 32 + b * 8 + 2 * t;
                    if (gj < N) epi(gi, gj, acc[a][b][0]);
                    if (gj + 1 < N) epi(gi, gj + 1, acc[a][b][1]);
                }
            }
        }
    }
    __syncthreads();
}

// In-place blocked right-looking Cholesky of the lower triangle of the n x n matrix S (leading dimension lds), factoring
// only the first ncf columns; the trailing updates cover all n rows and the columns below ncu.  With ncu = n, on return
// S[:, :ncf] holds the first ncf columns of L and S[ncf:, ncf:] holds the Schur complement.  A non-positive pivot zeroes its column (positive
// SEMI-definite input, e.g. a covariance with exactly correlated clones).  Linv (may be null) receives the inverse of
// every NB x NB diagonal block of L, [block][NB * NB] row-major.
static __device__ __noinline__ void cholesky(Smem& sm, double* S, int lds, int n, int ncf, int ncu, double* Linv) {
    const int tid = threadIdx.x;
    double(*D)[NB + 1] = sm.D;
    double(*Di)[NB + 1] = sm.Di;
    for (int kb = 0; kb < ncf; kb += NB) {
        const int nb = min(NB, ncf - kb);
        for (int e = tid; e < NB * NB; e += NT) {
            const int r = e / NB, c = e % NB;
            D[r][c] = (r < nb && c <= r) ? S[(size_t)(kb + r) * lds + kb + c] : ((r == c) ? 1.0 : 0.0);
        }
        __syncthreads();
        if (tid < 32) {   // unblocked factorisation of the diagonal block by one warp
            for (int k = 0; k < nb; ++k) {
                const double p = D[k][k];
                const double dk = p > 0.0 ? sqrt(p) : 0.0;
                const double idk = p > 0.0 ? 1.0 / dk : 0.0;
                __syncwarp();
                if (tid == 0) D[k][k] = dk;
                for (int r = k + 1 + tid; r < nb; r += 32) D[r][k] *= idk;
                __syncwarp();
                const int m = nb - k - 1;
                for (int e = tid; e < m * m; e += 32) {
                    const int r = k + 1 + e / m, c = k + 1 + e % m;
                    if (c <= r) D[r][c] -= D[r][k] * D[c][k];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (tid < NB) {   // inverse of the diagonal block: thread t solves L x = e_t (zero pivots -> zero rows)
            const int tcol = tid;
            double x[NB];
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                double v = (r == tcol) ? 1.0 : 0.0;
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c < r) v -= D[r][c] * x[c];
                const double dd = D[r][r];
                x[r] = dd > 0.0 ? v / dd : 0.0;
            }
#pragma unroll
            for (int r = 0; r < NB; ++r) Di[r][tcol] = (r >= tcol) ? x[r] : 0.0;
        }
        __syncthreads();
        for (int e = tid; e < NB * NB; e += NT) {
            const int r = e / NB, c = e % NB;
            if (Linv) Linv[(size_t)(kb / NB) * NB * NB + e] = Di[r][c];
            if (r < nb && c <= r) S[(size_t)(kb + r) * lds + kb + c] = D[r][c];
        }
        // panel below the diagonal block: S[i, kb:kb+nb] <- S[i, kb:kb+nb] L_kk^-T
        for (int i = kb + nb + tid; i < n; i += NT) {
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
        const int n2 = n - kb - nb, c2 = ncu - kb - nb;
        if (n2 > 0 && c2 > 0) {   // trailing update, lower triangle, columns < ncu: S22 -= L21 L21^T
            double* S22 = S + (size_t)(kb + nb) * lds + kb + nb;
            const double* L21 = S + (size_t)(kb + nb) * lds + kb;
            gemm<true, true>(sm, n2, c2, nb, L21, lds, L21, lds, MASK_NONE, MASK_NONE, true,
                             [=](int i, int j, double v) { if (j <= i) S22[(size_t)i * lds + j] -= v; });
        }
    }
    __syncthreads();
}

// X <- X C^-T in place: X is m x n (leading dimension ldx), C the n x n lower-triangular Cholesky factor (leading
// dimension ldc) whose diagonal-block inverses are in Linv (as written by cholesky()).
static __device__ __noinline__ void trsm_right_lt(Smem& sm, double* X, int ldx, int m, const double* C, int ldc, int n, const double* Linv) {
    const int tid = threadIdx.x;
    double(*Di)[NB + 1] = sm.Di;
    for (int kb = 0; kb < n; kb += NB) {
        const int nb = min(NB, n - kb);
        for (int e = tid; e < NB * NB; e += NT) Di[e / NB][e % NB] = Linv[(size_t)(kb / NB) * NB * NB + e];
        __syncthreads();
        for (int i = tid; i < m; i += NT) {
            double x[NB], y[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) x[c] = (c < nb) ? X[(size_t)i * ldx + kb + c] : 0.0;
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
                if (c < nb) X[(size_t)i * ldx + kb + c] = y[c];
        }
        __syncthreads();
        const int n2 = n - kb - nb;
        if (n2 > 0) {   // X[:, kb+nb:] -= X[:, kb:kb+nb] C[kb+nb:, kb:kb+nb]^T
            double* X2 = X + kb + nb;
            gemm<true, true>(sm, m, n2, nb, X + kb, ldx, C + (size_t)(kb + nb) * ldc + kb, ldc, MASK_NONE, MASK_NONE,
                             false, [=](int i, int j, double v) { X2[(size_t)i * ldx + j] -= v; });
        }
    }
    __syncthreads();
}

constexpr int WB = 64;   // block-column width of the left-looking drivers (one 64-wide tile column of gemm())

// Left-looking blocked Cholesky (same contract as cholesky() with ncu = n): every 64-wide block column is first updated
// with ALL previously factored columns in one long-K GEMM (pipelined, written once), then factored by the right-looking
// routine restricted to the block column.  Three sequential GEMM phases at n = 195 instead of twelve rank-16
// read-modify-write sweeps of the trailing matrix.
static __device__ __noinline__ void cholesky_ll(Smem& sm, double* S, int lds, int n, int ncf, double* Linv) {
    for (int kb = 0; kb < ncf; kb += WB) {
        const int w = min(WB, ncf - kb);
        double* Skk = S + (size_t)kb * lds + kb;
        if (kb > 0)
            gemm<true, true>(sm, n - kb, w, kb, S + (size_t)kb * lds, lds, S + (size_t)kb * lds, lds, MASK_NONE, MASK_NONE,
                             true, [=](int i, int j, double v) { if (j <= i) Skk[(size_t)i * lds + j] -= v; });
        cholesky(sm, Skk, lds, n - kb, w, w, Linv ? Linv + (size_t)(kb / NB) * NB * NB : nullptr);
    }
    if (n > ncf) {   // Schur complement of the unfactored trailing block
        double* S22 = S + (size_t)ncf * lds + ncf;
        const double* L21 = S + (size_t)ncf * lds;
        gemm<true, true>(sm, n - ncf, n - ncf, ncf, L21, lds, L21, lds, MASK_NONE, MASK_NONE, true,
                         [=](int i, int j, double v) { if (j <= i) S22[(size_t)i * lds + j] -= v; });
    }
}

// Left-looking X <- X C^-T (same contract as trsm_right_lt()).
static __device__ __noinline__ void trsm_right_lt_ll(Smem& sm, double* X, int ldx, int m, const double* C, int ldc, int n,
                                                     const double* Linv) {
    for (int kb = 0; kb < n; kb += WB) {
        const int w = min(WB, n - kb);
        double* Xk = X + kb;
        if (kb > 0)
            gemm<true, true>(sm, m, w, kb, X, ldx, C + (size_t)kb * ldc, ldc, MASK_NONE, MASK_NONE, false,
                             [=](int i, int j, double v) { Xk[(size_t)i * ldx + j] -= v; });
        trsm_right_lt(sm, Xk, ldx, m, C + (size_t)kb * ldc + kb, ldc, w, Linv + (size_t)(kb / NB) * NB * NB);
    }
}

// y <- C^-1 y for the lower-triangular Cholesky factor C (n x n), using the diagonal-block inverses.  y in global/shared.
static __device__ __noinline__ void trsv_lower(Smem& sm, const double* C, int ldc, int n, const double* Linv, double* y) {
    const int tid = threadIdx.x;
    double(*Di)[NB + 1] = sm.Di;
    double* tmp = &sm.D[0][0];
    for (int kb = 0; kb < n; kb += NB) {
        const int nb = min(NB, n - kb);
        for (int e = tid; e < NB * NB; e += NT) Di[e / NB][e % NB] = Linv[(size_t)(kb / NB) * NB * NB + e];
        __syncthreads();
        if (tid < NB) {
            double v = 0.0;
            for (int k = 0; k <= tid && k < nb; ++k) v += Di[tid][k] * y[kb + k];
            tmp[tid] = (tid < nb) ? v : 0.0;
        }
        __syncthreads();
        if (tid < nb) y[kb + tid] = tmp[tid];
        for (int i = kb + nb + tid; i < n; i += NT) {
            double v = 0.0;
            for (int k = 0; k < nb; ++k) v += C[(size_t)i * ldc + kb + k] * tmp[k];
            y[i] -= v;
        }
        __syncthreads();
    }
}

}  // namespace dense
