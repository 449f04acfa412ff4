// Factorisation + gate (rows 7-8 of SURVEY.md section 8a: MSCKF.project_on_nullspace / gating_test, reference
// src/msckf/MSCKF.py:554-568) for tracks of MORE than 32 observations (up to 64: the stress shape of BASELINE.json has 60),
// where the padded K of a feature has up to 17 tile columns and no longer fits the register file of one warp
// (front_split.cu's kernel stops at 9).  One 256-thread CTA per pending feature, the lower tile triangle of K (<= 153
// tiles = 78 KB) resident in shared memory, XOR-swizzled so that every FP64 tensor-instruction fragment load is
// bank-conflict free; right-looking tile Cholesky, two tile columns per block step, all O(n^3) work on DMMA.8x8x4, the 8x8
// diagonal tiles factored by one warp with shuffles (the routine of update_smem.cu, here for 8 warps).  The appended rows
// [r~; Q1^T] ride along as ordinary rows of the last tile row; after the measurement columns are factored, the block
// under them holds minus the Gram matrix of [y; V], from which gamma follows (front_gate.cuh).
// The kernel that runs is the LEFT-looking one further down (front_chol_ll_kernel: four warps per feature, four features per
// SM, look-ahead); the right-looking one stays as the reference variant (CHOL_LL=0).
#include "msckf_common.cuh"
#include "dense.cuh"
#include "front_tiles.cuh"
#include "front_gate.cuh"

namespace {

using dense::dmma;
using namespace ftile;

constexpr int CT = 256;
constexpr int CW = CT / 32;
constexpr int TILE = 64;
constexpr int FLAG_KEEP = MSCKF_FEAT_LOST_FLAG | MSCKF_FEAT_TRIANGULATED;

__device__ __forceinline__ int swz(int r, int c) { return r * 8 + (c ^ ((r & 2) << 1)); }
__device__ __forceinline__ int toff(int I, int J) { return (((I * (I + 1)) >> 1) + J) * TILE; }
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }
__device__ __forceinline__ int tri_row_f(int e) {
    int r = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
    if (tri(r) > e) --r;
    else if (tri(r + 1) <= e) ++r;
    return r;
}

// One warp factors the 8x8 SPD tile held in accumulator layout (lane (g, t): D0 = S[g][2t], D1 = S[g][2t+1]; only the
// lower triangle is read): L (lower, zeros above the diagonal) and X = L^-T come back in the same layout.
__device__ __forceinline__ void factor_tile(double D0, double D1, double& L0, double& L1, double& x0, double& x1, int g, int t) {
    x0 = (g == 2 * t) ? 1.0 : 0.0;
    x1 = (g == 2 * t + 1) ? 1.0 : 0.0;
    L0 = 0.0;
    L1 = 0.0;
#pragma unroll
    for (int tk = 0; tk < 4; ++tk) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = 2 * tk + h;
            double& Dk = h ? D1 : D0;
            double& xk = h ? x1 : x0;
            double& Lk = h ? L1 : L0;
            const double piv = shfl_d(Dk, k * 4 + tk);
            const double rs = fast_rsqrt_spd(piv);
            const double colv = (g > k) ? Dk * rs : 0.0;
            const double xsel = xk * rs;
            if (t == tk) { xk = xsel; Lk = (g > k) ? colv : ((g == k) ? Dk * rs : 0.0); }
            const double lg = shfl_d(colv, g * 4 + tk);
            const double xg = shfl_d(xsel, g * 4 + tk);
            const double lc0 = shfl_d(colv, (2 * t) * 4 + tk);
            const double lc1 = shfl_d(colv, (2 * t + 1) * 4 + tk);
            D0 -= lg * lc0; x0 -= xg * lc0;
            D1 -= lg * lc1; x1 -= xg * lc1;
        }
    }
}

__device__ __forceinline__ void diag_factor(double D0, double D1, double* tile, double* Xs, int g, int t) {
    double L0, L1, x0, x1;
    factor_tile(D0, D1, L0, L1, x0, x1, g, t);
    st2(tile + swz(g, 2 * t), L0, L1);
    st2(Xs + swz(g, 2 * t), x0, x1);
}

// Right-looking tile Cholesky of the first ncf tile columns of the nt x nt lower tile triangle T (two tile columns per block
// step); the trailing updates cover every tile row, so T[ncf:, ncf:] ends up holding the Schur complement.
__device__ void chol_tiles(double* T, const int nt, const int ncf, double* Xs) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int k0 = 0; k0 < ncf; k0 += 2) {
        const bool two = k0 + 1 < ncf;
        if (warp == 0) {
            double* tile = T + toff(k0, k0);
            const double2 v = ld2(tile + swz(g, 2 * t));
            diag_factor(v.x, v.y, tile, Xs, g, t);
        }
        __syncthreads();
        {   // panel of tile column k0
            const double bf0 = Xs[swz(2 * t, g)], bf1 = Xs[swz(2 * t + 1, g)];
            for (int i = k0 + 1 + warp; i < nt; i += CW) {
                double* p = T + toff(i, k0) + swz(g, 2 * t);
                const double2 v = ld2(p);
                double n0 = 0.0, n1 = 0.0;
                dmma(n0, n1, v.x, bf0);
                dmma(n0, n1, v.y, bf1);
                st2(p, n0, n1);
            }
        }
        __syncthreads();
        if (two) {
            const int k1 = k0 + 1;
            if (warp == 0) {   // diagonal tile of column k1: rank-8 update from column k0, then factor
                double* tile = T + toff(k1, k1);
                double2 v = ld2(tile + swz(g, 2 * t));
                const double* Lk = T + toff(k1, k0);
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const double av = Lk[swz(g, 4 * s + t)];
                    dmma(v.x, v.y, -av, av);
                }
                diag_factor(v.x, v.y, tile, Xs, g, t);
            }
            __syncthreads();
            {   // tile column k1 below the diagonal: update from column k0 fused with the panel solve
                const double bf0 = Xs[swz(2 * t, g)], bf1 = Xs[swz(2 * t + 1, g)];
                const double* Lk = T + toff(k1, k0);
                const double b0 = Lk[swz(g, t)], b1 = Lk[swz(g, 4 + t)];
                for (int i = k1 + 1 + warp; i < nt; i += CW) {
                    double* p = T + toff(i, k1) + swz(g, 2 * t);
                    double2 v = ld2(p);
                    const double* Li = T + toff(i, k0);
                    dmma(v.x, v.y, -Li[swz(g, t)], b0);
                    dmma(v.x, v.y, -Li[swz(g, 4 + t)], b1);
                    double n0 = 0.0, n1 = 0.0;
                    dmma(n0, n1, v.x, bf0);
                    dmma(n0, n1, v.y, bf1);
                    st2(p, n0, n1);
                }
            }
            __syncthreads();
        }
        // trailing update by 2x2 super-tiles: T[i][j] -= sum_kc L[i][kc] L[j][kc]^T, kc in {k0, k1}
        const int j0 = k0 + (two ? 2 : 1);
        const int m = nt - j0;
        if (m > 0) {
            const int ns = (m + 1) >> 1;
            const int cnt = ns * (ns + 1) / 2;
            const int nk = two ? 2 : 1;
            for (int q = warp; q < cnt; q += CW) {
                const int si = tri_row_f(q), sj = q - tri(si);
                const int I0 = j0 + 2 * si, J0 = j0 + 2 * sj;
                const bool vi1 = I0 + 1 < nt, vj1 = (J0 + 1 < nt) && (J0 + 1 <= I0 + 1);
                const bool v01 = (J0 + 1 <= I0), v10 = vi1, v11 = vi1 && vj1;
                double2 c00 = ld2(T + toff(I0, J0) + swz(g, 2 * t));
                double2 c01 = v01 ? ld2(T + toff(I0, J0 + 1) + swz(g, 2 * t)) : make_double2(0.0, 0.0);
                double2 c10 = v10 ? ld2(T + toff(I0 + 1, J0) + swz(g, 2 * t)) : make_double2(0.0, 0.0);
                double2 c11 = v11 ? ld2(T + toff(I0 + 1, J0 + 1) + swz(g, 2 * t)) : make_double2(0.0, 0.0);
                for (int kk = 0; kk < nk; ++kk) {
                    const int kc = k0 + kk;
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const int fo = swz(g, 4 * s + t);
                        const double a0 = -T[toff(I0, kc) + fo];
                        const double a1 = vi1 ? -T[toff(I0 + 1, kc) + fo] : 0.0;
                        const double b0 = T[toff(J0, kc) + fo];
                        const double b1 = (J0 + 1 < nt) ? T[toff(J0 + 1, kc) + fo] : 0.0;
                        dmma(c00.x, c00.y, a0, b0);
                        if (v01) dmma(c01.x, c01.y, a0, b1);
                        if (v10) dmma(c10.x, c10.y, a1, b0);
                        if (v11) dmma(c11.x, c11.y, a1, b1);
                    }
                }
                st2(T + toff(I0, J0) + swz(g, 2 * t), c00.x, c00.y);
                if (v01) st2(T + toff(I0, J0 + 1) + swz(g, 2 * t), c01.x, c01.y);
                if (v10) st2(T + toff(I0 + 1, J0) + swz(g, 2 * t), c10.x, c10.y);
                if (v11) st2(T + toff(I0 + 1, J0 + 1) + swz(g, 2 * t), c11.x, c11.y);
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(CT, 3)
front_chol_smem_kernel(const double2* __restrict__ Kbuf, const int kt_d2, const int* __restrict__ pend_list,
                       const int* __restrict__ pend_count, const int* __restrict__ feat_len, const int* __restrict__ fac_rank,
                       const int* __restrict__ dof, const double* __restrict__ chi2_table, const int chi2_len,
                       unsigned char* __restrict__ feat_status, double* __restrict__ gamma, const size_t fbase) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int count = *pend_count;
    for (int q = blockIdx.x; q < count; q += gridDim.x) {
        const int fl = pend_list[q];
        const size_t fj = fbase + (size_t)fl;
        const int n2 = 2 * feat_len[fj];
        const int a4 = (n2 + 3) & ~3;
        const int Tn = (a4 + 4 + 7) >> 3;
        const int rr = a4 - 8 * (Tn - 1);
        const int ntile = Tn * (Tn + 1) / 2;
        double* T = sm;
        double* Xs = sm + (size_t)ntile * TILE;
        const double2* Kt = Kbuf + (size_t)fl * kt_d2;
        // global tiles are row-major inside the tile (double2 number 4 g + t = elements (g, 2 t), (g, 2 t + 1)): swizzle in
        for (int e = tid; e < ntile * 32; e += CT) {
            const double2 v = __ldcs(Kt + e);
            const int tile = e >> 5, l = e & 31;
            st2(T + (size_t)tile * TILE + swz(l >> 2, 2 * (l & 3)), v.x, v.y);
        }
        __syncthreads();
        chol_tiles(T, Tn, Tn - 1, Xs);
        if (warp == 0) {
            const double2 v = ld2(T + toff(Tn - 1, Tn - 1) + swz(g, 2 * t));
            const int rank = fac_rank[fj], dofv = dof[fj];
            const int st = feat_status[fj];
            const double gam = gamma_from_last_tile(v.x, v.y, g, t, rr, rank);
            if (lane == 0) {
                const double crit = (dofv < chi2_len) ? chi2_table[dofv] : nan("");
                const bool pass = gam <= crit;
                feat_status[fj] = (unsigned char)((pass ? MSCKF_FEAT_ACCEPTED : MSCKF_FEAT_GATED_OUT) | (st & FLAG_KEEP));
                gamma[fj] = gam;
            }
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// LEFT-looking variant: four warps per feature, five features per SM.  The right-looking kernel above keeps the whole lower
// triangle of a feature in shared memory (70 KB at 60 observations: three features per SM) and its profile is the sixteen
// serial in-tile factorisations of a feature (1400 clocks each, seven warps waiting): 0.23 of the FP64 peak.  Left-looking,
// block column k needs the finished tiles L[I][j] with j < k <= I only -- at most max_k k (T - k) = 64 tiles = 32 KB, in slots
// assigned by interval colouring (a tile is live from the column that creates it to the column of its row) -- plus the current
// column (8 KB), so five CTAs fit an SM and their serial phases interleave.  Column k of K is read from the tile buffer when
// it is needed (and only then: the trailing matrix is never held).
constexpr int LT = 128;           // threads per CTA
constexpr int LWARPS = LT / 32;
constexpr int LMAXT = 17;         // tile rows at 64 observations

__host__ __device__ inline int ll_slots(int T) {
    int m = 0;
    for (int k = 0; k < T; ++k) m = (k * (T - k) > m) ? k * (T - k) : m;
    return m;
}

__global__ void __launch_bounds__(LT, 5)
front_chol_ll_kernel(const double2* __restrict__ Kbuf, const int kt_d2, const int* __restrict__ pend_list,
                     const int* __restrict__ pend_count, const int* __restrict__ feat_len, const int* __restrict__ fac_rank,
                     const int* __restrict__ dof, const double* __restrict__ chi2_table, const int chi2_len,
                     unsigned char* __restrict__ feat_status, double* __restrict__ gamma, const size_t fbase, const int Tcap) {
    extern __shared__ double sm[];
    const int nslots = ll_slots(Tcap);
    double* Ls = sm;                                   // [nslots][64] finished tiles
    double* const Cbase = Ls + (size_t)nslots * TILE;  // [2][Tcap][64] block columns k and k + 1 (tile row k first)
    double* Xs = Cbase + (size_t)2 * Tcap * TILE;      // [64]
    __shared__ unsigned char s_slot[LMAXT][LMAXT];     // slot of tile (I, j), I > j, for THIS feature's tile count
    __shared__ int s_T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int count = *pend_count;
    if (tid == 0) s_T = -1;
    __syncthreads();
    for (int q = blockIdx.x; q < count; q += gridDim.x) {
        const int fl = pend_list[q];
        const size_t fj = fbase + (size_t)fl;
        const int n2 = 2 * feat_len[fj];
        const int a4 = (n2 + 3) & ~3;
        const int Tn = (a4 + 4 + 7) >> 3;
        const int rr = a4 - 8 * (Tn - 1);
        const double2* Kt = Kbuf + (size_t)fl * kt_d2;
        double* Cc = Cbase;
        if (s_T != Tn) {                               // (interval colouring of the tiles of a Tn-row triangle; CTA-uniform test)
            __syncthreads();
            if (tid == 0) {
                unsigned long long fr[2] = {~0ull, ~0ull};             // free slots (bit set = free), <= 128
                for (int c = 0; c < Tn; ++c) {
                    for (int j = 0; j < c; ++j) { const int sl = s_slot[c][j]; fr[sl >> 6] |= 1ull << (sl & 63); }
                    for (int I = c + 1; I < Tn; ++I) {
                        const int w = fr[0] ? 0 : 1;
                        const int b = __ffsll((long long)fr[w]) - 1;
                        fr[w] &= ~(1ull << b);
                        s_slot[I][c] = (unsigned char)(64 * w + b);
                    }
                }
                s_T = Tn;
            }
            __syncthreads();
        }
        // column k of K: tiles (I, k), I = k .. Tn - 1 (row-major inside the tile: double2 number 4 g + t), requested two
        // columns ahead and parked in registers: the read of the tile buffer is off the critical path
        constexpr int LPRE = (LMAXT * 32 + LT - 1) / LT;      // double2 per thread of one column
        double2 pre[LPRE];
        auto fetch_column = [&](int k) {
            const int nt = Tn - k;
#pragma unroll
            for (int u = 0; u < LPRE; ++u) {
                const int e = tid + u * LT;
                if (e < nt * 32) pre[u] = __ldcs(Kt + (size_t)(tri(k + (e >> 5)) + k) * 32 + (e & 31));
            }
        };
        auto park_column = [&](double* C, int k) {
            const int nt = Tn - k;
#pragma unroll
            for (int u = 0; u < LPRE; ++u) {
                const int e = tid + u * LT;
                if (e < nt * 32) {
                    const int l = e & 31;
                    st2(C + (size_t)(e >> 5) * TILE + swz(l >> 2, 2 * (l & 3)), pre[u].x, pre[u].y);
                }
            }
        };
        // C[I - kk] -= sum_{jlo <= j < jhi} L[I][j] L[kk][j]^T for the tile rows I = kk + w, kk + w + nw, ... of column kk
        auto update_tiles = [&](double* C, int kk, int jlo, int jhi, int w, int nw) {
            for (int I = kk + w; I < Tn; I += nw) {
                double* p = C + (size_t)(I - kk) * TILE + swz(g, 2 * t);
                double2 c = ld2(p);
                double2 c2 = make_double2(0.0, 0.0);           // second accumulator: two independent DMMA chains
                int j = jlo;
                for (; j + 1 < jhi; j += 2) {
                    const double* Aa = Ls + (size_t)s_slot[I][j] * TILE;
                    const double* Ba = Ls + (size_t)s_slot[kk][j] * TILE;
                    const double* Ab = Ls + (size_t)s_slot[I][j + 1] * TILE;
                    const double* Bb = Ls + (size_t)s_slot[kk][j + 1] * TILE;
#pragma unroll
                    for (int sx = 0; sx < 2; ++sx) {
                        const int fo = swz(g, 4 * sx + t);
                        dmma(c.x, c.y, -Aa[fo], Ba[fo]);
                        dmma(c2.x, c2.y, -Ab[fo], Bb[fo]);
                    }
                }
                if (j < jhi) {
                    const double* Aa = Ls + (size_t)s_slot[I][j] * TILE;
                    const double* Ba = Ls + (size_t)s_slot[kk][j] * TILE;
#pragma unroll
                    for (int sx = 0; sx < 2; ++sx) {
                        const int fo = swz(g, 4 * sx + t);
                        dmma(c.x, c.y, -Aa[fo], Ba[fo]);
                    }
                }
                st2(p, c.x + c2.x, c.y + c2.y);
            }
        };
        // Look-ahead: while one warp factors the diagonal tile of column k (a chain of ~1400 clocks, longer under the other
        // CTAs' tensor work), the other three already subtract the finished columns j < k from column k + 1; column k's own
        // contribution follows when its panel exists.  (Without it 40 % of the warp time was the three warps waiting.)
        fetch_column(0);
        park_column(Cc, 0);
        if (Tn > 1) fetch_column(1);
        __syncthreads();
        for (int k = 0; k < Tn; ++k) {
            double* cur = Cc + (size_t)(k & 1) * Tcap * TILE;
            double* nxt = Cc + (size_t)((k + 1) & 1) * Tcap * TILE;
            if (k >= 1) update_tiles(cur, k, k - 1, k, warp, LWARPS);        // column k - 1's share of column k
            if (k + 1 < Tn) park_column(nxt, k + 1);
            __syncthreads();
            if (k + 2 < Tn) fetch_column(k + 2);
            if (k == Tn - 1) break;
            const int dw = k & (LWARPS - 1);
            if (warp == dw) {
                const double2 v = ld2(cur + swz(g, 2 * t));
                diag_factor(v.x, v.y, cur, Xs, g, t);
            } else if (k >= 1) {
                update_tiles(nxt, k + 1, 0, k, (warp - dw - 1) & (LWARPS - 1), LWARPS - 1);
            }
            __syncthreads();
            {   // panel: L[I][k] = C[I] X into its slot
                const double bf0 = Xs[swz(2 * t, g)], bf1 = Xs[swz(2 * t + 1, g)];
                for (int I = k + 1 + warp; I < Tn; I += LWARPS) {
                    const double2 v = ld2(cur + (size_t)(I - k) * TILE + swz(g, 2 * t));
                    double n0 = 0.0, n1 = 0.0;
                    dmma(n0, n1, v.x, bf0);
                    dmma(n0, n1, v.y, bf1);
                    st2(Ls + (size_t)s_slot[I][k] * TILE + swz(g, 2 * t), n0, n1);
                }
            }
            __syncthreads();
        }
        Cc = Cc + (size_t)((Tn - 1) & 1) * Tcap * TILE;       // (the last column's buffer holds the appended rows' tile)
        if (warp == 0) {
            const double2 v = ld2(Cc + swz(g, 2 * t));
            const int rank = fac_rank[fj], dofv = dof[fj];
            const int st = feat_status[fj];
            const double gam = gamma_from_last_tile(v.x, v.y, g, t, rr, rank);
            if (lane == 0) {
                const double crit = (dofv < chi2_len) ? chi2_table[dofv] : nan("");
                const bool pass = gam <= crit;
                feat_status[fj] = (unsigned char)((pass ? MSCKF_FEAT_ACCEPTED : MSCKF_FEAT_GATED_OUT) | (st & FLAG_KEEP));
                gamma[fj] = gam;
            }
        }
        __syncthreads();
    }
}

}  // namespace

size_t front_chol_smem_bytes(int Mcap) {
    return ((size_t)ftile::tiles_for_capacity(Mcap) * TILE + TILE) * sizeof(double);
}

#ifndef CHOL_LL
#define CHOL_LL 1
#endif

cudaError_t launch_front_chol_smem(const double2* Kbuf, int kt_d2, const int* pend_list, const int* pend_count, const int* feat_len,
                                   const int* fac_rank, const int* dof, const double* chi2_table, int chi2_len,
                                   unsigned char* feat_status, double* gamma, size_t fbase, int Mcap, cudaStream_t s) {
    if (CHOL_LL) {
        const int Tcap = ftile::tile_rows(Mcap);
        const size_t smem = ((size_t)ll_slots(Tcap) * TILE + (size_t)2 * Tcap * TILE + TILE) * sizeof(double);
        static size_t configured_ll = 0;
        if (smem > configured_ll) {
            cudaError_t e = cudaFuncSetAttribute(front_chol_ll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured_ll = smem;
        }
        int per_sm = (int)((227 * 1024) / (smem + 1024 + 512));
        if (per_sm > 5) per_sm = 5;
        if (per_sm < 1) per_sm = 1;
        front_chol_ll_kernel<<<148 * per_sm, LT, smem, s>>>(Kbuf, kt_d2, pend_list, pend_count, feat_len, fac_rank, dof, chi2_table,
                                                           chi2_len, feat_status, gamma, fbase, Tcap);
        msckf_count_launch(1);
        return cudaGetLastError();
    }
    const size_t smem = front_chol_smem_bytes(Mcap);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(front_chol_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const int per_sm = (int)((227 * 1024) / (smem + 1024));
    const int ctas = 148 * (per_sm < 1 ? 1 : (per_sm > 3 ? 3 : per_sm));
    front_chol_smem_kernel<<<ctas, CT, smem, s>>>(Kbuf, kt_d2, pend_list, pend_count, feat_len, fac_rank, dof, chi2_table, chi2_len,
                                                 feat_status, gamma, fbase);
    msckf_count_launch(1);
    return cudaGetLastError();
}
