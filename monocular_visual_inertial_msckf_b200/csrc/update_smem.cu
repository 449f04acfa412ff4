// Rows 9b-10 of SURVEY.md section 8a for MSCKF_COMPRESS_GRAM, shared-memory-resident version for windows of up to 30
// poses (d <= 195): the Kalman update of reference src/msckf/MSCKF.py:604-607 and MSCKF.correct (611-661) driven by the
// information pair (G, g) = (T_H^T T_H, T_H^T r_n) that gram.cu accumulates.  Same algebra as update_info.cu (push-through
// identity, whitened coordinates, state ordered [C; I]), arranged so that every factor stays TRIANGULAR:
//     P'_CC = L L^T,  U = P_IC L^-T,  S_I = P_II - U U^T          (tile Cholesky, the I rows ride along as extra tile rows)
//     B  = s^2 I + L^T G L                                         (accumulated in registers, G streamed once from L2)
//     B  = Y^T Y with Y LOWER triangular                           (= Cholesky of the index-reversed B, so the same routine)
//     W  = [L; U] Y^-1                                             (lower trapezoidal: W_C is the Cholesky factor of P+_CC / s^2)
//     P+ = s^2 W W^T + [0 0; 0 S_I],   dx = W Y^-T L^T g           (== Joseph form of the reference, algebraically)
// About 2.7 c^3 flops (15 MFLOP at c = 180) instead of 25 MFLOP for the dense-W form of update_info.cu.
//
// One persistent 512-thread CTA per SM walks the filters.  The working matrix lives in shared memory as 8x8 tiles of the
// lower triangle (325 tiles = 166 KB at d = 195, XOR-swizzled so that every FP64 tensor-instruction fragment load is
// bank-conflict free); P' -> L' is spilled to the CTA's L2-resident scratch while B is factored in the same tiles, and
// comes back for the in-place triangular solve.  All O(n^3) work is DMMA.8x8x4 (mma.sync.m8n8k4.f64); the 8x8 diagonal
// tiles are factored by one warp with shuffles (same routine as the gate's factorisation kernel in front_split.cu).
#include "msckf_common.cuh"
#include "dense.cuh"

namespace {

using dense::dmma;

constexpr int UT = 512;                    // threads per CTA
constexpr int UW = UT / 32;
constexpr int MAX_NCP = 24;                // tile columns of the padded camera block: 6 N <= 192  ->  N <= 32
constexpr int DUAL_NCP = 23;               // up to here both row-block stages (G and T1 = G L) fit beside the tiles
constexpr int MAX_NT = MAX_NCP + 2;        // + 2 tile rows for the 15 IMU error states (padded to 16)
constexpr int MAX_TILES = MAX_NT * (MAX_NT + 1) / 2;
constexpr int TILE = 64;
constexpr int KBR = 16;                    // rows of G per streamed row block
constexpr int LDG = 196;                   // leading dimension of the row-block stages (= 4 mod 16: conflict-free fragments)
constexpr int NSUP = 5;                    // B super-tiles (2x2 tiles) owned by each warp: 16 * 5 >= 12 * 13 / 2
constexpr int VEC = 192;

__device__ __forceinline__ int swz(int r, int c) { return r * 8 + (c ^ ((r & 2) << 1)); }
__device__ __forceinline__ int toff(int I, int J) { return (((I * (I + 1)) >> 1) + J) * TILE; }
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ void cp_async16(void* s, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ int tri_row_f(int e) {
    int r = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
    if (tri(r) > e) --r;
    else if (tri(r + 1) <= e) ++r;
    return r;
}

// 1/sqrt(x) for x > 0, 0 otherwise (a non-positive pivot zeroes its column: positive SEMI-definite input)
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // one cubically convergent step: e = 1 - x y^2, y <- y (1 + e/2 + 3 e^2/8); seed error 2^-22 -> ~2^-64
    const double e = fma(-(x * y), y, 1.0);
    y = fma(y * e, fma(0.375, e, 0.5), y);
    return x > 0.0 ? y : 0.0;
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
        {
            const int k = 2 * tk;
            const double piv = shfl_d(D0, k * 4 + tk);
            const double rs = fast_rsqrt(piv);
            const double colv = (g > k) ? D0 * rs : 0.0;
            const double xsel = x0 * rs;
            if (t == tk) { x0 = xsel; L0 = (g > k) ? colv : ((g == k) ? D0 * rs : 0.0); }
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
            if (t == tk) { x1 = xsel; L1 = (g > k) ? colv : ((g == k) ? D1 * rs : 0.0); }
            const double lg = shfl_d(colv, g * 4 + tk);
            const double xg = shfl_d(xsel, g * 4 + tk);
            const double lc0 = shfl_d(colv, (2 * t) * 4 + tk);
            const double lc1 = shfl_d(colv, (2 * t + 1) * 4 + tk);
            D0 -= lg * lc0; x0 -= xg * lc0;
            D1 -= lg * lc1; x1 -= xg * lc1;
        }
    }
}

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// warp 0: factor the diagonal tile in place, leave X = L^-T in Xs (and in Xg when given)
__device__ __forceinline__ void diag_factor(double D0, double D1, double* tile, double* Xs, double* Xg, int g, int t) {
    double L0, L1, x0, x1;
    factor_tile(D0, D1, L0, L1, x0, x1, g, t);
    st2(tile + swz(g, 2 * t), L0, L1);
    st2(Xs + swz(g, 2 * t), x0, x1);
    if (Xg) st2(Xg + swz(g, 2 * t), x0, x1);
}

// Right-looking tile Cholesky of the first ncf tile columns of the nt x nt lower tile triangle T, two tile columns per
// block step; the trailing updates cover every tile row, so T[ncf:, ncf:] ends up holding the Schur complement and the
// rows below ncf the panel L21 = S21 L11^-T.  Xg (may be null) receives X = L_kk^-T of every diagonal tile.
__device__ void chol_tiles(double* T, const int nt, const int ncf, double* Xs, double* Xg) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int k0 = 0; k0 < ncf; k0 += 2) {
        const bool two = k0 + 1 < ncf;
        if (warp == 0) {
            double* tile = T + toff(k0, k0);
            const double2 v = ld2(tile + swz(g, 2 * t));
            diag_factor(v.x, v.y, tile, Xs, Xg ? Xg + k0 * TILE : nullptr, g, t);
        }
        __syncthreads();
        {   // panel of tile column k0
            const double bf0 = Xs[swz(2 * t, g)], bf1 = Xs[swz(2 * t + 1, g)];
            for (int i = k0 + 1 + warp; i < nt; i += UW) {
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
                diag_factor(v.x, v.y, tile, Xs, Xg ? Xg + k1 * TILE : nullptr, g, t);
            }
            __syncthreads();
            {   // tile column k1 below the diagonal: update from column k0 fused with the panel solve
                const double bf0 = Xs[swz(2 * t, g)], bf1 = Xs[swz(2 * t + 1, g)];
                const double* Lk = T + toff(k1, k0);
                const double b0 = Lk[swz(g, t)], b1 = Lk[swz(g, 4 + t)];
                for (int i = k1 + 1 + warp; i < nt; i += UW) {
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
            for (int q = warp; q < cnt; q += UW) {
                const int si = tri_row_f(q), sj = q - tri(si);
                const int I0 = j0 + 2 * si, J0 = j0 + 2 * sj;
                const bool vi1 = I0 + 1 < nt, vj1 = (J0 + 1 < nt) && (J0 + 1 <= I0 + 1);
                // tiles: (I0,J0) always valid (J0 <= I0); (I0,J0+1) valid iff J0+1 <= I0; (I0+1,J0) iff vi1; (I0+1,J0+1) iff vi1 && vj1
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

template <bool single>
__global__ void __launch_bounds__(UT, 1) update_smem_kernel(const ApplyArgs a) {
    extern __shared__ double sm[];
    // Capacity-dependent carve-up: up to 23 tile columns (N <= 30) the row block of G and the row block of T1 = G L have a
    // stage each (the next block of G streams in while B accumulates); with 24 (N = 31, 32) the tiles leave room for ONE
    // stage, T1 overwrites the block of G it was computed from and the next block is fetched after the accumulation.
    const int ncp_max = (6 * a.dims.N + 7) >> 3;
    const int stage_doubles = single ? max(KBR * LDG, 2 * ncp_max * TILE + 2 * TILE) : 2 * KBR * LDG;
    double* T = sm;                                  // [tri(ncp_max + 2)][64]
    double* Gs = T + (size_t)tri(ncp_max + 2) * TILE;   // [KBR][LDG]  row block of G          | Y row buffers of the solve
    double* T1s = single ? Gs : Gs + KBR * LDG;      // [KBR][LDG]  row block of T1 = G L   |
    double* Xs = Gs + stage_doubles;                 // [2][64]     X = L_kk^-T of the current diagonal tile(s)
    double* gvec = Xs + 2 * TILE;                    // [VEC] g (padded)
    double* yvec = gvec + VEC;                       // [VEC] y = L^T g
    double* zvec = yvec + VEC;                       // [VEC] z = Y^-T y
    double* dxs = zvec + VEC;                        // [208] dx in the reference's state order

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ldp = a.dims.ldp, Ncap = a.dims.N, ccap = 6 * Ncap;
    const double s2 = a.sigma * a.sigma;
    double* scratch = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;
    const int ncp_cap = (ccap + 7) >> 3;
    double* Lg = scratch;                                                  // spilled L' tiles
    double* Yg = Lg + (size_t)tri(ncp_cap + 2) * TILE;                     // Cholesky factor of the index-reversed B
    double* Xg = Yg + (size_t)tri(ncp_cap) * TILE;                         // its diagonal-tile inverses

    for (int b = blockIdx.x; b < a.dims.B; b += gridDim.x) {
        double* P = a.P + (size_t)b * ldp * ldp;
        double* dxo = a.delta_x + (size_t)b * ldp;
        const int nc = min(a.n_cams[b], Ncap);
        if (a.n_accepted[b] <= 0 || nc <= 0) {   // nothing survived gating: the reference returns before touching the state
            for (int i = tid; i < ldp; i += UT) dxo[i] = 0.0;
            continue;
        }
        const int c = 6 * nc, cp = (c + 7) & ~7, ncp = cp >> 3, nt = ncp + 2;
        const int ntiles = nt * (nt + 1) / 2;
        const double* G = a.T_H + (size_t)b * ccap * ccap;
        const double* gv = a.r_n + (size_t)b * ccap;
        // padded [C; I] index -> index in the reference's state order (-1: padding)
        auto smap = [=](int i) { return i < c ? 15 + i : (i < cp ? -1 : (i < cp + 15 ? i - cp : -1)); };

        // ---- S1: P' = P in [C; I] order, lower tile triangle ----
        for (int i = warp; i < nt * 8; i += UW) {
            const int si = smap(i), I = i >> 3;
            const double* Prow = P + (size_t)(si < 0 ? 0 : si) * ldp;
            // all the loads of the row are issued before the first store (one dependent load -> store per iteration left
            // the global-memory latency exposed: 7 % of the kernel's stall samples, profiles/r01q)
            constexpr int JU = (MAX_NT * 8 + 31) / 32;
            double vv[JU];
#pragma unroll
            for (int u = 0; u < JU; ++u) {
                const int j = lane + 32 * u;
                const int sj = smap(j);
                double v = 0.0;
                if (j < (I + 1) * 8 && j <= i) v = (si < 0 || sj < 0) ? ((i == j) ? 1.0 : 0.0) : Prow[sj];
                vv[u] = v;
            }
#pragma unroll
            for (int u = 0; u < JU; ++u) {
                const int j = lane + 32 * u;
                if (j < (I + 1) * 8) T[toff(I, j >> 3) + swz(i & 7, j & 7)] = vv[u];
            }
        }
        for (int e = tid; e < stage_doubles; e += UT) Gs[e] = 0.0;   // both stages (the solve below reuses them)
        for (int e = tid; e < VEC; e += UT) gvec[e] = (e < c) ? gv[e] : 0.0;
        __syncthreads();

        // ---- S2: L' (tile rows 0..nt-1, columns < ncp) and the Schur complement S_I (tiles [ncp:, ncp:]) ----
        chol_tiles(T, nt, ncp, Xs, nullptr);

        // ---- S3: y = L^T g;  B = L^T (G L) accumulated in registers over 16-row blocks of G ----
        for (int j = tid; j < cp; j += UT) {
            double acc = 0.0;
            for (int k = j; k < cp; ++k) acc += T[toff(k >> 3, j >> 3) + swz(k & 7, j & 7)] * gvec[k];
            yvec[j] = acc;
        }
        const int nsp = (ncp + 1) >> 1;
        const int nsup = nsp * (nsp + 1) / 2;
        int SI[NSUP], SJ[NSUP];
        double acc[NSUP][2][2][2];
#pragma unroll
        for (int s = 0; s < NSUP; ++s) {
            const int q = warp + UW * s;
            SI[s] = (q < nsup) ? tri_row_f(q) : -1;
            SJ[s] = (q < nsup) ? q - tri(SI[s]) : 0;
#pragma unroll
            for (int x = 0; x < 8; ++x) acc[s][x >> 2][(x >> 1) & 1][x & 1] = 0.0;
        }
        auto load_G_rows = [&](int Kb) {
            // rows 16 Kb .. 16 Kb + 15 of G (zero beyond c); columns [c, cp) of the stage stay zero
            const int chunks = c >> 1;
            for (int r = warp; r < KBR; r += UW) {
                const int gr = KBR * Kb + r;
                double* dst = Gs + r * LDG;
                if (gr < c) {
                    const double* src = G + (size_t)gr * ccap;
                    for (int x = lane; x < chunks; x += 32) cp_async16(dst + 2 * x, src + 2 * x);
                } else {
                    for (int x = lane; x < cp; x += 32) dst[x] = 0.0;
                }
            }
            cp_async_commit();
        };
        load_G_rows(0);
        for (int Kb = 0; Kb < nsp; ++Kb) {
            cp_async_wait_all();
            __syncthreads();
            // T1[rows of this block][:] = G[rows][:] L ; warp w < ceil(ncp / 2) takes the tile columns w and ncp-1-w
            if (!single) {
                if (warp < ((ncp + 1) >> 1)) {
#pragma unroll 1
                    for (int u = 0; u < 2; ++u) {
                        const int j = u == 0 ? warp : ncp - 1 - warp;
                        if (u == 1 && j <= warp) break;
                        double r00 = 0.0, r01 = 0.0, r10 = 0.0, r11 = 0.0;
                        double q00 = 0.0, q01 = 0.0, q10 = 0.0, q11 = 0.0;   // second accumulator set (odd k tiles): shorter chains
                        int kt = j;
                        for (; kt + 1 < ncp; kt += 2) {
                            const double* La = T + toff(kt, j);
                            const double* Lb = T + toff(kt + 1, j);
#pragma unroll
                            for (int s = 0; s < 2; ++s) {
                                const double bfa = La[swz(4 * s + t, g)];
                                const double bfb = Lb[swz(4 * s + t, g)];
                                const int ca = kt * 8 + 4 * s + t;
                                dmma(r00, r01, Gs[g * LDG + ca], bfa);
                                dmma(r10, r11, Gs[(8 + g) * LDG + ca], bfa);
                                dmma(q00, q01, Gs[g * LDG + ca + 8], bfb);
                                dmma(q10, q11, Gs[(8 + g) * LDG + ca + 8], bfb);
                            }
                        }
                        if (kt < ncp) {
                            const double* La = T + toff(kt, j);
#pragma unroll
                            for (int s = 0; s < 2; ++s) {
                                const double bfa = La[swz(4 * s + t, g)];
                                const int ca = kt * 8 + 4 * s + t;
                                dmma(r00, r01, Gs[g * LDG + ca], bfa);
                                dmma(r10, r11, Gs[(8 + g) * LDG + ca], bfa);
                            }
                        }
                        st2(T1s + g * LDG + j * 8 + 2 * t, r00 + q00, r01 + q01);
                        st2(T1s + (8 + g) * LDG + j * 8 + 2 * t, r10 + q10, r11 + q11);
                    }
                }
                __syncthreads();
            } else {
                // single stage: the results wait in registers until every warp is done with this block of G
                double t1r[2][4];
                int t1j[2] = {-1, -1};
                if (warp < ((ncp + 1) >> 1)) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int j = u == 0 ? warp : ncp - 1 - warp;
                        if (u == 1 && j <= warp) break;
                        t1j[u] = j;
                        double r00 = 0.0, r01 = 0.0, r10 = 0.0, r11 = 0.0;
                        double q00 = 0.0, q01 = 0.0, q10 = 0.0, q11 = 0.0;   // second accumulator set (odd k tiles): shorter chains
                        int kt = j;
                        for (; kt + 1 < ncp; kt += 2) {
                            const double* La = T + toff(kt, j);
                            const double* Lb = T + toff(kt + 1, j);
#pragma unroll
                            for (int s = 0; s < 2; ++s) {
                                const double bfa = La[swz(4 * s + t, g)];
                                const double bfb = Lb[swz(4 * s + t, g)];
                                const int ca = kt * 8 + 4 * s + t;
                                dmma(r00, r01, Gs[g * LDG + ca], bfa);
                                dmma(r10, r11, Gs[(8 + g) * LDG + ca], bfa);
                                dmma(q00, q01, Gs[g * LDG + ca + 8], bfb);
                                dmma(q10, q11, Gs[(8 + g) * LDG + ca + 8], bfb);
                            }
                        }
                        if (kt < ncp) {
                            const double* La = T + toff(kt, j);
#pragma unroll
                            for (int s = 0; s < 2; ++s) {
                                const double bfa = La[swz(4 * s + t, g)];
                                const int ca = kt * 8 + 4 * s + t;
                                dmma(r00, r01, Gs[g * LDG + ca], bfa);
                                dmma(r10, r11, Gs[(8 + g) * LDG + ca], bfa);
                            }
                        }
                        t1r[u][0] = r00 + q00; t1r[u][1] = r01 + q01; t1r[u][2] = r10 + q10; t1r[u][3] = r11 + q11;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (t1j[u] >= 0) {
                        st2(T1s + g * LDG + t1j[u] * 8 + 2 * t, t1r[u][0], t1r[u][1]);
                        st2(T1s + (8 + g) * LDG + t1j[u] * 8 + 2 * t, t1r[u][2], t1r[u][3]);
                    }
                }
                __syncthreads();
            }
            if (!single && Kb + 1 < nsp) load_G_rows(Kb + 1);
            // B[I][J] += L[kt][I]^T T1[kt][J] for the two tile rows kt of this block
#pragma unroll
            for (int s = 0; s < NSUP; ++s) {
                if (SI[s] < 0 || SI[s] > Kb) continue;
                const int I0 = 2 * SI[s], J0 = 2 * SJ[s];
#pragma unroll
                for (int ktl = 0; ktl < 2; ++ktl) {
                    const int kt = 2 * Kb + ktl;
                    if (kt >= ncp) continue;
                    const bool i0 = I0 <= kt, i1 = I0 + 1 <= kt;
#pragma unroll
                    for (int st = 0; st < 2; ++st) {
                        const int fo = swz(4 * st + t, g);
                        const double a0 = i0 ? T[toff(kt, I0) + fo] : 0.0;
                        const double a1 = i1 ? T[toff(kt, I0 + 1) + fo] : 0.0;
                        const double* trow = T1s + (ktl * 8 + 4 * st + t) * LDG + g;
                        const double b0 = trow[J0 * 8];
                        const double b1 = trow[J0 * 8 + 8];
                        if (i0) {
                            dmma(acc[s][0][0][0], acc[s][0][0][1], a0, b0);
                            dmma(acc[s][0][1][0], acc[s][0][1][1], a0, b1);
                        }
                        if (i1) {
                            dmma(acc[s][1][0][0], acc[s][1][0][1], a1, b0);
                            dmma(acc[s][1][1][0], acc[s][1][1][1], a1, b1);
                        }
                    }
                }
            }
            if (single && Kb + 1 < nsp) {
                __syncthreads();                      // T1 (in the only stage) has been consumed
                load_G_rows(Kb + 1);
            }
        }
        __syncthreads();

        // ---- S4: spill L' to scratch; the tiles now receive the index-reversed B (+ s^2 I) and the appended row y~ ----
        for (int e = tid * 2; e < ntiles * TILE; e += UT * 2) st2(Lg + e, T[e], T[e + 1]);
        __syncthreads();
#pragma unroll
        for (int s = 0; s < NSUP; ++s) {
            if (SI[s] < 0) continue;
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int I = 2 * SI[s] + ii, J = 2 * SJ[s] + jj;
                    if (I >= ncp || J > I) continue;
                    double* dst = T + toff(ncp - 1 - J, ncp - 1 - I);
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int col = 2 * t + r;
                        double v = acc[s][ii][jj][r];
                        if (I == J && g == col) v += (I * 8 + g < c) ? s2 : 1.0;
                        dst[swz(7 - col, 7 - g)] = v;
                    }
                }
        }
        for (int e = tid; e < (ncp + 1) * TILE; e += UT) {
            const int Jp = e >> 6, r = (e >> 3) & 7, bb = e & 7;
            double v;
            if (Jp < ncp) v = (r == 0) ? yvec[cp - 1 - Jp * 8 - bb] : 0.0;
            else v = (r == bb) ? 1.0 : 0.0;
            T[toff(ncp, Jp) + swz(r, bb)] = v;
        }
        __syncthreads();

        // ---- S5: reversed B = C C^T (C left in the tiles), appended row -> C^-1 y~ = reversed z ----
        chol_tiles(T, ncp + 1, ncp, Xs, Xg);
        for (int k = tid; k < cp; k += UT) {
            const int m = cp - 1 - k;
            zvec[k] = T[toff(ncp, m >> 3) + swz(0, m & 7)];
        }
        {
            const int n = tri(ncp) * TILE;
            for (int e = tid * 2; e < n; e += UT * 2) st2(Yg + e, T[e], T[e + 1]);
        }
        __syncthreads();
        // ---- S6: L' back into the tiles ----
        for (int e = tid * 2; e < ntiles * TILE; e += UT * 2) {
            const double2 v = ld2(Lg + e);
            st2(T + e, v.x, v.y);
        }
        __threadfence_block();
        __syncthreads();

        // ---- S7: W = L' Y^-1 in place, tile column j = ncp-1 .. 0; row jp = ncp-1-j of C (tiles (jp, 0..jp)) holds the
        //      tiles Y[k][j] = flip-transpose(C[jp][ncp-1-k]) and streams through two shared-memory buffers ----
        {
            double* Ybuf = Gs;                         // [2][MAX_NCP * 64] + [2][64] for the diagonal inverses
            double* Xbuf = Gs + 2 * ncp_max * TILE;
            auto load_Y_row = [&](int jp) {
                double* dst = Ybuf + (jp & 1) * ncp_max * TILE;
                const double* src = Yg + (size_t)tri(jp) * TILE;
                const int n = (jp + 1) * TILE / 2;
                for (int x = tid; x < n; x += UT) cp_async16(dst + 2 * x, src + 2 * x);
                if (tid < 32) cp_async16(Xbuf + (jp & 1) * TILE + 2 * tid, Xg + (size_t)jp * TILE + 2 * tid);
                cp_async_commit();
            };
            load_Y_row(0);
            for (int jp = 0; jp < ncp; ++jp) {
                const int j = ncp - 1 - jp;
                cp_async_wait_all();
                __syncthreads();
                if (jp + 1 < ncp) load_Y_row(jp + 1);
                const double* Yrow = Ybuf + (jp & 1) * ncp_max * TILE;
                const double* Xb = Xbuf + (jp & 1) * TILE;
                const double xb0 = Xb[swz(7 - 2 * t, 7 - g)], xb1 = Xb[swz(6 - 2 * t, 7 - g)];
                // rows owned by this warp: i >= 16 -> warp nt-1-i, i < 16 -> warp 9 + (i % 7); enumerated directly
                for (int q = 0; q < 4; ++q) {
                    const int i = (q < 3) ? ((warp >= 9 && warp - 9 + 7 * q < 16) ? warp - 9 + 7 * q : -1)
                                          : ((nt - 1 - warp >= 16) ? nt - 1 - warp : -1);
                    if (i < j || i >= nt) continue;
                    double* p = T + toff(i, j) + swz(g, 2 * t);
                    double2 v = ld2(p);
                    double w0 = 0.0, w1 = 0.0;
                    const int kend = min(i, ncp - 1);
                    int k = j + 1;
                    for (; k + 1 <= kend; k += 2) {
                        const double* Wa = T + toff(i, k);
                        const double* Wb = T + toff(i, k + 1);
                        const double* Ya = Yrow + (ncp - 1 - k) * TILE;
                        const double* Yb = Yrow + (ncp - 2 - k) * TILE;
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            const int yo = swz(7 - g, 7 - 4 * s - t);
                            dmma(v.x, v.y, -Wa[swz(g, 4 * s + t)], Ya[yo]);
                            dmma(w0, w1, -Wb[swz(g, 4 * s + t)], Yb[yo]);
                        }
                    }
                    if (k <= kend) {
                        const double* Wa = T + toff(i, k);
                        const double* Ya = Yrow + (ncp - 1 - k) * TILE;
#pragma unroll
                        for (int s = 0; s < 2; ++s) dmma(v.x, v.y, -Wa[swz(g, 4 * s + t)], Ya[swz(7 - g, 7 - 4 * s - t)]);
                    }
                    v.x += w0; v.y += w1;
                    double n0 = 0.0, n1 = 0.0;
                    dmma(n0, n1, v.x, xb0);
                    dmma(n0, n1, v.y, xb1);
                    st2(p, n0, n1);
                }
            }
            __syncthreads();
        }

        // ---- dx = W z, in the reference's state order ----
        for (int i = tid; i < nt * 8; i += UT) {
            const int si = smap(i);
            if (si < 0) continue;
            const int I = i >> 3, kend = min(i, cp - 1);
            double accd = 0.0;
            for (int k = 0; k <= kend; ++k) accd += T[toff(I, k >> 3) + swz(i & 7, k & 7)] * zvec[k];
            dxs[si] = accd;
        }
        __syncthreads();
        for (int i = tid; i < ldp; i += UT) dxo[i] = (i < 15 + c) ? dxs[i] : 0.0;

        // ---- S8: P+ = s^2 W W^T + [0 0; 0 S_I], written symmetric in the reference's state order ----
        {
            const int nsr = (nt + 1) >> 1;
            const int cnt = nsr * (nsr + 1) / 2;
            for (int q = warp; q < cnt; q += UW) {
                const int si = tri_row_f(q), sj = q - tri(si);
                const int I0 = 2 * si, J0 = 2 * sj;
                const bool vi1 = I0 + 1 < nt, vj1 = J0 + 1 < nt;
                double o[2][2][2];
#pragma unroll
                for (int x = 0; x < 8; ++x) o[x >> 2][(x >> 1) & 1][x & 1] = 0.0;
                const int kmax = min(J0 + 1, ncp - 1);
                for (int k = 0; k <= kmax; ++k) {
                    const bool j0k = k <= J0;                 // tile (J0, k) exists
                    const bool j1k = vj1 && k <= J0 + 1;
                    const bool i0k = k <= I0;
                    const bool i1k = vi1 && k <= I0 + 1;
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const int fo = swz(g, 4 * s + t);
                        const double a0 = i0k ? T[toff(I0, k) + fo] : 0.0;
                        const double a1 = i1k ? T[toff(I0 + 1, k) + fo] : 0.0;
                        const double b0 = j0k ? T[toff(J0, k) + fo] : 0.0;
                        const double b1 = j1k ? T[toff(J0 + 1, k) + fo] : 0.0;
                        if (i0k && j0k) dmma(o[0][0][0], o[0][0][1], a0, b0);
                        if (i0k && j1k) dmma(o[0][1][0], o[0][1][1], a0, b1);
                        if (i1k && j0k) dmma(o[1][0][0], o[1][0][1], a1, b0);
                        if (i1k && j1k) dmma(o[1][1][0], o[1][1][1], a1, b1);
                    }
                }
#pragma unroll
                for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const int I = I0 + ii, J = J0 + jj;
                        if (I >= nt || J > I) continue;
                        const int pi = I * 8 + g, spi = smap(pi);
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            const int pj = J * 8 + 2 * t + r, spj = smap(pj);
                            if (pj > pi || spi < 0 || spj < 0) continue;
                            double v = s2 * o[ii][jj][r];
                            if (J >= ncp) v += T[toff(I, J) + swz(g, 2 * t + r)];
                            P[(size_t)spi * ldp + spj] = v;
                            P[(size_t)spj * ldp + spi] = v;
                        }
                    }
            }
        }

        // ---- state correction (MSCKF.py:616-661) ----
        if (tid == 0) {
            double* st = a.imu_state + (size_t)b * MSCKF_IMU_STATE_DOUBLES;
            apply_rotation_correction(st, dxs);
            for (int k = 0; k < 3; ++k) {
                st[9 + k] += dxs[12 + k];    // position
                st[12 + k] += dxs[6 + k];    // velocity
                st[15 + k] += dxs[3 + k];    // gyroscope bias
                st[18 + k] += dxs[9 + k];    // accelerometer bias
            }
        }
        if (tid >= 32 && tid - 32 < nc) {
            const int ci = tid - 32;
            double* Rc = a.cam_R + ((size_t)b * Ncap + ci) * 9;
            double* tc = a.cam_t + ((size_t)b * Ncap + ci) * 3;
            apply_rotation_correction(Rc, dxs + 15 + 6 * ci);
            for (int k = 0; k < 3; ++k) tc[k] += dxs[15 + 6 * ci + 3 + k];
        }
        __syncthreads();
    }
}

size_t us_smem_bytes(const msckf_dims& d) {
    const size_t ncp = (6 * (size_t)d.N + 7) / 8, nt = ncp + 2;
    const size_t stage = ncp > DUAL_NCP ? (KBR * LDG > 2 * ncp * TILE + 2 * TILE ? KBR * LDG : 2 * ncp * TILE + 2 * TILE) : 2 * KBR * LDG;
    return (nt * (nt + 1) / 2 * TILE + stage + 2 * TILE + 3 * VEC + 208) * sizeof(double);
}
static_assert(2 * KBR * LDG >= 2 * DUAL_NCP * TILE + 2 * TILE, "Y row buffers must fit the G/T1 stages");

}  // namespace

bool update_smem_supported(const msckf_dims& d) {
    return 6 * d.N <= 8 * MAX_NCP && d.ldp >= 15 + 6 * d.N && us_smem_bytes(d) <= 227 * 1024;
}

size_t update_smem_scratch_doubles(const msckf_dims& d) {
    const size_t ncp = (6 * (size_t)d.N + 7) / 8, nt = ncp + 2;
    return (nt * (nt + 1) / 2 + ncp * (ncp + 1) / 2 + ncp) * TILE + 64;
}

cudaError_t launch_update_smem(const ApplyArgs& a, int n_ctas, cudaStream_t s) {
    static int sms = 0;
    static size_t configured = 0;
    const size_t smem = us_smem_bytes(a.dims);
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const bool single = (6 * a.dims.N + 7) / 8 > DUAL_NCP;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(update_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(update_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    int grid = a.dims.B < sms ? a.dims.B : sms;
    if (grid > n_ctas) grid = n_ctas;
    if (single) update_smem_kernel<true><<<grid, UT, smem, s>>>(a);
    else update_smem_kernel<false><<<grid, UT, smem, s>>>(a);
    msckf_count_launch(1);
    return cudaGetLastError();
}
