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
#include "front_geom.cuh"

namespace {

using namespace fgeom;

constexpr int FW_THREADS = 256;
constexpr int FW_WARPS = FW_THREADS / 32;
constexpr int TC = 8;                         // 8x8 tile columns -> up to 64 measurement rows (32 observations)
constexpr int NTILE = TC * (TC + 1) / 2;

struct WarpScratch {
    double A[2 * FW_MAX_OBS][6];     // observability-projected Jacobian rows
    double aug[2 * FW_MAX_OBS][4];   // r~, Q1[:, 0..2] per measurement row
    double2 rowbuf[TC][32];          // one tile row of K in accumulator-fragment layout (staging for the registers)
};

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

    const GeoConsts gc = make_geo_consts(a);
    const double s2 = gc.s2;

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
        ObsRegs o = nxt;
        const int M = Mn;
        mask_obs(o, M, lane);
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


        // ---- selection, triangulation, residual / Jacobian blocks, basis of range(H_f), r~ (front_geom.cuh) ----
        double Hx[2][6], Hf[2][3], rt0, rt1;
        int rank;
        const int code = feature_geometry(a, gc, fj, M, lane, o, sR, sT, Hx, Hf, rt0, rt1, rank);
        if ((code & 0xF) != MSCKF_FEAT_ACCEPTED) {
            if (lane == 0) {
                a.feat_status[fj] = (unsigned char)code;
                a.gamma[fj] = nan("");
                a.dof[fj] = 0;
                a.fac_rank[fj] = 0;
            }
            continue;
        }

        const int dofv = n2 - rank;

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
