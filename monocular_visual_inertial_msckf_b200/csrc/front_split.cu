// Per-feature front end as TWO kernels (rows 3-8 of SURVEY.md section 8a) for tracks of up to 32 observations:
//   selection + closed-form triangulation     reference src/msckf/MSCKF.py:458-495, src/utils/geometry.py:274-303
//   residual + Jacobian blocks                MSCKF.py:497-548, src/msckf/Camera.py:54-67
//   left-null-space projection of H_f         MSCKF.py:554-559
//   chi-square Mahalanobis gate               MSCKF.py:561-568
//
// A single fused kernel would have to keep the 2M x 2M matrix K = H_x P H_x^T + sigma^2 I of a feature in registers while
// the camera block of P occupies the SM's shared memory: 255 registers x 8 warps, two warps per scheduler, and every
// phase of it is latency-bound (measured in round 1: 42.6 ms against 18.6 ms, profiles/r01n; that kernel was removed).
// Here the two halves get the resources each one needs:
//
//   front_asm_kernel   CTA = (filter, feature chunk: a whole filter when the batch fills the machine), 16 warps, <= 128
//                      registers.  P's camera block (lower block triangle, 6x6 blocks padded to 38 doubles -> conflict-free
//                      16-byte reads) + window poses staged once per CTA.  Per feature (one warp, lane = observation):
//                      geometry (front_geom.cuh), factors for the compression stage, then the M (M + 1) / 2 lower 2x2
//                      blocks  A_i P[s_i, s_k] A_k^T  as a FLAT list, one block per lane per pass (97 % lane use at
//                      M = 30), stored straight into the accumulator-fragment tile layout of the FP64 tensor instruction
//                      in a global (L2 / HBM) buffer.  The four appended rows [r~; Q1^T] are ordinary rows of the padded
//                      matrix (they fill the padding rows of the last tile row at M = 30): 18 KB per feature, written
//                      once, read once.  The warps of a CTA start staggered so that the geometry phases of some overlap
//                      the shared-memory-bound pair loops of the others.
//   front_chol_kernel  one warp per pending feature, 16 warps per SM, <= 128 registers: LEFT-looking blocked Cholesky
//                      over 8x8 tiles in accumulator-fragment layout (DMMA.8x8x4 updates / panel solves, shuffles inside
//                      the diagonal tiles); only the finished tiles still needed and the current block column are live,
//                      the next block column's tiles are in flight (coalesced 16-byte loads); fully unrolled per tile
//                      count with the matrix aligned to the bottom-right corner, so every register index is static and no
//                      tile is ever moved.  The block under the appended rows ends up holding minus the Gram matrix of
//                      [y; V]: gamma and the gate decision follow.
//
// gamma = y.y - (V^T y)^T (V^T V)^-1 (V^T y),  y = L^-1 r~,  V = L^-1 Q1  (basis-free form of the projected Mahalanobis
// distance).
#include "msckf_common.cuh"
#include "dense.cuh"
#include "front_geom.cuh"

namespace {

using namespace fgeom;

constexpr int FA_THREADS = 512;
constexpr int FA_WARPS = FA_THREADS / 32;
constexpr int FC_THREADS = 128;
constexpr int FC_WARPS = FC_THREADS / 32;
constexpr int FLAG_KEEP = MSCKF_FEAT_LOST_FLAG | MSCKF_FEAT_TRIANGULATED;   // selection flags carried into the final status
constexpr int TC = 9;                          // 8x8 tile columns: up to 64 measurement rows (32 observations) + 4 appended rows
constexpr int NTILE = TC * (TC + 1) / 2;
constexpr int PB = 38;                         // doubles per staged 6x6 block of P (stride 76 words = 12 mod 32)
constexpr int AR = 14;                         // doubles per observation's 2x6 block in the warp scratch (28 words)
constexpr int KT_D2 = NTILE * 32;               // double2 per feature slot in the tile buffer (45 tiles; M = 30 uses 36)

__device__ __forceinline__ int tile_index(int I, int J) { return I * (I + 1) / 2 + J; }

// ---------------------------------------------------------------------------------------------------------------------
// kernel A: geometry + assembly of K into the tile buffer
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FA_THREADS, 1)
front_asm_kernel(const FrontArgs a, double2* __restrict__ Kbuf, int* __restrict__ pend_list, int* __restrict__ pend_count,
                 const int b0, const int feats_per_cta, const unsigned stagger_ns) {
    const int bl = blockIdx.y, b = b0 + bl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = a.dims.F, Mcap = a.dims.M, Ncap = a.dims.N, ldp = a.dims.ldp;
    const int nc = min(a.n_cams[b], Ncap);

    extern __shared__ double smem[];
    double* sP = smem;                                              // [nblk(Ncap)][PB] lower block triangle of P_CC
    double* sR = sP + (size_t)(Ncap * (Ncap + 1) / 2) * PB;         // [Ncap][9]
    double* sT = sR + (size_t)Ncap * 9;                             // [Ncap][3]
    double* wsA = sT + (size_t)Ncap * 3 + (size_t)warp * (FW_MAX_OBS * AR);   // 16-byte aligned (12 Ncap doubles of poses)
    const int nthreads = blockDim.x, nwarps = blockDim.x >> 5;

    const int f_begin = blockIdx.x * feats_per_cta;
    const int f_end = min(F, f_begin + feats_per_cta);
    if (f_begin >= f_end) return;

    // ---- stage the window once per CTA ----
    {
        const double* P = a.P + (size_t)b * ldp * ldp;
        const int c = 6 * nc;
        for (int row = warp; row < c; row += nwarps) {
            const int ba = row / 6, r = row - 6 * ba;
            const double* Prow = P + (size_t)(15 + row) * ldp + 15;
            double* dst = sP + (size_t)tile_index(ba, 0) * PB + r * 6;
            // (all the loads of the row first: a load -> store pair per iteration leaves the global latency exposed)
            constexpr int CU = (6 * 32 + 31) / 32;     // rows of the supported windows (N <= 32) span at most 6 x 32 columns
            double pv[CU];
#pragma unroll
            for (int u = 0; u < CU; ++u) {
                const int col = lane + 32 * u;
                pv[u] = (col < 6 * ba + 6) ? Prow[col] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < CU; ++u) {
                const int col = lane + 32 * u;
                const int bb = col / 6;
                if (col < 6 * ba + 6) dst[bb * PB + (col - 6 * bb)] = pv[u];
            }
        }
        const double* camR = a.cam_R + (size_t)b * Ncap * 9;
        const double* camt = a.cam_t + (size_t)b * Ncap * 3;
        for (int e = tid; e < nc * 9; e += nthreads) sR[e] = camR[e];
        for (int e = tid; e < nc * 3; e += nthreads) sT[e] = camt[e];
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
    // Every warp alternates between the geometry (no shared-memory traffic) and the pair loop (shared-memory bound), and
    // all the warps of the CTA leave the barrier above in the same phase: staggering their start spreads the phases, so
    // that the pair loops of some warps run under the geometry of the others instead of all of them queueing on the
    // shared-memory pipe at once.
    if (stagger_ns) __nanosleep(warp * stagger_ns);

    for (int j = f_begin + warp; j < f_end; j += nwarps) {
        const size_t fj = (size_t)b * F + j;
        ObsRegs o = nxt;
        const int M = Mn;
        mask_obs(o, M, lane);
        {   // prefetch the next feature of this warp
            const int jn = j + nwarps;
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

        int rank, code;
        double rt0, rt1;
        {
            double Hx[2][6], Hf[2][3];
            code = feature_geometry(a, gc, fj, M, lane, o, sR, sT, Hx, Hf, rt0, rt1, rank);
            if ((code & 0xF) != MSCKF_FEAT_ACCEPTED) {
                if (lane == 0) {
                    a.feat_status[fj] = (unsigned char)code;
                    a.gamma[fj] = nan("");
                    a.dof[fj] = 0;
                    a.fac_rank[fj] = 0;
                }
                continue;
            }
            // ---- factors for the compression stage + the warp's shared scratch ----
            __syncwarp();
            if (live) {
                double2* fA = reinterpret_cast<double2*>(a.fac_A + (fj * Mcap + lane) * 12);
                double2* fQ = reinterpret_cast<double2*>(a.fac_Q + (fj * Mcap + lane) * 6);
                double2* sA = reinterpret_cast<double2*>(wsA + lane * AR);
#pragma unroll
                for (int c2 = 0; c2 < 6; ++c2) {
                    const double2 v = make_double2(Hx[c2 / 3][2 * (c2 % 3)], Hx[c2 / 3][2 * (c2 % 3) + 1]);
                    fA[c2] = v;
                    sA[c2] = v;
                }
                fQ[0] = make_double2(Hf[0][0], Hf[0][1]);
                fQ[1] = make_double2(Hf[0][2], Hf[1][0]);
                fQ[2] = make_double2(Hf[1][1], Hf[1][2]);
                *reinterpret_cast<double2*>(a.fac_rt + (fj * Mcap + lane) * 2) = make_double2(rt0, rt1);
            }
            const int dofv = n2 - rank;
            if (dofv <= 0) {   // empty null space: gamma = 0 against chi2.ppf(0.95, 0) = NaN -> rejected (MSCKF.py:566-568)
                if (lane == 0) {
                    const double crit = a.chi2_table[0];
                    const bool pass = 0.0 <= crit;
                    a.feat_status[fj] = (unsigned char)((pass ? MSCKF_FEAT_ACCEPTED : MSCKF_FEAT_GATED_OUT) | (code & FLAG_KEEP));
                    a.gamma[fj] = 0.0;
                    a.dof[fj] = 0;
                    a.fac_rank[fj] = rank;
                }
                continue;
            }
            __syncwarp();

            // ---- exact early-out (MSCKF_GATE_BOUNDS): |r~|^2 / (s^2 + tr S_x) <= gamma <= |r~|^2 / s^2 ----
            if (a.par.gate_mode == MSCKF_GATE_BOUNDS) {
                double trs = 0.0;
                if (live) {
                    const double* Pd = sP + (size_t)tile_index(o.slot, o.slot) * PB;
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
                                                            (code & FLAG_KEEP) | MSCKF_FEAT_GAMMA_BOUND);
                        a.gamma[fj] = acc ? gub : glb;
                        a.dof[fj] = dofv;
                        a.fac_rank[fj] = rank;
                    }
                    continue;
                }
            }

            // ---- pending: the factorisation kernel finishes this feature ----
            const int fl = bl * F + j;
            int pend_pos = 0;     // (the list position is consumed after the pair loop: the atomic's latency stays hidden)
            if (lane == 0) {
                a.feat_status[fj] = (unsigned char)(MSCKF_FEAT_PENDING | (code & FLAG_KEEP));
                a.dof[fj] = dofv;
                a.fac_rank[fj] = rank;
                pend_pos = atomicAdd(pend_count, 1);
            }
            double2* Kt = Kbuf + (size_t)fl * KT_D2;
            // Padded matrix of the factorisation kernel: rows [0, n2) measurement rows, [n2, a4) identity (a4 = n2 rounded up
            // to a multiple of 4), [a4, a4 + 4) the appended rows [r~; Q1^T] (they never straddle a tile row and their
            // diagonal block is zero: after the factorisation of the columns < a4 it holds minus their Gram matrix), rest
            // zero.  At M = 30 the appended rows fill the four padding rows of the eighth tile row.
            const int a4 = (n2 + 3) & ~3;
            const int Tn = (a4 + 4 + 7) >> 3, Mp = 4 * Tn;
            const int rr = a4 - 8 * (Tn - 1);          // local row of the first appended row in the last tile row: 0 or 4
            const int ia = a4 >> 1;                     // the appended rows as two observation-sized row pairs ia, ia + 1
            if (live) {   // columns (2 lane, 2 lane + 1) of the four appended rows
                double2* Gt = Kt + tile_index(Tn - 1, lane >> 2) * 32 + 4 * rr + (lane & 3);
                Gt[0] = make_double2(rt0, rt1);
#pragma unroll
                for (int q = 0; q < 3; ++q) Gt[4 * (q + 1)] = make_double2(Hf[0][q], Hf[1][q]);
                if ((lane >> 2) == Tn - 1) {   // measurement rows of the same (diagonal) tile: define the mirrored blocks
                    double2* Tm = Kt + tile_index(Tn - 1, Tn - 1) * 32 + 8 * (lane & 3) + (ia & 3);
                    Tm[0] = Tm[1] = Tm[4] = Tm[5] = make_double2(0.0, 0.0);
                }
            }

            // ---- K = H_x P H_x^T + sigma^2 I, one lower 2x2 block per lane per pass ----
            const int npairs = M * (M + 1) / 2;     // measurement rows only; the few padding blocks follow below
            // Pair p = i (i + 1) / 2 + k and its window slots.  Under load every dependent shared-memory / shuffle access
            // costs a full trip through the saturated memory-input queue, so the descriptor is pure arithmetic whenever
            // the track's slots are consecutive (slot[i] = slot[0] + i: the normal case, a track is observed from every
            // camera since its first one); otherwise the slots come from the lanes that hold them.
            const int s0 = __shfl_sync(0xffffffffu, o.slot, 0);
            const bool affine = __all_sync(0xffffffffu, !live || o.slot == s0 + lane);
#pragma unroll 1
            for (int p0 = 0; p0 < npairs; p0 += 32) {
                const int p = p0 + lane;
                const bool act = p < npairs;
                int i = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
                i -= (i * (i + 1) / 2 > p) ? 1 : 0;
                i += ((i + 1) * (i + 2) / 2 <= p) ? 1 : 0;
                const int k = p - i * (i + 1) / 2;
                int si, sk;
                if (affine) {
                    si = s0 + i;
                    sk = s0 + k;
                } else {
                    si = __shfl_sync(0xffffffffu, o.slot, i & 31);
                    sk = __shfl_sync(0xffffffffu, o.slot, k & 31);
                }
                if (!act) continue;
                double o00 = 0.0, o01 = 0.0, o10 = 0.0, o11 = 0.0;
                {
                    double ai[2][6];
                    {
                        const double2* Ai = reinterpret_cast<const double2*>(wsA + i * AR);
#pragma unroll
                        for (int c2 = 0; c2 < 6; ++c2) {
                            const double2 x = Ai[c2];
                            ai[c2 / 3][2 * (c2 % 3)] = x.x;
                            ai[c2 / 3][2 * (c2 % 3) + 1] = x.y;
                        }
                    }
                    double v0[6], v1[6];
#pragma unroll
                    for (int c = 0; c < 6; ++c) { v0[c] = 0.0; v1[c] = 0.0; }
                    if (si >= sk) {
                        const double2* Pb = reinterpret_cast<const double2*>(sP + (size_t)tile_index(si, sk) * PB);
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
                        const double2* Pb = reinterpret_cast<const double2*>(sP + (size_t)tile_index(sk, si) * PB);
#pragma unroll
                        for (int c = 0; c < 6; ++c) {
#pragma unroll
                            for (int r2 = 0; r2 < 3; ++r2) {
                                const double2 pv = Pb[c * 3 + r2];
                                v0[c] = fma(ai[0][2 * r2 + 1], pv.y, fma(ai[0][2 * r2], pv.x, v0[c]));
                                v1[c] = fma(ai[1][2 * r2 + 1], pv.y, fma(ai[1][2 * r2], pv.x, v1[c]));
                            }
                        }
                    }
                    const double2* Ak = reinterpret_cast<const double2*>(wsA + k * AR);
#pragma unroll
                    for (int c2 = 0; c2 < 3; ++c2) {
                        const double2 x0 = Ak[c2], x1 = Ak[3 + c2];
                        o00 = fma(v0[2 * c2 + 1], x0.y, fma(v0[2 * c2], x0.x, o00));
                        o01 = fma(v0[2 * c2 + 1], x1.y, fma(v0[2 * c2], x1.x, o01));
                        o10 = fma(v1[2 * c2 + 1], x0.y, fma(v1[2 * c2], x0.x, o10));
                        o11 = fma(v1[2 * c2 + 1], x1.y, fma(v1[2 * c2], x1.x, o11));
                    }
                    if (i == k) { o00 += s2; o11 += s2; }
                }
                // rows 2 i, 2 i + 1 / columns 2 k, 2 k + 1  ->  tile (i / 4, k / 4), lanes 4 (2 (i & 3) + row) + (k & 3)
                const int I = i >> 2, J = k >> 2;
                double2* T = Kt + tile_index(I, J) * 32;
                const int l0 = 8 * (i & 3) + (k & 3);
                T[l0] = make_double2(o00, o01);
                T[l0 + 4] = make_double2(o10, o11);
                if (I == J && i != k) {   // the mirrored block of a diagonal tile is never read as data: keep it defined
                    const int l1 = 8 * (k & 3) + (i & 3);
                    T[l1] = make_double2(0.0, 0.0);
                    T[l1 + 4] = make_double2(0.0, 0.0);
                }
            }
            // ---- the padding row pairs i = M .. Mp-1 (at most 7): identity before the appended rows, zero after; the
            //      blocks (appended rows x measurement columns) were written above ----
            for (int i = M; i < Mp; ++i) {
                const bool appended = i >= ia && i < ia + 2;
                for (int k = (appended ? M : 0) + lane; k <= i; k += 32) {
                    const double dg = (i == k && i < ia) ? 1.0 : 0.0;
                    const int I = i >> 2, J = k >> 2;
                    double2* T = Kt + tile_index(I, J) * 32;
                    const int l0 = 8 * (i & 3) + (k & 3);
                    T[l0] = make_double2(dg, 0.0);
                    T[l0 + 4] = make_double2(0.0, dg);
                    if (I == J && i != k) {
                        const int l1 = 8 * (k & 3) + (i & 3);
                        T[l1] = make_double2(0.0, 0.0);
                        T[l1 + 4] = make_double2(0.0, 0.0);
                    }
                }
            }
            if (lane == 0) pend_list[pend_pos] = fl;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// kernel B: register-resident blocked Cholesky of every pending feature's K + the gate
// ---------------------------------------------------------------------------------------------------------------------
// In-tile factorisation of an 8x8 diagonal tile held in accumulator-fragment layout (D0, D1): two columns per pass so the
// register parity is static; column entries of finished rows are zeroed at the source lane, which makes every rank-1
// update unconditional.  Returns the B fragments of X = L^-T for the panel solve: B[slot t][n = g] = X[2 t + step][g].
__device__ __forceinline__ void diag_tile_inverse(double D0, double D1, const int g, const int t, double& bf0, double& bf1) {
    double x0 = (g == 2 * t) ? 1.0 : 0.0, x1 = (g == 2 * t + 1) ? 1.0 : 0.0;
#pragma unroll 1
    for (int tk = 0; tk < 4; ++tk) {
        {
            const int k = 2 * tk;
            const double piv = shfl_d(D0, k * 4 + tk);
            const double rs = fast_rsqrt_spd(piv);
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
            const double rs = fast_rsqrt_spd(piv);
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
    const int src0 = (2 * t) * 4 + (g >> 1), src1 = (2 * t + 1) * 4 + (g >> 1);
    const double v00 = shfl_d(x0, src0), v01 = shfl_d(x1, src0);
    const double v10 = shfl_d(x0, src1), v11 = shfl_d(x1, src1);
    bf0 = (g & 1) ? v01 : v00;
    bf1 = (g & 1) ? v11 : v10;
}

// the first `npass` column pairs of a diagonal tile, no inverse (last tile row: the columns left of the appended rows)
__device__ __forceinline__ void diag_tile_partial(double& D0, double& D1, const int g, const int t, const int npass) {
#pragma unroll 1
    for (int tk = 0; tk < npass; ++tk) {
        {
            const int k = 2 * tk;
            const double rs = fast_rsqrt_spd(shfl_d(D0, k * 4 + tk));
            const double colv = (g > k) ? D0 * rs : 0.0;
            const double lg = shfl_d(colv, g * 4 + tk);
            const double lc0 = shfl_d(colv, (2 * t) * 4 + tk);
            const double lc1 = shfl_d(colv, (2 * t + 1) * 4 + tk);
            D0 -= lg * lc0;
            D1 -= lg * lc1;
        }
        {
            const int k = 2 * tk + 1;
            const double rs = fast_rsqrt_spd(shfl_d(D1, k * 4 + tk));
            const double colv = (g > k) ? D1 * rs : 0.0;
            const double lg = shfl_d(colv, g * 4 + tk);
            const double lc0 = shfl_d(colv, (2 * t) * 4 + tk);
            const double lc1 = shfl_d(colv, (2 * t + 1) * 4 + tk);
            D0 -= lg * lc0;
            D1 -= lg * lc1;
        }
    }
}

struct CholRegs {
    double L[TC][TC][2];   // finished tiles L[I][c], I > c (statically indexed; live only while later columns need them)
    double Cn[TC][2];      // prefetched tiles of the next block column (rows I >= column)
};

// loads block column BK (static position; the feature's own tile column is BK - off) into the prefetch registers
template <int BK, int off>
__device__ __forceinline__ void load_column(CholRegs& r, const double2* __restrict__ Kt) {
    constexpr int J = BK - off;
#pragma unroll
    for (int I = BK; I < TC; ++I) {
        const double2 v = __ldcs(Kt + ((I - off) * (I - off + 1) / 2 + J) * 32);
        r.Cn[I][0] = v.x;
        r.Cn[I][1] = v.y;
    }
}

// Left-looking block column BK: C[I] = K[I][BK] - sum_{c < BK} L[I][c] L[BK][c]^T, then (BK < TC - 1) factor the diagonal
// tile and solve the panel.  Only BK (9 - BK) + (9 - BK) tiles are live (<= 25 of 45), and the next column's tiles are in
// flight while this one is computed.  The last block column (BK = TC - 1) is the tile that holds the appended rows: its
// updated diagonal tile is returned in (D0, D1).
template <int BK, int off>
__device__ __forceinline__ void chol_block_column(CholRegs& r, const double2* __restrict__ Kt, const int g, const int t,
                                                  double& D0, double& D1) {
    if (BK < off) return;
    double C[TC][2];
#pragma unroll
    for (int I = BK; I < TC; ++I) { C[I][0] = r.Cn[I][0]; C[I][1] = r.Cn[I][1]; }
    if (BK + 1 < TC) load_column<(BK + 1 < TC) ? BK + 1 : BK, off>(r, Kt);
#pragma unroll
    for (int c = 0; c < BK; ++c) {
        if (c >= off) {   // columns left of the feature's first one were never produced
            const double nb0 = -r.L[BK][c][0], nb1 = -r.L[BK][c][1];
#pragma unroll
            for (int I = BK; I < TC; ++I) {
                dense::dmma(C[I][0], C[I][1], r.L[I][c][0], nb0);
                dense::dmma(C[I][0], C[I][1], r.L[I][c][1], nb1);
            }
        }
    }
    if (BK == TC - 1) {
        D0 = C[BK][0];
        D1 = C[BK][1];
        return;
    }
    double bf0, bf1;
    diag_tile_inverse(C[BK][0], C[BK][1], g, t, bf0, bf1);
#pragma unroll
    for (int I = BK + 1; I < TC; ++I) {
        double n0 = 0.0, n1 = 0.0;
        dense::dmma(n0, n1, C[I][0], bf0);
        dense::dmma(n0, n1, C[I][1], bf1);
        r.L[I][BK][0] = n0;
        r.L[I][BK][1] = n1;
    }
}

// the whole factorisation of a feature whose first tile column sits at the static position OFF = 9 - (tile rows)
template <int OFF>
__device__ __forceinline__ void chol_feature(const double2* __restrict__ Kt, const int g, const int t, double& D0, double& D1) {
    CholRegs r;
    load_column<OFF, OFF>(r, Kt);
    chol_block_column<0, OFF>(r, Kt, g, t, D0, D1);
    chol_block_column<1, OFF>(r, Kt, g, t, D0, D1);
    chol_block_column<2, OFF>(r, Kt, g, t, D0, D1);
    chol_block_column<3, OFF>(r, Kt, g, t, D0, D1);
    chol_block_column<4, OFF>(r, Kt, g, t, D0, D1);
    chol_block_column<5, OFF>(r, Kt, g, t, D0, D1);
    chol_block_column<6, OFF>(r, Kt, g, t, D0, D1);
    chol_block_column<7, OFF>(r, Kt, g, t, D0, D1);
    chol_block_column<8, OFF>(r, Kt, g, t, D0, D1);
}

__global__ void __launch_bounds__(FC_THREADS, 4)
front_chol_kernel(const double2* __restrict__ Kbuf, const int* __restrict__ pend_list, const int* __restrict__ pend_count,
                  const int* __restrict__ feat_len, const int* __restrict__ fac_rank, const int* __restrict__ dof,
                  const double* __restrict__ chi2_table, const int chi2_len, unsigned char* __restrict__ feat_status,
                  double* __restrict__ gamma, const size_t fbase) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int count = *pend_count;
    const int nwarps = gridDim.x * FC_WARPS;
    for (int q = blockIdx.x * FC_WARPS + (threadIdx.x >> 5); q < count; q += nwarps) {
        const int fl = pend_list[q];
        const size_t fj = fbase + (size_t)fl;
        const int n2 = 2 * feat_len[fj];
        const int a4 = (n2 + 3) & ~3;                       // first appended row (front_asm_kernel)
        const int Tn = (a4 + 4 + 7) >> 3, off = TC - Tn;
        const int rr = a4 - 8 * (Tn - 1);                   // its local row in the last tile row: 0 or 4
        const double2* Kt = Kbuf + (size_t)fl * KT_D2 + lane;
        // tile (I, J) of the feature sits at the static position (I + off, J + off): the factorisation then ends at the
        // bottom-right corner for every track length and every register index is static
        double D0 = 0.0, D1 = 0.0;
        switch (off) {
            case 0: chol_feature<0>(Kt, g, t, D0, D1); break;
            case 1: chol_feature<1>(Kt, g, t, D0, D1); break;
            case 2: chol_feature<2>(Kt, g, t, D0, D1); break;
            case 3: chol_feature<3>(Kt, g, t, D0, D1); break;
            case 4: chol_feature<4>(Kt, g, t, D0, D1); break;
            case 5: chol_feature<5>(Kt, g, t, D0, D1); break;
            case 6: chol_feature<6>(Kt, g, t, D0, D1); break;
            case 7: chol_feature<7>(Kt, g, t, D0, D1); break;
            default: chol_feature<8>(Kt, g, t, D0, D1); break;
        }
        // the measurement columns of the last tile row, then its 4x4 block at (rr, rr) holds MINUS the Gram matrix of
        // [y; V] = [r~; Q1^T] L^-T:  yy = (0,0), b_q = (q,0), M_qp = (q,p), q >= p (lower triangle of the tile)
        diag_tile_partial(D0, D1, g, t, rr >> 1);
        const int l0 = 4 * rr + (rr >> 1);                  // lane of element (rr, rr)
        const double yy = -shfl_d(D0, l0);
        double bq[3], Mq[6];
        bq[0] = -shfl_d(D0, l0 + 4); bq[1] = -shfl_d(D0, l0 + 8); bq[2] = -shfl_d(D0, l0 + 12);
        Mq[0] = -shfl_d(D1, l0 + 4);                               // (1,1)
        Mq[1] = -shfl_d(D1, l0 + 8);                               // (2,1)
        Mq[2] = -shfl_d(D1, l0 + 12);                              // (3,1)
        Mq[3] = -shfl_d(D0, l0 + 9);                               // (2,2)
        Mq[4] = -shfl_d(D0, l0 + 13);                              // (3,2)
        Mq[5] = -shfl_d(D1, l0 + 13);                              // (3,3)
        const int rank = fac_rank[fj], dofv = dof[fj];
        const int st = feat_status[fj];
        double gam = yy;
        {   // Cholesky solve on the leading rank x rank block of V V^T (reciprocal square roots: 0 for a non-positive pivot)
            const double r0 = (rank >= 1) ? fast_rsqrt(Mq[0]) : 0.0;
            const double z0 = bq[0] * r0;
            const double l10 = Mq[1] * r0, l20 = Mq[2] * r0;
            const double r1 = (rank >= 2 && r0 > 0.0) ? fast_rsqrt(Mq[3] - l10 * l10) : 0.0;
            const double z1 = (bq[1] - l10 * z0) * r1;
            const double l21 = (Mq[4] - l20 * l10) * r1;
            const double r2 = (rank >= 3 && r1 > 0.0) ? fast_rsqrt(Mq[5] - l20 * l20 - l21 * l21) : 0.0;
            const double z2 = (bq[2] - l20 * z0 - l21 * z1) * r2;
            gam = yy - (z0 * z0 + z1 * z1 + z2 * z2);
        }
        if (lane == 0) {
            const double crit = (dofv < chi2_len) ? chi2_table[dofv] : nan("");
            const bool pass = gam <= crit;
            feat_status[fj] = (unsigned char)((pass ? MSCKF_FEAT_ACCEPTED : MSCKF_FEAT_GATED_OUT) | (st & FLAG_KEEP));
            gamma[fj] = gam;
        }
    }
}

size_t front_asm_smem(const msckf_dims& d, int warps = FA_WARPS) {
    return ((size_t)(d.N * (d.N + 1) / 2) * PB + (size_t)d.N * 12 + (size_t)warps * FW_MAX_OBS * AR) * sizeof(double);
}

}  // namespace

bool front_split_supported(const msckf_dims& d) { return d.M <= fgeom::FW_MAX_OBS && front_asm_smem(d) <= 227 * 1024; }

// filters handled per assembly / factorisation kernel pair: bounds the tile buffer (20 KB per feature)
int front_split_sub_filters(const msckf_dims& d) {
    const size_t budget = (size_t)6 << 30;
    const size_t per_filter = (size_t)(d.F > 0 ? d.F : 1) * KT_D2 * sizeof(double2);
    long long sub = (long long)(budget / per_filter);
    if (sub < 1) sub = 1;
    if (sub > d.B) sub = d.B;
    const int n_sub = (int)((d.B + sub - 1) / sub);
    return (d.B + n_sub - 1) / n_sub;
}

size_t front_split_tile_bytes(const msckf_dims& d) {
    return (size_t)front_split_sub_filters(d) * d.F * KT_D2 * sizeof(double2);
}

// optional per-kernel timing (msckf_set_stage_timing / msckf_front_stage_ms): CUDA events on the launching stream around
// the two kernels of every sub-batch
namespace {
constexpr int MAX_SUB = 64;
bool g_timing = false;
cudaEvent_t g_ev[MAX_SUB][3];
int g_ev_created = 0, g_ev_used = 0;
}  // namespace

extern "C" int msckf_set_stage_timing(int on) {
    g_timing = on != 0;
    g_ev_used = 0;
    return 0;
}

extern "C" int msckf_front_stage_ms(double* asm_ms, double* chol_ms) {
    double a = 0.0, c = 0.0;
    for (int k = 0; k < g_ev_used; ++k) {
        float t0 = 0.f, t1 = 0.f;
        cudaError_t e = cudaEventSynchronize(g_ev[k][2]);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t0, g_ev[k][0], g_ev[k][1]);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t1, g_ev[k][1], g_ev[k][2]);
        if (e != cudaSuccess) return (int)e;
        a += t0;
        c += t1;
    }
    if (asm_ms) *asm_ms = a;
    if (chol_ms) *chol_ms = c;
    return g_ev_used;
}

cudaError_t launch_front_split(const FrontArgs& a, double2* Kbuf, int* pend_list, int* pend_count, cudaStream_t s) {
    const size_t smem = front_asm_smem(a.dims);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(front_asm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const int sub = front_split_sub_filters(a.dims);
    const int F = a.dims.F;
    if (g_timing) g_ev_used = 0;
    for (int b0 = 0; b0 < a.dims.B; b0 += sub) {
        const int nb = (b0 + sub <= a.dims.B) ? sub : a.dims.B - b0;
        cudaError_t e = cudaMemsetAsync(pend_count, 0, sizeof(int), s);
        if (e != cudaSuccess) return e;
        const bool timed = g_timing && g_ev_used < MAX_SUB;
        const int slot = g_ev_used;
        if (timed) {
            for (; g_ev_created <= slot; ++g_ev_created)
                for (int q = 0; q < 3; ++q)
                    if ((e = cudaEventCreate(&g_ev[g_ev_created][q])) != cudaSuccess) return e;
            cudaEventRecord(g_ev[slot][0], s);
        }
        // feature chunks per filter so that the CTAs come in >= ~6 waves of 148 (each warp keeps >= 2 features): long CTAs
        // amortise the staggered start of their warps (measured: 13.3 ms -> 12.0 ms per 4096 filters with both)
        constexpr int waves = 6;
        constexpr unsigned stagger_ns = 1500u;
        int chunks = (waves * 148 + nb - 1) / nb;
        int fpc = (F + chunks - 1) / chunks;
        if (fpc < 2 * FA_WARPS) fpc = 2 * FA_WARPS;
        chunks = (F + fpc - 1) / fpc;
        dim3 grid(chunks, nb);
        front_asm_kernel<<<grid, FA_THREADS, smem, s>>>(a, Kbuf, pend_list, pend_count, b0, fpc, stagger_ns);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (timed) cudaEventRecord(g_ev[slot][1], s);
        front_chol_kernel<<<148 * 4, FC_THREADS, 0, s>>>(Kbuf, pend_list, pend_count, a.feat_len, a.fac_rank, a.dof,
                                                        a.chi2_table, a.chi2_len, a.feat_status, a.gamma, (size_t)b0 * F);
        msckf_count_launch(2);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (timed) {
            cudaEventRecord(g_ev[slot][2], s);
            ++g_ev_used;
        }
    }
    return cudaSuccess;
}
