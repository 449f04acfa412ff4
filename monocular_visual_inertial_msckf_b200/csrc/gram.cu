// Row 9 of SURVEY.md section 8a, information-form flavour (MSCKF_COMPRESS_GRAM): instead of "vstack the surviving H_o
// blocks and QR them" (reference src/msckf/MSCKF.py:587-598) the kernel accumulates, per filter, the information matrix
// and vector of the same stacked system,
//     G = sum_j H_o,j^T H_o,j = sum_j (D_j - W_j^T W_j),        g = sum_j H_o,j^T r_o,j = sum_j H_x,j^T r~_j,
// with D_j = blockdiag_slot(A_i^T A_i) (6 x 6 per observing camera), W_j = Q1_j^T H_x,j (3 x 6N, rank(H_f) rows) and
// r~_j = (I - Q1 Q1^T) r_j -- the basis-free form of the left-null-space projection (A A^T = I - Q1 Q1^T).  G is what
// T_H^T T_H of the reference's QR equals, g is T_H^T r_n; the update kernel (update_info.cu) consumes them directly.
//
// One 512-thread CTA per filter, a software pipeline over panels of PF accepted features with ONE CTA barrier per panel:
//   * cp.async (LDGSTS) prefetches the compact factors (A_i, Q1 rows, r~_i, slots) of panel p + 2 into one of two staging
//     buffers; the warp that fetched a feature turns its slot list into a slot -> observation map when its copies land;
//   * the 3 PF rows of W of panel p + 1 (zero where a camera does not observe the feature) are written into one of two
//     shared K-panels, one (feature, slot, row) unit per thread and round, and D / g of that panel are accumulated on the
//     tensor pipe straight from the staging buffer (operand A == operand B: the tile [A_i | r~_i | 0] of a window slot);
//   * every warp folds panel p into its share of the lower-triangular 8 x 8 output tiles of W^T W with FP64 tensor
//     instructions (DMMA.8x8x4), accumulators resident in registers for the whole filter (20 tiles per warp);
//   * two of the four warps of every scheduler take the tensor loop first and the scatter second, the other two the other
//     way round, so the tensor pipe is fed while half of the CTA moves operands.
// The stacked matrix H_X (22.8 MB per filter at 256 features x 30 poses) is never formed.
#include "msckf_common.cuh"
#include "dense.cuh"

namespace {

constexpr int GT = 512;        // threads per CTA (16 warps, 4 per SM sub-partition; <= 128 registers per thread)
constexpr int GW = GT / 32;
constexpr int UG = 5;          // units per branch-free group of the tensor loop
constexpr int UMAX = 10;       // (2 tile rows x 1 tile column) units per warp per pass -> 20 accumulator tiles

__device__ __forceinline__ void cp_async16(void* s, const void* g) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g));
}
__device__ __forceinline__ void cp_async4(void* s, const void* g) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(g));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// shared-memory fragment load at a 32-bit shared address plus a compile-time byte offset.  Volatile on purpose: the
// loads of a group of units must be ISSUED before the group's tensor instructions (the compiler otherwise sinks each
// load to just before its use and the warp stalls on every one of them).
template <int IMM>
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(IMM));
    return v;
}

__device__ __forceinline__ int strip_len(int p, int T) { return min(2 * p + 2, T); }

// Panel leading dimensions are = 4 mod 16: the DMMA fragment loads [k = t][x = g] of a half-warp then touch 16 distinct
// 8-byte bank pairs (t * 4 + g); the r01k profile of the previous layout (= 8 mod 16) showed 42% conflicted wavefronts.

// Folds one k-panel (KW rows) into this warp's accumulator tiles.  wbase = shared address of panel element [t][g].
template <int KW, int LDW, int UM = UMAX, int UGR = UG, int KK = 0>
__device__ __forceinline__ void gram_tensor_loop(unsigned wbase, int p0, int J0, int nu, unsigned newstrip,
                                                 double (&acc)[UM][2][2]) {
    if constexpr (KK < KW) {
        constexpr int KOFF = KK * LDW * 8;
        unsigned abase = wbase + 128u * p0;
        unsigned offb = 64u * J0;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int u0g = 0; u0g < UM; u0g += UGR) {
            if (u0g < nu) {
                double bf[UGR];
#pragma unroll
                for (int v = 0; v < UGR; ++v) {
                    const int u = u0g + v;
                    if (u < UM) {                        // (the last group may be short)
                        if (u > 0 && ((newstrip >> u) & 1u)) offb = 0;
                        bf[v] = lds_f64<KOFF>(wbase + offb);
                        offb += 64u;
                    }
                }
#pragma unroll
                for (int v = 0; v < UGR; ++v) {
                    const int u = u0g + v;
                    if (u < UM) {
                        constexpr int UL = UM - 1;
                        const int uu = u < UM ? u : UL;
                        if ((newstrip >> u) & 1u) {
                            a0 = lds_f64<KOFF>(abase);
                            a1 = lds_f64<KOFF + 64>(abase);
                            abase += 128u;
                        }
                        dense::dmma(acc[uu][0][0], acc[uu][0][1], a0, bf[v]);
                        dense::dmma(acc[uu][1][0], acc[uu][1][1], a1, bf[v]);
                    }
                }
            }
        }
        gram_tensor_loop<KW, LDW, UM, UGR, KK + 4>(wbase, p0, J0, nu, newstrip, acc);
    }
}

// NMAX = pose capacity of this instantiation (32 or 64): fixes the panels' leading dimensions at compile time, so the
// k step of every operand load is an immediate offset.  PF = features per panel (8, or 4 when the tracks are long).
template <int NMAX, int PF>
__global__ void __launch_bounds__(GT, 1) gram_kernel(const CompressArgs a) {
    constexpr int SMAX = NMAX / GW;                     // window slots per warp for the D / g tiles
    constexpr int LDW = 6 * NMAX + 4;                   // mod 16 == 4 (see below)
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int F = a.dims.F, Mcap = a.dims.M, Ncap = a.dims.N;
    const int nc = a.n_cams[b];
    const int c = 6 * nc;
    const int ccap = 6 * Ncap;
    const int T = (c + 7) >> 3;            // 8 x 8 output tiles per side
    const int T2 = (T + 1) >> 1;           // 2-row strips

    // shared-memory carve-up (offsets, not pointers: a table of generic pointers would pin two dozen registers)
    //   panelW[2] [3 PF][LDW]  rows of W = Q1^T H_x of a panel's features (k-panel of W^T W), double-buffered
    //   stage[2]               prefetched factors of a panel: A [PF][Mcap][12], Q [PF][Mcap][6], r~ [PF][Mcap][2], slot [PF][Mcap]
    //   Dg [Ncap][48]          D_s (6 x 6, row stride 8; column 6 = g_s)
    //   acc_list [F] (int), inv[2] [PF][NMAX] (bytes: observation index held by a window slot, 0xFF = none)
    extern __shared__ double smem[];
    const int PM = PF * Mcap;
    constexpr int WSZ = 3 * PF * LDW;
    auto rawA = [&](int st) { return smem + 2 * WSZ + (size_t)st * PM * 21; };
    auto rawQ = [&](int st) { return rawA(st) + PM * 12; };
    auto rawR = [&](int st) { return rawA(st) + PM * 18; };
    auto rawS = [&](int st) { return reinterpret_cast<int*>(rawA(st) + PM * 20); };
    double* const Dg = rawA(2);
    int* const acc_list = reinterpret_cast<int*>(Dg + Ncap * 48);
    unsigned char* const inv = reinterpret_cast<unsigned char*>(acc_list + F);
    __shared__ int s_wcnt[GW];
    __shared__ int s_tot[2];

    // ---- ordered list of the features that passed the gate (stacking order = feature order, MSCKF.py:573-588) ----
    if (tid < 2) s_tot[tid] = 0;
    __syncthreads();
    int nacc = 0, nrej_local = 0;
    for (int base = 0; base < F; base += GT) {
        const int j = base + tid;
        const int st = (j < F) ? (a.feat_status[(size_t)b * F + j] & 0xF) : 0;
        const bool ok = st == MSCKF_FEAT_ACCEPTED;
        nrej_local += (st == MSCKF_FEAT_GATED_OUT);
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = nacc;
        for (int w = 0; w < warp; ++w) off += s_wcnt[w];
        if (ok) acc_list[off + __popc(bal & ((1u << lane) - 1u))] = j;
        for (int w = 0; w < GW; ++w) nacc += s_wcnt[w];
        __syncthreads();
    }
    if (nrej_local) atomicAdd(&s_tot[1], nrej_local);
    __syncthreads();
    if (tid == 0) {
        a.n_accepted[b] = nacc;
        a.n_rejected[b] = s_tot[1];
    }
    double* Gout = a.T_H + (size_t)b * ccap * ccap;
    double* gout = a.r_n + (size_t)b * ccap;
    if (nacc == 0) {   // nothing survived: the update kernel skips this filter (MSCKF.py:584-585)
        for (int e = tid; e < ccap; e += GT) gout[e] = 0.0;
        return;
    }

    const int npanels = (nacc + PF - 1) / PF;
    int U = 0;
    for (int p = 0; p < T2; ++p) U += strip_len(p, T);
    const int npass = (U + GW * UMAX - 1) / (GW * UMAX);
    const int upp = (U + npass - 1) / npass;
    constexpr int KW = 3 * PF, KX = 2 * PF;      // k extents of the two panels (multiples of 4)

    // asynchronous prefetch of one panel's compact factors: warp f fetches feature f (and remembers its length)
    int my_len = 0;
    auto prefetch = [&](int pi, int st) {
        if (warp < PF) {
            const int f = warp;
            const int k = pi * PF + f;
            const int j = (k < nacc) ? acc_list[k] : -1;
            const int M = (j >= 0) ? min(a.feat_len[(size_t)b * F + j], Mcap) : 0;
            my_len = M;
            if (M > 0) {
                const size_t fj = (size_t)b * F + j;
                const double* gA = a.fac_A + fj * (size_t)Mcap * 12;
                const double* gQ = a.fac_Q + fj * (size_t)Mcap * 6;
                const double* gR = a.fac_rt + fj * (size_t)Mcap * 2;
                const int* gS = a.obs_slot + fj * (size_t)Mcap;
                double* sA = rawA(st) + (size_t)f * Mcap * 12;
                double* sQ = rawQ(st) + (size_t)f * Mcap * 6;
                double* sR = rawR(st) + (size_t)f * Mcap * 2;
                int* sS = rawS(st) + f * Mcap;
                for (int it = lane; it < 6 * M; it += 32) cp_async16(sA + 2 * it, gA + 2 * it);
                for (int it = lane; it < 3 * M; it += 32) cp_async16(sQ + 2 * it, gQ + 2 * it);
                for (int it = lane; it < M; it += 32) {
                    cp_async16(sR + 2 * it, gR + 2 * it);
                    cp_async4(sS + it, gS + it);
                }
            }
        }
        cp_async_commit();
    };
    // once its copies have landed, the fetching warp inverts the slot list of its feature: inv[slot] = observation
    auto land = [&](int st) {
        cp_async_wait_all();
        if (warp < PF) {
            __syncwarp();
            unsigned char* iv = inv + (st * PF + warp) * NMAX;
            for (int sl = lane; sl < NMAX; sl += 32) iv[sl] = 0xFF;
            __syncwarp();
            const int* sS = rawS(st) + warp * Mcap;
            for (int i = lane; i < my_len; i += 32) {
                const int sl = sS[i];
                if (sl >= 0 && sl < nc) iv[sl] = (unsigned char)i;
            }
        }
    };
    constexpr int WPF = GW / PF;                 // warps that share a feature in the scatter
    static_assert(GW % PF == 0, "the scatter assigns whole warps to features");

    for (int pass = 0; pass < npass; ++pass) {
        // this warp's contiguous range of units in the linearised (strip, column) order; the per-unit operand offsets
        // are the same for every k step of every panel, so they are worked out once, into registers
        const int ubeg = pass * upp, uend = min(U, ubeg + upp);
        const int upw = (uend - ubeg + GW - 1) / GW;
        const int u0 = min(uend, ubeg + warp * upw);
        const int nu = min(uend, u0 + upw) - u0;
        int p0 = 0, J0 = u0;
        while (p0 < T2 && J0 >= strip_len(p0, T)) { J0 -= strip_len(p0, T); ++p0; }
        unsigned newstrip = 0;       // bit u: unit u is the first of a strip (reload the two A fragments, column 0)
        {
            int p = p0, J = J0;
#pragma unroll
            for (int u = 0; u < UMAX; ++u) {
                if (u == 0 || J == 0) newstrip |= 1u << u;
                if (u < nu && ++J == strip_len(p, T)) { J = 0; ++p; }
            }
        }
        const bool doX = pass == 0;

        double acc[UMAX][2][2];
#pragma unroll
        for (int u = 0; u < UMAX; ++u) acc[u][0][0] = acc[u][0][1] = acc[u][1][0] = acc[u][1][1] = 0.0;
        double accX[SMAX][2];
#pragma unroll
        for (int k = 0; k < SMAX; ++k) accX[k][0] = accX[k][1] = 0.0;

        // ---- rows of W of one panel: W[3f+q][6 s + e] = Q1[2i][q] A[2i][e] + Q1[2i+1][q] A[2i+1][e], i = the observation of
        //      feature f in window slot s (zero row segment when there is none); one (slot, row) unit per thread and round
        auto scatter = [&](int st) {
            const int f = warp % PF;
            const unsigned char* iv = inv + (st * PF + f) * NMAX;
            double* Wp = smem + (size_t)st * WSZ + (size_t)(3 * f) * LDW;
            for (int e = (warp / PF) * 32 + lane; e < 3 * nc; e += 32 * WPF) {
                const int sl = e / 3, q = e - 3 * sl;
                const int i = iv[sl];
                double2* dst = reinterpret_cast<double2*>(Wp + (size_t)q * LDW + 6 * sl);
                if (i == 0xFF) {
                    dst[0] = dst[1] = dst[2] = make_double2(0.0, 0.0);
                } else {
                    const double2* A2 = reinterpret_cast<const double2*>(rawA(st) + ((size_t)f * Mcap + i) * 12);
                    const double* Q = rawQ(st) + ((size_t)f * Mcap + i) * 6;
                    const double q0 = Q[q], q1 = Q[3 + q];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double2 x = A2[k], y = A2[3 + k];
                        dst[k] = make_double2(q0 * x.x + q1 * y.x, q0 * x.y + q1 * y.y);
                    }
                }
            }
        };
        // ---- D_s = sum A_i^T A_i and g_s = sum A_i^T r~_i: one 8 x 8 tile per window slot, operand A == operand B; the
        //      operand tile [A_i | r~_i | 0] (rows = the 2 PF measurement rows of the panel) is gathered from the stage
        auto fold_X = [&](int st) {
            const double* sA = rawA(st);
            const double* sR = rawR(st);
            const unsigned char* iv = inv + st * PF * NMAX;
#pragma unroll
            for (int kk = 0; kk < KX; kk += 4) {
                const int f = (kk >> 1) + (t >> 1), r = t & 1;
#pragma unroll
                for (int k = 0; k < SMAX; ++k) {
                    const int sl = warp + GW * k;
                    if (sl < nc) {
                        const int i = iv[f * NMAX + sl];
                        double x = 0.0;
                        if (i != 0xFF && g < 7) {
                            const size_t o = (size_t)f * Mcap + i;
                            x = (g < 6) ? sA[o * 12 + 6 * r + g] : sR[o * 2 + r];
                        }
                        dense::dmma(accX[k][0], accX[k][1], x, x);
                    }
                }
            }
        };
        // ---- W^T W on the tensor pipe: per k step, the B fragments of a group of units are loaded first, then the tensor
        //      instructions are issued back to back (no branch per unit; units beyond nu work on dead tiles)
        auto fold_W = [&](int st) {
            gram_tensor_loop<KW, LDW>((unsigned)__cvta_generic_to_shared(smem + (size_t)st * WSZ) + (unsigned)(t * LDW + g) * 8u,
                                      p0, J0, nu, newstrip, acc);
        };

        __syncthreads();                 // (the previous pass is done with every buffer)
        prefetch(0, 0);
        land(0);
        __syncthreads();
        if (npanels > 1) prefetch(1, 1);
        scatter(0);
        if (doX) fold_X(0);
        land(1);
        __syncthreads();

        const bool tensor_first = (warp >> 2) & 1;
        for (int pi = 0; pi < npanels; ++pi) {
            const int st = pi & 1;
            // stage st was consumed (scatter, D / g) during the previous iteration: refill it with panel pi + 2
            if (pi + 2 < npanels) prefetch(pi + 2, st);
#pragma unroll 1
            for (int ph = 0; ph < 2; ++ph) {
#ifndef GRAM_DIAG
#define GRAM_DIAG 0
#endif
                if ((ph == 0) == tensor_first) {
                    if (GRAM_DIAG != 2) fold_W(st);
                } else if (pi + 1 < npanels) {
                    if (GRAM_DIAG != 1 && GRAM_DIAG != 3) scatter(st ^ 1);
                    if (doX && GRAM_DIAG != 1 && GRAM_DIAG != 4) fold_X(st ^ 1);
                }
            }
            land(st);
            __syncthreads();
        }

        if (doX) {
#pragma unroll
            for (int k = 0; k < SMAX; ++k) {
                const int s = warp + GW * k;
                if (s < nc && g < 6) {   // rows 6, 7 of the tile are by-products (r~^T A, 0)
                    Dg[s * 48 + g * 8 + 2 * t] = accX[k][0];
                    Dg[s * 48 + g * 8 + 2 * t + 1] = accX[k][1];
                }
            }
            __syncthreads();
            for (int e = tid; e < ccap; e += GT) {
                const int s = e / 6, aa = e - 6 * s;
                gout[e] = (s < nc) ? Dg[s * 48 + aa * 8 + 6] : 0.0;
            }
        }
        // ---- G = D - W^T W: lower tiles written with their mirror image ----
        {
            int p = p0, J = J0;
#pragma unroll
            for (int u = 0; u < UMAX; ++u) {
                if (u < nu) {
                    if (u > 0 && ((newstrip >> u) & 1u)) { ++p; J = 0; }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int I = 2 * p + h;
                        if (I < T && J <= I) {
                            const int r = 8 * I + g;
#pragma unroll
                            for (int x = 0; x < 2; ++x) {
                                const int col = 8 * J + 2 * t + x;
                                if (r < c && col <= r) {
                                    double v = -acc[u][h][x];
                                    const int s = r / 6;
                                    if (col >= 6 * s) {
                                        const int ea = r - 6 * s, eb = col - 6 * s;
                                        v += Dg[s * 48 + ea * 8 + eb];
                                    }
                                    Gout[(size_t)r * ccap + col] = v;
                                    Gout[(size_t)col * ccap + r] = v;
                                }
                            }
                        }
                    }
                    ++J;
                }
            }
        }
    }
    // zero padding beyond the live block (capacity > n_cams)
    if (c < ccap) {
        for (int e = tid; e < ccap * ccap; e += GT) {
            const int r = e / ccap, cc = e - r * ccap;
            if (r >= c || cc >= c) Gout[e] = 0.0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Warp-specialised flavour for windows of up to 32 poses (the benchmark shape).  The profile of gram_kernel above showed the
// tensor pipe busy 52 % of the time with the DMMAs only 8 % of the instructions: every warp spent two thirds of its issue
// slots on the operand movement around them (scatter, D / g gathers, prefetch, addresses re-derived under register
// pressure) and a sixth of its time at the per-panel CTA barrier.  Here the roles are split:
//   * warps 0-3 (one warpgroup = one warp per scheduler) are PRODUCERS: cp.async prefetch of the factors two panels ahead,
//     slot -> observation maps, the rows of W of the next panel into one of two shared K-panels, and D / g of the panel
//     (8 window slots per warp) on the tensor pipe; they give registers back (setmaxnreg.dec);
//   * warps 4-15 are CONSUMERS: nothing but the W^T W loop, 13 units = 26 accumulator tiles per warp (setmaxnreg.inc),
//     resident for the whole filter.
// The two sides meet at named barriers only (full[2]: producers arrive, consumers wait; empty[2]: the other way round),
// so a consumer never waits for another consumer and the producers run a panel ahead.
constexpr int WS_PW = 4;            // producer warps
constexpr int WS_CW = GW - WS_PW;   // consumer warps
constexpr int WS_UM = 13;           // units per consumer warp: 12 * 13 = 156 = the units of a 24-tile triangle
#ifndef WS_UG_N
#define WS_UG_N 5
#endif
constexpr int WS_UG = WS_UG_N;      // units whose B fragments are in flight together
constexpr int BAR_FULL = 1, BAR_EMPTY = 3, BAR_PROD = 5, BAR_ALL = 6;

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// Build switches of the measurements in profiles/r02_summary.md section 5 (tools/build_variant.sh): GRAM_WS_REGS = 1 moves
// registers from the producers to the consumers with setmaxnreg (3 % slower, off); GRAM_DIAG != 0 SKIPS a stage (1: scatter and
// D / g, 2: the tensor loop, 3: scatter, 4: D / g) to time the rest -- wrong results, never shipped.
#ifndef GRAM_WS_REGS
#define GRAM_WS_REGS 0
#endif
#ifndef GRAM_DIAG
#define GRAM_DIAG 0
#endif

template <int PF, bool YSPLIT>
__global__ void __launch_bounds__(GT, 1) gram_ws_kernel(const CompressArgs a) {
    constexpr int NMAX = 32;
    constexpr int LDW = 6 * NMAX + 4;
    constexpr int SMAXP = NMAX / WS_PW;                 // window slots per producer warp for the D / g tiles
    constexpr int FPW = PF / WS_PW;                     // features fetched / scattered per producer warp
    static_assert(PF % WS_PW == 0, "whole features per producer warp");
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int F = a.dims.F, Mcap = a.dims.M, Ncap = a.dims.N;
    const int nc = a.n_cams[b];
    const int c = 6 * nc;
    const int ccap = 6 * Ncap;
    const int T = (c + 7) >> 3;
    const int T2 = (T + 1) >> 1;

    extern __shared__ double smem[];
    const int PM = PF * Mcap;
    constexpr int WSZ = 3 * PF * LDW;
    auto rawA = [&](int st) { return smem + 2 * WSZ + (size_t)st * PM * 21; };
    auto rawQ = [&](int st) { return rawA(st) + PM * 12; };
    auto rawR = [&](int st) { return rawA(st) + PM * 18; };
    auto rawS = [&](int st) { return reinterpret_cast<int*>(rawA(st) + PM * 20); };
    double* const Dg = rawA(2);
    int* const acc_list = reinterpret_cast<int*>(Dg + Ncap * 48);
    unsigned char* const inv = reinterpret_cast<unsigned char*>(acc_list + F);
    __shared__ int s_wcnt[GW];
    __shared__ int s_tot[2];
    __shared__ int s_len[2][PF];

    // ---- ordered list of the features that passed the gate (stacking order = feature order, MSCKF.py:573-588) ----
    if (tid < 2) s_tot[tid] = 0;
    __syncthreads();
    int nacc = 0, nrej_local = 0;
    for (int base = 0; base < F; base += GT) {
        const int j = base + tid;
        const int st = (j < F) ? (a.feat_status[(size_t)b * F + j] & 0xF) : 0;
        const bool ok = st == MSCKF_FEAT_ACCEPTED;
        nrej_local += (st == MSCKF_FEAT_GATED_OUT);
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = nacc;
        for (int w = 0; w < warp; ++w) off += s_wcnt[w];
        if (ok) acc_list[off + __popc(bal & ((1u << lane) - 1u))] = j;
        for (int w = 0; w < GW; ++w) nacc += s_wcnt[w];
        __syncthreads();
    }
    if (nrej_local) atomicAdd(&s_tot[1], nrej_local);
    // the track length rides along in the list entry (j | len << 20): the producers' prefetch then has no dependent global load
    for (int k = tid; k < nacc; k += GT) {
        const int j = acc_list[k];
        acc_list[k] = j | (min(a.feat_len[(size_t)b * F + j], Mcap) << 20);
    }
    __syncthreads();
    // gridDim.y CTAs may share a filter (fewer filters than SMs): each runs the whole producer side and owns a contiguous
    // share of the output units; CTA 0 of the filter also writes g, the counters and the padding
    const int nsplit = YSPLIT ? (int)gridDim.y : 1, split = YSPLIT ? (int)blockIdx.y : 0;      // (compile-time 1 / 0 for full batches)
    if (tid == 0 && split == 0) {
        a.n_accepted[b] = nacc;
        a.n_rejected[b] = s_tot[1];
    }
    double* Gout = a.T_H + (size_t)b * ccap * ccap;
    double* gout = a.r_n + (size_t)b * ccap;
    if (nacc == 0) {   // nothing survived: the update kernel skips this filter (MSCKF.py:584-585)
        if (split == 0)
            for (int e = tid; e < ccap; e += GT) gout[e] = 0.0;
        return;
    }
    const int npanels = (nacc + PF - 1) / PF;
    constexpr int KW = 3 * PF, KX = 2 * PF;

    if (warp < WS_PW) {
        // =================================================== producers ===================================================
#if GRAM_WS_REGS
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
#endif
        const int ptid = tid;                            // 0 .. 127
        double accX[SMAXP][2];
#pragma unroll
        for (int k = 0; k < SMAXP; ++k) accX[k][0] = accX[k][1] = 0.0;

        auto prefetch = [&](int pi, int st) {
            if (pi < npanels) {
#pragma unroll
                for (int h = 0; h < FPW; ++h) {
                    const int f = warp * FPW + h;
                    const int k = pi * PF + f;
                    const int packed = (k < nacc) ? acc_list[k] : 0;
                    const int j = packed & 0xFFFFF, M = packed >> 20;
                    if (lane == 0) s_len[st][f] = M;
                    if (M > 0) {
                        const size_t fj = (size_t)b * F + j;
                        const double* gA = a.fac_A + fj * (size_t)Mcap * 12;
                        const double* gQ = a.fac_Q + fj * (size_t)Mcap * 6;
                        const double* gR = a.fac_rt + fj * (size_t)Mcap * 2;
                        const int* gS = a.obs_slot + fj * (size_t)Mcap;
                        double* sA = rawA(st) + (size_t)f * Mcap * 12;
                        double* sQ = rawQ(st) + (size_t)f * Mcap * 6;
                        double* sR = rawR(st) + (size_t)f * Mcap * 2;
                        int* sS = rawS(st) + f * Mcap;
                        for (int it = lane; it < 6 * M; it += 32) cp_async16(sA + 2 * it, gA + 2 * it);
                        for (int it = lane; it < 3 * M; it += 32) cp_async16(sQ + 2 * it, gQ + 2 * it);
                        for (int it = lane; it < M; it += 32) {
                            cp_async16(sR + 2 * it, gR + 2 * it);
                            cp_async4(sS + it, gS + it);
                        }
                    }
                }
            }
            cp_async_commit();                           // (an empty group past the last panel keeps the wait count fixed)
        };
        prefetch(0, 0);
        prefetch(1, 1);
        for (int q = 0; q < npanels; ++q) {
            const int st = q & 1;
            cp_async_wait_1();                           // panel q's factors of this warp's features have landed
            __syncwarp();
            // slot -> observation maps of this warp's features
#pragma unroll
            for (int h = 0; h < FPW; ++h) {
                const int f = warp * FPW + h;
                unsigned char* iv = inv + (st * PF + f) * NMAX;
                iv[lane] = 0xFF;
                __syncwarp();
                const int* sS = rawS(st) + f * Mcap;
                const int len = s_len[st][f];
                for (int i = lane; i < len; i += 32) {
                    const int sl = sS[i];
                    if (sl >= 0 && sl < nc) iv[sl] = (unsigned char)i;
                }
            }
            __syncwarp();
            if (q >= 2) bar_sync(BAR_EMPTY + st, GT);    // the consumers are done with panel q - 2 in this K-panel
            // rows of W: W[3f+r][6 s + e] = Q1[2i][r] A[2i][e] + Q1[2i+1][r] A[2i+1][e], i = the observation of feature f in
            // window slot s (a zero segment when there is none); the three (slot, row) units of a lane are independent chains
#pragma unroll
            for (int h = 0; h < (GRAM_DIAG == 1 || GRAM_DIAG == 3 ? 0 : FPW); ++h) {
                const int f = warp * FPW + h;
                const unsigned char* iv = inv + (st * PF + f) * NMAX;
                double* Wp = smem + (size_t)st * WSZ + (size_t)(3 * f) * LDW;
                const double* fA = rawA(st) + (size_t)f * Mcap * 12;
                const double* fQ = rawQ(st) + (size_t)f * Mcap * 6;
#pragma unroll
                for (int rd = 0; rd < (3 * NMAX) / 32; ++rd) {
                    const int e = lane + 32 * rd;
                    const int sl = e / 3, r = e - 3 * sl;
                    if (e < 3 * nc) {
                        const int i = iv[sl];
                        const int ii = (i == 0xFF) ? 0 : i;
                        const double2* A2 = reinterpret_cast<const double2*>(fA + (size_t)ii * 12);
                        const double* Q = fQ + (size_t)ii * 6;
                        const double q0 = (i == 0xFF) ? 0.0 : Q[r], q1 = (i == 0xFF) ? 0.0 : Q[3 + r];
                        double2* dst = reinterpret_cast<double2*>(Wp + (size_t)r * LDW + 6 * sl);
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const double2 x = A2[k], y = A2[3 + k];
                            dst[k] = make_double2(q0 * x.x + q1 * y.x, q0 * x.y + q1 * y.y);
                        }
                    }
                }
            }
            __threadfence_block();
            bar_arrive(BAR_FULL + st, GT);
            bar_sync(BAR_PROD, WS_PW * 32);              // every producer's factors and maps of panel q are visible
            // D_s += A_i^T A_i, g_s += A_i^T r~_i: one 8 x 8 tile per window slot, operand A == operand B = [A_i | r~_i | 0];
            // the eight gathers of a k step are issued before its tensor instructions (a slot nobody observes adds zeros)
            if (GRAM_DIAG != 1 && GRAM_DIAG != 4) {
                const double* sA = rawA(st);
                const double* sR = rawR(st);
                const unsigned char* iv = inv + st * PF * NMAX;
#pragma unroll
                for (int kk = 0; kk < KX; kk += 4) {
                    const int f = (kk >> 1) + (t >> 1), r = t & 1;
                    double x[SMAXP];
#pragma unroll
                    for (int k = 0; k < SMAXP; ++k) {
                        const int i = iv[f * NMAX + warp + WS_PW * k];
                        const size_t o = (size_t)f * Mcap + (i == 0xFF ? 0 : i);
                        const double* src = (g < 6) ? sA + o * 12 + 6 * r + g : sR + o * 2 + r;
                        const double v = *src;
                        x[k] = (i != 0xFF && g < 7) ? v : 0.0;
                    }
#pragma unroll
                    for (int k = 0; k < SMAXP; ++k) dense::dmma(accX[k][0], accX[k][1], x[k], x[k]);
                }
            }
            bar_sync(BAR_PROD, WS_PW * 32);              // every producer is done with stage st: refill it
            prefetch(q + 2, st);
        }
#pragma unroll
        for (int k = 0; k < SMAXP; ++k) {
            const int s = warp + WS_PW * k;
            if (s < nc && g < 6) {   // rows 6, 7 of the tile are by-products (r~^T A, 0)
                Dg[s * 48 + g * 8 + 2 * t] = accX[k][0];
                Dg[s * 48 + g * 8 + 2 * t + 1] = accX[k][1];
            }
        }
        bar_sync(BAR_ALL, GT);
        if (split == 0)
            for (int e = tid; e < ccap; e += GT) {
                const int s = e / 6, aa = e - 6 * s;
                gout[e] = (s < nc) ? Dg[s * 48 + aa * 8 + 6] : 0.0;
            }
        (void)ptid;
    } else {
        // =================================================== consumers ===================================================
#if GRAM_WS_REGS
        asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
#endif
        const int cw = warp - WS_PW;
        int U = 0;
        for (int p = 0; p < T2; ++p) U += strip_len(p, T);
        const int ub = (int)(((long long)U * split) / nsplit), ue = (int)(((long long)U * (split + 1)) / nsplit);
        const int upw = (ue - ub + WS_CW - 1) / WS_CW;   // <= WS_UM for T <= 24
        const int u0 = min(ue, ub + cw * upw);
        const int nu = min(ue, u0 + upw) - u0;
        int p0 = 0, J0 = u0;
        while (p0 < T2 && J0 >= strip_len(p0, T)) { J0 -= strip_len(p0, T); ++p0; }
        unsigned newstrip = 0;       // bit u: unit u is the first of a strip (reload the two A fragments, column 0)
        {
            int p = p0, J = J0;
#pragma unroll
            for (int u = 0; u < WS_UM; ++u) {
                if (u == 0 || J == 0) newstrip |= 1u << u;
                if (u < nu && ++J == strip_len(p, T)) { J = 0; ++p; }
            }
        }
        double acc[WS_UM][2][2];
#pragma unroll
        for (int u = 0; u < WS_UM; ++u) acc[u][0][0] = acc[u][0][1] = acc[u][1][0] = acc[u][1][1] = 0.0;
        const unsigned w0 = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)(t * LDW + g) * 8u;
        for (int pi = 0; pi < npanels; ++pi) {
            const int st = pi & 1;
            bar_sync(BAR_FULL + st, GT);
            if (GRAM_DIAG != 2) gram_tensor_loop<KW, LDW, WS_UM, WS_UG>(w0 + (unsigned)(st * WSZ) * 8u, p0, J0, nu, newstrip, acc);
            if (pi + 2 < npanels) bar_arrive(BAR_EMPTY + st, GT);
        }
        bar_sync(BAR_ALL, GT);
        if (split == 0)
            for (int e = tid; e < ccap; e += GT) {
                const int s = e / 6, aa = e - 6 * s;
                gout[e] = (s < nc) ? Dg[s * 48 + aa * 8 + 6] : 0.0;
            }
        // ---- G = D - W^T W: lower tiles written with their mirror image ----
        int p = p0, J = J0;
#pragma unroll
        for (int u = 0; u < WS_UM; ++u) {
            if (u < nu) {
                if (u > 0 && ((newstrip >> u) & 1u)) { ++p; J = 0; }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int I = 2 * p + h;
                    if (I < T && J <= I) {
                        const int r = 8 * I + g;
#pragma unroll
                        for (int x = 0; x < 2; ++x) {
                            const int col = 8 * J + 2 * t + x;
                            if (r < c && col <= r) {
                                double v = -acc[u][h][x];
                                const int s = r / 6;
                                if (col >= 6 * s) {
                                    const int ea = r - 6 * s, eb = col - 6 * s;
                                    v += Dg[s * 48 + ea * 8 + eb];
                                }
                                Gout[(size_t)r * ccap + col] = v;
                                Gout[(size_t)col * ccap + r] = v;
                            }
                        }
                    }
                }
                ++J;
            }
        }
    }
    // zero padding beyond the live block (capacity > n_cams): the rows below it, then the columns to its right
    if (c < ccap && split == 0) {
        const int pad = ccap - c;
        for (int e = tid; e < pad * ccap; e += GT) Gout[(size_t)c * ccap + e] = 0.0;
        for (int e = tid; e < c * pad; e += GT) {
            const int r = e / pad, cc = e - r * pad;
            Gout[(size_t)r * ccap + c + cc] = 0.0;
        }
    }
}

struct GramPlan { int NMAX, PF; size_t smem; };

GramPlan gram_plan(const msckf_dims& d) {
    GramPlan best{0, 0, 0};
    if (d.N > 64) return best;
    const int NMAX = d.N <= 32 ? 32 : 64;
    const int LDW = 6 * NMAX + 4;
    for (int PF = 8; PF >= 4; PF >>= 1) {
        size_t doubles = (size_t)2 * 3 * PF * LDW + (size_t)2 * PF * d.M * 21 + (size_t)d.N * 48;
        size_t bytes = doubles * 8 + (size_t)d.F * 4 + (size_t)2 * PF * NMAX + 64;
        if (bytes <= 227 * 1024) { best = GramPlan{NMAX, PF, bytes}; break; }
    }
    return best;
}

template <int NMAX, int PF>
cudaError_t gram_launch(const CompressArgs& a, size_t smem, cudaStream_t s) {
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(gram_kernel<NMAX, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    gram_kernel<NMAX, PF><<<a.dims.B, GT, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace

bool gram_supported(const msckf_dims& d) { return gram_plan(d).PF >= 4; }

cudaError_t launch_compress_gram(const CompressArgs& a, cudaStream_t s) {
    const GramPlan pl = gram_plan(a.dims);
    if (pl.PF < 4) return cudaErrorInvalidConfiguration;
    msckf_count_launch(1);
#ifndef GRAM_WS
#define GRAM_WS 1
#endif
    if (GRAM_WS && pl.NMAX == 32 && pl.PF == 8) {
        static size_t configured = 0;
        if (pl.smem > configured) {
            cudaError_t e = cudaFuncSetAttribute(gram_ws_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(gram_ws_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
            if (e != cudaSuccess) return e;
            configured = pl.smem;
        }
        int ny = a.dims.B > 0 ? 148 / a.dims.B : 1;      // few filters: several CTAs per filter, each a share of the output tiles
        if (ny > 6) ny = 6;
        if (ny > 1) gram_ws_kernel<8, true><<<dim3(a.dims.B, ny), GT, pl.smem, s>>>(a);
        else gram_ws_kernel<8, false><<<a.dims.B, GT, pl.smem, s>>>(a);
        return cudaGetLastError();
    }
    if (pl.NMAX == 32) return pl.PF == 8 ? gram_launch<32, 8>(a, pl.smem, s) : gram_launch<32, 4>(a, pl.smem, s);
    return pl.PF == 8 ? gram_launch<64, 8>(a, pl.smem, s) : gram_launch<64, 4>(a, pl.smem, s);
}
