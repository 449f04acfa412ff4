// Per-feature front end as THREE kernels (rows 3-8 of SURVEY.md section 8a) for tracks of up to 32 observations:
//   selection + closed-form triangulation     reference src/msckf/MSCKF.py:458-495, src/utils/geometry.py:274-303
//   residual + Jacobian blocks                MSCKF.py:497-548, src/msckf/Camera.py:54-67
//   left-null-space projection of H_f         MSCKF.py:554-559
//   chi-square Mahalanobis gate               MSCKF.py:561-568
//
// A single fused kernel would have to keep the 2M x 2M matrix K = H_x P H_x^T + sigma^2 I of a feature in registers while
// the camera block of P occupies the SM's shared memory: 255 registers x 8 warps, two warps per scheduler, and every
// phase of it is latency-bound (measured in round 1: 42.6 ms against 18.6 ms, profiles/r01n; that kernel was removed).
// Round 1 ran geometry + assembly in one kernel: the geometry is a chain of warp reductions whose every shuffle then waits
// in the same memory-input queue as the shared-memory-bound pair loops of the other warps (~9 us per feature and warp).
// Here every phase gets the resources it needs:
//
//   front_geo_kernel   one warp per feature, lane = observation, many warps per SM (no covariance staged, the window
//                      poses are the only shared memory): geometry (front_geom.cuh), the factors A_i / Q1 / r~ written to
//                      the workspace (the compression stage needs them anyway), the exact gate bounds when requested, and
//                      a key per pending feature (first window slot, length, "slots are consecutive"), its slot ->
//                      observation map, and every part of the padded K that is not a measurement block.
//   front_asm_kernel   CTA = filter, 12 warps.  P's camera block (lower block triangle, 6x6 blocks padded to 38 doubles ->
//                      conflict-free 16-byte reads) staged once per CTA; the filter's pending features are bucketed by
//                      key in shared memory and taken in PAIRS: two features with the same slot pattern share every
//                      staged 6x6 block of P (one 16-byte shared-memory read feeds the multiply-adds of both), which is
//                      what bounds this kernel (shared-memory wavefronts).  The M (M + 1) / 2 lower 2x2 blocks
//                      A_i P[s_i, s_k] A_k^T are a FLAT list, one block per lane per pass (97 % lane use at M = 30), stored
//                      straight into the accumulator-fragment tile layout of the FP64 tensor instruction in a global
//                      (L2 / HBM) buffer.  The four appended rows [r~; Q1^T] are ordinary rows of the padded matrix (they
//                      fill the padding rows of the last tile row at M = 30): 18 KB per feature, written once, read once.
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
#include "front_tiles.cuh"
#include "front_gate.cuh"

namespace {

using namespace fgeom;
using namespace ftile;

constexpr int FG_THREADS = 256;                // geometry kernel: 8 warps per CTA, 2+ CTAs per SM
constexpr int FG_WARPS = FG_THREADS / 32;
constexpr int FA_WARPS_MAX = 12;               // assembly kernel: up to 12 warps (170 registers each), 2 features per warp
constexpr int FA_THREADS_MAX = FA_WARPS_MAX * 32;
constexpr int FC_THREADS = 128;
constexpr int FC_WARPS = FC_THREADS / 32;
constexpr int FLAG_KEEP = MSCKF_FEAT_LOST_FLAG | MSCKF_FEAT_TRIANGULATED;   // selection flags carried into the final status
constexpr int TC = 9;                          // 8x8 tile columns: up to 64 measurement rows (32 observations) + 4 appended rows
constexpr int NTILE = TC * (TC + 1) / 2;
constexpr int KT_D2 = NTILE * 32;               // double2 per feature slot in the tile buffer (45 tiles; M = 30 uses 36)
constexpr int KEY_BINS = 32 * 33 + 1;          // (first slot, length) of a consecutive-slot track; last bin: anything else

// key of a pending feature: bin of the assembly kernel's bucketing, or -1 (not pending)
__device__ __forceinline__ int make_key(bool affine, int s0, int M) { return affine ? (s0 * 33 + M) : (KEY_BINS - 1); }

// ---------------------------------------------------------------------------------------------------------------------
// kernel G: geometry, factors, gate bounds, keys
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FG_THREADS, 2)
front_geo_kernel(const FrontArgs a, int* __restrict__ pend_key, unsigned char* __restrict__ slot_map, double2* __restrict__ Kbuf,
                 int* __restrict__ pend_list, int* __restrict__ pend_count, const int b0, const int feats_per_cta) {
    const int bl = blockIdx.y, b = b0 + bl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = a.dims.F, Mcap = a.dims.M, Ncap = a.dims.N, ldp = a.dims.ldp;
    const int nc = min(a.n_cams[b], Ncap);

    extern __shared__ double smem[];
    double* sR = smem;                       // [Ncap][9]
    double* sT = sR + (size_t)Ncap * 9;      // [Ncap][3]

    const int f_begin = blockIdx.x * feats_per_cta;
    const int f_end = min(F, f_begin + feats_per_cta);
    if (f_begin >= f_end) return;
    {   // a CTA whose tracks are all empty (an update over a subset of the tracks: the pruning update of a filter that does not
        // prune this frame) only reports them skipped -- one coalesced pass instead of a walk by eight warps
        int live = 0;
        for (int j = f_begin + tid; j < f_end; j += FG_THREADS) live |= a.feat_len[(size_t)b * F + j] > 0;
        if (!__syncthreads_or(live)) {
            for (int j = f_begin + tid; j < f_end; j += FG_THREADS) {
                const size_t fj = (size_t)b * F + j;
                a.feat_status[fj] = MSCKF_FEAT_SKIPPED;
                a.gamma[fj] = nan("");
                a.dof[fj] = 0;
                a.fac_rank[fj] = 0;
                pend_key[(size_t)bl * F + j] = -1;
            }
            return;
        }
    }
    {
        const double* camR = a.cam_R + (size_t)b * Ncap * 9;
        const double* camt = a.cam_t + (size_t)b * Ncap * 3;
        for (int e = tid; e < nc * 9; e += FG_THREADS) sR[e] = camR[e];
        for (int e = tid; e < nc * 3; e += FG_THREADS) sT[e] = camt[e];
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
    for (int j = f_begin + warp; j < f_end; j += FG_WARPS) {
        const size_t fj = (size_t)b * F + j;
        int* key_out = pend_key + (size_t)bl * F + j;
        ObsRegs o = nxt;
        const int M = Mn;
        mask_obs(o, M, lane);
        {   // prefetch the next feature of this warp
            const int jn = j + FG_WARPS;
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
                *key_out = -1;
            }
            continue;
        }
        const int n2 = 2 * M;
        const bool live = lane < M;

        int rank, code;
        double rt0, rt1;
        double Hx[2][6], Hf[2][3];
        code = feature_geometry(a, gc, fj, M, lane, o, sR, sT, Hx, Hf, rt0, rt1, rank);
        if ((code & 0xF) != MSCKF_FEAT_ACCEPTED) {
            if (lane == 0) {
                a.feat_status[fj] = (unsigned char)code;
                a.gamma[fj] = nan("");
                a.dof[fj] = 0;
                a.fac_rank[fj] = 0;
                *key_out = -1;
            }
            continue;
        }
        // ---- factors for the assembly kernel and the compression stage ----
        if (live) {
            double2* fA = reinterpret_cast<double2*>(a.fac_A + (fj * Mcap + lane) * 12);
            double2* fQ = reinterpret_cast<double2*>(a.fac_Q + (fj * Mcap + lane) * 6);
#pragma unroll
            for (int c2 = 0; c2 < 6; ++c2) fA[c2] = make_double2(Hx[c2 / 3][2 * (c2 % 3)], Hx[c2 / 3][2 * (c2 % 3) + 1]);
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
                *key_out = -1;
            }
            continue;
        }

        // ---- exact early-out (MSCKF_GATE_BOUNDS): |r~|^2 / (s^2 + tr S_x) <= gamma <= |r~|^2 / s^2 ----
        if (a.par.gate_mode == MSCKF_GATE_BOUNDS) {
            double trs = 0.0;
            if (live) {
                const double* Pd = a.P + (size_t)b * ldp * ldp + (size_t)(15 + 6 * o.slot) * ldp + 15 + 6 * o.slot;
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    double v0 = 0.0, v1 = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        // (lower triangle of the symmetric block, as the assembly kernel stages it)
                        const double pv = (c <= r) ? Pd[(size_t)r * ldp + c] : Pd[(size_t)c * ldp + r];
                        v0 += pv * Hx[0][c];
                        v1 += pv * Hx[1][c];
                    }
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
                    *key_out = -1;
                }
                continue;
            }
        }

        // ---- pending: the assembly + factorisation kernels finish this feature ----
        const int s0 = __shfl_sync(0xffffffffu, o.slot, 0);
        const bool affine = __all_sync(0xffffffffu, !live || o.slot == s0 + lane) && s0 >= 0 && s0 < 32;
        const int fl = bl * F + j;
        if (lane == 0) {
            a.feat_status[fj] = (unsigned char)(MSCKF_FEAT_PENDING | (code & FLAG_KEEP));
            a.dof[fj] = dofv;
            a.fac_rank[fj] = rank;
            *key_out = make_key(affine, s0, M);
            pend_list[atomicAdd(pend_count, 1)] = fl;
        }
        // everything of the padded matrix that is not a measurement block: appended rows [r~; Q1^T], identity / zero padding
        double2* Kt = Kbuf + (size_t)fl * KT_D2;
        if (live) write_appended_obs(Kt, M, lane, rt0, rt1, Hf[0][0], Hf[0][1], Hf[0][2], Hf[1][0], Hf[1][1], Hf[1][2]);
        write_padding_rows(Kt, M, lane, 32);
        // window slot -> observation index of this track (0xFF: not observed from that slot)
        unsigned char* mp = slot_map + (size_t)fl * 32;
        mp[lane] = 0xFF;
        __syncwarp();
        if (live && o.slot >= 0 && o.slot < 32) mp[o.slot] = (unsigned char)lane;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// kernel A: assembly of K into the tile buffer, two features of equal slot pattern per warp
// ---------------------------------------------------------------------------------------------------------------------
struct AsmFeature {
    int j;            // feature index in the filter (-1: none)
    int M, s0;
    bool affine;
    int slot;         // this lane's observation slot
    double2* Kt;      // its tile buffer
    const double* wsA;
};

// K = H_x P H_x^T + sigma^2 I of NF features with the SAME consecutive slot pattern (s0, M): one lower 2x2 block per lane
// per pass and feature; every 16-byte read of a staged block of P feeds the multiply-adds of all NF features.  The
// arithmetic of a feature does not depend on NF (same operations in the same order).
template <int NF>
__device__ __forceinline__ void pair_loop_affine(const double* __restrict__ sP, const AsmFeature (&f)[NF], const int lane,
                                                 const double s2) {
    const int M = f[0].M, s0 = f[0].s0;
    const int npairs = M * (M + 1) / 2;
#pragma unroll 1
    for (int p0 = 0; p0 < npairs; p0 += 32) {
        const int p = p0 + lane;
        if (p >= npairs) continue;
        int i, k;
        pair_of(p, i, k);
        double ai[NF][2][6];
#pragma unroll
        for (int q = 0; q < NF; ++q) {
            const double2* Ai = reinterpret_cast<const double2*>(f[q].wsA + i * AR);
#pragma unroll
            for (int c2 = 0; c2 < 6; ++c2) {
                const double2 x = Ai[c2];
                ai[q][c2 / 3][2 * (c2 % 3)] = x.x;
                ai[q][c2 / 3][2 * (c2 % 3) + 1] = x.y;
            }
        }
        double v0[NF][6], v1[NF][6];
#pragma unroll
        for (int q = 0; q < NF; ++q)
#pragma unroll
            for (int c = 0; c < 6; ++c) { v0[q][c] = 0.0; v1[q][c] = 0.0; }
        const double2* Pb = reinterpret_cast<const double2*>(sP + (size_t)tile_index(s0 + i, s0 + k) * PB);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int c2 = 0; c2 < 3; ++c2) {
                const double2 pv = Pb[r * 3 + c2];
#pragma unroll
                for (int q = 0; q < NF; ++q) {
                    v0[q][2 * c2] += ai[q][0][r] * pv.x;
                    v0[q][2 * c2 + 1] += ai[q][0][r] * pv.y;
                    v1[q][2 * c2] += ai[q][1][r] * pv.x;
                    v1[q][2 * c2 + 1] += ai[q][1][r] * pv.y;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NF; ++q) {
            double o00 = 0.0, o01 = 0.0, o10 = 0.0, o11 = 0.0;
            const double2* Ak = reinterpret_cast<const double2*>(f[q].wsA + k * AR);
#pragma unroll
            for (int c2 = 0; c2 < 3; ++c2) {
                const double2 x0 = Ak[c2], x1 = Ak[3 + c2];
                o00 = fma(v0[q][2 * c2 + 1], x0.y, fma(v0[q][2 * c2], x0.x, o00));
                o01 = fma(v0[q][2 * c2 + 1], x1.y, fma(v0[q][2 * c2], x1.x, o01));
                o10 = fma(v1[q][2 * c2 + 1], x0.y, fma(v1[q][2 * c2], x0.x, o10));
                o11 = fma(v1[q][2 * c2 + 1], x1.y, fma(v1[q][2 * c2], x1.x, o11));
            }
            if (i == k) { o00 += s2; o11 += s2; }
            store_block(f[q].Kt, i, k, o00, o01, o10, o11);
        }
    }
}

// the same for ONE feature whose slots are arbitrary (gaps, or stored newest-first): slots from the lanes that hold them,
// transposed block when the slots do not increase with the observation index
__device__ __forceinline__ void pair_loop_general(const double* __restrict__ sP, const AsmFeature& f, const int lane,
                                                  const double s2) {
    const int M = f.M;
    const int npairs = M * (M + 1) / 2;
#pragma unroll 1
    for (int p0 = 0; p0 < npairs; p0 += 32) {
        const int p = p0 + lane;
        const bool act = p < npairs;
        int i, k;
        pair_of(p, i, k);
        const int si = __shfl_sync(0xffffffffu, f.slot, i & 31);
        const int sk = __shfl_sync(0xffffffffu, f.slot, k & 31);
        if (!act) continue;
        double ai[2][6];
        {
            const double2* Ai = reinterpret_cast<const double2*>(f.wsA + i * AR);
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
        } else {   // transposed block
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
        double o00 = 0.0, o01 = 0.0, o10 = 0.0, o11 = 0.0;
        const double2* Ak = reinterpret_cast<const double2*>(f.wsA + k * AR);
#pragma unroll
        for (int c2 = 0; c2 < 3; ++c2) {
            const double2 x0 = Ak[c2], x1 = Ak[3 + c2];
            o00 = fma(v0[2 * c2 + 1], x0.y, fma(v0[2 * c2], x0.x, o00));
            o01 = fma(v0[2 * c2 + 1], x1.y, fma(v0[2 * c2], x1.x, o01));
            o10 = fma(v1[2 * c2 + 1], x0.y, fma(v1[2 * c2], x0.x, o10));
            o11 = fma(v1[2 * c2 + 1], x1.y, fma(v1[2 * c2], x1.x, o11));
        }
        if (i == k) { o00 += s2; o11 += s2; }
        store_block(f.Kt, i, k, o00, o01, o10, o11);
    }
}

__global__ void __launch_bounds__(FA_THREADS_MAX, 1)
front_asm_kernel(const FrontArgs a, const int* __restrict__ pend_key, double2* __restrict__ Kbuf, const int b0) {
    const int bl = blockIdx.x, b = b0 + bl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = a.dims.F, Mcap = a.dims.M, Ncap = a.dims.N, ldp = a.dims.ldp;
    const int nc = min(a.n_cams[b], Ncap);
    const int nthreads = blockDim.x, nwarps = blockDim.x >> 5;

    extern __shared__ double smem[];
    double* sP = smem;                                                        // [nblk(Ncap)][PB] lower block triangle of P_CC
    double* wsAll = sP + (size_t)(Ncap * (Ncap + 1) / 2) * PB;                // [nwarps][2][FW_MAX_OBS * AR]
    unsigned short* s_list = reinterpret_cast<unsigned short*>(wsAll + (size_t)nwarps * 2 * FW_MAX_OBS * AR);   // [F] bucketed
    int* s_hist = reinterpret_cast<int*>(wsAll);                              // [KEY_BINS] (aliases the scratch: start-up only)
    int* s_cur = s_hist + KEY_BINS;                                           // [KEY_BINS]
    __shared__ int s_npend, s_next;

    const int* keys = pend_key + (size_t)bl * F;

    // ---- bucket the filter's pending features by slot pattern ----
    for (int e = tid; e < 2 * KEY_BINS; e += nthreads) s_hist[e] = 0;
    if (tid == 0) { s_npend = 0; s_next = 0; }
    __syncthreads();
    for (int j = tid; j < F; j += nthreads) {
        const int key = keys[j];
        if (key >= 0) atomicAdd(&s_hist[key], 1);
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the bins, 34 per lane
        constexpr int PER = (KEY_BINS + 31) / 32;
        int sum = 0;
        for (int q = 0; q < PER; ++q) { const int e = lane * PER + q; sum += (e < KEY_BINS) ? s_hist[e] : 0; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        int run = incl - sum;
        for (int q = 0; q < PER; ++q) {
            const int e = lane * PER + q;
            if (e < KEY_BINS) { s_cur[e] = run; run += s_hist[e]; }
        }
        if (lane == 31) s_npend = incl;
    }
    __syncthreads();
    for (int j = tid; j < F; j += nthreads) {
        const int key = keys[j];
        if (key >= 0) s_list[atomicAdd(&s_cur[key], 1)] = (unsigned short)j;
    }
    __syncthreads();
    const int npend = s_npend;
    if (npend == 0) return;

    // ---- stage the camera block of P (the scratch aliasing ends here: barrier above) ----
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
    }
    __syncthreads();

    const double s2 = a.par.sigma_image * a.par.sigma_image;
    double* ws0 = wsAll + (size_t)warp * 2 * FW_MAX_OBS * AR;
    double* ws1 = ws0 + FW_MAX_OBS * AR;
    const int n_units = (npend + 1) >> 1;

    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(&s_next, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_units) break;
        AsmFeature f[2];
        int key[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int e = 2 * u + q;
            AsmFeature& g = f[q];
            g.j = (e < npend) ? (int)s_list[e] : -1;
            g.wsA = q ? ws1 : ws0;
            g.M = 0; g.s0 = 0; g.affine = false; g.slot = 0; g.Kt = nullptr;
            key[q] = -1;
            if (g.j < 0) continue;
            const size_t fj = (size_t)b * F + g.j;
            key[q] = keys[g.j];
            g.M = a.feat_len[fj];
            g.affine = key[q] != KEY_BINS - 1;
            g.Kt = Kbuf + ((size_t)bl * F + g.j) * KT_D2;
            const size_t e0 = fj * Mcap + (lane < g.M ? lane : 0);
            g.slot = a.obs_slot[e0];
            const double2* fA = reinterpret_cast<const double2*>(a.fac_A + e0 * 12);
            double2 av[6];
#pragma unroll
            for (int c2 = 0; c2 < 6; ++c2) av[c2] = fA[c2];
            g.s0 = __shfl_sync(0xffffffffu, g.slot, 0);
            __syncwarp();      // the previous unit's pair loop no longer reads this scratch
            if (lane < g.M) {
                double2* sA = reinterpret_cast<double2*>(const_cast<double*>(g.wsA) + lane * AR);
#pragma unroll
                for (int c2 = 0; c2 < 6; ++c2) sA[c2] = av[c2];
            }
        }
        __syncwarp();
        if (f[1].j >= 0 && key[0] == key[1] && f[0].affine) {
            pair_loop_affine<2>(sP, f, lane, s2);
        } else {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (f[q].j < 0) continue;
                if (f[q].affine) {
                    const AsmFeature one[1] = {f[q]};
                    pair_loop_affine<1>(sP, one, lane, s2);
                } else {
                    pair_loop_general(sP, f[q], lane, s2);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// kernel A (windows of up to 30 / 31 poses): P-STATIONARY assembly.  The lower block triangle of P's camera block lives in
// the REGISTER FILE of the CTA: thread p owns the 6x6 block P[s, t] of the slot pair (s, t), p = s (s + 1) / 2 + t, for the
// whole life of the CTA (465 blocks = 15 warps at a 30-pose window), and the filter's pending features stream through
// shared memory (cp.async, double-buffered batches of their 2x6 Jacobian blocks + slot maps).  For every feature each
// thread computes the one 2x2 block of K that involves its block of P: A_i P[s, t] A_k^T with (i, k) the track's
// observations from the slots (s, t).  Per block that is 12 shared-memory reads (the two Jacobian blocks) instead of 30
// (Jacobians + the staged block of P), the shared memory holds ~56 features in flight instead of 24, and tracks with
// gaps or stored newest-first need no special path (the slot map says which observation, if any, a slot holds; a block
// above the diagonal is stored transposed).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int NLEFT>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NLEFT) : "memory"); }

constexpr int PS_FEAT_BYTES = FW_MAX_OBS * AR * 8 + 32;   // one feature in a batch buffer: Jacobian blocks + slot map

// WIDE (tracks of up to 64 observations, windows of up to 64 poses): the block triangle no longer fits one CTA's threads, so
// gridDim.y CTAs share a filter, CTA y owning the blocks [y NT, (y + 1) NT); every CTA streams all of the filter's pending
// features (64-observation slots, 64-byte maps) and writes its own blocks of their K.
template <int NWARPS, bool WIDE, bool ZSPLIT = false>
__global__ void __launch_bounds__(NWARPS * 32, 1)
front_asm_ps_kernel(const FrontArgs a, const unsigned char* __restrict__ slot_map, double2* __restrict__ Kbuf, const int b0,
                    const int batch, const int kt_d2) {
    constexpr int NT = NWARPS * 32;
    constexpr int OBS = WIDE ? 64 : FW_MAX_OBS;          // observation slots per feature in a batch buffer
    constexpr int MAPB = WIDE ? 64 : 32;                 // bytes of a slot map
    constexpr int FEATB = OBS * AR * 8 + MAPB;
    const int bl = blockIdx.x, b = b0 + bl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = a.dims.F, Mcap = a.dims.M, Ncap = a.dims.N, ldp = a.dims.ldp;
    const int nc = min(a.n_cams[b], Ncap);
    const int nblk = nc * (nc + 1) / 2;

    extern __shared__ double smem[];
    unsigned char* sbuf = reinterpret_cast<unsigned char*>(smem);                        // [2][batch][FEATB]
    unsigned short* s_list = reinterpret_cast<unsigned short*>(sbuf + (size_t)2 * batch * FEATB);   // [F] pending, in order
    unsigned char* s_len = reinterpret_cast<unsigned char*>(s_list + ((F + 7) & ~7));    // [F] their lengths
    __shared__ int s_warp_cnt[NWARPS + 1];

    // ---- the filter's pending features, in feature order (ballot scan) ----
    int npend = 0;
    for (int base = 0; base < F; base += NT) {
        const int j = base + tid;
        const bool pend = j < F && (a.feat_status[(size_t)b * F + j] & 0xF) == MSCKF_FEAT_PENDING;
        const unsigned bal = __ballot_sync(0xffffffffu, pend);
        if (lane == 0) s_warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = npend;
        for (int w = 0; w < warp; ++w) off += s_warp_cnt[w];
        if (pend) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));
            s_list[pos] = (unsigned short)j;
            s_len[pos] = (unsigned char)a.feat_len[(size_t)b * F + j];
        }
        int tot = 0;
        for (int w = 0; w < NWARPS; ++w) tot += s_warp_cnt[w];
        npend += tot;
        __syncthreads();
    }
    if (npend == 0) return;

    // ---- teams: ceil(nblk / 32) warps cover the block triangle once; several teams work on different features ----
    const int team_warps = WIDE ? NWARPS : (nblk + 31) >> 5;
    const int n_teams = WIDE ? 1 : NWARPS / team_warps;           // >= 1 (the launcher guarantees nblk <= NT per CTA row)
    const int team = WIDE ? 0 : warp / team_warps;
    const int p = WIDE ? (int)blockIdx.y * NT + tid : (warp - team * team_warps) * 32 + lane;
    const bool has_block = team < n_teams && p < nblk;
    int s = 0, t = 0;
    if (has_block) pair_of(p, s, t);
    double Pb[6][6];
    {
        const double* P = a.P + (size_t)b * ldp * ldp + (size_t)(15 + 6 * s) * ldp + 15 + 6 * t;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) Pb[r][c] = has_block ? P[(size_t)r * ldp + c] : 0.0;
    }
    const double s2 = a.par.sigma_image * a.par.sigma_image;
    const int n_batches = (npend + batch - 1) / batch;

    auto issue_batch = [&](int bt, int buf) {
        unsigned char* dst = sbuf + (size_t)buf * batch * FEATB;
        const int e_end = min(batch, npend - bt * batch);
        // 16-byte chunks: 6 per observation of Jacobian blocks (AR = 14 doubles per observation in shared memory) + 2 of map
        for (int e = warp; e < e_end; e += NWARPS) {
            const int pos = bt * batch + e;
            const int j = s_list[pos], M = s_len[pos];
            const size_t fj = (size_t)b * F + j;
            const double* gA = a.fac_A + fj * Mcap * 12;
            unsigned char* d = dst + (size_t)e * FEATB;
            for (int it = lane; it < 6 * M; it += 32) {
                const int ob = it / 6, c2 = it - 6 * ob;
                cp_async16(d + (size_t)(ob * AR + 2 * c2) * 8, gA + (size_t)it * 2);
            }
            if (lane < MAPB / 16) cp_async16(d + OBS * AR * 8 + 16 * lane, slot_map + ((size_t)bl * F + j) * MAPB + 16 * lane);
        }
        cp_async_commit();
    };

    // gridDim.z CTAs share a filter's pending features batch by batch (z, z + gridDim.z, ...): with fewer filters than SMs
    // (a single filter: the drop-in's own case) one CTA per filter left the assembly at the latency of 256 features in a row
    const int zs = ZSPLIT ? (int)gridDim.z : 1, z0 = ZSPLIT ? (int)blockIdx.z : 0;      // (compile-time 1 / 0 for full batches)
    if (z0 < n_batches) issue_batch(z0, 0);
    for (int bt = z0, it = 0; bt < n_batches; bt += zs, ++it) {
        if (bt + zs < n_batches) {
            issue_batch(bt + zs, (it + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const unsigned char* src = sbuf + (size_t)(it & 1) * batch * FEATB;
        const int e_end = min(batch, npend - bt * batch);
        if (has_block) {
            // (the slot map of the NEXT feature is read one iteration ahead: its latency is off the dependency chain)
            int i_n = 0xFF, k_n = 0xFF;
            if (team < e_end) {
                const unsigned char* mp = src + (size_t)team * FEATB + OBS * AR * 8;
                i_n = mp[s];
                k_n = mp[t];
            }
#pragma unroll 1
            for (int e = team; e < e_end; e += n_teams) {
                const unsigned char* fb = src + (size_t)e * FEATB;
                const int i = i_n, k = k_n;
                if (e + n_teams < e_end) {
                    const unsigned char* mp = fb + (size_t)n_teams * FEATB + OBS * AR * 8;
                    i_n = mp[s];
                    k_n = mp[t];
                }
                if (i == 0xFF || k == 0xFF) continue;
                const double* sA = reinterpret_cast<const double*>(fb);
                const double2* Ai = reinterpret_cast<const double2*>(sA + i * AR);
                const double2* Ak = reinterpret_cast<const double2*>(sA + k * AR);
                double ak[2][6];
#pragma unroll
                for (int c2 = 0; c2 < 6; ++c2) {
                    const double2 x = Ak[c2];
                    ak[c2 / 3][2 * (c2 % 3)] = x.x;
                    ak[c2 / 3][2 * (c2 % 3) + 1] = x.y;
                }
                double o[2][2];
                // one row of A_i at a time: six independent accumulation chains over the rows of the block of P (the same
                // operations in the same order as the shared-memory kernel: identical bits)
#pragma unroll
                for (int row = 0; row < 2; ++row) {
                    double ar[6];
#pragma unroll
                    for (int c2 = 0; c2 < 3; ++c2) {
                        const double2 x = Ai[3 * row + c2];
                        ar[2 * c2] = x.x;
                        ar[2 * c2 + 1] = x.y;
                    }
                    double v[6];
#pragma unroll
                    for (int c = 0; c < 6; ++c) v[c] = 0.0;
#pragma unroll
                    for (int r = 0; r < 6; ++r)
#pragma unroll
                        for (int c = 0; c < 6; ++c) v[c] += ar[r] * Pb[r][c];
                    double q0 = 0.0, q1 = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        q0 = fma(v[c], ak[0][c], q0);
                        q1 = fma(v[c], ak[1][c], q1);
                    }
                    o[row][0] = q0;
                    o[row][1] = q1;
                }
                double2* Kt = Kbuf + ((size_t)bl * F + s_list[bt * batch + e]) * (WIDE ? (size_t)kt_d2 : (size_t)KT_D2);
                if (i == k) { o[0][0] += s2; o[1][1] += s2; }
                if (i >= k) store_block(Kt, i, k, o[0][0], o[0][1], o[1][0], o[1][1]);
                else store_block(Kt, k, i, o[0][0], o[1][0], o[0][1], o[1][1]);      // block above the diagonal: its transpose below it
            }
        }
        __syncthreads();      // the buffer is free for batch bt + 2
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
#pragma unroll 1      // rolled on purpose: 81 instantiations (tile column x tile count) of the unrolled chain overflow the
                      // instruction cache as soon as the tracks of a launch differ in length (window 31: 11.8 vs 8.9 ms)
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
        const int rank = fac_rank[fj], dofv = dof[fj];
        const int st = feat_status[fj];
        const double gam = gamma_from_last_tile(D0, D1, g, t, rr, rank);
        if (lane == 0) {
            const double crit = (dofv < chi2_len) ? chi2_table[dofv] : nan("");
            const bool pass = gam <= crit;
            feat_status[fj] = (unsigned char)((pass ? MSCKF_FEAT_ACCEPTED : MSCKF_FEAT_GATED_OUT) | (st & FLAG_KEEP));
            gamma[fj] = gam;
        }
    }
}

int front_asm_warps(const msckf_dims& d) {
    const size_t p_bytes = (size_t)(d.N * (d.N + 1) / 2) * PB * sizeof(double);
    const size_t list_bytes = ((size_t)d.F * sizeof(unsigned short) + 15) & ~(size_t)15;
    const size_t per_warp = (size_t)2 * FW_MAX_OBS * AR * sizeof(double);
    const size_t cap = 227 * 1024;
    if (p_bytes + list_bytes + 2 * per_warp > cap) return 0;
    size_t w = (cap - p_bytes - list_bytes) / per_warp;
    return (int)(w > (size_t)FA_WARPS_MAX ? FA_WARPS_MAX : w);
}

size_t front_asm_smem(const msckf_dims& d, int warps) {
    return (size_t)(d.N * (d.N + 1) / 2) * PB * sizeof(double) + (size_t)warps * 2 * FW_MAX_OBS * AR * sizeof(double) +
           (((size_t)d.F * sizeof(unsigned short) + 15) & ~(size_t)15);
}

}  // namespace

static bool front_narrow(const msckf_dims& d) {
    // (the start-up histogram of the shared-memory assembly kernel aliases two warps' scratch; 16-bit feature indices)
    return d.M <= fgeom::FW_MAX_OBS && d.N <= 32 && d.F <= 65535 && front_asm_warps(d) >= 2;
}

// tracks of up to 64 observations / windows of up to 64 poses: geometry by front.cu's CTA-per-feature kernel, P-stationary
// assembly over several CTAs per filter, shared-memory factorisation (front_chol_smem.cu)
static bool front_wide(const msckf_dims& d) { return d.M <= 64 && d.N <= 64 && d.F <= 65535; }

bool front_split_supported(const msckf_dims& d) { return front_narrow(d) || front_wide(d); }

// double2 per feature slot of the tile buffer
int front_split_kt_d2(const msckf_dims& d) { return front_narrow(d) ? KT_D2 : ftile::tiles_for_capacity(d.M) * 32; }
int front_split_map_bytes(const msckf_dims& d) { return front_narrow(d) ? 32 : 64; }

// filters handled per kernel triple: bounds the tile buffer (23 KB per feature for short tracks, 78 KB at 64 observations)
int front_split_sub_filters(const msckf_dims& d) {
#ifndef MSCKF_TILE_BUDGET_GB
#define MSCKF_TILE_BUDGET_GB 32
#endif
    const size_t budget = (size_t)MSCKF_TILE_BUDGET_GB << 30;
    const size_t per_filter = (size_t)(d.F > 0 ? d.F : 1) * front_split_kt_d2(d) * sizeof(double2);
    long long sub = (long long)(budget / per_filter);
    if (sub < 1) sub = 1;
    if (sub >= d.B) return d.B > 0 ? d.B : 1;
    // the assembly kernel runs one CTA (= filter) per SM: whole waves of the 148 SMs per sub-batch (820 filters = 5.5 waves
    // cost as much as 6; 888 = 6 waves exactly), the remainder in the last one
    constexpr int SMS = 148;
    if (sub >= SMS) return (int)(sub / SMS) * SMS;
    const int n_sub = (int)((d.B + sub - 1) / sub);
    return (d.B + n_sub - 1) / n_sub;
}

size_t front_split_tile_bytes(const msckf_dims& d) {
    return (size_t)front_split_sub_filters(d) * d.F * front_split_kt_d2(d) * sizeof(double2);
}

// optional per-kernel timing (msckf_set_stage_timing / msckf_front_stage_ms): CUDA events on the launching stream around
// the three kernels of every sub-batch
namespace {
constexpr int MAX_SUB = 64;
bool g_timing = false;
cudaEvent_t g_ev[MAX_SUB][4];
int g_ev_created = 0, g_ev_used = 0;
}  // namespace

extern "C" int msckf_set_stage_timing(int on) {
    g_timing = on != 0;
    g_ev_used = 0;
    return 0;
}

extern "C" int msckf_front_stage_ms(double* geo_ms, double* asm_ms, double* chol_ms) {
    double g = 0.0, a = 0.0, c = 0.0;
    for (int k = 0; k < g_ev_used; ++k) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
        cudaError_t e = cudaEventSynchronize(g_ev[k][3]);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t0, g_ev[k][0], g_ev[k][1]);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t1, g_ev[k][1], g_ev[k][2]);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t2, g_ev[k][2], g_ev[k][3]);
        if (e != cudaSuccess) return (int)e;
        g += t0;
        a += t1;
        c += t2;
    }
    if (geo_ms) *geo_ms = g;
    if (asm_ms) *asm_ms = a;
    if (chol_ms) *chol_ms = c;
    return g_ev_used;
}

int front_ps_batch(const msckf_dims& d) {
    const size_t cap = 227 * 1024;
    const size_t lists = (((size_t)d.F + 7) & ~(size_t)7) * 2 + (((size_t)d.F + 15) & ~(size_t)15) + 1024;
    if (lists + 2 * 2 * (size_t)PS_FEAT_BYTES > cap) return 0;
    size_t nb = (cap - lists) / (2 * (size_t)PS_FEAT_BYTES);
    return (int)(nb > 32 ? 32 : nb);
}

template <int NWARPS>
cudaError_t launch_asm_ps(const FrontArgs& a, const unsigned char* slot_map, double2* Kbuf, int b0, int nb, cudaStream_t s) {
    const int batch = front_ps_batch(a.dims);
    const size_t smem = (size_t)2 * batch * PS_FEAT_BYTES + (((size_t)a.dims.F + 7) & ~(size_t)7) * 2 +
                        (((size_t)a.dims.F + 15) & ~(size_t)15);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(front_asm_ps_kernel<NWARPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    // few filters: several CTAs per filter, each taking every nz-th batch of its pending features (<= one CTA per SM in all)
    int nz = 1;
    if (nb < 148) {
        const int nbatch = (a.dims.F + batch - 1) / batch;
        nz = 148 / (nb > 0 ? nb : 1);
        if (nz > nbatch) nz = nbatch;
        if (nz < 1) nz = 1;
    }
    if (nz > 1) {
        static size_t configured_z = 0;
        if (smem > configured_z) {
            cudaError_t e = cudaFuncSetAttribute(front_asm_ps_kernel<NWARPS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured_z = smem;
        }
        front_asm_ps_kernel<NWARPS, false, true><<<dim3(nb, 1, nz), NWARPS * 32, smem, s>>>(a, slot_map, Kbuf, b0, batch, KT_D2);
    } else {
        front_asm_ps_kernel<NWARPS, false, false><<<nb, NWARPS * 32, smem, s>>>(a, slot_map, Kbuf, b0, batch, KT_D2);
    }
    return cudaGetLastError();
}

cudaError_t launch_front_geo_wide(const FrontArgs& a, double2* Kbuf, int kt_d2, unsigned char* slot_map, int* pend_list,
                                  int* pend_count, int b0, int nb, cudaStream_t s);
cudaError_t launch_front_chol_smem(const double2* Kbuf, int kt_d2, const int* pend_list, const int* pend_count, const int* feat_len,
                                   const int* fac_rank, const int* dof, const double* chi2_table, int chi2_len,
                                   unsigned char* feat_status, double* gamma, size_t fbase, int Mcap, cudaStream_t s);

static cudaError_t launch_front_wide(const FrontArgs& a, double2* Kbuf, int* pend_list, int* pend_count, unsigned char* slot_map,
                                     cudaStream_t s) {
    constexpr int NW = 15;
    constexpr size_t featb = 64 * AR * 8 + 64;
    const int F = a.dims.F;
    const int kt_d2 = front_split_kt_d2(a.dims);
    const size_t lists = (((size_t)F + 7) & ~(size_t)7) * 2 + (((size_t)F + 15) & ~(size_t)15);
    int batch = (int)(((size_t)227 * 1024 - lists - 1024) / (2 * featb));
    if (batch > 32) batch = 32;
    if (batch < 1) return cudaErrorInvalidValue;
    const size_t smem = (size_t)2 * batch * featb + lists;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(front_asm_ps_kernel<NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const int nblk = a.dims.N * (a.dims.N + 1) / 2;
    const int parts = (nblk + NW * 32 - 1) / (NW * 32);
    const int sub = front_split_sub_filters(a.dims);
    if (g_timing) g_ev_used = 0;
    for (int b0 = 0; b0 < a.dims.B; b0 += sub) {
        const int nb = (b0 + sub <= a.dims.B) ? sub : a.dims.B - b0;
        cudaError_t e = cudaMemsetAsync(pend_count, 0, sizeof(int), s);
        if (e != cudaSuccess) return e;
        const bool timed = g_timing && g_ev_used < MAX_SUB;
        const int slot = g_ev_used;
        if (timed) {
            for (; g_ev_created <= slot; ++g_ev_created)
                for (int q = 0; q < 4; ++q)
                    if ((e = cudaEventCreate(&g_ev[g_ev_created][q])) != cudaSuccess) return e;
            cudaEventRecord(g_ev[slot][0], s);
        }
        e = launch_front_geo_wide(a, Kbuf, kt_d2, slot_map, pend_list, pend_count, b0, nb, s);
        if (e != cudaSuccess) return e;
        if (timed) cudaEventRecord(g_ev[slot][1], s);
        front_asm_ps_kernel<NW, true><<<dim3(nb, parts), NW * 32, smem, s>>>(a, slot_map, Kbuf, b0, batch, kt_d2);
        msckf_count_launch(1);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (timed) cudaEventRecord(g_ev[slot][2], s);
        e = launch_front_chol_smem(Kbuf, kt_d2, pend_list, pend_count, a.feat_len, a.fac_rank, a.dof, a.chi2_table, a.chi2_len,
                                   a.feat_status, a.gamma, (size_t)b0 * F, a.dims.M, s);
        if (e != cudaSuccess) return e;
        if (timed) {
            cudaEventRecord(g_ev[slot][3], s);
            ++g_ev_used;
        }
    }
    return cudaSuccess;
}

cudaError_t launch_front_split(const FrontArgs& a, double2* Kbuf, int* pend_list, int* pend_count, int* pend_key,
                               unsigned char* slot_map, cudaStream_t s) {
    if (!front_narrow(a.dims)) return launch_front_wide(a, Kbuf, pend_list, pend_count, slot_map, s);
    // P-stationary assembly when the block triangle of the window fits one CTA's threads and the batch buffers fit its
    // shared memory; P staged in shared memory otherwise
    const int nblk = a.dims.N * (a.dims.N + 1) / 2;
    const int ps_warps = (front_ps_batch(a.dims) >= 2) ? (nblk <= 15 * 32 ? 15 : (nblk <= 16 * 32 ? 16 : 0)) : 0;
    const int warps = front_asm_warps(a.dims);
    const size_t smem = front_asm_smem(a.dims, warps);
    static size_t configured = 0;
    if (!ps_warps && smem > configured) {
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
                for (int q = 0; q < 4; ++q)
                    if ((e = cudaEventCreate(&g_ev[g_ev_created][q])) != cudaSuccess) return e;
            cudaEventRecord(g_ev[slot][0], s);
        }
        {   // geometry: 8 features per warp and CTA keep the prefetch chain going; small CTAs fill the machine at any B
            const int fpc = 8 * FG_WARPS;
            dim3 grid((F + fpc - 1) / fpc, nb);
            const size_t gsm = (size_t)a.dims.N * 12 * sizeof(double);
            front_geo_kernel<<<grid, FG_THREADS, gsm, s>>>(a, pend_key, slot_map, Kbuf, pend_list, pend_count, b0, fpc);
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
        if (timed) cudaEventRecord(g_ev[slot][1], s);
        if (ps_warps == 15) e = launch_asm_ps<15>(a, slot_map, Kbuf, b0, nb, s);
        else if (ps_warps == 16) e = launch_asm_ps<16>(a, slot_map, Kbuf, b0, nb, s);
        else {
            front_asm_kernel<<<nb, warps * 32, smem, s>>>(a, pend_key, Kbuf, b0);
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) return e;
        if (timed) cudaEventRecord(g_ev[slot][2], s);
        front_chol_kernel<<<148 * 4, FC_THREADS, 0, s>>>(Kbuf, pend_list, pend_count, a.feat_len, a.fac_rank, a.dof,
                                                        a.chi2_table, a.chi2_len, a.feat_status, a.gamma, (size_t)b0 * F);
        msckf_count_launch(3);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (timed) {
            cudaEventRecord(g_ev[slot][3], s);
            ++g_ev_used;
        }
    }
    return cudaSuccess;
}
