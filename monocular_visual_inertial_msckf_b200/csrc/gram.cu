// Row 9 of SURVEY.md section 8a, information-form flavour (MSCKF_COMPRESS_GRAM): instead of "vstack the surviving H_o
// blocks and QR them" (reference src/msckf/MSCKF.py:587-598) the kernel accumulates, per filter, the information matrix
// and vector of the same stacked system,
//     G = sum_j H_o,j^T H_o,j = sum_j (D_j - W_j^T W_j),        g = sum_j H_o,j^T r_o,j = sum_j H_x,j^T r~_j,
// with D_j = blockdiag_slot(A_i^T A_i) (6 x 6 per observing camera), W_j = Q1_j^T H_x,j (3 x 6N, rank(H_f) rows) and
// r~_j = (I - Q1 Q1^T) r_j -- the basis-free form of the left-null-space projection (A A^T = I - Q1 Q1^T).  G is what
// T_H^T T_H of the reference's QR equals, g is T_H^T r_n; the update kernel (update_info.cu) consumes them directly.
//
// One 256-thread CTA per filter.  Per panel of PF accepted features:
//   * cp.async (LDGSTS) prefetches the next panel's compact factors (A_i, Q1 rows, r~_i, slots) into shared memory;
//   * the 3 PF rows of W (zero where a camera does not observe the feature) are scattered into a shared K-panel;
//   * every warp folds the panel into its share of the lower-triangular 8 x 8 output tiles of W^T W with FP64 tensor
//     instructions (DMMA.8x8x4), accumulators resident in registers for the whole filter (<= 40 tiles per warp);
//   * D and g are accumulated by owner threads in registers in feature order (deterministic, no atomics).
// The stacked matrix H_X (22.8 MB per filter at 256 features x 30 poses) is never formed.
#include "msckf_common.cuh"
#include "dense.cuh"

namespace {

constexpr int GT = 256;        // threads per CTA (8 warps, 2 per SM sub-partition)
constexpr int UMAX = 20;       // (2 tile rows x 1 tile column) units per warp per pass -> 40 accumulator tiles
constexpr int IMAX = 7;        // D/g items per thread: 27 N_cams / 256 <= 7 for N <= 64

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

struct GramSmem {
    double* panel;      // [3 PF][LD]
    double* rawA[2];    // [PF][Mcap][12]
    double* rawQ[2];    // [PF][Mcap][6]
    double* rawR[2];    // [PF][Mcap][2]
    int* rawS[2];       // [PF][Mcap]
    double* Dg;         // [Ncap][27]
    int* acc_list;      // [F]
    int* flen;          // [2][PF]
    signed char* inv;   // [PF][Ncap]
};

__device__ __forceinline__ int strip_len(int p, int T) { return min(2 * p + 2, T); }

__global__ void __launch_bounds__(GT, 1) gram_kernel(const CompressArgs a, const int PF, const int LD) {
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = a.dims.F, Mcap = a.dims.M, Ncap = a.dims.N;
    const int nc = a.n_cams[b];
    const int c = 6 * nc;
    const int ccap = 6 * Ncap;
    const int T = (c + 7) >> 3;            // 8 x 8 output tiles per side
    const int T2 = (T + 1) >> 1;           // 2-row strips

    extern __shared__ double smem[];
    GramSmem S;
    {
        double* p = smem;
        S.panel = p;            p += (size_t)3 * PF * LD;
        for (int s = 0; s < 2; ++s) { S.rawA[s] = p; p += (size_t)PF * Mcap * 12; }
        for (int s = 0; s < 2; ++s) { S.rawQ[s] = p; p += (size_t)PF * Mcap * 6; }
        for (int s = 0; s < 2; ++s) { S.rawR[s] = p; p += (size_t)PF * Mcap * 2; }
        S.Dg = p;               p += (size_t)Ncap * 27;
        int* q = (int*)p;
        for (int s = 0; s < 2; ++s) { S.rawS[s] = q; q += PF * Mcap; }
        S.acc_list = q;         q += F;
        S.flen = q;             q += 2 * PF;
        S.inv = (signed char*)q;
    }
    __shared__ int s_wcnt[GT / 32];
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
        if (ok) S.acc_list[off + __popc(bal & ((1u << lane) - 1u))] = j;
        for (int w = 0; w < GT / 32; ++w) nacc += s_wcnt[w];
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
    const int npass = (U + 8 * UMAX - 1) / (8 * UMAX);
    const int upp = (U + npass - 1) / npass;

    // asynchronous prefetch of one panel's compact factors
    auto prefetch = [&](int pi, int st) {
        for (int f = 0; f < PF; ++f) {
            const int k = pi * PF + f;
            const int j = (k < nacc) ? S.acc_list[k] : -1;
            const int M = (j >= 0) ? min(a.feat_len[(size_t)b * F + j], Mcap) : 0;
            if (tid == 0) S.flen[st * PF + f] = M;
            if (M <= 0) continue;
            const size_t fj = (size_t)b * F + j;
            const double* gA = a.fac_A + fj * (size_t)Mcap * 12;
            const double* gQ = a.fac_Q + fj * (size_t)Mcap * 6;
            const double* gR = a.fac_rt + fj * (size_t)Mcap * 2;
            const int* gS = a.obs_slot + fj * (size_t)Mcap;
            double* sA = S.rawA[st] + (size_t)f * Mcap * 12;
            double* sQ = S.rawQ[st] + (size_t)f * Mcap * 6;
            double* sR = S.rawR[st] + (size_t)f * Mcap * 2;
            int* sS = S.rawS[st] + f * Mcap;
            for (int it = tid; it < 11 * M; it += GT) {
                if (it < 6 * M) cp_async16(sA + 2 * it, gA + 2 * it);
                else if (it < 9 * M) cp_async16(sQ + 2 * (it - 6 * M), gQ + 2 * (it - 6 * M));
                else if (it < 10 * M) cp_async16(sR + 2 * (it - 9 * M), gR + 2 * (it - 9 * M));
                else cp_async4(sS + (it - 10 * M), gS + (it - 10 * M));
            }
        }
        cp_async_commit();
    };

    double dg[IMAX];
    const int nitems = nc * 27;

    for (int pass = 0; pass < npass; ++pass) {
        // this warp's contiguous range of units in the linearised (strip, column) order
        const int ubeg = pass * upp, uend = min(U, ubeg + upp);
        const int upw = (uend - ubeg + 7) >> 3;
        const int u0 = min(uend, ubeg + warp * upw);
        const int nu = min(uend, u0 + upw) - u0;
        int p0 = 0, J0 = u0;
        while (p0 < T2 && J0 >= strip_len(p0, T)) { J0 -= strip_len(p0, T); ++p0; }

        double acc[UMAX][2][2];
#pragma unroll
        for (int u = 0; u < UMAX; ++u) acc[u][0][0] = acc[u][0][1] = acc[u][1][0] = acc[u][1][1] = 0.0;
        if (pass == 0) {
#pragma unroll
            for (int k = 0; k < IMAX; ++k) dg[k] = 0.0;
        }

        __syncthreads();
        prefetch(0, 0);
        cp_async_wait_all();
        __syncthreads();

        for (int pi = 0; pi < npanels; ++pi) {
            const int st = pi & 1;
            if (pi + 1 < npanels) prefetch(pi + 1, st ^ 1);
            // ---- zero the K-panel and the slot -> observation map ----
            for (int e = tid; e < 3 * PF * LD; e += GT) S.panel[e] = 0.0;
            for (int e = tid; e < PF * Ncap; e += GT) S.inv[e] = -1;
            __syncthreads();
            // ---- scatter W rows: W[q][6 s + e] = Q1[2i][q] A[2i][e] + Q1[2i+1][q] A[2i+1][e] ----
            for (int it = tid; it < PF * Mcap; it += GT) {
                const int f = it / Mcap, i = it - f * Mcap;
                if (i >= S.flen[st * PF + f]) continue;
                const int s = S.rawS[st][f * Mcap + i];
                if (s < 0 || s >= nc) continue;
                const double* A = S.rawA[st] + ((size_t)f * Mcap + i) * 12;
                const double* Q = S.rawQ[st] + ((size_t)f * Mcap + i) * 6;
                S.inv[f * Ncap + s] = (signed char)i;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    double* dst = S.panel + (size_t)(3 * f + q) * LD + 6 * s;
                    const double q0 = Q[q], q1 = Q[3 + q];
#pragma unroll
                    for (int e = 0; e < 6; ++e) dst[e] = q0 * A[e] + q1 * A[6 + e];
                }
            }
            __syncthreads();
            // ---- W^T W on the tensor pipe ----
            {
                const int g = lane >> 2, t = lane & 3;
                for (int kk = 0; kk < 3 * PF; kk += 4) {
                    const double* row = S.panel + (size_t)(kk + t) * LD + g;
                    int p = p0, J = J0;
                    double a0 = row[16 * p], a1 = row[16 * p + 8];
#pragma unroll
                    for (int u = 0; u < UMAX; ++u) {
                        if (u < nu) {
                            const double bf = row[8 * J];
                            dense::dmma(acc[u][0][0], acc[u][0][1], a0, bf);
                            dense::dmma(acc[u][1][0], acc[u][1][1], a1, bf);
                            if (++J == strip_len(p, T)) {
                                J = 0;
                                ++p;
                                if (p < T2) { a0 = row[16 * p]; a1 = row[16 * p + 8]; }
                            }
                        }
                    }
                }
            }
            // ---- D = sum A_i^T A_i (6 x 6 lower per slot) and g = sum A_i^T r~_i, in feature order ----
            if (pass == 0) {
#pragma unroll
                for (int k = 0; k < IMAX; ++k) {
                    const int item = tid + k * GT;
                    if (item < nitems) {
                        const int s = item / 27, e = item - 27 * s;
                        int ea = 0, eb = 0;
                        if (e < 21) { while ((ea + 1) * (ea + 2) / 2 <= e) ++ea; eb = e - ea * (ea + 1) / 2; } else ea = e - 21;
                        double v = dg[k];
                        for (int f = 0; f < PF; ++f) {
                            const int i = S.inv[f * Ncap + s];
                            if (i < 0) continue;
                            const double* A = S.rawA[st] + ((size_t)f * Mcap + i) * 12;
                            if (e < 21) v += A[ea] * A[eb] + A[6 + ea] * A[6 + eb];
                            else {
                                const double* R = S.rawR[st] + ((size_t)f * Mcap + i) * 2;
                                v += A[ea] * R[0] + A[6 + ea] * R[1];
                            }
                        }
                        dg[k] = v;
                    }
                }
            }
            cp_async_wait_all();
            __syncthreads();
        }

        if (pass == 0) {
#pragma unroll
            for (int k = 0; k < IMAX; ++k) {
                const int item = tid + k * GT;
                if (item < nitems) S.Dg[item] = dg[k];
            }
            __syncthreads();
            for (int e = tid; e < ccap; e += GT) {
                const int s = e / 6, aa = e - 6 * s;
                gout[e] = (s < nc) ? S.Dg[s * 27 + 21 + aa] : 0.0;
            }
        }
        // ---- G = D - W^T W: lower tiles written with their mirror image ----
        {
            const int g = lane >> 2, t = lane & 3;
            int p = p0, J = J0;
#pragma unroll
            for (int u = 0; u < UMAX; ++u) {
                if (u < nu) {
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
                                        v += S.Dg[s * 27 + ea * (ea + 1) / 2 + eb];
                                    }
                                    Gout[(size_t)r * ccap + col] = v;
                                    Gout[(size_t)col * ccap + r] = v;
                                }
                            }
                        }
                    }
                    if (++J == strip_len(p, T)) { J = 0; ++p; }
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

struct GramPlan { int PF, LD; size_t smem; };

GramPlan gram_plan(const msckf_dims& d) {
    GramPlan best{0, 0, 0};
    const int T = (6 * d.N + 7) / 8, T2 = (T + 1) / 2;
    const int LD = 16 * T2 + 8;                        // LD mod 16 == 8: conflict-free DMMA fragment loads
    for (int PF = 8; PF >= 1; PF >>= 1) {
        if ((3 * PF) % 4 != 0 && PF != 4 && PF != 8) continue;
        size_t doubles = (size_t)3 * PF * LD + (size_t)2 * PF * d.M * 20 + (size_t)d.N * 27;
        size_t bytes = doubles * 8 + ((size_t)2 * PF * d.M + d.F + 2 * PF) * 4 + (size_t)PF * d.N + 64;
        if (bytes <= 227 * 1024) { best = GramPlan{PF, LD, bytes}; break; }
    }
    return best;
}

}  // namespace

bool gram_supported(const msckf_dims& d) { return gram_plan(d).PF >= 4; }

cudaError_t launch_compress_gram(const CompressArgs& a, cudaStream_t s) {
    const GramPlan pl = gram_plan(a.dims);
    if (pl.PF < 4) return cudaErrorInvalidConfiguration;
    static size_t configured = 0;
    if (pl.smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e != cudaSuccess) return e;
        configured = pl.smem;
    }
    gram_kernel<<<a.dims.B, GT, pl.smem, s>>>(a, pl.PF, pl.LD);
    msckf_count_launch(1);
    return cudaGetLastError();
}
