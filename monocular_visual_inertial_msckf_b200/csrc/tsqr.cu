// Row 9 of SURVEY.md section 8a, QR flavour: "vstack the surviving H_o blocks, np.linalg.qr, T_H = R, r_n = Q^T r_o"
// (reference src/msckf/MSCKF.py:587-598) -- done WITHOUT ever materialising the stacked matrix H_X (22.8 MB per filter at
// 256 features x 30 poses).  One CTA per filter folds one feature at a time into a resident upper-triangular R:
//
//   * the rows a feature contributes are regenerated on the fly from its compact factors as
//         H~ = blockdiag_slot(A_i) - Q1 (Q1^T H_x)        (2M x 6N),      r~ = (I - Q1 Q1^T) r,
//     which has the same Gram matrix / same R factor as the reference's null-space-projected block A^T H_x
//     (A A^T = I - Q1 Q1^T), so T_H and r_n are the reference's up to the sign convention of each row of R;
//   * the block lives in REGISTERS, one matrix column per 4 lanes (each lane holds both rows of every 4th
//     observation), the right-hand side is carried as column 6N;
//   * R (6N+1 columns, packed upper triangle, <= 140 KB at N = 31) lives in SHARED memory;
//   * each Householder step: the 4 owner lanes build the reflector from R[k,k] and their column (two shuffles),
//     publish v through a double-buffered shared vector, one __syncthreads, then every column j > k does a
//     16-term dot product + 2 shuffles + rank-1 update of its registers and of R[k,j].
#include "msckf_common.cuh"

namespace {

constexpr int OPL = 8;                 // observations per lane  -> 4 lanes cover a chunk of 32 observations (64 rows)
constexpr int CHUNK_OBS = 4 * OPL;
constexpr int CHUNK_ROWS = 2 * CHUNK_OBS;

__device__ __forceinline__ int roff(int k, int NC) { return k * NC - ((k * (k - 1)) >> 1); }

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) tsqr_kernel(const CompressArgs a) {
    const int b = blockIdx.x;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int lane = tid & 31;
    const int F = a.dims.F, Mcap = a.dims.M, Ncap = a.dims.N;
    const int nc = a.n_cams[b];
    const int c = 6 * nc, NC = c + 1;
    const int ccap = 6 * Ncap;
    const int col = tid >> 2, q = tid & 3;
    const unsigned mask4 = 0xFu << (lane & 28);

    extern __shared__ double smem[];
    double* Rs = smem;                                        // packed upper triangle, rows of length NC - k
    const int nR = ((6 * Ncap + 1) * (6 * Ncap + 2)) / 2;
    double* vbuf = Rs + nR + (nR & 1);                        // [2][CHUNK_ROWS], 16-byte aligned
    double* taub = vbuf + 2 * CHUNK_ROWS;                     // [2]
    double* fA = taub + 2;                                    // [2 Mcap][6]
    double* fQ = fA + 2 * Mcap * 6;                           // [2 Mcap][3]
    double* frt = fQ + 2 * Mcap * 3;                          // [2 Mcap]
    int* inv = (int*)(frt + 2 * Mcap);                        // [Ncap] slot -> observation index of this feature
    __shared__ int s_cnt[2];

    for (int e = tid; e < nR; e += nth) Rs[e] = 0.0;
    if (tid < 2) s_cnt[tid] = 0;
    __syncthreads();
    {
        int acc = 0, rej = 0;
        for (int j = tid; j < F; j += nth) {
            const int st = a.feat_status[(size_t)b * F + j] & 0xF;
            acc += (st == MSCKF_FEAT_ACCEPTED);
            rej += (st == MSCKF_FEAT_GATED_OUT);
        }
        if (acc) atomicAdd(&s_cnt[0], acc);
        if (rej) atomicAdd(&s_cnt[1], rej);
    }

    for (int j = 0; j < F; ++j) {
        const size_t fj = (size_t)b * F + j;
        if ((a.feat_status[fj] & 0xF) != MSCKF_FEAT_ACCEPTED) continue;   // CTA-uniform
        const int M = a.feat_len[fj];
        __syncthreads();   // previous feature fully folded before its factors are overwritten
        {
            const double* gA = a.fac_A + fj * (size_t)Mcap * 12;
            const double* gQ = a.fac_Q + fj * (size_t)Mcap * 6;
            const double* gr = a.fac_rt + fj * (size_t)Mcap * 2;
            for (int e = tid; e < 12 * M; e += nth) fA[e] = gA[e];
            for (int e = tid; e < 6 * M; e += nth) fQ[e] = gQ[e];
            for (int e = tid; e < 2 * M; e += nth) frt[e] = gr[e];
            for (int e = tid; e < nc; e += nth) inv[e] = -1;
        }
        __syncthreads();
        {
            const int* gs = a.obs_slot + fj * (size_t)Mcap;
            for (int e = tid; e < M; e += nth) {
                const int s = gs[e];
                if (s >= 0 && s < nc) inv[s] = e;
            }
        }
        __syncthreads();

        // column constants: W[:, col] = Q1[rows of i_s]^T A_{i_s}[:, e6]
        double W0 = 0.0, W1 = 0.0, W2 = 0.0;
        int i_s = -1, e6 = 0;
        if (col < c) {
            const int s = col / 6;
            e6 = col - 6 * s;
            i_s = inv[s];
            if (i_s >= 0) {
                const double a0 = fA[(2 * i_s) * 6 + e6], a1 = fA[(2 * i_s + 1) * 6 + e6];
                W0 = fQ[(2 * i_s) * 3 + 0] * a0 + fQ[(2 * i_s + 1) * 3 + 0] * a1;
                W1 = fQ[(2 * i_s) * 3 + 1] * a0 + fQ[(2 * i_s + 1) * 3 + 1] * a1;
                W2 = fQ[(2 * i_s) * 3 + 2] * a0 + fQ[(2 * i_s + 1) * 3 + 2] * a1;
            }
        }

        for (int chunk0 = 0; chunk0 < M; chunk0 += CHUNK_OBS) {
            // ---- regenerate this chunk of the feature's rows into registers ----
            double Bv[2 * OPL];
#pragma unroll
            for (int e = 0; e < OPL; ++e) {
                const int i = chunk0 + q + 4 * e;
                double v0 = 0.0, v1 = 0.0;
                if (i < M && col < NC) {
                    if (col < c) {
                        const double* q0 = fQ + (2 * i) * 3;
                        v0 = -(q0[0] * W0 + q0[1] * W1 + q0[2] * W2);
                        v1 = -(q0[3] * W0 + q0[4] * W1 + q0[5] * W2);
                        if (i == i_s) {
                            v0 += fA[(2 * i) * 6 + e6];
                            v1 += fA[(2 * i + 1) * 6 + e6];
                        }
                    } else {
                        v0 = frt[2 * i];
                        v1 = frt[2 * i + 1];
                    }
                }
                Bv[2 * e] = v0;
                Bv[2 * e + 1] = v1;
            }

            // ---- fold the chunk into R: c Householder steps ----
            for (int k = 0; k < c; ++k) {
                double* vb = vbuf + (k & 1) * CHUNK_ROWS;
                if (col == k) {
                    double s = 0.0;
#pragma unroll
                    for (int e = 0; e < 2 * OPL; ++e) s += Bv[e] * Bv[e];
                    s += __shfl_xor_sync(mask4, s, 1);
                    s += __shfl_xor_sync(mask4, s, 2);
                    const int o = roff(k, NC);
                    const double alpha = Rs[o];
                    double tau = 0.0, scal = 0.0;
                    if (s > 0.0) {
                        const double beta = -copysign(sqrt(alpha * alpha + s), alpha);
                        tau = (beta - alpha) / beta;
                        scal = 1.0 / (alpha - beta);
                        if (q == 0) Rs[o] = beta;
                    }
                    if (q == 0) taub[k & 1] = tau;
#pragma unroll
                    for (int e = 0; e < OPL; ++e) {
                        double2 v2;
                        v2.x = Bv[2 * e] * scal;
                        v2.y = Bv[2 * e + 1] * scal;
                        *reinterpret_cast<double2*>(vb + 2 * (q + 4 * e)) = v2;
                    }
                }
                __syncthreads();
                const double tau = taub[k & 1];
                // warps whose 8 columns are all <= k have nothing left to do in this fold (warp-uniform test)
                if (tau != 0.0 && ((tid | 31) >> 2) > k) {
                    const bool active = (col > k) && (col < NC);
                    const int o = roff(k, NC) + (col - k);
                    double dot = 0.0;
                    const double2* vp = reinterpret_cast<const double2*>(vb) + q;
                    if (active) {
                        double d0 = 0.0, d1 = 0.0;
#pragma unroll
                        for (int e = 0; e < OPL; ++e) {
                            const double2 vv = vp[4 * e];
                            d0 += vv.x * Bv[2 * e];
                            d1 += vv.y * Bv[2 * e + 1];
                        }
                        dot = d0 + d1;
                        if (q == 0) dot += Rs[o];
                    }
                    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                    if (active) {
                        const double w = tau * dot;
                        if (q == 0) Rs[o] -= w;
#pragma unroll
                        for (int e = 0; e < OPL; ++e) {
                            const double2 vv = vp[4 * e];
                            Bv[2 * e] -= w * vv.x;
                            Bv[2 * e + 1] -= w * vv.y;
                        }
                    }
                }
            }
            __syncthreads();   // vbuf / R row reuse across chunks
        }
    }
    __syncthreads();

    // ---- T_H (upper triangular, zero padded to the capacity) and r_n ----
    double* T = a.T_H + (size_t)b * ccap * ccap;
    double* rn = a.r_n + (size_t)b * ccap;
    for (int e = tid; e < ccap * ccap; e += nth) {
        const int r = e / ccap, cc = e - r * ccap;
        double v = 0.0;
        if (r < c && cc < c && cc >= r) v = Rs[roff(r, NC) + (cc - r)];
        T[e] = v;
    }
    for (int r = tid; r < ccap; r += nth) rn[r] = (r < c) ? Rs[roff(r, NC) + (c - r)] : 0.0;
    if (tid == 0) {
        a.n_accepted[b] = s_cnt[0];
        a.n_rejected[b] = s_cnt[1];
    }
}

}  // namespace

static size_t tsqr_smem_bytes(int Ncap, int Mcap) {
    const size_t nR = ((size_t)(6 * Ncap + 1) * (6 * Ncap + 2)) / 2;
    const size_t doubles = nR + (nR & 1) + 2 * CHUNK_ROWS + 2 + (size_t)2 * Mcap * 10;
    return doubles * sizeof(double) + (size_t)Ncap * sizeof(int) + 16;
}

int tsqr_max_cams() {
    int n = 1;
    while (tsqr_smem_bytes(n + 1, 32) <= 227 * 1024 && 4 * (6 * (n + 1) + 1) <= 1024) ++n;
    return n;
}

cudaError_t launch_compress_qr(const CompressArgs& a, cudaStream_t s) {
    const size_t smem = tsqr_smem_bytes(a.dims.N, a.dims.M);
    int threads = 4 * (6 * a.dims.N + 1);
    threads = (threads + 31) & ~31;
    static size_t configured[2] = {0, 0};
    const int v = threads <= 768 ? 0 : 1;
    if (smem > configured[v]) {
        cudaError_t e = v == 0 ? cudaFuncSetAttribute(tsqr_kernel<768>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                               : cudaFuncSetAttribute(tsqr_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[v] = smem;
    }
    if (v == 0) tsqr_kernel<768><<<a.dims.B, threads, smem, s>>>(a);
    else tsqr_kernel<1024><<<a.dims.B, threads, smem, s>>>(a);
    msckf_count_launch(1);
    return cudaGetLastError();
}
