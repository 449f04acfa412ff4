// Track descriptors of the matching front with the reference's semantics:
//   score filter of the frame's detections                 reference src/msckf/MSCKF.py:279-284
//   per-observation descriptor / score lists of a track    Feature.descriptors / Feature.scores, MSCKF.py:407-408, 425-427
//   their deletion when a camera leaves the window          MSCKF.py:764-765
//   the score-weighted mean descriptor the matcher sees     np.average(feature.descriptors, axis=0, weights=feature.scores),
//                                                           MSCKF.py:436-441 (NOT re-normalised)
// Descriptors are float32 as XFeat delivers them; the per-observation scores are the tracks' obs_conf (Line.confidence is
// the same number, MSCKF.py:410).  The mean is accumulated in FP64 and rounded once.  All kernels are one pass over the
// arrays they touch (HBM-bound by construction).
#include "msckf_common.cuh"

namespace {

constexpr int DT = 256;

// ---- score filter: keep detections with score >= 0.5 mean(score), in order (MSCKF.py:279-284) ----
__global__ void __launch_bounds__(DT) score_filter_kernel(int F2, int D, const int* __restrict__ n_in,
                                                          const double* __restrict__ scores, const double* __restrict__ kp,
                                                          const float* __restrict__ desc, double* __restrict__ scores_out,
                                                          double* __restrict__ kp_out, float* __restrict__ desc_out,
                                                          int* __restrict__ n_out) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = n_in ? min(max(n_in[b], 0), F2) : F2;
    __shared__ double s_part[DT / 32];
    __shared__ int s_cnt[DT / 32], s_base;
    __shared__ double s_thr;
    // np.mean: pairwise summation in NumPy; a tree reduction in FP64 differs from it by rounding only, and the comparison
    // below is >= against half of it: detections within an ulp of the threshold are the caller's tie to break
    double acc = 0.0;
    for (int j = tid; j < n; j += DT) acc += scores[(size_t)b * F2 + j];
    acc = warp_sum(acc);
    if (lane == 0) s_part[warp] = acc;
    if (tid == 0) s_base = 0;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < DT / 32; ++w) t += s_part[w];
        s_thr = n > 0 ? 0.5 * (t / (double)n) : 0.0;
    }
    __syncthreads();
    const double thr = s_thr;
    for (int j0 = 0; j0 < n; j0 += DT) {
        const int j = j0 + tid;
        const bool keep = j < n && scores[(size_t)b * F2 + j] >= thr;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_cnt[w];
        const int pos = off + __popc(bal & ((1u << lane) - 1u));
        // (the descriptor rows are moved by the whole warp below: positions through shared memory would cost more than this)
        if (keep) {
            scores_out[(size_t)b * F2 + pos] = scores[(size_t)b * F2 + j];
            kp_out[((size_t)b * F2 + pos) * 2] = kp[((size_t)b * F2 + j) * 2];
            kp_out[((size_t)b * F2 + pos) * 2 + 1] = kp[((size_t)b * F2 + j) * 2 + 1];
            const float* s = desc + ((size_t)b * F2 + j) * D;
            float* d = desc_out + ((size_t)b * F2 + pos) * D;
            for (int e = 0; e < D; e += 4) *reinterpret_cast<float4*>(d + e) = *reinterpret_cast<const float4*>(s + e);
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < DT / 32; ++w) tot += s_cnt[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (tid == 0) n_out[b] = s_base;
}

// ---- append: the accepted match / the new track stores the detection's descriptor as its newest / first observation ----
// Runs BEFORE msckf_append_observations_batched (feat_len still holds the lengths before the frame) with the FINAL actions.
__global__ void __launch_bounds__(DT) append_descriptors_kernel(msckf_dims dims, int D, int F2, const unsigned char* __restrict__ action,
                                                                const int* __restrict__ match12, const int* __restrict__ match21,
                                                                const int* __restrict__ n_tracks, const int* __restrict__ n_new,
                                                                const float* __restrict__ desc_new, const int* __restrict__ feat_len,
                                                                float* __restrict__ obs_desc, const double* __restrict__ new_conf,
                                                                double* __restrict__ desc_sum, const int R,
                                                                const int* __restrict__ n_cams, const int* __restrict__ cam_row) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = dims.F, M = dims.M;
    const int n1 = n_tracks ? min(max(n_tracks[b], 0), F) : F;
    const int n2 = n_new ? min(max(n_new[b], 0), F2) : F2;
    const int q4 = D / 4, DS = D + 2;
    // the frame's observations belong to the newest camera (slot n_cams - 1, msckf_append_observations_batched): its row
    const int nc = min(n_cams[b], dims.N);
    const int row = nc > 0 ? cam_row[(size_t)b * dims.N + nc - 1] : -1;
    const bool row_ok = row >= 0 && row < R;
    // matched tracks: q4 = D / 4 lanes per track (16-byte pieces), 32 / q4 tracks per warp at a time; the running sums
    // sum_i w_i d_i | sum_i w_i follow
    {
        const int rpp = (q4 <= 32 && (32 % q4) == 0) ? 32 / q4 : 1;
        const int piece = lane % q4, sub = lane / q4;
        const bool lane_on = rpp > 1 || lane < q4;
        for (int j = warp * rpp + sub; j < n1 && lane_on; j += (DT / 32) * rpp) {
            const size_t fj = (size_t)b * F + j;
            if (action[fj] != 2) continue;
            const int src = match12[fj], len = feat_len[fj];
            if (src < 0 || len >= M || !row_ok) continue;
            const float4 v = *reinterpret_cast<const float4*>(desc_new + ((size_t)b * F2 + src) * D + 4 * piece);
            *reinterpret_cast<float4*>(obs_desc + ((fj * R) + row) * D + 4 * piece) = v;
            if (desc_sum) {
                const double w = new_conf[fj];
                double* a = desc_sum + fj * DS + 4 * piece;
                if (len == 0) { a[0] = w * (double)v.x; a[1] = w * (double)v.y; a[2] = w * (double)v.z; a[3] = w * (double)v.w; }
                else { a[0] += w * (double)v.x; a[1] += w * (double)v.y; a[2] += w * (double)v.z; a[3] += w * (double)v.w; }
                if (piece == 0) desc_sum[fj * DS + D] = (len == 0 ? 0.0 : desc_sum[fj * DS + D]) + w;
            }
        }
    }
    // new tracks: the r-th unmatched detection (ascending) went to slot n1 + r (np.setdiff1d order, FeatureExtractor.py:73)
    __shared__ int s_cnt[DT / 32], s_base;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int j0 = 0; j0 < n2; j0 += DT) {
        const int j = j0 + tid;
        const bool un = j < n2 && match21[(size_t)b * F2 + j] < 0;
        const unsigned bal = __ballot_sync(0xffffffffu, un);
        if (lane == 0) s_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_cnt[w];
        if (un) {
            const int slot = n1 + off + __popc(bal & ((1u << lane) - 1u));
            if (slot < F && action[(size_t)b * F + slot] == 3 && row_ok) {
                const size_t fs = (size_t)b * F + slot;
                const float4* s = reinterpret_cast<const float4*>(desc_new + ((size_t)b * F2 + j) * D);
                float4* d = reinterpret_cast<float4*>(obs_desc + (fs * R + row) * D);
                const double w = desc_sum ? new_conf[fs] : 0.0;
                for (int e = 0; e < q4; ++e) {
                    const float4 v = s[e];
                    d[e] = v;
                    if (desc_sum) {
                        double* a = desc_sum + fs * DS + 4 * e;
                        a[0] = w * (double)v.x; a[1] = w * (double)v.y; a[2] = w * (double)v.z; a[3] = w * (double)v.w;
                    }
                }
                if (desc_sum) desc_sum[fs * DS + D] = w;
            }
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < DT / 32; ++w) tot += s_cnt[w];
            s_base += tot;
        }
        __syncthreads();
    }
}

// ---- deletion: the descriptors observed from the cameras flagged in remove_mask leave their tracks (MSCKF.py:764-765) ----
// Runs BEFORE msckf_drop_observations_batched (same mask; obs_slot / feat_len still describe the window before the deletion).
__global__ void __launch_bounds__(DT) drop_descriptors_kernel(msckf_dims dims, int D, const unsigned char* __restrict__ remove_mask,
                                                              const int* __restrict__ feat_len, const int* __restrict__ obs_slot,
                                                              const float* __restrict__ obs_desc, const double* __restrict__ obs_conf,
                                                              double* __restrict__ desc_sum, const int R,
                                                              const int* __restrict__ cam_row) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int F = dims.F, M = dims.M, N = dims.N;
    {   // nothing leaves this filter's window: its sums stay as they are (the usual case)
        const unsigned char* mk = remove_mask + (size_t)b * N;
        int any = 0;
        for (int s = lane; s < N; s += 32) any |= mk[s];
        if (!__any_sync(0xffffffffu, any != 0)) return;
    }
    const int q4 = D / 4, DS = D + 2;
    for (int j = blockIdx.x * (DT / 32) + (threadIdx.x >> 5); j < F; j += gridDim.x * (DT / 32)) {
    const size_t fj = (size_t)b * F + j;
    const int len = min(feat_len[fj], M);
    if (len <= 0) continue;
    int kept = 0;
    double wgone = 0.0;
    bool any_gone = false;
    for (int i0 = 0; i0 < len; i0 += 32) {
        const int i = i0 + lane;
        const int sl = (i < len) ? obs_slot[fj * M + i] : -1;
        const bool gone = i < len && sl >= 0 && sl < N && remove_mask[(size_t)b * N + sl] != 0;
        const int rw = gone ? cam_row[(size_t)b * N + sl] : -1;
        const unsigned gm = __ballot_sync(0xffffffffu, gone);
        const unsigned lm = __ballot_sync(0xffffffffu, i < len);
        for (unsigned m = gm; m; m &= m - 1) {      // the rows that leave also leave the running sums
            const int src = __ffs(m) - 1;
            const int r = __shfl_sync(0xffffffffu, rw, src);
            const double w = obs_conf[fj * M + i0 + src];
            if (lane < q4 && r >= 0 && r < R) {
                const float4 v = *reinterpret_cast<const float4*>(obs_desc + (fj * R + r) * D + 4 * lane);
                double* a = desc_sum + fj * DS + 4 * lane;
                a[0] -= w * (double)v.x; a[1] -= w * (double)v.y; a[2] -= w * (double)v.z; a[3] -= w * (double)v.w;
            }
            wgone += w;
        }
        any_gone |= gm != 0;
        kept += __popc(lm & ~gm);
    }
    if (any_gone && lane == 0) {
        if (kept == 0) desc_sum[fj * DS + D] = 0.0;
        else desc_sum[fj * DS + D] -= wgone;
    }
    if (any_gone && kept == 0 && lane < q4) {          // an emptied track: exact zeros, not rounding residue
        double* a = desc_sum + fj * DS + 4 * lane;
        a[0] = a[1] = a[2] = a[3] = 0.0;
    }
    }
}

// ---- whole tracks follow the slot map of msckf_compact_tracks_batched ----
__global__ void __launch_bounds__(DT) compact_descriptors_kernel(msckf_dims dims, int D, const int* __restrict__ dest,
                                                                 const int* __restrict__ feat_len_dst, const float* __restrict__ src,
                                                                 float* __restrict__ dst) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int j = blockIdx.x * (DT / 32) + (threadIdx.x >> 5);
    const int F = dims.F, M = dims.M;
    if (j >= F) return;
    const int to = dest[(size_t)b * F + j];
    if (to < 0 || to >= F) return;
    const int len = min(feat_len_dst[(size_t)b * F + to], M);
    const size_t n4 = (size_t)len * D / 4;
    const float4* s = reinterpret_cast<const float4*>(src + ((size_t)b * F + j) * M * D);
    float4* d = reinterpret_cast<float4*>(dst + ((size_t)b * F + to) * M * D);
    for (size_t e = lane; e < n4; e += 32) d[e] = s[e];
}

// the same IN PLACE, one CTA per filter, staged through a scratch buffer (rows of M observations x D floats; M = 1 for the
// rows of last_camera_measurement); a filter whose map moves nothing returns after reading its map
__global__ void __launch_bounds__(DT) compact_descriptors_inplace_kernel(msckf_dims dims, int D, const int* __restrict__ dest,
                                                                         const int* __restrict__ feat_len_new, float* __restrict__ desc,
                                                                         float* __restrict__ tmp, int zero_deleted) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = dims.F, M = dims.M;
    __shared__ int s_moves, s_kept;
    if (tid == 0) { s_moves = 0; s_kept = 0; }
    __syncthreads();
    int moves = 0, kept = 0;
    for (int j = tid; j < F; j += DT) {
        const int d = dest[(size_t)b * F + j];
        if (d >= 0) { ++kept; if (d != j) moves = 1; }
    }
    if (moves) atomicOr(&s_moves, 1);
    if (kept) atomicAdd(&s_kept, kept);
    __syncthreads();
    if (s_moves) {
        for (int pass = 0; pass < 2; ++pass) {
            const float* from = pass == 0 ? desc : tmp;
            float* to = pass == 0 ? tmp : desc;
            for (int j = warp; j < F; j += DT / 32) {
                const int d = dest[(size_t)b * F + j];
                if (d < 0 || d == j || d >= F) continue;
                const int len = min(feat_len_new[(size_t)b * F + d], M);
                const size_t n4 = (size_t)len * D / 4;
                const float4* s4 = reinterpret_cast<const float4*>(from + ((size_t)b * F + (pass == 0 ? j : d)) * M * D);
                float4* d4 = reinterpret_cast<float4*>(to + ((size_t)b * F + d) * M * D);
                for (size_t e = lane; e < n4; e += 32) d4[e] = s4[e];
            }
            __threadfence_block();
            __syncthreads();
        }
    }
    if (zero_deleted) {   // rows behind the survivors hold no track any more
        const int n_kept = s_kept;
        const size_t row4 = (size_t)M * D / 4;
        for (int j = n_kept + warp; j < F; j += DT / 32) {
            float4* d4 = reinterpret_cast<float4*>(desc + ((size_t)b * F + j) * M * D);
            for (size_t e = lane; e < row4; e += 32) d4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// ---- the matcher's view: score-weighted mean descriptor of every track (MSCKF.py:439), zero rows for empty slots ----
__global__ void __launch_bounds__(DT) track_descriptors_kernel(msckf_dims dims, int D, const int* __restrict__ feat_len,
                                                               const double* __restrict__ obs_conf, const float* __restrict__ obs_desc,
                                                               const unsigned char* __restrict__ keep_old,
                                                               float* __restrict__ desc_tracks, double* __restrict__ desc_sum_out,
                                                               const int R, const int* __restrict__ obs_slot,
                                                               const int* __restrict__ cam_row) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int j = blockIdx.x * (DT / 32) + (threadIdx.x >> 5);
    const int F = dims.F, M = dims.M;
    if (j >= F) return;
    const bool skip_rows = keep_old && keep_old[b];
    if (skip_rows && !desc_sum_out) return;
    const size_t fj = (size_t)b * F + j;
    const int len = min(feat_len[fj], M);
    // a descriptor row is D / 4 float4 pieces: lane l handles piece l % q4 of the rows l / q4, l / q4 + 32 / q4, ... (D = 32:
    // four rows per 512-byte warp load); the partial sums of the lanes that share a piece are combined by shuffles
    const int q4 = D / 4;                     // 8, 16 or 32 for D = 32, 64, 128 (other D: one row per pass, lanes beyond q4 idle)
    const int rows_per_pass = q4 <= 32 && (32 % q4) == 0 ? 32 / q4 : 1;
    const int piece = lane % q4, sub = lane / q4;
    const bool lane_on = (rows_per_pass > 1) || lane < q4;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, wsum = 0.0;
    if (lane_on) {
        for (int i = sub; i < len; i += rows_per_pass) {
            const int sl = obs_slot[fj * M + i];
            const int r = (sl >= 0 && sl < dims.N) ? cam_row[(size_t)b * dims.N + sl] : -1;
            if (r < 0 || r >= R) continue;
            const double w = obs_conf[fj * M + i];
            const float4 v = *reinterpret_cast<const float4*>(obs_desc + (fj * R + r) * D + 4 * piece);
            wsum += w;
            a0 += w * (double)v.x; a1 += w * (double)v.y; a2 += w * (double)v.z; a3 += w * (double)v.w;
        }
    }
    for (int o = q4; o < 32; o <<= 1) {       // lanes with the same piece: l, l + q4, l + 2 q4, ...
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        a3 += __shfl_xor_sync(0xffffffffu, a3, o);
        wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    }
    if (lane < q4) {
        if (desc_sum_out) {                   // (re)build the running sums from the lists
            double* a = desc_sum_out + fj * (D + 2) + 4 * lane;
            a[0] = a0; a[1] = a1; a[2] = a2; a[3] = a3;
            if (lane == 0) desc_sum_out[fj * (D + 2) + D] = wsum;
        }
        if (!skip_rows && desc_tracks) {
            const double inv = (len > 0 && wsum != 0.0) ? 1.0 / wsum : 0.0;
            *reinterpret_cast<float4*>(desc_tracks + fj * D + 4 * lane) =
                make_float4((float)(a0 * inv), (float)(a1 * inv), (float)(a2 * inv), (float)(a3 * inv));
        }
    }
}

// the same rows from the running sums (sum_i w_i d_i | sum_i w_i, FP64) the append / drop kernels maintain: D + 2 doubles per
// track instead of its whole list (128 bytes per observation at D = 32)
__global__ void __launch_bounds__(DT) track_descriptors_sums_kernel(msckf_dims dims, int D, const int* __restrict__ feat_len,
                                                                    const double* __restrict__ desc_sum,
                                                                    const unsigned char* __restrict__ keep_old,
                                                                    float* __restrict__ desc_tracks) {
    const int b = blockIdx.y;
    if (keep_old && keep_old[b]) return;
    const int q4 = D / 4;
    const int F = dims.F;
    for (int e = blockIdx.x * DT + threadIdx.x; e < F * q4; e += gridDim.x * DT) {
        const int j = e / q4, p = e - j * q4;
        const size_t fj = (size_t)b * F + j;
        const double* a = desc_sum + fj * (D + 2);
        const double wsum = a[D];
        const double inv = (feat_len[fj] > 0 && wsum != 0.0) ? 1.0 / wsum : 0.0;
        *reinterpret_cast<float4*>(desc_tracks + fj * D + 4 * p) =
            make_float4((float)(a[4 * p] * inv), (float)(a[4 * p + 1] * inv), (float)(a[4 * p + 2] * inv), (float)(a[4 * p + 3] * inv));
    }
}

// ---- the camera ring: cam_row[B][N] = descriptor row (0 .. R-1) of the camera in window slot s.  A camera keeps its row for
//      its whole life, so the deletion of cameras moves no descriptor; rows of deleted cameras are handed out again. ----
__global__ void ring_assign_kernel(msckf_dims dims, int R, const int* __restrict__ n_cams, const unsigned char* __restrict__ active,
                                   int* __restrict__ cam_row) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= dims.B || (active && !active[b])) return;
    const int n = n_cams[b];
    if (n < 0 || n >= dims.N) return;                 // (a full window: msckf_augment_batched leaves it untouched as well)
    unsigned long long used = 0ull;
    for (int s = 0; s < n; ++s) {
        const int r = cam_row[(size_t)b * dims.N + s];
        if (r >= 0 && r < 64) used |= 1ull << r;
    }
    const int fr = __ffsll((long long)~used) - 1;     // lowest free row
    cam_row[(size_t)b * dims.N + n] = (fr >= 0 && fr < R) ? fr : -1;
}

__global__ void ring_remove_kernel(msckf_dims dims, const int* __restrict__ n_cams, const unsigned char* __restrict__ remove_mask,
                                   int* __restrict__ cam_row) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= dims.B) return;
    const int n = min(max(n_cams[b], 0), dims.N);
    int k = 0;
    for (int s = 0; s < n; ++s)
        if (!remove_mask[(size_t)b * dims.N + s]) cam_row[(size_t)b * dims.N + k++] = cam_row[(size_t)b * dims.N + s];
}

}  // namespace

static int check_desc(const msckf_dims* d, int D, int Dmax = 128) {
    if (!d || d->B <= 0 || d->F <= 0 || d->M <= 0 || d->N <= 0) return MSCKF_EINVAL;
    if (D <= 0 || D > Dmax || (D & 3)) return MSCKF_ECAPACITY;
    if (d->B > 65535) return MSCKF_ECAPACITY;
    return 0;
}

extern "C" {

int msckf_score_filter_batched(int B, int F2, int D, const int* n_in, const double* scores, const double* kp, const float* desc,
                               double* scores_out, double* kp_out, float* desc_out, int* n_out, void* stream) {
    if (B <= 0 || F2 <= 0 || !scores || !kp || !desc || !scores_out || !kp_out || !desc_out || !n_out) return MSCKF_EINVAL;
    if (D <= 0 || (D & 3)) return MSCKF_ECAPACITY;
    score_filter_kernel<<<B, DT, 0, (cudaStream_t)stream>>>(F2, D, n_in, scores, kp, desc, scores_out, kp_out, desc_out, n_out);
    msckf_count_launch(1);
    return (int)cudaGetLastError();
}

int msckf_append_descriptors_batched(const msckf_dims* dims, int D, int F_new, const unsigned char* action, const int* match12,
                                     const int* match21, const int* n_tracks, const int* n_new, const float* desc_new,
                                     const int* feat_len, float* obs_desc, const double* new_conf, double* desc_sum, int R,
                                     const int* n_cams, const int* cam_row, void* stream) {
    int rc = check_desc(dims, D);
    if (rc) return rc;
    if (F_new <= 0 || !action || !match12 || !match21 || !desc_new || !feat_len || !obs_desc || !n_cams || !cam_row) return MSCKF_EINVAL;
    if (desc_sum && !new_conf) return MSCKF_EINVAL;
    if (R <= 0 || R > 64) return MSCKF_ECAPACITY;
    append_descriptors_kernel<<<dims->B, DT, 0, (cudaStream_t)stream>>>(*dims, D, F_new, action, match12, match21, n_tracks, n_new,
                                                                       desc_new, feat_len, obs_desc, new_conf, desc_sum, R, n_cams,
                                                                       cam_row);
    msckf_count_launch(1);
    return (int)cudaGetLastError();
}

int msckf_drop_descriptors_batched(const msckf_dims* dims, int D, const unsigned char* remove_mask, const int* feat_len,
                                   const int* obs_slot, const float* obs_desc, const double* obs_conf, double* desc_sum, int R,
                                   const int* cam_row, void* stream) {
    int rc = check_desc(dims, D);
    if (rc) return rc;
    if (!remove_mask || !feat_len || !obs_slot || !obs_desc || !obs_conf || !desc_sum || !cam_row) return MSCKF_EINVAL;
    if (R <= 0 || R > 64) return MSCKF_ECAPACITY;
    dim3 grid(min(4, (dims->F + DT / 32 - 1) / (DT / 32)), dims->B);
    drop_descriptors_kernel<<<grid, DT, 0, (cudaStream_t)stream>>>(*dims, D, remove_mask, feat_len, obs_slot, obs_desc, obs_conf,
                                                                   desc_sum, R, cam_row);
    msckf_count_launch(1);
    return (int)cudaGetLastError();
}

int msckf_compact_descriptors_batched(const msckf_dims* dims, int D, const int* dest, const int* feat_len_dst, const float* src,
                                      float* dst, void* stream) {
    int rc = check_desc(dims, D);
    if (rc) return rc;
    if (!dest || !feat_len_dst || !src || !dst || src == dst) return MSCKF_EINVAL;
    dim3 grid((dims->F + DT / 32 - 1) / (DT / 32), dims->B);
    compact_descriptors_kernel<<<grid, DT, 0, (cudaStream_t)stream>>>(*dims, D, dest, feat_len_dst, src, dst);
    msckf_count_launch(1);
    return (int)cudaGetLastError();
}

int msckf_compact_descriptors_inplace_batched(const msckf_dims* dims, int D, const int* dest, const int* feat_len_new, float* desc,
                                              float* scratch, int zero_deleted, void* stream) {
    int rc = check_desc(dims, D, 1024);      // (also moves the rows of the running sums: 2 (D + 2) floats each)
    if (rc) return rc;
    if (!dest || !feat_len_new || !desc || !scratch || desc == scratch) return MSCKF_EINVAL;
    compact_descriptors_inplace_kernel<<<dims->B, DT, 0, (cudaStream_t)stream>>>(*dims, D, dest, feat_len_new, desc, scratch, zero_deleted);
    msckf_count_launch(1);
    return (int)cudaGetLastError();
}

int msckf_track_descriptors_batched(const msckf_dims* dims, int D, const int* feat_len, const double* obs_conf,
                                    const float* obs_desc, const unsigned char* keep_old, float* desc_tracks, double* desc_sum_out,
                                    int R, const int* obs_slot, const int* cam_row, void* stream) {
    int rc = check_desc(dims, D);
    if (rc) return rc;
    if (!feat_len || !obs_conf || !obs_desc || (!desc_tracks && !desc_sum_out) || !obs_slot || !cam_row) return MSCKF_EINVAL;
    if (R <= 0 || R > 64) return MSCKF_ECAPACITY;
    dim3 grid((dims->F + DT / 32 - 1) / (DT / 32), dims->B);
    track_descriptors_kernel<<<grid, DT, 0, (cudaStream_t)stream>>>(*dims, D, feat_len, obs_conf, obs_desc, keep_old, desc_tracks,
                                                                    desc_sum_out, R, obs_slot, cam_row);
    msckf_count_launch(1);
    return (int)cudaGetLastError();
}

int msckf_track_descriptors_from_sums_batched(const msckf_dims* dims, int D, const int* feat_len, const double* desc_sum,
                                              const unsigned char* keep_old, float* desc_tracks, void* stream) {
    int rc = check_desc(dims, D);
    if (rc) return rc;
    if (!feat_len || !desc_sum || !desc_tracks) return MSCKF_EINVAL;
    const int work = dims->F * (D / 4);
    dim3 grid(min(64, (work + DT - 1) / DT), dims->B);
    track_descriptors_sums_kernel<<<grid, DT, 0, (cudaStream_t)stream>>>(*dims, D, feat_len, desc_sum, keep_old, desc_tracks);
    msckf_count_launch(1);
    return (int)cudaGetLastError();
}

int msckf_descriptor_rows_assign_batched(const msckf_dims* dims, int R, const int* n_cams, const unsigned char* active, int* cam_row,
                                         void* stream) {
    if (!dims || dims->B <= 0 || dims->N <= 0 || !n_cams || !cam_row) return MSCKF_EINVAL;
    if (R <= 0 || R > 64 || R < dims->N) return MSCKF_ECAPACITY;
    ring_assign_kernel<<<(dims->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*dims, R, n_cams, active, cam_row);
    msckf_count_launch(1);
    return (int)cudaGetLastError();
}

int msckf_descriptor_rows_remove_batched(const msckf_dims* dims, const int* n_cams, const unsigned char* remove_mask, int* cam_row,
                                         void* stream) {
    if (!dims || dims->B <= 0 || dims->N <= 0 || !n_cams || !remove_mask || !cam_row) return MSCKF_EINVAL;
    ring_remove_kernel<<<(dims->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*dims, n_cams, remove_mask, cam_row);
    msckf_count_launch(1);
    return (int)cudaGetLastError();
}

}  // extern "C"
