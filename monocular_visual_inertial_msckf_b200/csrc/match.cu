// "Next rows" of SURVEY.md section 8f, row 3: what add_camera_measurements does between the feature extractor and the
// track bookkeeping, per filter and per camera frame, with the tracks resident in HBM:
//   match_descriptors_kernel  <-> XFeat.match (reference src/accelerated_features/modules/xfeat.py:327-348) as called from
//                                 FeatureExtractor.match (src/msckf/FeatureExtractor.py:62-84): mutual nearest neighbours by
//                                 cosine similarity above min_cossim; plus the proposal list the reference builds from it
//                                 (MSCKF.py:317-318, :400-434: matched -> candidate append, unmatched track -> lost, unmatched
//                                 detection -> new track, in ascending detection order like np.setdiff1d)
//   verify_matches_kernel     <-> the epipolar / homography walk over a track's past observations (MSCKF.py:333-401)
// FP32 dot products (XFeat's dtype) for the matcher, FP64 for the geometry.  Neither is GEMM-shaped enough to leave the
// CUDA cores: 64-dimensional dot products with an argmax epilogue, and a handful of 3x3 products per observation.
#include "msckf_common.cuh"

namespace {

constexpr int MT = 256;
constexpr int MM = 512;      // threads of a matcher CTA (thread grid 16 x 32, 128 registers each)

struct Intrinsics { double k[9], ki[9]; };

// ---- geometric verification -----------------------------------------------------------------------------------------
// CTA = (filter, 128 consecutive tracks).  Four warps stage the proposals of those tracks, two others put the relative
// geometry of every window camera against the newest one into shared memory (thread = camera): R = R_s^T R_new, the branch
// flag |t| < 0.01 and the fundamental matrix in pixel coordinates F = K^-T [t]x R K^-1 (MSCKF.py:360-362, :387).  Then the
// CTA walks the 128 x M observation slots of its tracks as ONE flat, contiguous range (coalesced 16-byte uv and 4-byte slot
// loads, all independent): every valid observation is tested -- 11 multiply-adds on the epipolar branch -- and a failure
// enters a shared-memory atomicMin over (observation index, test); the minimum is the FIRST failing observation, which is
// what the reference's sequential walk with `break` reports (MSCKF.py:381, :398).
// HBM: 20 B per past observation (uv + slot) + 18 B per proposal.
constexpr int V_TRACKS = 128;

__global__ void __launch_bounds__(MT, 4) verify_matches_kernel(msckf_dims dims, Intrinsics Km, double epipolar_thr,
                                                            double homography_thr, const int* __restrict__ n_cams,
                                                            const double* __restrict__ cam_R, const double* __restrict__ cam_t,
                                                            const int* __restrict__ feat_len, const double* __restrict__ obs_uv,
                                                            const int* __restrict__ obs_slot, const double* __restrict__ new_uv,
                                                            unsigned char* action, unsigned char* __restrict__ verdict,
                                                            int* n_homography, int* n_epipolar) {
    __shared__ double s_F[MSCKF_MAX_CAMS][9];
    __shared__ double s_R[MSCKF_MAX_CAMS][9];
    __shared__ unsigned char s_h[MSCKF_MAX_CAMS];
    __shared__ double s_uv[V_TRACKS][2];
    __shared__ int s_len[V_TRACKS], s_first[V_TRACKS];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int F = dims.F, M = dims.M, N = dims.N;
    const int nc = min(n_cams[b], N), s2 = nc - 1;
    const int j0 = blockIdx.x * V_TRACKS, nt = min(V_TRACKS, F - j0);
    const double* Ki = Km.ki;
    const double* K = Km.k;
    if (tid < nt) {
        const size_t fj = (size_t)b * F + j0 + tid;
        const bool proposed = action[fj] == 2;
        s_len[tid] = proposed ? (s2 >= 0 ? min(feat_len[fj], M) : 0) : -1;          // -1: not proposed
        s_uv[tid][0] = new_uv[fj * 2];
        s_uv[tid][1] = new_uv[fj * 2 + 1];
        s_first[tid] = 0x7fffffff;
    } else if (tid >= V_TRACKS && tid - V_TRACKS < nc) {
        const int s1 = tid - V_TRACKS;
        const double* R1 = cam_R + ((size_t)b * N + s1) * 9;
        const double* t1 = cam_t + ((size_t)b * N + s1) * 3;
        const double* R2 = cam_R + ((size_t)b * N + s2) * 9;
        const double* t2 = cam_t + ((size_t)b * N + s2) * 3;
        // T_C1_C2 = T_W_C1^-1 T_W_C2 (MSCKF.py:360): R = R1^T R2, t = R1^T (t2 - t1)
        double R[9], t[3], E[9], G[9];
        const double d0 = t2[0] - t1[0], d1 = t2[1] - t1[1], d2 = t2[2] - t1[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            t[r] = R1[r] * d0 + R1[3 + r] * d1 + R1[6 + r] * d2;
#pragma unroll
            for (int c = 0; c < 3; ++c) R[r * 3 + c] = R1[r] * R2[c] + R1[3 + r] * R2[3 + c] + R1[6 + r] * R2[6 + c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {            // E = [t]x R
            E[c] = t[1] * R[6 + c] - t[2] * R[3 + c];
            E[3 + c] = t[2] * R[c] - t[0] * R[6 + c];
            E[6 + c] = t[0] * R[3 + c] - t[1] * R[c];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) G[r * 3 + c] = E[r * 3] * Ki[c] + E[r * 3 + 1] * Ki[3 + c] + E[r * 3 + 2] * Ki[6 + c];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) s_F[s1][r * 3 + c] = Ki[r] * G[c] + Ki[3 + r] * G[3 + c] + Ki[6 + r] * G[6 + c];
        const bool near = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]) < 0.01;
        s_h[s1] = near ? 1 : 0;
        if (near) {
            // homography branch (MSCKF.py:362-363): H = K R K^-1 into s_R, H^-1 = K R^T K^-1 into s_F (F is not needed)
            double A[9], Bt[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    A[r * 3 + c] = R[r * 3] * Ki[c] + R[r * 3 + 1] * Ki[3 + c] + R[r * 3 + 2] * Ki[6 + c];      // R K^-1
                    Bt[r * 3 + c] = R[r] * Ki[c] + R[3 + r] * Ki[3 + c] + R[6 + r] * Ki[6 + c];                 // R^T K^-1
                }
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    s_R[s1][r * 3 + c] = K[r * 3] * A[c] + K[r * 3 + 1] * A[3 + c] + K[r * 3 + 2] * A[6 + c];
                    s_F[s1][r * 3 + c] = K[r * 3] * Bt[c] + K[r * 3 + 1] * Bt[3 + c] + K[r * 3 + 2] * Bt[6 + c];
                }
        }
    }
    __syncthreads();
    const size_t base = ((size_t)b * F + j0) * M;
    const double2* uv2 = reinterpret_cast<const double2*>(obs_uv) + base;
    // four slots per thread and pass: all eight loads are issued before the first test (the walk is bound by load latency)
    constexpr int VU = 4;
    for (int e0 = tid; e0 < nt * M; e0 += VU * MT) {
        int jl_[VU], i_[VU], s_[VU];
        double2 f_[VU];
#pragma unroll
        for (int u = 0; u < VU; ++u) {
            const int e = e0 + u * MT;
            s_[u] = -1;
            jl_[u] = 0;
            i_[u] = 0;
            f_[u] = make_double2(0.0, 0.0);
            if (e < nt * M) {
                jl_[u] = e / M;
                i_[u] = e - jl_[u] * M;
                if (i_[u] < s_len[jl_[u]]) {
                    s_[u] = obs_slot[base + e];
                    f_[u] = uv2[e];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < VU; ++u) {
            const int s1 = s_[u], jl = jl_[u], i = i_[u];
            const double2 f = f_[u];
            if (s1 < 0 || s1 > s2) continue;
            const double fu = f.x, fv = f.y, mu = s_uv[jl][0], mv = s_uv[jl][1];
            int fail = 0;
            if (s_h[s1]) {
                // homography H = K R K^-1 (MSCKF.py:363), H^-1 = K R^T K^-1.  Literal reference form (:364-368): the new
                // detection is compared with ITS OWN image under H^-1, the old observation with its own image under H (not
                // with each other); mean of the two distances
                const double* H = s_R[s1];
                const double* Hi = s_F[s1];
                const double p0 = H[0] * fu + H[1] * fv + H[2], p1 = H[3] * fu + H[4] * fv + H[5], p2 = H[6] * fu + H[7] * fv + H[8];
                const double q0 = Hi[0] * mu + Hi[1] * mv + Hi[2], q1 = Hi[3] * mu + Hi[4] * mv + Hi[5],
                             q2 = Hi[6] * mu + Hi[7] * mv + Hi[8];
                const double e1x = mu - q0 / q2, e1y = mv - q1 / q2;       // matched_keypoint - x1_pred
                const double e2x = fu - p0 / p2, e2y = fv - p1 / p2;       // feature_keypoint - x2_pred
                const double score = 0.5 * (sqrt(e1x * e1x + e1y * e1y) + sqrt(e2x * e2x + e2y * e2y));
                if (score > homography_thr) fail = 2;
            } else {
                // SIGNED score [x_new, 1]^T F [x_old, 1] > threshold rejects (MSCKF.py:388-390)
                const double* Fm = s_F[s1];
                const double v0 = Fm[0] * fu + Fm[1] * fv + Fm[2], v1 = Fm[3] * fu + Fm[4] * fv + Fm[5],
                             v2 = Fm[6] * fu + Fm[7] * fv + Fm[8];
                const double score = mu * v0 + mv * v1 + v2;
                if (score > epipolar_thr) fail = 3;
            }
            if (fail) atomicMin(&s_first[jl], i * 4 + fail);
        }
    }
    __syncthreads();
    if (tid < nt) {
        const size_t fj = (size_t)b * F + j0 + tid;
        int code = 0;
        if (s_len[tid] >= 0) {
            const int first = s_first[tid];
            code = first == 0x7fffffff ? 1 : (first & 3);
            if (code != 1) {
                action[fj] = 1;               // rejected proposal: the track is only marked lost this frame (MSCKF.py:402-404)
                if (code == 2 && n_homography) atomicAdd(n_homography + b, 1);
                if (code == 3 && n_epipolar) atomicAdd(n_epipolar + b, 1);
            }
        }
        verdict[fj] = (unsigned char)code;
    }
}

// ---- mutual-nearest-neighbour matcher ---------------------------------------------------------------------------------
// One CTA per filter, both descriptor sets staged in shared memory (row stride D + 4 floats).  The similarity matrix
// S = D1 D2^T is computed ONCE in 8 x 4 register tiles (thread grid 16 x 32, rows ty + 32 i, columns tx + 16 j: every
// 16-byte operand read is shared by 16 threads of the warp, 8 multiply-adds per shared-memory wavefront) and never stored:
// each tile goes straight into the two arg-max reductions -- rows (D1 D2^T, xfeat.py:330/333) through registers and
// half-warp shuffles, columns (the reference's second product D2 D1^T, xfeat.py:331/334, which is the same matrix
// transposed) through 64-bit shared-memory atomicMax.  Every S[i][j] is one fixed-order FP32 fmaf chain over k, so the two
// reductions see identical values.  Keys = (order-preserving bits of the similarity << 32) | ~index: the maximum key is
// the maximum similarity and, among equals, the lowest index (first arg-max).
__device__ __forceinline__ unsigned long long sim_key(float v, int idx) {
    const unsigned bits = __float_as_uint(v + 0.0f);
    const unsigned ord = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
    return ((unsigned long long)ord << 32) | (unsigned)(~idx);
}
__device__ __forceinline__ float key_sim(unsigned long long key) {
    const unsigned ord = (unsigned)(key >> 32);
    return __uint_as_float((ord & 0x80000000u) ? (ord ^ 0x80000000u) : ~ord);
}
__device__ __forceinline__ int key_idx(unsigned long long key) { return (int)~(unsigned)(key & 0xffffffffu); }
__device__ __forceinline__ unsigned long long kmax(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

struct MatchArgs {
    int B, F1, F2;                 // filters, track capacity, detection capacity
    float min_cossim;
    const int* n_tracks;           // [B] or null (= F1)
    const int* n_new;              // [B] or null (= F2)
    const float* desc_tracks;      // [B][F1][D]
    const float* desc_new;         // [B][F2][D]
    int* match12;                  // [B][F1]  detection index or -1
    int* match21;                  // [B][F2]  track index or -1
    // optional proposal list for verify_matches / append_observations (all null to skip)
    const double* kp_new;          // [B][F2][2]
    const double* score_new;       // [B][F2]
    unsigned char* action;         // [B][F1]
    double* new_uv;                // [B][F1][2]
    double* new_conf;              // [B][F1]
    int* n_tracks_out;             // [B] tracks after the new ones were given slots (may be null)
    int* overflow;                 // detections that found no free track slot (incremented)
};

template <int D>
__global__ void __launch_bounds__(MM) match_descriptors_kernel(MatchArgs a) {
    constexpr int LD = D + 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n1 = a.n_tracks ? min(max(a.n_tracks[b], 0), a.F1) : a.F1;
    const int n2 = a.n_new ? min(max(a.n_new[b], 0), a.F2) : a.F2;
    unsigned long long* s_rowkey = reinterpret_cast<unsigned long long*>(smem_raw);
    unsigned long long* s_colkey = s_rowkey + a.F1;
    float* s_d1 = reinterpret_cast<float*>(smem_raw + (((size_t)(a.F1 + a.F2) * sizeof(unsigned long long) + 15) & ~(size_t)15));
    float* s_d2 = s_d1 + (size_t)a.F1 * LD;
    int* s_m12 = reinterpret_cast<int*>(s_d2 + (size_t)a.F2 * LD);
    int* s_m21 = s_m12 + a.F1;
    int* s_scan = s_m21 + a.F2;                 // [MM / 32 + 1]
    // stage both sets (coalesced 16-byte loads)
    const float4* g1 = reinterpret_cast<const float4*>(a.desc_tracks + (size_t)b * a.F1 * D);
    const float4* g2 = reinterpret_cast<const float4*>(a.desc_new + (size_t)b * a.F2 * D);
    for (int e = tid; e < n1 * (D / 4); e += MM) {
        const int r = e / (D / 4), k = e - r * (D / 4);
        *reinterpret_cast<float4*>(s_d1 + (size_t)r * LD + 4 * k) = g1[e];
    }
    for (int e = tid; e < n2 * (D / 4); e += MM) {
        const int r = e / (D / 4), k = e - r * (D / 4);
        *reinterpret_cast<float4*>(s_d2 + (size_t)r * LD + 4 * k) = g2[e];
    }
    for (int j = tid; j < a.F2; j += MM) s_colkey[j] = 0ull;
    for (int i = tid; i < a.F1; i += MM) s_rowkey[i] = 0ull;
    __syncthreads();
    if (n1 > 0 && n2 > 0) {
        constexpr int RB = 8 * (MM / 16);       // rows per pass of the thread grid: 8 per thread x (MM / 16) row groups
        for (int rb = 0; rb < n1; rb += RB) {
            float rbest[8];
            int ridx[8];
            const float* pa[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                rbest[i] = -INFINITY;
                ridx[i] = 0;
                // rows / columns beyond n1 / n2 are clamped to the last valid one: duplicates with a HIGHER index, which lose
                // every tie against the original and are never written out
                pa[i] = s_d1 + (size_t)min(rb + i * (MM / 16) + ty, n1 - 1) * LD;
            }
            for (int cb = 0; cb < n2; cb += 64) {
                const float* pb[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) pb[j] = s_d2 + (size_t)min(cb + j * 16 + tx, n2 - 1) * LD;
                float acc[8][4];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
                for (int k = 0; k < D; k += 4) {
                    float4 av[8], bv[4];
#pragma unroll
                    for (int i = 0; i < 8; ++i) av[i] = *reinterpret_cast<const float4*>(pa[i] + k);
#pragma unroll
                    for (int j = 0; j < 4; ++j) bv[j] = *reinterpret_cast<const float4*>(pb[j] + k);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            acc[i][j] = fmaf(av[i].x, bv[j].x, acc[i][j]);
                            acc[i][j] = fmaf(av[i].y, bv[j].y, acc[i][j]);
                            acc[i][j] = fmaf(av[i].z, bv[j].z, acc[i][j]);
                            acc[i][j] = fmaf(av[i].w, bv[j].w, acc[i][j]);
                        }
                }
                // a thread meets its columns (j, cb) and rows (i) in ascending index order: strict > keeps the first maximum
                float cbest[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                int cidx[4] = {0, 0, 0, 0};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rb + i * (MM / 16) + ty;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float v = acc[i][j];
                        if (v > rbest[i]) { rbest[i] = v; ridx[i] = cb + j * 16 + tx; }
                        if (v > cbest[j]) { cbest[j] = v; cidx[j] = r; }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned long long mine = sim_key(cbest[j], cidx[j]);
                    const unsigned long long m = kmax(mine, __shfl_xor_sync(0xffffffffu, mine, 16));
                    const int c = cb + j * 16 + tx;
                    if ((tid & 16) == 0 && c < n2) atomicMax(&s_colkey[c], m);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                unsigned long long m = sim_key(rbest[i], ridx[i]);
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) m = kmax(m, __shfl_xor_sync(0xffffffffu, m, o));
                const int r = rb + i * (MM / 16) + ty;
                if (tx == 0 && r < n1) s_rowkey[r] = m;
            }
        }
    }
    __syncthreads();
    // mutual & good (xfeat.py:336-343)
    for (int i = tid; i < a.F1; i += MM) {
        int j = -1;
        if (i < n1 && n2 > 0) {
            const unsigned long long rk = s_rowkey[i];
            const int c = key_idx(rk);
            if (key_idx(s_colkey[c]) == i && (!(a.min_cossim > 0.f) || key_sim(rk) > a.min_cossim)) j = c;
        }
        a.match12[(size_t)b * a.F1 + i] = j;
        s_m12[i] = j;
    }
    for (int j = tid; j < a.F2; j += MM) s_m21[j] = -1;
    __syncthreads();
    for (int i = tid; i < n1; i += MM)
        if (s_m12[i] >= 0) s_m21[s_m12[i]] = i;
    __syncthreads();
    for (int j = tid; j < a.F2; j += MM) a.match21[(size_t)b * a.F2 + j] = s_m21[j];
    if (!a.action) return;
    // proposals: matched track -> 2 with the detection's keypoint / score, unmatched track -> 1 (lost this frame,
    // MSCKF.py:437), free slots from n1 on -> 3 for the unmatched detections in ascending order (np.setdiff1d, :414-434)
    for (int i = tid; i < a.F1; i += MM) {
        const size_t o = (size_t)b * a.F1 + i;
        unsigned char act = 0;
        double u = 0.0, v = 0.0, cf = 0.0;
        if (i < n1) {
            const int j = s_m12[i];
            act = j >= 0 ? 2 : 1;
            if (j >= 0) {
                u = a.kp_new[((size_t)b * a.F2 + j) * 2]; v = a.kp_new[((size_t)b * a.F2 + j) * 2 + 1];
                cf = a.score_new[(size_t)b * a.F2 + j];
            }
        }
        a.action[o] = act; a.new_uv[o * 2] = u; a.new_uv[o * 2 + 1] = v; a.new_conf[o] = cf;
    }
    __syncthreads();                             // the new-track writes below land on slots >= n1 of the same arrays
    int running = 0;                             // unmatched detections before the current chunk (same in every thread)
    for (int base = 0; base < n2; base += MM) {
        const int j = base + tid;
        const bool un = j < n2 && s_m21[j] < 0;
        const unsigned ball = __ballot_sync(0xffffffffu, un);
        const int lane = tid & 31, w = tid >> 5;
        if (lane == 0) s_scan[w] = __popc(ball);
        __syncthreads();
        int before = 0, total = 0;
        for (int k = 0; k < MM / 32; ++k) {
            const int cnt = s_scan[k];
            if (k < w) before += cnt;
            total += cnt;
        }
        if (un) {
            const int slot = n1 + running + before + __popc(ball & ((1u << lane) - 1u));
            if (slot < a.F1) {
                const size_t o = (size_t)b * a.F1 + slot;
                a.action[o] = 3;
                a.new_uv[o * 2] = a.kp_new[((size_t)b * a.F2 + j) * 2];
                a.new_uv[o * 2 + 1] = a.kp_new[((size_t)b * a.F2 + j) * 2 + 1];
                a.new_conf[o] = a.score_new[(size_t)b * a.F2 + j];
            } else if (a.overflow) {
                atomicAdd(a.overflow, 1);
            }
        }
        running += total;
        __syncthreads();
    }
    if (tid == 0 && a.n_tracks_out) a.n_tracks_out[b] = min(n1 + running, a.F1);
}

size_t match_smem_bytes(int F1, int F2, int D) {
    return (size_t)(F1 + F2) * (sizeof(unsigned long long) + (D + 4) * sizeof(float) + sizeof(int)) + (MM / 32 + 1) * sizeof(int) + 16;
}

template <int D>
cudaError_t launch_match_t(const MatchArgs& a, cudaStream_t s) {
    const size_t smem = match_smem_bytes(a.F1, a.F2, D);
    cudaError_t e = cudaFuncSetAttribute(match_descriptors_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    match_descriptors_kernel<D><<<a.B, MM, smem, s>>>(a);
    msckf_count_launch(1);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_verify_matches(const msckf_dims& d, const double* K_host, const double* Kinv_host, double epipolar_thr,
                                  double homography_thr, const int* n_cams, const double* cam_R, const double* cam_t,
                                  const int* feat_len, const double* obs_uv, const int* obs_slot, const double* new_uv,
                                  unsigned char* action, unsigned char* verdict, int* n_homography, int* n_epipolar,
                                  cudaStream_t s) {
    if (d.F == 0) return cudaSuccess;
    Intrinsics Km;
    for (int k = 0; k < 9; ++k) { Km.k[k] = K_host[k]; Km.ki[k] = Kinv_host[k]; }
    dim3 grid((d.F + V_TRACKS - 1) / V_TRACKS, d.B);
    verify_matches_kernel<<<grid, MT, 0, s>>>(d, Km, epipolar_thr, homography_thr, n_cams, cam_R, cam_t, feat_len, obs_uv,
                                              obs_slot, new_uv, action, verdict, n_homography, n_epipolar);
    msckf_count_launch(1);
    return cudaGetLastError();
}

// returns -2 (MSCKF_ECAPACITY) for an unsupported descriptor length or sets that do not fit the shared memory of one SM
int launch_match_descriptors(int B, int F1, int F2, int D, double min_cossim, const int* n_tracks, const int* n_new,
                             const float* desc_tracks, const float* desc_new, int* match12, int* match21,
                             const double* kp_new, const double* score_new, unsigned char* action, double* new_uv,
                             double* new_conf, int* n_tracks_out, int* overflow, cudaStream_t s) {
    MatchArgs a{B, F1, F2, (float)min_cossim, n_tracks, n_new, desc_tracks, desc_new, match12, match21,
                kp_new, score_new, action, new_uv, new_conf, n_tracks_out, overflow};
    if (match_smem_bytes(F1, F2, D) > 227 * 1024) return MSCKF_ECAPACITY;
    switch (D) {
        case 32: return (int)launch_match_t<32>(a, s);
        case 64: return (int)launch_match_t<64>(a, s);
        case 128: return (int)launch_match_t<128>(a, s);
        default: return MSCKF_ECAPACITY;
    }
}
