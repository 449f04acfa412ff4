// "Next rows" of SURVEY.md section 8f, rows 1-2: the window selection and the track surgery that sit between two updates,
// on the struct-of-arrays tracks in HBM (integer / byte work, one pass over the tracks each):
//   select_poorest_kernel     <-> MSCKF.prune_poorest_camera_states, selection part (reference src/msckf/MSCKF.py:712-733)
//   drop_observations_kernel  <-> MSCKF.remove_cameras, track part (MSCKF.py:760-770) + the slot renumbering implied by
//                                 `del self.state.cameras[...]` (MSCKF.py:758, slots are dict positions, MSCKF.py:539)
#include "msckf_common.cuh"

namespace {

constexpr int TT = 256;
#ifndef TRACKS_KT
#define TRACKS_KT 2
#endif

// Bit sets of the window slots that observe KS tracks of one filter (tracks j0, j0 + stride, ...): the KS length loads and then
// the KS slot loads of a round are independent, so a warp has them all in flight (one track at a time left the kernels at the
// latency of two dependent loads per track).  Lane = observation; the sets come back reduced over the warp.
template <int KS>
__device__ __forceinline__ void slot_sets(const int* __restrict__ len, const int* __restrict__ slot, int j0, int stride, int F, int M,
                                          int nc, int lane, unsigned (&lo)[KS], unsigned (&hi)[KS]) {
    int n[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        const int j = j0 + k * stride;
        n[k] = j < F ? min(len[j], M) : 0;
        lo[k] = hi[k] = 0;
    }
    for (int i0 = 0; i0 < M; i0 += 32) {
        int sl[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) sl[k] = (i0 + lane < n[k]) ? slot[(size_t)(j0 + k * stride) * M + i0 + lane] : -1;
#pragma unroll
        for (int k = 0; k < KS; ++k)
            if (sl[k] >= 0 && sl[k] < nc) {
                if (sl[k] < 32) lo[k] |= 1u << sl[k];
                else hi[k] |= 1u << (sl[k] - 32);
            }
    }
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        lo[k] = __reduce_or_sync(0xffffffffu, lo[k]);
        hi[k] = __reduce_or_sync(0xffffffffu, hi[k]);
    }
}

// One CTA per filter.  Every track is first reduced to the bit set of the window slots that observe it (warp per track,
// lane = observation, one OR-reduction); thread s then counts the tracks whose set holds slot s and finds the slot's first
// appearance in track order (the reference builds a dict while walking the tracks, so cameras enter it in that order, and
// Python's stable `sorted` keeps it among equal counts); the `max_select` smallest (count, first appearance) pairs win, and
// a track is touched iff its set meets the selected slots.  (A camera observes a track at most once, so counting sets is
// counting observations.  The previous version counted with shared-memory atomics -- 9216 of them onto ~30 addresses per
// filter -- and ran at 12 % of the HBM roofline.)
__global__ void __launch_bounds__(TT) select_poorest_kernel(msckf_dims dims, int max_select, int prune_above,
                                                            const int* __restrict__ n_cams,
                                                            const int* __restrict__ feat_len, const int* __restrict__ obs_slot,
                                                            unsigned char* __restrict__ remove_mask,
                                                            unsigned char* __restrict__ feat_touch) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = dims.F, M = dims.M, N = dims.N;
    extern __shared__ unsigned long long s_set[];          // [F] slots observing each track
    __shared__ int s_count[MSCKF_MAX_CAMS], s_first[MSCKF_MAX_CAMS];
    __shared__ unsigned char s_sel[MSCKF_MAX_CAMS];
    __shared__ unsigned long long s_selset;
    const int nc = min(n_cams[b], N);
    const int* len = feat_len + (size_t)b * F;
    const int* slot = obs_slot + (size_t)b * F * M;
    constexpr int KS = 4;
    for (int j0 = warp; j0 < F; j0 += KS * (TT / 32)) {
        unsigned lo[KS], hi[KS];
        slot_sets<KS>(len, slot, j0, TT / 32, F, M, nc, lane, lo, hi);
#pragma unroll
        for (int k = 0; k < KS; ++k)
            if (lane == 0 && j0 + k * (TT / 32) < F) s_set[j0 + k * (TT / 32)] = ((unsigned long long)hi[k] << 32) | lo[k];
    }
    for (int sl = tid; sl < N; sl += TT) s_sel[sl] = 0;
    __syncthreads();
    if (tid < nc) {
        int cnt = 0, jf = -1;
        for (int j = 0; j < F; ++j) {
            const int hit = (int)((s_set[j] >> tid) & 1ull);
            if (hit && jf < 0) jf = j;
            cnt += hit;
        }
        int first = 0x7fffffff;
        if (jf >= 0) {
            const int n = min(len[jf], M);
            for (int i = 0; i < n; ++i)
                if (slot[(size_t)jf * M + i] == tid) { first = jf * M + i; break; }
        }
        s_count[tid] = cnt;
        s_first[tid] = first;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long sel = 0;
        if (nc > prune_above) {      // MSCKF.feature_callback prunes only windows longer than the maximum (MSCKF.py:157-158)
            for (int pick = 0; pick < max_select; ++pick) {
                int best = -1;
                for (int sl = 0; sl < nc; ++sl) {
                    if (s_sel[sl] || s_count[sl] == 0) continue;
                    if (best < 0 || s_count[sl] < s_count[best] || (s_count[sl] == s_count[best] && s_first[sl] < s_first[best])) best = sl;
                }
                if (best < 0) break;
                s_sel[best] = 1;
                sel |= 1ull << best;
            }
        }
        s_selset = sel;
    }
    __syncthreads();
    for (int sl = tid; sl < N; sl += TT) remove_mask[(size_t)b * N + sl] = s_sel[sl];
    const unsigned long long sel = s_selset;
    for (int j = tid; j < F; j += TT) feat_touch[(size_t)b * F + j] = (s_set[j] & sel) ? 1 : 0;
}

// One warp per track, lane = observation (chunks of 32 for longer tracks): observations of removed slots leave the track,
// the rest keep their order (ballot + popc compaction, in place) and are renumbered; an inverse-depth anchor on a removed
// camera is frozen at that camera's position (InverseDepthPoint.base keeps aliasing the deleted camera's array,
// src/utils/geometry.py:55).  Must run BEFORE msckf_remove_cameras_batched compacts cam_t / n_cams.
__global__ void __launch_bounds__(TT) drop_observations_kernel(msckf_dims dims, const unsigned char* __restrict__ remove_mask,
                                                               const int* __restrict__ n_cams, const double* __restrict__ cam_t,
                                                               int* feat_len, double* obs_uv, double* obs_dir, double* obs_conf,
                                                               int* obs_slot, double* idp_base, int* idp_anchor_slot) {
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int F = dims.F, M = dims.M, N = dims.N;
    __shared__ unsigned char s_rm[MSCKF_MAX_CAMS];
    __shared__ unsigned char s_shift[MSCKF_MAX_CAMS];
    const int nc = min(n_cams[b], N);
    if (tid < N) s_rm[tid] = (tid < nc) ? remove_mask[(size_t)b * N + tid] : 0;
    __syncthreads();
    if (tid < N) {
        int c = 0;
        for (int s = 0; s <= tid; ++s) c += s_rm[s] ? 1 : 0;
        s_shift[tid] = (unsigned char)c;
    }
    __syncthreads();
    if (nc == 0 || s_shift[nc - 1] == 0) return;     // nothing leaves this filter's window: its tracks stay as they are
    // (a few CTAs per filter, each walking its share of the tracks: with one CTA per eight tracks the 147 k early exits of
    //  a frame in which no window loses a camera cost 0.23 ms)
    for (int j = blockIdx.x * (TT / 32) + (tid >> 5); j < F; j += gridDim.x * (TT / 32)) {
    const size_t fj = (size_t)b * F + j;
    const int n = min(feat_len[fj], M);
    int off = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const bool live = i < n;
        const size_t e = fj * M + i;
        int s = 0;
        double u = 0, v = 0, d0 = 0, d1 = 0, d2 = 0, cf = 0;
        if (live) {
            s = obs_slot[e];
            u = obs_uv[e * 2]; v = obs_uv[e * 2 + 1];
            d0 = obs_dir[e * 3]; d1 = obs_dir[e * 3 + 1]; d2 = obs_dir[e * 3 + 2];
            cf = obs_conf[e];
        }
        const bool keep = live && !(s >= 0 && s < nc && s_rm[s]);
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) {
            const size_t w = fj * M + off + __popc(km & ((1u << lane) - 1u));
            obs_slot[w] = (s >= 0 && s < nc) ? s - (int)s_shift[s] : s;
            obs_uv[w * 2] = u; obs_uv[w * 2 + 1] = v;
            obs_dir[w * 3] = d0; obs_dir[w * 3 + 1] = d1; obs_dir[w * 3 + 2] = d2;
            obs_conf[w] = cf;
        }
        off += __popc(km);
    }
    if (lane == 0) {
        feat_len[fj] = off;
        const int a = idp_anchor_slot[fj];
        if (a >= 0 && a < nc) {
            if (s_rm[a]) {
                idp_anchor_slot[fj] = -1;
                for (int k = 0; k < 3; ++k) idp_base[fj * 3 + k] = cam_t[((size_t)b * N + a) * 3 + k];
            } else {
                idp_anchor_slot[fj] = a - (int)s_shift[a];
            }
        }
    }
    }
}

// One thread per track slot: what add_camera_measurements does to a track once the matcher has decided (MSCKF.py:403-412
// append, :400 / :437-438 lost this frame, :414-434 new track).  The new ray is R_newest K^-1 [u, v, 1] (Camera.py:29-44);
// a new track's inverse-depth point is anchored on the newest camera with m from the ray's azimuth / elevation and
// rho = 0.1 (geometry.py:53-59).  HBM traffic: 25 B in + 52 B out per touched track (one new observation each).
struct Kinv9 { double k[9]; };
__global__ void __launch_bounds__(TT) append_observations_kernel(msckf_dims dims, Kinv9 Ki, const int* __restrict__ n_cams,
                                                                 const double* __restrict__ cam_R, const double* __restrict__ cam_t,
                                                                 const unsigned char* __restrict__ action,
                                                                 const double* __restrict__ new_uv, const double* __restrict__ new_conf,
                                                                 int* feat_len, double* obs_uv, double* obs_dir, double* obs_conf,
                                                                 int* obs_slot, double* idp_base, int* idp_anchor_slot, double* idp_m,
                                                                 double* idp_rho, int* lost_cnt, int* tracked_cnt, int* overflow) {
    const int b = blockIdx.y, j = blockIdx.x * TT + threadIdx.x;
    const int F = dims.F, M = dims.M, N = dims.N;
    if (j >= F) return;
    const size_t fj = (size_t)b * F + j;
    // everything a track may need is loaded up front, before the first branch: coalesced, and ONE round trip instead of four
    // dependent ones (the kernel issued 2.8 % of its cycles and sat on the long scoreboard for the rest)
    const int act = action[fj];
    const int len_in = feat_len[fj];
    const int lost_in = lost_cnt[fj], tracked_in = tracked_cnt[fj];
    const double u = new_uv[fj * 2], v = new_uv[fj * 2 + 1];
    const double conf_in = new_conf[fj];
    const int s = min(n_cams[b], N) - 1;
    double R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = cam_R[((size_t)b * N + (s < 0 ? 0 : s)) * 9 + k];
    if (act == 0) return;
    if (act == 1) { lost_cnt[fj] = lost_in + 1; return; }
    if (s < 0) return;
    const int n = (act == 2) ? len_in : 0;
    if (n >= M) { atomicAdd(overflow, 1); return; }      // capacity of the track arrays exceeded: reported, track untouched
    const double c0 = Ki.k[0] * u + Ki.k[1] * v + Ki.k[2];
    const double c1 = Ki.k[3] * u + Ki.k[4] * v + Ki.k[5];
    const double c2 = Ki.k[6] * u + Ki.k[7] * v + Ki.k[8];
    const double w0 = R[0] * c0 + R[1] * c1 + R[2] * c2;
    const double w1 = R[3] * c0 + R[4] * c1 + R[5] * c2;
    const double w2 = R[6] * c0 + R[7] * c1 + R[8] * c2;
    const size_t e = fj * M + n;
    obs_uv[e * 2] = u; obs_uv[e * 2 + 1] = v;
    obs_dir[e * 3] = w0; obs_dir[e * 3 + 1] = w1; obs_dir[e * 3 + 2] = w2;
    obs_conf[e] = conf_in;
    obs_slot[e] = s;
    feat_len[fj] = n + 1;
    if (act == 2) {
        tracked_cnt[fj] = tracked_in + 1;
        lost_cnt[fj] = 0;
    } else {
        const double theta = atan2(w0, w2);
        const double phi = atan2(-w1, sqrt(w0 * w0 + w2 * w2));
        idp_m[fj * 3] = cos(phi) * sin(theta);
        idp_m[fj * 3 + 1] = -sin(phi);
        idp_m[fj * 3 + 2] = cos(phi) * cos(theta);
        idp_rho[fj] = 0.1;
        idp_anchor_slot[fj] = s;
        for (int k = 0; k < 3; ++k) idp_base[fj * 3 + k] = cam_t[((size_t)b * N + s) * 3 + k];
        tracked_cnt[fj] = 1;
        lost_cnt[fj] = 0;
    }
}

// Window slots of a filter that no track observes any more (MSCKF.get_cameras_without_features, called by remove_features,
// MSCKF.py:739-749): the union of the tracks' slot sets (warp per track, OR-reductions), mask = slot not in the union.  A
// filter that is not `active` (no lost track was deleted this frame) does not read its tracks at all.
__global__ void __launch_bounds__(TT) unobserved_cameras_kernel(msckf_dims dims, const int* __restrict__ n_cams,
                                                                const int* __restrict__ feat_len, const int* __restrict__ obs_slot,
                                                                const unsigned char* __restrict__ active,
                                                                unsigned char* __restrict__ remove_mask) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = dims.F, M = dims.M, N = dims.N;
    __shared__ unsigned s_lo, s_hi;
    const bool on = active == nullptr || active[b] != 0;
    if (!on) {
        for (int s = tid; s < N; s += TT) remove_mask[(size_t)b * N + s] = 0;
        return;
    }
    const int nc = min(n_cams[b], N);
    if (tid == 0) { s_lo = 0; s_hi = 0; }
    __syncthreads();
    const int* len = feat_len + (size_t)b * F;
    const int* slot = obs_slot + (size_t)b * F * M;
    unsigned lo = 0, hi = 0;
    constexpr int KS = 4;
    for (int j0 = warp; j0 < F; j0 += KS * (TT / 32)) {
        unsigned l4[KS], h4[KS];
        slot_sets<KS>(len, slot, j0, TT / 32, F, M, nc, lane, l4, h4);
#pragma unroll
        for (int k = 0; k < KS; ++k) { lo |= l4[k]; hi |= h4[k]; }
    }
    if (lane == 0) { atomicOr(&s_lo, lo); atomicOr(&s_hi, hi); }
    __syncthreads();
    const unsigned long long seen = ((unsigned long long)s_hi << 32) | s_lo;
    for (int s = tid; s < N; s += TT) remove_mask[(size_t)b * N + s] = (s < nc && !((seen >> s) & 1ull)) ? 1 : 0;
}

// Stable compaction of the track slots of every filter from one set of track arrays into another (the host swaps the two
// sets afterwards): the reference deletes tracks from its insertion-ordered dict, so the survivors keep their order --
// which the pruning selection's tie-break depends on.  A slot survives when it holds a track (feat_len > 0) and, with
// drop_lost, when the track is not in the `lost_features` dict of the last update (MSCKF_FEAT_LOST_FLAG); drop_lost only
// applies to filters that had at least one valid feature in that update (MSCKF.process_features, MSCKF.py:454-456).
// active_out[b] = 1 when the filter deleted lost tracks (its unobserved cameras are removed next).
struct TrackSet {
    int* feat_len; double* obs_uv; double* obs_dir; double* obs_conf; int* obs_slot; double* idp_base; int* idp_anchor_slot;
    double* idp_m; double* idp_rho; int* lost_cnt; int* tracked_cnt;
};
__global__ void __launch_bounds__(TT) compact_tracks_kernel(msckf_dims dims, int drop_lost, const unsigned char* __restrict__ feat_status,
                                                            TrackSet src, TrackSet dst, unsigned char* __restrict__ active_out,
                                                            int* __restrict__ dest_out) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = dims.F, M = dims.M;
    extern __shared__ int s_dest[];            // [F] destination slot of each surviving track, -1 otherwise
    __shared__ int s_wsum[TT / 32];
    __shared__ int s_base, s_any;
    const size_t fb = (size_t)b * F;
    if (tid == 0) { s_base = 0; s_any = 0; }
    __syncthreads();
    bool lost_rule = false;
    if (drop_lost) {
        int any = 0;
        for (int j = tid; j < F; j += TT) {
            const int code = feat_status[fb + j] & 0xF;
            if (src.feat_len[fb + j] > 0 && (code == MSCKF_FEAT_ACCEPTED || code == MSCKF_FEAT_GATED_OUT)) any = 1;
        }
        if (any) atomicOr(&s_any, 1);
        __syncthreads();
        lost_rule = s_any != 0;
    }
    if (tid == 0 && active_out && blockIdx.y == 0) active_out[b] = lost_rule ? 1 : 0;
    for (int j0 = 0; j0 < F; j0 += TT) {          // block-wide exclusive scan of the keep flags, TT slots at a time
        const int j = j0 + tid;
        int keep = 0;
        if (j < F) keep = (src.feat_len[fb + j] > 0 && !(lost_rule && (feat_status[fb + j] & MSCKF_FEAT_LOST_FLAG))) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_wsum[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_wsum[w];
        if (j < F) s_dest[j] = keep ? off + __popc(bal & ((1u << lane) - 1u)) : -1;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < TT / 32; ++w) t += s_wsum[w]; s_base += t; }
        __syncthreads();
    }
    const int n_kept = s_base;
    if (dest_out && blockIdx.y == 0)
        for (int j = tid; j < F; j += TT) dest_out[fb + j] = s_dest[j];
    // every CTA of the filter repeats the (cheap) scan above and moves its own slice of the tracks
    for (int j = blockIdx.y * (TT / 32) + warp; j < F; j += gridDim.y * (TT / 32)) {      // one warp per track
        const int d = s_dest[j];
        if (d < 0) continue;
        const size_t sj = fb + j, dj = fb + d;
        const int n = min(src.feat_len[sj], M);
        // every array of a track is one contiguous range: linear, fully coalesced copies
        {
            const double* su = src.obs_uv + sj * M * 2; double* du = dst.obs_uv + dj * M * 2;
            for (int x = lane; x < 2 * n; x += 32) du[x] = su[x];
            const double* sd = src.obs_dir + sj * M * 3; double* dd = dst.obs_dir + dj * M * 3;
            for (int x = lane; x < 3 * n; x += 32) dd[x] = sd[x];
            const double* sc = src.obs_conf + sj * M; double* dc = dst.obs_conf + dj * M;
            const int* ss = src.obs_slot + sj * M; int* ds = dst.obs_slot + dj * M;
            for (int x = lane; x < n; x += 32) { dc[x] = sc[x]; ds[x] = ss[x]; }
        }
        if (lane == 0) {
            dst.feat_len[dj] = n;
            dst.idp_anchor_slot[dj] = src.idp_anchor_slot[sj];
            dst.idp_rho[dj] = src.idp_rho[sj];
            dst.lost_cnt[dj] = src.lost_cnt[sj];
            dst.tracked_cnt[dj] = src.tracked_cnt[sj];
        }
        if (lane < 3) { dst.idp_base[dj * 3 + lane] = src.idp_base[sj * 3 + lane]; dst.idp_m[dj * 3 + lane] = src.idp_m[sj * 3 + lane]; }
    }
    if (blockIdx.y == 0)
        for (int j = n_kept + tid; j < F; j += TT) dst.feat_len[fb + j] = 0;
}

// The same compaction IN PLACE, one CTA per filter: a filter in which no track moves (the usual case: nothing was
// deleted, or only tracks at the end) costs one scan of its F lengths; otherwise only the tracks behind the first
// deleted slot are moved, staged through a scratch set of track arrays (pass 1: track -> scratch at its new slot;
// barrier; pass 2: scratch -> the live arrays), so no second full copy of the tracks exists and the caller swaps nothing.
__global__ void __launch_bounds__(TT) compact_tracks_inplace_kernel(msckf_dims dims, int drop_lost,
                                                                    const unsigned char* __restrict__ feat_status, TrackSet trk,
                                                                    TrackSet tmp, unsigned char* __restrict__ active_out,
                                                                    int* __restrict__ dest_out) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = dims.F, M = dims.M;
    extern __shared__ int s_dest[];
    __shared__ int s_wsum[TT / 32];
    __shared__ int s_base, s_any, s_moves;
    const size_t fb = (size_t)b * F;
    if (tid == 0) { s_base = 0; s_any = 0; s_moves = 0; }
    __syncthreads();
    bool lost_rule = false;
    if (drop_lost) {
        int any = 0;
        for (int j = tid; j < F; j += TT) {
            const int code = feat_status[fb + j] & 0xF;
            if (trk.feat_len[fb + j] > 0 && (code == MSCKF_FEAT_ACCEPTED || code == MSCKF_FEAT_GATED_OUT)) any = 1;
        }
        if (any) atomicOr(&s_any, 1);
        __syncthreads();
        lost_rule = s_any != 0;
    }
    if (tid == 0 && active_out) active_out[b] = lost_rule ? 1 : 0;
    int moves = 0;
    for (int j0 = 0; j0 < F; j0 += TT) {
        const int j = j0 + tid;
        int keep = 0;
        if (j < F) keep = (trk.feat_len[fb + j] > 0 && !(lost_rule && (feat_status[fb + j] & MSCKF_FEAT_LOST_FLAG))) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_wsum[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_wsum[w];
        if (j < F) {
            const int d = keep ? off + __popc(bal & ((1u << lane) - 1u)) : -1;
            s_dest[j] = d;
            if (d >= 0 && d != j) moves = 1;
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < TT / 32; ++w) t += s_wsum[w]; s_base += t; }
        __syncthreads();
    }
    if (moves) atomicOr(&s_moves, 1);
    __syncthreads();
    const int n_kept = s_base;
    if (dest_out)
        for (int j = tid; j < F; j += TT) dest_out[fb + j] = s_dest[j];
    if (s_moves && M <= 32) {
        // Single pass through registers.  A kept track only ever moves DOWN (its new slot d <= its old slot j), so with the
        // tracks taken in rounds of increasing j -- every warp first reads its tracks of the round, one CTA barrier, then
        // writes them -- a write can only land on a slot that was read in this round (before the barrier) or in an earlier
        // one, and never on a slot a later round still has to read.  (The two-pass route through the scratch set moved every
        // byte twice: 1.0 -> 0.6 ms per frame of 4096 filters.)
        constexpr int NW = TT / 32, KT = TRACKS_KT;
        for (int j0 = 0; j0 < F; j0 += NW * KT) {
            double uv[KT][2], dr[KT][3], cf[KT], rho[KT], bm[KT];
            int sl[KT], dd[KT], nn[KT], ia[KT];
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                const int j = j0 + warp * KT + k;
                const int d = j < F ? s_dest[j] : -1;
                dd[k] = (d >= 0 && d != j) ? d : -1;
                nn[k] = 0;
                if (dd[k] >= 0) {
                    const size_t sj = fb + j;
                    const int n = min(trk.feat_len[sj], M);
                    nn[k] = n;
                    const double* su = trk.obs_uv + sj * M * 2;
                    const double* sd = trk.obs_dir + sj * M * 3;
#pragma unroll
                    for (int q = 0; q < 2; ++q) uv[k][q] = (lane + 32 * q < 2 * n) ? su[lane + 32 * q] : 0.0;
#pragma unroll
                    for (int q = 0; q < 3; ++q) dr[k][q] = (lane + 32 * q < 3 * n) ? sd[lane + 32 * q] : 0.0;
                    cf[k] = lane < n ? trk.obs_conf[sj * M + lane] : 0.0;
                    sl[k] = lane < n ? trk.obs_slot[sj * M + lane] : 0;
                    // lanes 0-2: idp_base, lanes 3-5: idp_m, lane 6: rho; lane 0: anchor, lane 1: lost, lane 2: tracked
                    bm[k] = lane < 3 ? trk.idp_base[sj * 3 + lane] : (lane < 6 ? trk.idp_m[sj * 3 + lane - 3] : 0.0);
                    rho[k] = lane == 6 ? trk.idp_rho[sj] : 0.0;
                    ia[k] = lane == 0 ? trk.idp_anchor_slot[sj] : (lane == 1 ? trk.lost_cnt[sj] : (lane == 2 ? trk.tracked_cnt[sj] : 0));
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                if (dd[k] < 0) continue;
                const size_t dj = fb + dd[k];
                const int n = nn[k];
                double* du = trk.obs_uv + dj * M * 2;
                double* dx = trk.obs_dir + dj * M * 3;
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (lane + 32 * q < 2 * n) du[lane + 32 * q] = uv[k][q];
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    if (lane + 32 * q < 3 * n) dx[lane + 32 * q] = dr[k][q];
                if (lane < n) { trk.obs_conf[dj * M + lane] = cf[k]; trk.obs_slot[dj * M + lane] = sl[k]; }
                if (lane < 3) trk.idp_base[dj * 3 + lane] = bm[k];
                else if (lane < 6) trk.idp_m[dj * 3 + lane - 3] = bm[k];
                else if (lane == 6) trk.idp_rho[dj] = rho[k];
                if (lane == 0) { trk.idp_anchor_slot[dj] = ia[k]; trk.feat_len[dj] = n; }
                else if (lane == 1) trk.lost_cnt[dj] = ia[k];
                else if (lane == 2) trk.tracked_cnt[dj] = ia[k];
            }
        }
        __threadfence_block();
        __syncthreads();
    } else if (s_moves) {
        for (int pass = 0; pass < 2; ++pass) {
            const TrackSet& from = pass == 0 ? trk : tmp;
            const TrackSet& to = pass == 0 ? tmp : trk;
            for (int j = warp; j < F; j += TT / 32) {
                const int d = s_dest[j];
                if (d < 0 || d == j) continue;
                const size_t sj = fb + (pass == 0 ? j : d), dj = fb + d;
                const int n = min(from.feat_len[sj], M);
                const double* su = from.obs_uv + sj * M * 2; double* du = to.obs_uv + dj * M * 2;
                for (int x = lane; x < 2 * n; x += 32) du[x] = su[x];
                const double* sd = from.obs_dir + sj * M * 3; double* dd = to.obs_dir + dj * M * 3;
                for (int x = lane; x < 3 * n; x += 32) dd[x] = sd[x];
                const double* sc = from.obs_conf + sj * M; double* dc = to.obs_conf + dj * M;
                const int* ss = from.obs_slot + sj * M; int* ds = to.obs_slot + dj * M;
                for (int x = lane; x < n; x += 32) { dc[x] = sc[x]; ds[x] = ss[x]; }
                if (lane == 0) {
                    to.idp_anchor_slot[dj] = from.idp_anchor_slot[sj];
                    to.idp_rho[dj] = from.idp_rho[sj];
                    to.lost_cnt[dj] = from.lost_cnt[sj];
                    to.tracked_cnt[dj] = from.tracked_cnt[sj];
                }
                if (lane < 3) { to.idp_base[dj * 3 + lane] = from.idp_base[sj * 3 + lane]; to.idp_m[dj * 3 + lane] = from.idp_m[sj * 3 + lane]; }
                __syncwarp();
                if (lane == 0) to.feat_len[dj] = n;       // (pass 0 reads from.feat_len[sj] of OTHER tracks only through tmp: trk untouched)
            }
            __threadfence_block();
            __syncthreads();
        }
    }
    for (int j = n_kept + tid; j < F; j += TT) trk.feat_len[fb + j] = 0;
}

}  // namespace

cudaError_t launch_unobserved_cameras(const msckf_dims& d, const int* n_cams, const int* feat_len, const int* obs_slot,
                                      const unsigned char* active, unsigned char* remove_mask, cudaStream_t s) {
    unobserved_cameras_kernel<<<d.B, TT, 0, s>>>(d, n_cams, feat_len, obs_slot, active, remove_mask);
    msckf_count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_compact_tracks(const msckf_dims& d, int drop_lost, const unsigned char* feat_status, void* const* src,
                                  void* const* dst, unsigned char* active_out, int* dest_out, cudaStream_t s) {
    TrackSet a{(int*)src[0], (double*)src[1], (double*)src[2], (double*)src[3], (int*)src[4], (double*)src[5], (int*)src[6],
               (double*)src[7], (double*)src[8], (int*)src[9], (int*)src[10]};
    TrackSet b{(int*)dst[0], (double*)dst[1], (double*)dst[2], (double*)dst[3], (int*)dst[4], (double*)dst[5], (int*)dst[6],
               (double*)dst[7], (double*)dst[8], (int*)dst[9], (int*)dst[10]};
    int slices = (d.F + 63) / 64;                       // ~64 tracks (8 per warp) per CTA
    if (slices < 1) slices = 1;
    if (slices > 16) slices = 16;
    compact_tracks_kernel<<<dim3(d.B, slices), TT, (size_t)d.F * sizeof(int), s>>>(d, drop_lost, feat_status, a, b, active_out, dest_out);
    msckf_count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_append_observations(const msckf_dims& d, const double* Kinv_host, const int* n_cams, const double* cam_R,
                                       const double* cam_t, const unsigned char* action, const double* new_uv,
                                       const double* new_conf, int* feat_len, double* obs_uv, double* obs_dir, double* obs_conf,
                                       int* obs_slot, double* idp_base, int* idp_anchor_slot, double* idp_m, double* idp_rho,
                                       int* lost_cnt, int* tracked_cnt, int* overflow, cudaStream_t s) {
    if (d.F == 0) return cudaSuccess;
    Kinv9 Ki;
    for (int k = 0; k < 9; ++k) Ki.k[k] = Kinv_host[k];
    dim3 grid((d.F + TT - 1) / TT, d.B);
    append_observations_kernel<<<grid, TT, 0, s>>>(d, Ki, n_cams, cam_R, cam_t, action, new_uv, new_conf, feat_len, obs_uv, obs_dir,
                                                   obs_conf, obs_slot, idp_base, idp_anchor_slot, idp_m, idp_rho, lost_cnt,
                                                   tracked_cnt, overflow);
    msckf_count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_select_poorest(const msckf_dims& d, int max_select, int prune_above, const int* n_cams, const int* feat_len,
                                  const int* obs_slot, unsigned char* remove_mask, unsigned char* feat_touch, cudaStream_t s) {
    if (d.N > 64) return cudaErrorInvalidValue;      // (slot sets are 64-bit; MSCKF_MAX_CAMS)
    select_poorest_kernel<<<d.B, TT, (size_t)d.F * sizeof(unsigned long long), s>>>(d, max_select, prune_above, n_cams, feat_len,
                                                                                    obs_slot, remove_mask, feat_touch);
    msckf_count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_drop_observations(const msckf_dims& d, const unsigned char* remove_mask, const int* n_cams,
                                     const double* cam_t, int* feat_len, double* obs_uv, double* obs_dir, double* obs_conf,
                                     int* obs_slot, double* idp_base, int* idp_anchor_slot, cudaStream_t s) {
    if (d.F == 0) return cudaSuccess;
    dim3 grid(min(4, (d.F + TT / 32 - 1) / (TT / 32)), d.B);
    drop_observations_kernel<<<grid, TT, 0, s>>>(d, remove_mask, n_cams, cam_t, feat_len, obs_uv, obs_dir, obs_conf, obs_slot,
                                                 idp_base, idp_anchor_slot);
    msckf_count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_compact_tracks_inplace(const msckf_dims& d, int drop_lost, const unsigned char* feat_status, void* const* trk,
                                          void* const* tmp, unsigned char* active_out, int* dest_out, cudaStream_t s) {
    TrackSet a{(int*)trk[0], (double*)trk[1], (double*)trk[2], (double*)trk[3], (int*)trk[4], (double*)trk[5], (int*)trk[6],
               (double*)trk[7], (double*)trk[8], (int*)trk[9], (int*)trk[10]};
    TrackSet b{(int*)tmp[0], (double*)tmp[1], (double*)tmp[2], (double*)tmp[3], (int*)tmp[4], (double*)tmp[5], (int*)tmp[6],
               (double*)tmp[7], (double*)tmp[8], (int*)tmp[9], (int*)tmp[10]};
    compact_tracks_inplace_kernel<<<d.B, TT, (size_t)d.F * sizeof(int), s>>>(d, drop_lost, feat_status, a, b, active_out, dest_out);
    msckf_count_launch(1);
    return cudaGetLastError();
}
