// C-ABI entry points of libmsckf_b200.so (declared in include/msckf_b200.h).  Argument checking + workspace carve-up +
// kernel launches on the caller's stream; no allocation, no synchronisation, no torch types.
#include <atomic>
#include <cstdlib>
#include "msckf_common.cuh"

static std::atomic<long long> g_launches{0};
void msckf_count_launch(int n) { g_launches += n; }
int tsqr_max_cams();
bool gram_supported(const msckf_dims& d);
bool front_split_supported(const msckf_dims& d);
int front_split_sub_filters(const msckf_dims& d);
size_t front_split_tile_bytes(const msckf_dims& d);
int front_split_map_bytes(const msckf_dims& d);
cudaError_t launch_front_split(const FrontArgs& a, double2* Kbuf, int* pend_list, int* pend_count, int* pend_key,
                               unsigned char* slot_map, cudaStream_t s);
size_t info_scratch_doubles(const msckf_dims& d);
cudaError_t launch_update_info(const ApplyArgs& a, int n_ctas, cudaStream_t s);
bool update_smem_supported(const msckf_dims& d);
size_t update_smem_scratch_doubles(const msckf_dims& d);
cudaError_t launch_update_smem(const ApplyArgs& a, int n_ctas, cudaStream_t s);
cudaError_t launch_select_poorest(const msckf_dims& d, int max_select, int prune_above, const int* n_cams, const int* feat_len,
                                  const int* obs_slot, unsigned char* remove_mask, unsigned char* feat_touch, cudaStream_t s);
cudaError_t launch_unobserved_cameras(const msckf_dims& d, const int* n_cams, const int* feat_len, const int* obs_slot,
                                      const unsigned char* active, unsigned char* remove_mask, cudaStream_t s);
cudaError_t launch_compact_tracks(const msckf_dims& d, int drop_lost, const unsigned char* feat_status, void* const* src,
                                  void* const* dst, unsigned char* active_out, int* dest_out, cudaStream_t s);
cudaError_t launch_compact_tracks_inplace(const msckf_dims& d, int drop_lost, const unsigned char* feat_status, void* const* trk,
                                          void* const* tmp, unsigned char* active_out, int* dest_out, cudaStream_t s);
cudaError_t launch_append_observations(const msckf_dims& d, const double* Kinv_host, const int* n_cams, const double* cam_R,
                                       const double* cam_t, const unsigned char* action, const double* new_uv,
                                       const double* new_conf, int* feat_len, double* obs_uv, double* obs_dir, double* obs_conf,
                                       int* obs_slot, double* idp_base, int* idp_anchor_slot, double* idp_m, double* idp_rho,
                                       int* lost_cnt, int* tracked_cnt, int* overflow, cudaStream_t s);
cudaError_t launch_drop_observations(const msckf_dims& d, const unsigned char* remove_mask, const int* n_cams,
                                     const double* cam_t, int* feat_len, double* obs_uv, double* obs_dir, double* obs_conf,
                                     int* obs_slot, double* idp_base, int* idp_anchor_slot, cudaStream_t s);
cudaError_t launch_verify_matches(const msckf_dims& d, const double* K_host, const double* Kinv_host, double epipolar_thr,
                                  double homography_thr, const int* n_cams, const double* cam_R, const double* cam_t,
                                  const int* feat_len, const double* obs_uv, const int* obs_slot, const double* new_uv,
                                  unsigned char* action, unsigned char* verdict, int* n_homography, int* n_epipolar,
                                  cudaStream_t s);
int launch_match_descriptors(int B, int F1, int F2, int D, double min_cossim, const int* n_tracks, const int* n_new,
                             const float* desc_tracks, const float* desc_new, int* match12, int* match21,
                             const double* kp_new, const double* score_new, unsigned char* action, double* new_uv,
                             double* new_conf, int* n_tracks_out, int* overflow, cudaStream_t s);

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

WsLayout msckf_ws_layout(const msckf_dims& d, int mode) {
    WsLayout w{};
    const size_t BF = (size_t)d.B * d.F, ccap = 6 * (size_t)d.N;
    size_t off = 0;
    w.fac_A = off;    off = align256(off + BF * d.M * 12 * sizeof(double));
    w.fac_Q = off;    off = align256(off + BF * d.M * 6 * sizeof(double));
    w.fac_rt = off;   off = align256(off + BF * d.M * 2 * sizeof(double));
    w.fac_rank = off; off = align256(off + BF * sizeof(int));
    w.T_H = off;      off = align256(off + (size_t)d.B * ccap * ccap * sizeof(double));
    w.r_n = off;      off = align256(off + (size_t)d.B * ccap * sizeof(double));
    w.nacc = off;     off = align256(off + (size_t)d.B * sizeof(int));
    w.nrej = off;     off = align256(off + (size_t)d.B * sizeof(int));
    w.n_update_ctas = apply_num_ctas(d);
    size_t sd = apply_scratch_doubles(d) > info_scratch_doubles(d) ? apply_scratch_doubles(d) : info_scratch_doubles(d);
    if (update_smem_supported(d) && update_smem_scratch_doubles(d) > sd) sd = update_smem_scratch_doubles(d);
    w.scratch_per_cta = align256(sd * sizeof(double));
    w.scratch = off;  off = align256(off + (size_t)w.n_update_ctas * w.scratch_per_cta);
    if (front_split_supported(d)) {
        w.k_tiles = off;    off = align256(off + front_split_tile_bytes(d));
        w.pend_list = off;  off = align256(off + (size_t)front_split_sub_filters(d) * d.F * sizeof(int));
        w.pend_count = off; off = align256(off + sizeof(int));
        w.pend_key = off;   off = align256(off + (size_t)front_split_sub_filters(d) * d.F * sizeof(int));
        w.slot_map = off;   off = align256(off + (size_t)front_split_sub_filters(d) * d.F * front_split_map_bytes(d));
    }
    w.total = off;
    (void)mode;
    return w;
}

static int check_dims(const msckf_dims* d) {
    if (!d) return MSCKF_EINVAL;
    if (d->B <= 0 || d->F < 0 || d->M <= 0 || d->N <= 0 || d->ldp < 15 + 6 * d->N) return MSCKF_EINVAL;
    if (d->B > 65535 || d->M > MSCKF_MAX_OBS || d->N > MSCKF_MAX_CAMS) return MSCKF_ECAPACITY;
    return 0;
}

extern "C" {

int msckf_abi_version(void) { return MSCKF_B200_ABI_VERSION; }
int msckf_max_obs_per_feature(void) { return MSCKF_MAX_OBS; }
int msckf_max_cameras(int mode) { return mode == MSCKF_COMPRESS_QR ? tsqr_max_cams() : MSCKF_MAX_CAMS; }
long long msckf_launch_count(void) { return g_launches.load(); }

size_t msckf_update_workspace_bytes(const msckf_dims* dims, int mode) {
    if (check_dims(dims) != 0) return 0;
    return msckf_ws_layout(*dims, mode).total;
}

static FrontArgs make_front(const msckf_dims* dims, const msckf_update_params* params, const double* chi2_table,
                            int chi2_len, const double* P, const double* cam_R, const double* cam_t, const int* n_cams,
                            const int* feat_len, const double* obs_uv, const double* obs_dir, const double* obs_conf,
                            const int* obs_slot, const double* idp_base, const int* idp_anchor_slot, double* idp_m,
                            double* idp_rho, const int* lost_cnt, const int* tracked_cnt, unsigned char* feat_status,
                            double* gamma, int* dof, char* ws, const WsLayout& w) {
    FrontArgs a{};
    a.dims = *dims; a.par = *params; a.chi2_table = chi2_table; a.chi2_len = chi2_len;
    a.P = P; a.cam_R = cam_R; a.cam_t = cam_t; a.n_cams = n_cams; a.feat_len = feat_len;
    a.obs_uv = obs_uv; a.obs_dir = obs_dir; a.obs_conf = obs_conf; a.obs_slot = obs_slot;
    a.idp_base = idp_base; a.idp_anchor_slot = idp_anchor_slot; a.idp_m = idp_m; a.idp_rho = idp_rho;
    a.lost_cnt = lost_cnt; a.tracked_cnt = tracked_cnt; a.feat_status = feat_status; a.gamma = gamma; a.dof = dof;
    a.fac_A = (double*)(ws + w.fac_A); a.fac_Q = (double*)(ws + w.fac_Q); a.fac_rt = (double*)(ws + w.fac_rt);
    a.fac_rank = (int*)(ws + w.fac_rank);
    return a;
}

int msckf_front_batched(const msckf_dims* dims, const msckf_update_params* params, const double* chi2_table,
                        int chi2_len, const double* P, const double* cam_R, const double* cam_t, const int* n_cams,
                        const int* feat_len, const double* obs_uv, const double* obs_dir, const double* obs_conf,
                        const int* obs_slot, const double* idp_base, const int* idp_anchor_slot, double* idp_m,
                        double* idp_rho, const int* lost_cnt, const int* tracked_cnt, unsigned char* feat_status,
                        double* gamma, int* dof, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_dims(dims);
    if (rc) return rc;
    if (!params || !chi2_table || !P || !cam_R || !cam_t || !n_cams || !feat_len || !obs_uv || !obs_dir || !obs_conf ||
        !obs_slot || !idp_base || !idp_anchor_slot || !idp_m || !idp_rho || !lost_cnt || !tracked_cnt || !feat_status ||
        !gamma || !dof || !workspace)
        return MSCKF_EINVAL;
    if (chi2_len < 2 * dims->M + 1) return MSCKF_EINVAL;
    if (dims->F == 0) return 0;
    const WsLayout w = msckf_ws_layout(*dims, params->compress_mode);
    if (workspace_bytes < w.total) return MSCKF_EWORKSPACE;
    FrontArgs a = make_front(dims, params, chi2_table, chi2_len, P, cam_R, cam_t, n_cams, feat_len, obs_uv, obs_dir,
                             obs_conf, obs_slot, idp_base, idp_anchor_slot, idp_m, idp_rho, lost_cnt, tracked_cnt,
                             feat_status, gamma, dof, (char*)workspace, w);
    // warp-per-feature pair (assembly kernel + factorisation kernel) when the track length / window fit its tile budget,
    // CTA-per-feature kernel otherwise
    if (front_split_supported(*dims))
        return (int)launch_front_split(a, (double2*)((char*)workspace + w.k_tiles), (int*)((char*)workspace + w.pend_list),
                                       (int*)((char*)workspace + w.pend_count), (int*)((char*)workspace + w.pend_key),
                                       (unsigned char*)workspace + w.slot_map, (cudaStream_t)stream);
    return (int)launch_front(a, (cudaStream_t)stream);
}

int msckf_compress_batched(const msckf_dims* dims, const msckf_update_params* params, const int* n_cams,
                           const int* feat_len, const int* obs_slot, const unsigned char* feat_status, double* T_H,
                           double* r_n, int* n_accepted, int* n_rejected, void* workspace, size_t workspace_bytes,
                           void* stream) {
    int rc = check_dims(dims);
    if (rc) return rc;
    if (!params || !n_cams || !feat_len || !obs_slot || !feat_status || !T_H || !r_n || !n_accepted || !n_rejected ||
        !workspace)
        return MSCKF_EINVAL;
    if (dims->N > msckf_max_cameras(params->compress_mode)) return MSCKF_ECAPACITY;
    if (params->compress_mode != MSCKF_COMPRESS_QR && !gram_supported(*dims)) return MSCKF_ECAPACITY;
    const WsLayout w = msckf_ws_layout(*dims, params->compress_mode);
    if (workspace_bytes < w.total) return MSCKF_EWORKSPACE;
    char* ws = (char*)workspace;
    CompressArgs c{};
    c.dims = *dims; c.mode = params->compress_mode; c.n_cams = n_cams; c.feat_len = feat_len; c.obs_slot = obs_slot;
    c.feat_status = feat_status; c.fac_A = (double*)(ws + w.fac_A); c.fac_Q = (double*)(ws + w.fac_Q);
    c.fac_rt = (double*)(ws + w.fac_rt); c.T_H = T_H; c.r_n = r_n; c.n_accepted = n_accepted; c.n_rejected = n_rejected;
    if (params->compress_mode == MSCKF_COMPRESS_QR) return (int)launch_compress_qr(c, (cudaStream_t)stream);
    return (int)launch_compress_gram(c, (cudaStream_t)stream);
}

int msckf_apply_update_batched(const msckf_dims* dims, const msckf_update_params* params, double* P, double* imu_state,
                               double* cam_R, double* cam_t, const int* n_cams, const double* T_H, const double* r_n,
                               const int* n_accepted, double* delta_x, void* workspace, size_t workspace_bytes,
                               void* stream) {
    int rc = check_dims(dims);
    if (rc) return rc;
    if (!params || !P || !imu_state || !cam_R || !cam_t || !n_cams || !T_H || !r_n || !n_accepted || !delta_x || !workspace)
        return MSCKF_EINVAL;
    const WsLayout w = msckf_ws_layout(*dims, params->compress_mode);
    if (workspace_bytes < w.total) return MSCKF_EWORKSPACE;
    ApplyArgs a{};
    a.dims = *dims; a.sigma = params->sigma_image; a.P = P; a.imu_state = imu_state; a.cam_R = cam_R; a.cam_t = cam_t;
    a.n_cams = n_cams; a.T_H = T_H; a.r_n = r_n; a.n_accepted = n_accepted; a.delta_x = delta_x;
    a.scratch = (double*)((char*)workspace + w.scratch);
    a.scratch_per_cta = w.scratch_per_cta / sizeof(double);
    if (params->compress_mode == MSCKF_COMPRESS_QR) return (int)launch_apply(a, w.n_update_ctas, (cudaStream_t)stream);
    if (update_smem_supported(*dims)) return (int)launch_update_smem(a, w.n_update_ctas, (cudaStream_t)stream);
    return (int)launch_update_info(a, w.n_update_ctas, (cudaStream_t)stream);
}

int msckf_update_batched(const msckf_dims* dims, const msckf_update_params* params, const double* chi2_table,
                         int chi2_len, double* P, double* imu_state, double* cam_R, double* cam_t, const int* n_cams,
                         const int* feat_len, const double* obs_uv, const double* obs_dir, const double* obs_conf,
                         const int* obs_slot, const double* idp_base, const int* idp_anchor_slot, double* idp_m,
                         double* idp_rho, const int* lost_cnt, const int* tracked_cnt, unsigned char* feat_status,
                         double* gamma, int* dof, double* delta_x, int* n_accepted, int* n_rejected, void* workspace,
                         size_t workspace_bytes, void* stream) {
    int rc = msckf_front_batched(dims, params, chi2_table, chi2_len, P, cam_R, cam_t, n_cams, feat_len, obs_uv, obs_dir,
                                 obs_conf, obs_slot, idp_base, idp_anchor_slot, idp_m, idp_rho, lost_cnt, tracked_cnt,
                                 feat_status, gamma, dof, workspace, workspace_bytes, stream);
    if (rc) return rc;
    if (!imu_state || !delta_x || !n_accepted || !n_rejected) return MSCKF_EINVAL;
    const WsLayout w = msckf_ws_layout(*dims, params->compress_mode);
    char* ws = (char*)workspace;
    double* T_H = (double*)(ws + w.T_H);
    double* r_n = (double*)(ws + w.r_n);
    rc = msckf_compress_batched(dims, params, n_cams, feat_len, obs_slot, feat_status, T_H, r_n, n_accepted, n_rejected,
                                workspace, workspace_bytes, stream);
    if (rc) return rc;
    return msckf_apply_update_batched(dims, params, P, imu_state, cam_R, cam_t, n_cams, T_H, r_n, n_accepted, delta_x,
                                      workspace, workspace_bytes, stream);
}

int msckf_propagate_batched(const msckf_dims* dims, int n_samples, const double* imu_samples, double* imu_time,
                            const double* Qc_diag, const double* gravity_host, double* imu_state, double* imu_null,
                            double* P, const int* n_cams, void* stream) {
    int rc = check_dims(dims);
    if (rc) return rc;
    if (n_samples < 0 || !imu_samples || !imu_time || !Qc_diag || !gravity_host || !imu_state || !P || !n_cams)
        return MSCKF_EINVAL;
    if (n_samples == 0) return 0;
    return (int)launch_propagate(*dims, n_samples, imu_samples, imu_time, Qc_diag, gravity_host, imu_state, imu_null, P,
                                 n_cams, (cudaStream_t)stream);
}

int msckf_augment_batched(const msckf_dims* dims, const double* T_I_C_host, const double* imu_state, double* cam_R,
                          double* cam_t, double* P, int* n_cams, const unsigned char* active, int* overflow_count,
                          void* stream) {
    int rc = check_dims(dims);
    if (rc) return rc;
    if (!T_I_C_host || !imu_state || !cam_R || !cam_t || !P || !n_cams) return MSCKF_EINVAL;
    if (15 + 6 * dims->N > 512) return MSCKF_ECAPACITY;
    return (int)launch_augment(*dims, T_I_C_host, imu_state, cam_R, cam_t, P, n_cams, active, overflow_count,
                               (cudaStream_t)stream);
}

size_t msckf_remove_workspace_bytes(const msckf_dims* dims) {
    if (check_dims(dims) != 0) return 0;
    const int ctas = dims->B < 148 * 4 ? dims->B : 148 * 4;
    return (size_t)ctas * dims->ldp * dims->ldp * sizeof(double);
}

int msckf_remove_cameras_batched(const msckf_dims* dims, const unsigned char* remove_mask, double* cam_R, double* cam_t,
                                 double* P, int* n_cams, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_dims(dims);
    if (rc) return rc;
    if (!remove_mask || !cam_R || !cam_t || !P || !n_cams || !workspace) return MSCKF_EINVAL;
    if (dims->N > 128) return MSCKF_ECAPACITY;
    if (workspace_bytes < msckf_remove_workspace_bytes(dims)) return MSCKF_EWORKSPACE;
    return (int)launch_remove(*dims, remove_mask, cam_R, cam_t, P, n_cams, (double*)workspace, (cudaStream_t)stream);
}

int msckf_select_poorest_cameras_batched(const msckf_dims* dims, int max_select, int prune_above, const int* n_cams,
                                         const int* feat_len, const int* obs_slot, unsigned char* remove_mask,
                                         unsigned char* feat_touch, void* stream) {
    int rc = check_dims(dims);
    if (rc) return rc;
    if (max_select < 0 || !n_cams || !feat_len || !obs_slot || !remove_mask || !feat_touch) return MSCKF_EINVAL;
    return (int)launch_select_poorest(*dims, max_select, prune_above, n_cams, feat_len, obs_slot, remove_mask, feat_touch,
                                      (cudaStream_t)stream);
}

int msckf_unobserved_cameras_batched(const msckf_dims* dims, const int* n_cams, const int* feat_len, const int* obs_slot,
                                     const unsigned char* active, unsigned char* remove_mask, void* stream) {
    int rc = check_dims(dims);
    if (rc) return rc;
    if (!n_cams || !feat_len || !obs_slot || !remove_mask) return MSCKF_EINVAL;
    return (int)launch_unobserved_cameras(*dims, n_cams, feat_len, obs_slot, active, remove_mask, (cudaStream_t)stream);
}

int msckf_compact_tracks_batched(const msckf_dims* dims, int drop_lost, const unsigned char* feat_status,
                                 const msckf_track_arrays* src, const msckf_track_arrays* dst, unsigned char* active_out,
                                 int* dest_out, void* stream) {
    int rc = check_dims(dims);
    if (rc) return rc;
    if (!src || !dst || (drop_lost && !feat_status)) return MSCKF_EINVAL;
    const void* const* ps = (const void* const*)src;
    const void* const* pq = (const void* const*)dst;
    for (int k = 0; k < 11; ++k)
        if (!ps[k] || !pq[k] || ps[k] == pq[k]) return MSCKF_EINVAL;
    if ((size_t)dims->F * sizeof(int) > 48 * 1024) return MSCKF_ECAPACITY;
    return (int)launch_compact_tracks(*dims, drop_lost, feat_status, (void* const*)src, (void* const*)dst, active_out,
                                      dest_out, (cudaStream_t)stream);
}

int msckf_compact_tracks_inplace_batched(const msckf_dims* dims, int drop_lost, const unsigned char* feat_status,
                                         const msckf_track_arrays* tracks, const msckf_track_arrays* scratch,
                                         unsigned char* active_out, int* dest_out, void* stream) {
    int rc = check_dims(dims);
    if (rc) return rc;
    if (!tracks || !scratch || (drop_lost && !feat_status)) return MSCKF_EINVAL;
    const void* const* ps = (const void* const*)tracks;
    const void* const* pq = (const void* const*)scratch;
    for (int k = 0; k < 11; ++k)
        if (!ps[k] || !pq[k] || ps[k] == pq[k]) return MSCKF_EINVAL;
    if ((size_t)dims->F * sizeof(int) > 48 * 1024) return MSCKF_ECAPACITY;
    return (int)launch_compact_tracks_inplace(*dims, drop_lost, feat_status, (void* const*)tracks, (void* const*)scratch, active_out,
                                              dest_out, (cudaStream_t)stream);
}

int msckf_append_observations_batched(const msckf_dims* dims, const double* Kinv_host, const int* n_cams, const double* cam_R,
                                      const double* cam_t, const unsigned char* action, const double* new_uv,
                                      const double* new_conf, int* feat_len, double* obs_uv, double* obs_dir, double* obs_conf,
                                      int* obs_slot, double* idp_base, int* idp_anchor_slot, double* idp_m, double* idp_rho,
                                      int* lost_cnt, int* tracked_cnt, int* overflow_count, void* stream) {
    int rc = check_dims(dims);
    if (rc) return rc;
    if (!Kinv_host || !n_cams || !cam_R || !cam_t || !action || !new_uv || !new_conf || !feat_len || !obs_uv || !obs_dir ||
        !obs_conf || !obs_slot || !idp_base || !idp_anchor_slot || !idp_m || !idp_rho || !lost_cnt || !tracked_cnt ||
        !overflow_count)
        return MSCKF_EINVAL;
    return (int)launch_append_observations(*dims, Kinv_host, n_cams, cam_R, cam_t, action, new_uv, new_conf, feat_len, obs_uv,
                                           obs_dir, obs_conf, obs_slot, idp_base, idp_anchor_slot, idp_m, idp_rho, lost_cnt,
                                           tracked_cnt, overflow_count, (cudaStream_t)stream);
}

int msckf_drop_observations_batched(const msckf_dims* dims, const unsigned char* remove_mask, const int* n_cams,
                                    const double* cam_t, int* feat_len, double* obs_uv, double* obs_dir, double* obs_conf,
                                    int* obs_slot, double* idp_base, int* idp_anchor_slot, void* stream) {
    int rc = check_dims(dims);
    if (rc) return rc;
    if (!remove_mask || !n_cams || !cam_t || !feat_len || !obs_uv || !obs_dir || !obs_conf || !obs_slot || !idp_base ||
        !idp_anchor_slot)
        return MSCKF_EINVAL;
    return (int)launch_drop_observations(*dims, remove_mask, n_cams, cam_t, feat_len, obs_uv, obs_dir, obs_conf, obs_slot,
                                         idp_base, idp_anchor_slot, (cudaStream_t)stream);
}

int msckf_verify_matches_batched(const msckf_dims* dims, const double* K_host, const double* Kinv_host,
                                 double epipolar_rejection_threshold, double homography_rejection_threshold,
                                 const int* n_cams, const double* cam_R, const double* cam_t, const int* feat_len,
                                 const double* obs_uv, const int* obs_slot, const double* new_uv, unsigned char* action,
                                 unsigned char* verdict, int* n_homography_rejected, int* n_epipolar_rejected, void* stream) {
    int rc = check_dims(dims);
    if (rc) return rc;
    if (!K_host || !Kinv_host || !n_cams || !cam_R || !cam_t || !feat_len || !obs_uv || !obs_slot || !new_uv || !action ||
        !verdict)
        return MSCKF_EINVAL;
    return (int)launch_verify_matches(*dims, K_host, Kinv_host, epipolar_rejection_threshold, homography_rejection_threshold,
                                      n_cams, cam_R, cam_t, feat_len, obs_uv, obs_slot, new_uv, action, verdict,
                                      n_homography_rejected, n_epipolar_rejected, (cudaStream_t)stream);
}

int msckf_match_descriptors_batched(int B, int F_tracks, int F_new, int D, double min_cossim, const int* n_tracks,
                                    const int* n_new, const float* desc_tracks, const float* desc_new, int* match12,
                                    int* match21, const double* kp_new, const double* score_new, unsigned char* action,
                                    double* new_uv, double* new_conf, int* n_tracks_out, int* overflow_count, void* stream) {
    if (B <= 0 || F_tracks <= 0 || F_new <= 0 || D <= 0 || !desc_tracks || !desc_new || !match12 || !match21) return MSCKF_EINVAL;
    if (B > 2147483647 / 2) return MSCKF_ECAPACITY;
    if (action && (!kp_new || !score_new || !new_uv || !new_conf)) return MSCKF_EINVAL;
    return launch_match_descriptors(B, F_tracks, F_new, D, min_cossim, n_tracks, n_new, desc_tracks, desc_new, match12, match21,
                                    kp_new, score_new, action, new_uv, new_conf, n_tracks_out, overflow_count,
                                    (cudaStream_t)stream);
}

}  // extern "C"

// ---- host-buffer step: chunked upload / compute / download pipeline ------------------------------------------------
namespace {
constexpr int HS_MAX_CHUNKS = 256;
struct HostStepCtx {
    int device = -1;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_entry = nullptr, ev_out_done = nullptr;
    cudaEvent_t ev_in[HS_MAX_CHUNKS] = {}, ev_done[HS_MAX_CHUNKS] = {}, ev_out[HS_MAX_CHUNKS] = {};
    int n_events = 0;
    int last_chunk_filters = -1, last_B = -1;   // chunk geometry of the previous call (the per-chunk events refer to it)
};
HostStepCtx g_hs[16];

cudaError_t hs_ctx(HostStepCtx** out, int n_chunks) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 16) return cudaErrorInvalidDevice;
    HostStepCtx& c = g_hs[dev];
    if (c.device != dev) {
        if ((e = cudaStreamCreateWithFlags(&c.s_in, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&c.s_out, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&c.ev_entry, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&c.ev_out_done, cudaEventDisableTiming)) != cudaSuccess) return e;
        c.device = dev;
    }
    for (; c.n_events < n_chunks; ++c.n_events) {
        if ((e = cudaEventCreateWithFlags(&c.ev_in[c.n_events], cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&c.ev_done[c.n_events], cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&c.ev_out[c.n_events], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    *out = &c;
    return cudaSuccess;
}

template <class T>
cudaError_t hs_copy(T* dst, const T* src, size_t per_filter, int b0, int nb, cudaMemcpyKind kind, cudaStream_t s) {
    if (!dst || !src || per_filter == 0) return cudaSuccess;
    return cudaMemcpyAsync(dst + (size_t)b0 * per_filter, src + (size_t)b0 * per_filter, (size_t)nb * per_filter * sizeof(T),
                           kind, s);
}
}  // namespace

extern "C" int msckf_filter_step_host(const msckf_dims* dims, const msckf_update_params* params, const double* chi2_table,
                                      int chi2_len, const msckf_device_buffers* dev, const msckf_host_step* io,
                                      int chunk_filters, int overlap_previous, void* workspace, size_t workspace_bytes,
                                      void* stream) {
    int rc = check_dims(dims);
    if (rc) return rc;
    if (!params || !chi2_table || !dev || !io || !workspace) return MSCKF_EINVAL;
    if (io->n_samples < 0 || (io->n_samples > 0 && (!dev->imu_samples || !dev->imu_time || !dev->Qc_diag))) return MSCKF_EINVAL;
    const int B = dims->B, F = dims->F, M = dims->M, N = dims->N, ldp = dims->ldp;
    if (chunk_filters <= 0 || chunk_filters > B) chunk_filters = B;
    const int n_chunks = (B + chunk_filters - 1) / chunk_filters;
    if (n_chunks > HS_MAX_CHUNKS) return MSCKF_ECAPACITY;
    HostStepCtx* cx = nullptr;
    cudaError_t e = hs_ctx(&cx, n_chunks);
    if (e != cudaSuccess) return (int)e;
    cudaStream_t sc = (cudaStream_t)stream;
    const size_t FM = (size_t)F * M;
#define HS_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)
    // The per-chunk events order this call's uploads after the previous call's kernels AND downloads of the same chunk;
    // they mean that only while the chunk geometry is unchanged -- otherwise fall back to "after everything enqueued".
    const bool same_chunks = cx->last_chunk_filters == chunk_filters && cx->last_B == B;
    if (!overlap_previous || !same_chunks) {
        HS_TRY(cudaEventRecord(cx->ev_entry, sc));
        HS_TRY(cudaStreamWaitEvent(cx->s_in, cx->ev_entry, 0));
        HS_TRY(cudaStreamWaitEvent(cx->s_in, cx->ev_out_done, 0));   // (a never-recorded event is complete)
    }
    cx->last_chunk_filters = chunk_filters;
    cx->last_B = B;
    for (int k = 0; k < n_chunks; ++k) {
        const int b0 = k * chunk_filters, nb = (b0 + chunk_filters <= B) ? chunk_filters : B - b0;
        // ---- uploads of this chunk (after the previous call's kernels of the same chunk released the buffers) ----
        HS_TRY(cudaStreamWaitEvent(cx->s_in, cx->ev_done[k], 0));
        HS_TRY(cudaStreamWaitEvent(cx->s_in, cx->ev_out[k], 0));   // idp_m / idp_rho are upload targets AND download sources
        const cudaMemcpyKind H2D = cudaMemcpyHostToDevice;
        HS_TRY(hs_copy(dev->feat_len, io->feat_len, F, b0, nb, H2D, cx->s_in));
        HS_TRY(hs_copy(dev->obs_uv, io->obs_uv, FM * 2, b0, nb, H2D, cx->s_in));
        HS_TRY(hs_copy(dev->obs_dir, io->obs_dir, FM * 3, b0, nb, H2D, cx->s_in));
        HS_TRY(hs_copy(dev->obs_conf, io->obs_conf, FM, b0, nb, H2D, cx->s_in));
        HS_TRY(hs_copy(dev->obs_slot, io->obs_slot, FM, b0, nb, H2D, cx->s_in));
        HS_TRY(hs_copy(dev->idp_base, io->idp_base, (size_t)F * 3, b0, nb, H2D, cx->s_in));
        HS_TRY(hs_copy(dev->idp_anchor_slot, io->idp_anchor_slot, F, b0, nb, H2D, cx->s_in));
        HS_TRY(hs_copy(dev->idp_m, io->idp_m, (size_t)F * 3, b0, nb, H2D, cx->s_in));
        HS_TRY(hs_copy(dev->idp_rho, io->idp_rho, F, b0, nb, H2D, cx->s_in));
        HS_TRY(hs_copy(dev->lost_cnt, io->lost_cnt, F, b0, nb, H2D, cx->s_in));
        HS_TRY(hs_copy(dev->tracked_cnt, io->tracked_cnt, F, b0, nb, H2D, cx->s_in));
        if (io->n_samples > 0) HS_TRY(hs_copy(dev->imu_samples, io->imu_samples, (size_t)io->n_samples * 7, b0, nb, H2D, cx->s_in));
        HS_TRY(cudaEventRecord(cx->ev_in[k], cx->s_in));
        // ---- kernels of this chunk on the caller's stream ----
        HS_TRY(cudaStreamWaitEvent(sc, cx->ev_in[k], 0));
        msckf_dims cd = *dims;
        cd.B = nb;
        const size_t o1 = b0;
        if (io->n_samples > 0) {
            rc = msckf_propagate_batched(&cd, io->n_samples, dev->imu_samples + o1 * io->n_samples * 7, dev->imu_time + o1,
                                         dev->Qc_diag + o1 * 4, io->gravity, dev->imu_state + o1 * MSCKF_IMU_STATE_DOUBLES,
                                         (io->use_null_state && dev->imu_null) ? dev->imu_null + o1 * MSCKF_IMU_NULL_DOUBLES : nullptr,
                                         dev->P + o1 * ldp * ldp, dev->n_cams + o1, stream);
            if (rc) return rc;
        }
        rc = msckf_update_batched(&cd, params, chi2_table, chi2_len, dev->P + o1 * ldp * ldp,
                                  dev->imu_state + o1 * MSCKF_IMU_STATE_DOUBLES, dev->cam_R + o1 * N * 9, dev->cam_t + o1 * N * 3,
                                  dev->n_cams + o1, dev->feat_len + o1 * F, dev->obs_uv + o1 * FM * 2, dev->obs_dir + o1 * FM * 3,
                                  dev->obs_conf + o1 * FM, dev->obs_slot + o1 * FM, dev->idp_base + o1 * F * 3,
                                  dev->idp_anchor_slot + o1 * F, dev->idp_m + o1 * F * 3, dev->idp_rho + o1 * F,
                                  dev->lost_cnt + o1 * F, dev->tracked_cnt + o1 * F, dev->feat_status + o1 * F,
                                  dev->gamma + o1 * F, dev->dof + o1 * F, dev->delta_x + o1 * ldp, dev->n_accepted + o1,
                                  dev->n_rejected + o1, workspace, workspace_bytes, stream);
        if (rc) return rc;
        HS_TRY(cudaEventRecord(cx->ev_done[k], sc));
        // ---- downloads of this chunk ----
        HS_TRY(cudaStreamWaitEvent(cx->s_out, cx->ev_done[k], 0));
        const cudaMemcpyKind D2H = cudaMemcpyDeviceToHost;
        HS_TRY(hs_copy(io->imu_state, (const double*)dev->imu_state, MSCKF_IMU_STATE_DOUBLES, b0, nb, D2H, cx->s_out));
        HS_TRY(hs_copy(io->cam_R, (const double*)dev->cam_R, (size_t)N * 9, b0, nb, D2H, cx->s_out));
        HS_TRY(hs_copy(io->cam_t, (const double*)dev->cam_t, (size_t)N * 3, b0, nb, D2H, cx->s_out));
        HS_TRY(hs_copy(io->idp_m_out, (const double*)dev->idp_m, (size_t)F * 3, b0, nb, D2H, cx->s_out));
        HS_TRY(hs_copy(io->idp_rho_out, (const double*)dev->idp_rho, F, b0, nb, D2H, cx->s_out));
        HS_TRY(hs_copy(io->feat_status, (const unsigned char*)dev->feat_status, F, b0, nb, D2H, cx->s_out));
        HS_TRY(hs_copy(io->gamma, (const double*)dev->gamma, F, b0, nb, D2H, cx->s_out));
        HS_TRY(hs_copy(io->delta_x, (const double*)dev->delta_x, ldp, b0, nb, D2H, cx->s_out));
        HS_TRY(hs_copy(io->n_accepted, (const int*)dev->n_accepted, 1, b0, nb, D2H, cx->s_out));
        HS_TRY(hs_copy(io->n_rejected, (const int*)dev->n_rejected, 1, b0, nb, D2H, cx->s_out));
        HS_TRY(cudaEventRecord(cx->ev_out[k], cx->s_out));
    }
    HS_TRY(cudaEventRecord(cx->ev_out_done, cx->s_out));
    HS_TRY(cudaStreamWaitEvent(sc, cx->ev_out_done, 0));
#undef HS_TRY
    return 0;
}
