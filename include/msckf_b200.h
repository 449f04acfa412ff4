/*
 * msckf_b200.h -- C ABI of the B200-native (sm_100a, FP64) batched MSCKF measurement-update hot path.
 *
 * The reference (ValerioSpagnoli/Monocular-Visual-Inertial-MSCKF) is pure Python and has no FFI/plugin layer
 * (SURVEY.md section 8b); its boundary for this path is the method set of class MSCKF in src/msckf/MSCKF.py.
 * Each entry point below names the reference method(s) it replaces (file:line relative to the reference root).
 * The reference-side binding a maintainer would add (a ctypes stub) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the function name ends in _host.  FP64 everywhere.
 *   - Batch-major struct-of-arrays: B independent filters, each with capacity F features, M observations per
 *     feature, N camera poses in the sliding window.  Filter b uses n_cams[b] <= N poses; its error-state
 *     dimension is d_b = 15 + 6 n_cams[b] in the reference's order [dtheta, dbg, dv, dba, dp | (dtheta, dp) per
 *     camera] (MSCKF.py:171, 258-262).  P is [B][ldp][ldp] row-major, ldp >= 15 + 6 N.
 *   - No hidden allocations: the caller passes a workspace obtained from the matching *_workspace_bytes query.
 *   - Every function returns 0 on success, a negative MSCKF_E* code on bad arguments, or a positive cudaError_t
 *     if a launch failed.  Nothing throws.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - Error behaviour mirrors the reference: a filter whose features are all rejected (or that has none) is left
 *     untouched and reports n_accepted = 0 (MSCKF.py:584-585); no exception is raised for degenerate geometry.
 */
#ifndef MSCKF_B200_H
#define MSCKF_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSCKF_B200_ABI_VERSION 10

/* error codes (negative); positive return values are cudaError_t */
#define MSCKF_EINVAL (-1)      /* null pointer / non-positive dimension */
#define MSCKF_ECAPACITY (-2)   /* a dimension exceeds what this build's kernels support */
#define MSCKF_EWORKSPACE (-3)  /* workspace too small */

/* feat_status codes written per feature (rows 3 and 8 of SURVEY.md section 8a) */
#define MSCKF_FEAT_SKIPPED 0       /* not selected: neither lost nor enough parallax (MSCKF.py:479), or empty slot */
#define MSCKF_FEAT_ACCEPTED 1      /* valid and passed the chi-square gate: its rows entered the stacked update */
#define MSCKF_FEAT_GATED_OUT 2     /* valid but rejected by gating_test (MSCKF.py:577-579) */
#define MSCKF_FEAT_LOST_DISCARD 3  /* lost before it was tracked long enough (MSCKF.py:467-469) */
#define MSCKF_FEAT_LOST_FLAG 0x10  /* OR-ed in when the feature is in the reference's `lost_features` dict */
#define MSCKF_FEAT_GAMMA_BOUND 0x20 /* OR-ed in when the gate was decided by the exact bounds below (gate_mode 1):
                                      gamma[] then holds the bound that decided it, not the Mahalanobis distance */

#define MSCKF_FEAT_TRIANGULATED 0x40 /* OR-ed in when the re-triangulated point projected inside the first observing
                                       camera's image, i.e. idp_m / idp_rho were rewritten (MSCKF.py:484-490: the features
                                       whose W_p_hat the reference appends to its debug lists estimated_world_points /
                                       currently_processed_world_points) */

/* how the chi-square gate (MSCKF.gating_test, MSCKF.py:561-568) is evaluated */
#define MSCKF_GATE_EXACT 0     /* gamma = r_o^T (H_o P H_o^T + sigma^2 I)^-1 r_o for every selected feature */
#define MSCKF_GATE_BOUNDS 1    /* same DECISIONS: |r_o|^2 / (sigma^2 + tr(H_x P H_x^T)) <= gamma <= |r_o|^2 / sigma^2 holds
                                  exactly (H_o P H_o^T is positive semi-definite with largest eigenvalue below that trace),
                                  so a feature whose upper bound passes / whose lower bound fails is decided without the
                                  factorisation; every other feature gets the exact test */

/* how the surviving rows are compressed before the Kalman update (row 9) */
#define MSCKF_COMPRESS_QR 0    /* Householder TSQR of the stacked rows: T_H = R, r_n = Q^T r (MSCKF.py:594-598) */
#define MSCKF_COMPRESS_GRAM 1  /* information form: G = H^T H, g = H^T r, solved through (sigma^2 I + G P)^-1 */

typedef struct msckf_dims {
    int B;    /* filters in the batch */
    int F;    /* feature capacity per filter */
    int M;    /* observation capacity per feature */
    int N;    /* camera-pose capacity per filter */
    int ldp;  /* row stride (in doubles) of each P matrix, >= 15 + 6 N */
} msckf_dims;

typedef struct msckf_update_params {
    double K[9];             /* pinhole intrinsics, row-major (MSCKFParameters.K, MSCKF.py:21-23) */
    double Kinv[9];          /* inverse of K as the host computes it (the reference calls np.linalg.inv(K)) */
    int width, height;       /* image bounds test of Camera.project_point (Camera.py:18-26) */
    double sigma_image;      /* MSCKF.py:92 */
    double gravity[3];       /* W_gravity, MSCKF.py:32 */
    int min_lost;            /* min_number_of_frames_to_be_lost (clamped >= 1, MSCKF.py:119) */
    int min_tracked;         /* min_number_of_frames_to_be_tracked (clamped >= 2, MSCKF.py:120) */
    int use_parallax;        /* MSCKF.py:114 */
    double min_parallax_deg; /* MSCKF.py:115 */
    int compress_mode;       /* MSCKF_COMPRESS_* */
    int gate_mode;           /* MSCKF_GATE_* */
} msckf_update_params;

/* imu_state: [B][21] = R_W_I (9, row-major), t (3), v (3), b_g (3), b_a (3)       (IMU.py:13-41)
 * imu_null:  [B][15] = R_null (9), t_null (3), v_null (3)                          (IMU.py:38-39)
 * cam_R:     [B][N][9] row-major R_W_C, cam_t: [B][N][3]; a camera's null pose is its pose (Camera.py:11)       */
#define MSCKF_IMU_STATE_DOUBLES 21
#define MSCKF_IMU_NULL_DOUBLES 15

int msckf_abi_version(void);

/* Largest supported capacities of this build (queried by the host wrappers and the tests). */
int msckf_max_obs_per_feature(void);
int msckf_max_cameras(int compress_mode);

/* ---------------------------------------------------------------------------------------------------------------
 * msckf_update_batched  <->  MSCKF.get_valid_features (MSCKF.py:458-495) + MSCKF.update (570-609)
 *                            + MSCKF.correct (611-661), for B filters at once.
 *
 * Inputs per filter b / feature j / observation i:
 *   feat_len[B][F]            number of observations of the track (0 = empty slot)
 *   obs_uv[B][F][M][2]        keypoints in pixels (Feature.keypoints)
 *   obs_dir[B][F][M][3]       world ray directions frozen at capture (Line.direction, MSCKF.py:403-410)
 *   obs_conf[B][F][M]         Line.confidence (detector score)
 *   obs_slot[B][F][M]         position of the observing camera in the window (MSCKF.py:539); Line.base is that
 *                             camera's LIVE position (aliasing, MSCKF.py:305,410), so it is read from cam_t
 *   idp_anchor_slot[B][F]     window slot whose position InverseDepthPoint.base aliases, or -1: use idp_base
 *   idp_base[B][F][3]         explicit base (anchor camera already deleted)
 *   idp_m[B][F][3], idp_rho[B][F]   inverse-depth parameters; re-triangulated IN PLACE (MSCKF.py:485-488)
 *   lost_cnt, tracked_cnt[B][F]     Feature.lost_for_n_frames / tracked_for_n_frames
 *   chi2_table[chi2_len]      chi2.ppf(0.95, k), k = 0..chi2_len-1 (entry 0 must be NaN), chi2_len >= 2 M + 1
 * In/out: P, imu_state, cam_R, cam_t.   n_cams[B] is read only.
 * Outputs:
 *   feat_status[B][F]         MSCKF_FEAT_* | MSCKF_FEAT_LOST_FLAG
 *   gamma[B][F], dof[B][F]    Mahalanobis distance (or deciding bound, see MSCKF_FEAT_GAMMA_BOUND) and degrees of
 *                             freedom of each valid feature (NaN / 0 else)
 *   delta_x[B][ldp]           the correction that was applied (zeros when nothing was accepted)
 *   n_accepted[B], n_rejected[B]   features stacked / gated out (the reference's discard counter, MSCKF.py:578)
 * ------------------------------------------------------------------------------------------------------------- */
size_t msckf_update_workspace_bytes(const msckf_dims* dims, int compress_mode);

int msckf_update_batched(const msckf_dims* dims, const msckf_update_params* params,
                         const double* chi2_table, int chi2_len,
                         double* P, double* imu_state, double* cam_R, double* cam_t, const int* n_cams,
                         const int* feat_len, const double* obs_uv, const double* obs_dir, const double* obs_conf,
                         const int* obs_slot, const double* idp_base, const int* idp_anchor_slot,
                         double* idp_m, double* idp_rho, const int* lost_cnt, const int* tracked_cnt,
                         unsigned char* feat_status, double* gamma, int* dof, double* delta_x,
                         int* n_accepted, int* n_rejected,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Fine-grained stages of the same call (function-level parity tests; same argument meaning as above).
 *
 * msckf_front_batched      <-> get_valid_features + compute_residual_and_jacobians (MSCKF.py:497-552)
 *                              + project_on_nullspace (554-559) + gating_test (561-568), per feature.
 *   Also leaves the per-feature factors (A blocks, range basis Q1 of H_f, projected residual) in the workspace for
 *   the compression stage.
 * msckf_compress_batched   <-> the vstack + np.linalg.qr block of update (MSCKF.py:587-602):
 *   T_H[B][6N][6N] (upper triangular over the camera columns; the 15 IMU columns of H are identically zero) and
 *   r_n[B][6N].  Must follow msckf_front_batched on the same workspace.
 * msckf_apply_update_batched <-> S, K, delta_x (MSCKF.py:604-607) + correct (611-661).
 * ------------------------------------------------------------------------------------------------------------- */
int msckf_front_batched(const msckf_dims* dims, const msckf_update_params* params,
                        const double* chi2_table, int chi2_len,
                        const double* P, const double* cam_R, const double* cam_t, const int* n_cams,
                        const int* feat_len, const double* obs_uv, const double* obs_dir, const double* obs_conf,
                        const int* obs_slot, const double* idp_base, const int* idp_anchor_slot,
                        double* idp_m, double* idp_rho, const int* lost_cnt, const int* tracked_cnt,
                        unsigned char* feat_status, double* gamma, int* dof,
                        void* workspace, size_t workspace_bytes, void* stream);

int msckf_compress_batched(const msckf_dims* dims, const msckf_update_params* params,
                           const int* n_cams, const int* feat_len, const int* obs_slot,
                           const unsigned char* feat_status, double* T_H, double* r_n,
                           int* n_accepted, int* n_rejected,
                           void* workspace, size_t workspace_bytes, void* stream);

int msckf_apply_update_batched(const msckf_dims* dims, const msckf_update_params* params,
                               double* P, double* imu_state, double* cam_R, double* cam_t, const int* n_cams,
                               const double* T_H, const double* r_n, const int* n_accepted, double* delta_x,
                               void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * msckf_propagate_batched  <->  MSCKF.process_imu (MSCKF.py:160-248) + IMU.integrate (IMU.py:78-100), applied
 * n_samples times per filter in one launch (the samples between two camera frames).
 *
 *   imu_samples[B][n_samples][7]  timestamp, angular velocity (3), linear acceleration (3)  (IMUMeasurement)
 *   imu_time[B]                   in/out: IMU.timestamp
 *   Qc_diag[B][4]                 sigma_g^2, sigma_bg^2, sigma_a^2, sigma_ba^2 (MSCKF.py:99-103)
 *   gravity[3]                    host pointer
 * In/out: imu_state, imu_null (null state <- state after every sample, MSCKF.py:247-248), P rows/cols 0:15.
 * ------------------------------------------------------------------------------------------------------------- */
int msckf_propagate_batched(const msckf_dims* dims, int n_samples, const double* imu_samples, double* imu_time,
                            const double* Qc_diag, const double* gravity_host,
                            double* imu_state, double* imu_null, double* P, const int* n_cams, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Next rows (window surgery kept on the device so P never leaves HBM between calls):
 * msckf_augment_batched  <-> MSCKF.state_augmentation (MSCKF.py:250-265): appends camera pose slot n_cams[b]
 *   (T_W_C = T_W_I * T_I_C), P <- [I; J] P [I; J]^T symmetrised, n_cams[b] += 1.   T_I_C: host R (9) + t (3).
 *   A filter whose window already holds dims->N poses is left untouched and counted in *overflow_count (device int,
 *   incremented; may be NULL): the reference's dict has no capacity, so the caller must size N for the longest window
 *   (max_number_of_camera_states + 1 + the frames that can pass without a valid feature) and treat a non-zero count as
 *   an error.
 * msckf_remove_cameras_batched <-> MSCKF.remove_cameras (MSCKF.py:751-758): deletes the poses flagged in
 *   remove_mask[B][N] (and their covariance rows/columns), compacting slots; n_cams updated.
 * ------------------------------------------------------------------------------------------------------------- */
int msckf_augment_batched(const msckf_dims* dims, const double* T_I_C_host, const double* imu_state,
                          double* cam_R, double* cam_t, double* P, int* n_cams, const unsigned char* active,
                          int* overflow_count, void* stream);

size_t msckf_remove_workspace_bytes(const msckf_dims* dims);
int msckf_remove_cameras_batched(const msckf_dims* dims, const unsigned char* remove_mask,
                                 double* cam_R, double* cam_t, double* P, int* n_cams,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * msckf_select_poorest_cameras_batched  <->  MSCKF.prune_poorest_camera_states, selection part (MSCKF.py:712-733)
 *
 * Per filter: observations per window slot are counted over all tracks; the (up to) `max_select` slots (the reference
 * uses 2) with the fewest observations -- ties in the order in which the cameras first appear while walking the tracks,
 * as the reference's dict + stable sort produce -- are marked in remove_mask[B][N]; feat_touch[B][F] marks the tracks
 * observing a selected camera (the reference's `features_to_process`, which get one last update before the cameras go).
 *
 * msckf_drop_observations_batched  <->  MSCKF.remove_cameras, track part (MSCKF.py:760-770)
 *
 * The observations made from the slots marked in remove_mask leave every track (order kept, in place), the remaining
 * obs_slot values and idp_anchor_slot are renumbered to the window after the deletion, an anchor on a deleted camera
 * becomes -1 with idp_base frozen at that camera's position, feat_len is updated (0 = the reference deletes the track).
 * Call it BEFORE msckf_remove_cameras_batched (it reads cam_t / n_cams of the window before the deletion).
 * ------------------------------------------------------------------------------------------------------------- */
/* msckf_append_observations_batched  <->  MSCKF.add_camera_measurements, track bookkeeping after the matcher
 * (MSCKF.py:400-438; the descriptor matching and the epipolar / homography tests stay on the host side)
 *
 * action[B][F] (u8) per track slot: 0 untouched, 1 lost this frame (lost_cnt += 1), 2 matched: (new_uv, new_conf) is
 * appended with the ray R_newest K^-1 [u, v, 1] and the newest window slot, tracked_cnt += 1, lost_cnt = 0, 3 new track
 * in this (empty) slot: one observation, inverse-depth point anchored on the newest camera, rho = 0.1, tracked_cnt = 1.
 * Kinv_host: 9 doubles on the HOST.  A track already holding dims->M observations is left untouched and counted in
 * *overflow_count (device int, incremented).  With this call a frame costs 25 B of input per track instead of the whole
 * track. */
int msckf_append_observations_batched(const msckf_dims* dims, const double* Kinv_host, const int* n_cams, const double* cam_R,
                                      const double* cam_t, const unsigned char* action, const double* new_uv,
                                      const double* new_conf, int* feat_len, double* obs_uv, double* obs_dir, double* obs_conf,
                                      int* obs_slot, double* idp_base, int* idp_anchor_slot, double* idp_m, double* idp_rho,
                                      int* lost_cnt, int* tracked_cnt, int* overflow_count, void* stream);
/* ---------------------------------------------------------------------------------------------------------------
 * Matching front of add_camera_measurements, per filter and camera frame:
 *
 * msckf_match_descriptors_batched  <->  FeatureExtractor.match (src/msckf/FeatureExtractor.py:62-84) = XFeat.match
 *   (src/accelerated_features/modules/xfeat.py:327-348): cossim = D1 D2^T in FP32, match12 = first row arg-max,
 *   match21 = first row arg-max of D2 D1^T, kept where mutual and (min_cossim <= 0 or the row maximum > min_cossim).
 *   desc_tracks[B][F_tracks][D] (rows 0 .. n_tracks[b]-1 are the filter's tracks in slot order = the reference's dict order;
 *   the running weighted descriptor average of MSCKF.py:439 is supplied by the caller), desc_new[B][F_new][D] (the frame's
 *   detections, rows 0 .. n_new[b]-1); n_tracks / n_new may be NULL (= full capacity).  D in {32, 64, 128}; both sets of
 *   one filter must fit the shared memory of one SM ((F_tracks + F_new) (D + 4) 4 B + indices <= 227 KB), else
 *   MSCKF_ECAPACITY.  Outputs: match12[B][F_tracks] = detection index or -1, match21[B][F_new] = track slot or -1.
 *   With action != NULL the proposal list of MSCKF.py:317-318 / :400-434 is written as well, in the form
 *   msckf_verify_matches_batched and msckf_append_observations_batched consume: action[B][F_tracks] = 2 + (new_uv,
 *   new_conf) taken from kp_new[B][F_new][2] / score_new[B][F_new] for a matched track, 1 for an unmatched one (lost this
 *   frame), 3 in the free slots n_tracks[b] ... for the unmatched detections in ascending detection order (the order of
 *   np.setdiff1d, FeatureExtractor.py:73), 0 elsewhere; n_tracks_out[b] (may be NULL) = slots in use afterwards;
 *   detections that find no free slot are counted in *overflow_count (device int, incremented; may be NULL).
 *
 * msckf_verify_matches_batched  <->  the epipolar / homography walk of MSCKF.add_camera_measurements (MSCKF.py:333-401):
 *   for every track with action == 2 the proposal new_uv is tested against each past observation in order, relative
 *   pose T_C1_C2 = T_W_C1^-1 T_W_Cnewest from the LIVE window poses: baseline < 0.01 -> mean symmetric transfer error
 *   under H = K R K^-1 > homography threshold rejects; otherwise the SIGNED score x_new^T K^-T [t]x R K^-1 x_old >
 *   epipolar threshold rejects (no absolute value, as in the reference); the first failing observation decides.
 *   verdict[B][F] = 0 not proposed, 1 passed, 2 rejected by the homography test, 3 by the epipolar test; a rejected
 *   proposal's action becomes 1 (the reference only bumps lost_for_n_frames, MSCKF.py:402-404); the per-filter counters
 *   n_homography_rejected[B] / n_epipolar_rejected[B] (may be NULL) are incremented like the reference's
 *   number_of_features_discarder_for_homography_test / number_of_features_discarded_for_epipolar_test.
 *   K_host / Kinv_host: 9 doubles each on the HOST.
 * ------------------------------------------------------------------------------------------------------------- */
int msckf_match_descriptors_batched(int B, int F_tracks, int F_new, int D, double min_cossim, const int* n_tracks,
                                    const int* n_new, const float* desc_tracks, const float* desc_new, int* match12,
                                    int* match21, const double* kp_new, const double* score_new, unsigned char* action,
                                    double* new_uv, double* new_conf, int* n_tracks_out, int* overflow_count, void* stream);
int msckf_verify_matches_batched(const msckf_dims* dims, const double* K_host, const double* Kinv_host,
                                 double epipolar_rejection_threshold, double homography_rejection_threshold,
                                 const int* n_cams, const double* cam_R, const double* cam_t, const int* feat_len,
                                 const double* obs_uv, const int* obs_slot, const double* new_uv, unsigned char* action,
                                 unsigned char* verdict, int* n_homography_rejected, int* n_epipolar_rejected, void* stream);
int msckf_select_poorest_cameras_batched(const msckf_dims* dims, int max_select, int prune_above, const int* n_cams,
                                         const int* feat_len, const int* obs_slot, unsigned char* remove_mask,
                                         unsigned char* feat_touch, void* stream);
/* (prune_above: filters whose window holds at most that many cameras select nothing -- MSCKF.feature_callback prunes only
 * when len(cameras) > max_number_of_camera_states, MSCKF.py:157-158; pass -1 to select in every filter) */

/* msckf_unobserved_cameras_batched  <->  MSCKF.get_cameras_without_features (called by remove_features, MSCKF.py:739-749):
 * remove_mask[B][N] marks the window slots no track observes; filters with active[b] == 0 (active may be NULL) mark none. */
int msckf_unobserved_cameras_batched(const msckf_dims* dims, const int* n_cams, const int* feat_len, const int* obs_slot,
                                     const unsigned char* active, unsigned char* remove_mask, void* stream);

/* msckf_compact_tracks_batched  <->  the `del self.features[...]` of MSCKF.remove_features (MSCKF.py:739-741) and of
 * remove_cameras (tracks left without observations, MSCKF.py:768-772): the surviving tracks of every filter are copied, in
 * order, to the front of a SECOND set of track arrays (the caller swaps the two sets); the rest of dst gets feat_len 0.
 * drop_lost = 0: only empty slots (feat_len 0) go.  drop_lost = 1: in filters whose last update had at least one valid
 * feature (feat_status ACCEPTED / GATED_OUT) the tracks flagged MSCKF_FEAT_LOST_FLAG go too (process_features,
 * MSCKF.py:450-456); active_out[b] (may be NULL) = 1 for those filters.  dest_out[B][F] (may be NULL) receives the new
 * slot of every old slot, -1 for a deleted track, so that a caller can keep its own per-track records aligned. */
typedef struct msckf_track_arrays {
    int* feat_len; double* obs_uv; double* obs_dir; double* obs_conf; int* obs_slot; double* idp_base; int* idp_anchor_slot;
    double* idp_m; double* idp_rho; int* lost_cnt; int* tracked_cnt;
} msckf_track_arrays;
int msckf_compact_tracks_batched(const msckf_dims* dims, int drop_lost, const unsigned char* feat_status,
                                 const msckf_track_arrays* src, const msckf_track_arrays* dst, unsigned char* active_out,
                                 int* dest_out, void* stream);
/* The same compaction IN PLACE: `tracks` is compacted where it lies, `scratch` (a second set of track arrays of the same
 * shapes) only stages the tracks that move.  A filter in which nothing moves costs one scan of its F lengths, so the frame
 * loop pays for deletions only where they happen; nothing is swapped by the caller. */
int msckf_compact_tracks_inplace_batched(const msckf_dims* dims, int drop_lost, const unsigned char* feat_status,
                                         const msckf_track_arrays* tracks, const msckf_track_arrays* scratch,
                                         unsigned char* active_out, int* dest_out, void* stream);
int msckf_drop_observations_batched(const msckf_dims* dims, const unsigned char* remove_mask, const int* n_cams,
                                    const double* cam_t, int* feat_len, double* obs_uv, double* obs_dir, double* obs_conf,
                                    int* obs_slot, double* idp_base, int* idp_anchor_slot, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * msckf_filter_step_host  <->  one camera-frame step of B filters driven from HOST buffers: the n_samples
 * MSCKF.process_imu calls since the last frame (MSCKF.py:160-248) followed by get_valid_features + update + correct
 * (MSCKF.py:458-495, 570-661).  This is the call a host-side owner of the tracks (the reference keeps them in Python
 * objects) makes once per frame.
 *
 * The batch is cut into chunks of `chunk_filters` filters.  For every chunk the library copies the host inputs to the
 * caller's device buffers on an internal copy stream, runs propagate + update of that chunk on `stream`, and copies the
 * requested results back on a second internal stream, so the PCIe transfers of chunk k+1 / k-1 overlap the kernels of
 * chunk k.  On return everything is enqueued; `stream` has been made to wait for the last device-to-host copy, so
 * synchronising `stream` makes all host outputs valid.
 *
 *   dev   device buffers of the whole batch (layouts as in msckf_update_batched / msckf_propagate_batched)
 *   io    host buffers in the same layouts.  A NULL input pointer means "already resident in dev, do not copy";
 *         a NULL output pointer means "not wanted".  Pinned memory is needed for the copies to be asynchronous.
 *   chunk_filters  <= 0: one chunk.  A multiple of the resident update-CTA count (2 x SM count) avoids tail waves.
 *   overlap_previous  non-zero: the uploads of this call may start while work enqueued on `stream` BEFORE the call is
 *         still running, ordered only after the previous msckf_filter_step_host call's kernels of the same chunk (the
 *         caller promises nothing else on `stream` still reads the device track buffers).  Zero: uploads start after
 *         everything already enqueued on `stream`.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct msckf_device_buffers {
    double* P; double* imu_state; double* imu_null; double* imu_time; const double* Qc_diag;
    double* cam_R; double* cam_t; const int* n_cams;
    int* feat_len; double* obs_uv; double* obs_dir; double* obs_conf; int* obs_slot;
    double* idp_base; int* idp_anchor_slot; double* idp_m; double* idp_rho; int* lost_cnt; int* tracked_cnt;
    unsigned char* feat_status; double* gamma; int* dof; double* delta_x; int* n_accepted; int* n_rejected;
    double* imu_samples;     /* [B][n_samples][7] device staging for the IMU samples (may be NULL if n_samples == 0) */
} msckf_device_buffers;

typedef struct msckf_host_step {
    /* inputs (host) */
    const int* feat_len; const double* obs_uv; const double* obs_dir; const double* obs_conf; const int* obs_slot;
    const double* idp_base; const int* idp_anchor_slot; const double* idp_m; const double* idp_rho;
    const int* lost_cnt; const int* tracked_cnt;
    const double* imu_samples; int n_samples;     /* n_samples == 0: no propagation */
    int use_null_state;                           /* 0: null state = state on entry (see msckf_propagate_batched) */
    double gravity[3];
    /* outputs (host) */
    double* imu_state; double* cam_R; double* cam_t; double* idp_m_out; double* idp_rho_out;
    unsigned char* feat_status; double* gamma; double* delta_x; int* n_accepted; int* n_rejected;
} msckf_host_step;

int msckf_filter_step_host(const msckf_dims* dims, const msckf_update_params* params,
                           const double* chi2_table, int chi2_len,
                           const msckf_device_buffers* dev, const msckf_host_step* io,
                           int chunk_filters, int overlap_previous,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Track descriptors of the matching front with the reference's list semantics.
 * obs_desc[B][F][R][D] float32 holds Feature.descriptors, one row per CAMERA of the window (R >= N rows, <= 64): the descriptor
 * a track observed from the camera in window slot s sits in row cam_row[B][N][s] -- a camera keeps its row for its whole life
 * (the "camera ring"), so deleting cameras moves no descriptor (rewriting every list cost 2 ms per pruning frame of 4096
 * filters); the list of a track is its observations in order, obs_slot[i] -> cam_row -> row.  obs_conf IS Feature.scores
 * (Line.confidence is the same number, MSCKF.py:408-410).  D a multiple of 4, <= 128.
 *
 * msckf_score_filter_batched  <->  MSCKF.add_camera_measurements' score filter (MSCKF.py:279-284): per filter the detections
 *   with score >= 0.5 mean(score) are kept, in order: (scores, kp[.][2], desc[.][D]) of n_in[b] detections (NULL = F_new) ->
 *   compacted copies + n_out[b].  Source and destination must not alias.
 * msckf_append_descriptors_batched  <->  feature.descriptors.append(...) of MSCKF.py:407 (accepted match, action 2: the
 *   detection match12[slot]) and :425 (new track, action 3: the r-th unmatched detection in ascending order feeds slot
 *   n_tracks[b] + r, as msckf_match_descriptors_batched assigns them).  Call it with the FINAL actions (after
 *   msckf_verify_matches_batched) and BEFORE msckf_append_observations_batched: feat_len must still hold the lengths before
 *   the frame.
 * msckf_drop_descriptors_batched  <->  del feature.descriptors[i] / del feature.scores[i] of MSCKF.remove_cameras
 *   (MSCKF.py:764-765): the rows of the deleted cameras leave the running sums (they simply stop being referenced in
 *   obs_desc).  Same remove_mask as msckf_drop_observations_batched, called BEFORE it and before
 *   msckf_descriptor_rows_remove_batched (obs_slot / feat_len / cam_row still describe the window before the deletion).
 * msckf_compact_descriptors_batched  <->  the track deletions (MSCKF.py:739-741, 768-772) for the descriptor lists: every
 *   track moves to the slot dest[B][F] that msckf_compact_tracks_batched returned (-1: deleted); feat_len_dst = the lengths
 *   AFTER the compaction.  src and dst are two different buffers (the caller swaps them like the track arrays).
 * msckf_track_descriptors_batched  <->  np.average(feature.descriptors, axis=0, weights=feature.scores) for every track
 *   (MSCKF.py:436-441): desc_tracks[B][F][D], the rows msckf_match_descriptors_batched matches against (accumulated in FP64,
 *   NOT re-normalised, zero rows for empty slots).  The reference rebuilds these rows at the END of add_camera_measurements
 *   (after the frame's appends, before process_features / pruning) and matches the NEXT frame against them, so a pruned
 *   camera's descriptors leave a track's mean one frame after they left its list; a filter that returned early from
 *   add_camera_measurements (no detections / no match, MSCKF.py:286, 320) keeps its old rows: keep_old[B] (may be NULL)
 *   marks those filters.  Rows of deleted tracks follow msckf_compact_descriptors_batched (called with dims->M = 1).
 * Running sums (optional, NULL = not kept): desc_sum[B][F][D + 2] float64 = (sum_i w_i d_i | sum_i w_i | pad) per track, w_i =
 *   the observation's score.  msckf_append_descriptors_batched adds the new observation (new_conf[B][F] = its score per track
 *   slot, the array msckf_append_observations_batched stores as obs_conf), msckf_drop_descriptors_batched subtracts the rows
 *   that leave (obs_conf = the scores of the lists), the in-place compaction moves the records with their tracks (dims->M = 1,
 *   D' = 2 (D + 2) floats), msckf_track_descriptors_batched(desc_sum_out) rebuilds them from the lists, and
 *   msckf_track_descriptors_from_sums_batched produces the same rows as msckf_track_descriptors_batched from D + 2 doubles per
 *   track instead of its whole list (the two differ by FP64 summation order only, i.e. below the float32 rounding of the rows).
 * ------------------------------------------------------------------------------------------------------------- */
int msckf_score_filter_batched(int B, int F_new, int D, const int* n_in, const double* scores, const double* kp,
                               const float* desc, double* scores_out, double* kp_out, float* desc_out, int* n_out,
                               void* stream);
int msckf_append_descriptors_batched(const msckf_dims* dims, int D, int F_new, const unsigned char* action,
                                     const int* match12, const int* match21, const int* n_tracks, const int* n_new,
                                     const float* desc_new, const int* feat_len, float* obs_desc, const double* new_conf,
                                     double* desc_sum, int R, const int* n_cams, const int* cam_row, void* stream);
int msckf_drop_descriptors_batched(const msckf_dims* dims, int D, const unsigned char* remove_mask, const int* feat_len,
                                   const int* obs_slot, const float* obs_desc, const double* obs_conf, double* desc_sum, int R,
                                   const int* cam_row, void* stream);
int msckf_compact_descriptors_batched(const msckf_dims* dims, int D, const int* dest, const int* feat_len_dst,
                                      const float* src, float* dst, void* stream);
/* in-place variant for msckf_compact_tracks_inplace_batched: desc is compacted where it lies, scratch stages the rows that
 * move; zero_deleted != 0 clears the rows behind the survivors (the rows of last_camera_measurement). */
int msckf_compact_descriptors_inplace_batched(const msckf_dims* dims, int D, const int* dest, const int* feat_len_new,
                                              float* desc, float* scratch, int zero_deleted, void* stream);
int msckf_track_descriptors_batched(const msckf_dims* dims, int D, const int* feat_len, const double* obs_conf,
                                    const float* obs_desc, const unsigned char* keep_old, float* desc_tracks,
                                    double* desc_sum_out, int R, const int* obs_slot, const int* cam_row, void* stream);
int msckf_track_descriptors_from_sums_batched(const msckf_dims* dims, int D, const int* feat_len, const double* desc_sum,
                                              const unsigned char* keep_old, float* desc_tracks, void* stream);
/* the camera ring: msckf_descriptor_rows_assign_batched gives the camera msckf_augment_batched is ABOUT to append (same
 * `active`, call it first) the lowest free row; msckf_descriptor_rows_remove_batched compacts cam_row like
 * msckf_remove_cameras_batched compacts the window (same mask, call it first: n_cams still counts the deleted cameras). */
int msckf_descriptor_rows_assign_batched(const msckf_dims* dims, int R, const int* n_cams, const unsigned char* active,
                                         int* cam_row, void* stream);
int msckf_descriptor_rows_remove_batched(const msckf_dims* dims, const int* n_cams, const unsigned char* remove_mask,
                                         int* cam_row, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Data synthesis and metrics on the device: a Monte-Carlo run builds its streams and reduces
 * its errors where the filters live; the final gather carries statistics only.
 *
 * msckf_synth_imu_batched  <->  IMU synthesis from ground-truth poses:
 *   PhotorealisticGenerator.process_data, dataset/tools/dataset_generators/photorealistic_generator.py:145-216 (interp = 10:
 *   every key interval is sampled at np.linspace(0, 1, 10) -- linear position, Slerp rotation -- and differentiated:
 *   v = dp/dt, a = R_prev^T (dv/dt + g), w = (2/dt) vec(q_prev^* q_cur) with the sign of q_cur fixed; noisy = gt + white
 *   noise + random-walk INCREMENT) and MeasurementsGenerator.generate, synthetic_generator.py:145-186 (interp = 1: the poses
 *   are at IMU rate; sign_fix = 0; accumulate_bias = 1: noisy = gt + white noise + ACCUMULATED walk).
 *   key_poses[B][K][12] (R row-major, t), key_times[B][K]; T = (interp > 1 ? interp - 1 : 1) (K - 1) samples.
 *   noise[B][T][12] (may be NULL): acc white (3), acc walk increment (3), gyro white (3), gyro walk increment (3), already
 *   scaled; with noise == NULL and sigmas[B][4] (acc, gyro, acc walk, gyro walk) != NULL the draws come from a
 *   counter-based generator (Philox4x32-10 + Box-Muller, keyed by seed / filter / sample); both NULL: noisy = gt.
 *   Outputs: poses_out[B][T+1][12] / times_out[B][T+1] (may be NULL), imu_gt / imu_noisy[B][T+1][7] = (timestamp, w, a) in
 *   the sample layout of msckf_propagate_batched; row 0 is the reference's zero row at the first timestamp.
 *
 * msckf_project_landmarks_batched  <->  the camera-measurement loop of MeasurementsGenerator.generate,
 *   synthetic_generator.py:189-223 (Camera.W2Ci / project_point, src/msckf/Camera.py:13-52): body_poses[B][n_frames][12],
 *   T_I_C_host (R 9 + t 3) and K_host (9) on the HOST, landmarks[B][L][3].  Per (filter, frame) the landmarks in front of
 *   the camera and inside the image are listed in id order: kp_out[B][n_frames][L][2] (pixel + noise), id_out[B][n_frames][L],
 *   count_out[B][n_frames], score_out (may be NULL) = 1 / (1 + 2 sigma_px^2).  noise_uv[B][n_frames][L][2] indexed by
 *   landmark id (may be NULL: drawn on the device with sigma_px, or none when sigma_px == 0).
 *
 * msckf_metrics_step_batched / msckf_metrics_finalize_batched  <->  the error block of the reference driver, main.py:312-391
 *   and 436-437, and its summary, main.py:440-445.  One call per step with the step's ground-truth body pose gt_pose[B][12];
 *   the estimate is read from imu_state / P.  metrics_state[B][msckf_metrics_state_doubles()] is the running state (zero it
 *   before the first step): previous poses, the two 10-sample deques of the smoothing (:352-361), running sums.
 *   record_out[B][14] (may be NULL) = rte_smoothed, roe_smoothed, ATE x y z, AOE roll pitch yaw (intrinsic XYZ Euler angles of
 *   the absolute rotation error, R2euler), the six 3-sigma bounds (:337-345).  summary_out[B][8] = rmse_position,
 *   rmse_orientation, mean_rte, std_rte, mean_roe, std_roe (unrounded; the reference rounds to 8 decimals), rmse of |ATE|, n.
 * ------------------------------------------------------------------------------------------------------------- */
int msckf_synth_imu_batched(int B, int K, int interp, const double* key_poses, const double* key_times,
                            const double* gravity_host, const double* noise, const double* sigmas, unsigned long long seed,
                            int accumulate_bias, int sign_fix, double* poses_out, double* times_out, double* imu_gt,
                            double* imu_noisy, void* stream);
int msckf_project_landmarks_batched(int B, int L, int n_frames, const double* body_poses, const double* T_I_C_host,
                                    const double* K_host, int width, int height, const double* landmarks,
                                    const double* noise_uv, double sigma_px, unsigned long long seed, double* kp_out,
                                    int* id_out, int* count_out, double* score_out, void* stream);
int msckf_metrics_state_doubles(void);
int msckf_metrics_step_batched(const msckf_dims* dims, const double* gt_pose, const double* imu_state, const double* P,
                               double* metrics_state, double* record_out, void* stream);
int msckf_metrics_finalize_batched(int B, const double* metrics_state, double* summary_out, void* stream);

/* Diagnostic: launches n_ctas x 256 threads, every warp running 8 independent FP64 tensor-instruction (DMMA.8x8x4) chains
 * of length iters (the measured FP64-pipe roofline denominator; MEASURED_PEAKS.json has no FP64 figure).
 * *flops_out_host receives the FLOP count. */
int msckf_probe_fp64(int n_ctas, int iters, double* sink, double* flops_out_host, void* stream);

/* Number of kernels launched by this library since load (the benchmark reports it as gpu_launches). */
long long msckf_launch_count(void);

/* Optional per-kernel timing of the split front end (measurement aid for bench.py; not part of the reference's surface):
 * with timing on, msckf_front_batched / msckf_update_batched record CUDA events on the caller's stream around the geometry,
 * the assembly and the factorisation kernel of every sub-batch; msckf_front_stage_ms waits for the last recorded call and
 * returns the summed durations in milliseconds.  Returns the number of sub-batches timed (0: timing off, or the general
 * front-end kernel ran instead). */
int msckf_set_stage_timing(int on);
int msckf_front_stage_ms(double* geo_ms, double* asm_ms, double* chol_ms);

#ifdef __cplusplus
}
#endif
#endif /* MSCKF_B200_H */
