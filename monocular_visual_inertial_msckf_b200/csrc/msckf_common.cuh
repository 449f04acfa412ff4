// Shared device/host declarations of the sm_100a MSCKF kernels (internal; the public surface is include/msckf_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include "../../include/msckf_b200.h"

#define MSCKF_MAX_OBS 64          // observation capacity per feature supported by the front-end kernel
#define MSCKF_TSQR_MAX_COLS 256   // 6 N + 1 columns (R | rhs) handled by one TSQR CTA  -> N <= 42
#define MSCKF_FEAT_PENDING 0x04   // internal feat_status CODE (low nibble) between the kernels of the split front end
#define MSCKF_MAX_CAMS 64         // pose capacity of the dense-update kernels (d <= 399)

// Workspace carve-up shared by msckf_*_workspace_bytes and the launches.  All offsets in bytes, 256-aligned.
struct WsLayout {
    size_t fac_A;     // [B][F][M][12]  observability-projected 2x6 Jacobian blocks A_i (MSCKF.py:532-540)
    size_t fac_Q;     // [B][F][M][6]   orthonormal basis Q1 (2M x 3, zero columns beyond rank) of range(H_f)
    size_t fac_rt;    // [B][F][M][2]   projected residual (I - Q1 Q1^T) r
    size_t fac_rank;  // [B][F] int     rank(H_f)
    size_t T_H;       // [B][c][c]      compressed measurement matrix over the camera columns (c = 6 N)
    size_t r_n;       // [B][c]
    size_t nacc;      // [B] int
    size_t nrej;      // [B] int
    size_t scratch;   // [n_cta][scratch_per_cta] dense-update scratch of the persistent update CTAs
    size_t scratch_per_cta;
    int n_update_ctas;
    size_t k_tiles;   // [sub][F][1280] double2: per-feature K tiles of the split front end (front_split.cu); 0 bytes if unused
    size_t pend_list; // [sub][F] int   features waiting for the factorisation kernel
    size_t pend_count;// int
    size_t pend_key;  // [sub][F] int   slot-pattern key of every pending feature (-1: not pending)
    size_t slot_map;  // [sub][F][32] u8 window slot -> observation index of every pending feature (0xFF: none)
    size_t total;
};

WsLayout msckf_ws_layout(const msckf_dims& d, int compress_mode);

struct FrontArgs {
    msckf_dims dims;
    msckf_update_params par;
    const double* chi2_table;
    int chi2_len;
    const double* P;
    const double* cam_R;
    const double* cam_t;
    const int* n_cams;
    const int* feat_len;
    const double* obs_uv;
    const double* obs_dir;
    const double* obs_conf;
    const int* obs_slot;
    const double* idp_base;
    const int* idp_anchor_slot;
    double* idp_m;
    double* idp_rho;
    const int* lost_cnt;
    const int* tracked_cnt;
    unsigned char* feat_status;
    double* gamma;
    int* dof;
    double* fac_A;
    double* fac_Q;
    double* fac_rt;
    int* fac_rank;
};

struct CompressArgs {
    msckf_dims dims;
    int mode;
    const int* n_cams;
    const int* feat_len;
    const int* obs_slot;
    const unsigned char* feat_status;
    const double* fac_A;
    const double* fac_Q;
    const double* fac_rt;
    double* T_H;   // [B][c][c], c = 6 N (capacity)
    double* r_n;   // [B][c]
    int* n_accepted;
    int* n_rejected;
};

struct ApplyArgs {
    msckf_dims dims;
    double sigma;
    double* P;
    double* imu_state;
    double* cam_R;
    double* cam_t;
    const int* n_cams;
    const double* T_H;
    const double* r_n;
    const int* n_accepted;
    double* delta_x;
    double* scratch;
    size_t scratch_per_cta;   // in doubles
};

cudaError_t launch_front(const FrontArgs& a, cudaStream_t s);
cudaError_t launch_compress_qr(const CompressArgs& a, cudaStream_t s);
cudaError_t launch_compress_gram(const CompressArgs& a, cudaStream_t s);
cudaError_t launch_apply(const ApplyArgs& a, int n_ctas, cudaStream_t s);
size_t apply_scratch_doubles(const msckf_dims& d);
int apply_num_ctas(const msckf_dims& d);
cudaError_t launch_propagate(const msckf_dims& d, int n_samples, const double* samples, double* imu_time,
                             const double* Qc_diag, const double* g, double* imu_state, double* imu_null, double* P,
                             const int* n_cams, cudaStream_t s);
cudaError_t launch_augment(const msckf_dims& d, const double* T_I_C, const double* imu_state, double* cam_R,
                           double* cam_t, double* P, int* n_cams, const unsigned char* active, int* overflow,
                           cudaStream_t s);
cudaError_t launch_remove(const msckf_dims& d, const unsigned char* mask, double* cam_R, double* cam_t, double* P,
                          int* n_cams, double* scratch, cudaStream_t s);

void msckf_count_launch(int n);

// ---- small device helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int tri(int r) { return (r * (r + 1)) >> 1; }

// row of packed-lower flat index e (e = tri(r) + c, 0 <= c <= r)
__device__ __forceinline__ int tri_row(int e) {
    int r = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
    while (tri(r) > e) --r;
    while (tri(r + 1) <= e) ++r;
    return r;
}

// R <- polar orthogonal factor of the 3x3 matrix R (== U V^T of its SVD, MSCKF.py:633-634) by Newton iteration
// X <- (X + X^-T)/2, quadratically convergent from the near-orthogonal matrices this path produces.
__device__ inline void polar_orthogonalize3(double* X) {
    for (int it = 0; it < 6; ++it) {
        const double a = X[0], b = X[1], c = X[2], d = X[3], e = X[4], f = X[5], g = X[6], h = X[7], i = X[8];
        const double c00 = e * i - f * h, c01 = f * g - d * i, c02 = d * h - e * g;
        const double c10 = c * h - b * i, c11 = a * i - c * g, c12 = b * g - a * h;
        const double c20 = b * f - c * e, c21 = c * d - a * f, c22 = a * e - b * d;
        const double det = a * c00 + b * c01 + c * c02;
        const double id = 1.0 / det;
        // X^-T = cofactor matrix / det
        const double n0 = 0.5 * (a + c00 * id), n1 = 0.5 * (b + c01 * id), n2 = 0.5 * (c + c02 * id);
        const double n3 = 0.5 * (d + c10 * id), n4 = 0.5 * (e + c11 * id), n5 = 0.5 * (f + c12 * id);
        const double n6 = 0.5 * (g + c20 * id), n7 = 0.5 * (h + c21 * id), n8 = 0.5 * (i + c22 * id);
        const double diff = fabs(n0 - a) + fabs(n1 - b) + fabs(n2 - c) + fabs(n3 - d) + fabs(n4 - e) + fabs(n5 - f) +
                            fabs(n6 - g) + fabs(n7 - h) + fabs(n8 - i);
        X[0] = n0; X[1] = n1; X[2] = n2; X[3] = n3; X[4] = n4; X[5] = n5; X[6] = n6; X[7] = n7; X[8] = n8;
        if (diff < 1e-17) break;
    }
}

// R <- reorthogonalise(R * exp(dtheta)^T) as MSCKF.correct does for the IMU and every camera (MSCKF.py:625-635):
// Rodrigues only above np.isclose's absolute tolerance 1e-8 on |dtheta|.
__device__ inline void apply_rotation_correction(double* R, const double* dth) {
    const double n2 = dth[0] * dth[0] + dth[1] * dth[1] + dth[2] * dth[2];
    const double n = sqrt(n2);
    double E[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (!(fabs(n) <= 1e-8)) {
        const double a = sin(n) / n, bb = (1.0 - cos(n)) / (n * n);
        const double x = dth[0], y = dth[1], z = dth[2];
        // S = skew(dth), S^2 = dth dth^T - |dth|^2 I
        const double S[9] = {0, -z, y, z, 0, -x, -y, x, 0};
        const double S2[9] = {-(y * y + z * z), x * y, x * z, x * y, -(x * x + z * z), y * z, x * z, y * z, -(x * x + y * y)};
#pragma unroll
        for (int k = 0; k < 9; ++k) E[k] = E[k] + a * S[k] + bb * S2[k];
    }
    double Rn[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Rn[r * 3 + c] = R[r * 3 + 0] * E[c * 3 + 0] + R[r * 3 + 1] * E[c * 3 + 1] + R[r * 3 + 2] * E[c * 3 + 2];
    polar_orthogonalize3(Rn);
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = Rn[k];
}
