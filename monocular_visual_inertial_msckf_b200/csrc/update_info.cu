// Rows 9b-10 of SURVEY.md section 8a for MSCKF_COMPRESS_GRAM: the Kalman update of reference src/msckf/MSCKF.py:604-607
// and MSCKF.correct (611-661), driven by the information pair (G, g) = (T_H^T T_H, T_H^T r_n) that gram.cu accumulates.
//
// The reference evaluates  K = P T_H^T (T_H P T_H^T + s^2 I)^-1,  dx = K r_n,  P+ = (I - K T_H) P (I - K T_H)^T + K R_n K^T.
// With the push-through identity  T_H^T (T_H P T_H^T + s^2 I)^-1 = (G P + s^2 I)^-1 T_H^T  both dx and P+ depend on
// (T_H, r_n) only through (G, g).  This kernel evaluates them in whitened coordinates, which keeps every intermediate
// symmetric and P+ positive semi-definite by construction (the property the Joseph form is used for):
//     order the state as [C; I] (the 6N camera errors first: H has no IMU columns, MSCKF.py:538-540)
//     L'  = first c columns of the Cholesky factor of P'          (d x c;  rows C: L,  rows I: U = P_IC L^-T)
//     S_I = P_II - U U^T                                          (Schur complement, left in place by the factorisation)
//     B   = s^2 I + L^T G L  = C_B C_B^T                          (c x c, SPD, eigenvalues >= s^2)
//     W   = L' C_B^-T                                             (d x c)
//     P+  = s^2 W W^T + [0 0; 0 S_I]                              (== Joseph form of the reference, algebraically)
//     dx  = W C_B^-1 L^T g
// About 26 MFLOP per filter at d = 195 against 71 MFLOP for the S-form and 940 MFLOP for the explicit QR it replaces.
// All GEMM-shaped steps run on the FP64 tensor pipe (dense.cuh).  One persistent 256-thread CTA per filter; operands
// live in per-CTA scratch (L2-resident), staged through shared memory.
#include "msckf_common.cuh"
#include "dense.cuh"

namespace {

using dense::NT;
using dense::NB;

__global__ void __launch_bounds__(NT, 2) update_info_kernel(const ApplyArgs a) {
    extern __shared__ double smem_raw[];
    dense::Smem& sm = *reinterpret_cast<dense::Smem*>(smem_raw);
    const int tid = threadIdx.x;
    const int ldp = a.dims.ldp, Ncap = a.dims.N, ccap = 6 * Ncap;
    double* scratch = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;
    double* Pp = scratch;                                     // [ldp][ldp]  P in [C; I] order -> L' -> W (+ S_I)
    double* T1 = Pp + (size_t)ldp * ldp;                      // [ccap][ccap] G L
    double* Bm = T1 + (size_t)ccap * ccap;                    // [ccap][ccap] B -> C_B
    double* Linv = Bm + (size_t)ccap * ccap;                  // [ceil(ccap / NB)][NB * NB]
    double* vec = Linv + (size_t)((ccap + NB - 1) / NB) * NB * NB;   // [ldp] y, [ldp] dx
    const double s2 = a.sigma * a.sigma;

    for (int b = blockIdx.x; b < a.dims.B; b += gridDim.x) {
        double* P = a.P + (size_t)b * ldp * ldp;
        double* dxo = a.delta_x + (size_t)b * ldp;
        if (a.n_accepted[b] <= 0) {   // nothing survived gating: the reference returns before touching the state
            for (int i = tid; i < ldp; i += NT) dxo[i] = 0.0;
            continue;
        }
        const int nc = a.n_cams[b];
        const int c = 6 * nc, d = 15 + c;
        const double* G = a.T_H + (size_t)b * ccap * ccap;
        const double* gv = a.r_n + (size_t)b * ccap;
        double* y = vec;
        double* dx = vec + ldp;

        // ---- P' = P in [C; I] order (lower triangle) ----
        for (int e = tid; e < d * d; e += NT) {
            const int i = e / d, j = e - i * d;
            if (j > i) continue;
            const int si = i < c ? 15 + i : i - c, sj = j < c ? 15 + j : j - c;
            Pp[(size_t)i * ldp + j] = P[(size_t)si * ldp + sj];
        }
        __syncthreads();
        // ---- L' and the Schur complement S_I ----
        dense::cholesky_ll(sm, Pp, ldp, d, c, nullptr);
        // ---- y = L^T g ; make L explicitly lower triangular (the solve below treats L' as a full d x c matrix) ----
        for (int j = tid; j < c; j += NT) {
            double acc = 0.0;
            for (int k = j; k < c; ++k) acc += Pp[(size_t)k * ldp + j] * gv[k];
            y[j] = acc;
        }
        for (int e = tid; e < c * c; e += NT) {
            const int i = e / c, j = e - i * c;
            if (j > i) Pp[(size_t)i * ldp + j] = 0.0;
        }
        __syncthreads();
        // ---- T1 = G L ----
        dense::gemm<true, false>(sm, c, c, c, G, ccap, Pp, ldp, dense::MASK_NONE, dense::MASK_K_GE_X, false,
                                 [=](int i, int j, double v) { T1[(size_t)i * ccap + j] = v; });
        // ---- B = s^2 I + L^T T1 (lower) ----
        dense::gemm<false, false>(sm, c, c, c, Pp, ldp, T1, ccap, dense::MASK_K_GE_X, dense::MASK_NONE, true,
                                  [=](int i, int j, double v) { if (j <= i) Bm[(size_t)i * ccap + j] = v + (i == j ? s2 : 0.0); });
        // ---- B = C_B C_B^T ; y <- C_B^-1 y ; W = L' C_B^-T (in place) ----
        dense::cholesky_ll(sm, Bm, ccap, c, c, Linv);
        dense::trsv_lower(sm, Bm, ccap, c, Linv, y);
        dense::trsm_right_lt_ll(sm, Pp, ldp, d, Bm, ccap, c, Linv);
        // ---- dx = W y (back to the reference's state order) ----
        for (int i = tid; i < ldp; i += NT) {
            double acc = 0.0;
            if (i < d) {
                const int pi = i >= 15 ? i - 15 : c + i;
                for (int k = 0; k < c; ++k) acc += Pp[(size_t)pi * ldp + k] * y[k];
            }
            dx[i] = acc;
            dxo[i] = acc;
        }
        // ---- P+ = s^2 W W^T + [0 0; 0 S_I], written symmetric in the reference's state order ----
        dense::gemm<true, true>(sm, d, d, c, Pp, ldp, Pp, ldp, dense::MASK_NONE, dense::MASK_NONE, true,
                                [=](int i, int j, double v) {
                                    if (j > i) return;
                                    double r = s2 * v;
                                    if (j >= c) r += Pp[(size_t)i * ldp + j];
                                    const int si = i < c ? 15 + i : i - c, sj = j < c ? 15 + j : j - c;
                                    P[(size_t)si * ldp + sj] = r;
                                    P[(size_t)sj * ldp + si] = r;
                                });
        // ---- state correction (MSCKF.py:616-661) ----
        if (tid == 0) {
            double* st = a.imu_state + (size_t)b * MSCKF_IMU_STATE_DOUBLES;
            apply_rotation_correction(st, dx);
            for (int k = 0; k < 3; ++k) {
                st[9 + k] += dx[12 + k];    // position
                st[12 + k] += dx[6 + k];    // velocity
                st[15 + k] += dx[3 + k];    // gyroscope bias
                st[18 + k] += dx[9 + k];    // accelerometer bias
            }
        }
        for (int i = tid; i < nc; i += NT) {
            double* Rc = a.cam_R + ((size_t)b * Ncap + i) * 9;
            double* tc = a.cam_t + ((size_t)b * Ncap + i) * 3;
            apply_rotation_correction(Rc, dx + 15 + 6 * i);
            for (int k = 0; k < 3; ++k) tc[k] += dx[15 + 6 * i + 3 + k];
        }
        __syncthreads();
    }
}

}  // namespace

size_t info_scratch_doubles(const msckf_dims& d) {
    const size_t ccap = 6 * (size_t)d.N, ldp = d.ldp;
    return ldp * ldp + 2 * ccap * ccap + ((ccap + NB - 1) / NB) * NB * NB + 2 * ldp + 64;
}

cudaError_t launch_update_info(const ApplyArgs& a, int n_ctas, cudaStream_t s) {
    static int resident = 0;   // persistent grid = exactly the CTAs that fit on the device at once (no second wave)
    const size_t smem = sizeof(dense::Smem);
    if (resident == 0) {
        cudaError_t e = cudaFuncSetAttribute(update_info_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, update_info_kernel, NT, smem);
        resident = sms * (per_sm > 0 ? per_sm : 1);
    }
    if (n_ctas > resident) n_ctas = resident;
    update_info_kernel<<<n_ctas, NT, smem, s>>>(a);
    msckf_count_launch(1);
    return cudaGetLastError();
}
