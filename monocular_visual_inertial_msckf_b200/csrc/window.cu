// "Next rows" of SURVEY.md section 8f kept on the device so the covariance never leaves HBM between calls:
//   augment_kernel  <-> MSCKF.state_augmentation (reference src/msckf/MSCKF.py:250-265)
//   remove_kernel   <-> the covariance/pose part of MSCKF.remove_cameras (MSCKF.py:751-758)
#include "msckf_common.cuh"

namespace {

constexpr int WT = 256;

// P <- [I; J] P [I; J]^T, symmetrised; J (6 x d) has non-zero columns 0:3 and 12:15 only (MSCKF.py:258-261).
__global__ void __launch_bounds__(WT) augment_kernel(msckf_dims dims, const double* __restrict__ imu_state, double* cam_R,
                                                     double* cam_t, double* Pall, int* n_cams,
                                                     const unsigned char* __restrict__ active, int* overflow, double r0, double r1,
                                                     double r2, double r3, double r4, double r5, double r6, double r7,
                                                     double r8, double t0, double t1, double t2) {
    const int b = blockIdx.x, tid = threadIdx.x;
    if (active && !active[b]) return;
    const int n = n_cams[b];
    if (n >= dims.N) {   // window full: nothing is appended; the caller must not go on as if the pose existed
        if (tid == 0 && overflow) atomicAdd(overflow, 1);
        return;
    }
    const int ldp = dims.ldp, d = 15 + 6 * n;
    double* P = Pall + (size_t)b * ldp * ldp;
    const double* st = imu_state + (size_t)b * 21;
    const double Ric[9] = {r0, r1, r2, r3, r4, r5, r6, r7, r8};
    const double tic[3] = {t0, t1, t2};
    __shared__ double J[6][6];     // columns 0:3 then 12:15
    __shared__ double JP[6][512];  // J P, d <= 512
    __shared__ double JPJ[6][6];
    if (tid == 0) {
        double Rc[9], tc[3], rt[3];
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) Rc[r * 3 + c] = st[r * 3] * Ric[c] + st[r * 3 + 1] * Ric[3 + c] + st[r * 3 + 2] * Ric[6 + c];
            rt[r] = st[r * 3] * tic[0] + st[r * 3 + 1] * tic[1] + st[r * 3 + 2] * tic[2];
            tc[r] = rt[r] + st[9 + r];
        }
        for (int k = 0; k < 9; ++k) cam_R[((size_t)b * dims.N + n) * 9 + k] = Rc[k];
        for (int k = 0; k < 3; ++k) cam_t[((size_t)b * dims.N + n) * 3 + k] = tc[k];
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) J[r][c] = 0.0;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) J[r][c] = Ric[c * 3 + r];          // T_I_C.R^T
        J[3][1] = -rt[2]; J[3][2] = rt[1];                                  // skew(R_W_I t_I_C)
        J[4][0] = rt[2];  J[4][2] = -rt[0];
        J[5][0] = -rt[1]; J[5][1] = rt[0];
        J[3][3] = 1.0; J[4][4] = 1.0; J[5][5] = 1.0;
    }
    // thread i: column i of the six rows of P that J touches (rows 0:3 and 12:15) -- the loads do not depend on J, so they are
    // issued before the barrier that publishes it (the previous version walked 6 d elements with six dependent loads each
    // behind thread 0's prologue: 59 % of the kernel's stall samples on those loads)
    double pc[2][6];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = tid + h * WT;
#pragma unroll
        for (int k = 0; k < 6; ++k) pc[h][k] = (i < d) ? P[(size_t)(k < 3 ? k : 9 + k) * ldp + i] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = tid + h * WT;
        if (i < d) {
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) acc += J[a][k] * pc[h][k];
                JP[a][i] = acc;
            }
        }
    }
    __syncthreads();
    if (tid < 36) {
        const int a = tid / 6, c = tid % 6;
        double acc = 0.0;
        for (int k = 0; k < 6; ++k) {
            const int col = k < 3 ? k : 9 + k;
            acc += JP[a][col] * J[c][k];
        }
        JPJ[a][c] = acc;
    }
    __syncthreads();
    for (int e = tid; e < 6 * d; e += WT) {
        const int a = e / d, i = e - a * d;
        P[(size_t)(d + a) * ldp + i] = JP[a][i];
        P[(size_t)i * ldp + d + a] = JP[a][i];
    }
    if (tid < 36) {
        const int a = tid / 6, c = tid % 6;
        P[(size_t)(d + a) * ldp + d + c] = (JPJ[a][c] + JPJ[c][a]) / 2;
    }
    if (tid == 0) n_cams[b] = n + 1;
}

constexpr int RCH = 16;

__global__ void __launch_bounds__(WT) remove_kernel(msckf_dims dims, const unsigned char* __restrict__ mask, double* cam_R,
                                                    double* cam_t, double* Pall, int* n_cams, double* scratch) {
    const int tid = threadIdx.x;
    const int ldp = dims.ldp;
    __shared__ int keep[128];
    __shared__ int s_nkeep;
    extern __shared__ double rows[];            // [RCH][d] staging of the rows being moved
    double* tmp = scratch + (size_t)blockIdx.x * ldp * ldp;
    for (int b = blockIdx.x; b < dims.B; b += gridDim.x) {
        const int n = n_cams[b];
        if (tid == 0) {
            int nk = 0;
            for (int s = 0; s < n; ++s)
                if (!mask[(size_t)b * dims.N + s]) keep[nk++] = s;
            s_nkeep = nk;
        }
        __syncthreads();
        const int nk = s_nkeep;
        if (nk != n) {
            double* P = Pall + (size_t)b * ldp * ldp;
            const int dn = 15 + 6 * nk;
            // In place, one pass: destination rows in increasing order, RCH rows at a time through shared memory.  A
            // destination element (i, j) comes from (si, sj) with si >= i, sj >= j, so a chunk only overwrites rows whose
            // content has already been consumed (by an earlier chunk, or by the loads of this chunk before the barrier).
            for (int i0 = 0; i0 < dn; i0 += RCH) {
                const int nr = min(RCH, dn - i0);
                for (int e = tid; e < nr * dn; e += WT) {
                    const int i = i0 + e / dn, j = e % dn;
                    const int si = i < 15 ? i : 15 + 6 * keep[(i - 15) / 6] + (i - 15) % 6;
                    const int sj = j < 15 ? j : 15 + 6 * keep[(j - 15) / 6] + (j - 15) % 6;
                    rows[e] = P[(size_t)si * ldp + sj];
                }
                __syncthreads();
                for (int e = tid; e < nr * dn; e += WT) P[(size_t)(i0 + e / dn) * ldp + e % dn] = rows[e];
                __syncthreads();
            }
            // poses: compact through the scratch as well
            for (int e = tid; e < n * 9; e += WT) tmp[e] = cam_R[(size_t)b * dims.N * 9 + e];
            for (int e = tid; e < n * 3; e += WT) tmp[n * 9 + e] = cam_t[(size_t)b * dims.N * 3 + e];
            __syncthreads();
            for (int e = tid; e < nk * 9; e += WT) cam_R[(size_t)b * dims.N * 9 + e] = tmp[keep[e / 9] * 9 + e % 9];
            for (int e = tid; e < nk * 3; e += WT) cam_t[(size_t)b * dims.N * 3 + e] = tmp[n * 9 + keep[e / 3] * 3 + e % 3];
            if (tid == 0) n_cams[b] = nk;
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t launch_augment(const msckf_dims& d, const double* T, const double* imu_state, double* cam_R, double* cam_t,
                           double* P, int* n_cams, const unsigned char* active, int* overflow, cudaStream_t s) {
    augment_kernel<<<d.B, WT, 0, s>>>(d, imu_state, cam_R, cam_t, P, n_cams, active, overflow, T[0], T[1], T[2], T[3], T[4], T[5],
                                      T[6], T[7], T[8], T[9], T[10], T[11]);
    msckf_count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_remove(const msckf_dims& d, const unsigned char* mask, double* cam_R, double* cam_t, double* P,
                          int* n_cams, double* scratch, cudaStream_t s) {
    const int ctas = d.B < 148 * 4 ? d.B : 148 * 4;
    const size_t smem = (size_t)RCH * (15 + 6 * d.N) * sizeof(double);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(remove_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    remove_kernel<<<ctas, WT, smem, s>>>(d, mask, cam_R, cam_t, P, n_cams, scratch);
    msckf_count_launch(1);
    return cudaGetLastError();
}
