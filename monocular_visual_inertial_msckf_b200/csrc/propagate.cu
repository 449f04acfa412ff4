// Row 11 of SURVEY.md section 8a: MSCKF.process_imu (reference src/msckf/MSCKF.py:160-248) + IMU.integrate
// (src/msckf/IMU.py:78-100), n_samples per launch, one CTA per filter.
//
// The 15x15 IMU block and the composed transition matrix stay in shared memory across all the samples between two camera
// frames; HBM is touched once: read P[0:15, :] (15 d doubles), write P[0:15, :] and its transpose P[:, 0:15].  The camera block of P is not touched: the reference's global (P + P^T)/2 (MSCKF.py:244) is the
// identity on it because every producer of P leaves it exactly symmetric.
#include "msckf_common.cuh"

namespace {

constexpr int PT = 128;

// ---------------------------------------------------------------------------------------------------------------------
// The arithmetic is organised so that nothing but a 3x3 recurrence is serial.
//   A  thread k: sample k's dt, bias-corrected w / a and its rotation increment exp(w dt) (the sin / cos / sqrt)
//   B  thread 0: the state recurrence R <- R dR, p, v over the chunk (the only sequential part: ~60 flops per sample)
//   C  warp w, samples w, w + 4, ...: F dt, Phi = I + F + F^2/2 + F^3/6, the observability fix-up and Q_k, all from the
//      stored per-sample states -- independent across samples
//   D  all threads, sample by sample: P_II <- Phi P_II Phi^T + Q (symmetrised), Phi_tot <- Phi Phi_tot
//   E  once per launch: P_IC <- Phi_tot P_IC straight from / to global memory (one thread per column; the reference's
//      per-sample products Phi_k (... (Phi_1 P_IC)) are the same matrix up to rounding), the strip and its transpose written once
// ---------------------------------------------------------------------------------------------------------------------
constexpr int PCH = 12;   // samples per chunk (per-sample Phi / Q live in shared memory)

struct SampleState {      // what phase C needs of sample k
    double dt, w[3], a[3];
    double R[9], p[3], v[3];          // state AFTER integrating the sample
    double Rn[9], tn[3], vn[3];       // null state = state before it (imu_null for the first sample of the launch)
    double dR[9];
};

// element (r, c) of skew(v) = [[0, -v2, v1], [v2, 0, -v0], [-v1, v0, 0]]
__device__ __forceinline__ double skew_elem(const double* v, int r, int c) {
    const int d = (c - r + 3) % 3;
    return d == 0 ? 0.0 : (d == 1 ? -v[(r + 2) % 3] : v[(r + 1) % 3]);
}

__device__ __forceinline__ void warp_mm15(double* C, const double* A, const double* B, int lane) {   // C = A B, 15x15
    for (int e = lane; e < 225; e += 32) {
        const int i = e / 15, j = e - 15 * i;
        double acc = 0.0;
#pragma unroll 5
        for (int l = 0; l < 15; ++l) acc += A[i * 15 + l] * B[l * 15 + j];
        C[e] = acc;
    }
}

__global__ void __launch_bounds__(PT, 4) propagate_kernel_v2(msckf_dims dims, int n_samples, const double* __restrict__ samples,
                                                          double* imu_time, const double* __restrict__ Qc_diag, double g0,
                                                          double g1, double g2, double* imu_state, double* imu_null,
                                                          double* Pall, const int* __restrict__ n_cams, int chunk_cap) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ldp = dims.ldp;
    const int c = 6 * n_cams[b];
    double* P = Pall + (size_t)b * ldp * ldp;
    const double g[3] = {g0, g1, g2};

    extern __shared__ double sm[];
    double* sPhiK = sm;                          // [chunk_cap][225]
    double* sQK = sPhiK + chunk_cap * 225;       // [chunk_cap][225]
    double* sW = sQK + chunk_cap * 225;          // [4 warps][2][225] scratch of phase C
    double* sPii = sW + 4 * 450;
    double* sT = sPii + 225;
    double* sTot = sT + 225;                     // Phi_tot (double-buffered)
    double* sTot2 = sTot + 225;
    double* st = sTot2 + 225;                    // [21] running state  R t v bg ba
    double* sn = st + 21;                        // [15] running null state
    double* sc = sn + 15;                        // [1] time
    SampleState* ss = reinterpret_cast<SampleState*>(sc + 1);   // [chunk_cap]

    for (int e = tid; e < 225; e += PT) {
        sPii[e] = P[(size_t)(e / 15) * ldp + (e % 15)];
        sTot[e] = ((e / 15) == (e % 15)) ? 1.0 : 0.0;
    }
    if (tid < 21) st[tid] = imu_state[(size_t)b * 21 + tid];
    if (tid < 15) sn[tid] = imu_null ? imu_null[(size_t)b * 15 + tid] : imu_state[(size_t)b * 21 + tid];
    if (tid == 0) sc[0] = imu_time[b];
    __syncthreads();

    for (int k0 = 0; k0 < n_samples; k0 += chunk_cap) {
        const int nk = min(chunk_cap, n_samples - k0);
        // ---- A: per-sample quantities that do not depend on the state ----
        if (tid < nk) {
            const double* z = samples + ((size_t)b * n_samples + k0 + tid) * 7;
            const double tprev = (tid == 0) ? sc[0] : z[-7];
            SampleState& q = ss[tid];
            const double dt = z[0] - tprev;
            q.dt = dt;
            double w[3], a[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                w[k] = z[1 + k] - st[15 + k];
                a[k] = z[4 + k] - st[18 + k];
                q.w[k] = w[k];
                q.a[k] = a[k];
            }
            // IMU.integrate rotation increment (IMU.py:78-100)
            const double wn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
            const double theta = wn * dt;
            double dR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            if (theta > 0.0) {
                const double x = w[0] / wn, y = w[1] / wn, zz = w[2] / wn;
                const double K[9] = {0, -zz, y, zz, 0, -x, -y, x, 0};
                double K2[9];
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) K2[r * 3 + cc] = K[r * 3] * K[cc] + K[r * 3 + 1] * K[3 + cc] + K[r * 3 + 2] * K[6 + cc];
                const double sn_ = sin(theta), cs = 1.0 - cos(theta);
#pragma unroll
                for (int k = 0; k < 9; ++k) dR[k] = dR[k] + sn_ * K[k] + cs * K2[k];
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) q.dR[k] = dR[k];
        }
        __syncthreads();
        // ---- B: the state recurrence ----
        if (tid == 0) {
            double* R = st;
            double* p = st + 9;
            double* v = st + 12;
            for (int k = 0; k < nk; ++k) {
                SampleState& q = ss[k];
                for (int e = 0; e < 9; ++e) q.Rn[e] = sn[e];
                for (int e = 0; e < 3; ++e) { q.tn[e] = sn[9 + e]; q.vn[e] = sn[12 + e]; }
                const double dt = q.dt;
                double Rn[9];
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) Rn[r * 3 + cc] = R[r * 3] * q.dR[cc] + R[r * 3 + 1] * q.dR[3 + cc] + R[r * 3 + 2] * q.dR[6 + cc];
                double aw[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) aw[r] = (R[r * 3] * q.a[0] + R[r * 3 + 1] * q.a[1] + R[r * 3 + 2] * q.a[2]) - g[r];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    p[r] = p[r] + v[r] * dt + 0.5 * aw[r] * (dt * dt);
                    v[r] = v[r] + aw[r] * dt;
                }
#pragma unroll
                for (int e = 0; e < 9; ++e) { R[e] = Rn[e]; q.R[e] = Rn[e]; }
                for (int e = 0; e < 3; ++e) { q.p[e] = p[e]; q.v[e] = v[e]; }
                for (int e = 0; e < 15; ++e) sn[e] = st[e];          // null <- current (MSCKF.py:247-248)
            }
            sc[0] = samples[((size_t)b * n_samples + k0 + nk - 1) * 7];
        }
        __syncthreads();
        // ---- C: Phi_k and Q_k, one warp per sample ----
        for (int k = warp; k < nk; k += 4) {
            const SampleState& q = ss[k];
            double* F = sW + warp * 450;
            double* F2 = F + 225;
            double* Phi = sPhiK + k * 225;
            double* Qk = sQK + k * 225;
            const double dt = q.dt;
            for (int e = lane; e < 225; e += 32) F[e] = 0.0;
            __syncwarp();
            if (lane < 9) {   // F dt (MSCKF.py:179-192, 215) with the post-integration rotation
                const int r = lane / 3, cc = lane - 3 * r;
                F[r * 15 + cc] = (-skew_elem(q.w, r, cc)) * dt;
                F[r * 15 + 3 + cc] = (r == cc ? -1.0 : -0.0) * dt;
                const double ra = q.R[r * 3] * skew_elem(q.a, 0, cc) + q.R[r * 3 + 1] * skew_elem(q.a, 1, cc) +
                                  q.R[r * 3 + 2] * skew_elem(q.a, 2, cc);
                F[(6 + r) * 15 + cc] = (-ra) * dt;
                F[(6 + r) * 15 + 9 + cc] = (-q.R[r * 3 + cc]) * dt;
                F[(12 + r) * 15 + 6 + cc] = (r == cc ? 1.0 : 0.0) * dt;
            }
            __syncwarp();
            warp_mm15(F2, F, F, lane);
            __syncwarp();
            for (int e = lane; e < 225; e += 32) {
                const int i = e / 15, j = e - 15 * i;
                double f3 = 0.0;
#pragma unroll 5
                for (int l = 0; l < 15; ++l) f3 += F2[i * 15 + l] * F[l * 15 + j];
                const double eye = (i == j) ? 1.0 : 0.0;
                Phi[e] = ((eye + F[e]) + 0.5 * F2[e]) + (1.0 / 6.0) * f3;
            }
            __syncwarp();
            // observability fix-up (MSCKF.py:221-233); null state = state before this integration
            if (lane < 9) {
                const int r = lane / 3, cc = lane - 3 * r;
                Phi[r * 15 + cc] = q.R[r * 3] * q.Rn[cc * 3] + q.R[r * 3 + 1] * q.Rn[cc * 3 + 1] + q.R[r * 3 + 2] * q.Rn[cc * 3 + 2];
            }
            if (lane >= 16 && lane < 19) {
                const int r = lane - 16;
                double u[3], s[3];
#pragma unroll
                for (int e = 0; e < 3; ++e) u[e] = q.Rn[e * 3] * g[0] + q.Rn[e * 3 + 1] * g[1] + q.Rn[e * 3 + 2] * g[2];
                const double uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
#pragma unroll
                for (int e = 0; e < 3; ++e) s[e] = u[e] / uu;
                const double d1[3] = {q.vn[0] - q.v[0], q.vn[1] - q.v[1], q.vn[2] - q.v[2]};
                const double d2[3] = {dt * q.vn[0] + q.tn[0] - q.p[0], dt * q.vn[1] + q.tn[1] - q.p[1], dt * q.vn[2] + q.tn[2] - q.p[2]};
                const double w1[3] = {d1[1] * g[2] - d1[2] * g[1], d1[2] * g[0] - d1[0] * g[2], d1[0] * g[1] - d1[1] * g[0]};
                const double w2[3] = {d2[1] * g[2] - d2[2] * g[1], d2[2] * g[0] - d2[0] * g[2], d2[0] * g[1] - d2[1] * g[0]};
                double* av = Phi + (6 + r) * 15;
                double* ap = Phi + (12 + r) * 15;
                const double ev = (av[0] * u[0] + av[1] * u[1] + av[2] * u[2]) - w1[r];
                const double ep = (ap[0] * u[0] + ap[1] * u[1] + ap[2] * u[2]) - w2[r];
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    av[cc] = av[cc] - ev * s[cc];
                    ap[cc] = ap[cc] - ep * s[cc];
                }
            }
            __syncwarp();
            // Q = Phi G Qc G^T Phi^T dt, evaluated left to right like the reference (MSCKF.py:237)
            {
                const double* R = q.R;
                const double* qc = Qc_diag + (size_t)b * 4;
                double* T1 = F;      // 15 x 12
                double* T2 = F2;     // 15 x 15
                for (int e = lane; e < 15 * 12; e += 32) {
                    const int i = e / 12, j = e % 12;
                    const int blk = j / 3, jj = j % 3;
                    double v;
                    if (blk == 0) v = -Phi[i * 15 + jj];
                    else if (blk == 1) v = Phi[i * 15 + 3 + jj];
                    else if (blk == 2) v = Phi[i * 15 + 6] * (-R[jj]) + Phi[i * 15 + 7] * (-R[3 + jj]) + Phi[i * 15 + 8] * (-R[6 + jj]);
                    else v = Phi[i * 15 + 9 + jj];
                    T1[i * 12 + j] = v * qc[blk];
                }
                __syncwarp();
                for (int e = lane; e < 225; e += 32) {
                    const int i = e / 15, j = e % 15;
                    const int blk = j / 3, jj = j % 3;
                    double v;
                    if (blk == 0) v = -T1[i * 12 + jj];
                    else if (blk == 1) v = T1[i * 12 + 3 + jj];
                    else if (blk == 2) v = T1[i * 12 + 6] * (-R[jj * 3]) + T1[i * 12 + 7] * (-R[jj * 3 + 1]) + T1[i * 12 + 8] * (-R[jj * 3 + 2]);
                    else if (blk == 3) v = T1[i * 12 + 9 + jj];
                    else v = 0.0;
                    T2[e] = v;
                }
                __syncwarp();
                for (int e = lane; e < 225; e += 32) {
                    const int i = e / 15, j = e % 15;
                    double acc = 0.0;
#pragma unroll 5
                    for (int l = 0; l < 15; ++l) acc += T2[i * 15 + l] * Phi[j * 15 + l];
                    Qk[e] = acc * dt;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // ---- D: the 15x15 recurrences, sample by sample ----
        for (int k = 0; k < nk; ++k) {
            const double* Phi = sPhiK + k * 225;
            const double* Qk = sQK + k * 225;
            for (int e = tid; e < 225; e += PT) {
                const int i = e / 15, j = e - 15 * i;
                double acc = 0.0, acc2 = 0.0;
#pragma unroll 5
                for (int l = 0; l < 15; ++l) {
                    acc += Phi[i * 15 + l] * sPii[l * 15 + j];
                    acc2 += Phi[i * 15 + l] * sTot[l * 15 + j];
                }
                sT[e] = acc;
                sTot2[e] = acc2;
            }
            __syncthreads();
            double r0 = 0.0, r1 = 0.0;
            const int e0 = tid, e1 = tid + PT;
            {
                const int i = e0 / 15, j = e0 - 15 * i;
                double x = 0.0, y = 0.0;
#pragma unroll 5
                for (int l = 0; l < 15; ++l) { x += sT[i * 15 + l] * Phi[j * 15 + l]; y += sT[j * 15 + l] * Phi[i * 15 + l]; }
                r0 = ((x + Qk[i * 15 + j]) + (y + Qk[j * 15 + i])) / 2;
            }
            if (e1 < 225) {
                const int i = e1 / 15, j = e1 - 15 * i;
                double x = 0.0, y = 0.0;
#pragma unroll 5
                for (int l = 0; l < 15; ++l) { x += sT[i * 15 + l] * Phi[j * 15 + l]; y += sT[j * 15 + l] * Phi[i * 15 + l]; }
                r1 = ((x + Qk[i * 15 + j]) + (y + Qk[j * 15 + i])) / 2;
            }
            sPii[e0] = r0;
            if (e1 < 225) sPii[e1] = r1;
            { double* tmp = sTot; sTot = sTot2; sTot2 = tmp; }
            __syncthreads();
        }
    }

    // ---- E: P_IC <- Phi_tot P_IC, one thread per column; row strip and its transpose written once ----
    const volatile double* vTot = sTot;   // (volatile: keeps the 225 loop-invariant loads out of the registers)
    for (int j = tid; j < c; j += PT) {
        double col[15];
#pragma unroll
        for (int i = 0; i < 15; ++i) col[i] = P[(size_t)i * ldp + 15 + j];
        double* prow = P + (size_t)(15 + j) * ldp;
#pragma unroll 1
        for (int i = 0; i < 15; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int l = 0; l < 15; ++l) acc += vTot[i * 15 + l] * col[l];
            P[(size_t)i * ldp + 15 + j] = acc;
            prow[i] = acc;
        }
    }
    for (int e = tid; e < 225; e += PT) P[(size_t)(e / 15) * ldp + (e % 15)] = sPii[e];
    if (tid < 21) imu_state[(size_t)b * 21 + tid] = st[tid];
    if (tid < 15 && imu_null) imu_null[(size_t)b * 15 + tid] = sn[tid];
    if (tid == 0) imu_time[b] = sc[0];
}

}  // namespace

cudaError_t launch_propagate(const msckf_dims& d, int n_samples, const double* samples, double* imu_time,
                             const double* Qc_diag, const double* g, double* imu_state, double* imu_null, double* P,
                             const int* n_cams, cudaStream_t s) {
    const int cap = n_samples < PCH ? n_samples : PCH;
    const size_t smem = ((size_t)cap * 450 + 4 * 450 + 4 * 225 + 21 + 15 + 1) * sizeof(double) + (size_t)cap * sizeof(SampleState);
    static size_t configured2 = 0;
    if (smem > 48 * 1024 && smem > configured2) {
        cudaError_t e = cudaFuncSetAttribute(propagate_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured2 = smem;
    }
    propagate_kernel_v2<<<d.B, PT, smem, s>>>(d, n_samples, samples, imu_time, Qc_diag, g[0], g[1], g[2], imu_state,
                                               imu_null, P, n_cams, cap);
    msckf_count_launch(1);
    return cudaGetLastError();
}
