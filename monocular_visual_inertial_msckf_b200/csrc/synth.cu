// Data synthesis + metrics on the device, so that a Monte-Carlo run over thousands of filters
// builds its streams and reduces its errors where the filters live, and the final gather carries statistics only:
//
//   synth_imu_kernel            IMU synthesis from ground-truth poses
//                               reference dataset/tools/dataset_generators/photorealistic_generator.py:145-216 (key poses
//                               interpolated: linear position + Slerp; finite differences; white noise + walk increment)
//                               and dataset/tools/dataset_generators/synthetic_generator.py:145-186 (poses at IMU rate,
//                               accumulated random-walk bias)
//   project_landmarks_kernel    landmark projection into the camera of a ground-truth pose, visible ones compacted in id order
//                               synthetic_generator.py:189-223 (Camera.W2Ci / project_point, src/msckf/Camera.py:13-52)
//   metrics_step / finalize     RTE / ROE with the reference's 10-sample smoothing, ATE / AOE and their 3-sigma bounds, and
//                               the run summary -- reference main.py:312-391, 436-437, 440-445
//
// All three are bandwidth / latency trivial next to the update path; they exist to keep host Python out of the frame loop.
// Quaternions follow scipy.spatial.transform.Rotation (x, y, z, w; from_matrix picks the largest of trace / diagonal),
// because the reference's R2quaternion / Slerp / R2euler wrap it (src/utils/geometry.py:132-194).
#include "msckf_common.cuh"

namespace {

// ---- counter-based normal draws (Philox4x32-10 + Box-Muller) for callers that do not supply their own noise ----
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// two independent N(0, 1) draws for (seed, stream, index)
__device__ __forceinline__ void normal_pair(unsigned long long seed, uint32_t stream, uint32_t index, double& z0, double& z1) {
    uint32_t c[4] = {index, stream, 0x5EEDu, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u0 = ((double)(((uint64_t)c[0] << 21) ^ (uint64_t)c[1]) + 0.5) * (1.0 / 9007199254740992.0);   // (0, 1), 53 bits
    const double u1 = ((double)(((uint64_t)c[2] << 21) ^ (uint64_t)c[3]) + 0.5) * (1.0 / 9007199254740992.0);
    const double r = sqrt(-2.0 * log(u0));
    double s, co;
    sincospi(2.0 * u1, &s, &co);
    z0 = r * co;
    z1 = r * s;
}

// ---- rotations ----
__device__ __forceinline__ void quat_from_matrix(const double* R, double* q) {   // (x, y, z, w), scipy's branch rule
    const double tr = R[0] + R[4] + R[8];
    // argmax over (R00, R11, R22, trace) with numpy's first-maximum tie rule
    int c = 0;
    double best = R[0];
    if (R[4] > best) { best = R[4]; c = 1; }
    if (R[8] > best) { best = R[8]; c = 2; }
    if (tr > best) { best = tr; c = 3; }
    // scipy: i = choice, j = i + 1, k = j + 1 (mod 3): q[i] = 1 - tr + 2 R[i][i], q[j] = R[j][i] + R[i][j], q[k] = R[k][i] + R[i][k],
    // q[3] = R[k][j] - R[j][k]  (written out per case: static register indices)
    if (c == 0) {
        q[0] = 1.0 - tr + 2.0 * R[0];
        q[1] = R[3] + R[1];
        q[2] = R[6] + R[2];
        q[3] = R[7] - R[5];
    } else if (c == 1) {
        q[1] = 1.0 - tr + 2.0 * R[4];
        q[2] = R[7] + R[5];
        q[0] = R[1] + R[3];
        q[3] = R[2] - R[6];
    } else if (c == 2) {
        q[2] = 1.0 - tr + 2.0 * R[8];
        q[0] = R[2] + R[6];
        q[1] = R[5] + R[7];
        q[3] = R[3] - R[1];
    } else {
        q[0] = R[7] - R[5];
        q[1] = R[2] - R[6];
        q[2] = R[3] - R[1];
        q[3] = 1.0 + tr;
    }
    const double n = 1.0 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
#pragma unroll
    for (int e = 0; e < 4; ++e) q[e] *= n;
}

__device__ __forceinline__ void quat_mul(const double* a, const double* b, double* o) {
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
    o[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}

__device__ __forceinline__ void quat_to_matrix(const double* q, double* R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

// Slerp([0, 1], [R0, R1])(t) as a matrix: R0 exp(t log(R0^T R1))
__device__ __forceinline__ void slerp_matrix(const double* R0, const double* R1, double t, double* R) {
    double q0[4], q1[4], qc[4], d[4];
    quat_from_matrix(R0, q0);
    quat_from_matrix(R1, q1);
    qc[0] = -q0[0]; qc[1] = -q0[1]; qc[2] = -q0[2]; qc[3] = q0[3];
    quat_mul(qc, q1, d);
    if (d[3] < 0.0) { d[0] = -d[0]; d[1] = -d[1]; d[2] = -d[2]; d[3] = -d[3]; }
    const double nv = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const double ang = 2.0 * atan2(nv, d[3]);
    double dq[4] = {0.0, 0.0, 0.0, 1.0};
    if (nv > 1e-300 && ang > 0.0) {
        const double a = t * ang;
        const double sc = sin(0.5 * a) / nv;
        dq[0] = d[0] * sc; dq[1] = d[1] * sc; dq[2] = d[2] * sc; dq[3] = cos(0.5 * a);
    }
    double q[4];
    quat_mul(q0, dq, q);
    quat_to_matrix(q, R);
}

// pose of interpolation point j (0 .. sub) of key interval i (1 .. K-1): R[9], p[3], time
__device__ __forceinline__ void interval_pose(const double* kp, const double* kt, int i, int j, int sub, int interp, double* R,
                                              double* p, double& tm) {
    const double* A = kp + (size_t)(i - 1) * 12;
    const double* Bp = kp + (size_t)i * 12;
    if (interp <= 1) {
        const double* S = j == 0 ? A : Bp;
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = S[e];
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = S[9 + e];
        tm = j == 0 ? kt[i - 1] : kt[i];
        return;
    }
    // np.linspace(0, 1, interp)[j]: j * step, the last point set to 1 exactly
    const double t = (j == sub) ? 1.0 : (double)j * (1.0 / (double)sub);
    slerp_matrix(A, Bp, t, R);
#pragma unroll
    for (int e = 0; e < 3; ++e) p[e] = (1.0 - t) * A[9 + e] + t * Bp[9 + e];
    tm = (1.0 - t) * kt[i - 1] + t * kt[i];
}

__global__ void synth_imu_kernel(int B, int K, int interp, const double* __restrict__ key_poses,
                                 const double* __restrict__ key_times, double g0, double g1, double g2,
                                 const double* __restrict__ noise, const double* __restrict__ sigmas, unsigned long long seed,
                                 int accumulate_bias, int sign_fix, double* __restrict__ poses_out, double* __restrict__ times_out,
                                 double* __restrict__ imu_gt, double* __restrict__ imu_noisy) {
    const int sub = interp > 1 ? interp - 1 : 1;
    const int T = sub * (K - 1);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * (T + 1)) return;
    const int b = (int)(gid / (T + 1)), k = (int)(gid - (long long)b * (T + 1));
    const double* kp = key_poses + (size_t)b * K * 12;
    const double* kt = key_times + (size_t)b * K;
    double* po = poses_out ? poses_out + ((size_t)b * (T + 1) + k) * 12 : nullptr;
    double* og = imu_gt + ((size_t)b * (T + 1) + k) * 7;
    double* on = imu_noisy + ((size_t)b * (T + 1) + k) * 7;
    if (k == 0) {   // the reference's first row: the first pose, zero measurements (photorealistic_generator.py:124-135)
        if (po)
            for (int e = 0; e < 12; ++e) po[e] = kp[e];
        if (times_out) times_out[(size_t)b * (T + 1)] = kt[0];
        og[0] = kt[0]; on[0] = kt[0];
        for (int e = 1; e < 7; ++e) { og[e] = 0.0; on[e] = 0.0; }
        return;
    }
    const int i = (k - 1) / sub + 1, j = (k - 1) % sub + 1;
    double Rp[9], Rc[9], pp[3], pc[3], tp, tc;
    interval_pose(kp, kt, i, j - 1, sub, interp, Rp, pp, tp);
    interval_pose(kp, kt, i, j, sub, interp, Rc, pc, tc);
    const double dt = tc - tp;
    double v[3], vprev[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int e = 0; e < 3; ++e) v[e] = (pc[e] - pp[e]) / dt;
    if (k > 1) {   // the velocity of the previous sample (it may lie in the previous key interval)
        double Rq[9], pq[3], pr[3], tq, tr;
        const int i2 = (k - 2) / sub + 1, j2 = (k - 2) % sub + 1;
        interval_pose(kp, kt, i2, j2 - 1, sub, interp, Rq, pq, tq);
        interval_pose(kp, kt, i2, j2, sub, interp, Rq, pr, tr);
        const double dt2 = tr - tq;
#pragma unroll
        for (int e = 0; e < 3; ++e) vprev[e] = (pr[e] - pq[e]) / dt2;
    }
    double al[3], acc[3], om[3];
    al[0] = (v[0] - vprev[0]) / dt + g0;
    al[1] = (v[1] - vprev[1]) / dt + g1;
    al[2] = (v[2] - vprev[2]) / dt + g2;
#pragma unroll
    for (int e = 0; e < 3; ++e) acc[e] = Rp[e] * al[0] + Rp[3 + e] * al[1] + Rp[6 + e] * al[2];     // prev_rotation^T (a + g)
    {
        double qp[4], qc[4];
        quat_from_matrix(Rp, qp);
        quat_from_matrix(Rc, qc);
        if (sign_fix && qp[0] * qc[0] + qp[1] * qc[1] + qp[2] * qc[2] + qp[3] * qc[3] < 0.0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) qc[e] = -qc[e];
        }
        const double w1 = qp[3], x1 = qp[0], y1 = qp[1], z1 = qp[2];
        const double w2 = qc[3], x2 = qc[0], y2 = qc[1], z2 = qc[2];
        const double s = 2.0 / dt;
        om[0] = s * (w1 * x2 - x1 * w2 - y1 * z2 + z1 * y2);
        om[1] = s * (w1 * y2 + x1 * z2 - y1 * w2 - z1 * x2);
        om[2] = s * (w1 * z2 - x1 * y2 + y1 * x2 - z1 * w2);
    }
    if (po) {
#pragma unroll
        for (int e = 0; e < 9; ++e) po[e] = Rc[e];
#pragma unroll
        for (int e = 0; e < 3; ++e) po[9 + e] = pc[e];
    }
    if (times_out) times_out[(size_t)b * (T + 1) + k] = tc;
    og[0] = tc; on[0] = tc;
#pragma unroll
    for (int e = 0; e < 3; ++e) { og[1 + e] = om[e]; og[4 + e] = acc[e]; }
    // noise [B][T][12]: acc white, acc walk increment, gyro white, gyro walk increment (already scaled); drawn here otherwise
    double nz[12];
    if (noise) {
        const double* src = noise + ((size_t)b * T + (k - 1)) * 12;
#pragma unroll
        for (int e = 0; e < 12; ++e) nz[e] = src[e];
    } else if (sigmas) {
        const double* sg = sigmas + (size_t)b * 4;      // acc, gyro, acc walk, gyro walk
        const double sc[4] = {sg[0], sg[2], sg[1], sg[3]};
#pragma unroll
        for (int h = 0; h < 6; ++h) {
            double z0, z1;
            normal_pair(seed, (uint32_t)b, (uint32_t)((k - 1) * 6 + h), z0, z1);
            nz[2 * h] = z0 * sc[(2 * h) / 3];
            nz[2 * h + 1] = z1 * sc[(2 * h + 1) / 3];
        }
    } else {
#pragma unroll
        for (int e = 0; e < 12; ++e) nz[e] = 0.0;
    }
    if (accumulate_bias) {   // synthetic_generator.py:168-181: white noise now, the accumulated walk in bias_kernel
#pragma unroll
        for (int e = 0; e < 3; ++e) { on[1 + e] = om[e] + nz[6 + e]; on[4 + e] = acc[e] + nz[e]; }
    } else {                 // photorealistic_generator.py:189-198: walk INCREMENT + white noise
#pragma unroll
        for (int e = 0; e < 3; ++e) { on[1 + e] = om[e] + nz[9 + e] + nz[6 + e]; on[4 + e] = acc[e] + nz[3 + e] + nz[e]; }
    }
}

// accumulated random-walk biases (synthetic_generator.py:175-181): one thread per (filter, component), a running sum over time
__global__ void synth_bias_kernel(int B, int T, const double* __restrict__ noise, const double* __restrict__ sigmas,
                                  unsigned long long seed, double* __restrict__ imu_noisy) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= B * 6) return;
    const int b = gid / 6, c = gid - 6 * b;        // c 0-2: gyro, 3-5: acc
    const int e = c < 3 ? 9 + c : 3 + (c - 3);     // walk increment index in the 12-vector
    double run = 0.0;
    for (int k = 1; k <= T; ++k) {
        double inc;
        if (noise) {
            inc = noise[((size_t)b * T + (k - 1)) * 12 + e];
        } else {
            double z0, z1;
            normal_pair(seed, (uint32_t)b, (uint32_t)((k - 1) * 6 + e / 2), z0, z1);
            inc = ((e & 1) ? z1 : z0) * sigmas[(size_t)b * 4 + (c < 3 ? 3 : 2)];
        }
        run += inc;
        imu_noisy[((size_t)b * (T + 1) + k) * 7 + 1 + c] += run;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// landmark projection: one CTA per (filter, frame), landmarks in chunks of blockDim.x, compaction in id order
// ---------------------------------------------------------------------------------------------------------------------
struct ProjConsts {
    double Ric[9], tic[3], K[9];
};

__global__ void __launch_bounds__(256) project_landmarks_kernel(int L, int n_frames, const double* __restrict__ body_poses,
                                                                ProjConsts pc, int width, int height,
                                                                const double* __restrict__ landmarks,
                                                                const double* __restrict__ noise_uv, double sigma_px,
                                                                unsigned long long seed, double* __restrict__ kp_out,
                                                                int* __restrict__ id_out, int* __restrict__ count_out,
                                                                double* __restrict__ score_out) {
    const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t bf = (size_t)b * n_frames + f;
    const double* T = body_poses + bf * 12;
    __shared__ double sRc[9], stc[3];
    __shared__ int s_cnt[8], s_base;
    if (tid < 9) {   // T_W_C = T_W_I T_I_C
        const int r = tid / 3, c = tid % 3;
        sRc[tid] = T[r * 3] * pc.Ric[c] + T[r * 3 + 1] * pc.Ric[3 + c] + T[r * 3 + 2] * pc.Ric[6 + c];
    }
    if (tid < 3) stc[tid] = T[tid * 3] * pc.tic[0] + T[tid * 3 + 1] * pc.tic[1] + T[tid * 3 + 2] * pc.tic[2] + T[9 + tid];
    if (tid == 0) s_base = 0;
    __syncthreads();
    const double score = 1.0 / (1.0 + 2.0 * sigma_px * sigma_px);      // 1 / (1 + trace(cov)), synthetic_generator.py:219
    for (int j0 = 0; j0 < L; j0 += blockDim.x) {
        const int j = j0 + tid;
        bool vis = false;
        double u = 0.0, v = 0.0;
        if (j < L) {
            const double* W = landmarks + ((size_t)b * L + j) * 3;
            const double d0 = W[0] - stc[0], d1 = W[1] - stc[1], d2 = W[2] - stc[2];
            const double cx = sRc[0] * d0 + sRc[3] * d1 + sRc[6] * d2;
            const double cy = sRc[1] * d0 + sRc[4] * d1 + sRc[7] * d2;
            const double cz = sRc[2] * d0 + sRc[5] * d1 + sRc[8] * d2;
            if (cz > 0.0) {
                const double iw = pc.K[6] * cx + pc.K[7] * cy + pc.K[8] * cz;
                u = (pc.K[0] * cx + pc.K[1] * cy + pc.K[2] * cz) / iw;
                v = (pc.K[3] * cx + pc.K[4] * cy + pc.K[5] * cz) / iw;
                vis = !(u < 0.0 || u >= (double)width || v < 0.0 || v >= (double)height);
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, vis);
        if (lane == 0) s_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_cnt[w];
        if (vis) {
            double n0 = 0.0, n1 = 0.0;
            if (noise_uv) {
                n0 = noise_uv[(bf * L + j) * 2];
                n1 = noise_uv[(bf * L + j) * 2 + 1];
            } else if (sigma_px > 0.0) {
                normal_pair(seed, (uint32_t)bf, (uint32_t)j, n0, n1);
                n0 *= sigma_px;
                n1 *= sigma_px;
            }
            const size_t o = bf * L + off + __popc(bal & ((1u << lane) - 1u));
            kp_out[o * 2] = u + n0;
            kp_out[o * 2 + 1] = v + n1;
            id_out[o] = j;
            if (score_out) score_out[o] = score;
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_cnt[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (tid == 0) count_out[bf] = s_base;
}

// ---------------------------------------------------------------------------------------------------------------------
// metrics (main.py:312-391): one thread per filter; the running state lives in a 64-double record per filter
// ---------------------------------------------------------------------------------------------------------------------
// state layout: [0..11] previous gt pose (R, t)  [12..23] previous estimate  [24..33] deque rte  [34..43] deque roe
//               [44] samples in the deques  [45] n  [46] sum rte^2  [47] sum roe^2  [48] mean rte_s  [49] M2 rte_s
//               [50] mean roe_s  [51] M2 roe_s  [52] sum |ate|^2
constexpr int MS = 64;

__device__ __forceinline__ void rel_pose(const double* Ra, const double* pa, const double* Rb, const double* pb, double* R,
                                         double* p) {   // T_a^-1 T_b
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) R[r * 3 + c] = Ra[r] * Rb[c] + Ra[3 + r] * Rb[3 + c] + Ra[6 + r] * Rb[6 + c];
        p[r] = Ra[r] * (pb[0] - pa[0]) + Ra[3 + r] * (pb[1] - pa[1]) + Ra[6 + r] * (pb[2] - pa[2]);
    }
}

__device__ __forceinline__ double rot_angle(const double* R) {
    return acos(fmin(1.0, fmax(-1.0, ((R[0] + R[4] + R[8]) - 1.0) / 2.0)));
}

__global__ void metrics_step_kernel(msckf_dims dims, const double* __restrict__ gt_pose, const double* __restrict__ imu_state,
                                    const double* __restrict__ P, double* __restrict__ state, double* __restrict__ record) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= dims.B) return;
    double* st = state + (size_t)b * MS;
    const double* g = gt_pose + (size_t)b * 12;
    const double* e = imu_state + (size_t)b * MSCKF_IMU_STATE_DOUBLES;       // R (9), t (3), ...
    const bool first = st[45] == 0.0;
    double R0g[9], p0g[3], R0e[9], p0e[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        R0g[k] = first ? ((k % 4 == 0) ? 1.0 : 0.0) : st[k];
        R0e[k] = first ? ((k % 4 == 0) ? 1.0 : 0.0) : st[12 + k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { p0g[k] = first ? 0.0 : st[9 + k]; p0e[k] = first ? 0.0 : st[21 + k]; }
    double Rrg[9], prg[3], Rre[9], pre[3], Ra[9], pa[3], Rr[9], pr[3];
    rel_pose(R0g, p0g, g, g + 9, Rrg, prg);          // T_I0_I1_gt
    rel_pose(R0e, p0e, e, e + 9, Rre, pre);          // T_I0_I1_est
    rel_pose(g, g + 9, e, e + 9, Ra, pa);            // abs_T_error
    rel_pose(Rrg, prg, Rre, pre, Rr, pr);            // rel_T_error
    const double rel_t_gt = sqrt(prg[0] * prg[0] + prg[1] * prg[1] + prg[2] * prg[2]);
    const double rel_o_gt = rot_angle(Rrg);
    const double rte = rel_t_gt != 0.0 ? sqrt(pr[0] * pr[0] + pr[1] * pr[1] + pr[2] * pr[2]) / rel_t_gt : 0.0;
    const double roe = rel_o_gt != 0.0 ? rot_angle(Rr) / rel_o_gt : 0.0;
    const int nd = (int)st[44];
    double mp = 0.0, mo = 0.0;
    for (int k = 0; k < nd; ++k) { mp += st[24 + k]; mo += st[34 + k]; }
    if (nd > 0) { mp /= nd; mo /= nd; }
    const double rte_s = (nd > 0 && rte > 10.0 * mp) ? 0.001 * rte + 0.999 * mp : rte;
    const double roe_s = (nd > 0 && roe > 10.0 * mo) ? 0.001 * roe + 0.999 * mo : roe;
    if (nd < 10) {
        st[24 + nd] = rte_s; st[34 + nd] = roe_s; st[44] = nd + 1;
    } else {      // deque(maxlen=10): drop the oldest
        for (int k = 0; k < 9; ++k) { st[24 + k] = st[25 + k]; st[34 + k] = st[35 + k]; }
        st[33] = rte_s; st[43] = roe_s;
    }
    const double n = st[45] + 1.0;
    st[45] = n;
    st[46] += rte * rte;
    st[47] += roe * roe;
    {   // Welford: np.std is the population standard deviation
        const double d1 = rte_s - st[48];
        st[48] += d1 / n;
        st[49] += d1 * (rte_s - st[48]);
        const double d2 = roe_s - st[50];
        st[50] += d2 / n;
        st[51] += d2 * (roe_s - st[50]);
    }
    st[52] += pa[0] * pa[0] + pa[1] * pa[1] + pa[2] * pa[2];
#pragma unroll
    for (int k = 0; k < 12; ++k) { st[k] = g[k]; st[12 + k] = e[k]; }
    if (record) {   // rte_s, roe_s, ate xyz, aoe (intrinsic XYZ euler of the absolute rotation error), 3-sigma bounds
        double* r = record + (size_t)b * 14;
        const double* Pb = P + (size_t)b * dims.ldp * dims.ldp;
        r[0] = rte_s; r[1] = roe_s;
        r[2] = pa[0]; r[3] = pa[1]; r[4] = pa[2];
        r[5] = atan2(-Ra[5], Ra[8]);
        r[6] = asin(fmin(1.0, fmax(-1.0, Ra[2])));
        r[7] = atan2(-Ra[1], Ra[0]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            r[8 + k] = 3.0 * sqrt(Pb[(size_t)(12 + k) * dims.ldp + 12 + k]);
            r[11 + k] = 3.0 * sqrt(Pb[(size_t)k * dims.ldp + k]);
        }
    }
}

__global__ void metrics_finalize_kernel(int B, const double* __restrict__ state, double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* st = state + (size_t)b * MS;
    double* o = out + (size_t)b * 8;
    const double n = st[45] > 0.0 ? st[45] : 1.0;
    o[0] = sqrt(st[46] / n);      // rmse_position  (main.py:440)
    o[1] = sqrt(st[47] / n);      // rmse_orientation
    o[2] = st[48];                // mean_rte
    o[3] = sqrt(st[49] / n);      // std_rte
    o[4] = st[50];                // mean_roe
    o[5] = sqrt(st[51] / n);      // std_roe
    o[6] = sqrt(st[52] / n);      // rmse of the absolute translation error
    o[7] = st[45];
}

}  // namespace

extern "C" {

int msckf_synth_imu_batched(int B, int K, int interp, const double* key_poses, const double* key_times, const double* gravity_host,
                            const double* noise, const double* sigmas, unsigned long long seed, int accumulate_bias, int sign_fix,
                            double* poses_out, double* times_out, double* imu_gt, double* imu_noisy, void* stream) {
    if (B <= 0 || K < 2 || interp < 1 || !key_poses || !key_times || !gravity_host || !imu_gt || !imu_noisy) return MSCKF_EINVAL;
    const int sub = interp > 1 ? interp - 1 : 1;
    const long long T = (long long)sub * (K - 1);
    const long long n = (long long)B * (T + 1);
    if (n > 2147483647LL * 64 || T > 2147483647LL / 16) return MSCKF_ECAPACITY;
    cudaStream_t s = (cudaStream_t)stream;
    synth_imu_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(B, K, interp, key_poses, key_times, gravity_host[0], gravity_host[1],
                                                              gravity_host[2], noise, sigmas, seed, accumulate_bias, sign_fix,
                                                              poses_out, times_out, imu_gt, imu_noisy);
    msckf_count_launch(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (accumulate_bias && (noise || sigmas)) {
        synth_bias_kernel<<<(B * 6 + 127) / 128, 128, 0, s>>>(B, (int)T, noise, sigmas, seed, imu_noisy);
        msckf_count_launch(1);
        e = cudaGetLastError();
    }
    return (int)e;
}

int msckf_project_landmarks_batched(int B, int L, int n_frames, const double* body_poses, const double* T_I_C_host,
                                    const double* K_host, int width, int height, const double* landmarks, const double* noise_uv,
                                    double sigma_px, unsigned long long seed, double* kp_out, int* id_out, int* count_out,
                                    double* score_out, void* stream) {
    if (B <= 0 || L <= 0 || n_frames <= 0 || !body_poses || !T_I_C_host || !K_host || !landmarks || !kp_out || !id_out || !count_out)
        return MSCKF_EINVAL;
    if (B > 65535) return MSCKF_ECAPACITY;
    ProjConsts pc;
    for (int e = 0; e < 9; ++e) { pc.Ric[e] = T_I_C_host[e]; pc.K[e] = K_host[e]; }
    for (int e = 0; e < 3; ++e) pc.tic[e] = T_I_C_host[9 + e];
    dim3 grid(n_frames, B);
    project_landmarks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(L, n_frames, body_poses, pc, width, height, landmarks, noise_uv,
                                                                    sigma_px, seed, kp_out, id_out, count_out, score_out);
    msckf_count_launch(1);
    return (int)cudaGetLastError();
}

int msckf_metrics_state_doubles(void) { return MS; }

int msckf_metrics_step_batched(const msckf_dims* dims, const double* gt_pose, const double* imu_state, const double* P,
                               double* metrics_state, double* record_out, void* stream) {
    if (!dims || dims->B <= 0 || !gt_pose || !imu_state || !metrics_state || (record_out && !P)) return MSCKF_EINVAL;
    metrics_step_kernel<<<(dims->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*dims, gt_pose, imu_state, P, metrics_state, record_out);
    msckf_count_launch(1);
    return (int)cudaGetLastError();
}

int msckf_metrics_finalize_batched(int B, const double* metrics_state, double* summary_out, void* stream) {
    if (B <= 0 || !metrics_state || !summary_out) return MSCKF_EINVAL;
    metrics_finalize_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, metrics_state, summary_out);
    msckf_count_launch(1);
    return (int)cudaGetLastError();
}

}  // extern "C"
