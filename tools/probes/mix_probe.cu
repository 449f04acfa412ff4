// Do DFMA and DMMA.8x8x4 share one pipe on sm_100a?  A warp issues, per iteration, 4 independent DMMAs (2048 flop) and
// 32 DFMAs in 8 chains (2048 flop).  Shared pipe: the mixed kernel takes the SUM of the two pure kernels; separate pipes:
// their maximum.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix_probe mix_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <bool DM, bool DF>
__global__ void __launch_bounds__(256) k_mix(int iters, double* sink) {
    double c0[4], c1[4], a[8];
    for (int i = 0; i < 4; ++i) { c0[i] = 0; c1[i] = 0; }
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double m = 1.0000001, c = 1e-9;
    double fa = threadIdx.x * 1e-6, fb = 1.0 + threadIdx.x * 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (DM)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0[j]), "+d"(c1[j]) : "d"(fa), "d"(fb));
            if (DF) {
#pragma unroll
                for (int q = 0; q < 8; ++q) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[q]) : "d"(m), "d"(c));
            }
        }
    }
    double s = 0;
    for (int j = 0; j < 4; ++j) s += c0[j] + c1[j];
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 12345.678) sink[0] = s;
}

template <typename F>
static float run(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    double* sink; cudaMalloc(&sink, 64);
    const int it = 20000;
    for (int thr : {128, 256, 512}) {
        const float tm = run([&] { k_mix<true, false><<<148, thr>>>(it, sink); });
        const float tf = run([&] { k_mix<false, true><<<148, thr>>>(it, sink); });
        const float tx = run([&] { k_mix<true, true><<<148, thr>>>(it, sink); });
        printf("threads/SM %4d: DMMA only %.3f ms, DFMA only %.3f ms, both %.3f ms  (sum %.3f, max %.3f) %s\n", thr, tm, tf, tx,
               tm + tf, tm > tf ? tm : tf, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
