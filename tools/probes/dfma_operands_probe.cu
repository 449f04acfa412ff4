// DFMA issue rate against operand pattern on sm_100a (one warp per scheduler and four): shared multiplier / addend
// (fma(a_j, m, c): what the peak probe measures) vs three distinct register operands per instruction (fma(a_j, b_j, c_j),
// what a dense block product issues) vs two distinct + accumulator (a_j = fma(x_j, y_j, a_j)).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_operands_probe dfma_operands_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(int iters, double* sink) {
    double a[8], b[8], c[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3 + i; b[i] = 1.0 + 1e-9 * (i + threadIdx.x); c[i] = 1e-9 * (i + 1); }
    const double m = 1.0000001, cc = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[j]) : "d"(m), "d"(cc));
            if (MODE == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[j]) : "d"(b[j]), "d"(c[j]));
            if (MODE == 2) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[j]) : "d"(b[j]), "d"(c[(j + 3) & 7]));
            if (MODE == 3) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[j]) : "d"(b[j]), "d"(m));
        }
    }
    double s = 0;
    for (int j = 0; j < 8; ++j) s += a[j] + b[j] + c[j];
    if (s == 12345.678) sink[0] = s;
}

template <typename F>
static float run(F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    double* sink; cudaMalloc(&sink, 64);
    const int it = 20000;
    const char* names[4] = {"a = fma(a, m, c)      shared m, c", "a = fma(a, b_j, c_j)  three distinct", "a = fma(b_j, c_k, a)  two distinct + acc",
                            "a = fma(b_j, m, a)    one distinct + shared + acc"};
    for (int thr : {128, 512}) {
        float t[4];
        t[0] = run([&] { k<0><<<148, thr>>>(it, sink); });
        t[1] = run([&] { k<1><<<148, thr>>>(it, sink); });
        t[2] = run([&] { k<2><<<148, thr>>>(it, sink); });
        t[3] = run([&] { k<3><<<148, thr>>>(it, sink); });
        for (int q = 0; q < 4; ++q) {
            const double flops = 2.0 * 8 * it * 148.0 * thr;
            printf("threads/SM %4d  %-48s %.3f ms  %6.2f TFLOP/s  %.2f clk per warp-DFMA per scheduler\n", thr, names[q], t[q], flops / t[q] / 1e9,
                   t[q] * 1e-3 * 1.965e9 / (8.0 * it * (thr / 128)));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
