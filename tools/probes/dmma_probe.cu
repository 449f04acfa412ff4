// Throughput probe for the FP64 paths on sm_100a: DFMA chains, mma.sync m8n8k4 / m16n8k8 / m16n8k16 f64.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_dfma(int iters, double* sink) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fma(a[j], m, c);
    double s = 0;
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 12345.678) sink[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_m8n8k4(int iters, double* sink) {
    double c0[NACC], c1[NACC];
    for (int i = 0; i < NACC; ++i) { c0[i] = 0; c1[i] = 0; }
    double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-7;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < NACC; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[j]), "+d"(c1[j]) : "d"(a), "d"(b));
    double s = 0;
    for (int j = 0; j < NACC; ++j) s += c0[j] + c1[j];
    if (s == 12345.678) sink[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_m16n8k8(int iters, double* sink) {
    double c[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int k = 0; k < 4; ++k) c[i][k] = 0;
    double a0 = threadIdx.x * 1e-6, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 1.0 + threadIdx.x * 1e-7, b1 = b0 + 1;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < NACC; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    double s = 0;
    for (int j = 0; j < NACC; ++j) for (int k = 0; k < 4; ++k) s += c[j][k];
    if (s == 12345.678) sink[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_m16n8k16(int iters, double* sink) {
    double c[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int k = 0; k < 4; ++k) c[i][k] = 0;
    double a[8], b[4];
    for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-6 + k;
    for (int k = 0; k < 4; ++k) b[k] = 1.0 + threadIdx.x * 1e-7 + k;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < NACC; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    double s = 0;
    for (int j = 0; j < NACC; ++j) for (int k = 0; k < 4; ++k) s += c[j][k];
    if (s == 12345.678) sink[0] = s;
}

template <typename F>
static void run(const char* name, double flops_per_thread_iter_x32, int ctas, int threads, int iters, F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double warps = (double)ctas * threads / 32.0;
    const double flops = flops_per_thread_iter_x32 * warps * iters;
    printf("%-28s ctas=%5d thr=%4d  %8.3f ms  %8.2f TFLOP/s  (%s)\n", name, ctas, threads, best, flops / best / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double* sink; cudaMalloc(&sink, 64);
    const int it = 20000;
    for (int cps : {1, 2, 4, 8}) {
        const int ctas = 148 * cps;
        run("dfma x8 chains", 2.0 * 8 * 32, ctas, 256, it, [&] { k_dfma<<<ctas, 256>>>(it, sink); });
    }
    for (int cps : {1, 2, 4}) {
        const int ctas = 148 * cps;
        run("dmma m8n8k4 x4", 512.0 * 4, ctas, 256, it, [&] { k_m8n8k4<4><<<ctas, 256>>>(it, sink); });
        run("dmma m8n8k4 x8", 512.0 * 8, ctas, 256, it, [&] { k_m8n8k4<8><<<ctas, 256>>>(it, sink); });
        run("dmma m8n8k4 x16", 512.0 * 16, ctas, 256, it, [&] { k_m8n8k4<16><<<ctas, 256>>>(it, sink); });
        run("dmma m16n8k8 x4", 2048.0 * 4, ctas, 256, it, [&] { k_m16n8k8<4><<<ctas, 256>>>(it, sink); });
        run("dmma m16n8k8 x8", 2048.0 * 8, ctas, 256, it, [&] { k_m16n8k8<8><<<ctas, 256>>>(it, sink); });
        run("dmma m16n8k16 x4", 4096.0 * 4, ctas, 256, it, [&] { k_m16n8k16<4><<<ctas, 256>>>(it, sink); });
        run("dmma m16n8k16 x8", 4096.0 * 8, ctas, 256, it, [&] { k_m16n8k16<8><<<ctas, 256>>>(it, sink); });
    }
    // one warp per SM sub-partition only: issue-limited view
    run("dmma m8n8k4 x8 (128 thr)", 512.0 * 8, 148, 128, it, [&] { k_m8n8k4<8><<<148, 128>>>(it, sink); });
    run("dfma x8 (128 thr)", 2.0 * 8 * 32, 148, 128, it, [&] { k_dfma<<<148, 128>>>(it, sink); });
    return 0;
}
