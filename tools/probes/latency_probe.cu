// Dependent-chain latencies of the instructions on the critical path of the in-tile Cholesky factorisations (one warp,
// clock64 around 256 dependent operations).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency_probe latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void probe(double* out, long long* clk, double seed) {
    const int lane = threadIdx.x & 31;
    double x = seed + lane * 1e-3, y = 1.0 + seed;
    long long t0, t1;
    constexpr int N = 256;
    // DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = fma(x, y, seed);
    t1 = clock64();
    if (threadIdx.x == 0) clk[0] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = x * y;
    t1 = clock64();
    if (threadIdx.x == 0) clk[1] = t1 - t0;
    // SHFL chain (64-bit = two 32-bit shuffles)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 5) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) clk[2] = t1 - t0;
    // rsqrt seed chain (MUFU.RSQ64H)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double r;
        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        x = r;
    }
    t1 = clock64();
    if (threadIdx.x == 0) clk[3] = t1 - t0;
    // DMMA chain on one accumulator
    double c0 = x, c1 = y;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) dmma(c0, c1, y, seed);
    t1 = clock64();
    if (threadIdx.x == 0) clk[4] = t1 - t0;
    // DMMA: accumulator of one feeds the A operand of the next (panel-solve pattern)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) { double n0 = 0, n1 = 0; dmma(n0, n1, c0, seed); c0 = n0 + n1; }
    t1 = clock64();
    if (threadIdx.x == 0) clk[5] = t1 - t0;
    // shared-memory store -> load round trip (same thread)
    __shared__ double sm[64];
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31]; __syncwarp(); }
    t1 = clock64();
    if (threadIdx.x == 0) clk[6] = t1 - t0;
    // 8 independent DMMA accumulators (throughput of one warp)
    double a[8][2];
    for (int i = 0; i < 8; ++i) a[i][0] = a[i][1] = x + i;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N / 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma(a[j][0], a[j][1], y, seed);
    t1 = clock64();
    if (threadIdx.x == 0) clk[7] = t1 - t0;
    double s = x + c0 + c1;
    for (int i = 0; i < 8; ++i) s += a[i][0] + a[i][1];
    out[threadIdx.x] = s;
}

int main() {
    double* out; long long* clk;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&clk, 8 * 8);
    for (int r = 0; r < 2; ++r) probe<<<1, 32>>>(out, clk, 1e-9);
    long long h[8];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    const char* nm[8] = {"DFMA dependent", "DMUL dependent", "SHFL(64-bit) dependent", "MUFU.RSQ64H dependent", "DMMA same accumulator",
                         "DMMA acc -> A operand (+DADD)", "STS -> LDS round trip (+2 syncwarp)", "DMMA 8 independent accumulators (per DMMA)"};
    for (int i = 0; i < 8; ++i) printf("%-48s %7.1f clk\n", nm[i], h[i] / 256.0);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
