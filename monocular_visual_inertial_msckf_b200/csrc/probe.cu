// FP64 throughput probe: the measured denominator of the FP64-pipe roofline (MEASURED_PEAKS.json carries no FP64
// figure).  Every warp runs 8 independent DMMA.8x8x4 accumulator chains (mma.sync.m8n8k4.f64); flops = 512 per DMMA.
// DFMA chains reach the same 36-37 TFLOP/s on B200 (tools/probes/dmma_probe.cu): FP64 vector and tensor share one peak.
#include "msckf_common.cuh"
#include "dense.cuh"

namespace {
__global__ void __launch_bounds__(256) fp64_probe_kernel(int iters, double* sink) {
    double c0[8], c1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c0[j] = 0.0; c1[j] = 0.0; }
    const double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-7;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dense::dmma(c0[j], c1[j], a, b);
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c0[j] + c1[j];
    if (s == 12345.678) sink[0] = s;
}
}  // namespace

extern "C" int msckf_probe_fp64(int n_ctas, int iters, double* sink, double* flops_out_host, void* stream) {
    if (n_ctas <= 0 || iters <= 0 || !sink) return MSCKF_EINVAL;
    fp64_probe_kernel<<<n_ctas, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    msckf_count_launch(1);
    if (flops_out_host) *flops_out_host = 512.0 * 8.0 * (double)iters * 8.0 * (double)n_ctas;
    return (int)cudaGetLastError();
}
