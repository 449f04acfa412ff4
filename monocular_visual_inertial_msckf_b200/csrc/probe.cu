// FP64 FMA throughput probe: the measured denominator of the FP64-pipe roofline (MEASURED_PEAKS.json carries no FP64
// figure).  Every thread runs 8 independent DFMA chains; flops = 2 * 8 * iters * threads.
#include "msckf_common.cuh"

namespace {
__global__ void __launch_bounds__(256) fp64_probe_kernel(int iters, double* sink) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) sink[0] = s;
}
}  // namespace

extern "C" int msckf_probe_fp64(int n_ctas, int iters, double* sink, double* flops_out_host, void* stream) {
    if (n_ctas <= 0 || iters <= 0 || !sink) return MSCKF_EINVAL;
    fp64_probe_kernel<<<n_ctas, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    msckf_count_launch(1);
    if (flops_out_host) *flops_out_host = 2.0 * 8.0 * (double)iters * 256.0 * (double)n_ctas;
    return (int)cudaGetLastError();
}
