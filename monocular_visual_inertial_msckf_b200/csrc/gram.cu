// Information-form compression (MSCKF_COMPRESS_GRAM) -- placeholder until the Gram kernels land.
#include "msckf_common.cuh"
cudaError_t launch_compress_gram(const CompressArgs&, cudaStream_t) { return cudaErrorNotSupported; }
