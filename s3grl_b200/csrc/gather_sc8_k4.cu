// Instantiations of the gather kernel for 8-row CCN work items (PoS Plus union), K+1 = 4,
// one float4 column per thread (see gather_kernel.cuh). One file per K: these kernels are big.
#include "gather_kernel.cuh"

namespace s3 {
cudaError_t launch_gather_sc8_k4(const GatherParams& p, int C, dim3 grid, size_t smem, cudaStream_t st) {
    return launch_k1<8, 4>(p, C, grid, smem, st);
}
}  // namespace s3
