// Instantiations of the gather kernel for SC = 1, K+1 in [7, 8] (see gather_kernel.cuh).
#include "gather_kernel.cuh"

namespace s3 {
cudaError_t launch_gather_sc1_hi(const GatherParams& p, int K1, int C, dim3 grid, size_t smem, cudaStream_t st) {
    switch (K1) {
        case 7: return launch_k1<1, 7>(p, C, grid, smem, st);
        case 8: return launch_k1<1, 8>(p, C, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}
}  // namespace s3
