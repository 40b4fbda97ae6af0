// Instantiations of the gather kernel for SC = 1, K+1 in [5, 6] (see gather_kernel.cuh).
#include "gather_kernel.cuh"

namespace s3 {
cudaError_t launch_gather_sc1_mid(const GatherParams& p, int K1, int C, dim3 grid, size_t smem, cudaStream_t st) {
    switch (K1) {
        case 5: return launch_k1<1, 5>(p, C, grid, smem, st);
        case 6: return launch_k1<1, 6>(p, C, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}
}  // namespace s3
