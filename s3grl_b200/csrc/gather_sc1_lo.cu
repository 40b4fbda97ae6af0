// Instantiations of the gather kernel for SC = 1, K+1 in [2, 3, 4] (see gather_kernel.cuh).
#include "gather_kernel.cuh"

namespace s3 {
cudaError_t launch_gather_sc1_lo(const GatherParams& p, int K1, int C, dim3 grid, size_t smem, cudaStream_t st) {
    switch (K1) {
        case 2: return launch_k1<1, 2>(p, C, grid, smem, st);
        case 3: return launch_k1<1, 3>(p, C, grid, smem, st);
        case 4: return launch_k1<1, 4>(p, C, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}
}  // namespace s3
