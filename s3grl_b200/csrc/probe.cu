// Ceilings the path is measured against besides the HBM copy peak (bench.py `roofline.kernels`): PubMed's
// feature matrix (39 MB) is L2-resident, so kernel 3 is bound by the L2 -> SM read bandwidth and the FP32
// issue rate, not by HBM. Two micro-kernels measure those two ceilings on the GPU the bench runs on:
//   s3_probe_l2_read : the grid streams an L2-sized buffer `iters` times with 128-bit L1-bypassing loads
//   s3_probe_fma     : 8 independent FFMA chains per thread
// Diagnostics only: nothing on the product path calls them.
#include "common.cuh"

namespace s3 {
namespace {

__global__ void __launch_bounds__(256) probe_l2_read_kernel(const float4* __restrict__ buf, int64_t n4, int iters, float* sink) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it) {
        // grid-stride pass over the buffer, rotated per iteration
        int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x + (int64_t)it * 977 * blockDim.x) % n4;
        for (int64_t done = 0; done < n4; done += 4 * stride) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int64_t j = i + u * stride;
                if (j >= n4) j -= n4;
                v[u] = __ldcg(buf + j);  // L2 only: no L1 allocation
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc.x += v[u].x;
                acc.y += v[u].y;
                acc.z += v[u].z;
                acc.w += v[u].w;
            }
            i += 4 * stride;
            if (i >= n4) i -= n4;
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 1234.5f) sink[0] = acc.x;  // keeps the loads alive
}

__global__ void __launch_bounds__(256) probe_fma_kernel(int iters, float* sink) {
    float a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = 1.0f + 0.001f * (threadIdx.x + q);
    const float m = 0.999f, c = 0.001f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = fmaf(a[q], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += a[q];
    if (s == 1234.5f) sink[0] = s;
}

// the same chains as packed FP32 pairs: fma.rn.f32x2 (SASS FFMA2), two FMAs per issued instruction
__global__ void __launch_bounds__(256) probe_fma2_kernel(int iters, float* sink) {
    unsigned long long a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float2 v = make_float2(1.0f + 0.001f * (threadIdx.x + q), 1.0f + 0.002f * (threadIdx.x + q));
        a[q] = *reinterpret_cast<const unsigned long long*>(&v);
    }
    const float2 m2 = make_float2(0.999f, 0.998f), c2 = make_float2(0.001f, 0.002f);
    const unsigned long long m = *reinterpret_cast<const unsigned long long*>(&m2), c = *reinterpret_cast<const unsigned long long*>(&c2);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < 8; ++q) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[q]) : "l"(m), "l"(c));
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float2 v = *reinterpret_cast<const float2*>(&a[q]);
        s += v.x + v.y;
    }
    if (s == 1234.5f) sink[0] = s;
}

}  // namespace

cudaError_t launch_probe_fma2(int iters, float* sink, int ctas, cudaStream_t st) {
    probe_fma2_kernel<<<ctas, 256, 0, st>>>(iters, sink);   // 128 FMAs per thread per iteration, as probe_fma
    return cudaGetLastError();
}

cudaError_t launch_probe_l2_read(const float* buf, int64_t bytes, int iters, float* sink, int ctas, cudaStream_t st) {
    probe_l2_read_kernel<<<ctas, 256, 0, st>>>(reinterpret_cast<const float4*>(buf), bytes / 16, iters, sink);
    return cudaGetLastError();
}

cudaError_t launch_probe_fma(int iters, float* sink, int ctas, cudaStream_t st) {
    probe_fma_kernel<<<ctas, 256, 0, st>>>(iters, sink);
    return cudaGetLastError();
}

}  // namespace s3
