// Kernel 8 — CCN segment pooling (north_star kernel 3's "center / CCN pooling"; reference models.py:339-367,
// SIGNNet._centre_pool_helper with k_heuristic set: global_mean_pool / global_add_pool of the rows beyond the two
// targets of every link, next to the center product h_src * h_dst).
//
// Input: a row-stacked matrix [R, ld] whose rows are grouped per link by row_ptr (rows 0, 1 of a link = src, dst;
// rows 2.. = its CCN rows) — the hidden rows h that s3_sign_head(pool = 0) leaves, or an operator / joint matrix of
// the precompute path. Output per link, one row of
//   S3_POOL_OUT_CENTER : [ src * dst | pool(rows 2..) ]            (2 * cols; what link_pred_mlp consumes)
//   S3_POOL_OUT_ROWS   : [ src | dst | pool(rows 2..) ]            (3 * cols; the pooled output mode of SURVEY §7)
// pool = sum or mean over the link's CCN rows in ascending row order (fixed order: bit-reproducible; an empty
// segment gives zeros, as torch_scatter does). One CTA per link, columns strided over the threads, 128-bit loads when
// the row stride allows. Pure data movement: HBM-bound, 4 * (R + 2..3 * L) * cols bytes.
#include "common.cuh"

namespace s3 {
namespace {

constexpr int kPoolThreads = 128;

template <bool VEC>
__global__ void __launch_bounds__(kPoolThreads) segment_pool_kernel(const float* __restrict__ src, int64_t ld, int cols,
                                                                    const int64_t* __restrict__ row_ptr, int mode, int layout,
                                                                    float* __restrict__ out, int64_t ld_out) {
    const int64_t link = blockIdx.x;
    const int64_t r0 = row_ptr[link], r1 = row_ptr[link + 1];
    const int s = (int)(r1 - r0);
    float* orow = out + link * ld_out;
    const float inv = (mode == S3_POOL_MEAN && s > 2) ? 1.0f / (float)(s - 2) : 1.0f;
    if (VEC) {
        const int c4n = cols >> 2;
        for (int c4 = threadIdx.x; c4 < c4n; c4 += kPoolThreads) {
            const float4* col = reinterpret_cast<const float4*>(src + r0 * ld) + c4;
            const int64_t ld4 = ld >> 2;
            float4 a = s > 0 ? __ldg(col) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 b = s > 1 ? __ldg(col + ld4) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 2; r < s; ++r) {
                const float4 v = __ldg(col + (int64_t)r * ld4);
                acc.x += v.x;
                acc.y += v.y;
                acc.z += v.z;
                acc.w += v.w;
            }
            acc.x *= inv;
            acc.y *= inv;
            acc.z *= inv;
            acc.w *= inv;
            float4* o4 = reinterpret_cast<float4*>(orow);
            if (layout == S3_POOL_OUT_CENTER) {
                o4[c4] = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
                o4[c4n + c4] = acc;
            } else {
                o4[c4] = a;
                o4[c4n + c4] = b;
                o4[2 * c4n + c4] = acc;
            }
        }
    } else {
        for (int c = threadIdx.x; c < cols; c += kPoolThreads) {
            const float* col = src + r0 * ld + c;
            const float a = s > 0 ? __ldg(col) : 0.f;
            const float b = s > 1 ? __ldg(col + ld) : 0.f;
            float acc = 0.f;
            for (int r = 2; r < s; ++r) acc += __ldg(col + (int64_t)r * ld);
            acc *= inv;
            if (layout == S3_POOL_OUT_CENTER) {
                orow[c] = a * b;
                orow[cols + c] = acc;
            } else {
                orow[c] = a;
                orow[cols + c] = b;
                orow[2 * cols + c] = acc;
            }
        }
    }
}

}  // namespace

cudaError_t launch_segment_pool(const float* src, int64_t ld, int64_t cols, const int64_t* row_ptr, int64_t num_links, int mode,
                                int layout, float* out, int64_t ld_out, cudaStream_t st) {
    if (num_links == 0) return cudaSuccess;
    const bool vec = (cols % 4 == 0) && (ld % 4 == 0) && (ld_out % 4 == 0) && !(reinterpret_cast<uintptr_t>(src) & 15) &&
                     !(reinterpret_cast<uintptr_t>(out) & 15);
    if (vec)
        segment_pool_kernel<true><<<(unsigned)num_links, kPoolThreads, 0, st>>>(src, ld, (int)cols, row_ptr, mode, layout, out, ld_out);
    else
        segment_pool_kernel<false><<<(unsigned)num_links, kPoolThreads, 0, st>>>(src, ld, (int)cols, row_ptr, mode, layout, out, ld_out);
    return cudaGetLastError();
}

}  // namespace s3
