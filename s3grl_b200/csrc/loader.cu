// Kernel 5 — GPU-resident batch assembly: replaces, for the SIGN flows, PyG's
// DataLoader(dataset, batch_size, shuffle, follow_batch=[x1..xK]) + Batch.from_data_list (reference
// sgrl_link_pred.py:1253-1269) and the feature-wise concat at the top of SIGNNet.forward
// (models.py:372, xs_cat = torch.cat(xs, dim=-1)).
//
// Input: the collated dataset as the precompute path leaves it in HBM — K+1 row-stacked operator
// matrices [R, F'] and row_ptr [L+1] — and a list of link indices (one batch, or a whole shuffled
// epoch). Output: the JOINT matrix of those links, [R_out, (K+1)*F'], row r = [x | x1 | .. | xK] of one
// selected row, links in list order, plus the `batch` vector PyG would build (position of the link in
// the list, repeated once per row; every x{k}_batch of follow_batch is this same vector because all
// operators of a link have the same number of rows — SURVEY §8a row 10b).
// Pure data movement, HBM-bound: 2 * 4 * R_out * (K+1) * F' bytes. One CTA per output link; the
// (row, operator, column) space of the link is flattened and strided by the CTA with four independent
// loads in flight per thread. Rows are only 4-byte aligned (F' = F + 1 is odd for every dataset of the
// reference), so accesses are 32-bit, one coalesced 128-byte line per warp instruction.
#include "common.cuh"

namespace s3 {
namespace {

constexpr int kJointThreads = 256;

struct JointParams {
    OutPtrs src;
    int num_ops, cols;
    int64_t ld_src;
    const int64_t* __restrict__ row_ptr;
    const int64_t* __restrict__ link_idx;
    const int64_t* __restrict__ out_row_ptr;  // may be null: rows_per_link rows per link
    int rows_per_link;
    float* __restrict__ dst;
    int64_t ld_dst;
    int64_t* __restrict__ batch_vec;  // may be null
};

// NOPS > 0: operator count known at compile time (all loads of a column position in flight at once);
// NOPS == 0: run-time count (hybrid flows with many operators).
template <int NOPS>
__global__ void __launch_bounds__(kJointThreads) joint_rows_kernel(JointParams p) {
    const int64_t b = blockIdx.x;
    const int64_t link = p.link_idx[b];
    const int64_t r0 = p.row_ptr[link];
    const int s = (int)(p.row_ptr[link + 1] - r0);
    const int64_t o0 = p.out_row_ptr ? p.out_row_ptr[b] : b * p.rows_per_link;
    const int tid = threadIdx.x;
    if (p.batch_vec)
        for (int r = tid; r < s; r += kJointThreads) p.batch_vec[o0 + r] = b;
    const int cols = p.cols;
    int r = 0;
    if (NOPS > 0) {
        // two rows at a time: 2 * NOPS independent loads in flight per thread
        for (; r + 1 < s; r += 2) {
            const int64_t soff = (r0 + r) * p.ld_src;
            float* drow = p.dst + (o0 + r) * p.ld_dst;
            for (int c = tid; c < cols; c += kJointThreads) {
                float v[2][NOPS > 0 ? NOPS : 1];
#pragma unroll
                for (int op = 0; op < NOPS; ++op) {
                    v[0][op] = __ldg(p.src.p[op] + soff + c);
                    v[1][op] = __ldg(p.src.p[op] + soff + p.ld_src + c);
                }
#pragma unroll
                for (int op = 0; op < NOPS; ++op) {
                    drow[op * cols + c] = v[0][op];
                    drow[p.ld_dst + op * cols + c] = v[1][op];
                }
            }
        }
    }
    for (; r < s; ++r) {
        const int64_t soff = (r0 + r) * p.ld_src;
        float* drow = p.dst + (o0 + r) * p.ld_dst;
        for (int c = tid; c < cols; c += kJointThreads) {
            if (NOPS > 0) {
                float v[NOPS > 0 ? NOPS : 1];
#pragma unroll
                for (int op = 0; op < NOPS; ++op) v[op] = __ldg(p.src.p[op] + soff + c);
#pragma unroll
                for (int op = 0; op < NOPS; ++op) drow[op * cols + c] = v[op];
            } else {
                for (int op = 0; op < p.num_ops; ++op) drow[op * cols + c] = __ldg(p.src.p[op] + soff + c);
            }
        }
    }
}

}  // namespace

cudaError_t launch_joint_rows(const OutPtrs& src, int num_ops, int64_t cols, int64_t ld_src, const int64_t* row_ptr,
                              const int64_t* link_idx, int64_t num_links, const int64_t* out_row_ptr, int rows_per_link,
                              float* dst, int64_t ld_dst, int64_t* batch_vec, cudaStream_t st) {
    if (num_links == 0) return cudaSuccess;
    if (num_links > 0x7fffffff) return cudaErrorInvalidValue;
    JointParams p;
    p.src = src;
    p.num_ops = num_ops;
    p.cols = (int)cols;
    p.ld_src = ld_src;
    p.row_ptr = row_ptr;
    p.link_idx = link_idx;
    p.out_row_ptr = out_row_ptr;
    p.rows_per_link = rows_per_link;
    p.dst = dst;
    p.ld_dst = ld_dst;
    p.batch_vec = batch_vec;
    const unsigned grid = (unsigned)num_links;
    switch (num_ops) {
        case 2: joint_rows_kernel<2><<<grid, kJointThreads, 0, st>>>(p); break;
        case 3: joint_rows_kernel<3><<<grid, kJointThreads, 0, st>>>(p); break;
        case 4: joint_rows_kernel<4><<<grid, kJointThreads, 0, st>>>(p); break;
        case 5: joint_rows_kernel<5><<<grid, kJointThreads, 0, st>>>(p); break;
        case 6: joint_rows_kernel<6><<<grid, kJointThreads, 0, st>>>(p); break;
        case 8: joint_rows_kernel<8><<<grid, kJointThreads, 0, st>>>(p); break;
        default: joint_rows_kernel<0><<<grid, kJointThreads, 0, st>>>(p); break;
    }
    return cudaGetLastError();
}

}  // namespace s3
