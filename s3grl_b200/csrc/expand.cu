// Link pairing for the flows with data-dependent row counts (PoS Plus), and the assembly of their per-batch
// output pieces into the collated matrices.
//
// The reference precomputes every training positive as (u,v) AND (v,u) (sgrl_link_pred.py:193-204 over PyG's
// train_test_split_edges). For PoS Plus the two records differ in nothing but the order of rows 0
// and 1: the enclosing subgraph (utils.py:53-80) is symmetric in src / dst, and the CCN rows — common neighbours
// or the union of the neighbourhoods, tuned_SIGN.py:228-238 — are listed in ascending local id, which both
// directions share beyond the two seeds. So the host runs the path on ONE link per unordered node pair (the head
// of its chain in s3_pair_links' table) and this file writes every link's rows from its head's record:
//   * pair_heads_kernel: head and direction of every link from the chain table;
//   * scatter_rows_kernel: one CTA per computed record copies its s rows of all K+1 operators to the record's own
//     position in the collated output and to the positions of its chain members (rows 0 / 1 exchanged for the
//     opposite direction). Without a table it is a plain placement of a batch's piece — the step PyG's collate
//     (sgrl_link_pred.py:204) performs on the host, and torch.cat performed here until round 2.
// Pure data movement, HBM-bound; rows are 4-byte aligned (F' = F + 1 is odd), so accesses are 32-bit, one coalesced
// 128-byte line per warp instruction, four rows in flight per thread (cf. loader.cu).
#include "common.cuh"

namespace s3 {
namespace {

constexpr int kScatterThreads = 256;

__global__ void pair_heads_kernel(const long long* __restrict__ mirror, int64_t L, long long* __restrict__ head_code) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    long long m = mirror[i];
    if (m < -1) return;  // a chain member: written by its head's thread
    head_code[i] = 2 * i;
    while (m >= 0) {
        const long long enc = -2 - mirror[m];  // ((next + 1) << 1) | swap, pair.cu
        head_code[m] = 2 * i + (enc & 1);
        m = (enc >> 1) - 1;
    }
}

struct ScatterParams {
    OutPtrs src, dst;
    int num_ops, cols;
    int64_t ld_src, ld_dst;
    const int64_t* __restrict__ src_row_ptr;  // [num_records + 1] rows of the piece
    const int64_t* __restrict__ link_idx;     // [num_records] global link index of every record, or null: link_base + r
    int64_t link_base;
    const long long* __restrict__ mirror;     // chain table over the whole link list, or null
    const int64_t* __restrict__ dst_row_ptr;  // [links + 1] rows of the collated output
    int lead;  // the record's first `lead` rows are written a second time in front of its rows (s3_scatter_rows_lead)
};

// one record's rows at destination row d0: [its first `lead` rows again |] its s rows
__device__ __forceinline__ void place_record(const ScatterParams& p, int64_t s0, int s, int64_t d0, bool swap);

__device__ __forceinline__ void copy_rows(const ScatterParams& p, int64_t s0, int s, int64_t d0, bool swap) {
    const int tid = threadIdx.x, cols = p.cols;
    for (int op = 0; op < p.num_ops; ++op) {
        const float* __restrict__ src = p.src.p[op];
        float* __restrict__ dst = p.dst.p[op];
        int r = 0;
        for (; r + 3 < s; r += 4) {  // four rows in flight
            const float* a[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int q = r + t;
                a[t] = src + (s0 + ((swap && q < 2) ? 1 - q : q)) * p.ld_src;
            }
            float* drow = dst + (d0 + r) * p.ld_dst;
            for (int c = tid; c < cols; c += kScatterThreads) {
                float v[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) v[t] = __ldg(a[t] + c);
#pragma unroll
                for (int t = 0; t < 4; ++t) drow[t * p.ld_dst + c] = v[t];
            }
        }
        for (; r < s; ++r) {
            const float* a = src + (s0 + ((swap && r < 2) ? 1 - r : r)) * p.ld_src;
            float* drow = dst + (d0 + r) * p.ld_dst;
            for (int c = tid; c < cols; c += kScatterThreads) drow[c] = __ldg(a + c);
        }
    }
}

__device__ __forceinline__ void place_record(const ScatterParams& p, int64_t s0, int s, int64_t d0, bool swap) {
    const int lead = min(p.lead, s);
    if (lead > 0) copy_rows(p, s0, lead, d0, swap);
    copy_rows(p, s0, s, d0 + lead, swap);
}

__global__ void __launch_bounds__(kScatterThreads) scatter_rows_kernel(ScatterParams p) {
    const int64_t r = blockIdx.x;
    const int64_t s0 = p.src_row_ptr[r];
    const int s = (int)(p.src_row_ptr[r + 1] - s0);
    if (s <= 0) return;
    const int64_t link = p.link_idx ? p.link_idx[r] : p.link_base + r;
    place_record(p, s0, s, p.dst_row_ptr[link], false);
    if (!p.mirror) return;
    for (long long m = p.mirror[link]; m >= 0;) {
        const long long enc = -2 - p.mirror[m];
        place_record(p, s0, s, p.dst_row_ptr[m], (enc & 1) != 0);
        m = (enc >> 1) - 1;
    }
}

}  // namespace

cudaError_t launch_pair_heads(const int64_t* mirror, int64_t L, int64_t* head_code, cudaStream_t st) {
    if (L == 0) return cudaSuccess;
    pair_heads_kernel<<<(unsigned)((L + 255) / 256), 256, 0, st>>>(reinterpret_cast<const long long*>(mirror), L,
                                                                    reinterpret_cast<long long*>(head_code));
    return cudaGetLastError();
}

cudaError_t launch_scatter_rows(const OutPtrs& src, int64_t ld_src, const int64_t* src_row_ptr, int64_t num_records,
                                const int64_t* link_idx, int64_t link_base, const int64_t* mirror, const int64_t* dst_row_ptr,
                                const OutPtrs& dst, int64_t ld_dst, int num_ops, int64_t cols, cudaStream_t st, int lead) {
    if (num_records == 0) return cudaSuccess;
    if (num_records > 0x7fffffff) return cudaErrorInvalidValue;
    ScatterParams p;
    p.src = src;
    p.dst = dst;
    p.num_ops = num_ops;
    p.cols = (int)cols;
    p.ld_src = ld_src;
    p.ld_dst = ld_dst;
    p.src_row_ptr = src_row_ptr;
    p.link_idx = link_idx;
    p.link_base = link_base;
    p.mirror = reinterpret_cast<const long long*>(mirror);
    p.dst_row_ptr = dst_row_ptr;
    p.lead = lead;
    scatter_rows_kernel<<<(unsigned)num_records, kScatterThreads, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace s3
