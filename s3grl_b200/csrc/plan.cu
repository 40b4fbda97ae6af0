// Plan: output-row offsets and work-item lists from the per-record selected-row counts —
// the `slices` PyG's collate records for x, x1..xK (reference sgrl_link_pred.py:204,
// SURVEY.md §8a row 10b) — plus the canonical edge-list dump used by the parity tests.
#include "common.cuh"

namespace s3 {
namespace {

// Single CTA: exclusive scans of s (rows) and of the CCN work items ceil((s - seeds) / 8) over all records.
__device__ __forceinline__ int record_items(const int32_t* c, int nseed, int cr, int flags, int strategy, int K) {
    const int s = c[S3_CNT_S];
    const int n1 = c[S3_CNT_HOP0] + c[S3_CNT_HOP0 + 1];
    if (chain_eligible(flags, strategy, c[S3_CNT_N], c[S3_CNT_M], n1)) return 0;
    if (chain_spill_eligible(flags, strategy, c[S3_CNT_N], c[S3_CNT_M], n1, s, K)) return 0;
    return ccn_items(s, nseed, cr);
}

__global__ void __launch_bounds__(1024) plan_scan_kernel(const int32_t* __restrict__ cnt, int64_t num_records, int nseed, int cr,
                                                         int flags, int strategy, int K,
                                                         int64_t* __restrict__ row_ptr, int64_t* __restrict__ item_ptr,
                                                         unsigned long long* counters) {
    __shared__ long long s_rows[1024], s_items[1024];
    const int tid = threadIdx.x, T = blockDim.x;
    const int64_t per = (num_records + T - 1) / T;
    const int64_t r0 = min(num_records, (int64_t)tid * per), r1 = min(num_records, r0 + per);
    long long rows = 0, items = 0;
    for (int64_t r = r0; r < r1; ++r) {
        const int s = cnt[r * S3_NCNT + S3_CNT_S];
        rows += s;
        items += record_items(cnt + r * S3_NCNT, nseed, cr, flags, strategy, K);
    }
    s_rows[tid] = rows;
    s_items[tid] = items;
    __syncthreads();
    // Hillis-Steele over 1024 partials (tiny; runs once per batch)
    for (int d = 1; d < T; d <<= 1) {
        long long a = 0, b = 0;
        if (tid >= d) {
            a = s_rows[tid - d];
            b = s_items[tid - d];
        }
        __syncthreads();
        s_rows[tid] += a;
        s_items[tid] += b;
        __syncthreads();
    }
    long long row_run = s_rows[tid] - rows, item_run = s_items[tid] - items;
    for (int64_t r = r0; r < r1; ++r) {
        const int s = cnt[r * S3_NCNT + S3_CNT_S];
        row_ptr[r] = row_run;
        item_ptr[r] = item_run;
        row_run += s;
        item_run += record_items(cnt + r * S3_NCNT, nseed, cr, flags, strategy, K);
    }
    if (tid == T - 1) {
        row_ptr[num_records] = s_rows[tid];
        item_ptr[num_records] = s_items[tid];
        counters[S3_CTR_ROWS] = (unsigned long long)s_rows[tid];
        counters[S3_CTR_ITEMS] = (unsigned long long)s_items[tid];
    }
}

__global__ void plan_items_kernel(const int64_t* __restrict__ item_ptr, int64_t num_records, int32_t* __restrict__ item_rec) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= num_records) return;
    for (int64_t i = item_ptr[r]; i < item_ptr[r + 1]; ++i) item_rec[i] = (int32_t)r;
}

// edges_out[2*e], edges_out[2*e+1] = global (row, col) of local edge e, in local CSR order.
__global__ void dump_edges_kernel(const int32_t* __restrict__ arena, const int64_t* __restrict__ off,
                                  const int32_t* __restrict__ cnt, const int64_t* __restrict__ edge_ptr,
                                  int32_t* __restrict__ edges_out) {
    const int64_t rec = blockIdx.x;
    if (cnt[rec * S3_NCNT + S3_CNT_STATUS] != S3_REC_OK) return;
    const int n = cnt[rec * S3_NCNT + S3_CNT_NSTORE];  // == cnt[S3_CNT_N] with S3_BATCH_STORE_ALL_ROWS
    const int32_t* nodes = arena + off[rec * S3_NOFF + S3_OFF_NODES];
    const int32_t* rowptr = arena + off[rec * S3_NOFF + S3_OFF_ROWPTR];
    const int32_t* rowlen = arena + off[rec * S3_NOFF + S3_OFF_ROWLEN];
    const int32_t* lcol = arena + off[rec * S3_NOFF + S3_OFF_LCOL];
    int32_t* out = edges_out + 2 * edge_ptr[rec];
    // compact position of row j = sum of rowlen[0..j): serial prefix by one thread per tile
    __shared__ int s_run;
    __shared__ int s_pos[128];
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int j = base + threadIdx.x;
        if (threadIdx.x == 0) {
            int run = s_run;
            for (int t = 0; t < (int)blockDim.x && base + t < n; ++t) {
                s_pos[t] = run;
                run += rowlen[base + t];
            }
            s_run = run;
        }
        __syncthreads();
        if (j < n) {
            const int g = nodes[j];
            const int pos = s_pos[threadIdx.x];
            int k = 0;
            for (int e = rowptr[j]; e < rowptr[j + 1]; ++e) {
                const int c = lcol[e];
                if (c < 0) continue;  // hole of the padded CSR
                out[2 * (int64_t)(pos + k)] = g;
                out[2 * (int64_t)(pos + k) + 1] = nodes[c];
                ++k;
            }
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t launch_plan(const s3_batch& b, cudaStream_t st) {
    const int64_t R = s3_num_records(&b);
    plan_scan_kernel<<<1, 1024, 0, st>>>(b.cnt, R, num_seeds(b.flow), ccn_rows(b.strategy), b.flags, b.strategy, b.sign_k, b.row_ptr, b.item_ptr,
                                         reinterpret_cast<unsigned long long*>(b.counters));
    return cudaGetLastError();
}

cudaError_t launch_plan_items(const s3_batch& b, cudaStream_t st) {
    const int64_t R = s3_num_records(&b);
    if (R == 0) return cudaSuccess;
    plan_items_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(b.item_ptr, R, b.item_rec);
    return cudaGetLastError();
}

cudaError_t launch_dump_edges(const s3_batch& b, const int64_t* edge_ptr, int32_t* edges_out, cudaStream_t st) {
    const int64_t R = s3_num_records(&b);
    if (R == 0) return cudaSuccess;
    dump_edges_kernel<<<(unsigned)R, 128, 0, st>>>(b.arena, b.off, b.cnt, edge_ptr, edges_out);
    return cudaGetLastError();
}

}  // namespace s3
