// Kernel 4 — the NON-optimised SIGN + SEAL flow: PyG's SIGN transform on the
// WHOLE enclosing subgraph, x_k = S x_{k-1} for ALL n rows, S = D^-1/2 A_sub D^-1/2.
//
// Replaces reference utils.py:497-520 (k_hop_subgraph -> construct_pyg_graph -> TunedSIGN),
// utils.py:281-316 (construct_pyg_graph: the labeling-trick column z prepended to the features),
// utils.py:211-236 (drnl_node_labeling) and tuned_SIGN.py:18-23 (TunedSIGN.__call__ = PyG SIGN:
// unweighted adj_t, deg = row count, inf -> 0, x_k = adj_t @ x_{k-1}).
//
// This is north_star's kernel (2): a segmented, per-subgraph CSR SpMM applied K times. One CTA per
// (record, chunk of 128 output columns): operator columns are independent, so a CTA carries its
// column chunk through all K operators on its own and needs no grid-wide synchronisation.
// x_{k-1} is read back from the operator matrix the same CTA has just written (L1/L2 resident; a
// __syncthreads() orders the global writes within the CTA), one warp per subgraph row, lanes along
// the columns: every neighbour costs four coalesced 128-byte loads per warp and the row reduction
// stays in registers — no atomics, fixed summation order (slots ascending), results independent
// of scheduling. The padded local CSR of the front kernel (S3_BATCH_STORE_ALL_ROWS) is consumed as
// is: 32 slots per step, holes skipped through a ballot.
//
// Output rows of record r: [row_base + row_ptr[r], + n), local node j -> row j (canonical order:
// src, dst, then ascending (hop, global id)). Column 0 is the labeling-trick value z_j.
// HBM-write-bound: 4*(K+1)*n*(F+1) bytes written per record against 4*F*n read.
#include "common.cuh"

namespace s3 {
namespace {

constexpr int kFullThreads = 256;
constexpr int kFullCols = 128;  // output columns per CTA: 32 lanes x 4

struct FullParams {
    const float* __restrict__ x;
    int64_t ldx;
    int F1;  // F + 1 output columns
    int32_t* arena;
    const int64_t* __restrict__ off;
    const int32_t* __restrict__ cnt;
    const int64_t* __restrict__ row_ptr;
    const int32_t* __restrict__ order;  // may be null
    int sign_k, label;
    OutPtrs out;
    int64_t ldo, row_base;
    int64_t* node_out;  // may be null: global id of every output row
};

// DRNL (utils.py:211-236): z = 1 + min(ds, dd) + (d/2)*((d/2) + d%2 - 1), d = ds + dd, where ds is
// the distance to local 0 in the subgraph WITHOUT local 1 and dd the distance to local 1 WITHOUT
// local 0; z[0] = z[1] = 1; unreachable -> 0.  Level-synchronous pull BFS over the padded CSR, one
// thread per row, both searches in the same sweep.
__device__ void drnl_labels(int n, const int32_t* rowptr, const int32_t* lcol, int* ds, int* dd, float* z, int* s_flag) {
    const int tid = threadIdx.x, T = blockDim.x;
    constexpr int INF = 0x3fffffff;
    for (int j = tid; j < n; j += T) {
        ds[j] = j == 0 ? 0 : INF;
        dd[j] = j == 1 ? 0 : INF;
    }
    __syncthreads();
    for (int level = 1; level < n; ++level) {
        if (tid == 0) *s_flag = 0;
        __syncthreads();
        bool changed = false;
        for (int j = tid; j < n; j += T) {
            const bool need_s = j != 1 && ds[j] == INF, need_d = j != 0 && dd[j] == INF;
            if (!need_s && !need_d) continue;
            bool hit_s = false, hit_d = false;
            for (int e = rowptr[j]; e < rowptr[j + 1]; ++e) {
                const int i = lcol[e];
                if (i < 0) continue;
                if (i != 1 && ds[i] == level - 1) hit_s = true;
                if (i != 0 && dd[i] == level - 1) hit_d = true;
            }
            if (need_s && hit_s) {
                ds[j] = level;
                changed = true;
            }
            if (need_d && hit_d) {
                dd[j] = level;
                changed = true;
            }
        }
        if (changed) *s_flag = 1;
        __syncthreads();
        const bool any = *s_flag != 0;
        __syncthreads();
        if (!any) break;
    }
    for (int j = tid; j < n; j += T) {
        float v;
        if (j < 2) {
            v = 1.0f;
        } else if (ds[j] == INF || dd[j] == INF) {
            v = 0.0f;
        } else {
            const int d = ds[j] + dd[j], h2 = d / 2;
            v = (float)(1 + min(ds[j], dd[j]) + h2 * (h2 + (d & 1) - 1));
        }
        z[j] = v;
    }
}

__global__ void __launch_bounds__(kFullThreads) sign_full_kernel(FullParams p) {
    __shared__ int s_flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = kFullThreads / 32;
    const int32_t rec = p.order ? p.order[blockIdx.x] : (int32_t)blockIdx.x;
    if (rec < 0) return;
    const int32_t* cnt = p.cnt + (int64_t)rec * S3_NCNT;
    if (cnt[S3_CNT_STATUS] != S3_REC_OK) return;
    const int n = cnt[S3_CNT_N];
    const int64_t* off = p.off + (int64_t)rec * S3_NOFF;
    const int32_t* nodes = p.arena + off[S3_OFF_NODES];
    const int32_t* rowptr = p.arena + off[S3_OFF_ROWPTR];
    const int32_t* rowlen = p.arena + off[S3_OFF_ROWLEN];
    const int32_t* lcol = p.arena + off[S3_OFF_LCOL];
    // float scratch of the record (the front kernel's weights are not needed by this flow):
    // [dis n | z n | ds n | dd n]; every column-chunk CTA of the record writes the SAME values.
    const int K = p.sign_k;
    const int NWP = (2 * (K + 1) + 3) & ~3;
    float* scratch = reinterpret_cast<float*>(p.arena + off[S3_OFF_F32]) + NWP;
    float* dis = scratch;
    float* z = scratch + n;
    const int64_t row0 = p.row_base + p.row_ptr[rec];
    const int c0 = blockIdx.y * kFullCols;

    for (int j = tid; j < n; j += kFullThreads) {
        const int deg = rowlen[j];
        dis[j] = deg > 0 ? 1.0f / sqrtf((float)deg) : 0.0f;  // PyG SIGN: deg.pow(-0.5), inf -> 0
    }
    if (blockIdx.y == 0) {
        // labeling-trick column (construct_pyg_graph, utils.py:296-310)
        if (p.label == S3_LABEL_DRNL) {
            drnl_labels(n, rowptr, lcol, reinterpret_cast<int*>(scratch + 2 * (int64_t)n),
                        reinterpret_cast<int*>(scratch + 3 * (int64_t)n), z, &s_flag);
        } else {
            int hop_end[S3_MAX_HOPS + 1];
            int acc = 0;
#pragma unroll
            for (int l = 0; l <= S3_MAX_HOPS; ++l) {
                acc += cnt[S3_CNT_HOP0 + l];
                hop_end[l] = acc;
            }
            for (int j = tid; j < n; j += kFullThreads) {
                float v = 0.0f;
                if (p.label == S3_LABEL_ZO) {
                    v = j < 2 ? 1.0f : 0.0f;  // (dists == 0)
                } else if (p.label == S3_LABEL_HOP) {
                    int hop = 0;
#pragma unroll
                    for (int l = 0; l < S3_MAX_HOPS; ++l) hop += j >= hop_end[l] ? 1 : 0;
                    v = (float)hop;
                } else if (p.label == S3_LABEL_DEGREE) {
                    v = (float)min(rowlen[j], 100);  // adj.sum(axis=0), capped at 100
                }
                z[j] = v;
            }
        }
        if (p.node_out)
            for (int j = tid; j < n; j += kFullThreads) p.node_out[row0 + j] = nodes[j];
    }
    __syncthreads();

    // operator 0: x = [z | X[nodes]] (exact copy)
    {
        float* o0 = p.out.p[0];
        for (int j = warp; j < n; j += NWARP) {
            const float* xr = p.x + (int64_t)nodes[j] * p.ldx;
            float* orow = o0 + (row0 + j) * p.ldo;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int c = c0 + lane + 32 * t;
                if (c < p.F1) orow[c] = c == 0 ? z[j] : xr[c - 1];
            }
        }
    }
    __syncthreads();

    bool ok[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) ok[t] = c0 + lane + 32 * t < p.F1;

    for (int k = 1; k <= K; ++k) {
        const float* prev = p.out.p[k - 1] + row0 * p.ldo + c0 + lane;
        float* cur = p.out.p[k] + row0 * p.ldo + c0 + lane;
        for (int j = warp; j < n; j += NWARP) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const int e0 = rowptr[j], e1 = rowptr[j + 1];
            for (int eb = e0; eb < e1; eb += 32) {
                const int e = eb + lane;
                const int mine = e < e1 ? lcol[e] : -1;
                const float mydis = mine >= 0 ? dis[mine] : 0.0f;
                unsigned live = __ballot_sync(0xffffffffu, mine >= 0);
                while (live) {
                    const int a = __ffs(live) - 1;
                    live &= live - 1;
                    const int i0 = __shfl_sync(0xffffffffu, mine, a);
                    const float d0 = __shfl_sync(0xffffffffu, mydis, a);
                    int i1 = i0;
                    float d1 = 0.0f;
                    if (live) {  // warp-uniform
                        const int b = __ffs(live) - 1;
                        live &= live - 1;
                        i1 = __shfl_sync(0xffffffffu, mine, b);
                        d1 = __shfl_sync(0xffffffffu, mydis, b);
                    }
                    const float* r0 = prev + (int64_t)i0 * p.ldo;
                    const float* r1 = prev + (int64_t)i1 * p.ldo;
                    float v0[4], v1[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        v0[t] = ok[t] ? r0[32 * t] : 0.0f;
                        v1[t] = ok[t] ? r1[32 * t] : 0.0f;
                    }
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        acc[t] = fmaf(d0, v0[t], acc[t]);
                        acc[t] = fmaf(d1, v1[t], acc[t]);
                    }
                }
            }
            const float dj = dis[j];
            float* orow = cur + (int64_t)j * p.ldo;
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (ok[t]) orow[32 * t] = dj * acc[t];
        }
        __syncthreads();  // x_k of this column chunk is complete and visible to the whole CTA
    }
}

// row_ptr[r] = exclusive scan of n over the records (0 rows for a failed record); total -> S3_CTR_ROWS
__global__ void __launch_bounds__(1024) plan_full_kernel(const int32_t* __restrict__ cnt, int64_t num_records,
                                                         int64_t* __restrict__ row_ptr, unsigned long long* counters) {
    __shared__ long long s_rows[1024];
    const int tid = threadIdx.x, T = blockDim.x;
    const int64_t per = (num_records + T - 1) / T;
    const int64_t r0 = min(num_records, (int64_t)tid * per), r1 = min(num_records, r0 + per);
    auto rows_of = [&](int64_t r) {
        return cnt[r * S3_NCNT + S3_CNT_STATUS] == S3_REC_OK ? (long long)cnt[r * S3_NCNT + S3_CNT_N] : 0ll;
    };
    long long rows = 0;
    for (int64_t r = r0; r < r1; ++r) rows += rows_of(r);
    s_rows[tid] = rows;
    __syncthreads();
    for (int d = 1; d < T; d <<= 1) {
        long long a = tid >= d ? s_rows[tid - d] : 0;
        __syncthreads();
        s_rows[tid] += a;
        __syncthreads();
    }
    long long run = s_rows[tid] - rows;
    for (int64_t r = r0; r < r1; ++r) {
        row_ptr[r] = run;
        run += rows_of(r);
    }
    if (tid == T - 1) {
        row_ptr[num_records] = s_rows[tid];
        counters[S3_CTR_ROWS] = (unsigned long long)s_rows[tid];
    }
}

}  // namespace

cudaError_t launch_plan_full(const s3_batch& b, cudaStream_t st) {
    plan_full_kernel<<<1, 1024, 0, st>>>(b.cnt, s3_num_records(&b), b.row_ptr, reinterpret_cast<unsigned long long*>(b.counters));
    return cudaGetLastError();
}

cudaError_t launch_sign_full(const s3_graph& g, const s3_batch& b, int64_t num_records, int label, const OutPtrs& out,
                             int64_t ldo, int64_t row_base, int64_t* node_out, cudaStream_t st) {
    if (num_records == 0) return cudaSuccess;
    FullParams p;
    p.x = g.x;
    p.ldx = g.ldx;
    p.F1 = (int)g.num_feat + 1;
    p.arena = b.arena;
    p.off = b.off;
    p.cnt = b.cnt;
    p.row_ptr = b.row_ptr;
    p.order = b.order;
    p.sign_k = b.sign_k;
    p.label = label;
    p.out = out;
    p.ldo = ldo;
    p.row_base = row_base;
    p.node_out = node_out;
    dim3 grid((unsigned)num_records, (unsigned)((p.F1 + kFullCols - 1) / kFullCols));
    sign_full_kernel<<<grid, kFullThreads, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace s3
