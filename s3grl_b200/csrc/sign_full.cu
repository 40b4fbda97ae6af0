// Kernel 4 — the NON-optimised SIGN + SEAL flow: PyG's SIGN transform on the
// WHOLE enclosing subgraph, x_k = S x_{k-1} for ALL n rows, S = D^-1/2 A_sub D^-1/2.
//
// Replaces reference utils.py:497-520 (k_hop_subgraph -> construct_pyg_graph -> TunedSIGN),
// utils.py:281-316 (construct_pyg_graph: the labeling-trick column z prepended to the features),
// utils.py:211-236 (drnl_node_labeling) and tuned_SIGN.py:18-23 (TunedSIGN.__call__ = PyG SIGN:
// unweighted adj_t, deg = row count, inf -> 0, x_k = adj_t @ x_{k-1}).
//
// This is north_star's kernel (2): a segmented, per-subgraph CSR SpMM applied K times. One CTA (512 threads)
// per (record, chunk of 128 output columns), the chunks of a record adjacent in a 1-D grid: operator columns are
// independent, so a CTA carries its columns through all K operators on its own and needs no grid-wide
// synchronisation. Per CTA the subgraph's compact CSR — row starts, (column, dis[column]) pairs, dis per row,
// node ids, built from the front kernel's padded CSR (S3_BATCH_STORE_ALL_ROWS) by one block scan over its slots —
// and x_{k-1} of the current 32 / 64 / 128 columns live in shared memory (112 KB, two CTAs per SM):
//   * two x buffers (ping-pong) when they fit: the inner loop is LDS.64 (pair, broadcast) + LDS + FFMA per
//     neighbour and global memory only sees the streaming stores of the results;
//   * one buffer for larger subgraphs (n <= ~800): x_k is re-loaded from out[k] (coalesced, L2-hot) after a barrier;
//   * read-back of x_{k-1} from out[k-1] through L2 for anything larger.
// One warp per row in work-balanced contiguous row ranges, lanes along the columns, row reduction in registers — no
// atomics, fixed summation order (slots ascending): results are independent of scheduling.
//
// Output rows of record r: [row_base + row_ptr[r], + n), local node j -> row j (canonical order:
// src, dst, then ascending (hop, global id)). Column 0 is the labeling-trick value z_j.
// HBM-write-bound: 4*(K+1)*n*(F+1) bytes written per record against 4*F*n read.
#include "common.cuh"

namespace s3 {
namespace {

constexpr int kFullThreads = 512;
constexpr int kFullCols = 128;             // output columns per CTA
constexpr int kFullSmemBytes = 112 * 1024;  // shared x_{k-1} buffer: two CTAs per SM

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
    int smem_floats;    // capacity of the shared x_{k-1} buffer
    int chunks;         // column chunks (CTAs) per record
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

// One sub-chunk of CPL*32 output columns starting at cs. x_{k-1} and x_k ([n][CPL*32] floats each,
// ping-pong) AND the subgraph's compact CSR (row starts, (column, dis[column]) pairs, dis per row, node
// ids) live in shared memory, so the per-row dependency chain never leaves the SM and global memory
// only sees the streaming stores of the results: operator 0 is copied from X into shared memory and
// out[0] (four rows in flight per warp); operator k is computed row by row (one warp per row; the pair
// is a broadcast read, the neighbour's row a conflict-free one) and written to out[k] and to the other
// shared buffer.
template <int CPL, bool PING>
__device__ __forceinline__ void chain_shared(const FullParams& p, int n, const int* cnode, const int* crow,
                                             const int2* cpair, const float* cdis, const float* z, int64_t row0, int cs,
                                             float* bufA, float* bufB) {
    constexpr int CW = CPL * 32, NWARP = kFullThreads / 32, U = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool ok[CPL];
#pragma unroll
    for (int t = 0; t < CPL; ++t) ok[t] = cs + lane + 32 * t < p.F1;
    __syncthreads();  // the previous sub-chunk is done with the buffers
    for (int jb = warp; jb < n; jb += U * NWARP) {
        float v[U][CPL];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = jb + u * NWARP;
            const float* xr = p.x + (int64_t)cnode[min(j, n - 1)] * p.ldx;
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                const int c = cs + lane + 32 * t;
                v[u][t] = (ok[t] && c > 0) ? __ldg(xr + c - 1) : 0.0f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = jb + u * NWARP;
            if (j < n) {
                float* orow = p.out.p[0] + (row0 + j) * p.ldo + cs + lane;
#pragma unroll
                for (int t = 0; t < CPL; ++t) {
                    const float val = (cs + lane + 32 * t == 0) ? z[j] : v[u][t];
                    bufA[j * CW + lane + 32 * t] = val;
                    if (ok[t]) orow[32 * t] = val;
                }
            }
        }
    }
    __syncthreads();
    float* prev = bufA;
    float* next = bufB;
    // rows are dealt to the warps in contiguous ranges of equal work (neighbours + 1 per row), so a hub
    // row does not leave the other warps waiting at the barrier: boundary b = first row with crow[j] + j >= b * W
    int jlo, jhi;
    {
        const int total = crow[n] + n;
        auto first_row = [&](int target) {
            int lo = 0, hi = n;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (crow[mid] + mid < target) lo = mid + 1; else hi = mid;
            }
            return lo;
        };
        jlo = first_row((int)(((int64_t)total * warp) / NWARP));
        jhi = warp == NWARP - 1 ? n : first_row((int)(((int64_t)total * (warp + 1)) / NWARP));
    }
    for (int k = 1; k <= p.sign_k; ++k) {
        float* cur = p.out.p[k] + row0 * p.ldo + cs + lane;
        const bool keep = k < p.sign_k;
        for (int j = jlo; j < jhi; ++j) {
            float acc[CPL];
#pragma unroll
            for (int t = 0; t < CPL; ++t) acc[t] = 0.0f;
            const int e0 = crow[j], e1 = crow[j + 1];
#pragma unroll 4
            for (int e = e0; e < e1; ++e) {
                const int2 pr = cpair[e];
                const float d = __int_as_float(pr.y);
                const float* r = prev + pr.x * CW + lane;
#pragma unroll
                for (int t = 0; t < CPL; ++t) acc[t] = fmaf(d, r[32 * t], acc[t]);
            }
            const float dj = cdis[j];
            float* orow = cur + (int64_t)j * p.ldo;
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                const float val = dj * acc[t];
                if (PING && keep) next[j * CW + lane + 32 * t] = val;
                if (ok[t]) orow[32 * t] = val;
            }
        }
        __syncthreads();  // x_k complete; nobody reads x_{k-1} any more
        if (PING) {
            float* tmp = prev;
            prev = next;
            next = tmp;
        } else if (keep) {
            // one buffer only (larger subgraphs): x_k is re-loaded from out[k] (coalesced, L2-hot), U rows in flight
            for (int jb = warp; jb < n; jb += U * NWARP) {
                float v[U][CPL];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float* orow = cur + (int64_t)min(jb + u * NWARP, n - 1) * p.ldo;
#pragma unroll
                    for (int t = 0; t < CPL; ++t) v[u][t] = ok[t] ? orow[32 * t] : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = jb + u * NWARP;
                    if (j < n) {
#pragma unroll
                        for (int t = 0; t < CPL; ++t) prev[j * CW + lane + 32 * t] = v[u][t];
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Subgraphs too large for shared memory: x_{k-1} is read back from the operator matrix the CTA has
// just written (L2), 128 columns per CTA, two neighbours in flight per warp.
__device__ __forceinline__ void chain_global(const FullParams& p, int n, const int32_t* nodes, const int32_t* rowptr,
                                             const int32_t* lcol, const float* dis, const float* z, int64_t row0, int c0) {
    constexpr int NWARP = kFullThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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

    for (int k = 1; k <= p.sign_k; ++k) {
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

__global__ void __launch_bounds__(kFullThreads) sign_full_kernel(FullParams p) {
    extern __shared__ __align__(16) float s_dyn[];
    __shared__ int s_flag;
    __shared__ int s_scan[33];
    const int tid = threadIdx.x;
    // 1-D grid, column chunks of a record adjacent: they run at the same time, so the cache lines they
    // share at chunk boundaries are completed in L2 and the record's X rows / CSR are fetched once
    const int ridx = blockIdx.x / p.chunks, chunk = blockIdx.x - ridx * p.chunks;
    const int32_t rec = p.order ? p.order[ridx] : (int32_t)ridx;
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
    const int c0 = chunk * kFullCols;

    for (int j = tid; j < n; j += kFullThreads) {
        const int deg = rowlen[j];
        dis[j] = deg > 0 ? 1.0f / sqrtf((float)deg) : 0.0f;  // PyG SIGN: deg.pow(-0.5), inf -> 0
    }
    if (chunk == 0) {
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

    // shared memory: [row starts n+1 (+pad) | (col, dis) pairs 2m | dis n | node ids n | x ping | x pong]
    const int m = cnt[S3_CNT_M];
    const int64_t idx_words = ((n + 2) & ~1) + 2 * (int64_t)m + 2 * (int64_t)n;
    const int64_t S1 = (int64_t)p.smem_floats - idx_words;  // floats for the x buffer(s)
    if ((int64_t)n * 32 <= S1) {
        int* crow = reinterpret_cast<int*>(s_dyn);
        int2* cpair = reinterpret_cast<int2*>(crow + ((n + 2) & ~1));
        float* cdis = reinterpret_cast<float*>(cpair + m);
        int* cnode = reinterpret_cast<int*>(cdis + n);
        float* bufA = reinterpret_cast<float*>(cnode + n);
        // compact row starts: exclusive scan of the induced degrees, tile by tile
        int running = 0;
        for (int base = 0; base < n; base += kFullThreads) {
            const int j = base + tid;
            const int d = j < n ? rowlen[j] : 0;
            int tile_total;
            const int ex = block_exclusive_scan(d, s_scan, &tile_total);
            if (j < n) {
                crow[j] = running + ex;
                cdis[j] = d > 0 ? 1.0f / sqrtf((float)d) : 0.0f;
                cnode[j] = nodes[j];
            }
            running += tile_total;
            __syncthreads();
        }
        if (tid == 0) crow[n] = running;
        // compaction of the padded slots (holes dropped): slot order == compact order, so the position of
        // a kept slot is the exclusive scan of the keep flags; one coalesced pass, dis[col] folded in
        const int Dslots = rowptr[n];
        running = 0;
        for (int base = 0; base < Dslots; base += kFullThreads) {
            const int e = base + tid;
            const int col = e < Dslots ? lcol[e] : -1;
            int tile_total;
            const int ex = block_exclusive_scan(col >= 0 ? 1 : 0, s_scan, &tile_total);
            if (col >= 0) {
                const int dc = rowlen[col];
                cpair[running + ex] = make_int2(col, __float_as_int(dc > 0 ? 1.0f / sqrtf((float)dc) : 0.0f));
            }
            running += tile_total;
            __syncthreads();
        }
        // two x buffers (ping-pong) at the widest sub-chunk that fits; one buffer + re-load for larger subgraphs
        const bool ping = (int64_t)n * 64 <= S1;
        const int64_t S = ping ? S1 / 2 : S1;
        const int cw = (int64_t)n * 128 <= S ? 128 : ((int64_t)n * 64 <= S ? 64 : 32);
        float* bufB = bufA + (int64_t)n * cw;
        for (int cs = c0; cs < min(c0 + kFullCols, p.F1); cs += cw) {
            if (!ping) chain_shared<1, false>(p, n, cnode, crow, cpair, cdis, z, row0, cs, bufA, bufA);
            else if (cw == 128) chain_shared<4, true>(p, n, cnode, crow, cpair, cdis, z, row0, cs, bufA, bufB);
            else if (cw == 64) chain_shared<2, true>(p, n, cnode, crow, cpair, cdis, z, row0, cs, bufA, bufB);
            else chain_shared<1, true>(p, n, cnode, crow, cpair, cdis, z, row0, cs, bufA, bufB);
        }
    } else {
        chain_global(p, n, nodes, rowptr, lcol, dis, z, row0, c0);
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
    static LaunchCache cache;  // the shared-memory opt-in is per device
    cudaError_t e = cache.get(reinterpret_cast<const void*>(sign_full_kernel), kFullThreads, kFullSmemBytes, nullptr, nullptr);
    if (e != cudaSuccess) return e;
    p.smem_floats = kFullSmemBytes / 4;
    p.chunks = (p.F1 + kFullCols - 1) / kFullCols;
    if (num_records * p.chunks > 0x7fffffff) return cudaErrorInvalidValue;
    sign_full_kernel<<<(unsigned)(num_records * p.chunks), kFullThreads, kFullSmemBytes, st>>>(p);
    return cudaGetLastError();
}

}  // namespace s3
