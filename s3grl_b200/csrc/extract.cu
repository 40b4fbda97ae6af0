// Kernel 1 — enclosing-subgraph extraction, bitmap tier (graphs whose node-id bitmaps fit in
// shared memory: every dataset bundled with the reference; N up to a few hundred thousand).
//
// Replaces, per record: reference utils.py:33-44 (neighbors), utils.py:53-74 (BFS),
// utils.py:76 (A[nodes,:][:,nodes]), utils.py:79-80 (target-link mask), the ssp.find at
// tuned_SIGN.py:153/:208 and the PoS-Plus row selection tuned_SIGN.py:228-238.
//
// One CTA per record. Shared memory holds a visited bitmap V, one bitmap per hop level and a
// per-word exclusive popcount prefix per level. A level bitmap enumerated in word order IS the
// level sorted by global id, so the canonical order [seeds, then (hop, global id) ascending]
// needs no sort, and  local_id(g) = level_base + prefix[word(g)] + popc(bits below g)  needs
// no hash table. Results go to a bump-allocated arena in global memory (count -> allocate ->
// fill, one atomicAdd per allocation); placement depends on scheduling, contents do not.
#include "common.cuh"

namespace s3 {
namespace {

struct ExtractParams {
    const int64_t* __restrict__ indptr;
    const int32_t* __restrict__ indices;
    int64_t num_nodes;
    const int64_t* __restrict__ link_src;
    const int64_t* __restrict__ link_dst;
    int64_t num_records;
    int flow, strategy, radius, sign_k;
    int W;  // bitmap words
    int32_t* arena;
    int64_t arena_words;
    int64_t* off;
    int32_t* cnt;
    unsigned long long* counters;
};

__device__ __forceinline__ bool test_bit(const uint32_t* bm, int g) { return (bm[g >> 5] >> (g & 31)) & 1u; }

// local id of global node g, which must be in the subgraph
__device__ __forceinline__ int local_id(int g, int s0, int s1, int nseed, const uint32_t* Lb, const uint32_t* pre,
                                        const int* lvl_base, int nlev, int W) {
    if (g == s0) return 0;
    if (nseed == 2 && g == s1) return 1;
    const int w = g >> 5;
    const uint32_t below = (1u << (g & 31)) - 1u;
    for (int l = 0; l < nlev; ++l) {
        const uint32_t bits = Lb[l * W + w];
        if ((bits >> (g & 31)) & 1u) return lvl_base[l] + (int)pre[l * W + w] + __popc(bits & below);
    }
    return -1;  // unreachable for members of V
}

__global__ void __launch_bounds__(kExtractThreads) extract_bitmap_kernel(ExtractParams p) {
    extern __shared__ uint32_t sm[];
    const int W = p.W, h = p.radius, T = blockDim.x, tid = threadIdx.x;
    uint32_t* V = sm;
    uint32_t* Lb = V + W;              // [h][W]
    uint32_t* pre = Lb + (size_t)h * W;  // [h][W]
    __shared__ int s_scan[33];
    __shared__ int s_lvl_base[S3_MAX_HOPS + 2];
    __shared__ int s_lvl_cnt[S3_MAX_HOPS + 1];
    __shared__ long long s_base;
    __shared__ unsigned long long s_sumdeg;

    const int64_t rec = blockIdx.x;
    const int nseed = num_seeds(p.flow);
    const bool mask_target = p.flow == S3_FLOW_POS;
    int64_t a, b;
    if (p.flow == S3_FLOW_POS) {
        a = p.link_src[rec];
        b = p.link_dst[rec];
    } else {
        const int64_t l = rec >> 1;
        a = (rec & 1) ? p.link_dst[l] : p.link_src[l];
        b = (rec & 1) ? p.link_src[l] : p.link_dst[l];
    }
    int32_t* cnt = p.cnt + rec * S3_NCNT;
    int64_t* off = p.off + rec * S3_NOFF;
    if (a < 0 || b < 0 || a >= p.num_nodes || b >= p.num_nodes || a == b) {
        if (tid == 0) {
            for (int i = 0; i < S3_NCNT; ++i) cnt[i] = 0;
            for (int i = 0; i < S3_NOFF; ++i) off[i] = 0;
            cnt[S3_CNT_STATUS] = S3_REC_BAD_LINK;
            cnt[S3_CNT_PARTNER] = -1;
            atomicAdd(&p.counters[S3_CTR_ERRORS], 1ull);
        }
        return;
    }
    const int s0 = (int)a, s1 = (int)b;  // s1 is the second seed (PoS) or the partner (SoP)

    for (int i = tid; i < (1 + 2 * h) * W; i += T) sm[i] = 0u;
    if (tid == 0) s_sumdeg = 0ull;
    __syncthreads();
    if (tid == 0) {
        V[s0 >> 5] |= 1u << (s0 & 31);
        if (nseed == 2) V[s1 >> 5] |= 1u << (s1 & 31);
        s_lvl_cnt[0] = nseed;
    }
    __syncthreads();

    // ---------------- BFS: h rounds of frontier expansion (utils.py:57-74) ----------------
    const int chunk = (W + T - 1) / T;
    const int w0 = min(W, tid * chunk), w1 = min(W, w0 + chunk);
    int nlev = 0, n = nseed;
    unsigned long long my_deg = 0;  // sum of global degrees of this thread's subgraph nodes (D of SURVEY 8d)
    for (int l = 0; l < h; ++l) {
        uint32_t* cur = Lb + (size_t)l * W;
        if (l == 0) {
            // seeds: one warp per seed, lanes stride the adjacency list
            const int wid = tid >> 5, lane = tid & 31;
            if (wid < nseed) {
                const int g = wid == 0 ? s0 : s1;
                const int64_t e0 = p.indptr[g], e1 = p.indptr[g + 1];
                for (int64_t e = e0 + lane; e < e1; e += 32) {
                    const int c = p.indices[e];
                    if (!test_bit(V, c)) atomicOr(&cur[c >> 5], 1u << (c & 31));
                }
            }
        } else {
            const uint32_t* prev = Lb + (size_t)(l - 1) * W;
            for (int w = tid; w < W; w += T) {
                uint32_t bits = prev[w];
                while (bits) {
                    const int bit = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int g = w * 32 + bit;
                    const int64_t e0 = p.indptr[g], e1 = p.indptr[g + 1];
                    for (int64_t e = e0; e < e1; ++e) {
                        const int c = p.indices[e];
                        if (!test_bit(V, c)) atomicOr(&cur[c >> 5], 1u << (c & 31));
                    }
                }
            }
        }
        __syncthreads();
        // fold the level into V, count it, build its per-word prefix
        int local = 0;
        for (int w = w0; w < w1; ++w) {
            const uint32_t bits = cur[w];
            V[w] |= bits;
            local += __popc(bits);
        }
        int total;
        int run = block_exclusive_scan(local, s_scan, &total);
        uint32_t* pl = pre + (size_t)l * W;
        for (int w = w0; w < w1; ++w) {
            pl[w] = (uint32_t)run;
            run += __popc(cur[w]);
        }
        if (tid == 0) {
            s_lvl_base[l] = n;
            s_lvl_cnt[l + 1] = total;
        }
        if (total == 0) break;  // block-uniform (utils.py:71-72)
        n += total;
        nlev = l + 1;
        __syncthreads();
    }
    __syncthreads();

    // ---------------- allocation 1: nodes[n] + rowptr[n+1] ----------------
    const int64_t words1 = ((int64_t)n + (n + 1) + 31) & ~int64_t(31);
    if (tid == 0) s_base = (long long)atomicAdd(&p.counters[S3_CTR_CURSOR], (unsigned long long)words1);
    __syncthreads();
    const int64_t base1 = s_base;
    bool overflow = base1 + words1 > p.arena_words;
    int32_t* nodes = p.arena + base1;
    int32_t* rowptr = nodes + n;

    if (!overflow) {
        if (tid == 0) {
            nodes[0] = s0;
            if (nseed == 2) nodes[1] = s1;
            rowptr[0] = 0;
        }
        for (int l = 0; l < nlev; ++l) {
            const uint32_t* cur = Lb + (size_t)l * W;
            const uint32_t* pl = pre + (size_t)l * W;
            const int lb = s_lvl_base[l];
            for (int w = tid; w < W; w += T) {
                uint32_t bits = cur[w];
                int r = lb + (int)pl[w];
                while (bits) {
                    const int bit = __ffs(bits) - 1;
                    bits &= bits - 1;
                    nodes[r++] = w * 32 + bit;
                }
            }
        }
    }
    __syncthreads();

    // ---------------- count pass: induced, masked degree of every node ----------------
    if (!overflow) {
        for (int j = tid; j < n; j += T) {
            const int g = nodes[j];
            const int64_t e0 = p.indptr[g], e1 = p.indptr[g + 1];
            my_deg += (unsigned long long)(e1 - e0);
            int deg = 0;
            for (int64_t e = e0; e < e1; ++e) {
                const int c = p.indices[e];
                if (!test_bit(V, c)) continue;
                if (mask_target && ((j == 0 && c == s1) || (j == 1 && c == s0))) continue;  // utils.py:79-80
                ++deg;
            }
            rowptr[j + 1] = deg;
        }
    }
    __syncthreads();
    // in-place inclusive scan of rowptr[1..n] (thread-contiguous chunks)
    int m = 0;
    if (!overflow) {
        const int jc = (n + T - 1) / T;
        const int j0 = min(n, tid * jc), j1 = min(n, j0 + jc);
        int local = 0;
        for (int j = j0; j < j1; ++j) local += rowptr[j + 1];
        int run = block_exclusive_scan(local, s_scan, &m);
        for (int j = j0; j < j1; ++j) {
            run += rowptr[j + 1];
            rowptr[j + 1] = run;
        }
    }
    __syncthreads();

    // ---------------- allocation 2: lcol[m] + extra selected rows ----------------
    int sel_bound = 0;
    if (!overflow && p.strategy != S3_STRATEGY_NONE) {
        const int d0 = rowptr[1], d1 = rowptr[2] - rowptr[1];
        sel_bound = p.strategy == S3_STRATEGY_INTERSECTION ? min(d0, d1) : d0 + d1;
    }
    const int64_t words2 = ((int64_t)m + sel_bound + 31) & ~int64_t(31);
    int64_t base2 = 0;
    if (!overflow) {
        if (tid == 0) s_base = (long long)atomicAdd(&p.counters[S3_CTR_CURSOR], (unsigned long long)words2);
        __syncthreads();
        base2 = s_base;
        overflow = base2 + words2 > p.arena_words;
    }
    int32_t* lcol = p.arena + base2;
    int32_t* sel = lcol + m;

    // ---------------- fill pass: local column ids, ascending global id per row ----------------
    if (!overflow) {
        for (int j = tid; j < n; j += T) {
            const int g = nodes[j];
            int pos = rowptr[j];
            const int64_t e0 = p.indptr[g], e1 = p.indptr[g + 1];
            for (int64_t e = e0; e < e1; ++e) {
                const int c = p.indices[e];
                if (!test_bit(V, c)) continue;
                if (mask_target && ((j == 0 && c == s1) || (j == 1 && c == s0))) continue;
                lcol[pos++] = local_id(c, s0, s1, nseed, Lb, pre, s_lvl_base, nlev, W);
            }
        }
    }
    int partner_local = -1;
    if (p.flow == S3_FLOW_SOP && test_bit(V, s1)) partner_local = local_id(s1, s0, s1, nseed, Lb, pre, s_lvl_base, nlev, W);
    __syncthreads();

    // ---------------- row selection (PoS Plus): tuned_SIGN.py:228-238 ----------------
    // Neighbours of local 0 / 1 are hop-1 nodes, local ids [2, 2 + cnt1). Two flag bitmaps over
    // that range (re-using V and the level-1 bitmap, both dead now) give the intersection or
    // union in ascending local id.
    int s = nseed;
    if (!overflow && p.strategy != S3_STRATEGY_NONE) {
        const int cnt1 = nlev > 0 ? s_lvl_cnt[1] : 0;
        const int FW = (cnt1 + 31) >> 5;
        uint32_t* F0 = V;
        uint32_t* F1 = Lb;
        for (int w = tid; w < FW; w += T) {
            F0[w] = 0u;
            F1[w] = 0u;
        }
        __syncthreads();
        for (int e = rowptr[0] + tid; e < rowptr[1]; e += T) {
            const int c = lcol[e] - 2;
            if (c >= 0) atomicOr(&F0[c >> 5], 1u << (c & 31));
        }
        for (int e = rowptr[1] + tid; e < rowptr[2]; e += T) {
            const int c = lcol[e] - 2;
            if (c >= 0) atomicOr(&F1[c >> 5], 1u << (c & 31));
        }
        __syncthreads();
        const int fc = (FW + T - 1) / T;
        const int f0 = min(FW, tid * fc), f1 = min(FW, f0 + fc);
        const bool inter = p.strategy == S3_STRATEGY_INTERSECTION;
        int local = 0;
        for (int w = f0; w < f1; ++w) local += __popc(inter ? (F0[w] & F1[w]) : (F0[w] | F1[w]));
        int extra;
        int run = block_exclusive_scan(local, s_scan, &extra);
        for (int w = f0; w < f1; ++w) {
            uint32_t bits = inter ? (F0[w] & F1[w]) : (F0[w] | F1[w]);
            while (bits) {
                const int bit = __ffs(bits) - 1;
                bits &= bits - 1;
                sel[run++] = 2 + w * 32 + bit;
            }
        }
        s = nseed + extra;
    }

    // ---------------- allocation 3: float scratch of the record's work items ----------------
    const int sc = sel_chunk(p.flow);
    const int items = (s + sc - 1) / sc;
    const int64_t words3 = ((int64_t)items * item_words(p.flow, p.sign_k, n) + 31) & ~int64_t(31);
    int64_t base3 = 0;
    if (!overflow) {
        __syncthreads();
        if (tid == 0) s_base = (long long)atomicAdd(&p.counters[S3_CTR_CURSOR], (unsigned long long)words3);
        __syncthreads();
        base3 = s_base;
        overflow = base3 + words3 > p.arena_words;
    }

    // D accounting (block reduction through one shared atomic per warp)
    for (int d = 16; d > 0; d >>= 1) my_deg += __shfl_down_sync(0xffffffffu, my_deg, d);
    if ((tid & 31) == 0) atomicAdd(&s_sumdeg, my_deg);
    __syncthreads();

    if (tid == 0) {
        off[S3_OFF_NODES] = base1;
        off[S3_OFF_ROWPTR] = base1 + n;
        off[S3_OFF_LCOL] = base2;
        off[S3_OFF_SEL] = base2 + m;
        off[S3_OFF_F32] = base3;
        cnt[S3_CNT_N] = n;
        cnt[S3_CNT_M] = m;
        cnt[S3_CNT_S] = overflow ? 0 : s;
        cnt[S3_CNT_STATUS] = overflow ? S3_REC_ARENA_OVERFLOW : S3_REC_OK;
        cnt[S3_CNT_PARTNER] = partner_local;
        for (int l = 0; l <= S3_MAX_HOPS; ++l) cnt[S3_CNT_HOP0 + l] = (l <= nlev) ? s_lvl_cnt[l] : 0;
        if (overflow) atomicAdd(&p.counters[S3_CTR_ERRORS], 1ull);
        atomicMax(&p.counters[S3_CTR_MAX_N], (unsigned long long)n);
        atomicAdd(&p.counters[S3_CTR_SUM_N], (unsigned long long)n);
        atomicAdd(&p.counters[S3_CTR_SUM_D], s_sumdeg);
    }
}

}  // namespace

cudaError_t launch_extract_bitmap(const s3_graph& g, const s3_batch& b, cudaStream_t st) {
    ExtractParams p;
    p.indptr = g.indptr;
    p.indices = g.indices;
    p.num_nodes = g.num_nodes;
    p.link_src = b.link_src;
    p.link_dst = b.link_dst;
    p.num_records = s3_num_records(&b);
    p.flow = b.flow;
    p.strategy = b.flow == S3_FLOW_POS ? b.strategy : S3_STRATEGY_NONE;
    p.radius = b.flow == S3_FLOW_POS ? b.num_hops : b.sign_k;
    p.sign_k = b.sign_k;
    p.W = (int)((g.num_nodes + 31) / 32);
    p.arena = b.arena;
    p.arena_words = b.arena_words;
    p.off = b.off;
    p.cnt = b.cnt;
    p.counters = reinterpret_cast<unsigned long long*>(b.counters);
    if (p.num_records == 0) return cudaSuccess;
    const size_t smem = (size_t)s3_extract_smem_bytes(g.num_nodes, p.radius);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(extract_bitmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    extract_bitmap_kernel<<<(unsigned)p.num_records, kExtractThreads, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace s3
