// Kernel 6 — PoS Plus CCN rows by a hop-limited SpMM chain (used for the `union` strategy, whose
// selected rows are ALL hop-1 nodes: ~20 rows per PubMed link).
//
// Replaces reference tuned_SIGN.py:210-258 for the extra selected rows (x_k[sel] = S^k[sel] @ subg_x),
// like kernels 2 + 3 (diffuse.cu + gather_kernel.cuh) do, with a different factorisation. The weight
// formulation of kernels 2 + 3 costs s * n * (K+1) * F' FMAs per record (every selected row times every
// subgraph node): 41 MFMA for a mean PubMed union record, FP32-issue bound. Here the operators are
// propagated as whole matrices but only on the rows that can still reach a selected row:
//      x_k = S x_{k-1}   on rows of hop <= 1 + K - k      (selected rows are hop <= 1)
// i.e. x_1 on the whole subgraph, x_2 on hops <= K-1, ... x_K on hops <= 1: (m_1 + m_2 + .. ) * F' FMAs,
// 3.3 MFMA for the same record. One CTA per record walks the F' columns in sub-chunks of CW columns;
// per sub-chunk the scaled operator y_{k-1} = D^-1/2 x_{k-1} lives in one buffer and y_k in another
// (the inner loop is a bare sum over the neighbours' rows), and the compact CSR is staged once per CTA.
// Everything a record needs lives in shared memory (compact CSR + two [n][CW] buffers, CW = 32/16/8 by
// subgraph size); records too large for that (n > ~2.3 k) keep the work-item path of kernels 2 + 3 —
// s3_plan counts no CCN items for the records this kernel serves (S3_BATCH_CCN_CHAIN). CW/4 lanes cooperate
// on a row with one float4 each, 128/CW rows per warp step; sums run in slot order: results do not depend
// on scheduling.
//
// Rows 0 and 1 of every record still come from kernels 1 + 3 (bit-identical for every strategy); this
// kernel writes rows 2.. (the CCN rows, ascending local id) of all K+1 operators.
#include "common.cuh"

namespace s3 {
namespace {

constexpr int kChainThreads = 1024;

struct ChainParams {
    const float* __restrict__ x;
    int64_t ldx;
    int F;  // feature columns; column index F of the chain is the label column (output column 0)
    int32_t* arena;
    const int64_t* __restrict__ off;
    const int32_t* __restrict__ cnt;
    const int64_t* __restrict__ row_ptr;
    const int32_t* __restrict__ order;  // may be null
    int sign_k, strategy, flags;
    OutPtrs out;
    int64_t ldo, row_base;
};

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

// store the four chain columns f0..f0+3 of one output row: feature f -> output column f + 1, f == F (label) -> column 0
__device__ __forceinline__ void store_row(float* orow, int f0, int F, float4 v) {
    const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int f = f0 + t;
        if (f < F) orow[f + 1] = a[t];
        else if (f == F) orow[0] = a[t];
    }
}

constexpr int kHeavyDeg = 24;  // rows with more neighbours are processed by a whole warp, edges split over lane groups

__device__ __forceinline__ float4 f4_shfl_xor(float4 a, int d) {
    return make_float4(__shfl_xor_sync(0xffffffffu, a.x, d), __shfl_xor_sync(0xffffffffu, a.y, d),
                       __shfl_xor_sync(0xffffffffu, a.z, d), __shfl_xor_sync(0xffffffffu, a.w, d));
}

// All K levels for the chain columns [cs, cs + CW) of every sub-chunk. CW/4 lanes cooperate on a row (one float4
// each), 128/CW rows per warp step; rows with more than kHeavyDeg neighbours (hubs: one of them would hold up the
// whole level) are taken out of that schedule and processed one per warp, their edges split over the 128/CW lane
// groups and combined by a fixed shuffle tree. The X rows of the NEXT sub-chunk are fetched into registers before
// the levels of the current one, so their latency hides behind the shared-memory work.
// smem: crow/ccol compact CSR, cdis, cnode, cpos, heavy-row list, and two [n][CW] buffers.
template <int CW>
__device__ __forceinline__ void chain_columns(const ChainParams& p, int n, int n1, const int* hop_end, const int* crow,
                                              const int* ccol, const float* cdis, const int* cpos, const int* cnode,
                                              const int* heavy, int nheavy, float4* bufA, float4* bufB, int64_t orow0) {
    constexpr int LPR = CW / 4, RPW = 32 / LPR;  // lanes per row, rows per warp step
    constexpr int MAXR = 8;                      // rows of level 0 a thread may own (n <= MAXR * groups, checked by the caller)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NWARP = blockDim.x >> 5;
    const int sub = lane / LPR, l = lane - sub * LPR;
    const int K = p.sign_k, F = p.F;
    const int G = NWARP * RPW, g = warp * RPW + sub;  // row groups of the CTA
    const int R0 = hop_end[min(1 + K, S3_MAX_HOPS + 1)];
    float4 pre[MAXR];
    auto fetch = [&](int cs) {
        const int f0 = cs + 4 * l;
#pragma unroll
        for (int u = 0; u < MAXR; ++u) {
            const int j = g + u * G;
            pre[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < R0 && f0 < p.ldx && cs <= F) pre[u] = __ldg(reinterpret_cast<const float4*>(p.x + (int64_t)cnode[j] * p.ldx + f0));
        }
    };
    fetch(0);
    for (int cs = 0; cs <= F; cs += CW) {
        const int f0 = cs + 4 * l;
        __syncthreads();  // the previous sub-chunk is done with the buffers
        // level 0: y_0 = D^-1/2 [X | label] on the rows x_1 needs (hop <= 1 + K); x itself for the CCN rows
#pragma unroll
        for (int u = 0; u < MAXR; ++u) {
            const int j = g + u * G;
            if (j < R0) {
                float4 v = pre[u];
                const float lab = j < 2 ? 1.0f : 0.0f;  // zero-one label, tuned_SIGN.py:234
                if (f0 == F) v.x = lab;
                else if (f0 + 1 == F) v.y = lab;
                else if (f0 + 2 == F) v.z = lab;
                else if (f0 + 3 == F) v.w = lab;
                bufA[j * LPR + l] = f4_scale(v, cdis[j]);
                if (j >= 2 && j < n1) {
                    const int pos = cpos[j];
                    if (pos >= 0) store_row(p.out.p[0] + (orow0 + pos) * p.ldo, f0, F, v);
                }
            }
        }
        fetch(cs + CW);  // in flight during the K levels below
        __syncthreads();
        for (int k = 1; k <= K; ++k) {
            const float4* prev = (k & 1) ? bufA : bufB;
            float4* next = (k & 1) ? bufB : bufA;
            const int Rk = hop_end[min(1 + K - k, S3_MAX_HOPS + 1)];
            for (int j = g; j < Rk; j += G) {
                const int e0 = crow[j], e1 = crow[j + 1];
                if (e1 - e0 > kHeavyDeg) continue;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int e = e0; e < e1; ++e) acc = f4_add(acc, prev[ccol[e] * LPR + l]);
                const float dj = cdis[j];
                const float4 xv = f4_scale(acc, dj);  // x_k[j]
                next[j * LPR + l] = f4_scale(xv, dj);
                if (j >= 2 && j < n1) {
                    const int pos = cpos[j];
                    if (pos >= 0) store_row(p.out.p[k] + (orow0 + pos) * p.ldo, f0, F, xv);
                }
            }
            for (int hi = warp; hi < nheavy; hi += NWARP) {
                const int j = heavy[hi];
                if (j >= Rk) break;  // the list ascends
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                const int e1 = crow[j + 1];
#pragma unroll 2
                for (int e = crow[j] + sub; e < e1; e += RPW) acc = f4_add(acc, prev[ccol[e] * LPR + l]);
#pragma unroll
                for (int d = LPR; d < 32; d <<= 1) acc = f4_add(acc, f4_shfl_xor(acc, d));
                if (sub == 0) {
                    const float dj = cdis[j];
                    const float4 xv = f4_scale(acc, dj);
                    next[j * LPR + l] = f4_scale(xv, dj);
                    if (j >= 2 && j < n1) {
                        const int pos = cpos[j];
                        if (pos >= 0) store_row(p.out.p[k] + (orow0 + pos) * p.ldo, f0, F, xv);
                    }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(kChainThreads, 1) ccn_chain_kernel(ChainParams p) {
    extern __shared__ __align__(16) int s_dyn[];
    __shared__ int s_scan[33];
    __shared__ int s_hop_end[S3_MAX_HOPS + 2];
    const int tid = threadIdx.x;
    const int32_t rec = p.order ? p.order[blockIdx.x] : (int32_t)blockIdx.x;
    if (rec < 0) return;
    const int32_t* cnt = p.cnt + (int64_t)rec * S3_NCNT;
    if (cnt[S3_CNT_STATUS] != S3_REC_OK) return;
    const int n = cnt[S3_CNT_N], s = cnt[S3_CNT_S], m = cnt[S3_CNT_M];
    const int n1 = cnt[S3_CNT_HOP0] + cnt[S3_CNT_HOP0 + 1];
    if (s <= 2) return;  // no CCN rows
    // records that do not fit the shared-memory placement keep the work-item path (s3_plan counted their items)
    if (!chain_eligible(p.flags, p.strategy, n, m, n1)) return;
    const int64_t* off = p.off + (int64_t)rec * S3_NOFF;
    const int32_t* nodes = p.arena + off[S3_OFF_NODES];
    const int32_t* rowptr = p.arena + off[S3_OFF_ROWPTR];  // padded
    const int32_t* rowlen = p.arena + off[S3_OFF_ROWLEN];
    const int32_t* lcol = p.arena + off[S3_OFF_LCOL];      // padded, -1 holes
    const int32_t* sel = p.arena + off[S3_OFF_SEL];
    if (tid == 0) {
        int acc = 0;
        for (int l = 0; l <= S3_MAX_HOPS; ++l) {
            acc += cnt[S3_CNT_HOP0 + l];
            s_hop_end[l] = acc;
        }
        s_hop_end[S3_MAX_HOPS + 1] = acc;
    }
    const int64_t orow0 = p.row_base + p.row_ptr[rec];

    // shared memory: [bufA n*CW | bufB n*CW | cdis n | cnode n | cpos n1 | crow n+1 | ccol m | heavy rows m/24]
    const int64_t S = kChainSmemBytes / 4;
    const int64_t fixed = chain_fixed_words(n, m, n1);
    const int cw = fixed + 2 * (int64_t)n * 32 <= S ? 32 : (fixed + 2 * (int64_t)n * 16 <= S ? 16 : 8);
    float4* bufA = reinterpret_cast<float4*>(s_dyn);
    float4* bufB = bufA + (int64_t)n * (cw / 4);
    float* cdis = reinterpret_cast<float*>(bufB + (int64_t)n * (cw / 4));
    int* cnode = reinterpret_cast<int*>(cdis + n);
    int* cpos = cnode + n;
    int* crow = cpos + n1;
    int* ccol = crow + n + 1;
    int* heavy = ccol + m;  // rows with more than kHeavyDeg neighbours, ascending (at most m / kHeavyDeg of them)
    __shared__ int s_nheavy;

    for (int j = tid; j < n; j += kChainThreads) {
        const int d = rowlen[j];
        cdis[j] = d > 0 ? 1.0f / sqrtf((float)d) : 0.0f;  // tuned_SIGN.py:212-216, inf -> 0
        cnode[j] = nodes[j];
    }
    for (int j = tid; j < n1; j += kChainThreads) cpos[j] = -1;
    __syncthreads();
    for (int q = tid; q < s - 2; q += kChainThreads) cpos[sel[q]] = 2 + q;  // output row of every CCN node
    {
        // compact CSR: row starts = scan of the induced degrees; columns = the non-hole slots in slot order
        int running = 0;
        for (int base = 0; base < n; base += kChainThreads) {
            const int j = base + tid;
            const int d = j < n ? rowlen[j] : 0;
            int tile_total;
            const int ex = block_exclusive_scan(d, s_scan, &tile_total);
            if (j < n) crow[j] = running + ex;
            running += tile_total;
            __syncthreads();
        }
        if (tid == 0) crow[n] = running;
        running = 0;
        for (int base = 0; base < n; base += kChainThreads) {
            const int j = base + tid;
            const bool hv = j < n && rowlen[j] > kHeavyDeg;
            int tile_total;
            const int ex = block_exclusive_scan(hv ? 1 : 0, s_scan, &tile_total);
            if (hv) heavy[running + ex] = j;
            running += tile_total;
            __syncthreads();
        }
        if (tid == 0) s_nheavy = running;
        const int Dslots = rowptr[n];
        running = 0;
        for (int base = 0; base < Dslots; base += kChainThreads) {
            const int e = base + tid;
            const int col = e < Dslots ? lcol[e] : -1;
            int tile_total;
            const int ex = block_exclusive_scan(col >= 0 ? 1 : 0, s_scan, &tile_total);
            if (col >= 0) ccol[running + ex] = col;
            running += tile_total;
            __syncthreads();
        }
    }
    __syncthreads();

    if (cw == 32) chain_columns<32>(p, n, n1, s_hop_end, crow, ccol, cdis, cpos, cnode, heavy, s_nheavy, bufA, bufB, orow0);
    else if (cw == 16) chain_columns<16>(p, n, n1, s_hop_end, crow, ccol, cdis, cpos, cnode, heavy, s_nheavy, bufA, bufB, orow0);
    else chain_columns<8>(p, n, n1, s_hop_end, crow, ccol, cdis, cpos, cnode, heavy, s_nheavy, bufA, bufB, orow0);
}

}  // namespace

cudaError_t launch_ccn_chain(const s3_graph& g, const s3_batch& b, int64_t num_records, const OutPtrs& out, int64_t ldo,
                             int64_t row_base, cudaStream_t st) {
    if (num_records == 0) return cudaSuccess;
    if (!b.row_ptr || b.flow != S3_FLOW_POS || b.strategy != S3_STRATEGY_UNION || !(b.flags & S3_BATCH_CCN_CHAIN))
        return cudaErrorInvalidValue;
    static LaunchCache cache;  // the shared-memory opt-in is per device
    cudaError_t e = cache.get(reinterpret_cast<const void*>(ccn_chain_kernel), kChainThreads, kChainSmemBytes, nullptr, nullptr);
    if (e != cudaSuccess) return e;
    ChainParams p;
    p.x = g.x;
    p.ldx = g.ldx;
    p.F = (int)g.num_feat;
    p.arena = b.arena;
    p.off = b.off;
    p.cnt = b.cnt;
    p.row_ptr = b.row_ptr;
    p.order = b.order;
    p.sign_k = b.sign_k;
    p.strategy = b.strategy;
    p.flags = b.flags;
    p.out = out;
    p.ldo = ldo;
    p.row_base = row_base;
    ccn_chain_kernel<<<(unsigned)num_records, kChainThreads, kChainSmemBytes, st>>>(p);
    return cudaGetLastError();
}

}  // namespace s3
