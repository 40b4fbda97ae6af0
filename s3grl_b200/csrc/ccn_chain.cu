// Kernel 6 — PoS Plus CCN rows by a hop-limited SpMM chain: the default route of the `union` strategy, whose
// selected rows are ALL hop-1 nodes (~20 rows per PubMed link).
//
// Replaces reference tuned_SIGN.py:210-258 for the extra selected rows (x_k[sel] = S^k[sel] @ subg_x),
// like kernels 2 + 3 (diffuse.cu + gather_kernel.cuh) do, with a different factorisation. The weight
// formulation of kernels 2 + 3 costs s * n * (K+1) * F' FMAs per record (every selected row times every
// subgraph node): 41 MFMA for a mean PubMed union record, FP32-issue bound. Here the operators are
// propagated as whole matrices but only on the rows that can still reach a selected row:
//      x_k = S x_{k-1}   on rows of hop <= 1 + K - k      (selected rows are hop <= 1)
// i.e. x_1 on the whole subgraph, x_2 on hops <= K-1, ... x_K on hops <= 1: (m_1 + m_2 + ..) * F' additions,
// 3.3 M for the same record — and no weights, no s3_diffuse. The bound is shared-memory bandwidth: every
// induced edge of a level reads one row segment of the previous level's buffer (round 2, second session;
// round 1's version of this kernel was one 1024-thread CTA per record with K + 2 barriers per 8-32-column
// sub-chunk, per-CTA compaction of the padded CSR and rows handed to lane groups in node order: 323-455 ms per
// PubMed step, barrier / divergence bound).
//
// Two kernels:
//  * chain_prep_kernel, one WARP per record, in place in the record's arena: the padded local CSR (rows own
//    deg_G slots, holes are -1) becomes  crow int32[n+1] | ccol uint16[m]  and the rows of every hop are
//    ordered by descending induced degree (counting sort, 32 buckets) into rorder uint16[n] — lane groups of
//    a warp then work on rows of the same length. Heavy rows (degree >= 31) lead their hop's segment.
//  * chain_kernel<T>, one CTA per (record, column slab): CSR and order staged once, then per sub-chunk of
//    CW columns  y_0 = D^-1/2 [X | label]  ->  K levels ping-ponging between two [n][CW] shared buffers.
//    The LAST level of a sub-chunk writes nothing to shared memory, so it runs in the same barrier interval
//    as the y_0 fill of the next sub-chunk: K barriers per sub-chunk instead of K + 2. CW/4 lanes cooperate on a
//    row (one float4 each); heavy rows and all rows of the last level (few rows, long dependent chains) are
//    taken by a whole warp, edges split over its lane groups and combined by a fixed shuffle tree.
//    Records are classed by the shared memory they need: CW = 32 wherever it fits (a 128-byte row segment per
//    quarter warp is bank-conflict free; 64-byte segments conflict 1.5x, 32-byte ones 2.1x on random rows), in
//    256-thread CTAs (<= 54 KB, 4 per SM), 512 (<= 110 KB, 2 per SM) or 1024 (<= 222 KB); then CW = 16 / 8 / 4
//    in 1024-thread CTAs. Records beyond that (n > ~4 000) keep their CSR in shared memory and the two buffers in the
//    float scratch of their work items ("spill" class, CW = 32, global memory).
//  * pooled route (s3_ccn_chain_pooled, third session of round 2): records that would run at CW <= pool_cw in shared
//    memory (default 8: n > ~1 700, 21 % of PubMed's records with 73 % of the traffic) run the spill kernel instead with
//    their two [n][32] buffers in a slot of a caller-owned pool — 16 sub-chunks with 128-byte row segments (no bank
//    conflicts, 4-8x fewer barriers and fills) against 63-126, paid for with L2 latency. The launch is split by the
//    shared memory the CSR needs (54 / 110 / 222 KB) so that what is left of the SM's 256 KB serves the pool as L1.
//    Measured on the PubMed union step (chain ms): all in shared memory 356, pool_cw 8: 322 (319 with the split),
//    16: 338, 32: 447 (398 with the split) — profiles/README.md.
//  * level 1 from the feature matrix (chain_columns_x, default for the records with global-memory buffers): y_0 is not
//    stored; level 1 forms D^-1/2 [X | label] per neighbour from the graph's X, which is L2-resident and shared by all
//    CTAs, where a stored y_0 is private to its CTA (148 x 0.3-0.6 MB next to X's 39 MB in a 126 MB L2). No fill phase,
//    one barrier less per sub-chunk, same bits (the product is rounded before the sum). pool_cw 8: 319 -> 285,
//    16: 321 -> 280 (default), every record pooled: 330; the same for the shared-memory classes: slower (280 -> 321).
//
// Sums run in a fixed order per record (slot order inside a lane group, fixed shuffle tree for warp-wide rows):
// results do not depend on scheduling, batch composition or slab count. Rows 0 and 1 of every record still
// come from kernels 1 + 3 (bit-identical for every strategy); this kernel writes rows 2.. (the CCN rows,
// ascending local id) of all K+1 operators.
#include <stdlib.h>

#include "common.cuh"

namespace s3 {
namespace {

constexpr int kHeavyDeg = 31;  // rows with at least this many neighbours: bucket 0 of the degree sort, one warp each
constexpr int kPrefetchRows = 6;  // X rows of the next sub-chunk a lane group keeps in registers

struct ChainParams {
    const float* __restrict__ x;
    int64_t ldx;
    int F;  // feature columns; column index F of the chain is the label column (output column 0)
    int32_t* arena;
    const int64_t* __restrict__ off;
    int32_t* cnt;
    unsigned long long* counters;
    const int64_t* __restrict__ row_ptr;
    const int32_t* __restrict__ order;  // may be null
    int64_t num_records;
    int sign_k, strategy, flags, policy, cls;
    OutPtrs out;
    int64_t ldo, row_base;
    // pooled route (s3_ccn_chain_pooled): records whose shared-memory placement would run at CW <= pool_cw keep their CSR in
    // shared memory and take two [n][32] operator buffers from a slot of a caller-owned global pool (read through L2 / L1)
    float* pool;
    int* pool_busy;  // [pool_slots] 0 = free
    int64_t slot_floats;
    int pool_slots, pool_cw;
    int pool_x;      // 1: the records with buffers in global memory form y_0 from the feature matrix instead of storing it (chain_columns_x)
    int smem_x;      // 1: so do the records with buffers in shared memory
    int pool_split;  // pooled records run in the smallest of 1: {110, 222} KB / 2: {54, 110, 222} KB launches that holds their CSR
};

__device__ __forceinline__ bool chain_pooled(const ChainParams& p, int n, int cw) {
    return p.pool != nullptr && cw <= p.pool_cw && 64 * (int64_t)n <= p.slot_floats;
}

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_shfl_xor(float4 a, int d) {
    return make_float4(__shfl_xor_sync(0xffffffffu, a.x, d), __shfl_xor_sync(0xffffffffu, a.y, d),
                       __shfl_xor_sync(0xffffffffu, a.z, d), __shfl_xor_sync(0xffffffffu, a.w, d));
}

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

// ------------------------------------------------------------------------------------------------------------
// prep: one warp per record, in place. rowptr[n+1] -> crow (compact row starts), lcol[D slots] -> ccol uint16[m]
// (slot order: rows ascending, neighbours ascending), rowlen[n] -> rorder uint16[n] (per hop: descending degree
// bucket min(deg, 31), node order inside a bucket). cnt[S3_CNT_NSTORE] = -1 marks the record as converted.
// ------------------------------------------------------------------------------------------------------------
constexpr int kPrepWarps = 8;

__global__ void __launch_bounds__(kPrepWarps * 32) chain_prep_kernel(ChainParams p) {
    __shared__ int s_off[kPrepWarps][32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t rec = (int64_t)blockIdx.x * kPrepWarps + w;
    if (rec >= p.num_records) return;
    int32_t* cnt = p.cnt + rec * S3_NCNT;
    if (cnt[S3_CNT_STATUS] != S3_REC_OK) return;
    const int n = cnt[S3_CNT_N], s = cnt[S3_CNT_S], m = cnt[S3_CNT_M];
    const int n1 = cnt[S3_CNT_HOP0] + cnt[S3_CNT_HOP0 + 1];
    if (s <= 2 || cnt[S3_CNT_NSTORE] < 0) return;  // no CCN rows / already converted
    if (!chain_eligible(p.flags, p.strategy, n, m, n1) && !chain_spill_eligible(p.flags, p.strategy, n, m, n1, s, p.sign_k)) return;
    const int64_t* off = p.off + rec * S3_NOFF;
    int32_t* rowptr = p.arena + off[S3_OFF_ROWPTR];
    int32_t* rowlen = p.arena + off[S3_OFF_ROWLEN];
    int32_t* lcol = p.arena + off[S3_OFF_LCOL];
    uint16_t* ccol = reinterpret_cast<uint16_t*>(lcol);
    uint16_t* rorder = reinterpret_cast<uint16_t*>(rowlen);

    // 1. columns: stream the padded slots, keep the non-holes. The compact position of a slot never exceeds its
    //    padded position and a uint16 is half an int32, so every store lands below the next step's loads.
    const int Dslots = rowptr[n];
    int running = 0;
    int c = lane < Dslots ? lcol[lane] : -1;
    for (int base = 0; base < Dslots; base += 32) {
        const int e_next = base + 32 + lane;
        const int c_next = e_next < Dslots ? lcol[e_next] : -1;  // loaded before this step's stores (higher addresses)
        const unsigned keep = __ballot_sync(full, c >= 0);
        if (c >= 0) ccol[running + __popc(keep & lt)] = (uint16_t)c;
        running += __popc(keep);
        c = c_next;
    }
    __syncwarp();
    // 2. row starts of the compact CSR: exclusive scan of the induced degrees
    running = 0;
    for (int base = 0; base < n; base += 32) {
        const int j = base + lane;
        const int d = j < n ? rowlen[j] : 0;
        int inc = d;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) {
            const int t = __shfl_up_sync(full, inc, sft);
            if (lane >= sft) inc += t;
        }
        if (j < n) rowptr[j] = running + inc - d;
        running += __shfl_sync(full, inc, 31);
    }
    if (lane == 0) rowptr[n] = running;
    __syncwarp();  // crow is read back below by other lanes
    // 3. per hop: rows in descending degree bucket (counting sort; ranks inside a step from match_any: deterministic)
    int a = 0;
    for (int l = 0; l <= S3_MAX_HOPS; ++l) {
        const int b = a + cnt[S3_CNT_HOP0 + l];
        if (b > a) {
            s_off[w][lane] = 0;
            __syncwarp();
            for (int j = a + lane; j < b; j += 32) {
                const int d = rowptr[j + 1] - rowptr[j];
                atomicAdd(&s_off[w][31 - min(d, 31)], 1);
            }
            __syncwarp();
            const int v = s_off[w][lane];
            int inc = v;
#pragma unroll
            for (int sft = 1; sft < 32; sft <<= 1) {
                const int t = __shfl_up_sync(full, inc, sft);
                if (lane >= sft) inc += t;
            }
            __syncwarp();
            s_off[w][lane] = inc - v;
            __syncwarp();
            for (int base = a; base < b; base += 32) {
                const int j = base + lane;
                const bool valid = j < b;
                const int bucket = valid ? 31 - min(rowptr[j + 1] - rowptr[j], 31) : 32 + lane;
                const unsigned same = __match_any_sync(full, bucket);
                const int rank = __popc(same & lt);
                const int o = valid ? s_off[w][bucket] : 0;
                __syncwarp();
                if (valid && rank == 0) s_off[w][bucket] = o + __popc(same);
                __syncwarp();
                if (valid) rorder[a + o + rank] = (uint16_t)j;
            }
            __syncwarp();
        }
        a = b;
    }
    if (lane == 0) {
        cnt[S3_CNT_NSTORE] = -1;
        // accounting for the kernel's roofline: row segments the K levels read (bench.py)
        int hop_end[S3_MAX_HOPS + 2], acc = 0;
        for (int l = 0; l <= S3_MAX_HOPS; ++l) {
            acc += cnt[S3_CNT_HOP0 + l];
            hop_end[l] = acc;
        }
        hop_end[S3_MAX_HOPS + 1] = acc;
        unsigned long long reads = 0;
        for (int k = 1; k <= p.sign_k; ++k) reads += (unsigned long long)rowptr[hop_end[min(1 + p.sign_k - k, S3_MAX_HOPS + 1)]];
        atomicAdd(&p.counters[S3_CTR_CHAIN_READS], reads);
        atomicAdd(&p.counters[S3_CTR_CHAIN_RECORDS], 1ull);
        atomicAdd(&p.counters[S3_CTR_CHAIN_N], (unsigned long long)n);
    }
}

// ------------------------------------------------------------------------------------------------------------
// the chain
// ------------------------------------------------------------------------------------------------------------
struct ChainRec {
    int n, n1, m;
    int64_t orow0;
    const int32_t* nodes;  // global
    const uint2* meta;     // shared from here on: rows in processing order, {first edge, local id | degree << 16}
    const float* cdis;     // D^-1/2 by local id
    const int* cpos;
    const uint16_t* ccol;
    const int* hop_end;  // [S3_MAX_HOPS + 2]
    const int* nheavy;   // [S3_MAX_HOPS + 1] heavy rows leading every hop's segment of the processing order
    const float* dtab;   // [2][32]: 1/sqrt(d) and its square for d < 32
    int* ctr;            // [2][S3_MAX_K + 1] row hand-out counters of the middle levels
};

#ifndef S3_CHAIN_GRAB
#define S3_CHAIN_GRAB 1
#endif
constexpr int kGrabSteps = S3_CHAIN_GRAB;  // warp steps a warp takes from the hand-out counter at once (0: static round robin)

// The previous level's row of local node c, this lane's four columns. FROM_X (level 1 of the pooled route, S3GRL_CHAIN_POOL_X):
// y_0[c] = D^-1/2 [X | label] is not stored at all but formed from the graph's feature matrix — L2-resident and shared by
// every record, where a stored y_0 is private to its CTA — with the two roundings the stored value has (the product is
// rounded before it is added: __fmul_rn is never contracted into an FMA), so both routes give the same bits.
template <int LPR, bool FROM_X>
struct PrevRows {
    const float4* prev;
    const float* x;
    const int32_t* nodes;
    const float* cdis;
    int64_t ldx;
    int f0, l, labpos;  // labpos: position of the label column among this lane's four columns, or -1
    __device__ __forceinline__ float4 operator()(int c) const {
        if (!FROM_X) return prev[c * LPR + l];
        float4 v = f4_zero();
        if (f0 < ldx) v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)__ldg(nodes + c) * ldx + f0));
        if (labpos >= 0) {
            const float lab = c < 2 ? 1.0f : 0.0f;  // zero-one label, tuned_SIGN.py:234
            if (labpos == 0) v.x = lab;
            else if (labpos == 1) v.y = lab;
            else if (labpos == 2) v.z = lab;
            else v.w = lab;
        }
        const float dc = cdis[c];
        return make_float4(__fmul_rn(v.x, dc), __fmul_rn(v.y, dc), __fmul_rn(v.z, dc), __fmul_rn(v.w, dc));
    }
};

// D neighbours of one row, D a compile-time bound shared by the rows of the warp step (they are sorted by degree)
template <int D, class Prev>
__device__ __forceinline__ float4 row_sum(const Prev& prev, const uint16_t* cc, int d) {
    int c[D];
#pragma unroll
    for (int t = 0; t < D; ++t) c[t] = t < d ? (int)cc[t] : -1;
    float4 acc = f4_zero();
#pragma unroll
    for (int t = 0; t < D; ++t)
        if (c[t] >= 0) acc = f4_add(acc, prev(c[t]));
    return acc;
}

// One level for the chain columns [cs, cs + CW): x_k = D^-1/2 (sum over neighbours of y_{k-1}), y_k = D^-1/2 x_k.
// `next` is null for the last level (k == K), which only writes the output rows.
template <int CW, int T, bool FROM_X = false>
// (prev / next are NOT __restrict__: in the spill class they are global memory written earlier in the same kernel, which the
// non-coherent load path a const __restrict__ pointer invites must not serve)
__device__ __forceinline__ void chain_level(const ChainParams& p, const ChainRec& r, int k, const float4* prev_buf, float4* next, int cs,
                                            int* ctr) {
    constexpr int LPR = CW / 4, RPW = 32 / LPR, NWARP = T / 32;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPR, l = lane - sub * LPR;
    const int K = p.sign_k, F = p.F;
    const int f0 = cs + 4 * l;
    const int top = min(1 + K - k, S3_MAX_HOPS + 1);
    const int Rk = r.hop_end[top];
    const bool last = next == nullptr;
    float* outk = p.out.p[k];
    PrevRows<LPR, FROM_X> prev;
    prev.prev = prev_buf;
    prev.x = p.x;
    prev.nodes = r.nodes;
    prev.cdis = r.cdis;
    prev.ldx = p.ldx;
    prev.f0 = f0;
    prev.l = l;
    prev.labpos = (f0 + 3 >= F && f0 <= F) ? F - f0 : -1;

    // warp-wide rows: the heavy rows of every hop in range — and every row of the last level (hop <= 1: a few
    // rows whose dependent chains would otherwise leave the SM idle). Edges split over the lane groups.
    int a = 0;
    for (int h = 0; h <= min(top, S3_MAX_HOPS); ++h) {
        const int b = r.hop_end[h];
        const int cntw = last ? b - a : r.nheavy[h];
        for (int hi = warp; hi < cntw; hi += NWARP) {
            const uint2 mt = r.meta[a + hi];
            const int j = (int)(mt.y & 0xffffu), d = (int)(mt.y >> 16);
            if (last && j < 2) continue;  // rows 0 and 1 of the last level are neither read nor written
            const uint16_t* cc = r.ccol + mt.x;
            float4 acc = f4_zero();
#pragma unroll 2
            for (int e = sub; e < d; e += RPW) acc = f4_add(acc, prev((int)cc[e]));
#pragma unroll
            for (int sft = LPR; sft < 32; sft <<= 1) acc = f4_add(acc, f4_shfl_xor(acc, sft));
            if (sub == 0) {
                const float dj = r.cdis[j];
                const float4 xv = f4_scale(acc, dj);  // x_k[j]
                if (!last) next[j * LPR + l] = f4_scale(xv, dj);
                if (j >= 2 && j < r.n1) {
                    const int pos = r.cpos[j];
                    if (pos >= 0) store_row(outk + (r.orow0 + pos) * p.ldo, f0, F, xv);
                }
            }
        }
        a = b;
    }
    if (last) return;

    // light rows: one lane group per row, rows handed out kGrabSteps warp steps at a time from a shared counter
    // (the processing order is by descending degree inside a hop: the rows of a warp step have the same length,
    // and the warps that were busy with heavy rows simply take fewer steps)
    constexpr int kSteps = kGrabSteps > 0 ? kGrabSteps : 1;
    for (int base = warp * RPW;;) {
        if (kGrabSteps > 0) {
            if (lane == 0) base = atomicAdd(ctr, kSteps * RPW);
            base = __shfl_sync(full, base, 0);
        }
        if (base >= Rk) break;
#pragma unroll 1
        for (int st = 0; st < kSteps; ++st) {
            const int idx = base + st * RPW + sub;
            uint2 mt = make_uint2(0u, 0u);
            if (idx < Rk) mt = r.meta[idx];
            const int j = (int)(mt.y & 0xffffu);
            int d = (int)(mt.y >> 16);
            const bool act = idx < Rk && d < kHeavyDeg;
            if (!act) d = 0;
            const int dmax = __reduce_max_sync(full, d);
            if (base + st * RPW >= Rk) break;
            const uint16_t* cc = r.ccol + mt.x;
            float4 acc;
            switch (dmax) {
                case 0: acc = f4_zero(); break;
                case 1: acc = row_sum<1>(prev, cc, d); break;
                case 2: acc = row_sum<2>(prev, cc, d); break;
                case 3: acc = row_sum<3>(prev, cc, d); break;
                case 4: acc = row_sum<4>(prev, cc, d); break;
                case 5: acc = row_sum<5>(prev, cc, d); break;
                case 6: acc = row_sum<6>(prev, cc, d); break;
                case 7: acc = row_sum<7>(prev, cc, d); break;
                case 8: acc = row_sum<8>(prev, cc, d); break;
                default: {
                    acc = f4_zero();
                    for (int t0 = 0; t0 < dmax; t0 += 4) acc = f4_add(acc, row_sum<4>(prev, cc + t0, d - t0));
                }
            }
            if (act) {
                next[j * LPR + l] = f4_scale(acc, r.dtab[32 + d]);  // y_k[j] = x_k[j] / sqrt(d) = acc / d
                if (j >= 2 && j < r.n1) {
                    const int pos = r.cpos[j];
                    if (pos >= 0) store_row(outk + (r.orow0 + pos) * p.ldo, f0, F, f4_scale(acc, r.dtab[d]));
                }
            }
        }
        if (kGrabSteps == 0) base += NWARP * RPW;
    }
}

// All sub-chunks [c0, c1) of a record's column slab.
template <int CW, int T>
__device__ __forceinline__ void chain_columns(const ChainParams& p, const ChainRec& r, float4* buf0, int c0, int c1) {
    constexpr int LPR = CW / 4, RPW = 32 / LPR, G = T / LPR;
    constexpr int MAXR = kPrefetchRows;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPR, l = lane - sub * LPR;
    const int g = warp * RPW + sub;
    const int K = p.sign_k, F = p.F;
    const int R0 = r.hop_end[min(1 + K, S3_MAX_HOPS + 1)];
    const int bstride = r.n * LPR;  // float4 per buffer
    auto buf = [&](int i) -> float4* { return buf0 + (i & 1) * bstride; };

    auto load_x = [&](int j, int cs) -> float4 {  // columns cs + 4l .. of [X | label] of local node j; the label is patched in
        const int f0 = cs + 4 * l;
        float4 v = f4_zero();
        if (f0 < p.ldx) v = __ldg(reinterpret_cast<const float4*>(p.x + (int64_t)r.nodes[j] * p.ldx + f0));
        return v;
    };
    auto put_y0 = [&](int j, float4 v, float4* dst, int cs) {
        const int f0 = cs + 4 * l;
        if (f0 + 3 >= F && f0 <= F) {  // the sub-chunk that holds the label column
            const float lab = j < 2 ? 1.0f : 0.0f;  // zero-one label, tuned_SIGN.py:234
            if (f0 == F) v.x = lab;
            else if (f0 + 1 == F) v.y = lab;
            else if (f0 + 2 == F) v.z = lab;
            else v.w = lab;
        }
        dst[j * LPR + l] = f4_scale(v, r.cdis[j]);
        if (j >= 2 && j < r.n1) {  // operator 0 of a CCN row: the row itself
            const int pos = r.cpos[j];
            if (pos >= 0) store_row(p.out.p[0] + (r.orow0 + pos) * p.ldo, f0, F, v);
        }
    };

    float4 pre[MAXR];
    auto fetch = [&](int cs) {
#pragma unroll
        for (int u = 0; u < MAXR; ++u) {
            const int j = g + u * G;
            pre[u] = j < R0 ? load_x(j, cs) : f4_zero();
        }
    };
    if (c0 >= c1) return;
    fetch(c0 * CW);
    int w = 0;  // buffer that receives y_0 of the current sub-chunk
    for (int ci = c0; ci <= c1; ++ci) {
        const int cs = ci * CW;
        int* ctr = r.ctr + (ci & 1) * (S3_MAX_K + 1);
        if (ci < c1) {
            // the hand-out counters of the next sub-chunk's middle levels (last used two sub-chunks ago)
            if (threadIdx.x < S3_MAX_K + 1) r.ctr[((ci + 1) & 1) * (S3_MAX_K + 1) + threadIdx.x] = 0;
            // y_0 of this sub-chunk (the buffer was last read two barriers ago)
#pragma unroll
            for (int u = 0; u < MAXR; ++u) {
                const int j = g + u * G;
                if (j < R0) put_y0(j, pre[u], buf(w), cs);
            }
            for (int j = g + MAXR * G; j < R0; j += G) put_y0(j, load_x(j, cs), buf(w), cs);
            if (ci + 1 < c1) fetch(cs + CW);  // in flight during the levels below
        }
        // last level of the previous sub-chunk: reads the other buffer, writes global memory only
        if (ci > c0) chain_level<CW, T>(p, r, K, buf(w ^ 1), nullptr, cs - CW, nullptr);
        if (ci == c1) break;
        __syncthreads();
        for (int k = 1; k < K; ++k) {
            chain_level<CW, T>(p, r, k, buf(w ^ ((k - 1) & 1)), buf(w ^ (k & 1)), cs, ctr + k);
            __syncthreads();
        }
        w ^= (K & 1);
    }
}

// The pooled route without a stored y_0 (S3GRL_CHAIN_POOL_X; CW = 32, sign_k >= 2): level 1 forms y_0 on the fly from the
// graph's feature matrix (PrevRows<.., true>), so a sub-chunk has no fill phase, one barrier less, and the CTA's private
// footprint in L2 is the two buffers of levels 1.. only. The last level of a sub-chunk still shares a barrier interval
// with the next sub-chunk's first phase: level 1 writes the buffer that last level does not read.
template <int CW, int T>
__device__ __forceinline__ void chain_columns_x(const ChainParams& p, const ChainRec& r, float4* buf0, int c0, int c1) {
    constexpr int LPR = CW / 4, RPW = 32 / LPR, G = T / LPR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPR, l = lane - sub * LPR;
    const int g = warp * RPW + sub;
    const int K = p.sign_k, F = p.F;
    const int bstride = r.n * LPR;  // float4 per buffer
    auto buf = [&](int i) -> float4* { return buf0 + (i & 1) * bstride; };
    if (c0 >= c1) return;
    int w = 0;   // buffer level 1 of the current sub-chunk writes
    int lr = 0;  // buffer the last level of the previous sub-chunk reads
    for (int ci = c0; ci <= c1; ++ci) {
        const int cs = ci * CW;
        int* ctr = r.ctr + (ci & 1) * (S3_MAX_K + 1);
        if (ci < c1) {
            if (threadIdx.x < S3_MAX_K + 1) r.ctr[((ci + 1) & 1) * (S3_MAX_K + 1) + threadIdx.x] = 0;
            // operator 0 of the CCN rows: the row itself, [label = 0 | X[node]]
            const int f0 = cs + 4 * l;
            for (int j = 2 + g; j < r.n1; j += G) {
                const int pos = r.cpos[j];
                if (pos < 0) continue;
                float4 v = f4_zero();
                if (f0 < p.ldx) v = __ldg(reinterpret_cast<const float4*>(p.x + (int64_t)r.nodes[j] * p.ldx + f0));
                if (f0 + 3 >= F && f0 <= F) {
                    if (f0 == F) v.x = 0.0f;
                    else if (f0 + 1 == F) v.y = 0.0f;
                    else if (f0 + 2 == F) v.z = 0.0f;
                    else v.w = 0.0f;
                }
                store_row(p.out.p[0] + (r.orow0 + pos) * p.ldo, f0, F, v);
            }
        }
        if (ci > c0) chain_level<CW, T>(p, r, K, buf(lr), nullptr, cs - CW, nullptr);  // last level of the previous sub-chunk
        if (ci == c1) break;
        chain_level<CW, T, true>(p, r, 1, nullptr, buf(w), cs, ctr + 1);
        __syncthreads();
        for (int k = 2; k < K; ++k) {
            chain_level<CW, T>(p, r, k, buf(w ^ (k & 1)), buf(w ^ ((k - 1) & 1)), cs, ctr + k);
            __syncthreads();
        }
        lr = w ^ (K & 1);  // level K reads what level K - 1 wrote: buf(w ^ ((K - 2) & 1))
        w = lr ^ 1;
    }
}

// SPILL: the records of class 3 — operator buffers in the record's global float scratch instead of shared memory
// XR: level 1 from the feature matrix (chain_columns_x) — a separate instantiation, so the code of the other route is untouched
template <int T, bool SPILL, bool XR>
__global__ void __launch_bounds__(T, 1024 / T) chain_kernel(ChainParams p) {
    extern __shared__ __align__(16) int s_dyn[];
    __shared__ int s_hop_end[S3_MAX_HOPS + 2];
    __shared__ int s_nheavy[S3_MAX_HOPS + 1];
    __shared__ int s_ctr[2 * (S3_MAX_K + 1)];
    __shared__ float s_dtab[64];
    const int tid = threadIdx.x;
    const int32_t rec = p.order ? p.order[blockIdx.x] : (int32_t)blockIdx.x;
    if (rec < 0) return;
    const int32_t* cnt = p.cnt + (int64_t)rec * S3_NCNT;
    if (cnt[S3_CNT_STATUS] != S3_REC_OK) return;
    const int n = cnt[S3_CNT_N], s = cnt[S3_CNT_S], m = cnt[S3_CNT_M];
    const int n1 = cnt[S3_CNT_HOP0] + cnt[S3_CNT_HOP0 + 1];
    if (s <= 2) return;  // no CCN rows
    // records that fit neither placement keep the work-item path (s3_plan counted their items)
    int cw = 32;
    bool pooled = false;
    if (SPILL) {
        if (!chain_spill_eligible(p.flags, p.strategy, n, m, n1, s, p.sign_k)) {
            if (!chain_eligible(p.flags, p.strategy, n, m, n1) || !chain_pooled(p, n, chain_shape(n, m, n1, p.policy) & 255)) return;
            pooled = true;
            // the smallest launch whose shared memory holds the record's CSR: what it leaves of the 256 KB is L1 for the pool
            const int64_t need = 4 * (3 * (int64_t)n + n1 + (m + 2) / 2 + 8);
            const int target = (p.pool_split >= 2 && need <= chain_class_bytes(0)) ? 5 : ((p.pool_split >= 1 && need <= chain_class_bytes(1)) ? 4 : 3);
            if (target != p.cls) return;  // served by another pooled launch
        } else if (p.cls != 3) {
            return;
        }
    } else {
        if (!chain_eligible(p.flags, p.strategy, n, m, n1)) return;
        const int shape = chain_shape(n, m, n1, p.policy);
        if ((shape >> 8) != p.cls) return;  // served by the launch of another CTA size
        cw = shape & 255;
        if (chain_pooled(p, n, cw)) return;  // served by the pooled launch
    }
    // column slab of this CTA
    const int nchunks = p.F / cw + 1;  // chain columns 0..F
    const int per = (nchunks + (int)gridDim.y - 1) / (int)gridDim.y;
    const int c0 = (int)blockIdx.y * per, c1 = min(nchunks, c0 + per);
    if (c0 >= c1) return;

    const int64_t* off = p.off + (int64_t)rec * S3_NOFF;
    const int32_t* g_crow = p.arena + off[S3_OFF_ROWPTR];                                        // compact (chain_prep_kernel)
    const uint16_t* g_rorder = reinterpret_cast<const uint16_t*>(p.arena + off[S3_OFF_ROWLEN]);  // uint16[n]
    const uint32_t* g_ccol = reinterpret_cast<const uint32_t*>(p.arena + off[S3_OFF_LCOL]);      // uint16[m]
    const int32_t* sel = p.arena + off[S3_OFF_SEL];

    // shared memory: [buf0 n*CW | buf1 n*CW | meta 2n | cdis n | cpos n1 | ccol (m+1)/2 words]; SPILL: the two buffers sit
    // behind the record's own work item in its float scratch (128-byte aligned, 2 * n * 32 floats)
    float4* buf0 = SPILL ? reinterpret_cast<float4*>(p.arena + off[S3_OFF_F32] + item_words(S3_FLOW_POS, p.sign_k, n))
                         : reinterpret_cast<float4*>(s_dyn);
    __shared__ int s_slot;
    if (SPILL && pooled) {
        // a free slot of the pool, starting from this SM's own pair (the spill launch holds one CTA per SM, so the first try
        // succeeds and a slot's lines stay in the L2 partition of the SM that keeps using it)
        if (tid == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            int slot = (int)((2u * smid) % (unsigned)p.pool_slots);
            while (atomicCAS(&p.pool_busy[slot], 0, 1) != 0) slot = slot + 1 == p.pool_slots ? 0 : slot + 1;
            s_slot = slot;
        }
        __syncthreads();
        buf0 = reinterpret_cast<float4*>(p.pool + (int64_t)s_slot * p.slot_floats);
    }
    uint2* meta = SPILL ? reinterpret_cast<uint2*>(s_dyn) : reinterpret_cast<uint2*>(buf0 + 2 * (int64_t)n * (cw / 4));
    float* cdis = reinterpret_cast<float*>(meta + n);
    int* cpos = reinterpret_cast<int*>(cdis + n);
    uint32_t* ccol32 = reinterpret_cast<uint32_t*>(cpos + n1);

    if (tid == 0) {
        int acc = 0;
        for (int l = 0; l <= S3_MAX_HOPS; ++l) {
            acc += cnt[S3_CNT_HOP0 + l];
            s_hop_end[l] = acc;
            s_nheavy[l] = 0;
        }
        s_hop_end[S3_MAX_HOPS + 1] = acc;
    }
    if (tid < 2 * (S3_MAX_K + 1)) s_ctr[tid] = 0;
    if (tid < 32) {
        const float dis = tid > 0 ? 1.0f / sqrtf((float)tid) : 0.0f;  // tuned_SIGN.py:212-216, inf -> 0
        s_dtab[tid] = dis;
        s_dtab[32 + tid] = dis * dis;
    }
    for (int j = tid; j < n1; j += T) cpos[j] = -1;
    for (int i = tid; i < (m + 1) / 2; i += T) ccol32[i] = g_ccol[i];
    __syncthreads();
    for (int idx = tid; idx < n; idx += T) {  // rows in processing order
        const int j = g_rorder[idx];
        const int e0 = g_crow[j], d = g_crow[j + 1] - e0;
        meta[idx] = make_uint2((unsigned)e0, (unsigned)j | ((unsigned)d << 16));
        if (d >= kHeavyDeg) {
            int h = 0;
            while (idx >= s_hop_end[h]) ++h;  // the order permutes rows inside their hop
            atomicAdd(&s_nheavy[h], 1);
        }
    }
    for (int j = tid; j < n; j += T) {
        const int d = g_crow[j + 1] - g_crow[j];
        cdis[j] = d > 0 ? 1.0f / sqrtf((float)d) : 0.0f;
    }
    for (int q = tid; q < s - 2; q += T) cpos[sel[q]] = 2 + q;  // output row of every CCN node
    __syncthreads();

    ChainRec r;
    r.n = n;
    r.n1 = n1;
    r.m = m;
    r.orow0 = p.row_base + p.row_ptr[rec];
    r.nodes = p.arena + off[S3_OFF_NODES];
    r.meta = meta;
    r.cdis = cdis;
    r.cpos = cpos;
    r.ccol = reinterpret_cast<const uint16_t*>(ccol32);
    r.hop_end = s_hop_end;
    r.nheavy = s_nheavy;
    r.dtab = s_dtab;
    r.ctr = s_ctr;
    if (XR) {  // level 1 from the feature matrix, no stored y_0
        if (cw == 32) chain_columns_x<32, T>(p, r, buf0, c0, c1);
        else if (cw == 16) chain_columns_x<16, T>(p, r, buf0, c0, c1);
        else if (cw == 8) chain_columns_x<8, T>(p, r, buf0, c0, c1);
        else chain_columns_x<4, T>(p, r, buf0, c0, c1);
    } else {
        if (cw == 32) chain_columns<32, T>(p, r, buf0, c0, c1);
        else if (cw == 16) chain_columns<16, T>(p, r, buf0, c0, c1);
        else if (cw == 8) chain_columns<8, T>(p, r, buf0, c0, c1);
        else chain_columns<4, T>(p, r, buf0, c0, c1);
    }
    if (SPILL && pooled) {
        __threadfence();  // this thread's stores to the slot are performed before the slot is handed to another CTA
        __syncthreads();  // every thread's last reads of the slot
        if (tid == 0) atomicExch(&p.pool_busy[s_slot], 0);
    }
}

int env_int(const char* name, int dflt, int lo, int hi) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    const int x = atoi(v);
    return x < lo ? lo : (x > hi ? hi : x);
}

template <int T, bool SPILL, bool XR>
cudaError_t launch_class_x(const ChainParams& p, int cls, int slabs, cudaStream_t st) {
    static LaunchCache cache;  // the shared-memory opt-in is per device
    const size_t smem = (size_t)chain_class_bytes(cls == 4 ? 1 : (cls == 5 ? 0 : cls));  // 4 / 5: pooled launches with less shared memory
    cudaError_t e = cache.get(reinterpret_cast<const void*>(chain_kernel<T, SPILL, XR>), T, smem, nullptr, nullptr);
    if (e != cudaSuccess) return e;
    ChainParams q = p;
    q.cls = cls;
    chain_kernel<T, SPILL, XR><<<dim3((unsigned)p.num_records, (unsigned)slabs), T, smem, st>>>(q);
    return cudaGetLastError();
}

template <int T, bool SPILL>
cudaError_t launch_class(const ChainParams& p, int cls, int slabs, cudaStream_t st) {
    const bool xr = p.sign_k >= 2 && (SPILL ? p.pool_x : p.smem_x);
    return xr ? launch_class_x<T, SPILL, true>(p, cls, slabs, st) : launch_class_x<T, SPILL, false>(p, cls, slabs, st);
}

}  // namespace

cudaError_t launch_ccn_chain(const s3_graph& g, const s3_batch& b, int64_t num_records, const OutPtrs& out, int64_t ldo,
                             int64_t row_base, cudaStream_t st, float* pool, int* pool_busy, int64_t slot_floats, int pool_slots,
                             int pool_cw) {
    if (num_records == 0) return cudaSuccess;
    if (!b.row_ptr || b.flow != S3_FLOW_POS || b.strategy != S3_STRATEGY_UNION || !(b.flags & S3_BATCH_CCN_CHAIN))
        return cudaErrorInvalidValue;
    ChainParams p;
    p.x = g.x;
    p.ldx = g.ldx;
    p.F = (int)g.num_feat;
    p.arena = b.arena;
    p.off = b.off;
    p.cnt = b.cnt;
    p.counters = reinterpret_cast<unsigned long long*>(b.counters);
    p.row_ptr = b.row_ptr;
    p.order = b.order;
    p.num_records = num_records;
    p.sign_k = b.sign_k;
    p.strategy = b.strategy;
    p.flags = b.flags;
    // tuning knobs (A/B runs): which (CW, CTA size) a record gets, and how many CTAs share a record's columns
    p.policy = env_int("S3GRL_CHAIN_POLICY", 0, 0, 2);
    p.cls = 0;
    p.pool = pool_slots > 0 ? pool : nullptr;
    p.pool_busy = pool_busy;
    p.slot_floats = slot_floats;
    p.pool_slots = pool_slots;
    p.pool_cw = pool_cw;
    p.pool_split = 0;
    p.pool_x = p.pool ? env_int("S3GRL_CHAIN_POOL_X", 1, 0, 1) : 0;
    p.smem_x = env_int("S3GRL_CHAIN_SMEM_X", 0, 0, 1);
    p.out = out;
    p.ldo = ldo;
    p.row_base = row_base;
    const int slabs = env_int("S3GRL_CHAIN_SLABS", 1, 1, 16);
    chain_prep_kernel<<<(unsigned)((num_records + kPrepWarps - 1) / kPrepWarps), kPrepWarps * 32, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // largest CTAs first: their records are the long ones
    p.pool_split = p.pool ? env_int("S3GRL_CHAIN_POOL_SPLIT", 2, 0, 2) : 0;
    e = launch_class<1024, true>(p, 3, 1, st);  // the records whose buffers live in global memory: longest of all
    if (e != cudaSuccess) return e;
    for (int cls = 4; cls < 4 + p.pool_split; ++cls) {
        e = launch_class<1024, true>(p, cls, 1, st);
        if (e != cudaSuccess) return e;
    }
    e = launch_class<1024, false>(p, 2, slabs, st);
    if (e != cudaSuccess) return e;
    e = launch_class<512, false>(p, 1, slabs, st);
    if (e != cudaSuccess) return e;
    return launch_class<256, false>(p, 0, slabs, st);
}

}  // namespace s3
