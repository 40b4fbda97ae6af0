// Kernel 3 — fused row-select + weighted feature gather, writing the operator matrices.
//
// Replaces reference tuned_SIGN.py:177-187 / :240-260 (subg_x = [label | X_sub];
// x = subg_x[sel]; x_k = P_k[sel] @ subg_x) and, in SoP flow, tuned_SIGN.py:94-133
// (g = rows @ X, prepend the self-return weight, pack x / x1..xK).
//
// For one work item (up to SC selected rows of one record) and all K+1 operators at once:
//      out_k[row(c), 1 + f] = sum_j  w[j][k*SC + c] * X[node_j][f]      (k = 0 is the one-hot: x itself)
//      out_k[row(c), 0]     = label weight computed by kernel 2
// Every subgraph node's feature row is read exactly ONCE (4*F*n bytes, the dominant term of
// the roofline in SURVEY.md §8d) with 128-bit loads; NW = (K+1)*SC accumulator rows live in
// registers; weights and row offsets are staged through shared memory in tiles. Feature rows are
// padded by the host (DeviceGraph) so that every lane owns a valid 16-byte column: the inner
// loop is LDS(offset) + LDG.128 + LDS(weights) + FFMAs, with no predicates or 64-bit index math.
//
// Structural sparsity: a k-step walk cannot reach a node more than k hops away, so a node at
// hop l of the canonical (hop-major) order has w_k = 0 for k < l (k < l-1 when the selected
// rows are hop-1 CCN nodes). The node list is walked hop range by hop range with the inner
// loop specialised on the first live operator, which removes ~2/3 of the FMAs and weight
// loads on 3-hop subgraphs (most nodes sit on the outermost hop) and skips nodes beyond hop K.
//
// The sum over j runs in canonical node order within each row group and groups are combined
// in fixed order, so results are independent of scheduling, batch composition and GPU count.
// HBM/L2-bound streaming gather: no tensor cores (M = NW <= 16 rows, fp32 required by the
// 1e-5 tolerance).

#pragma once
#include "common.cuh"

namespace s3 {

struct GatherParams {
    const float* __restrict__ x;
    int64_t ldx;
    int F, F4;  // features, float4 columns per row (ldx / 4)
    const int32_t* __restrict__ arena;
    const int64_t* __restrict__ off;
    const int32_t* __restrict__ cnt;
    const int64_t* __restrict__ row_ptr;   // may be null: row = rec * num_seeds
    const int64_t* __restrict__ item_ptr;  // may be null
    const int32_t* __restrict__ item_rec;  // may be null
    const int32_t* __restrict__ order;     // may be null: largest-first schedule of the records
    int flow, sign_k, tpr;                 // tpr = threads per feature row (32/64/128)
    int ccn;                               // 1: CCN work items (item_rec/item_ptr), 0: the records' own rows
    OutPtrs out;
    int64_t ldo, row_base;
    // link pairing and explicit output placement (s3_batch.out_link / mirror / link_base); null / 0 when unused
    const int64_t* __restrict__ out_link;
    const int64_t* __restrict__ mirror;
    int64_t link_base;
    // s3_gather_peers: every row goes to num_dst buffers (this GPU's and its NVLink peers'), operator k of buffer
    // d at dst_base[d] + k * op_stride; num_dst == 0: the single local destination `out`
    int num_dst;
    float* dst_base[S3_MAX_PEERS];
    int64_t op_stride;
    // s3_gather_peers: operator 0 (x itself, an exact copy of [1 | X[node]]) is not sent over NVLink — every GPU holds
    // X and writes those rows locally for the whole link list (s3_fill_x0): a quarter less traffic at K = 3
    int skip_op0;
    // s3_gather_peers: rows of paired links (chain members) are not sent either — every GPU copies them from the
    // first link's rows after the exchange (s3_fill_mirrors): 23 % fewer rows over NVLink on the PubMed list
    int skip_chain;
};
// per-(SC, K1 range) translation units, so that the ~130 instantiations compile in parallel
#define S3_DECL_GATHER_TU(name) \
    cudaError_t name(const GatherParams& p, int K1, int C, dim3 grid, size_t smem, cudaStream_t st)
S3_DECL_GATHER_TU(launch_gather_sc1_lo);
S3_DECL_GATHER_TU(launch_gather_sc1_mid);
S3_DECL_GATHER_TU(launch_gather_sc1_hi);
S3_DECL_GATHER_TU(launch_gather_sc2_lo);
S3_DECL_GATHER_TU(launch_gather_sc2_mid);
S3_DECL_GATHER_TU(launch_gather_sc2_hi);
cudaError_t launch_gather_sc8_k2(const GatherParams& p, int C, dim3 grid, size_t smem, cudaStream_t st);
cudaError_t launch_gather_sc8_k3(const GatherParams& p, int C, dim3 grid, size_t smem, cudaStream_t st);
cudaError_t launch_gather_sc8_k4(const GatherParams& p, int C, dim3 grid, size_t smem, cudaStream_t st);
cudaError_t launch_gather_sc8_k5(const GatherParams& p, int C, dim3 grid, size_t smem, cudaStream_t st);
cudaError_t launch_gather_sc8_k6(const GatherParams& p, int C, dim3 grid, size_t smem, cudaStream_t st);

namespace {


#ifndef S3_KTILE
#define S3_KTILE 128
#endif
#ifndef S3_GATHER_BLOCKS
#define S3_GATHER_BLOCKS 8
#endif
constexpr int kTile = S3_KTILE;  // nodes staged per tile
// feature rows in flight per row group, by accumulator footprint (float4 per thread). Measured on PubMed PoS (8 float4):
// 2 rows 14.2 ms per step, 4 13.0, 8 12.2, 16 12.0 (with 128-node tiles) — the loads are L2 hits ~600 cycles away and the
// compiler interleaves them with the FFMA2s inside the 64-register budget
__host__ __device__ constexpr int rows_in_flight(int acc4) { return acc4 <= 8 ? 16 : acc4 <= 16 ? 8 : 4; }


template <int C>
struct GatherCtx {
    const int32_t* __restrict__ nodes;
    const float4* __restrict__ wgt4;
    const float4* xcol[C];  // x + this lane's float4 column(s)
    uint32_t ldx4;
    bool colok[C];
    float* s_w;
    uint32_t* s_off;  // float4 index of every staged node's feature row
    int tid, grp, G;
    // PoS records: the two seeds are accumulated in ascending GLOBAL id order, so that (u,v) and (v,u) produce
    // bit-identical rows (link pairing, pair.cu); true when local 0 has the larger global id
    bool swap01;
};

// Packed FP32 pairs (sm_100a FFMA2, PTX fma.rn.f32x2): two IEEE fused multiply-adds per issued instruction — the same
// bits as two scalar FFMAs, half the issue slots. The kernel is issue-bound on the L2-resident workloads (ncu: 67 %
// issue active, FFMA 42 % of the instructions), not FP32-pipe-bound (0.26 of the FFMA ceiling).
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
constexpr bool kUseF32x2 = true;
// accumulator of one float4 column: two packed pairs (x, y), (z, w) kept packed for the whole kernel
struct Acc {
    unsigned long long lo, hi;
};
__device__ __forceinline__ float4 acc_to_float4(const Acc& a) {
    float4 v;
    unpack2(a.lo, v.x, v.y);
    unpack2(a.hi, v.z, v.w);
    return v;
}

template <int K1, int SC, int KMIN>
__device__ __forceinline__ void load_weights(float (&w)[K1 * SC], const float* wrow) {
    if (SC == 2) {
#pragma unroll
        for (int k = KMIN; k < K1; ++k) {
            const float2 v = reinterpret_cast<const float2*>(wrow)[k];
            w[2 * k] = v.x;
            w[2 * k + 1] = v.y;
        }
    } else if (SC % 4 == 0) {
#pragma unroll
        for (int q4 = KMIN * SC / 4; q4 < K1 * SC / 4; ++q4) {
            const float4 v = reinterpret_cast<const float4*>(wrow)[q4];
            w[4 * q4] = v.x;
            w[4 * q4 + 1] = v.y;
            w[4 * q4 + 2] = v.z;
            w[4 * q4 + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int q = KMIN * SC; q < K1 * SC; ++q) w[q] = wrow[q];
    }
}

// Accumulate nodes [lo, hi) of the record; only operators k >= KMIN carry weight there.
// FULL: every lane owns a valid column (rows are padded), so the loads carry no predicates.
template <int K1, int SC, int C, int KMIN, bool FULL>
__device__ __forceinline__ void accumulate_range(Acc (&acc)[K1 * SC][C], int lo, int hi, const GatherCtx<C>& cx) {
    constexpr int NW = K1 * SC, NWP = (NW + 3) & ~3, Q0 = KMIN * SC;
    constexpr int kU = FULL ? rows_in_flight(NW * C) : 4;  // predicated loads (foreign row strides) keep round 1's depth
    float* s_w = cx.s_w;
    uint32_t* s_off = cx.s_off;
    const int tid = cx.tid, grp = cx.grp, G = cx.G;
    for (int base = lo; base < hi; base += kTile) {
        const int tn = min(kTile, hi - base);
        __syncthreads();
        const bool sw = cx.swap01 && base == 0;  // the seeds sit at the head of the first range
        if (tid < tn) s_off[tid] = (uint32_t)cx.nodes[base + ((sw && tid < 2) ? 1 - tid : tid)] * cx.ldx4;
        {
            const float4* src = cx.wgt4 + (int64_t)base * (NWP / 4);
            float4* dst = reinterpret_cast<float4*>(s_w);
            for (int i = tid; i < tn * (NWP / 4); i += kGatherThreads) {
                int si = i;
                if (sw && i < 2 * (NWP / 4)) si = i < NWP / 4 ? i + NWP / 4 : i - NWP / 4;
                dst[i] = src[si];
            }
        }
        __syncthreads();

        int t = grp;
        // main loop: kU rows in flight per row group, no bounds checks
        for (; t + (kU - 1) * G < tn; t += kU * G) {
            float4 xv[kU][C];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const uint32_t o = s_off[t + u * G];
#pragma unroll
                for (int i = 0; i < C; ++i)
                    xv[u][i] = (FULL || cx.colok[i]) ? __ldg(cx.xcol[i] + o) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                float w[NW];
                load_weights<K1, SC, KMIN>(w, s_w + (t + u * G) * NWP);
                if (kUseF32x2) {
                    unsigned long long xlo[C], xhi[C];
#pragma unroll
                    for (int i = 0; i < C; ++i) {
                        xlo[i] = pack2(xv[u][i].x, xv[u][i].y);
                        xhi[i] = pack2(xv[u][i].z, xv[u][i].w);
                    }
#pragma unroll
                    for (int q = Q0; q < NW; ++q) {
                        const unsigned long long ww = pack2(w[q], w[q]);
#pragma unroll
                        for (int i = 0; i < C; ++i) {
                            fma2(acc[q][i].lo, ww, xlo[i]);
                            fma2(acc[q][i].hi, ww, xhi[i]);
                        }
                    }
                }
            }
        }
        // tail of the tile, one row at a time (same order of accumulation: ascending node index)
        for (; t < tn; t += G) {
            const uint32_t o = s_off[t];
            float4 xv[C];
#pragma unroll
            for (int i = 0; i < C; ++i)
                xv[i] = (FULL || cx.colok[i]) ? __ldg(cx.xcol[i] + o) : make_float4(0.f, 0.f, 0.f, 0.f);
            float w[NW];
            load_weights<K1, SC, KMIN>(w, s_w + t * NWP);
            if (kUseF32x2) {
#pragma unroll
                for (int q = Q0; q < NW; ++q) {
                    const unsigned long long ww = pack2(w[q], w[q]);
#pragma unroll
                    for (int i = 0; i < C; ++i) {
                        fma2(acc[q][i].lo, ww, pack2(xv[i].x, xv[i].y));
                        fma2(acc[q][i].hi, ww, pack2(xv[i].z, xv[i].w));
                    }
                }
            }
        }
    }
}

template <int K1, int SC, int C, int KMIN, bool FULL>
struct RangeDispatch {
    __device__ __forceinline__ static void run(int kmin, Acc (&acc)[K1 * SC][C], int lo, int hi, const GatherCtx<C>& cx) {
        if (kmin == KMIN)
            accumulate_range<K1, SC, C, KMIN, FULL>(acc, lo, hi, cx);
        else
            RangeDispatch<K1, SC, C, KMIN + 1, FULL>::run(kmin, acc, lo, hi, cx);
    }
};
template <int K1, int SC, int C, bool FULL>
struct RangeDispatch<K1, SC, C, K1, FULL> {
    __device__ __forceinline__ static void run(int, Acc (&)[K1 * SC][C], int, int, const GatherCtx<C>&) {}
};

// Occupancy target by accumulator footprint (K1*SC*C float4 per thread): the kernel is latency /
// L2-throughput bound, and 8 CTAs of 4 warps per SM stream 16 TB/s where 5 CTAs streamed 12 TB/s.
constexpr int gather_min_blocks(int acc4) { return acc4 <= 8 ? S3_GATHER_BLOCKS : acc4 <= 12 ? 6 : acc4 <= 16 ? 5 : acc4 <= 24 ? 3 : 2; }

template <int K1, int SC, int C, bool FULL>
__global__ void __launch_bounds__(kGatherThreads, gather_min_blocks(K1 * SC * C)) gather_kernel(GatherParams p) {
    constexpr int NW = K1 * SC, NWP = (NW + 3) & ~3;
    extern __shared__ float4 smem4[];
    float* s_w = reinterpret_cast<float*>(smem4);                        // [kTile][NWP]
    uint32_t* s_off = reinterpret_cast<uint32_t*>(s_w + kTile * NWP);    // [kTile]

    const int tid = threadIdx.x;
    // record mode: one CTA per record (largest first when an order is given), its seed rows;
    // CCN mode: one CTA per CCN work item = up to SC extra selected rows of a record
    int64_t rec;
    int chunk = 0;
    if (p.ccn) {
        const int64_t item = blockIdx.x;
        rec = p.item_rec[item];
        chunk = (int)(item - p.item_ptr[rec]);
    } else {
        rec = p.order ? (int64_t)p.order[blockIdx.x] : (int64_t)blockIdx.x;
        if (rec < 0) return;  // slot of an invalid record
    }
    const int32_t* cnt = p.cnt + rec * S3_NCNT;
    if (cnt[S3_CNT_STATUS] != S3_REC_OK) return;
    const int n = cnt[S3_CNT_N], s = cnt[S3_CNT_S];
    const int nseed = num_seeds(p.flow);
    const int64_t* off = p.off + rec * S3_NOFF;
    const int32_t* nodes = p.arena + off[S3_OFF_NODES];
    const float* item_f = reinterpret_cast<const float*>(p.arena + off[S3_OFF_F32]) +
                          (p.ccn ? item_words(p.flow, p.sign_k, n) + (int64_t)chunk * ccn_item_words(p.sign_k, n, SC) : 0);
    const float* lab = item_f;
    const float4* wgt4 = reinterpret_cast<const float4*>(item_f + NWP);

    const int tpr = p.tpr, G = kGatherThreads / tpr;
    const int grp = tid / tpr, lane = tid - grp * tpr;
    GatherCtx<C> cx;
    int col[C];
    bool (&colok)[C] = cx.colok;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        col[i] = (blockIdx.y * C + i) * tpr + lane;
        colok[i] = col[i] < p.F4;
        cx.xcol[i] = reinterpret_cast<const float4*>(p.x) + (colok[i] ? col[i] : 0);
    }
    cx.nodes = nodes;
    cx.wgt4 = wgt4;
    cx.ldx4 = (uint32_t)(p.ldx >> 2);
    cx.s_w = s_w;
    cx.s_off = s_off;
    cx.tid = tid;
    cx.grp = grp;
    cx.G = G;
    cx.swap01 = SC == 2 && !p.ccn && p.flow == S3_FLOW_POS && n >= 2 && nodes[0] > nodes[1];

    Acc acc[NW][C];
#pragma unroll
    for (int q = 0; q < NW; ++q)
#pragma unroll
        for (int i = 0; i < C; ++i) acc[q][i].lo = acc[q][i].hi = 0ull;  // (0.f, 0.f)

    // hop ranges of the canonical node order; the records' own rows are the seeds, CCN rows are
    // hop-1 nodes (one hop closer to everything: kmin shifts down by one)
    const int shift = p.ccn ? 1 : 0;
    int lo = 0;
    for (int l = 0; l <= S3_MAX_HOPS && lo < n; ++l) {
        const int hi = lo + cnt[S3_CNT_HOP0 + l];
        const int kmin = max(0, l - shift);
        if (kmin >= K1) break;  // farther than K hops: no operator reaches these nodes
        if (hi > lo)
            RangeDispatch<K1, SC, C, 0, FULL>::run(kmin, acc, lo, hi, cx);
        lo = hi;
    }

    // combine the G row groups in fixed order (group 0 accumulates groups 1..G-1)
    if (G > 1) {
        float4* red = smem4;  // [(G-1)][NW][C][tpr]
        __syncthreads();
        if (grp > 0) {
#pragma unroll
            for (int q = 0; q < NW; ++q)
#pragma unroll
                for (int i = 0; i < C; ++i) red[(((grp - 1) * NW + q) * C + i) * tpr + lane] = acc_to_float4(acc[q][i]);
        }
        __syncthreads();
        if (grp == 0) {
            for (int g = 1; g < G; ++g)
#pragma unroll
                for (int q = 0; q < NW; ++q)
#pragma unroll
                    for (int i = 0; i < C; ++i) {
                        const float4 v = red[(((g - 1) * NW + q) * C + i) * tpr + lane];
                        float4 a = acc_to_float4(acc[q][i]);
                        a.x += v.x;
                        a.y += v.y;
                        a.z += v.z;
                        a.w += v.w;
                        acc[q][i].lo = pack2(a.x, a.y);
                        acc[q][i].hi = pack2(a.z, a.w);
                    }
        }
    }

    // ---- epilogue: every output row is staged through shared memory exactly as it lies in memory ([label | features],
    // rows are only 4-byte aligned: ldo = F + 1) and stored with 16-byte vectors from the first 16-byte boundary on
    // (scalar head / tail), to this GPU's operator matrices or, for s3_gather_peers, to every GPU's — full 128-bit
    // stores are what NVLink needs: 4-byte stores reached 150 GB/s per GPU on 8 x B200. Chain members of a paired
    // link get the same rows (seed rows exchanged for the opposite direction).
    const int first_sel = p.ccn ? nseed + chunk * SC : 0;  // index of this item's first selected row
    const int rpl = p.flow == S3_FLOW_SOP ? 2 : 1;  // records per link (SoP: one per endpoint)
    const int64_t gl = p.out_link ? p.out_link[rec / rpl] : p.link_base + rec / rpl;  // global link index (fixed-row flows)
    const int64_t row0 = (p.out_link ? (gl * rpl + rec % rpl) * nseed
                                     : p.row_base + (p.row_ptr ? p.row_ptr[rec] : rec * (int64_t)nseed)) + first_sel;
    const long long chain = (p.mirror && !p.ccn && !p.skip_chain) ? (long long)p.mirror[gl] : -1;
    float* s_row = reinterpret_cast<float*>(smem4);  // element e of the output row sits at s_row[e - f0]
    const int f0 = blockIdx.y * C * tpr * 4;         // first feature of this CTA's column chunk
    const int nfl = min(C * tpr * 4, p.F - f0);
    const int e_lo = blockIdx.y == 0 ? 0 : 1 + f0;   // elements [e_lo, e_hi) of every row are this CTA's
    const int e_hi = 1 + f0 + nfl;
    const int ndst = p.num_dst > 0 ? p.num_dst : 1;
#pragma unroll
    for (int q = 0; q < NW; ++q) {
        const int k = q / SC, c = q - k * SC;
        const bool live = first_sel + c < s && !(p.skip_op0 && k == 0);  // uniform over the CTA
        __syncthreads();
        if (live && grp == 0) {
#pragma unroll
            for (int i = 0; i < C; ++i) {
                float* d = s_row + 1 + 4 * (i * tpr + lane);
                const float4 a = acc_to_float4(acc[q][i]);
                d[0] = a.x;
                d[1] = a.y;
                d[2] = a.z;
                d[3] = a.w;
            }
            if (lane == 0) s_row[0] = lab[q];  // the label / self-return column (read by chunk 0 only)
        }
        __syncthreads();
        if (!live || e_hi <= e_lo) continue;
        int64_t row = row0 + c;
        long long m = chain;
        for (;;) {
            for (int d = 0; d < ndst; ++d) {
                float* orow = (p.num_dst > 0 ? p.dst_base[d] + (int64_t)k * p.op_stride : p.out.p[k]) + row * p.ldo;
                const int mis = (int)((reinterpret_cast<uintptr_t>(orow + e_lo) >> 2) & 3);
                const int head = min((4 - mis) & 3, e_hi - e_lo);
                const int nvec = (e_hi - e_lo - head) >> 2;
                const int ev = e_lo + head;  // first element on a 16-byte boundary
                for (int v = tid; v < nvec; v += kGatherThreads) {
                    const float* sp = s_row + (ev + 4 * v - f0);
                    *reinterpret_cast<float4*>(orow + ev + 4 * v) = make_float4(sp[0], sp[1], sp[2], sp[3]);
                }
                const int et = ev + 4 * nvec;  // scalar head [e_lo, ev) and tail [et, e_hi)
                if (tid < head) orow[e_lo + tid] = s_row[e_lo + tid - f0];
                if (tid >= 32 && tid - 32 < e_hi - et) orow[et + tid - 32] = s_row[et + tid - 32 - f0];
            }
            if (m < 0) break;
            const long long v = -2 - (long long)p.mirror[m];
            row = (int64_t)m * nseed + ((v & 1) ? SC - 1 - c : c);
            m = (v >> 1) - 1;
        }
    }
}

template <int K1, int SC, int C, bool FULL>
cudaError_t launch_full(const GatherParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(gather_kernel<K1, SC, C, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    gather_kernel<K1, SC, C, FULL><<<grid, kGatherThreads, smem, st>>>(p);
    return cudaGetLastError();
}

template <int K1, int SC, int C>
cudaError_t launch_one(const GatherParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    const bool full = p.F4 % (p.tpr * C) == 0;  // every lane of every column chunk owns a valid column
    return full ? launch_full<K1, SC, C, true>(p, grid, smem, st) : launch_full<K1, SC, C, false>(p, grid, smem, st);
}

template <int K1, int SC>
cudaError_t launch_c(const GatherParams& p, int C, dim3 grid, size_t smem, cudaStream_t st) {
    if (SC == 8) return C == 1 ? launch_one<K1, SC, 1>(p, grid, smem, st) : cudaErrorInvalidValue;  // 8-row items: 1 column/thread
    switch (C) {
        case 1: return launch_one<K1, SC, 1>(p, grid, smem, st);
        case 2: return launch_one<K1, SC, 2>(p, grid, smem, st);
        default: return launch_one<K1, SC, 3>(p, grid, smem, st);
    }
}

template <int SC, int K1>
cudaError_t launch_k1(const GatherParams& p, int C, dim3 grid, size_t smem, cudaStream_t st) {
    return launch_c<K1, SC>(p, C, grid, smem, st);
}

}  // namespace
}  // namespace s3
