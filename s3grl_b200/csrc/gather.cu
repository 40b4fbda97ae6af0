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
// registers; weights and node ids are staged through shared memory in tiles. The sum over j
// runs in canonical node order within each row group, groups are combined in fixed order, so
// results are independent of scheduling, batch composition and GPU count.
//
// HBM/L2-bound streaming gather: no tensor cores (M = NW <= 16 rows, fp32 required by the
// 1e-5 tolerance).
#include "common.cuh"

namespace s3 {
namespace {

constexpr int kTile = 64;  // nodes staged per tile

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
    int flow, sign_k, sc, tpr;             // tpr = threads per feature row (32/64/128)
    OutPtrs out;
    int64_t ldo, row_base;
};

template <int NW, int C>
__global__ void __launch_bounds__(kGatherThreads) gather_kernel(GatherParams p) {
    constexpr int NWP = (NW + 3) & ~3;
    extern __shared__ float4 smem4[];
    float* s_w = reinterpret_cast<float*>(smem4);                 // [kTile][NWP]
    int* s_gid = reinterpret_cast<int*>(s_w + kTile * NWP);       // [kTile]

    const int tid = threadIdx.x;
    const int64_t item = blockIdx.x;
    const int64_t rec = p.item_rec ? p.item_rec[item] : item;
    const int32_t* cnt = p.cnt + rec * S3_NCNT;
    if (cnt[S3_CNT_STATUS] != S3_REC_OK) return;
    const int chunk = p.item_ptr ? (int)(item - p.item_ptr[rec]) : 0;
    const int n = cnt[S3_CNT_N], s = cnt[S3_CNT_S];
    const int64_t* off = p.off + rec * S3_NOFF;
    const int32_t* nodes = p.arena + off[S3_OFF_NODES];
    const float* item_f = reinterpret_cast<const float*>(p.arena + off[S3_OFF_F32]) +
                          (int64_t)chunk * item_words(p.flow, p.sign_k, n);
    const float* lab = item_f;
    const float4* wgt4 = reinterpret_cast<const float4*>(item_f + NWP);

    const int tpr = p.tpr, G = kGatherThreads / tpr;
    const int grp = tid / tpr, lane = tid - grp * tpr;
    int col[C];
    bool colok[C];
#pragma unroll
    for (int i = 0; i < C; ++i) {
        col[i] = (blockIdx.y * C + i) * tpr + lane;
        colok[i] = col[i] < p.F4;
    }

    float4 acc[NW][C];
#pragma unroll
    for (int q = 0; q < NW; ++q)
#pragma unroll
        for (int i = 0; i < C; ++i) acc[q][i] = make_float4(0.f, 0.f, 0.f, 0.f);

    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(p.x);
    const int64_t ldx4 = p.ldx >> 2;

    for (int base = 0; base < n; base += kTile) {
        const int tn = min(kTile, n - base);
        __syncthreads();
        if (tid < tn) s_gid[tid] = nodes[base + tid];
        {
            const float4* src = wgt4 + (int64_t)base * (NWP / 4);
            float4* dst = reinterpret_cast<float4*>(s_w);
            for (int i = tid; i < tn * (NWP / 4); i += kGatherThreads) dst[i] = src[i];
        }
        __syncthreads();

        constexpr int U = 4;
        for (int t0 = grp; t0 < tn; t0 += U * G) {
            float4 xv[U][C];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = t0 + u * G;
                const bool ok = t < tn;
                const int64_t rowoff = ok ? (int64_t)s_gid[t] * ldx4 : 0;
#pragma unroll
                for (int i = 0; i < C; ++i)
                    xv[u][i] = (ok && colok[i]) ? __ldg(x4 + rowoff + col[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = min(t0 + u * G, tn - 1);  // out-of-range slots carry x = 0
                const float* wrow = s_w + t * NWP;
                float w[NWP];
#pragma unroll
                for (int q4 = 0; q4 < NWP / 4; ++q4) {
                    const float4 v = reinterpret_cast<const float4*>(wrow)[q4];
                    w[4 * q4] = v.x;
                    w[4 * q4 + 1] = v.y;
                    w[4 * q4 + 2] = v.z;
                    w[4 * q4 + 3] = v.w;
                }
#pragma unroll
                for (int q = 0; q < NW; ++q)
#pragma unroll
                    for (int i = 0; i < C; ++i) {
                        acc[q][i].x = fmaf(w[q], xv[u][i].x, acc[q][i].x);
                        acc[q][i].y = fmaf(w[q], xv[u][i].y, acc[q][i].y);
                        acc[q][i].z = fmaf(w[q], xv[u][i].z, acc[q][i].z);
                        acc[q][i].w = fmaf(w[q], xv[u][i].w, acc[q][i].w);
                    }
            }
        }
    }

    // combine the G row groups in fixed order (group 0 accumulates groups 1..G-1)
    if (G > 1) {
        static_assert(C == 1 || true, "");
        float4* red = smem4;  // [(G-1)][NW][C][tpr]
        __syncthreads();
        if (grp > 0) {
#pragma unroll
            for (int q = 0; q < NW; ++q)
#pragma unroll
                for (int i = 0; i < C; ++i) red[(((grp - 1) * NW + q) * C + i) * tpr + lane] = acc[q][i];
        }
        __syncthreads();
        if (grp == 0) {
            for (int g = 1; g < G; ++g)
#pragma unroll
                for (int q = 0; q < NW; ++q)
#pragma unroll
                    for (int i = 0; i < C; ++i) {
                        const float4 v = red[(((g - 1) * NW + q) * C + i) * tpr + lane];
                        acc[q][i].x += v.x;
                        acc[q][i].y += v.y;
                        acc[q][i].z += v.z;
                        acc[q][i].w += v.w;
                    }
        }
    }
    if (grp != 0) return;

    const int sc = p.sc;
    const int64_t row0 = p.row_base + (p.row_ptr ? p.row_ptr[rec] : rec * (int64_t)num_seeds(p.flow)) + (int64_t)chunk * sc;
#pragma unroll
    for (int q = 0; q < NW; ++q) {
        const int k = q / sc, c = q - k * sc;
        if (chunk * sc + c >= s) continue;
        float* orow = p.out.p[k] + (row0 + c) * p.ldo;
        if (blockIdx.y == 0 && lane == 0) orow[0] = lab[q];
#pragma unroll
        for (int i = 0; i < C; ++i) {
            if (!colok[i]) continue;
            const int f = 4 * col[i];
            float* o = orow + 1 + f;
            if (f + 0 < p.F) o[0] = acc[q][i].x;
            if (f + 1 < p.F) o[1] = acc[q][i].y;
            if (f + 2 < p.F) o[2] = acc[q][i].z;
            if (f + 3 < p.F) o[3] = acc[q][i].w;
        }
    }
}

template <int NW, int C>
cudaError_t launch_one(const GatherParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(gather_kernel<NW, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    gather_kernel<NW, C><<<grid, kGatherThreads, smem, st>>>(p);
    return cudaGetLastError();
}

template <int NW>
cudaError_t launch_c(const GatherParams& p, int C, dim3 grid, size_t smem, cudaStream_t st) {
    switch (C) {
        case 1: return launch_one<NW, 1>(p, grid, smem, st);
        case 2: return launch_one<NW, 2>(p, grid, smem, st);
        default: return launch_one<NW, 3>(p, grid, smem, st);
    }
}

}  // namespace

cudaError_t launch_gather(const s3_graph& g, const s3_batch& b, int64_t num_items, const OutPtrs& out, int64_t ldo,
                          int64_t row_base, cudaStream_t st) {
    if (num_items == 0) return cudaSuccess;
    GatherParams p;
    p.x = g.x;
    p.ldx = g.ldx;
    p.F = (int)g.num_feat;
    p.F4 = (int)(g.ldx / 4);
    p.arena = b.arena;
    p.off = b.off;
    p.cnt = b.cnt;
    p.row_ptr = b.row_ptr;
    p.item_ptr = b.item_rec ? b.item_ptr : nullptr;
    p.item_rec = b.item_rec;
    p.flow = b.flow;
    p.sign_k = b.sign_k;
    p.sc = sel_chunk(b.flow);
    p.out = out;
    p.ldo = ldo;
    p.row_base = row_base;

    int C, tpr;
    const int F4 = p.F4;
    if (F4 <= 32) { C = 1; tpr = 32; }
    else if (F4 <= 64) { C = 1; tpr = 64; }
    else if (F4 <= 128) { C = 1; tpr = 128; }
    else if (F4 <= 256) { C = 2; tpr = 128; }
    else { C = 3; tpr = 128; }
    p.tpr = tpr;
    const int colchunks = (F4 + tpr * C - 1) / (tpr * C);
    const int NW = weight_cols(b.flow, b.sign_k);
    const int NWP = (NW + 3) & ~3;
    const int G = kGatherThreads / tpr;
    const size_t tile_bytes = (size_t)kTile * NWP * 4 + kTile * 4;
    const size_t red_bytes = (size_t)(G - 1) * NW * C * tpr * 16;
    const size_t smem = tile_bytes > red_bytes ? tile_bytes : red_bytes;
    dim3 grid((unsigned)num_items, (unsigned)colchunks);
    switch (NW) {
        case 2: return launch_c<2>(p, C, grid, smem, st);
        case 3: return launch_c<3>(p, C, grid, smem, st);
        case 4: return launch_c<4>(p, C, grid, smem, st);
        case 5: return launch_c<5>(p, C, grid, smem, st);
        case 6: return launch_c<6>(p, C, grid, smem, st);
        case 7: return launch_c<7>(p, C, grid, smem, st);
        case 8: return launch_c<8>(p, C, grid, smem, st);
        case 10: return launch_c<10>(p, C, grid, smem, st);
        case 12: return launch_c<12>(p, C, grid, smem, st);
        case 14: return launch_c<14>(p, C, grid, smem, st);
        case 16: return launch_c<16>(p, C, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace s3
