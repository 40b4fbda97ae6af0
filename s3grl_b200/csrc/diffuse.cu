// Kernel 2 — per-subgraph diffusion weights: the K row vectors e_sel^T S^k, S = D^-1/2 A_sub D^-1/2.
//
// Replaces reference tuned_SIGN.py:155-161 (degree normalisation), :168-170 (S^2..S^K by
// SpGEMM) and :173-175 (keep the selected rows); in SoP flow sgrl_link_pred.py:161-178
// (global normalised powers) and tuned_SIGN.py:60-86, :106-113 (rows of the endpoints, partner
// entry zeroed, self-return weight).
//
// The reference forms whole n x n powers and keeps 2 (+CCN) rows. Only those rows are needed:
// with w_0 = e_sel and z_k = w_k D^-1/2,
//      t_j = sum_{i in N(j)} z_{k-1}[i],   w_k[j] = dis_j t_j,   z_k[j] = dis_j^2 t_j
// (A_sub symmetric), i.e. K segmented CSR SpMV sweeps per selected row instead of K-1 SpGEMMs.
// A k-step walk cannot leave the k-hop ball of the selected rows, so sweep k only visits the
// nodes of hops <= k (+1 when the selected rows are hop-1 CCN nodes) of the hop-major node
// order; everything beyond is an exact zero that kernel 3 never reads.
//
// The rows of the two targets are diffused inside the front kernel (extract.cu); this kernel
// handles the PoS Plus CCN rows — extra work items of up to 8 selected rows each — over the fully
// stored padded CSR.
// One CTA per CCN work item, one 8-lane group per node:
// lanes stride the node's local row, partial sums are combined by a fixed shuffle tree, so
// results do not depend on scheduling. z ping-pong buffers live in shared memory when the
// subgraph fits, else in the item's global scratch. Output per item, in the record's float
// scratch:  labels[NWP] | weights[n][NWP] | z ping | z pong,  weight column q = k*SC + c for
// operator k (k = 0 is the one-hot selecting x itself) and selected row c.
#include "common.cuh"

namespace s3 {
namespace {

constexpr int kDiffLanes = 4;  // lanes per row (rows of the reference's graphs are short: 4 leave fewer lanes idle than 8)
// floats per shared z buffer (two buffers, dynamic shared memory): n <= 2048 with 2 rows, 512 with 8. Measured in
// round 2: 96 KB for the 8-row union items (n <= 1536, two CTAs per SM) was SLOWER than 32 KB with z of the larger
// items in global memory (306 vs 276 ms per PubMed step): the sweeps are latency bound and want resident CTAs.
__host__ __device__ constexpr int z_cap(int) { return 4096; }

struct DiffuseParams {
    const int64_t* __restrict__ indptr;  // SoP: global degrees
    int32_t* arena;
    const int64_t* __restrict__ off;
    const int32_t* __restrict__ cnt;
    const int64_t* __restrict__ item_ptr;  // may be null (one item per record)
    const int32_t* __restrict__ item_rec;  // may be null
    int flow, sign_k;
};

template <int SC>
__global__ void __launch_bounds__(kDiffuseThreads) diffuse_kernel(DiffuseParams p) {
    extern __shared__ float s_zbuf[];  // [2][z_cap(SC)]
    constexpr int kZCap = z_cap(SC);
    const int tid = threadIdx.x, T = kDiffuseThreads;
    const int l8 = tid & (kDiffLanes - 1), grp = tid / kDiffLanes;  // lane group of a row
    constexpr int NG = kDiffuseThreads / kDiffLanes;
    const int64_t item = blockIdx.x;
    const int64_t rec = p.item_rec ? p.item_rec[item] : item;
    const int32_t* cnt = p.cnt + rec * S3_NCNT;
    if (cnt[S3_CNT_STATUS] != S3_REC_OK) return;
    const int chunk = (int)(item - p.item_ptr[rec]);  // CCN chunk: selected rows nseed + chunk*8 ...
    const int n = cnt[S3_CNT_N], s = cnt[S3_CNT_S];
    const int K = p.sign_k, nseed = num_seeds(p.flow);
    const int NW = (K + 1) * SC, NWP = (NW + 3) & ~3;
    const int64_t* off = p.off + rec * S3_NOFF;
    const int32_t* nodes = p.arena + off[S3_OFF_NODES];
    const int32_t* rowptr = p.arena + off[S3_OFF_ROWPTR];
    const int32_t* rowlen = p.arena + off[S3_OFF_ROWLEN];
    const int32_t* lcol = p.arena + off[S3_OFF_LCOL];
    const int32_t* sel = p.arena + off[S3_OFF_SEL];
    float* item_f = reinterpret_cast<float*>(p.arena + off[S3_OFF_F32]) + item_words(p.flow, K, n) +
                    (int64_t)chunk * ccn_item_words(K, n, SC);
    float* lab = item_f;
    float* wgt = item_f + NWP;
    const bool z_shared = (int64_t)n * SC <= kZCap;
    float* zA = z_shared ? s_zbuf : wgt + (int64_t)n * NWP;
    float* zB = z_shared ? s_zbuf + kZCap : wgt + (int64_t)n * NWP + (int64_t)n * SC;

    // selected local rows of this item
    int r[SC];
#pragma unroll
    for (int c = 0; c < SC; ++c) {
        const int i = nseed + chunk * SC + c;
        r[c] = i >= s ? -1 : sel[i - nseed];
    }
    // reach[k] = number of leading nodes (hop-major order) a k-step walk from the selected rows
    // can touch: hops <= k for seed rows, hops <= k+1 for hop-1 (CCN) rows
    const int shift = 1;  // CCN rows are hop-1 nodes
    int hop_end[S3_MAX_HOPS + 2];
    {
        int acc = 0;
#pragma unroll
        for (int l = 0; l <= S3_MAX_HOPS; ++l) {
            acc += cnt[S3_CNT_HOP0 + l];
            hop_end[l] = acc;
        }
        hop_end[S3_MAX_HOPS + 1] = acc;
    }
    auto reach = [&](int k) { return hop_end[min(k + shift, S3_MAX_HOPS + 1)]; };

    // k = 0: one-hot weights on the reachable prefix, z_0 = e_sel D^-1/2 on ALL nodes (both
    // buffers must be zero beyond what later sweeps overwrite)
    const int n0 = reach(0);
    for (int j = tid; j < n; j += T) {
        float zv[SC];
#pragma unroll
        for (int c = 0; c < SC; ++c) zv[c] = 0.0f;
        if (j < n0) {
            int deg;
            if (p.flow == S3_FLOW_SOP) {
                const int g = nodes[j];
                deg = (int)(p.indptr[g + 1] - p.indptr[g]);  // global degree (sgrl_link_pred.py:165-168)
            } else {
                deg = rowlen[j];  // induced, masked degree (tuned_SIGN.py:158)
            }
            const float dis = deg > 0 ? 1.0f / sqrtf((float)deg) : 0.0f;  // inf -> 0 (tuned_SIGN.py:159-160)
            float* wj = wgt + (int64_t)j * NWP;
#pragma unroll
            for (int c = 0; c < SC; ++c) {
                const float one = (j == r[c]) ? 1.0f : 0.0f;
                wj[c] = one;
                zv[c] = one * dis;
            }
        }
#pragma unroll
        for (int c = 0; c < SC; ++c) {
            zA[(int64_t)j * SC + c] = zv[c];
            zB[(int64_t)j * SC + c] = 0.0f;
        }
    }
    __syncthreads();

    float* zprev = zA;
    float* znext = zB;
    for (int k = 1; k <= K; ++k) {
        const int nk = reach(k);
        for (int jb = 0; jb < nk; jb += NG) {
            const int j = jb + grp;
            const bool valid = j < nk;
            int e0 = 0, e1 = 0;
            if (valid) {
                e0 = rowptr[j];
                e1 = rowptr[j + 1];
            }
            float t[SC];
#pragma unroll
            for (int c = 0; c < SC; ++c) t[c] = 0.0f;
            for (int e = e0 + l8; e < e1; e += kDiffLanes) {
                const int i = lcol[e];  // -1: neighbour outside the subgraph / masked target link
                if (i >= 0) {
#pragma unroll
                    for (int c = 0; c < SC; ++c) t[c] += zprev[(int64_t)i * SC + c];
                }
            }
#pragma unroll
            for (int c = 0; c < SC; ++c) {
#pragma unroll
                for (int d = kDiffLanes / 2; d > 0; d >>= 1) t[c] += __shfl_xor_sync(0xffffffffu, t[c], d);
            }
            if (valid && l8 == 0) {
                int deg = rowlen[j];
                if (p.flow == S3_FLOW_SOP) {
                    const int g = nodes[j];
                    deg = (int)(p.indptr[g + 1] - p.indptr[g]);
                }
                const float dis = deg > 0 ? 1.0f / sqrtf((float)deg) : 0.0f;
#pragma unroll
                for (int c = 0; c < SC; ++c) {
                    const float w = dis * t[c];
                    wgt[(int64_t)j * NWP + k * SC + c] = w;
                    znext[(int64_t)j * SC + c] = dis * w;
                }
            }
        }
        __syncthreads();
        float* tmp = zprev;
        zprev = znext;
        znext = tmp;
    }

    // label / self-return column of every operator, then (SoP) drop the partner's weight
    if (p.flow == S3_FLOW_POS) {
        // x_k[sel, 0] = sum_j w_k[j] * label_j, label = 1 on local 0 and 1 (tuned_SIGN.py:177).
        // Weights of operator k exist on the first reach(k) nodes only.
        for (int q = tid; q < NWP; q += T) {
            float v = 0.0f;
            if (q < NW) {
                const int nk = reach(q / SC);
                if (nk > 0) v += wgt[q];
                if (nk > 1) v += wgt[NWP + q];
            }
            lab[q] = v;
        }
    } else {
        // x_k[., 0] = A^k[u,u] (tuned_SIGN.py:106-113); x[., 0] = 1 (tuned_SIGN.py:119-124)
        for (int q = tid; q < NWP; q += T) lab[q] = q < NW ? wgt[q] : 0.0f;
        __syncthreads();
        const int partner = cnt[S3_CNT_PARTNER];
        if (partner >= 0)  // r_u[v] = 0 (tuned_SIGN.py:73-76); columns k < hop(partner) are never read
            for (int q = SC + tid; q < NW; q += T) wgt[(int64_t)partner * NWP + q] = 0.0f;
    }
}

}  // namespace

cudaError_t launch_diffuse(const s3_graph& g, const s3_batch& b, int64_t num_items, cudaStream_t st) {
    if (num_items == 0) return cudaSuccess;
    DiffuseParams p;
    p.indptr = g.indptr;
    p.arena = b.arena;
    p.off = b.off;
    p.cnt = b.cnt;
    if (!b.item_rec || !b.item_ptr) return cudaErrorInvalidValue;  // CCN items need the plan
    p.item_ptr = b.item_ptr;
    p.item_rec = b.item_rec;
    p.flow = b.flow;
    p.sign_k = b.sign_k;
    if (ccn_rows(b.strategy) == 8) {
        static LaunchCache cache;  // the > 48 KB shared-memory opt-in is per device
        constexpr size_t smem = 2 * (size_t)z_cap(8) * sizeof(float);
        cudaError_t e = cache.get(reinterpret_cast<const void*>(diffuse_kernel<8>), kDiffuseThreads, smem, nullptr, nullptr);
        if (e != cudaSuccess) return e;
        diffuse_kernel<8><<<(unsigned)num_items, kDiffuseThreads, smem, st>>>(p);
    } else {
        diffuse_kernel<2><<<(unsigned)num_items, kDiffuseThreads, 2 * (size_t)z_cap(2) * sizeof(float), st>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace s3
