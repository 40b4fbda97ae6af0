// Kernel 3 launcher: parameter set-up and dispatch to the per-SC translation units
// (gather_sc*.cu). The kernel itself is in gather_kernel.cuh.
#include "gather_kernel.cuh"

namespace s3 {

cudaError_t launch_gather(const s3_graph& g, const s3_batch& b, int64_t num_items, const OutPtrs& out, int64_t ldo,
                          int64_t row_base, bool ccn, cudaStream_t st, const PeerDst* peers) {
    if (num_items == 0) return cudaSuccess;
    if (g.num_nodes * (g.ldx / 4) >= (int64_t(1) << 32)) return cudaErrorInvalidValue;  // 32-bit row offsets
    GatherParams p;
    p.x = g.x;
    p.ldx = g.ldx;
    p.F = (int)g.num_feat;
    p.F4 = (int)(g.ldx / 4);
    p.arena = b.arena;
    p.off = b.off;
    p.cnt = b.cnt;
    p.row_ptr = b.row_ptr;
    if (ccn && (!b.item_rec || !b.item_ptr || !b.row_ptr || b.flow != S3_FLOW_POS)) return cudaErrorInvalidValue;
    p.item_ptr = b.item_ptr;
    p.item_rec = b.item_rec;
    p.order = b.order;
    p.ccn = ccn ? 1 : 0;
    p.flow = b.flow;
    p.sign_k = b.sign_k;
    p.out = out;
    p.ldo = ldo;
    p.row_base = row_base;
    p.out_link = ccn ? nullptr : b.out_link;
    p.mirror = (!ccn && batch_pairing(b)) ? b.mirror : nullptr;
    p.link_base = b.link_base;
    p.num_dst = 0;
    p.skip_op0 = 0;
    p.skip_chain = 0;
    p.op_stride = 0;
    for (int d = 0; d < S3_MAX_PEERS; ++d) p.dst_base[d] = nullptr;
    if (peers) {
        if (ccn || peers->num_dst < 1 || peers->num_dst > S3_MAX_PEERS) return cudaErrorInvalidValue;
        p.num_dst = peers->num_dst;
        p.op_stride = peers->op_stride;
        p.skip_op0 = peers->skip_op0;
        p.skip_chain = peers->skip_chain;
        for (int d = 0; d < peers->num_dst; ++d) p.dst_base[d] = peers->base[d];
    }

    int C, tpr;
    const int F4 = p.F4;
    const bool wide_items = ccn && ccn_rows(b.strategy) == 8;  // 8-row items: one column per thread, more column chunks
    if (F4 <= 32) { C = 1; tpr = 32; }
    else if (F4 <= 64) { C = 1; tpr = 64; }
    else if (F4 <= 128) { C = 1; tpr = 128; }
    else if (wide_items) { C = 1; tpr = 128; }
    else if (F4 <= 256) { C = 2; tpr = 128; }
    else { C = 3; tpr = 128; }
    p.tpr = tpr;
    const int colchunks = (F4 + tpr * C - 1) / (tpr * C);
    const int sc = ccn ? ccn_rows(b.strategy) : sel_chunk(b.flow);
    const int NW = (b.sign_k + 1) * sc;
    const int NWP = (NW + 3) & ~3;
    const int G = kGatherThreads / tpr;
    const size_t tile_bytes = (size_t)kTile * NWP * 4 + kTile * 4;
    const size_t red_bytes = (size_t)(G - 1) * NW * C * tpr * 16;
    const size_t row_bytes = (size_t)C * tpr * 16 + 32;  // one staged output row of the column chunk (+ label, slack)
    size_t smem = tile_bytes > red_bytes ? tile_bytes : red_bytes;
    if (row_bytes > smem) smem = row_bytes;
    dim3 grid((unsigned)num_items, (unsigned)colchunks);
    const int K1 = b.sign_k + 1;
    if (sc == 8) {  // union with sign_k > 5 is not instantiated
        switch (K1) {
            case 2: return launch_gather_sc8_k2(p, C, grid, smem, st);
            case 3: return launch_gather_sc8_k3(p, C, grid, smem, st);
            case 4: return launch_gather_sc8_k4(p, C, grid, smem, st);
            case 5: return launch_gather_sc8_k5(p, C, grid, smem, st);
            case 6: return launch_gather_sc8_k6(p, C, grid, smem, st);
            default: return cudaErrorInvalidValue;
        }
    }
    if (sc == 2)
        return K1 <= 4 ? launch_gather_sc2_lo(p, K1, C, grid, smem, st)
                       : K1 <= 6 ? launch_gather_sc2_mid(p, K1, C, grid, smem, st) : launch_gather_sc2_hi(p, K1, C, grid, smem, st);
    return K1 <= 4 ? launch_gather_sc1_lo(p, K1, C, grid, smem, st)
                   : K1 <= 6 ? launch_gather_sc1_mid(p, K1, C, grid, smem, st) : launch_gather_sc1_hi(p, K1, C, grid, smem, st);
}

namespace {
// x (operator 0) of the fixed-row flows for a whole link list: row 2i = [1 | X[src_i]], row 2i+1 = [1 | X[dst_i]]
// (tuned_SIGN.py:181 x = subg_x[[0,1]] with the zero-one label; SoP tuned_SIGN.py:119-124). Plain copy, one CTA per row.
__global__ void __launch_bounds__(128) fill_x0_kernel(const float* __restrict__ x, int64_t ldx, int F, int64_t num_nodes,
                                                      const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                                      float* __restrict__ out, int64_t ldo) {
    const int64_t row = blockIdx.x, link = row >> 1;
    const int64_t a = src[link], b = dst[link];
    if (a < 0 || b < 0 || a >= num_nodes || b >= num_nodes || a == b) return;  // invalid link: flagged by s3_extract
    const float* xr = x + ((row & 1) ? b : a) * ldx;
    float* o = out + row * ldo;
    if (threadIdx.x == 0) o[0] = 1.0f;
    for (int f = threadIdx.x; f < F; f += 128) o[1 + f] = __ldg(xr + f);
}
}  // namespace

namespace {
// Rows of paired links, locally: one CTA per link that heads a chain copies its two rows of operators first_op..
// to every chain member (rows exchanged for the opposite direction). Plain copy, consecutive lanes on consecutive
// floats.
__global__ void __launch_bounds__(128) fill_mirrors_kernel(const int64_t* __restrict__ mirror, OutPtrs ops, int first_op,
                                                           int num_ops, int cols, int64_t ldo) {
    const int64_t p = blockIdx.x;
    long long m = mirror[p];
    if (m < 0) return;  // unpaired, or a member itself
    while (m >= 0) {
        const long long v = -2 - (long long)mirror[m];
        const int swap = (int)(v & 1);
        for (int k = first_op; k < num_ops; ++k) {
            const float* src = ops.p[k] + 2 * p * ldo;
            float* dst = ops.p[k] + 2 * m * ldo;
            for (int i = threadIdx.x; i < 2 * cols; i += 128) {
                const int r = i >= cols ? 1 : 0, c = i - r * cols;
                dst[(swap ? 1 - r : r) * ldo + c] = src[r * ldo + c];
            }
        }
        m = (v >> 1) - 1;
    }
}
}  // namespace

cudaError_t launch_fill_mirrors(const int64_t* mirror, int64_t num_links, const OutPtrs& ops, int first_op, int num_ops,
                                int64_t cols, int64_t ldo, cudaStream_t st) {
    if (num_links == 0) return cudaSuccess;
    if (num_links > 0x7fffffff) return cudaErrorInvalidValue;
    fill_mirrors_kernel<<<(unsigned)num_links, 128, 0, st>>>(mirror, ops, first_op, num_ops, (int)cols, ldo);
    return cudaGetLastError();
}

cudaError_t launch_fill_x0(const s3_graph& g, const int64_t* src, const int64_t* dst, int64_t num_links, float* out, int64_t ldo,
                           cudaStream_t st) {
    if (num_links == 0) return cudaSuccess;
    if (2 * num_links > 0x7fffffff) return cudaErrorInvalidValue;
    fill_x0_kernel<<<(unsigned)(2 * num_links), 128, 0, st>>>(g.x, g.ldx, (int)g.num_feat, g.num_nodes, src, dst, out, ldo);
    return cudaGetLastError();
}

}  // namespace s3
