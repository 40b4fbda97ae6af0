// Shared device helpers and launch declarations for libs3grl_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <mutex>

#include "../../include/s3grl_b200.h"

namespace s3 {

constexpr int kExtractThreads = 256;
constexpr int kDiffuseThreads = 256;
constexpr int kGatherThreads = 128;
constexpr int kMaxSmemBytes = 200 * 1024;  // bitmap tier budget (227 KB is the hardware limit)

// selected rows a work item carries; (K+1)*kSC weight columns per subgraph node
__host__ __device__ inline int sel_chunk(int flow) { return flow == S3_FLOW_SOP ? 1 : 2; }
__host__ __device__ inline int num_seeds(int flow) { return flow == S3_FLOW_SOP ? 1 : 2; }
__host__ __device__ inline int weight_cols(int flow, int K) { return (K + 1) * sel_chunk(flow); }
__host__ __device__ inline int weight_stride(int flow, int K) { return (weight_cols(flow, K) + 3) & ~3; }
// float words of one work item's scratch: [labels NWP][weights n*NWP][z ping n*SC][z pong n*SC]
__host__ __device__ inline int64_t item_words(int flow, int K, int64_t n) {
    const int64_t nwp = weight_stride(flow, K), sc = sel_chunk(flow);
    return (nwp + n * nwp + 2 * n * sc + 3) & ~int64_t(3);
}

// ---- block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32) ----
// `warp_sums` is caller-provided shared scratch of 33 ints. Returns the exclusive prefix of
// `v`; *total (same for all threads) is the block sum. Contains two __syncthreads().
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int s = lane < nw ? warp_sums[lane] : 0;
        int si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, si, d);
            if (lane >= d) si += t;
        }
        if (lane < nw) warp_sums[lane] = si - s;  // exclusive warp offsets
        if (lane == 31) warp_sums[32] = si;
    }
    __syncthreads();
    *total = warp_sums[32];
    const int out = warp_sums[wid] + inc - v;
    return out;
}

// PoS Plus: the CCN rows (selected rows beyond the seeds) are processed kCcnRows at a time as
// extra work items of the record; their scratch follows the record's own item.
// Intersections are small (mean 0.4 rows on PubMed): 2 rows per item; unions are large (mean 20):
// 8 rows per item, so that the subgraph's features are re-read 4x less often.
__host__ __device__ inline int ccn_rows(int strategy) { return strategy == S3_STRATEGY_UNION ? 8 : 2; }
__host__ __device__ inline int ccn_items(int s, int nseed, int cr) { return s > nseed ? (s - nseed + cr - 1) / cr : 0; }
__host__ __device__ inline int64_t ccn_item_words(int K, int64_t n, int cr) {
    const int64_t nwp = ((int64_t)(K + 1) * cr + 3) & ~int64_t(3);
    return (nwp + n * nwp + 2 * n * cr + 3) & ~int64_t(3);
}

// s3_ccn_chain (ccn_chain.cu): a record is served by the chain when its compact CSR (uint16 columns) and two
// [n][CW] operator buffers fit a CTA's shared memory with CW >= 4:
//   [2 * n * CW | row records 2n | dis n | pos n1 | cols m/2]   words
// CTA sizes ("classes"): 256 threads / 54 KB (4 per SM), 512 / 110 KB (2 per SM), 1024 / 222 KB.
__host__ __device__ inline int64_t chain_words(int64_t n, int64_t m, int64_t n1, int cw) {
    return 2 * n * cw + 2 * n + n + n1 + (m + 2) / 2 + 8;
}
__host__ __device__ inline int chain_class_bytes(int cls) { return cls == 0 ? 54 * 1024 : (cls == 1 ? 110 * 1024 : 222 * 1024); }
__host__ __device__ inline bool chain_fits(int64_t n, int64_t m, int64_t n1, int cw, int cls) {
    return 4 * chain_words(n, m, n1, cw) <= chain_class_bytes(cls);
}
__host__ __device__ inline bool chain_eligible(int flags, int strategy, int64_t n, int64_t m, int64_t n1) {
    return (flags & S3_BATCH_CCN_CHAIN) && strategy == S3_STRATEGY_UNION && n < 65536 && chain_fits(n, m, n1, 4, 2);
}
// Records too large for that placement ("spill" class 3: n up to ~9 000) keep the CSR in shared memory and the two
// [n][32] buffers in the float scratch the front kernel reserved for their CCN work items — global memory, read
// through L2 — provided that scratch is large enough (two or more work items' worth).
__host__ __device__ inline bool chain_spill_eligible(int flags, int strategy, int64_t n, int64_t m, int64_t n1, int s, int K) {
    if (!(flags & S3_BATCH_CCN_CHAIN) || strategy != S3_STRATEGY_UNION || n >= 65536 || chain_fits(n, m, n1, 4, 2)) return false;
    if (4 * (3 * n + n1 + (m + 2) / 2 + 8) > chain_class_bytes(2)) return false;
    return (int64_t)ccn_items(s, 2, 8) * ccn_item_words(K, n, 8) >= 64 * n + 8;
}
// (CW, class) of an eligible record as CW | class << 8. policy 0: the widest sub-chunk first (128-byte row segments
// are free of bank conflicts), in the smallest CTA that holds it; 1: two CTAs per SM at CW = 16 before one at CW = 32;
// 2: occupancy first.
__host__ __device__ inline int chain_shape(int64_t n, int64_t m, int64_t n1, int policy) {
#define S3_CHAIN_TRY(cw, cls) \
    if (chain_fits(n, m, n1, cw, cls)) return (cw) | ((cls) << 8)
    S3_CHAIN_TRY(32, 0);
    if (policy == 2) S3_CHAIN_TRY(16, 0);
    S3_CHAIN_TRY(32, 1);
    if (policy != 0) S3_CHAIN_TRY(16, 1);
    if (policy == 2) S3_CHAIN_TRY(8, 1);
    if (policy != 1) S3_CHAIN_TRY(32, 2);
    S3_CHAIN_TRY(16, 2);
    S3_CHAIN_TRY(8, 2);
#undef S3_CHAIN_TRY
    return 4 | (2 << 8);
}

// ---- deterministic per-hop cap (s3_batch.ratio_per_hop / max_nodes_per_hop / cap_seed; reference utils.py:66-70) ----
__host__ __device__ inline uint32_t fmix32(uint32_t h) {  // murmur3 finaliser: a bijection of the 32-bit ids
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
// nodes a hop keeps out of `count` new ones: min(int(ratio * count), max) as the reference computes it
__host__ __device__ inline int cap_keep(int count, double ratio, int max_nodes) {
    int k = count;
    if (ratio > 0.0 && ratio < 1.0) k = (int)(ratio * (double)count);
    if (max_nodes > 0 && max_nodes < k) k = max_nodes;
    return k;
}
inline bool batch_caps(const s3_batch& b) {
    return b.flow == S3_FLOW_POS && !b.walk_sets && ((b.ratio_per_hop > 0.0 && b.ratio_per_hop < 1.0) || b.max_nodes_per_hop > 0);
}
// One pass of the 4 x 8-bit radix select over 32-bit rank keys, run by the whole CTA after its threads have filled
// hist[256] with the counts of the keys that match `prefix` on the higher bytes: thread 0 finds the byte of the
// krem-th smallest key. sel[0] = prefix (in/out), sel[1] = krem (in/out). Contains two __syncthreads().
__device__ __forceinline__ void radix_pick(int* hist, int* sel) {
    __syncthreads();
    if (threadIdx.x == 0) {
        int krem = sel[1], cum = 0, b = 0;
        for (; b < 255; ++b) {
            if (cum + hist[b] >= krem) break;
            cum += hist[b];
        }
        sel[0] = (int)(((uint32_t)sel[0] << 8) | (uint32_t)b);
        sel[1] = krem - cum;
    }
    __syncthreads();
}

// link pairing (pair.cu) applies to fixed-row PoS batches whose rows are not all stored (no parity dump)
inline bool batch_pairing(const s3_batch& b) {
    return b.mirror && b.flow == S3_FLOW_POS && b.strategy == S3_STRATEGY_NONE && !(b.flags & S3_BATCH_STORE_ALL_ROWS) &&
           !b.walk_sets;
}

struct OutPtrs {
    float* p[2 * S3_MAX_K];  // K+1 operators; 2K for the hybrid flow (reference utils.py:454-480)
};

// Launch configuration that depends on the device: the opt-in to > 48 KB of dynamic shared memory
// (cudaFuncSetAttribute is per context) and the SM count / occupancy of a persistent kernel. Cached per
// (device, shared-memory size) because the queries are slow host calls; safe for several devices and
// threads in one process.
struct LaunchCache {
    static constexpr int kDevices = 64;
    std::mutex mu;
    size_t configured[kDevices] = {};
    size_t occ_smem[kDevices] = {};
    int sms[kDevices] = {};
    int occ[kDevices] = {};
    bool have[kDevices] = {};
    // func: the kernel; threads / smem: its launch shape. sms_out / occ_out may be null (attribute only).
    cudaError_t get(const void* func, int threads, size_t smem, int* sms_out, int* occ_out) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= kDevices) return cudaErrorInvalidDevice;
        std::lock_guard<std::mutex> lock(mu);
        if (smem > configured[dev] || (smem > 0 && configured[dev] == 0)) {
            e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured[dev] = smem;
        }
        if (sms_out || occ_out) {
            if (!have[dev] || occ_smem[dev] != smem) {
                e = cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
                if (e != cudaSuccess) return e;
                e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[dev], func, threads, smem);
                if (e != cudaSuccess) return e;
                have[dev] = true;
                occ_smem[dev] = smem;
            }
            if (sms_out) *sms_out = sms[dev];
            if (occ_out) *occ_out = occ[dev];
        }
        return cudaSuccess;
    }
};

struct PeerDst {  // s3_gather_peers destinations
    float* base[S3_MAX_PEERS];
    int num_dst;
    int64_t op_stride;
    int skip_op0;
    int skip_chain;
};

// launchers (defined in the .cu files, called from c_abi.cu)
cudaError_t launch_extract_bitmap(const s3_graph& g, const s3_batch& b, cudaStream_t st, int* rc_out);
cudaError_t launch_extract_sorted(const s3_graph& g, const s3_batch& b, cudaStream_t st, int* rc_out);
cudaError_t launch_order(const s3_batch& b, int64_t num_records, cudaStream_t st);
cudaError_t launch_build_hub_bits(const s3_graph& g, cudaStream_t st);
cudaError_t launch_node_proxy(const s3_graph& g, int32_t* out, cudaStream_t st);
cudaError_t launch_walk_sets(const s3_graph& g, const int64_t* starts, int64_t num_starts, int rw_m, int rw_M, uint64_t seed,
                             int cap, int32_t* sets, int32_t* counts, cudaStream_t st);
cudaError_t launch_plan(const s3_batch& b, cudaStream_t st);
cudaError_t launch_plan_items(const s3_batch& b, cudaStream_t st);
cudaError_t launch_diffuse(const s3_graph& g, const s3_batch& b, int64_t num_items, cudaStream_t st);
cudaError_t launch_gather(const s3_graph& g, const s3_batch& b, int64_t num_items, const OutPtrs& out,
                          int64_t ldo, int64_t row_base, bool ccn, cudaStream_t st, const PeerDst* peers = nullptr);
cudaError_t launch_fill_mirrors(const int64_t* mirror, int64_t num_links, const OutPtrs& ops, int first_op, int num_ops,
                                int64_t cols, int64_t ldo, cudaStream_t st);
cudaError_t launch_fill_x0(const s3_graph& g, const int64_t* src, const int64_t* dst, int64_t num_links, float* out, int64_t ldo,
                           cudaStream_t st);
cudaError_t launch_pair_links(const int64_t* src, const int64_t* dst, int64_t L, int64_t N, int64_t* table, int64_t slots,
                              int64_t* mirror, cudaStream_t st);
cudaError_t launch_pair_heads(const int64_t* mirror, int64_t L, int64_t* head_code, cudaStream_t st);
cudaError_t launch_scatter_rows(const OutPtrs& src, int64_t ld_src, const int64_t* src_row_ptr, int64_t num_records,
                                const int64_t* link_idx, int64_t link_base, const int64_t* mirror, const int64_t* dst_row_ptr,
                                const OutPtrs& dst, int64_t ld_dst, int num_ops, int64_t cols, cudaStream_t st, int lead = 0);
cudaError_t launch_segment_pool(const float* src, int64_t ld, int64_t cols, const int64_t* row_ptr, int64_t num_links, int mode,
                                int layout, float* out, int64_t ld_out, cudaStream_t st);
cudaError_t launch_probe_l2_read(const float* buf, int64_t bytes, int iters, float* sink, int ctas, cudaStream_t st);
cudaError_t launch_probe_fma(int iters, float* sink, int ctas, cudaStream_t st);
cudaError_t launch_probe_fma2(int iters, float* sink, int ctas, cudaStream_t st);
cudaError_t launch_negative_candidates(const s3_graph& g, int64_t M, uint64_t seed, int64_t* table, int64_t slots, int64_t* src,
                                       int64_t* dst, uint8_t* valid, cudaStream_t st);
cudaError_t peer_alloc(int64_t bytes, void** ptr);
cudaError_t peer_free(void* ptr);
cudaError_t peer_export(void* ptr, unsigned char* handle);
cudaError_t peer_open(const unsigned char* handle, void** ptr);
cudaError_t peer_close(void* ptr);
cudaError_t launch_ccn_chain(const s3_graph& g, const s3_batch& b, int64_t num_records, const OutPtrs& out, int64_t ldo,
                             int64_t row_base, cudaStream_t st, float* pool = nullptr, int* pool_busy = nullptr,
                             int64_t slot_floats = 0, int pool_slots = 0, int pool_cw = 0);
cudaError_t launch_plan_full(const s3_batch& b, cudaStream_t st);
cudaError_t launch_sign_full(const s3_graph& g, const s3_batch& b, int64_t num_records, int label, const OutPtrs& out,
                             int64_t ldo, int64_t row_base, int64_t* node_out, cudaStream_t st);
cudaError_t launch_joint_rows(const OutPtrs& src, int num_ops, int64_t cols, int64_t ld_src, const int64_t* row_ptr,
                              const int64_t* link_idx, int64_t num_links, const int64_t* out_row_ptr, int rows_per_link,
                              float* dst, int64_t ld_dst, int64_t* batch_vec, cudaStream_t st);
cudaError_t launch_sign_head(const float* x, int64_t rows, int64_t kdim, int64_t ldx, const float* w, int64_t ldw,
                             const float* bias, const float* scale, const float* shift, float* pooled, int pool, cudaStream_t st);
cudaError_t launch_dump_edges(const s3_batch& b, const int64_t* edge_ptr, int32_t* edges_out, cudaStream_t st);

}  // namespace s3
