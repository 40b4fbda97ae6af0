// ScaLed random-walk node sets (first "next" row).
//
// Replaces reference utils.py:425-443 (create_rw_cache): for every start node, M uniform random
// walks of length m (torch_cluster.random_walk with p = q = 1: each step moves to a uniformly
// chosen neighbour, a node without neighbours stays put) and the sorted set of the nodes they
// visit — what the reference keeps per node as `torch.unique(cat(walks))`.  The enclosing subgraph
// of link (u, v) is then the union of the sets of u and v (utils.py:102-105), which is exactly
// what the sorted-set front kernel (extract_sorted.cu) merges.
//
// One CTA per start node. Walks are counter-based: step t of walk w of start node s draws
// hash(seed, s, w, t), so the sets do not depend on scheduling or on which other nodes are in
// the call (the reference's RNG stream cannot be reproduced; parity is distributional, while
// everything downstream of the sets is exact). The visited nodes are sorted by a bitonic
// network in shared memory and de-duplicated with a block scan.
#include <climits>

#include "common.cuh"

namespace s3 {
namespace {

constexpr int kWalkThreads = 128;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(kWalkThreads) walk_sets_kernel(const int64_t* __restrict__ indptr,
                                                                 const int32_t* __restrict__ indices, int64_t num_nodes,
                                                                 const int64_t* __restrict__ starts, int rw_m, int rw_M,
                                                                 uint64_t seed, int cap, int P, int32_t* __restrict__ sets,
                                                                 int32_t* __restrict__ counts) {
    extern __shared__ int s_v[];  // [P] visited nodes, padded with INT_MAX to a power of two
    __shared__ int s_scan[33];
    const int tid = threadIdx.x, T = kWalkThreads;
    const int64_t si = blockIdx.x;
    const int64_t s = starts[si];
    if (s < 0 || s >= num_nodes) {  // invalid start: empty set (the front kernel flags the link)
        if (tid == 0) counts[si] = 0;
        return;
    }
    for (int i = tid; i < P; i += T) s_v[i] = INT_MAX;
    __syncthreads();
    for (int w = tid; w < rw_M; w += T) {
        int cur = (int)s;
        s_v[w * (rw_m + 1)] = cur;
        for (int t = 1; t <= rw_m; ++t) {
            const int64_t e0 = indptr[cur];
            const int d = (int)(indptr[cur + 1] - e0);
            if (d > 0) {
                const uint64_t r = mix64(mix64(seed ^ (uint64_t)s * 0xD1B54A32D192ED03ull) + ((uint64_t)w << 20) + (uint64_t)t);
                cur = indices[e0 + (int64_t)__umul64hi(r, (uint64_t)d)];  // floor(r / 2^64 * d): uniform in [0, d) for any degree
            }
            s_v[w * (rw_m + 1) + t] = cur;
        }
    }
    __syncthreads();
    // bitonic sort of P values
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += T) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const int a = s_v[i], b = s_v[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        s_v[i] = b;
                        s_v[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    // unique: keep v[i] != v[i-1], compact in order
    int running = 0;
    for (int base = 0; base < P; base += T) {
        const int i = base + tid;
        int keep = 0, val = INT_MAX;
        if (i < P) {
            val = s_v[i];
            keep = (val != INT_MAX && (i == 0 || s_v[i - 1] != val)) ? 1 : 0;
        }
        int tot;
        const int ex = block_exclusive_scan(keep, s_scan, &tot);
        if (keep && running + ex < cap) sets[si * cap + running + ex] = val;
        running += tot;
        __syncthreads();
    }
    if (tid == 0) counts[si] = min(running, cap);
}

}  // namespace

cudaError_t launch_walk_sets(const s3_graph& g, const int64_t* starts, int64_t num_starts, int rw_m, int rw_M, uint64_t seed,
                             int cap, int32_t* sets, int32_t* counts, cudaStream_t st) {
    if (num_starts == 0) return cudaSuccess;
    int P = 1;
    while (P < rw_M * (rw_m + 1)) P <<= 1;
    walk_sets_kernel<<<(unsigned)num_starts, kWalkThreads, (size_t)P * sizeof(int), st>>>(
        g.indptr, g.indices, g.num_nodes, starts, rw_m, rw_M, seed, cap, P, sets, counts);
    return cudaGetLastError();
}

}  // namespace s3
