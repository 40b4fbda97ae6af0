// Link pairing: one record per unordered node pair.
//
// The reference precomputes every training positive twice, as (u,v) and as (v,u)
// (sgrl_link_pred.py:193-204 over PyG's train_test_split_edges, which stores both directions of
// every training edge: SURVEY.md A.7). The enclosing subgraph of (v,u) is the subgraph of (u,v)
// with local rows 0 and 1 exchanged (utils.py:53-80 is symmetric in src / dst), so the second one
// is a row swap of the first. s3_pair_links finds, for every unordered pair that occurs more
// than once in the link list of a call, the link with the lowest index (it keeps the work) and
// chains the others behind it; s3_extract skips chain members and s3_gather writes their rows.
//
// Open-addressing hash table keyed by min(u,v) * N + max(u,v); value = lowest link index
// (atomicMin). No sort, no host round trip; the chain order is scheduling dependent, the rows
// written are not.
#include "common.cuh"

namespace s3 {
namespace {

constexpr unsigned long long kEmpty = ~0ull;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

__device__ __forceinline__ bool pair_key(const int64_t* src, const int64_t* dst, int64_t i, int64_t N, unsigned long long* key) {
    const int64_t u = src[i], v = dst[i];
    if (u < 0 || v < 0 || u >= N || v >= N || u == v) return false;  // left to s3_extract (S3_REC_BAD_LINK)
    const int64_t lo = u < v ? u : v, hi = u < v ? v : u;
    *key = (unsigned long long)(lo * N + hi);
    return true;
}

__global__ void pair_insert_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t L, int64_t N,
                                   unsigned long long* keys, unsigned long long* first, uint64_t mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    unsigned long long key;
    if (!pair_key(src, dst, i, N, &key)) return;
    uint64_t h = mix64(key) & mask;
    for (;;) {
        const unsigned long long prev = atomicCAS(&keys[h], kEmpty, key);
        if (prev == kEmpty || prev == key) {
            atomicMin(&first[h], (unsigned long long)i);
            return;
        }
        h = (h + 1) & mask;
    }
}

__global__ void pair_chain_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t L, int64_t N,
                                  const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ first,
                                  uint64_t mask, long long* mirror) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    unsigned long long key;
    if (!pair_key(src, dst, i, N, &key)) return;
    uint64_t h = mix64(key) & mask;
    while (keys[h] != key) h = (h + 1) & mask;
    const int64_t p = (int64_t)first[h];
    if (p == i) return;  // keeps the work; its slot is the chain head, touched by atomicExch only
    const long long swap = src[i] != src[p] ? 1 : 0;
    const long long next = (long long)atomicExch(reinterpret_cast<unsigned long long*>(&mirror[p]), (unsigned long long)i);
    mirror[i] = -2 - (((next + 1) << 1) | swap);
}

}  // namespace

cudaError_t launch_pair_links(const int64_t* src, const int64_t* dst, int64_t L, int64_t N, int64_t* table, int64_t slots,
                              int64_t* mirror, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(table, 0xff, (size_t)slots * 16, st);  // keys = empty, first = +inf
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(mirror, 0xff, (size_t)L * 8, st);  // -1: no chain
    if (e != cudaSuccess || L == 0) return e;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(table);
    unsigned long long* first = keys + slots;
    const unsigned grid = (unsigned)((L + 255) / 256);
    pair_insert_kernel<<<grid, 256, 0, st>>>(src, dst, L, N, keys, first, (uint64_t)slots - 1);
    pair_chain_kernel<<<grid, 256, 0, st>>>(src, dst, L, N, keys, first, (uint64_t)slots - 1,
                                            reinterpret_cast<long long*>(mirror));
    return cudaGetLastError();
}

}  // namespace s3
