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

// ---------------------------------------------------------------------------------------------------------
// Negative sampling on the GPU (the input side of the path): replaces the
// torch_geometric.utils.negative_sampling call of reference utils.py:645-648 / the NumPy rejection loop of the
// host mirror. Candidate i is the ordered pair (u, v) = counter-based hash of (seed, i): it is kept when u != v,
// (u, v) is not a stored entry of the CSR (binary search in row u) and no earlier candidate is the same pair
// (open-addressing table keyed by u * N + v, value = lowest candidate index). A candidate's fate depends on
// (seed, i) and on the candidates before it only, so the kept list — the first `count` kept candidates in index
// order — is the same on every run and on every GPU.
// ---------------------------------------------------------------------------------------------------------
namespace s3 {
namespace {

__device__ __forceinline__ void neg_candidate(uint64_t seed, int64_t i, int64_t N, int64_t* u, int64_t* v) {
    const uint64_t r = mix64(seed ^ mix64((uint64_t)i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull));
    *u = (int64_t)__umul64hi(r, (uint64_t)N);
    *v = (int64_t)__umul64hi(mix64(r ^ 0xD6E8FEB86659FD93ull), (uint64_t)N);
}

__device__ __forceinline__ bool is_edge(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t u, int64_t v) {
    int64_t lo = indptr[u], hi = indptr[u + 1];
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (indices[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo < indptr[u + 1] && indices[lo] == (int32_t)v;
}

__global__ void neg_insert_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t N, int64_t M,
                                  uint64_t seed, unsigned long long* keys, unsigned long long* first, uint64_t mask,
                                  int64_t* src, int64_t* dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    int64_t u, v;
    neg_candidate(seed, i, N, &u, &v);
    src[i] = u;
    dst[i] = v;
    if (u == v || is_edge(indptr, indices, u, v)) return;
    const unsigned long long key = (unsigned long long)(u * N + v);
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

__global__ void neg_valid_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t N, int64_t M,
                                 const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ first,
                                 uint64_t mask, const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                 uint8_t* __restrict__ valid) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int64_t u = src[i], v = dst[i];
    uint8_t ok = 0;
    if (u != v && !is_edge(indptr, indices, u, v)) {
        const unsigned long long key = (unsigned long long)(u * N + v);
        uint64_t h = mix64(key) & mask;
        while (keys[h] != key) h = (h + 1) & mask;
        ok = first[h] == (unsigned long long)i ? 1 : 0;
    }
    valid[i] = ok;
}

}  // namespace

cudaError_t launch_negative_candidates(const s3_graph& g, int64_t M, uint64_t seed, int64_t* table, int64_t slots, int64_t* src,
                                       int64_t* dst, uint8_t* valid, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(table, 0xff, (size_t)slots * 16, st);
    if (e != cudaSuccess || M == 0) return e;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(table);
    unsigned long long* first = keys + slots;
    const unsigned grid = (unsigned)((M + 255) / 256);
    neg_insert_kernel<<<grid, 256, 0, st>>>(g.indptr, g.indices, g.num_nodes, M, seed, keys, first, (uint64_t)slots - 1, src, dst);
    neg_valid_kernel<<<grid, 256, 0, st>>>(g.indptr, g.indices, g.num_nodes, M, keys, first, (uint64_t)slots - 1, src, dst, valid);
    return cudaGetLastError();
}

}  // namespace s3
