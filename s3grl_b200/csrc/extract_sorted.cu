// Kernel 1, sorted-set tier — the front kernel for graphs too large for shared-memory bitmaps
// (N up to 2^31: the 10 M-node R-MAT configuration), one-hop enclosing subgraphs.
//
// Replaces the same reference lines as extract.cu (utils.py:33-85, tuned_SIGN.py:153-175 for the
// rows of the two targets). With num_hops = 1 the subgraph of link (u, v) is
//      S = {u, v} ∪ N(u) ∪ N(v)
// and both adjacency lists are already sorted, so the canonical order [u, v, ascending id] comes
// out of a parallel MERGE — exclusive prefix sums of "keep" flags over N(u) and N(v) plus one
// binary search per element give every node its final position; no hash table, no sort.
// Membership / local-id lookups afterwards are binary searches in that sorted node list
// (shared memory when it fits).
//
// The induced adjacency N(g_j) ∩ S of row j is computed by whichever side is cheaper:
//   (i)  stream N(g_j) and look every entry up in S          deg_j · log n
//   (ii) look every node of S up in N(g_j)                    n · log deg_j   (hub rows)
// one warp per row, exact count -> block scan -> fill (ballot-ordered), so the local CSR is
// dense and the arena only holds the m real edges. All K sweeps then run over that CSR
// (num_hops = 1 < K: every row is read by every sweep), 8-lane group per row, z in shared
// memory when the subgraph fits.
//
// Larger radii on graphs of this size are rejected (S3_ERR_UNSUPPORTED): a 2-hop ball of an
// R-MAT hub is most of the graph, and the reference itself only handles it with random caps.
#include <climits>

#include "common.cuh"

namespace s3 {
namespace {

struct SortedParams {
    const int64_t* __restrict__ indptr;
    const int32_t* __restrict__ indices;
    int64_t num_nodes;
    const int32_t* __restrict__ hub_id;     // optional hub index (s3_graph): null when absent
    const uint32_t* __restrict__ hub_bits;
    int hub_words;                          // words per row of the hub bit matrix
    const int64_t* __restrict__ link_src;
    const int64_t* __restrict__ link_dst;
    int64_t num_records;
    int sign_k, store_all, flags;
    int strategy;  // PoS Plus row selection (S3_STRATEGY_*): needs store_all
    // ScaLed: node lists come from per-node random-walk sets instead of the adjacency lists
    const int32_t* __restrict__ walk_sets;    // [num_sets, walk_cap] ascending unique, or null
    const int32_t* __restrict__ walk_counts;  // [num_sets]
    const int64_t* __restrict__ src_set;      // [num_records] row of walk_sets for the source / destination
    const int64_t* __restrict__ dst_set;
    int walk_cap;
    int caps;  // per-hop cap of the one hop (s3_batch.ratio_per_hop / max_nodes_per_hop / cap_seed)
    double cap_ratio;
    int cap_max;
    uint32_t cap_seed;
    int32_t* arena;
    int64_t arena_words;
    int64_t slab_stride;  // words per CTA slab: prefix arrays of both adjacency lists
    int64_t slab_words;
    int64_t* off;
    int32_t* cnt;
    unsigned long long* counters;
};

constexpr int kNodeCap = 2048;  // nodes cached in shared memory (8 KB)
constexpr int kZCapS = 2048;    // floats per shared z buffer (2 buffers; also scratch for the big-node list)
constexpr int kBitCap = 2048;   // subgraphs up to this many nodes use the n x n bit-matrix path
constexpr int kMCapS = 3072;    // bit-matrix words kept in shared memory (n <= 313), else in the arena

// number of elements of the ascending array a[0..n) that are < x
__device__ __forceinline__ int lower_bound(const int32_t* __restrict__ a, int n, int x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ bool contains(const int32_t* __restrict__ a, int n, int x) {
    const int p = lower_bound(a, n, x);
    return p < n && a[p] == x;
}

// local id of global node c in [u, v, sorted rest...] or -1
__device__ __forceinline__ int lookup(const int32_t* nodes, int n, int u, int v, int c) {
    if (c == u) return 0;
    if (c == v) return 1;
    const int p = lower_bound(nodes + 2, n - 2, c);
    return (p < n - 2 && nodes[2 + p] == c) ? 2 + p : -1;
}


// ---- PoS Plus on this tier (reference tuned_SIGN.py:228-238) + the float scratch of the record's work items ----
// With num_hops = 1 every non-seed node of the subgraph is a neighbour of u or of v, and the target-link mask only
// touches the seed-seed entry, so the selected rows beyond [0, 1] are: intersection — the nodes that sit in BOTH
// adjacency lists A = N(u) and B = N(v); union — every node c >= 2. They are written in ascending local id behind
// `sel`, then ONE allocation takes [sel | float scratch of item 0 | scratch of the CCN work items] (the CCN items must
// follow item 0 contiguously: diffuse.cu / gather_kernel.cuh address them from OFF_F32). Returns false on overflow.
__device__ __forceinline__ bool select_and_allocate(const SortedParams& p, const int32_t* nodes, int n, const int32_t* __restrict__ A,
                                                    int du, const int32_t* __restrict__ B, int dv, int* s_scan, long long* s_base,
                                                    int& s_out, int64_t& base_sel, int64_t& base3, int m) {
    const int T = kExtractThreads, tid = threadIdx.x, K = p.sign_k;
    int extra = 0;
    if (p.strategy == S3_STRATEGY_UNION) {
        extra = n - 2;
    } else if (p.strategy == S3_STRATEGY_INTERSECTION) {
        for (int base = 2; base < n; base += T) {
            const int c = base + tid;
            const int f = (c < n && contains(A, du, nodes[c]) && contains(B, dv, nodes[c])) ? 1 : 0;
            int tot;
            block_exclusive_scan(f, s_scan, &tot);
            extra += tot;
            __syncthreads();
        }
    }
    const int s = 2 + extra, cr = ccn_rows(p.strategy);
    const int64_t wsel = ((int64_t)extra + 31) & ~int64_t(31);
    // records whose CCN rows go through s3_ccn_chain need no work-item scratch for them (one hop: every node is hop <= 1)
    const int64_t items = chain_eligible(p.flags, p.strategy, n, m, n) ? 0 : ccn_items(s, 2, cr);
    const int64_t wf = (item_words(S3_FLOW_POS, K, n) + items * ccn_item_words(K, n, cr) + 31) & ~int64_t(31);
    __syncthreads();
    if (tid == 0) *s_base = (long long)atomicAdd(&p.counters[S3_CTR_CURSOR], (unsigned long long)(wsel + wf));
    __syncthreads();
    base_sel = p.slab_words + *s_base;
    base3 = base_sel + wsel;
    s_out = s;
    if (base3 + wf > p.arena_words) return false;
    int32_t* sel = p.arena + base_sel;
    if (p.strategy == S3_STRATEGY_UNION) {
        for (int c = 2 + tid; c < n; c += T) sel[c - 2] = c;
    } else if (p.strategy == S3_STRATEGY_INTERSECTION) {
        int run = 0;
        for (int base = 2; base < n; base += T) {
            const int c = base + tid;
            const int f = (c < n && contains(A, du, nodes[c]) && contains(B, dv, nodes[c])) ? 1 : 0;
            int tot;
            const int ex = block_exclusive_scan(f, s_scan, &tot);
            if (f) sel[run + ex] = c;
            run += tot;
            __syncthreads();
        }
    }
    __syncthreads();
    return true;
}

// ---- bit-matrix path (n <= kBitCap) ---------------------------------------------------------
// The induced adjacency is an n x n symmetric bit matrix M. Every unordered pair is decided ONCE
// and both bits are set:
//   * a "small" node j (deg_G <= 2n) streams its adjacency list and looks every entry up in S;
//   * a pair of "big" nodes is decided by one binary search of the one in the adjacency list of
//     the other (the shorter list) — flat over all threads of the CTA, no per-row imbalance.
// The K sweeps then run straight over the bit rows (ascending local id: the canonical order), so
// no CSR is materialised unless a dump asks for it.
struct BitCtx {
    const int32_t* nodes;  // shared
    int n, u, v;
    int* s_deg;            // [kBitCap] shared
    uint32_t* s_M;         // [kMCapS] shared
    float* s_z;            // [2 * kZCapS] shared
    int* s_scan;
    long long* s_base;
    const int32_t* A;  // the two merged lists (row selection of PoS Plus)
    const int32_t* B;
    int du, dv;
};

__device__ __forceinline__ bool bitmatrix_record(const SortedParams& p, const BitCtx& cx, int32_t* rowptr, int32_t* rowlen,
                                                 int& m_out, int64_t& base2, int64_t& base3, unsigned long long& my_deg,
                                                 unsigned long long& my_read, int& s_out, int64_t& base_sel) {
    constexpr int SC = 2;
    const int T = kExtractThreads, tid = threadIdx.x, K = p.sign_k;
    const int lane = tid & 31, wid = tid >> 5, l8 = tid & 7, grp = tid >> 3;
    constexpr int NG = kExtractThreads / 8, NWARP = kExtractThreads / 32;
    const int NW = (K + 1) * SC, NWP = (NW + 3) & ~3;
    const int n = cx.n, u = cx.u, v = cx.v;
    const int32_t* nodes = cx.nodes;
    const int nw = (n + 31) >> 5;
    const int mwords = n * nw;
    const bool m_shared = mwords <= kMCapS;

    // allocation: the bit matrix, when it does not fit in shared memory (the float scratch follows the row selection)
    const int64_t wordsM = m_shared ? 0 : (((int64_t)mwords + 31) & ~int64_t(31));
    __syncthreads();
    if (tid == 0) *cx.s_base = (long long)atomicAdd(&p.counters[S3_CTR_CURSOR], (unsigned long long)wordsM);
    __syncthreads();
    const int64_t baseM = p.slab_words + *cx.s_base;
    base3 = baseM + wordsM;
    if (baseM + wordsM > p.arena_words) return true;  // overflow
    uint32_t* M = m_shared ? cx.s_M : reinterpret_cast<uint32_t*>(p.arena + baseM);
    for (int i = tid; i < mwords; i += T) M[i] = 0u;

    // global degrees; list of big nodes (scratch: the z buffers, not live yet)
    int* s_big = reinterpret_cast<int*>(cx.s_z);
    const int tau = 2 * n;
    int nb = 0;
    for (int base = 0; base < n; base += T) {
        const int j = base + tid;
        int d = 0, big = 0;
        if (j < n) {
            const int g = nodes[j];
            d = (int)(p.indptr[g + 1] - p.indptr[g]);
            cx.s_deg[j] = d;
            my_deg += (unsigned long long)d;
            big = d > tau ? 1 : 0;
            if (!big) my_read += (unsigned long long)d;  // adjacency entries this method really streams
        }
        int tot;
        const int ex = block_exclusive_scan(big, cx.s_scan, &tot);
        if (big) s_big[nb + ex] = j;
        nb += tot;
        __syncthreads();
    }

    // phase A: small rows stream their adjacency list
    for (int j = wid; j < n; j += NWARP) {
        const int d = cx.s_deg[j];
        if (d > tau) continue;
        const int32_t* __restrict__ Nj = p.indices + p.indptr[nodes[j]];
        for (int e = lane; e < d; e += 32) {
            const int lid = lookup(nodes, n, u, v, Nj[e]);
            if (lid >= 0) {
                atomicOr(&M[j * nw + (lid >> 5)], 1u << (lid & 31));
                atomicOr(&M[lid * nw + (j >> 5)], 1u << (j & 31));
            }
        }
    }
    // phase B: big-big pairs, one binary search each, in the shorter of the two lists
    for (int q = tid; q < nb * nb; q += T) {
        const int ia = q / nb, ib = q - ia * nb;
        if (ia >= ib) continue;
        int ja = s_big[ia], jb = s_big[ib];
        if (cx.s_deg[ja] > cx.s_deg[jb]) {
            const int t = ja;
            ja = jb;
            jb = t;
        }
        const int ga = nodes[ja], gb = nodes[jb];
        bool adj;
        int ha = -1, hb = -1;
        if (p.hub_id) {
            ha = p.hub_id[ga];
            hb = p.hub_id[gb];
        }
        if (ha >= 0 && hb >= 0)  // both among the graph's hubs: one bit probe
            adj = (p.hub_bits[(size_t)ha * p.hub_words + (hb >> 5)] >> (hb & 31)) & 1u;
        else                      // search g_b in N(g_a), the shorter list
            adj = contains(p.indices + p.indptr[ga], cx.s_deg[ja], gb);
        if (adj) {
            atomicOr(&M[ja * nw + (jb >> 5)], 1u << (jb & 31));
            atomicOr(&M[jb * nw + (ja >> 5)], 1u << (ja & 31));
        }
    }
    __syncthreads();
    if (tid == 0) {  // mask the target link (utils.py:79-80)
        M[0 * nw + 0] &= ~2u;
        M[1 * nw + 0] &= ~1u;
    }
    __syncthreads();
    // induced degrees
    int my_m = 0;
    for (int j = tid; j < n; j += T) {
        int c = 0;
        for (int w = 0; w < nw; ++w) c += __popc(M[j * nw + w]);
        cx.s_deg[j] = c;
        my_m += c;
    }
    __syncthreads();

    // optional CSR for dumps / CCN work items: rowlen, rowptr, lcol (ascending local id)
    base2 = baseM;
    m_out = 0;
    if (p.store_all) {
        int m = 0;
        for (int base = 0; base < n; base += T) {
            const int j = base + tid;
            const int d = j < n ? cx.s_deg[j] : 0;
            int tot;
            const int ex = block_exclusive_scan(d, cx.s_scan, &tot);
            if (j < n) {
                rowptr[j] = m + ex;
                rowlen[j] = d;
            }
            m += tot;
            __syncthreads();
        }
        if (tid == 0) rowptr[n] = m;
        m_out = m;
        const int64_t wordsC = ((int64_t)m + 31) & ~int64_t(31);
        if (tid == 0) *cx.s_base = (long long)atomicAdd(&p.counters[S3_CTR_CURSOR], (unsigned long long)wordsC);
        __syncthreads();
        base2 = p.slab_words + *cx.s_base;
        if (base2 + wordsC > p.arena_words) return true;
        int32_t* lcol = p.arena + base2;
        for (int j = tid; j < n; j += T) {  // one thread per row: dumps are not on the hot path
            int o = rowptr[j];
            for (int w = 0; w < nw; ++w) {
                uint32_t bits = M[j * nw + w];
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    lcol[o++] = w * 32 + b;
                }
            }
        }
        __syncthreads();
    } else {
        for (int d = 16; d > 0; d >>= 1) my_m += __shfl_down_sync(0xffffffffu, my_m, d);
        m_out = 0;  // not reduced across the block: CNT_M is exact only for dumps
    }

    // ---- selected rows (PoS Plus) and the float scratch of the work items ----
    if (!select_and_allocate(p, nodes, n, cx.A, cx.du, cx.B, cx.dv, cx.s_scan, cx.s_base, s_out, base_sel, base3, m_out)) return true;

    // ---- diffusion of the target rows over the bit rows ----
    float* item_f = reinterpret_cast<float*>(p.arena + base3);
    float* lab = item_f;
    float* wgt = item_f + NWP;
    const bool z_shared = n * SC <= kZCapS;
    float* zprev = z_shared ? cx.s_z : wgt + (int64_t)n * NWP;
    float* znext = z_shared ? cx.s_z + kZCapS : wgt + (int64_t)n * NWP + (int64_t)n * SC;
    __syncthreads();  // s_big (aliasing the z buffers) is dead
    for (int j = tid; j < n; j += T) {
#pragma unroll
        for (int c = 0; c < SC; ++c) {
            float zv = 0.0f;
            if (j == c) {
                const int deg = cx.s_deg[j];
                zv = deg > 0 ? 1.0f / sqrtf((float)deg) : 0.0f;
            }
            zprev[(int64_t)j * SC + c] = zv;
            znext[(int64_t)j * SC + c] = 0.0f;
        }
    }
    for (int i = tid; i < 2 * NWP; i += T) {
        const int j = i / NWP, q = i - j * NWP;
        wgt[i] = (q < SC && q == j) ? 1.0f : 0.0f;
    }
    __syncthreads();
    for (int k = 1; k <= K; ++k) {
        for (int jb0 = 0; jb0 < n; jb0 += NG) {
            const int j = jb0 + grp;
            const bool valid = j < n;
            float t[SC];
#pragma unroll
            for (int c = 0; c < SC; ++c) t[c] = 0.0f;
            if (valid) {
                for (int w = l8; w < nw; w += 8) {
                    uint32_t bits = M[j * nw + w];
                    while (bits) {
                        const int i = w * 32 + __ffs(bits) - 1;
                        bits &= bits - 1;
#pragma unroll
                        for (int c = 0; c < SC; ++c) t[c] += zprev[(int64_t)i * SC + c];
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < SC; ++c) {
                t[c] += __shfl_xor_sync(0xffffffffu, t[c], 4);
                t[c] += __shfl_xor_sync(0xffffffffu, t[c], 2);
                t[c] += __shfl_xor_sync(0xffffffffu, t[c], 1);
            }
            if (valid && l8 == 0) {
                const int deg = cx.s_deg[j];
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
    for (int q = tid; q < NWP; q += T) lab[q] = q < NW ? wgt[q] + wgt[NWP + q] : 0.0f;
    return false;
}

__global__ void __launch_bounds__(kExtractThreads, 4) front_sorted_kernel(SortedParams p) {
    const int T = kExtractThreads, tid = threadIdx.x, K = p.sign_k;
    constexpr int SC = 2;
    const int lane = tid & 31, wid = tid >> 5, l8 = tid & 7, grp = tid >> 3;
    constexpr int NG = kExtractThreads / 8, NWARP = kExtractThreads / 32;
    const int NW = (K + 1) * SC, NWP = (NW + 3) & ~3;
    __shared__ int s_scan[33];
    __shared__ long long s_base;
    __shared__ long long s_rec;
    __shared__ int s_nodes[kNodeCap];
    __shared__ float s_z[2][kZCapS];
    __shared__ int s_deg[kBitCap];
    __shared__ uint32_t s_M[kMCapS];

    int32_t* PA = p.arena + (int64_t)blockIdx.x * p.slab_stride;  // [du + 1]

    for (;;) {
        __syncthreads();
        if (tid == 0) s_rec = (long long)atomicAdd(&p.counters[S3_CTR_WORK], 1ull);
        __syncthreads();
        const int64_t rec = s_rec;
        if (rec >= p.num_records) break;
        const int64_t a64 = p.link_src[rec], b64 = p.link_dst[rec];
        int32_t* cnt = p.cnt + rec * S3_NCNT;
        int64_t* off = p.off + rec * S3_NOFF;
        if (a64 < 0 || b64 < 0 || a64 >= p.num_nodes || b64 >= p.num_nodes || a64 == b64) {
            if (tid == 0) {
                for (int i = 0; i < S3_NCNT; ++i) cnt[i] = 0;
                for (int i = 0; i < S3_NOFF; ++i) off[i] = 0;
                cnt[S3_CNT_STATUS] = S3_REC_BAD_LINK;
                cnt[S3_CNT_PARTNER] = -1;
                atomicAdd(&p.counters[S3_CTR_ERRORS], 1ull);
            }
            continue;
        }
        const int u = (int)a64, v = (int)b64;
        // the two sorted node lists whose union (plus u, v) is the subgraph: adjacency lists
        // (num_hops = 1) or random-walk sets (ScaLed, utils.py:102-105)
        const int32_t* __restrict__ A;
        const int32_t* __restrict__ B;
        int du, dv;
        if (p.walk_sets) {
            const int64_t ia = p.src_set[rec], ib = p.dst_set[rec];
            A = p.walk_sets + ia * p.walk_cap;
            B = p.walk_sets + ib * p.walk_cap;
            du = p.walk_counts[ia];
            dv = p.walk_counts[ib];
        } else {
            const int64_t eu0 = p.indptr[u], ev0 = p.indptr[v];
            A = p.indices + eu0;
            B = p.indices + ev0;
            du = (int)(p.indptr[u + 1] - eu0);
            dv = (int)(p.indptr[v + 1] - ev0);
        }
        int32_t* PB = PA + du + 1;  // [dv + 1]

        // ---- keep flags and their exclusive prefix sums: A minus {u,v}; B minus {u,v} minus A ----
        // (with a per-hop cap: a second round keeps only the nodes whose rank key is <= the selected threshold)
        int nA = 0, nB = 0;
        bool use_thr = false;
        uint32_t thr = 0u;
        for (int round = 0; round < 2; ++round) {
            nA = 0;
            nB = 0;
            for (int base = 0; base < du; base += T) {
                const int i = base + tid;
                int f = 0;
                if (i < du) {
                    const int x = A[i];
                    f = (x != u && x != v && (!use_thr || fmix32((uint32_t)x ^ p.cap_seed) <= thr)) ? 1 : 0;
                }
                int tot;
                const int ex = block_exclusive_scan(f, s_scan, &tot);
                if (i < du) PA[i] = nA + ex;
                nA += tot;
                __syncthreads();
            }
            for (int base = 0; base < dv; base += T) {
                const int k = base + tid;
                int f = 0;
                if (k < dv) {
                    const int y = B[k];
                    f = (y != u && y != v && (!use_thr || fmix32((uint32_t)y ^ p.cap_seed) <= thr) && !contains(A, du, y)) ? 1 : 0;
                }
                int tot;
                const int ex = block_exclusive_scan(f, s_scan, &tot);
                if (k < dv) PB[k] = nB + ex;
                nB += tot;
                __syncthreads();
            }
            if (tid == 0) {
                PA[du] = nA;
                PB[dv] = nB;
            }
            const int keep = (p.caps && !p.walk_sets && round == 0) ? cap_keep(nA + nB, p.cap_ratio, p.cap_max) : nA + nB;
            if (keep >= nA + nB) break;  // block-uniform: no cap, or nothing to drop
            // utils.py:66-70 with the deterministic rule: radix select of the keep-th smallest rank key among the
            // kept elements of both lists (scratch: the z buffers, not live yet)
            int* hist = reinterpret_cast<int*>(&s_z[0][0]);
            int* sel = hist + 256;
            __syncthreads();  // PA[du], PB[dv]
            if (tid == 0) {
                sel[0] = 0;
                sel[1] = keep;
            }
            for (int pass = 0; pass < 4; ++pass) {
                const int shift = 24 - 8 * pass;
                for (int i = tid; i < 256; i += T) hist[i] = 0;
                __syncthreads();
                const uint32_t prefix = (uint32_t)sel[0];
                for (int i = tid; i < du; i += T) {
                    if (PA[i + 1] == PA[i]) continue;
                    const uint32_t hsh = fmix32((uint32_t)A[i] ^ p.cap_seed);
                    if (pass == 0 || (hsh >> (shift + 8)) == prefix) atomicAdd(&hist[(hsh >> shift) & 255u], 1);
                }
                for (int k = tid; k < dv; k += T) {
                    if (PB[k + 1] == PB[k]) continue;
                    const uint32_t hsh = fmix32((uint32_t)B[k] ^ p.cap_seed);
                    if (pass == 0 || (hsh >> (shift + 8)) == prefix) atomicAdd(&hist[(hsh >> shift) & 255u], 1);
                }
                radix_pick(hist, sel);
            }
            thr = (uint32_t)sel[0];
            use_thr = true;
            __syncthreads();
        }
        const int n = 2 + nA + nB;

        // ---- allocation 1a: nodes | rowptr | rowlen ----
        const int64_t words1 = (3 * (int64_t)n + 1 + 31) & ~int64_t(31);
        __syncthreads();
        if (tid == 0) s_base = (long long)atomicAdd(&p.counters[S3_CTR_CURSOR], (unsigned long long)words1);
        __syncthreads();
        const int64_t base1 = p.slab_words + s_base;
        bool overflow = base1 + words1 > p.arena_words;
        int32_t* nodes_g = p.arena + base1;
        int32_t* rowptr = nodes_g + n;
        int32_t* rowlen = rowptr + n + 1;
        int64_t base2 = 0, base3 = 0, base_sel = 0;
        int m = 0, s_sel = 2;
        unsigned long long my_deg = 0;
        unsigned long long my_read = tid == 0 ? (unsigned long long)(du + dv) : 0ull;  // the two merged lists

        if (!overflow) {
            // ---- merge into the canonical order ----
            if (tid == 0) {
                nodes_g[0] = u;
                nodes_g[1] = v;
            }
            for (int i = tid; i < du; i += T) {
                if (PA[i + 1] != PA[i]) {  // kept (not a target, not dropped by the per-hop cap)
                    const int x = A[i];
                    nodes_g[2 + PA[i] + PB[lower_bound(B, dv, x)]] = x;
                }
            }
            for (int k = tid; k < dv; k += T) {
                if (PB[k + 1] != PB[k]) {  // kept
                    const int y = B[k];
                    nodes_g[2 + PB[k] + PA[lower_bound(A, du, y)]] = y;
                }
            }
            __syncthreads();
            const bool cached = n <= kNodeCap;
            if (cached)
                for (int j = tid; j < n; j += T) s_nodes[j] = nodes_g[j];
            __syncthreads();
            const int32_t* nodes = cached ? s_nodes : nodes_g;

            if (n <= kBitCap) {
                BitCtx cx;
                cx.nodes = s_nodes;
                cx.n = n;
                cx.u = u;
                cx.v = v;
                cx.s_deg = s_deg;
                cx.s_M = s_M;
                cx.s_z = &s_z[0][0];
                cx.s_scan = s_scan;
                cx.s_base = &s_base;
                cx.A = A;
                cx.B = B;
                cx.du = du;
                cx.dv = dv;
                overflow = bitmatrix_record(p, cx, rowptr, rowlen, m, base2, base3, my_deg, my_read, s_sel, base_sel);
            } else {
            // ======== large subgraphs: per-row intersections, exact count -> scan -> fill ========
            // ---- count pass: |N(g_j) ∩ S| minus the masked target link, one warp per row ----
            for (int j = wid; j < n; j += NWARP) {
                const int g = nodes[j];
                const int64_t e0 = p.indptr[g];
                const int d = (int)(p.indptr[g + 1] - e0);
                if (lane == 0) my_deg += (unsigned long long)d;
                const int32_t* __restrict__ Nj = p.indices + e0;
                const bool by_nodes = (int64_t)n * (32 - __clz(d | 1)) * 2 < d;  // (ii) for hub rows
                if (lane == 0 && !by_nodes) my_read += 2ull * (unsigned long long)d;  // count pass + fill pass
                int c_row = 0;
                if (!by_nodes) {
                    for (int e = lane; e < ((d + 31) & ~31); e += 32) {
                        int lid = -1;
                        if (e < d) lid = lookup(nodes, n, u, v, Nj[e]);
                        const bool ok = lid >= 0 && !((j == 0 && lid == 1) || (j == 1 && lid == 0));  // utils.py:79-80
                        c_row += __popc(__ballot_sync(0xffffffffu, ok));
                    }
                } else {
                    for (int t = lane; t < ((n + 31) & ~31); t += 32) {
                        bool ok = false;
                        if (t < n) ok = contains(Nj, d, nodes[t]) && !((j == 0 && t == 1) || (j == 1 && t == 0));
                        c_row += __popc(__ballot_sync(0xffffffffu, ok));
                    }
                }
                if (lane == 0) rowlen[j] = c_row;
            }
            __syncthreads();
            // ---- row starts: exclusive scan of rowlen ----
            for (int base = 0; base < n; base += T) {
                const int j = base + tid;
                const int d = j < n ? rowlen[j] : 0;
                int tot;
                const int ex = block_exclusive_scan(d, s_scan, &tot);
                if (j < n) rowptr[j] = m + ex;
                m += tot;
                __syncthreads();
            }
            if (tid == 0) rowptr[n] = m;

            // ---- allocation 1b: lcol[m] (the float scratch follows the row selection) ----
            const int64_t words2 = ((int64_t)m + 31) & ~int64_t(31);
            if (tid == 0) s_base = (long long)atomicAdd(&p.counters[S3_CTR_CURSOR], (unsigned long long)words2);
            __syncthreads();
            base2 = p.slab_words + s_base;
            overflow = base2 + words2 > p.arena_words;

            if (!overflow) {
                int32_t* lcol = p.arena + base2;
                // ---- fill pass: same intersections, ballot-ordered ----
                for (int j = wid; j < n; j += NWARP) {
                    const int g = nodes[j];
                    const int64_t e0 = p.indptr[g];
                    const int d = (int)(p.indptr[g + 1] - e0);
                    const int32_t* __restrict__ Nj = p.indices + e0;
                    const bool by_nodes = (int64_t)n * (32 - __clz(d | 1)) * 2 < d;
                    int outp = rowptr[j];
                    if (!by_nodes) {
                        for (int e = lane; e < ((d + 31) & ~31); e += 32) {
                            int lid = -1;
                            if (e < d) lid = lookup(nodes, n, u, v, Nj[e]);
                            const bool ok = lid >= 0 && !((j == 0 && lid == 1) || (j == 1 && lid == 0));
                            const unsigned ball = __ballot_sync(0xffffffffu, ok);
                            if (ok) lcol[outp + __popc(ball & ((1u << lane) - 1u))] = lid;
                            outp += __popc(ball);
                        }
                    } else {
                        for (int t = lane; t < ((n + 31) & ~31); t += 32) {
                            bool ok = false;
                            if (t < n) ok = contains(Nj, d, nodes[t]) && !((j == 0 && t == 1) || (j == 1 && t == 0));
                            const unsigned ball = __ballot_sync(0xffffffffu, ok);
                            if (ok) lcol[outp + __popc(ball & ((1u << lane) - 1u))] = t;
                            outp += __popc(ball);
                        }
                    }
                }
                __syncthreads();
                // ---- selected rows (PoS Plus) and the float scratch of the work items ----
                overflow = !select_and_allocate(p, nodes, n, A, du, B, dv, s_scan, &s_base, s_sel, base_sel, base3, m);
              if (!overflow) {
                // ---- diffusion of the target rows: K sweeps over every row (num_hops = 1 <= K) ----
                float* item_f = reinterpret_cast<float*>(p.arena + base3);
                float* lab = item_f;
                float* wgt = item_f + NWP;
                const bool z_shared = n * SC <= kZCapS;
                float* zprev = z_shared ? s_z[0] : wgt + (int64_t)n * NWP;
                float* znext = z_shared ? s_z[1] : wgt + (int64_t)n * NWP + (int64_t)n * SC;
                for (int j = tid; j < n; j += T) {
#pragma unroll
                    for (int c = 0; c < SC; ++c) {
                        float zv = 0.0f;
                        if (j == c) {
                            const int deg = rowlen[j];
                            zv = deg > 0 ? 1.0f / sqrtf((float)deg) : 0.0f;
                        }
                        zprev[(int64_t)j * SC + c] = zv;
                        znext[(int64_t)j * SC + c] = 0.0f;
                    }
                }
                for (int i = tid; i < 2 * NWP; i += T) {
                    const int j = i / NWP, q = i - j * NWP;
                    wgt[i] = (q < SC && q == j) ? 1.0f : 0.0f;
                }
                __syncthreads();
                for (int k = 1; k <= K; ++k) {
                    const int nk = n;  // hop <= 1 <= k: every row
                    for (int jb0 = 0; jb0 < nk; jb0 += NG) {
                        const int j = jb0 + grp;
                        const bool valid = j < nk;
                        int e0 = 0, e1 = 0;
                        if (valid) {
                            e0 = rowptr[j];
                            e1 = e0 + rowlen[j];
                        }
                        float t[SC];
#pragma unroll
                        for (int c = 0; c < SC; ++c) t[c] = 0.0f;
                        for (int e = e0 + l8; e < e1; e += 8) {
                            const int i = lcol[e];
#pragma unroll
                            for (int c = 0; c < SC; ++c) t[c] += zprev[(int64_t)i * SC + c];
                        }
#pragma unroll
                        for (int c = 0; c < SC; ++c) {
                            t[c] += __shfl_xor_sync(0xffffffffu, t[c], 4);
                            t[c] += __shfl_xor_sync(0xffffffffu, t[c], 2);
                            t[c] += __shfl_xor_sync(0xffffffffu, t[c], 1);
                        }
                        if (valid && l8 == 0) {
                            const int deg = e1 - e0;
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
                for (int q = tid; q < NWP; q += T) lab[q] = q < NW ? wgt[q] + wgt[NWP + q] : 0.0f;
              }
            }
            }  // per-row path
        }

        for (int d = 16; d > 0; d >>= 1) my_deg += __shfl_down_sync(0xffffffffu, my_deg, d);
        if (lane == 0 && my_deg) atomicAdd(&p.counters[S3_CTR_SUM_D], my_deg);
        for (int d = 16; d > 0; d >>= 1) my_read += __shfl_down_sync(0xffffffffu, my_read, d);
        if (lane == 0 && my_read) atomicAdd(&p.counters[S3_CTR_SUM_READ], my_read);
        __syncthreads();
        if (tid == 0) {
            off[S3_OFF_NODES] = base1;
            off[S3_OFF_ROWPTR] = base1 + n;
            off[S3_OFF_ROWLEN] = base1 + 2 * (int64_t)n + 1;
            off[S3_OFF_LCOL] = base2;
            off[S3_OFF_SEL] = base_sel;
            off[S3_OFF_F32] = base3;
            for (int i = 0; i < S3_NCNT; ++i) cnt[i] = 0;
            cnt[S3_CNT_N] = n;
            cnt[S3_CNT_M] = m;
            cnt[S3_CNT_S] = overflow ? 0 : s_sel;
            cnt[S3_CNT_STATUS] = overflow ? S3_REC_ARENA_OVERFLOW : S3_REC_OK;
            cnt[S3_CNT_PARTNER] = -1;
            cnt[S3_CNT_HOP0] = 2;
            cnt[S3_CNT_HOP0 + 1] = n - 2;
            cnt[S3_CNT_NSTORE] = (n <= kBitCap && !p.store_all) ? 0 : n;  // bit-matrix records keep no CSR
            cnt[S3_CNT_CLASSPOS] = (int)atomicAdd(&p.counters[S3_CTR_CLASS0 + (31 - __clz(n))], 1ull);
            if (overflow) atomicAdd(&p.counters[S3_CTR_ERRORS], 1ull);
            atomicMax(&p.counters[S3_CTR_MAX_N], (unsigned long long)n);
            atomicAdd(&p.counters[S3_CTR_SUM_N], (unsigned long long)n);
        }
    }
}

// hub index: one warp per hub row sets the bits of its hub neighbours
__global__ void __launch_bounds__(256) hub_bits_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                       int64_t num_nodes, const int32_t* __restrict__ hub_id, uint32_t* hub_bits,
                                                       int hub_words) {
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= num_nodes) return;
    const int h = hub_id[v];
    if (h < 0) return;
    const int64_t e0 = indptr[v], e1 = indptr[v + 1];
    for (int64_t e = e0 + lane; e < e1; e += 32) {
        const int hc = hub_id[indices[e]];
        if (hc >= 0) atomicOr(&hub_bits[(size_t)h * hub_words + (hc >> 5)], 1u << (hc & 31));
    }
}

}  // namespace

cudaError_t launch_build_hub_bits(const s3_graph& g, cudaStream_t st) {
    if (!g.hub_id || !g.hub_bits || g.num_hubs <= 0) return cudaErrorInvalidValue;
    const int64_t threads = g.num_nodes * 32;
    hub_bits_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(g.indptr, g.indices, g.num_nodes, g.hub_id,
                                                                       const_cast<uint32_t*>(g.hub_bits), (int)((g.num_hubs + 31) / 32));
    return cudaGetLastError();
}

// defined in extract.cu
cudaError_t launch_order(const s3_batch& b, int64_t num_records, cudaStream_t st);

cudaError_t launch_extract_sorted(const s3_graph& g, const s3_batch& b, cudaStream_t st, int* rc_out) {
    *rc_out = S3_OK;
    SortedParams p;
    p.indptr = g.indptr;
    p.indices = g.indices;
    p.num_nodes = g.num_nodes;
    p.hub_id = (g.hub_id && g.hub_bits && g.num_hubs > 0) ? g.hub_id : nullptr;
    p.hub_bits = g.hub_bits;
    p.hub_words = (int)((g.num_hubs + 31) / 32);
    p.link_src = b.link_src;
    p.link_dst = b.link_dst;
    p.num_records = b.num_links;
    p.sign_k = b.sign_k;
    p.strategy = b.flow == S3_FLOW_POS ? b.strategy : S3_STRATEGY_NONE;
    p.store_all = ((b.flags & S3_BATCH_STORE_ALL_ROWS) || p.strategy != S3_STRATEGY_NONE) ? 1 : 0;  // CCN items sweep the CSR
    p.flags = b.flags;
    p.walk_sets = b.walk_sets;
    p.walk_counts = b.walk_counts;
    p.src_set = b.link_src_set;
    p.dst_set = b.link_dst_set;
    p.walk_cap = b.walk_cap;
    p.caps = batch_caps(b) ? 1 : 0;
    p.cap_ratio = b.ratio_per_hop;
    p.cap_max = b.max_nodes_per_hop;
    p.cap_seed = b.cap_seed;
    p.arena = b.arena;
    p.arena_words = b.arena_words;
    p.off = b.off;
    p.cnt = b.cnt;
    p.counters = reinterpret_cast<unsigned long long*>(b.counters);
    if (p.num_records == 0) return cudaSuccess;
    static LaunchCache cache;  // per device: SM count and occupancy of the persistent kernel
    int c_sms = 0, c_occ = 0;
    cudaError_t e = cache.get(reinterpret_cast<const void*>(front_sorted_kernel), kExtractThreads, 0, &c_sms, &c_occ);
    if (e != cudaSuccess) return e;
    p.slab_stride = (2 * ((b.walk_sets ? (int64_t)b.walk_cap : g.max_degree) + 1) + 31) & ~int64_t(31);
    int64_t grid = (int64_t)c_sms * (c_occ > 0 ? c_occ : 1);
    if (grid > p.num_records) grid = p.num_records;
    const int64_t fit = (b.arena_words / 2) / p.slab_stride;
    if (grid > fit) grid = fit;
    if (grid < 1) {
        *rc_out = S3_ERR_WORKSPACE;
        return cudaSuccess;
    }
    p.slab_words = grid * p.slab_stride;
    front_sorted_kernel<<<(unsigned)grid, kExtractThreads, 0, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess || !b.order) return e;
    return launch_order(b, p.num_records, st);
}

}  // namespace s3
