// Kernel 1 — "front" kernel, bitmap tier: enclosing-subgraph extraction FUSED with the diffusion
// sweeps of the record's first work item (graphs whose node-id bitmaps fit in shared memory:
// every dataset bundled with the reference; N up to a few hundred thousand).
//
// Replaces, per record: reference utils.py:33-44 (neighbors), utils.py:53-74 (BFS),
// utils.py:76 (A[nodes,:][:,nodes]), utils.py:79-80 (target-link mask), the ssp.find at
// tuned_SIGN.py:153/:208, the PoS-Plus row selection tuned_SIGN.py:228-238, and for the rows of
// the two targets tuned_SIGN.py:155-175 (normalise, powers, row select) — in SoP flow
// sgrl_link_pred.py:161-178 + tuned_SIGN.py:60-86, :106-113.
//
// Persistent CTAs (one wave, grid = SMs x resident CTAs) pull records from an atomic work
// counter. Shared memory holds a visited bitmap V, one bitmap per hop level and a per-word
// exclusive popcount prefix per level. A level bitmap enumerated in word order IS the level
// sorted by global id, so the canonical order [seeds, then (hop, global id) ascending] needs no
// sort, and  local_id(g) = level_base + prefix[word(g)] + popc(bits below g)  needs no hash.
// Each CTA owns a slab [nodes | degrees | adjacency offsets | row starts] at the head of the
// arena for the growing node list.
//
// Which rows of the induced adjacency are ever needed?  With w_0 = e_sel, z_k = w_k D^-1/2:
//      t_j = sum_{i in N(j)} z_{k-1}[i],   w_k[j] = dis_j t_j,   z_k[j] = dis_j^2 t_j
// and a k-step walk stays inside the k-hop ball, sweep k touches rows of hop <= k and reads z
// only on hop <= k-1. Rows of hop <= K-1 are read by several sweeps: they are STORED, as a
// padded CSR (row j owns deg_G(j) slots at the prefix sum of global degrees; a neighbour
// outside the subgraph or the masked target link leaves a -1 hole), filled edge-parallel, one
// lane per slot. Rows of hop == K are read exactly once, by the last sweep: they are STREAMED
// — their adjacency is scanned once by 8-lane groups, accumulating z_{K-1} of the inner
// neighbours and the induced degree on the fly — and never stored. Nodes beyond hop K get no
// weight at all. On a 3-hop PubMed subgraph (n ~ 930, of which ~770 on hop 3) this stores
// ~2 k of ~9 k adjacency slots and removes the largest sweep. When later work items (PoS Plus
// CCN rows) or a parity dump need every row, S3_BATCH_STORE_ALL_ROWS stores them all and the
// stand-alone diffuse kernel (diffuse.cu) handles the remaining items.
//
// Results go to a bump-allocated arena in global memory (one atomicAdd per allocation);
// placement depends on scheduling, contents do not. Sums run in fixed order (row slots by
// lane, fixed shuffle tree), so results are independent of scheduling.
#include <climits>

#include "common.cuh"

namespace s3 {
namespace {

struct ExtractParams {
    const int64_t* __restrict__ indptr;
    const int32_t* __restrict__ indices;
    int64_t num_nodes;
    const int64_t* __restrict__ link_src;
    const int64_t* __restrict__ link_dst;
    int64_t num_records;
    int flow, strategy, radius, sign_k, store_all, flags;
    int W;  // bitmap words
    int32_t* arena;
    int64_t arena_words;
    int64_t slab_stride;  // words per CTA slab, slabs sit at the arena head
    int64_t slab_words;   // gridDim.x * slab_stride: bump allocations start here
    int64_t* off;
    int32_t* cnt;
    unsigned long long* counters;
    const int64_t* __restrict__ out_link;  // link pairing (s3_batch): global link index of a record
    const int64_t* __restrict__ mirror;    // chain table of s3_pair_links, or null
    int64_t link_base;
    const int32_t* __restrict__ front_order;  // records in descending size proxy, or null: link order
    int caps;                              // per-hop caps active (s3_batch.ratio_per_hop / max_nodes_per_hop)
    double cap_ratio;
    int cap_max;
    uint32_t cap_seed;
};

#ifndef S3_STREAM_LANES
#define S3_STREAM_LANES 4
#endif
#ifdef S3_V_LDG
#define S3_IDX(ptr, i) __ldg((ptr) + (i))
#else
#define S3_IDX(ptr, i) (ptr)[i]
#endif
#ifndef S3_V_UNROLL
#define S3_V_UNROLL 2  // streamed rows: two adjacency loads in flight per lane (measured: 14.11 -> 13.94 ms per PubMed step;
                       // 3 and 4 no better, `unroll 1` on the sweep / BFS loops 5 % worse than the compiler's own choice)
#endif
#define S3_PRAGMA_(x) _Pragma(#x)
#define S3_UNROLL(n) S3_PRAGMA_(unroll n)
constexpr int kStreamLanes = S3_STREAM_LANES;  // lanes per streamed (hop-K) row
constexpr int kBfsLanes = 4;     // lanes per frontier node in the BFS expansion
constexpr int kSweepLanes = 4;   // lanes per stored row in the diffusion sweeps
constexpr int kRowCap = 1536;  // rows whose (start, adjacency offset) are cached in shared memory
constexpr int kZCap = 1024;    // floats per shared z buffer

__device__ __forceinline__ bool test_bit(const uint32_t* bm, int g) { return (bm[g >> 5] >> (g & 31)) & 1u; }

// local id of global node g if it sits on one of the first `levels` hop levels (or is a seed),
// else -1. Block-uniform trip count, no early exit: no divergence.
__device__ __forceinline__ int local_id(int g, int s0, int s1, int nseed, const uint32_t* Lb, const uint32_t* pre,
                                        const int* lvl_base, int levels, int W) {
    const int w = g >> 5, b = g & 31;
    const uint32_t below = (1u << b) - 1u;
    int lid = -1;
    for (int l = 0; l < levels; ++l) {
        const uint32_t bits = Lb[l * W + w];
        const int cand = lvl_base[l] + (int)pre[l * W + w] + __popc(bits & below);
        lid = ((bits >> b) & 1u) ? cand : lid;
    }
    if (nseed == 2 && g == s1) lid = 1;
    if (g == s0) lid = 0;
    return lid;
}

// Deterministic per-hop cap (reference utils.py:66-70): keep the `keep` nodes of the level bitmap `cur` with the
// smallest fmix32(node ^ seed) — 4 x 8-bit radix select over the rank keys (a bijection of the ids: no ties), then
// the dropped bits are cleared. Whole CTA; `hist` = 256 + 2 ints of shared scratch. Not on the hot path of the
// reference's configurations (none caps), hence not inlined.
__device__ __noinline__ void cap_level(uint32_t* cur, int w0, int w1, int keep, uint32_t seed, int* hist) {
    int* sel = hist + 256;
    if (threadIdx.x == 0) {
        sel[0] = 0;
        sel[1] = keep;
    }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = (uint32_t)sel[0];
        for (int w = w0; w < w1; ++w) {
            uint32_t bb = cur[w];
            while (bb) {
                const int bit = __ffs(bb) - 1;
                bb &= bb - 1;
                const uint32_t hsh = fmix32((uint32_t)(w * 32 + bit) ^ seed);
                if (pass == 0 || (hsh >> (shift + 8)) == prefix) atomicAdd(&hist[(hsh >> shift) & 255u], 1);
            }
        }
        radix_pick(hist, sel);
    }
    const uint32_t thr = (uint32_t)sel[0];  // rank key of the keep-th smallest
    for (int w = w0; w < w1; ++w) {
        uint32_t bb = cur[w], kept = 0u;
        while (bb) {
            const int bit = __ffs(bb) - 1;
            bb &= bb - 1;
            if (fmix32((uint32_t)(w * 32 + bit) ^ seed) <= thr) kept |= 1u << bit;
        }
        cur[w] = kept;
    }
    __syncthreads();
}

template <int SC>  // selected rows of the first work item == number of seeds: 2 (PoS), 1 (SoP)
#ifndef S3_FRONT_BLOCKS
#define S3_FRONT_BLOCKS 5  // resident CTAs per SM the register budget is sized for (48 registers)
#endif
__global__ void __launch_bounds__(kExtractThreads, S3_FRONT_BLOCKS) front_kernel(ExtractParams p) {
    extern __shared__ uint32_t sm[];
    const int W = p.W, h = p.radius, T = kExtractThreads, tid = threadIdx.x, K = p.sign_k;
    constexpr int nseed = SC;
    uint32_t* V = sm;
    uint32_t* Lb = V + W;                // [h][W]
    uint32_t* pre = Lb + (size_t)h * W;  // [h][W]
    __shared__ int s_scan[33];
    __shared__ int s_lvl_base[S3_MAX_HOPS + 2];
    __shared__ int s_lvl_cnt[S3_MAX_HOPS + 1];
    __shared__ int s_hop_end[S3_MAX_HOPS + 2];
    __shared__ long long s_base;
    __shared__ long long s_rec;
    __shared__ unsigned long long s_sumdeg;
    __shared__ int s_m;
    __shared__ int s_rowptr[kRowCap + 1];   // row starts of the current record when n_rows <= kRowCap
    __shared__ uint32_t s_estart[kRowCap];  // indptr[node] of every row
    __shared__ float s_z[2][kZCap];

    int32_t* slab_nodes = p.arena + (int64_t)blockIdx.x * p.slab_stride;
    int32_t* slab_deg = slab_nodes + p.num_nodes;
    uint32_t* slab_estart = reinterpret_cast<uint32_t*>(slab_deg + p.num_nodes);
    int32_t* slab_rp = slab_deg + 2 * p.num_nodes;  // [N + 1]
    const bool pos_flow = p.flow == S3_FLOW_POS;
    const int lane = tid & 31;
    // streamed hop-K rows are mostly low-degree: narrower lane groups leave fewer lanes idle
    constexpr int NGW = kExtractThreads / kSweepLanes;
    const int lw = tid & (kSweepLanes - 1), grpw = tid / kSweepLanes;
    constexpr int NGB = kExtractThreads / kBfsLanes;
    const int lb = tid & (kBfsLanes - 1), grpb = tid / kBfsLanes;
    constexpr int NGS = kExtractThreads / kStreamLanes;
    const int ls = tid & (kStreamLanes - 1), grps = tid / kStreamLanes;
    const int NW = (K + 1) * SC, NWP = (NW + 3) & ~3;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_rec = (long long)atomicAdd(&p.counters[S3_CTR_WORK], 1ull);
        __syncthreads();
        if (s_rec >= p.num_records) break;
        const int64_t rec = p.front_order ? (int64_t)p.front_order[s_rec] : (int64_t)s_rec;  // longest first

        int64_t a, b;
        if (pos_flow) {
            a = p.link_src[rec];
            b = p.link_dst[rec];
        } else {
            const int64_t l = rec >> 1;
            a = (rec & 1) ? p.link_dst[l] : p.link_src[l];
            b = (rec & 1) ? p.link_src[l] : p.link_dst[l];
        }
        int32_t* cnt = p.cnt + rec * S3_NCNT;
        int64_t* off = p.off + rec * S3_NOFF;
        const int64_t gl = !p.mirror ? 0 : (p.out_link ? p.out_link[rec] : p.link_base + rec);  // pairing: PoS only
        if (p.mirror && p.mirror[gl] <= -2) {
            // an earlier link over the same node pair does the work and writes this link's rows (pair.cu)
            if (tid == 0) {
                for (int i = 0; i < S3_NCNT; ++i) cnt[i] = 0;
                for (int i = 0; i < S3_NOFF; ++i) off[i] = 0;
                cnt[S3_CNT_STATUS] = S3_REC_MIRROR;
                cnt[S3_CNT_PARTNER] = -1;
                atomicAdd(&p.counters[S3_CTR_MIRRORS], 1ull);
            }
            continue;
        }
        if (a < 0 || b < 0 || a >= p.num_nodes || b >= p.num_nodes || a == b) {
            if (tid == 0) {
                for (int i = 0; i < S3_NCNT; ++i) cnt[i] = 0;
                for (int i = 0; i < S3_NOFF; ++i) off[i] = 0;
                cnt[S3_CNT_STATUS] = S3_REC_BAD_LINK;
                cnt[S3_CNT_PARTNER] = -1;
                atomicAdd(&p.counters[S3_CTR_ERRORS], 1ull);
            }
            continue;
        }
        const int s0 = (int)a, s1 = (int)b;  // s1 is the second seed (PoS) or the partner (SoP)

        for (int i = tid; i < (1 + h) * W; i += T) sm[i] = 0u;  // V and the level bitmaps
        __syncthreads();
        unsigned long long my_deg = 0;  // global degrees of the nodes this thread emitted (D of SURVEY 8d)
        if (tid == 0) {
            V[s0 >> 5] |= 1u << (s0 & 31);
            slab_nodes[0] = s0;
            const int d0 = (int)(p.indptr[s0 + 1] - p.indptr[s0]);
            slab_deg[0] = d0;
            slab_estart[0] = (uint32_t)p.indptr[s0];
            my_deg += d0;
            if (nseed == 2) {
                V[s1 >> 5] |= 1u << (s1 & 31);
                slab_nodes[1] = s1;
                const int d1 = (int)(p.indptr[s1 + 1] - p.indptr[s1]);
                slab_deg[1] = d1;
                slab_estart[1] = (uint32_t)p.indptr[s1];
                my_deg += d1;
            }
            s_lvl_cnt[0] = nseed;
            s_sumdeg = 0ull;
            s_m = 0;
        }
        __syncthreads();

        // ---------------- BFS: h rounds of frontier expansion (utils.py:57-74) ----------------
        const int chunk = (W + T - 1) / T;
        const int w0 = min(W, tid * chunk), w1 = min(W, w0 + chunk);
        int nlev = 0, n = nseed, flo = 0;
        bool capped = false;
        for (int l = 0; l < h; ++l) {
            uint32_t* cur = Lb + (size_t)l * W;
            // frontier = slab_nodes[flo, n): one 8-lane group per node, 32 B of column ids per step
            for (int j = flo + grpb; j < n; j += NGB) {
                const int64_t e0 = slab_estart[j], e1 = e0 + slab_deg[j];
                for (int64_t e = e0 + lb; e < e1; e += kBfsLanes) {
                    const int c = S3_IDX(p.indices, e);
                    if (!test_bit(V, c)) atomicOr(&cur[c >> 5], 1u << (c & 31));
                }
            }
            __syncthreads();
            // fold the level into V, count it, build its per-word prefix
            int local = 0;
            for (int w = w0; w < w1; ++w) {
                const uint32_t bits = cur[w];
                V[w] |= bits;
                local += __popc(bits);
            }
            int total;
            int run = block_exclusive_scan(local, s_scan, &total);
            if (p.caps) {  // utils.py:66-70; V keeps the dropped nodes (visited), the level bitmap loses them
                const int keep = cap_keep(total, p.cap_ratio, p.cap_max);
                if (keep < total) {  // block-uniform
                    cap_level(cur, w0, w1, keep, p.cap_seed, reinterpret_cast<int*>(&s_z[0][0]));
                    capped = true;
                    local = 0;
                    for (int w = w0; w < w1; ++w) local += __popc(cur[w]);
                    run = block_exclusive_scan(local, s_scan, &total);
                }
            }
            uint32_t* pl = pre + (size_t)l * W;
            for (int w = w0; w < w1; ++w) {
                const uint32_t bits = cur[w];
                pl[w] = (uint32_t)run;
                // emit the level's nodes (ascending global id) and their global degrees
                uint32_t bb = bits;
                int r = n + run;
                while (bb) {
                    const int bit = __ffs(bb) - 1;
                    bb &= bb - 1;
                    const int g = w * 32 + bit;
                    const int64_t ge0 = p.indptr[g];
                    const int d = (int)(p.indptr[g + 1] - ge0);
                    slab_nodes[r] = g;
                    slab_deg[r] = d;
                    slab_estart[r] = (uint32_t)ge0;
                    my_deg += d;
                    ++r;
                }
                run += __popc(bits);
            }
            if (tid == 0) {
                s_lvl_base[l] = n;
                s_lvl_cnt[l + 1] = total;
            }
            if (total == 0) break;  // block-uniform (utils.py:71-72)
            flo = n;
            n += total;
            nlev = l + 1;
            __syncthreads();
        }
        if (capped) {  // V = members only from here on (it was "visited" during the BFS): seeds + kept levels
            __syncthreads();
            for (int i = tid; i < W; i += T) {
                uint32_t bits = 0u;
                for (int l = 0; l < h; ++l) bits |= Lb[(size_t)l * W + i];
                V[i] = bits;
            }
            __syncthreads();
            if (tid == 0) {
                V[s0 >> 5] |= 1u << (s0 & 31);
                if (nseed == 2) V[s1 >> 5] |= 1u << (s1 & 31);
            }
            __syncthreads();
        }
        // D = sum of global degrees (block reduction); hop_end[l] = nodes of hop <= l
        for (int d = 16; d > 0; d >>= 1) my_deg += __shfl_down_sync(0xffffffffu, my_deg, d);
        if (lane == 0 && my_deg) atomicAdd(&s_sumdeg, my_deg);
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int l = 0; l <= S3_MAX_HOPS + 1; ++l) {
                if (l <= nlev) acc += s_lvl_cnt[l];
                s_hop_end[l] = acc;
            }
        }
        __syncthreads();
        const int64_t D = (int64_t)s_sumdeg;
        const int n_reach = s_hop_end[min(K, S3_MAX_HOPS + 1)];  // nodes that get any weight
        const int nz = s_hop_end[min(K - 1, S3_MAX_HOPS + 1)];   // support of z_k, k <= K-1
        const int n_store = p.store_all ? n : nz;                // rows kept in the padded CSR
        const int n_rows = max(n_store, n_reach);                // rows whose start is needed
        const bool cached = n_rows <= kRowCap;

        // ---------------- row starts: exclusive scan of global degrees, tile by tile ----------------
        {
            int running = 0;
            for (int base = 0; base < n_rows; base += T) {
                const int j = base + tid;
                const int d = j < n_rows ? slab_deg[j] : 0;
                int tile_total;
                const int ex = block_exclusive_scan(d, s_scan, &tile_total);
                if (j < n_rows) {
                    slab_rp[j] = running + ex;
                    if (cached) {
                        s_rowptr[j] = running + ex;
                        s_estart[j] = slab_estart[j];
                    }
                }
                running += tile_total;
                __syncthreads();  // s_scan is reused by the next tile
            }
            if (tid == 0) {
                slab_rp[n_rows] = running;
                if (cached) s_rowptr[n_rows] = running;
            }
            __syncthreads();
        }
        const int* rp = cached ? s_rowptr : slab_rp;
        const uint32_t* es = cached ? s_estart : slab_estart;
        const int Ds = rp[n_store];  // slots of the stored rows

        // ---------------- allocation 1: every integer array of the record ----------------
        int sel_bound = 0;
        if (p.strategy != S3_STRATEGY_NONE) {
            const int d0 = slab_deg[0], d1 = slab_deg[1];
            sel_bound = p.strategy == S3_STRATEGY_INTERSECTION ? min(d0, d1) : d0 + d1;
        }
        const int64_t words1 = ((int64_t)n + (n_store + 1) + n_store + Ds + sel_bound + 31) & ~int64_t(31);
        if (tid == 0) s_base = (long long)atomicAdd(&p.counters[S3_CTR_CURSOR], (unsigned long long)words1);
        __syncthreads();
        const int64_t base1 = p.slab_words + s_base;
        bool overflow = base1 + words1 > p.arena_words;
        int32_t* nodes = p.arena + base1;
        int32_t* rowptr = nodes + n;             // [n_store+1] row starts (prefix of global degrees)
        int32_t* rowlen = rowptr + n_store + 1;  // [n_store]   induced, masked degree
        int32_t* lcol = rowlen + n_store;        // [Ds]        padded local column ids, -1 = hole
        int32_t* sel = lcol + Ds;

        int s = nseed, partner_local = -1;
        int64_t base3 = 0;
        if (!overflow) {
            for (int j = tid; j < n; j += T) nodes[j] = slab_nodes[j];
            for (int j = tid; j <= n_store; j += T) {
                rowptr[j] = rp[j];
                if (j < n_store) rowlen[j] = 0;
            }
            __syncthreads();

            // ---------------- fill: one lane per adjacency slot of the stored rows ----------------
            // Each warp streams a contiguous chunk of slots, 32 per step. The row of a step's first
            // slot is carried from the previous step (one binary search per chunk); the other lanes
            // find theirs from a 32-bit mask of the row starts inside the step's window (rows 2..
            // own >= 1 slot, so at most 32 rows start in 32 slots).
            int my_m = 0;
            const int chunk_slots = (((Ds + (T >> 5) - 1) / (T >> 5)) + 31) & ~31;
            const int c_lo = (tid >> 5) * chunk_slots, c_hi = min(Ds, c_lo + chunk_slots);
            int jb = 0;
            if (c_lo < c_hi) {
                int lo = 0, hi = n_store;  // rp[lo] <= c_lo < rp[hi]; picks the last of equal starts
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (rp[mid] <= c_lo) lo = mid; else hi = mid;
                }
                jb = lo;
            }
            for (int slot0 = c_lo; slot0 < c_hi; slot0 += 32) {
                const int slot = slot0 + lane;
                const bool ok = slot < c_hi;
                const int r = jb + 1 + lane;
                const int pos = (r <= n_store ? rp[r] : INT_MAX) - slot0;  // start of row r relative to the window
                const unsigned heads = __reduce_or_sync(0xffffffffu, (pos >= 0 && pos < 32) ? (1u << pos) : 0u);
                int j = jb + __popc(heads & (0xffffffffu >> (31 - lane)));
                if (j + 1 < n_store && rp[j + 1] <= slot) ++j;  // a degree-0 seed (row 1) shares its start with row 2
                jb = __shfl_sync(0xffffffffu, j, 31);
                int lid = -1;
                bool in = false;
                if (ok) {
                    const int c = S3_IDX(p.indices, (int64_t)es[j] + (slot - rp[j]));
                    in = test_bit(V, c);
                    if (pos_flow && ((j == 0 && c == s1) || (j == 1 && c == s0))) in = false;  // utils.py:79-80
                    if (in) lid = local_id(c, s0, s1, nseed, Lb, pre, s_lvl_base, nlev, W);
                    lcol[slot] = lid;
                }
                // induced degree: one integer atomic per (row, warp step)
                const unsigned seg = __match_any_sync(0xffffffffu, ok ? j : -1 - lane);
                const int kept = __popc(__ballot_sync(0xffffffffu, in) & seg);
                if (ok && kept && (__ffs(seg) - 1) == lane) atomicAdd(&rowlen[j], kept);
                my_m += in ? 1 : 0;
            }
            for (int d = 16; d > 0; d >>= 1) my_m += __shfl_down_sync(0xffffffffu, my_m, d);
            if (lane == 0 && my_m) atomicAdd(&s_m, my_m);
            if (!pos_flow && test_bit(V, s1)) partner_local = local_id(s1, s0, s1, nseed, Lb, pre, s_lvl_base, nlev, W);
            __syncthreads();

            // ---------------- row selection (PoS Plus): tuned_SIGN.py:228-238 ----------------
            // Neighbours of local 0 / 1 are hop-1 nodes, local ids [2, 2 + cnt1). Two flag bitmaps
            // over that range (carved from the z scratch, or the dead row-start slab when huge)
            // give the intersection or union in ascending local id.
            if (p.strategy != S3_STRATEGY_NONE) {
                const int cnt1 = nlev > 0 ? s_lvl_cnt[1] : 0;
                const int FW = (cnt1 + 31) >> 5;
                const bool fshared = 2 * FW <= 2 * kZCap;
                uint32_t* F0 = fshared ? reinterpret_cast<uint32_t*>(&s_z[0][0]) : reinterpret_cast<uint32_t*>(slab_nodes);
                uint32_t* F1 = F0 + FW;
                for (int w = tid; w < 2 * FW; w += T) F0[w] = 0u;
                __syncthreads();
                for (int e = rowptr[0] + tid; e < rowptr[1]; e += T) {
                    const int c = lcol[e] - 2;  // holes are -1, seeds 0/1: both excluded
                    if (c >= 0) atomicOr(&F0[c >> 5], 1u << (c & 31));
                }
                for (int e = rowptr[1] + tid; e < rowptr[2]; e += T) {
                    const int c = lcol[e] - 2;
                    if (c >= 0) atomicOr(&F1[c >> 5], 1u << (c & 31));
                }
                __syncthreads();
                const int fc = (FW + T - 1) / T;
                const int f0 = min(FW, tid * fc), f1 = min(FW, f0 + fc);
                const bool inter = p.strategy == S3_STRATEGY_INTERSECTION;
                int local = 0;
                for (int w = f0; w < f1; ++w) local += __popc(inter ? (F0[w] & F1[w]) : (F0[w] | F1[w]));
                int extra;
                int run = block_exclusive_scan(local, s_scan, &extra);
                for (int w = f0; w < f1; ++w) {
                    uint32_t bits = inter ? (F0[w] & F1[w]) : (F0[w] | F1[w]);
                    while (bits) {
                        const int bit = __ffs(bits) - 1;
                        bits &= bits - 1;
                        sel[run++] = 2 + w * 32 + bit;
                    }
                }
                s = nseed + extra;
                __syncthreads();
            }

            // ---------------- allocation 2: float scratch of the record's work items ----------------
            // records whose CCN rows go through s3_ccn_chain need no work-item scratch for them (s3_plan counts none)
            const bool chained = chain_eligible(p.flags, p.strategy, n, s_m, s_hop_end[1]);
            const int64_t items3 = chained ? 0 : ccn_items(s, nseed, ccn_rows(p.strategy));
            const int64_t words3 = (item_words(p.flow, K, n) + items3 * ccn_item_words(K, n, ccn_rows(p.strategy)) + 31) & ~int64_t(31);
            if (tid == 0) s_base = (long long)atomicAdd(&p.counters[S3_CTR_CURSOR], (unsigned long long)words3);
            __syncthreads();
            base3 = p.slab_words + s_base;
            overflow = base3 + words3 > p.arena_words;
        }

        // ---------------- diffusion of work item 0 (selected rows = the seeds) ----------------
        if (!overflow) {
            float* item_f = reinterpret_cast<float*>(p.arena + base3);
            float* lab = item_f;
            float* wgt = item_f + NWP;
            const bool z_shared = nz * SC <= kZCap;
            float* zprev = z_shared ? s_z[0] : wgt + (int64_t)n * NWP;
            float* znext = z_shared ? s_z[1] : wgt + (int64_t)n * NWP + (int64_t)n * SC;

            // k = 0: one-hot rows of the seeds; z_0 = e_sel D^-1/2 on the support, z buffers zeroed
            for (int j = tid; j < nz; j += T) {
#pragma unroll
                for (int c = 0; c < SC; ++c) {
                    float zv = 0.0f;
                    if (j == c) {
                        const int deg = pos_flow ? rowlen[j] : slab_deg[j];  // tuned_SIGN.py:158 / sgrl_link_pred.py:165
                        zv = deg > 0 ? 1.0f / sqrtf((float)deg) : 0.0f;     // inf -> 0 (tuned_SIGN.py:159-160)
                    }
                    zprev[j * SC + c] = zv;
                    znext[j * SC + c] = 0.0f;
                }
            }
            for (int i = tid; i < nseed * NWP; i += T) {
                const int j = i / NWP, q = i - j * NWP;
                wgt[i] = (q < SC && q == j) ? 1.0f : 0.0f;
            }
            __syncthreads();

            for (int k = 1; k <= K; ++k) {
                // stored rows reached by a k-step walk
                const int nk = min(s_hop_end[min(k, S3_MAX_HOPS + 1)], n_store);
                for (int jb0 = 0; jb0 < nk; jb0 += NGW) {
                    const int j = jb0 + grpw;
                    const bool valid = j < nk;
                    int e0 = 0, e1 = 0;
                    if (valid) {
                        e0 = rp[j];
                        e1 = rp[j + 1];
                    }
                    float t[SC];
#pragma unroll
                    for (int c = 0; c < SC; ++c) t[c] = 0.0f;
                    for (int e = e0 + lw; e < e1; e += kSweepLanes) {
                        const int i = lcol[e];  // -1: hole; >= nz: z is an exact zero there
                        if (i >= 0 && i < nz) {
#pragma unroll
                            for (int c = 0; c < SC; ++c) t[c] += zprev[i * SC + c];
                        }
                    }
#pragma unroll
                    for (int c = 0; c < SC; ++c) {
#pragma unroll
                        for (int d = kSweepLanes / 2; d > 0; d >>= 1) t[c] += __shfl_xor_sync(0xffffffffu, t[c], d);
                    }
                    if (valid && lw == 0) {
                        const int deg = pos_flow ? rowlen[j] : slab_deg[j];
                        const float dis = deg > 0 ? 1.0f / sqrtf((float)deg) : 0.0f;
#pragma unroll
                        for (int c = 0; c < SC; ++c) {
                            const float w = dis * t[c];
                            wgt[(int64_t)j * NWP + k * SC + c] = w;
                            if (j < nz) znext[j * SC + c] = dis * w;
                        }
                    }
                }
                if (k == K) {
                    // streamed rows: hop-K nodes, read once, never stored. Their neighbours sit on
                    // hops K-1, K, K+1 and z_{K-1} lives on hop K-1 only: ONE level bitmap (or, for
                    // K = 1, the seed ids) decides whether a neighbour contributes.
                    const int zl = K - 2;  // level index of hop K-1 (levels are hops 1..h); -1: the seeds
                    const uint32_t* Lz = Lb + (size_t)max(zl, 0) * W;
                    const uint32_t* Pz = pre + (size_t)max(zl, 0) * W;
                    const int zbase = zl >= 0 ? s_lvl_base[zl] : 0;
                    for (int jb0 = n_store; jb0 < n_reach; jb0 += NGS) {
                        const int j = jb0 + grps;
                        const bool valid = j < n_reach;
                        int64_t e0 = 0;
                        int len = 0;
                        if (valid) {
                            e0 = es[j];
                            len = rp[j + 1] - rp[j];
                        }
                        float t[SC];
#pragma unroll
                        for (int c = 0; c < SC; ++c) t[c] = 0.0f;
                        int cntv = 0;
S3_UNROLL(S3_V_UNROLL)
                        for (int idx = ls; idx < len; idx += kStreamLanes) {
                            const int c_ = S3_IDX(p.indices, e0 + idx);
                            const int w = c_ >> 5, b = c_ & 31;
                            cntv += (V[w] >> b) & 1u;
                            int i = -1;
                            if (zl >= 0) {
                                const uint32_t bits = Lz[w];
                                if ((bits >> b) & 1u) i = zbase + (int)Pz[w] + __popc(bits & ((1u << b) - 1u));
                            } else {
                                i = c_ == s0 ? 0 : ((nseed == 2 && c_ == s1) ? 1 : -1);
                            }
                            if (i >= 0) {
#pragma unroll
                                for (int c = 0; c < SC; ++c) t[c] += zprev[i * SC + c];
                            }
                        }
#pragma unroll
                        for (int c = 0; c < SC; ++c) {
#pragma unroll
                            for (int d = kStreamLanes / 2; d > 0; d >>= 1) t[c] += __shfl_xor_sync(0xffffffffu, t[c], d);
                        }
#pragma unroll
                        for (int d = kStreamLanes / 2; d > 0; d >>= 1) cntv += __shfl_xor_sync(0xffffffffu, cntv, d);
                        if (valid && ls == 0) {
                            const int deg = pos_flow ? cntv : len;
                            const float dis = deg > 0 ? 1.0f / sqrtf((float)deg) : 0.0f;
#pragma unroll
                            for (int c = 0; c < SC; ++c) wgt[(int64_t)j * NWP + K * SC + c] = dis * t[c];
                        }
                    }
                }
                __syncthreads();
                float* tmp = zprev;
                zprev = znext;
                znext = tmp;
            }

            // label / self-return column of every operator, then (SoP) drop the partner's weight
            if (pos_flow) {
                // x_k[sel, 0] = sum_j w_k[j] * label_j, label = 1 on local 0 and 1 (tuned_SIGN.py:177)
                for (int q = tid; q < NWP; q += T) lab[q] = q < NW ? wgt[q] + wgt[NWP + q] : 0.0f;
            } else {
                // x_k[., 0] = A^k[u,u] (tuned_SIGN.py:106-113); x[., 0] = 1 (tuned_SIGN.py:119-124)
                for (int q = tid; q < NWP; q += T) lab[q] = q < NW ? wgt[q] : 0.0f;
                __syncthreads();
                if (partner_local >= 0 && partner_local < n_reach)  // r_u[v] = 0 (tuned_SIGN.py:73-76)
                    for (int q = SC + tid; q < NW; q += T) wgt[(int64_t)partner_local * NWP + q] = 0.0f;
            }
        }

        __syncthreads();
        if (tid == 0) {
            off[S3_OFF_NODES] = base1;
            off[S3_OFF_ROWPTR] = base1 + n;
            off[S3_OFF_ROWLEN] = base1 + n + n_store + 1;
            off[S3_OFF_LCOL] = base1 + n + 2 * (int64_t)n_store + 1;
            off[S3_OFF_SEL] = base1 + n + 2 * (int64_t)n_store + 1 + Ds;
            off[S3_OFF_F32] = base3;
            cnt[S3_CNT_N] = n;
            cnt[S3_CNT_M] = s_m;
            cnt[S3_CNT_S] = overflow ? 0 : s;
            cnt[S3_CNT_STATUS] = overflow ? S3_REC_ARENA_OVERFLOW : S3_REC_OK;
            cnt[S3_CNT_PARTNER] = partner_local;
            for (int l = 0; l <= S3_MAX_HOPS; ++l) cnt[S3_CNT_HOP0 + l] = (l <= nlev) ? s_lvl_cnt[l] : 0;
            cnt[S3_CNT_NSTORE] = n_store;
            // size class for the largest-first schedule of kernel 3
            cnt[S3_CNT_CLASSPOS] = (int)atomicAdd(&p.counters[S3_CTR_CLASS0 + (31 - __clz(n))], 1ull);
            if (overflow) atomicAdd(&p.counters[S3_CTR_ERRORS], 1ull);
            atomicMax(&p.counters[S3_CTR_MAX_N], (unsigned long long)n);
            atomicAdd(&p.counters[S3_CTR_SUM_N], (unsigned long long)n);
            atomicAdd(&p.counters[S3_CTR_SUM_D], (unsigned long long)D);
            // per-LINK figures (SURVEY 8d sums over links): this record also serves its chain members
            unsigned long long served = 1;
            if (p.mirror) {
                for (long long m = p.mirror[gl]; m >= 0; m = ((-2 - (long long)p.mirror[m]) >> 1) - 1) ++served;
            }
            atomicAdd(&p.counters[S3_CTR_SUM_N_ALL], served * (unsigned long long)n);
            atomicAdd(&p.counters[S3_CTR_SUM_D_ALL], served * (unsigned long long)D);
        }
    }
}

// order[] = records sorted by descending size class (counting sort over the class counters the
// front kernel filled; arrival order inside a class is scheduling dependent, results are not).
__global__ void order_kernel(const int32_t* __restrict__ cnt, const unsigned long long* __restrict__ counters,
                             int64_t num_records, int32_t* __restrict__ order) {
    __shared__ int s_off[32];
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int c = 31; c >= 0; --c) {
            s_off[c] = acc;
            acc += (int)counters[S3_CTR_CLASS0 + c];
        }
    }
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= num_records) return;
    const int32_t* c = cnt + r * S3_NCNT;
    if (c[S3_CNT_STATUS] == S3_REC_BAD_LINK || c[S3_CNT_STATUS] == S3_REC_MIRROR) return;  // not classed; slots stay -1
    const int n = c[S3_CNT_N];
    order[s_off[31 - __clz(n)] + c[S3_CNT_CLASSPOS]] = (int32_t)r;
}

// ---- longest-first hand-out of the records: counting sort by a size proxy over 128 logarithmic classes ----
__device__ __forceinline__ int proxy_class(long long key) {  // 4 * floor(log2 key) + the next two bits, 0..127
    if (key < 4) return (int)key;
    const int hb = 63 - __clzll(key);
    return min(127, 4 * hb + (int)((key >> (hb - 2)) & 3));
}

__device__ __forceinline__ long long record_key(const ExtractParams& p, const int32_t* __restrict__ proxy, int64_t r) {
    const int64_t l = p.flow == S3_FLOW_POS ? r : (r >> 1);
    const int64_t a = p.link_src[l], b = p.link_dst[l];
    if (a < 0 || b < 0 || a >= p.num_nodes || b >= p.num_nodes) return 0;
    if (p.flow != S3_FLOW_POS) return proxy[(r & 1) ? b : a];
    if (p.mirror) {
        const int64_t gl = p.out_link ? p.out_link[r] : p.link_base + r;
        if (p.mirror[gl] <= -2) return 0;  // served by another record: no work
    }
    return (long long)proxy[a] + proxy[b];
}

__global__ void front_hist_kernel(ExtractParams p, const int32_t* __restrict__ proxy, int* hist) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < p.num_records) atomicAdd(&hist[proxy_class(record_key(p, proxy, r))], 1);
}

__global__ void front_scatter_kernel(ExtractParams p, const int32_t* __restrict__ proxy, int* hist, int32_t* order) {
    __shared__ int s_start[128];
    if (threadIdx.x == 0) {  // descending classes; hist[128 + c] is the running cursor of class c
        int acc = 0;
        for (int c = 127; c >= 0; --c) {
            s_start[c] = acc;
            acc += hist[c];
        }
    }
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.num_records) return;
    const int c = proxy_class(record_key(p, proxy, r));
    order[s_start[c] + atomicAdd(&hist[128 + c], 1)] = (int32_t)r;
}

// size proxy of every node: its degree plus its neighbours' degrees (one warp per node)
__global__ void __launch_bounds__(256) node_proxy_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                         int64_t num_nodes, int32_t* __restrict__ out) {
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= num_nodes) return;
    const int64_t e0 = indptr[v], e1 = indptr[v + 1];
    long long acc = 0;
    for (int64_t e = e0 + lane; e < e1; e += 32) {
        const int c = indices[e];
        acc += indptr[c + 1] - indptr[c];
    }
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    if (lane == 0) out[v] = (int32_t)min((long long)INT_MAX, acc + (e1 - e0));
}

}  // namespace

cudaError_t launch_node_proxy(const s3_graph& g, int32_t* out, cudaStream_t st) {
    const int64_t threads = g.num_nodes * 32;
    node_proxy_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(g.indptr, g.indices, g.num_nodes, out);
    return cudaGetLastError();
}

cudaError_t launch_order(const s3_batch& b, int64_t num_records, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(b.order, 0xff, (size_t)num_records * 4, st);  // -1: skipped by kernel 3
    if (e != cudaSuccess) return e;
    order_kernel<<<(unsigned)((num_records + 255) / 256), 256, 0, st>>>(
        b.cnt, reinterpret_cast<const unsigned long long*>(b.counters), num_records, b.order);
    return cudaGetLastError();
}

namespace {

template <int SC>
cudaError_t launch_front(ExtractParams& p, const s3_graph& g, const s3_batch& b, cudaStream_t st, int* rc_out) {
    const size_t smem = (size_t)s3_extract_smem_bytes(g.num_nodes, p.radius);
    // function attributes and occupancy are per device: cache them per (device, smem), under a lock
    static LaunchCache cache;
    int sms = 0, occ = 0;
    cudaError_t e = cache.get(reinterpret_cast<const void*>(front_kernel<SC>), kExtractThreads, smem, &sms, &occ);
    if (e != cudaSuccess) return e;
    p.slab_stride = (4 * g.num_nodes + 1 + 31) & ~int64_t(31);
    // S3_BATCH_SHARE_SMS: the caller runs kernel 3 of the previous batch beside this launch (two-stream schedule of the
    // multi-GPU exchange, where kernel 3 waits on NVLink): leave two CTA slots per SM's worth of registers to it
    if ((b.flags & S3_BATCH_SHARE_SMS) && occ > 3) occ = 3;
    int64_t grid = (int64_t)sms * (occ > 0 ? occ : 1);
    if (grid > p.num_records) grid = p.num_records;
    const int64_t fit = (b.arena_words / 2) / p.slab_stride;  // slabs may take at most half of the arena
    if (grid > fit) grid = fit;
    if (grid < 1) {
        *rc_out = S3_ERR_WORKSPACE;
        return cudaSuccess;
    }
    p.slab_words = grid * p.slab_stride;
    p.front_order = nullptr;
    if (b.front_order && g.size_proxy && p.num_records > 2 * grid && b.arena_words >= 256) {
        // the histogram lives at the head of the arena: the slabs there are not written before the front kernel starts
        int* hist = reinterpret_cast<int*>(b.arena);
        e = cudaMemsetAsync(hist, 0, 256 * sizeof(int), st);
        if (e != cudaSuccess) return e;
        const unsigned blocks = (unsigned)((p.num_records + 255) / 256);
        front_hist_kernel<<<blocks, 256, 0, st>>>(p, g.size_proxy, hist);
        front_scatter_kernel<<<blocks, 256, 0, st>>>(p, g.size_proxy, hist, b.front_order);
        p.front_order = b.front_order;
    }
    front_kernel<SC><<<(unsigned)grid, kExtractThreads, smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess || !b.order) return e;
    return launch_order(b, p.num_records, st);
}

}  // namespace

cudaError_t launch_extract_bitmap(const s3_graph& g, const s3_batch& b, cudaStream_t st, int* rc_out) {
    *rc_out = S3_OK;
    ExtractParams p;
    p.indptr = g.indptr;
    p.indices = g.indices;
    p.num_nodes = g.num_nodes;
    p.link_src = b.link_src;
    p.link_dst = b.link_dst;
    p.num_records = s3_num_records(&b);
    p.flow = b.flow;
    p.strategy = b.flow == S3_FLOW_POS ? b.strategy : S3_STRATEGY_NONE;
    p.radius = b.flow == S3_FLOW_POS ? b.num_hops : b.sign_k;
    p.sign_k = b.sign_k;
    p.store_all = (p.strategy != S3_STRATEGY_NONE) || (b.flags & S3_BATCH_STORE_ALL_ROWS);
    p.flags = b.flags;
    p.W = (int)((g.num_nodes + 31) / 32);
    p.arena = b.arena;
    p.arena_words = b.arena_words;
    p.off = b.off;
    p.cnt = b.cnt;
    p.counters = reinterpret_cast<unsigned long long*>(b.counters);
    p.caps = batch_caps(b) ? 1 : 0;
    p.cap_ratio = b.ratio_per_hop;
    p.cap_max = b.max_nodes_per_hop;
    p.cap_seed = b.cap_seed;
    p.out_link = b.out_link;
    p.mirror = batch_pairing(b) ? b.mirror : nullptr;
    p.link_base = b.link_base;
    if (p.num_records == 0) return cudaSuccess;
    return b.flow == S3_FLOW_POS ? launch_front<2>(p, g, b, st, rc_out) : launch_front<1>(p, g, b, st, rc_out);
}

}  // namespace s3
