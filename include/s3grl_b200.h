/*
 * s3grl_b200.h — C ABI of the B200-native S3GRL precompute library (libs3grl_b200.so).
 *
 * The reference (venomouscyanide/S3GRL) has no FFI layer: its hot path is the Python
 * call  extract_enclosing_subgraphs(link_index, A, x, y, num_hops, ..., sign_kwargs,
 * powers_of_A, data)  (reference utils.py:446-449) and beneath it the three static methods
 * OptimizedSignOperations.get_PoS_prepped_ds / get_PoS_Plus_prepped_ds / get_SoP_prepped_ds
 * (reference tuned_SIGN.py:137, :192, :49).  This header is what a binding for that path
 * binds instead: plain pointers and sizes, no torch types, every pointer a DEVICE pointer
 * owned by the caller, every call asynchronous on the given cudaStream_t (passed as void*).
 * The library allocates nothing: scratch comes from a caller-supplied arena.
 *
 * One precompute call over a batch of B "records" is three stages (north_star kernels 1-3):
 *
 *   s3_extract  : per record, h-hop frontier expansion over the device-resident CSR, dedup,
 *                 canonical renumbering, induced + target-masked local CSR, row selection,
 *                 AND the diffusion weights of the record's first work item (the rows of the
 *                 two targets) — the "front" kernel.
 *                 replaces reference utils.py:33-85 (neighbors, k_hop_subgraph BFS branch),
 *                 the ssp.find at tuned_SIGN.py:153/:208, the CCN row selection
 *                 tuned_SIGN.py:228-238 and tuned_SIGN.py:155-175 for rows [0,1].
 *   s3_plan     : row_ptr / work-item lists from the per-record selected-row counts
 *                 (the slices PyG's collate would record, sgrl_link_pred.py:204).
 *   s3_diffuse  : per CCN work item (PoS Plus: up to 8 extra selected rows of a record),
 *                 S = D^-1/2 A_sub D^-1/2 and the K row vectors e_sel^T S^k
 *                 replaces tuned_SIGN.py:155-175 (normalise, SpGEMM powers, row select)
 *                 and, in SoP flow, sgrl_link_pred.py:161-178 + tuned_SIGN.py:60-86, :106-113.
 *   s3_gather   : per work item, x_k[sel] = (e_sel^T S^k) [label | X_sub] for k = 0..K written
 *                 straight into the K+1 row-stacked operator matrices the SIGNNet models
 *                 consume (tuned_SIGN.py:177-187, :94-133; layout models.py:372).
 *
 * A "record" is one target link (PoS / PoS Plus: two seeds, target edge masked, induced
 * degrees) or one link endpoint (SoP: one seed, K-hop ball, global degrees; record 2i is
 * the source of link i, 2i+1 its destination).
 *
 * Canonical order (bit-exact contract): local node ids are [seeds..., then ascending
 * (hop, global id)]; row j of the local CSR lists j's neighbours in ascending GLOBAL id;
 * extra selected rows ascend by local id.
 */
#ifndef S3GRL_B200_H
#define S3GRL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S3_VERSION 200

/* return codes (the reference raises Python exceptions; see INTEGRATION.md for the mapping) */
#define S3_OK 0
#define S3_ERR_INVALID_ARG 1     /* null pointer, negative size, unsupported K / hops        */
#define S3_ERR_UNSUPPORTED 2     /* graph too large for the selected extraction tier        */
#define S3_ERR_CUDA 3            /* a CUDA runtime call failed; see s3_last_cuda_error()     */
#define S3_ERR_NOT_IMPLEMENTED 4 /* unknown strategy / flow (reference: NotImplementedError) */
#define S3_ERR_WORKSPACE 5       /* arena smaller than s3_min_arena_words(): nothing was launched */

/* flows (reference sign_type) and row-selection strategies (reference k_node_set_strategy) */
#define S3_FLOW_POS 0
#define S3_FLOW_SOP 1
#define S3_STRATEGY_NONE 0          /* PoS: rows [0,1]              tuned_SIGN.py:173     */
#define S3_STRATEGY_INTERSECTION 1  /* PoS Plus, common neighbours  tuned_SIGN.py:232-233 */
#define S3_STRATEGY_UNION 2         /* PoS Plus, union minus {0,1}  tuned_SIGN.py:230-231 */

/* labeling-trick column of the non-optimised SIGN flow (reference node_label, utils.py:296-310) */
#define S3_LABEL_ZERO 0    /* any other value of node_label: zeros                 */
#define S3_LABEL_ZO 1      /* 'zo': 1 on the two targets                           */
#define S3_LABEL_HOP 2     /* 'hop': BFS hop of the node                           */
#define S3_LABEL_DRNL 3    /* 'drnl': double-radius node labeling, utils.py:211-236 */
#define S3_LABEL_DEGREE 4  /* 'degree': induced degree capped at 100               */

#define S3_MAX_HOPS 8
#define S3_MAX_PEERS 8 /* GPUs of one NVLink domain served by s3_gather_peers */
#define S3_MAX_K 7
#define S3_MAX_K_UNION 5 /* S3_STRATEGY_UNION: larger sign_k returns S3_ERR_NOT_IMPLEMENTED from every entry point */

/* per-record status written by s3_extract into cnt[S3_CNT_STATUS] */
#define S3_REC_OK 0
#define S3_REC_ARENA_OVERFLOW 1  /* arena too small: re-run the batch with a larger arena            */
#define S3_REC_BAD_LINK 2        /* node id out of range or src == dst               */
#define S3_REC_MIRROR 3          /* s3_pair_links found an earlier link over the same node pair: nothing is    */
                                 /* extracted; the earlier link's s3_gather writes this link's rows as well     */

/* per-record int64 offsets (in 4-byte words from the arena base), off[rec*S3_NOFF + i] */
/* Only rows j < cnt[S3_CNT_NSTORE] are stored (see s3_extract).                                */
/* The local CSR is padded: row j owns the slots lcol[rowptr[j] .. rowptr[j+1]), one per entry */
/* of the node's GLOBAL adjacency list in its order (rowptr = prefix sum of global degrees,    */
/* rowptr[n] = D); a slot holds the neighbour's local id, or -1 when the neighbour is outside  */
/* the subgraph or the entry is the masked target link. rowlen[j] counts the non-holes.        */
#define S3_OFF_NODES 0   /* int32 global ids, n                                   */
#define S3_OFF_ROWPTR 1  /* int32 row starts into lcol, n+1                       */
#define S3_OFF_ROWLEN 2  /* int32 induced + masked degree of every row, n         */
#define S3_OFF_LCOL 3    /* int32 local column ids, D slots of which m are used   */
#define S3_OFF_SEL 4     /* int32 extra selected local ids, s - num_seeds         */
#define S3_OFF_F32 5     /* float scratch of the record's work items              */
#define S3_NOFF 6

/* per-record int32 counts, cnt[rec*S3_NCNT + i] */
#define S3_CNT_N 0        /* subgraph nodes                                       */
#define S3_CNT_M 1        /* directed induced edges after masking, over the STORED rows */
#define S3_CNT_S 2        /* selected rows                                        */
#define S3_CNT_STATUS 3
#define S3_CNT_PARTNER 4  /* SoP: local id of the other endpoint in the ball, or -1 */
#define S3_CNT_HOP0 5     /* S3_CNT_HOP0 + l = number of nodes at hop l, l = 0..S3_MAX_HOPS */
#define S3_CNT_NSTORE 14  /* rows kept in the padded CSR: hops <= K-1, or all n with S3_BATCH_STORE_ALL_ROWS */
#define S3_CNT_CLASSPOS 15 /* arrival index of the record inside its size class           */
#define S3_NCNT 16

/* int64 counters[S3_NCTR]; the caller zeroes them before s3_extract */
#define S3_CTR_CURSOR 0    /* arena words requested so far (may exceed the capacity)      */
#define S3_CTR_ERRORS 1    /* number of records whose status != S3_REC_OK                 */
#define S3_CTR_ROWS 2      /* total selected rows   (written by s3_plan)                  */
#define S3_CTR_ITEMS 3     /* total CCN work items: 8 extra selected rows each (s3_plan)  */
#define S3_CTR_MAX_N 4     /* largest subgraph in the batch                               */
#define S3_CTR_SUM_N 5     /* sum of n  (roofline accounting: 4*F*sum_n feature bytes)    */
#define S3_CTR_SUM_D 6     /* sum over subgraph nodes of their global degree (4*D bytes)  */
#define S3_CTR_WORK 7      /* work-queue head of the persistent extraction CTAs           */
#define S3_CTR_SUM_N_ALL 8 /* sum of n over every LINK served (a record counts once per paired link too) */
#define S3_CTR_SUM_D_ALL 9 /* same for D: the SURVEY 8d per-link figures with pairing switched on        */
#define S3_CTR_MIRRORS 10  /* links served by another link's record (s3_pair_links)                      */
#define S3_CTR_SUM_READ 11 /* sorted tier: adjacency entries the method really streams (the two merged lists + */
                           /* the rows of low-degree nodes; hub rows are probed by binary search instead): the  */
                           /* index bytes of its roofline, in place of the 4*D of the SURVEY formula            */
#define S3_CTR_CHAIN_READS 12   /* s3_ccn_chain: row segments its levels read, summed over the records it served:   */
                               /* sum_k (induced edges of the rows of hop <= 1 + K - k) — times 4 * (F + 1) bytes = */
                               /* the shared-memory traffic that bounds the kernel (its roofline in bench.py)       */
#define S3_CTR_CHAIN_RECORDS 13 /* records s3_ccn_chain served (the others took the CCN work items)                */
#define S3_CTR_CHAIN_N 14      /* sum of n over those records                                                       */
#define S3_CTR_CLASS0 16   /* + c: records whose n has floor(log2 n) == c (size classes)  */
#define S3_NCTR 48

/* Device-resident graph: CSR of the training graph (both directions stored, columns
 * ascending and unique per row — what scipy's csr_matrix gives the reference at
 * sgrl_link_pred.py:111-114; stored values are not needed, tuned_SIGN.py:153 drops them)
 * and the feature matrix, rows padded to a multiple of 4 floats for 128-bit loads. */
typedef struct s3_graph {
    const int64_t* indptr;  /* [num_nodes + 1]                              */
    const int32_t* indices; /* [indptr[num_nodes]]                          */
    const float* x;         /* [num_nodes, ldx] row-major, 16-byte aligned  */
    int64_t num_nodes;
    int64_t num_feat;       /* F                                            */
    int64_t ldx;            /* row stride in floats, multiple of 4, >= F    */
    int64_t num_edges;      /* indptr[num_nodes]; < 2^32 for the bitmap tier */
    int64_t max_degree;     /* largest row of the CSR (sizes the sorted tier's slabs) */
    /* Optional hub index of the sorted-set tier (all three 0 / NULL: not used). hub_id[v] = index of v among the
     * num_hubs highest-degree nodes, or -1; hub_bits is their num_hubs x ceil(num_hubs / 32) adjacency bit matrix
     * (s3_build_hub_bits). Whether two hubs are adjacent is then ONE bit probe instead of a binary search in an
     * adjacency list of 10^3..10^5 entries — the probe chains that bound the R-MAT configuration in round 1. */
    const int32_t* hub_id;    /* [num_nodes]                                  */
    const uint32_t* hub_bits; /* [num_hubs * ((num_hubs + 31) / 32)], zeroed by the caller before s3_build_hub_bits */
    int64_t num_hubs;
    /* Optional size proxy per node (NULL: not used), filled by s3_node_proxy: deg(v) + sum of its neighbours' degrees —
     * an estimate of how much of the graph a BFS from v touches. With s3_batch.front_order set, s3_extract hands the
     * records to its persistent CTAs in descending proxy(src) + proxy(dst) (longest first), which trims the tail of
     * the launch; results do not depend on it. */
    const int32_t* size_proxy; /* [num_nodes]                                 */
} s3_graph;

/* s3_batch.flags: keep every row of the induced adjacency (parity dumps). Without it rows of
 * hop == K are streamed once and never stored (PoS Plus always stores all rows). */
#define S3_BATCH_STORE_ALL_ROWS 1
/* Use the sorted-set extraction tier even when the bitmap tier would fit (tests). The sorted
 * tier serves graphs of any size but only PoS with num_hops == 1 and S3_STRATEGY_NONE. */
#define S3_BATCH_FORCE_SORTED_TIER 2

/* PoS Plus union: records whose subgraph fits the shared-memory placement of s3_ccn_chain get their CCN rows
 * from it and count no CCN work items in s3_plan; the others keep the work-item path (s3_plan_items +
 * s3_diffuse + s3_gather_ccn). Set the bit for s3_plan AND s3_ccn_chain of the same batch. */
#define S3_BATCH_CCN_CHAIN 4

/* The front kernel takes 3 instead of 5 CTA slots per SM, so that kernel 3 of the previous batch (on another stream,
 * waiting on NVLink stores in the multi-GPU exchange) stays resident beside it. */
#define S3_BATCH_SHARE_SMS 8

/* One batch of records and its scratch. */
typedef struct s3_batch {
    const int64_t* link_src; /* [num_links] device                                         */
    const int64_t* link_dst; /* [num_links] device                                         */
    int64_t num_links;
    int32_t flow;            /* S3_FLOW_*                                                   */
    int32_t strategy;        /* S3_STRATEGY_* (PoS flow only)                               */
    int32_t num_hops;        /* PoS: h.  SoP: ignored (the ball radius is sign_k)           */
    int32_t sign_k;          /* K operators beyond x                                        */
    int32_t flags;           /* S3_BATCH_* bits                                             */
    int32_t reserved;
    int32_t* arena;          /* scratch, 16-byte aligned                                    */
    int64_t arena_words;     /* capacity in 4-byte words                                    */
    int64_t* off;            /* [num_records * S3_NOFF]                                     */
    int32_t* cnt;            /* [num_records * S3_NCNT]                                     */
    int64_t* counters;       /* [S3_NCTR]                                                   */
    /* The next three may be NULL when every record has exactly num_seeds selected rows
     * (PoS with S3_STRATEGY_NONE, SoP): then record r is work item r and owns output rows
     * [r*num_seeds, (r+1)*num_seeds). */
    int64_t* row_ptr;        /* [num_records + 1] output-row offset of each record          */
    int64_t* item_ptr;       /* [num_records + 1] first CCN work item of each record        */
    int32_t* item_rec;       /* [total CCN items] record of each CCN work item              */
    /* Optional [num_records]: s3_extract fills it with the records in descending size class
     * (largest subgraphs first) and s3_gather schedules its CTAs in that order, which trims the
     * tail of the launch. Results do not depend on it. */
    int32_t* order;
    /* ScaLed random-walk subgraphs (reference utils.py:86-150): when walk_sets is not NULL the
     * subgraph of link i is {src, dst} ∪ set[link_src_set[i]] ∪ set[link_dst_set[i]] instead of an
     * h-hop ball; sets come from s3_walk_sets (or from the caller). PoS, S3_STRATEGY_NONE only;
     * num_hops is ignored (the reference passes 0). */
    const int32_t* walk_sets;     /* [num_sets, walk_cap] ascending unique node ids          */
    const int32_t* walk_counts;   /* [num_sets]                                              */
    const int64_t* link_src_set;  /* [num_links] row of walk_sets of every link's source     */
    const int64_t* link_dst_set;  /* [num_links] ... destination                             */
    int32_t walk_cap;
    int32_t reserved2;
    /* Link pairing (fixed-row PoS only; all three NULL / 0 otherwise). Training positives arrive as (u,v) AND
     * (v,u) (PyG train_test_split_edges): same enclosing subgraph, rows 0 and 1 swapped. `mirror`
     * is the chain table s3_pair_links built over the WHOLE link list of the call (indexed by global link
     * index); record r serves global link  out_link ? out_link[r] : link_base + r.  A record whose link is a
     * chain member is skipped by s3_extract (S3_REC_MIRROR); s3_gather of the chain's first link writes the
     * members' rows too, at row 2 * (member's global link index). With out_link the output rows of record r
     * are 2 * out_link[r] (+0, +1) instead of row_base + 2r: any partition of the link list over GPUs writes
     * straight into the layout of the whole list. */
    const int64_t* out_link;      /* [num_links] global link index of every record, or NULL   */
    const int64_t* mirror;        /* [links of the whole call] chain table, or NULL           */
    int64_t link_base;            /* global link index of record 0 when out_link is NULL      */
    /* Per-hop caps of the BFS (reference utils.py:66-70: sample_ratio, max_nodes_per_hop), PoS flows. A hop with
     * c new nodes keeps k = min(int(ratio_per_hop * c), max_nodes_per_hop) of them — the reference's counts. The
     * reference picks them with random.sample (no reproducible semantics; raises on Python >= 3.11); here they are
     * the k nodes with the smallest  fmix32(node XOR cap_seed)  (murmur3's 32-bit finaliser, a bijection: no
     * ties), so the subgraph is a function of (link, seed) alone and the oracle restates it exactly. As in the
     * reference the dropped nodes stay visited and do not return at a later hop. ratio_per_hop <= 0 or >= 1 and
     * max_nodes_per_hop <= 0 switch the respective cap off (a zeroed struct caps nothing). */
    double ratio_per_hop;
    int32_t max_nodes_per_hop;
    uint32_t cap_seed;
    /* Optional [num_records] scratch: with s3_graph.size_proxy, s3_extract first sorts the records by descending
     * size proxy into it (counting sort over 128 logarithmic classes) and its CTAs take them in that order. */
    int32_t* front_order;
} s3_batch;

int s3_version(void);
const char* s3_error_string(int code);
const char* s3_last_cuda_error(void);

/* records in a batch: num_links (PoS) or 2*num_links (SoP) */
int64_t s3_num_records(const s3_batch* b);
/* upper bound of work items s3_plan can produce is not known before extract; after
 * s3_plan + a stream sync the exact numbers are counters[S3_CTR_ROWS] / [S3_CTR_ITEMS]. */

/* smallest arena s3_extract accepts for this graph (per-CTA slabs live at its head); a useful
 * arena is much larger: roughly 14*n + D_store + 12*n*ceil(s/2) words per record. */
int64_t s3_min_arena_words(const s3_graph* g, const s3_batch* b);

/* 0: shared-memory bitmap tier, 1: sorted-set tier, -1: unsupported combination. */
int s3_extract_tier(const s3_graph* g, const s3_batch* b);

/* dynamic shared memory the bitmap extraction tier needs for this graph, or -1 if the
 * graph is too large for it (then S3_ERR_UNSUPPORTED from s3_extract). */
int64_t s3_extract_smem_bytes(int64_t num_nodes, int32_t radius);

int s3_extract(const s3_graph* g, const s3_batch* b, void* stream);
/* s3_plan: exclusive scans -> row_ptr, item_ptr, counters[S3_CTR_ROWS|ITEMS]. The caller then
 * synchronises the stream, reads the two totals, allocates item_rec and the outputs, and calls
 * s3_plan_items to fill item_rec. */
int s3_plan(const s3_batch* b, void* stream);
int s3_plan_items(const s3_batch* b, void* stream);
int s3_diffuse(const s3_graph* g, const s3_batch* b, int64_t num_items, void* stream);

/* s3_gather writes the rows of the records themselves (row 0 = src, row 1 = dst; SoP: one row per
 * record), one CTA per record; s3_gather_ccn writes the CCN rows of PoS Plus, one CTA per CCN
 * work item (s3_plan / s3_plan_items).
 * out: HOST array of sign_k+1 device pointers; out[k], k = 0..sign_k to the operator matrices, [*, ldo] row-major
 * float32; record r's rows land at row_base + row_ptr[r] .. ; column 0 is the label /
 * self-return column, columns 1..F the features (reference tuned_SIGN.py:177-187). */
int s3_gather(const s3_graph* g, const s3_batch* b, int64_t num_records,
              float* const* out, int64_t ldo, int64_t row_base, void* stream);
int s3_gather_ccn(const s3_graph* g, const s3_batch* b, int64_t num_items,
                  float* const* out, int64_t ldo, int64_t row_base, void* stream);

/* s3_ccn_chain: the CCN rows (rows 2.. of every record) of PoS Plus with S3_STRATEGY_UNION by a hop-limited
 * SpMM chain over the record's stored CSR instead of s3_plan_items + s3_diffuse + s3_gather_ccn (same
 * results within fp32 rounding, ~10x fewer FMAs when a record has ~20 CCN rows). Call after s3_plan (row_ptr)
 * and s3_gather on the same stream; one CTA per record. replaces tuned_SIGN.py:210-258 for the extra rows. */
int s3_ccn_chain(const s3_graph* g, const s3_batch* b, int64_t num_records, float* const* out, int64_t ldo,
                 int64_t row_base, void* stream);
/* s3_ccn_chain with a caller-owned pool for the large records: a record whose shared-memory placement would run at a
 * sub-chunk width CW <= pool_cw (4, 8, 16 or 32: s3_chain_shape & 255) keeps its CSR in shared memory and takes its two
 * [n][32] operator buffers from one of pool_slots slots of slot_floats floats each (a slot serves records with
 * 64 * n <= slot_floats; slot_floats a multiple of 32; pool 128-byte aligned), read and written through L2. pool_busy
 * [pool_slots] int32 must be zero on entry and is zero again when the kernels have finished. pool_slots == 0: s3_ccn_chain.
 * Records with buffers in global memory (pooled or in their own arena scratch) form level 1's input D^-1/2 [X | label] from
 * the feature matrix instead of storing it (same bits). A/B knobs read from the environment at every call, defaults in
 * brackets: S3GRL_CHAIN_POOL_SPLIT [2] pooled launches by CSR size, S3GRL_CHAIN_POOL_X [1] / S3GRL_CHAIN_SMEM_X [0] level 1
 * from the feature matrix for global- / shared-memory buffers, S3GRL_CHAIN_POLICY [0], S3GRL_CHAIN_SLABS [1]. */
int s3_ccn_chain_pooled(const s3_graph* g, const s3_batch* b, int64_t num_records, float* const* out, int64_t ldo,
                        int64_t row_base, float* pool, int32_t* pool_busy, int64_t slot_floats, int32_t pool_slots,
                        int32_t pool_cw, void* stream);
/* Placement of a record in s3_ccn_chain by its size (n nodes, m directed induced edges, n1 nodes of hop <= 1):
 * -1 if the record does not fit the chain's shared memory (it takes the CCN work items), else
 * columns_per_sub_chunk | cta_class << 8 with cta_class 0 / 1 / 2 = 256 / 512 / 1024 threads. Host-side helper
 * (no GPU): the same function the kernels and s3_plan evaluate. */
int s3_chain_shape(int64_t n, int64_t m, int64_t n1);

/* Non-optimised SIGN + SEAL flow (reference utils.py:497-520, tuned_SIGN.py:18-23,
 * i.e. PyG's SIGN transform on the whole subgraph): every subgraph node is an output row,
 * x = [z | X_sub], x_k = S x_{k-1}, S = D^-1/2 A_sub D^-1/2. The batch must have been extracted with
 * S3_BATCH_STORE_ALL_ROWS (PoS flow, S3_STRATEGY_NONE).
 * s3_plan_full: row_ptr[r] = exclusive scan of the subgraph sizes n, counters[S3_CTR_ROWS] = total.
 * s3_sign_full: out[k], k = 0..sign_k, [*, ldo] row-major; local node j of record r (canonical order)
 * lands on row row_base + row_ptr[r] + j. label = S3_LABEL_*. node_out (optional, int64 per output
 * row) receives the node's global id (the reference keeps it as data.node_id). */
int s3_plan_full(const s3_batch* b, void* stream);
int s3_sign_full(const s3_graph* g, const s3_batch* b, int64_t num_records, int32_t label, float* const* out,
                 int64_t ldo, int64_t row_base, int64_t* node_out, void* stream);

/* ScaLed: sorted node sets of rw_M uniform random walks of length rw_m from every
 * start node — replaces reference utils.py:425-443 (create_rw_cache). sets is [num_starts, cap]
 * with cap >= 1 + rw_M*rw_m (a set can not be larger); counts[i] receives the size of set i.
 * Counter-based RNG: set i depends only on (seed, starts[i]). */
int s3_walk_sets(const s3_graph* g, const int64_t* starts, int64_t num_starts, int32_t rw_m, int32_t rw_M,
                 uint64_t seed, int32_t cap, int32_t* sets, int32_t* counts, void* stream);

/* GPU-resident batch assembly: replaces PyG's DataLoader(..., follow_batch=[x1..xK]) /
 * Batch.from_data_list for the SIGN flows (reference sgrl_link_pred.py:1253-1269) and the feature-wise
 * concat at the top of SIGNNet.forward (models.py:372).
 * src: HOST array of num_ops device pointers to the collated operator matrices [R, ld_src] (num_cols
 * valid columns each); row_ptr [L+1]; link_idx [num_links] the links of the batch (or of a whole shuffled
 * epoch) in output order. dst [R_out, ld_dst], ld_dst >= num_ops*num_cols: row = [x | x1 | .. | xK] of one
 * selected row. out_row_ptr [num_links] = first output row of every listed link (exclusive scan of their
 * row counts), or NULL when every link has exactly rows_per_link rows. batch_vec (optional, [R_out]
 * int64) receives the position of the row's link in link_idx — PyG's data.batch / x{k}_batch. */
int s3_joint_rows(const float* const* src, int32_t num_ops, int64_t num_cols, int64_t ld_src, const int64_t* row_ptr,
                  const int64_t* link_idx, int64_t num_links, const int64_t* out_row_ptr, int32_t rows_per_link,
                  float* dst, int64_t ld_dst, int64_t* batch_vec, void* stream);

/* Fused scoring head of SIGNNet for the fixed-row flows, evaluation mode: replaces
 * models.py:370-376 (operator_diff = Linear -> ELU -> BatchNorm) + models.py:339-346 (center pooling
 * h[2i] * h[2i+1], k_heuristic = 0) on the loader's joint matrix. tcgen05 TF32 tensor-core GEMM with the
 * epilogue fused; 256 hidden channels (every paper config).
 * joint [rows, ld_joint] fp32 (rows even, ld_joint*4 a multiple of 16 bytes, 16-byte aligned base), weight
 * [256, ld_w] (torch Linear layout), bias / bn_scale / bn_shift [256] with bn_scale = gamma / sqrt(var + eps),
 * bn_shift = beta - mean * bn_scale. pool != 0: pooled [rows / 2, 256] = h[2i] * h[2i+1]; pool == 0: pooled [rows, 256]
 * = h itself (PoS Plus: the CCN pooling of models.py:347-367 is then done by the caller). */
int s3_sign_head(const float* joint, int64_t rows, int64_t kdim, int64_t ld_joint, const float* weight, int64_t ld_w,
                 int64_t hidden, const float* bias, const float* bn_scale, const float* bn_shift, float* pooled, int32_t pool,
                 void* stream);

/* Link pairing (see s3_batch.mirror): for every unordered node pair {u,v} that occurs more than once in the link
 * list (both directions of a training edge, or repeats), the link with the lowest index keeps the work and the
 * others become members of its chain:
 *   mirror[i] >= 0   link i is the first of its pair; mirror[i] is the first member of its chain
 *   mirror[i] == -1  link i has no other link over its pair (or is invalid: out of range / src == dst)
 *   mirror[i] <= -2  link i is a member: v = -2 - mirror[i], bit 0 of v = rows swapped relative to the first
 *                    link (opposite direction), (v >> 1) - 1 = next member or -1.
 * Which member follows which is scheduling dependent; the rows written are not. table: scratch of
 * 2 * s3_pair_table_slots(num_links) int64 words. Results are exact: the extraction / diffusion / gather of
 * (u,v) and (v,u) are bit-identical up to the row swap (the two seed rows are accumulated in ascending
 * global-id order). */
int64_t s3_pair_table_slots(int64_t num_links);
int s3_pair_links(const int64_t* link_src, const int64_t* link_dst, int64_t num_links, int64_t num_nodes,
                  int64_t* table, int64_t table_slots, int64_t* mirror, void* stream);

/* size_proxy[v] = min(INT32_MAX, deg(v) + sum over v's neighbours of their degree). out: [num_nodes] int32. */
int s3_node_proxy(const s3_graph* g, int32_t* out, void* stream);

/* Fills g->hub_bits from the CSR: bit (hub_id[u], hub_id[v]) for every stored entry (u, v) between two hubs.
 * hub_id and num_hubs come from the caller (any choice of hubs is valid: the index only short-cuts look-ups). */
int s3_build_hub_bits(const s3_graph* g, void* stream);

/* Negative sampling on the GPU (input side; replaces torch_geometric.utils.negative_sampling at
 * reference utils.py:645-648): candidate i is the ordered pair (u, v) drawn from a counter-based hash of (seed, i);
 * valid[i] = 1 when u != v, (u, v) is not a stored entry of g (columns ascending: binary search) and no candidate
 * j < i is the same pair. The caller keeps the first `count` valid candidates in index order (deterministic) and
 * asks for more candidates if there are too few. table: 2 * table_slots int64 words, table_slots a power of two
 * >= 2 * num_candidates (s3_pair_table_slots). g.x is not read. */
int s3_negative_candidates(const s3_graph* g, int64_t num_candidates, uint64_t seed, int64_t* table, int64_t table_slots,
                           int64_t* cand_src, int64_t* cand_dst, uint8_t* valid, void* stream);

/* Fused gather + all-gather over NVLink peer memory: as s3_gather for the fixed-row flows, but every
 * output row is stored into the operator matrices of ALL num_dst GPUs (this one included) instead of one local
 * copy followed by an NCCL all-gather. dst_bases: HOST array of num_dst (<= 8) device pointers, each the base of
 * one GPU's buffer mapped into this process (s3_peer_open); operator k of a buffer starts at base + k * op_stride
 * floats and is [*, ldo] row-major. Rows land at 2 * (global link index) (out_link / link_base / mirror as in
 * s3_batch). No flag is spun on: completion is the kernel boundary followed by the caller's barrier.
 * flags: S3_PEERS_LOCAL_X0 — operator 0 (x = [1 | X[node]], an exact copy of the features every GPU already holds)
 * is not stored at all; every GPU writes those rows itself for the whole link list with s3_fill_x0 (a quarter less
 * NVLink traffic at sign_k = 3).  S3_PEERS_LOCAL_MIRRORS — the rows of paired links (chain members of
 * s3_pair_links) are not stored either; after the exchange (barrier) every GPU copies them from the first link's
 * rows in its own memory with s3_fill_mirrors (23 % fewer rows over NVLink on the PubMed link list). */
#define S3_PEERS_LOCAL_X0 1
#define S3_PEERS_LOCAL_MIRRORS 2
int s3_gather_peers(const s3_graph* g, const s3_batch* b, int64_t num_records, float* const* dst_bases, int32_t num_dst,
                    int64_t op_stride, int64_t ldo, int32_t flags, void* stream);
/* Rows of every chain member of `mirror` (s3_pair_links over the whole list) copied from the chain's first link,
 * seed rows exchanged for the opposite direction: ops[k], k = first_op..num_ops-1, [2 * num_links, ldo] row-major. */
int s3_fill_mirrors(const int64_t* mirror, int64_t num_links, float* const* ops, int32_t first_op, int32_t num_ops,
                    int64_t num_cols, int64_t ldo, void* stream);
/* Link pairing for the flows with data-dependent row counts (PoS Plus; reference sgrl_link_pred.py:193-204 precomputes
 * (u,v) and (v,u), which differ only in the order of rows 0 and 1: the CCN rows of tuned_SIGN.py:228-238 are the same
 * set in the same ascending local order). The caller runs the path on the chain heads only (mirror[i] >= -1) and
 * places every link's rows from its head's record.
 * s3_pair_heads: head_code[i] = 2 * head(i) + swap(i) for every link of the table (swap: opposite direction).
 * s3_scatter_rows: record r of a piece (rows src_row_ptr[r] .. src_row_ptr[r+1] of src[k], k < num_ops, [*, ld_src])
 * is copied to rows dst_row_ptr[link] .. of dst[k] for link = link_idx ? link_idx[r] : link_base + r and, with a
 * table, for every member of that link's chain (rows 0 / 1 exchanged where swap is set). mirror == NULL: a plain
 * placement of the piece — PyG's collate of reference sgrl_link_pred.py:204 for one batch of records. */
int s3_pair_heads(const int64_t* mirror, int64_t num_links, int64_t* head_code, void* stream);
int s3_scatter_rows(float* const* src, int64_t ld_src, const int64_t* src_row_ptr, int64_t num_records,
                    const int64_t* link_idx, int64_t link_base, const int64_t* mirror, const int64_t* dst_row_ptr,
                    float* const* dst, int64_t ld_dst, int32_t num_ops, int64_t num_cols, void* stream);
/* s3_scatter_rows with the record's first lead_rows (0..2) rows written a second time in front of its rows: destination
 * rows [r_0 .. r_{lead-1} | r_0 .. r_{s-1}] (exchanged like rows 0 / 1 for an opposite-direction chain member), so
 * dst_row_ptr must count s + min(lead_rows, s) rows per link. lead_rows = 2 is the `compat_explicit_zero` layout of the
 * PoS Plus union: the reference (tuned_SIGN.py:230-231 on the masked subgraph of utils.py:78-79, whose explicit zeros
 * `neighbors` reports) selects src and dst a second time among the extra rows — SURVEY.md A.4. */
int s3_scatter_rows_lead(float* const* src, int64_t ld_src, const int64_t* src_row_ptr, int64_t num_records,
                         const int64_t* link_idx, int64_t link_base, const int64_t* mirror, const int64_t* dst_row_ptr,
                         float* const* dst, int64_t ld_dst, int32_t num_ops, int64_t num_cols, int32_t lead_rows, void* stream);
/* x (operator 0) of the fixed-row flows for a whole link list: out0 [2 * num_links, ldo], row 2i = [1 | X[src_i]],
 * row 2i+1 = [1 | X[dst_i]] (reference tuned_SIGN.py:181 / :119-124); rows of invalid links are left untouched. */
int s3_fill_x0(const s3_graph* g, const int64_t* link_src, const int64_t* link_dst, int64_t num_links, float* out0, int64_t ldo,
               void* stream);

/* Peer memory for s3_gather_peers: one cudaMalloc'ed buffer per GPU, exported as a 64-byte CUDA IPC handle that
 * the other ranks of the node open (peer access is enabled lazily by the open). Plain CUDA runtime IPC; the
 * handles travel through the caller's own channel (torch.distributed all_gather_object in parallel.py). */
#define S3_PEER_HANDLE_BYTES 64
int s3_peer_alloc(int64_t bytes, void** ptr);
int s3_peer_free(void* ptr);
int s3_peer_export(void* ptr, unsigned char* handle /* [S3_PEER_HANDLE_BYTES] */);
int s3_peer_open(const unsigned char* handle, void** ptr);
int s3_peer_close(void* ptr);

/* CCN segment pooling on the GPU (north_star kernel 3 "center / CCN pooling"; reference models.py:339-367,
 * SIGNNet._centre_pool_helper with k_heuristic: the center product h_src * h_dst next to global_mean_pool /
 * global_add_pool of a link's rows beyond its two targets). src [R, ld_src] row-major with num_cols valid columns,
 * rows grouped per link by row_ptr [num_links + 1] (row 0 = src, row 1 = dst, rows 2.. = CCN rows). One output row
 * per link, out [num_links, ld_out]:
 *   S3_POOL_OUT_CENTER  [ src * dst | pool(rows 2..) ]   2 * num_cols   (the input of link_pred_mlp, models.py:357-362)
 *   S3_POOL_OUT_ROWS    [ src | dst | pool(rows 2..) ]   3 * num_cols   (pooled output mode of the precompute path:
 *                                                                        pooling operator rows BEFORE the MLP is a
 *                                                                        different model than the reference's)
 * mode S3_POOL_SUM / S3_POOL_MEAN (k_pool_strategy 'sum' / 'mean'; an empty segment gives zeros, as torch_scatter
 * does; 'concat' needs exactly k_heuristic extra rows per link, which PoS Plus does not guarantee: not built).
 * Rows are added in ascending order: results are bit-reproducible. */
#define S3_POOL_SUM 1
#define S3_POOL_MEAN 2
#define S3_POOL_OUT_CENTER 0
#define S3_POOL_OUT_ROWS 1
int s3_segment_pool(const float* src, int64_t ld_src, int64_t num_cols, const int64_t* row_ptr, int64_t num_links, int32_t mode,
                    int32_t layout, float* out, int64_t ld_out, void* stream);

/* Measurement aids (bench.py roofline): the L2 -> SM read bandwidth (the grid streams an L2-sized buffer `iters`
 * times with 128-bit L1-bypassing loads: bytes * iters per launch) and the FP32 FMA issue rate
 * (ctas * 256 threads * iters * 128 FMAs per launch). Nothing on the product path calls them. */
int s3_probe_l2_read(const float* buf, int64_t bytes, int32_t iters, float* sink, int32_t ctas, void* stream);
int s3_probe_fma(int32_t iters, float* sink, int32_t ctas, void* stream);
int s3_probe_fma2(int32_t iters, float* sink, int32_t ctas, void* stream); /* the same FMAs as packed pairs (fma.rn.f32x2 / FFMA2) */

/* Optional dumps for parity checks: canonical global-id edge list of every record,
 * edges[e] = (global row, global col), e in [edge_ptr[r], edge_ptr[r+1]). */
int s3_dump_edges(const s3_batch* b, const int64_t* edge_ptr, int32_t* edges_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S3GRL_B200_H */
