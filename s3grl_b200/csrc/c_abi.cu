// extern "C" boundary of libs3grl_b200.so — see include/s3grl_b200.h for the contract and the
// reference interfaces each entry point replaces.
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace {
thread_local char g_cuda_err[256] = "";

int cuda_fail(cudaError_t e) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    return S3_ERR_CUDA;
}

int check_batch(const s3_batch* b) {
    if (!b || b->num_links < 0) return S3_ERR_INVALID_ARG;
    if (b->flow != S3_FLOW_POS && b->flow != S3_FLOW_SOP) return S3_ERR_NOT_IMPLEMENTED;
    if (b->flow == S3_FLOW_POS && (b->strategy < S3_STRATEGY_NONE || b->strategy > S3_STRATEGY_UNION))
        return S3_ERR_NOT_IMPLEMENTED;  // reference: NotImplementedError(f"check strat {strat}"), tuned_SIGN.py:235
    if (b->sign_k < 1 || b->sign_k > S3_MAX_K) return S3_ERR_INVALID_ARG;
    // the 8-row CCN work items of `union` are instantiated up to sign_k = 5 (gather_sc8_k*.cu)
    if (b->flow == S3_FLOW_POS && b->strategy == S3_STRATEGY_UNION && b->sign_k > S3_MAX_K_UNION) return S3_ERR_NOT_IMPLEMENTED;
    const int radius = b->flow == S3_FLOW_POS ? b->num_hops : b->sign_k;
    if (radius < 0 || radius > S3_MAX_HOPS) return S3_ERR_INVALID_ARG;
    if (b->num_links > 0 && (!b->link_src || !b->link_dst || !b->arena || !b->off || !b->cnt || !b->counters))
        return S3_ERR_INVALID_ARG;
    if (b->arena_words < 0 || (reinterpret_cast<uintptr_t>(b->arena) & 15)) return S3_ERR_INVALID_ARG;
    return S3_OK;
}

int check_graph(const s3_graph* g, bool need_x) {
    if (!g || !g->indptr || !g->indices || g->num_nodes <= 0 || g->num_nodes > INT32_MAX) return S3_ERR_INVALID_ARG;
    if (g->num_edges < 0 || g->max_degree < 0) return S3_ERR_INVALID_ARG;
    if (need_x) {
        if (!g->x || g->num_feat <= 0 || g->ldx < g->num_feat || (g->ldx & 3)) return S3_ERR_INVALID_ARG;
        if (reinterpret_cast<uintptr_t>(g->x) & 15) return S3_ERR_INVALID_ARG;
    }
    return S3_OK;
}
}  // namespace

extern "C" {

int s3_version(void) { return S3_VERSION; }

const char* s3_error_string(int code) {
    switch (code) {
        case S3_OK: return "ok";
        case S3_ERR_INVALID_ARG: return "invalid argument";
        case S3_ERR_UNSUPPORTED: return "graph too large for the bitmap tier and the request is not PoS / num_hops 1 / no CCN (sorted tier)";
        case S3_ERR_CUDA: return "CUDA error";
        case S3_ERR_NOT_IMPLEMENTED: return "unknown flow or strategy";
        case S3_ERR_WORKSPACE: return "arena smaller than s3_min_arena_words()";
        default: return "unknown error code";
    }
}

const char* s3_last_cuda_error(void) { return g_cuda_err; }

int64_t s3_num_records(const s3_batch* b) { return b->flow == S3_FLOW_SOP ? 2 * b->num_links : b->num_links; }

int s3_extract_tier(const s3_graph* g, const s3_batch* b) {
    if (!g || !b) return -1;
    if (b->walk_sets)  // ScaLed: always the sorted-set tier
        return (b->flow == S3_FLOW_POS && b->strategy == S3_STRATEGY_NONE && b->walk_counts && b->link_src_set &&
                b->link_dst_set && b->walk_cap > 0) ? 1 : -1;
    const int radius = b->flow == S3_FLOW_POS ? b->num_hops : b->sign_k;
    const bool bitmap_ok = s3_extract_smem_bytes(g->num_nodes, radius) >= 0 && g->num_edges < (int64_t(1) << 32);
    const bool sorted_ok = b->flow == S3_FLOW_POS && b->num_hops == 1;  // PoS and PoS Plus (round 2)
    if (sorted_ok && ((b->flags & S3_BATCH_FORCE_SORTED_TIER) || !bitmap_ok)) return 1;
    return bitmap_ok ? 0 : -1;
}

int64_t s3_min_arena_words(const s3_graph* g, const s3_batch* b) {
    if (s3_extract_tier(g, b) == 1)
        return 2 * ((2 * ((b->walk_sets ? (int64_t)b->walk_cap : g->max_degree) + 1) + 31) & ~int64_t(31));
    return 2 * ((4 * g->num_nodes + 1 + 31) & ~int64_t(31));
}

int64_t s3_extract_smem_bytes(int64_t num_nodes, int32_t radius) {
    if (num_nodes <= 0 || radius < 0 || radius > S3_MAX_HOPS) return -1;
    const int64_t W = (num_nodes + 31) / 32;
    const int64_t bytes = (1 + 2 * (int64_t)radius) * W * 4;
    return bytes <= s3::kMaxSmemBytes ? bytes : -1;
}

int s3_extract(const s3_graph* g, const s3_batch* b, void* stream) {
    int rc = check_graph(g, false);
    if (rc != S3_OK) return rc;
    rc = check_batch(b);
    if (rc != S3_OK) return rc;
    const int tier = s3_extract_tier(g, b);
    if (tier < 0) return S3_ERR_UNSUPPORTED;
    int launch_rc = S3_OK;
    cudaError_t e = tier == 1 ? s3::launch_extract_sorted(*g, *b, static_cast<cudaStream_t>(stream), &launch_rc)
                              : s3::launch_extract_bitmap(*g, *b, static_cast<cudaStream_t>(stream), &launch_rc);
    if (e != cudaSuccess) return cuda_fail(e);
    return launch_rc;
}

int s3_plan(const s3_batch* b, void* stream) {
    int rc = check_batch(b);
    if (rc != S3_OK) return rc;
    if (!b->row_ptr || !b->item_ptr) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_plan(*b, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_plan_items(const s3_batch* b, void* stream) {
    int rc = check_batch(b);
    if (rc != S3_OK) return rc;
    if (!b->item_ptr || !b->item_rec) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_plan_items(*b, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_diffuse(const s3_graph* g, const s3_batch* b, int64_t num_items, void* stream) {
    int rc = check_graph(g, false);
    if (rc != S3_OK) return rc;
    rc = check_batch(b);
    if (rc != S3_OK) return rc;
    if (num_items < 0) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_diffuse(*g, *b, num_items, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

static int gather_impl(const s3_graph* g, const s3_batch* b, int64_t num_items, float* const* out, int64_t ldo,
                       int64_t row_base, int ccn, void* stream) {
    int rc = check_graph(g, true);
    if (rc != S3_OK) return rc;
    rc = check_batch(b);
    if (rc != S3_OK) return rc;
    if (num_items < 0 || !out || ldo < g->num_feat + 1 || row_base < 0) return S3_ERR_INVALID_ARG;
    s3::OutPtrs o;
    memset(&o, 0, sizeof(o));
    for (int k = 0; k <= b->sign_k; ++k) {
        if (!out[k]) return S3_ERR_INVALID_ARG;
        o.p[k] = out[k];
    }
    cudaError_t e = s3::launch_gather(*g, *b, num_items, o, ldo, row_base, ccn != 0, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_gather(const s3_graph* g, const s3_batch* b, int64_t num_records, float* const* out, int64_t ldo, int64_t row_base,
              void* stream) {
    return gather_impl(g, b, num_records, out, ldo, row_base, 0, stream);
}

int s3_gather_ccn(const s3_graph* g, const s3_batch* b, int64_t num_items, float* const* out, int64_t ldo, int64_t row_base,
                  void* stream) {
    return gather_impl(g, b, num_items, out, ldo, row_base, 1, stream);
}

int s3_gather_peers(const s3_graph* g, const s3_batch* b, int64_t num_records, float* const* dst_bases, int32_t num_dst,
                    int64_t op_stride, int64_t ldo, int32_t flags, void* stream) {
    int rc = check_graph(g, true);
    if (rc != S3_OK) return rc;
    rc = check_batch(b);
    if (rc != S3_OK) return rc;
    if (b->strategy != S3_STRATEGY_NONE && b->flow == S3_FLOW_POS) return S3_ERR_NOT_IMPLEMENTED;  // fixed-row flows only
    if (num_records < 0 || !dst_bases || num_dst < 1 || num_dst > S3_MAX_PEERS || ldo < g->num_feat + 1 || op_stride < 0)
        return S3_ERR_INVALID_ARG;
    s3::PeerDst peers;
    peers.num_dst = num_dst;
    peers.op_stride = op_stride;
    peers.skip_op0 = (flags & S3_PEERS_LOCAL_X0) ? 1 : 0;
    peers.skip_chain = (flags & S3_PEERS_LOCAL_MIRRORS) ? 1 : 0;
    for (int d = 0; d < S3_MAX_PEERS; ++d) peers.base[d] = nullptr;
    for (int d = 0; d < num_dst; ++d) {
        if (!dst_bases[d]) return S3_ERR_INVALID_ARG;
        peers.base[d] = dst_bases[d];
    }
    s3::OutPtrs o;
    memset(&o, 0, sizeof(o));
    cudaError_t e = s3::launch_gather(*g, *b, num_records, o, ldo, b->link_base * 2 /* rows per link */, false,
                                      static_cast<cudaStream_t>(stream), &peers);
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_fill_x0(const s3_graph* g, const int64_t* link_src, const int64_t* link_dst, int64_t num_links, float* out0, int64_t ldo,
               void* stream) {
    int rc = check_graph(g, true);
    if (rc != S3_OK) return rc;
    if (num_links < 0 || ldo < g->num_feat + 1) return S3_ERR_INVALID_ARG;
    if (num_links > 0 && (!link_src || !link_dst || !out0)) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_fill_x0(*g, link_src, link_dst, num_links, out0, ldo, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_fill_mirrors(const int64_t* mirror, int64_t num_links, float* const* ops, int32_t first_op, int32_t num_ops, int64_t num_cols,
                    int64_t ldo, void* stream) {
    if (num_links < 0 || first_op < 0 || num_ops < first_op || num_ops > 2 * S3_MAX_K || num_cols < 1 || ldo < num_cols)
        return S3_ERR_INVALID_ARG;
    if (num_links == 0) return S3_OK;
    if (!mirror || !ops) return S3_ERR_INVALID_ARG;
    s3::OutPtrs o;
    memset(&o, 0, sizeof(o));
    for (int k = first_op; k < num_ops; ++k) {
        if (!ops[k] && num_links > 0) return S3_ERR_INVALID_ARG;
        o.p[k] = ops[k];
    }
    cudaError_t e = s3::launch_fill_mirrors(mirror, num_links, o, first_op, num_ops, num_cols, ldo, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int64_t s3_pair_table_slots(int64_t num_links) {
    int64_t slots = 64;
    while (slots < 2 * num_links) slots <<= 1;
    return slots;
}

int s3_pair_links(const int64_t* link_src, const int64_t* link_dst, int64_t num_links, int64_t num_nodes, int64_t* table,
                  int64_t table_slots, int64_t* mirror, void* stream) {
    if (num_links < 0 || num_nodes <= 0 || num_nodes > INT32_MAX) return S3_ERR_INVALID_ARG;
    if (num_links == 0) return S3_OK;
    if (!link_src || !link_dst || !table || !mirror) return S3_ERR_INVALID_ARG;
    if (table_slots < 2 * num_links || (table_slots & (table_slots - 1))) return S3_ERR_WORKSPACE;
    cudaError_t e = s3::launch_pair_links(link_src, link_dst, num_links, num_nodes, table, table_slots, mirror,
                                          static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_pair_heads(const int64_t* mirror, int64_t num_links, int64_t* head_code, void* stream) {
    if (num_links < 0) return S3_ERR_INVALID_ARG;
    if (num_links == 0) return S3_OK;
    if (!mirror || !head_code) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_pair_heads(mirror, num_links, head_code, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_scatter_rows_lead(float* const* src, int64_t ld_src, const int64_t* src_row_ptr, int64_t num_records, const int64_t* link_idx,
                         int64_t link_base, const int64_t* mirror, const int64_t* dst_row_ptr, float* const* dst, int64_t ld_dst,
                         int32_t num_ops, int64_t num_cols, int32_t lead_rows, void* stream) {
    if (num_records < 0 || num_ops < 1 || num_ops > 2 * S3_MAX_K || num_cols < 1 || num_cols > INT32_MAX || ld_src < num_cols ||
        ld_dst < num_cols || link_base < 0 || num_records > INT32_MAX || lead_rows < 0 || lead_rows > 2)
        return S3_ERR_INVALID_ARG;
    if (num_records == 0) return S3_OK;
    if (!src || !dst || !src_row_ptr || !dst_row_ptr) return S3_ERR_INVALID_ARG;
    s3::OutPtrs a, b;
    memset(&a, 0, sizeof(a));
    memset(&b, 0, sizeof(b));
    for (int k = 0; k < num_ops; ++k) {
        if (!src[k] || !dst[k]) return S3_ERR_INVALID_ARG;
        a.p[k] = src[k];
        b.p[k] = dst[k];
    }
    cudaError_t e = s3::launch_scatter_rows(a, ld_src, src_row_ptr, num_records, link_idx, link_base, mirror, dst_row_ptr, b, ld_dst,
                                            num_ops, num_cols, static_cast<cudaStream_t>(stream), lead_rows);
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_scatter_rows(float* const* src, int64_t ld_src, const int64_t* src_row_ptr, int64_t num_records, const int64_t* link_idx,
                    int64_t link_base, const int64_t* mirror, const int64_t* dst_row_ptr, float* const* dst, int64_t ld_dst,
                    int32_t num_ops, int64_t num_cols, void* stream) {
    return s3_scatter_rows_lead(src, ld_src, src_row_ptr, num_records, link_idx, link_base, mirror, dst_row_ptr, dst, ld_dst, num_ops,
                                num_cols, 0, stream);
}

int s3_segment_pool(const float* src, int64_t ld_src, int64_t num_cols, const int64_t* row_ptr, int64_t num_links, int32_t mode,
                    int32_t layout, float* out, int64_t ld_out, void* stream) {
    if (mode != S3_POOL_SUM && mode != S3_POOL_MEAN) return S3_ERR_NOT_IMPLEMENTED;  // reference: "Check pool strat" (models.py:333)
    if (layout != S3_POOL_OUT_CENTER && layout != S3_POOL_OUT_ROWS) return S3_ERR_INVALID_ARG;
    const int64_t width = (layout == S3_POOL_OUT_CENTER ? 2 : 3) * num_cols;
    if (num_links < 0 || num_cols < 1 || num_cols > INT32_MAX / 4 || ld_src < num_cols || ld_out < width) return S3_ERR_INVALID_ARG;
    if (num_links > 0 && (!src || !row_ptr || !out)) return S3_ERR_INVALID_ARG;
    if (num_links > INT32_MAX) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_segment_pool(src, ld_src, num_cols, row_ptr, num_links, mode, layout, out, ld_out,
                                            static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_probe_l2_read(const float* buf, int64_t bytes, int32_t iters, float* sink, int32_t ctas, void* stream) {
    if (!buf || !sink || bytes < 16 * 1024 || (bytes & 15) || iters < 1 || ctas < 1 || (reinterpret_cast<uintptr_t>(buf) & 15))
        return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_probe_l2_read(buf, bytes, iters, sink, ctas, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_probe_fma(int32_t iters, float* sink, int32_t ctas, void* stream) {
    if (!sink || iters < 1 || ctas < 1) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_probe_fma(iters, sink, ctas, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_probe_fma2(int32_t iters, float* sink, int32_t ctas, void* stream) {
    if (!sink || iters < 1 || ctas < 1) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_probe_fma2(iters, sink, ctas, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_node_proxy(const s3_graph* g, int32_t* out, void* stream) {
    int rc = check_graph(g, false);
    if (rc != S3_OK) return rc;
    if (!out) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_node_proxy(*g, out, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_build_hub_bits(const s3_graph* g, void* stream) {
    int rc = check_graph(g, false);
    if (rc != S3_OK) return rc;
    if (!g->hub_id || !g->hub_bits || g->num_hubs <= 0 || g->num_hubs > (int64_t(1) << 20)) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_build_hub_bits(*g, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_negative_candidates(const s3_graph* g, int64_t num_candidates, uint64_t seed, int64_t* table, int64_t table_slots,
                           int64_t* cand_src, int64_t* cand_dst, uint8_t* valid, void* stream) {
    int rc = check_graph(g, false);
    if (rc != S3_OK) return rc;
    if (num_candidates < 0) return S3_ERR_INVALID_ARG;
    if (num_candidates == 0) return S3_OK;
    if (!table || !cand_src || !cand_dst || !valid) return S3_ERR_INVALID_ARG;
    if (table_slots < 2 * num_candidates || (table_slots & (table_slots - 1))) return S3_ERR_WORKSPACE;
    cudaError_t e = s3::launch_negative_candidates(*g, num_candidates, seed, table, table_slots, cand_src, cand_dst, valid,
                                                   static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_peer_alloc(int64_t bytes, void** ptr) {
    if (bytes <= 0 || !ptr) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::peer_alloc(bytes, ptr);
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_peer_free(void* ptr) {
    if (!ptr) return S3_OK;
    cudaError_t e = s3::peer_free(ptr);
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_peer_export(void* ptr, unsigned char* handle) {
    if (!ptr || !handle) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::peer_export(ptr, handle);
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_peer_open(const unsigned char* handle, void** ptr) {
    if (!handle || !ptr) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::peer_open(handle, ptr);
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_peer_close(void* ptr) {
    if (!ptr) return S3_OK;
    cudaError_t e = s3::peer_close(ptr);
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_ccn_chain_pooled(const s3_graph* g, const s3_batch* b, int64_t num_records, float* const* out, int64_t ldo, int64_t row_base,
                        float* pool, int32_t* pool_busy, int64_t slot_floats, int32_t pool_slots, int32_t pool_cw, void* stream) {
    int rc = check_graph(g, true);
    if (rc != S3_OK) return rc;
    rc = check_batch(b);
    if (rc != S3_OK) return rc;
    if (b->flow != S3_FLOW_POS || b->strategy != S3_STRATEGY_UNION) return S3_ERR_NOT_IMPLEMENTED;
    if (num_records < 0 || !out || !b->row_ptr || ldo < g->num_feat + 1 || row_base < 0) return S3_ERR_INVALID_ARG;
    if (pool_slots < 0 || (pool_slots > 0 && (!pool || !pool_busy || slot_floats < 64 || (slot_floats & 31) || pool_cw < 4 || pool_cw > 32)))
        return S3_ERR_INVALID_ARG;
    s3::OutPtrs o;
    memset(&o, 0, sizeof(o));
    for (int k = 0; k <= b->sign_k; ++k) {
        if (!out[k]) return S3_ERR_INVALID_ARG;
        o.p[k] = out[k];
    }
    cudaError_t e = s3::launch_ccn_chain(*g, *b, num_records, o, ldo, row_base, static_cast<cudaStream_t>(stream), pool, pool_busy,
                                         slot_floats, pool_slots, pool_cw);
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_ccn_chain(const s3_graph* g, const s3_batch* b, int64_t num_records, float* const* out, int64_t ldo, int64_t row_base,
                 void* stream) {
    return s3_ccn_chain_pooled(g, b, num_records, out, ldo, row_base, nullptr, nullptr, 0, 0, 0, stream);
}

int s3_chain_shape(int64_t n, int64_t m, int64_t n1) {
    if (n < 2 || m < 0 || n1 < 2 || n1 > n) return -1;
    if (!s3::chain_eligible(S3_BATCH_CCN_CHAIN, S3_STRATEGY_UNION, n, m, n1)) return -1;
    return s3::chain_shape(n, m, n1, 0);
}

int s3_plan_full(const s3_batch* b, void* stream) {
    int rc = check_batch(b);
    if (rc != S3_OK) return rc;
    if (!b->row_ptr) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_plan_full(*b, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_sign_full(const s3_graph* g, const s3_batch* b, int64_t num_records, int32_t label, float* const* out, int64_t ldo,
                 int64_t row_base, int64_t* node_out, void* stream) {
    int rc = check_graph(g, true);
    if (rc != S3_OK) return rc;
    rc = check_batch(b);
    if (rc != S3_OK) return rc;
    if (label < S3_LABEL_ZERO || label > S3_LABEL_DEGREE) return S3_ERR_NOT_IMPLEMENTED;
    if (b->flow != S3_FLOW_POS || b->strategy != S3_STRATEGY_NONE || !(b->flags & S3_BATCH_STORE_ALL_ROWS)) return S3_ERR_INVALID_ARG;
    if (num_records < 0 || !out || !b->row_ptr || ldo < g->num_feat + 1 || row_base < 0) return S3_ERR_INVALID_ARG;
    s3::OutPtrs o;
    memset(&o, 0, sizeof(o));
    for (int k = 0; k <= b->sign_k; ++k) {
        if (!out[k]) return S3_ERR_INVALID_ARG;
        o.p[k] = out[k];
    }
    cudaError_t e = s3::launch_sign_full(*g, *b, num_records, label, o, ldo, row_base, node_out, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_walk_sets(const s3_graph* g, const int64_t* starts, int64_t num_starts, int32_t rw_m, int32_t rw_M, uint64_t seed,
                 int32_t cap, int32_t* sets, int32_t* counts, void* stream) {
    int rc = check_graph(g, false);
    if (rc != S3_OK) return rc;
    if (num_starts < 0 || rw_m < 0 || rw_M < 1 || (int64_t)rw_M * (rw_m + 1) > 8192 || cap < 1 + rw_M * rw_m) return S3_ERR_INVALID_ARG;
    if (num_starts > 0 && (!starts || !sets || !counts)) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_walk_sets(*g, starts, num_starts, rw_m, rw_M, seed, cap, sets, counts,
                                         static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_joint_rows(const float* const* src, int32_t num_ops, int64_t num_cols, int64_t ld_src, const int64_t* row_ptr,
                  const int64_t* link_idx, int64_t num_links, const int64_t* out_row_ptr, int32_t rows_per_link, float* dst,
                  int64_t ld_dst, int64_t* batch_vec, void* stream) {
    if (!src || num_ops < 1 || num_ops > 2 * S3_MAX_K || num_cols < 1 || num_cols > INT32_MAX / 64 || ld_src < num_cols)
        return S3_ERR_INVALID_ARG;
    if (num_links < 0 || ld_dst < (int64_t)num_ops * num_cols || (!out_row_ptr && rows_per_link < 1)) return S3_ERR_INVALID_ARG;
    if (num_links > 0 && (!row_ptr || !link_idx || !dst)) return S3_ERR_INVALID_ARG;
    s3::OutPtrs o;
    memset(&o, 0, sizeof(o));
    for (int k = 0; k < num_ops; ++k) {
        if (!src[k] && num_links > 0) return S3_ERR_INVALID_ARG;
        o.p[k] = const_cast<float*>(src[k]);
    }
    cudaError_t e = s3::launch_joint_rows(o, num_ops, num_cols, ld_src, row_ptr, link_idx, num_links, out_row_ptr,
                                          rows_per_link, dst, ld_dst, batch_vec, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_sign_head(const float* joint, int64_t rows, int64_t kdim, int64_t ld_joint, const float* weight, int64_t ld_w,
                 int64_t hidden, const float* bias, const float* bn_scale, const float* bn_shift, float* pooled, int32_t pool,
                 void* stream) {
    if (hidden != 256) return S3_ERR_UNSUPPORTED;
    if (rows < 0 || (pool && (rows & 1)) || kdim < 1 || ld_joint < kdim || ld_w < kdim || (ld_joint & 3) || (ld_w & 3)) return S3_ERR_INVALID_ARG;
    if (rows > 0 && (!joint || !weight || !bias || !bn_scale || !bn_shift || !pooled)) return S3_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(joint) & 15) || (reinterpret_cast<uintptr_t>(weight) & 15) ||
        (reinterpret_cast<uintptr_t>(pooled) & 15) || rows > INT32_MAX)
        return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_sign_head(joint, rows, kdim, ld_joint, weight, ld_w, bias, bn_scale, bn_shift, pooled, pool ? 1 : 0,
                                         static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

int s3_dump_edges(const s3_batch* b, const int64_t* edge_ptr, int32_t* edges_out, void* stream) {
    int rc = check_batch(b);
    if (rc != S3_OK) return rc;
    if (!edge_ptr || !edges_out) return S3_ERR_INVALID_ARG;
    cudaError_t e = s3::launch_dump_edges(*b, edge_ptr, edges_out, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? S3_OK : cuda_fail(e);
}

}  // extern "C"
