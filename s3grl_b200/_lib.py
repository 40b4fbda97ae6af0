"""ctypes binding of libs3grl_b200.so (include/s3grl_b200.h).  There is no fallback: if the
library cannot be loaded, or a call fails, the product path raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('S3GRL_LIB') or os.path.join(HERE, 'lib', 'libs3grl_b200.so')     # S3GRL_LIB: A/B builds

# constants mirrored from the header
S3_OK, S3_ERR_INVALID_ARG, S3_ERR_UNSUPPORTED, S3_ERR_CUDA, S3_ERR_NOT_IMPLEMENTED, S3_ERR_WORKSPACE = 0, 1, 2, 3, 4, 5
FLOW_POS, FLOW_SOP = 0, 1
STRATEGY_NONE, STRATEGY_INTERSECTION, STRATEGY_UNION = 0, 1, 2
MAX_HOPS, MAX_K, MAX_K_UNION, MAX_PEERS, PEER_HANDLE_BYTES = 8, 7, 5, 8, 64
VERSION = 200
PEERS_LOCAL_X0, PEERS_LOCAL_MIRRORS = 1, 2
POOL_SUM, POOL_MEAN, POOL_OUT_CENTER, POOL_OUT_ROWS = 1, 2, 0, 1
LABEL_ZERO, LABEL_ZO, LABEL_HOP, LABEL_DRNL, LABEL_DEGREE = 0, 1, 2, 3, 4
REC_OK, REC_ARENA_OVERFLOW, REC_BAD_LINK, REC_MIRROR = 0, 1, 2, 3
OFF_NODES, OFF_ROWPTR, OFF_ROWLEN, OFF_LCOL, OFF_SEL, OFF_F32, NOFF = 0, 1, 2, 3, 4, 5, 6
CNT_N, CNT_M, CNT_S, CNT_STATUS, CNT_PARTNER, CNT_HOP0, CNT_NSTORE, NCNT = 0, 1, 2, 3, 4, 5, 14, 16
BATCH_STORE_ALL_ROWS, BATCH_FORCE_SORTED_TIER, BATCH_CCN_CHAIN, BATCH_SHARE_SMS = 1, 2, 4, 8
CTR_CURSOR, CTR_ERRORS, CTR_ROWS, CTR_ITEMS, CTR_MAX_N, CTR_SUM_N, CTR_SUM_D, CTR_WORK, NCTR = 0, 1, 2, 3, 4, 5, 6, 7, 48
CTR_SUM_N_ALL, CTR_SUM_D_ALL, CTR_MIRRORS, CTR_SUM_READ = 8, 9, 10, 11
CTR_CHAIN_READS, CTR_CHAIN_RECORDS, CTR_CHAIN_N = 12, 13, 14

EXPORTS = ['s3_version', 's3_error_string', 's3_last_cuda_error', 's3_num_records', 's3_extract_smem_bytes',
           's3_min_arena_words', 's3_extract_tier',
           's3_extract', 's3_plan', 's3_plan_items', 's3_diffuse', 's3_gather', 's3_gather_ccn', 's3_ccn_chain', 's3_ccn_chain_pooled', 's3_chain_shape', 's3_plan_full', 's3_sign_full', 's3_joint_rows', 's3_sign_head', 's3_walk_sets', 's3_dump_edges',
           's3_pair_table_slots', 's3_pair_links', 's3_pair_heads', 's3_scatter_rows', 's3_scatter_rows_lead', 's3_gather_peers', 's3_fill_x0', 's3_fill_mirrors', 's3_peer_alloc', 's3_peer_free', 's3_peer_export',
           's3_peer_open', 's3_peer_close', 's3_probe_l2_read', 's3_probe_fma', 's3_probe_fma2', 's3_segment_pool', 's3_negative_candidates', 's3_build_hub_bits', 's3_node_proxy']


class Graph(C.Structure):
    _fields_ = [('indptr', C.c_void_p), ('indices', C.c_void_p), ('x', C.c_void_p),
                ('num_nodes', C.c_int64), ('num_feat', C.c_int64), ('ldx', C.c_int64), ('num_edges', C.c_int64),
                ('max_degree', C.c_int64), ('hub_id', C.c_void_p), ('hub_bits', C.c_void_p), ('num_hubs', C.c_int64),
                ('size_proxy', C.c_void_p)]


class Batch(C.Structure):
    _fields_ = [('link_src', C.c_void_p), ('link_dst', C.c_void_p), ('num_links', C.c_int64),
                ('flow', C.c_int32), ('strategy', C.c_int32), ('num_hops', C.c_int32), ('sign_k', C.c_int32),
                ('flags', C.c_int32), ('reserved', C.c_int32),
                ('arena', C.c_void_p), ('arena_words', C.c_int64),
                ('off', C.c_void_p), ('cnt', C.c_void_p), ('counters', C.c_void_p),
                ('row_ptr', C.c_void_p), ('item_ptr', C.c_void_p), ('item_rec', C.c_void_p), ('order', C.c_void_p),
                ('walk_sets', C.c_void_p), ('walk_counts', C.c_void_p), ('link_src_set', C.c_void_p),
                ('link_dst_set', C.c_void_p), ('walk_cap', C.c_int32), ('reserved2', C.c_int32),
                ('out_link', C.c_void_p), ('mirror', C.c_void_p), ('link_base', C.c_int64),
                ('ratio_per_hop', C.c_double), ('max_nodes_per_hop', C.c_int32), ('cap_seed', C.c_uint32),
                ('front_order', C.c_void_p)]


class S3Error(RuntimeError):
    def __init__(self, code, where, detail=''):
        self.code = code
        super().__init__(f"{where}: {detail}" if detail else where)


_lib = None


def lib():
    """Load the shared library once. Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m s3grl_b200.build` "
                "(the precompute path has no CPU or PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        L.s3_version.restype = C.c_int
        L.s3_error_string.restype = C.c_char_p
        L.s3_error_string.argtypes = [C.c_int]
        L.s3_last_cuda_error.restype = C.c_char_p
        L.s3_num_records.restype = C.c_int64
        L.s3_num_records.argtypes = [C.POINTER(Batch)]
        L.s3_min_arena_words.restype = C.c_int64
        L.s3_min_arena_words.argtypes = [C.POINTER(Graph), C.POINTER(Batch)]
        L.s3_extract_tier.restype = C.c_int
        L.s3_extract_tier.argtypes = [C.POINTER(Graph), C.POINTER(Batch)]
        L.s3_extract_smem_bytes.restype = C.c_int64
        L.s3_extract_smem_bytes.argtypes = [C.c_int64, C.c_int32]
        L.s3_extract.argtypes = [C.POINTER(Graph), C.POINTER(Batch), C.c_void_p]
        L.s3_plan.argtypes = [C.POINTER(Batch), C.c_void_p]
        L.s3_plan_items.argtypes = [C.POINTER(Batch), C.c_void_p]
        L.s3_diffuse.argtypes = [C.POINTER(Graph), C.POINTER(Batch), C.c_int64, C.c_void_p]
        L.s3_gather.argtypes = [C.POINTER(Graph), C.POINTER(Batch), C.c_int64, C.POINTER(C.c_void_p),
                                C.c_int64, C.c_int64, C.c_void_p]
        L.s3_gather_ccn.argtypes = L.s3_gather.argtypes
        L.s3_ccn_chain.argtypes = L.s3_gather.argtypes
        L.s3_ccn_chain_pooled.argtypes = L.s3_gather.argtypes[:-1] + [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
        L.s3_chain_shape.argtypes = [C.c_int64, C.c_int64, C.c_int64]
        L.s3_chain_shape.restype = C.c_int
        L.s3_joint_rows.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.s3_sign_head.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.s3_plan_full.argtypes = [C.POINTER(Batch), C.c_void_p]
        L.s3_sign_full.argtypes = [C.POINTER(Graph), C.POINTER(Batch), C.c_int64, C.c_int32, C.POINTER(C.c_void_p),
                                   C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.s3_walk_sets.argtypes = [C.POINTER(Graph), C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_uint64, C.c_int32,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
        L.s3_dump_edges.argtypes = [C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p]
        L.s3_pair_table_slots.restype = C.c_int64
        L.s3_pair_table_slots.argtypes = [C.c_int64]
        L.s3_pair_links.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.s3_pair_heads.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.s3_scatter_rows.argtypes = [C.POINTER(C.c_void_p), C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.POINTER(C.c_void_p), C.c_int64, C.c_int32, C.c_int64, C.c_void_p]
        L.s3_scatter_rows_lead.argtypes = [C.POINTER(C.c_void_p), C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.POINTER(C.c_void_p), C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_void_p]
        L.s3_gather_peers.argtypes = [C.POINTER(Graph), C.POINTER(Batch), C.c_int64, C.POINTER(C.c_void_p), C.c_int32,
                                      C.c_int64, C.c_int64, C.c_int32, C.c_void_p]
        L.s3_fill_mirrors.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                      C.c_void_p]
        L.s3_fill_x0.argtypes = [C.POINTER(Graph), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.s3_probe_l2_read.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        L.s3_segment_pool.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_int64, C.c_void_p]
        L.s3_negative_candidates.argtypes = [C.POINTER(Graph), C.c_int64, C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p]
        L.s3_build_hub_bits.argtypes = [C.POINTER(Graph), C.c_void_p]
        L.s3_node_proxy.argtypes = [C.POINTER(Graph), C.c_void_p, C.c_void_p]
        L.s3_probe_fma.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        L.s3_probe_fma2.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        L.s3_peer_alloc.argtypes = [C.c_int64, C.POINTER(C.c_void_p)]
        L.s3_peer_free.argtypes = [C.c_void_p]
        L.s3_peer_export.argtypes = [C.c_void_p, C.c_char_p]
        L.s3_peer_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.s3_peer_close.argtypes = [C.c_void_p]
        for fn in ('s3_extract', 's3_plan', 's3_plan_items', 's3_diffuse', 's3_gather', 's3_gather_ccn', 's3_ccn_chain', 's3_ccn_chain_pooled', 's3_plan_full', 's3_sign_full', 's3_joint_rows', 's3_sign_head', 's3_walk_sets',
                   's3_dump_edges', 's3_pair_links', 's3_pair_heads', 's3_scatter_rows', 's3_scatter_rows_lead', 's3_gather_peers', 's3_fill_x0', 's3_fill_mirrors', 's3_peer_alloc', 's3_peer_free', 's3_peer_export',
                   's3_peer_open', 's3_peer_close', 's3_probe_l2_read', 's3_probe_fma', 's3_probe_fma2', 's3_segment_pool', 's3_negative_candidates', 's3_build_hub_bits', 's3_node_proxy'):
            getattr(L, fn).restype = C.c_int
        _lib = L
    return _lib


def check(code, where):
    """Map a C return code to the exception the reference would raise at the same point."""
    if code == S3_OK:
        return
    L = lib()
    msg = L.s3_error_string(code).decode()
    if code == S3_ERR_NOT_IMPLEMENTED:
        raise NotImplementedError(f"{where}: {msg}")      # reference tuned_SIGN.py:235, utils.py:553
    if code == S3_ERR_CUDA:
        raise S3Error(code, where, f"{msg}: {L.s3_last_cuda_error().decode()}")
    raise S3Error(code, where, msg)
