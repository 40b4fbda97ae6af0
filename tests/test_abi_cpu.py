"""CPU-side checks of the drop-in boundary: the shared library loads, exports every function
include/s3grl_b200.h declares, constants agree with the header, and argument validation that
needs no GPU behaves (no compute calls here)."""
import ctypes
import os
import re

import pytest

from s3grl_b200 import _lib as L

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
HEADER = os.path.join(ROOT, 'include', 's3grl_b200.h')


@pytest.fixture(scope='module')
def lib():
    from s3grl_b200.build import build
    build()
    return L.lib()


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(s3_[a-z_0-9]+)\s*\(', src)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared_functions()
    assert set(names) == set(L.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n


def test_constants_match_header():
    src = open(HEADER).read()
    defs = dict(re.findall(r'#define\s+(S3_[A-Z_0-9]+)\s+(-?\d+)', src))
    pairs = {'S3_FLOW_POS': L.FLOW_POS, 'S3_FLOW_SOP': L.FLOW_SOP, 'S3_STRATEGY_UNION': L.STRATEGY_UNION,
             'S3_STRATEGY_INTERSECTION': L.STRATEGY_INTERSECTION, 'S3_MAX_HOPS': L.MAX_HOPS, 'S3_MAX_K': L.MAX_K,
             'S3_NOFF': L.NOFF, 'S3_NCNT': L.NCNT, 'S3_NCTR': L.NCTR, 'S3_OFF_F32': L.OFF_F32,
             'S3_CNT_HOP0': L.CNT_HOP0, 'S3_CNT_PARTNER': L.CNT_PARTNER, 'S3_CTR_SUM_D': L.CTR_SUM_D,
             'S3_CTR_ITEMS': L.CTR_ITEMS, 'S3_REC_BAD_LINK': L.REC_BAD_LINK, 'S3_ERR_NOT_IMPLEMENTED': L.S3_ERR_NOT_IMPLEMENTED,
             'S3_BATCH_STORE_ALL_ROWS': L.BATCH_STORE_ALL_ROWS, 'S3_BATCH_FORCE_SORTED_TIER': L.BATCH_FORCE_SORTED_TIER,
             'S3_BATCH_CCN_CHAIN': L.BATCH_CCN_CHAIN, 'S3_BATCH_SHARE_SMS': L.BATCH_SHARE_SMS, 'S3_LABEL_ZO': L.LABEL_ZO, 'S3_LABEL_HOP': L.LABEL_HOP,
             'S3_LABEL_DRNL': L.LABEL_DRNL, 'S3_LABEL_DEGREE': L.LABEL_DEGREE, 'S3_LABEL_ZERO': L.LABEL_ZERO,
             'S3_REC_MIRROR': L.REC_MIRROR, 'S3_CTR_SUM_N_ALL': L.CTR_SUM_N_ALL, 'S3_CTR_SUM_D_ALL': L.CTR_SUM_D_ALL,
             'S3_CTR_MIRRORS': L.CTR_MIRRORS, 'S3_CTR_SUM_READ': L.CTR_SUM_READ, 'S3_CTR_CHAIN_READS': L.CTR_CHAIN_READS,
             'S3_CTR_CHAIN_RECORDS': L.CTR_CHAIN_RECORDS, 'S3_CTR_CHAIN_N': L.CTR_CHAIN_N, 'S3_MAX_PEERS': L.MAX_PEERS, 'S3_MAX_K_UNION': L.MAX_K_UNION,
             'S3_PEER_HANDLE_BYTES': L.PEER_HANDLE_BYTES, 'S3_VERSION': L.VERSION}
    for k, v in pairs.items():
        assert int(defs[k]) == v, k
    assert ctypes.sizeof(L.Graph) == 96 and ctypes.sizeof(L.Batch) == 208


def test_version_and_error_strings(lib):
    assert lib.s3_version() == L.VERSION == 200
    assert lib.s3_error_string(0) == b'ok'
    assert b'strategy' in lib.s3_error_string(L.S3_ERR_NOT_IMPLEMENTED)


def test_smem_sizing(lib):
    # PubMed, h=3: (1 + 2*3) * ceil(19717/32) * 4 bytes
    assert lib.s3_extract_smem_bytes(19717, 3) == 7 * 617 * 4
    assert lib.s3_extract_smem_bytes(10_000_000, 1) == -1     # needs the hash tier
    assert lib.s3_extract_smem_bytes(100, 9) == -1


def test_argument_validation_without_gpu(lib):
    g = L.Graph(0, 0, 0, 10, 4, 4, 20, 5)
    b = L.Batch()
    assert lib.s3_extract(ctypes.byref(g), ctypes.byref(b), None) == L.S3_ERR_INVALID_ARG      # null graph arrays
    g = L.Graph(16, 16, 16, 10, 4, 4, 20, 5)
    b.flow, b.strategy, b.sign_k, b.num_hops = 7, 0, 3, 2
    assert lib.s3_extract(ctypes.byref(g), ctypes.byref(b), None) == L.S3_ERR_NOT_IMPLEMENTED  # unknown flow
    b.flow, b.strategy = L.FLOW_POS, 9
    assert lib.s3_extract(ctypes.byref(g), ctypes.byref(b), None) == L.S3_ERR_NOT_IMPLEMENTED  # unknown strategy
    b.strategy, b.sign_k = 0, 0
    assert lib.s3_extract(ctypes.byref(g), ctypes.byref(b), None) == L.S3_ERR_INVALID_ARG      # K < 1
    b.sign_k, b.num_hops = 3, 99
    assert lib.s3_extract(ctypes.byref(g), ctypes.byref(b), None) == L.S3_ERR_INVALID_ARG      # too many hops


def test_argument_validation_of_the_wider_entry_points(lib):
    """Invalid arguments are rejected before anything is launched (no GPU needed): loader, scoring head, non-optimised
    flow, union chain."""
    P = ctypes.c_void_p
    g = L.Graph(16, 16, 16, 10, 4, 4, 20, 5)
    b = L.Batch()
    b.flow, b.strategy, b.sign_k, b.num_hops = L.FLOW_POS, L.STRATEGY_NONE, 3, 2
    out = (ctypes.c_void_p * 4)(16, 16, 16, 16)
    # s3_sign_full: unknown label -> NotImplementedError at the Python level; batch without S3_BATCH_STORE_ALL_ROWS -> invalid
    assert lib.s3_sign_full(ctypes.byref(g), ctypes.byref(b), 0, 9, out, 5, 0, None, None) == L.S3_ERR_NOT_IMPLEMENTED
    assert lib.s3_sign_full(ctypes.byref(g), ctypes.byref(b), 0, L.LABEL_DRNL, out, 5, 0, None, None) == L.S3_ERR_INVALID_ARG
    assert lib.s3_plan_full(ctypes.byref(b), None) == L.S3_ERR_INVALID_ARG                      # no row_ptr
    # s3_ccn_chain serves the union strategy only
    assert lib.s3_ccn_chain(ctypes.byref(g), ctypes.byref(b), 0, out, 5, 0, None) == L.S3_ERR_NOT_IMPLEMENTED
    # s3_ccn_chain_pooled: the same rules, then the pool (slots of a multiple of 32 floats, flags, threshold 4..32)
    pooled = lambda *pool: lib.s3_ccn_chain_pooled(ctypes.byref(g), ctypes.byref(b), 0, out, 5, 0, *pool, None)      # noqa: E731
    assert pooled(P(128), P(16), 64 * 100, 4, 8) == L.S3_ERR_NOT_IMPLEMENTED
    b.strategy, b.row_ptr = L.STRATEGY_UNION, 16
    assert pooled(None, None, 0, 0, 0) == L.S3_OK                                   # no pool, no records: s3_ccn_chain
    assert pooled(P(128), P(16), 64 * 100, 4, 8) == L.S3_OK
    assert pooled(None, P(16), 64 * 100, 4, 8) == L.S3_ERR_INVALID_ARG              # slots without a pool
    assert pooled(P(128), None, 64 * 100, 4, 8) == L.S3_ERR_INVALID_ARG             # ... without flags
    assert pooled(P(128), P(16), 6401, 4, 8) == L.S3_ERR_INVALID_ARG                # slot size not a multiple of 32 floats
    assert pooled(P(128), P(16), 32, 4, 8) == L.S3_ERR_INVALID_ARG                  # no record fits a slot of 32 floats
    assert pooled(P(128), P(16), 64 * 100, 4, 2) == L.S3_ERR_INVALID_ARG            # threshold below the narrowest sub-chunk
    assert pooled(P(128), P(16), 64 * 100, 4, 64) == L.S3_ERR_INVALID_ARG
    assert pooled(P(128), P(16), 64 * 100, -1, 8) == L.S3_ERR_INVALID_ARG
    b.strategy, b.row_ptr = L.STRATEGY_NONE, None
    # s3_joint_rows: operator count, leading dimensions
    assert lib.s3_joint_rows(out, 0, 5, 5, P(16), P(16), 1, None, 2, P(16), 20, None, None) == L.S3_ERR_INVALID_ARG
    assert lib.s3_joint_rows(out, 4, 5, 4, P(16), P(16), 1, None, 2, P(16), 20, None, None) == L.S3_ERR_INVALID_ARG   # ld_src < cols
    assert lib.s3_joint_rows(out, 4, 5, 5, P(16), P(16), 1, None, 2, P(16), 19, None, None) == L.S3_ERR_INVALID_ARG   # ld_dst < 4*5
    assert lib.s3_joint_rows(out, 4, 5, 5, P(16), P(16), 1, None, 0, P(16), 20, None, None) == L.S3_ERR_INVALID_ARG   # no row counts
    assert lib.s3_joint_rows(out, 4, 5, 5, None, None, 0, None, 2, None, 20, None, None) == L.S3_OK                    # empty: nothing to do
    # s3_sign_head: hidden size, parity of rows when pooling, 16-byte row strides (TMA)
    head = lambda rows, kd, ld, ldw, hidden, pool: lib.s3_sign_head(P(16), rows, kd, ld, P(16), ldw, hidden, P(16), P(16), P(16),  # noqa: E731
                                                                     P(16), pool, None)
    assert head(4, 8, 8, 8, 128, 1) == L.S3_ERR_UNSUPPORTED
    assert head(3, 8, 8, 8, 256, 1) == L.S3_ERR_INVALID_ARG
    assert head(4, 8, 6, 8, 256, 1) == L.S3_ERR_INVALID_ARG
    assert head(4, 6, 6, 8, 256, 1) == L.S3_ERR_INVALID_ARG
    assert head(0, 8, 8, 8, 256, 1) == L.S3_OK


def test_argument_validation_of_the_round2_entry_points(lib):
    """Link pairing, the peer exchange, pooling, negative sampling, the graph-level helpers: invalid arguments come back
    as error codes before anything is launched (no GPU needed)."""
    P = ctypes.c_void_p
    g = L.Graph(16, 16, 16, 10, 4, 4, 20, 5)
    b = L.Batch()
    b.flow, b.strategy, b.sign_k, b.num_hops = L.FLOW_POS, L.STRATEGY_NONE, 3, 2
    # s3_pair_links: table must hold >= 2 * L slots, a power of two
    assert lib.s3_pair_table_slots(0) == 64 and lib.s3_pair_table_slots(100) == 256 and lib.s3_pair_table_slots(164000) == 524288
    assert lib.s3_pair_links(P(16), P(16), 100, 50, P(16), 128, P(16), None) == L.S3_ERR_WORKSPACE
    assert lib.s3_pair_links(P(16), P(16), 100, 50, P(16), 300, P(16), None) == L.S3_ERR_WORKSPACE       # not a power of two
    assert lib.s3_pair_links(None, P(16), 100, 50, P(16), 256, P(16), None) == L.S3_ERR_INVALID_ARG
    assert lib.s3_pair_links(None, None, 0, 50, None, 0, None, None) == L.S3_OK                               # empty list
    # s3_gather_peers: fixed-row flows only, 1..8 destinations
    dst = (ctypes.c_void_p * 9)(*([16] * 9))
    assert lib.s3_gather_peers(ctypes.byref(g), ctypes.byref(b), 0, dst, 9, 100, 5, 0, None) == L.S3_ERR_INVALID_ARG
    assert lib.s3_gather_peers(ctypes.byref(g), ctypes.byref(b), 0, dst, 0, 100, 5, 0, None) == L.S3_ERR_INVALID_ARG
    assert lib.s3_gather_peers(ctypes.byref(g), ctypes.byref(b), 0, dst, 2, 100, 4, 0, None) == L.S3_ERR_INVALID_ARG   # ldo < F + 1
    b.strategy = L.STRATEGY_UNION
    assert lib.s3_gather_peers(ctypes.byref(g), ctypes.byref(b), 0, dst, 2, 100, 5, 0, None) == L.S3_ERR_NOT_IMPLEMENTED
    b.sign_k = L.MAX_K_UNION + 1          # union is built for sign_k <= 5: rejected by every entry point, up front
    assert lib.s3_extract(ctypes.byref(g), ctypes.byref(b), None) == L.S3_ERR_NOT_IMPLEMENTED
    b.strategy, b.sign_k = L.STRATEGY_NONE, 3
    # s3_fill_x0 / s3_fill_mirrors
    assert lib.s3_fill_x0(ctypes.byref(g), P(16), P(16), 4, P(16), 4, None) == L.S3_ERR_INVALID_ARG            # ldo < F + 1
    assert lib.s3_fill_x0(ctypes.byref(g), None, None, 0, None, 5, None) == L.S3_OK
    ops = (ctypes.c_void_p * 4)(16, 16, 16, 16)
    assert lib.s3_fill_mirrors(P(16), 4, ops, 5, 4, 5, 5, None) == L.S3_ERR_INVALID_ARG                         # first_op > num_ops
    assert lib.s3_fill_mirrors(None, 0, None, 1, 4, 5, 5, None) == L.S3_OK
    # s3_segment_pool: k_pool_strategy and layout
    assert lib.s3_segment_pool(P(16), 8, 8, P(16), 4, 3, L.POOL_OUT_CENTER, P(16), 16, None) == L.S3_ERR_NOT_IMPLEMENTED   # 'concat'
    assert lib.s3_segment_pool(P(16), 8, 8, P(16), 4, L.POOL_SUM, 7, P(16), 16, None) == L.S3_ERR_INVALID_ARG
    assert lib.s3_segment_pool(P(16), 8, 8, P(16), 4, L.POOL_SUM, L.POOL_OUT_ROWS, P(16), 16, None) == L.S3_ERR_INVALID_ARG  # ld_out < 3 * cols
    assert lib.s3_segment_pool(None, 8, 8, None, 0, L.POOL_MEAN, L.POOL_OUT_CENTER, None, 16, None) == L.S3_OK
    # s3_negative_candidates / s3_build_hub_bits / s3_node_proxy / probes / peer memory
    assert lib.s3_negative_candidates(ctypes.byref(g), 100, 1, P(16), 128, P(16), P(16), P(16), None) == L.S3_ERR_WORKSPACE
    assert lib.s3_negative_candidates(ctypes.byref(g), 0, 1, None, 0, None, None, None, None) == L.S3_OK
    assert lib.s3_build_hub_bits(ctypes.byref(g), None) == L.S3_ERR_INVALID_ARG                                 # no hub index attached
    assert lib.s3_node_proxy(ctypes.byref(g), None, None) == L.S3_ERR_INVALID_ARG
    assert lib.s3_probe_l2_read(None, 1 << 20, 1, P(16), 8, None) == L.S3_ERR_INVALID_ARG
    assert lib.s3_probe_fma(0, P(16), 8, None) == L.S3_ERR_INVALID_ARG
    assert lib.s3_peer_alloc(0, ctypes.byref(ctypes.c_void_p())) == L.S3_ERR_INVALID_ARG
    assert lib.s3_peer_export(None, ctypes.create_string_buffer(64)) == L.S3_ERR_INVALID_ARG
    assert lib.s3_peer_free(None) == L.S3_OK and lib.s3_peer_close(None) == L.S3_OK
    # pairing of the flows with data-dependent row counts (csrc/expand.cu)
    arr = (ctypes.c_void_p * 4)(16, 16, 16, 16)
    assert lib.s3_pair_heads(None, 0, None, None) == L.S3_OK and lib.s3_pair_heads(None, 3, P(16), None) == L.S3_ERR_INVALID_ARG
    assert lib.s3_pair_heads(P(16), -1, P(16), None) == L.S3_ERR_INVALID_ARG
    assert lib.s3_scatter_rows(arr, 5, P(16), 0, None, 0, None, P(16), arr, 5, 4, 5, None) == L.S3_OK       # nothing to place
    assert lib.s3_scatter_rows(arr, 4, P(16), 2, None, 0, None, P(16), arr, 5, 4, 5, None) == L.S3_ERR_INVALID_ARG   # ld < cols
    assert lib.s3_scatter_rows(arr, 5, None, 2, None, 0, None, P(16), arr, 5, 4, 5, None) == L.S3_ERR_INVALID_ARG    # no row_ptr
    assert lib.s3_scatter_rows(arr, 5, P(16), 2, None, 0, None, P(16), arr, 5, 0, 5, None) == L.S3_ERR_INVALID_ARG   # no operators
    assert lib.s3_scatter_rows(arr, 5, P(16), 2, None, -1, None, P(16), arr, 5, 4, 5, None) == L.S3_ERR_INVALID_ARG
    assert lib.s3_scatter_rows_lead(arr, 5, P(16), 0, None, 0, None, P(16), arr, 5, 4, 5, 2, None) == L.S3_OK
    assert lib.s3_scatter_rows_lead(arr, 5, P(16), 2, None, 0, None, P(16), arr, 5, 4, 5, 3, None) == L.S3_ERR_INVALID_ARG   # lead_rows > 2
    assert lib.s3_scatter_rows_lead(arr, 5, P(16), 2, None, 0, None, P(16), arr, 5, 4, 5, -1, None) == L.S3_ERR_INVALID_ARG


def test_chain_placement_by_record_size(lib):
    """s3_chain_shape (host-side, the function s3_plan and the chain kernels evaluate): the widest sub-chunk that fits,
    in the smallest CTA; wider never follows narrower as records grow; beyond ~4 000 nodes the work items keep the record."""
    shape = lambda n, m, n1: lib.s3_chain_shape(n, m, n1)
    assert shape(1, 0, 1) == -1 and shape(100, 400, 200) == -1                      # not a subgraph
    assert shape(120, 500, 30) == 32 | (0 << 8)                                      # 256-thread CTA, 4 per SM
    assert shape(300, 1500, 40) == 32 | (1 << 8) and shape(700, 3500, 50) == 32 | (2 << 8)
    assert shape(1200, 6000, 60) == 16 | (2 << 8) and shape(2500, 14000, 90) == 8 | (2 << 8)
    assert shape(3600, 20000, 120) == 4 | (2 << 8) and shape(9000, 50000, 300) == -1 and shape(70000, 70000, 5) == -1
    last = (32, 0)
    for n in range(20, 6000, 37):
        sh = shape(n, 5 * n, min(n, 40))
        if sh < 0:
            assert n > 3000
            last = (0, 3)
            continue
        cw, cls = sh & 255, sh >> 8
        assert cw in (32, 16, 8, 4) and cls in (0, 1, 2) and (cw < last[0] or (cw == last[0] and cls >= last[1])), (n, cw, cls, last)
        bytes_needed = 4 * (2 * n * cw + 3 * n + min(n, 40) + (5 * n + 2) // 2 + 8)
        assert bytes_needed <= (54, 110, 222)[cls] * 1024
        last = (cw, cls)


def test_chain_pool_slot_holds_every_record_the_chain_serves(lib):
    """engine.DeviceGraph.CHAIN_SLOT_FLOATS: a pool slot takes two [n][32] buffers of the largest record s3_chain_shape
    places in shared memory (those are the records the pooled route can take over)."""
    from s3grl_b200.engine import DeviceGraph
    assert DeviceGraph.CHAIN_SLOT_FLOATS % 32 == 0
    largest = max(n for n in range(2, 7000) if lib.s3_chain_shape(n, 0, 2) >= 0)     # no edges, two hop-<=1 nodes: the largest n
    assert 64 * largest <= DeviceGraph.CHAIN_SLOT_FLOATS, largest


def test_per_hop_cap_counts_match_the_reference_formula():
    """k = min(int(ratio * c), max) (reference utils.py:66-70) — the oracle's cap_fringe keeps exactly that many nodes,
    the same ones for the same seed, and a subset relation holds between caps."""
    import numpy as np
    from oracle import s3grl_oracle as orc
    fringe = np.arange(100, 1100, 3)
    for ratio, mx in ((1.0, None), (0.5, None), (1.0, 40), (0.3, 1000), (0.9, 7), (0.001, None)):
        kept = orc.cap_fringe(fringe, ratio, mx, seed=5)
        k = fringe.size if ratio >= 1.0 else int(ratio * fringe.size)
        k = min(k, mx) if mx is not None else k
        assert kept.size == k and np.all(np.diff(kept) > 0) and np.isin(kept, fringe).all()
        assert np.array_equal(kept, orc.cap_fringe(fringe, ratio, mx, seed=5))
    small, large = orc.cap_fringe(fringe, 1.0, 10, seed=1), orc.cap_fringe(fringe, 1.0, 50, seed=1)
    assert np.isin(small, large).all()              # the k smallest rank keys are nested
    assert not np.array_equal(orc.cap_fringe(fringe, 1.0, 10, seed=1), orc.cap_fringe(fringe, 1.0, 10, seed=2))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(L, '_lib', None)
    monkeypatch.setattr(L, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(RuntimeError, match='no CPU or PyTorch fallback'):
        L.lib()


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, 's3grl_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f


def test_library_is_sm100a_only_and_the_head_uses_tcgen05(lib, tmp_path):
    """Static evidence from the built binary (no GPU): every embedded cubin targets sm_100a, and the scoring head's SASS
    holds the 5th-generation tensor-core / TMA / TMEM instructions (B200_PROFILING.md: UTCHMMA = tcgen05.mma,
    UTMALDG = cp.async.bulk.tensor, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit)."""
    import shutil
    import subprocess
    if not shutil.which('cuobjdump'):
        pytest.skip('cuobjdump not on PATH')
    elfs = subprocess.run(['cuobjdump', '-lelf', L.LIB_PATH], capture_output=True, text=True).stdout
    names = re.findall(r'ELF file\s+\d+:\s+(\S+)', elfs)
    assert names and all(n.endswith('.sm_100a.cubin') for n in names), names
    head = [n for n in names if n.startswith('head.')]
    assert head, names
    subprocess.run(['cuobjdump', '-xelf', head[0], L.LIB_PATH], cwd=tmp_path, check=True, capture_output=True)
    sass = subprocess.run(['cuobjdump', '-sass', str(tmp_path / head[0])], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'UTMALDG.2D', 'LDTM', 'UTCBAR'):
        assert mnemonic in sass, mnemonic
