"""LIVE check of the oracle against the UNMODIFIED reference (imported from /root/reference behind oracle/ref_stub
by oracle/ref_runner.py) on generated graphs — SURVEY.md §4 / §8c.  The committed goldens pin fixed cases; this test
draws new graphs and links every run (hypothesis, derandomised so that CI is reproducible) and compares node sets,
hop labels, induced + masked edge sets (bit-exact) and operators (1e-5 of max|ref|) for PoS, PoS Plus intersection,
PoS Plus union (the reference's ragged label-column literal at tuned_SIGN.py:243 repaired at run time, and — unmodified —
on the 3-node subgraphs it happens to accept), the per-hop caps (the reference's `random.sample` replaced by the rank rule),
SoP, hybrid and the non-optimised flow.  Build-container only: skipped where the reference checkout is absent
(the GPU box), and never imported by the product."""
import numpy as np
import pytest
import scipy.sparse as ssp
from hypothesis import HealthCheck, given, settings, strategies as st

from golden_util import assert_features_close, drop_duplicated_seed_rows
from oracle import ref_runner as rr
from oracle import s3grl_oracle as orc

pytestmark = pytest.mark.skipif(not rr.available(), reason="reference checkout not present (GPU box)")
SETTINGS = dict(max_examples=12, deadline=None, derandomize=True,
                suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])


@st.composite
def graph_and_links(draw, max_nodes=40):
    n = draw(st.integers(6, max_nodes))
    m = draw(st.integers(n // 2, 3 * n))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    e = rng.integers(0, n, (m, 2))
    e = e[e[:, 0] != e[:, 1]]
    e = np.unique(np.sort(e, axis=1), axis=0)
    row = np.concatenate([e[:, 0], e[:, 1]])
    col = np.concatenate([e[:, 1], e[:, 0]])
    A = ssp.csr_matrix((np.ones(row.size, dtype=np.int64), (row, col)), shape=(n, n))
    A.sort_indices()
    F = draw(st.integers(1, 6))
    X = rng.random((n, F), dtype=np.float32)
    L = draw(st.integers(1, 8))
    links = rng.integers(0, n, (2, 3 * L))
    links = links[:, links[0] != links[1]][:, :L]
    if e.shape[0] and draw(st.booleans()):          # a few true edges: the target-link mask matters
        pick = e[rng.integers(0, e.shape[0], 2)].T
        links = np.concatenate([links, pick, pick[::-1]], axis=1)
    return A, X, np.ascontiguousarray(links, dtype=np.int64)


@settings(**SETTINGS)
@given(graph_and_links(), st.integers(1, 3), st.integers(1, 4))
def test_k_hop_subgraph_sets(gl, num_hops, _k):
    A, X, links = gl
    for u, v in links.T.tolist():
        nodes, hops, edges = rr.ref_k_hop(u, v, num_hops, A)
        gn, gh, lrowptr, lcol = orc.k_hop_subgraph(u, v, num_hops, A)
        assert np.array_equal(gn, nodes) and np.array_equal(gh, hops)
        rows = np.repeat(np.arange(gn.size), np.diff(lrowptr))
        e = np.stack([gn[rows], gn[lcol]], 1) if lcol.size else np.zeros((0, 2), np.int64)
        assert np.array_equal(e, edges)


@settings(**SETTINGS)
@given(graph_and_links(), st.integers(1, 3), st.sampled_from([(0.5, None), (1.0, 3), (0.7, 4), (0.34, 1), (1.0, 1)]), st.integers(0, 5))
def test_per_hop_caps_with_the_reference_sampler_ranked(gl, num_hops, cap, seed):
    """ratio_per_hop / max_nodes_per_hop (utils.py:66-70).  The reference's `random.sample` on a set cannot be reproduced
    (and raises on Python >= 3.11); with ONLY that call replaced by the framework's rank rule (ref_runner.cap_sampler_ranked)
    the reference's capped BFS — its counts, their order, dropped nodes staying visited, the early exit — and the oracle's
    agree bit for bit on node sets, hop labels, edges, and within 1e-5 on the PoS operators."""
    A, X, links = gl
    ratio, max_nodes = cap
    caps = dict(ratio_per_hop=ratio, max_nodes_per_hop=max_nodes, cap_seed=seed)
    for u, v in links.T.tolist():
        nodes, hops, edges = rr.ref_k_hop(u, v, num_hops, A, ratio, max_nodes, seed)
        gn, gh, lrowptr, lcol = orc.k_hop_subgraph(u, v, num_hops, A, **caps)
        assert np.array_equal(gn, nodes) and np.array_equal(gh, hops)
        rows = np.repeat(np.arange(gn.size), np.diff(lrowptr))
        e = np.stack([gn[rows], gn[lcol]], 1) if lcol.size else np.zeros((0, 2), np.int64)
        assert np.array_equal(e, edges)
    ref = rr.ref_pos(links, num_hops, A, X, 3, None, caps=caps)
    out = orc.pos_precompute(links, num_hops, A, X, 3, None, caps=caps)
    for k in range(4):
        assert_features_close(out['xs'][k], ref['xs'][k], what=f'x{k}')


@settings(**SETTINGS)
@given(graph_and_links(), st.integers(1, 3), st.integers(1, 4), st.sampled_from([None, 'intersection']))
def test_pos_and_pos_plus(gl, num_hops, K, strategy):
    A, X, links = gl
    ref = rr.ref_pos(links, num_hops, A, X, K, strategy)
    out = orc.pos_precompute(links, num_hops, A, X, K, strategy, keep_graphs=True)
    assert np.array_equal(out['row_ptr'], ref['row_ptr'])
    gid = np.concatenate([g['nodes'][g['sel']] for g in out['graphs']])
    assert np.array_equal(gid, ref['row_gid'])
    for k in range(K + 1):
        assert_features_close(out['xs'][k], ref['xs'][k], what=f'x{k}')


@settings(**SETTINGS)
@given(graph_and_links(), st.integers(1, 3), st.integers(1, 4))
def test_pos_plus_union_against_the_repaired_reference(gl, num_hops, K):
    """BASELINE config 3.  The reference's union branch with its one broken literal repaired (ref_runner.union_typo_repaired);
    its rows are the framework's plus src and dst selected a second time (golden_util.drop_duplicated_seed_rows)."""
    A, X, links = gl
    ref = rr.ref_pos(links, num_hops, A, X, K, 'union', repair_union_typo=True)
    row_ptr, row_gid, xs = drop_duplicated_seed_rows(links, ref['row_ptr'], ref['row_gid'], ref['xs'])
    out = orc.pos_precompute(links, num_hops, A, X, K, 'union', keep_graphs=True)
    assert np.array_equal(out['row_ptr'], row_ptr)
    assert np.array_equal(np.concatenate([g['nodes'][g['sel']] for g in out['graphs']]), row_gid)
    for k in range(K + 1):
        assert_features_close(out['xs'][k], xs[k], what=f'x{k}')


def test_union_on_the_unmodified_reference():
    """What the reference does for `union` as it stands: ValueError (the ragged literal of tuned_SIGN.py:243) unless the
    subgraph has exactly 3 nodes, where `[[1]] + [[1]] + [[0] * 1]` happens to be rectangular — those links pin the oracle's
    union against the UNMODIFIED code."""
    # wedges a - c - b (h = 1: nodes {a, b, c}) next to a 4-cycle (4 nodes) and an isolated pair (2 nodes)
    e = np.array([(0, 2), (1, 2), (3, 5), (4, 5), (6, 7), (7, 8), (8, 9), (9, 6)])
    n = 12
    row, col = np.concatenate([e[:, 0], e[:, 1]]), np.concatenate([e[:, 1], e[:, 0]])
    A = ssp.csr_matrix((np.ones(row.size, dtype=np.int64), (row, col)), shape=(n, n))
    A.sort_indices()
    X = np.random.default_rng(0).random((n, 4), dtype=np.float32)
    wedges = np.array([[0, 3, 1, 2], [1, 4, 0, 0]])          # (0,1), (3,4), (1,0): 3 nodes; (2,0) at h=1: {2, 0, 1}
    ref = rr.ref_pos(wedges, 1, A, X, 3, 'union')             # no repair
    row_ptr, row_gid, xs = drop_duplicated_seed_rows(wedges, ref['row_ptr'], ref['row_gid'], ref['xs'])
    out = orc.pos_precompute(wedges, 1, A, X, 3, 'union', keep_graphs=True)
    assert np.array_equal(out['row_ptr'], row_ptr)
    assert np.array_equal(np.concatenate([g['nodes'][g['sel']] for g in out['graphs']]), row_gid)
    for k in range(4):
        assert_features_close(out['xs'][k], xs[k], what=f'x{k}')
    for bad in ([[6], [8]], [[10], [11]]):                    # 4 nodes / 2 nodes
        with pytest.raises(ValueError):
            rr.ref_pos(np.array(bad), 1, A, X, 3, 'union')


@settings(**SETTINGS)
@given(graph_and_links(), st.integers(1, 4))
def test_sop(gl, K):
    A, X, links = gl
    ref = rr.ref_sop(links, A, X, K)
    out = orc.sop_precompute(links, A, X, K)
    for k in range(K + 1):
        assert_features_close(out['xs'][k], ref['xs'][k], what=f'x{k}')


@settings(**SETTINGS)
@given(graph_and_links(), st.integers(1, 2), st.integers(2, 3))
def test_hybrid(gl, num_hops, K):
    A, X, links = gl
    ref = rr.ref_hybrid(links, num_hops, A, X, K)
    out = orc.hybrid_precompute(links, num_hops, A, X, K)
    assert len(out['xs']) == len(ref['xs']) == 2 * K
    for k in range(2 * K):
        assert_features_close(out['xs'][k], ref['xs'][k], what=f'x{k}')


@settings(**dict(SETTINGS, max_examples=8))
@given(graph_and_links(max_nodes=24), st.integers(1, 2), st.integers(1, 3), st.sampled_from(['zo', 'hop', 'drnl', 'degree']))
def test_non_optimised_flow(gl, num_hops, K, node_label):
    A, X, links = gl
    ref = rr.ref_full(links, num_hops, A, X, K, node_label)
    out = orc.full_precompute(links, num_hops, A, X, K, node_label)
    assert np.array_equal(out['row_ptr'], ref['row_ptr']) and np.array_equal(out['node_id'], ref['node_id'])
    for k in range(K + 1):
        assert_features_close(out['xs'][k], ref['xs'][k], what=f'{node_label} x{k}')
