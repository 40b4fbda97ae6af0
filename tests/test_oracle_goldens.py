"""The oracle (oracle/s3grl_oracle.py) against the reference's own outputs committed under
tests/golden/ — this is what pins the oracle (the reference has no tests)."""
import numpy as np
import pytest

from golden_util import Case, assert_features_close, case_names
from oracle import s3grl_oracle as orc


@pytest.mark.parametrize('name', case_names('pos'))
def test_pos_flows_match_reference(name):
    c = Case(name)
    out = orc.pos_precompute(c.links, c.num_hops, c.A, c.X, c.K, c.strategy, keep_graphs=True)
    assert np.array_equal(out['row_ptr'], c.row_ptr)
    for k in range(c.K + 1):
        assert_features_close(out['xs'][k], c.xs[k], what=f'{name} x{k}')
    # index parity: node lists, hop labels, induced + masked edge lists, bit-exact
    for i, g in enumerate(out['graphs']):
        a, b = c.node_ptr[i], c.node_ptr[i + 1]
        assert np.array_equal(g['nodes'], c.nodes[a:b]), f'{name} link {i} nodes'
        assert np.array_equal(g['hops'], c.hops[a:b]), f'{name} link {i} hops'
        rows = np.repeat(np.arange(g['nodes'].size), np.diff(g['lrowptr']))
        e = np.stack([g['nodes'][rows], g['nodes'][g['lcol']]], 1)
        assert np.array_equal(e, c.edges[c.edge_ptr[i]:c.edge_ptr[i + 1]]), f'{name} link {i} edges'
        assert np.array_equal(g['nodes'][g['sel']], c.row_gid[c.row_ptr[i]:c.row_ptr[i + 1]])


@pytest.mark.parametrize('name', case_names('pos_caps'))
def test_per_hop_caps_match_the_reference_with_its_sampler_ranked(name):
    """ratio_per_hop / max_nodes_per_hop: the fixture is the reference's own capped BFS + PoS with only `random.sample`
    replaced by the rank rule (oracle/make_goldens.caps_case); node lists, hop labels, edges bit-exact, operators 1e-5."""
    c = Case(name)
    assert 'random.sample' in c.reference_repair
    out = orc.pos_precompute(c.links, c.num_hops, c.A, c.X, c.K, None, keep_graphs=True, caps=c.caps)
    assert np.array_equal(out['row_ptr'], c.row_ptr)
    uncapped = 0
    for i, g in enumerate(out['graphs']):
        a, b = c.node_ptr[i], c.node_ptr[i + 1]
        assert np.array_equal(g['nodes'], c.nodes[a:b]) and np.array_equal(g['hops'], c.hops[a:b]), f'{name} link {i}'
        rows = np.repeat(np.arange(g['nodes'].size), np.diff(g['lrowptr']))
        e = np.stack([g['nodes'][rows], g['nodes'][g['lcol']]], 1)
        assert np.array_equal(e, c.edges[c.edge_ptr[i]:c.edge_ptr[i + 1]]), f'{name} link {i} edges'
        uncapped += orc.k_hop_subgraph(int(c.links[0, i]), int(c.links[1, i]), c.num_hops, c.A)[0].size
    assert uncapped > 2 * int(c.node_ptr[-1]), "the caps must bite on this fixture"
    for k in range(c.K + 1):
        assert_features_close(out['xs'][k], c.xs[k], what=f'{name} x{k}')


@pytest.mark.parametrize('name', case_names('sop'))
def test_sop_matches_reference(name):
    c = Case(name)
    out = orc.sop_precompute(c.links, c.A, c.X, c.K)
    assert np.array_equal(out['row_ptr'], c.row_ptr)
    for k in range(c.K + 1):
        assert_features_close(out['xs'][k], c.xs[k], what=f'{name} x{k}')


@pytest.mark.parametrize('name', case_names('hybrid'))
def test_hybrid_matches_reference(name):
    """the reference's own dispatcher (utils.py:454-480) with sign_type='hybrid' — PoS x, x1..xK
    then the SoP operators x2..xK as x{K+1}..x{2K-1}."""
    c = Case(name)
    out = orc.hybrid_precompute(c.links, c.num_hops, c.A, c.X, c.K)
    assert len(out['xs']) == 2 * c.K == len(c.xs)
    assert np.array_equal(out['row_ptr'], c.row_ptr)
    for k in range(2 * c.K):
        assert_features_close(out['xs'][k], c.xs[k], what=f'{name} x{k}')


@pytest.mark.parametrize('name', [n for n in case_names('pos') if 'union' in n])
def test_union_literal_rows_of_the_repaired_reference(name):
    """The `union` fixtures hold the reference's LITERAL rows (its label-column literal repaired): src and dst selected a
    second time.  The oracle's compat_explicit_zero restates that selection as [0, 1, 0, 1, CCN rows]; the fixtures list the
    rows beyond the first two in ascending global id."""
    c = Case(name)
    assert c.reference_repair.startswith('tuned_SIGN.py:243')
    out = orc.pos_precompute(c.links, c.num_hops, c.A, c.X, c.K, 'union', keep_graphs=True, compat_explicit_zero=True)
    assert np.array_equal(out['row_ptr'], c.literal_row_ptr)
    for i, g in enumerate(out['graphs']):
        a, b = int(c.literal_row_ptr[i]), int(c.literal_row_ptr[i + 1])
        gid = g['nodes'][g['sel']]
        order = np.concatenate([[0, 1], 2 + np.argsort(gid[2:], kind='stable')])
        assert np.array_equal(gid[order], c.literal_row_gid[a:b])
        for k in range(c.K + 1):
            assert_features_close(g['xs'][k][order], c.literal_xs[k][a:b], what=f'{name} link {i} x{k}')


def test_union_rule_is_superset_of_intersection():
    c = Case('usair_posplus')
    inter = orc.pos_precompute(c.links[:, :20], c.num_hops, c.A, c.X, c.K, 'intersection', keep_graphs=True)
    union = orc.pos_precompute(c.links[:, :20], c.num_hops, c.A, c.X, c.K, 'union', keep_graphs=True)
    for gi, gu in zip(inter['graphs'], union['graphs']):
        assert set(gi['sel']) <= set(gu['sel'])
        assert list(gu['sel'][:2]) == [0, 1] and np.all(np.diff(gu['sel'][2:]) > 0)
        # rows 0,1 do not depend on the strategy
        for k in range(c.K + 1):
            assert np.array_equal(gi['xs'][k][:2], gu['xs'][k][:2])


def test_src_equals_dst_rejected():
    c = Case('tiny_pos_h1')
    with pytest.raises(ValueError):
        orc.k_hop_subgraph(3, 3, 2, c.A)


def test_float64_oracle_agrees():
    c = Case('cora_pos')
    a = orc.pos_precompute(c.links[:, :30], c.num_hops, c.A, c.X, c.K)
    b = orc.pos_precompute(c.links[:, :30], c.num_hops, c.A, c.X, c.K, dtype=np.float64)
    for k in range(c.K + 1):
        assert_features_close(a['xs'][k], b['xs'][k], what=f'x{k}')


def test_non_optimised_flow_is_an_independent_cross_check():
    """Flow 9 (SIGN on every row, SpMM chain) and flow 6 (SpGEMM powers then 2 rows) are two
    algebraically different routes; rows 0,1 must agree."""
    c = Case('cora_pos')
    for i in range(12):
        src, dst = int(c.links[0, i]), int(c.links[1, i])
        full = orc.sign_all_rows(src, dst, c.num_hops, c.A, c.X, c.K)
        opt = orc.pos_link(src, dst, c.num_hops, c.A, c.X, c.K)
        for k in range(c.K + 1):
            assert_features_close(full[k][:2], opt['xs'][k], what=f'link {i} x{k}')
            a, b = c.row_ptr[i], c.row_ptr[i + 1]
            assert_features_close(full[k][:2], c.xs[k][a:b], what=f'link {i} x{k} vs reference golden')


@pytest.mark.parametrize('name', case_names('full'))
def test_non_optimised_flow_matches_reference(name):
    """the reference's optimize_sign=False PoS branch (k_hop_subgraph ->
    construct_pyg_graph(node_label) -> TunedSIGN) for every labeling trick it accepts."""
    c = Case(name)
    out = orc.full_precompute(c.links, c.num_hops, c.A, c.X, c.K, c.node_label)
    assert np.array_equal(out['row_ptr'], c.row_ptr)
    assert np.array_equal(out['node_id'], c.node_id)
    assert np.array_equal(out['xs'][0], c.xs[0]), 'x = [z | X_sub] is an exact copy'
    for k in range(1, c.K + 1):
        assert_features_close(out['xs'][k], c.xs[k], what=f'{name} x{k}')


def test_scaled_random_walk_flow_matches_reference():
    """ScaLed: given the same walk sets, the oracle and the reference's own
    k_hop_subgraph (random-walk branch) + get_PoS_prepped_ds agree."""
    c = Case('cora_scaled')
    out = orc.scaled_pos_precompute(c.links, c.sets, c.A, c.X, c.K)
    assert np.array_equal(out['row_ptr'], c.row_ptr)
    for k in range(c.K + 1):
        assert_features_close(out['xs'][k], c.xs[k], what=f'scaled x{k}')


def test_oracle_random_walk_sets_are_walks():
    c = Case('cora_scaled')
    starts = c.links.reshape(-1)[:40]
    sets = orc.random_walk_sets(c.A, starts, 3, 20, seed=1)
    import scipy.sparse.csgraph as csg
    for s, nodes in sets.items():
        assert s in nodes and nodes.size <= 61 and np.all(np.diff(nodes) > 0)
        dist = csg.shortest_path(c.A, unweighted=True, indices=s)
        assert np.all(dist[nodes] <= 3)
