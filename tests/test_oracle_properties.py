"""Size-independent properties of the oracle (CPU): what must hold for ANY correct implementation of the path, used to
cross-check the restatement beyond the reference goldens.  The same properties are asserted for the CUDA path at full
size in tests/test_gpu_parity.py::test_pubmed_full_size_sample_and_properties."""
import numpy as np
import pytest
import scipy.sparse as ssp

from golden_util import Case, assert_features_close
from oracle import s3grl_oracle as orc


def _relabel(A, X, links, perm):
    """Rename node i -> perm[i]."""
    inv = np.argsort(perm)
    P = ssp.csr_matrix((np.ones(A.shape[0], dtype=A.dtype), (perm, np.arange(A.shape[0]))), shape=A.shape)
    A2 = (P @ A @ P.T).tocsr()
    A2.sort_indices()
    return A2, X[inv], perm[links]


@pytest.mark.parametrize('strategy', [None, 'intersection', 'union'])
def test_features_are_linear_in_x(strategy):
    """x_k[sel] = P_k[sel] [label | X]: columns 1.. are linear in X, column 0 does not depend on X at all."""
    c = Case('cora_posplus')
    links = c.links[:, :12]
    rng = np.random.default_rng(0)
    X1, X2 = c.X.astype(np.float64), rng.random(c.X.shape)
    a = orc.pos_precompute(links, c.num_hops, c.A, X1, c.K, strategy, dtype=np.float64)
    b = orc.pos_precompute(links, c.num_hops, c.A, X2, c.K, strategy, dtype=np.float64)
    s = orc.pos_precompute(links, c.num_hops, c.A, 2.0 * X1 - 3.0 * X2, c.K, strategy, dtype=np.float64)
    assert np.array_equal(a['row_ptr'], b['row_ptr'])
    for k in range(c.K + 1):
        np.testing.assert_allclose(s['xs'][k][:, 1:], 2.0 * a['xs'][k][:, 1:] - 3.0 * b['xs'][k][:, 1:], rtol=1e-10, atol=1e-12)
        np.testing.assert_array_equal(a['xs'][k][:, 0], b['xs'][k][:, 0])


def test_rows_of_the_targets_do_not_depend_on_node_names():
    """Relabelling the graph's nodes permutes canonical positions beyond the seeds but leaves rows 0, 1 of every
    operator unchanged (sums over the same neighbourhoods; only the fp32 summation order may differ)."""
    c = Case('cora_pos')
    links = c.links[:, :16]
    perm = np.random.default_rng(3).permutation(c.N)
    A2, X2, links2 = _relabel(c.A, c.X, links, perm)
    a = orc.pos_precompute(links, c.num_hops, c.A, c.X, c.K)
    b = orc.pos_precompute(links2, c.num_hops, A2, X2, c.K)
    for k in range(c.K + 1):
        assert_features_close(b['xs'][k], a['xs'][k], what=f'x{k}')
    a = orc.sop_precompute(links, c.A, c.X, c.K)
    b = orc.sop_precompute(links2, A2, X2, c.K)
    for k in range(c.K + 1):
        assert_features_close(b['xs'][k], a['xs'][k], what=f'SoP x{k}')


def test_reversed_link_swaps_the_two_rows():
    c = Case('cora_pos')
    links = c.links[:, :16]
    a = orc.pos_precompute(links, c.num_hops, c.A, c.X, c.K)
    b = orc.pos_precompute(links[::-1], c.num_hops, c.A, c.X, c.K)
    for k in range(c.K + 1):
        assert_features_close(b['xs'][k].reshape(-1, 2, c.X.shape[1] + 1)[:, ::-1], a['xs'][k].reshape(-1, 2, c.X.shape[1] + 1), what=f'x{k}')


def test_operator_zero_is_a_copy_and_label_column_of_x_is_one():
    c = Case('usair_posplus')
    out = orc.pos_precompute(c.links[:, :20], c.num_hops, c.A, c.X, c.K, 'intersection', keep_graphs=True)
    for i, g in enumerate(out['graphs']):
        a, b = out['row_ptr'][i], out['row_ptr'][i + 1]
        assert np.array_equal(out['xs'][0][a:b, 1:], c.X[g['nodes'][g['sel']]])
        assert out['xs'][0][a:b, 0].tolist() == [1.0, 1.0] + [0.0] * (b - a - 2)


def test_deeper_extraction_changes_nothing_within_k_steps():
    """A k-step walk from the targets stays inside the k-hop ball: with K = 2, h = 2 and h = 3 subgraphs differ only by
    nodes that can carry no weight... except through degrees of hop-2 nodes, which DO change — so x1 (touching hop <= 1
    rows whose degrees are complete at h = 2) is identical while x2 is not required to be."""
    c = Case('cora_pos')
    links = c.links[:, :10]
    a = orc.pos_precompute(links, 2, c.A, c.X, 1)
    b = orc.pos_precompute(links, 3, c.A, c.X, 1)
    for k in range(2):
        assert_features_close(b['xs'][k], a['xs'][k], what=f'x{k}')


def test_full_flow_label_columns_are_integers_and_drnl_is_symmetric_in_the_targets():
    c = Case('cora_full_drnl')
    a = orc.full_precompute(c.links[:, :6], c.num_hops, c.A, c.X, 1, 'drnl')
    b = orc.full_precompute(c.links[::-1, :6], c.num_hops, c.A, c.X, 1, 'drnl')
    za, zb = a['xs'][0][:, 0], b['xs'][0][:, 0]
    assert np.array_equal(za, np.round(za)) and za.min() >= 0
    # swapping src and dst swaps local rows 0 and 1 only; DRNL is symmetric in (src, dst)
    for i in range(6):
        ra, rb = slice(a['row_ptr'][i], a['row_ptr'][i + 1]), slice(b['row_ptr'][i], b['row_ptr'][i + 1])
        assert np.array_equal(a['node_id'][ra][2:], b['node_id'][rb][2:]) and np.array_equal(za[ra][2:], zb[rb][2:])


def test_sop_is_linear_in_x_and_its_first_column_is_the_return_probability():
    c = Case('usair_sop')
    links = c.links[:, :10]
    rng = np.random.default_rng(1)
    X1, X2 = c.X.astype(np.float64), rng.random(c.X.shape)
    powers = orc.sop_powers(c.A, c.K, np.float64)
    a = orc.sop_precompute(links, c.A, X1, c.K, np.float64, powers)
    b = orc.sop_precompute(links, c.A, X2, c.K, np.float64, powers)
    s = orc.sop_precompute(links, c.A, X1 + 0.5 * X2, c.K, np.float64, powers)
    for k in range(c.K + 1):
        np.testing.assert_allclose(s['xs'][k][:, 1:], a['xs'][k][:, 1:] + 0.5 * b['xs'][k][:, 1:], rtol=1e-10, atol=1e-12)
        np.testing.assert_array_equal(a['xs'][k][:, 0], b['xs'][k][:, 0])
    assert np.all(a['xs'][0][:, 0] == 1.0)                                   # tuned_SIGN.py:119-124
    for i in range(links.shape[1]):                                          # x_k[., 0] = A_hat^k[u, u] (tuned_SIGN.py:106-113)
        for k in range(1, c.K + 1):
            assert a['xs'][k][2 * i, 0] == powers[k - 1][links[0, i], links[0, i]]
            assert a['xs'][k][2 * i + 1, 0] == powers[k - 1][links[1, i], links[1, i]]
    assert np.all(a['xs'][1][:, 0] == 0.0)                                   # no self loops: no 1-step return


def test_hybrid_is_pos_followed_by_sop_operators_two_to_k():
    c = Case('cora_pos')
    links = c.links[:, :8]
    h = orc.hybrid_precompute(links, c.num_hops, c.A, c.X, c.K)
    p = orc.pos_precompute(links, c.num_hops, c.A, c.X, c.K)
    s = orc.sop_precompute(links, c.A, c.X, c.K)
    assert len(h['xs']) == 2 * c.K
    for k in range(c.K + 1):
        assert np.array_equal(h['xs'][k], p['xs'][k])
    for j, k in enumerate(range(2, c.K + 1)):
        assert np.array_equal(h['xs'][c.K + 1 + j], s['xs'][k])


def test_scaled_subgraph_is_a_subset_of_the_m_hop_ball():
    """Walks of length m never leave the m-hop ball of their start: the ScaLed node set of (u, v) is a subset of the
    m-hop enclosing subgraph's, and with exhaustive walk sets (= the whole ball) both flows coincide."""
    c = Case('cora_scaled')
    links = c.links[:, :10]
    for i in range(links.shape[1]):
        u, v = int(links[0, i]), int(links[1, i])
        sc = orc.scaled_pos_link(u, v, c.sets, c.A, c.X, c.K)
        ball, _, _, _ = orc.k_hop_subgraph(u, v, c.rw_m, c.A)
        assert set(sc['nodes'].tolist()) <= set(ball.tolist())
    import scipy.sparse.csgraph as csg
    u, v = int(links[0, 0]), int(links[1, 0])
    dist = csg.shortest_path(c.A, unweighted=True, indices=[u, v])
    full_sets = {u: np.flatnonzero(dist[0] <= 1), v: np.flatnonzero(dist[1] <= 1)}
    sc = orc.scaled_pos_link(u, v, full_sets, c.A, c.X, c.K)
    bf = orc.pos_link(u, v, 1, c.A, c.X, c.K)
    assert np.array_equal(np.sort(sc['nodes'][2:]), np.sort(bf['nodes'][2:]))
    for k in range(c.K + 1):
        assert_features_close(sc['xs'][k], bf['xs'][k], what=f'x{k}')
