"""Parity tests proper (run on the B200 box): the CUDA path, called through the C ABI, against
(a) the reference's own outputs committed under tests/golden/ and (b) the oracle on seeded
inputs.  Indices (node lists, hop labels, induced + masked edge lists, selected rows) must be
bit-exact; features within 1e-5 of max|ref| per tensor (north_star tolerance)."""
import numpy as np
import pytest
import torch

from golden_util import Case, assert_features_close, case_names
from oracle import s3grl_oracle as orc
from s3grl_b200 import DeviceGraph, datasets as ds, precompute, precompute_full

pytestmark = pytest.mark.gpu


def _graph_edges(g):
    rows = np.repeat(np.arange(g['nodes'].size), np.diff(g['lrowptr']))
    e = np.stack([g['nodes'][rows], g['nodes'][g['lcol']]], 1)
    return e


def _check_indices(got, ref, what):
    assert np.array_equal(got['nodes'], ref['nodes']), f'{what}: nodes'
    assert np.array_equal(got['hops'], ref['hops']), f'{what}: hops'
    assert np.array_equal(got['lrowptr'], ref['lrowptr']), f'{what}: rowptr'
    # the CUDA path lists a row's columns in ascending GLOBAL id, the oracle in ascending local
    # id: same set, compare after canonical (row, global col) ordering
    ge, re_ = _graph_edges(got), _graph_edges(ref)
    assert ge.shape == re_.shape, f'{what}: edge count'
    if ge.size:
        ge = ge[np.lexsort((ge[:, 1], ge[:, 0]))]
        re_ = re_[np.lexsort((re_[:, 1], re_[:, 0]))]
        assert np.array_equal(ge, re_), f'{what}: edges'
    assert np.array_equal(got['sel'], ref['sel']), f'{what}: selected rows'


@pytest.mark.parametrize('name', case_names('pos'))
def test_pos_against_reference_goldens(name):
    c = Case(name)
    g = DeviceGraph(c.A, c.X)
    res = precompute(g, c.links, c.num_hops, c.K, 'PoS', c.strategy, return_graphs=True)
    assert np.array_equal(res.row_ptr.cpu().numpy(), c.row_ptr)
    for k in range(c.K + 1):
        assert_features_close(res.xs[k].cpu().numpy(), c.xs[k], what=f'{name} x{k}')
    assert np.array_equal(res.xs[0].cpu().numpy(), c.xs[0]), 'x (operator 0) must be an exact copy'
    for i, gr in enumerate(res.graphs):
        a, b = c.node_ptr[i], c.node_ptr[i + 1]
        assert np.array_equal(gr['nodes'], c.nodes[a:b]), f'{name} link {i}: nodes vs reference'
        assert np.array_equal(gr['hops'], c.hops[a:b]), f'{name} link {i}: hops vs reference'
        e = _graph_edges(gr)
        ref_e = c.edges[c.edge_ptr[i]:c.edge_ptr[i + 1]].astype(np.int64)
        assert e.shape == ref_e.shape
        if e.size:
            e = e[np.lexsort((e[:, 1], e[:, 0]))]
            ref_e = ref_e[np.lexsort((ref_e[:, 1], ref_e[:, 0]))]
            assert np.array_equal(e, ref_e), f'{name} link {i}: edges vs reference'
        assert np.array_equal(gr['nodes'][gr['sel']], c.row_gid[c.row_ptr[i]:c.row_ptr[i + 1]])


@pytest.mark.parametrize('name', case_names('sop'))
def test_sop_against_reference_goldens(name):
    c = Case(name)
    g = DeviceGraph(c.A, c.X)
    res = precompute(g, c.links, 0, c.K, 'SoP')
    assert np.array_equal(res.row_ptr.cpu().numpy(), c.row_ptr)
    for k in range(c.K + 1):
        assert_features_close(res.xs[k].cpu().numpy(), c.xs[k], what=f'{name} x{k}')


def _random_graph(rng, N, E, self_loops=False):
    u = rng.integers(0, N, E)
    v = rng.integers(0, N, E)
    keep = (u != v) | self_loops
    e = np.unique(np.stack([np.minimum(u, v)[keep], np.maximum(u, v)[keep]], 1), axis=0)
    return ds.adjacency(e, N)


@pytest.mark.parametrize('seed,N,E,F,h,K,strategy', [
    (0, 60, 90, 7, 2, 3, None), (1, 200, 500, 33, 3, 3, 'intersection'), (2, 200, 900, 12, 2, 3, 'union'),
    (3, 1000, 1500, 130, 3, 2, None), (4, 333, 2000, 5, 1, 5, 'union'), (5, 97, 300, 64, 2, 4, 'intersection'),
    (6, 50, 40, 1, 3, 1, None), (7, 3000, 9000, 260, 2, 3, None), (8, 40, 500, 9, 2, 7, None),
    (9, 400, 450, 6, 5, 2, None), (10, 300, 400, 10, 4, 6, 'intersection'), (11, 150, 200, 3, 8, 3, None),
    (12, 5000, 6000, 20, 3, 3, 'union'),
])
def test_pos_against_oracle_random_graphs(seed, N, E, F, h, K, strategy):
    rng = np.random.default_rng(seed)
    A = _random_graph(rng, N, E)
    X = rng.random((N, F), dtype=np.float32)
    links = rng.integers(0, N, (2, 64))
    links = links[:, links[0] != links[1]]
    ref = orc.pos_precompute(links, h, A, X, K, strategy, keep_graphs=True)
    res = precompute(DeviceGraph(A, X), links, h, K, 'PoS', strategy, return_graphs=True)
    assert np.array_equal(res.row_ptr.cpu().numpy(), ref['row_ptr'])
    for i, (g, r) in enumerate(zip(res.graphs, ref['graphs'])):
        _check_indices(g, r, f'link {i}')
    for k in range(K + 1):
        assert_features_close(res.xs[k].cpu().numpy(), ref['xs'][k], what=f'x{k}')


@pytest.mark.parametrize('seed,N,E,F,K', [(0, 80, 150, 6, 3), (1, 500, 1200, 40, 2), (2, 120, 900, 129, 3), (3, 64, 100, 3, 5)])
def test_sop_against_oracle_random_graphs(seed, N, E, F, K):
    rng = np.random.default_rng(100 + seed)
    A = _random_graph(rng, N, E)
    X = rng.random((N, F), dtype=np.float32)
    links = rng.integers(0, N, (2, 48))
    links = links[:, links[0] != links[1]]
    ref = orc.sop_precompute(links, A, X, K)
    res = precompute(DeviceGraph(A, X), links, 0, K, 'SoP')
    for k in range(K + 1):
        assert_features_close(res.xs[k].cpu().numpy(), ref['xs'][k], what=f'x{k}')


def test_batching_and_arena_growth_do_not_change_results():
    """Per-link results must not depend on batch composition (SURVEY.md §7 determinism): one
    big batch, many small batches and a run that starts with a far-too-small arena (forcing the
    overflow/retry path) are bit-identical."""
    c = Case('cora_posplus')
    for strategy in (None, 'intersection'):
        g = DeviceGraph(c.A, c.X)
        a = precompute(g, c.links, c.num_hops, c.K, 'PoS', strategy, batch_records=8192)
        b = precompute(g, c.links, c.num_hops, c.K, 'PoS', strategy, batch_records=7)
        g2 = DeviceGraph(c.A, c.X)
        d = precompute(g2, c.links, c.num_hops, c.K, 'PoS', strategy, batch_records=64, arena_words=4096)
        assert d.stats['retries'] > 0
        for k in range(c.K + 1):
            assert torch.equal(a.xs[k], b.xs[k]) and torch.equal(a.xs[k], d.xs[k])
        assert torch.equal(a.row_ptr, b.row_ptr) and torch.equal(a.row_ptr, d.row_ptr)


def test_invalid_links_and_strategy_raise():
    c = Case('tiny_pos_h2')
    g = DeviceGraph(c.A, c.X)
    with pytest.raises(ValueError):
        precompute(g, np.array([[0, 3], [1, 3]]), 2, 3)            # src == dst
    with pytest.raises(ValueError):
        precompute(g, np.array([[0], [c.N]]), 2, 3)                # out of range
    with pytest.raises(NotImplementedError):
        precompute(g, c.links, 2, 3, 'PoS', 'both')                # reference: check strat
    res = precompute(g, np.zeros((2, 0), dtype=np.int64), 2, 3)    # empty call
    assert res.xs[0].shape == (0, c.X.shape[1] + 1) and res.row_ptr.tolist() == [0]


def test_reference_interface_mirror():
    """The drop-in boundary: same call the reference's SEALDataset.process makes
    (sgrl_link_pred.py:195-203), returning a sequence of Data with x, y, x1..xK."""
    from s3grl_b200 import extract_enclosing_subgraphs
    c = Case('cora_posplus')
    x = torch.from_numpy(c.X)
    link_index = torch.from_numpy(c.links[:, :40])
    sign_kwargs = dict(sign_k=c.K, use_feature=True, sign_type='PoS', optimize_sign=True, k_heuristic=1,
                       k_node_set_strategy='intersection')
    pos = extract_enclosing_subgraphs(link_index, c.A, x, 1, c.num_hops, 'zo', 1.0, None, False, None, None,
                                      sign_kwargs, powers_of_A=[], data=None)
    neg = extract_enclosing_subgraphs(link_index, c.A, x, 0, c.num_hops, 'zo', 1.0, None, False, None, None,
                                      sign_kwargs, powers_of_A=[], data=None)
    both = pos + neg
    assert len(pos) == 40 and len(both) == 80
    d = both[3]
    a, b = c.row_ptr[3], c.row_ptr[4]
    assert d.y == 1 and both[43].y == 0 and not d.x.is_cuda
    for k, key in enumerate(['x', 'x1', 'x2', 'x3']):
        assert_features_close(d[key].numpy(), c.xs[k][a:b], what=key)
    data, slices = both.collate()
    assert data.x.shape[0] == 2 * c.row_ptr[40] and slices['x2'].shape[0] == 81 and data.y.tolist() == [1] * 40 + [0] * 40
    # hybrid (utils.py:454-480): K PoS operators + SoP x2..xK
    sign_kwargs.update(sign_type='hybrid', k_heuristic=0)
    hyb = extract_enclosing_subgraphs(link_index, c.A, x, 1, c.num_hops, 'zo', 1.0, None, False, None, None,
                                      sign_kwargs, powers_of_A=[1, 2, 3], data=None)
    ref = orc.hybrid_precompute(c.links[:, :40], c.num_hops, c.A, c.X, c.K)
    assert len(hyb.xs) == 2 * c.K
    for k in range(2 * c.K):
        assert_features_close(hyb.xs[k].numpy(), ref['xs'][k], what=f'hybrid x{k}')


def test_dump_edges_entry_point():
    """s3_dump_edges returns the canonical global-id edge list of every record."""
    import ctypes as C
    from s3grl_b200 import _lib as L
    c = Case('tiny_pos_h2')
    g = DeviceGraph(c.A, c.X)
    dev = g.device
    links = torch.from_numpy(c.links).to(dev)
    nrec = c.L
    arena = torch.empty(1 << 20, dtype=torch.int32, device=dev)
    off = torch.empty((nrec, L.NOFF), dtype=torch.int64, device=dev)
    cnt = torch.empty((nrec, L.NCNT), dtype=torch.int32, device=dev)
    ctr = torch.zeros(L.NCTR, dtype=torch.int64, device=dev)
    b = L.Batch(links[0].data_ptr(), links[1].data_ptr(), nrec, L.FLOW_POS, 0, 2, 3, L.BATCH_STORE_ALL_ROWS, 0,
                arena.data_ptr(), arena.numel(),
                off.data_ptr(), cnt.data_ptr(), ctr.data_ptr(), None, None, None, None)
    lib = L.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(lib.s3_extract(C.byref(g._c), C.byref(b), st), 'extract')
    m = cnt[:, L.CNT_M].to(torch.int64)
    edge_ptr = torch.zeros(nrec + 1, dtype=torch.int64, device=dev)
    edge_ptr[1:] = torch.cumsum(m, 0)
    edges = torch.empty((int(edge_ptr[-1]), 2), dtype=torch.int32, device=dev)
    L.check(lib.s3_dump_edges(C.byref(b), C.c_void_p(edge_ptr.data_ptr()), C.c_void_p(edges.data_ptr()), st), 'dump')
    got = edges.cpu().numpy()
    ep = edge_ptr.cpu().numpy()
    assert np.array_equal(ep, c.edge_ptr)
    for i in range(nrec):
        e, r = got[ep[i]:ep[i + 1]], c.edges[ep[i]:ep[i + 1]]
        if e.size:
            assert np.array_equal(e[np.lexsort((e[:, 1], e[:, 0]))], r[np.lexsort((r[:, 1], r[:, 0]))])


def test_pubmed_full_size_sample_and_properties():
    """BASELINE config 3 at full size: PubMed training graph, F = 500, h = 3, K = 3.  A seeded
    sample of links is checked against the oracle; the full run is checked through
    size-independent properties: x rows are exact copies of [1 | X[u]], operator rows are
    non-negative with row sums <= label + 1 (S is sub-stochastic up to D^1/2 scaling ... so only
    finiteness and the label-column identity are asserted), and link (u,v) equals link (v,u)
    with its two rows swapped."""
    edges, N, _ = ds.load_graph('pubmed')
    A, splits = ds.split_links(edges, N, seed=1)
    X = ds.synthetic_features(N, 500, 0.1, 0)
    links = ds.all_links(splits)
    g = DeviceGraph(A, X)
    res = precompute(g, links, 3, 3)
    assert res.xs[0].shape == (2 * links.shape[1], 501)
    x0 = res.xs[0].cpu().numpy()
    assert np.array_equal(x0[:, 0], np.ones(x0.shape[0], np.float32))
    assert np.array_equal(x0[0::2, 1:], X[links[0]]) and np.array_equal(x0[1::2, 1:], X[links[1]])
    for k in range(1, 4):
        assert bool(torch.isfinite(res.xs[k]).all()) and float(res.xs[k].min()) >= 0.0
    # symmetry: train positives hold both directions of every edge
    rng = np.random.default_rng(0)
    pick = rng.choice(links.shape[1], 300, replace=False)
    sub = links[:, pick]
    fwd = precompute(g, sub, 3, 3)
    rev = precompute(g, sub[::-1].copy(), 3, 3)
    for k in range(4):
        a = fwd.xs[k].view(-1, 2, 501)
        b = rev.xs[k].view(-1, 2, 501).flip(1)
        assert torch.equal(a, b), f"swap x{k}: (u,v) and (v,u) must be the same bits (link pairing relies on it)"
    # oracle on a sample
    ref = orc.pos_precompute(sub[:, :48], 3, A, X, 3, keep_graphs=True)
    got = precompute(g, sub[:, :48], 3, 3, return_graphs=True)
    for i, (gg, r) in enumerate(zip(got.graphs, ref['graphs'])):
        _check_indices(gg, r, f'pubmed link {i}')
    for k in range(4):
        assert_features_close(got.xs[k].cpu().numpy(), ref['xs'][k], what=f'pubmed x{k}')
    # and the sampled rows of the full run are the same bits as the sampled run
    for k in range(4):
        full = res.xs[k].view(-1, 2, 501)[torch.from_numpy(pick).to(res.xs[k].device)]
        assert torch.equal(full, fwd.xs[k].view(-1, 2, 501))


@pytest.mark.parametrize('seed,N,E,F,K', [(0, 300, 900, 9, 3), (1, 2000, 30000, 33, 3), (2, 500, 6000, 128, 5), (3, 64, 80, 4, 1)])
def test_sorted_tier_against_oracle(seed, N, E, F, K):
    """The large-graph extraction tier (sorted merge, num_hops = 1), forced on small graphs so the
    oracle can check it: bit-exact indices, features within tolerance, and bit-identical to the
    bitmap tier's operators is NOT required (different summation order) but both must agree."""
    rng = np.random.default_rng(200 + seed)
    A = _random_graph(rng, N, E)
    hub = int(np.argmax(np.diff(A.indptr)))
    X = rng.random((N, F), dtype=np.float32)
    links = rng.integers(0, N, (2, 60))
    links[0, :6] = hub                       # hub rows exercise the search-in-adjacency side
    links = links[:, links[0] != links[1]]
    ref = orc.pos_precompute(links, 1, A, X, K, None, keep_graphs=True)
    g = DeviceGraph(A, X)
    res = precompute(g, links, 1, K, 'PoS', None, return_graphs=True, force_sorted_tier=True)
    for i, (gg, r) in enumerate(zip(res.graphs, ref['graphs'])):
        _check_indices(gg, r, f'link {i}')
    for k in range(K + 1):
        assert_features_close(res.xs[k].cpu().numpy(), ref['xs'][k], what=f'x{k}')
    bitmap = precompute(g, links, 1, K, 'PoS', None)
    for k in range(K + 1):
        assert_features_close(res.xs[k].cpu().numpy(), bitmap.xs[k].cpu().numpy(), tol=2e-6, what=f'tiers x{k}')


def test_sorted_tier_on_reference_golden():
    c = Case('tiny_pos_h1')
    res = precompute(DeviceGraph(c.A, c.X), c.links, 1, c.K, 'PoS', None, force_sorted_tier=True)
    for k in range(c.K + 1):
        assert_features_close(res.xs[k].cpu().numpy(), c.xs[k], what=f'x{k}')


def test_rmat_large_graph_uses_sorted_tier():
    """A graph beyond the bitmap tier (1 M nodes): R-MAT, hub links included, against the oracle."""
    import ctypes as C
    from s3grl_b200 import _lib as L
    edges = ds.rmat_edges(20, 1_000_000, seed=42)
    N = 1 << 20
    A = ds.adjacency(edges, N)
    X = ds.synthetic_features(N, 16, 1.0, 43)
    rng = np.random.default_rng(44)
    deg = np.diff(A.indptr)
    hubs = np.argsort(deg)[-1:]              # one hub link (n ~ 9 000): the oracle needs seconds for it
    pos = edges[rng.choice(edges.shape[0], 30, replace=False)].T
    neg = rng.integers(0, N, (2, 30))
    hub_links = np.stack([hubs, A.indices[A.indptr[hubs]]])     # each hub with its first neighbour
    links = np.concatenate([pos, neg[:, neg[0] != neg[1]], hub_links], axis=1)
    g = DeviceGraph(A, X, check_symmetric=False)
    probe = L.Batch(None, None, 0, L.FLOW_POS, 0, 1, 3, 0, 0, None, 0, None, None, None, None, None, None, None)
    assert L.lib().s3_extract_tier(C.byref(g._c), C.byref(probe)) == 1
    res = precompute(g, links, 1, 3)
    ref = orc.pos_precompute(links, 1, A, X, 3, None, keep_graphs=True)
    assert res.stats['max_n'] == max(gr['nodes'].size for gr in ref['graphs']) > 1000
    for k in range(4):
        assert_features_close(res.xs[k].cpu().numpy(), ref['xs'][k], what=f'rmat x{k}')
    with pytest.raises(Exception):
        precompute(g, links[:, :4], 2, 3)        # 2 hops on a graph of this size: unsupported


def test_schedules_are_bit_identical():
    """Synchronous, deferred, two-stream overlapped and host-pipelined runs give the same bits."""
    c = Case('cora_pos')
    g = DeviceGraph(c.A, c.X)
    base = precompute(g, c.links, c.num_hops, c.K, batch_records=40)
    deferred = precompute(g, c.links, c.num_hops, c.K, batch_records=40, defer=True).finalize()
    overlapped = precompute(g, c.links, c.num_hops, c.K, batch_records=40, overlap=True)
    host = [torch.empty((2 * c.L, c.X.shape[1] + 1), dtype=torch.float32, pin_memory=True) for _ in range(c.K + 1)]
    piped = precompute(g, c.links, c.num_hops, c.K, batch_records=40, host_out=host)
    for k in range(c.K + 1):
        assert torch.equal(base.xs[k], deferred.xs[k]) and torch.equal(base.xs[k], overlapped.xs[k])
        assert torch.equal(base.xs[k].cpu(), host[k]) and torch.equal(piped.xs[k], base.xs[k])
    assert deferred.stats['sum_n'] == base.stats['sum_n'] > 0


def test_sop_and_k5_through_reference_interface():
    """get_SoP_prepped_ds mirror on host tensors, and K = 5 / num_hops = 2 (BASELINE config 4)."""
    from s3grl_b200 import OptimizedSignOperations as Ops
    c = Case('usair_sop')
    out = Ops.get_SoP_prepped_ds([None] * c.K, torch.from_numpy(c.links), c.A, torch.from_numpy(c.X), 1)
    assert len(out) == c.L and not out.xs[0].is_cuda
    for k in range(c.K + 1):
        assert_features_close(out.xs[k].numpy(), c.xs[k], what=f'sop x{k}')
    c = Case('yeast_pos_k5')
    kw = dict(sign_k=5, use_feature=True, sign_type='PoS', optimize_sign=True, k_heuristic=0, k_node_set_strategy=None)
    out = Ops.get_PoS_prepped_ds(torch.from_numpy(c.links), 2, c.A, 1.0, None, False, None, torch.from_numpy(c.X), 0, kw, None)
    for k in range(6):
        assert_features_close(out.xs[k].numpy(), c.xs[k], what=f'yeast x{k}')
    with pytest.raises(NotImplementedError):      # directed BFS (utils.py:58-63) stays out of scope
        Ops.get_PoS_prepped_ds(torch.from_numpy(c.links), 2, c.A, 1.0, None, True, None, torch.from_numpy(c.X), 0, kw, None)
    capped = Ops.get_PoS_prepped_ds(torch.from_numpy(c.links), 2, c.A, 1.0, 50, False, None, torch.from_numpy(c.X), 0, kw, None)
    ref = orc.pos_precompute(c.links, 2, c.A, c.X, 5, None, caps=dict(ratio_per_hop=1.0, max_nodes_per_hop=50, cap_seed=0))
    for k in range(6):
        assert_features_close(capped.xs[k].numpy(), ref['xs'][k], what=f'yeast capped x{k}')


def test_self_loops_follow_the_reference_semantics():
    """A[nodes,:][:,nodes] keeps diagonal entries (utils.py:76): a self loop is an edge (j, j) of the
    induced subgraph, counts in the degree and carries weight.  Both tiers, against the oracle."""
    rng = np.random.default_rng(77)
    u = rng.integers(0, 120, 400)
    v = rng.integers(0, 120, 400)
    v[:60] = u[:60]                                   # 60 self loops
    row, col = np.r_[u, v], np.r_[v, u]
    import scipy.sparse as ssp
    A = ssp.csr_matrix((np.ones(row.size, np.int64), (row, col)), shape=(120, 120))
    A.sum_duplicates()
    A.data[:] = 1
    X = rng.random((120, 11), dtype=np.float32)
    links = rng.integers(0, 120, (2, 50))
    links = links[:, links[0] != links[1]]
    g = DeviceGraph(A, X)
    for h, K, kw in ((2, 3, {}), (1, 3, {}), (1, 3, dict(force_sorted_tier=True)), (3, 2, {})):
        ref = orc.pos_precompute(links, h, A, X, K, None, keep_graphs=True)
        res = precompute(g, links, h, K, return_graphs=True, **kw)
        for i, (gg, r) in enumerate(zip(res.graphs, ref['graphs'])):
            _check_indices(gg, r, f'h={h} link {i}')
        for k in range(K + 1):
            assert_features_close(res.xs[k].cpu().numpy(), ref['xs'][k], what=f'h={h} x{k}')


def test_scaled_subgraphs_from_a_walk_cache_match_reference():
    """ScaLed with the walk sets of the golden fixture (the form the reference's create_rw_cache
    hands over): everything downstream of the walks is exact — parity with the reference's output,
    through the tensor API and through the rw_kwargs of the reference-facing call."""
    from s3grl_b200 import OptimizedSignOperations as Ops
    c = Case('cora_scaled')
    g = DeviceGraph(c.A, c.X)
    cache = {k: torch.from_numpy(v) for k, v in c.sets.items()}
    res = precompute(g, c.links, 0, c.K, walk=dict(cache=cache), return_graphs=True)
    ref = orc.scaled_pos_precompute(c.links, c.sets, c.A, c.X, c.K, keep_graphs=True)
    for i, (gg, r) in enumerate(zip(res.graphs, ref['graphs'])):
        _check_indices(gg, r, f'scaled link {i}')
    for k in range(c.K + 1):
        assert_features_close(res.xs[k].cpu().numpy(), c.xs[k], what=f'scaled x{k}')
    kw = dict(sign_k=c.K, use_feature=True, sign_type='PoS', optimize_sign=True, k_heuristic=0, k_node_set_strategy=None)
    rw_kwargs = dict(rw_m=c.rw_m, rw_M=c.rw_M, cached_pos_rws=cache, cached_neg_rws=None, sign=True)
    out = Ops.get_PoS_prepped_ds(torch.from_numpy(c.links), 0, c.A, 1.0, None, False, None, torch.from_numpy(c.X), 1, kw, rw_kwargs)
    for k in range(c.K + 1):
        assert_features_close(out.xs[k].numpy(), c.xs[k], what=f'scaled mirror x{k}')


def test_gpu_random_walk_sets():
    """The CUDA sampler (s3_walk_sets): every set is sorted, unique, contains its start node, has at
    most 1 + M*m nodes all within m steps of the start; sets depend only on (seed, node); and the
    set-size distribution matches a NumPy simulation of the same process (the reference's
    torch_cluster RNG stream cannot be reproduced: parity is distributional)."""
    import scipy.sparse.csgraph as csg
    from s3grl_b200 import walk_sets
    c = Case('cora_scaled')
    g = DeviceGraph(c.A, c.X)
    starts = torch.arange(c.N)
    sets, counts = walk_sets(g, starts, 3, 20, seed=7)
    sets, counts = sets.cpu().numpy(), counts.cpu().numpy()
    again, again_n = walk_sets(g, starts[100:200], 3, 20, seed=7)
    again, again_n = again.cpu().numpy(), again_n.cpu().numpy()
    assert np.array_equal(again_n, counts[100:200])                                 # independent of the call
    assert all(np.array_equal(again[i, :again_n[i]], sets[100 + i, :again_n[i]]) for i in range(100))
    _, other_n = walk_sets(g, starts, 3, 20, seed=8)
    assert not np.array_equal(other_n.cpu().numpy(), counts)
    dist = csg.shortest_path(c.A, unweighted=True, indices=np.arange(0, c.N, 37))
    for row, s in enumerate(range(0, c.N, 37)):
        nodes = sets[s, :counts[s]]
        assert s in nodes and counts[s] <= 61 and np.all(np.diff(nodes) > 0) and np.all(dist[row][nodes] <= 3)
    deg = np.diff(c.A.indptr)
    assert np.all(counts[deg == 0] == 1)
    sim = orc.random_walk_sets(c.A, np.arange(c.N), 3, 20, seed=3)
    sim_mean = np.mean([v.size for v in sim.values()])
    assert abs(counts.mean() - sim_mean) / sim_mean < 0.03, (counts.mean(), sim_mean)
    # end to end on sampled walks: identical to the oracle fed with the same sets
    links = c.links[:, :30]
    res = precompute(g, links, 0, c.K, walk=dict(m=3, M=20, seed=7))
    as_dict = {int(s): sets[s, :counts[s]].astype(np.int64) for s in np.unique(links)}
    ref = orc.scaled_pos_precompute(links, as_dict, c.A, c.X, c.K)
    for k in range(c.K + 1):
        assert_features_close(res.xs[k].cpu().numpy(), ref['xs'][k], what=f'sampled x{k}')


@pytest.mark.parametrize('name', case_names('full'))
def test_non_optimised_flow_against_reference_goldens(name):
    """SURVEY.md §8a row 9 (reference utils.py:497-520, optimize_sign=False): SIGN on the whole subgraph,
    all labeling tricks; node ids bit-exact, x an exact copy, x1..xK within tolerance."""
    c = Case(name)
    res = precompute_full(DeviceGraph(c.A, c.X), c.links, c.num_hops, c.K, node_label=c.node_label)
    assert np.array_equal(res.row_ptr.cpu().numpy(), c.row_ptr)
    assert np.array_equal(res.node_id.cpu().numpy(), c.node_id)
    assert np.array_equal(res.xs[0].cpu().numpy(), c.xs[0])
    for k in range(1, c.K + 1):
        assert_features_close(res.xs[k].cpu().numpy(), c.xs[k], what=f'{name} x{k}')


@pytest.mark.parametrize('seed,N,E,F,h,K,label,batch', [
    (0, 80, 160, 7, 2, 3, 'drnl', None), (1, 400, 900, 130, 3, 2, 'zo', 16), (2, 300, 1500, 200, 2, 4, 'drnl', 7),
    (3, 1000, 1400, 33, 4, 3, 'hop', None), (4, 64, 500, 1, 1, 5, 'degree', 5), (5, 2000, 3000, 513, 3, 3, 'drnl', 32),
    (6, 700, 3000, 70, 3, 3, 'drnl', None), (7, 1300, 6000, 40, 3, 2, 'hop', None), (8, 150, 2500, 300, 2, 3, 'drnl', None),
])
def test_non_optimised_flow_against_oracle(seed, N, E, F, h, K, label, batch):
    rng = np.random.default_rng(300 + seed)
    A = _random_graph(rng, N, E)
    X = rng.random((N, F), dtype=np.float32)
    links = rng.integers(0, N, (2, 40))
    links = links[:, links[0] != links[1]]
    ref = orc.full_precompute(links, h, A, X, K, label)
    g = DeviceGraph(A, X)
    res = precompute_full(g, links, h, K, node_label=label, batch_records=batch)
    assert np.array_equal(res.row_ptr.cpu().numpy(), ref['row_ptr'])
    assert np.array_equal(res.node_id.cpu().numpy(), ref['node_id'])
    assert np.array_equal(res.xs[0].cpu().numpy(), ref['xs'][0])
    for k in range(1, K + 1):
        assert_features_close(res.xs[k].cpu().numpy(), ref['xs'][k], what=f'x{k}')
    if label == 'zo':
        # rows 0, 1 of the SpMM chain == the optimised PoS flow (row-vector propagation): two
        # algebraically different CUDA routes to the same numbers
        opt = precompute(g, links, h, K, 'PoS')
        rp = res.row_ptr[:-1]
        for k in range(K + 1):
            both = torch.stack([res.xs[k][rp], res.xs[k][rp + 1]], 1).reshape(-1, F + 1)
            assert_features_close(both.cpu().numpy(), opt.xs[k].cpu().numpy(), what=f'full vs optimised x{k}')


def test_non_optimised_flow_through_reference_interface():
    from s3grl_b200 import extract_enclosing_subgraphs
    c = Case('cora_full_drnl')
    kw = {'sign_k': c.K, 'use_feature': True, 'sign_type': 'PoS', 'optimize_sign': False, 'k_heuristic': 0,
          'k_node_set_strategy': None}
    out = extract_enclosing_subgraphs(torch.as_tensor(c.links), c.A, torch.as_tensor(c.X), 1, c.num_hops, 'drnl',
                                      sign_kwargs=kw, powers_of_A=[])
    assert len(out) == c.L
    d = out[3]
    a, b = c.row_ptr[3], c.row_ptr[4]
    assert np.array_equal(d['node_id'].numpy(), c.node_id[a:b]) and d['y'] == 1
    assert np.array_equal(d['x'].numpy(), c.xs[0][a:b])
    assert_features_close(d[f'x{c.K}'].numpy(), c.xs[c.K][a:b], what='last operator of link 3')
    with pytest.raises(NotImplementedError):
        extract_enclosing_subgraphs(torch.as_tensor(c.links), c.A, torch.as_tensor(c.X), 1, c.num_hops, 'de',
                                    sign_kwargs=kw, powers_of_A=[])
    with pytest.raises(NotImplementedError):
        extract_enclosing_subgraphs(torch.as_tensor(c.links), c.A, torch.as_tensor(c.X), 1, c.num_hops, 'drnl',
                                    sign_kwargs=kw, powers_of_A=[1, 2, 3])


def test_non_optimised_flow_on_the_sorted_tier():
    rng = np.random.default_rng(77)
    A = _random_graph(rng, 500, 2500)
    X = rng.random((500, 40), dtype=np.float32)
    links = rng.integers(0, 500, (2, 30))
    links = links[:, links[0] != links[1]]
    ref = orc.full_precompute(links, 1, A, X, 3, 'drnl')
    res = precompute_full(DeviceGraph(A, X), links, 1, 3, node_label='drnl', force_sorted_tier=True)
    assert np.array_equal(res.node_id.cpu().numpy(), ref['node_id'])
    for k in range(4):
        assert_features_close(res.xs[k].cpu().numpy(), ref['xs'][k], what=f'x{k}')


@pytest.mark.parametrize('N,E,F,h,K,L', [(300, 900, 40, 2, 3, 40), (2500, 9000, 70, 3, 3, 12), (5000, 40000, 33, 3, 2, 6),
                                         (9000, 70000, 20, 3, 3, 4), (600, 2400, 500, 3, 5, 16), (400, 1500, 9, 1, 3, 30),
                                         (1500, 6000, 127, 3, 3, 20), (700, 2000, 128, 3, 3, 20), (1000, 2600, 3, 4, 4, 30),
                                         (3600, 12000, 9, 4, 3, 5), (900, 2000, 64, 3, 1, 25), (1200, 9000, 31, 2, 2, 10)])
def test_union_chain_matches_work_item_path(N, E, F, h, K, L):
    """PoS Plus union: the default hop-limited SpMM chain (s3_ccn_chain, for records that fit shared memory;
    larger records fall to work items inside the same call) against the all-work-item path
    (s3_diffuse + s3_gather_ccn, round 1's default), both also against the oracle: CW = 32 / 16 / 8 / 4, all three
    CTA sizes and mixed batches.  The chain's results do not depend on the batch composition."""
    rng = np.random.default_rng(N)
    A = _random_graph(rng, N, E)
    X = rng.random((N, F), dtype=np.float32)
    links = rng.integers(0, N, (2, L))
    links = links[:, links[0] != links[1]]
    g = DeviceGraph(A, X)
    a = precompute(g, links, h, K, 'PoS', 'union')                     # default: chain where it fits shared memory
    b = precompute(g, links, h, K, 'PoS', 'union', ccn_mode='items')   # work items for every record
    c = precompute(g, links, h, K, 'PoS', 'union', ccn_mode='chain', batch_records=5)
    assert torch.equal(a.row_ptr, c.row_ptr) and all(torch.equal(x, y) for x, y in zip(a.xs, c.xs))
    ref = orc.pos_precompute(links, h, A, X, K, 'union')
    assert np.array_equal(a.row_ptr.cpu().numpy(), ref['row_ptr'])
    for k in range(K + 1):
        assert_features_close(a.xs[k].cpu().numpy(), ref['xs'][k], what=f'N={N} chain vs oracle x{k}')
    assert torch.equal(a.row_ptr, b.row_ptr) and a.stats['max_n'] == b.stats['max_n']
    assert torch.equal(a.xs[0], b.xs[0])                 # x: exact copies on both routes
    for k in range(1, K + 1):
        assert_features_close(a.xs[k].cpu().numpy(), b.xs[k].cpu().numpy(), what=f'N={N} x{k}')
    rp = a.row_ptr[:-1]
    for k in range(K + 1):                               # rows 0, 1 are produced by kernels 1 + 3 on both routes
        assert torch.equal(a.xs[k][rp], b.xs[k][rp]) and torch.equal(a.xs[k][rp + 1], b.xs[k][rp + 1])
    print(f"N={N}: max_n={a.stats['max_n']} rows={a.stats['rows']}")


@pytest.mark.parametrize('pool_cw,split', [(4, 0), (8, 1), (16, 2), (32, 0), (32, 2), (0, 0)])
@pytest.mark.parametrize('N,E,F,h,K,L', [(2500, 9000, 70, 3, 3, 12), (5000, 40000, 33, 3, 2, 6), (600, 2400, 500, 3, 5, 16),
                                         (3600, 12000, 9, 4, 3, 5)])
def test_union_chain_pooled_route(monkeypatch, pool_cw, split, N, E, F, h, K, L):
    """s3_ccn_chain_pooled: records that would run at a sub-chunk width <= pool_cw in shared memory take two [n][32]
    buffers from the global pool instead.  Same sums in a different lane-group split for the warp-wide rows: within fp32
    rounding of the shared-memory route, 1e-5 against the oracle, and independent of the batch composition."""
    from s3grl_b200 import engine
    rng = np.random.default_rng(N)
    A = _random_graph(rng, N, E)
    X = rng.random((N, F), dtype=np.float32)
    links = rng.integers(0, N, (2, L))
    links = links[:, links[0] != links[1]]
    g = DeviceGraph(A, X)
    a = precompute(g, links, h, K, 'PoS', 'union')                     # the default threshold
    monkeypatch.setattr(engine, '_CHAIN_POOL_CW', pool_cw)             # 0: every record in shared memory
    monkeypatch.setenv('S3GRL_CHAIN_POOL_SPLIT', str(split))           # pooled launches by the shared memory the CSR needs
    b = precompute(g, links, h, K, 'PoS', 'union')
    c = precompute(g, links, h, K, 'PoS', 'union', batch_records=5)
    assert int(g.chain_pool()[1].sum()) == 0                           # every slot released
    assert torch.equal(b.row_ptr, c.row_ptr) and all(torch.equal(x, y) for x, y in zip(b.xs, c.xs))
    # level 1 straight from the feature matrix instead of a stored y_0: the same bits (the product is rounded before the sum)
    monkeypatch.setenv('S3GRL_CHAIN_POOL_X', '1')
    d = precompute(g, links, h, K, 'PoS', 'union', batch_records=7)
    monkeypatch.setenv('S3GRL_CHAIN_SMEM_X', '1')                      # ... and for the records with shared-memory buffers
    d = precompute(g, links, h, K, 'PoS', 'union', batch_records=7)
    monkeypatch.setenv('S3GRL_CHAIN_POOL_X', '0')
    monkeypatch.setenv('S3GRL_CHAIN_SMEM_X', '0')
    assert torch.equal(b.row_ptr, d.row_ptr) and all(torch.equal(x, y) for x, y in zip(b.xs, d.xs))
    assert torch.equal(a.row_ptr, b.row_ptr) and torch.equal(a.xs[0], b.xs[0])
    ref = orc.pos_precompute(links, h, A, X, K, 'union')
    for k in range(1, K + 1):
        assert_features_close(b.xs[k].cpu().numpy(), a.xs[k].cpu().numpy(), tol=2e-6, what=f'pooled vs shared x{k}')
        assert_features_close(b.xs[k].cpu().numpy(), ref['xs'][k], what=f'pooled vs oracle x{k}')


@pytest.mark.parametrize('strategy', ['intersection', 'union'])
def test_pos_plus_pairing_places_both_directions(strategy):
    """PoS Plus with link pairing (csrc/expand.cu): a list holding (u,v), (v,u) and exact repeats runs the path once per
    unordered node pair; every link still gets its own rows — rows 0 / 1 exchanged for the opposite direction (bit-exact:
    they come from kernels 1 + 3), the CCN rows within fp32 rounding of computing the link on its own — and the oracle
    agrees with all of them. Several batches: the head of a chain and its members sit in different pieces."""
    rng = np.random.default_rng(91)
    A = _random_graph(rng, 500, 2200)
    X = rng.random((500, 37), dtype=np.float32)
    e = np.stack(A.nonzero(), 1)
    base = np.concatenate([e[rng.choice(e.shape[0], 25, replace=False)].T, rng.integers(0, 500, (2, 15))], axis=1)
    base = base[:, base[0] != base[1]]
    links = np.concatenate([base, base[::-1, :20], base[:, 5:12], rng.integers(0, 500, (2, 6))], axis=1)
    links = links[:, links[0] != links[1]]
    links = links[:, rng.permutation(links.shape[1])]
    g = DeviceGraph(A, X)
    ref = orc.pos_precompute(links, 2, A, X, 3, strategy)
    for batch in (None, 7):
        on = precompute(g, links, 2, 3, 'PoS', strategy, batch_records=batch)
        off = precompute(g, links, 2, 3, 'PoS', strategy, batch_records=batch, pair=False)
        assert on.stats['mirrors'] >= 27 and off.stats['mirrors'] == 0
        assert np.array_equal(on.row_ptr.cpu().numpy(), ref['row_ptr']) and torch.equal(on.row_ptr, off.row_ptr)
        rp = on.row_ptr[:-1]
        for k in range(4):
            assert_features_close(on.xs[k].cpu().numpy(), ref['xs'][k], what=f'{strategy} paired x{k}')
            assert_features_close(on.xs[k].cpu().numpy(), off.xs[k].cpu().numpy(), tol=2e-6, what=f'{strategy} pair on/off x{k}')
            assert torch.equal(on.xs[k][rp], off.xs[k][rp]) and torch.equal(on.xs[k][rp + 1], off.xs[k][rp + 1])
        assert torch.equal(on.xs[0], off.xs[0])


def test_scatter_rows_entry_point():
    """s3_pair_heads / s3_scatter_rows through the raw C ABI against a NumPy restatement of the placement."""
    import ctypes as C
    from s3grl_b200 import _lib as L
    from s3grl_b200.engine import pair_links
    lib = L.lib()
    rng = np.random.default_rng(5)
    u = rng.integers(0, 30, 200)
    v = (u + 1 + rng.integers(0, 28, 200)) % 30
    links = torch.from_numpy(np.stack([u, v])).cuda()
    Lk = links.shape[1]
    mirror, _table = pair_links(links, 30)
    code = torch.empty(Lk, dtype=torch.int64, device='cuda')
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(lib.s3_pair_heads(C.c_void_p(mirror.data_ptr()), Lk, C.c_void_p(code.data_ptr()), st), 's3_pair_heads')
    code_h, m_h = code.cpu().numpy(), mirror.cpu().numpy()
    first = {}
    for i in range(Lk):
        key = (min(u[i], v[i]), max(u[i], v[i]))
        first.setdefault(key, i)
        h = first[key]
        assert code_h[i] >> 1 == h and (code_h[i] & 1) == int(u[i] != u[h]) and (m_h[i] >= -1) == (h == i)
    heads = np.array(sorted(set(first.values())))
    counts = rng.integers(2, 6, heads.size)
    prp = np.concatenate([[0], np.cumsum(counts)])
    cols, ops = 11, 3
    src = [torch.from_numpy(rng.random((int(prp[-1]), cols), dtype=np.float32)).cuda() for _ in range(ops)]
    rank = {int(h): r for r, h in enumerate(heads)}
    full_counts = np.array([counts[rank[int(c >> 1)]] for c in code_h])
    frp = np.concatenate([[0], np.cumsum(full_counts)])
    dst = [torch.zeros((int(frp[-1]), cols), dtype=torch.float32, device='cuda') for _ in range(ops)]
    sp = (C.c_void_p * ops)(*[t.data_ptr() for t in src])
    dp = (C.c_void_p * ops)(*[t.data_ptr() for t in dst])
    t_prp, t_heads, t_frp = torch.from_numpy(prp).cuda(), torch.from_numpy(heads).cuda(), torch.from_numpy(frp).cuda()
    L.check(lib.s3_scatter_rows(sp, cols, C.c_void_p(t_prp.data_ptr()), heads.size, C.c_void_p(t_heads.data_ptr()), 0,
                                C.c_void_p(mirror.data_ptr()), C.c_void_p(t_frp.data_ptr()), dp, cols, ops, cols, st), 's3_scatter_rows')
    for k in range(ops):
        got, s_h = dst[k].cpu().numpy(), src[k].cpu().numpy()
        for i in range(Lk):
            r = rank[int(code_h[i] >> 1)]
            rows = s_h[prp[r]:prp[r + 1]].copy()
            if code_h[i] & 1:
                rows[[0, 1]] = rows[[1, 0]]
            assert np.array_equal(got[frp[i]:frp[i + 1]], rows), (k, i)
    # s3_scatter_rows_lead: rows 0 / 1 of every record written a second time in front (compat_explicit_zero layout)
    frp2 = frp + 2 * np.arange(Lk + 1)
    dst2 = [torch.zeros((int(frp2[-1]), cols), dtype=torch.float32, device='cuda') for _ in range(ops)]
    dp2 = (C.c_void_p * ops)(*[t.data_ptr() for t in dst2])
    t_frp2 = torch.from_numpy(frp2).cuda()
    L.check(lib.s3_scatter_rows_lead(sp, cols, C.c_void_p(t_prp.data_ptr()), heads.size, C.c_void_p(t_heads.data_ptr()), 0,
                                     C.c_void_p(mirror.data_ptr()), C.c_void_p(t_frp2.data_ptr()), dp2, cols, ops, cols, 2, st),
            's3_scatter_rows_lead')
    for k in range(ops):
        got, ref = dst2[k].cpu().numpy(), dst[k].cpu().numpy()
        for i in range(Lk):
            rows = ref[frp[i]:frp[i + 1]]
            assert np.array_equal(got[frp2[i]:frp2[i + 1]], np.concatenate([rows[:2], rows])), (k, i)


@pytest.mark.parametrize('name', [n for n in case_names('pos') if 'union' in n])
def test_union_compat_explicit_zero_reproduces_the_literal_reference_rows(name):
    """with compat_explicit_zero the union rows are the code-literal ones of the reference (its label-column
    literal repaired, oracle/ref_runner.union_typo_repaired) — src and dst selected a second time.  The fixtures hold the
    literal rows with the extra rows in ascending global id; the framework writes [0, 1, 0, 1, CCN rows]."""
    c = Case(name)
    g = DeviceGraph(c.A, c.X)
    for pair in (True, False):
        res = precompute(g, c.links, c.num_hops, c.K, 'PoS', 'union', compat_explicit_zero=True, pair=pair)
        rp = res.row_ptr.cpu().numpy()
        assert np.array_equal(rp, c.literal_row_ptr)
        for i in range(c.L):
            a, b = int(rp[i]), int(rp[i + 1])
            gid = c.row_gid[c.row_ptr[i]:c.row_ptr[i + 1]]                      # framework selection: src, dst, CCN ascending
            gid = np.concatenate([gid[:2], gid[:2], gid[2:]])
            order = np.concatenate([[0, 1], 2 + np.argsort(gid[2:], kind='stable')])
            assert np.array_equal(gid[order], c.literal_row_gid[a:b])
            for k in range(c.K + 1):
                got = res.xs[k][a:b].cpu().numpy()[order]
                assert_features_close(got, c.literal_xs[k][a:b], what=f'{name} link {i} x{k} (pair={pair})')
    plain = precompute(g, c.links, c.num_hops, c.K, 'PoS', 'union')
    assert np.array_equal(plain.row_ptr.cpu().numpy(), c.row_ptr)
    with pytest.raises(ValueError):
        precompute(g, c.links, c.num_hops, c.K, 'PoS', 'intersection', compat_explicit_zero=True)


def test_empty_inputs_on_every_entry_point():
    """Edge case: an empty link list (a split without negatives, an empty shard on a rank) goes through every flow."""
    from s3grl_b200 import JointLoader, PrecomputedList, joint_rows, sign_head
    c = Case('tiny_pos_h2')
    g = DeviceGraph(c.A, c.X)
    none = np.zeros((2, 0), dtype=np.int64)
    F1 = c.X.shape[1] + 1
    for flow, strat in (('PoS', None), ('PoS', 'intersection'), ('PoS', 'union'), ('SoP', None)):
        res = precompute(g, none, 2, 3, flow, strat)
        assert res.row_ptr.tolist() == [0] and all(x.shape == (0, F1) for x in res.xs) and len(res.xs) == 4
    full = precompute_full(g, none, 2, 3, node_label='drnl')
    assert full.row_ptr.tolist() == [0] and full.xs[0].shape == (0, F1) and full.node_id.numel() == 0
    res = precompute(g, c.links[:, :5], 2, 3, 'PoS')
    joint, batch, ptr = joint_rows(res.xs, res.row_ptr, torch.zeros(0, dtype=torch.int64, device='cuda'), 2)
    assert joint.shape == (0, 4 * F1) and batch.numel() == 0 and ptr.tolist() == [0]
    assert list(JointLoader(PrecomputedList([x[:0] for x in res.xs], res.row_ptr[:1], torch.zeros(0, dtype=torch.long)).to('cuda'), 8)) == []
    z = torch.zeros
    out = sign_head(z((0, 24), device='cuda'), z((256, 24), device='cuda'), z(256, device='cuda'), z(256, device='cuda'), z(256, device='cuda'))
    assert out.shape == (0, 256)


# ------------------------------------------------------------------------------------------------
# link pairing (csrc/pair.cu): (u,v) and (v,u) share one record, the second one is an exact row swap
# ------------------------------------------------------------------------------------------------
def _pair_reference(links, N):
    """NumPy restatement of the chain table's contract: classes of links over the same unordered node pair."""
    first, members = {}, {}
    for i, (u, v) in enumerate(links.T.tolist()):
        if u < 0 or v < 0 or u >= N or v >= N or u == v:
            continue
        key = (min(u, v), max(u, v))
        if key in first:
            members[first[key]].append(i)
        else:
            first[key] = i
            members[i] = []
    return members


def test_pair_links_chain_table():
    from s3grl_b200.engine import pair_links
    rng = np.random.default_rng(7)
    N = 50
    base = rng.integers(0, N, (2, 400))
    links = np.concatenate([base, base[::-1][:, ::3], base[:, ::5], np.array([[3, -1, N, 9], [3, 4, 5, 9]])], axis=1)
    links = links[:, rng.permutation(links.shape[1])]
    mirror, _table = pair_links(torch.from_numpy(links).cuda(), N)
    m = mirror.cpu().numpy()
    ref = _pair_reference(links, N)
    seen = set()
    for p, mem in ref.items():
        chain, j = [], m[p]
        assert j >= -1, f'link {p} keeps the work'
        while j >= 0:
            v = -2 - m[j]
            assert v >= 0
            assert (v & 1) == int(links[0, j] != links[0, p]), 'swap bit = opposite direction'
            chain.append(int(j))
            j = (v >> 1) - 1
        assert sorted(chain) == mem, f'chain of link {p}'
        seen.update(chain)
        seen.add(p)
    for i in range(links.shape[1]):          # invalid links stay unpaired
        if i not in seen:
            assert m[i] == -1


def test_pairing_is_bit_exact_and_independent_of_batching():
    """PubMed-shaped train positives hold both directions of every edge.  With pairing the reverse link's rows are
    written by the forward link's record; the result must be the same BITS as without pairing, for any batch
    size, and a sub-list (different pairs available) must reproduce the same bits as the full list."""
    edges, N, X = ds.load_graph('cora')
    X = ds.normalize_features(X)
    A, splits = ds.split_links(edges, N, seed=1)
    links = ds.all_links(splits)[:, :6000]
    g = DeviceGraph(A, X)
    plain = precompute(g, links, 3, 3, pair=False)
    paired = precompute(g, links, 3, 3, pair=True)
    assert paired.stats['mirrors'] > 500 and plain.stats['mirrors'] == 0
    assert paired.stats['sum_n_links'] == plain.stats['sum_n'] == plain.stats['sum_n_links']
    assert paired.stats['sum_n'] < plain.stats['sum_n']
    small = precompute(g, links, 3, 3, pair=True, batch_records=700)
    for k in range(4):
        assert torch.equal(plain.xs[k], paired.xs[k]), f'x{k}: pairing changed bits'
        assert torch.equal(plain.xs[k], small.xs[k]), f'x{k}: batch size changed bits'
    sub = np.ascontiguousarray(links[:, 1000:3000])
    part = precompute(g, sub, 3, 3)
    for k in range(4):
        assert torch.equal(part.xs[k], paired.xs[k][2000:6000])
    # (u,v) against (v,u): exact row swap
    rev = precompute(g, np.ascontiguousarray(links[::-1]), 3, 3, pair=False)
    for k in range(4):
        assert torch.equal(plain.xs[k].view(-1, 2, X.shape[1] + 1), rev.xs[k].view(-1, 2, X.shape[1] + 1).flip(1))


def test_pairing_with_host_output_and_overlap():
    """Rows of a paired link are written by an EARLIER batch: the pipelined device-to-host copy of its own batch
    and the two-stream schedule must still see them."""
    edges, N, X = ds.load_graph('cora')
    X = ds.normalize_features(X)
    A, splits = ds.split_links(edges, N, seed=1)
    links = ds.all_links(splits)[:, :5000]
    g = DeviceGraph(A, X)
    want = precompute(g, links, 3, 3, pair=False)
    host = [torch.empty((2 * links.shape[1], X.shape[1] + 1), dtype=torch.float32, pin_memory=True) for _ in range(4)]
    precompute(g, links, 3, 3, host_out=host, batch_records=1024)
    torch.cuda.synchronize()
    over = precompute(g, links, 3, 3, overlap=True, batch_records=1024)
    for k in range(4):
        assert torch.equal(host[k], want.xs[k].cpu())
        assert torch.equal(over.xs[k], want.xs[k])


@pytest.mark.parametrize('name', case_names('hybrid'))
def test_hybrid_against_reference_goldens(name):
    """SURVEY.md §8a row 11 pinned by the reference itself: utils.py:454-480 (sign_type='hybrid') ran through
    oracle/ref_runner.py produced the fixture; here the CUDA path goes through the mirrored dispatcher."""
    from s3grl_b200 import extract_enclosing_subgraphs
    c = Case(name)
    sign_kwargs = dict(sign_k=c.K, use_feature=True, sign_type='hybrid', optimize_sign=True, k_heuristic=0,
                       k_node_set_strategy=None)
    out = extract_enclosing_subgraphs(torch.from_numpy(c.links), c.A, torch.from_numpy(c.X), 1, c.num_hops, 'zo', 1.0, None,
                                      False, None, None, sign_kwargs, powers_of_A=[None] * c.K, data=None)
    assert len(out.xs) == 2 * c.K and np.array_equal(out.row_ptr.cpu().numpy(), c.row_ptr)
    for k in range(2 * c.K):
        assert_features_close(out.xs[k].cpu().numpy(), c.xs[k], what=f'{name} x{k}')


def test_explicit_device_and_output_arguments():
    """device= / output_device= / graph= replace the process-global switches of round 1."""
    from s3grl_b200 import extract_enclosing_subgraphs
    c = Case('cora_pos')
    kw = dict(sign_k=c.K, use_feature=True, sign_type='PoS', optimize_sign=True, k_heuristic=0, k_node_set_strategy=None)
    args = (torch.from_numpy(c.links[:, :20]), c.A, torch.from_numpy(c.X), 1, c.num_hops, 'zo', 1.0, None, False, None, None, kw)
    on_host = extract_enclosing_subgraphs(*args, powers_of_A=[], data=None)
    on_dev = extract_enclosing_subgraphs(*args, powers_of_A=[], data=None, device='cuda:0', output_device='cuda')
    g = DeviceGraph(c.A, c.X)
    with_graph = extract_enclosing_subgraphs(*args, powers_of_A=[], data=None, graph=g, output_device='cuda')
    assert not on_host.xs[0].is_cuda and on_dev.xs[0].is_cuda and with_graph.xs[0].is_cuda
    for k in range(c.K + 1):
        assert torch.equal(on_host.xs[k], on_dev.xs[k].cpu()) and torch.equal(on_dev.xs[k], with_graph.xs[k])


# ------------------------------------------------------------------------------------------------
# per-hop caps (reference utils.py:66-70) with the deterministic rank rule
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('ratio,max_nodes,seed,hops', [(1.0, 5, 0, 3), (0.5, None, 7, 2), (0.7, 12, 3, 3), (1.0, 1, 9, 2),
                                                          (0.3, 40, 1, 3), (1.0, 10000, 0, 2)])
def test_per_hop_caps_bitmap_tier(ratio, max_nodes, seed, hops):
    """Capped BFS on the bitmap tier against the oracle's restatement of the same rule: node lists, hop labels and
    induced edges bit-exact, operators within tolerance; PoS Plus and the non-optimised flow take the same subgraphs."""
    c = Case('cora_pos')
    links = c.links[:, :60]
    caps = dict(ratio_per_hop=ratio, max_nodes_per_hop=max_nodes, cap_seed=seed)
    g = DeviceGraph(c.A, c.X)
    for strategy in (None, 'intersection'):
        ref = orc.pos_precompute(links, hops, c.A, c.X, c.K, strategy, keep_graphs=True, caps=caps)
        res = precompute(g, links, hops, c.K, 'PoS', strategy, return_graphs=True, **caps)
        assert np.array_equal(res.row_ptr.cpu().numpy(), ref['row_ptr'])
        for i, (gg, r) in enumerate(zip(res.graphs, ref['graphs'])):
            _check_indices(gg, r, f'caps {caps} link {i}')
        for k in range(c.K + 1):
            assert_features_close(res.xs[k].cpu().numpy(), ref['xs'][k], what=f'x{k}')
    # without return_graphs (streamed hop-K rows, link pairing) the same numbers come out
    plain = precompute(g, links, hops, c.K, 'PoS', None, **caps)
    ref = orc.pos_precompute(links, hops, c.A, c.X, c.K, None, caps=caps)
    for k in range(c.K + 1):
        assert_features_close(plain.xs[k].cpu().numpy(), ref['xs'][k], what=f'x{k} (hot path)')
    if max_nodes is not None and max_nodes < 100:
        n_max = 2 + hops * max_nodes
        assert plain.stats['max_n'] <= n_max


@pytest.mark.parametrize('ratio,max_nodes,seed', [(1.0, 8, 0), (0.5, None, 5), (0.9, 30, 2), (1.0, 1, 4)])
def test_per_hop_caps_sorted_tier(ratio, max_nodes, seed):
    """The large-graph tier (num_hops = 1) with a cap: the hub rows of an R-MAT-like graph shrink to the cap."""
    rng = np.random.default_rng(300 + seed)
    A = _random_graph(rng, 2000, 30000)
    hub = int(np.argmax(np.diff(A.indptr)))
    X = rng.random((2000, 17), dtype=np.float32)
    links = rng.integers(0, 2000, (2, 80))
    links[0, :8] = hub
    links = links[:, links[0] != links[1]]
    caps = dict(ratio_per_hop=ratio, max_nodes_per_hop=max_nodes, cap_seed=seed)
    ref = orc.pos_precompute(links, 1, A, X, 3, None, keep_graphs=True, caps=caps)
    res = precompute(DeviceGraph(A, X), links, 1, 3, 'PoS', None, return_graphs=True, force_sorted_tier=True, **caps)
    for i, (gg, r) in enumerate(zip(res.graphs, ref['graphs'])):
        _check_indices(gg, r, f'link {i}')
    for k in range(4):
        assert_features_close(res.xs[k].cpu().numpy(), ref['xs'][k], what=f'x{k}')
    both = precompute(DeviceGraph(A, X), links, 1, 3, 'PoS', None, **caps)          # bitmap tier: same subgraphs
    for k in range(4):
        assert_features_close(both.xs[k].cpu().numpy(), ref['xs'][k], what=f'bitmap x{k}')


def test_per_hop_caps_through_the_reference_interface():
    """ratio_per_hop / max_nodes_per_hop of extract_enclosing_subgraphs (utils.py:446) are honoured, not rejected."""
    from s3grl_b200 import extract_enclosing_subgraphs
    c = Case('cora_pos')
    kw = dict(sign_k=c.K, use_feature=True, sign_type='PoS', optimize_sign=True, k_heuristic=0, k_node_set_strategy=None)
    out = extract_enclosing_subgraphs(torch.from_numpy(c.links[:, :30]), c.A, torch.from_numpy(c.X), 1, 3, 'zo', 0.6, 25, False,
                                      None, None, kw, powers_of_A=[], data=None, cap_seed=11)
    ref = orc.pos_precompute(c.links[:, :30], 3, c.A, c.X, c.K, None, caps=dict(ratio_per_hop=0.6, max_nodes_per_hop=25, cap_seed=11))
    for k in range(c.K + 1):
        assert_features_close(out.xs[k].numpy(), ref['xs'][k], what=f'x{k}')


def test_hub_index_changes_nothing_but_the_lookups():
    """The sorted tier's hub x hub bit matrix (s3_build_hub_bits) only short-cuts adjacency look-ups between
    high-degree nodes: same bits with and without it, and the same subgraphs as the oracle."""
    rng = np.random.default_rng(77)
    N = 3000
    A = _random_graph(rng, N, 20000)
    hubs = rng.choice(N, 40, replace=False)                      # 40 hubs, densely connected to everything and each other
    extra = np.stack([np.repeat(hubs, 400), rng.integers(0, N, 40 * 400)], 1)
    hh = np.stack(np.meshgrid(hubs, hubs), -1).reshape(-1, 2)
    hh = hh[rng.random(hh.shape[0]) < 0.5]
    e = np.concatenate([extra, hh])
    e = e[e[:, 0] != e[:, 1]]
    import scipy.sparse as ssp
    B = ssp.csr_matrix((np.ones(e.shape[0], np.int64), (e[:, 0], e[:, 1])), shape=(N, N))
    A = ((A + B + B.T) > 0).astype(np.int64).tocsr()
    A.sort_indices()
    X = rng.random((N, 12), dtype=np.float32)
    links = rng.integers(0, N, (2, 120))
    links[0, :10] = hubs[:10]                                     # hub targets: many hub-hub pairs per subgraph
    links[1, 10:20] = hubs[10:20]
    links = links[:, links[0] != links[1]]
    g_plain = DeviceGraph(A, X)
    g_plain._hub = (None, None, 0)                                # index switched off
    g_hub = DeviceGraph(A, X)
    assert g_hub.ensure_hub_index(max_hubs=64, min_degree=100) >= 40
    a = precompute(g_plain, links, 1, 3, force_sorted_tier=True, return_graphs=True)
    b = precompute(g_hub, links, 1, 3, force_sorted_tier=True, return_graphs=True)
    for k in range(4):
        assert torch.equal(a.xs[k], b.xs[k])
    ref = orc.pos_precompute(links[:, :30], 1, A, X, 3, None, keep_graphs=True)
    for i, r in enumerate(ref['graphs']):
        _check_indices(b.graphs[i], r, f'hub link {i}')


def test_walks_reach_every_neighbour_of_a_hub():
    """the neighbour pick must be uniform for degrees >= 2048 (a 53-bit product overflowed there and hubs only
    ever stepped to their first 2048 neighbours).  Star graph of degree 10 000: one-step walks from the centre, under
    many seeds, must land on late neighbours as often as on early ones."""
    import scipy.sparse as ssp
    from s3grl_b200 import walk_sets
    D = 10000
    row = np.concatenate([np.zeros(D, np.int64), np.arange(1, D + 1)])
    col = np.concatenate([np.arange(1, D + 1), np.zeros(D, np.int64)])
    A = ssp.csr_matrix((np.ones(2 * D, np.int64), (row, col)), shape=(D + 1, D + 1))
    g = DeviceGraph(A, np.zeros((D + 1, 4), np.float32))
    hit = np.zeros(D + 1, np.int64)
    for seed in range(40):
        sets, counts = walk_sets(g, torch.zeros(1, dtype=torch.int64), 1, 200, seed=seed)      # 200 one-step walks
        nodes = sets[0, :int(counts[0])].cpu().numpy()
        hit[nodes] += 1
    leaves = hit[1:]
    assert leaves[2048:].sum() > 0.7 * leaves.sum() * (D - 2048) / D          # late neighbours are reachable, in proportion
    quart = leaves.reshape(4, -1).sum(1)
    assert quart.min() > 0.8 * quart.mean()


@pytest.mark.parametrize('strategy', ['intersection', 'union'])
@pytest.mark.parametrize('seed,N,E,F,K', [(0, 300, 900, 9, 3), (1, 2000, 30000, 33, 3), (3, 64, 80, 4, 2)])
def test_sorted_tier_pos_plus(strategy, seed, N, E, F, K):
    """PoS Plus (tuned_SIGN.py:192-262) on the large-graph tier, forced on small graphs so the oracle can check it:
    selected rows, indices and operators; and the same rows as the bitmap tier produces."""
    rng = np.random.default_rng(400 + seed)
    A = _random_graph(rng, N, E)
    hub = int(np.argmax(np.diff(A.indptr)))
    X = rng.random((N, F), dtype=np.float32)
    links = rng.integers(0, N, (2, 50))
    links[0, :5] = hub
    links = links[:, links[0] != links[1]]
    e = np.stack(A.nonzero(), 1)
    links = np.concatenate([links, e[rng.choice(e.shape[0], 10, replace=False)].T], axis=1)     # true edges: the mask matters
    ref = orc.pos_precompute(links, 1, A, X, K, strategy, keep_graphs=True)
    g = DeviceGraph(A, X)
    res = precompute(g, links, 1, K, 'PoS', strategy, return_graphs=True, force_sorted_tier=True)
    assert np.array_equal(res.row_ptr.cpu().numpy(), ref['row_ptr'])
    for i, (gg, r) in enumerate(zip(res.graphs, ref['graphs'])):
        _check_indices(gg, r, f'{strategy} link {i}')
    for k in range(K + 1):
        assert_features_close(res.xs[k].cpu().numpy(), ref['xs'][k], what=f'{strategy} x{k}')
    bitmap = precompute(g, links, 1, K, 'PoS', strategy)
    assert torch.equal(bitmap.row_ptr, res.row_ptr)
    for k in range(K + 1):
        # union: `res` (parity dump) took the CCN work items, `bitmap` the SpMM chain — another summation order
        assert_features_close(res.xs[k].cpu().numpy(), bitmap.xs[k].cpu().numpy(), tol=1e-5 if strategy == 'union' else 2e-6,
                              what=f'tiers {strategy} x{k}')
    plain = precompute(g, links, 1, K, 'PoS', strategy, force_sorted_tier=True)      # union: the chain over the tier's CSR
    assert torch.equal(plain.row_ptr, res.row_ptr)
    for k in range(K + 1):
        assert_features_close(plain.xs[k].cpu().numpy(), ref['xs'][k], what=f'sorted tier, no dump, {strategy} x{k}')
    capped = precompute(g, links, 1, K, 'PoS', strategy, force_sorted_tier=True, max_nodes_per_hop=6, cap_seed=2)
    refc = orc.pos_precompute(links, 1, A, X, K, strategy, caps=dict(max_nodes_per_hop=6, cap_seed=2))
    assert np.array_equal(capped.row_ptr.cpu().numpy(), refc['row_ptr'])
    for k in range(K + 1):
        assert_features_close(capped.xs[k].cpu().numpy(), refc['xs'][k], what=f'capped {strategy} x{k}')
