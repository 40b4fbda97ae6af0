"""Run the REAL reference functions (imported, unmodified, from /root/reference behind the
stand-ins in oracle/ref_stub) and canonicalise their outputs.  TEST INFRASTRUCTURE ONLY and
build-container only: /root/reference does not exist on the GPU box, so nothing under
`-m gpu`, smoke() or bench.py may import this module.  Used by oracle/make_goldens.py and by
tests/test_oracle_vs_reference.py (skipped when the reference is absent).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get('S3GRL_REFERENCE', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE, 'tuned_SIGN.py'))


_mods = None


def load_reference():
    """-> (utils, tuned_SIGN) modules of the reference."""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError(f"reference not found at {REFERENCE}")
        for p in (REFERENCE, os.path.join(_HERE, 'ref_stub')):
            if p not in sys.path:
                sys.path.insert(0, p)
        import tqdm as _tqdm_mod  # silence the reference's progress bars
        import tuned_SIGN
        import utils
        quiet = lambda it=None, *a, **k: it  # noqa: E731
        tuned_SIGN.tqdm = quiet
        utils.tqdm = quiet
        _mods = (utils, tuned_SIGN)
    return _mods


@contextlib.contextmanager
def _quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


class _RandomWithRankRule:
    """Stand-in for the name `random` inside the reference's utils module while its per-hop caps are exercised.
    utils.py:66-70 draws the nodes a hop keeps with `random.sample(fringe, k)` on a SET — no reproducible semantics, and a
    TypeError on Python >= 3.11.  This proxy forwards everything to `random` and makes `sample` return the k elements of
    the population with the smallest fmix32(node XOR seed): the framework's deterministic rule (include/s3grl_b200.h,
    oracle.hash_rank).  Everything else of the capped BFS — the counts int(ratio * len) and max_nodes_per_hop, their
    order, that dropped nodes stay visited, the early exit — is the reference's own code."""

    def __init__(self, real, seed):
        self._real, self._seed = real, int(seed)
        self.calls = 0

    def __getattr__(self, name):
        return getattr(self._real, name)

    def sample(self, population, k):
        from oracle import s3grl_oracle as orc
        self.calls += 1
        pop = np.asarray(sorted(int(v) for v in population), dtype=np.int64)
        if k > pop.size:
            raise ValueError("Sample larger than population or is negative")
        keep = np.argsort(orc.hash_rank(pop, self._seed), kind='stable')[:k]
        return [int(v) for v in pop[keep]]


@contextlib.contextmanager
def cap_sampler_ranked(seed=0):
    """Within the block the reference's `random.sample` follows the rank rule (see _RandomWithRankRule)."""
    utils, _ = load_reference()
    real = utils.random
    proxy = _RandomWithRankRule(real, seed)
    utils.random = proxy
    try:
        yield proxy
    finally:
        utils.random = real


def ref_k_hop(src, dst, num_hops, A, ratio_per_hop=1.0, max_nodes_per_hop=None, cap_seed=0):
    """Reference k_hop_subgraph (utils.py:47-85) -> canonical
    (nodes int64 [n], hops int32 [n], edges int64 [m,2] of GLOBAL ids sorted by canonical
    (local row, local col)).  Edges are the non-zero entries `ssp.find` returns
    (tuned_SIGN.py:153), i.e. after the target-link mask.  With caps the reference's sampler is the rank rule
    (cap_sampler_ranked); without, the unmodified code runs."""
    import scipy.sparse as ssp
    utils, _ = load_reference()
    capped = (ratio_per_hop is not None and ratio_per_hop < 1.0) or max_nodes_per_hop is not None
    with (cap_sampler_ranked(cap_seed) if capped else contextlib.nullcontext()):
        nodes, sub, dists, _, _ = utils.k_hop_subgraph(src, dst, num_hops, A, ratio_per_hop, max_nodes_per_hop)
    nodes = np.asarray(nodes, dtype=np.int64)
    dists = np.asarray(dists, dtype=np.int32)
    order = np.concatenate([[0, 1], 2 + np.lexsort((nodes[2:], dists[2:]))]).astype(np.int64)
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    u, v, _ = ssp.find(sub)
    lu, lv = rank[u], rank[v]
    eo = np.lexsort((lv, lu))
    cn = nodes[order]
    edges = np.stack([cn[lu[eo]], cn[lv[eo]]], axis=1) if eo.size else np.zeros((0, 2), np.int64)
    return cn, dists[order], edges


class _TorchWithLabelColumnRepaired:
    """Stand-in for the name `torch` inside the reference's tuned_SIGN module while a `union` golden is generated.
    tuned_SIGN.py:243 builds the zero-one label column of the union branch from the ragged literal
    `[[1]] + [[1]] + [[0] * (n - 2)]` — one inner list of n - 2 zeros where the intersection branch four lines below
    (`[[0]] * (n - 2)`: n - 2 inner lists) shows what was meant — so `torch.tensor` raises ValueError for every
    subgraph that does not have exactly 3 nodes.  This proxy forwards everything to torch and only makes `tensor()` of
    exactly that literal return the intended (n, 1) column.  /root/reference itself is never modified; fixtures made
    this way carry `reference_repair` in the file and "union" in their name."""

    def __init__(self, real):
        self._real = real
        self.repairs = 0

    def __getattr__(self, name):
        return getattr(self._real, name)

    def tensor(self, data, *args, **kwargs):
        if (isinstance(data, list) and len(data) == 3 and data[0] == [1] and data[1] == [1] and isinstance(data[2], list)
                and len(data[2]) != 1 and all(v == 0 for v in data[2])):
            self.repairs += 1
            return self._real.tensor([[1], [1]] + [[0]] * len(data[2]), *args, **kwargs)
        return self._real.tensor(data, *args, **kwargs)


@contextlib.contextmanager
def union_typo_repaired():
    """Within the block the reference's PoS Plus `union` branch builds its label column as its `intersection` branch does
    (see _TorchWithLabelColumnRepaired).  Yields the proxy (its `repairs` counts the literals it replaced)."""
    _, tuned = load_reference()
    real = tuned.torch
    proxy = _TorchWithLabelColumnRepaired(real)
    tuned.torch = proxy
    try:
        yield proxy
    finally:
        tuned.torch = real


def _sign_kwargs(K, strategy):
    return {'sign_k': K, 'use_feature': True, 'sign_type': 'PoS', 'optimize_sign': True,
            'k_heuristic': 0 if strategy is None else 1, 'k_node_set_strategy': strategy}


def ref_pos(links, num_hops, A, X, K, strategy=None, repair_union_typo=False, caps=None):
    """get_PoS_prepped_ds / get_PoS_Plus_prepped_ds (tuned_SIGN.py:137-262) through
    extract_enclosing_subgraphs (utils.py:446-496) -> dict(xs, row_ptr, row_gid) with the rows
    of every link put in canonical order: row 0 = src, row 1 = dst, extra rows by ascending
    global id (== ascending canonical local id, all extra rows being hop-1 nodes).
    strategy='union' runs the UNMODIFIED reference unless repair_union_typo=True (union_typo_repaired above): unmodified,
    it raises ValueError at tuned_SIGN.py:243 for every link whose subgraph does not have exactly 3 nodes.
    caps = dict(ratio_per_hop=, max_nodes_per_hop=, cap_seed=): the reference's per-hop caps with its sampler replaced by
    the rank rule (cap_sampler_ranked); PoS only (the row recovery of PoS Plus below re-runs an uncapped BFS)."""
    utils, tuned = load_reference()
    caps = dict(caps or {})
    ratio, max_nodes = caps.get('ratio_per_hop', 1.0), caps.get('max_nodes_per_hop')
    if caps and strategy is not None:
        raise NotImplementedError("caps with PoS Plus are not wired through ref_pos")
    link_index = torch.as_tensor(np.asarray(links), dtype=torch.long)
    x = torch.as_tensor(np.asarray(X), dtype=torch.float32)
    ctx = union_typo_repaired() if (repair_union_typo and strategy == 'union') else contextlib.nullcontext()
    ctx2 = cap_sampler_ranked(caps.get('cap_seed', 0)) if caps else contextlib.nullcontext()
    with _quiet(), ctx, ctx2:
        data_list = utils.extract_enclosing_subgraphs(
            link_index, A, x, 1, num_hops, 'zo', ratio, max_nodes, False, None, None,
            _sign_kwargs(K, strategy), powers_of_A=[], data=None)
    xs = [[] for _ in range(K + 1)]
    row_ptr, row_gid = [0], []
    for i, d in enumerate(data_list):
        src, dst = int(links[0][i]), int(links[1][i])
        s = d['x'].shape[0]
        gids = [src, dst]
        if strategy is not None:
            # recover which node each extra row is: same calls, same (deterministic) set order
            nodes, sub, _, _, _ = utils.k_hop_subgraph(src, dst, num_hops, A)
            if strategy == 'intersection':
                sel = utils.neighbors({0}, sub).intersection(utils.neighbors({1}, sub))
            else:
                sel = utils.neighbors({0}, sub).union(utils.neighbors({1}, sub))
            gids += [int(nodes[j]) for j in list(sel)]
        assert len(gids) == s
        gids = np.asarray(gids, dtype=np.int64)
        order = np.concatenate([[0, 1], 2 + np.argsort(gids[2:], kind='stable')]).astype(np.int64)
        row_gid.append(gids[order])
        keys = ['x'] + [f'x{k}' for k in range(1, K + 1)]
        for k, key in enumerate(keys):
            xs[k].append(d[key].numpy().astype(np.float32)[order])
        row_ptr.append(row_ptr[-1] + s)
    F1 = x.shape[1] + 1
    xs = [np.concatenate(v, 0) if v else np.zeros((0, F1), np.float32) for v in xs]
    return dict(xs=xs, row_ptr=np.asarray(row_ptr, np.int64),
                row_gid=np.concatenate(row_gid) if row_gid else np.zeros(0, np.int64))


def ref_sop_powers(A, K):
    """Global normalised powers exactly as SEALDataset.process builds them
    (sgrl_link_pred.py:161-178; that file cannot be imported without real PyG, so its ten
    lines of SparseTensor calls are issued here against the same stand-in)."""
    load_reference()
    from torch_sparse import SparseTensor
    coo = A.tocoo()
    # edge_index holds one column per stored edge; multiplicity -> repeated columns
    rep = np.asarray(coo.data, dtype=np.int64)
    row = torch.as_tensor(np.repeat(coo.row, rep), dtype=torch.long)
    col = torch.as_tensor(np.repeat(coo.col, rep), dtype=torch.long)
    adj_t = SparseTensor(row=row, col=col, sparse_sizes=A.shape)
    deg = adj_t.sum(dim=1).to(torch.float)
    dis = deg.pow(-0.5)
    dis[dis == float('inf')] = 0
    adj_t = dis.view(-1, 1) * adj_t * dis.view(1, -1)
    powers = [adj_t]
    for _ in range(2, K + 1):
        powers += [adj_t @ powers[-1]]
    return powers


def ref_sop(links, A, X, K):
    """get_SoP_prepped_ds (tuned_SIGN.py:49-134) -> dict(xs, row_ptr)."""
    utils, tuned = load_reference()
    link_index = torch.as_tensor(np.asarray(links), dtype=torch.long)
    x = torch.as_tensor(np.asarray(X), dtype=torch.float32)
    powers = ref_sop_powers(A, K)
    with _quiet():
        data_list = tuned.OptimizedSignOperations.get_SoP_prepped_ds(powers, link_index, A, x, 1)
    keys = ['x'] + [f'x{k}' for k in range(1, K + 1)]
    xs = [np.concatenate([np.asarray(d[key], dtype=np.float32) for d in data_list], 0) for key in keys]
    return dict(xs=xs, row_ptr=np.arange(len(data_list) + 1, dtype=np.int64) * 2)


def ref_hybrid(links, num_hops, A, X, K):
    """The hybrid flow (utils.py:454-480) through the reference's own dispatcher: PoS x, x1..xK followed by the SoP
    operators x2..xK stored as x{K+1}..x{2K-1} -> dict(xs (2K operators), row_ptr)."""
    utils, tuned = load_reference()
    link_index = torch.as_tensor(np.asarray(links), dtype=torch.long)
    x = torch.as_tensor(np.asarray(X), dtype=torch.float32)
    kw = {'sign_k': K, 'use_feature': True, 'sign_type': 'hybrid', 'optimize_sign': True, 'k_heuristic': 0,
          'k_node_set_strategy': None}
    powers = ref_sop_powers(A, K)
    with _quiet():
        data_list = utils.extract_enclosing_subgraphs(link_index, A, x, 1, num_hops, 'zo', 1.0, None, False, None, None, kw,
                                                      powers_of_A=powers, data=None)
    keys = ['x'] + [f'x{k}' for k in range(1, 2 * K)]
    xs = [np.concatenate([np.asarray(d[key], dtype=np.float32) for d in data_list], 0) for key in keys]
    return dict(xs=xs, row_ptr=np.arange(len(data_list) + 1, dtype=np.int64) * 2)


def ref_scaled_pos(links, A, X, K, sets):
    """ScaLed through the reference's own code: get_PoS_prepped_ds with rw_kwargs carrying a walk
    cache (utils.py:94-105 picks cached_pos_rws for y = 1) -> dict(xs, row_ptr).  `sets` is
    {node: array}; the walks themselves are an input (the reference draws them with
    torch_cluster.random_walk, which is not available here)."""
    utils, tuned = load_reference()
    from torch_geometric.data import Data
    link_index = torch.as_tensor(np.asarray(links), dtype=torch.long)
    x = torch.as_tensor(np.asarray(X), dtype=torch.float32)
    cache = {int(k): torch.as_tensor(v, dtype=torch.long) for k, v in sets.items()}
    data = Data(x=x, num_nodes=A.shape[0])
    rw_kwargs = dict(rw_m=3, rw_M=20, sparse_adj=None, edge_index=None, device='cpu', data=data, node_label='zo',
                     cached_pos_rws=cache, cached_neg_rws=cache, sign=True)
    with _quiet():
        data_list = tuned.OptimizedSignOperations.get_PoS_prepped_ds(
            link_index, 0, A, 1.0, None, False, None, x, 1, _sign_kwargs(K, None), rw_kwargs)
    keys = ['x'] + [f'x{k}' for k in range(1, K + 1)]
    xs = [np.concatenate([np.asarray(d[key], dtype=np.float32) for d in data_list], 0) for key in keys]
    return dict(xs=xs, row_ptr=np.arange(len(data_list) + 1, dtype=np.int64) * 2)


def ref_full(links, num_hops, A, X, K, node_label):
    """The reference's NON-optimised PoS flow (utils.py:497-520: k_hop_subgraph -> construct_pyg_graph ->
    TunedSIGN) through extract_enclosing_subgraphs with optimize_sign=False -> dict(xs, row_ptr, node_id)
    with every link's rows permuted into canonical node order (src, dst, ascending (hop, global id))."""
    utils, tuned = load_reference()
    link_index = torch.as_tensor(np.asarray(links), dtype=torch.long)
    x = torch.as_tensor(np.asarray(X), dtype=torch.float32)
    kw = {'sign_k': K, 'use_feature': True, 'sign_type': 'PoS', 'optimize_sign': False, 'k_heuristic': 0,
          'k_node_set_strategy': None}
    with _quiet():
        data_list = utils.extract_enclosing_subgraphs(link_index, A, x, 1, num_hops, node_label, 1.0, None, False, None,
                                                      None, kw, powers_of_A=[], data=None)
    keys = ['x'] + [f'x{k}' for k in range(1, K + 1)]
    xs = [[] for _ in keys]
    row_ptr, node_id = [0], []
    for i, d in enumerate(data_list):
        src, dst = int(links[0][i]), int(links[1][i])
        nodes = d['node_id'].numpy().astype(np.int64)
        # hop of every node: the same call again (deterministic) for its dists list
        nodes2, _, dists, _, _ = utils.k_hop_subgraph(src, dst, num_hops, A)
        assert list(nodes2) == list(nodes)
        dists = np.asarray(dists)
        order = np.concatenate([[0, 1], 2 + np.lexsort((nodes[2:], dists[2:]))]).astype(np.int64)
        node_id.append(nodes[order])
        for k, key in enumerate(keys):
            xs[k].append(np.asarray(d[key], dtype=np.float32)[order])
        row_ptr.append(row_ptr[-1] + nodes.size)
    return dict(xs=[np.concatenate(v, 0) for v in xs], row_ptr=np.asarray(row_ptr, np.int64),
                node_id=np.concatenate(node_id))
