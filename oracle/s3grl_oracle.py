"""CPU oracle for S3GRL's precompute hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the product package `s3grl_b200` never does (its CUDA path
fails loudly when the extension is missing).

It restates, in NumPy/SciPy, the algorithm of the reference's hot path, with a CANONICAL
node order so results can be compared bit-exactly on indices:

  reference                                   restated here
  ------------------------------------------  ---------------------------------------
  utils.py:33-44   neighbors                  `_neighbors`
  utils.py:47-85   k_hop_subgraph (BFS)       `k_hop_subgraph`
  tuned_SIGN.py:137-189 get_PoS_prepped_ds    `pos_link(..., strategy=None)`
  tuned_SIGN.py:192-262 get_PoS_Plus_prepped_ds  `pos_link(..., strategy='intersection'|'union')`
  sgrl_link_pred.py:161-178 SoP global powers `sop_powers`
  tuned_SIGN.py:49-134  get_SoP_prepped_ds    `sop_link`
  utils.py:454-480 hybrid                     `hybrid_precompute`
  utils.py:497-520 non-optimised PoS flow     `sign_all_rows`, `full_precompute`, `node_labels`
  utils.py:86-150, :425-443 ScaLed walks      `random_walk_sets`, `scaled_pos_link`
  models.py:372    joint matrix layout        `joint_matrix`

Pinning (see tests/test_oracle_vs_reference.py, oracle/make_goldens.py): the reference has
NO tests or golden vectors for this path, so the oracle is pinned
against outputs of the reference's own functions, imported unmodified from /root/reference
behind the stand-in packages in oracle/ref_stub, run in the build container and committed as
tests/golden/ref_*.npz.  The `union` strategy: the unmodified reference raises ValueError at
tuned_SIGN.py:243 (a ragged literal for the label column) unless the subgraph has exactly 3 nodes.
It is pinned (a) unmodified, on such 3-node subgraphs (tests/test_oracle_vs_reference.py), and
(b) against the reference with that ONE literal repaired at run time
(oracle/ref_runner.union_typo_repaired; fixtures ref_*_union*.npz, plus the live hypothesis test).
The repaired reference selects src and dst a second time among the extra rows (its target-link
mask leaves explicit zeros that `neighbors` reports); the rule here is the paper's,
[0, 1] + (N(0) ∪ N(1)) − {0, 1}, and the tests check that the two additional reference rows are
bit-identical copies of rows 0 / 1 before dropping them (tests/golden_util.drop_duplicated_seed_rows).

Canonical order: nodes = [src, dst] then ascending (hop, global id);
induced edges sorted by (local row, local col); extra selected rows ascending local id.
The reference's order of nodes 2.. is CPython set-iteration order, so comparisons with it are
made after mapping through global ids.
"""
import numpy as np
import scipy.sparse as ssp

# --------------------------------------------------------------------------------------
# graph helpers
# --------------------------------------------------------------------------------------


def csr_from_edges(edges_undirected, num_nodes):
    """Both directions of every undirected edge, values = multiplicity, as the reference
    builds `A` at sgrl_link_pred.py:107-114 from `data.edge_index`."""
    e = np.asarray(edges_undirected, dtype=np.int64)
    row = np.concatenate([e[:, 0], e[:, 1]])
    col = np.concatenate([e[:, 1], e[:, 0]])
    A = ssp.csr_matrix((np.ones(row.shape[0], dtype=np.int64), (row, col)),
                       shape=(num_nodes, num_nodes))
    A.sum_duplicates()
    A.sort_indices()
    return A


def _neighbors(fringe, indptr, indices):
    """utils.py:33-44 — union of the stored column ids of the fringe rows."""
    if len(fringe) == 0:
        return np.empty(0, dtype=np.int64)
    parts = [indices[indptr[v]:indptr[v + 1]] for v in fringe]
    return np.unique(np.concatenate(parts)) if parts else np.empty(0, dtype=np.int64)


def hash_rank(nodes, seed=0):
    """Rank key of the deterministic per-hop cap: murmur3's 32-bit finaliser of (node XOR seed).  It is a
    bijection of the 32-bit ids, so distinct nodes never tie."""
    h = (np.asarray(nodes, dtype=np.uint64) ^ np.uint64(seed & 0xFFFFFFFF)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    return h


def cap_fringe(fringe, ratio_per_hop, max_nodes_per_hop, seed=0):
    """utils.py:66-70 with a DETERMINISTIC rule in place of random.sample (which has no reproducible semantics and
    raises on Python >= 3.11): the hop keeps k = min(int(ratio * len), max) nodes — the reference's
    counts — and they are the k nodes of the fringe with the smallest hash_rank.  Returned ascending.
    Pinned by tests/test_oracle_vs_reference.py against the reference's own capped BFS with its `random.sample` replaced
    by this rule (oracle/ref_runner.cap_sampler_ranked)."""
    k = fringe.size
    if ratio_per_hop is not None and ratio_per_hop < 1.0:
        k = int(ratio_per_hop * fringe.size)
    if max_nodes_per_hop is not None and max_nodes_per_hop < k:
        k = int(max_nodes_per_hop)
    if k >= fringe.size:
        return fringe
    keep = np.argsort(hash_rank(fringe, seed), kind='stable')[:k]
    return np.sort(fringe[keep])


def k_hop_subgraph(src, dst, num_hops, A, ratio_per_hop=1.0, max_nodes_per_hop=None, cap_seed=0):
    """utils.py:47-85 (undirected, no random walks).  Per-hop caps follow `cap_fringe`; as in the reference the
    nodes a cap drops stay `visited` (utils.py:64-65 runs before the sampling) and never come back at a later hop.

    Returns (nodes[n] int64 canonical, hops[n] int32, lrowptr[n+1] int64, lcol[m] int32) where
    (lrowptr, lcol) is the induced adjacency on `nodes` in local ids with the target link
    (0,1)/(1,0) removed and columns ascending.  Stored values are ignored, as
    tuned_SIGN.py:153 discards them and SciPy's explicit zeros at (0,1),(1,0)
    (utils.py:79-80) are dropped by `ssp.find`."""
    if src == dst:
        raise ValueError("src == dst is not a valid target link (SURVEY.md A.2)")
    indptr, indices = A.indptr, A.indices
    nodes = [np.array([src, dst], dtype=np.int64)]
    hops = [np.zeros(2, dtype=np.int32)]
    visited = np.zeros(A.shape[0], dtype=bool)
    visited[[src, dst]] = True
    fringe = np.array([src, dst], dtype=np.int64)
    for dist in range(1, num_hops + 1):
        cand = _neighbors(fringe, indptr, indices)
        fringe = cand[~visited[cand]]              # np.unique output is ascending
        visited[fringe] = True                     # before the cap: utils.py:64-65
        fringe = cap_fringe(fringe, ratio_per_hop, max_nodes_per_hop, cap_seed)
        if fringe.size == 0:
            break
        nodes.append(fringe.astype(np.int64))
        hops.append(np.full(fringe.size, dist, dtype=np.int32))
    nodes = np.concatenate(nodes)
    hops = np.concatenate(hops)
    lrowptr, lcol = _induced_masked(nodes, A)
    return nodes, hops, lrowptr, lcol


def _induced_masked(nodes, A):
    """utils.py:76-80 — A[nodes,:][:,nodes] with the target link (0,1)/(1,0) removed, as a local CSR
    with ascending columns (pattern only)."""
    indptr, indices = A.indptr, A.indices
    n = nodes.size
    local = np.full(A.shape[0], -1, dtype=np.int64)
    local[nodes] = np.arange(n)
    lrowptr = np.zeros(n + 1, dtype=np.int64)
    cols = []
    for i, g in enumerate(nodes):
        c = local[indices[indptr[g]:indptr[g + 1]]]
        c = np.unique(c[c >= 0])                   # unweighted: one entry per (row, col)
        if i == 0:
            c = c[c != 1]
        elif i == 1:
            c = c[c != 0]
        cols.append(c)
        lrowptr[i + 1] = lrowptr[i] + c.size
    lcol = np.concatenate(cols).astype(np.int32) if cols else np.empty(0, np.int32)
    return lrowptr, lcol


def select_rows(lrowptr, lcol, strategy, compat_explicit_zero=False):
    """Row selection, tuned_SIGN.py:173 (PoS) and :228-238 (PoS Plus).
    strategy None -> [0,1]; 'intersection' -> [0,1] + common neighbours of local 0 and 1 in
    the masked subgraph; 'union' -> [0,1] + (N(0) ∪ N(1)) − {0,1} (paper semantics; the
    reference code raises for union, SURVEY.md A.4, and with its literal repaired additionally repeats
    rows 0 and 1 — module docstring; compat_explicit_zero=True restates that literal selection,
    [0,1] + sorted({0,1} ∪ N(0) ∪ N(1)): utils.py:78-79 leave explicit zeros at [0,1] and [1,0], which
    `neighbors` reports).  Extra rows ascending local id."""
    if strategy is None:
        return np.array([0, 1], dtype=np.int32)
    n0 = lcol[lrowptr[0]:lrowptr[1]]
    n1 = lcol[lrowptr[1]:lrowptr[2]]
    if strategy == 'intersection':
        extra = np.intersect1d(n0, n1)
    elif strategy == 'union':
        extra = np.union1d(n0, n1)
    else:
        raise NotImplementedError(f"check strat {strategy}")
    extra = extra[extra > 1]
    if compat_explicit_zero and strategy == 'union':
        extra = np.concatenate([[0, 1], extra])
    return np.concatenate([[0, 1], extra]).astype(np.int32)


def normalized_subgraph(lrowptr, lcol, dtype=np.float32):
    """tuned_SIGN.py:155-161 — S = D^-1/2 A_sub D^-1/2, deg = stored entries per row, inf -> 0."""
    n = lrowptr.size - 1
    deg = np.diff(lrowptr).astype(dtype)
    with np.errstate(divide='ignore'):
        dis = np.power(deg, dtype(-0.5), dtype=dtype)
    dis[np.isinf(dis)] = 0
    rows = np.repeat(np.arange(n), np.diff(lrowptr))
    vals = (dis[rows] * dis[lcol]).astype(dtype)
    return ssp.csr_matrix((vals, lcol, lrowptr), shape=(n, n)), dis


def pos_link(src, dst, num_hops, A, X, K, strategy=None, dtype=np.float32, caps=None, compat_explicit_zero=False):
    """One link through the optimised PoS / PoS-Plus flow (tuned_SIGN.py:147-187, :202-260).
    caps: None or dict(ratio_per_hop=, max_nodes_per_hop=, cap_seed=) for k_hop_subgraph.

    Returns dict: nodes, hops, lrowptr, lcol, sel (local ids), xs = [x, x1..xK] each [s, F+1]."""
    nodes, hops, lrowptr, lcol = k_hop_subgraph(src, dst, num_hops, A, **(caps or {}))
    n = nodes.size
    S, _ = normalized_subgraph(lrowptr, lcol, dtype)
    sel = select_rows(lrowptr, lcol, strategy, compat_explicit_zero)
    label = np.zeros((n, 1), dtype=dtype)
    label[:2] = 1                                   # zero-one label, tuned_SIGN.py:177
    subg_x = np.hstack([label, np.asarray(X[nodes], dtype=dtype)])
    xs = [subg_x[sel]]
    P = S
    for k in range(1, K + 1):
        if k > 1:
            P = (S @ P).astype(dtype)               # powers by SpGEMM, tuned_SIGN.py:168-170
        xs.append(np.asarray(P[sel] @ subg_x, dtype=dtype))
    return dict(nodes=nodes, hops=hops, lrowptr=lrowptr, lcol=lcol, sel=sel, xs=xs)


def random_walk_sets(A, starts, rw_m, rw_M, seed=0):
    """utils.py:425-443 (create_rw_cache): for every start node the sorted set of nodes visited by
    rw_M uniform random walks of length rw_m (torch_cluster.random_walk, p = q = 1; a node without
    neighbours stays put).  NumPy RNG: same distribution as the reference and as the CUDA kernel,
    not the same stream.  Returns {node: int64 array}."""
    rng = np.random.default_rng(seed)
    indptr, indices = A.indptr, A.indices
    out = {}
    for s in np.unique(np.asarray(starts)):
        seen = {int(s)}
        for _ in range(rw_M):
            cur = int(s)
            for _ in range(rw_m):
                d = indptr[cur + 1] - indptr[cur]
                if d > 0:
                    cur = int(indices[indptr[cur] + rng.integers(0, d)])
                seen.add(cur)
        out[int(s)] = np.array(sorted(seen), dtype=np.int64)
    return out


def scaled_pos_link(src, dst, sets, A, X, K, dtype=np.float32):
    """ScaLed + optimised PoS (utils.py:94-150 with rw_kwargs['sign'], then tuned_SIGN.py:153-187):
    nodes = [src, dst] + ascending(set(src) ∪ set(dst) − {src, dst}); induced, masked subgraph; the
    usual diffusion with rows [0, 1]."""
    rest = np.union1d(sets[int(src)], sets[int(dst)])
    rest = rest[(rest != src) & (rest != dst)]
    nodes = np.concatenate([[src, dst], rest]).astype(np.int64)
    lrowptr, lcol = _induced_masked(nodes, A)
    S, _ = normalized_subgraph(lrowptr, lcol, dtype)
    label = np.zeros((nodes.size, 1), dtype=dtype)
    label[:2] = 1
    subg_x = np.hstack([label, np.asarray(X[nodes], dtype=dtype)])
    xs = [subg_x[:2]]
    P = S
    for k in range(1, K + 1):
        if k > 1:
            P = (S @ P).astype(dtype)
        xs.append(np.asarray(P[:2] @ subg_x, dtype=dtype))
    hops = np.r_[0, 0, np.ones(nodes.size - 2, dtype=np.int32)].astype(np.int32)
    return dict(nodes=nodes, hops=hops, lrowptr=lrowptr, lcol=lcol, sel=np.array([0, 1], np.int32), xs=xs)


def scaled_pos_precompute(links, sets, A, X, K, dtype=np.float32, keep_graphs=False):
    links = np.asarray(links)
    per = [scaled_pos_link(int(links[0, i]), int(links[1, i]), sets, A, X, K, dtype) for i in range(links.shape[1])]
    xs = [np.concatenate([p['xs'][k] for p in per], 0) for k in range(K + 1)]
    out = dict(xs=xs, row_ptr=np.arange(links.shape[1] + 1, dtype=np.int64) * 2)
    if keep_graphs:
        out['graphs'] = per
    return out


def pos_precompute(links, num_hops, A, X, K, strategy=None, dtype=np.float32, keep_graphs=False, caps=None,
                   compat_explicit_zero=False):
    """Whole call of get_PoS_prepped_ds / get_PoS_Plus_prepped_ds over `links` [2, L], in the
    collated layout PyG's InMemoryDataset.collate produces:
    K+1 row-stacked [R, F+1] arrays and row_ptr [L+1]."""
    links = np.asarray(links)
    L = links.shape[1]
    per = [pos_link(int(links[0, i]), int(links[1, i]), num_hops, A, X, K, strategy, dtype, caps, compat_explicit_zero)
           for i in range(L)]
    row_ptr = np.zeros(L + 1, dtype=np.int64)
    row_ptr[1:] = np.cumsum([p['sel'].size for p in per])
    F1 = X.shape[1] + 1
    xs = [np.concatenate([p['xs'][k] for p in per], 0) if L else np.zeros((0, F1), dtype)
          for k in range(K + 1)]
    out = dict(xs=xs, row_ptr=row_ptr)
    if keep_graphs:
        out['graphs'] = per
    return out


# --------------------------------------------------------------------------------------
# SoP
# --------------------------------------------------------------------------------------


def sop_powers(A, K, dtype=np.float32):
    """sgrl_link_pred.py:161-178 — Â = D^-1/2 A D^-1/2 on the whole training graph, powers
    Â^1..Â^K by SpGEMM.  `edge_index` is given to SparseTensor without values, so a duplicated
    edge is stored twice: deg counts it twice and the products sum it twice — i.e. A carries
    the multiplicity, which is exactly the `data` of the SciPy `A` (sgrl_link_pred.py:111)."""
    n = A.shape[0]
    Au = ssp.csr_matrix((A.data.astype(dtype), A.indices, A.indptr), shape=A.shape)
    deg = np.asarray(A.sum(axis=1)).reshape(-1).astype(dtype)   # duplicates counted
    with np.errstate(divide='ignore'):
        dis = np.power(deg, dtype(-0.5), dtype=dtype)
    dis[np.isinf(dis)] = 0
    Ahat = (ssp.diags(dis).astype(dtype) @ Au @ ssp.diags(dis).astype(dtype)).astype(dtype).tocsr()
    powers = [Ahat]
    for _ in range(2, K + 1):
        powers.append((Ahat @ powers[-1]).astype(dtype).tocsr())
    assert powers[0].shape == (n, n)
    return powers


def sop_link(src, dst, powers, X, dtype=np.float32):
    """tuned_SIGN.py:60-133 for one link: x = [[1, X[u]], [1, X[v]]];
    x_k = [[Â^k[u,u], r_u·X], [Â^k[v,v], r_v·X]] with r_u = Â^k[u,:] minus entry v, r_v likewise."""
    X = np.asarray(X, dtype=dtype)
    xs = [np.hstack([np.ones((2, 1), dtype=dtype), X[[src, dst]]])]
    for P in powers:
        rows = []
        for a, b in ((src, dst), (dst, src)):
            r = P.getrow(a).toarray().astype(dtype).reshape(-1)
            self_w = r[a]
            r[b] = 0
            rows.append(np.concatenate([[self_w], r @ X]).astype(dtype))
        xs.append(np.stack(rows))
    return xs


def sop_precompute(links, A, X, K, dtype=np.float32, powers=None):
    links = np.asarray(links)
    L = links.shape[1]
    powers = sop_powers(A, K, dtype) if powers is None else powers
    per = [sop_link(int(links[0, i]), int(links[1, i]), powers, X, dtype) for i in range(L)]
    F1 = X.shape[1] + 1
    xs = [np.concatenate([p[k] for p in per], 0) if L else np.zeros((0, F1), dtype)
          for k in range(K + 1)]
    return dict(xs=xs, row_ptr=np.arange(L + 1, dtype=np.int64) * 2)


def hybrid_precompute(links, num_hops, A, X, K, dtype=np.float32):
    """utils.py:454-480 — PoS x, x1..xK, then SoP x2..xK appended as x{K+1}..x{2K-1}."""
    pos = pos_precompute(links, num_hops, A, X, K, None, dtype)
    if K == 1:
        return pos
    sop = sop_precompute(links, A, X, K, dtype)
    return dict(xs=pos['xs'] + sop['xs'][2:], row_ptr=pos['row_ptr'])


def node_labels(node_label, hops, lrowptr, lcol):
    """construct_pyg_graph's labeling-trick column (utils.py:296-310) in canonical local ids.
    'drnl' restates drnl_node_labeling (utils.py:211-236): BFS distance to local 0 with local 1 removed
    and vice versa, z = 1 + min + (d//2)*((d//2) + d%2 - 1), z[0] = z[1] = 1, unreachable -> 0."""
    n = hops.size
    if node_label == 'zo':
        return (hops == 0).astype(np.int64)
    if node_label == 'hop':
        return hops.astype(np.int64)
    if node_label == 'degree':
        return np.minimum(np.diff(lrowptr), 100).astype(np.int64)
    if node_label == 'drnl':
        def bfs(start, removed):
            dist = np.full(n, -1, dtype=np.int64)
            dist[start] = 0
            frontier = [start]
            while frontier:
                nxt = []
                for j in frontier:
                    for i in lcol[lrowptr[j]:lrowptr[j + 1]]:
                        if i != removed and dist[i] < 0:
                            dist[i] = dist[j] + 1
                            nxt.append(int(i))
                frontier = nxt
            return dist
        ds, dd = bfs(0, 1), bfs(1, 0)
        d = ds + dd
        z = 1 + np.minimum(ds, dd) + (d // 2) * ((d // 2) + d % 2 - 1)
        z[(ds < 0) | (dd < 0)] = 0
        z[:2] = 1
        return z.astype(np.int64)
    if node_label in ('de', 'de+'):
        raise NotImplementedError("two-column labels cannot go through the SIGN flow (utils.py:314)")
    return np.zeros(n, dtype=np.int64)


def sign_all_rows(src, dst, num_hops, A, X, K, dtype=np.float32, node_label='zo', with_nodes=False):
    """The NON-optimised flow (utils.py:497-520 + construct_pyg_graph utils.py:281-316 +
    TunedSIGN.__call__, tuned_SIGN.py:18-23 — PyG's SIGN transform on the whole subgraph):
    x = [z | X_sub], x_k = S x_{k-1} for ALL n rows, rows in canonical node order.  With
    node_label == 'zo' its rows 0 and 1 must equal the optimised PoS flow's output
    (SURVEY.md §8a row 9)."""
    nodes, hops, lrowptr, lcol = k_hop_subgraph(src, dst, num_hops, A)
    S, _ = normalized_subgraph(lrowptr, lcol, dtype)
    label = node_labels(node_label, hops, lrowptr, lcol).astype(dtype).reshape(-1, 1)
    xs = [np.hstack([label, np.asarray(X[nodes], dtype=dtype)])]
    for _ in range(K):
        xs.append(np.asarray(S @ xs[-1], dtype=dtype))
    return (xs, nodes) if with_nodes else xs


def full_precompute(links, num_hops, A, X, K, node_label='drnl', dtype=np.float32):
    """Whole call of the non-optimised PoS flow in the collated layout: K+1 row-stacked [sum n, F+1]
    arrays, row_ptr [L+1] and node_id [sum n]."""
    links = np.asarray(links)
    per = [sign_all_rows(int(links[0, i]), int(links[1, i]), num_hops, A, X, K, dtype, node_label, True)
           for i in range(links.shape[1])]
    row_ptr = np.zeros(links.shape[1] + 1, dtype=np.int64)
    row_ptr[1:] = np.cumsum([p[1].size for p in per])
    F1 = X.shape[1] + 1
    xs = [np.concatenate([p[0][k] for p in per], 0) if per else np.zeros((0, F1), dtype) for k in range(K + 1)]
    return dict(xs=xs, row_ptr=row_ptr, node_id=np.concatenate([p[1] for p in per]) if per else np.zeros(0, np.int64))


def joint_matrix(xs):
    """models.py:372 — feature-wise concat of the operators, [R, (K+1)(F+1)]."""
    return np.concatenate(xs, axis=-1)


# --------------------------------------------------------------------------------------
# roofline accounting
# --------------------------------------------------------------------------------------


def algorithmic_bytes(graph, A, F, K):
    """B(link) = 4·D + 8·n + 4·F·n + 4·s·(K+1)·(F+1), D = Σ_{v in V_sub} deg_G(v)."""
    nodes = graph['nodes']
    D = int(np.diff(A.indptr)[nodes].sum())
    n = nodes.size
    s = graph['sel'].size
    return 4 * D + 8 * n + 4 * F * n + 4 * s * (K + 1) * (F + 1)
