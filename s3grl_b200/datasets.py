"""Workload builders for tests and bench.py: fixture graphs, the 85/5/10 link split, synthetic
features and the R-MAT generator.  Host-side NumPy only; nothing here is on
the hot path.  No file under /root/reference is read at run time — graphs come from
tests/golden/graphs/*.npz (made by oracle/make_fixtures.py).
"""
import os

import numpy as np
import scipy.sparse as ssp

_GRAPHS = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'graphs')


def load_graph(name):
    """-> (edges int64 [E,2] with u<v unique, num_nodes, X csr float32 or None)."""
    d = np.load(os.path.join(_GRAPHS, f'{name}.npz'))
    X = None
    if 'x_data' in d.files:
        X = ssp.csr_matrix((d['x_data'], d['x_indices'], d['x_indptr']), shape=tuple(d['x_shape']))
    return d['edges'].astype(np.int64), int(d['num_nodes']), X


def normalize_features(X):
    """PyG NormalizeFeatures: rows divided by their sum, clamped at 1 from below
    (applied by the reference at sgrl_link_pred.py:851 and :1000-1003)."""
    X = np.asarray(X.todense() if ssp.issparse(X) else X, dtype=np.float32)
    s = X.sum(axis=1, keepdims=True)
    s[s < 1] = 1  # clamp(min=1)
    return (X / s).astype(np.float32)


def synthetic_features(num_nodes, F, density=0.1, seed=0):
    """Row-normalised non-negative features with ~`density` non-zeros."""
    rng = np.random.default_rng(seed)
    X = rng.random((num_nodes, F), dtype=np.float32)
    X *= (rng.random((num_nodes, F), dtype=np.float32) < density)
    s = X.sum(axis=1, keepdims=True)
    s[s == 0] = 1
    return (X / s).astype(np.float32)


def degree_one_hot(A, max_degree=1024):
    """PyG OneHotDegree(max_degree) on the training graph: F = max_degree + 1
    (reference: sgrl_link_pred.py:961-963)."""
    deg = np.minimum(np.diff(A.indptr), max_degree)
    X = np.zeros((A.shape[0], max_degree + 1), dtype=np.float32)
    X[np.arange(A.shape[0]), deg] = 1
    return X


def adjacency(edges_undirected, num_nodes):
    """SciPy CSR with both directions of each edge and int64 multiplicity values — what
    SEALDataset.process builds (reference: sgrl_link_pred.py:107-114)."""
    e = np.asarray(edges_undirected, dtype=np.int64)
    row = np.concatenate([e[:, 0], e[:, 1]])
    col = np.concatenate([e[:, 1], e[:, 0]])
    A = ssp.csr_matrix((np.ones(row.shape[0], dtype=np.int64), (row, col)), shape=(num_nodes, num_nodes))
    A.sum_duplicates()
    A.sort_indices()
    return A


def _sample_non_edges(rng, forbidden_keys, num_nodes, count):
    """Uniform ordered pairs (u != v) whose key u*N+v is not in `forbidden_keys` (sorted)."""
    out = np.empty(0, dtype=np.int64)
    while out.size < count:
        need = int((count - out.size) * 1.2) + 16
        u = rng.integers(0, num_nodes, need)
        v = rng.integers(0, num_nodes, need)
        key = u * num_nodes + v
        ok = (u != v)
        pos = np.searchsorted(forbidden_keys, key)
        pos[pos >= forbidden_keys.size] = forbidden_keys.size - 1
        ok &= forbidden_keys[pos] != key
        out = np.unique(np.concatenate([out, key[ok]]))
    out = rng.permutation(out)[:count]
    return np.stack([out // num_nodes, out % num_nodes])


def split_links(edges_undirected, num_nodes, val_ratio=0.05, test_ratio=0.10, seed=1):
    """The reference's 85/5/10 split (utils.py:588-634 via PyG train_test_split_edges +
    negative_sampling), re-done with NumPy RNG (PyG's RNG stream is not reproducible here).

    Returns (A_train, splits) where splits[name] = (pos [2,L], neg [2,L]) int64.  Training
    positives contain BOTH directions of every training edge and as many
    training negatives; val/test positives are one direction with equally many negatives."""
    rng = np.random.default_rng(seed)
    e = np.asarray(edges_undirected, dtype=np.int64)
    E = e.shape[0]
    e = e[rng.permutation(E)]
    n_v, n_t = int(np.floor(val_ratio * E)), int(np.floor(test_ratio * E))
    val, test, train = e[:n_v], e[n_v:n_v + n_t], e[n_v + n_t:]
    A_train = adjacency(train, num_nodes)
    all_keys = np.unique(np.concatenate([e[:, 0] * num_nodes + e[:, 1], e[:, 1] * num_nodes + e[:, 0]]))
    vt_neg = _sample_non_edges(rng, all_keys, num_nodes, n_v + n_t)
    train_pos = np.concatenate([train.T, train.T[::-1]], axis=1)
    tr = A_train.tocoo()
    train_keys = np.unique(tr.row.astype(np.int64) * num_nodes + tr.col)
    train_neg = _sample_non_edges(rng, train_keys, num_nodes, train_pos.shape[1])
    # get_pos_neg_edges permutes each list (utils.py:651-659)
    splits = {
        'train': (train_pos[:, rng.permutation(train_pos.shape[1])], train_neg),
        'valid': (val.T[:, rng.permutation(n_v)], vt_neg[:, :n_v]),
        'test': (test.T[:, rng.permutation(n_t)], vt_neg[:, n_v:]),
    }
    return A_train, splits


def all_links(splits):
    """Every link the three SEALDataset.process calls precompute, in call order
    (train pos, train neg, valid pos, valid neg, test pos, test neg)."""
    parts = []
    for name in ('train', 'valid', 'test'):
        parts.extend(splits[name])
    return np.ascontiguousarray(np.concatenate(parts, axis=1))


def rmat_edges(scale, num_edges, num_nodes=None, abcd=(0.57, 0.19, 0.19, 0.05), seed=42, chunk=1 << 24):
    """R-MAT edge samples: ids scrambled by a fixed odd multiplier
    permutation of [0, 2^scale), reduced mod num_nodes, self loops dropped, symmetrised and
    de-duplicated by `adjacency`.  Returns unique undirected edges int64 [E,2] (u<v)."""
    rng = np.random.default_rng(seed)
    a, b, c, _ = abcd
    n = 1 << scale
    num_nodes = n if num_nodes is None else num_nodes
    keys = []
    done = 0
    while done < num_edges:
        m = min(chunk, num_edges - done)
        u = np.zeros(m, dtype=np.int64)
        v = np.zeros(m, dtype=np.int64)
        for _ in range(scale):
            r = rng.random(m, dtype=np.float32)
            right = (r >= a) & (r < a + b) | (r >= a + b + c)       # quadrants b, d
            down = r >= a + b                                        # quadrants c, d
            u = (u << 1) | down
            v = (v << 1) | right
        mult = 0x9E3779B1 | 1
        u = ((u * mult) & (n - 1)) % num_nodes
        v = ((v * mult) & (n - 1)) % num_nodes
        keep = u != v
        lo, hi = np.minimum(u[keep], v[keep]), np.maximum(u[keep], v[keep])
        keys.append(np.unique(lo * num_nodes + hi))
        done += m
    key = np.unique(np.concatenate(keys))
    return np.stack([key // num_nodes, key % num_nodes], axis=1)


def rmat_csr_torch(scale, num_edges, num_nodes=None, abcd=(0.57, 0.19, 0.19, 0.05), seed=42, device='cuda',
                   chunk=1 << 26):
    """R-MAT graph built on the GPU (SURVEY.md §8d config 5 is 200 M edge samples on 10 M nodes:
    minutes in NumPy, seconds here).  Same construction as `rmat_edges`: ids scrambled by an odd
    multiplier on [0, 2^scale), reduced mod num_nodes, self loops dropped, symmetrised,
    de-duplicated.  Returns (indptr int64 [N+1], indices int32 [nnz]) on `device`, columns
    ascending.  (Different RNG stream from the NumPy generator: same distribution, not the same graph.)"""
    import torch
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    a, b, c, _ = abcd
    n = 1 << scale
    num_nodes = n if num_nodes is None else int(num_nodes)
    keys = []
    done = 0
    while done < num_edges:
        m = min(chunk, num_edges - done)
        u = torch.zeros(m, dtype=torch.int64, device=dev)
        v = torch.zeros(m, dtype=torch.int64, device=dev)
        for _ in range(scale):
            r = torch.rand(m, device=dev, generator=gen)
            right = ((r >= a) & (r < a + b)) | (r >= a + b + c)
            down = r >= a + b
            u = (u << 1) | down.to(torch.int64)
            v = (v << 1) | right.to(torch.int64)
        mult = 0x9E3779B1 | 1
        u = ((u * mult) & (n - 1)) % num_nodes
        v = ((v * mult) & (n - 1)) % num_nodes
        keep = u != v
        u, v = u[keep], v[keep]
        keys.append(torch.cat([u * num_nodes + v, v * num_nodes + u]))     # both directions
        done += m
        del r, right, down, keep
    key = torch.unique(torch.cat(keys))           # sorted: (row, col) ascending
    del keys
    row = key // num_nodes
    indices = (key % num_nodes).to(torch.int32)
    counts = torch.bincount(row, minlength=num_nodes)
    indptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=indptr[1:])
    return indptr, indices
