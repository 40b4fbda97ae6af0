#!/usr/bin/env python
"""Build the small graph fixtures under tests/golden/graphs/ from the reference's bundled
DATA files (TEST INFRASTRUCTURE ONLY; run once in the build container, outputs committed).

    python oracle/make_fixtures.py [--reference /root/reference]

Nothing on the GPU box can read /root/reference, so tests and bench.py use these instead.
Only data (graph topology / Cora's bag-of-words rows) is converted; no reference source
code is read or copied.

* cora.npz    — Planetoid Cora as PyG's loader builds it (self loops dropped, undirected,
                de-duplicated; rows of allx/tx re-ordered by test.index), X stored sparse.
                What the reference loads at sgrl_link_pred.py:849-859.
* pubmed.npz  — Planetoid PubMed topology only (ind.pubmed.allx is absent from the
                reference checkout, see its .MISSING_LARGE_BLOBS); X is synthesised by
                s3grl_b200.datasets.synthetic_features.
* usair/yeast/power/router/ns/celegans.npz — SEAL edge lists with the reference's
                string-sorted node ids (data_utils.py:76-93).
Each file holds `edges` int32 [E,2] (u<v, unique, undirected) and `num_nodes`.
"""
import argparse
import os
import pickle
import sys

import numpy as np
import scipy.sparse as ssp

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '..', 'tests', 'golden', 'graphs')


def _undirected_unique(row, col, n):
    row, col = np.asarray(row, np.int64), np.asarray(col, np.int64)
    keep = row != col
    row, col = row[keep], col[keep]
    lo, hi = np.minimum(row, col), np.maximum(row, col)
    key = np.unique(lo * n + hi)
    return np.stack([key // n, key % n], 1).astype(np.int32)


def _load_pickle(path):
    with open(path, 'rb') as f:
        return pickle.load(f, encoding='latin1')


def planetoid(ref, name, with_x):
    raw = os.path.join(ref, 'data', name, 'raw')
    graph = _load_pickle(os.path.join(raw, f'ind.{name}.graph'))
    test_index = np.loadtxt(os.path.join(raw, f'ind.{name}.test.index'), dtype=np.int64)
    rows, cols = [], []
    for k, vs in graph.items():
        rows.extend([k] * len(vs))
        cols.extend(vs)
    n = max(max(rows), max(cols)) + 1
    out = {'edges': _undirected_unique(rows, cols, n), 'num_nodes': np.int64(n)}
    if with_x:
        allx = _load_pickle(os.path.join(raw, f'ind.{name}.allx')).tocsr()
        tx = _load_pickle(os.path.join(raw, f'ind.{name}.tx')).tocsr()
        x = ssp.vstack([allx, tx]).tolil()
        sorted_test = np.sort(test_index)
        x[test_index, :] = x[sorted_test, :]          # PyG read_planetoid_data re-ordering
        x = x.tocsr().astype(np.float32)
        x.sort_indices()
        assert x.shape[0] == n
        out.update(x_indptr=x.indptr.astype(np.int32), x_indices=x.indices.astype(np.int32),
                   x_data=x.data.astype(np.float32), x_shape=np.asarray(x.shape, np.int64))
    return out


def seal_edge_list(ref, name):
    path = os.path.join(ref, 'data', 'link_prediction', name, 'edges.txt')
    pairs = []
    with open(path) as f:
        for line in f:
            a, b = line.strip().split()[:2]
            pairs.append((a, b))
    ids = sorted({t for p in pairs for t in p})      # string sort, as data_utils.read_label
    m = {s: i for i, s in enumerate(ids)}
    row = [m[a] for a, _ in pairs]
    col = [m[b] for _, b in pairs]
    n = len(ids)
    return {'edges': _undirected_unique(row, col, n), 'num_nodes': np.int64(n)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    jobs = {'cora': lambda: planetoid(args.reference, 'cora', True),
            'pubmed': lambda: planetoid(args.reference, 'pubmed', False)}
    for nm in ['usair', 'yeast', 'power', 'router', 'ns', 'celegans']:
        jobs[nm] = (lambda nm=nm: seal_edge_list(args.reference, nm))
    for nm, fn in jobs.items():
        d = fn()
        np.savez_compressed(os.path.join(OUT, f'{nm}.npz'), **d)
        print(f"{nm}: N={int(d['num_nodes'])} undirected edges={d['edges'].shape[0]}"
              + (f" X={tuple(d['x_shape'])} nnz={d['x_data'].size}" if 'x_data' in d else ''))


if __name__ == '__main__':
    sys.exit(main())
