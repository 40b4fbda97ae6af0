"""Load the committed golden cases (tests/golden/ref_*.npz, made by oracle/make_goldens.py
from the reference's own functions) and rebuild their inputs."""
import glob
import os

import numpy as np
import scipy.sparse as ssp

from s3grl_b200 import datasets as ds

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def case_names(flow=None):
    names = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, 'ref_*.npz')))
    if flow is None:
        return names
    return [n for n in names if str(np.load(os.path.join(GOLDEN, f'ref_{n}.npz'))['flow']) == flow]


def features_from_spec(spec, A, num_nodes):
    kind = spec.split(':')[0]
    if kind == 'cora':
        _, _, X = ds.load_graph('cora')
        return ds.normalize_features(X)
    if kind == 'degree':
        return ds.normalize_features(ds.degree_one_hot(A, int(spec.split(':')[1])))
    if kind == 'synthetic':
        _, F, density, seed = spec.split(':')
        return ds.synthetic_features(num_nodes, int(F), float(density), int(seed))
    raise ValueError(spec)


class Case:
    def __init__(self, name):
        d = np.load(os.path.join(GOLDEN, f'ref_{name}.npz'))
        self.name = name
        self.N = int(d['num_nodes'])
        self.A = ssp.csr_matrix((d['adata'], d['indices'], d['indptr']), shape=(self.N, self.N))
        self.links = d['links']
        self.L = self.links.shape[1]
        self.num_hops = int(d['num_hops'])
        self.K = int(d['K'])
        self.flow = str(d['flow'])
        self.strategy = str(d['strategy']) or None
        self.X = d['X'] if 'X' in d.files else features_from_spec(str(d['x_spec']), self.A, self.N)
        self.row_ptr = d['row_ptr']
        self.num_ops = 2 * self.K if self.flow == 'hybrid' else self.K + 1      # hybrid: x, x1..xK, then SoP x2..xK
        self.xs = [d[f'x{k}'] for k in range(self.num_ops)]
        if self.flow == 'scaled':
            self.sets = {int(k): row[row >= 0].astype(np.int64) for k, row in zip(d['set_nodes'], d['set_table'])}
            self.rw_m, self.rw_M = int(d['rw_m']), int(d['rw_M'])
        if self.flow == 'full':
            self.node_label, self.node_id = str(d['node_label']), d['node_id']
        if self.flow == 'pos_caps':
            self.caps = dict(ratio_per_hop=float(d['ratio_per_hop']), max_nodes_per_hop=int(d['max_nodes_per_hop']),
                             cap_seed=int(d['cap_seed']))
        if self.flow in ('pos', 'pos_caps'):
            self.row_gid = d['row_gid']
            self.node_ptr, self.edge_ptr = d['node_ptr'], d['edge_ptr']
            self.nodes, self.hops, self.edges = d['nodes'], d['hops'], d['edges']
        self.reference_repair = str(d['reference_repair']) if 'reference_repair' in d.files else ''
        if self.strategy == 'union':      # fixtures hold the reference's LITERAL rows: see drop_duplicated_seed_rows
            self._drop_duplicated_seed_rows()

    def _drop_duplicated_seed_rows(self):
        self.literal_row_ptr, self.literal_xs, self.literal_row_gid = self.row_ptr, self.xs, self.row_gid
        self.row_ptr, self.row_gid, self.xs = drop_duplicated_seed_rows(self.links, self.row_ptr, self.row_gid, self.xs, self.name)


def drop_duplicated_seed_rows(links, row_ptr, row_gid, xs, what=''):
    """`union` outputs of the reference (its tuned_SIGN.py:243 typo repaired, see oracle/ref_runner.union_typo_repaired)
    -> the same in the framework's row selection.  k_hop_subgraph masks the target link by ASSIGNING 0 to [0, 1] and
    [1, 0] (utils.py:78-79), which leaves two explicitly stored zeros in the CSR, and `neighbors` (utils.py:40) returns
    stored column ids — so N(0) contains 1, N(1) contains 0, and the union selects src and dst a second time among the
    extra rows.  The framework (like the paper) selects [0, 1] + (N(0) ∪ N(1)) − {0, 1}.  This checks that the two extra
    rows are bit-identical copies of rows 0 / 1 in every operator and removes them.  -> (row_ptr, row_gid, xs)"""
    L = links.shape[1]
    keep = np.ones(row_gid.size, dtype=bool)
    for i in range(L):
        a, b = int(row_ptr[i]), int(row_ptr[i + 1])
        src, dst = int(links[0][i]), int(links[1][i])
        assert row_gid[a] == src and row_gid[a + 1] == dst
        for seed_row, gid in ((a, src), (a + 1, dst)):
            dup = a + 2 + np.flatnonzero(row_gid[a + 2:b] == gid)
            assert dup.size == 1, f'{what} link {i}: expected exactly one duplicate of seed {gid}'
            for x in xs:
                assert np.array_equal(x[dup[0]], x[seed_row]), f'{what} link {i}: duplicate row differs'
            keep[dup[0]] = False
    return row_ptr - 2 * np.arange(L + 1, dtype=row_ptr.dtype), row_gid[keep], [x[keep] for x in xs]


def assert_features_close(got, ref, tol=1e-5, what=''):
    """north_star tolerance: 1e-5 relative fp32, taken per tensor against max|ref|
    (SURVEY.md §7 'fp32 tolerance'), plus a 1e-7 absolute floor for all-zero tensors."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} != {ref.shape}"
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    err = float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()) if ref.size else 0.0
    assert err <= tol * scale + 1e-7, f"{what}: max abs err {err:.3e} > {tol}*{scale:.3e}"
