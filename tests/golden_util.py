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
        if self.flow == 'pos':
            self.row_gid = d['row_gid']
            self.node_ptr, self.edge_ptr = d['node_ptr'], d['edge_ptr']
            self.nodes, self.hops, self.edges = d['nodes'], d['hops'], d['edges']


def assert_features_close(got, ref, tol=1e-5, what=''):
    """north_star tolerance: 1e-5 relative fp32, taken per tensor against max|ref|
    (SURVEY.md §7 'fp32 tolerance'), plus a 1e-7 absolute floor for all-zero tensors."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} != {ref.shape}"
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    err = float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()) if ref.size else 0.0
    assert err <= tol * scale + 1e-7, f"{what}: max abs err {err:.3e} > {tol}*{scale:.3e}"
