"""Small run of every flow / tier for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from golden_util import Case
from s3grl_b200 import DeviceGraph, precompute

for name, kw in [('tiny_pos_h3', {}), ('tiny_posplus_h2', {}), ('tiny_sop', {}), ('tiny_pos_h1', dict(force_sorted_tier=True)),
                 ('cora_posplus', {}), ('usair_posplus', {}), ('yeast_pos_k5', {})]:
    c = Case(name)
    g = DeviceGraph(c.A, c.X)
    links = c.links[:, :24]
    flow = 'SoP' if c.flow == 'sop' else 'PoS'
    for strat in ([c.strategy] if c.strategy else [None]) + (['union'] if c.strategy else []):
        res = precompute(g, links, c.num_hops, c.K, flow, strat, batch_records=16, **kw)
        torch.cuda.synchronize()
        print(name, strat, kw, 'rows', res.stats['rows'], 'max_n', res.stats['max_n'], flush=True)
rng = np.random.default_rng(0)
import scipy.sparse as ssp
N = 3000
u = rng.integers(0, N, 40000); v = rng.integers(0, N, 40000); k = u != v
A = ssp.csr_matrix((np.ones(k.sum() * 2), (np.r_[u[k], v[k]], np.r_[v[k], u[k]])), shape=(N, N))
X = rng.random((N, 20), dtype=np.float32)
links = rng.integers(0, N, (2, 32)); links = links[:, links[0] != links[1]]
g = DeviceGraph(A, X)
for kw in (dict(), dict(force_sorted_tier=True), dict(return_graphs=True, force_sorted_tier=True)):
    r = precompute(g, links, 1, 3, **kw); torch.cuda.synchronize(); print('random h=1', kw, r.stats['max_n'], flush=True)
print('done')
