#!/usr/bin/env python
"""Where does a host-to-host call (bench.py e2e) spend its time?  Per iteration: graph upload, pinned output
buffers, precompute with the pipelined device->host copy, finish, release.   python tools/e2e_diag.py
Measured on the B200 pool: graph 2.2 ms, buffers 0 (cached), precompute 63-108 ms (41 ms of kernels overlapped with
2.63 GB over PCIe; the spread is the copy, i.e. host / PCIe load of the shared box), finish 0.2 ms."""
import sys, time, os, numpy as np, torch
sys.path.insert(0, '.')
import bench
from s3grl_b200 import tuned_sign, DeviceGraph, precompute
from s3grl_b200.tuned_sign import _host_buffers, _finish
w = bench.build_workload('pubmed_pos')
x_host = torch.from_numpy(w['X']).pin_memory()
links = torch.from_numpy(np.ascontiguousarray(w['links'])).pin_memory()
rows = []
for it in range(16):
    t0 = time.perf_counter()
    g = DeviceGraph(w['A'], x_host)
    t1 = time.perf_counter()
    host = _host_buffers(links.shape[1], g.num_feat, 3, 'cpu')
    t2 = time.perf_counter()
    res = precompute(g, links, 3, 3, flow='PoS', host_out=host)
    t3 = time.perf_counter()
    out = _finish(res, 1, host)
    t4 = time.perf_counter()
    del g, host, res, out
    t5 = time.perf_counter()
    rows.append([1000 * (b - a) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5), (t0, t5))])
for r in rows:
    print(' '.join(f'{v:7.1f}' for v in r))
print('cols: graph  hostbuf  precompute  finish  del  total')
