#!/usr/bin/env python
"""Time the UNMODIFIED reference (imported from /root/reference behind oracle/ref_stub, see oracle/ref_runner.py) next to
the oracle port on the same links of the bench workload, single process, in the BUILD container (the reference cannot
travel to the GPU box, so `bench.py`'s CPU legs run the port; this file records how the two relate).

    python tools/time_reference.py [links] > profiles/r2_reference_cpu_timing.json

PubMed training graph, F = 500 synthetic features (the bench's spec), h = 3, K = 3; PoS, PoS Plus intersection and
PoS Plus union (the reference's ragged label-column literal repaired at run time, ref_runner.union_typo_repaired).
Test / measurement infrastructure only."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from oracle import ref_runner as rr                 # noqa: E402
from oracle import s3grl_oracle as orc              # noqa: E402
from s3grl_b200 import datasets as ds               # noqa: E402


def main():
    n_links = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    edges, N, _ = ds.load_graph('pubmed')
    A, splits = ds.split_links(edges, N, seed=1)
    X = ds.synthetic_features(N, 500, 0.1, 0)
    links = ds.all_links(splits)
    total = links.shape[1]
    pick = np.sort(np.random.default_rng(123).choice(total, n_links, replace=False))
    links = np.ascontiguousarray(links[:, pick])
    out = dict(workload="PubMed training graph, synthetic X F=500 (bench spec), h=3, K=3", links=n_links,
               sample=f"{n_links} links sampled uniformly (seed 123) of {total}",
               host=dict(cpu_count=os.cpu_count()), process="single process, single thread pool as torch / SciPy default",
               flows={})
    for name, strategy in (('pos', None), ('posplus_intersection', 'intersection'), ('posplus_union', 'union')):
        t0 = time.perf_counter()
        ref = rr.ref_pos(links, 3, A, X, 3, strategy, repair_union_typo=strategy == 'union')
        t_ref = time.perf_counter() - t0
        t0 = time.perf_counter()
        port = orc.pos_precompute(links, 3, A, X, 3, strategy, compat_explicit_zero=strategy == 'union')
        t_port = time.perf_counter() - t0
        err = 0.0
        if strategy != 'union':          # union rows are compared in tests/ (the reference's order differs)
            err = max(float(np.abs(a - b).max()) for a, b in zip(ref['xs'], port['xs']))
        out['flows'][name] = dict(reference_links_per_s=n_links / t_ref, port_links_per_s=n_links / t_port,
                                  port_over_reference=t_ref / t_port, rows=int(ref['row_ptr'][-1]), max_abs_diff=err)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
