#!/usr/bin/env python
"""Generate tests/golden/ref_*.npz by running the REFERENCE's own functions (imported from
/root/reference behind oracle/ref_stub, see oracle/ref_runner.py) on small seeded inputs.
TEST INFRASTRUCTURE ONLY; run in the build container, outputs are committed:

    python oracle/make_goldens.py

Every file is self-contained: inputs (CSR of the training graph, features or a feature
spec, links, num_hops, K, flow, strategy) and the reference's outputs in canonical order
(operators x0..xK row-stacked, row_ptr, the global id of every output row, and for PoS
flows the node / hop / induced-edge lists of every link from the reference's
k_hop_subgraph).  The `union` strategy has no golden: the reference raises for it
(tuned_SIGN.py:243).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, '..')
sys.path.insert(0, ROOT)

from oracle import ref_runner as rr          # noqa: E402
from s3grl_b200 import datasets as ds        # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def features_from_spec(spec, A, num_nodes):
    """Feature matrices are rebuilt from a spec where storing them would bloat the fixture."""
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


def _graph_dump(links, h, A):
    node_ptr, edge_ptr, nodes, hops, edges = [0], [0], [], [], []
    for i in range(links.shape[1]):
        cn, hp, ed = rr.ref_k_hop(int(links[0, i]), int(links[1, i]), h, A)
        nodes.append(cn)
        hops.append(hp)
        edges.append(ed)
        node_ptr.append(node_ptr[-1] + cn.size)
        edge_ptr.append(edge_ptr[-1] + ed.shape[0])
    return dict(node_ptr=np.asarray(node_ptr, np.int64), edge_ptr=np.asarray(edge_ptr, np.int64),
                nodes=np.concatenate(nodes).astype(np.int32), hops=np.concatenate(hops).astype(np.int8),
                edges=np.concatenate(edges, 0).astype(np.int32))


def make_case(name, A, links, flow, K, num_hops=0, strategy=None, X=None, x_spec=None):
    A = A.tocsr()
    A.sort_indices()
    N = A.shape[0]
    feats = X if X is not None else features_from_spec(x_spec, A, N)
    links = np.ascontiguousarray(links, dtype=np.int64)
    if flow == 'pos':
        r = rr.ref_pos(links, num_hops, A, feats, K, strategy, repair_union_typo=strategy == 'union')
        extra = _graph_dump(links, num_hops, A)
        extra['row_gid'] = r['row_gid']
        if strategy == 'union':
            extra['reference_repair'] = np.str_(
                "tuned_SIGN.py:243 `[[0] * (csr_shape - 2)]` read as `[[0]] * (csr_shape - 2)` (the intersection branch's "
                "line 247) by oracle/ref_runner.union_typo_repaired at run time; /root/reference unmodified on disk")
    elif flow == 'sop':
        r = rr.ref_sop(links, A, feats, K)
        extra = {}
    elif flow == 'hybrid':
        r = rr.ref_hybrid(links, num_hops, A, feats, K)
        extra = {}
    else:
        raise ValueError(flow)
    out = dict(indptr=A.indptr.astype(np.int64), indices=A.indices.astype(np.int32),
               adata=A.data.astype(np.int64), num_nodes=np.int64(N), links=links,
               num_hops=np.int64(num_hops), K=np.int64(K), flow=np.str_(flow),
               strategy=np.str_(strategy or ''), row_ptr=r['row_ptr'], **extra)
    if X is not None:
        out['X'] = np.asarray(X, np.float32)
    else:
        out['x_spec'] = np.str_(x_spec)
    for k, x in enumerate(r['xs']):
        out[f'x{k}'] = x.astype(np.float32)
    path = os.path.join(OUT, f'ref_{name}.npz')
    np.savez_compressed(path, **out)
    print(f"{name}: L={links.shape[1]} R={int(r['row_ptr'][-1])} F'={r['xs'][0].shape[1]} "
          f"-> {os.path.getsize(path) / 1024:.0f} KiB")


def make_scaled_case(name, A, links, K, x_spec, rw_m=3, rw_M=20, seed=5):
    """ScaLed (configs/paper/scaled.json: m = 3, M = 20, num_hops = 0): the walk sets are drawn here
    with NumPy and stored in the fixture — the reference's own sampler (torch_cluster.random_walk) is
    not available — and the REFERENCE code turns them into subgraphs and operators."""
    from oracle import s3grl_oracle as orc
    A = A.tocsr()
    A.sort_indices()
    N = A.shape[0]
    feats = features_from_spec(x_spec, A, N)
    links = np.ascontiguousarray(links, dtype=np.int64)
    sets = orc.random_walk_sets(A, links.reshape(-1), rw_m, rw_M, seed)
    r = rr.ref_scaled_pos(links, A, feats, K, sets)
    keys = np.array(sorted(sets), dtype=np.int64)
    cap = max(v.size for v in sets.values())
    table = np.full((keys.size, cap), -1, dtype=np.int32)
    for i, k in enumerate(keys):
        table[i, :sets[int(k)].size] = sets[int(k)]
    out = dict(indptr=A.indptr.astype(np.int64), indices=A.indices.astype(np.int32), adata=A.data.astype(np.int64),
               num_nodes=np.int64(N), links=links, num_hops=np.int64(0), K=np.int64(K), flow=np.str_('scaled'),
               strategy=np.str_(''), row_ptr=r['row_ptr'], x_spec=np.str_(x_spec), set_nodes=keys, set_table=table,
               rw_m=np.int64(rw_m), rw_M=np.int64(rw_M))
    for k, x in enumerate(r['xs']):
        out[f'x{k}'] = x.astype(np.float32)
    path = os.path.join(OUT, f'ref_{name}.npz')
    np.savez_compressed(path, **out)
    print(f"{name}: L={links.shape[1]} sets={keys.size} cap={cap} -> {os.path.getsize(path) / 1024:.0f} KiB")


def make_full_case(name, A, links, K, num_hops, node_label, X=None, x_spec=None):
    """The non-optimised PoS flow (optimize_sign=False, utils.py:497-520) through the reference."""
    A = A.tocsr()
    A.sort_indices()
    N = A.shape[0]
    feats = X if X is not None else features_from_spec(x_spec, A, N)
    links = np.ascontiguousarray(links, dtype=np.int64)
    r = rr.ref_full(links, num_hops, A, feats, K, node_label)
    out = dict(indptr=A.indptr.astype(np.int64), indices=A.indices.astype(np.int32), adata=A.data.astype(np.int64),
               num_nodes=np.int64(N), links=links, num_hops=np.int64(num_hops), K=np.int64(K), flow=np.str_('full'),
               strategy=np.str_(''), node_label=np.str_(node_label), row_ptr=r['row_ptr'], node_id=r['node_id'])
    if X is not None:
        out['X'] = np.asarray(X, np.float32)
    else:
        out['x_spec'] = np.str_(x_spec)
    for k, x in enumerate(r['xs']):
        out[f'x{k}'] = x.astype(np.float32)
    path = os.path.join(OUT, f'ref_{name}.npz')
    np.savez_compressed(path, **out)
    print(f"{name}: L={links.shape[1]} R={int(r['row_ptr'][-1])} -> {os.path.getsize(path) / 1024:.0f} KiB")


def make_posneg_case():
    """get_pos_neg_edges (reference utils.py:637-678) on seeded split dictionaries, np.random.seed(11): the link
    ORDER it produces fixes the output row order.  The 'edge' format carries pre-sampled negatives (edge_neg), so
    PyG's negative_sampling is not needed; the 'source_node' train split draws negatives with torch.randint and is
    left out."""
    import torch
    utils, _ = rr.load_reference()
    rng = np.random.default_rng(3)
    N = 50
    mk = lambda n: torch.as_tensor(rng.integers(0, N, (n, 2)))      # noqa: E731
    split_edge = {s: {'edge': mk(n), 'edge_neg': mk(n2)} for s, n, n2 in (('train', 40, 40), ('valid', 9, 11), ('test', 13, 13))}
    src = {s: {'source_node': torch.as_tensor(rng.integers(0, N, n)), 'target_node': torch.as_tensor(rng.integers(0, N, n)),
               'target_node_neg': torch.as_tensor(rng.integers(0, N, (n, 3)))} for s, n in (('train', 30), ('valid', 8), ('test', 10))}
    out = {}
    for name, se in (('edge', split_edge), ('src', src)):
        for split in ('train', 'valid', 'test'):
            for percent in (100, 50):
                if name == 'src' and split == 'train':
                    continue
                np.random.seed(11)
                p, n = utils.get_pos_neg_edges(split, se, torch.zeros((2, 0), dtype=torch.long), N, percent)
                out[f'{name}_{split}_{percent}_pos'] = p.numpy()
                out[f'{name}_{split}_{percent}_neg'] = n.numpy()
    for s in split_edge:
        out[f'in_edge_{s}'] = split_edge[s]['edge'].numpy()
        out[f'in_edgeneg_{s}'] = split_edge[s]['edge_neg'].numpy()
    for s in src:
        for k, v in src[s].items():
            out[f'in_{k}_{s}'] = v.numpy()
    np.savez_compressed(os.path.join(OUT, 'posneg_edges_ref.npz'), **out)
    print(f"posneg_edges_ref: {len(out)} arrays")


def tiny_graphs():
    """Hand graphs for the edge cases of SURVEY.md A.6: isolated endpoints, n == 2, an
    endpoint whose only neighbour is the other endpoint, pendant paths, a hub, two components."""
    E = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5),          # path 0..5
         (6, 7), (6, 8), (7, 8), (8, 9),                  # triangle + pendant
         (10, 11), (10, 12), (10, 13), (10, 14), (10, 15), (11, 12),   # hub 10
         (16, 17),                                        # lone edge
         (5, 6)]                                          # bridge
    N = 21                                                # 18, 19, 20 isolated
    e = np.asarray([(min(a, b), max(a, b)) for a, b in E], dtype=np.int64)
    A = ds.adjacency(e, N)
    links = np.asarray([(0, 1), (1, 0), (0, 5), (2, 4), (6, 7), (7, 9), (10, 11), (11, 12), (13, 14),
                        (16, 17), (18, 19), (18, 0), (3, 20), (10, 8), (5, 6), (12, 15), (17, 4)]).T
    rng = np.random.default_rng(7)
    X = rng.random((N, 5), dtype=np.float32)
    return A, links, X


def sample_links(splits, count, seed):
    links = ds.all_links(splits)
    pick = np.sort(np.random.default_rng(seed).choice(links.shape[1], count, replace=False))
    return links[:, pick]


def round2_cases():
    """Fixtures added in round 2 (VERDICT r1 "parity holes"): the hybrid flow, the headline PubMed graph with the
    bench's F = 500 feature spec (PoS and PoS Plus intersection), and Router (BASELINE config 4)."""
    A, links, X = tiny_graphs()
    make_case('tiny_hybrid', A, links, 'hybrid', 3, 2, None, X=X)
    edges, N, _ = ds.load_graph('usair')
    A, splits = ds.split_links(edges, N, seed=1)
    make_case('usair_hybrid', A, sample_links(splits, 40, 5), 'hybrid', 3, 2, None, x_spec='synthetic:16:0.5:6')
    edges, N, _ = ds.load_graph('pubmed')
    A, splits = ds.split_links(edges, N, seed=1)
    make_case('pubmed_pos', A, sample_links(splits, 24, 11), 'pos', 3, 3, None, x_spec='synthetic:500:0.1:0')
    make_case('pubmed_posplus', A, sample_links(splits, 24, 12), 'pos', 3, 3, 'intersection', x_spec='synthetic:500:0.1:0')
    edges, N, _ = ds.load_graph('router')
    A, splits = ds.split_links(edges, N, seed=1)
    make_case('router_pos_k5', A, sample_links(splits, 60, 13), 'pos', 5, 2, None, x_spec='synthetic:32:0.5:8')


def caps_case():
    """Per-hop caps (utils.py:66-70): the reference's capped BFS + PoS with ONLY its `random.sample` replaced by the
    framework's rank rule (oracle/ref_runner.cap_sampler_ranked).  flow = 'pos_caps', so the uncapped PoS tests skip it."""
    edges, N, _ = ds.load_graph('usair')
    A, splits = ds.split_links(edges, N, seed=1)
    A = A.tocsr()
    A.sort_indices()
    links = np.ascontiguousarray(sample_links(splits, 40, 31), dtype=np.int64)
    caps = dict(ratio_per_hop=0.6, max_nodes_per_hop=12, cap_seed=5)
    h, K, x_spec = 2, 3, 'synthetic:16:0.5:6'
    feats = features_from_spec(x_spec, A, N)
    r = rr.ref_pos(links, h, A, feats, K, None, caps=caps)
    node_ptr, edge_ptr, nodes, hops, edge_list = [0], [0], [], [], []
    for i in range(links.shape[1]):
        cn, hp, ed = rr.ref_k_hop(int(links[0, i]), int(links[1, i]), h, A, caps['ratio_per_hop'], caps['max_nodes_per_hop'],
                                  caps['cap_seed'])
        nodes.append(cn)
        hops.append(hp)
        edge_list.append(ed)
        node_ptr.append(node_ptr[-1] + cn.size)
        edge_ptr.append(edge_ptr[-1] + ed.shape[0])
    out = dict(indptr=A.indptr.astype(np.int64), indices=A.indices.astype(np.int32), adata=A.data.astype(np.int64),
               num_nodes=np.int64(N), links=links, num_hops=np.int64(h), K=np.int64(K), flow=np.str_('pos_caps'),
               strategy=np.str_(''), row_ptr=r['row_ptr'], row_gid=r['row_gid'], x_spec=np.str_(x_spec),
               node_ptr=np.asarray(node_ptr, np.int64), edge_ptr=np.asarray(edge_ptr, np.int64),
               nodes=np.concatenate(nodes).astype(np.int32), hops=np.concatenate(hops).astype(np.int8),
               edges=np.concatenate(edge_list, 0).astype(np.int32), ratio_per_hop=np.float64(caps['ratio_per_hop']),
               max_nodes_per_hop=np.int64(caps['max_nodes_per_hop']), cap_seed=np.int64(caps['cap_seed']),
               reference_repair=np.str_("utils.py:67,70 `random.sample` replaced by the rank rule (smallest fmix32(node ^ seed)) by "
                                        "oracle/ref_runner.cap_sampler_ranked at run time; /root/reference unmodified on disk"))
    for k, x in enumerate(r['xs']):
        out[f'x{k}'] = x.astype(np.float32)
    path = os.path.join(OUT, 'ref_usair_pos_caps.npz')
    np.savez_compressed(path, **out)
    print(f"usair_pos_caps: L={links.shape[1]} nodes={node_ptr[-1]} -> {os.path.getsize(path) / 1024:.0f} KiB")


def union_cases():
    """PoS Plus `union` (BASELINE config 3) against the reference with its label-column literal repaired: hand graphs,
    USAir, Cora and the PubMed graph with the bench's F = 500 feature spec."""
    A, links, X = tiny_graphs()
    make_case('tiny_posplus_union_h2', A, links, 'pos', 3, 2, 'union', X=X)
    make_case('tiny_posplus_union_h1_k5', A, links, 'pos', 5, 1, 'union', X=X)
    edges, N, _ = ds.load_graph('usair')
    A, splits = ds.split_links(edges, N, seed=1)
    make_case('usair_posplus_union', A, sample_links(splits, 40, 21), 'pos', 3, 2, 'union', x_spec='synthetic:16:0.5:6')
    edges, N, _ = ds.load_graph('cora')
    A, splits = ds.split_links(edges, N, seed=1)
    make_case('cora_posplus_union', A, sample_links(splits, 60, 22), 'pos', 3, 3, 'union', x_spec='synthetic:24:0.3:5')
    edges, N, _ = ds.load_graph('pubmed')
    A, splits = ds.split_links(edges, N, seed=1)
    make_case('pubmed_posplus_union', A, sample_links(splits, 12, 23), 'pos', 3, 3, 'union', x_spec='synthetic:500:0.1:0')


def main():
    os.makedirs(OUT, exist_ok=True)
    if '--round2' in sys.argv:          # only the fixtures added in round 2 (the others are unchanged)
        round2_cases()
        return
    if '--union' in sys.argv:           # only the union fixtures (round 2, third session)
        union_cases()
        return
    if '--caps' in sys.argv:
        caps_case()
        return
    make_posneg_case()
    A, links, X = tiny_graphs()
    for h in (1, 2, 3):
        make_case(f'tiny_pos_h{h}', A, links, 'pos', 3, h, None, X=X)
    make_case('tiny_posplus_h2', A, links, 'pos', 3, 2, 'intersection', X=X)
    make_case('tiny_sop', A, links, 'sop', 3, X=X)
    for lab in ('zo', 'hop', 'drnl', 'degree', 'none'):
        make_full_case(f'tiny_full_{lab}', A, links, 3, 2, lab, X=X)

    edges, N, _ = ds.load_graph('cora')
    A, splits = ds.split_links(edges, N, seed=1)
    make_case('cora_pos_fullF', A, sample_links(splits, 20, 0), 'pos', 3, 3, None, x_spec='cora')
    make_case('cora_pos', A, sample_links(splits, 160, 1), 'pos', 3, 3, None, x_spec='synthetic:24:0.3:5')
    make_case('cora_posplus', A, sample_links(splits, 160, 2), 'pos', 3, 3, 'intersection',
              x_spec='synthetic:24:0.3:5')

    make_full_case('cora_full_drnl', A, sample_links(splits, 16, 9), 3, 3, 'drnl', x_spec='synthetic:24:0.3:5')
    make_full_case('cora_full_zo_h2', A, sample_links(splits, 24, 10), 2, 2, 'zo', x_spec='synthetic:24:0.3:5')

    make_scaled_case('cora_scaled', A, sample_links(splits, 120, 6), 3, 'synthetic:24:0.3:5')

    edges, N, _ = ds.load_graph('usair')
    A, splits = ds.split_links(edges, N, seed=1)
    make_case('usair_sop_degree', A, sample_links(splits, 12, 0), 'sop', 3, x_spec='degree:1024')
    make_case('usair_sop', A, sample_links(splits, 160, 1), 'sop', 3, x_spec='synthetic:16:0.5:6')
    make_case('usair_posplus', A, sample_links(splits, 120, 2), 'pos', 3, 2, 'intersection',
              x_spec='synthetic:16:0.5:6')

    edges, N, _ = ds.load_graph('yeast')
    A, splits = ds.split_links(edges, N, seed=1)
    make_case('yeast_pos_k5', A, sample_links(splits, 100, 3), 'pos', 5, 2, None, x_spec='synthetic:32:0.5:8')

    edges, N, _ = ds.load_graph('power')
    A, splits = ds.split_links(edges, N, seed=1)
    make_case('power_pos_k5', A, sample_links(splits, 100, 4), 'pos', 5, 2, None, x_spec='synthetic:8:1.0:9')
    round2_cases()
    union_cases()
    caps_case()


if __name__ == '__main__':
    main()
