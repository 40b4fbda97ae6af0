"""host orchestration mirrors — get_pos_neg_edges (reference utils.py:637-678) against
outputs of the reference's own function (tests/golden/posneg_edges_ref.npz, generated through oracle/ref_runner
with np.random.seed(11)) and SEALDataset.process (sgrl_link_pred.py:96-220) end to end on the GPU."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from s3grl_b200 import get_pos_neg_edges
from s3grl_b200.dataset import sample_negative_edges

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'posneg_edges_ref.npz'))


def _inputs():
    edge = {s: {'edge': torch.as_tensor(G[f'in_edge_{s}']), 'edge_neg': torch.as_tensor(G[f'in_edgeneg_{s}'])}
            for s in ('train', 'valid', 'test')}
    src = {s: {k: torch.as_tensor(G[f'in_{k}_{s}']) for k in ('source_node', 'target_node', 'target_node_neg')}
           for s in ('train', 'valid', 'test')}
    return edge, src


@pytest.mark.parametrize('fmt,split,percent', [(f, s, p) for f in ('edge', 'src') for s in ('train', 'valid', 'test')
                                               for p in (100, 50) if not (f == 'src' and s == 'train')])
def test_get_pos_neg_edges_matches_reference(fmt, split, percent):
    edge, src = _inputs()
    np.random.seed(11)
    pos, neg = get_pos_neg_edges(split, edge if fmt == 'edge' else src, torch.zeros((2, 0), dtype=torch.long), 50, percent)
    assert np.array_equal(pos.numpy(), G[f'{fmt}_{split}_{percent}_pos'])      # link ORDER fixes the output row order
    assert np.array_equal(neg.numpy(), G[f'{fmt}_{split}_{percent}_neg'])


def test_do_edge_split_structure():
    """reference utils.py:588-634: 85/5/10 split, 1:1 negatives, training positives in both directions, no
    val/test positive left in the training graph, negatives are non-edges of the full graph."""
    from s3grl_b200 import datasets as ds, do_edge_split
    edges, N, _ = ds.load_graph('usair')
    und = np.asarray(edges)
    data = SimpleNamespace(edge_index=torch.as_tensor(np.concatenate([und.T, und.T[::-1]], 1).astype(np.int64)), num_nodes=N)
    se = do_edge_split(data, seed=1)
    E = np.unique(np.sort(und, 1), axis=0).shape[0]
    n_v, n_t = int(np.floor(0.05 * E)), int(np.floor(0.1 * E))
    assert se['valid']['edge'].shape == (n_v, 2) and se['test']['edge'].shape == (n_t, 2)
    assert se['train']['edge'].shape == (2 * (E - n_v - n_t), 2) == tuple(se['train']['edge_neg'].shape)
    key = lambda t: set((t[:, 0] * N + t[:, 1]).tolist())           # noqa: E731
    train = key(data.edge_index.t())
    assert train == key(se['train']['edge']) and all((b * N + a) in train for a, b in se['train']['edge'].tolist())
    full = key(torch.as_tensor(np.concatenate([und, und[:, ::-1]], 0).astype(np.int64)))
    for s in ('valid', 'test'):
        assert not (key(se[s]['edge']) & train) and key(se[s]['edge']) <= full
        assert not (key(se[s]['edge_neg']) & full)
    assert not (key(se['train']['edge_neg']) & train)
    with pytest.raises(NotImplementedError):
        do_edge_split(data, fast_split=True)


def test_sampled_negatives_are_non_edges():
    ei = torch.as_tensor(np.random.default_rng(0).integers(0, 30, (2, 200)))
    neg = sample_negative_edges(ei, 30, 150, seed=4).numpy()
    have = set((ei[0] * 30 + ei[1]).tolist())
    keys = neg[0] * 30 + neg[1]
    assert neg.shape == (2, 150) and len(set(keys.tolist())) == 150
    assert not (set(keys.tolist()) & have) and np.all(neg[0] != neg[1])
    assert np.array_equal(neg, sample_negative_edges(ei, 30, 150, seed=4).numpy())


@pytest.mark.gpu
@pytest.mark.parametrize('sign_type,k_heuristic,optimize', [('PoS', 0, True), ('PoS', 1, True), ('SoP', 0, True), ('PoS', 0, False)])
def test_seal_dataset_process_end_to_end(tmp_path, sign_type, k_heuristic, optimize):
    from golden_util import assert_features_close
    from oracle import s3grl_oracle as orc
    from s3grl_b200 import JointLoader, SEALDataset, datasets as ds
    edges, N, _ = ds.load_graph('usair')
    A, splits = ds.split_links(edges, N, seed=1)
    X = ds.synthetic_features(N, 12, 0.5, 2)
    coo = A.tocoo()
    data = SimpleNamespace(edge_index=torch.as_tensor(np.stack([coo.row, coo.col]).astype(np.int64)),
                           x=torch.as_tensor(X), num_nodes=N)
    pick = lambda a, n: torch.as_tensor(np.ascontiguousarray(a[:, :n].T))   # noqa: E731
    split_edge = {s: {'edge': pick(splits[s][0], 20), 'edge_neg': pick(splits[s][1], 20)} for s in ('train', 'valid', 'test')}
    args = SimpleNamespace(model='SIGN', sign_k=3, optimize_sign=optimize, k_heuristic=k_heuristic,
                           k_node_set_strategy='intersection' if k_heuristic else '', seed=1)
    np.random.seed(5)
    dset = SEALDataset(str(tmp_path), data, split_edge, 2, split='valid', node_label='zo', use_feature=True,
                       sign_type=sign_type, args=args)
    assert os.path.isfile(tmp_path / 'processed' / 'SEAL_valid_data.pt') and len(dset) == 40
    assert dset.num_features == 13 and dset.lists.y.tolist() == [1] * 20 + [0] * 20
    # the same links in the same order through the oracle
    np.random.seed(5)
    pos, neg = get_pos_neg_edges('valid', split_edge, data.edge_index, N, 100)
    links = torch.cat([pos, neg], 1).numpy()
    if not optimize:
        ref = orc.full_precompute(links, 2, A, X, 3, 'zo')
    elif sign_type == 'SoP':
        ref = orc.sop_precompute(links, A, X, 3)
    else:
        ref = orc.pos_precompute(links, 2, A, X, 3, 'intersection' if k_heuristic else None)
    assert np.array_equal(dset.slices['x'].cpu().numpy(), ref['row_ptr'])
    for k in range(4):
        assert_features_close(dset.lists.xs[k].cpu().numpy(), ref['xs'][k], what=f'{sign_type} x{k}')
    # a second construction re-uses the processed file, and the loader walks it
    again = SEALDataset(str(tmp_path), data, split_edge, 2, split='valid', node_label='zo', use_feature=True,
                        sign_type=sign_type, args=args)
    assert torch.equal(again.lists.xs[2].cpu(), dset.lists.xs[2].cpu())
    seen = sum(b.num_graphs for b in JointLoader(dset.lists, 16, shuffle=True, seed=0))
    assert seen == 40


@pytest.mark.gpu
def test_gpu_negative_sampling_and_edge_split():
    """SURVEY.md §8f row 4 (input side): s3_negative_candidates / do_edge_split_gpu against the properties the
    reference's split guarantees (utils.py:588-634, :645-648)."""
    from s3grl_b200 import DeviceGraph, datasets as ds, do_edge_split_gpu, precompute, sample_negative_edges_gpu
    edges, N, X = ds.load_graph('cora')
    A = ds.adjacency(edges, N)
    neg = sample_negative_edges_gpu(A.indptr, A.indices, N, 20000, seed=5)
    again = sample_negative_edges_gpu(A.indptr, A.indices, N, 20000, seed=5)
    more = sample_negative_edges_gpu(A.indptr, A.indices, N, 30000, seed=5)
    other = sample_negative_edges_gpu(A.indptr, A.indices, N, 20000, seed=6)
    assert torch.equal(neg, again) and torch.equal(more[:, :20000], neg) and not torch.equal(neg, other)
    n = neg.cpu().numpy()
    assert n.shape == (2, 20000) and (n[0] != n[1]).all() and n.min() >= 0 and n.max() < N
    assert np.unique(n[0] * N + n[1]).size == 20000                                # distinct ordered pairs
    assert not np.asarray(A[n[0], n[1]]).any()                                      # none is an edge
    hist = np.bincount(n.reshape(-1), minlength=N)                                  # endpoints roughly uniform
    assert hist.min() > 0 and hist.max() < 5 * hist.mean()
    # the whole split on the device, then straight into the hot path
    e2 = torch.as_tensor(np.concatenate([edges.T, edges.T[::-1]], 1))
    indptr, indices, se = do_edge_split_gpu(e2, N, seed=1)
    E = edges.shape[0]
    n_v, n_t = int(0.05 * E), int(0.1 * E)
    assert se['valid']['edge'].shape == (n_v, 2) and se['test']['edge'].shape == (n_t, 2)
    assert se['train']['edge'].shape == (2 * (E - n_v - n_t), 2) == tuple(se['train']['edge_neg'].shape)
    key = lambda t: set((t[:, 0] * N + t[:, 1]).cpu().tolist())           # noqa: E731
    full = set((edges[:, 0] * N + edges[:, 1]).tolist()) | set((edges[:, 1] * N + edges[:, 0]).tolist())
    rows = torch.repeat_interleave(torch.arange(N, device=indptr.device), indptr[1:] - indptr[:-1])
    train = set((rows * N + indices.long()).cpu().tolist())
    assert train == key(se['train']['edge']) and int(indptr[-1]) == 2 * (E - n_v - n_t)
    for s_ in ('valid', 'test'):
        assert not (key(se[s_]['edge']) & train) and key(se[s_]['edge']) <= full
        assert not (key(se[s_]['edge_neg']) & full)
    assert not (key(se['train']['edge_neg']) & train)
    g = DeviceGraph.from_device_csr(indptr, indices, torch.as_tensor(ds.normalize_features(X)[:, :64].copy()).to(indptr.device))
    links = torch.cat([se['valid']['edge'], se['valid']['edge_neg']]).t().contiguous()
    res = precompute(g, links, 2, 3)
    assert res.xs[0].shape == (2 * links.shape[1], 65) and bool(torch.isfinite(res.xs[3]).all())
