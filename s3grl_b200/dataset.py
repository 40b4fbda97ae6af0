"""Host-side mirror of the reference's dataset layer for the SIGN models:
`get_pos_neg_edges` (reference utils.py:637-678) and `SEALDataset` (sgrl_link_pred.py:54-220) — the
caller of the hot path.  Same constructor arguments, same `process()` sequence (build A from
`data.edge_index`, pick the split's positive / negative links with the same `np.random.permutation`
calls, run `extract_enclosing_subgraphs` on positives (y = 1) then negatives (y = 0), collate and
`torch.save` to `processed_paths[0]`), same file name, and the `(data, slices)` file layout.  The
subgraph extraction and operator construction run on the GPU (no CPU path); PyG itself is not needed.
"""
import os

import numpy as np
import scipy.sparse as ssp
import torch

from .data import PrecomputedList
from .loader import load_collated, save_collated
from .utils import extract_enclosing_subgraphs


def sample_negative_edges(edge_index, num_nodes, num_neg_samples, seed=0):
    """Stand-in for torch_geometric.utils.negative_sampling(add_self_loops(edge_index), ...) at reference
    utils.py:645-648: `num_neg_samples` distinct ordered pairs (u, v), u != v, that are not stored edges,
    drawn uniformly by rejection.  Same distribution, different RNG stream than PyG's sampler."""
    ei = np.asarray(edge_index)
    N = int(num_nodes)
    taken = set((ei[0].astype(np.int64) * N + ei[1].astype(np.int64)).tolist())
    rng = np.random.default_rng(seed)
    out = []
    need = int(num_neg_samples)
    if need > N * (N - 1) - len(taken):
        raise ValueError("not enough non-edges")
    while need > 0:
        u = rng.integers(0, N, 2 * need + 16)
        v = rng.integers(0, N, 2 * need + 16)
        for a, b in zip(u.tolist(), v.tolist()):
            k = a * N + b
            if a != b and k not in taken:
                taken.add(k)
                out.append((a, b))
                need -= 1
                if need == 0:
                    break
    return torch.as_tensor(np.asarray(out, dtype=np.int64).reshape(-1, 2).T.copy())


def sample_negative_edges_gpu(indptr, indices, num_nodes, num_neg_samples, seed=0, device='cuda'):
    """GPU negative sampling (csrc/pair.cu, s3_negative_candidates; reference utils.py:645-648): `num_neg_samples`
    distinct ordered pairs (u, v), u != v, that are not stored entries of the CSR (int64 indptr, int32 indices with
    ascending columns, on `device`).  Candidates come from a counter-based hash of (seed, index); the first
    `num_neg_samples` valid ones in index order are kept, so the result depends on (graph, seed) only.
    Returns an int64 [2, num_neg_samples] tensor on the device.  No CPU path."""
    import ctypes as C
    from . import _lib as L
    lib = L.lib()
    dev = torch.device(device)
    indptr = torch.as_tensor(indptr).to(device=dev, dtype=torch.int64).contiguous()
    indices = torch.as_tensor(indices).to(device=dev, dtype=torch.int32).contiguous()
    N, need = int(num_nodes), int(num_neg_samples)
    if need > N * (N - 1) - int(indices.numel()):
        raise ValueError("not enough non-edges")
    g = L.Graph(C.c_void_p(indptr.data_ptr()), C.c_void_p(indices.data_ptr()), None, N, 0, 0, int(indices.numel()), 0)
    density = float(indices.numel()) / max(1.0, float(N) * N)
    M = max(1024, int(need * (1.1 + 2 * density)) + 64)
    st = torch.cuda.current_stream(dev)
    while True:
        slots = int(lib.s3_pair_table_slots(M))
        table = torch.empty(2 * slots, dtype=torch.int64, device=dev)
        src = torch.empty(M, dtype=torch.int64, device=dev)
        dst = torch.empty(M, dtype=torch.int64, device=dev)
        valid = torch.empty(M, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.s3_negative_candidates(C.byref(g), M, int(seed) & (2**64 - 1), C.c_void_p(table.data_ptr()), slots,
                                               C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), C.c_void_p(valid.data_ptr()),
                                               C.c_void_p(st.cuda_stream)), 's3_negative_candidates')
        keep = torch.nonzero(valid, as_tuple=False).flatten()          # ascending candidate index
        if int(keep.numel()) >= need:
            keep = keep[:need]
            return torch.stack([src[keep], dst[keep]])
        M *= 2          # a prefix of a longer candidate list: the kept pairs so far stay the same


def do_edge_split_gpu(edge_index, num_nodes, val_ratio=0.05, test_ratio=0.1, seed=1, device='cuda'):
    """The link split of reference utils.py:588-634 on the GPU: undirected edges are permuted
    (torch.randperm on the device, seeded), cut 5 / 10 / 85 %, the training graph's CSR is built on the device and all
    negatives come from s3_negative_candidates — validation / test negatives avoid every edge of the full graph,
    training negatives the training edges, as `datasets.split_links` does on the host (same construction; the random
    stream is this implementation's, as PyG's is not reproducible here).
    Returns (indptr int64 [N+1], indices int32 [nnz] of the training graph, split_edge) with split_edge in the
    reference's layout {'train'|'valid'|'test': {'edge': [L, 2], 'edge_neg': [L, 2]}}, everything on the device;
    training positives hold both directions of every training edge."""
    dev = torch.device(device)
    ei = torch.as_tensor(edge_index).to(dev)
    N = int(num_nodes)
    lo, hi = torch.minimum(ei[0], ei[1]), torch.maximum(ei[0], ei[1])
    key = torch.unique(lo[lo != hi] * N + hi[lo != hi])
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    key = key[torch.randperm(key.numel(), device=dev, generator=gen)]
    E = int(key.numel())
    n_v, n_t = int(val_ratio * E), int(test_ratio * E)
    und = torch.stack([key // N, key % N])
    val, test, train = und[:, :n_v], und[:, n_v:n_v + n_t], und[:, n_v + n_t:]

    def csr(u2):
        k = torch.unique(torch.cat([u2[0] * N + u2[1], u2[1] * N + u2[0]]))
        row = k // N
        indptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        torch.cumsum(torch.bincount(row, minlength=N), 0, out=indptr[1:])
        return indptr, (k % N).to(torch.int32)
    full_ptr, full_idx = csr(und)
    tr_ptr, tr_idx = csr(train)
    vt_neg = sample_negative_edges_gpu(full_ptr, full_idx, N, n_v + n_t, seed=int(seed) * 3 + 1, device=dev)
    train_pos = torch.cat([train, train.flip(0)], 1)
    train_neg = sample_negative_edges_gpu(tr_ptr, tr_idx, N, train_pos.shape[1], seed=int(seed) * 3 + 2, device=dev)
    split_edge = {'train': {'edge': train_pos.t().contiguous(), 'edge_neg': train_neg.t().contiguous()},
                  'valid': {'edge': val.t().contiguous(), 'edge_neg': vt_neg[:, :n_v].t().contiguous()},
                  'test': {'edge': test.t().contiguous(), 'edge_neg': vt_neg[:, n_v:].t().contiguous()}}
    return tr_ptr, tr_idx, split_edge


def do_edge_split(data, fast_split=False, val_ratio=0.05, test_ratio=0.1, neg_ratio=1, seed=1):
    """reference utils.py:588-634 (`data_passed=True` form): the 85/5/10 link split + 1:1 negatives, returned as the
    same `split_edge` dictionary ({'train'|'valid'|'test': {'edge': [L,2], 'edge_neg': [L,2]}}) and leaving
    `data.edge_index` = both directions of the TRAINING edges (what SEALDataset builds A from).  PyG's
    train_test_split_edges / negative_sampling draw from torch's RNG and are not importable here, so the split is
    drawn with NumPy (`datasets.split_links`, seeded): same construction, different random stream.  Training
    positives hold both directions of every training edge."""
    from . import datasets as ds
    if fast_split:
        raise NotImplementedError('Fast split is untested and unsupported.')      # reference utils.py:601
    if neg_ratio != 1:
        raise NotImplementedError("neg_ratio != 1 is not supported")
    ei = data.edge_index.cpu().numpy()
    und = np.unique(np.stack([np.minimum(ei[0], ei[1]), np.maximum(ei[0], ei[1])], 1), axis=0)
    und = und[und[:, 0] != und[:, 1]]
    A_train, splits = ds.split_links(und, int(data.num_nodes), val_ratio, test_ratio, seed)
    coo = A_train.tocoo()
    data.edge_index = torch.as_tensor(np.stack([coo.row, coo.col]).astype(np.int64))
    split_edge = {}
    for name in ('train', 'valid', 'test'):
        pos, neg = splits[name]
        split_edge[name] = {'edge': torch.as_tensor(np.ascontiguousarray(pos.T)),
                            'edge_neg': torch.as_tensor(np.ascontiguousarray(neg.T))}
    data.train_pos_edge_index = split_edge['train']['edge'].t()
    data.train_neg_edge_index = split_edge['train']['edge_neg'].t()
    data.val_pos_edge_index, data.val_neg_edge_index = split_edge['valid']['edge'].t(), split_edge['valid']['edge_neg'].t()
    data.test_pos_edge_index, data.test_neg_edge_index = split_edge['test']['edge'].t(), split_edge['test']['edge_neg'].t()
    return split_edge


def _keep(count, percent):
    """Indices kept from a list of `count` links: a NumPy permutation (global RNG, as the reference draws it)
    cut to `percent` per cent.  One draw per call: the order of the calls fixes the link order."""
    order = np.random.permutation(count)
    return order[:int(percent / 100 * count)]


def get_pos_neg_edges(split, split_edge, edge_index, num_nodes, percent=100, neg_ratio=1, neg_seed=0):
    """Positive and negative target links of one split, [2, L] each, in the order the hot path will emit rows.

    Mirrors the behaviour of reference utils.py:637-678 for both split-dictionary layouts:
      * 'edge' layout — positives from `edge`, negatives from `edge_neg` when present (else sampled by
        `sample_negative_edges`); positives are subsampled first, then negatives, each with its own permutation;
      * 'source_node' layout (OGB citation style) — one permutation shared by source / target / negative targets;
        training negatives are drawn uniformly per source with torch.randint, evaluation negatives come from
        `target_node_neg`; every source is repeated once per negative target.
    With the same NumPy seed the link ORDER equals the reference's (tests/golden/posneg_edges_ref.npz)."""
    layout = split_edge['train']
    part = split_edge[split]
    if 'edge' in layout:
        pos = part['edge'].t()
        if 'edge_neg' in layout:
            neg = part['edge_neg'].t()
        else:
            neg = sample_negative_edges(edge_index, num_nodes, pos.size(1) * neg_ratio, neg_seed)
        pos = pos[:, _keep(pos.size(1), percent)]
        neg = neg[:, _keep(neg.size(1), percent)]
        return pos, neg
    if 'source_node' in layout:
        src, dst = part['source_node'], part['target_node']
        if split == 'train':
            dst_neg = torch.randint(0, num_nodes, [dst.size(0), 1], dtype=torch.long)
        else:
            dst_neg = part['target_node_neg']
        kept = _keep(src.size(0), percent)
        src, dst, dst_neg = src[kept], dst[kept], dst_neg[kept, :]
        pos = torch.stack([src, dst])
        neg = torch.stack([src.repeat_interleave(dst_neg.size(1)), dst_neg.reshape(-1)])
        return pos, neg
    raise KeyError("split_edge has neither 'edge' nor 'source_node' entries")


class SEALDataset:
    """sgrl_link_pred.py:54-220 for `args.model == 'SIGN'`.  `data` needs `.edge_index` [2, E] (both directions
    of the training edges), `.x` [N, F], `.num_nodes` and optionally `.edge_weight`; `args` needs `model`,
    `sign_k`, `optimize_sign`, `k_heuristic`, `k_node_set_strategy` (and `seed` for sampled negatives).
    After construction `self.lists` is the collated PrecomputedList (on `device`); `self.data, self.slices`
    are what the reference's `torch.load(self.processed_paths[0])` returns."""

    def __init__(self, root, data, split_edge, num_hops, percent=100, split='train', use_coalesce=False,
                 node_label='drnl', ratio_per_hop=1.0, max_nodes_per_hop=None, directed=False, rw_kwargs=None,
                 device='cuda', pairwise=False, pos_pairwise=False, neg_ratio=1, use_feature=False, sign_type="",
                 args=None):
        self.root = root
        self.data_in = data
        self.split_edge = split_edge
        self.num_hops = num_hops
        self.percent = int(percent) if percent >= 1.0 else percent
        self.split = split
        self.use_coalesce = use_coalesce
        self.node_label = node_label
        self.ratio_per_hop = ratio_per_hop
        self.max_nodes_per_hop = max_nodes_per_hop
        self.directed = directed
        self.device = device
        self.rw_kwargs = rw_kwargs or {}
        self.pairwise = pairwise
        self.pos_pairwise = pos_pairwise
        self.neg_ratio = neg_ratio
        self.use_feature = use_feature
        self.sign_type = sign_type
        self.args = args
        if getattr(args, 'model', 'SIGN') != 'SIGN':
            raise NotImplementedError("only the SIGN datasets are on the accelerated path (SURVEY.md §2)")
        os.makedirs(self.processed_dir, exist_ok=True)
        if not os.path.isfile(self.processed_paths[0]):
            self.process()
        self.lists = load_collated(self.processed_paths[0], device=device if str(device) != 'cpu' else None)
        self.data, self.slices = self.lists.collate()

    # --- InMemoryDataset surface the reference relies on ---
    @property
    def processed_dir(self):
        return os.path.join(self.root, 'processed')

    @property
    def processed_file_names(self):       # sgrl_link_pred.py:87-94
        name = f'SEAL_{self.split}_data' if self.percent == 100 else f'SEAL_{self.split}_data_{self.percent}'
        return [name + '.pt']

    @property
    def processed_paths(self):
        return [os.path.join(self.processed_dir, f) for f in self.processed_file_names]

    @property
    def num_features(self):               # sizes the MLP, models.py:316-320
        return int(self.lists.xs[0].shape[1])

    def __len__(self):
        return len(self.lists)

    def get(self, idx):
        return self.lists[idx]

    __getitem__ = get

    def process(self):                    # sgrl_link_pred.py:96-220
        d, args = self.data_in, self.args
        pos_edge, neg_edge = get_pos_neg_edges(self.split, self.split_edge, d.edge_index, d.num_nodes, self.percent,
                                               neg_ratio=self.neg_ratio, neg_seed=getattr(args, 'seed', 0))
        ei = d.edge_index.cpu().numpy()
        ew = getattr(d, 'edge_weight', None)
        edge_weight = np.ones(ei.shape[1], dtype=np.int64) if ew is None else np.asarray(ew.cpu()).reshape(-1)
        # csr_matrix sums duplicate entries, which is also what `coalesce` does when use_coalesce is set (:102-105)
        A = ssp.csr_matrix((edge_weight, (ei[0], ei[1])), shape=(d.num_nodes, d.num_nodes))
        rw_kwargs = None
        if self.rw_kwargs.get('m'):       # ScaLed: walks are sampled on the GPU inside the call (create_rw_cache, :121-127)
            rw_kwargs = {"rw_m": self.rw_kwargs.get('m'), "rw_M": self.rw_kwargs.get('M'), "sign": True,
                         "seed": getattr(args, 'seed', 0), "node_label": self.node_label}
        sign_kwargs = {"sign_k": args.sign_k, "use_feature": self.use_feature, "sign_type": self.sign_type,
                       "optimize_sign": args.optimize_sign, "k_heuristic": args.k_heuristic,
                       "k_node_set_strategy": args.k_node_set_strategy}
        # the global powers of sgrl_link_pred.py:161-178 are never formed: only their count is used (tuned_sign.py)
        powers_of_A = [None] * args.sign_k if self.sign_type in ('SoP', 'hybrid') else []

        def run(edges, y):
            return extract_enclosing_subgraphs(edges, A, d.x, y, self.num_hops, self.node_label, self.ratio_per_hop,
                                               self.max_nodes_per_hop, self.directed, None, rw_kwargs, sign_kwargs,
                                               powers_of_A=powers_of_A, data=d,
                                               device=None if str(self.device) == 'cpu' else self.device,
                                               # the operator matrices stay in HBM between the two calls and the collate
                                               output_device='cuda' if str(self.device) != 'cpu' else 'cpu')
        if not self.pairwise:
            out = run(pos_edge, 1) + run(neg_edge, 0)
        elif self.pos_pairwise:
            out = run(pos_edge, 1)
        else:
            out = run(neg_edge, 0)
        assert isinstance(out, PrecomputedList)
        save_collated(out, self.processed_paths[0])
