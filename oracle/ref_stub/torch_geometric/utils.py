"""Names the reference's utils.py imports at module load; none is on the hot path."""


def _unavailable(*a, **k):
    raise NotImplementedError("torch_geometric.utils stub: not on the hot path")


negative_sampling = add_self_loops = train_test_split_edges = to_networkx = subgraph = _unavailable
to_scipy_sparse_matrix = k_hop_subgraph = to_undirected = _unavailable
