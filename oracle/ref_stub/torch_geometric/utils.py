"""Names the reference's utils.py imports at module load; none is on the hot path."""


def _unavailable(*a, **k):
    raise NotImplementedError("torch_geometric.utils stub: not on the hot path")


negative_sampling = add_self_loops = train_test_split_edges = to_networkx = subgraph = _unavailable
to_scipy_sparse_matrix = to_undirected = _unavailable


def k_hop_subgraph(node_idx, num_hops, edge_index, relabel_nodes=False, num_nodes=None, **kwargs):
    """Only the num_hops == 0 use of the reference's ScaLed branch (utils.py:124): the subset is
    the given node list itself; the SIGN path reads nothing but `subset` from the result."""
    import torch
    assert num_hops == 0
    subset = torch.as_tensor(node_idx, dtype=torch.long)
    return subset, torch.zeros((2, 0), dtype=torch.long), torch.arange(subset.numel()), None
