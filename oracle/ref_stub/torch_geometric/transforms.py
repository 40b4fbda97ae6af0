"""`SIGN` transform stand-in (base class of the reference's TunedSIGN, tuned_SIGN.py:13):
x_k = (D^-1/2 A D^-1/2)^k x over the stored edge_index, PyG semantics."""
import torch
from torch_sparse import SparseTensor


class SIGN:
    def __init__(self, K):
        self.K = K

    def __call__(self, data):
        assert data.edge_index is not None
        row, col = data.edge_index
        n = data.num_nodes
        adj_t = SparseTensor(row=col, col=row, sparse_sizes=(n, n))
        deg = adj_t.sum(dim=1).to(torch.float)
        dis = deg.pow(-0.5)
        dis[dis == float('inf')] = 0
        adj_t = dis.view(-1, 1) * adj_t * dis.view(1, -1)
        xs = [data.x]
        for i in range(1, self.K + 1):
            xs.append(adj_t @ xs[-1])
            data[f'x{i}'] = xs[-1]
        return data
