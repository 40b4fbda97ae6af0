"""Output containers of the drop-in boundary.

The reference returns a Python list of PyG `Data` objects (keys x, y, x1..xK; tuned_SIGN.py:
117-133, :182-187) which SEALDataset.process concatenates and collates into K+1 row-stacked
matrices plus per-key slices (sgrl_link_pred.py:204).  The CUDA path produces that collated
layout directly; `PrecomputedList` presents it as the list the reference's callers expect
(len / index / iterate / `pos_list + neg_list`) without materialising one object per link,
and `collate()` hands back (data, slices) as InMemoryDataset.collate would.
"""
import torch

try:  # real PyG if present, else a minimal attribute bag with the same access pattern
    from torch_geometric.data import Data  # type: ignore
except Exception:  # pragma: no cover - PyG is absent in this image
    class Data:
        def __init__(self, **kwargs):
            self.__dict__.update(kwargs)

        def __getitem__(self, key):
            return self.__dict__[key]

        def __setitem__(self, key, value):
            self.__dict__[key] = value

        def __contains__(self, key):
            return key in self.__dict__

        def keys(self):
            return list(self.__dict__.keys())

        @property
        def num_nodes(self):
            return self.__dict__['x'].shape[0]

        def __repr__(self):
            return 'Data(' + ', '.join(f"{k}={list(v.shape) if torch.is_tensor(v) else v}" for k, v in self.__dict__.items()) + ')'


class PrecomputedList:
    """Sequence of per-link `Data` backed by collated tensors.

    xs      : list of K+1 tensors [R, F+1] (x, x1..xK)
    row_ptr : int64 [L+1]
    y       : int label shared by the call (reference passes one y per call, utils.py:446) or
              an int64 tensor [L] after concatenation."""

    def __init__(self, xs, row_ptr, y, stats=None, extras=None):
        self.xs = list(xs)
        self.extras = dict(extras or {})     # row-aligned extra keys (non-optimised flow: node_id)
        self.row_ptr = row_ptr
        L = int(row_ptr.shape[0]) - 1
        self.y = y if torch.is_tensor(y) else torch.full((L,), int(y), dtype=torch.long)
        self.stats = stats
        self._rp = None

    @property
    def keys(self):
        return ['x'] + [f'x{k}' for k in range(1, len(self.xs))]

    def __len__(self):
        return int(self.row_ptr.shape[0]) - 1

    def _bounds(self, i):
        if self._rp is None:
            self._rp = self.row_ptr.cpu().tolist()
        return self._rp[i], self._rp[i + 1]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        a, b = self._bounds(i)
        d = Data(x=self.xs[0][a:b], y=int(self.y[i]))
        for k in range(1, len(self.xs)):
            d[f'x{k}'] = self.xs[k][a:b]
        for key, t in self.extras.items():
            d[key] = t[a:b]
        return d

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __add__(self, other):
        """`pos_list + neg_list` (sgrl_link_pred.py:204)."""
        if not isinstance(other, PrecomputedList):
            return list(self) + list(other)
        assert len(self.xs) == len(other.xs)
        dev = self.xs[0].device
        xs = [torch.cat([a, b.to(dev)], 0) for a, b in zip(self.xs, other.xs)]
        rp = torch.cat([self.row_ptr, other.row_ptr[1:].to(self.row_ptr.device) + self.row_ptr[-1]])
        extras = {k: torch.cat([t, other.extras[k].to(t.device)]) for k, t in self.extras.items() if k in other.extras}
        return PrecomputedList(xs, rp, torch.cat([self.y, other.y]), extras=extras)

    def collate(self):
        """-> (data, slices) in the layout of InMemoryDataset.collate: every operator
        concatenated along dim 0 with its own [L+1] slice vector; y as an [L] tensor."""
        data = Data(x=self.xs[0], y=self.y)
        slices = {'x': self.row_ptr, 'y': torch.arange(len(self) + 1, dtype=torch.long)}
        for k in range(1, len(self.xs)):
            data[f'x{k}'] = self.xs[k]
            slices[f'x{k}'] = self.row_ptr
        for key, t in self.extras.items():
            data[key] = t
            slices[key] = self.row_ptr
        return data, slices

    def to(self, device, non_blocking=False):
        return PrecomputedList([x.to(device, non_blocking=non_blocking) for x in self.xs],
                               self.row_ptr.to(device), self.y, self.stats,
                               {k: t.to(device, non_blocking=non_blocking) for k, t in self.extras.items()})
