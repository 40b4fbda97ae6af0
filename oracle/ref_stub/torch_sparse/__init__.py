"""SciPy-backed stand-in for the handful of `torch_sparse` calls the reference's hot path
makes (TEST INFRASTRUCTURE ONLY — see oracle/ref_stub/README.md).

Semantics followed (torch-sparse 0.6.13, the version the reference pins in
quick_install.sh:7), as used at the reference call sites:

* ``SparseTensor(row=, col=, value=None, sparse_sizes=)``           tuned_SIGN.py:155, :210
* ``.sum(dim=1)`` with no value  -> number of stored entries per row tuned_SIGN.py:158
* ``dense[n,1] * A`` scales rows, ``A * dense[1,n]`` scales columns; when A has no value
  the scale becomes the value                                        tuned_SIGN.py:161
* ``A @ B`` sparse-sparse (fp32 SpGEMM), ``A @ dense`` (fp32 SpMM)   tuned_SIGN.py:170, :185
* ``A[list_of_rows]`` row index_select                               tuned_SIGN.py:175
* ``A[i, j].to_dense()`` -> 1x1 dense                                tuned_SIGN.py:110
* ``.to_scipy()``; ``from_scipy``; ``spspmm``                        tuned_SIGN.py:70, :88, :94
"""
import numpy as np
import scipy.sparse as ssp
import torch


class _LilWithLilGetrow(ssp.lil_matrix):
    """SciPy 1.9.3 (the reference's pin) returns a 1xN *LIL* copy from ``getrow``; newer
    SciPy returns CSR, which breaks ``item.rows`` at tuned_SIGN.py:84."""

    def getrow(self, i):
        out = ssp.lil_matrix((1, self.shape[1]), dtype=self.dtype)
        out.rows[0] = list(self.rows[i])
        out.data[0] = list(self.data[i])
        return out


class _CooToLil(ssp.coo_matrix):
    def tolil(self, copy=False):
        base = ssp.coo_matrix(self).tolil()
        out = _LilWithLilGetrow(base.shape, dtype=base.dtype)
        out.rows, out.data = base.rows, base.data
        return out


class SparseTensor:
    def __init__(self, row=None, col=None, value=None, sparse_sizes=None, _csr=None, _has_value=None):
        if _csr is not None:
            self._m = _csr
            self._has_value = bool(_has_value)
            return
        row = np.asarray(row, dtype=np.int64)
        col = np.asarray(col, dtype=np.int64)
        if value is None:
            data = np.ones(row.shape[0], dtype=np.float32)
            self._has_value = False
        else:
            data = np.asarray(value, dtype=np.float32)
            self._has_value = True
        m = ssp.coo_matrix((data, (row, col)), shape=tuple(sparse_sizes)).tocsr()
        m.sort_indices()
        self._m = m

    # -- helpers ---------------------------------------------------------------------
    @classmethod
    def _wrap(cls, m, has_value=True):
        m = ssp.csr_matrix(m, dtype=np.float32)
        m.sort_indices()
        return cls(_csr=m, _has_value=has_value)

    def sizes(self):
        return list(self._m.shape)

    def size(self, dim):
        return self._m.shape[dim]

    def nnz(self):
        return int(self._m.nnz)

    # -- reductions ------------------------------------------------------------------
    def sum(self, dim=None):
        if dim is None:
            return torch.tensor(float(self._m.sum()))
        # no value -> torch_sparse counts stored entries; here duplicates were merged into one
        # entry whose data is the multiplicity, so summing the data gives the same count.
        out = np.asarray(self._m.sum(axis=dim)).reshape(-1)
        return torch.from_numpy(out.astype(np.float32))

    # -- elementwise scaling ---------------------------------------------------------
    def _scale(self, other):
        other = other.detach().cpu().numpy().astype(np.float32)
        m = self._m.copy()
        if other.ndim == 2 and other.shape[1] == 1 and other.shape[0] == m.shape[0]:
            m = ssp.diags(other[:, 0].astype(np.float32)).astype(np.float32) @ m
        elif other.ndim == 2 and other.shape[0] == 1 and other.shape[1] == m.shape[1]:
            m = m @ ssp.diags(other[0].astype(np.float32)).astype(np.float32)
        else:
            raise NotImplementedError(f"stub: unsupported broadcast shape {other.shape}")
        m = ssp.csr_matrix(m, dtype=np.float32)
        # keep explicitly stored zeros (a zero scale must not drop the entry's position)
        return SparseTensor._wrap(m, True)

    def __mul__(self, other):
        return self._scale(other)

    def __rmul__(self, other):
        return self._scale(other)

    # -- products --------------------------------------------------------------------
    def __matmul__(self, other):
        if isinstance(other, SparseTensor):
            return SparseTensor._wrap((self._m @ other._m).astype(np.float32), True)
        if isinstance(other, torch.Tensor):
            dense = other.detach().cpu().numpy().astype(np.float32)
            return torch.from_numpy(np.asarray(self._m @ dense, dtype=np.float32))
        return NotImplemented

    # -- indexing --------------------------------------------------------------------
    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            r, c = idx
            r = [int(r)] if np.isscalar(r) or isinstance(r, (int, np.integer)) else list(r)
            c = [int(c)] if np.isscalar(c) or isinstance(c, (int, np.integer)) else list(c)
            return SparseTensor._wrap(self._m[r, :][:, c], True)
        if isinstance(idx, torch.Tensor):
            idx = idx.tolist()
        if isinstance(idx, (int, np.integer)):
            idx = [int(idx)]
        return SparseTensor._wrap(self._m[list(idx), :], self._has_value)

    def to_dense(self):
        return torch.from_numpy(np.asarray(self._m.todense(), dtype=np.float32))

    def to_scipy(self, layout=None):
        return _CooToLil(self._m.tocoo())

    def coo(self):
        c = self._m.tocoo()
        return (torch.from_numpy(c.row.astype(np.int64)), torch.from_numpy(c.col.astype(np.int64)),
                torch.from_numpy(c.data.astype(np.float32)))


def from_scipy(m):
    c = ssp.coo_matrix(m)
    index = torch.from_numpy(np.vstack([c.row, c.col]).astype(np.int64))
    value = torch.from_numpy(np.asarray(c.data))
    return index, value


def spspmm(indexA, valueA, indexB, valueB, m, k, n, coalesced=False):
    A = ssp.coo_matrix((valueA.numpy().astype(np.float32), (indexA[0].numpy(), indexA[1].numpy())),
                       shape=(m, k)).tocsr()
    B = ssp.coo_matrix((valueB.numpy().astype(np.float32), (indexB[0].numpy(), indexB[1].numpy())),
                       shape=(k, n)).tocsr()
    C = (A @ B).astype(np.float32).tocoo()
    order = np.lexsort((C.col, C.row))
    index = torch.from_numpy(np.vstack([C.row[order], C.col[order]]).astype(np.int64))
    value = torch.from_numpy(C.data[order].astype(np.float32))
    return index, value
