"""GPU-resident loader and on-disk format for the precomputed SIGN datasets.

The reference stores the collated dataset with `torch.save(self.collate(pos_list + neg_list), path)`
(sgrl_link_pred.py:204), reloads it as `(data, slices)` (:85) and trains from
`DataLoader(dataset, batch_size, shuffle, follow_batch=['x1'..'xK'])` (:1253-1269): every batch is
re-assembled on the CPU by Batch.from_data_list, moved to the GPU (:445) and concatenated feature-wise
at the top of SIGNNet.forward (models.py:372).  Here the collated dataset stays in HBM as the precompute
path wrote it and `JointLoader` assembles batches there with one CUDA kernel (`s3_joint_rows`,
csrc/loader.cu) that writes the joint matrix [rows, (K+1)*F'] directly — one launch per epoch, batches
are then zero-copy slices.  There is no CPU path.
"""
import ctypes as C

import torch

from . import _lib as L
from .data import Data, PrecomputedList


def save_collated(dataset, path):
    """`torch.save` of (data, slices) in the layout InMemoryDataset.collate produces (x, x1..xK
    concatenated along dim 0, y [L], one [L+1] slice vector per key).  PyG's `Data` class is not
    importable in this image, so `data` is stored as a plain dict of tensors; on a machine with PyG,
    `Data.from_dict(data)` is the object the reference's `torch.load(self.processed_paths[0])` returns."""
    data, slices = dataset.collate()
    fields = data.to_dict() if hasattr(data, 'to_dict') else data.__dict__
    payload = ({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in fields.items()},
               {k: v.cpu() for k, v in slices.items()})
    torch.save(payload, path)


def load_collated(path, device=None):
    """Inverse of save_collated -> PrecomputedList (on `device` if given)."""
    data, slices = torch.load(path)
    keys = ['x'] + [f'x{k}' for k in range(1, 64) if f'x{k}' in data]
    xs = [data[k] for k in keys]
    extras = {k: v for k, v in data.items() if k not in keys and k != 'y' and torch.is_tensor(v)}
    out = PrecomputedList(xs, slices['x'], data['y'], extras=extras)
    return out.to(device) if device is not None else out


def joint_rows(xs, row_ptr, link_idx, rows_per_link=None, out=None, want_batch=True, stream=None):
    """Joint matrix of the listed links: -> (joint [R_out, (K+1)*F'] (a view whose row stride is padded to a multiple
    of 4 floats), batch int64 [R_out] or None, ptr int64 [B+1] first row of every listed link).  `xs`: K+1 collated [R, F'] float32 CUDA tensors,
    `row_ptr` int64 [L+1], `link_idx` int64 [B] (device).  `rows_per_link`: pass 2 for the fixed-row
    flows (PoS, SoP) to skip the scan and the host sync."""
    lib = L.lib()
    dev = xs[0].device
    if dev.type != 'cuda':
        raise RuntimeError("joint_rows needs CUDA tensors: there is no CPU path")
    F1 = int(xs[0].shape[1])
    nops = len(xs)
    if not all(x.dtype == torch.float32 and x.is_contiguous() and x.shape == xs[0].shape for x in xs):
        raise ValueError("xs must be contiguous float32 tensors of one shape")
    link_idx = link_idx.to(device=dev, dtype=torch.int64).contiguous()
    row_ptr = row_ptr.to(device=dev, dtype=torch.int64).contiguous()
    B = int(link_idx.numel())
    if rows_per_link:
        ptr = torch.arange(B + 1, dtype=torch.int64, device=dev) * int(rows_per_link)
        R, optr = B * int(rows_per_link), None
    else:
        counts = row_ptr[link_idx + 1] - row_ptr[link_idx]
        ptr = torch.zeros(B + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=ptr[1:])
        R, optr = int(ptr[-1]), ptr        # host sync: data-dependent row count
    kd = nops * F1
    ld = (kd + 3) // 4 * 4          # 16-byte row stride: what TMA (the fused scoring head) needs
    if out is None:
        out = torch.empty((R, ld), dtype=torch.float32, device=dev)
    elif out.dim() != 2 or out.shape[0] < R or out.shape[1] < kd or out.stride(1) != 1 or out.dtype != torch.float32:
        raise ValueError("out must be a [>= R_out, >= (K+1)*F'] float32 tensor with unit column stride")
    ld = int(out.stride(0))
    batch = torch.empty(R, dtype=torch.int64, device=dev) if want_batch else None
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    ptrs = (C.c_void_p * nops)(*[x.data_ptr() for x in xs])
    with torch.cuda.device(dev):
        L.check(lib.s3_joint_rows(ptrs, nops, F1, F1, C.c_void_p(row_ptr.data_ptr()), C.c_void_p(link_idx.data_ptr()), B,
                                  C.c_void_p(optr.data_ptr()) if optr is not None else C.c_void_p(0),
                                  int(rows_per_link or 0), C.c_void_p(out.data_ptr()), ld,
                                  C.c_void_p(batch.data_ptr()) if batch is not None else C.c_void_p(0),
                                  C.c_void_p(st.cuda_stream)), 's3_joint_rows')
    return out[:R, :kd], batch, ptr


class JointLoader:
    """Iterates a device-resident `PrecomputedList` in batches the way the reference's
    DataLoader(..., shuffle, follow_batch=[x1..xK]) does, without leaving the GPU.

    Every epoch: one permutation (torch.randperm on the device, seeded), ONE `s3_joint_rows` launch for
    the whole epoch, then batches are slices of the epoch's joint matrix.  A batch is a `Data` with
        joint       [rows, (K+1)*F']   == torch.cat([x, x1, .., xK], -1) of models.py:372
        x, x1..xK   column views of `joint`
        batch, x{k}_batch              link-in-batch index of every row (PyG follow_batch)
        ptr         [B+1]              first row of every link (== np.unique(batch, return_index) of
                                       models.py:341, so center pooling needs no host round trip)
        y           [B]                labels,  num_graphs = B
    """

    def __init__(self, dataset, batch_size, shuffle=False, seed=0, drop_last=False, reuse_epoch_buffer=False):
        if dataset.xs[0].device.type != 'cuda':
            raise RuntimeError("JointLoader needs the dataset on a CUDA device (PrecomputedList.to('cuda'))")
        self.ds, self.batch_size, self.shuffle, self.drop_last = dataset, int(batch_size), bool(shuffle), bool(drop_last)
        self.dev = dataset.xs[0].device
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(int(seed))
        self.row_ptr = dataset.row_ptr.to(self.dev)
        self.y = dataset.y.to(self.dev)
        counts = self.row_ptr[1:] - self.row_ptr[:-1]
        L_ = len(dataset)
        self.fixed = int(counts[0]) if L_ and bool((counts == counts[0]).all()) else None
        # Batches are views of the epoch's joint matrix.  By default every __iter__ allocates a fresh matrix, so batches
        # (and tensors saved for backward) stay valid across epochs and concurrent iterators, as PyG's DataLoader
        # batches do; reuse_epoch_buffer=True keeps ONE matrix for all epochs (no allocation per epoch) and the
        # batches of an epoch are invalidated by the next __iter__.
        self.reuse_epoch_buffer = bool(reuse_epoch_buffer)
        self._epoch_buf = None

    def __len__(self):
        n = len(self.ds)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n, bs = len(self.ds), self.batch_size
        perm = torch.randperm(n, device=self.dev, generator=self.gen) if self.shuffle else torch.arange(n, device=self.dev)
        K1, F1 = len(self.ds.xs), int(self.ds.xs[0].shape[1])
        buf = None
        if self.fixed and self.reuse_epoch_buffer:
            if self._epoch_buf is None or self._epoch_buf.shape[0] < n * self.fixed:
                self._epoch_buf = torch.empty((n * self.fixed, (K1 * F1 + 3) // 4 * 4), dtype=torch.float32, device=self.dev)
            buf = self._epoch_buf
        joint, _, ptr = joint_rows(self.ds.xs, self.row_ptr, perm, self.fixed, out=buf, want_batch=False)
        y = self.y[perm]
        ptr_host = None if self.fixed else ptr[::bs].cpu().tolist() + [int(ptr[-1])]
        for bi in range(len(self)):
            l0, l1 = bi * bs, min(n, (bi + 1) * bs)
            if self.fixed:
                r0, r1 = l0 * self.fixed, l1 * self.fixed
            else:
                r0, r1 = ptr_host[bi], (ptr_host[bi + 1] if l1 < n else ptr_host[-1])
            bptr = ptr[l0:l1 + 1] - r0
            counts = bptr[1:] - bptr[:-1]
            batch = torch.repeat_interleave(torch.arange(l1 - l0, device=self.dev), counts, output_size=r1 - r0)
            jm = joint[r0:r1]
            d = Data(joint=jm, x=jm[:, :F1], y=y[l0:l1], batch=batch, ptr=bptr, num_graphs=l1 - l0)
            for k in range(1, K1):
                d[f'x{k}'] = jm[:, k * F1:(k + 1) * F1]
                d[f'x{k}_batch'] = batch
            yield d
