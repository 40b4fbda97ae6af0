"""Multi-GPU plumbing: one process per GPU, target links sharded in contiguous blocks, graph
replicated, and ONE collective — an allgather of the precomputed operator rows so every rank
holds the whole joint matrix for training.  The reference has no distributed
code at all; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is plumbing only.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist


def shard_range(num_links, rank, world_size):
    """Contiguous, balanced block of the link list owned by `rank`: sizes differ by at most 1
    and concatenating the shards in rank order restores the original order."""
    base, rem = divmod(int(num_links), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allgather_rows(xs, row_ptr, group=None):
    """All-gather row-stacked operator matrices and their row_ptr across ranks.

    xs: list of [R_r, C] tensors on this rank; row_ptr: int64 [L_r + 1].
    Returns (xs_full, row_ptr_full) identical on every rank, rows in rank order.  Equal
    shard sizes take the single-call all_gather_into_tensor path; ragged shards (PoS Plus)
    are padded to the largest shard and trimmed."""
    world = dist.get_world_size(group)
    dev = xs[0].device
    sizes = torch.tensor([xs[0].shape[0], row_ptr.shape[0] - 1], dtype=torch.int64, device=dev)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    rows = [int(s[0]) for s in all_sizes]
    nlinks = [int(s[1]) for s in all_sizes]

    def gather(t, counts):
        mx = max(counts)
        if all(c == mx for c in counts):
            full = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            dist.all_gather_into_tensor(full, t.contiguous(), group=group)
            return full
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pad[:t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        return torch.cat([p[:c] for p, c in zip(parts, counts)], 0)

    xs_full = [gather(x, rows) for x in xs]
    counts = gather((row_ptr[1:] - row_ptr[:-1]).contiguous(), nlinks)
    row_ptr_full = torch.zeros(counts.shape[0] + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=row_ptr_full[1:])
    return xs_full, row_ptr_full


def precompute_sharded(graph, links, num_hops, sign_k, flow='PoS', strategy=None, gather=True, group=None, **kw):
    """Each rank precomputes its contiguous shard of `links` on its own GPU (graph replicated),
    then (optionally) all ranks all-gather the result.  Per-link results do not depend on the
    shard, so the gathered matrices are bit-identical to a single-GPU run."""
    from .engine import precompute
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    a, b = shard_range(links.shape[1], rank, world)
    res = precompute(graph, links[:, a:b], num_hops, sign_k, flow, strategy, **kw)
    if not gather:
        return res.xs, res.row_ptr, res.stats
    xs, row_ptr = allgather_rows(res.xs, res.row_ptr, group)
    return xs, row_ptr, res.stats


# ----------------------------------------------------------------------------------------------
# Fused gather + all-gather over NVLink peer memory (the path's one exchange step)
# ----------------------------------------------------------------------------------------------
def cyclic_shard(num_links, rank, world_size):
    """Global link indices owned by `rank` under the cyclic partition i = rank (mod world): balanced over the
    link list's blocks (train positives, negatives, ...), over the skew of subgraph sizes and over the paired
    links that cost nothing.  The shards of all ranks partition range(num_links)."""
    return range(int(rank), int(num_links), int(world_size))


class _DevMem:
    """Raw device allocation presented through __cuda_array_interface__ so that torch can view it."""

    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = dict(shape=(int(nfloats),), typestr='<f4', data=(int(ptr), False), version=2)


class PeerBuffers:
    """The K+1 operator matrices [2 * num_links, F+1] of the WHOLE link list on every GPU of the node, each
    rank's copy reachable from every other rank.  s3_gather_peers stores every output row into all of them, so
    when the last kernel of a step has finished and the ranks have met at `barrier()`, every GPU holds the
    complete matrices: the all-gather of SURVEY.md §8e happens inside kernel 3, row by row, overlapped with the
    computation.  Two ways to reach the peers:

    'ipc'        (default) cudaMalloc + CUDA IPC handles (csrc/peer.cu) exchanged through torch.distributed: one
                 store per peer over NVLink P2P.  Measured steady (2 x B200: 15.1 ms per PubMed step, +-0.05).
    'multicast'  torch symmetric memory (cuMem VMM allocation + NVSwitch multicast object, plumbing only): ONE
                 store to the multicast address lands in every GPU's copy.  Bit-exact too, but the 4-byte-aligned row
                 stores through the switch were erratic on this pool (15.9 .. 48 ms per step on 2 GPUs), and a GPU
                 still has to RECEIVE (N-1)/N of the matrices either way: opt-in.  'auto' tries it first.

    Operator 0 (x itself = [1 | X[node]]) never crosses NVLink when local_x0 is set: every GPU holds X and fills
    those rows for the whole list with s3_fill_x0 (a quarter less traffic at sign_k = 3).

    .local     K+1 torch views [2 * num_links, F+1] of this GPU's copy
    .dst       device pointers kernel 3 stores to: [multicast address] or all ranks' copies"""

    def __init__(self, num_links, num_feat, sign_k, device, group=None, backend=None, local_x0=None, local_mirrors=None):
        from . import _lib as L
        # NVLink ingress only binds from 4 GPUs on (measured: profiles/README.md); on 2 GPUs the two extra local
        # passes cost more than the bytes they save
        # measured on 8 x B200: 5.37 vs 5.43 ms per PubMed step — the senders, not the receivers, are the limit; opt-in
        self.antiphase = os.environ.get('S3GRL_ANTIPHASE', '0') != '0'
        many = dist.get_world_size(group) >= 4
        self.local_x0 = many if local_x0 is None else bool(local_x0)        # operator 0 written locally (s3_fill_x0)
        self.local_mirrors = many if local_mirrors is None else bool(local_mirrors)   # paired links' rows copied locally
        self.flags = (L.PEERS_LOCAL_X0 if self.local_x0 else 0) | (L.PEERS_LOCAL_MIRRORS if self.local_mirrors else 0)
        if backend is None:    # measured on this pool (profiles/): P2P stores are steady and fastest on 2 GPUs, the
            backend = 'ipc' if dist.get_world_size(group) <= 2 else 'auto'     # multicast path wins from 4 GPUs on
        self._L, self._lib = L, L.lib()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > L.MAX_PEERS:
            raise NotImplementedError(f"s3_gather_peers serves up to {L.MAX_PEERS} GPUs of one NVLink domain")
        self.device = torch.device(device)
        self.num_links, self.cols, self.num_ops = int(num_links), int(num_feat) + 1, int(sign_k) + 1
        self.rows = 2 * self.num_links
        # row stride of the exchanged matrices: F + 1 (PyG's contiguous layout) or padded to a multiple of 32 floats, so
        # that every row starts on a 128-byte line and remote stores are whole lines (S3GRL_PEER_ROWS=pad; the .local
        # views are then strided [rows, F + 1] windows of [rows, ld])
        self.ld = (self.cols + 31) // 32 * 32 if os.environ.get('S3GRL_PEER_ROWS', 'packed') == 'pad' else self.cols
        self.op_stride = (self.rows * self.ld + 31) // 32 * 32        # floats; operators start on 128-byte lines
        self._nfloats = max(self.num_ops * self.op_stride, 64)
        self._opened, self._ptr, self._symm = [], None, None
        self.backend = None
        if backend in ('auto', 'multicast'):
            try:
                self._init_multicast()
            except Exception as ex:          # every rank must take the same branch: agree below
                self._symm, self._why_not_multicast = None, f"{type(ex).__name__}: {ex}"[:200]
            ok = torch.tensor([1 if self._symm is not None else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok) == 1:
                self.backend = 'multicast'
            else:
                self._symm = None
                if backend == 'multicast':
                    raise RuntimeError("NVSwitch multicast is not available: " + getattr(self, '_why_not_multicast', 'a peer failed'))
        if self.backend is None:
            self._init_ipc()
            self.backend = 'ipc'
        self.world_dst = len(self.dst)
        self.base_array = (C.c_void_p * self.world_dst)(*self.dst)
        self.local = [self._flat[k * self.op_stride:k * self.op_stride + self.rows * self.ld].view(self.rows, self.ld)[:, :self.cols]
                      for k in range(self.num_ops)]
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.barrier()

    def _init_multicast(self):
        import torch.distributed._symmetric_memory as symm_mem
        grp = self.group if self.group is not None else dist.group.WORLD
        t = symm_mem.empty(self._nfloats, dtype=torch.float32, device=self.device)
        hdl = symm_mem.rendezvous(t, grp)
        mc = int(hdl.multicast_ptr or 0)
        if not mc:
            raise RuntimeError("the symmetric-memory handle has no multicast address")
        self._symm, self._flat = hdl, t
        self.bases = [int(p) for p in hdl.buffer_ptrs]
        self.dst = [mc]

    def _init_ipc(self):
        L = self._L
        with torch.cuda.device(self.device):
            ptr = C.c_void_p()
            L.check(self._lib.s3_peer_alloc(self._nfloats * 4, C.byref(ptr)), 's3_peer_alloc')
            self._ptr = ptr.value
            handle = C.create_string_buffer(L.PEER_HANDLE_BYTES)
            L.check(self._lib.s3_peer_export(C.c_void_p(self._ptr), handle), 's3_peer_export')
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=self.group)
            self.bases = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.bases.append(self._ptr)
                    continue
                p = C.c_void_p()
                L.check(self._lib.s3_peer_open(C.create_string_buffer(h, L.PEER_HANDLE_BYTES), C.byref(p)), 's3_peer_open')
                self._opened.append(p.value)
                self.bases.append(p.value)
        self.dst = list(self.bases)
        self._mem = _DevMem(self._ptr, self._nfloats)
        self._flat = torch.as_tensor(self._mem, device=self.device)

    def barrier(self):
        """All ranks' kernels enqueued so far (on the current stream) have finished before any rank's later work
        on its current stream starts: one tiny NCCL all-reduce, no flag is spun on."""
        dist.all_reduce(self._flag, group=self.group)

    def close(self):
        if self._flat is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)           # nobody may still be storing into a buffer that goes away
        self.local = None
        if self.backend == 'ipc':
            with torch.cuda.device(self.device):
                for p in self._opened:
                    self._lib.s3_peer_close(C.c_void_p(p))
                self._flat = None
                self._lib.s3_peer_free(C.c_void_p(self._ptr))
        self._flat, self._symm, self._ptr, self._opened = None, None, None, []


def precompute_exchange(graph, links, num_hops, sign_k, buffers, flow='PoS', defer=False, **kw):
    """Fixed-row flows on all GPUs of the node, with the all-gather fused into kernel 3: every rank gets the
    WHOLE link list, pairs it (same table on every rank), takes the cyclic shard `rank (mod world)` of it and
    stores each of its output rows — and the rows of the links paired with them — into every rank's
    `buffers` at the row of the link's position in the whole list.  After `buffers.barrier()` (called here
    unless defer=True) buffers.local holds the complete operator matrices on every GPU, bit-identical to a
    single-GPU precompute of the whole list.  Returns (result of this rank's shard, mirror table)."""
    from .engine import pair_links, precompute
    dev = graph.device
    links = torch.as_tensor(links).to(device=dev, dtype=torch.int64).contiguous()
    n = int(links.shape[1])
    if n != buffers.num_links:
        raise ValueError("buffers were sized for another link list")
    mirror, table = (None, None)
    if flow == 'PoS' and n > 1 and not kw.get('walk') and kw.get('pair', True):
        mirror, table = pair_links(links, graph.num_nodes, kw.get('stream'))
    kw.pop('pair', None)
    if buffers.local_x0:       # x of every link, locally: the one operator that is a plain copy of resident data.
        # Independent of everything else in the step: it runs on a side stream beside the front / gather kernels
        # and is joined in exchange_finish.
        main = kw.get('stream') or torch.cuda.current_stream(dev)
        side = graph.streams()[0]
        side.wait_stream(main)
        with torch.cuda.device(dev):
            buffers._L.check(buffers._lib.s3_fill_x0(C.byref(graph._c), C.c_void_p(links[0].data_ptr()),
                                                     C.c_void_p(links[1].data_ptr()), n, C.c_void_p(buffers.local[0].data_ptr()),
                                                     buffers.ld, C.c_void_p(side.cuda_stream)), 's3_fill_x0')
        buffers._x0_done = torch.cuda.Event()
        buffers._x0_done.record(side)
        buffers._x0_links = links          # alive until the side stream has read them
    idx = torch.arange(buffers.rank, n, buffers.world, device=dev, dtype=torch.int64)
    if buffers.world >= 8 and kw.get('batch_records') is None and not kw.get('overlap') and buffers.antiphase:
        # NVLink ingress binds when all eight ranks run kernel 3 at once (eight senders ask a port for 0.98 GB/ms, it
        # takes 0.62): two batches per rank, even ranks front / gather / front / gather, odd ranks front / front /
        # gather / gather, so that mostly four ranks store at a time
        records = int(idx.numel()) * (2 if flow == 'SoP' else 1)
        kw['batch_records'] = (records + 1) // 2 + (1 if flow == 'SoP' else 0)
        kw['fronts_first'] = bool(buffers.rank & 1)
    res = precompute(graph, links[:, idx].contiguous(), num_hops, sign_k, flow, None, out_link=idx, mirror=mirror,
                     peers=buffers, pair=False, defer=defer, **kw)
    res._pair_table = table
    if not defer:
        exchange_finish(buffers, mirror, kw.get('stream'))
    return res, mirror


def exchange_finish(buffers, mirror, stream=None):
    """End of an exchange step, on the stream: the barrier (every rank's rows have landed everywhere), then — with
    local_mirrors — every GPU copies the rows of the paired links from their first link's rows in its own memory."""
    dev = buffers.device
    st = stream or torch.cuda.current_stream(dev)
    if getattr(buffers, '_x0_done', None) is not None:      # join the side stream that filled operator 0
        st.wait_event(buffers._x0_done)
        buffers._x0_done = None
    buffers.barrier()
    if mirror is not None and buffers.local_mirrors and buffers.num_ops > 1:
        ptrs = (C.c_void_p * buffers.num_ops)(*[o.data_ptr() for o in buffers.local])
        with torch.cuda.device(dev):
            buffers._L.check(buffers._lib.s3_fill_mirrors(C.c_void_p(mirror.data_ptr()), buffers.num_links, ptrs,
                                                          1 if buffers.local_x0 else 0, buffers.num_ops, buffers.cols, buffers.ld,
                                                          C.c_void_p(st.cuda_stream)), 's3_fill_mirrors')
