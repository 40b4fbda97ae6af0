"""Host orchestration of the CUDA precompute path (tensor-level API).

`DeviceGraph` uploads the training graph once (what the reference rebuilds per split at
sgrl_link_pred.py:107-114 as a SciPy CSR) and `precompute` runs a whole
get_PoS_prepped_ds / get_PoS_Plus_prepped_ds / get_SoP_prepped_ds call
(reference tuned_SIGN.py:137, :192, :49) on the GPU in batches of records:

    s3_extract -> [s3_plan -> sync -> s3_plan_items] -> s3_diffuse -> s3_gather

PyTorch is used for device memory, streams and (in parallel.py) torch.distributed only; all
compute is in libs3grl_b200.so.  There is no CPU fallback.
"""
import ctypes as C
import time

import numpy as np
import scipy.sparse as ssp
import torch

from . import _lib as L

_STRATEGY = {None: L.STRATEGY_NONE, '': L.STRATEGY_NONE, 'intersection': L.STRATEGY_INTERSECTION,
             'union': L.STRATEGY_UNION}
_FLOW = {'PoS': L.FLOW_POS, 'SoP': L.FLOW_SOP}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class DeviceGraph:
    """CSR + features resident in HBM.

    A : scipy sparse matrix N x N (any values; only the pattern is used by PoS,
        tuned_SIGN.py:153).  Must be structurally symmetric — the reference's training graphs
        always are (both directions of every edge, sgrl_link_pred.py:849-859).
    x : [N, F] float32 array / tensor (rows are padded to a multiple of 4 floats on device so
        every feature row is read with 128-bit loads).
    """

    def __init__(self, A, x, device='cuda', check_symmetric=True):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError("s3grl_b200 needs a CUDA device: there is no CPU path")
        L.lib()  # fail early and loudly if the extension is missing
        A = ssp.csr_matrix(A)
        A.sum_duplicates()
        A.sort_indices()
        if A.shape[0] != A.shape[1]:
            raise ValueError("A must be square")
        if check_symmetric and A.nnz and ((A != 0) != (A.T != 0)).nnz != 0:
            raise NotImplementedError("directed / structurally asymmetric graphs are not supported "
                                      "(reference flag `directed`, utils.py:58-63, is out of scope)")
        self.num_nodes = int(A.shape[0])
        self.nnz = int(A.nnz)
        self.has_multi_edges = bool(A.nnz and A.data.max() > 1)
        self.max_degree = int(np.diff(A.indptr).max()) if A.shape[0] else 0
        self.indptr = torch.from_numpy(A.indptr.astype(np.int64)).to(self.device)
        self.indices = torch.from_numpy(A.indices.astype(np.int32)).to(self.device)
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x, dtype=torch.float32)
        if x.dim() != 2 or x.shape[0] != self.num_nodes:
            raise ValueError(f"x must be [N={self.num_nodes}, F], got {tuple(x.shape)}")
        self.num_feat = int(x.shape[1])
        self.ldx = self.padded_row_stride(self.num_feat)
        self.x = torch.zeros((self.num_nodes, self.ldx), dtype=torch.float32, device=self.device)
        self.x[:, :self.num_feat].copy_(x, non_blocking=True)
        self.h2d_bytes = self.indptr.numel() * 8 + self.indices.numel() * 4 + x.numel() * 4
        self._c = L.Graph(_ptr(self.indptr), _ptr(self.indices), _ptr(self.x), self.num_nodes, self.num_feat,
                          self.ldx, self.nnz, self.max_degree)
        self._arena = None
        self._arena2 = None
        self._streams = None

    @staticmethod
    def padded_row_stride(F):
        """Row stride (floats) that gives every lane of kernel 3 a valid 16-byte column and starts
        every row on a 128-byte line: the kernel's (threads per row) x (columns per thread) tiling
        — 32/64/128 lanes x 1..3 float4 — mirrored here (gather.cu, launch_gather)."""
        f4 = (F + 3) // 4
        if f4 <= 32:
            return 128
        if f4 <= 64:
            return 256
        if f4 <= 128:
            return 512
        if f4 <= 256:
            return 1024
        return (f4 + 383) // 384 * 384 * 4

    @classmethod
    def from_device_csr(cls, indptr, indices, x):
        """Wrap a CSR that already lives on the GPU (int64 indptr, int32 indices with ascending unique
        columns, symmetric pattern — not checked) and features [N, F] on the same device."""
        self = cls.__new__(cls)
        self.device = indptr.device
        L.lib()
        self.num_nodes = int(indptr.numel() - 1)
        self.nnz = int(indices.numel())
        self.has_multi_edges = False
        self.max_degree = int((indptr[1:] - indptr[:-1]).max()) if self.num_nodes else 0
        self.indptr = indptr.to(torch.int64).contiguous()
        self.indices = indices.to(torch.int32).contiguous()
        self.num_feat = int(x.shape[1])
        self.ldx = cls.padded_row_stride(self.num_feat)
        if x.shape[1] == self.ldx and x.is_contiguous() and x.dtype == torch.float32:
            self.x = x
        else:
            self.x = torch.zeros((self.num_nodes, self.ldx), dtype=torch.float32, device=self.device)
            self.x[:, :self.num_feat].copy_(x)
        self.h2d_bytes = 0
        self._c = L.Graph(_ptr(self.indptr), _ptr(self.indices), _ptr(self.x), self.num_nodes, self.num_feat,
                          self.ldx, self.nnz, self.max_degree)
        self._arena = None
        self._arena2 = None
        self._streams = None
        return self

    # scratch arenas (int32 words), grown on demand and kept across calls; the second one is
    # only allocated by the overlapped (two-stream, double-buffered) schedule
    def arena(self, words, slot=0):
        name = '_arena' if slot == 0 else '_arena2'
        cur = getattr(self, name)
        if cur is None or cur.numel() < words:
            setattr(self, name, None)
            cur = torch.empty(int(words), dtype=torch.int32, device=self.device)
            setattr(self, name, cur)
        return cur

    def streams(self):
        if self._streams is None:
            self._streams = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        return self._streams


class PrecomputeResult:
    """Collated outputs of one precompute call:
    xs[k] is the row-stacked operator matrix x_k, [R, F+1]; row_ptr[i]..row_ptr[i+1] are the
    rows of link i (row 0 = src, row 1 = dst, then CCN rows ascending local id)."""

    def __init__(self, xs, row_ptr, stats, graphs=None):
        self.xs, self.row_ptr, self.stats, self.graphs = xs, row_ptr, stats, graphs
        self._finalize = None

    def finalize(self):
        """Complete a deferred call: sync, validate, re-run overflowed batches. Idempotent."""
        if self._finalize is not None:
            fn, self._finalize = self._finalize, None
            fn()
        return self


def _records_per_link(flow):
    return 2 if flow == L.FLOW_SOP else 1


def precompute(graph, links, num_hops, sign_k, flow='PoS', strategy=None, batch_records=32768,
               out=None, return_graphs=False, arena_words=None, stream=None, profile=None, overlap=False, defer=False, host_out=None, force_sorted_tier=False):
    """Run the hot path for `links` ([2, L] int64, host or device) on `graph`.

    Returns PrecomputeResult with device tensors.  `out`, if given, is a list of K+1
    preallocated [>=R, F+1] float32 device tensors (fixed-row flows only).  `profile`, if a
    list, receives (stage, batch, start_event, end_event) for every kernel launch group so the
    caller can time each kernel on the launching stream with CUDA events.  `overlap` (fixed-row
    flows) runs extract+diffuse of batch i+1 on one stream while gather of batch i runs on
    another, with two arenas: the latency-bound front half hides under the bandwidth-bound gather.
    `host_out` (fixed-row flows): K+1 pinned host tensors [>=R, F+1]; every batch's rows are
    copied device->host on a side stream as soon as its gather finishes, so the D2H of batch i
    overlaps the kernels of batch i+1 (the reference returns CPU tensors).
    `defer` (fixed-row flows) returns right after enqueueing; the caller must call
    `result.finalize()` (stream sync + validation + re-run of overflowed batches) before using
    the outputs.  It lets several calls be queued back to back without a host round trip.
    Raises ValueError for invalid links (out of range, src == dst), NotImplementedError for an
    unknown strategy (as reference tuned_SIGN.py:235)."""
    lib = L.lib()
    if flow not in _FLOW:
        raise NotImplementedError(f"sign_type {flow!r}: no matching configuration (reference utils.py:553)")
    if strategy not in _STRATEGY:
        raise NotImplementedError(f"check strat {strategy}")   # reference tuned_SIGN.py:235
    cflow = _FLOW[flow]
    cstrat = _STRATEGY[strategy] if cflow == L.FLOW_POS else L.STRATEGY_NONE
    if cflow == L.FLOW_SOP and graph.has_multi_edges:
        raise NotImplementedError("SoP on a multigraph (duplicate edges) is not supported")
    dev = graph.device
    links = torch.as_tensor(links)
    if links.dim() != 2 or links.shape[0] != 2:
        raise ValueError("links must be [2, L]")
    links = links.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
    Lk = int(links.shape[1])
    K = int(sign_k)
    F1 = graph.num_feat + 1
    rpl = _records_per_link(cflow)
    nseed = 1 if cflow == L.FLOW_SOP else 2
    fixed_rows = cstrat == L.STRATEGY_NONE
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    st_ptr = C.c_void_p(st.cuda_stream)

    with torch.cuda.device(dev), torch.cuda.stream(st):
        copy_stream = None
        if host_out is not None:
            if not fixed_rows or overlap:
                raise ValueError("host_out needs a fixed-row flow without overlap")
            copy_stream = graph.streams()[1]
            ready = torch.cuda.Event()
            ready.record(st)
            copy_stream.wait_event(ready)
        if fixed_rows:
            R = 2 * Lk
            if out is None:
                out = [torch.empty((R, F1), dtype=torch.float32, device=dev) for _ in range(K + 1)]
            else:
                assert len(out) == K + 1 and all(o.shape[0] >= R and o.shape[1] == F1 and o.is_contiguous() for o in out)
            out_ptrs = (C.c_void_p * (K + 1))(*[o.data_ptr() for o in out])
        if not fixed_rows:
            # data-dependent row counts: every batch ends in a host sync and keeps its own output
            # pieces, and CCN work items multiply the float scratch — smaller batches
            batch_records = min(int(batch_records), 4096 if cstrat == L.STRATEGY_UNION else 16384)
        batch_links = max(1, int(batch_records) // rpl)
        nb = (Lk + batch_links - 1) // batch_links
        counters = torch.zeros((max(nb, 1), L.NCTR), dtype=torch.int64, device=dev)
        if arena_words:
            words = int(arena_words)
        else:   # ~128 KiB of scratch per record to start with (PubMed h=3 averages ~100 KiB), grown on overflow
            free, _ = torch.cuda.mem_get_info(dev)
            words = max(1 << 22, min(min(int(batch_records), Lk * rpl) * 32768, free // 16))
            if graph._arena is not None:
                words = max(words, graph._arena.numel())
        batch_flags = (L.BATCH_STORE_ALL_ROWS if return_graphs else 0) | (L.BATCH_FORCE_SORTED_TIER if force_sorted_tier else 0)
        probe = L.Batch(None, None, 0, cflow, cstrat, int(num_hops), K, batch_flags, 0, None, 0, None, None, None, None, None,
                        None, None)
        if lib.s3_extract_tier(C.byref(graph._c), C.byref(probe)) < 0:
            L.check(L.S3_ERR_UNSUPPORTED, 's3_extract')
        words = max(words, 4 * int(lib.s3_min_arena_words(C.byref(graph._c), C.byref(probe))))
        stats = dict(records=Lk * rpl, links=Lk, sum_n=0, sum_d=0, max_n=0, rows=0, retries=0, batches=nb, launches=0)
        pieces, row_counts, graphs = [], [], ([] if return_graphs else None)

        def make_batch(b0, b1, arena, off, cnt, ctr, row_ptr=None, item_ptr=None, item_rec=None, order=None):
            return L.Batch(_ptr(links[0, b0:b1]), _ptr(links[1, b0:b1]), b1 - b0, cflow, cstrat, int(num_hops), K,
                           batch_flags, 0, _ptr(arena), arena.numel(), _ptr(off), _ptr(cnt), _ptr(ctr),
                           _ptr(row_ptr), _ptr(item_ptr), _ptr(item_rec), _ptr(order))

        def timed(stage, bi, fn, on=None):
            if profile is None:
                return fn()
            on = st if on is None else on
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(on)
            r = fn()
            e1.record(on)
            profile.append((stage, bi, e0, e1))
            return r

        def run_fixed_overlapped(todo, words):
            """Two streams, two arenas: front (extract, diffuse) of batch i+1 overlaps back (gather)
            of batch i.  Returns [(bi, cnt)] for validation after the caller's sync."""
            sF, sB = graph.streams()
            arenas = (graph.arena(words, 0), graph.arena(words, 1))
            nrecs = [(min(Lk, (bi + 1) * batch_links) - bi * batch_links) * rpl for bi in todo]
            off_all = torch.empty((sum(nrecs), L.NOFF), dtype=torch.int64, device=dev)
            cnt_all = torch.empty((sum(nrecs), L.NCNT), dtype=torch.int32, device=dev)
            order_all = torch.empty(sum(nrecs), dtype=torch.int32, device=dev)
            counters[torch.as_tensor(todo, device=dev)] = 0
            start = torch.cuda.Event()
            start.record(st)
            sF.wait_event(start)
            sB.wait_event(start)
            pF, pB = C.c_void_p(sF.cuda_stream), C.c_void_p(sB.cuda_stream)
            metas, back_done, r0 = [], [], 0
            for idx, bi in enumerate(todo):
                b0, b1 = bi * batch_links, min(Lk, (bi + 1) * batch_links)
                nrec = nrecs[idx]
                off, cnt = off_all[r0:r0 + nrec], cnt_all[r0:r0 + nrec]
                r0 += nrec
                order = order_all[r0 - nrec:r0]
                batch = make_batch(b0, b1, arenas[idx % 2], off, cnt, counters[bi], order=order)
                if idx >= 2:
                    sF.wait_event(back_done[idx - 2])          # this arena is free again
                timed('extract', bi, lambda: L.check(lib.s3_extract(C.byref(graph._c), C.byref(batch), pF), 's3_extract'), sF)
                fd = torch.cuda.Event()
                fd.record(sF)
                sB.wait_event(fd)
                timed('gather', bi, lambda: L.check(
                    lib.s3_gather(C.byref(graph._c), C.byref(batch), nrec, out_ptrs, F1, b0 * rpl * nseed, pB), 's3_gather'), sB)
                bd = torch.cuda.Event()
                bd.record(sB)
                back_done.append(bd)
                stats['launches'] += 2
                metas.append((bi, cnt))
            if back_done:
                st.wait_event(back_done[-1])
            return metas, (off_all, cnt_all, order_all)

        def run_batch(bi, arena):
            """Enqueue one batch; returns (cnt, off, pending) where pending finishes Plus flows."""
            b0, b1 = bi * batch_links, min(Lk, (bi + 1) * batch_links)
            nrec = (b1 - b0) * rpl
            off = torch.empty((nrec, L.NOFF), dtype=torch.int64, device=dev)
            cnt = torch.empty((nrec, L.NCNT), dtype=torch.int32, device=dev)
            ctr = counters[bi]
            ctr.zero_()
            if fixed_rows:
                order = torch.empty(nrec, dtype=torch.int32, device=dev)
                batch = make_batch(b0, b1, arena, off, cnt, ctr, order=order)
                timed('extract', bi, lambda: L.check(lib.s3_extract(C.byref(graph._c), C.byref(batch), st_ptr), 's3_extract'))
                timed('gather', bi, lambda: L.check(
                    lib.s3_gather(C.byref(graph._c), C.byref(batch), nrec, out_ptrs, F1, b0 * rpl * nseed, st_ptr), 's3_gather'))
                stats['launches'] += 2
                if host_out is not None:      # pipelined D2H of this batch's rows
                    r0, r1 = b0 * rpl * nseed, b1 * rpl * nseed
                    done = torch.cuda.Event()
                    done.record(st)
                    copy_stream.wait_event(done)
                    with torch.cuda.stream(copy_stream):
                        for k in range(K + 1):
                            host_out[k][r0:r1].copy_(out[k][r0:r1], non_blocking=True)
                return cnt, off, None
            row_ptr = torch.empty(nrec + 1, dtype=torch.int64, device=dev)
            item_ptr = torch.empty(nrec + 1, dtype=torch.int64, device=dev)
            order = torch.empty(nrec, dtype=torch.int32, device=dev)
            batch = make_batch(b0, b1, arena, off, cnt, ctr, row_ptr, item_ptr, order=order)
            timed('extract', bi, lambda: L.check(lib.s3_extract(C.byref(graph._c), C.byref(batch), st_ptr), 's3_extract'))
            L.check(lib.s3_plan(C.byref(batch), st_ptr), 's3_plan')
            stats['launches'] += 2
            c = ctr.cpu()                                  # sync: rows / items / errors of this batch
            if int(c[L.CTR_ERRORS]) != 0:
                return cnt, off, 'retry'
            rows, items = int(c[L.CTR_ROWS]), int(c[L.CTR_ITEMS])
            item_rec = torch.empty(max(items, 1), dtype=torch.int32, device=dev)
            xs_b = [torch.empty((rows, F1), dtype=torch.float32, device=dev) for _ in range(K + 1)]
            batch = make_batch(b0, b1, arena, off, cnt, ctr, row_ptr, item_ptr, item_rec, order)
            ptrs = (C.c_void_p * (K + 1))(*[o.data_ptr() for o in xs_b])
            timed('gather', bi, lambda: L.check(lib.s3_gather(C.byref(graph._c), C.byref(batch), nrec, ptrs, F1, 0, st_ptr), 's3_gather'))
            stats['launches'] += 1
            if items:       # CCN rows: extra work items of up to 8 selected rows each
                L.check(lib.s3_plan_items(C.byref(batch), st_ptr), 's3_plan_items')
                timed('diffuse', bi, lambda: L.check(lib.s3_diffuse(C.byref(graph._c), C.byref(batch), items, st_ptr), 's3_diffuse'))
                timed('gather_ccn', bi, lambda: L.check(
                    lib.s3_gather_ccn(C.byref(graph._c), C.byref(batch), items, ptrs, F1, 0, st_ptr), 's3_gather_ccn'))
                stats['launches'] += 3
            pieces.append(xs_b)
            row_counts.append(row_ptr[1:] - row_ptr[:-1])
            return cnt, off, None

        def check_and_account(bi, cnt, host_counters=None):
            c = counters[bi].cpu() if host_counters is None else host_counters[bi]
            if int(c[L.CTR_ERRORS]) != 0:
                status = cnt[:, L.CNT_STATUS]
                if bool((status == L.REC_BAD_LINK).any()):
                    bad = int(torch.nonzero(status == L.REC_BAD_LINK)[0]) // rpl + bi * batch_links
                    raise ValueError(f"invalid target link at position {bad}: node id out of range or src == dst")
                return False
            stats['sum_n'] += int(c[L.CTR_SUM_N])
            stats['sum_d'] += int(c[L.CTR_SUM_D])
            stats['max_n'] = max(stats['max_n'], int(c[L.CTR_MAX_N]))
            return True

        def dump(bi, cnt, off, arena):
            cn, of, ar = cnt.cpu().numpy(), off.cpu().numpy(), arena.cpu().numpy()
            for r in range(cn.shape[0]):
                n, m, s = int(cn[r, L.CNT_N]), int(cn[r, L.CNT_M]), int(cn[r, L.CNT_S])
                assert int(cn[r, L.CNT_NSTORE]) == n
                nodes = ar[of[r, L.OFF_NODES]:of[r, L.OFF_NODES] + n].astype(np.int64)
                rstart = ar[of[r, L.OFF_ROWPTR]:of[r, L.OFF_ROWPTR] + n + 1].astype(np.int64)
                rlen = ar[of[r, L.OFF_ROWLEN]:of[r, L.OFF_ROWLEN] + n].astype(np.int64)
                padded = ar[of[r, L.OFF_LCOL]:of[r, L.OFF_LCOL] + int(rstart[n])]
                rowptr = np.zeros(n + 1, dtype=np.int64)      # compact the padded CSR (holes are -1)
                np.cumsum(rlen, out=rowptr[1:])
                lcol = padded[padded >= 0].copy()
                owner = np.repeat(np.arange(n), np.diff(rstart))[padded >= 0]
                assert int(rowptr[n]) == m == lcol.size and np.array_equal(np.bincount(owner, minlength=n), rlen)
                sel = np.concatenate([np.arange(nseed), ar[of[r, L.OFF_SEL]:of[r, L.OFF_SEL] + s - nseed]]).astype(np.int32)
                hop_cnt = cn[r, L.CNT_HOP0:L.CNT_HOP0 + L.MAX_HOPS + 1]
                hops = np.repeat(np.arange(L.MAX_HOPS + 1), hop_cnt).astype(np.int32)
                graphs.append(dict(nodes=nodes, hops=hops, lrowptr=rowptr, lcol=lcol, sel=sel,
                                   partner=int(cn[r, L.CNT_PARTNER])))

        def grow():
            nonlocal words
            stats['retries'] += 1
            words = int(words * 2)
            graph._arena = graph._arena2 = None       # hand the old arenas back to the driver first
            torch.cuda.empty_cache()
            free, _ = torch.cuda.mem_get_info(dev)
            if words * 4 > free:
                raise MemoryError("scratch arena does not fit in device memory; lower batch_records")

        if fixed_rows:
            # Enqueue every batch without a host sync, validate at the end; a batch whose arena
            # overflowed is re-run (its output rows are simply rewritten) with a larger arena.
            use_overlap = bool(overlap) and not return_graphs and nb > 1

            def enqueue(todo):
                t_enq = time.perf_counter()
                if use_overlap:
                    metas, keep = run_fixed_overlapped(todo, words)
                else:
                    metas, keep = [], None
                    arena = graph.arena(words)
                    for bi in todo:
                        cnt, off, _ = run_batch(bi, arena)
                        if return_graphs:   # the arena is recycled by the next batch: dump now
                            st.synchronize()
                            if check_and_account(bi, cnt):
                                dump(bi, cnt, off, arena)
                                continue
                        metas.append((bi, cnt))
                stats['host_enqueue_ms'] = 1000 * (time.perf_counter() - t_enq)
                return metas, keep

            def settle(metas):
                """After a stream sync: account finished batches, return the ones to re-run."""
                hc = counters.cpu()       # one D2H for every batch's counters
                return [bi for bi, cnt in metas if not (False if return_graphs else check_and_account(bi, cnt, hc))]

            def finalize(metas):
                with torch.cuda.device(dev), torch.cuda.stream(st):
                    st.synchronize()
                    if copy_stream is not None:
                        copy_stream.synchronize()
                    todo = settle(metas)
                    while todo:
                        if return_graphs and graphs:
                            raise RuntimeError("arena overflow while dumping graphs: pass a larger arena_words")
                        grow()
                        metas, _keep = enqueue(todo)
                        st.synchronize()
                        if copy_stream is not None:
                            copy_stream.synchronize()
                        todo = settle(metas)

            metas0, keep0 = enqueue(list(range(nb)))
            if not defer:
                finalize(metas0)
        else:
            # Row counts are data dependent: one host sync per batch (inside run_batch).
            for bi in range(nb):
                while True:
                    arena = graph.arena(words)
                    cnt, off, pending = run_batch(bi, arena)
                    st.synchronize()
                    if pending is None and check_and_account(bi, cnt):
                        if return_graphs:
                            dump(bi, cnt, off, arena)
                        break
                    check_and_account(bi, cnt)      # raises on bad links
                    grow()

        if fixed_rows:
            xs = [o[:2 * Lk] for o in out]
            row_ptr = torch.arange(Lk + 1, dtype=torch.int64, device=dev) * 2
            stats['rows'] = 2 * Lk
        else:
            xs = [torch.cat([p[k] for p in pieces], 0) if pieces else torch.empty((0, F1), device=dev) for k in range(K + 1)]
            counts = torch.cat(row_counts) if row_counts else torch.zeros(0, dtype=torch.int64, device=dev)
            row_ptr = torch.zeros(Lk + 1, dtype=torch.int64, device=dev)
            torch.cumsum(counts, 0, out=row_ptr[1:])
            stats['rows'] = int(xs[0].shape[0])
    result = PrecomputeResult(xs, row_ptr, stats, graphs)
    if fixed_rows and defer:
        result._finalize = lambda: finalize(metas0)
        result._keep = keep0
    return result


def algorithmic_bytes(stats, num_feat, sign_k):
    """SURVEY.md §8d:  sum over links of 4·D + 8·n + 4·F·n + 4·s·(K+1)·(F+1)."""
    return (4 * stats['sum_d'] + 8 * stats['sum_n'] + 4 * num_feat * stats['sum_n']
            + 4 * stats['rows'] * (sign_k + 1) * (num_feat + 1))
