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
import os
import time

import numpy as np
import scipy.sparse as ssp
import torch

from . import _lib as L

_NVTX = bool(os.environ.get('S3GRL_NVTX'))      # read once at import: NVTX ranges around every kernel launch
# union chain: records whose shared-memory placement would run at a sub-chunk width <= this (16: n > ~860) take their two
# operator buffers from a global pool at width 32 instead (s3_ccn_chain_pooled; read through L2) and form y_0 from the
# feature matrix on the fly.  Measured on the PubMed union step, chain ms: 0 (all in shared memory) 356, 8: 285, 16: 280,
# 32: 330 (profiles/README.md).  0 switches it off.
_CHAIN_POOL_CW = int(os.environ.get('S3GRL_CHAIN_POOL_CW', '16') or 0)
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

    def ensure_size_proxy(self):
        """Per-node size proxy (s3_node_proxy: degree + neighbours' degrees) for the longest-first hand-out of records
        in the front kernel.  Built once per graph on first use of the bitmap tier; results do not depend on it."""
        if getattr(self, '_proxy', None) is None:
            self._proxy = torch.empty(self.num_nodes, dtype=torch.int32, device=self.device)
            with torch.cuda.device(self.device):
                L.check(L.lib().s3_node_proxy(C.byref(self._c), _ptr(self._proxy),
                                              C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), 's3_node_proxy')
            self._c.size_proxy = self._proxy.data_ptr()
        return self._proxy

    def ensure_hub_index(self, max_hubs=None, min_degree=128, force=False):
        """Hub index of the sorted-set tier (include/s3grl_b200.h, s3_graph.hub_*): the `max_hubs` highest-degree
        nodes of degree >= min_degree get an id and a hub x hub adjacency bit matrix (max_hubs^2 / 8 bytes: 302 MB at
        the default), so that the adjacency of two hubs inside a subgraph is one bit probe instead of a binary search
        in a list of up to 10^5 entries.  Built once per graph, on first use of the sorted tier.  Results do not
        depend on it."""
        if getattr(self, '_hub', None) is not None and not force:
            return self._hub[2]
        if max_hubs is None:
            # bit matrix of max_hubs^2 / 8 bytes: 0.3 GB; 8.6 GB for graphs of millions of nodes when the GPU has the
            # room (B200: 180 GB).  Measured on the 10 M-node R-MAT: 49 152 hubs 2.73 M links/s, 131 072 3.19 M, 262 144 3.46 M
            big = self.num_nodes > 2_000_000 and torch.cuda.mem_get_info(self.device)[0] > (48 << 30)
            max_hubs = int(os.environ.get('S3GRL_MAX_HUBS', 262144 if big else 49152))
        deg = self.indptr[1:] - self.indptr[:-1]
        k = min(int(max_hubs), self.num_nodes)
        top, idx = torch.topk(deg, k)
        idx = idx[top >= int(min_degree)]
        H = int(idx.numel())
        if H < 2:
            self._hub = (None, None, 0)
            return 0
        hub_id = torch.full((self.num_nodes,), -1, dtype=torch.int32, device=self.device)
        hub_id[idx] = torch.arange(H, dtype=torch.int32, device=self.device)
        bits = torch.zeros(H * ((H + 31) // 32), dtype=torch.int32, device=self.device)
        self._hub = (hub_id, bits, H)
        self.hub_min_degree = int(top[H - 1])
        self._c.hub_id, self._c.hub_bits, self._c.num_hubs = hub_id.data_ptr(), bits.data_ptr(), H
        with torch.cuda.device(self.device):
            L.check(L.lib().s3_build_hub_bits(C.byref(self._c), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                    's3_build_hub_bits')
        return H

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

    # pool of the union chain's large records (s3_ccn_chain_pooled): two slots per SM, each for two [n][32] buffers of
    # the largest record the chain serves (n <= 5 200); allocated on first use, kept across calls
    CHAIN_SLOT_FLOATS = 64 * 5216

    def chain_pool(self):
        if getattr(self, '_chain_pool', None) is None:
            slots = 2 * torch.cuda.get_device_properties(self.device).multi_processor_count
            self._chain_pool = (torch.empty(slots * self.CHAIN_SLOT_FLOATS, dtype=torch.float32, device=self.device),
                                torch.zeros(slots, dtype=torch.int32, device=self.device), slots)
        return self._chain_pool

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


def walk_sets(graph, starts, rw_m, rw_M, seed=0, stream=None):
    """ScaLed (reference utils.py:425-443, create_rw_cache): sorted node sets of `rw_M` uniform random
    walks of length `rw_m` from every node in `starts` (int64 tensor).  Returns (sets int32 [S, cap],
    counts int32 [S]) on the graph's device.  Counter-based RNG: a node's set depends only on
    (seed, node), never on the other nodes of the call."""
    lib = L.lib()
    dev = graph.device
    starts = torch.as_tensor(starts).to(device=dev, dtype=torch.int64).contiguous()
    S = int(starts.numel())
    cap = 1 + int(rw_M) * int(rw_m)
    sets = torch.empty((S, cap), dtype=torch.int32, device=dev)
    counts = torch.empty(S, dtype=torch.int32, device=dev)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        L.check(lib.s3_walk_sets(C.byref(graph._c), _ptr(starts), S, int(rw_m), int(rw_M), int(seed) & (2**64 - 1), cap,
                                 _ptr(sets), _ptr(counts), C.c_void_p(st.cuda_stream)), 's3_walk_sets')
    return sets, counts


def walk_sets_from_cache(cache, nodes, device):
    """Table for `nodes` (1-D int64 tensor) from a reference-style cache {node id: 1-D tensor of the
    unique nodes its walks visited} (what create_rw_cache returns)."""
    lists = [torch.unique(torch.as_tensor(cache[int(v)]).flatten().to(torch.int64)).to(torch.int32) for v in nodes.tolist()]
    cap = max([int(t.numel()) for t in lists] + [1])
    sets = torch.zeros((len(lists), cap), dtype=torch.int32)
    for i, t in enumerate(lists):
        sets[i, :t.numel()] = t
    counts = torch.tensor([int(t.numel()) for t in lists], dtype=torch.int32)
    return sets.to(device), counts.to(device)


def _free_bytes(dev):
    """Free device memory; not queried while a CUDA graph is being captured (sizes were settled by the warm-up)."""
    if torch.cuda.is_current_stream_capturing():
        return 1 << 62
    return torch.cuda.mem_get_info(dev)[0]


def pair_links(links, num_nodes, stream=None):
    """s3_pair_links (csrc/pair.cu) over a device link list [2, L] int64 -> (mirror int64 [L], scratch table).
    mirror is the chain table described in include/s3grl_b200.h: >= 0 first link of a node pair with the first
    member of its chain, -1 unpaired, <= -2 member (served by the first link's record).  Keep `table` alive
    until the stream has run the two kernels."""
    lib = L.lib()
    if links.dim() != 2 or links.shape[0] != 2 or links.dtype != torch.int64 or not links.is_cuda:
        raise ValueError("links must be a [2, L] int64 CUDA tensor")
    links = links.contiguous()
    dev = links.device
    n = int(links.shape[1])
    slots = int(lib.s3_pair_table_slots(n))
    table = torch.empty(2 * slots, dtype=torch.int64, device=dev)
    mirror = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        L.check(lib.s3_pair_links(_ptr(links[0]), _ptr(links[1]), n, int(num_nodes), _ptr(table), slots, _ptr(mirror),
                                  C.c_void_p(st.cuda_stream)), 's3_pair_links')
    return mirror[:n] if n else mirror[:0], table


class _Call:
    """One precompute call: validated arguments, per-call device state and the batch schedule.

    Fixed-row flows (PoS, SoP: every record owns `num_seeds` output rows) enqueue all batches
    without a host sync, validate once at the end and re-run batches whose arena overflowed.
    PoS Plus has data-dependent row counts: every batch ends in a host sync (rows / CCN items),
    owns its output pieces, and the pieces are concatenated at the end."""

    def __init__(self, graph, links, num_hops, sign_k, flow, strategy, batch_records, out, return_graphs,
                 arena_words, stream, profile, overlap, host_out, force_sorted_tier, walk=None, ccn_mode=None,
                 pair=True, out_link=None, mirror=None, peers=None, ratio_per_hop=1.0, max_nodes_per_hop=None, cap_seed=0,
                 fronts_first=False, compat_explicit_zero=False):
        self.lib = L.lib()
        self.fronts_first = bool(fronts_first)
        self.compat_explicit_zero = bool(compat_explicit_zero)
        if self.compat_explicit_zero and not (flow == 'PoS' and strategy == 'union'):
            raise ValueError("compat_explicit_zero only changes the PoS Plus union row selection")
        # per-hop caps of the BFS (reference utils.py:66-70) with the deterministic rank rule of include/s3grl_b200.h
        self.cap_ratio = 1.0 if ratio_per_hop is None else float(ratio_per_hop)
        self.cap_max = 0 if max_nodes_per_hop is None else int(max_nodes_per_hop)
        self.cap_seed = int(cap_seed) & 0xFFFFFFFF
        if self.cap_ratio <= 0.0 or self.cap_max < 0:
            raise ValueError("ratio_per_hop must be in (0, 1] and max_nodes_per_hop positive (or None)")
        if ccn_mode not in (None, 'items', 'chain'):
            raise ValueError("ccn_mode must be None, 'items' or 'chain'")
        self.ccn_mode = ccn_mode
        if flow not in _FLOW:
            raise NotImplementedError(f"sign_type {flow!r}: no matching configuration (reference utils.py:553)")
        if strategy not in _STRATEGY:
            raise NotImplementedError(f"check strat {strategy}")   # reference tuned_SIGN.py:235
        self.graph, self.dev = graph, graph.device
        self.flow = _FLOW[flow]
        self.strategy = _STRATEGY[strategy] if self.flow == L.FLOW_POS else L.STRATEGY_NONE
        if self.flow == L.FLOW_SOP and graph.has_multi_edges:
            raise NotImplementedError("SoP on a multigraph (duplicate edges) is not supported")
        links = torch.as_tensor(links)
        if links.dim() != 2 or links.shape[0] != 2:
            raise ValueError("links must be [2, L]")
        self.links = links.to(device=self.dev, dtype=torch.int64, non_blocking=True).contiguous()
        self.num_links = int(self.links.shape[1])
        self.num_hops, self.K = int(num_hops), int(sign_k)
        if self.K > L.MAX_K_UNION and strategy == 'union' and flow == 'PoS':
            raise NotImplementedError(f"PoS Plus union is built for sign_k <= {L.MAX_K_UNION}")
        # ScaLed: per-endpoint random-walk sets replace the h-hop ball
        self.walk = None
        if walk is not None:
            if self.flow != L.FLOW_POS or self.strategy != L.STRATEGY_NONE:
                raise NotImplementedError("random-walk subgraphs are supported for PoS without CCN rows only")
            uniq, inv = torch.unique(self.links.reshape(-1), return_inverse=True)
            inv = inv.reshape(2, -1).contiguous()
            if 'cache' in walk:
                sets, counts = walk_sets_from_cache(walk['cache'], uniq.cpu(), self.dev)
            elif 'sets' in walk:      # table indexed by node id
                sets, counts = walk['sets'].to(self.dev)[uniq].contiguous(), walk['counts'].to(self.dev)[uniq].contiguous()
            else:
                sets, counts = walk_sets(graph, uniq, walk['m'], walk['M'], walk.get('seed', 0), stream)
            self.walk = dict(sets=sets.contiguous(), counts=counts.contiguous(), src=inv[0].contiguous(),
                             dst=inv[1].contiguous(), cap=int(sets.shape[1]))
        self.F1 = graph.num_feat + 1
        self.rpl = 2 if self.flow == L.FLOW_SOP else 1          # records per link
        self.nseed = 1 if self.flow == L.FLOW_SOP else 2        # output rows per record (fixed-row flows)
        self.fixed_rows = self.strategy == L.STRATEGY_NONE
        # CCN rows of `union`: the hop-limited SpMM chain (s3_ccn_chain) for every record that fits its shared-memory
        # placement (n up to ~4 000) and CCN work items (s3_diffuse + s3_gather_ccn) for the rest; ccn_mode='items'
        # sends everything through the work items (round 1's default, ~4x slower on PubMed).  The chain converts the
        # stored CSR in place, so a parity dump (return_graphs) keeps the work items.
        if self.ccn_mode == 'chain' and self.strategy != L.STRATEGY_UNION:
            raise NotImplementedError("ccn_mode='chain' serves the union strategy only")
        if self.ccn_mode == 'chain' and return_graphs:
            raise NotImplementedError("ccn_mode='chain' rewrites the stored CSR in place: not with return_graphs")
        self.ccn_chain = self.strategy == L.STRATEGY_UNION and (self.ccn_mode == 'chain' or (self.ccn_mode is None and not return_graphs))
        self.return_graphs, self.profile = bool(return_graphs), profile
        self.stream = stream if stream is not None else torch.cuda.current_stream(self.dev)
        self.stream_ptr = C.c_void_p(self.stream.cuda_stream)
        self.host_out, self.copy_stream = host_out, None
        if host_out is not None and (not self.fixed_rows or overlap):
            raise ValueError("host_out needs a fixed-row flow without overlap")
        if batch_records is None:
            # fixed rows: 32 Ki records amortise the launch tails.  PoS Plus ends every batch with a host
            # sync, so intersection (small scratch) takes everything in as few batches as memory allows;
            # union multiplies the float scratch by the CCN work items and stays small.
            free = _free_bytes(self.dev)
            batch_records = {L.STRATEGY_NONE: 32768, L.STRATEGY_INTERSECTION: 262144,
                             L.STRATEGY_UNION: 32768 if self.ccn_chain else 8192}[self.strategy]
            batch_records = max(1024, min(batch_records, int(free // 3 // (32768 * 4))))
        self.batch_links = max(1, int(batch_records) // self.rpl)
        self.set_batches()
        self.overlap = bool(overlap) and self.fixed_rows and not self.return_graphs and self.num_batches > 1
        self.flags = ((L.BATCH_SHARE_SMS if (self.overlap and peers is not None) else 0)
                      | (L.BATCH_STORE_ALL_ROWS if self.return_graphs else 0)
                      | (L.BATCH_FORCE_SORTED_TIER if force_sorted_tier else 0)
                      | (L.BATCH_CCN_CHAIN if self.ccn_chain else 0))
        probe = self.make_batch(0, 0, None, None, None, None)
        if self.lib.s3_extract_tier(C.byref(graph._c), C.byref(probe)) < 0:
            L.check(L.S3_ERR_UNSUPPORTED, 's3_extract')
        if arena_words:
            words = int(arena_words)
        else:   # ~128 KiB of scratch per record to start with (PubMed h=3 averages ~60 KiB), grown on overflow
            free = _free_bytes(self.dev)
            words = max(1 << 22, min(min(int(batch_records), self.num_links * self.rpl) * 32768, free // 16))
            if graph._arena is not None:
                words = max(words, graph._arena.numel())
        self.words = max(words, 4 * int(self.lib.s3_min_arena_words(C.byref(graph._c), C.byref(probe))))
        self.out, self.out_ptrs = out, None
        # link pairing (csrc/pair.cu): one record per unordered node pair, the other direction is a row swap.
        # Fixed-row PoS on the bitmap tier only; `out_link` / `mirror` / `peers` come from parallel.precompute_exchange
        # (this call then holds a subset of a larger link list and writes rows at their global positions).
        tier = int(self.lib.s3_extract_tier(C.byref(graph._c), C.byref(probe)))
        if tier == 1 and self.walk is None and not torch.cuda.is_current_stream_capturing():
            graph.ensure_hub_index()
        self.sorted_front = tier == 0 and not self.return_graphs and os.environ.get('S3GRL_FRONT_ORDER', '1') != '0'
        if self.sorted_front and not torch.cuda.is_current_stream_capturing():
            graph.ensure_size_proxy()
        can_pair = (self.flow == L.FLOW_POS and self.fixed_rows and not self.return_graphs and self.walk is None
                    and tier == 0)
        self.pair = bool(pair) and can_pair and mirror is None
        # PoS Plus: (u,v) and (v,u) share everything but the order of rows 0 and 1 (csrc/expand.cu): the path runs on
        # one link per unordered node pair and s3_scatter_rows places every link's rows.  Not with a parity dump.
        self.pair_var = (bool(pair) and self.flow == L.FLOW_POS and not self.fixed_rows and not self.return_graphs
                         and mirror is None and out_link is None and peers is None)
        if (mirror is not None or out_link is not None or peers is not None) and not self.fixed_rows:
            raise NotImplementedError("out_link / mirror / peers serve the fixed-row flows only")
        if mirror is not None and not can_pair:
            mirror = None       # every rank of an exchange takes the same decision: all links are computed directly
        self.mirror = None if mirror is None else mirror.to(device=self.dev, dtype=torch.int64).contiguous()
        self.out_link = None if out_link is None else out_link.to(device=self.dev, dtype=torch.int64).contiguous()
        if self.out_link is not None and self.out_link.numel() != self.num_links:
            raise ValueError("out_link must hold one global link index per link")
        self.peers = peers
        self.total_links = self.num_links      # links of the whole list the output is laid out for
        if self.mirror is not None:
            self.total_links = int(self.mirror.numel())
        elif peers is not None:
            self.total_links = int(peers.num_links)
        if self.out_link is None and self.total_links != self.num_links:
            raise ValueError("a call over a subset of the link list needs out_link")
        if peers is not None and (host_out is not None or out is not None):
            raise ValueError("peers owns the output buffers: out / host_out cannot be combined with it")
        self.stats = dict(records=self.num_links * self.rpl, links=self.num_links, sum_n=0, sum_d=0, max_n=0, rows=0,
                          retries=0, batches=self.num_batches, launches=0, sum_n_links=0, sum_d_links=0, mirrors=0)
        # inputs above were produced on the current stream; self.stream may be another one
        self._prep_done = torch.cuda.Event()
        self._prep_done.record(torch.cuda.current_stream(self.dev))
        self.pieces, self.row_counts = [], []
        self.graphs = [] if self.return_graphs else None
        self.counters = None

    # ------------------------------------------------------------------ helpers
    def set_batches(self):
        """Batch boundaries (links).  With a pipelined device->host copy the first batches are small (1/8, 1/4, 1/2 of a
        batch): the copy engine starts after ~1 ms of compute instead of a whole batch, and PCIe is the bottleneck
        of that path.  Results do not depend on the batch composition."""
        starts, pos = [0], 0
        ramp = [self.batch_links // 8, self.batch_links // 4, self.batch_links // 2] if self.host_out is not None else []
        for step in ramp:
            if step >= 512 and pos + step < self.num_links:
                pos += step
                starts.append(pos)
        rest = self.num_links - pos
        nb = max(1, int(round(rest / self.batch_links)))     # even batches: no tiny tail batch (164 000 = 5 x 32 800)
        size = (rest + nb - 1) // nb
        while pos + size < self.num_links:
            pos += size
            starts.append(pos)
        self.starts = starts + [self.num_links]
        self.num_batches = len(starts) if self.num_links else 0

    def bounds(self, bi):
        return self.starts[bi], self.starts[bi + 1]

    def make_batch(self, b0, b1, arena, off, cnt, ctr, row_ptr=None, item_ptr=None, item_rec=None, order=None):
        src = self.links[0, b0:b1] if b1 > b0 else None
        dst = self.links[1, b0:b1] if b1 > b0 else None
        w = self.walk
        ol = getattr(self, 'out_link', None)
        return L.Batch(_ptr(src), _ptr(dst), b1 - b0, self.flow, self.strategy, self.num_hops, self.K, self.flags, 0,
                       _ptr(arena), arena.numel() if arena is not None else 0, _ptr(off), _ptr(cnt), _ptr(ctr),
                       _ptr(row_ptr), _ptr(item_ptr), _ptr(item_rec), _ptr(order),
                       _ptr(w['sets']) if w else None, _ptr(w['counts']) if w else None,
                       _ptr(w['src'][b0:]) if w else None, _ptr(w['dst'][b0:]) if w else None, w['cap'] if w else 0, 0,
                       _ptr(ol[b0:]) if ol is not None and b1 > b0 else None, _ptr(getattr(self, 'mirror', None)), b0,
                       self.cap_ratio, self.cap_max, self.cap_seed,
                       _ptr(getattr(self, '_front_order', None)) if b1 > b0 else None)

    def launch(self, stage, bi, fn_name, *args, on=None):
        """Call one C entry point; with `profile`, bracket it with CUDA events on its stream.  With S3GRL_NVTX=1
        in the environment every launch sits in an NVTX range "s3:<stage>" (ranges around kernels 1-3
        for nsys / ncu --nvtx filtering)."""
        fn = getattr(self.lib, fn_name)
        if _NVTX:
            torch.cuda.nvtx.range_push(f"s3:{stage}:{bi}")
            try:
                return self._launch(stage, bi, fn, fn_name, args, on)
            finally:
                torch.cuda.nvtx.range_pop()
        return self._launch(stage, bi, fn, fn_name, args, on)

    def _launch(self, stage, bi, fn, fn_name, args, on):
        if self.profile is None:
            return L.check(fn(*args), fn_name)
        on = self.stream if on is None else on
        # inside a CUDA-graph capture the events become event-record nodes (external=True): after a replay they hold
        # that replay's timestamps, so the per-kernel times of a graph-replayed step can still be read back
        ext = torch.cuda.is_current_stream_capturing()
        e0, e1 = torch.cuda.Event(enable_timing=True, external=ext), torch.cuda.Event(enable_timing=True, external=ext)
        e0.record(on)
        L.check(fn(*args), fn_name)
        e1.record(on)
        self.profile.append((stage, bi, e0, e1))

    def hop_nodes(self):
        """Nodes per hop summed over all records of the call (needs profile=...); call after finalize()."""
        tot = [0] * (L.MAX_HOPS + 1)
        for cnt in getattr(self, '_hop_cnts', []):
            hops = cnt[:, L.CNT_HOP0:L.CNT_HOP0 + L.MAX_HOPS + 1].sum(0, dtype=torch.int64).cpu()
            tot = [a + int(b) for a, b in zip(tot, hops)]
        return tot

    def grow(self):
        self.stats['retries'] += 1
        self.words = int(self.words * 2)
        self.graph._arena = self.graph._arena2 = None       # hand the old arenas back to the driver first
        torch.cuda.empty_cache()
        free, _ = torch.cuda.mem_get_info(self.dev)
        if self.words * 4 > free:
            raise MemoryError("scratch arena does not fit in device memory; lower batch_records")

    def account(self, bi, cnt, host_counters=None):
        """True if batch `bi` finished cleanly (and add its counters to the stats); False if its arena
        overflowed; raises ValueError for invalid links."""
        c = self.counters[bi].cpu() if host_counters is None else host_counters[bi]
        if int(c[L.CTR_ERRORS]) != 0:
            status = cnt[:, L.CNT_STATUS]
            if bool((status == L.REC_BAD_LINK).any()):
                bad = int(torch.nonzero(status == L.REC_BAD_LINK)[0]) // self.rpl + self.starts[bi]
                raise ValueError(f"invalid target link at position {bad}: node id out of range or src == dst")
            return False
        self.stats['sum_n'] += int(c[L.CTR_SUM_N])
        self.stats['sum_d'] += int(c[L.CTR_SUM_D])
        self.stats['sum_n_links'] += int(c[L.CTR_SUM_N_ALL])
        self.stats['sum_d_links'] += int(c[L.CTR_SUM_D_ALL])
        self.stats['mirrors'] += int(c[L.CTR_MIRRORS])
        self.stats['sum_read'] = self.stats.get('sum_read', 0) + int(c[L.CTR_SUM_READ])
        for key, slot in (('chain_reads', L.CTR_CHAIN_READS), ('chain_records', L.CTR_CHAIN_RECORDS), ('chain_n', L.CTR_CHAIN_N)):
            self.stats[key] = self.stats.get(key, 0) + int(c[slot])
        self.stats['max_n'] = max(self.stats['max_n'], int(c[L.CTR_MAX_N]))
        return True

    def dump(self, cnt, off, arena):
        """Parity dump of one batch (S3_BATCH_STORE_ALL_ROWS): canonical nodes, hops, compact local CSR."""
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
            sel = np.concatenate([np.arange(self.nseed),
                                  ar[of[r, L.OFF_SEL]:of[r, L.OFF_SEL] + s - self.nseed]]).astype(np.int32)
            hop_cnt = cn[r, L.CNT_HOP0:L.CNT_HOP0 + L.MAX_HOPS + 1]
            hops = np.repeat(np.arange(L.MAX_HOPS + 1), hop_cnt).astype(np.int32)
            self.graphs.append(dict(nodes=nodes, hops=hops, lrowptr=rowptr, lcol=lcol, sel=sel,
                                    partner=int(cn[r, L.CNT_PARTNER])))

    def meta(self, nrec):
        off = torch.empty((nrec, L.NOFF), dtype=torch.int64, device=self.dev)
        cnt = torch.empty((nrec, L.NCNT), dtype=torch.int32, device=self.dev)
        both = torch.empty(2 * nrec, dtype=torch.int32, device=self.dev)     # one allocation: same lifetime
        order = both[:nrec]
        self._front_order = both[nrec:] if getattr(self, 'sorted_front', False) else None
        return off, cnt, order

    # ------------------------------------------------------------------ fixed-row flows
    def pair_links(self):
        """s3_pair_links over the call's link list -> self.mirror (chain table, see include/s3grl_b200.h)."""
        self.mirror, self._pair_table = pair_links(self.links, self.graph.num_nodes, self.stream)
        self.stats['launches'] += 2

    def gather_fixed(self, bi, batch, nrec, row_base, st, on=None):
        """Kernel 3 of a fixed-row batch: s3_gather into this GPU's matrices, or s3_gather_peers into every GPU's."""
        g = C.byref(self.graph._c)
        if self.peers is None:
            return self.launch('gather', bi, 's3_gather', g, C.byref(batch), nrec, self.out_ptrs, self.F1, row_base, st, on=on)
        pb = self.peers
        return self.launch('gather', bi, 's3_gather_peers', g, C.byref(batch), nrec, pb.base_array, pb.world_dst, pb.op_stride,
                           pb.ld, pb.flags, st, on=on)

    def enqueue_fixed_batch(self, bi, arena):
        g, st = C.byref(self.graph._c), self.stream_ptr
        b0, b1 = self.bounds(bi)
        nrec = (b1 - b0) * self.rpl
        off, cnt, order = self.meta(nrec)
        self.counters[bi].zero_()
        batch = self.make_batch(b0, b1, arena, off, cnt, self.counters[bi], order=order)
        row_base = b0 * self.rpl * self.nseed
        self.launch('extract', bi, 's3_extract', g, C.byref(batch), st)
        self.gather_fixed(bi, batch, nrec, row_base, st)
        self.stats['launches'] += 3         # front kernel, order kernel, gather kernel
        if self.host_out is not None:      # pipelined D2H of this batch's rows
            r0, r1 = row_base, b1 * self.rpl * self.nseed
            done = torch.cuda.Event()
            done.record(self.stream)
            self.copy_stream.wait_event(done)
            with torch.cuda.stream(self.copy_stream):
                for k in range(self.K + 1):
                    self.host_out[k][r0:r1].copy_(self.out[k][r0:r1], non_blocking=True)
        return off, cnt

    def enqueue_fixed_overlapped(self, todo):
        """Two streams, two arenas: the front kernel of batch i+1 overlaps the gather of batch i."""
        g = C.byref(self.graph._c)
        sF, sB = self.graph.streams()
        arenas = (self.graph.arena(self.words, 0), self.graph.arena(self.words, 1))
        for bi in todo:                      # per-batch memsets: no host synchronisation
            self.counters[bi].zero_()
        start = torch.cuda.Event()
        start.record(self.stream)
        sF.wait_event(start)
        sB.wait_event(start)
        pF, pB = C.c_void_p(sF.cuda_stream), C.c_void_p(sB.cuda_stream)
        metas, keep, back_done = [], [], []
        for idx, bi in enumerate(todo):
            b0, b1 = self.bounds(bi)
            nrec = (b1 - b0) * self.rpl
            off, cnt, order = self.meta(nrec)
            keep.append((off, cnt, order))          # allocated on self.stream, used on sF / sB: keep alive
            batch = self.make_batch(b0, b1, arenas[idx % 2], off, cnt, self.counters[bi], order=order)
            if idx >= 2:
                sF.wait_event(back_done[idx - 2])   # this arena is free again
            self.launch('extract', bi, 's3_extract', g, C.byref(batch), pF, on=sF)
            front_done = torch.cuda.Event()
            front_done.record(sF)
            sB.wait_event(front_done)
            self.gather_fixed(bi, batch, nrec, b0 * self.rpl * self.nseed, pB, on=sB)
            done = torch.cuda.Event()
            done.record(sB)
            back_done.append(done)
            self.stats['launches'] += 3
            metas.append((bi, cnt))
        if back_done:
            self.stream.wait_event(back_done[-1])
        return metas, keep

    def enqueue_fixed_fronts_first(self, todo):
        """Both front kernels, then both gathers (two arenas, one stream).  Used by the odd ranks of an 8-GPU exchange
        while the even ranks run front / gather / front / gather: the two halves of the node then store their rows
        over NVLink at different times, and a GPU's NVLink ingress (the limiter there) is not asked for all eight
        ranks' rows at once."""
        g, st = C.byref(self.graph._c), self.stream_ptr
        arenas = (self.graph.arena(self.words, 0), self.graph.arena(self.words, 1))
        metas, staged = [], []
        for idx, bi in enumerate(todo):
            b0, b1 = self.bounds(bi)
            nrec = (b1 - b0) * self.rpl
            off, cnt, order = self.meta(nrec)
            self.counters[bi].zero_()
            batch = self.make_batch(b0, b1, arenas[idx], off, cnt, self.counters[bi], order=order)
            self.launch('extract', bi, 's3_extract', g, C.byref(batch), st)
            staged.append((bi, batch, nrec, b0 * self.rpl * self.nseed))
            metas.append((bi, cnt))
        for bi, batch, nrec, row_base in staged:
            self.gather_fixed(bi, batch, nrec, row_base, st)
        self.stats['launches'] += 3 * len(todo)
        return metas, staged

    def enqueue_fixed(self, todo):
        t0 = time.perf_counter()
        if self.overlap:
            metas, keep = self.enqueue_fixed_overlapped(todo)
        elif self.fronts_first and len(todo) == 2 and self.host_out is None and not self.return_graphs:
            metas, keep = self.enqueue_fixed_fronts_first(todo)
        else:
            metas, keep = [], None
            arena = self.graph.arena(self.words)
            for bi in todo:
                off, cnt = self.enqueue_fixed_batch(bi, arena)
                if self.return_graphs:       # the arena is recycled by the next batch: dump now
                    self.stream.synchronize()
                    if self.account(bi, cnt):
                        self.dump(cnt, off, arena)
                        continue
                metas.append((bi, cnt))
        self.stats['host_enqueue_ms'] = 1000 * (time.perf_counter() - t0)
        return metas, keep

    def sync(self):
        self.stream.synchronize()
        if self.copy_stream is not None:
            self.copy_stream.synchronize()

    def finalize_fixed(self, metas):
        """Stream sync, validation, re-run of overflowed batches with a larger arena."""
        with torch.cuda.device(self.dev), torch.cuda.stream(self.stream):
            while True:
                self.sync()
                host_counters = self.counters.cpu()       # one D2H for every batch's counters
                todo = [bi for bi, cnt in metas if not self.account(bi, cnt, host_counters)]
                if self.profile is not None:       # per-hop node counts (FMA count of kernel 3) are summed lazily: hop_nodes()
                    self._hop_cnts = getattr(self, '_hop_cnts', []) + [cnt for bi, cnt in metas if bi not in todo]
                if not todo:
                    return
                if self.return_graphs and self.graphs:
                    raise RuntimeError("arena overflow while dumping graphs: pass a larger arena_words")
                self.grow()
                metas, _keep = self.enqueue_fixed(todo)

    # ------------------------------------------------------------------ PoS Plus (data-dependent rows)
    def run_variable_batch(self, bi, arena):
        """front -> plan -> host sync (rows, CCN items) -> gather [-> plan_items -> diffuse -> gather_ccn].
        Returns (off, cnt, ok); not ok means the arena overflowed."""
        g, st = C.byref(self.graph._c), self.stream_ptr
        b0, b1 = self.bounds(bi)
        nrec = (b1 - b0) * self.rpl
        off, cnt, order = self.meta(nrec)
        row_ptr = torch.empty(nrec + 1, dtype=torch.int64, device=self.dev)
        item_ptr = torch.empty(nrec + 1, dtype=torch.int64, device=self.dev)
        ctr = self.counters[bi]
        ctr.zero_()
        batch = self.make_batch(b0, b1, arena, off, cnt, ctr, row_ptr, item_ptr, order=order)
        self.launch('extract', bi, 's3_extract', g, C.byref(batch), st)
        L.check(self.lib.s3_plan(C.byref(batch), st), 's3_plan')
        self.stats['launches'] += 3         # front kernel, order kernel, plan scan
        c = ctr.cpu()                                  # sync: rows / items / errors of this batch
        if int(c[L.CTR_ERRORS]) != 0:
            return off, cnt, False
        rows, items = int(c[L.CTR_ROWS]), int(c[L.CTR_ITEMS])
        item_rec = torch.empty(max(items, 1), dtype=torch.int32, device=self.dev)
        xs = [torch.empty((rows, self.F1), dtype=torch.float32, device=self.dev) for _ in range(self.K + 1)]
        ptrs = (C.c_void_p * (self.K + 1))(*[o.data_ptr() for o in xs])
        batch = self.make_batch(b0, b1, arena, off, cnt, ctr, row_ptr, item_ptr, item_rec, order)
        self.launch('gather', bi, 's3_gather', g, C.byref(batch), nrec, ptrs, self.F1, 0, st)
        self.stats['launches'] += 1
        if self.ccn_chain and rows > 2 * nrec:
            # union: hop-limited SpMM chain over the stored CSR, one CTA per record; records too large for its
            # shared-memory placement were counted as work items by s3_plan and take the path below
            if _CHAIN_POOL_CW:
                pool, busy, slots = self.graph.chain_pool()
                self.launch('ccn_chain', bi, 's3_ccn_chain_pooled', g, C.byref(batch), nrec, ptrs, self.F1, 0, _ptr(pool), _ptr(busy),
                            DeviceGraph.CHAIN_SLOT_FLOATS, slots, _CHAIN_POOL_CW, st)
            else:
                self.launch('ccn_chain', bi, 's3_ccn_chain', g, C.byref(batch), nrec, ptrs, self.F1, 0, st)
            # kernels of one call: prep, the global-memory class (+ its launches by CSR size with a pool), three CTA sizes
            split = min(2, max(0, int(os.environ.get('S3GRL_CHAIN_POOL_SPLIT', '2') or 0))) if _CHAIN_POOL_CW else 0
            self.stats['launches'] += 5 + split
        if items:       # CCN rows: extra work items of 2 (intersection) or 8 (union) selected rows each
            L.check(self.lib.s3_plan_items(C.byref(batch), st), 's3_plan_items')
            self.launch('diffuse', bi, 's3_diffuse', g, C.byref(batch), items, st)
            self.launch('gather_ccn', bi, 's3_gather_ccn', g, C.byref(batch), items, ptrs, self.F1, 0, st)
            self.stats['launches'] += 3
        self.pieces.append(xs)
        self.row_counts.append(row_ptr[1:] - row_ptr[:-1])
        if getattr(self, '_rec_n', None) is not None:        # pairing: per-link accounting needs every record's n
            self._rec_n.append(cnt[:, L.CNT_N].to(torch.int64))
        return off, cnt, True

    def run_variable(self):
        for bi in range(self.num_batches):
            while True:
                arena = self.graph.arena(self.words)
                off, cnt, ok = self.run_variable_batch(bi, arena)
                self.stream.synchronize()
                if ok and self.account(bi, cnt):
                    if self.return_graphs:
                        self.dump(cnt, off, arena)
                    break
                self.account(bi, cnt)      # raises on bad links
                self.grow()

    # ------------------------------------------------------------------ driver
    def run(self, defer):
        dev, K, F1, Lk = self.dev, self.K, self.F1, self.num_links
        with torch.cuda.device(dev), torch.cuda.stream(self.stream):
            self.stream.wait_event(self._prep_done)
            self.counters = torch.zeros((max(self.num_batches, 1), L.NCTR), dtype=torch.int64, device=dev)
            if self.host_out is not None:
                self.copy_stream = self.graph.streams()[1]
                ready = torch.cuda.Event()
                ready.record(self.stream)
                self.copy_stream.wait_event(ready)
            if self.fixed_rows:
                R = 2 * self.total_links
                if self.pair and Lk > 1:
                    self.pair_links()
                if self.peers is not None:
                    if not (self.peers.rows >= R and self.peers.cols == F1 and self.peers.num_ops == K + 1):
                        raise ValueError("peer buffers do not match this call: need K+1 operators of [>= 2L, F+1]")
                    self.out = self.peers.local
                elif self.out is None:
                    self.out = [torch.empty((R, F1), dtype=torch.float32, device=dev) for _ in range(K + 1)]
                elif not (len(self.out) == K + 1 and all(o.shape[0] >= R and o.shape[1] == F1 and o.is_contiguous()
                                                         for o in self.out)):
                    raise ValueError("out must be K+1 contiguous [>= 2L, F+1] float32 tensors")
                self.out_ptrs = (C.c_void_p * (K + 1))(*[o.data_ptr() for o in self.out])
                metas, keep = self.enqueue_fixed(list(range(self.num_batches)))
                xs = [o[:R] for o in self.out]
                row_ptr = torch.arange(self.total_links + 1, dtype=torch.int64, device=dev) * 2
                self.stats['rows'] = R
                result = PrecomputeResult(xs, row_ptr, self.stats, self.graphs)
                result._keep = keep
                result._finalize = lambda: self.finalize_fixed(metas)
                result.hop_nodes = self.hop_nodes
                return result if defer else result.finalize()
            head_idx, head_code, var_mirror = None, None, None
            if self.pair_var and Lk > 1:
                # chain table of the whole list; the path then runs on the chain heads only
                var_mirror, self._pair_table = pair_links(self.links, self.graph.num_nodes, self.stream)
                is_head = var_mirror >= -1
                head_idx = torch.nonzero(is_head).flatten()
                self.stats['launches'] += 2
                if int(head_idx.numel()) < Lk:
                    head_code = torch.empty(Lk, dtype=torch.int64, device=dev)
                    L.check(self.lib.s3_pair_heads(_ptr(var_mirror), Lk, _ptr(head_code), self.stream_ptr), 's3_pair_heads')
                    self.stats['launches'] += 1
                    self.stats['mirrors'] = Lk - int(head_idx.numel())
                    self.links = self.links[:, head_idx].contiguous()
                    self.num_links = int(head_idx.numel())
                    self.stats['records'] = self.num_links * self.rpl
                    self.set_batches()
                    self.counters = torch.zeros((max(self.num_batches, 1), L.NCTR), dtype=torch.int64, device=dev)
                    self._rec_n = []
                else:
                    head_idx, var_mirror = None, None
            self.run_variable()
            counts = torch.cat(self.row_counts) if self.row_counts else torch.zeros(0, dtype=torch.int64, device=dev)
            if head_idx is not None:      # rows of every link of the list = rows of its head's record
                rank = torch.cumsum(is_head, 0) - 1
                counts = counts[rank[head_code >> 1]]
                # SURVEY 8d sums over LINKS: a paired link counts its own subgraph although its head's record serves it
                # (n exactly; D scaled by the same ratio — the per-record D is not kept)
                served = torch.bincount(head_code >> 1, minlength=Lk)[head_idx]
                n_links = int((torch.cat(self._rec_n) * served).sum()) if self._rec_n else 0
                self.stats['sum_d_links'] = int(self.stats['sum_d'] * n_links / max(1, self.stats['sum_n']))
                self.stats['sum_n_links'] = n_links
            # compat_explicit_zero: the reference's literal union selects src and dst a second time; the
            # collate below writes rows 0 / 1 of every record twice — [0, 1, 0, 1, CCN rows] — nothing is recomputed
            lead = 2 if self.compat_explicit_zero else 0
            if lead:
                counts = counts + lead
            row_ptr = torch.zeros(Lk + 1, dtype=torch.int64, device=dev)
            torch.cumsum(counts, 0, out=row_ptr[1:])
            self.stats['rows_computed'] = sum(int(p[0].shape[0]) for p in self.pieces)   # without the paired links' copies
            if not self.pieces:
                xs = [torch.empty((0, F1), dtype=torch.float32, device=dev) for _ in range(K + 1)]
            elif len(self.pieces) == 1 and head_idx is None and not lead:
                xs = self.pieces[0]
            else:
                # collate: every batch's piece placed (and, with pairing, replicated) by s3_scatter_rows
                R = int(row_ptr[-1])
                xs = [torch.empty((R, F1), dtype=torch.float32, device=dev) for _ in range(K + 1)]
                dst = (C.c_void_p * (K + 1))(*[o.data_ptr() for o in xs])
                for bi, piece in enumerate(self.pieces):
                    b0, b1 = self.bounds(bi)
                    prp = torch.zeros(b1 - b0 + 1, dtype=torch.int64, device=dev)
                    torch.cumsum(self.row_counts[bi], 0, out=prp[1:])
                    src = (C.c_void_p * (K + 1))(*[o.data_ptr() for o in piece])
                    self.launch('collate', bi, 's3_scatter_rows_lead', src, F1, _ptr(prp), b1 - b0,
                                _ptr(head_idx[b0:b1]) if head_idx is not None else None, b0, _ptr(var_mirror), _ptr(row_ptr),
                                dst, F1, K + 1, F1, lead, self.stream_ptr)
                    self.stats['launches'] += 1
                    self._keep_var = getattr(self, '_keep_var', []) + [prp]
                self.pieces = []
            self.stats['rows'] = int(xs[0].shape[0])
            self.stats['links'] = Lk
            return PrecomputeResult(xs, row_ptr, self.stats, self.graphs)


def precompute(graph, links, num_hops, sign_k, flow='PoS', strategy=None, batch_records=None, out=None,
               return_graphs=False, arena_words=None, stream=None, profile=None, overlap=False, defer=False,
               host_out=None, force_sorted_tier=False, walk=None, ccn_mode=None, pair=True, out_link=None, mirror=None,
               peers=None, ratio_per_hop=1.0, max_nodes_per_hop=None, cap_seed=0, fronts_first=False,
               compat_explicit_zero=False):
    """Run the hot path for `links` ([2, L] int64, host or device) on `graph`; returns a
    PrecomputeResult with device tensors.

    batch_records  records per batch (default: 32768 for fixed-row flows, as many as memory allows for
                   PoS Plus intersection, 32768 for union (8192 with ccn_mode='items')).
    out            K+1 preallocated [>= 2L, F+1] float32 device tensors (fixed-row flows only).
    profile        a list that receives (stage, batch, start_event, end_event) for every kernel launch,
                   so the caller can time each kernel on its launching stream with CUDA events.
    overlap        (fixed-row flows) front kernel of batch i+1 on one stream while the gather of batch i
                   runs on another, with two arenas.  Measured gain on PubMed: ~1.5 %; off by default.
    host_out       (fixed-row flows) K+1 pinned host tensors [>= 2L, F+1]: every batch's rows are copied
                   device->host on a side stream as soon as its gather finishes, so the D2H of batch i
                   overlaps the kernels of batch i+1 (the reference returns CPU tensors).
    defer          (fixed-row flows) return right after enqueueing; the caller must call
                   `result.finalize()` (stream sync + validation + re-run of overflowed batches) before
                   using the outputs.  Lets several calls be queued back to back.
    return_graphs  also return every record's canonical nodes / hops / local CSR (parity tests).
    walk           ScaLed subgraphs (reference utils.py:86-150): dict(m=, M=, seed=) to sample on the GPU,
                   dict(cache={node: tensor}) for a reference-style walk cache, or dict(sets=, counts=) for
                   a table indexed by node id.  The subgraph of (u, v) is {u, v} ∪ set(u) ∪ set(v).
    ccn_mode       PoS Plus union only: None / 'chain' (default) = s3_ccn_chain, a hop-limited SpMM chain per record
                   (12x fewer additions than the weight formulation, shared-memory-bandwidth bound) for the records that
                   fit its shared-memory placement (n up to ~4 000) and work items for the rest; 'items' = s3_diffuse +
                   s3_gather_ccn work items of 8 selected rows for every record (round 1's route; also what
                   return_graphs=True uses, because the chain converts the stored CSR in place).
    pair           (PoS without CCN rows, bitmap tier) link pairing: links over the same unordered node pair — both
                   directions of a training edge, SURVEY.md A.7 — share one record; the other direction's rows are
                   the same rows exchanged, bit for bit (csrc/pair.cu).  On by default; results do not depend on it.
    ratio_per_hop, max_nodes_per_hop, cap_seed   the reference's per-hop caps (utils.py:66-70) for the PoS flows: a hop
                   with c new nodes keeps min(int(ratio_per_hop * c), max_nodes_per_hop) of them.  The reference picks
                   them with random.sample (no reproducible semantics, raises on Python >= 3.11); here they are the
                   nodes with the smallest fmix32(node ^ cap_seed), so a subgraph depends on (link, seed) only and
                   the oracle restates it exactly.  Ignored by SoP and by ScaLed walk subgraphs, as in the reference.
    compat_explicit_zero   PoS Plus union only.  False (default, what is benchmarked): the paper's selection
                   [0, 1] + (N(0) ∪ N(1)) − {0, 1}.  True: the code-literal selection of the reference with its ragged
                   label-column literal (tuned_SIGN.py:243) repaired — its target-link mask leaves explicit zeros that
                   `neighbors` reports, so src and dst are selected a second time: every link gets the rows
                   [0, 1, 0, 1, CCN rows] (the reference's own order of the rows beyond the first two is CPython set order).
                   Pinned by tests/golden/ref_*_union*.npz, which hold the literal rows.
    out_link, mirror, peers   used by parallel.precompute_exchange: `links` is a subset of a larger list, out_link
                   its global link indices, mirror the chain table of the whole list and peers the PeerBuffers every
                   output row is stored into (this GPU's and its NVLink peers').
    Raises ValueError for invalid links (out of range, src == dst), NotImplementedError for an unknown
    strategy (as reference tuned_SIGN.py:235) or an unsupported combination."""
    call = _Call(graph, links, num_hops, sign_k, flow, strategy, batch_records, out, return_graphs, arena_words,
                 stream, profile, overlap, host_out, force_sorted_tier, walk, ccn_mode, pair, out_link, mirror, peers,
                 ratio_per_hop, max_nodes_per_hop, cap_seed, fronts_first, compat_explicit_zero)
    return call.run(defer)


_LABEL = {'zo': L.LABEL_ZO, 'hop': L.LABEL_HOP, 'drnl': L.LABEL_DRNL, 'degree': L.LABEL_DEGREE}


def precompute_full(graph, links, num_hops, sign_k, node_label='drnl', batch_records=None, arena_words=None,
                    stream=None, profile=None, force_sorted_tier=False, walk=None, ratio_per_hop=1.0, max_nodes_per_hop=None,
                    cap_seed=0):
    """The NON-optimised SIGN + SEAL flow (reference utils.py:497-520 with `powers_of_A` empty:
    k_hop_subgraph -> construct_pyg_graph(node_label) -> TunedSIGN(sign_k), i.e. PyG's SIGN on the
    whole enclosing subgraph): every subgraph node is an output row, x = [z | X_sub], x_k = S x_{k-1}.

    Returns a PrecomputeResult whose xs[k] are [sum_i n_i, F+1], row_ptr[i]..row_ptr[i+1] the rows of
    link i in canonical node order (src, dst, then ascending (hop, global id)), plus `.node_id`
    (int64 [sum n], the reference's data.node_id).  Rows 0 and 1 of every link equal the optimised
    PoS flow's output when node_label == 'zo' (SURVEY.md §8a row 9).

    node_label: 'zo' | 'hop' | 'drnl' | 'degree'; any other string gives a zero column, as the
    reference's else-branch does (utils.py:309-310); 'de' / 'de+' raise (two label columns: the
    reference's own reshape at utils.py:314 fails for them).
    Two passes: sizes first (extraction only), then one exact allocation and the SpMM chain."""
    if node_label in ('de', 'de+'):
        raise NotImplementedError(f"node_label {node_label!r} yields two columns; the reference's SIGN flow cannot "
                                  "reshape it either (utils.py:314)")
    if node_label == 'degree' and graph.has_multi_edges:
        raise NotImplementedError("node_label 'degree' on a multigraph is not supported")
    label = _LABEL.get(node_label, L.LABEL_ZERO)
    call = _Call(graph, links, num_hops, sign_k, 'PoS', None, batch_records or 4096, None, False, arena_words,
                 stream, profile, False, None, force_sorted_tier, walk, ratio_per_hop=ratio_per_hop,
                 max_nodes_per_hop=max_nodes_per_hop, cap_seed=cap_seed)
    call.flags |= L.BATCH_STORE_ALL_ROWS
    dev, K, F1, lib = call.dev, call.K, call.F1, call.lib
    g, st = C.byref(graph._c), call.stream_ptr

    def front(bi):
        """extract + plan_full of batch bi (retried with a larger arena on overflow) -> (batch, n per record, rows)"""
        b0, b1 = call.bounds(bi)
        while True:
            arena = graph.arena(call.words)
            off, cnt, order = call.meta(b1 - b0)
            row_ptr = torch.empty(b1 - b0 + 1, dtype=torch.int64, device=dev)
            ctr = call.counters[bi]
            ctr.zero_()
            batch = call.make_batch(b0, b1, arena, off, cnt, ctr, row_ptr=row_ptr, order=order)
            call.launch('extract', bi, 's3_extract', g, C.byref(batch), st)
            L.check(lib.s3_plan_full(C.byref(batch), st), 's3_plan_full')
            call.stats['launches'] += 3
            c = ctr.cpu()
            if int(c[L.CTR_ERRORS]) == 0:
                return batch, (off, cnt, order, row_ptr, arena), cnt[:, L.CNT_N].to(torch.int64), int(c[L.CTR_ROWS]), c
            call.account(bi, cnt, {bi: c})      # raises on bad links
            call.grow()

    with torch.cuda.device(dev), torch.cuda.stream(call.stream):
        call.counters = torch.zeros((max(call.num_batches, 1), L.NCTR), dtype=torch.int64, device=dev)
        sizes, rows_per_batch, last = [], [], None
        for bi in range(call.num_batches):
            last = front(bi)
            sizes.append(last[2])
            rows_per_batch.append(last[3])
        R = sum(rows_per_batch)
        need = R * F1 * 4 * (K + 1) + R * 8
        free, _ = torch.cuda.mem_get_info(dev)
        if need > free:
            raise MemoryError(f"the non-optimised flow needs {need / 2**30:.1f} GiB for {R} rows x {K + 1} operators "
                              f"({free / 2**30:.1f} GiB free): pass fewer links per call")
        xs = [torch.empty((R, F1), dtype=torch.float32, device=dev) for _ in range(K + 1)]
        node_id = torch.empty(R, dtype=torch.int64, device=dev)
        ptrs = (C.c_void_p * (K + 1))(*[o.data_ptr() for o in xs])
        row_base = 0
        for bi in range(call.num_batches):
            cur = last if call.num_batches == 1 else front(bi)
            batch, _keep, _, rows, c = cur
            call.launch('sign_full', bi, 's3_sign_full', g, C.byref(batch), int(batch.num_links), label, ptrs, F1,
                        row_base, _ptr(node_id), st)
            call.stats['launches'] += 1
            call.stats['sum_n'] += int(c[L.CTR_SUM_N])
            call.stats['sum_d'] += int(c[L.CTR_SUM_D])
            call.stats['max_n'] = max(call.stats['max_n'], int(c[L.CTR_MAX_N]))
            row_base += rows
            if call.num_batches > 1:
                call.stream.synchronize()     # the arena is recycled by the next batch
        counts = torch.cat(sizes) if sizes else torch.zeros(0, dtype=torch.int64, device=dev)
        row_ptr = torch.zeros(call.num_links + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=row_ptr[1:])
        call.stream.synchronize()
    call.stats['rows'] = R
    res = PrecomputeResult(xs, row_ptr, call.stats)
    res.node_id = node_id
    return res


def pool_rows(result, mode='sum'):
    """Pooled output mode (SURVEY.md §7 'pooling semantics'): every link keeps three rows per operator —
    src, dst and the sum / mean of its CCN rows — instead of 2 + #CCN.  Runs s3_segment_pool over each operator
    matrix of a PoS Plus result and returns a PrecomputeResult with xs[k] [3L, F+1], row_ptr = 3 * arange.
    NOTE: this pools operator rows BEFORE the model's MLP, which is a different model than the reference's
    (it pools hidden rows, models.py:347-362); the parity check is "reduce the oracle's rows"."""
    from .head import segment_pool
    nl = int(result.row_ptr.numel()) - 1
    xs = [segment_pool(x, result.row_ptr, mode, 'rows').view(3 * nl, x.shape[1]) for x in result.xs]
    row_ptr = torch.arange(nl + 1, dtype=torch.int64, device=result.row_ptr.device) * 3
    stats = dict(result.stats, rows=3 * nl, pooled=mode)
    return PrecomputeResult(xs, row_ptr, stats, result.graphs)


def algorithmic_bytes(stats, num_feat, sign_k, per_link=True):
    """SURVEY.md §8d:  sum over links of 4·D + 8·n + 4·F·n + 4·s·(K+1)·(F+1).
    per_link=True sums n and D over every LINK served (the survey's definition); per_link=False over the records
    actually extracted — smaller when link pairing served (v,u) from the record of (u,v)."""
    n = stats['sum_n_links'] if per_link and stats.get('sum_n_links') else stats['sum_n']
    d = stats['sum_d_links'] if per_link and stats.get('sum_d_links') else stats['sum_d']
    return 4 * d + 8 * n + 4 * num_feat * n + 4 * stats['rows'] * (sign_k + 1) * (num_feat + 1)
