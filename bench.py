#!/usr/bin/env python
"""Benchmark of the precompute hot path (BASELINE.json metric: precomputed target links/s,
extract + diffuse + select, PubMed PoS r=3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

One "step" = one pass of the hot path over the whole link list of the workload (what the
three SEALDataset.process calls of a run precompute: train/valid/test positives+negatives,
reference sgrl_link_pred.py:193-204).  Prints ONE JSON line on rank 0.

* value      : links/s with CSR, X, links and the output buffers resident in HBM.
* e2e        : the same metric through the reference-facing call
               s3grl_b200.extract_enclosing_subgraphs(link_index, A, x, ...) with HOST inputs
               (SciPy CSR, CPU tensors) and HOST outputs: graph + link H2D and the D2H of all
               K+1 operator matrices are inside the timed region.
* roofline   : the kernel with the largest share of the step, plus every kernel of the path with its own
               bytes / time / fraction and the measured L2-read and FP32 ceilings (roofline.kernels).
* cpu_baseline / --impl reference : the oracle port (oracle/s3grl_oracle.py — the reference
               itself is Python and cannot travel to the GPU box) on all host cores.

Multi-GPU (torchrun, one rank per GPU): STRONG scaling of the same step — the workload's link list is
sharded cyclically over the ranks, the graph is replicated, and every rank stores its output rows straight
into every rank's operator matrices over NVLink peer memory (s3_gather_peers: the all-gather of SURVEY.md
§8e fused into kernel 3), so the timed region ends with the complete matrices on every GPU.  After the
timed region every rank checks its copy against a single-GPU precompute of the whole list, bit for bit.
The default run also measures the other BASELINE configs that are not the headline (PoS Plus union,
R-MAT) with few steps and reports them under "configs".
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "precomputed target links/sec (extract+diffuse+pool), PubMed PoS r=3"   # default workload; see config.workload
UNIT = "links/s"


def metric_name(workload):
    return METRIC if workload == 'pubmed_pos' else f"precomputed target links/sec (extract+diffuse+pool), {workload}"


# ----------------------------------------------------------------------------------------
# workloads (host side, seeded; no file outside the repo is read)
# ----------------------------------------------------------------------------------------
def build_workload(name):
    from s3grl_b200 import datasets as ds
    if name == 'pubmed_scaled':      # ScaLed random-walk subgraphs, configs/paper/scaled.json (m=3, M=20)
        edges, N, _ = ds.load_graph('pubmed')
        A, splits = ds.split_links(edges, N, seed=1)
        X = ds.synthetic_features(N, 500, 0.1, 0)
        links = ds.all_links(splits)
        return dict(A=A, X=X, links=links, num_hops=0, K=3, flow='PoS', strategy=None, walk=dict(m=3, M=20, seed=0),
                    desc=f"PubMed graph, synthetic X F=500, all {links.shape[1]} links, ScaLed random-walk subgraphs "
                         f"m=3 M=20 (walks sampled inside every step), PoS sign_k=3")
    if name in ('pubmed_pos', 'pubmed_posplus_union', 'pubmed_posplus_intersection'):
        edges, N, _ = ds.load_graph('pubmed')
        A, splits = ds.split_links(edges, N, seed=1)
        X = ds.synthetic_features(N, 500, 0.1, 0)
        links = ds.all_links(splits)
        strategy = {'pubmed_pos': None, 'pubmed_posplus_union': 'union',
                    'pubmed_posplus_intersection': 'intersection'}[name]
        desc = (f"PubMed graph (tests/golden/graphs/pubmed.npz, N={N}, train nnz={A.nnz}), synthetic X F=500 "
                f"~10% nnz row-normalised seed 0 (real ind.pubmed.allx absent from the reference checkout), 85/5/10 "
                f"split seed 1, all {links.shape[1]} links of the 3 splits, PoS"
                f"{' Plus ' + strategy if strategy else ''} num_hops=3 sign_k=3")
        return dict(A=A, X=X, links=links, num_hops=3, K=3, flow='PoS', strategy=strategy, desc=desc)
    if name == 'usair_sop':          # BASELINE config 2
        edges, N, _ = ds.load_graph('usair')
        A, splits = ds.split_links(edges, N, seed=1)
        X = ds.normalize_features(ds.degree_one_hot(A, 1024))
        links = ds.all_links(splits)
        return dict(A=A, X=X, links=links, num_hops=0, K=3, flow='SoP', strategy=None,
                    desc=f"USAir (N={N}), init_features=degree one-hot F=1025, all {links.shape[1]} links, SoP sign_k=3")
    if name in ('yeast_pos_k5', 'power_pos_k5', 'router_pos_k5'):      # BASELINE config 4
        gname = name.split('_')[0]
        edges, N, _ = ds.load_graph(gname)
        A, splits = ds.split_links(edges, N, seed=1)
        X = ds.synthetic_features(N, 256, 1.0, 0)
        links = ds.all_links(splits)
        return dict(A=A, X=X, links=links, num_hops=2, K=5, flow='PoS', strategy=None,
                    desc=f"{gname} SEAL graph (N={N}), synthetic X F=256 row-normalised seed 0, all {links.shape[1]} links, "
                         f"PoS num_hops=2 sign_k=5")
    if name == 'cora_pos':
        edges, N, X = ds.load_graph('cora')
        A, splits = ds.split_links(edges, N, seed=1)
        X = ds.normalize_features(X)
        links = ds.all_links(splits)
        return dict(A=A, X=X, links=links, num_hops=3, K=3, flow='PoS', strategy=None,
                    desc=f"Cora (real X F=1433), all {links.shape[1]} links, PoS num_hops=3 sign_k=3")
    if name == 'cora_full_drnl':     # the non-optimised SIGN + SEAL flow (optimize_sign=False)
        edges, N, X = ds.load_graph('cora')
        A, splits = ds.split_links(edges, N, seed=1)
        X = ds.normalize_features(X)
        links = ds.all_links(splits)
        links = links[:, np.sort(np.random.default_rng(5).choice(links.shape[1], 2048, replace=False))]
        return dict(A=A, X=X, links=links, num_hops=3, K=3, flow='PoS', strategy=None, full='drnl',
                    desc=f"Cora (real X F=1433), 2048 links sampled (seed 5) of the 3 splits, NON-optimised flow "
                         f"(optimize_sign=False: SIGN on every subgraph row, DRNL label column) num_hops=3 sign_k=3")
    raise SystemExit(f"unknown workload {name}")


_RMAT = {}


def build_rmat_graph(args, dev):
    """BASELINE config 5: R-MAT (0.57, 0.19, 0.19, 0.05), ids reduced mod N, symmetrised, de-duplicated;
    F = 128 row-normalised features.  Built on the GPU once per process (seeds 42 / 43)."""
    import torch
    from s3grl_b200 import DeviceGraph, datasets as ds
    key = (args.rmat_nodes, args.rmat_edges, str(dev))
    if key not in _RMAT:
        _RMAT.clear()
        N, E = args.rmat_nodes, args.rmat_edges
        scale = int(np.ceil(np.log2(N)))
        indptr, indices = ds.rmat_csr_torch(scale, E, N, seed=42, device=dev)
        gen = torch.Generator(device=dev)
        gen.manual_seed(43)
        X = torch.rand((N, 128), device=dev, generator=gen)
        X /= X.sum(1, keepdim=True)
        g = DeviceGraph.from_device_csr(indptr, indices, X)
        _RMAT[key] = dict(g=g, indptr=indptr, indices=indices, deg=indptr[1:] - indptr[:-1], scale=scale)
    return _RMAT[key]


def build_rmat_workload(args, dev, degree_cap=None, num_links=None, max_nodes_per_hop=None):
    """Targets on the R-MAT graph: half stored edges, half random pairs (seed 44), num_hops = 1, sign_k = 3.
    degree_cap        restrict targets to endpoints of degree <= cap (0 / None: any degree).  The EXACT one-hop subgraph
                      of a hub link has 10^4..10^5 nodes and ~10^8 adjacency entries to intersect — the reference only
                      copes with such graphs through its per-hop caps (utils.py:66-70), see DESIGN.md §6.
    max_nodes_per_hop the reference's cap, with this implementation's deterministic rank rule: every subgraph has at
                      most 2 + cap nodes, so targets of ANY degree are served."""
    import torch
    G = build_rmat_graph(args, dev)
    g, indptr, indices, deg = G['g'], G['indptr'], G['indices'], G['deg']
    N, E = args.rmat_nodes, args.rmat_edges
    Lk = num_links or args.rmat_links
    cap = args.rmat_degree_cap if degree_cap is None else degree_cap
    lim = cap if cap and cap > 0 else int(deg.max())
    gen = torch.Generator(device=dev)
    gen.manual_seed(44)
    nnz = indices.numel()
    # positives: uniformly sampled stored entries (u, v) with both degrees <= cap
    pos = torch.empty((2, 0), dtype=torch.int64, device=dev)
    while pos.shape[1] < Lk // 2:
        e = torch.randint(0, nnz, (Lk,), device=dev, generator=gen)
        u = torch.searchsorted(indptr, e, right=True) - 1
        v = indices[e].to(torch.int64)
        ok = (deg[u] <= lim) & (deg[v] <= lim)
        pos = torch.cat([pos, torch.stack([u[ok], v[ok]])], 1)
    neg = torch.empty((2, 0), dtype=torch.int64, device=dev)
    while neg.shape[1] < Lk - Lk // 2:
        u = torch.randint(0, N, (Lk,), device=dev, generator=gen)
        v = torch.randint(0, N, (Lk,), device=dev, generator=gen)
        ok = (u != v) & (deg[u] <= lim) & (deg[v] <= lim)
        neg = torch.cat([neg, torch.stack([u[ok], v[ok]])], 1)
    links = torch.cat([pos[:, :Lk // 2], neg[:, :Lk - Lk // 2]], 1).contiguous()
    desc = (f"synthetic R-MAT (0.57,0.19,0.19,0.05) scale {G['scale']}, N={N}, {E} edge samples -> nnz={nnz}, max degree "
            f"{g.max_degree}; X F=128 uniform row-normalised; {Lk} targets (half edges, half random pairs) with endpoint "
            f"degree {'<= ' + str(cap) if cap and cap > 0 else 'unrestricted'}; PoS num_hops=1 sign_k=3"
            + (f", max_nodes_per_hop={max_nodes_per_hop} (deterministic rank rule)" if max_nodes_per_hop else ", exact subgraphs (no per-hop cap)")
            + "; graph and targets generated on the GPU (seeds 42/43/44)")
    return dict(graph=g, links_dev=links, links=links.cpu().numpy(), num_hops=1, K=3, flow='PoS', strategy=None, desc=desc,
                A=None, X=None, caps=dict(max_nodes_per_hop=max_nodes_per_hop) if max_nodes_per_hop else {})


# ----------------------------------------------------------------------------------------
# CPU baseline: the oracle port on all host cores (bounded sample)
# ----------------------------------------------------------------------------------------
_W = {}


def _cpu_init(w):
    os.environ['OMP_NUM_THREADS'] = '1'
    _W.update(w)


def _cpu_chunk(cols):
    from oracle import s3grl_oracle as orc
    w = _W
    if w.get('walk'):
        sub = w['links'][:, cols]
        sets = orc.random_walk_sets(w['A'], sub.reshape(-1), w['walk']['m'], w['walk']['M'], w['walk']['seed'])
        out = orc.scaled_pos_precompute(sub, sets, w['A'], w['X'], w['K'])
    elif w['flow'] == 'SoP':
        if 'sop_powers' not in w:       # the reference builds the global powers once per split
            w['sop_powers'] = orc.sop_powers(w['A'], w['K'])
        out = orc.sop_precompute(w['links'][:, cols], w['A'], w['X'], w['K'], powers=w['sop_powers'])
    elif w.get('full'):
        out = orc.full_precompute(w['links'][:, cols], w['num_hops'], w['A'], w['X'], w['K'], w['full'])
    else:
        out = orc.pos_precompute(w['links'][:, cols], w['num_hops'], w['A'], w['X'], w['K'], w['strategy'])
    return int(out['row_ptr'][-1])


def cpu_pass(w, sample_cols, procs):
    """One pass of the oracle over the sampled links with `procs` processes -> seconds."""
    import multiprocessing as mp
    chunks = [c for c in np.array_split(sample_cols, procs * 4) if c.size]
    ctx = mp.get_context('fork')
    with ctx.Pool(procs, initializer=_cpu_init, initargs=(w,)) as pool:
        pool.map(_cpu_chunk, [chunks[0][:1]] * procs)      # start-up outside the timing
        t0 = time.perf_counter()
        pool.map(_cpu_chunk, chunks)
        return time.perf_counter() - t0


def cpu_sample(w, per_core):
    procs = os.cpu_count() or 1
    n = min(w['links'].shape[1], max(procs, per_core * procs))
    cols = np.sort(np.random.default_rng(123).choice(w['links'].shape[1], n, replace=False))
    return cols, procs


def run_reference_arm(args, w, rank, world):
    if rank != 0:
        return
    cols, procs = cpu_sample(w, args.cpu_links_per_core)
    for _ in range(max(args.warmup, 0) and 1):
        cpu_pass(w, cols[:procs * 2], procs)
    t = sum(cpu_pass(w, cols, procs) for _ in range(args.steps))
    value = cols.size * args.steps / t
    sample = (f"{cols.size} links sampled uniformly (seed 123) from the workload's {w['links'].shape[1]}, "
              f"per step; links are independent, so links/s extrapolates linearly")
    line = dict(metric=metric_name(args.workload), value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000 * t / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=w['desc'], links_per_step=int(cols.size)),
                cpu_baseline=dict(value=value, unit=UNIT, cores=procs, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="oracle port (NumPy/SciPy restatement of reference utils.py:47-85 + tuned_SIGN.py:137-262, "
                     "validated against the reference in tests/golden) in a fork pool over all host cores; the "
                     "reference itself is single-threaded Python and cannot be shipped to the GPU box")
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons during the timed region.  NVML in a background thread (nvidia_ml_py) — an
    `nvidia-smi -lms` child process was measured to stall host-synchronising flows (PoS Plus) by 30-100 ms per
    query — with the nvidia-smi loop as the fallback when NVML bindings are missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        self.thread = None
        self.samples = []
        if os.environ.get('S3GRL_BENCH_NO_CLOCKS'):
            return
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            idx = str(index)
            h = pynvml.nvmlDeviceGetHandleByUUID(idx) if idx.startswith('GPU-') else pynvml.nvmlDeviceGetHandleByIndex(int(idx))
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.stop_flag = threading.Event()

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                             int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))))
                    except Exception:
                        pass
                    self.stop_flag.wait(0.1)
            self.pynvml = pynvml
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                       '-lms', '200', '-i', str(index)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            nv = self.pynvml
            names = [('hw_slowdown', nv.nvmlClocksEventReasonHwSlowdown), ('hw_thermal_slowdown', nv.nvmlClocksEventReasonHwThermalSlowdown),
                     ('sw_thermal_slowdown', nv.nvmlClocksEventReasonSwThermalSlowdown), ('sw_power_cap', nv.nvmlClocksEventReasonSwPowerCap)]
            sm = [s for s, _ in self.samples]
            if not sm:
                return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
            reasons = sorted({nm for _, r in self.samples for nm, bit in names if r & bit})
            busy = [s for s in sm if s >= 0.5 * max(sm)]
            return dict(sm_mhz=float(np.median(busy)), sm_max_mhz=self.mx, reasons=reasons, samples=len(sm), source="nvml")
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.p.terminate()
        out = self.p.communicate()[0]
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(',')]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        busy = [s for s in sm if s >= 0.5 * max(sm)]
        return dict(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                    samples=len(sm), source="nvidia-smi")


# ----------------------------------------------------------------------------------------
# measured ceilings besides HBM: L2 -> SM read bandwidth and FP32 FMA rate (csrc/probe.cu)
# ----------------------------------------------------------------------------------------
def measure_ceilings(dev):
    import ctypes as C
    import torch
    from s3grl_b200 import _lib as L
    lib = L.lib()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    buf = torch.rand(48 << 18, dtype=torch.float32, device=dev)        # 48 MB: stays in the 126 MB L2
    sink = torch.zeros(4, dtype=torch.float32, device=dev)
    out = {}
    for name, call, work in (
            ('l2_read_GBps', lambda it: lib.s3_probe_l2_read(C.c_void_p(buf.data_ptr()), buf.numel() * 4, it, C.c_void_p(sink.data_ptr()), sms * 8, st),
             lambda it: buf.numel() * 4 * it / 1e9),
            ('fp32_TFLOPs', lambda it: lib.s3_probe_fma(it * 64, C.c_void_p(sink.data_ptr()), sms * 8, st),
             lambda it: 2.0 * sms * 8 * 256 * it * 64 * 128 / 1e12),
            ('fp32x2_TFLOPs', lambda it: lib.s3_probe_fma2(it * 64, C.c_void_p(sink.data_ptr()), sms * 8, st),
             lambda it: 2.0 * sms * 8 * 256 * it * 64 * 128 / 1e12)):
        L.check(call(2), name)
        best = 0.0
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            L.check(call(20), name)
            b.record()
            b.synchronize()
            best = max(best, work(20) / (a.elapsed_time(b) / 1e3))
        out[name] = best
    return out


# ----------------------------------------------------------------------------------------
# one workload on this process group -> the fields of its JSON record
# ----------------------------------------------------------------------------------------
def measure(args, w, wname, steps, warmup, dev, rank, world, want_e2e, ceilings, cpu=None):
    import torch
    import torch.distributed as dist
    from s3grl_b200 import DeviceGraph, algorithmic_bytes, precompute, precompute_full
    from s3grl_b200 import tuned_sign
    from s3grl_b200.parallel import PeerBuffers, exchange_finish, precompute_exchange, shard_range

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    is_rmat = w.get('graph') is not None
    links_all = w['links']                       # host [2, L]: the whole list of the workload
    Lk = links_all.shape[1]
    K = w['K']
    g = w['graph'] if is_rmat else DeviceGraph(w['A'], w['X'], device=dev)
    F = g.num_feat
    full_flow = w.get('full')
    fixed = w['strategy'] is None and not full_flow
    exchange = world > 1 and fixed
    a, b = (0, Lk) if (world == 1 or exchange) else shard_range(Lk, rank, world)
    links_dev = (w['links_dev'] if is_rmat else torch.from_numpy(np.ascontiguousarray(links_all)).to(dev))[:, a:b].contiguous()
    Lmine = b - a
    caps = w.get('caps') or {}
    buffers = PeerBuffers(Lk, F, K, dev, backend=args.exchange_backend) if exchange else None
    exchange_backend = buffers.backend if exchange else None
    out = ([torch.empty((2 * Lk, F + 1), dtype=torch.float32, device=dev) for _ in range(K + 1)]
           if fixed and not exchange else None)
    # L2 between timed iterations: a step of the fixed-row flows WRITES its K+1 operator matrices — far more than the
    # 126 MB L2 (PubMed: 2.6 GB, on every rank of an exchange) — so every step starts with none of its inputs cached;
    # the flows with smaller outputs get an explicit 256 MiB fill
    out_bytes = 2 * Lk * (K + 1) * (F + 1) * 4
    flush = None if (fixed and out_bytes >= (512 << 20)) else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(profile=None):
        if flush is not None:
            flush.zero_()               # L2 flush between steps
        if full_flow:
            return precompute_full(g, links_dev, w['num_hops'], K, node_label=full_flow, batch_records=args.batch_records,
                                   profile=profile)
        if exchange:
            res, mirror = precompute_exchange(g, links_dev, w['num_hops'], K, buffers, flow=w['flow'], defer=True,
                                              batch_records=args.batch_records, profile=profile, overlap=args.overlap,
                                              walk=w.get('walk'), **caps)
            exchange_finish(buffers, mirror)    # on the stream: barrier (all rows have landed), then the paired links' rows
            return res
        return precompute(g, links_dev, w['num_hops'], K, w['flow'], w['strategy'], batch_records=args.batch_records, out=out,
                          profile=profile, overlap=args.overlap, defer=fixed, walk=w.get('walk'), pair=not args.no_pair, **caps)

    vis = [v for v in os.environ.get('CUDA_VISIBLE_DEVICES', '').split(',') if v.strip()]
    local_rank = dev.index
    clocks = ClockSampler(vis[local_rank] if local_rank < len(vis) else local_rank)
    res = None
    for _ in range(max(warmup, 3)):     # same code path as the timed steps: per-kernel events, deferred validation
        del res
        res = step([])
        res.finalize()
    if not fixed:
        res.xs = None
        if full_flow:
            res.node_id = None
    barrier()
    # One step of a fixed-row flow has no host-dependent control flow: capture it once as a CUDA graph and replay it
    # K times (one launch per step, so a descheduled host thread on a shared box cannot starve the GPU between the
    # ~20 kernel launches of a step). The events around every kernel are event-record nodes of the graph: after the
    # last replay they hold that step's per-kernel times. Falls back to eager launches if the capture fails.
    graph_step, graph_res, graph_profile, graph_note = None, None, [], None
    if fixed and not args.no_graph:
        try:
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                graph_res = step(graph_profile)
            cg.replay()
            torch.cuda.synchronize(dev)
            graph_step = cg
        except Exception as ex:
            graph_step, graph_res, graph_profile = None, None, []
            graph_note = f"capture failed, eager launches: {type(ex).__name__}: {ex}"[:200]
            torch.cuda.synchronize(dev)
        if world > 1:       # every rank replays a graph, or none does
            okg = torch.tensor([1 if graph_step is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(okg, op=dist.ReduceOp.MIN)
            if int(okg) == 0 and graph_step is not None:
                graph_step, graph_res, graph_profile, graph_note = None, None, [], "capture failed on another rank, eager launches"
        barrier()
    profile = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    launches = 0
    step_events = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    step_events[0].record()
    pending, r = [], None
    t_enq = time.perf_counter()
    for i in range(steps):
        if graph_step is not None:
            graph_step.replay()
            res = graph_res
        else:
            res = step(profile)     # K steps are queued back to back ...
            if not fixed:
                # flows whose outputs are allocated inside the call (data-dependent row counts): hand the rows back to
                # the allocator before the next step, as a caller that consumes each result would
                res.xs = None
                if full_flow:
                    res.node_id = None
            pending.append(res)
        step_events[i + 1].record()
        launches += res.stats['launches']               # kernels of libs3grl_b200.so only (not the L2 flush fill)
    host_enqueue_ms = 1000 * (time.perf_counter() - t_enq) / steps
    if graph_step is not None:
        torch.cuda.current_stream(dev).synchronize()
        graph_res.finalize()                # validation of the replayed step (its counters), inside the timed region
        profile = graph_profile             # per-kernel events of the LAST replayed step
    for r in pending:                       # ... then synchronised and validated, inside the timed region
        r.finalize()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = Lk * steps / (ms_max / 1e3)          # whole job: the list is processed once per step by all ranks together

    # ---- multi-GPU: every rank's matrices against a single-GPU precompute of the whole list ----
    exch = None
    if exchange:
        check = precompute(g, links_dev, w['num_hops'], K, w['flow'], None, pair=False, walk=w.get('walk'), **caps)
        same = torch.tensor([int(all(torch.equal(buffers.local[k], check.xs[k]) for k in range(K + 1)))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        rows_mine = 2 * (res.stats['links'])
        # NVLink egress of a rank: its rows once through the NVSwitch multicast object, or once per peer over P2P
        rows_sent = 2 * (res.stats['links'] - res.stats['mirrors'])     # paired links' rows and operator 0 stay local
        nv = rows_sent * K * (F + 1) * 4 * (1 if exchange_backend == 'multicast' else world - 1)
        exch = dict(kind="s3_gather_peers: kernel 3 stores every output row into all ranks' matrices ("
                         + ("one store to an NVSwitch multicast address, torch symmetric memory as plumbing" if exchange_backend == 'multicast'
                            else "one store per peer over NVLink P2P, cudaMalloc + CUDA IPC")
                         + "); the timed region ends with one 4-byte NCCL all-reduce per step as the barrier",
                    backend=exchange_backend,
                    equal_to_single_gpu_bitwise=bool(int(same)), nvlink_bytes_out_per_rank_per_step=int(nv),
                    nvlink_GBps_out_per_rank=nv * steps / (ms_max / 1e3) / 1e9, nvlink_peak_GBps=770.0,
                    nvlink_peak_source="B200_PROFILING.md: measured peer copy, per direction per GPU")
        del check

    # ---- per-kernel times from the events recorded on the launching stream during the timed steps ----
    stage_ms = {}
    for stage, bi, ea, eb in profile:
        stage_ms.setdefault(stage, []).append(ea.elapsed_time(eb))
    ev_steps = 1 if graph_step is not None else steps      # steps the per-kernel events cover
    if os.environ.get('S3GRL_BENCH_DEBUG'):
        for stage, v in stage_ms.items():
            print(stage, [round(t_, 2) for t_ in v[:3 * res.stats['batches']]], file=sys.stderr)
    st = res.stats
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if 'hbm_gbs' in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    rows_written = st['rows'] if not exchange else 2 * st['links']      # rows this rank produced (paired links included)
    names = dict(extract="front_sorted_kernel" if is_rmat else "front_kernel", gather="gather_kernel", diffuse="diffuse_kernel",
                 gather_ccn="gather_kernel (CCN work items)", ccn_chain="chain_kernel", sign_full="sign_full_kernel",
                 collate="scatter_rows_kernel")
    # SURVEY 8d terms per kernel, over the records this rank actually extracted
    alg = dict(extract=4 * st['sum_d'] + 8 * st['sum_n'],
               gather=4 * F * st['sum_n'] + 4 * rows_written * (K + 1) * (F + 1),
               sign_full=4 * F * st['sum_n'] + 4 * st['rows'] * (K + 1) * (F + 1),
               # union: the chain reads every subgraph's feature rows once and writes the CCN rows of all K+1 operators
               ccn_chain=4 * F * st['sum_n'] + 4 * max(0, (st.get('rows_computed') or st['rows']) - 2 * st['records']) * (K + 1) * (F + 1),
               collate=2 * 4 * st['rows'] * (K + 1) * (F + 1))
    bound_note = dict(extract="issue / latency bound integer work on shared-memory bitmaps (ncu: profiles/); its HBM fraction is "
                              "low by nature" if not is_rmat else "latency bound sorted-set intersections over HBM-resident adjacency lists",
                      gather="L2 -> SM bandwidth and FP32 issue when X is L2-resident (PubMed 39 MB), HBM otherwise",
                      sign_full="HBM writes",
                      collate="HBM copy (placement of the batches' pieces, replication of paired links' rows)",
                      ccn_chain="shared-memory bandwidth: every induced edge of a level reads one row segment of the previous "
                                "level's [n][CW] buffer (DESIGN.md section 4); X is L2-resident on PubMed; the records that would "
                                "run at CW <= 8 keep their buffers in a global-memory pool and read them through L2 instead "
                                "(s3_ccn_chain_pooled) — the same row segments, counted in the same figure")
    kernels = {}
    for stage, v in stage_ms.items():
        tot = float(np.sum(v))
        k = dict(kernel=names.get(stage, stage), share_of_step=tot / (ms * ev_steps / steps), ms_per_step=tot / ev_steps, launches_timed=len(v),
                 avg_launch_ms=tot / len(v))
        if stage in alg and tot > 0:
            k['algorithmic_bytes_per_step'] = int(alg[stage])
            k['achieved_GBps'] = alg[stage] * ev_steps / (tot / 1e3) / 1e9
            k['frac_of_hbm'] = k['achieved_GBps'] / peak
            k['bound_by'] = bound_note.get(stage)
        if stage == 'ccn_chain' and tot > 0 and st.get('chain_reads'):
            # the chain's own roofline: every induced edge of a level reads one row segment of the previous level's
            # shared-memory buffer; over all sub-chunks that is 4 * (F + 1) bytes per (edge, level)
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            smem_peak = 128.0 * sms * float(clk.get('sm_max_mhz') or 1965.0) * 1e6 / 1e9      # 128 B / clk / SM
            sm_bytes = st['chain_reads'] * 4 * (F + 1)
            k['smem_bytes_per_step'] = int(sm_bytes)
            k['smem_GBps'] = sm_bytes * ev_steps / (tot / 1e3) / 1e9
            k['smem_peak_GBps'] = smem_peak
            k['frac_of_smem_peak'] = k['smem_GBps'] / smem_peak
            k['smem_peak_source'] = "128 B/clk/SM (B300_MICROARCH.md, LDS crossbar) x SMs x max SM clock; conflict-free"
            if ceilings and ceilings.get('l2_read_GBps'):       # the pooled records' segments come from L2: second ceiling
                k['l2_read_ceiling_GBps'] = ceilings['l2_read_GBps']
                k['frac_of_l2_read_ceiling'] = k['smem_GBps'] / ceilings['l2_read_GBps']
            k['records_chained'] = st.get('chain_records')
            k['mean_n_chained'] = st['chain_n'] / max(1, st.get('chain_records') or 1)
        if stage == 'gather' and tot > 0 and ceilings:
            # kernel 3 reads every feature row from L2 once per column pass and issues (K+1-kmin) * SC FMAs per float
            hop = res.hop_nodes() if hasattr(res, 'hop_nodes') else None
            k['l2_read_ceiling_GBps'] = ceilings['l2_read_GBps']
            k['frac_of_l2_read_ceiling'] = k['achieved_GBps'] / ceilings['l2_read_GBps']
            if hop:
                sc = 1 if w['flow'] == 'SoP' else 2
                fma = sum(n * max(0, K + 1 - l) * sc for l, n in enumerate(hop)) * g.ldx
                k['fp32_TFLOPs'] = 2.0 * fma * ev_steps / (tot / 1e3) / 1e12
                k['fp32_ceiling_TFLOPs'] = ceilings['fp32_TFLOPs']
                k['frac_of_fp32_ceiling'] = k['fp32_TFLOPs'] / ceilings['fp32_TFLOPs']
        kernels[stage] = k
    dom = max(kernels, key=lambda s_: kernels[s_]['share_of_step']) if kernels else None
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')))
        traffic = tj.get(wname, {}).get(dom)
        traffic_src = tj.get('_source')
    except Exception:
        pass
    path_bytes = algorithmic_bytes(st, F, K, per_link=True) if not exchange else None
    if exchange:      # per-link figures of the whole list: sum over ranks
        tt = torch.tensor([st['sum_n_links'] or st['sum_n'], st['sum_d_links'] or st['sum_d']], dtype=torch.int64, device=dev)
        dist.all_reduce(tt)
        path_bytes = 4 * int(tt[1]) + 8 * int(tt[0]) + 4 * F * int(tt[0]) + 4 * 2 * Lk * (K + 1) * (F + 1)
    elif world > 1:
        tt = torch.tensor([path_bytes], dtype=torch.int64, device=dev)
        dist.all_reduce(tt)
        path_bytes = int(tt[0])
    d = kernels.get(dom, {})
    streamed = None
    if is_rmat and st.get('sum_read'):
        # the sorted tier does not scan the adjacency list of a hub row (binary-search probes instead): index bytes it
        # really streams = the two merged lists + the rows of the low-degree nodes, against the 4*D of SURVEY 8d
        rd = torch.tensor([st['sum_read'], st['sum_n']], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(rd)
        sb = 4 * int(rd[0]) + 16 * int(rd[1]) + 4 * F * int(rd[1]) + 4 * 2 * Lk * (K + 1) * (F + 1)
        streamed = dict(bytes_per_link=sb / Lk, achieved=sb * steps / (ms_max / 1e3) / 1e9,
                        frac=sb * steps / (ms_max / 1e3) / 1e9 / (peak * world),
                        note="index bytes the bit-matrix method really streams (merged lists + low-degree rows; hub rows are "
                             "probed) + indptr + features + outputs: the honest denominator for this tier; path.frac above "
                             "keeps SURVEY 8d's 4*D, which this method never reads")
    roofline = dict(bound="hbm", kernel=d.get('kernel'), achieved=d.get('achieved_GBps', 0.0), peak=peak, unit="GB/s",
                    frac=d.get('frac_of_hbm', 0.0), traffic=traffic, traffic_source=traffic_src, peak_source=peak_src,
                    dominant_by="share of the step (CUDA events on the launching stream" + ("; event-record nodes of the replayed CUDA graph, last timed step)" if graph_step is not None else ")"),
                    note=d.get('bound_by'),
                    bytes_per_launch=d.get('algorithmic_bytes_per_step', 0) / max(st['batches'], 1),
                    launches_timed=d.get('launches_timed'), avg_launch_ms=d.get('avg_launch_ms'),
                    share_of_step={k_: v_['share_of_step'] for k_, v_ in kernels.items()},
                    kernels=kernels, dominant=d,
                    path=dict(achieved=path_bytes * steps / (ms_max / 1e3) / 1e9,
                              frac=path_bytes * steps / (ms_max / 1e3) / 1e9 / (peak * world),
                              bytes_per_link=path_bytes / Lk, peak=peak * world,
                              note="whole path: SURVEY 8d bytes per LINK (a paired link counts its own bytes although it "
                                   "is served by its partner's record) over the full step time, against N x the HBM peak"))
    if st.get('mirrors') is not None and fixed:
        roofline['path']['links_served_by_pairing_this_rank'] = st['mirrors']
    if streamed:
        roofline['path_streamed'] = streamed
    if getattr(g, '_hub', None) and g._hub[2]:
        roofline['hub_index'] = dict(hubs=g._hub[2], min_degree=getattr(g, 'hub_min_degree', None),
                                     bit_matrix_bytes=int(g._hub[1].numel()) * 4)

    # ---- end to end through the reference-facing call, host buffers in, host buffers out ----
    e2e = None
    if want_e2e and not is_rmat:
        from s3grl_b200 import extract_enclosing_subgraphs
        sa, sb = shard_range(Lk, rank, world)          # every rank: its contiguous shard, to its own host memory
        x_host = torch.from_numpy(w['X']).pin_memory()
        link_index = torch.from_numpy(np.ascontiguousarray(links_all[:, sa:sb])).pin_memory()
        sign_kwargs = dict(sign_k=K, use_feature=True, sign_type=w['flow'], optimize_sign=not full_flow,
                           k_heuristic=0 if w['strategy'] is None else 1, k_node_set_strategy=w['strategy'])
        del out, res, pending, r
        if buffers is not None:
            buffers.close()
            buffers = None
        torch.cuda.empty_cache()

        def e2e_step():
            tuned_sign._graph_cache.clear()       # the graph upload is part of every step
            rw_kwargs = dict(rw_m=w['walk']['m'], rw_M=w['walk']['M'], seed=w['walk']['seed'], sign=True) if w.get('walk') else None
            lst = extract_enclosing_subgraphs(link_index, w['A'], x_host, 1, w['num_hops'], full_flow or 'zo', 1.0, None, False,
                                              None, rw_kwargs, sign_kwargs, powers_of_A=[] if w['flow'] == 'PoS' else [None] * K,
                                              data=None, device=dev, output_device='cpu')
            d2h = sum(x.numel() * 4 for x in lst.xs) + lst.row_ptr.numel() * 8
            chk = float(lst.xs[-1][0, 0]) if lst.xs[-1].numel() else 0.0        # touch the host result
            return d2h, chk
        for _ in range(2):
            e2e_step()
        n_e2e = args.e2e_steps or max(2, min(steps, 10))
        import gc
        gc.collect()                 # start the timed region with a clean heap (host hiccups show up in e2e.step_ms)
        barrier()
        t0 = time.perf_counter()
        e2e_ms = []
        for _ in range(n_e2e):
            ts = time.perf_counter()
            d2h, _ = e2e_step()
            e2e_ms.append(round(1000 * (time.perf_counter() - ts), 2))
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        h2d = g.h2d_bytes + link_index.numel() * 8
        e2e = dict(value=Lk * n_e2e / float(tt.item()), unit=UNIT, h2d_bytes_per_step=int(h2d),
                   d2h_bytes_per_step=int(d2h), steps=n_e2e, ms_per_step=1000 * float(tt.item()) / n_e2e, step_ms=e2e_ms,
                   api="s3grl_b200.extract_enclosing_subgraphs(link_index, A, x, y, num_hops, ..., sign_kwargs, output_device='cpu') "
                       "-> pinned host tensors" + ("" if world == 1 else f"; every rank uploads the graph, takes the contiguous shard "
                       f"rank/{world} of the link list and returns its rows to its own host (bytes are per rank)"))
    if buffers is not None:
        buffers.close()

    par = ("1 GPU" if world == 1 else
           (f"strong scaling: link list sharded cyclically x{world}, graph replicated, all-gather of the operator rows fused "
            f"into kernel 3 (NVLink peer stores) inside the timed region" if exchange else
            f"strong scaling: link list sharded in contiguous blocks x{world}, graph replicated, rows stay on their rank "
            f"(data-dependent row counts: no exchange step)"))
    line = dict(metric=metric_name(wname), value=value, unit=UNIT, n_gpus=world, steps=steps, warmup=max(warmup, 3),
                ms_per_step=ms_max / steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload=w['desc'], links_per_step=Lk, links_per_step_this_rank=st['links'],
                            batch_records=args.batch_records or "auto (32768 for fixed-row flows)",
                            l2=(f"every step writes {out_bytes / 2**30:.1f} GiB of operator rows (>> 126 MB L2) before the next one starts: "
                                "no input survives in L2 between steps" if flush is None else "flushed between steps (256 MiB memset)")
                               + "; PubMed's X (39 MB) is L2-resident WITHIN a step by nature of the workload, the R-MAT X (5 GB) is not",
                            pairing=("on: links over the same unordered node pair share one record (both directions of a "
                                     "training edge); bit-identical to computing each" if fixed and not args.no_pair and not is_rmat
                                     else ("on: one record per unordered node pair, the other direction's rows are placed by "
                                           "s3_scatter_rows (rows 0 / 1 exchanged; CCN rows within fp32 rounding of computing each)"
                                           if not fixed and not full_flow and not args.no_pair and not is_rmat else "off")),
                            parallelism=par),
                clocks=clk, e2e=e2e, gpu_launches=launches, roofline=roofline, cpu_baseline=cpu,
                timed_region=("one CUDA graph replay per step (captured from the same engine call after the warm-up)" if graph_step is not None
                              else "eager kernel launches" + (f" ({graph_note})" if graph_note else "")),
                host_enqueue_ms_per_step=host_enqueue_ms,
                step_ms=[round(step_events[i].elapsed_time(step_events[i + 1]), 3) for i in range(steps)])
    if exch:
        line['exchange'] = exch
    return line


def rmat_cpu_baseline(w, g, K):
    """single-process oracle on a small sample (the fork pool cannot be used once CUDA is up)"""
    import scipy.sparse as ssp
    from oracle import s3grl_oracle as orc
    F, Lk = g.num_feat, w['links'].shape[1]
    A_host = ssp.csr_matrix((np.ones(g.nnz, np.int8), g.indices.cpu().numpy(), g.indptr.cpu().numpy()),
                            shape=(g.num_nodes, g.num_nodes))
    X_host = g.x[:, :F].cpu().numpy()
    cols = np.random.default_rng(123).choice(Lk, min(Lk, 200), replace=False)
    t0 = time.perf_counter()
    orc.pos_precompute(w['links'][:, cols], 1, A_host, X_host, K, caps=w.get('caps') or None)
    dt = time.perf_counter() - t0
    return dict(value=cols.size / dt, unit=UNIT, cores=1, kind="port",
                sample=f"{cols.size} links sampled uniformly (seed 123), one pass, {dt:.1f} s; single-process oracle port")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='pubmed_pos')
    ap.add_argument('--configs', default='auto', help="'auto': with the default workload also measure pubmed_posplus_union and rmat "
                    "(few steps) and report them under \"configs\"; 'none'; or a comma-separated list of workloads")
    ap.add_argument('--batch-records', type=int, default=None)
    ap.add_argument('--links', type=int, default=None, help='use only the first N links of the workload (profiling runs)')
    ap.add_argument('--cpu-links-per-core', type=int, default=400)
    ap.add_argument('--e2e-steps', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='eager kernel launches in the timed region instead of one CUDA graph replay per step')
    ap.add_argument('--no-pair', action='store_true', help='switch link pairing off (every link extracts its own subgraph)')
    ap.add_argument('--rmat-nodes', type=int, default=10_000_000)
    ap.add_argument('--rmat-edges', type=int, default=200_000_000)
    ap.add_argument('--rmat-links', type=int, default=4_000_000)
    ap.add_argument('--rmat-degree-cap', type=int, default=512, help='targets restricted to endpoints of at most this degree (0: any)')
    ap.add_argument('--max-nodes-per-hop', type=int, default=None, help="rmat: the reference's per-hop cap (deterministic rank rule)")
    ap.add_argument('--exchange-backend', default=None, choices=['auto', 'multicast', 'ipc'],
                    help='default: ipc on 2 GPUs, multicast (with ipc as fallback) from 4 GPUs on')
    ap.add_argument('--overlap', action='store_true', help='two-stream schedule: front kernel of batch i+1 beside kernel 3 of batch i')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    is_rmat = args.workload.startswith('rmat')
    w = None if is_rmat else build_workload(args.workload)
    if w is not None and args.links:
        w['links'] = np.ascontiguousarray(w['links'][:, :args.links])
        w['desc'] += f' (first {args.links} links only)'

    if args.impl == 'reference':
        if is_rmat:
            if rank == 0:
                print(json.dumps(dict(impl="reference", unavailable="the rmat workload is generated on the GPU; run the "
                                      "default PubMed workload for the reference arm")))
            return
        run_reference_arm(args, w, rank, world)
        return

    # CPU baseline first: the fork pool must not inherit an initialised CUDA context
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not is_rmat:
        cols, procs = cpu_sample(w, args.cpu_links_per_core)
        t = cpu_pass(w, cols, procs)
        cpu = dict(value=cols.size / t, unit=UNIT, cores=procs, kind="port",
                   sample=f"{cols.size} links sampled uniformly (seed 123) of {w['links'].shape[1]}, one pass, "
                          f"{t:.1f} s; oracle port in a {procs}-process fork pool")

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the box exports NCCL_DEBUG=VERSION: send fd 1 to stderr while
        # the communicator comes up, so that stdout carries the one JSON line and nothing else
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    ceilings = measure_ceilings(dev)
    if is_rmat:
        w = (build_rmat_workload(args, dev, degree_cap=0, max_nodes_per_hop=args.max_nodes_per_hop or 1024) if args.workload == 'rmat_hopcap'
             else build_rmat_workload(args, dev, max_nodes_per_hop=args.max_nodes_per_hop))
        if not args.no_cpu_baseline and rank == 0 and world == 1:
            cpu = rmat_cpu_baseline(w, w['graph'], w['K'])
    line = measure(args, w, args.workload, args.steps, args.warmup, dev, rank, world, not args.no_e2e, ceilings, cpu)
    line['ceilings'] = dict(ceilings, hbm_copy_GBps=line['roofline']['peak'],
                            note="measured in this run by csrc/probe.cu (L2-resident 48 MB buffer streamed with 128-bit loads; "
                                 "8 independent FFMA chains per thread); HBM copy peak from MEASURED_PEAKS.json")

    # ---- the BASELINE configs that are not the headline, few steps each ----
    extra = []
    if args.configs == 'auto':
        extra = (['pubmed_posplus_union', 'rmat', 'rmat_deg4096', 'rmat_hopcap'] if args.workload == 'pubmed_pos' and not args.links
                 else [])
    elif args.configs != 'none':
        extra = [c for c in args.configs.split(',') if c]
    configs = []
    for name in extra:
        del w
        torch.cuda.empty_cache()
        try:
            if name == 'rmat':                 # BASELINE config 5 as in round 1: endpoint degree <= 512, exact subgraphs
                w = build_rmat_workload(args, dev)
            elif name == 'rmat_deg4096':       # larger hubs, exact subgraphs, a 400 k-link sample
                w = build_rmat_workload(args, dev, degree_cap=4096, num_links=min(args.rmat_links, 400_000))
            elif name == 'rmat_hopcap':        # targets of ANY degree through the reference's per-hop cap
                w = build_rmat_workload(args, dev, degree_cap=0, num_links=min(args.rmat_links, 1_000_000), max_nodes_per_hop=1024)
            else:
                w = build_workload(name)
            sub = measure(args, w, name, 2, 3, dev, rank, world, False, ceilings, None)
            keep = {k_: sub[k_] for k_ in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'ms_per_step', 'scaling', 'gpu_launches')}
            keep['config'] = sub['config']
            keep['roofline'] = {k_: sub['roofline'][k_] for k_ in ('kernel', 'achieved', 'peak', 'frac', 'share_of_step', 'path', 'path_streamed', 'hub_index', 'dominant')
                                if k_ in sub['roofline']}
            if 'exchange' in sub:
                keep['exchange'] = sub['exchange']
            configs.append(keep)
        except Exception as ex:      # a secondary config must not take the headline line down with it
            configs.append(dict(metric=metric_name(name), error=f"{type(ex).__name__}: {ex}"[:300]))
            w = None
    if rank == 0:
        if configs:
            line['configs'] = configs
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
