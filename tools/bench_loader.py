#!/usr/bin/env python
"""Loader kernel (s3_joint_rows) against the HBM roofline on a PubMed-PoS-shaped collated dataset:
164 000 links x 2 rows, K+1 = 4 operators of F' = 501 columns (2.6 GB), shuffled epoch in ONE launch.
Prints one JSON line.   python tools/bench_loader.py [--links N] [--steps K]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from s3grl_b200 import joint_rows  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--links', type=int, default=164000)
    ap.add_argument('--cols', type=int, default=501)
    ap.add_argument('--ops', type=int, default=4)
    ap.add_argument('--steps', type=int, default=10)
    a = ap.parse_args()
    dev = torch.device('cuda')
    R = 2 * a.links
    xs = [torch.rand((R, a.cols), device=dev) for _ in range(a.ops)]
    row_ptr = torch.arange(a.links + 1, device=dev, dtype=torch.int64) * 2
    out = torch.empty((R, a.ops * a.cols), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    times = []
    for it in range(3 + a.steps):
        perm = torch.randperm(a.links, device=dev, generator=gen)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        joint_rows(xs, row_ptr, perm, 2, out=out, want_batch=False)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    nbytes = 2 * R * a.ops * a.cols * 4
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    # check against torch indexing (bit-exact)
    rows = (perm[:, None] * 2 + torch.arange(2, device=dev)[None]).reshape(-1)
    ref = torch.cat([x[rows] for x in xs], 1)
    ok = bool(torch.equal(ref, out))
    print(json.dumps(dict(kernel='joint_rows_kernel', links=a.links, rows=R, ops=a.ops, cols=a.cols, ms_per_epoch=ms,
                          links_per_s=a.links / (ms / 1e3), bytes=nbytes, achieved_GBps=nbytes / (ms / 1e3) / 1e9, peak_GBps=peak,
                          frac=nbytes / (ms / 1e3) / 1e9 / peak, matches_torch_indexing=ok,
                          note='read + write of every joint-matrix byte; L2 flushed between launches; CUDA events')))


if __name__ == '__main__':
    main()
