#!/usr/bin/env python
"""Fused scoring head (s3_sign_head, tcgen05 TF32) on a PubMed-PoS-shaped epoch: 328 000 rows x 2004 columns
-> 164 000 x 256 pooled, against the HBM roofline (the GEMM is bound by reading the joint matrix) and next to
the same op in plain PyTorch (fp32 cuBLAS linear + elu + batch-norm affine + pooling).  One JSON line."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from s3grl_b200 import sign_head  # noqa: E402


def timed(fn, steps, flush):
    ts = []
    for it in range(3 + steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', type=int, default=328000)
    ap.add_argument('--kd', type=int, default=2004)
    ap.add_argument('--steps', type=int, default=10)
    a = ap.parse_args()
    dev = torch.device('cuda')
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.rand((a.rows, a.kd), device=dev, generator=g) / 8
    W = (torch.rand((256, a.kd), device=dev, generator=g) - 0.5) * 0.2
    b = torch.rand(256, device=dev, generator=g) - 0.5
    scale = torch.rand(256, device=dev, generator=g) + 0.5
    shift = torch.rand(256, device=dev, generator=g) - 0.5
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = torch.empty((a.rows // 2, 256), device=dev)
    ms, got = timed(lambda: sign_head(X, W, b, scale, shift, out=out), a.steps, flush)

    def torch_ref():
        h = torch.nn.functional.elu(torch.nn.functional.linear(X, W, b)) * scale + shift
        return h[0::2] * h[1::2]
    torch.backends.cuda.matmul.allow_tf32 = False
    ms_fp32, ref = timed(torch_ref, max(2, a.steps // 3), flush)
    torch.backends.cuda.matmul.allow_tf32 = True
    ms_tf32, _ = timed(torch_ref, max(2, a.steps // 3), flush)
    err = float((got - ref).abs().max() / ref.abs().max())
    nbytes = a.rows * a.kd * 4 + (a.rows // 2) * 256 * 4 + 256 * a.kd * 4
    flops = 2.0 * a.rows * a.kd * 256
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    print(json.dumps(dict(kernel='sign_head_kernel (tcgen05.mma kind::tf32, TMA, TMEM)', rows=a.rows, kdim=a.kd, hidden=256,
                          ms=ms, links_per_s=a.rows / 2 / (ms / 1e3), algorithmic_bytes=nbytes,
                          achieved_GBps=nbytes / (ms / 1e3) / 1e9, peak_GBps=peak, frac=nbytes / (ms / 1e3) / 1e9 / peak,
                          tflops=flops / (ms / 1e3) / 1e12, torch_fp32_ms=ms_fp32, torch_tf32_ms=ms_tf32,
                          speedup_vs_torch_fp32=ms_fp32 / ms, speedup_vs_torch_tf32=ms_tf32 / ms,
                          max_rel_err_vs_torch_fp32=err,
                          note='bound: HBM read of the joint matrix; L2 flushed between launches; CUDA events')))


if __name__ == '__main__':
    main()
