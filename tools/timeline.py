"""Print the in-stream timeline of one precompute call (event timestamps per launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from s3grl_b200 import DeviceGraph, precompute
w = bench.build_workload('pubmed_pos')
dev = torch.device('cuda', 0)
g = DeviceGraph(w['A'], w['X'], device=dev)
links = torch.from_numpy(w['links']).to(dev)
K, F = w['K'], w['X'].shape[1]
out = [torch.empty((2 * links.shape[1], F + 1), dtype=torch.float32, device=dev) for _ in range(K + 1)]
for _ in range(3):
    precompute(g, links, 3, K, out=out)
torch.cuda.synchronize()
prof = []
s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
s0.record()
res = precompute(g, links, 3, K, out=out, profile=prof)
s1.record(); torch.cuda.synchronize()
print('total', s0.elapsed_time(s1))
for stage, bi, a, b in prof[:12] + prof[-4:]:
    print(f"{stage:8s} batch {bi:3d} start {s0.elapsed_time(a):8.3f} end {s0.elapsed_time(b):8.3f} dur {a.elapsed_time(b):7.3f}")
