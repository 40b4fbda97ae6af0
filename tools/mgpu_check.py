"""Multi-GPU parity of the fused gather + all-gather (s3_gather_peers over NVLink peer memory).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/mgpu_check.py

Every rank takes its cyclic shard of the link list and stores its rows into every rank's operator matrices;
after the barrier every rank compares its own copy with a single-GPU precompute of the whole list: same bits.
Prints MGPU_OK on rank 0 (tests/test_multigpu.py looks for it)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from s3grl_b200 import DeviceGraph, datasets as ds, precompute  # noqa: E402
from s3grl_b200.parallel import PeerBuffers, precompute_exchange  # noqa: E402


def main():
    rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    ok = torch.ones(1, dtype=torch.int32, device=dev)
    for name, hops, K, flow, nl in (('cora', 3, 3, 'PoS', 5000), ('usair', 0, 3, 'SoP', 1500), ('cora', 2, 2, 'PoS', 777)):
        edges, N, X = ds.load_graph(name)
        A, splits = ds.split_links(edges, N, seed=1)
        X = ds.normalize_features(X) if X is not None else ds.synthetic_features(N, 40, 0.5, 0)
        links = ds.all_links(splits)[:, :nl]
        g = DeviceGraph(A, X, device=dev)
        for backend in ('auto', 'ipc'):
            buf = PeerBuffers(links.shape[1], g.num_feat, K, dev, backend=backend)
            if rank == 0:
                print(f"PeerBuffers backend {backend!r} -> {buf.backend}"
                      + (f" ({getattr(buf, '_why_not_multicast', '')})" if buf.backend != 'multicast' and backend == 'auto' else ''), flush=True)
            for o in buf.local:
                o.fill_(float('nan'))
            torch.cuda.synchronize(dev)
            buf.barrier()
            half = (links.shape[1] // dist.get_world_size() + 1) // 2 * (2 if flow == 'SoP' else 1) + 2
            for rep, kw in enumerate((dict(), dict(batch_records=512, overlap=True),
                                      dict(batch_records=half, fronts_first=bool(rank & 1)))):     # the 8-GPU anti-phase schedule
                res, mirror = precompute_exchange(g, links, hops, K, buf, flow=flow, **kw)
                torch.cuda.synchronize(dev)
                want = precompute(g, links, hops, K, flow, pair=False)
                same = all(torch.equal(buf.local[k], want.xs[k]) for k in range(K + 1))
                if not same:
                    bad = [int((buf.local[k] != want.xs[k]).any(1).sum()) for k in range(K + 1)]
                    print(f"rank {rank}: {name} {flow} {buf.backend} rep {rep}: rows differing per operator {bad}", flush=True)
                    ok.zero_()
                elif rank == 0:
                    print(f"{name} {flow} h={hops} K={K} {links.shape[1]} links, world {dist.get_world_size()}, {buf.backend}, rep {rep}: "
                          f"every rank's matrices equal the single-GPU result bit for bit "
                          f"(this rank extracted {res.stats['links'] - res.stats['mirrors']} records)", flush=True)
                buf.barrier()
            buf.close()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('MGPU_OK' if int(ok) else 'MGPU_FAIL', flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(ok) else 1)


if __name__ == '__main__':
    main()
