"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding + allgather plumbing.  Each
rank 'precomputes' its shard with the oracle (a GPU is not available here; on the GPU box the
same functions run over NCCL with the CUDA path, see bench.py --gpus N) and the gathered
result must equal the single-process result bit for bit, for fixed-size (PoS) and ragged
(PoS Plus) shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from s3grl_b200.parallel import shard_range


def test_shard_range_partitions_in_order():
    for L in (0, 1, 7, 8, 164000, 164001):
        for world in (1, 2, 3, 8):
            got = [shard_range(L, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == L
            assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, strategy, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from golden_util import Case
        from oracle import s3grl_oracle as orc
        from s3grl_b200.parallel import allgather_rows
        c = Case('cora_posplus')
        links = c.links[:, :51]                       # odd count: ragged shards
        a, b = shard_range(links.shape[1], rank, world)
        part = orc.pos_precompute(links[:, a:b], c.num_hops, c.A, c.X, c.K, strategy)
        xs = [torch.from_numpy(x) for x in part['xs']]
        xs_full, rp_full = allgather_rows(xs, torch.from_numpy(part['row_ptr']))
        if rank == 0:
            whole = orc.pos_precompute(links, c.num_hops, c.A, c.X, c.K, strategy)
            ok = np.array_equal(rp_full.numpy(), whole['row_ptr']) and all(
                np.array_equal(g.numpy(), w) for g, w in zip(xs_full, whole['xs']))
            open(os.path.join(out_dir, f'ok_{strategy}'), 'w').write('1' if ok else '0')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('strategy', [None, 'intersection'])
def test_two_rank_gloo_allgather_matches_single_process(strategy, tmp_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    mp.spawn(_worker, args=(2, _free_port(), strategy, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / f'ok_{strategy}').read() == '1'
