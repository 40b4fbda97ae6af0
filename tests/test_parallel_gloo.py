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


# ---- the exchange of SURVEY 8e on CPU: cyclic shards + link pairing cover every row exactly once ----
def pair_table_reference(links, num_nodes):
    """NumPy restatement of s3_pair_links' contract (include/s3grl_b200.h): mirror[i] as the CUDA kernel defines it,
    with chain members in ascending order (the kernel's order is scheduling dependent)."""
    L = links.shape[1]
    mirror = np.full(L, -1, dtype=np.int64)
    first, chains = {}, {}
    for i, (u, v) in enumerate(links.T.tolist()):
        if u < 0 or v < 0 or u >= num_nodes or v >= num_nodes or u == v:
            continue
        key = (min(u, v), max(u, v))
        if key in first:
            chains[first[key]].append(i)
        else:
            first[key] = i
            chains[i] = []
    for p, mem in chains.items():
        nxt = -1
        for i in reversed(mem):
            swap = int(links[0, i] != links[0, p])
            mirror[i] = -2 - (((nxt + 1) << 1) | swap)
            nxt = i
        mirror[p] = nxt
    return mirror


def _exchange_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from golden_util import Case
        from oracle import s3grl_oracle as orc
        from s3grl_b200.parallel import cyclic_shard
        c = Case('cora_pos')
        base = c.links[:, :24]
        links = np.concatenate([base, base[::-1][:, ::2], base[:, :3]], axis=1)      # reverses and repeats
        L, K, F1 = links.shape[1], c.K, c.X.shape[1] + 1
        mirror = pair_table_reference(links, c.A.shape[0])
        mine = [i for i in cyclic_shard(L, rank, world) if mirror[i] >= -1]          # members are skipped
        part = orc.pos_precompute(links[:, mine], c.num_hops, c.A, c.X, K, None)
        full = [torch.zeros((2 * L, F1), dtype=torch.float64) for _ in range(K + 1)]
        written = torch.zeros(2 * L, dtype=torch.int64)
        for j, i in enumerate(mine):          # what s3_gather_peers does: own rows, then the chain's rows
            rows = [torch.from_numpy(part['xs'][k][2 * j:2 * j + 2]).double() for k in range(K + 1)]
            m, swap = i, 0
            while True:
                for k in range(K + 1):
                    full[k][2 * m:2 * m + 2] = rows[k].flip(0) if swap else rows[k]
                written[2 * m:2 * m + 2] += 1
                nxt = mirror[i] if m == i else (((-2 - mirror[m]) >> 1) - 1)
                if nxt < 0:
                    break
                swap = (-2 - mirror[nxt]) & 1
                m = nxt
        for t in full + [written]:
            dist.all_reduce(t)                # stands in for the peer stores: every row has exactly one writer
        if rank == 0:
            whole = orc.pos_precompute(links, c.num_hops, c.A, c.X, K, None)
            ok = bool((written == 1).all()) and all(
                np.allclose(full[k].numpy(), whole['xs'][k], rtol=0, atol=1e-6) for k in range(K + 1))
            open(os.path.join(out_dir, 'ok_exchange'), 'w').write('1' if ok else '0')
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_exchange_covers_every_row_once(tmp_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    mp.spawn(_exchange_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / 'ok_exchange').read() == '1'


def test_cyclic_shard_partitions():
    from s3grl_b200.parallel import cyclic_shard
    for L in (0, 1, 7, 164000):
        for world in (1, 2, 8):
            seen = sorted(i for r in range(world) for i in cyclic_shard(L, r, world))
            assert seen == list(range(L))
