"""Multi-GPU plumbing: one process per GPU, target links sharded in contiguous blocks, graph
replicated, and ONE collective — an allgather of the precomputed operator rows so every rank
holds the whole joint matrix for training.  The reference has no distributed
code at all; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is plumbing only.
"""
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
