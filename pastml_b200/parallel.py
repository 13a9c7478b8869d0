"""Multi-GPU plumbing: one process per GPU (torch.distributed), independent work items per rank.

The ACR path shards naturally: (tree x column x parameter-set) batches are independent, so a
rank owns a contiguous slice of the work items and no collective touches the data path; only results / timings are
gathered at the end.  ``torch.distributed`` is used for exactly that (NCCL on the GPU box, gloo in CPU tests).
"""


def partition(n_items, world_size, rank):
    """Contiguous, balanced slice [lo, hi) of n_items work items for this rank (the first n_items % world_size ranks
    get one extra item)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError('bad rank / world size')
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_columns(columns, world_size, rank):
    """The columns this rank reconstructs."""
    lo, hi = partition(len(columns), world_size, rank)
    return list(columns[lo:hi])


def gather_objects(local, group=None):
    """All ranks' python objects, in rank order, on every rank (results are small: dictionaries, scalars)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return [local]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local, group=group)
    return out


def max_over_ranks(value, device=None, group=None):
    """Maximum of a scalar over the ranks (how multi-GPU timings are reported)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sharded_acr(run_shard, columns, group=None):
    """Runs ``run_shard(columns_of_this_rank) -> list of results`` on every rank and returns the concatenation of all
    ranks' results in column order (every rank gets the full list)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    mine = shard_columns(list(columns), world, rank)
    local = run_shard(mine) if mine else []
    merged = []
    for part in gather_objects(local, group):
        merged.extend(part)
    return merged
