"""Page-level data parallelism: the model replicates, pages shard.

One process per GPU; page i of a job goes to rank ``i % world`` (equal-size synthetic pages) or, for
mixed-size pages, longest-first greedy by ViT token count.  There is no collective on the data path:
generated ids return to rank 0 with one ``gather_object`` after the job (a few KB per page).
"""
from __future__ import annotations

from typing import List, Optional, Sequence


def shard_round_robin(n_pages: int, world: int, rank: int) -> List[int]:
    return list(range(rank, n_pages, world))


def shard_by_cost(costs: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first greedy: returns the page indices of every rank."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += costs[i]
    return [sorted(x) for x in out]


def gather_pages(local_indices: Sequence[int], local_results: Sequence, n_pages: int, group=None) -> Optional[list]:
    """Collect per-page results on rank 0 in page order (others get None).  Works with gloo or nccl."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = [None] * n_pages
        for i, r in zip(local_indices, local_results):
            out[i] = r
        return out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bucket = [None] * world if rank == 0 else None
    dist.gather_object((list(local_indices), list(local_results)), bucket, dst=0, group=group)
    if rank != 0:
        return None
    out = [None] * n_pages
    for idx, res in bucket:
        for i, r in zip(idx, res):
            out[i] = r
    assert all(o is not None for o in out), "a page was not processed by any rank"
    return out
