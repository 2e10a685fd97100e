"""Data-parallel merge of locally accumulated curvature state.

One process per GPU accumulates factor sums over its own batches; a single all-reduce(SUM) of the flat
arena per estimation pass merges them (torch.distributed, NCCL over NVLink on GPUs, gloo in the CPU tests).
Summing over ranks is exactly the reference iterating over the same shards as consecutive batches
(curvatures.py:346-350 is a plain running sum), for KFAC, Diagonal and EFB alike.  The reference itself has no
multi-GPU path (its DataParallel wrapper breaks KFAC's module-keyed hooks).
"""
from typing import Iterable, Optional

import torch
import torch.distributed as dist


def allreduce_arena(estimator, group: Optional["dist.ProcessGroup"] = None, async_op: bool = False):
    """Sum the estimator's flat state arena over all ranks (in place).  Exactly one collective."""
    if estimator.arena is None:
        raise RuntimeError("nothing to reduce: call 'update' first")
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    return dist.all_reduce(estimator.arena.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def shard_indices(count: int, rank: int, world_size: int, costs: Optional[Iterable[float]] = None):
    """Indices of the work items (layers to invert, posterior samples to draw) owned by `rank`.
    With `costs` (e.g. D^3 per matrix) items are balanced greedily, largest first (LPT); otherwise
    round-robin.  No communication is involved."""
    if costs is None:
        return list(range(rank, count, world_size))
    costs = list(costs)
    order = sorted(range(count), key=lambda i: -costs[i])
    load = [0.0] * world_size
    owner = [0] * count
    for i in order:
        r = min(range(world_size), key=lambda q: load[q])
        owner[i] = r
        load[r] += costs[i]
    return [i for i in range(count) if owner[i] == rank]
