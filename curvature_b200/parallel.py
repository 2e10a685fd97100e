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


def invert_plan(dims, world_size: int, align: int = 64):
    """Layer-sharded `KFAC.invert`: which rank inverts which factor, and where every inverse lives.

    `dims[i]` is the order of factor i (2 per layer, in state order).  Factors are assigned greedily, largest first, by
    their D^3 cost (LPT).  The inverse arena is laid out RANK-MAJOR: rank r's matrices are contiguous in segment r, every
    segment padded to the same `segment` floats, so that one in-place all-gather (`all_gather_into_tensor` of the arena
    with each rank's own segment as its input) gives every rank every inverse factor -- the one exchange step.
    Returns dict(owner=[rank per factor], offset=[float offset per factor], segment=floats per rank, total=floats)."""
    count = len(dims)
    order = sorted(range(count), key=lambda i: (-int(dims[i]) ** 3, i))
    load = [0] * world_size
    owner = [0] * count
    for i in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += int(dims[i]) ** 3
    fill = [0] * world_size
    local = [0] * count
    for i in range(count):                      # state order within a segment
        local[i] = fill[owner[i]]
        n = int(dims[i]) ** 2
        fill[owner[i]] += (n + align - 1) // align * align
    segment = max(max(fill), align)
    return {"owner": owner, "offset": [owner[i] * segment + local[i] for i in range(count)], "segment": segment,
            "total": segment * world_size}


def invert_plan_two_rounds(dims, world_size: int, align: int = 64):
    """`invert_plan` with the exchange step overlapped with compute.  A rank's time is dominated by its largest matrix (the
    blocked Cholesky is a sequential chain of panels: a 4608^2 factor takes 10 ms on a whole B200 however little else the
    rank has to do), while the bytes to exchange are dominated by everything else (ResNet-152: 1.6 of 1.9 GB).  So every
    rank inverts its matrices in two rounds -- all but its largest, then its largest -- and the arena has two rank-major
    regions: region 1 (segments of `segment` floats) is all-gathered WHILE round 2 computes, region 2 (one matrix per rank,
    `segment2` floats each, starting at float `base2`) after it.  Same owners as `invert_plan`.
    Returns dict(owner, offset, late=[bool per factor], segment, base2, segment2, total)."""
    base = invert_plan(dims, world_size, align)
    owner = base["owner"]
    count = len(dims)
    late = [False] * count
    for r in range(world_size):
        mine = [i for i in range(count) if owner[i] == r]
        if mine:
            late[max(mine, key=lambda i: (int(dims[i]), -i))] = True
    fill = [0] * world_size
    local = [0] * count
    for i in range(count):
        if not late[i]:
            local[i] = fill[owner[i]]
            n = int(dims[i]) ** 2
            fill[owner[i]] += (n + align - 1) // align * align
    segment = max(max(fill), align)
    segment2 = max([(int(dims[i]) ** 2 + align - 1) // align * align for i in range(count) if late[i]] + [align])
    base2 = segment * world_size
    offset = [(base2 + owner[i] * segment2) if late[i] else (owner[i] * segment + local[i]) for i in range(count)]
    return {"owner": owner, "offset": offset, "late": late, "segment": segment, "base2": base2, "segment2": segment2,
            "total": base2 + segment2 * world_size}


def allgather_segments(flat: torch.Tensor, segment: int, group: Optional["dist.ProcessGroup"] = None):
    """In-place all-gather of a rank-major arena: rank r contributes flat[r * segment : (r + 1) * segment].  One collective."""
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    assert flat.numel() == segment * world
    mine = flat[rank * segment:(rank + 1) * segment]
    if flat.is_cuda:
        return dist.all_gather_into_tensor(flat, mine, group=group)
    # (gloo has no all_gather_into_tensor on every build: gather into views of the same buffer)
    return dist.all_gather([flat[r * segment:(r + 1) * segment] for r in range(world)], mine.clone(), group=group)
