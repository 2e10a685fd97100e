"""CPU, world_size 2, gloo: the multi-process merge of locally accumulated state (one all-reduce of the
flat arena) equals the reference iterating over both shards as consecutive batches."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import orc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import curvature_b200 as cb

    class Fake:      # stands in for an estimator whose arena was filled by the CUDA kernels
        arena = cb.FactorArena([(4, 4), (3, 3)], "cpu")
    torch.manual_seed(100 + rank)
    Fake.arena.views[0].copy_(torch.randn(4, 4))
    Fake.arena.views[1].copy_(torch.randn(3, 3))
    local = Fake.arena.flat.clone()
    cb.allreduce_arena(Fake)
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(Fake.arena.flat, sum(gathered))
    # views still alias the reduced buffer
    assert torch.equal(Fake.arena.views[1].reshape(-1), Fake.arena.flat[64:73])
    # sample / layer sharding is communication-free and covers everything exactly once
    mine = cb.shard_indices(7, rank, world)
    all_idx = [None] * world
    dist.all_gather_object(all_idx, mine)
    assert sorted(sum(all_idx, [])) == list(range(7))
    torch.save(Fake.arena.flat, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_gloo_allreduce_arena(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(tmp_path / "r0.pt")
    b = torch.load(tmp_path / "r1.pt")
    assert torch.equal(a, b)


def test_shard_sum_equals_consecutive_batches():
    """The identity the multi-GPU design rests on, checked with the oracle on CPU:
    factors(shard0) + factors(shard1) == reference run over shard0 then shard1."""
    torch.manual_seed(0)
    model = orc.lenet5()
    xs = [torch.rand(8, 1, 28, 28), torch.rand(8, 1, 28, 28)]
    labs = [torch.randint(0, 10, (8,)), torch.randint(0, 10, (8,))]
    seq = orc.KFAC(model)
    for x, l in zip(xs, labs):
        orc.fisher_step(model, x, labels=l)
        seq.update(8)
    parts = []
    for x, l in zip(xs, labs):
        k = orc.KFAC(model)
        orc.fisher_step(model, x, labels=l)
        k.update(8)
        parts.append(k)
        for h in k.hooks:
            h.remove()
    for layer in seq.state:
        for f in range(2):
            assert torch.allclose(seq.state[layer][f], parts[0].state[layer][f] + parts[1].state[layer][f],
                                  rtol=1e-6, atol=1e-8)


def _invert_worker(rank, world, port, out_dir):
    """Layer-sharded invert on two ranks (gloo): every rank fills ONLY the inverse factors the plan assigns to it (here
    with the oracle's curvatures.py:368-379 restatement standing in for the CUDA kernel), one all-gather of the
    rank-major arena, and every rank ends up with every inverse factor."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import curvature_b200 as cb
    torch.manual_seed(0)                                   # same (already merged) factors on every rank
    dims = [26, 6, 151, 16, 401, 120, 121, 84, 85, 10]     # LeNet-5's factor orders
    factors = []
    for D in dims:
        X = torch.randn(D, 2 * D + 3)
        factors.append(X @ X.t() / X.shape[1])
    plan = cb.invert_plan(dims, world)
    flat = torch.zeros(plan["total"])
    views = [flat[o:o + D * D].view(D, D) for o, D in zip(plan["offset"], dims)]
    for i, owner in enumerate(plan["owner"]):
        if owner == rank:
            views[i].copy_(orc.damped_inverse_cholesky(factors[i], 0.5, 1.0))
    cb.allgather_segments(flat, plan["segment"])
    for i, D in enumerate(dims):
        want = orc.damped_inverse_cholesky(factors[i], 0.5, 1.0)
        assert torch.equal(views[i], want), (rank, i)
    torch.save(flat, os.path.join(out_dir, f"inv{rank}.pt"))
    dist.destroy_process_group()


def test_gloo_sharded_invert_allgather(tmp_path):
    port = _free_port()
    mp.spawn(_invert_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert torch.equal(torch.load(tmp_path / "inv0.pt"), torch.load(tmp_path / "inv1.pt"))


def test_invert_plan_is_a_balanced_partition():
    import curvature_b200 as cb
    # ResNet-152's factor orders (SURVEY 8(a) multiplicities): 312 matrices
    shapes = [(147, 64)] + [(64, 64)] + [(576, 64)] * 3 + [(64, 256)] * 4 + [(256, 64)] * 2 + [(256, 128)] + [(1152, 128)] * 8 + \
             [(128, 512)] * 8 + [(256, 512)] + [(512, 128)] * 7 + [(512, 256)] + [(2304, 256)] * 36 + [(256, 1024)] * 36 + \
             [(512, 1024)] + [(1024, 256)] * 35 + [(1024, 512)] + [(4608, 512)] * 3 + [(512, 2048)] * 3 + [(1024, 2048)] + \
             [(2048, 512)] * 2 + [(2049, 1000)]
    dims = [d for kd in shapes for d in kd]
    for world in (1, 2, 4, 8):
        plan = cb.invert_plan(dims, world)
        assert len(plan["owner"]) == len(dims) and set(plan["owner"]) == set(range(world))
        spans = sorted((o, o + d * d) for o, d in zip(plan["offset"], dims))
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))                       # no two inverses overlap
        for i, (o, d) in enumerate(zip(plan["offset"], dims)):                           # each inside its owner's segment
            r = plan["owner"][i]
            assert r * plan["segment"] <= o and o + d * d <= (r + 1) * plan["segment"]
        load = [sum(d ** 3 for d, r in zip(dims, plan["owner"]) if r == q) for q in range(world)]
        assert max(load) <= 1.1 * sum(load) / world, (world, load)                       # LPT by D^3: within 10 % of even
        # two rounds (exchange overlapped with every rank's largest matrix): same owners, two rank-major regions
        p2 = cb.invert_plan_two_rounds(dims, world)
        assert p2["owner"] == plan["owner"] and sum(p2["late"]) == world
        spans = sorted((o, o + d * d) for o, d in zip(p2["offset"], dims))
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= p2["total"]
        for i, (o, d) in enumerate(zip(p2["offset"], dims)):
            r = p2["owner"][i]
            if p2["late"][i]:
                assert d == max(dd for dd, rr in zip(dims, p2["owner"]) if rr == r)
                assert p2["base2"] + r * p2["segment2"] <= o and o + d * d <= p2["base2"] + (r + 1) * p2["segment2"]
            else:
                assert r * p2["segment"] <= o and o + d * d <= (r + 1) * p2["segment"] <= p2["base2"]
