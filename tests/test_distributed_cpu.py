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
