"""GPU, world_size 2, NCCL: per-rank KFAC/Diagonal accumulation through the CUDA kernels, ONE all-reduce of the flat
arena, result equal to the oracle iterating over both shards as consecutive batches.  Skipped on
boxes with fewer than two GPUs (the CPU/gloo twin of this test is tests/test_distributed_cpu.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import orc, rel_fro

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import curvature_b200 as cb
    torch.backends.cudnn.allow_tf32 = False
    dev = f"cuda:{rank}"
    torch.manual_seed(0)
    model = orc.lenet5().to(dev)
    gen = torch.Generator().manual_seed(1000 + rank)
    x = torch.rand(16, 1, 28, 28, generator=gen)
    labels = torch.randint(0, 10, (16,), generator=gen)
    kfac, diag = cb.KFAC(model, precision="tf32"), cb.Diagonal(model)
    orc.fisher_step(model, x.to(dev), labels=labels.to(dev))
    kfac.update(16)
    diag.update(16)
    cb.allreduce_arena(kfac)          # exactly one collective per estimator
    cb.allreduce_arena(diag)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"kfac": [[f.cpu() for f in v] for v in kfac.state.values()],
                    "diag": [v.cpu() for v in diag.state.values()]}, os.path.join(out_dir, "reduced.pt"))
    # layer-sharded invert + one all-gather == every rank inverting everything
    kfac.invert(0.5, 1.0)                                   # sharded: torch.distributed is initialised, world 2
    sharded = [[t.clone() for t in v] for v in kfac.inv_state.values()]
    import curvature_b200.curvatures as cv
    old_floats, cv._TWO_ROUND_FLOATS = cv._TWO_ROUND_FLOATS, 0          # ... and the two-round plan (exchange overlapped
    kfac.invert(0.5, 1.0)                                               # with every rank's largest matrix) on this model
    cv._TWO_ROUND_FLOATS = old_floats
    two_round = [[t.clone() for t in v] for v in kfac.inv_state.values()]
    kfac.invert(0.5, 1.0, shard=False)
    for a, c, b in zip(sharded, two_round, kfac.inv_state.values()):
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
        assert torch.equal(c[0], b[0]) and torch.equal(c[1], b[1])
    kfac.sample_and_replace()                               # posterior samples: per-rank RNG, no communication
    # a model on a device that is not the current one (the C ABI switches to the operand's device itself)
    other = (rank + 1) % world
    torch.cuda.set_device(other)
    k2 = cb.KFAC(model)                                     # model lives on `dev`, current device is the other GPU
    k2.record = kfac.record
    k2.update(16)
    k2.invert(0.5, 1.0, shard=False)
    torch.cuda.synchronize(dev)
    torch.cuda.set_device(rank)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_arena_allreduce_equals_sequential_reference(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    red = torch.load(tmp_path / "reduced.pt")
    torch.manual_seed(0)
    model = orc.lenet5()
    kfac, diag = orc.KFAC(model), orc.Diagonal(model)
    for rank in range(2):
        gen = torch.Generator().manual_seed(1000 + rank)
        x = torch.rand(16, 1, 28, 28, generator=gen)
        labels = torch.randint(0, 10, (16,), generator=gen)
        orc.fisher_step(model, x, labels=labels)
        kfac.update(16)
        diag.update(16)
    for got, want in zip(red["kfac"], kfac.state.values()):
        assert rel_fro(got[0], want[0]) <= 1e-3 and rel_fro(got[1], want[1]) <= 1e-3   # tf32 tier
    for got, want in zip(red["diag"], diag.state.values()):
        assert rel_fro(got, want) <= 1e-5
