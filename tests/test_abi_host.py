"""CPU: the C-ABI library loads and exports every symbol include/*.h declares; host-side logic of the
drop-in classes (argument handling, assertion messages, index selection, arena, sharding).  No kernel runs."""
import ctypes
import os
import re

import pytest
import torch

from helpers import ROOT, orc

import curvature_b200 as cb
from curvature_b200 import _native as nat


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "curvature_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(crv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(nat.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 13
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/curvature_b200.h but not exported"
    assert sorted(nat.EXPORTED_SYMBOLS) == declared
    assert nat.ABI_VERSION == 5


def test_workspace_queries_are_host_only():
    assert nat.workspace_bytes(nat.OP_EFB_PROJECT, [10, 20]) == 10 * 20 * 4
    assert nat.workspace_bytes(nat.OP_SAMPLE_MN, [10, 20]) == 10 * 20 * 4 * 2
    assert nat.workspace_bytes(nat.OP_CHOL_INV, [2, 6, 401]) >= 2 * (36 + 401 * 401) * 4
    assert nat.workspace_bytes(nat.OP_SYRK_CONV, [2, 3, 8, 8, 3, 3, 1, 1, 1, 1, 1, nat.PREC_FP32]) == 0


def test_no_cpu_fallback():
    """CPU tensors are refused loudly instead of being routed to some other implementation."""
    model = torch.nn.Sequential(torch.nn.Linear(3, 2))
    kfac = cb.KFAC(model)
    model(torch.randn(4, 3)).sum().backward()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kfac.update(4)
    diag = cb.Diagonal(model)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        diag.update(4)


def test_layer_types_and_messages():
    model = orc.lenet5()
    assert cb.Diagonal(model).layer_types == ['Linear', 'Conv2d', 'MultiheadAttention']
    assert cb.Diagonal(model, []).layer_types == ['Linear', 'Conv2d', 'MultiheadAttention']
    assert cb.Diagonal(model, 'Linear').layer_types == ['Linear']
    assert cb.Diagonal(model, ['Conv2d']).layer_types == ['Conv2d']
    with pytest.raises(TypeError):
        cb.Diagonal(model, 3)
    with pytest.raises(AssertionError):
        cb.Diagonal(model, ['BatchNorm2d'])
    for cls in (cb.Diagonal, cb.KFAC):
        est = cls(model)
        with pytest.raises(AssertionError, match="State dict is empty. Did you call 'update' prior to this\\?"):
            est.invert()
        with pytest.raises(AssertionError, match="Inverse state dict is empty. Did you call 'invert' prior to this\\?"):
            est.sample(model[0])
    kfac = cb.KFAC(model, 'Conv2d')
    assert len(kfac.record) == 2 and len(kfac.hooks) == 4
    assert cb.KFAC._save_grad_output is cb.KFAC._save_output
    mha = torch.nn.Sequential(torch.nn.MultiheadAttention(8, 2))
    with pytest.raises(NotImplementedError):
        cb.KFAC(mha)
    with pytest.raises(NotImplementedError):
        cb.KFAC(torch.nn.Sequential(torch.nn.Conv2d(4, 4, 3, groups=2)))


def test_hooks_are_pointer_stashes():
    model = orc.lenet5()
    kfac = cb.KFAC(model)
    x = torch.rand(3, 1, 28, 28)
    out = model(x)
    assert kfac.record[model[0]][0] is x                       # reference: no copy (curvatures.py:307)
    out.sum().backward()
    g = kfac.record[model[11]][1]
    assert g.shape == (3, 10)
    assert torch.equal(kfac.scaled_record(model[11]), g * 3)   # what the reference stores (curvatures.py:310)
    with torch.no_grad():
        model(x)                                               # eval-style forward still only stashes


def test_dim_reduction_matches_oracle():
    torch.manual_seed(3)
    qa, qg = torch.randn(12, 12), torch.randn(7, 7)
    lam = torch.randn(84).abs()
    for rank in (5, 20, 83, 84, 200):
        a, b, c = cb.INF._dim_reduction(qa, qg, lam, rank)
        ao, bo, co = orc.INF._dim_reduction(qa, qg, lam, rank)
        assert torch.equal(a, ao) and torch.equal(b, bo) and torch.equal(c, co)


def test_kron_and_eigen_helpers():
    a = torch.tensor([[1, 2], [3, 4]])
    b = torch.tensor([[0, 5], [6, 7]])
    assert torch.equal(cb.kron(a, b), orc.kron(a, b))
    torch.manual_seed(0)
    A = torch.randn(9, 9); A = A @ A.t()
    G = torch.randn(5, 5); G = G @ G.t()
    ev = cb.get_eigenvectors({"l": (A, G)})["l"]
    evo = orc.eigenvectors_of_factors({"l": (A, G)})["l"]
    for q, qo, Fm in ((ev[0], evo[0], A), (ev[1], evo[1], G)):
        assert q.is_contiguous()
        assert torch.allclose(q.abs(), qo.abs(), atol=1e-5)
        assert torch.allclose(q.t() @ (2 * Fm) @ q, torch.diag(torch.diag(q.t() @ (2 * Fm) @ q)), atol=1e-3)
    vals = cb.get_eigenvalues([(A, G), torch.ones(3, 2)])
    assert vals.numel() == 45 + 6
    assert torch.allclose(vals, orc.eigenvalues_of_factors([(A, G), torch.ones(3, 2)]), rtol=1e-4, atol=1e-4)


def test_arena_views_alias_flat_buffer():
    arena = cb.FactorArena([(3, 3), (5, 2), (1,)], "cpu")
    assert arena.flat.numel() % cb.FactorArena.ALIGN == 0
    arena.views[1].fill_(2.0)
    assert arena.flat.sum().item() == 20.0
    for v in arena.views:
        assert v.data_ptr() % 256 == arena.flat.data_ptr() % 256


def test_shard_indices():
    assert cb.shard_indices(5, 0, 2) == [0, 2, 4]
    assert cb.shard_indices(5, 1, 2) == [1, 3]
    costs = [100, 1, 1, 1, 50, 50]
    parts = [cb.shard_indices(6, r, 2, costs) for r in range(2)]
    assert sorted(parts[0] + parts[1]) == list(range(6))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) <= 103


def test_factor_file_round_trip_is_name_keyed_and_pickle_free(tmp_path):
    """save_factors / load_factors: plain tensors + layer names, loadable with weights_only=True into
    a fresh estimator of another model instance; mismatching layers are refused; accumulate=True merges shards."""
    import torch
    import curvature_b200 as cb

    def make():
        torch.manual_seed(0)
        return torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, bias=True), torch.nn.ReLU(), torch.nn.Flatten(),
                                   torch.nn.Linear(4 * 6 * 6, 5, bias=False))
    for cls in (cb.KFAC, cb.Diagonal):
        a, b = cls(make()), cls(make())
        a._ensure_arena()
        a.arena.flat.copy_(torch.arange(a.arena.flat.numel(), dtype=torch.float32))
        for key, views in a._views.items():
            a.state[key] = views
        path = str(tmp_path / f"{cls.__name__}.pt")
        cb.save_factors(a, path)
        blob = torch.load(path, weights_only=True)           # no pickled modules inside
        assert blob["format"] == "curvature_b200.factors.v1" and [e["layer"] for e in blob["entries"]] == ["0", "3"]
        cb.load_factors(b, path)
        assert torch.equal(a.arena.flat, b.arena.flat) and len(b.state) == 2
        for (ka, va), (kb, vb) in zip(a.state.items(), b.state.items()):
            va = va if isinstance(va, (list, tuple)) else [va]
            vb = vb if isinstance(vb, (list, tuple)) else [vb]
            assert all(torch.equal(x, y) for x, y in zip(va, vb))
        cb.load_factors(b, path, accumulate=True)
        assert torch.equal(b.arena.flat, 2 * a.arena.flat)
        other = cls(torch.nn.Sequential(torch.nn.Linear(7, 5)))
        with pytest.raises(ValueError):
            cb.load_factors(other, path)
    with pytest.raises(ValueError):
        cb.load_factors(cb.Diagonal(make()), str(tmp_path / "KFAC.pt"))


def _resnet50_items(batch=256):
    """(N, C, H, W, kh, kw, sh, sw, ph, pw) of every channels-last factor operand of a ResNet-50 KFAC.update (fc excluded:
    its A factor has a bias row), from a 1-image CPU forward (shapes only)."""
    import torch
    import torchvision
    model = torchvision.models.resnet50(weights=None).eval()
    shapes = {}
    hooks = [m.register_forward_hook(lambda mod, inp, out: shapes.__setitem__(mod, (inp[0].shape, out.shape)))
             for m in model.modules() if isinstance(m, torch.nn.Conv2d)]
    with torch.no_grad():
        model(torch.zeros(1, 3, 224, 224))
    for h in hooks:
        h.remove()
    items = []
    for m in model.modules():
        if isinstance(m, torch.nn.Conv2d):
            (_, C, H, W), (_, M, OH, OW) = shapes[m]
            items.append((batch, C, H, W, *m.kernel_size, *m.stride, *m.padding))      # A factor
            items.append((batch, M, 1, OH * OW, 1, 1, 1, 1, 0, 0))                      # G factor (rows operand)
    return items


def test_stream_k_partition_of_a_resnet50_update_is_a_partition():
    """Host logic of K1e (no GPU): crv_debug_partition exposes how the 106 channels-last factors of a ResNet-50 update are
    cut into launches and each launch's work list into CTA ranges.  Invariants: re-read operands get a launch each and
    read-once ones ride in groups of <= 48; boundaries start at (0, 0), end at (pairs, 0), are strictly increasing, lie
    inside their pair on whole pipeline stages, and there are at most `sms` CTAs."""
    items = _resnet50_items()
    assert len(items) == 106
    first = nat.debug_partition(items, nat.PREC_BF16, sms=148, which=0)
    nl = first["n_launches"]
    sizes = [first["launch_of_item"].count(l) for l in range(nl)]
    assert sum(sizes) == len(items) and max(sizes) <= 48
    assert sum(1 for s in sizes if s > 1) >= 2 and sizes.count(1) == 37          # two groups of read-once factors, 37 re-read operands with a launch each
    for which in range(nl):
        part = nat.debug_partition(items, nat.PREC_BF16, sms=148, which=which)
        bd, nbox, nb = part["boundaries"], part["nbox"], part["nb"]
        pairs = len(nbox)
        assert bd[0] == (0, 0) and bd[-1] == (pairs, 0) and len(bd) - 1 <= 148
        assert all(x < y for x, y in zip(bd, bd[1:])), which
        for q, b in bd[:-1]:
            assert 0 <= q < pairs and 0 <= b < nbox[q] and b % nb[q] == 0, (which, q, b)
    # fewer SMs: still a partition with at most that many CTAs
    part = nat.debug_partition(items, nat.PREC_BF16, sms=20, which=nl - 1)
    assert len(part["boundaries"]) - 1 <= 20 and part["boundaries"][-1] == (len(part["nbox"]), 0)
