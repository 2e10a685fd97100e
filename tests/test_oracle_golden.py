"""CPU: the oracle restatement against the fixtures the REAL reference produced
(tests/golden/make_golden.py), plus the reference's one known-answer test (kron doctest)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import orc, rel_fro, model_from_golden, selected_layers, n_batches

TOL = 1e-6   # restatement vs reference fixtures: same fp32 op sequence, BLAS threading may reorder sums


def test_kron_known_answer():
    # curvature/utils.py:301-308 -- the only known-answer test in the reference
    a = torch.tensor([[1, 2], [3, 4]])
    b = torch.tensor([[0, 5], [6, 7]])
    want = torch.tensor([[0, 5, 0, 10], [6, 7, 12, 14], [0, 15, 0, 20], [18, 21, 24, 28]])
    assert torch.equal(orc.kron(a, b), want)


@pytest.mark.parametrize("C,H,W,k,p,s", [
    (1, 5, 5, (3, 3), (1, 1), (1, 1)), (3, 7, 6, (3, 2), (1, 0), (2, 1)), (2, 8, 8, (1, 1), (0, 0), (2, 2)),
    (4, 9, 11, (5, 5), (2, 2), (1, 1)), (3, 12, 12, (7, 7), (3, 3), (2, 2)), (2, 6, 9, (1, 3), (0, 2), (1, 2)),
    (5, 4, 4, (4, 4), (0, 0), (1, 1)), (1, 10, 3, (3, 3), (2, 2), (3, 1))])
def test_im2col_index_map_bit_exact(C, H, W, k, p, s):
    """The restated index map reproduces F.unfold (what the reference calls, curvatures.py:329) bit for bit."""
    x = torch.arange(2 * C * H * W, dtype=torch.float32).reshape(2, C, H, W) + 1.0
    want = F.unfold(x, k, padding=p, stride=s)
    got = orc.unfold_patches(x, k, p, s)
    assert torch.equal(got, want)


@pytest.mark.parametrize("name", ["convzoo", "lenet5"])
def test_oracle_matches_reference_fixtures(name, golden):
    g = golden(name)
    torch.manual_seed(0)
    model = model_from_golden(name, g)
    layers = selected_layers(model)
    N = int(g["meta/batch"])
    kfac, diag = orc.KFAC(model), orc.Diagonal(model)
    for b in range(n_batches(g)):
        orc.fisher_step(model, torch.from_numpy(g[f"x/{b}"]), labels=torch.from_numpy(g[f"labels/{b}"]))
        kfac.update(N)
        diag.update(N)
    for li, l in enumerate(layers):
        assert rel_fro(kfac.state[l][0], g[f"kfac_A/{li}"]) <= TOL
        assert rel_fro(kfac.state[l][1], g[f"kfac_G/{li}"]) <= TOL
        assert rel_fro(diag.state[l], g[f"diag/{li}"]) <= 1e-5   # via conv weight-grad: summation order is thread-count dependent
    eig = {l: (torch.from_numpy(g[f"eig_QA/{li}"]), torch.from_numpy(g[f"eig_QG/{li}"])) for li, l in enumerate(layers)}
    efb = orc.EFB(model, kfac.state, eigvecs=eig)
    for b in range(n_batches(g)):
        orc.fisher_step(model, torch.from_numpy(g[f"x/{b}"]), labels=torch.from_numpy(g[f"labels/{b}"]))
        efb.update(N)
    for li, l in enumerate(layers):
        assert rel_fro(efb.state[l], g[f"efb_lambda/{li}"]) <= 1e-5
    inf = orc.INF(model, diag.state, kfac.state, efb.state, eigvecs=eig)
    inf.update(rank=int(g["meta/rank"]))
    for li, l in enumerate(layers):
        assert inf.state[l][0].shape == g[f"inf_state_lrQA/{li}"].shape
        assert inf.state[l][1].shape == g[f"inf_state_lrQG/{li}"].shape
        assert rel_fro(inf.state[l][2], g[f"inf_state_lrlambda/{li}"]) <= 1e-5
        assert rel_fro(inf.state[l][3], g[f"inf_state_correction/{li}"]) <= 1e-4
    kfac.invert(*g["meta/kfac_damp"].tolist())
    efb.invert(*g["meta/diag_damp"].tolist())
    for li, l in enumerate(layers):
        assert rel_fro(kfac.inv_state[l][0], g[f"kfac_LA/{li}"]) <= 1e-4
        assert rel_fro(kfac.inv_state[l][1], g[f"kfac_LG/{li}"]) <= 1e-4
        z = torch.from_numpy(g[f"noise_KM/{li}"])
        assert rel_fro(kfac.sample(l, z), g[f"kfac_sample/{li}"]) <= 1e-4
        assert rel_fro(efb.sample(l, z), g[f"efb_sample/{li}"]) <= 1e-4


def test_flip_cholesky_identity(golden):
    """chol_lower(inv(reg)) == (J chol(J reg J) J)^{-T}; the identity K4 is built on."""
    g = golden("convzoo")
    for li in range(int(g["meta/n_layers"])):
        for key, Lkey in (("kfac_A", "kfac_LA"), ("kfac_G", "kfac_LG")):
            Fm = torch.from_numpy(g[f"{key}/{li}"]).double()
            n, s = g["meta/kfac_damp"].tolist()
            reg = s ** 0.5 * Fm + n ** 0.5 * torch.eye(Fm.shape[0], dtype=torch.float64)
            reg = (reg + reg.t()) / 2
            P = reg.flip(0).flip(1)
            C = torch.linalg.cholesky(P)
            Cinv = torch.linalg.inv(C)
            L = Cinv.t().flip(0).flip(1)
            assert rel_fro(L, g[f"{Lkey}/{li}"]) <= 1e-5


REF_WEIGHTS = "/root/reference/curvature/lenet5_mnist.pth"

# known answers of the REAL reference on its bundled LeNet-5 weights (BASELINE configs 1 / 2: three
# batches of 100, sampled-label Fisher, KFAC.invert(0.5, 1)).  Columns: tr A, |A|_F, tr G, |G|_F, sum Diagonal.state,
# tr L_A, tr L_G.
KNOWN_ANSWERS = [
    (25.935345, 19.379271, 1.258251e-3, 6.356086e-4, 11.94352, 25.098049, 7.134185),
    (558.588440, 521.087219, 5.661512e-3, 2.574870e-3, 188.7346, 157.176849, 19.022558),
    (28.024693, 9.075549, 2.322013, 0.6647527, 20.07698, 462.779999, 140.977631),
    (5.670164, 4.185143, 2.926824, 0.9636346, 4.899538, 141.880829, 97.875572),
    (5.301585, 4.446790, 2.671096, 0.9185417, 4.128474, 99.562302, 10.173796)]


@pytest.mark.skipif(not __import__("os").path.exists(REF_WEIGHTS),
                    reason="the reference tree (its bundled lenet5_mnist.pth is reference data, not copied here) is absent")
def test_oracle_reproduces_reference_known_answers_on_bundled_weights():
    """BASELINE config 1 proper: the reference's own weights (read in place from the reference tree, build container
    only), the recipe of SURVEY 8(c), every number of its table to 1e-5 relative."""
    from torch.distributions import Categorical
    torch.manual_seed(0)
    model = orc.lenet5()
    model.load_state_dict(torch.load(REF_WEIGHTS, weights_only=True, map_location="cpu"))
    kfac, diag = orc.KFAC(model), orc.Diagonal(model)
    gen = torch.Generator().manual_seed(123)
    crit = torch.nn.CrossEntropyLoss()
    for _ in range(3):
        x = torch.rand(100, 1, 28, 28, generator=gen)
        logits = model(x)
        labels = Categorical(logits=logits).sample()
        loss = crit(logits, labels)
        model.zero_grad()
        loss.backward()
        kfac.update(100)
        diag.update(100)
    assert labels[:5].tolist() == [9, 7, 8, 4, 0] and abs(loss.item() - 2.260454) < 1e-5
    kfac.invert(add=0.5, multiply=1)
    for layer, want in zip(kfac.state, KNOWN_ANSWERS):
        A, G = kfac.state[layer]
        LA, LG = kfac.inv_state[layer]
        got = (A.trace(), A.norm(), G.trace(), G.norm(), diag.state[layer].sum(), LA.trace(), LG.trace())
        assert A[-1, -1].item() == 3.0          # ones row x 3 summed updates: plain running sum
        for v, w in zip(got, want):
            assert abs(v.item() - w) <= 1e-5 * abs(w), (str(layer), v.item(), w)
