"""GPU (-m gpu): S stacked posterior draws (K5c, `KFAC.sample_many`) against `KFAC.sample` layer by layer with the same
noise, `replace_with` against `sample_and_replace`, and the evaluation loop `eval_bnn` (scripts/evaluate.py:121-152)."""
import numpy as np
import pytest
import torch

from helpers import orc, rel_fro

pytestmark = pytest.mark.gpu

import curvature_b200 as cb                      # noqa: E402

DEV = "cuda:0"


def small_net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 64, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(64, 128, 3, padding=1, bias=False),
                               torch.nn.ReLU(), torch.nn.AdaptiveAvgPool2d(2), torch.nn.Flatten(),
                               torch.nn.Linear(512, 10)).to(DEV)


@pytest.mark.parametrize("prec", ["tf32", "bf16x3"])
def test_sample_many_equals_single_draws_and_eval_bnn(prec):
    model = small_net()
    kfac = cb.KFAC(model, precision=prec)
    x = torch.randn(16, 3, 8, 8, device=DEV)
    orc.fisher_step(model, x)
    kfac.update(16)
    kfac.invert(1.0, 10.0)
    S = 5
    gen = torch.Generator(device=DEV).manual_seed(1)
    noise = {l: torch.randn(S, la.shape[0], lg.shape[0], device=DEV, generator=gen) for l, (la, lg) in kfac.inv_state.items()}
    draws = kfac.sample_many(S, noise=noise)
    tol = 1e-3 if prec == "tf32" else 1e-5
    for layer, d in draws.items():
        assert tuple(d.shape) == (S, kfac.inv_state[layer][1].shape[0], kfac.inv_state[layer][0].shape[0])
        for s in range(S):
            one = kfac.sample(layer, noise[layer][s].contiguous())
            assert rel_fro(d[s], one) <= tol, (str(layer), s, rel_fro(d[s], one))
            LA, LG = (t.double() for t in kfac.inv_state[layer])
            want = (LA @ noise[layer][s].double() @ LG.t()).t()
            assert rel_fro(d[s], want) <= max(tol, 1e-5)
    # replace_with(draws, s) leaves the model as sample_and_replace with the same noise does
    kfac.replace_with(draws, 2)
    got = {k: v.clone() for k, v in model.state_dict().items()}
    kfac.sample_and_replace(noise={l: noise[l][2].contiguous() for l in noise})
    for k, v in model.state_dict().items():
        assert rel_fro(got[k], v) <= tol, k
    # evaluation loop: ensemble mean of S sampled networks; statistics lists like the reference's
    data = [(torch.randn(12, 3, 8, 8), torch.randint(0, 10, (12,))) for _ in range(2)]
    torch.manual_seed(3)
    mean_pred, labels, stats = cb.eval_bnn(model, data, kfac, samples=4, stats=True, device=torch.device(DEV), verbose=False)
    assert mean_pred.shape == (24, 10) and labels.shape == (24,)
    assert np.allclose(mean_pred.sum(1), 1.0, atol=1e-5)
    assert all(len(v) == 4 for v in stats.values())
    # the generic path (an estimator without sample_many) gives a valid ensemble too
    diag = cb.Diagonal(model)
    orc.fisher_step(model, x)
    diag.update(16)
    diag.invert(1.0, 10.0)
    p2, l2, _ = cb.eval_bnn(model, data, diag, samples=3, device=torch.device(DEV), verbose=False)
    assert p2.shape == (24, 10) and np.array_equal(l2, labels)
    for h in kfac.hooks:
        h.remove()


def test_eigen_spectrum_tooling_shares_the_one_shot_eigensolve():
    """`get_eigenvalues` (utils.py:21-42) and `get_eigenvectors` (utils.py:45-60) on device factors
    against the oracle; `eigendecompose` gives both from ONE eigh per factor."""
    model = small_net()
    kfac = cb.KFAC(model)
    orc.fisher_step(model, torch.randn(16, 3, 8, 8, device=DEV))
    kfac.update(16)
    factors = list(kfac.state.values())
    host = [[f.cpu() for f in fs] for fs in factors]
    want = orc.eigenvalues_of_factors(host)
    got = cb.get_eigenvalues(factors)
    assert got.is_cuda and rel_fro(got, want) <= 1e-4
    eigvecs, eigvals = cb.eigendecompose(kfac.state)
    shared = cb.get_eigenvalues(factors, eigvals=list(eigvals.values()))
    assert rel_fro(shared, want) <= 1e-4
    for layer, (qa, qg) in eigvecs.items():
        A, G = kfac.state[layer]
        wa, wg = eigvals[layer]
        assert rel_fro(qa @ torch.diag(wa) @ qa.t(), A) <= 1e-4 and rel_fro(qg @ torch.diag(wg) @ qg.t(), G) <= 1e-4
    ref = cb.get_eigenvectors(kfac.state)
    for layer in ref:                       # same subspaces (eigenvectors are defined up to sign / rotation)
        assert rel_fro(ref[layer][0] @ ref[layer][0].t(), eigvecs[layer][0] @ eigvecs[layer][0].t()) <= 1e-4
    for h in kfac.hooks:
        h.remove()
