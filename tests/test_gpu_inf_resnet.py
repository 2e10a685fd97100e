"""GPU (-m gpu): INF (curvatures.py:463-672) at ResNet scale -- SURVEY 8(f) rank 1.  The reference materialises
kron(QA, QG) (tens of GB on a ResNet, it falls back to the CPU at :557-563); here every Kronecker product is the
equivalent pair of small GEMMs, so the whole Diagonal -> KFAC -> EFB -> INF(rank 100) -> invert -> sample sequence runs on
the device for all 21 layers of ResNet-18, including the 4608 x 512 ones.

Checks:
  * layers small enough for the oracle's literal kron formulation on the host (stem, 64-channel 1x1 / 3x3 blocks): state,
    inverse state and a same-noise sample against the oracle in fp64 (the reference's own fp32 chain is ill-conditioned,
    see test_gpu_parity.py::test_inf_matches_reference_fixtures: 2e-3 or 3x the reference's own fp32 error);
  * the largest layers: `_diagonal_accumulator` and the V^T V of `pre_sampler` against direct fp64 contractions on the
    device (no kron on either side);
  * every layer: finite samples of the right shape, `sample_and_replace` runs.
"""
import time

import pytest
import torch

from helpers import orc, rel_fro, selected_layers

pytestmark = pytest.mark.gpu

import curvature_b200 as cb                      # noqa: E402

DEV = "cuda:0"
N = 8
INF_DAMPING = (254.0, 206.0)          # README.rst:262 "INF Norm" / "INF Scale" for ResNet18


def test_inf_rank100_on_resnet18():
    import torchvision
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    model = torchvision.models.resnet18(weights=None).to(DEV).train()
    x = torch.randn(N, 3, 224, 224, device=DEV)
    layers = selected_layers(model)
    kfac, diag = cb.KFAC(model), cb.Diagonal(model)
    _, labels, _ = orc.fisher_step(model, x)
    kfac.update(N)
    diag.update(N)
    t0 = time.perf_counter()
    eig = cb.get_eigenvectors(kfac.state)
    torch.cuda.synchronize()
    t_eig = time.perf_counter() - t0
    efb = cb.EFB(model, kfac.state, eigvecs=eig)
    orc.fisher_step(model, x, labels=labels)
    efb.update(N)
    inf = cb.INF(model, diag.state, kfac.state, efb.state, eigvecs=eig)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    inf.update(rank=100)
    torch.cuda.synchronize()
    t_upd = time.perf_counter() - t0
    t0 = time.perf_counter()
    inf.invert(*INF_DAMPING)
    torch.cuda.synchronize()
    t_inv = time.perf_counter() - t0
    print(f"ResNet-18 INF rank 100: eigenbases {t_eig * 1e3:.0f} ms, update {t_upd * 1e3:.0f} ms, invert {t_inv * 1e3:.0f} ms; "
          f"largest pre-sample matrix {max(v[3].shape[0] for v in inf.inv_state.values())}^2")
    n_d, s_d = INF_DAMPING
    gen = torch.Generator(device=DEV).manual_seed(2)
    checked_small = 0
    for li, layer in enumerate(layers):
        qa, qg, lam, corr = inf.state[layer]
        K, M = qa.shape[0], qg.shape[0]
        assert qa.shape[1] * qg.shape[1] == lam.numel() and corr.numel() == K * M
        z = torch.randn(K * M, device=DEV, generator=gen)
        smp = inf.sample(layer, z)
        assert tuple(smp.shape) == (M, K) and torch.isfinite(smp).all(), (li, str(layer))
        # ---- small layers: the oracle's literal formulation (kron materialised) in fp64 on the host
        if K * M <= 40000 and checked_small < 4:
            checked_small += 1
            QA, QG = eig[layer][0].double().cpu(), eig[layer][1].double().cpu()
            lam_full = efb.state[layer].double().cpu().t().contiguous().view(-1)
            diag_vec = diag.state[layer].double().cpu().t().contiguous().view(-1)
            o_qa, o_qg, o_lam = orc.INF._dim_reduction(QA, QG, lam_full, 100)
            assert torch.equal(o_qa.float(), qa.cpu()) and torch.equal(o_qg.float(), qg.cpu())       # same index selection
            assert rel_fro(lam, o_lam) <= 1e-6
            # (`invert` has already clamped the correction at 0 in place, like the reference, curvatures.py:523)
            o_corr = (diag_vec - orc.INF._diagonal_accumulator(o_qa, o_qg, o_lam)).clamp_min(0)
            assert (corr.double().cpu() - o_corr).norm() <= 1e-5 * diag_vec.norm(), (li, "correction")
            o_corr = corr.double().cpu().clamp_min(0)                 # isolate invert / sample from that difference
            ric = torch.reciprocal(s_d * o_corr + n_d).sqrt()
            pre = orc.INF.pre_sampler(o_qa, o_qg, (s_d * o_lam).sqrt(), ric)
            want = orc.INF.sampler(o_qa, o_qg, ric, pre, z.double().cpu()).reshape(K, M).t()
            # the reference's own arithmetic (fp32) on the same inputs, for scale
            pre32 = orc.INF.pre_sampler(o_qa.float(), o_qg.float(), (s_d * o_lam).sqrt().float(), ric.float())
            ref32 = orc.INF.sampler(o_qa.float(), o_qg.float(), ric.float(), pre32, z.cpu()).reshape(K, M).t()
            ref_err = rel_fro(ref32, want)
            err = rel_fro(smp, want)
            assert err <= max(2e-3, 3 * ref_err), (li, str(layer), err, ref_err)
        # ---- the two largest layers: kron-free device formulas against direct fp64 contractions
        if K * M >= 4608 * 512:
            QA, QG = qa.double(), qg.double()
            ra, rg = QA.shape[1], QG.shape[1]
            want_acc = ((QA ** 2) @ lam.double().view(ra, rg) @ (QG ** 2).t()).reshape(-1)
            got_acc = diag.state[layer].t().contiguous().view(-1).double() - corr.double()
            keep = corr > 0                           # (`invert` clamped the negative corrections at 0 in place)
            assert keep.float().mean() > 0.5
            assert rel_fro(got_acc[keep], want_acc[keep]) <= 1e-4, (li, rel_fro(got_acc[keep], want_acc[keep]))
            c = inf.inv_state[layer][2].double()
            c2 = (c ** 2).view(K, M)
            inner = torch.einsum('km,mb,md->kbd', c2, QG, QG)                     # (K, rg, rg)
            vtv = torch.einsum('ka,kc,kbd->abcd', QA, QA, inner).reshape(ra * rg, ra * rg)
            rl = (s_d * lam.double()).sqrt()
            vtv = rl[:, None] * vtv * rl[None, :]
            eye = torch.eye(ra * rg, device=DEV, dtype=torch.float64)
            A_c = torch.linalg.inv(torch.linalg.cholesky((vtv + vtv.t()) / 2))
            B_c = torch.linalg.cholesky((vtv + vtv.t()) / 2 + eye)
            C = A_c.t() @ (B_c - eye) @ A_c
            want_pre = rl[:, None] * torch.linalg.inv(torch.linalg.inv(C) + vtv) * rl[None, :]
            assert rel_fro(inf.inv_state[layer][3], want_pre) <= 2e-3, (li, rel_fro(inf.inv_state[layer][3], want_pre))
    assert checked_small >= 2
    inf.sample_and_replace()
    for p in model.parameters():
        assert torch.isfinite(p).all()
    for h in kfac.hooks:
        h.remove()
