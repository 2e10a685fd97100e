"""GPU (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the golden fixtures
produced by the real reference.  Tolerances are the north-star's: bit-exact where the arithmetic is integer
(im2col indexing on integer-valued data), relative Frobenius 1e-5 for fp32 factors and diagonals, a stated
1e-3 for the tensor-core tiers; samples are compared with the same supplied Gaussian noise."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import orc, rel_fro, model_from_golden, selected_layers, n_batches, conv_zoo

pytestmark = pytest.mark.gpu

import curvature_b200 as cb                      # noqa: E402
from curvature_b200 import _native as nat        # noqa: E402

DEV = "cuda:0"
FACTOR_TOL = {nat.PREC_FP32: 1e-5, nat.PREC_TF32: 1e-3, nat.PREC_BF16X3: 1e-5, nat.PREC_BF16: 1e-3,
              nat.PREC_TF32_TMA: 1e-3}


@pytest.fixture(autouse=True)
def _strict_fp32():
    # the model's own conv fwd/bwd (cuDNN, not ours) must not run in TF32 or it perturbs the inputs
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def tiers():
    out = [nat.PREC_FP32]
    import os
    for name in os.environ.get("CURVATURE_B200_TEST_TIERS", "tf32,tf32_tma").split(","):
        if name and name != "fp32":
            out.append(nat.PRECISION_NAMES[name])
    return out


def oracle_A(x, k, s, p, has_bias, dtype=torch.float64):
    cols = orc.unfold_patches(x.to(dtype), k, p, s)
    X = cols.permute(1, 0, 2).reshape(cols.shape[1], -1)
    if has_bias:
        X = torch.cat([X, torch.ones_like(X[:1])], 0)
    return X @ X.t(), X.shape[1]


GEOMS = [
    # N, C, H, W, kernel, stride, padding, bias
    (2, 1, 5, 5, (3, 3), (1, 1), (1, 1), True),
    (3, 3, 7, 6, (3, 2), (2, 1), (1, 0), True),
    (2, 2, 8, 8, (1, 1), (2, 2), (0, 0), False),
    (2, 4, 9, 11, (5, 5), (1, 1), (2, 2), True),
    (1, 3, 12, 12, (7, 7), (2, 2), (3, 3), False),
    (2, 2, 6, 9, (1, 3), (1, 2), (0, 2), True),
    (2, 5, 4, 4, (4, 4), (1, 1), (0, 0), False),
    (1, 1, 10, 3, (3, 3), (3, 1), (2, 2), True),
    (5, 8, 14, 14, (3, 3), (1, 1), (1, 1), False),     # K = 72: two 64-row tiles
    (3, 16, 7, 7, (3, 3), (2, 2), (1, 1), True),       # K = 145: three tiles, odd sizes
    (70, 3, 6, 6, (3, 3), (1, 1), (1, 1), True),       # R = 2520: several contraction splits
]


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("prec", tiers())
def test_implicit_im2col_syrk_bit_exact_on_integers(geom, prec):
    """Integer-valued activations make every product and partial sum exactly representable (also in tf32 /
    bf16), so any indexing error -- wrong tap, wrong padding, wrong row order, lost bias row -- shows up as an
    exact mismatch.  This is the 'bit-exact for im2col indexing' gate."""
    N, C, H, W, k, s, p, bias = geom
    gen = torch.Generator().manual_seed(hash(geom) % (2 ** 31))
    x = torch.randint(0, 4, (N, C, H, W), generator=gen).float()
    want, R = oracle_A(x, k, s, p, bias)
    K = want.shape[0]
    out = torch.zeros(K, K, device=DEV)
    nat.syrk_conv_accum(x.to(DEV), k, s, p, bias, 1.0, out, prec)
    assert torch.equal(out.cpu().double(), want), f"max diff {(out.cpu().double() - want).abs().max()}"
    # running sum: a second call with alpha = 2 adds twice the factor (curvatures.py:346-350 is a plain +=)
    nat.syrk_conv_accum(x.to(DEV), k, s, p, bias, 2.0, out, prec)
    assert torch.equal(out.cpu().double(), 3 * want)


@pytest.mark.parametrize("shape,bias", [((4, 6), True), ((5, 10), False), ((3, 7, 5, 4), False), ((2, 70, 3, 3), False),
                                        ((9, 130), True), ((1, 1), True), ((300, 3, 2, 2), False)])
@pytest.mark.parametrize("prec", tiers())
def test_rows_syrk_bit_exact_on_integers(shape, bias, prec):
    gen = torch.Generator().manual_seed(len(shape) * 1000 + shape[1])
    g = torch.randint(-3, 4, shape, generator=gen).float()
    M = shape[1]
    X = g.reshape(shape[0], M, -1).permute(1, 0, 2).reshape(M, -1).double()
    if bias:
        X = torch.cat([X, torch.ones_like(X[:1])], 0)
    want = X @ X.t()
    out = torch.zeros(want.shape[0], want.shape[0], device=DEV)
    nat.syrk_rows_accum(g.to(DEV), bias, 1.0, out, prec)
    assert torch.equal(out.cpu().double(), want)


@pytest.mark.parametrize("prec", tiers())
def test_kfac_layer_factors_from_recorded_tensors(golden, prec):
    """Kernel-level parity on the exact tensors the reference's hooks recorded (convzoo: strides, asymmetric
    padding, non-square kernels, bias / no bias, Linear with and without bias)."""
    g = golden("convzoo")
    model = conv_zoo()
    for li, layer in enumerate(selected_layers(model)):
        x = torch.from_numpy(g[f"last_input/{li}"])
        go = torch.from_numpy(g[f"last_gradout/{li}"])          # already scaled by N, as the reference records it
        A_ref, G_ref = orc.kfac_factors(x, go, layer)
        bias = layer.bias is not None
        K, M = A_ref.shape[0], G_ref.shape[0]
        A = torch.zeros(K, K, device=DEV)
        G = torch.zeros(M, M, device=DEV)
        if layer.__class__.__name__ == "Conv2d":
            R = x.shape[0] * go.shape[2] * go.shape[3]
            nat.syrk_conv_accum(x.to(DEV), layer.kernel_size, layer.stride, layer.padding, bias, 1.0 / R, A, prec)
        else:
            R = x.shape[0]
            nat.syrk_rows_accum(x.to(DEV), bias, 1.0 / R, A, prec)
        nat.syrk_rows_accum(go.to(DEV).contiguous(), False, 1.0 / R, G, prec)
        assert rel_fro(A, A_ref) <= FACTOR_TOL[prec], (li, rel_fro(A, A_ref))
        assert rel_fro(G, G_ref) <= FACTOR_TOL[prec], (li, rel_fro(G, G_ref))
        assert torch.equal(A, A.t()) and torch.equal(G, G.t())


def run_estimation(name, g, prec):
    """The reference's estimation loop (scripts/factors.py:46-61) on the GPU with the fixture's inputs and labels."""
    model = model_from_golden(name, g, DEV)
    layers = selected_layers(model)
    N = int(g["meta/batch"])
    kfac = cb.KFAC(model, precision=prec)
    diag = cb.Diagonal(model)
    for b in range(n_batches(g)):
        orc.fisher_step(model, torch.from_numpy(g[f"x/{b}"]).to(DEV), labels=torch.from_numpy(g[f"labels/{b}"]).to(DEV))
        kfac.update(N)
        diag.update(N)
    eig = {l: (torch.from_numpy(g[f"eig_QA/{li}"]).to(DEV), torch.from_numpy(g[f"eig_QG/{li}"]).to(DEV))
           for li, l in enumerate(layers)}
    efb = cb.EFB(model, kfac.state, eigvecs=eig)
    for b in range(n_batches(g)):
        orc.fisher_step(model, torch.from_numpy(g[f"x/{b}"]).to(DEV), labels=torch.from_numpy(g[f"labels/{b}"]).to(DEV))
        efb.update(N)
    return model, layers, kfac, diag, efb, eig


@pytest.mark.parametrize("name", ["convzoo", "lenet5"])
@pytest.mark.parametrize("prec", tiers())
def test_estimators_match_reference_fixtures(name, golden, prec):
    g = golden(name)
    model, layers, kfac, diag, efb, eig = run_estimation(name, g, prec)
    assert list(kfac.state.keys()) == layers                     # model.modules() order, module-keyed
    for li, l in enumerate(layers):
        assert isinstance(kfac.state[l], list) and len(kfac.state[l]) == 2
        eA, eG = rel_fro(kfac.state[l][0], g[f"kfac_A/{li}"]), rel_fro(kfac.state[l][1], g[f"kfac_G/{li}"])
        assert eA <= FACTOR_TOL[prec] and eG <= FACTOR_TOL[prec], (name, li, eA, eG)
        assert rel_fro(diag.state[l], g[f"diag/{li}"]) <= 1e-5, (name, li, rel_fro(diag.state[l], g[f"diag/{li}"]))
        efb_tol = 5e-5 if prec == nat.PREC_FP32 else 1e-3       # tensor-core tiers: TF32 GEMMs where TMA can address them
        assert rel_fro(efb.state[l], g[f"efb_lambda/{li}"]) <= efb_tol, (name, li, rel_fro(efb.state[l], g[f"efb_lambda/{li}"]))
        assert rel_fro(efb.diags[l], g[f"diag/{li}"]) <= 1e-5
    # A[-1,-1] counts the updates (ones-row) wherever the layer has a bias: plain-sum accumulation
    for l in layers:
        if l.bias is not None:
            assert abs(kfac.state[l][0][-1, -1].item() - n_batches(g)) < 1e-5

    # ---- invert + samples with the supplied noise (fp32 kernels; the factors come from the fixture so that the
    #      comparison isolates K4 / K5 from the estimation tier) ----
    for li, l in enumerate(layers):
        kfac.state[l][0].copy_(torch.from_numpy(g[f"kfac_A/{li}"]))
        kfac.state[l][1].copy_(torch.from_numpy(g[f"kfac_G/{li}"]))
        efb.state[l].copy_(torch.from_numpy(g[f"efb_lambda/{li}"]))
        diag.state[l].copy_(torch.from_numpy(g[f"diag/{li}"]))
    kfac.invert(*g["meta/kfac_damp"].tolist())
    diag.invert(*g["meta/diag_damp"].tolist())
    efb.invert(*g["meta/diag_damp"].tolist())
    for li, l in enumerate(layers):
        LA, LG = kfac.inv_state[l]
        assert isinstance(kfac.inv_state[l], tuple)
        eA, eG = rel_fro(LA, g[f"kfac_LA/{li}"]), rel_fro(LG, g[f"kfac_LG/{li}"])
        assert eA <= 1e-4 and eG <= 1e-4, (name, li, eA, eG)
        assert torch.equal(LA, torch.tril(LA))
        z = torch.from_numpy(g[f"noise_KM/{li}"]).to(DEV)
        smp_tol = 1e-4 if prec == nat.PREC_FP32 else 1e-3
        assert rel_fro(kfac.sample(l, z), g[f"kfac_sample/{li}"]) <= smp_tol
        assert rel_fro(efb.sample(l, z), g[f"efb_sample/{li}"]) <= smp_tol
        if f"diag_sample/{li}" in g.files:
            zd = torch.from_numpy(g[f"noise_MK/{li}"]).to(DEV)
            assert rel_fro(diag.sample(l, zd), g[f"diag_sample/{li}"]) <= 1e-6
            assert rel_fro(diag.inv_state[l], g[f"diag_inv/{li}"]) <= 1e-6
            assert rel_fro(efb.inv_state[l], g[f"efb_inv/{li}"]) <= 1e-6

    # ---- sample_and_replace with the same noise: end state of every parameter ----
    noise = {l: torch.from_numpy(g[f"noise_KM/{li}"]).to(DEV) for li, l in enumerate(layers)}
    kfac.sample_and_replace(noise=noise)
    sd = model.state_dict()
    if any(k.startswith("replaced/") for k in g.files):
        for k, v in sd.items():
            assert rel_fro(v, g[f"replaced/{k}"]) <= (1e-5 if prec == nat.PREC_FP32 else 1e-3), k
    else:
        for li, l in enumerate(layers):
            s = torch.from_numpy(g[f"kfac_sample/{li}"])
            w0 = kfac.model_state[[k for k, v in model.state_dict(keep_vars=True).items() if v is l.weight][0]].cpu()
            want_w = w0 + s[:, :w0[0].numel()].reshape(w0.shape)
            assert rel_fro(l.weight.data, want_w) <= (1e-5 if prec == nat.PREC_FP32 else 1e-3)
    # a second call starts again from the mean (load_state_dict semantics, curvatures.py:119)
    kfac.sample_and_replace(noise=noise)
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]) or rel_fro(v, sd[k]) <= 1e-7


@pytest.mark.parametrize("name", ["convzoo", "lenet5"])
def test_inf_matches_reference_fixtures(name, golden):
    g = golden(name)
    model = model_from_golden(name, g, DEV)
    layers = selected_layers(model)
    dev = lambda key, li: torch.from_numpy(g[f"{key}/{li}"]).to(DEV)   # noqa: E731
    diags = {l: dev("diag", li) for li, l in enumerate(layers)}
    factors = {l: [dev("kfac_A", li), dev("kfac_G", li)] for li, l in enumerate(layers)}
    lambdas = {l: dev("efb_lambda", li) for li, l in enumerate(layers)}
    eig = {l: (dev("eig_QA", li), dev("eig_QG", li)) for li, l in enumerate(layers)}
    inf = cb.INF(model, diags, factors, lambdas, eigvecs=eig)
    inf.update(rank=int(g["meta/rank"]))
    for li, l in enumerate(layers):
        a, b, lam, corr = inf.state[l]
        assert torch.equal(a.cpu(), torch.from_numpy(g[f"inf_state_lrQA/{li}"]))       # pure index selection
        assert torch.equal(b.cpu(), torch.from_numpy(g[f"inf_state_lrQG/{li}"]))
        assert torch.equal(lam.cpu(), torch.from_numpy(g[f"inf_state_lrlambda/{li}"]))
        # the correction is a difference of two nearly equal diagonals: compare on the scale of the diagonal
        scale = np.linalg.norm(g[f"diag/{li}"])
        assert (corr.cpu() - torch.from_numpy(g[f"inf_state_correction/{li}"])).norm().item() <= 1e-5 * scale
    for li, l in enumerate(layers):     # isolate invert / sample from the small differences above
        inf.state[l] = (inf.state[l][0], inf.state[l][1], inf.state[l][2], dev("inf_state_correction", li))
    inf.invert(*g["meta/inf_damp"].tolist())
    # The reference's pre-sampler (Cholesky + three explicit inverses in fp32, curvatures.py:564-570) is
    # ill-conditioned: its own fp32 output is O(1) away from the fp64 evaluation of the same formulas on several
    # layers (measured in the build container).  Parity is therefore stated against the fp64 oracle: the CUDA path
    # must be within 2e-3, or within 3x of the reference's own fp32 error wherever that is larger.
    n_d, s_d = g["meta/inf_damp"].tolist()
    for li, l in enumerate(layers):
        qa64, qg64 = inf.state[l][0].double().cpu(), inf.state[l][1].double().cpu()
        lam64 = inf.state[l][2].double().cpu()
        corr64 = torch.from_numpy(g[f"inf_state_correction/{li}"]).double().clamp_min(0)
        ric64 = torch.reciprocal(s_d * corr64 + n_d).sqrt()
        pre64 = orc.INF.pre_sampler(qa64, qg64, (s_d * lam64).sqrt(), ric64)
        z = dev("noise_KM", li).reshape(-1)
        smp64 = orc.INF.sampler(qa64, qg64, ric64, pre64, z.double().cpu()).reshape(qa64.shape[0], qg64.shape[0]).t()
        ref_err = rel_fro(g[f"inf_sample/{li}"], smp64)
        got = inf.sample(l, z)
        assert got.shape == smp64.shape
        err = rel_fro(got, smp64)
        assert err <= max(2e-3, 3 * ref_err), (name, li, err, ref_err)
        if ref_err <= 1e-3:     # where the reference's own fp32 chain is accurate: parity against ITS fixture directly
            assert rel_fro(got, g[f"inf_sample/{li}"]) <= 2e-3, (name, li, rel_fro(got, g[f"inf_sample/{li}"]))
        pre = inf.inv_state[l][3]
        if f"inf_pre/{li}" in g.files:
            ref_pre_err = rel_fro(g[f"inf_pre/{li}"], pre64)
            assert rel_fro(pre, pre64) <= max(2e-3, 3 * ref_pre_err), (name, li, rel_fro(pre, pre64), ref_pre_err)
        assert rel_fro(inf.inv_state[l][2], ric64) <= 1e-6


RESNET_LAYERS = [
    # name, N, C, H, W, k, s, p
    ("stem 7x7 s2", 2, 3, 224, 224, 7, 2, 3),
    ("3x3 64ch 56^2", 2, 64, 56, 56, 3, 1, 1),
    ("3x3 s2 128ch", 2, 128, 56, 56, 3, 2, 1),
    ("1x1 256ch 56^2", 2, 256, 56, 56, 1, 1, 0),
    ("1x1 s2 256ch", 2, 256, 56, 56, 1, 2, 0),
    ("3x3 256ch 14^2", 4, 256, 14, 14, 3, 1, 1),
    ("3x3 512ch 7^2", 4, 512, 7, 7, 3, 1, 1),
]


@pytest.mark.parametrize("layer", RESNET_LAYERS, ids=[l[0] for l in RESNET_LAYERS])
@pytest.mark.parametrize("prec", tiers())
def test_resnet_shaped_factors_against_fp64(layer, prec):
    """BASELINE config shapes (ResNet-18/50/152 layer geometries) at a reduced batch, against an fp64
    unfold + matmul on the same device (size-independent check of the fused kernel on real tile counts)."""
    name, N, C, H, W, k, s, p = layer
    torch.manual_seed(1)
    x = torch.relu(torch.randn(N, C, H, W, device=DEV))            # post-ReLU-like activations
    cols = F.unfold(x.double(), k, padding=p, stride=s)
    X = cols.permute(1, 0, 2).reshape(cols.shape[1], -1)
    want = (X @ X.t()) / X.shape[1]
    out = torch.zeros_like(want, dtype=torch.float32)
    nat.syrk_conv_accum(x, (k, k), (s, s), (p, p), False, 1.0 / X.shape[1], out, prec)
    err = rel_fro(out, want)
    assert err <= FACTOR_TOL[prec], (name, err)
    if prec == nat.PREC_FP32:   # split-R partials meet in fp32 atomics: symmetric up to summation order
        assert rel_fro(out, out.t()) <= 1e-6
    else:                       # tensor-core tiers reduce partial tiles in a fixed order and mirror: exact
        assert torch.equal(out, out.t())
    # G-type operand of the same layer: (N, M, OH, OW) gradient-like tensor
    OH = (H + 2 * p - k) // s + 1
    gten = torch.randn(N, min(C, 256), OH, OH, device=DEV) * 1e-3
    Xg = gten.double().permute(1, 0, 2, 3).reshape(gten.shape[1], -1)
    wantg = (Xg @ Xg.t()) * (N * N / Xg.shape[1])
    outg = torch.zeros_like(wantg, dtype=torch.float32)
    nat.syrk_rows_accum(gten, False, N * N / Xg.shape[1], outg, prec)
    assert rel_fro(outg, wantg) <= FACTOR_TOL[prec], (name, rel_fro(outg, wantg))


@pytest.mark.parametrize("D", [1, 6, 31, 32, 33, 85, 401, 1000])
def test_damped_cholesky_of_inverse(D):
    torch.manual_seed(D)
    X = torch.randn(D, 2 * D + 3, device=DEV)
    Fm = (X @ X.t() / X.shape[1]).contiguous()
    Fm = Fm + 1e-3 * torch.randn_like(Fm)                 # slightly asymmetric, like a summed factor
    add, mul = 0.3, 2.5
    out = torch.empty_like(Fm)
    info = nat.chol_inv_batched([Fm], [add], [mul], [out])
    assert info.item() == 0
    reg = mul ** 0.5 * Fm.double() + add ** 0.5 * torch.eye(D, device=DEV, dtype=torch.float64)
    reg = (reg + reg.t()) / 2
    want = torch.linalg.cholesky(torch.linalg.inv(reg))
    assert rel_fro(out, want) <= 1e-4, rel_fro(out, want)
    resid = out.double() @ out.double().t() @ reg - torch.eye(D, device=DEV, dtype=torch.float64)
    assert resid.norm().item() / D ** 0.5 <= 1e-4
    assert torch.equal(out, torch.tril(out))


def test_cholesky_batch_mixed_sizes_and_failure_flag():
    torch.manual_seed(0)
    mats = []
    for D in (5, 70, 33, 128):
        X = torch.randn(D, 3 * D, device=DEV)
        mats.append((X @ X.t() / X.shape[1]).contiguous())
    mats.append(-torch.eye(4, device=DEV))                   # not positive definite after damping
    outs = [torch.empty_like(m) for m in mats]
    info = nat.chol_inv_batched(mats, [0.5] * 5, [1.0] * 5, outs).cpu()
    assert info[:4].tolist() == [0, 0, 0, 0] and info[4].item() > 0
    for m, o in zip(mats[:4], outs[:4]):
        want = orc.damped_inverse_cholesky(m.cpu(), 0.5, 1.0)
        assert rel_fro(o, want) <= 1e-4
    model = torch.nn.Sequential(torch.nn.Linear(3, 2)).to(DEV)
    kfac = cb.KFAC(model)
    model(torch.randn(4, 3, device=DEV)).sum().backward()
    kfac.update(4)
    kfac.state[model[0]][0].copy_(-torch.eye(4))
    with pytest.raises(RuntimeError, match="not positive definite"):
        kfac.invert(0.5, 1.0)


@pytest.mark.parametrize("M,K0,bias", [(6, 25, True), (16, 150, True), (7, 64, False), (1000, 2048, True), (3, 1, True)])
def test_streaming_kernels(M, K0, bias):
    torch.manual_seed(M)
    w = torch.randn(M, K0, device=DEV)
    b = torch.randn(M, device=DEV) if bias else None
    grads = torch.cat([w, b[:, None]], 1) if bias else w
    K = grads.shape[1]
    state = torch.rand(M, K, device=DEV)
    want = state + grads ** 2 * 32
    gout = torch.empty(M, K, device=DEV)
    nat.diag_accum(w, b, 32, state=state, grads_out=gout)
    assert rel_fro(state, want) <= 1e-6 and torch.equal(gout, grads)
    inv = torch.empty_like(state)
    nat.elementwise_inv_sqrt(state, 0.1, 1e3, inv)
    assert rel_fro(inv, torch.reciprocal(1e3 * state + 0.1).sqrt()) <= 1e-6
    z = torch.randn(M, K, device=DEV)
    mu_w, mu_b = torch.randn(M, K0, device=DEV), (torch.randn(M, device=DEV) if bias else None)
    w_out, b_out = torch.empty_like(mu_w), (torch.empty_like(mu_b) if bias else None)
    s_out = torch.empty_like(z)
    nat.diag_sample(z, inv, bias, mu_w=mu_w, mu_b=mu_b, w_out=w_out, b_out=b_out, s_out=s_out)
    assert torch.equal(s_out, z * inv)
    assert torch.equal(w_out, mu_w + (z * inv)[:, :K0])
    if bias:
        assert torch.equal(b_out, mu_b + (z * inv)[:, K0])


@pytest.mark.parametrize("M,K0,bias", [(10, 84, True), (120, 400, True), (64, 576, False), (256, 1152, False)])
def test_efb_projection_and_matrix_normal_draw(M, K0, bias):
    torch.manual_seed(K0)
    K = K0 + bias
    QA = torch.linalg.qr(torch.randn(K, K, device=DEV))[0].contiguous()
    QG = torch.linalg.qr(torch.randn(M, M, device=DEV))[0].contiguous()
    G = torch.randn(M, K, device=DEV)
    lam = torch.rand(M, K, device=DEV)
    want = lam.double() + (QG.double().t() @ G.double() @ QA.double()) ** 2
    nat.efb_project_accum(QG, QA, G, lam)
    assert rel_fro(lam, want) <= 1e-5
    LA = torch.tril(torch.randn(K, K, device=DEV)).contiguous()
    LG = torch.tril(torch.randn(M, M, device=DEV)).contiguous()
    z = torch.randn(K, M, device=DEV)
    S_want = (LA.double() @ z.double() @ LG.double().t()).t()             # curvatures.py:392
    mu_w, mu_b = torch.randn(M, K0, device=DEV), torch.randn(M, device=DEV)
    w_out, b_out, s_out = torch.empty_like(mu_w), torch.empty_like(mu_b), torch.empty(M, K, device=DEV)
    nat.sample_matrix_normal(LG, LA, z, bias, mu_w=mu_w, mu_b=mu_b if bias else None, w_out=w_out,
                             b_out=b_out if bias else None, s_out=s_out)
    assert rel_fro(s_out, S_want) <= 1e-5
    assert rel_fro(w_out, mu_w.double() + S_want[:, :K0]) <= 1e-5
    if bias:
        assert rel_fro(b_out, mu_b.double() + S_want[:, K0]) <= 1e-5
    rs = torch.rand(M, K, device=DEV)
    nat.sample_matrix_normal(QG, QA, z, False, row_scale=rs, s_out=s_out)   # EFB.sample, curvatures.py:458-460
    want_efb = (QA.double() @ (z.double() * rs.double().t()) @ QG.double().t()).t()
    assert rel_fro(s_out, want_efb) <= 1e-5


@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (200, 136, 72), (64, 576, 64), (512, 1152, 512), (100, 260, 36), (4, 8, 4)])
@pytest.mark.parametrize("transA,transB", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_tensor_core_operand_forms_exact_on_integers(m, n, k, transA, transB):
    """tcgen05 TF32 GEMM (gemm_tc.cu): K-major and MN-major operand forms in all four combinations, ragged tiles by
    TMA out-of-bounds fill.  Small integers are exact in TF32, so the product must be bit-exact."""
    gen = torch.Generator().manual_seed(m * 7 + n * 3 + k)
    A = torch.randint(-3, 4, (k, m) if transA else (m, k), generator=gen).float().to(DEV)
    B = torch.randint(-3, 4, (n, k) if transB else (k, n), generator=gen).float().to(DEV)
    want = (A.t() if transA else A).double() @ (B.t() if transB else B).double()
    got = nat.gemm(A, B, transA=transA, transB=transB, precision=nat.PREC_TF32)
    assert torch.equal(got.double(), want)
    C = torch.ones(m, n, device=DEV)
    nat.gemm(A, B, transA=transA, transB=transB, alpha=2.0, beta=3.0, out=C, precision=nat.PREC_TF32)
    assert torch.equal(C.double(), 2 * want + 3)


@pytest.mark.parametrize("M,K0", [(64, 576), (256, 1152), (128, 64), (512, 2304), (1000, 512)])
def test_efb_projection_and_matrix_normal_draw_tensor_core(M, K0):
    """K3 / K5 on the tensor cores (two chained TF32 GEMMs each, fused epilogues), stated 1e-3 tier."""
    torch.manual_seed(K0)
    K = K0
    QA = torch.linalg.qr(torch.randn(K, K, device=DEV))[0].contiguous()
    QG = torch.linalg.qr(torch.randn(M, M, device=DEV))[0].contiguous()
    G = torch.randn(M, K, device=DEV)
    lam = torch.rand(M, K, device=DEV)
    want = lam.double() + (QG.double().t() @ G.double() @ QA.double()) ** 2
    nat.efb_project_accum(nat.round_tf32(QG), nat.round_tf32(QA), nat.round_tf32(G), lam, nat.PREC_TF32)
    err = rel_fro(lam, want)
    assert err <= 1e-3, err
    LA = torch.tril(torch.randn(K, K, device=DEV)).contiguous()
    LG = torch.tril(torch.randn(M, M, device=DEV)).contiguous()
    z = torch.randn(K, M, device=DEV)
    S_want = (LA.double() @ z.double() @ LG.double().t()).t()
    mu_w = torch.randn(M, K0, device=DEV)
    w_out, s_out = torch.empty_like(mu_w), torch.empty(M, K, device=DEV)
    nat.sample_matrix_normal(nat.round_tf32(LG), nat.round_tf32(LA), nat.round_tf32(z), False, mu_w=mu_w, w_out=w_out,
                             s_out=s_out, precision=nat.PREC_TF32)
    e2 = rel_fro(s_out, S_want)
    assert e2 <= 1e-3, e2
    assert rel_fro(w_out, mu_w.double() + S_want) <= 1e-3
    print(f"[K3/K5 tensor core M={M} K={K}] lambda err {err:.2e}, sample err {e2:.2e}")


def test_diagonal_multihead_attention_keys():
    torch.manual_seed(0)
    mha = torch.nn.MultiheadAttention(16, 4).to(DEV)
    model = torch.nn.ModuleList([mha])
    x = torch.randn(5, 3, 16, device=DEV)
    out, _ = mha(x, x, x)
    out.sum().backward()
    diag = cb.Diagonal(model)
    diag.update(3)
    assert list(diag.state.keys()) == ['attn_in', 'attn_out']
    want_in = torch.cat([mha.in_proj_weight.grad, mha.in_proj_bias.grad[:, None]], 1) ** 2 * 3
    assert rel_fro(diag.state['attn_in'], want_in) <= 1e-6
    diag.invert(0.1, 10.0)
    diag.sample_and_replace()
    assert not torch.equal(mha.in_proj_weight.data, diag.model_state['0.in_proj_weight'])


def test_gemm_entry_point():
    torch.manual_seed(0)
    A, B = torch.randn(70, 33, device=DEV), torch.randn(33, 129, device=DEV)
    assert rel_fro(nat.gemm(A, B), A.double() @ B.double()) <= 1e-6
    assert rel_fro(nat.gemm(A, torch.randn(70, 5, device=DEV), transA=True).shape[0], 33) == 0
    Bt = B.t().contiguous()
    assert rel_fro(nat.gemm(A, Bt, transB=True), A.double() @ B.double()) <= 1e-6
    C = torch.ones(70, 129, device=DEV)
    nat.gemm(A, B, alpha=2.0, beta=3.0, out=C)
    assert rel_fro(C, 2 * (A.double() @ B.double()) + 3) <= 1e-6


# ---- channels-last (NHWC) operands: the TMA-fed MN-major tcgen05 kernel ---------------------------------------
TC_TIERS = [nat.PREC_TF32, nat.PREC_TF32_TMA, nat.PREC_BF16]

NHWC_GEOMS = [
    # N, C, H, W, kernel, stride, padding
    (2, 32, 8, 8, (3, 3), (1, 1), (1, 1)),       # 8x8 grid: boxes of 8x4
    (3, 64, 14, 14, (3, 3), (1, 1), (1, 1)),     # 14x14: boxes of 14x2 + 4 zero rows
    (2, 64, 7, 7, (3, 3), (1, 1), (1, 1)),       # 7x7: boxes of 7x1 + 1 zero row
    (2, 32, 28, 28, (3, 3), (2, 2), (1, 1)),     # stride 2: elementStrides in the tensor map
    (2, 96, 12, 16, (3, 3), (1, 1), (1, 1)),     # K = 864: three full 256-row blocks + 96 rows
    (2, 128, 14, 14, (3, 3), (1, 1), (1, 1)),    # K = 1152: bf16-eligible (C % 64 == 0), 14x14 boxes
    (3, 64, 28, 28, (3, 3), (1, 1), (1, 1)),     # K = 576 on a 28x28 grid
    (2, 64, 12, 10, (3, 5), (1, 2), (1, 2)),     # non-square kernel, mixed stride / padding
    (2, 128, 9, 9, (1, 1), (2, 2), (0, 0)),      # 1x1 stride 2 (ResNet downsample)
    (4, 40, 6, 6, (1, 1), (1, 1), (0, 0)),       # flat, C % 32 != 0: channel tail by out-of-bounds fill
    (3, 288, 5, 5, (1, 1), (1, 1), (0, 0)),      # flat, two blocks, R = 75 (ragged last box)
    (70, 32, 6, 6, (3, 3), (1, 1), (1, 1)),      # many boxes: several contraction splits
    (2, 32, 6, 6, (5, 5), (1, 1), (2, 2)),       # 25 taps
    (1, 64, 4, 4, (3, 3), (1, 1), (0, 0)),       # no padding: 2x2 output grid
]


@pytest.mark.parametrize("geom", NHWC_GEOMS)
@pytest.mark.parametrize("prec", TC_TIERS)
def test_nhwc_implicit_im2col_syrk_bit_exact_on_integers(geom, prec):
    """Same gate as the NCHW test, for channels-last activations: tap shifts are TMA box coordinates, padding is
    the out-of-bounds fill, strides are elementStrides; any indexing error is an exact mismatch."""
    N, C, H, W, k, s, p = geom
    gen = torch.Generator().manual_seed(hash(geom) % (2 ** 31))
    x = torch.randint(-2, 3, (N, C, H, W), generator=gen).float()
    want, R = oracle_A(x, k, s, p, False)
    K = want.shape[0]
    out = torch.zeros(K, K, device=DEV)
    xc = x.to(DEV).contiguous(memory_format=torch.channels_last)
    assert nat._is_channels_last(xc) or C == 1
    before = nat.launch_calls
    nat.syrk_conv_accum(xc, k, s, p, False, 1.0, out, prec)
    assert nat.workspace_bytes(nat.OP_SYRK_CONV_NHWC, [N, C, H, W, *k, *s, *p, 0, prec]) > 0   # took the TMA path
    assert nat.launch_calls - before >= 2
    assert torch.equal(out.cpu().double(), want), f"max diff {(out.cpu().double() - want).abs().max()}"
    nat.syrk_conv_accum(xc, k, s, p, False, 2.0, out, prec)
    assert torch.equal(out.cpu().double(), 3 * want)


@pytest.mark.parametrize("shape", [(256, 1000), (5, 64), (3, 48, 5, 5), (2, 256, 14, 14), (300, 36, 2, 2), (7, 2048, 1, 1),
                                   (40, 776, 3, 3)])
@pytest.mark.parametrize("prec", TC_TIERS)
def test_nhwc_rows_syrk_bit_exact_on_integers(shape, prec):
    gen = torch.Generator().manual_seed(len(shape) * 1000 + shape[1])
    g = torch.randint(-3, 4, shape, generator=gen).float()
    M = shape[1]
    X = g.reshape(shape[0], M, -1).permute(1, 0, 2).reshape(M, -1).double()
    want = X @ X.t()
    gd = g.to(DEV)
    if gd.dim() == 4:
        gd = gd.contiguous(memory_format=torch.channels_last)
    out = torch.zeros(M, M, device=DEV)
    nat.syrk_rows_accum(gd, False, 1.0, out, prec)
    assert torch.equal(out.cpu().double(), want)


@pytest.mark.parametrize("layer", RESNET_LAYERS[1:], ids=[l[0] for l in RESNET_LAYERS[1:]])
@pytest.mark.parametrize("prec", TC_TIERS)
def test_nhwc_resnet_shaped_factors_against_fp64(layer, prec):
    name, N, C, H, W, k, s, p = layer
    torch.manual_seed(1)
    x = torch.relu(torch.randn(N, C, H, W, device=DEV)).contiguous(memory_format=torch.channels_last)
    cols = F.unfold(x.double(), k, padding=p, stride=s)
    X = cols.permute(1, 0, 2).reshape(cols.shape[1], -1)
    want = (X @ X.t()) / X.shape[1]
    out = torch.zeros_like(want, dtype=torch.float32)
    nat.syrk_conv_accum(x, (k, k), (s, s), (p, p), False, 1.0 / X.shape[1], out, prec)
    err = rel_fro(out, want)
    # all of these are stated 1e-3 tiers (FACTOR_TOL); measured on B200 on these post-ReLU distributions (N = 2-4):
    # round-to-nearest TF32 operands (tier tf32) 1.2e-5 .. 7.6e-5, bf16 copies 1e-4 .. 7e-4, raw fp32 words (tier
    # tf32_tma, the tensor core truncates them) a systematic 7e-4.  The tf32 tier is held to 2e-4 here so that a
    # regression to truncation (7e-4) cannot hide behind the stated tier.
    assert err <= (2e-4 if prec == nat.PREC_TF32 else FACTOR_TOL[prec]), (name, err)
    print(f"[nhwc {name} tier {prec}] rel. Frobenius error {err:.3e}")
    assert torch.equal(out, out.t())
    OH = (H + 2 * p - k) // s + 1
    gten = (torch.randn(N, min(C, 256), OH, OH, device=DEV) * 1e-3).contiguous(memory_format=torch.channels_last)
    Xg = gten.double().permute(1, 0, 2, 3).reshape(gten.shape[1], -1)
    wantg = (Xg @ Xg.t()) * (N * N / Xg.shape[1])
    outg = torch.zeros_like(wantg, dtype=torch.float32)
    nat.syrk_rows_accum(gten, False, N * N / Xg.shape[1], outg, prec)
    assert rel_fro(outg, wantg) <= FACTOR_TOL[prec], (name, rel_fro(outg, wantg))


def test_kfac_channels_last_model_matches_nchw_model():
    """The same network run in torch.channels_last memory format records channels-last activations / gradients;
    the factors, the inverse factors and a posterior sample (same noise) must agree with the NCHW fp32 path."""
    torch.manual_seed(3)

    def make():
        return torch.nn.Sequential(
            torch.nn.Conv2d(3, 32, 3, padding=1, bias=False), torch.nn.ReLU(),
            torch.nn.Conv2d(32, 64, 3, stride=2, padding=1, bias=False), torch.nn.ReLU(),
            torch.nn.Conv2d(64, 64, 1, bias=False), torch.nn.ReLU(),
            torch.nn.Conv2d(64, 32, 3, padding=1, bias=True), torch.nn.ReLU(),
            torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(), torch.nn.Linear(32, 10))
    ref_model = make().to(DEV)
    cl_model = make().to(DEV)
    cl_model.load_state_dict(ref_model.state_dict())
    cl_model = cl_model.to(memory_format=torch.channels_last)
    x = torch.randn(8, 3, 16, 16, device=DEV)
    labels = torch.randint(0, 10, (8,), device=DEV)
    ref, cl = cb.KFAC(ref_model, precision="fp32"), cb.KFAC(cl_model, precision="tf32")
    for model, est, inp in ((ref_model, ref, x), (cl_model, cl, x.contiguous(memory_format=torch.channels_last))):
        for _ in range(2):
            model.zero_grad()
            torch.nn.functional.cross_entropy(model(inp), labels).backward()
            est.update(8)
    n_tma = 0
    for lr, lc in zip(ref.state, cl.state):
        xr = cl.record[lc][0]
        n_tma += int(xr.dim() == 4 and nat._is_channels_last(xr))
        for f in range(2):
            assert rel_fro(cl.state[lc][f], ref.state[lr][f]) <= FACTOR_TOL[nat.PREC_TF32], (lc, f)
    assert n_tma >= 3          # the inner convolutions really saw channels-last activations
    ref.invert(0.5, 10.0)
    cl.invert(0.5, 10.0)
    noise_r = {l: torch.randn(ref.inv_state[l][0].shape[0], ref.inv_state[l][1].shape[0], device=DEV) for l in ref.state}
    noise_c = {lc: noise_r[lr] for lr, lc in zip(ref.state, cl.state)}
    ref.sample_and_replace(noise_r)
    cl.sample_and_replace(noise_c)
    for (k1, v1), (k2, v2) in zip(ref_model.state_dict().items(), cl_model.state_dict().items()):
        assert v1.shape == v2.shape and rel_fro(v2, v1) <= 1e-3, (k1, rel_fro(v2, v1))


# ---- K1e: batch call (stream-K groups) and the packed small-C path -------------------------------------------------
PACK_GEOMS = [
    # N, C, H, W, kernel, stride, padding      (vertical stride 2, C <= 4, kw <= 8)
    (2, 3, 32, 32, (7, 7), (2, 2), (3, 3)),      # the ResNet stem's geometry on a small image
    (3, 3, 17, 23, (7, 7), (2, 2), (3, 3)),      # odd sizes: ragged last rows / columns
    (2, 1, 12, 12, (5, 5), (2, 2), (2, 2)),      # one channel
    (2, 4, 10, 14, (3, 8), (2, 1), (1, 4)),      # C = 4, widest filter, horizontal stride 1
    (5, 2, 9, 9, (2, 3), (2, 3), (0, 0)),        # kh = 2 (one row pair), no padding, horizontal stride 3
]


@pytest.mark.parametrize("geom", PACK_GEOMS)
@pytest.mark.parametrize("nchw", [False, True])
def test_packed_small_c_conv_bit_exact_on_integers(geom, nchw):
    """Small-C convolutions go through the pack pre-pass (row-parity x 8 taps x 4 channels = 64 packed channels) and the
    TMA-fed kernel as a ceil(kh/2) x 1 convolution; the reduction drops the padding rows.  Integer inputs: exact."""
    N, C, H, W, k, s, p = geom
    gen = torch.Generator().manual_seed(hash(geom) % (2 ** 31))
    x = torch.randint(-2, 3, (N, C, H, W), generator=gen).float()
    want, R = oracle_A(x, k, s, p, False)
    K = want.shape[0]
    out = torch.zeros(K, K, device=DEV)
    xd = x.to(DEV)
    if not nchw:
        xd = xd.contiguous(memory_format=torch.channels_last)
    item = nat.nhwc_item(xd, k, s, p, False, 1.0, out, nat.PREC_BF16)
    assert item is not None and item.nchw == int(nchw or C == 1)
    nat.syrk_batch_nhwc([item], nat.PREC_BF16, xd.device)
    assert torch.equal(out.cpu().double(), want), f"max diff {(out.cpu().double() - want).abs().max()}"
    nat.syrk_batch_nhwc([nat.nhwc_item(xd, k, s, p, False, 2.0, out, nat.PREC_BF16)], nat.PREC_BF16, xd.device)
    assert torch.equal(out.cpu().double(), 3 * want)


@pytest.mark.parametrize("prec", TC_TIERS)
def test_batch_call_equals_item_by_item_calls(prec):
    """crv_syrk_batch_nhwc over a mixed bag of operands (grouped read-once factors, multi-block factors with their own
    launch, more items than one group holds) gives exactly the per-item results on integer inputs."""
    gen = torch.Generator().manual_seed(11)
    specs = [(2, 32, 8, 8, (3, 3), (1, 1), (1, 1)), (2, 64, 7, 7, (1, 1), (1, 1), (0, 0)), (3, 288, 5, 5, (1, 1), (1, 1), (0, 0)),
             (2, 128, 9, 9, (1, 1), (2, 2), (0, 0)), (2, 96, 12, 16, (3, 3), (1, 1), (1, 1)), (4, 40, 6, 6, (1, 1), (1, 1), (0, 0)),
             (2, 512, 4, 4, (1, 1), (1, 1), (0, 0))]
    specs = specs + [(3, 64 + 32 * (i % 5), 5, 6, (1, 1), (1, 1), (0, 0)) for i in range(60)]    # > GRP_MAXF grouped items
    items, outs, wants, keep = [], [], [], []
    for (N, C, H, W, k, s, p) in specs:
        x = torch.randint(-2, 3, (N, C, H, W), generator=gen).float()
        want, _ = oracle_A(x, k, s, p, False)
        out = torch.zeros(want.shape[0], want.shape[0], device=DEV)
        xd = x.to(DEV).contiguous(memory_format=torch.channels_last)
        item = nat.nhwc_item(xd, k, s, p, False, 1.0, out, prec)
        assert item is not None
        items.append(item); outs.append(out); wants.append(want); keep.append(xd)
    for shape in [(5, 64), (300, 36, 2, 2), (7, 1000, 1, 1)]:
        g = torch.randint(-3, 4, shape, generator=gen).float()
        M = shape[1]
        X = g.reshape(shape[0], M, -1).permute(1, 0, 2).reshape(M, -1).double()
        gd = g.to(DEV)
        if gd.dim() == 4:
            gd = gd.contiguous(memory_format=torch.channels_last)
        out = torch.zeros(M, M, device=DEV)
        item = nat.nhwc_item(gd, None, None, None, False, 1.0, out, prec)
        assert item is not None
        items.append(item); outs.append(out); wants.append(X @ X.t()); keep.append(gd)
    nat.syrk_batch_nhwc(items, prec, DEV)
    for i, (out, want) in enumerate(zip(outs, wants)):
        assert torch.equal(out.cpu().double(), want), f"item {i}: max diff {(out.cpu().double() - want).abs().max()}"
    nat.syrk_batch_nhwc(items, prec, DEV)       # accumulates
    for out, want in zip(outs, wants):
        assert torch.equal(out.cpu().double(), 2 * want)


def test_dense_batch_one_launch_bit_exact_on_integers():
    """K1f (crv_syrk_batch_dense): every factor of a small model in ONE CUDA-core launch -- convolutions with bias rows,
    ragged geometries, rows operands, more items than one launch's table holds -- exact on integer inputs, accumulating."""
    gen = torch.Generator().manual_seed(13)
    items, outs, wants, keep = [], [], [], []
    for (N, C, H, W, k, s, p, bias) in GEOMS + [(100, 1, 28, 28, (5, 5), (1, 1), (2, 2), True), (100, 6, 14, 14, (5, 5), (1, 1), (0, 0), True)]:
        x = torch.randint(-2, 3, (N, C, H, W), generator=gen).float()
        want, _ = oracle_A(x, k, s, p, bias)
        out = torch.zeros(want.shape[0], want.shape[0], device=DEV)
        xd = x.to(DEV)
        it = nat.dense_item(xd, k, s, p, bias, 1.0, out)
        assert it is not None
        items.append(it); outs.append(out); wants.append(want); keep.append(xd)
    shapes = [(100, 400), (100, 120), (100, 84), (100, 10), (100, 16, 10, 10), (100, 6, 28, 28), (1, 7)]
    shapes += [(3 + i % 4, 5 + 7 * (i % 9)) for i in range(70)]                    # > 64 items: two launches
    for i, shape in enumerate(shapes):
        g = torch.randint(-3, 4, shape, generator=gen).float()
        M = shape[1]
        bias = (i % 2 == 0) and len(shape) == 2
        X = g.reshape(shape[0], M, -1).permute(1, 0, 2).reshape(M, -1).double()
        if bias:
            X = torch.cat([X, torch.ones_like(X[:1])], 0)
        gd = g.to(DEV)
        out = torch.zeros(X.shape[0], X.shape[0], device=DEV)
        it = nat.dense_item(gd, None, None, None, bias, 1.0, out)
        assert it is not None
        items.append(it); outs.append(out); wants.append(X @ X.t()); keep.append(gd)
    before = nat.launch_calls
    nat.syrk_batch_dense(nat.syrk_dense_array(items), len(items), DEV)
    assert nat.launch_calls == before + 1
    for i, (out, want) in enumerate(zip(outs, wants)):
        assert torch.equal(out.cpu().double(), want), f"item {i}: max diff {(out.cpu().double() - want).abs().max()}"
    nat.syrk_batch_dense(nat.syrk_dense_array(items), len(items), DEV)       # accumulates
    for out, want in zip(outs, wants):
        assert torch.equal(out.cpu().double(), 2 * want)
    assert nat.dense_item(keep[0].permute(0, 1, 3, 2), (3, 3), (1, 1), (1, 1), True, 1.0, outs[0]) is None      # not dense


def test_lenet_update_is_one_launch():
    """BASELINE configs[0]: the ten factors of a LeNet-5 update (C = 1 / 6 convolutions, bias rows) are one K1f launch
    (plus nothing else), whatever the tier, and equal the per-factor fp32 path to 1e-6."""
    import torch.nn as nn
    torch.manual_seed(0)
    model = nn.Sequential(nn.Conv2d(1, 6, 5, padding=2), nn.ReLU(), nn.MaxPool2d(2), nn.Conv2d(6, 16, 5), nn.ReLU(), nn.MaxPool2d(2),
                          nn.Flatten(), nn.Linear(400, 120), nn.ReLU(), nn.Linear(120, 84), nn.ReLU(), nn.Linear(84, 10)).to(DEV)
    x = torch.randn(100, 1, 28, 28, device=DEV)
    states = {}
    for prec in ("fp32", "bf16x3", "bf16"):
        kfac = cb.KFAC(model, precision=prec)
        out = model(x)
        lab = torch.distributions.Categorical(logits=out).sample()
        torch.manual_seed(1)
        F.cross_entropy(model(x), lab).backward()
        before = nat.launch_calls
        kfac.update(100)
        assert nat.launch_calls - before == 1, (prec, nat.launch_calls - before)
        states[prec] = [[f.clone() for f in v] for v in kfac.state.values()]
        for h in kfac.hooks:
            h.remove()
        model.zero_grad()
    import curvature_b200.curvatures as cv
    old, cv._DENSE_BATCH_FLOPS = cv._DENSE_BATCH_FLOPS, 0.0          # per-factor path
    try:
        kfac = cb.KFAC(model, precision="fp32")
        out = model(x)
        F.cross_entropy(out, torch.distributions.Categorical(logits=out).sample()).backward()
        kfac.update(100)
    finally:
        cv._DENSE_BATCH_FLOPS = old
    for h in kfac.hooks:
        h.remove()
    # (different sampled labels between the passes: compare the label-independent A factors)
    for a, b in zip(states["fp32"], [[f for f in v] for v in kfac.state.values()]):
        assert rel_fro(a[0], b[0]) <= 1e-6
    for a, b in zip(states["fp32"], states["bf16"]):
        assert rel_fro(a[0], b[0]) <= 1e-6


def test_resnet_stem_factor_against_fp64():
    """The 3 -> 64, 7x7, stride-2 stem at 224^2 through the packed path (bf16 operands, stated 1e-3 tier)."""
    torch.manual_seed(5)
    x = torch.randn(8, 3, 224, 224, device=DEV)
    cols = F.unfold(x.double(), 7, padding=3, stride=2)
    X = cols.permute(1, 0, 2).reshape(cols.shape[1], -1)
    want = (X @ X.t()) / X.shape[1]
    for xd in (x, x.contiguous(memory_format=torch.channels_last)):
        out = torch.zeros(147, 147, device=DEV)
        item = nat.nhwc_item(xd, (7, 7), (2, 2), (3, 3), False, 1.0 / X.shape[1], out, nat.PREC_BF16)
        assert item is not None
        nat.syrk_batch_nhwc([item], nat.PREC_BF16, xd.device)
        err = rel_fro(out, want)
        assert err <= FACTOR_TOL[nat.PREC_BF16], err
        assert torch.equal(out, out.t())


def test_diag_accum_batch_matches_per_layer_calls():
    """crv_diag_accum_batch (one launch for all parameter groups) == crv_diag_accum layer by layer, bit for bit."""
    torch.manual_seed(9)
    shapes = [(6, 25, True), (16, 150, True), (7, 64, False), (1000, 2048, True), (3, 1, True), (64, 147, False), (512, 4608, False),
              (5, 3, False), (2048, 512, False)]
    entries, wants_s, wants_g = [], [], []
    for M, K0, bias in shapes:
        wg = torch.randn(M, K0, device=DEV)
        bg = torch.randn(M, device=DEV) if bias else None
        K = K0 + int(bias)
        s0 = torch.rand(M, K, device=DEV)
        s_ref, g_ref = s0.clone(), torch.empty(M, K, device=DEV)
        nat.diag_accum(wg, bg, 37.0, state=s_ref, grads_out=g_ref)
        s_new, g_new = s0.clone(), torch.empty(M, K, device=DEV)
        entries.append((wg, bg, s_new, g_new))
        wants_s.append(s_ref); wants_g.append(g_ref)
    nat.diag_accum_batch(entries, 37.0)
    for (wg, bg, s_new, g_new), s_ref, g_ref in zip(entries, wants_s, wants_g):
        assert torch.equal(s_new, s_ref) and torch.equal(g_new, g_ref)
    # state only / grads only
    wg = torch.randn(40, 9, device=DEV)
    s1, s2 = torch.zeros(40, 9, device=DEV), torch.zeros(40, 9, device=DEV)
    g2 = torch.empty(40, 9, device=DEV)
    nat.diag_accum_batch([(wg, None, s1, None), (wg, None, None, g2)], 2.0)
    assert torch.equal(s1, 2.0 * wg * wg) and torch.equal(g2, wg)


@pytest.mark.parametrize("cls", ["KFAC", "EFB"])
def test_sample_and_replace_batched_draw_equals_layer_by_layer(cls, golden):
    """sample_and_replace goes through ONE batched C-ABI call (crv_sample_matrix_normal_batch, stream pool) and, without
    caller noise, one flat randn for the whole model: with the same noise it must equal the layer-by-layer path
    (`sample(layer, noise)` + mean), and without noise it must produce finite, fresh parameters each call."""
    import curvature_b200 as cb
    g = golden("convzoo")
    model = model_from_golden("convzoo", g, DEV)
    layers = selected_layers(model)
    kfac = cb.KFAC(model)
    for b in range(n_batches(g)):
        x = torch.from_numpy(g[f"x/{b}"]).to(DEV)
        logits = model(x)
        labels = torch.from_numpy(g[f"labels/{b}"]).to(DEV)
        loss = torch.nn.functional.cross_entropy(logits, labels)
        model.zero_grad()
        loss.backward()
        kfac.update(x.shape[0])
    est = kfac
    if cls == "EFB":
        est = cb.EFB(model, kfac.state)
        est.update(4)
    est.invert(1e-2, 1.0)
    torch.manual_seed(4)
    noise = {}
    for l in layers:
        K = l.weight[0].numel() + (l.bias is not None)
        noise[l] = torch.randn(K, l.weight.shape[0], device=DEV)
    mean = {k: v.clone() for k, v in model.state_dict().items()}
    est.sample_and_replace(noise={k: v.clone() for k, v in noise.items()})
    names = {id(v): k for k, v in model.state_dict(keep_vars=True).items()}
    for l in layers:
        s = est.sample(l, noise[l].clone())
        K0 = l.weight[0].numel()
        want_w = mean[names[id(l.weight)]] + s[:, :K0].reshape(l.weight.shape)
        assert torch.allclose(l.weight.data, want_w, rtol=1e-5, atol=1e-6)
        if l.bias is not None:
            assert torch.allclose(l.bias.data, mean[names[id(l.bias)]] + s[:, K0], rtol=1e-5, atol=1e-6)
    est.sample_and_replace()
    w1 = [l.weight.data.clone() for l in layers]
    est.sample_and_replace()
    for l, a in zip(layers, w1):
        assert torch.isfinite(l.weight.data).all() and not torch.equal(l.weight.data, a)


# ---- BASELINE full sizes (ResNet-50, batch 256): size-independent identities --------------------------------------
# every unique convolution geometry of torchvision's ResNet-50: C_in, H(=W), k, stride, padding, C_out
RESNET50_FULL = [(3, 224, 7, 2, 3, 64), (64, 56, 1, 1, 0, 64), (64, 56, 3, 1, 1, 64), (64, 56, 1, 1, 0, 256), (256, 56, 1, 1, 0, 64),
                 (256, 56, 1, 1, 0, 128), (128, 56, 3, 2, 1, 128), (128, 28, 1, 1, 0, 512), (256, 56, 1, 2, 0, 512),
                 (512, 28, 1, 1, 0, 128), (128, 28, 3, 1, 1, 128), (512, 28, 1, 1, 0, 256), (256, 28, 3, 2, 1, 256),
                 (256, 14, 1, 1, 0, 1024), (512, 28, 1, 2, 0, 1024), (1024, 14, 1, 1, 0, 256), (256, 14, 3, 1, 1, 256),
                 (1024, 14, 1, 1, 0, 512), (512, 14, 3, 2, 1, 512), (512, 7, 1, 1, 0, 2048), (1024, 14, 1, 2, 0, 2048),
                 (2048, 7, 1, 1, 0, 512), (512, 7, 3, 1, 1, 512)]


@pytest.mark.parametrize("geom", RESNET50_FULL, ids=[f"{g[0]}-{g[5]}ch_{g[1]}px_k{g[2]}s{g[3]}" for g in RESNET50_FULL])
def test_full_size_resnet50_factors_row_sum_and_trace_identities(geom):
    """At the benchmark's own sizes (N = 256) the oracle's unfold + matmul is out of reach, but two linear identities
    of A = X X^T / R pin EVERY row of the factor against an independent computation in fp64:
      row sums   A 1 = X s / R with s = X^T 1, i.e. s = conv2d(x, ones) and X s = conv2d_weight(x, s)   (tap by tap,
                 padding and stride included: an indexing or partition error in any tile changes some row sum),
      trace      tr A = sum_k ||X_k||^2 / R = conv2d_weight(x^2, ones),
    and the same for G = g g^T N^2 / R.  Both factors go through ONE crv_syrk_batch_nhwc call (bench tier, bf16)."""
    C, H, k, s, p, M = geom
    N = 256
    torch.manual_seed(C * 1000 + H + k)
    x = torch.relu(torch.randn(N, C, H, H, device=DEV)).contiguous(memory_format=torch.channels_last)
    OH = (H + 2 * p - k) // s + 1
    g = (torch.randn(N, M, OH, OH, device=DEV) * 1e-2).contiguous(memory_format=torch.channels_last)
    K, R = C * k * k, N * OH * OH
    A = torch.zeros(K, K, device=DEV)
    G = torch.zeros(M, M, device=DEV)
    items = [nat.nhwc_item(x, (k, k), (s, s), (p, p), False, 1.0 / R, A, nat.PREC_BF16),
             nat.nhwc_item(g, None, None, None, False, float(N) * N / R, G, nat.PREC_BF16)]
    assert all(it is not None for it in items)
    nat.syrk_batch_nhwc(items, nat.PREC_BF16, x.device)
    assert torch.equal(A, A.t()) and torch.equal(G, G.t())
    xd = x.double()
    ones = torch.ones(1, C, k, k, device=DEV, dtype=torch.float64)
    svec = F.conv2d(xd, ones, stride=s, padding=p)                                   # (N, 1, OH, OW): column sums of X
    rows = torch.nn.grad.conv2d_weight(xd, (1, C, k, k), svec, stride=s, padding=p).reshape(-1) / R
    assert rel_fro(A.double().sum(1), rows) <= 1e-3, rel_fro(A.double().sum(1), rows)
    tr = torch.nn.grad.conv2d_weight(xd * xd, (1, C, k, k), torch.ones_like(svec), stride=s, padding=p).sum() / R
    assert abs(A.double().trace().item() - tr.item()) <= 1e-3 * abs(tr.item())
    gd = g.double()
    sg = gd.sum(1, keepdim=True)
    rows_g = (gd * sg).sum((0, 2, 3)) * (float(N) * N / R)
    assert rel_fro(G.double().sum(1), rows_g) <= 1e-3, rel_fro(G.double().sum(1), rows_g)
    tr_g = (gd * gd).sum() * (float(N) * N / R)
    assert abs(G.double().trace().item() - tr_g.item()) <= 1e-3 * abs(tr_g.item())


def test_update_with_stream_overlap_equals_serialised_update():
    """The whole KFAC.update pipeline (pre-passes, contractions and reductions of ~100 launches on four prioritised
    streams over a double-buffered workspace) against the same update with everything serialised on one stream
    (crv_profile_enable switches the side streams off): the reductions are fixed-order, so the factors must be
    BIT-identical -- any workspace hazard between overlapping launches shows up as a difference.  ResNet-50 at batch
    64: long tap-aware reductions (2304^2, 4608^2 factors) followed by short launches, the pattern that once let a
    pre-pass overwrite partial tiles that were still being reduced (CURVATURE_B200_FAULT_COPY_LAYOUT=1 re-injects that
    layout: this test then fails on the symmetry check).  Also: every factor symmetric and, damped, PD."""
    import torchvision
    import curvature_b200 as cb
    torch.manual_seed(0)
    model = torchvision.models.resnet50(weights=None).to(DEV).train().to(memory_format=torch.channels_last)
    x = torch.randn(64, 3, 224, 224, device=DEV).contiguous(memory_format=torch.channels_last)
    a = cb.KFAC(model, precision="bf16")
    orc.fisher_step(model, x)
    for _ in range(3):                      # several back-to-back updates: the pipeline is in steady state
        a.update(64)
    torch.cuda.synchronize()
    b = cb.KFAC(model, precision="bf16")
    b.record = a.record
    nat.profile_enable(True)
    try:
        for _ in range(3):
            b.update(64)
        torch.cuda.synchronize()
    finally:
        nat.profile_collect()
        nat.profile_enable(False)
    worst = 0
    for (la, fa), (lb, fb) in zip(a.state.items(), b.state.items()):
        for u, v in zip(fa, fb):
            assert torch.equal(u, u.t())
            if not torch.equal(u, v):
                worst = max(worst, float((u - v).abs().max()))
    assert worst == 0, f"overlapped and serialised updates differ by up to {worst}"
    a.invert(add=1.0, multiply=1.0)         # sqrt(s) F + sqrt(n) I is PD for every PSD factor
    for h in a.hooks + b.hooks:
        h.remove()
