"""GPU (-m gpu): the K3 / K5 chain kernel (gemm_chain.cu) -- both GEMMs of every layer of a call in ONE persistent launch,
the second waiting on the first through device-side counters -- through the batched C-ABI calls, against fp64 and, on
integer-valued operands, bit-exactly (any mis-addressed tile, stale intermediate or missed dependency is an exact
mismatch)."""
import pytest
import torch

from helpers import rel_fro

pytestmark = pytest.mark.gpu

from curvature_b200 import _native as nat        # noqa: E402

DEV = "cuda:0"
TF32 = nat.PREC_TF32

# (M, K0, bias): K = K0 + bias.  TMA-addressable layers (K % 4 == 0, M % 4 == 0) ride in the chain launch; the others
# (147-wide stem, 2049-wide fc with its bias column) keep the per-layer path inside the same call.
LAYERS = [(64, 576, False), (256, 1152, False), (128, 64, False), (512, 2304, False), (1000, 512, False), (64, 64, False),
          (300, 260, False), (2048, 512, False), (64, 147, False), (1000, 2048, True), (8, 4, False), (516, 1028, False)]


def test_efb_projection_batch_against_fp64():
    torch.manual_seed(0)
    entries, wants, starts = [], [], []
    for M, K0, bias in LAYERS:
        K = K0 + bias
        QG = torch.linalg.qr(torch.randn(M, M, device=DEV))[0].contiguous()
        QA = torch.linalg.qr(torch.randn(K, K, device=DEV))[0].contiguous()
        G = torch.randn(M, K, device=DEV)
        lam = torch.rand(M, K, device=DEV)
        starts.append(lam.double().clone())
        wants.append(lam.double() + (QG.double().t() @ G.double() @ QA.double()) ** 2)
        entries.append((nat.round_tf32(QG), nat.round_tf32(QA), G.clone(), lam))
    nat.efb_project_batch(entries, TF32, round_g=True)
    torch.cuda.synchronize()
    for (M, K0, bias), e, w in zip(LAYERS, entries, wants):
        assert rel_fro(e[3], w) <= 1e-3, (M, K0, bias, rel_fro(e[3], w))
    # second call accumulates again (curvatures.py:433 is a +=)
    nat.efb_project_batch(entries, TF32, round_g=True)
    torch.cuda.synchronize()
    for e, w, s0 in zip(entries, wants, starts):
        assert rel_fro(e[3], 2 * w - s0) <= 1e-3


@pytest.mark.parametrize("repeat", [0, 1])
def test_matrix_normal_batch_bit_exact_on_integers(repeat):
    """S = LG z^T LA^T with entries in {-1, 0, 1}: the intermediate (|T| <= M <= 2048 fits TF32 exactly) and the result
    (< 2^24) are exact, so mean + S must match a float64 evaluation bit for bit."""
    gen = torch.Generator().manual_seed(11 + repeat)
    items, wants = [], []
    for M, K0, bias in LAYERS:
        if M > 2048:
            continue
        K = K0 + bias
        LG = torch.randint(-1, 2, (M, M), generator=gen).float().to(DEV)
        LA = torch.randint(-1, 2, (K, K), generator=gen).float().to(DEV)
        z = torch.randint(-1, 2, (K, M), generator=gen).float().to(DEV)
        mu_w = torch.randint(-5, 6, (M, K0), generator=gen).float().to(DEV)
        mu_b = torch.randint(-5, 6, (M,), generator=gen).float().to(DEV) if bias else None
        S = LG.double() @ z.double().t() @ LA.double().t()
        if S.abs().max() >= 2 ** 24:
            continue
        w_out = torch.empty(M, K0, device=DEV)
        b_out = torch.empty(M, device=DEV) if bias else None
        s_out = torch.empty(M, K, device=DEV)
        items.append(dict(LG=LG, LA=LA, z=z, has_bias=bias, mu_w=mu_w, mu_b=mu_b, w_out=w_out, b_out=b_out, s_out=s_out))
        wants.append(S)
    nat.sample_matrix_normal_batch(items, TF32)
    torch.cuda.synchronize()
    for it, S in zip(items, wants):
        K0 = it["mu_w"].shape[1]
        assert torch.equal(it["s_out"].double(), S), (tuple(S.shape), (it["s_out"].double() - S).abs().max().item())
        assert torch.equal(it["w_out"].double(), it["mu_w"].double() + S[:, :K0])
        if it["has_bias"]:
            assert torch.equal(it["b_out"].double(), it["mu_b"].double() + S[:, K0])


def test_matrix_normal_batch_with_row_scale_against_fp64():
    """EFB's draw (curvatures.py:453-460): the noise is scaled by inv_state^T before the two products."""
    torch.manual_seed(3)
    items, wants = [], []
    for M, K0, bias in [(64, 576, False), (256, 1152, False), (512, 260, False), (64, 147, False)]:
        K = K0 + bias
        QG = nat.round_tf32(torch.linalg.qr(torch.randn(M, M, device=DEV))[0].contiguous())
        QA = nat.round_tf32(torch.linalg.qr(torch.randn(K, K, device=DEV))[0].contiguous())
        z = torch.randn(K, M, device=DEV)
        scale = torch.rand(M, K, device=DEV) + 0.5
        s_out = torch.empty(M, K, device=DEV)
        items.append(dict(LG=QG, LA=QA, z=z, has_bias=False, row_scale=scale, s_out=s_out))
        wants.append((QA.double() @ (z.double() * scale.double().t()) @ QG.double().t()).t())
    nat.sample_matrix_normal_batch(items, TF32)
    torch.cuda.synchronize()
    for it, w in zip(items, wants):
        assert rel_fro(it["s_out"], w) <= 1e-3, rel_fro(it["s_out"], w)


def test_chain_launch_equals_per_layer_path(monkeypatch):
    """Same call with the chain kernel switched off (CURVATURE_B200_CHAIN=0 is read once per process, so the per-layer
    result comes from the single-item entry point): identical operand rounding, results within TF32 accumulation-order
    noise of each other."""
    torch.manual_seed(5)
    M, K = 256, 1152
    QG = nat.round_tf32(torch.linalg.qr(torch.randn(M, M, device=DEV))[0].contiguous())
    QA = nat.round_tf32(torch.linalg.qr(torch.randn(K, K, device=DEV))[0].contiguous())
    G = nat.round_tf32(torch.randn(M, K, device=DEV))
    a = torch.zeros(M, K, device=DEV)
    b = torch.zeros(M, K, device=DEV)
    nat.efb_project_batch([(QG, QA, G, a)], TF32, round_g=False)
    nat.efb_project_accum(QG, QA, G, b, TF32)
    assert rel_fro(a, b) <= 1e-5, rel_fro(a, b)
