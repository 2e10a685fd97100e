"""One launch each of the non-SYRK kernels at ResNet-50 sizes (for ncu captures): K2 diagonal accumulation / inverse
square root / diagonal sample, K3 EFB projection GEMMs, K4 batched Cholesky-of-inverse, K5 matrix-normal draw."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curvature_b200 import _native as nat  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
M, K0 = 512, 4608                       # the 512 -> 512 3x3 convolution
wg = torch.randn(M, K0, device=dev) * 1e-3
state = torch.zeros(M, K0, device=dev)
grads = torch.empty(M, K0, device=dev)
for _ in range(2):
    nat.diag_accum(wg, None, 256, state=state, grads_out=grads)                 # K2
inv = torch.empty_like(state)
nat.elementwise_inv_sqrt(state, 1e-3, 1.0, inv)
w_out = torch.empty(M, K0, device=dev)
nat.diag_sample(torch.randn(M, K0, device=dev), inv, False, mu_w=wg, w_out=w_out)
QA = torch.linalg.qr(torch.randn(K0, K0, device=dev))[0].contiguous()
QG = torch.linalg.qr(torch.randn(M, M, device=dev))[0].contiguous()
lam = torch.zeros(M, K0, device=dev)
for prec in (nat.PREC_TF32,):
    nat.round_tf32(grads, out=grads)
    for _ in range(2):
        nat.efb_project_accum(QG, QA, grads, lam, prec)                         # K3
    LA = torch.tril(torch.randn(K0, K0, device=dev)) / K0 ** 0.5
    LG = torch.tril(torch.randn(M, M, device=dev)) / M ** 0.5
    z = torch.randn(K0, M, device=dev)
    for _ in range(2):
        nat.sample_matrix_normal(LG, LA, z, False, mu_w=wg, w_out=w_out, precision=prec)   # K5
# K4: damped Cholesky of the inverse, a batch of ResNet-50-sized factors
fs = []
for D in (64, 256, 576, 1152, 2304):
    X = torch.randn(D, 2 * D, device=dev)
    fs.append((X @ X.t()) / (2 * D))
outs = [torch.empty_like(f) for f in fs]
info = nat.chol_inv_batched(fs, [1e-2] * len(fs), [1.0] * len(fs), outs)
torch.cuda.synchronize()
print("ok", info.tolist())
