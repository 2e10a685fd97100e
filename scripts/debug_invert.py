import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import curvature_b200 as cb
from curvature_b200 import _native as nat
dev = "cuda:0"
model = bench.make_model("resnet50")[0].to(dev).train().to(memory_format=torch.channels_last)
kfac = cb.KFAC(model, precision="bf16")
x = torch.randn(256, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
bench.fisher_step(model, x)
kfac.update(256)
torch.cuda.synchronize()
fs = [f for pair in kfac.state.values() for f in pair]
for idx in (50, 52, 56, 64, 96, 88, 102):
    F = fs[idx]
    D = F.shape[0]
    sym = bool(torch.equal(F, F.t()))
    fin = bool(torch.isfinite(F).all())
    ev = torch.linalg.eigvalsh(F.double())
    reg = (10.0 ** 0.5) * F.double() + torch.eye(D, device=dev, dtype=torch.float64)
    try:
        torch.linalg.cholesky(reg.float())
        tc = "torch fp32 chol ok"
    except Exception as e:
        tc = "torch fp32 chol FAILED"
    out = torch.empty_like(F)
    info = nat.chol_inv_batched([F], [1.0], [10.0], [out])
    print(idx, "D", D, "sym", sym, "finite", fin, "eig min/max", float(ev[0]), float(ev[-1]), tc, "ours info", info.tolist(),
          "diag min", float(F.diagonal().min()), flush=True)
