"""Where a CTA of the chain kernel spends its time (crv_debug_timeline buffer, per CTA: cycles waiting for dependencies,
for the accumulator drain, for operands, for the MMAs, draining; tiles; total)."""
import os, sys
import numpy as np
import torch, torchvision
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curvature_b200 import _native as nat  # noqa: E402
dev = "cuda:0"
torch.manual_seed(0)
model = torchvision.models.resnet50(weights=None)
shapes = [(m.weight.shape[0], m.weight[0].numel() + (m.bias is not None)) for m in model.modules()
          if m.__class__.__name__ in ("Conv2d", "Linear")]
efb = []
for M, K in shapes:
    efb.append((nat.round_tf32(torch.randn(M, M, device=dev) / M ** 0.5), nat.round_tf32(torch.randn(K, K, device=dev) / K ** 0.5),
                nat.round_tf32(torch.randn(M, K, device=dev)), torch.zeros(M, K, device=dev)))
for _ in range(2):
    nat.efb_project_batch(efb, nat.PREC_TF32, round_g=False)
torch.cuda.synchronize()
buf = torch.zeros(160 * 8, dtype=torch.int64, device=dev)
nat.debug_timeline(buf)
nat.efb_project_batch(efb, nat.PREC_TF32, round_g=False)
torch.cuda.synchronize()
nat.debug_timeline(None)
a = buf.view(160, 8).cpu().numpy()[:148]
names = ["dep wait (producer 0)", "MMA waits for drain", "MMA waits for operands", "epilogue waits for MMAs", "drain", "tiles", "total"]
for i, n in enumerate(names):
    print(f"{n:28s} min {a[:, i].min():9d}  median {int(np.median(a[:, i])):9d}  max {a[:, i].max():9d}  mean {a[:, i].mean():11.0f}")
