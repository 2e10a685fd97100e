"""Diagonal.update on ResNet-50 (K2): ms per update and achieved HBM GB/s (12 bytes per parameter: read grad, RMW state)."""
import json, os, sys
import torch, torchvision
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import curvature_b200 as cb  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
model = torchvision.models.resnet50(weights=None).to(dev).train()
for p in model.parameters():
    p.grad = torch.randn_like(p) * 1e-3
diag = cb.Diagonal(model)
nparam = sum(m.weight.numel() + (m.bias.numel() if m.bias is not None else 0) for m in model.modules()
             if m.__class__.__name__ in ("Conv2d", "Linear"))
for _ in range(3):
    diag.update(256)
torch.cuda.synchronize()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)     # > L2: the state must come from HBM
ms = []
for i in range(10):
    flush.fill_(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    diag.update(256)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms.sort()
med = ms[len(ms) // 2]
out = {"diagonal_update_ms": med, "parameters": nparam, "bytes": 12 * nparam, "gbs": 12 * nparam / med / 1e6,
       "note": "L2 flushed between updates (512 MB fill)"}
print(json.dumps(out))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
