"""ResNet-50: two KFAC.update and two Diagonal.update calls (for ncu captures of the HBM-bound kernels: the read-once
TF32 group launches of the channels-last SYRK and the whole-model diagonal accumulation)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import curvature_b200 as cb  # noqa: E402

dev = "cuda:0"
model = bench.make_model("resnet50")[0].to(dev).train().to(memory_format=torch.channels_last)
kfac = cb.KFAC(model, precision="bf16")
diag = cb.Diagonal(model)
x = torch.randn(256, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
bench.fisher_step(model, x)
for _ in range(2):
    kfac.update(256)
    diag.update(256)
torch.cuda.synchronize()
print("ok")
