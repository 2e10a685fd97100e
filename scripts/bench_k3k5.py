"""Time K3 (EFB.update: eigenbasis projection) and K5 (KFAC.sample_and_replace: matrix-normal draw) on ResNet-50, per tier.

    python scripts/bench_k3k5.py [out.json]
Random matrices stand in for the eigenbases / inverse factors (timing only; parity is in tests/)."""
import json
import os
import sys
import time

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import curvature_b200 as cb  # noqa: E402
from curvature_b200 import _native as nat  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = "cuda:0"
    torch.manual_seed(0)
    model = torchvision.models.resnet50(weights=None).to(dev).train()
    layers = [m for m in model.modules() if m.__class__.__name__ in ("Conv2d", "Linear")]
    for p in model.parameters():
        p.grad = torch.randn_like(p) * 1e-3
    flops = 0
    eig, inv = {}, {}
    for l in layers:
        M = l.weight.shape[0]
        K = l.weight[0].numel() + (l.bias is not None)
        flops += 2 * M * K * (K + M)
        eig[l] = (torch.randn(K, K, device=dev) / K ** 0.5, torch.randn(M, M, device=dev) / M ** 0.5)
        inv[l] = (torch.tril(torch.randn(K, K, device=dev)) / K ** 0.5, torch.tril(torch.randn(M, M, device=dev)) / M ** 0.5)
    out = {"flops_per_call": flops}
    for tier in ("fp32", "tf32"):
        efb = cb.EFB(model, None, precision=tier, eigvecs=eig)
        ms = timed(lambda: efb.update(256))
        out[f"efb_update_{tier}_ms"] = ms
        out[f"efb_update_{tier}_tflops"] = flops / ms / 1e9
        kfac = cb.KFAC(model, precision=tier)
        kfac.inv_state = dict(inv)
        ms = timed(lambda: kfac.sample_and_replace())
        out[f"kfac_sample_{tier}_ms"] = ms
        out[f"kfac_sample_{tier}_tflops"] = flops / ms / 1e9
        for h in kfac.hooks:
            h.remove()
    print(json.dumps(out))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
