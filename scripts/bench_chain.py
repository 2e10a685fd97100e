"""GPU time of the K3 / K5 batch calls alone (prebuilt argument lists, many back-to-back calls: the host runs ahead), chain
kernel vs the per-layer path (CURVATURE_B200_CHAIN=0), ResNet-50 layer shapes; and S stacked samples (K5c).

    python scripts/bench_chain.py [out.json]"""
import json
import os
import sys
import time

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curvature_b200 import _native as nat  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
model = torchvision.models.resnet50(weights=None)
shapes = [(m.weight.shape[0], m.weight[0].numel() + (m.bias is not None)) for m in model.modules()
          if m.__class__.__name__ in ("Conv2d", "Linear")]
flops = sum(2 * M * K * (M + K) for M, K in shapes)
efb, smp, multi = [], [], []
S = 8
for M, K in shapes:
    QG = nat.round_tf32(torch.randn(M, M, device=dev) / M ** 0.5)
    QA = nat.round_tf32(torch.randn(K, K, device=dev) / K ** 0.5)
    efb.append((QG, QA, nat.round_tf32(torch.randn(M, K, device=dev)), torch.zeros(M, K, device=dev)))
    smp.append(dict(LG=QG, LA=QA, z=nat.round_tf32(torch.randn(K, M, device=dev)), has_bias=False,
                    s_out=torch.empty(M, K, device=dev)))
    multi.append((QG, QA, nat.round_tf32(torch.randn(S * K, M, device=dev)), torch.empty(M, S, K, device=dev)))


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / reps * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, host


out = {"flops_per_call": flops, "chain": os.environ.get("CURVATURE_B200_CHAIN", "1")}
ms, host = timed(lambda: nat.efb_project_batch(efb, nat.PREC_TF32, round_g=False))
out.update(efb_batch_ms=ms, efb_batch_host_ms=host, efb_batch_tflops=flops / ms / 1e9)
ms, host = timed(lambda: nat.sample_matrix_normal_batch(smp, nat.PREC_TF32))
out.update(sample_batch_ms=ms, sample_batch_host_ms=host, sample_batch_tflops=flops / ms / 1e9)
ms, host = timed(lambda: nat.sample_matrix_normal_multi(multi, S, nat.PREC_TF32), reps=4)
out.update(sample_multi_S=S, sample_multi_ms=ms, sample_multi_ms_per_sample=ms / S, sample_multi_host_ms=host,
           sample_multi_tflops=S * flops / ms / 1e9)
print(json.dumps(out))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
