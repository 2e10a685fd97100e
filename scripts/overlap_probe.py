"""Does the library's stream pipeline (pre-pass i+1 || contraction i || reduction i-1) really overlap?  Times n
back-to-back channels-last SYRK calls between a fork and a join against the serialised per-kernel times."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curvature_b200 import _native as nat  # noqa: E402

dev = "cuda:0"
prec = nat.PREC_BF16
for name, (N, C, H, k, p) in {"a2304": (256, 256, 14, 3, 1), "a1024": (256, 1024, 14, 1, 0), "a576": (256, 64, 56, 3, 1)}.items():
    xs = [torch.relu(torch.randn(N, C, H, H, device=dev)).contiguous(memory_format=torch.channels_last) for _ in range(2)]
    K = C * k * k
    outs = [torch.zeros(K, K, device=dev) for _ in range(2)]
    R = N * H * H
    def run(n, fork):
        if fork:
            nat.stream_fork(dev)
        for i in range(n):
            nat.syrk_conv_accum(xs[i % 2], (k, k), (1, 1), (p, p), False, 1.0 / R, outs[i % 2], prec, join=False)
        nat.stream_join(dev)
    run(4, True)
    torch.cuda.synchronize()
    res = {}
    for fork in (False, True):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(20, fork)
        e1.record()
        torch.cuda.synchronize()
        res[fork] = e0.elapsed_time(e1) / 20 * 1e3
    nat.profile_enable(True)
    nat.profile_collect()
    run(4, True)
    torch.cuda.synchronize()
    prof = nat.profile_collect()
    nat.profile_enable(False)
    ser = {k_: v["ms"] / v["launches"] * 1e3 for k_, v in prof.items() if v["launches"]}
    print(f"{name}: per call {res[False]:.1f} us without fork, {res[True]:.1f} us forked; serialised kernels {ser} sum {sum(ser.values()):.1f} us")
