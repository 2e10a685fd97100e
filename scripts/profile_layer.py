"""Run single ResNet-shaped SYRK launches through the C ABI (for ncu captures and quick timing).

    python scripts/profile_layer.py [--prec tf32] [--reps 3] name[,name...]
Layer names: see LAYERS below.  Prints the median CUDA-event time and the achieved algorithmic TFLOP/s / GB/s.
"""
import argparse
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curvature_b200 import _native as nat  # noqa: E402

# name: (kind, N, C, H, W, k, s, p)   kind 'A' = conv input factor, 'G' = (N, M, OH, OW) gradient factor
LAYERS = {
    "a2304": ("A", 256, 256, 14, 14, 3, 1, 1),
    "a4608": ("A", 256, 512, 7, 7, 3, 1, 1),
    "a1152": ("A", 256, 128, 28, 28, 3, 1, 1),
    "a576": ("A", 256, 64, 56, 56, 3, 1, 1),
    "a1024": ("A", 256, 1024, 14, 14, 1, 1, 0),
    "a256": ("A", 256, 256, 56, 56, 1, 1, 0),
    "a64": ("A", 256, 64, 56, 56, 1, 1, 0),
    "stem": ("A", 256, 3, 224, 224, 7, 2, 3),
    "g64": ("G", 256, 64, 56, 56, 1, 1, 0),
    "g256": ("G", 256, 256, 56, 56, 1, 1, 0),
    "g1024": ("G", 256, 1024, 14, 14, 1, 1, 0),
    "g2048": ("G", 256, 2048, 7, 7, 1, 1, 0),
    "g256s": ("G", 256, 256, 14, 14, 1, 1, 0),
    "g128": ("G", 256, 128, 28, 28, 1, 1, 0),
    "g512": ("G", 256, 512, 28, 28, 1, 1, 0),
    "a512": ("A", 256, 512, 28, 28, 1, 1, 0),
    "a128": ("A", 256, 128, 28, 28, 1, 1, 0),
    "a2048": ("A", 256, 2048, 7, 7, 1, 1, 0),
    "a1152s2": ("A", 256, 128, 56, 56, 3, 2, 1),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names")
    ap.add_argument("--prec", default="tf32")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--layout", default="channels_last", choices=["channels_last", "nchw"])
    ap.add_argument("--timeline", action="store_true", help="print the per-CTA timeline of one launch")
    args = ap.parse_args()
    prec = nat.PRECISION_NAMES[args.prec]
    dev = "cuda:0"
    for name in args.names.split(","):
        kind, N, C, H, W, k, s, p = LAYERS[name]
        torch.manual_seed(0)
        x = torch.relu(torch.randn(N, C, H, W, device=dev))
        if args.layout == "channels_last":
            x = x.contiguous(memory_format=torch.channels_last)
        if kind == "A":
            K = C * k * k
            OH = (H + 2 * p - k) // s + 1
            R = N * OH * OH
            out = torch.zeros(K, K, device=dev)
            fn = lambda: nat.syrk_conv_accum(x, (k, k), (s, s), (p, p), False, 1.0 / R, out, prec)  # noqa: E731
        else:
            K = C
            R = N * H * W
            out = torch.zeros(K, K, device=dev)
            fn = lambda: nat.syrk_rows_accum(x, False, 1.0 / R, out, prec)  # noqa: E731
        times = []
        fn()
        torch.cuda.synchronize()
        nat.profile_enable(True)
        nat.profile_collect()
        for _ in range(max(args.reps, 1)):
            fn()
        torch.cuda.synchronize()
        prof = nat.profile_collect()
        nat.profile_enable(False)
        kms = "  ".join(f"{k.replace('syrk_', '')}={v['ms'] / v['launches'] * 1e3:.1f}us" for k, v in prof.items() if v["launches"])
        for _ in range(args.reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = statistics.median(times[1:] or times)
        if args.timeline:
            tl = torch.zeros(160 * 8, dtype=torch.int64, device=dev)
            nat.debug_timeline(tl)
            fn()
            torch.cuda.synchronize()
            nat.debug_timeline(None)
            t = tl.view(160, 8).cpu()
            t = t[t[:, 0] > 0]
            t0 = int(t[:, 0].min())
            import numpy as np
            a = t.numpy().astype(np.int64)
            rel = lambda c: (a[:, c] - t0) / 1e3
            print(f"   timeline ({len(a)} CTAs, us from first CTA start): start max {rel(0).max():.1f} | prologue {np.median(rel(1) - rel(0)):.1f} | "
                  f"first data {np.median(rel(2) - rel(1)):.1f} | last MMA issued min/med/max {rel(3).min():.1f}/{np.median(rel(3)):.1f}/{rel(3).max():.1f} | "
                  f"drained min/med/max {rel(4).min():.1f}/{np.median(rel(4)):.1f}/{rel(4).max():.1f} | segments max {a[:, 5].max()} | "
                  f"k-groups min/med/max {a[:, 6].min()}/{int(np.median(a[:, 6]))}/{a[:, 6].max()}")
            dur = rel(3) - rel(2)
            rate = dur * 1e3 / np.maximum(a[:, 6], 1)
            print(f"   ns per k-group min/med/max {rate.min():.1f}/{np.median(rate):.1f}/{rate.max():.1f}; by segments: " +
                  ", ".join(f"{k} seg: {int((a[:, 5] == k).sum())} CTAs, MMA phase {np.median(dur[a[:, 5] == k]):.1f} us" for k in sorted(set(a[:, 5]))))
        print(f"{name:6s} K={K:5d} R={R:8d}  {ms:8.3f} ms  {R * K * (K + 1) / ms / 1e9:7.1f} TFLOP/s (algorithmic)  "
              f"{4 * x.numel() / ms / 1e6:7.0f} GB/s (input once)  | {kms}")


if __name__ == "__main__":
    main()
