"""Launch-level trace of one ResNet KFAC.update WITH the stream overlap on (crv_debug_trace): for every launch of the
channels-last SYRK path the device-clock [first CTA start, last CTA end] of its pre-pass, contraction and reduction.

    python scripts/step_trace.py [--model resnet50] [--batch 256] [--precision bf16] [--out gpurun_out/trace.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import curvature_b200 as cb  # noqa: E402
from curvature_b200 import _native as nat  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="resnet50")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--layout", default="channels_last")
ap.add_argument("--out", default=None)
args = ap.parse_args()

dev = "cuda:0"
model = bench.make_model(args.model)[0].to(dev).train()
x = torch.randn(args.batch, 3, 224, 224, device=dev)
if args.layout == "channels_last":
    model = model.to(memory_format=torch.channels_last)
    x = x.contiguous(memory_format=torch.channels_last)
kfac = cb.KFAC(model, precision=args.precision)
bench.fisher_step(model, x)
for _ in range(3):
    kfac.update(args.batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
buf = nat.debug_trace(256, dev)
e0.record()
kfac.update(args.batch)
e1.record()
torch.cuda.synchronize()
n = nat.debug_trace_count()
nat.debug_trace(0)
a = buf[:n].cpu().numpy().astype(np.uint64)
NONE = np.uint64(0xFFFFFFFFFFFFFFFF)
starts = [a[i, w] for i in range(n) for w in (0, 2, 4) if a[i, w] != NONE]
t0 = min(starts)
rows = []


def span(i, w):
    if a[i, w] == NONE:
        return None
    return (float(a[i, w] - t0) / 1e3, float(a[i, w + 1] - t0) / 1e3)


print(f"{args.model} batch {args.batch} {args.precision} {args.layout}: update {e0.elapsed_time(e1):.3f} ms (event), {n} launches")
print(f"{'#':>3} {'D':>5} {'nf':>3} {'bf16':>4} {'pairs':>5} | {'prepass us':>19} | {'contraction us':>19} {'dur':>6} | {'reduction us':>19} {'dur':>6} | gap")
prev_end = 0.0
busy = {"pre": 0.0, "main": 0.0, "red": 0.0}
for i in range(n):
    D = int(a[i, 6] & np.uint64(0xFFFFF))
    nf = int((a[i, 6] >> np.uint64(20)) & np.uint64(0x3FF))
    bf = int((a[i, 6] >> np.uint64(30)) & np.uint64(1))
    pre, main, red = span(i, 4), span(i, 0), span(i, 2)
    f = lambda s: f"{s[0]:9.1f}-{s[1]:9.1f}" if s else " " * 19   # noqa: E731
    gap = main[0] - prev_end if main else 0.0
    prev_end = main[1] if main else prev_end
    for k, s in (("pre", pre), ("main", main), ("red", red)):
        if s:
            busy[k] += s[1] - s[0]
    print(f"{i:3d} {D:5d} {nf:3d} {bf:4d} {int(a[i, 7]):5d} | {f(pre)} | {f(main)} {main[1]-main[0] if main else 0:6.1f} | {f(red)} {red[1]-red[0] if red else 0:6.1f} | {gap:6.1f}")
    rows.append({"launch": i, "D": D, "factors": nf, "bf16": bf, "pairs": int(a[i, 7]), "prepass_us": pre, "contraction_us": main,
                 "reduction_us": red})
ends = [float(a[i, w + 1] - t0) / 1e3 for i in range(n) for w in (0, 2, 4) if a[i, w] != NONE]
print(f"span first start -> last end: {max(ends)/1e3:.3f} ms; summed kernel spans: pre-pass {busy['pre']/1e3:.3f} ms, "
      f"contraction {busy['main']/1e3:.3f} ms, reduction {busy['red']/1e3:.3f} ms")
if args.out:
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump({"model": args.model, "batch": args.batch, "precision": args.precision, "layout": args.layout,
               "update_ms": e0.elapsed_time(e1), "launches": rows}, open(args.out, "w"), indent=1)
