"""Per-CTA timeline of the big read-once group launch of a ResNet-50 KFAC.update (CURVATURE_B200_TL_MIN_NF=40)."""
import os, sys
os.environ.setdefault("CURVATURE_B200_TL_MIN_NF", "40")
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import curvature_b200 as cb  # noqa: E402
from curvature_b200 import _native as nat  # noqa: E402

dev = "cuda:0"
model = bench.make_model("resnet50")[0].to(dev).train().to(memory_format=torch.channels_last)
kfac = cb.KFAC(model, precision="bf16")
x = torch.randn(256, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
bench.fisher_step(model, x)
for _ in range(2):
    kfac.update(256)
torch.cuda.synchronize()
tl = torch.zeros(160 * 8, dtype=torch.int64, device=dev)
nat.debug_timeline(tl)
kfac.update(256)
torch.cuda.synchronize()
nat.debug_timeline(None)
a = tl.view(160, 8).cpu().numpy().astype(np.int64)
a = a[a[:, 0] > 0]
t0 = a[:, 0].min()
end = (a[:, 4] - t0) / 1e3
mma = (a[:, 3] - a[:, 2]) / 1e3
print(f"{len(a)} CTAs; drained (us) min/p10/med/p90/max: {end.min():.0f}/{np.percentile(end,10):.0f}/{np.median(end):.0f}/{np.percentile(end,90):.0f}/{end.max():.0f}")
print(f"segments per CTA min/med/max: {a[:,5].min()}/{int(np.median(a[:,5]))}/{a[:,5].max()}; k-groups min/med/max {a[:,6].min()}/{int(np.median(a[:,6]))}/{a[:,6].max()}")
order = np.argsort(end)
print("earliest 5 CTAs (cta, end us, segs, kgroups):", [(int(i), round(float(end[i])), int(a[i,5]), int(a[i,6])) for i in order[:5]])
print("latest   5 CTAs (cta, end us, segs, kgroups):", [(int(i), round(float(end[i])), int(a[i,5]), int(a[i,6])) for i in order[-5:]])
single = a[a[:, 5] == 1]
dur = (single[:, 4] - single[:, 2]) / 1e3
kg = single[:, 6]
import collections
buckets = collections.defaultdict(list)
for k, d in zip(kg, dur):
    buckets[int(round(k / 500.0)) * 500].append(d * 1e3 / k)
print("single-segment CTAs: k-groups bucket -> (CTAs, ns per k-group):", {k: (len(v), round(float(np.median(v)), 1)) for k, v in sorted(buckets.items())})
