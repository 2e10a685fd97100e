"""Host enqueue time vs device time of KFAC.update (is the update launch-bound on the CPU?)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from curvature_b200 import KFAC  # noqa: E402

model = bench.make_model("resnet50")[0].cuda().train().to(memory_format=torch.channels_last)
kfac = KFAC(model, precision="bf16")
x = torch.randn(256, 3, 224, 224, device="cuda").contiguous(memory_format=torch.channels_last)
bench.fisher_step(model, x)
for _ in range(3):
    kfac.update(256)
torch.cuda.synchronize()
host, dev = [], []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    kfac.update(256)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    host.append((t1 - t0) * 1e3)
    dev.append(e0.elapsed_time(e1))
print("host enqueue ms:", [round(h, 2) for h in host])
print("device ms      :", [round(d, 2) for d in dev])
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
kfac.update(256)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
