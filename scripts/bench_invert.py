"""KFAC.invert (K4: batched damped Cholesky of the inverse) on ResNet-50 / ResNet-152 after one update: ms per invert."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import curvature_b200 as cb  # noqa: E402

dev = "cuda:0"
out = {}
for name, batch in (("resnet50", 256), ("resnet152", 128)):
    model = bench.make_model(name)[0].to(dev).train().to(memory_format=torch.channels_last)
    kfac = cb.KFAC(model, precision="bf16")
    x = torch.randn(batch, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
    bench.fisher_step(model, x)
    kfac.update(batch)
    torch.cuda.synchronize()
    dims = [f.shape[0] for fs in kfac.state.values() for f in fs]
    ms = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        kfac.invert(add=1.0, multiply=10.0)
        torch.cuda.synchronize()
        ms.append((time.perf_counter() - t0) * 1e3)
    flops = sum(2.0 * d ** 3 / 3 * 3 for d in dims)     # factorisation + triangular inverse + product, ~2 D^3
    out[name] = {"invert_ms": min(ms), "matrices": len(dims), "largest": max(dims), "approx_gflop": flops / 1e9}
    print(name, out[name], flush=True)
    del kfac, model, x
    torch.cuda.empty_cache()
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
