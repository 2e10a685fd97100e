"""Small-model latency of the whole path (BASELINE configs[0]-[2]: LeNet-5, batch 100): ms per call of every estimator
method, host wall clock around a synchronised loop (these calls are launch- / host-bound, not throughput-bound)."""
import json, os, sys, time
import torch, torch.nn as nn, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import curvature_b200 as cb  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
model = nn.Sequential(nn.Conv2d(1, 6, 5, padding=2), nn.ReLU(), nn.MaxPool2d(2), nn.Conv2d(6, 16, 5), nn.ReLU(), nn.MaxPool2d(2),
                      nn.Flatten(), nn.Linear(400, 120), nn.ReLU(), nn.Linear(120, 84), nn.ReLU(), nn.Linear(84, 10)).to(dev)
x = torch.randn(100, 1, 28, 28, device=dev)
kfac, diag = cb.KFAC(model), cb.Diagonal(model)
out = model(x)
F.cross_entropy(out, torch.distributions.Categorical(logits=out).sample()).backward()


def timed(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


res = {"kfac_update_ms": timed(lambda: kfac.update(100)), "diag_update_ms": timed(lambda: diag.update(100))}
eig = cb.get_eigenvectors(kfac.state)
efb = cb.EFB(model, kfac.state, eigvecs=eig)
res["efb_update_ms"] = timed(lambda: efb.update(100))
res["kfac_invert_ms"] = timed(lambda: kfac.invert(0.5, 1.0))
res["diag_invert_ms"] = timed(lambda: diag.invert(0.5, 1.0))
res["efb_invert_ms"] = timed(lambda: efb.invert(0.5, 1.0))
res["kfac_sample_and_replace_ms"] = timed(lambda: kfac.sample_and_replace())
res["efb_sample_and_replace_ms"] = timed(lambda: efb.sample_and_replace())
res["diag_sample_and_replace_ms"] = timed(lambda: diag.sample_and_replace())
inf = cb.INF(model, diag.state, kfac.state, efb.state, eigvecs=eig)
res["inf_update_rank100_ms"] = timed(lambda: inf.update(rank=100), n=5, warm=1)
res["inf_invert_ms"] = timed(lambda: inf.invert(1e15, 1e20), n=5, warm=1)
res["inf_sample_and_replace_ms"] = timed(lambda: inf.sample_and_replace(), n=20, warm=2)
print(json.dumps(res))
