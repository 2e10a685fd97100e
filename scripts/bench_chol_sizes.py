"""K4 by matrix order: the batched Cholesky-of-inverse kernel on ONE matrix of each order (where the critical path of a
layer-sharded invert lies), on a batch of equal matrices, and torch.linalg (cuSOLVER potrf + trtri-like inverse) on the same
matrix as context (library code: not the product path).

    python scripts/bench_chol_sizes.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curvature_b200 import _native as nat  # noqa: E402

dev = "cuda:0"


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
for D in (64, 256, 512, 1024, 2048, 2304, 4608):
    torch.manual_seed(D)
    X = torch.randn(D, D + 64, device=dev)
    F = (X @ X.t() / X.shape[1]).contiguous()
    L = torch.empty_like(F)
    one = timed(lambda: nat.chol_inv_batched([F], [1.0], [10.0], [L]))
    nb = 8
    Fs = [F.clone() for _ in range(nb)]
    Ls = [torch.empty_like(F) for _ in range(nb)]
    many = timed(lambda: nat.chol_inv_batched(Fs, [1.0] * nb, [10.0] * nb, Ls))

    def lib():
        reg = 10.0 ** 0.5 * F + torch.eye(D, device=dev)
        return torch.linalg.cholesky(torch.linalg.inv((reg + reg.t()) / 2))
    lib_ms = timed(lib)
    out[D] = {"one_matrix_ms": one, "eight_matrices_ms": many, "torch_linalg_inv_cholesky_ms": lib_ms,
              "gflops_one": 2.0 * D ** 3 / one / 1e6}
    print(D, out[D], flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
