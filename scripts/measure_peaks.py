"""Measure this box's roofline denominators for context (the driver's MEASURED_PEAKS.json stays authoritative):
STREAM-style copy GB/s, cuBLAS bf16 and TF32 GEMM TFLOP/s (burst and a ~2 s sustained loop).

    python scripts/measure_peaks.py [out.json]

cuBLAS is used here ONLY as the yardstick the hand-written kernels are compared with; nothing in the product path
calls it."""
import json
import sys
import time

import torch


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = "cuda:0"
    out = {"device": torch.cuda.get_device_name(0)}
    a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    b = torch.empty_like(a)
    for _ in range(3):
        b.copy_(a)
    ms = timed(lambda: b.copy_(a), 20)
    out["hbm_copy_gbs"] = 2 * a.numel() / ms / 1e6
    del a, b
    n = 8192
    for name, dt, tf32 in (("bf16", torch.bfloat16, False), ("tf32", torch.float32, True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        x = torch.randn(n, n, device=dev, dtype=dt)
        y = torch.randn(n, n, device=dev, dtype=dt)
        for _ in range(3):
            x @ y
        ms = timed(lambda: x @ y, 10)
        out[f"{name}_tflops_burst"] = 2 * n ** 3 / ms / 1e9
        t0 = time.time()
        reps = 0
        torch.cuda.synchronize()
        while time.time() - t0 < 1.0:
            x @ y
            reps += 1
            if reps % 20 == 0:
                torch.cuda.synchronize()
        ms = timed(lambda: x @ y, 200)
        out[f"{name}_tflops_sustained"] = 2 * n ** 3 / ms / 1e9
    torch.backends.cuda.matmul.allow_tf32 = False
    print(json.dumps(out))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
