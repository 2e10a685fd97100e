#!/usr/bin/env python
"""Headline benchmark: ResNet-50 KFAC.update images/sec on synthetic ImageNet-shaped batches (BASELINE.json).

    python bench.py --gpus 1 --steps 10 --warmup 3                 # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 # the reference algorithm on the host CPU
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU, weak scaling

A "step" is one pass of the hot path -- `KFAC.update(batch)` over one recorded batch: for each of the 54
Conv2d/Linear layers the fused implicit-im2col SYRK for A and the SYRK for G, accumulated into the factor arena.
`value` times exactly K such steps with the recorded activations / output gradients already resident in HBM
(10.2 + 10.6 GiB per 256-batch: far larger than the 126 MB L2, so no flush is needed between iterations);
`e2e` times the user-visible loop of scripts/factors.py:46-61 (pinned host batch -> H2D -> forward -> sampled
labels -> backward -> update -> loss read back).  With N > 1 every rank processes its own batches and one
all-reduce of the arena per pass (inside the timed region) merges them; time = max over ranks.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import warnings

import torch

warnings.filterwarnings("ignore", category=FutureWarning)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ResNet-50 KFAC update images/sec"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="resnet50", choices=["resnet18", "resnet50", "resnet152", "lenet5"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default 256; 100 for lenet5)")
    ap.add_argument("--precision", default=os.environ.get("CURVATURE_B200_PRECISION", "bf16"),
                    help="arithmetic tier of the factor SYRKs: bf16 (default; stated 1e-3 tier: bf16 copy for re-read "
                         "operands, TF32 for read-once ones), tf32_tma (TF32 truncation, 1e-3), tf32 (round-to-nearest TF32 "
                         "pre-pass), fp32 (CUDA cores, 1e-5)")
    ap.add_argument("--layout", default="channels_last", choices=["channels_last", "nchw"],
                    help="memory format the model runs in (the logical tensors are identical; channels_last lets "
                         "every factor operand reach the tensor core through TMA)")
    ap.add_argument("--cpu-batch", type=int, default=16, help="batch of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--per-layer", default=None, help="write a per-layer SYRK timing table (JSON) to this path")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch images per GPU (default); strong: --batch images in total, split over the GPUs "
                         "(SURVEY 8(d) config 3's secondary variant)")
    ap.add_argument("--allreduce-every", type=int, default=0,
                    help="all-reduce the factor arena every this many updates inside the timed region (0 = once per pass, "
                         "the design point; 1 = the per-update stress variant of SURVEY 8(d) config 3)")
    ap.add_argument("--mode", default="update", choices=["update", "invert"],
                    help="update: the headline metric (KFAC.update images/s).  invert: BASELINE configs[4] -- KFAC.invert of "
                         "every factor of the model with the README damping, layer-sharded over the GPUs + one all-gather; "
                         "value = ms per invert (lower is better)")
    ap.add_argument("--no-context", action="store_true",
                    help="skip the context lines (eigensolver timing, the reference algorithm on torch-CUDA)")
    return ap.parse_args()


def make_model(name):
    torch.manual_seed(0)
    if name == "lenet5":
        nn = torch.nn                                 # curvature/lenet5.py:11-24 (architecture)
        return nn.Sequential(nn.Conv2d(1, 6, 5, padding=2), nn.ReLU(), nn.MaxPool2d(2, 2), nn.Conv2d(6, 16, 5),
                             nn.ReLU(), nn.MaxPool2d(2, 2), nn.Flatten(), nn.Linear(400, 120), nn.ReLU(),
                             nn.Linear(120, 84), nn.ReLU(), nn.Linear(84, 10)), (1, 28, 28)
    import torchvision
    return getattr(torchvision.models, name)(weights=None), (3, 224, 224)


def fisher_step(model, x):
    """Loop body of scripts/factors.py:51-59: forward, sample labels from the model's own predictive
    distribution, mean cross-entropy, zero_grad, backward."""
    logits = model(x)
    labels = torch.distributions.Categorical(logits=logits.detach()).sample()
    loss = torch.nn.functional.cross_entropy(logits, labels)
    model.zero_grad()
    loss.backward()
    return loss


def algorithmic_flops(kfac):
    """sum over layers of R*K(K+1) + R*M(M+1) (one multiply-add per unique element of each
    symmetric factor per contraction row)."""
    total = 0
    rows = []
    for layer, (x, g) in kfac.record.items():
        bias = layer.bias is not None
        if layer.__class__.__name__ == "Conv2d":
            R = g.shape[0] * g.shape[2] * g.shape[3]
            K = layer.weight[0].numel() + bias
        else:
            R = g.shape[0]
            K = layer.weight.shape[1] + bias
        M = layer.weight.shape[0]
        f = R * (K * (K + 1) + M * (M + 1))
        rows.append((layer, K, M, R, f))
        total += f
    return total, rows


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.nvml = None
        try:        # NVML in a sampling thread (every 10 ms): a K-step region is ~0.1 s, nvidia-smi -lms gives 1-2 samples
            import threading
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = {"sm": [], "mx": pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM), "mask": 0, "stop": False}
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self.nvml["stop"]:
                    try:
                        self.nvml["sm"].append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        self.nvml["mask"] |= int(get_reasons(h))
                    except Exception:
                        pass
                    time.sleep(0.01)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.nvml is not None:
            self.nvml["stop"] = True
            self.thread.join(timeout=2)
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            sm = [x for x in self.nvml["sm"] if x > 0]
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.nvml["mx"]),
                    "samples": len(self.nvml["sm"]), "reasons": sorted(n for n, b in bits.items() if self.nvml["mask"] & b),
                    "source": "nvml, 10 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(args, steps, warmup, model_name, batch):
    """The reference algorithm (oracle port of curvatures.py:312-350; the reference is pure Python + ATen and
    has no separately compilable kernel) on the host cores: `KFAC.update` on a bounded batch."""
    import oracle.curvature_oracle as orc
    model, shape = make_model(model_name)
    model.train()
    kfac = orc.KFAC(model)
    torch.manual_seed(1000)
    x = torch.randn(batch, *shape)
    fisher_step(model, x)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        kfac.update(batch)
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    total = sum(times)
    return {"value": batch * len(times) / total, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{model_name} KFAC.update on one {batch}-image {shape[-1]}^2 batch x {len(times)} timed calls "
                      f"({warmup} warm-up), torch {torch.__version__} CPU, {os.cpu_count()} logical cpus",
            "ms_per_step": 1e3 * total / len(times)}


def reference_cuda_run(model, kfac, batch, steps=2, warmup=1):
    """CONTEXT ONLY, not the target: the reference's own op sequence for `KFAC.update` (curvatures.py:321-350:
    F.unfold -> permute/contiguous -> torch.mm -> div -> add_) executed by torch on the SAME B200 on the tensors the hooks
    recorded, fp32 (`allow_tf32 = False`, torch's default for matmul).  Says how much of the speed-up over the CPU arm is
    "a GPU" and how much is this repo's kernels."""
    import torch.nn.functional as F
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    state = {}

    def update():
        for layer, (x, g) in kfac.record.items():
            x, g = x.detach(), g.detach() * g.size(0)
            if layer.__class__.__name__ == "Conv2d":
                x = F.unfold(x, layer.kernel_size, padding=layer.padding, stride=layer.stride)
                x = x.permute(1, 0, 2).contiguous().view(x.shape[1], -1)
                g = g.permute(1, 0, 2, 3).contiguous().view(g.shape[1], -1)
            else:
                x, g = x.t(), g.t()
            if layer.bias is not None:
                x = torch.cat([x, torch.ones_like(x[:1])], dim=0)
            first = torch.mm(x, x.t()) / float(x.shape[1])
            second = torch.mm(g, g.t()) / float(g.shape[1])
            if layer in state:
                state[layer][0] += first
                state[layer][1] += second
            else:
                state[layer] = [first, second]
    try:
        for _ in range(warmup):
            update()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            update()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": batch / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "kind": "port of curvatures.py:321-350 on torch-CUDA "
                "(F.unfold + torch.mm, fp32, allow_tf32=False)", "note": "context only, not the target"}
    except RuntimeError as exc:                              # e.g. out of memory on the unfolded matrices
        return {"unavailable": str(exc).splitlines()[0][:200]}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
        state.clear()
        torch.cuda.empty_cache()


def eigh_timing(kfac):
    """The one-shot eigenbases of EFB / INF (utils.py:45-60 -> cuSOLVER syevd through torch.linalg.eigh), timed on their
    own: all factors of the model, and the largest factor alone."""
    import curvature_b200 as cb
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    big = max((f for v in kfac.state.values() for f in v), key=lambda f: f.shape[0])
    torch.linalg.eigh(big + big.t(), UPLO="U")               # warm-up (cuSOLVER handle, workspace)
    torch.cuda.synchronize()
    e0.record()
    torch.linalg.eigh(big + big.t(), UPLO="U")
    e1.record()
    torch.cuda.synchronize()
    one = e0.elapsed_time(e1)
    e0.record()
    cb.get_eigenvectors(kfac.state)
    e1.record()
    torch.cuda.synchronize()
    return {"get_eigenvectors_ms": e0.elapsed_time(e1), "matrices": 2 * len(kfac.state),
            "largest_order": int(big.shape[0]), "largest_alone_ms": one,
            "what": "cuSOLVER syevd via torch.linalg.eigh (the library call the north star allows for the one-shot eigenbases)"}


README_DAMPING = {"resnet18": (1.0, 18916.0), "resnet50": (69.0, 25771.0), "resnet152": (2765.0, 10162.0),
                  "lenet5": (0.5, 1.0)}            # README.rst:262-264 "KFAC Norm" / "KFAC Scale"; README.rst:148 for LeNet-5


def invert_mode(args, rank, world, local):
    """BASELINE configs[4]: `KFAC.invert(add, multiply)` of all factors (ResNet-152: 312 matrices, orders 64 .. 4608) --
    sharded over the ranks by D^3 (LPT), every rank runs the batched Cholesky-of-inverse kernel on its own matrices, ONE
    all-gather of the rank-major inverse arena (curvature_b200/parallel.py).  Time = device time of K inverts, max over
    ranks.  The factors come from one update on a small synthetic batch (their values do not change the work)."""
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import curvature_b200 as cb
    model, shape = make_model(args.model)
    model = model.to(dev).train()
    if args.model != "lenet5":
        model = model.to(memory_format=torch.channels_last)
    kfac = cb.KFAC(model, precision=args.precision)
    torch.manual_seed(1000)                                  # same batch on every rank: the state is the merged one
    nb = args.batch or 32
    x = torch.randn(nb, *shape, device=dev)
    fisher_step(model, x)
    kfac.update(nb)
    kfac.record = {k: [None, None] for k in kfac.record}     # release the recorded activations
    del x
    torch.cuda.empty_cache()
    add, mul = README_DAMPING[args.model]
    dims = [f.shape[0] for v in kfac.state.values() for f in v]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        kfac.invert(add, mul)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    e0.record()
    for _ in range(args.steps):
        kfac.invert(add, mul)
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / args.steps
    # the same without sharding (every rank inverts everything): what the exchange step buys
    unsharded = None
    if world > 1:
        kfac.invert(add, mul, shard=False)
        barrier()
        e0.record()
        for _ in range(max(1, args.steps // 2)):
            kfac.invert(add, mul, shard=False)
        e1.record()
        barrier()
        unsharded = e0.elapsed_time(e1) / max(1, args.steps // 2)
    if rank == 0:
        plan = cb.invert_plan_two_rounds(dims, world) if world > 1 else cb.invert_plan(dims, world)
        line = {"metric": f"{args.model} KFAC invert ms (all factors, README damping)", "value": ms, "unit": "ms",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{args.model} KFAC.invert{(add, mul)}: {len(dims)} matrices, orders {min(dims)}..{max(dims)} "
                                       "(BASELINE configs[4])", "network": args.model,
                           "parallelism": f"layer-sharded x{world}" + (", exchange in two rounds, the first beside every rank's largest matrix" if world > 1 else ""),
                           "arena_bytes": 4 * plan["total"]},
                "unsharded_ms_per_invert_on_every_rank": unsharded, "clocks": clocks,
                "gpu_launches": None}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # the model's own forward / backward (torch + cuDNN, not this repo's code) is part of `e2e` only; let cuDNN pick
    # its fastest algorithms for the fixed shapes
    torch.backends.cudnn.benchmark = os.environ.get("CRV_BENCH_CUDNN_BENCHMARK", "1") != "0"
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    batch = args.batch or (100 if args.model == "lenet5" else 256)
    if args.scaling == "strong":
        assert batch % world == 0, "strong scaling: the total batch must divide over the GPUs"
        batch //= world
    workload = {"lenet5": "LeNet-5 KFAC.update, synthetic 28x28 batches of 100 (BASELINE configs[0])",
                "resnet18": "ResNet-18 KFAC.update, synthetic 224x224 batch 256 (BASELINE configs[2])",
                "resnet50": "ResNet-50 KFAC.update, synthetic 224x224 batch 256 per GPU (BASELINE metric config)",
                "resnet152": "ResNet-152 KFAC.update, synthetic 224x224 batch 256 (BASELINE configs[4])"}[args.model]
    metric = METRIC if args.model == "resnet50" else METRIC.replace("ResNet-50", args.model)
    config = {"workload": workload, "network": args.model, "batch_per_gpu": batch, "global_batch": batch * world,
              "image": "3x224x224" if args.model != "lenet5" else "1x28x28", "parallelism": f"dp{world}", "allreduce": "once per pass" if args.allreduce_every <= 0 else
              f"every {args.allreduce_every} update(s)",
              "memory_format": args.layout if args.model != "lenet5" else "nchw",
              "l2": "inputs larger than L2 (recorded activations + gradients >> 126 MB); no flush needed"
                    if args.model != "lenet5" else "inputs smaller than L2; 256 MB buffer written between iterations"}

    if args.impl == "reference":
        if rank != 0:
            return
        # torchrun exports OMP_NUM_THREADS=1 to every rank: the reference arm uses all host cores at every N
        torch.set_num_threads(os.cpu_count() or 1)
        if args.mode == "invert":
            # the reference's own invert (oracle port of curvatures.py:354-385: inverse + cholesky per matrix, serial loop)
            import oracle.curvature_oracle as orc
            model, shape = make_model(args.model)
            model.train()
            kfac = orc.KFAC(model)
            torch.manual_seed(1000)
            nb = min(args.batch or 8, 8)
            fisher_step(model, torch.randn(nb, *shape))
            kfac.update(nb)
            add, mul = README_DAMPING[args.model]
            t0 = time.perf_counter()
            kfac.invert(add, mul)
            ms = 1e3 * (time.perf_counter() - t0)
            dims = [f.shape[0] for v in kfac.state.values() for f in v]
            print(json.dumps({"impl": "reference", "metric": f"{args.model} KFAC invert ms (all factors, README damping)",
                              "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": ms,
                              "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                              "data": "synthetic", "config": {"workload": f"{args.model} KFAC.invert{(add, mul)}: {len(dims)} "
                                                              "matrices (BASELINE configs[4])", "network": args.model},
                              "cpu_baseline": {"value": ms, "unit": "ms", "cores": torch.get_num_threads(), "kind": "port",
                                               "sample": "one full invert (inverse + cholesky per matrix, fp32, torch CPU)"},
                              "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                              "gpu_launches": 0}))
            return
        cb_ = batch if args.model == "lenet5" else min(args.cpu_batch, batch)
        res = cpu_reference_run(args, args.steps, args.warmup, args.model, cb_)
        line = {"impl": "reference", "metric": metric, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": dict(config, cpu_sample_batch=cb_),
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    if args.mode == "invert":
        return invert_mode(args, rank, world, local)

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import curvature_b200 as cb
    from curvature_b200 import _native as nat

    model, shape = make_model(args.model)
    model = model.to(dev).train()
    if args.layout == "channels_last" and args.model != "lenet5":
        model = model.to(memory_format=torch.channels_last)
    kfac = cb.KFAC(model, precision=args.precision)
    torch.manual_seed(1000 + rank)
    x_host = torch.randn(batch, *shape).pin_memory()
    x_dev = torch.empty(batch, *shape, device=dev)
    x_dev.copy_(x_host, non_blocking=True)
    fisher_step(model, x_dev)            # records activations and output gradients (resident from here on)
    flops, rows = algorithmic_flops(kfac)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if args.model == "lenet5" else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- update-only (inputs resident in HBM) ----------------
    for _ in range(args.warmup):
        kfac.update(batch)
    if world > 1:                         # warm NCCL up on a scratch buffer of the arena's size (channels, NVLS setup)
        scratch = torch.zeros_like(kfac.arena.flat)
        dist.all_reduce(scratch)
        del scratch
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = nat.launch_calls
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 2)]
    ar_ev = []
    fc_layer = list(kfac.record.keys())[-1]
    updates_before = float(kfac.state[fc_layer][0][-1, -1].item()) if fc_layer.bias is not None else None
    barrier()
    t_wall0 = time.perf_counter()
    ev[0].record()
    n_allreduce = 0
    for i in range(args.steps):
        if flush is not None:
            flush.fill_(i)
        ev[1 + 2 * i].record()
        kfac.update(batch)
        ev[2 + 2 * i].record()
        if world > 1 and args.allreduce_every > 0 and (i + 1) % args.allreduce_every == 0 and i + 1 < args.steps:
            ar_ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
            ar_ev[-1][0].record()
            cb.allreduce_arena(kfac)
            ar_ev[-1][1].record()
            n_allreduce += 1
    if world > 1:
        ar_ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
        ar_ev[-1][0].record()
        cb.allreduce_arena(kfac)          # the one collective of the estimation pass
        ar_ev[-1][1].record()
        n_allreduce += 1
    ev[-1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    collective = None
    if world > 1:
        ar_ms = [a.elapsed_time(b) for a, b in ar_ev]
        nbytes = kfac.arena.flat.numel() * 4
        collective = {"allreduce_calls": n_allreduce, "allreduce_ms": statistics.mean(ar_ms), "bytes": nbytes,
                      "algbw_gbs": nbytes / (statistics.mean(ar_ms) / 1e3) / 1e9,
                      "busbw_gbs": nbytes / (statistics.mean(ar_ms) / 1e3) / 1e9 * 2 * (world - 1) / world,
                      "note": "device time of ncclAllReduce on the factor arena (event pair around the call; includes "
                              "waiting for the slowest rank to arrive)"}
        if updates_before is not None and args.allreduce_every <= 0:
            # the ones row of the fc layer's A factor counts updates: after the merge it must hold the global count
            got = float(kfac.state[fc_layer][0][-1, -1].item())
            want = (updates_before + args.steps) * world
            collective["collective_check"] = {"fc_A_bias_corner": got, "expected_updates_x_world": want, "ok": got == want}
            assert got == want, f"all-reduce check failed: fc A[-1,-1] = {got}, expected {want}"
    launches = nat.launch_calls - launches0
    clocks = sampler.stop() if sampler else None
    # the same K steps once more with one CUDA event pair around EVERY kernel launch (recorded by the library on the
    # launching stream): per-kernel durations for the roofline.  Kept out of the region above because the event
    # records between back-to-back launches cost ~8 % of the step.
    nat.profile_enable(True)
    for i in range(args.steps):
        if flush is not None:
            flush.fill_(i)
        kfac.update(batch)
    torch.cuda.synchronize(dev)
    kernels = nat.profile_collect()
    nat.profile_enable(False)
    total_ms = ev[0].elapsed_time(ev[-1])
    update_ms = [ev[1 + 2 * i].elapsed_time(ev[2 + 2 * i]) for i in range(args.steps)]
    if flush is not None:
        total_ms = sum(update_ms)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = t.item()
    value = world * batch * args.steps / (total_ms / 1e3)

    # ---------------- per-layer table (optional, outside the timed region) ----------------
    if args.per_layer and rank == 0:
        table = []
        for layer, K, M, R, f in rows:
            x, g = kfac.record[layer]
            xs, gs = x.detach(), g.detach()
            first, second = kfac.state[layer]
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            bias = layer.bias is not None
            ta, tg = [], []
            for rep in range(4):
                e[0].record()
                if layer.__class__.__name__ == "Conv2d":
                    nat.syrk_conv_accum(xs, layer.kernel_size, layer.stride, layer.padding, bias, 1.0 / R, first,
                                        kfac.precision)
                else:
                    nat.syrk_rows_accum(xs, bias, 1.0 / R, first, kfac.precision)
                e[1].record()
                nat.syrk_rows_accum(gs, False, 1.0 / R, second, kfac.precision)
                e[2].record()
                torch.cuda.synchronize(dev)
                if rep:
                    ta.append(e[0].elapsed_time(e[1]))
                    tg.append(e[1].elapsed_time(e[2]))
            a_ms, g_ms = statistics.median(ta), statistics.median(tg)
            table.append({"layer": str(layer), "K": K, "M": M, "R": R, "A_ms": a_ms, "G_ms": g_ms,
                          "A_tflops": R * K * (K + 1) / a_ms / 1e9, "G_tflops": R * M * (M + 1) / g_ms / 1e9,
                          "A_gbs": 4 * x.numel() / a_ms / 1e6, "G_gbs": 4 * g.numel() / g_ms / 1e6,
                          "x_channels_last": bool(x.dim() == 4 and nat._is_channels_last(x)),
                          "g_channels_last": bool(g.dim() == 4 and nat._is_channels_last(g))})
        os.makedirs(os.path.dirname(os.path.abspath(args.per_layer)), exist_ok=True)
        json.dump(table, open(args.per_layer, "w"), indent=1)

    # ---------------- end to end through the public API with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        # the user's loop with a standard input pipeline: the NEXT batch's host->device copy (pinned memory, its own
        # stream, double-buffered) overlaps the current step; every step still copies its own inputs inside the region
        copy_stream = torch.cuda.Stream(device=dev)
        bufs = [x_dev, torch.empty_like(x_dev)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i % 2])      # the step that last read this buffer is done with it
                bufs[i % 2].copy_(x_host, non_blocking=True)
                ready[i % 2].record(copy_stream)

        def e2e_step(i, more):
            torch.cuda.current_stream(dev).wait_event(ready[i % 2])
            if more:
                prefetch(i + 1)
            loss = fisher_step(model, bufs[i % 2])
            kfac.update(batch)
            consumed[i % 2].record(torch.cuda.current_stream(dev))
            return loss.item()        # device -> host read of the step's result

        for ev_ in consumed:
            ev_.record(torch.cuda.current_stream(dev))
        nwarm = max(1, min(args.warmup, 2))
        prefetch(0)
        for i in range(nwarm):
            e2e_step(i, True)
        barrier()
        t0 = time.perf_counter()
        for i in range(nwarm, nwarm + args.steps):
            e2e_step(i, True)       # (K copies inside the region: the one consumed first was issued before it, the last one is not consumed)
        if world > 1:
            cb.allreduce_arena(kfac)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * batch * args.steps / dt.item(), "unit": UNIT,
               "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 4,
               "what": "pinned host batch -> H2D (prefetched one step ahead on a copy stream) -> forward -> sampled labels -> "
                       "backward -> KFAC.update -> loss.item(); host wall clock over K steps"}

    # ---------------- context lines (rank 0 at N = 1 only; outside every timed region) ----------------
    context = None
    if world == 1 and not args.no_context and args.model != "lenet5":
        context = {"eigh": eigh_timing(kfac), "reference_cuda": reference_cuda_run(model, kfac, batch)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # Which measured peak is the denominator: the SUSTAINED cuBLAS figure was taken over 4 s at ~1.34 GHz, the BURST one is a
    # best-of-10 of single calls at boost clocks.  A timed region shorter than ~1 s runs at boost clocks (see `clocks`), so
    # the burst figure is the honest denominator there; both fractions are printed.  Fallbacks (B200_PROFILING.md): 1.4 /
    # 1.65 PFLOP/s.
    bf16_sus = peaks.get("bf16_tflops_sustained", 1400.0)
    bf16_burst = peaks.get("bf16_tflops", 1650.0)
    use_burst = t_wall < 1.0
    bf16 = bf16_burst if use_burst else bf16_sus
    which = "bf16_tflops (burst: timed region %.2f s < 1 s)" % t_wall if use_burst else "bf16_tflops_sustained (timed region %.2f s)" % t_wall
    peak_src = ("of measured (MEASURED_PEAKS.json " + which) if peaks else ("of fallback (B200_PROFILING.md " + which)
    tier = args.precision
    step_ms = statistics.mean(update_ms)
    # dominant kernel = the contraction kernel class with the largest share of the step's device time.  The split
    # reduction and the cast pre-pass run on side streams: their event brackets include the time they wait for an SM
    # next to the resident SYRK CTAs, so they are reported in step_breakdown but not ranked here (the serialised ncu
    # launch list under profiles/ gives their true share).
    busy = {k: v for k, v in kernels.items() if v["launches"]}
    ranked = {k: v for k, v in busy.items() if v["flops"]} or busy
    dom = max(ranked, key=lambda k: ranked[k]["ms"]) if ranked else None
    kern_total_ms = sum(v["ms"] for v in busy.values())
    roofline = None
    if dom is not None:
        d = busy[dom]
        dom_peak = bf16 if dom == "syrk_nhwc_bf16" else bf16 / 2.0
        note = peak_src + (")" if dom == "syrk_nhwc_bf16" else " / 2: kind::tf32 issues at half the bf16 rate)")
        achieved = d["flops"] / (d["ms"] / 1e3) / 1e12
        traffic = None
        try:
            tr_path = os.path.join(ROOT, "profiles", "r2_traffic.json")        # ncu dram__bytes per launch, final code
            if not os.path.exists(tr_path):
                tr_path = os.path.join(ROOT, "profiles", "r1_traffic.json")
            tr = json.load(open(tr_path))
            traffic = tr.get(dom, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        scale = 1.0 if dom == "syrk_nhwc_bf16" else 0.5
        roofline = {"bound": "tensor", "achieved": achieved, "peak": dom_peak, "unit": "TFLOP/s", "frac": achieved / dom_peak,
                    "frac_of_burst_peak": achieved / (bf16_burst * scale), "frac_of_sustained_peak": achieved / (bf16_sus * scale),
                    "traffic": traffic, "kernel": dom, "launches": d["launches"],
                    "avg_launch_ms": d["ms"] / d["launches"],
                    "algorithmic_flops_per_launch": d["flops"] / d["launches"],
                    "share_of_step_kernel_time": d["ms"] / kern_total_ms if kern_total_ms else None,
                    "peak_source": note, "tier": tier,
                    "how": "CUDA event pair recorded on the launching stream around every launch of this kernel inside "
                           "the timed region; algorithmic flops = R*D*(D+1) per factor (SURVEY 8d)"}
    whole_step = {"algorithmic_flops_per_step": flops, "achieved_tflops": flops / (step_ms / 1e3) / 1e12,
                  "frac_of_burst_bf16_peak": flops / (step_ms / 1e3) / 1e12 / bf16_burst,
                  "frac_of_sustained_bf16_peak": flops / (step_ms / 1e3) / 1e12 / bf16_sus,
                  "kernel_classes": {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                                         "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["ms"] and v["flops"] else None,
                                         "gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9) if v["ms"] else None}
                                     for k, v in busy.items()}}
    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "bf16x3": "bf16x3", "bf16": "bf16", "tf32_tma": "tf32"}[tier],
            "tolerance": "factors within 1e-3 relative Frobenius of the fp32 reference (north-star tensor-core tier)" if tier != "fp32" else "1e-5",
            "data": "synthetic", "config": config, "roofline": roofline, "step_breakdown": whole_step, "e2e": e2e,
            "gpu_launches": sum(v["launches"] for v in busy.values()) or launches,
            "clocks": clocks, "wall_s_timed_region": t_wall, "update_ms_mean": step_ms}
    if collective is not None:
        line["collective"] = collective
    if context is not None:
        line["context"] = context
    if world == 1 and not args.no_cpu_baseline:
        cb_ = batch if args.model == "lenet5" else min(args.cpu_batch, batch)
        res = cpu_reference_run(args, 3, 1, args.model, cb_)
        line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
