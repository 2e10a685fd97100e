"""Fisher information approximations with the reference's Python surface, computed by sm_100a kernels.

Drop-in for ``curvature/curvatures.py`` of DLR-RM/curvature: the class names, constructor
arguments, ``update`` / ``invert`` / ``sample`` / ``sample_and_replace``, the ``state`` / ``inv_state``
layouts (dicts keyed by the module objects, in ``model.modules()`` order) and the assertion messages
are the reference's (curvatures.py:17-672).  The bodies contain no ``torch.mm`` / ``unfold`` / ``cat``:
hooks stash device pointers and ``update`` hands them to the C-ABI kernels in ``libcurvature_b200.so``
(include/curvature_b200.h).  There is no CPU path: tensors must be CUDA fp32.

Differences a caller can observe (all documented in DESIGN.md):
  * ``KFAC.record[layer][1]`` holds the *unscaled* output gradient; the reference's ``* batch`` of
    curvatures.py:310 is folded into the SYRK epilogue scale.  ``KFAC.scaled_record(layer)`` returns the
    reference's tensor.
  * ``sample`` / ``sample_and_replace`` accept an optional ``noise`` argument (the Gaussian tensor the
    reference would draw) so that posterior samples can be compared with identical noise.
  * every estimator keeps its state in one flat fp32 arena (``.arena``); the per-layer tensors in ``state``
    are views into it, which is what a single NCCL all-reduce per estimation pass operates on.
  * ``add`` / ``multiply`` may be ints as well as floats for every estimator.
"""
from abc import ABC, abstractmethod
import copy
import warnings
from typing import Any, Dict, List, Optional, Sequence, Union

import os

import torch
from torch import Tensor
from torch.nn import Module, Sequential

from . import _native as nat
from .utils import get_eigenvectors, kron

# the reference registers the same (non-full) backward hook kind; torch warns about it on every forward
warnings.filterwarnings("ignore", message="Using a non-full backward hook")

_SUPPORTED = ['Linear', 'Conv2d', 'MultiheadAttention']
_SCALARS = (float, int)


def _pair(v):
    if isinstance(v, (tuple, list)):
        if len(v) == 1:
            return int(v[0]), int(v[0])
        return int(v[0]), int(v[1])
    return int(v), int(v)


class FactorArena:
    """One flat, zero-initialised fp32 buffer carved into per-layer matrices (256-byte aligned views)."""

    ALIGN = 64  # floats

    def __init__(self, shapes: Sequence[Sequence[int]], device, dtype=torch.float32):
        offsets, total = [], 0
        for shape in shapes:
            n = 1
            for s in shape:
                n *= int(s)
            offsets.append(total)
            total += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.flat = torch.zeros(max(total, self.ALIGN), dtype=dtype, device=device)
        self.views = []
        for off, shape in zip(offsets, shapes):
            n = 1
            for s in shape:
                n *= int(s)
            self.views.append(self.flat[off:off + n].view(*shape))


def _damping(add, multiply, index, count):
    """Scalar or per-layer damping (curvatures.py:183-187, 361-365)."""
    if not isinstance(add, _SCALARS) and not isinstance(multiply, _SCALARS):
        assert len(add) == len(multiply) == count
        return float(add[index]), float(multiply[index])
    return float(add), float(multiply)


def _grad_of(p: Tensor, what: str) -> Tensor:
    if p.grad is None:
        raise RuntimeError(f"{what}.grad is None. Did you call 'backward' prior to 'update'?")
    g = p.grad
    return g if g.is_contiguous() else g.contiguous()


def _gemm_tier(precision: int) -> int:
    """Tier of the chained GEMMs (K3 / K5): the 1e-5 tiers (fp32, bf16x3) use the fp32 CUDA-core GEMM, the 1e-3 tiers the
    TF32 tensor-core GEMM."""
    return nat.PREC_FP32 if precision in (nat.PREC_FP32, nat.PREC_BF16X3) else nat.PREC_TF32


def _rounded(cache: dict, key, tensors):
    """TF32-rounded copies of operands that stay constant over many GEMM calls (made once, cached)."""
    hit = cache.get(key)
    if hit is None or any(h[0] is not t for h, t in zip(hit, tensors)):
        hit = [(t, nat.round_tf32(t if t.is_contiguous() else t.contiguous())) for t in tensors]
        cache[key] = hit
    return [h[1] for h in hit]


def _flat_noise(shapes: dict, like: Tensor, tier: int) -> dict:
    """{key: (K, M) standard-normal matrix}, all views of ONE flat draw (rounded to TF32 once on tensor-core tiers);
    every view starts on a 16-byte boundary."""
    offs, total = {}, 0
    for key, (k, m) in shapes.items():
        offs[key] = total
        total += (k * m + 3) // 4 * 4
    flat = torch.randn(total, device=like.device, dtype=like.dtype)
    if tier != nat.PREC_FP32:
        nat.round_tf32(flat, out=flat)
    return {key: flat[offs[key]:offs[key] + k * m].view(k, m) for key, (k, m) in shapes.items()}


class _DenseTarget:
    """Stands in for a parameter whose storage is not row-major: `.data` is a dense buffer of the same shape."""

    def __init__(self, data: Tensor):
        self.data = data
        self.shape = data.shape


class Curvature(ABC):
    """Base class of all approximations (reference: curvatures.py:17-129)."""

    def __init__(self,
                 model: Union[Module, Sequential],
                 layer_types: Union[List[str], str] = None,
                 precision: Union[str, int, None] = None):
        """
        Args:
            model: Any (pre-trained) PyTorch model, on a CUDA device.
            layer_types: `Linear`, `Conv2d`, `MultiheadAttention`; all three if `None` or `[]`.
            precision: arithmetic tier of the dense contractions: `'bf16x3'` (default: two-term bf16 split on the tensor
                       cores, 1e-5 parity tier), `'fp32'` (CUDA cores, 1e-5), `'bf16'` / `'tf32'` / `'tf32_tma'` (tensor
                       cores, stated 1e-3 tier; `'bf16'` is the benchmark tier).  Default from
                       `CURVATURE_B200_PRECISION`, else `'bf16x3'`.
        """
        self.model = model
        self.model_state = copy.deepcopy(model.state_dict())
        self.layer_types = list()
        if isinstance(layer_types, str):
            self.layer_types.append(layer_types)
        elif isinstance(layer_types, list):
            if layer_types:
                self.layer_types.extend(layer_types)
            else:
                self.layer_types.extend(_SUPPORTED)
        elif layer_types is None:
            self.layer_types.extend(_SUPPORTED)
        else:
            raise TypeError
        for _type in self.layer_types:
            assert _type in _SUPPORTED
        self.state = dict()
        self.inv_state = dict()
        self.precision = nat.resolve_precision(precision)
        self.arena: Optional[FactorArena] = None

    # -- helpers ---------------------------------------------------------------------------------------
    def _selected(self):
        for layer in self.model.modules():
            name = layer.__class__.__name__
            if name in self.layer_types:
                yield name, layer

    @staticmethod
    def _layer_dims(layer: Module):
        """(M, K0, has_bias) of a Linear / Conv2d weight viewed as (M, K0)."""
        M = layer.weight.shape[0]
        return M, layer.weight.numel() // M, layer.bias is not None

    @staticmethod
    def _replace(sample: Tensor,
                 weight: Tensor,
                 bias: Tensor = None):
        """Adds `sample` (M, K) to the parameters: last column to `bias`, the rest to `weight`
        (reference: curvatures.py:67-82).  Kept for API compatibility; `sample_and_replace` itself writes
        mean + sample from inside the sampling kernel's epilogue."""
        if bias is not None:
            bias.data.add_(sample[:, -1].contiguous().view(*bias.shape))
            sample = sample[:, :-1]
        weight.data.add_(sample.contiguous().view(*weight.shape))

    @abstractmethod
    def update(self, *args: Any, **kwargs: Any):
        raise NotImplementedError

    @abstractmethod
    def invert(self,
               add: Union[float, list, tuple] = 0.,
               multiply: Union[float, list, tuple] = 1.):
        raise NotImplementedError

    @abstractmethod
    def sample(self, layer: Module, noise: Optional[Tensor] = None) -> Tensor:
        raise NotImplementedError

    def _sample_into(self, key, weight: Tensor, bias: Optional[Tensor], mean_w: Tensor, mean_b: Optional[Tensor],
                     noise: Optional[Tensor]):
        """Write mean + sample into (weight, bias).  Default: sample, then add (two passes)."""
        s = self.sample(key, noise)
        weight.data.copy_(mean_w)
        if bias is not None:
            bias.data.copy_(mean_b)
        self._replace(s, weight, bias)

    def _param_names(self):
        """Map id(parameter tensor) -> state_dict key, to find each layer's posterior mean."""
        return {id(v): k for k, v in self.model.state_dict(keep_vars=True).items()}

    def sample_and_replace(self, noise: Optional[Dict] = None):
        """Samples new model parameters and replaces old ones for selected layers, skipping all others
        (reference: curvatures.py:117-129).  Equivalent to reloading the mean (`load_state_dict`) and adding a
        fresh sample to every selected layer; parameters of selected layers are written once, as mean + sample,
        by the sampling kernel.  `noise` optionally maps layer (or 'attn_in'/'attn_out') -> Gaussian tensor."""
        names = self._param_names()
        current = self.model.state_dict(keep_vars=True)
        written = set()
        # estimators whose draw is the two-GEMM kernel (KFAC, EFB) queue their layers here; the whole model is then
        # one C-ABI call (crv_sample_matrix_normal_batch), the layers' GEMM chains overlapping on internal streams
        self._deferred = []
        self._noise_pool = self._draw_noise_pool() if noise is None else None
        copies = []
        try:
            self._sample_and_replace_layers(names, noise, written, copies)
            if self._deferred:
                nat.sample_matrix_normal_batch(self._deferred, _gemm_tier(self.precision))
        finally:
            self._deferred = None
            self._noise_pool = None
        for weight, dense_w in copies:
            weight.data.copy_(dense_w)
        with torch.no_grad():   # everything else: restore the mean, as load_state_dict would (multi-tensor copies)
            rest = [(v.data, self.model_state[k]) for k, v in current.items() if k not in written]
            by_type = {}
            for dst, src in rest:
                by_type.setdefault((dst.dtype, src.dtype), ([], []))
                by_type[(dst.dtype, src.dtype)][0].append(dst)
                by_type[(dst.dtype, src.dtype)][1].append(src)
            for dsts, srcs in by_type.values():
                torch._foreach_copy_(dsts, srcs)

    def _draw_noise_pool(self):
        """Estimators that draw (K, M) Gaussian matrices per layer may return {key: tensor} drawn in one go."""
        return None

    def _pooled_noise(self, key, first, second, noise):
        """(z, z_is_callers): the pre-drawn (and, on tensor-core tiers, already rounded) matrix of `key` if there is one."""
        pool = getattr(self, '_noise_pool', None)
        if noise is None and pool is not None and key in pool:
            return pool[key], None
        return self._noise(first, second, noise), noise is not None

    def _sample_and_replace_layers(self, names, noise, written, copies):
        for name, layer in self._selected():
            if name in ['Linear', 'Conv2d']:
                targets = [(layer, layer.weight, layer.bias)]
            else:
                targets = [('attn_in', layer.in_proj_weight, layer.in_proj_bias),
                           ('attn_out', layer.out_proj.weight, layer.out_proj.bias)]
            for key, weight, bias in targets:
                mean_w = self.model_state[names[id(weight)]]
                mean_b = self.model_state[names[id(bias)]] if bias is not None else None
                z = None if noise is None else noise[key]
                if weight.is_contiguous():
                    self._sample_into(key, weight, bias, mean_w, mean_b, z)
                else:
                    # e.g. a channels-last convolution weight: the kernel writes the logical (M, K0) row-major
                    # matrix, so sample into a dense buffer and let copy_ apply the parameter's strides
                    dense_w = torch.empty(weight.shape, dtype=weight.dtype, device=weight.device)
                    self._sample_into(key, _DenseTarget(dense_w), bias, mean_w.contiguous(), mean_b, z)
                    copies.append((weight, dense_w))
                written.add(names[id(weight)])
                if bias is not None:
                    written.add(names[id(bias)])


class Diagonal(Curvature):
    r"""Diagonal Fisher: `state[layer] += batch_size * [wgrad | bgrad]**2` (reference: curvatures.py:132-193),
    through the streaming kernel `crv_diag_accum` (K2)."""

    def _entries(self):
        """(key, weight, bias) for every selected parameter group, in `model.modules()` order."""
        for name, layer in self._selected():
            if name in ['Linear', 'Conv2d']:
                yield layer, layer.weight, layer.bias
            elif name == 'MultiheadAttention':
                yield 'attn_in', layer.in_proj_weight, layer.in_proj_bias
                yield 'attn_out', layer.out_proj.weight, layer.out_proj.bias

    def _ensure_arena(self):
        if self.arena is None:
            entries = list(self._entries())
            shapes = [(w.shape[0], w.numel() // w.shape[0] + (b is not None)) for _, w, b in entries]
            self.arena = FactorArena(shapes, entries[0][1].device)
            self._views = {key: v for (key, _, _), v in zip(entries, self.arena.views)}

    def update(self,
               batch_size: int):
        """Accumulates the squared gradients of all selected layers (call after `backward`)."""
        self._ensure_arena()
        entries = []
        for key, weight, bias in self._entries():
            wg = _grad_of(weight, 'weight')
            bg = _grad_of(bias, 'bias') if bias is not None else None
            if key not in self.state:
                self.state[key] = self._views[key]
            entries.append((wg, bg, self.state[key], None))
        nat.diag_accum_batch(entries, batch_size)      # one launch for the whole model (crv_diag_accum_batch)

    def invert(self,
               add: Union[float, list, tuple] = 0.,
               multiply: Union[float, list, tuple] = 1.):
        assert self.state, "State dict is empty. Did you call 'update' prior to this?"
        if self.inv_state:
            Warning("State has already been inverted. Is this expected?")
        for index, (layer, value) in enumerate(self.state.items()):
            n, s = _damping(add, multiply, index, len(self.state))
            out = torch.empty_like(value)
            nat.elementwise_inv_sqrt(value, n, s, out)
            self.inv_state[layer] = out

    def sample(self,
               layer: Union[Module, str],
               noise: Optional[Tensor] = None) -> Tensor:
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        inv = self.inv_state[layer]
        z = torch.empty_like(inv).normal_() if noise is None else noise
        out = torch.empty_like(inv)
        nat.diag_sample(z, inv, False, s_out=out)
        return out

    def _sample_into(self, key, weight, bias, mean_w, mean_b, noise):
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        inv = self.inv_state[key]
        z = torch.empty_like(inv).normal_() if noise is None else noise
        nat.diag_sample(z, inv, bias is not None, mu_w=mean_w, mu_b=mean_b, w_out=weight.data,
                        b_out=None if bias is None else bias.data)


# factors with fewer multiply-adds than this (R D^2) ride in the one-launch dense batch (K1f) when the TMA-fed kernel
# cannot take them; larger ones keep their own launch (tensor cores on the 1e-3 tiers)
_DENSE_BATCH_FLOPS = 5e8
# sharded `invert`: models whose inverse factors have at least this many floats (1 GiB) exchange them in two overlapped
# rounds (8 GPUs: ResNet-152, 1.9 GB, 19.6 -> 17.2 ms; ResNet-50, 0.7 GB, is 0.6 ms slower that way and stays on one)
_TWO_ROUND_FLOATS = 1 << 28


class KFAC(Curvature):
    r"""Kronecker-factored Fisher (reference: curvatures.py:264-392).

    `update` computes, per selected layer, A = X X^T / R from the recorded layer input (implicit im2col for
    convolutions, trailing ones row for the bias) and G = g g^T * N^2 / R from the recorded output gradient, and
    adds both to the running sums -- one fused kernel per factor (K1), accumulating straight into the arena.
    `invert` is one batched damped Cholesky-of-inverse call (K4); `sample` is the fused two-GEMM matrix-normal
    draw (K5)."""

    def __init__(self,
                 model: Union[Module, Sequential],
                 layer_types: Union[List[str], str] = None,
                 precision: Union[str, int, None] = None):
        super().__init__(model, layer_types, precision)
        self.hooks = list()
        self.record = dict()

        for layer in model.modules():
            if layer.__class__.__name__ in self.layer_types:
                if layer.__class__.__name__ in ['Linear', 'Conv2d']:
                    if layer.__class__.__name__ == 'Conv2d':
                        self._check_conv(layer)
                    self.record[layer] = [None, None]
                    self.hooks.append(layer.register_forward_pre_hook(self._save_input))
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        # same (non-full) hook kind as the reference, curvatures.py:302
                        self.hooks.append(layer.register_backward_hook(self._save_output))
                elif layer.__class__.__name__ == 'MultiheadAttention':
                    raise NotImplementedError

    @staticmethod
    def _check_conv(layer):
        # the reference calls F.unfold(x, kernel_size, padding=, stride=) only (curvatures.py:329)
        if _pair(layer.dilation) != (1, 1) or layer.groups != 1:
            raise NotImplementedError("KFAC supports Conv2d with dilation=1 and groups=1 only (as the reference)")
        if isinstance(layer.padding, str) or layer.padding_mode != 'zeros':
            raise NotImplementedError("KFAC supports integer zero padding only (as the reference)")

    # hooks: pointer stashes only -- no kernel launch, no copy (they also fire on eval forwards)
    def _save_input(self, module, input):
        self.record[module][0] = input[0]

    def _save_output(self, module, grad_input, grad_output):
        self.record[module][1] = grad_output[0]

    _save_grad_output = _save_output

    def scaled_record(self, layer: Module) -> Tensor:
        """The tensor the reference stores in `record[layer][1]` (curvatures.py:310)."""
        g = self.record[layer][1]
        return g * g.size(0)

    def _ensure_arena(self):
        if self.arena is None:
            layers = list(self.record.keys())
            shapes = []
            for layer in layers:
                M, K0, hb = self._layer_dims(layer)
                shapes += [(K0 + hb, K0 + hb), (M, M)]
            self.arena = FactorArena(shapes, layers[0].weight.device)
            self._views = {l: [self.arena.views[2 * i], self.arena.views[2 * i + 1]] for i, l in enumerate(layers)}

    def update(self,
               batch_size: int):
        """Adds the Kronecker factors of the recorded batch to the running sums of every selected layer.
        `batch_size` is accepted for API compatibility; like the reference the factors are normalised by the
        recorded tensors' own shapes."""
        self._ensure_arena()
        device = None
        if self.record:
            device = next(iter(self.record)).weight.device
        # The geometry of an estimation pass repeats every step: the C-ABI item array, which factors ride in the batch
        # call and the workspace size are built once per (shapes, strides, alignment, tier) signature; afterwards a step
        # only refreshes the operand pointers.
        recorded = []
        for layer in self.model.modules():
            module_class = layer.__class__.__name__
            if module_class in self.layer_types:
                if module_class in ['Linear', 'Conv2d']:
                    forward, backward = self.record[layer]
                    if forward is None or backward is None:
                        raise RuntimeError("KFAC.update: no recorded input / output gradient for "
                                           f"{module_class}; run a forward and backward pass first")
                    recorded.append((layer, forward.detach(), backward.detach()))   # any dense layout: the binding picks
                elif module_class == 'MultiheadAttention':                          # the kernel (and copies only if none applies)
                    raise NotImplementedError
        sig = (self.precision,) + tuple((x.shape, x.stride(), x.dtype, x.data_ptr() & 15, g.shape, g.stride(), g.dtype,
                                         g.data_ptr() & 15) for _, x, g in recorded)
        plan = self.__dict__.get('_update_plan')
        if plan is None or plan['sig'] != sig:
            plan = self._plan_update(recorded, sig)
            self._update_plan = plan
        arr = plan['arr']
        for slot, li, which in plan['slots']:                  # refresh the operand pointers of the batch items
            arr[slot].x = recorded[li][1 + which].data_ptr()
        try:
            if plan['dense'] is not None:                      # small factors the channels-last kernel cannot take: one launch
                darr, dslots = plan['dense']
                for slot, li, which in dslots:
                    darr[slot].x = recorded[li][1 + which].data_ptr()
                nat.syrk_batch_dense(darr, len(dslots), device)
            for li, which, args in plan['fallback']:           # the others, one by one
                t = recorded[li][1 + which]
                if args[0] == 'conv':
                    nat.syrk_conv_accum(t, *args[1:], join=False)
                else:
                    nat.syrk_rows_accum(t, *args[1:], join=False)
            if plan['n']:
                # every recorded tensor is complete on the current stream: between this fork and the join below the
                # library runs pre-passes, contractions and reductions on its own prioritised streams
                nat.stream_fork(device)
                nat.syrk_batch_arr(arr, plan['n'], plan['ws_bytes'], self.precision, device, join=False)
        finally:
            # the split reductions run on the library's side stream: order the caller's stream after them (also when
            # a launch failed: the fork must not stay open)
            if device is not None:
                nat.stream_join(device)

    def _plan_update(self, recorded, sig):
        """Which factor goes where (see `update`): the crv_syrk_item array of the batch call, the (slot, layer, operand)
        triples whose pointer must be refreshed every step, the one-launch dense batch of small factors and the
        per-factor calls for everything else."""
        specs = []        # (li, which, tensor, geometry or None, has_bias, alpha, out, zero_mean, R)
        for li, (layer, x, g) in enumerate(recorded):
            if layer not in self.state:
                self.state[layer] = self._views[layer]
            first, second = self.state[layer]
            has_bias = layer.bias is not None
            n_g = g.size(0)
            if layer.__class__.__name__ == 'Conv2d':
                N, _, H, W = x.shape
                kh, kw = _pair(layer.kernel_size)
                sh, sw = _pair(layer.stride)
                ph, pw = _pair(layer.padding)
                r_x = N * ((H + 2 * ph - kh) // sh + 1) * ((W + 2 * pw - kw) // sw + 1)
                specs.append((li, 0, x, ((kh, kw), (sh, sw), (ph, pw)), has_bias, 1.0 / r_x, first, None, r_x))
                r_g = n_g * g.shape[2] * g.shape[3]
            else:
                if x.dim() != 2:
                    raise NotImplementedError("KFAC supports 2-D Linear inputs only (as the reference)")
                specs.append((li, 0, x, None, has_bias, 1.0 / x.size(0), first, False, x.size(0)))
                r_g = n_g
            # reference: (g * N)(g * N)^T / R  ==  g g^T * N^2 / R
            specs.append((li, 1, g, None, False, float(n_g) * float(n_g) / float(r_g), second, None, r_g))

        def dense_of(spec):
            li, which, t, geom, has_bias, alpha, out, _, R = spec
            if float(R) * out.shape[0] * out.shape[0] > _DENSE_BATCH_FLOPS:
                return None
            return nat.dense_item(t, *(geom if geom is not None else (None, None, None)), has_bias, alpha, out)

        # A SMALL model (LeNet-5, BASELINE configs[0]: ten factors, 0.3 GFLOP per update) is launch-bound on any
        # per-factor or per-group path: all of its factors go into ONE launch of the exact-fp32 CUDA-core kernel (K1f;
        # exact products satisfy every tier).
        all_dense = [dense_of(sp) for sp in specs]
        if len(specs) >= 2 and all(d is not None for d in all_dense):
            arr, ws_bytes = nat.syrk_item_array([], self.precision)
            dslots = [(i, sp[0], sp[1]) for i, sp in enumerate(specs)]
            return {'sig': sig, 'arr': arr, 'n': 0, 'ws_bytes': ws_bytes, 'slots': [], 'fallback': [],
                    'dense': (nat.syrk_dense_array(all_dense), dslots)}

        items, slots, fallback, dense, dslots = [], [], [], [], []
        simt_tier = self.precision in (nat.PREC_FP32, nat.PREC_BF16X3)
        for sp, d in zip(specs, all_dense):
            li, which, t, geom, has_bias, alpha, out, zero_mean, R = sp
            kw = {} if zero_mean is None else {'zero_mean': zero_mean}
            item = nat.nhwc_item(t, *(geom if geom is not None else (None, None, None)), has_bias, alpha, out, self.precision, **kw)
            if item is not None:
                slots.append((len(items), li, which))
                items.append(item)
            elif d is not None and simt_tier:   # small dense operands the channels-last kernel cannot take share one launch
                                                # too (on the tiers whose per-factor fallback is that CUDA-core kernel anyway;
                                                # the 1e-3 tiers keep the thread-staged tensor-core kernel, whose fixed-order
                                                # reduction gives exactly symmetric factors)
                dslots.append((len(dense), li, which))
                dense.append(d)
            elif geom is not None:
                fallback.append((li, which, ('conv', *geom, has_bias, alpha, out, self.precision)))
            else:
                fallback.append((li, which, ('rows', has_bias, alpha, out, self.precision)))
        if len(dense) == 1:               # (a single one: its own call, which may use the tensor cores)
            _, li, which = dslots[0]
            sp = next(s_ for s_ in specs if s_[0] == li and s_[1] == which)
            geom = sp[3]
            fallback.append((li, which, ('conv', *geom, sp[4], sp[5], sp[6], self.precision) if geom is not None
                             else ('rows', sp[4], sp[5], sp[6], self.precision)))
            dense, dslots = [], []
        arr, ws_bytes = nat.syrk_item_array(items, self.precision)
        return {'sig': sig, 'arr': arr, 'n': len(items), 'ws_bytes': ws_bytes, 'slots': slots, 'fallback': fallback,
                'dense': (nat.syrk_dense_array(dense), dslots) if dense else None}

    def invert(self,
               add: Union[float, list, tuple] = 0.,
               multiply: Union[float, list, tuple] = 1.,
               group=None,
               shard: Optional[bool] = None):
        """Damped Cholesky factors of the inverse factors (reference: curvatures.py:354-385), one batched kernel call (K4).

        With `torch.distributed` initialised on more than one rank (and `shard` not False) the factors are sharded over
        the ranks of `group` by their D^3 cost, every rank inverts its own and ONE all-gather of the rank-major inverse
        arena gives every rank all of `inv_state` (the state must already be merged, see
        `allreduce_arena`).  Arenas of 1 GiB and more are exchanged in two rounds, the first beside the inversion of every
        rank's largest matrix (`parallel.invert_plan_two_rounds`).  The result is the same as every rank inverting
        everything."""
        assert self.state, "State dict is empty. Did you call 'update' prior to this?"
        if self.inv_state:
            Warning("State has already been inverted. Is this expected?")
        factors, adds, muls = [], [], []
        for index, (layer, value) in enumerate(self.state.items()):
            n, s = _damping(add, multiply, index, len(self.state))
            first, second = value
            factors += [first, second]
            adds += [n, n]
            muls += [s, s]
        import torch.distributed as dist
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        if shard is None:
            shard = world > 1
        if shard and world > 1:
            from .parallel import invert_plan, invert_plan_two_rounds, allgather_segments
            rank = dist.get_rank(group)
            dims = [f.shape[0] for f in factors]
            dev = factors[0].device
            # large models on GPUs: two rounds, the exchange of everything but each rank's largest inverse overlaps the
            # inversion of that largest matrix (see invert_plan_two_rounds); otherwise one round, one all-gather
            two = (dev.type == 'cuda' and sum(d * d for d in dims) >= _TWO_ROUND_FLOATS and
                   os.environ.get("CURVATURE_B200_INVERT_ROUNDS", "2") != "1")
            plan = invert_plan_two_rounds(dims, world) if two else invert_plan(dims, world)
            flat = torch.zeros(plan["total"], dtype=factors[0].dtype, device=dev)
            views = [flat[o:o + f.numel()].view(f.shape) for o, f in zip(plan["offset"], factors)]
            info = torch.zeros(len(factors), dtype=torch.int32, device=dev)

            def run(mine, ws=None):
                if mine:
                    info[mine] = nat.chol_inv_batched([factors[i] for i in mine], [adds[i] for i in mine],
                                                      [muls[i] for i in mine], [views[i] for i in mine], ws=ws)
            if two:
                late = plan["late"]
                mine_late = [i for i, r in enumerate(plan["owner"]) if r == rank and late[i]]
                main = torch.cuda.current_stream(dev)
                side = self.__dict__.get('_invert_streams')
                if side is None or side[0].device != dev:
                    side = self._invert_streams = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
                big, comm = side
                flat.record_stream(big)
                flat.record_stream(comm)
                # the rank's largest matrix on its own stream with its own workspace, BESIDE the others (together they
                # take what one batched call takes: the big one is a long chain of small panels, the others fill the GPU)
                nb = nat.chol_inv_workspace_bytes([dims[i] for i in mine_late]) if mine_late else 0
                ws_big = self.__dict__.get('_invert_ws')
                if mine_late and (ws_big is None or ws_big.numel() < nb or ws_big.device != dev):
                    ws_big = self._invert_ws = torch.empty(nb, dtype=torch.uint8, device=dev)
                big.wait_stream(main)
                info_big = None
                if mine_late:
                    with torch.cuda.stream(big):
                        info_big = nat.chol_inv_batched([factors[i] for i in mine_late], [adds[i] for i in mine_late],
                                                        [muls[i] for i in mine_late], [views[i] for i in mine_late], ws=ws_big)
                run([i for i, r in enumerate(plan["owner"]) if r == rank and not late[i]])
                comm.wait_stream(main)
                with torch.cuda.stream(comm):                    # exchange step 1 (everything but the largest matrices) ...
                    allgather_segments(flat[:plan["base2"]], plan["segment"], group)
                main.wait_stream(big)                            # ... while those are still being inverted
                if info_big is not None:
                    info[mine_late] = info_big
                allgather_segments(flat[plan["base2"]:], plan["segment2"], group)           # exchange step 2
                main.wait_stream(comm)
            else:
                run([i for i, r in enumerate(plan["owner"]) if r == rank])
                allgather_segments(flat, plan["segment"], group)          # the one exchange step
            dist.all_reduce(info, op=dist.ReduceOp.MAX, group=group)  # (count ints: every rank raises on any failure)
            inv_arena = FactorArena.__new__(FactorArena)
            inv_arena.flat, inv_arena.views = flat, views
        else:
            inv_arena = FactorArena([f.shape for f in factors], factors[0].device)
            info = nat.chol_inv_batched(factors, adds, muls, inv_arena.views)
        bad = torch.nonzero(info).flatten().tolist()   # one sync per invert; invert is not the hot loop
        if bad:
            raise RuntimeError("KFAC.invert: damped factor is not positive definite for matrices "
                               f"{bad[:8]} (factor index = 2*layer + {{0: A, 1: G}}); increase `add`. "
                               "(The reference falls back to numpy here; this implementation has no CPU path.)")
        self._inv_arena = inv_arena
        self.__dict__.pop('_inv_tf32', None)           # rounded copies of the previous inverse factors are stale
        for i, layer in enumerate(self.state.keys()):
            self.inv_state[layer] = (inv_arena.views[2 * i], inv_arena.views[2 * i + 1])

    def _draw_noise_pool(self):
        """One randn (and one TF32 rounding) for all layers of a sample_and_replace instead of one per layer."""
        return _flat_noise({k: (f.size(0), s.size(0)) for k, (f, s) in self.inv_state.items()},
                           next(iter(self.inv_state.values()))[0], _gemm_tier(self.precision)) if self.inv_state else None

    def _noise(self, first: Tensor, second: Tensor, noise: Optional[Tensor]) -> Tensor:
        if noise is None:   # same draw as the reference (curvatures.py:391)
            return torch.randn(first.size(0), second.size(0), device=first.device, dtype=first.dtype)
        return noise if noise.is_contiguous() else noise.contiguous()

    def sample(self,
               layer: Module,
               noise: Optional[Tensor] = None) -> Tensor:
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        first, second = self.inv_state[layer]
        z = self._noise(first, second, noise)
        out = torch.empty(second.size(0), first.size(0), device=first.device, dtype=first.dtype)
        tier, first, second, z = self._gemm_operands(layer, first, second, z, noise is not None)
        nat.sample_matrix_normal(second, first, z, False, s_out=out, precision=tier)
        return out

    def sample_and_replace(self, noise: Optional[Dict] = None):
        """As `Curvature.sample_and_replace` (reference: curvatures.py:117-129).  Without supplied noise the whole call
        is prebuilt once per `invert`: one persistent noise buffer (drawn and rounded in place), one C-ABI item array for
        the batched two-GEMM draw of all layers, the multi-tensor copies that restore every other parameter and buffer
        to its mean -- a posterior sample then costs four launches and no per-layer Python."""
        if noise is not None:
            return super().sample_and_replace(noise)
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        tier = _gemm_tier(self.precision)
        key = (id(self.__dict__.get('_inv_arena')), tier, tuple(p.data_ptr() for p in self.model.parameters()))
        plan = self.__dict__.get('_sample_plan')
        if plan is None or plan['key'] != key:
            plan = self._plan_sample(tier, key)
            self._sample_plan = plan
        flat = plan['noise']
        flat.normal_()                               # same distribution as the reference's per-layer randn (:391)
        if tier != nat.PREC_FP32:
            nat.round_tf32(flat, out=flat)
        nat.run_sample_batch(plan['arr'], plan['n'], plan['ws_bytes'], tier, flat)
        with torch.no_grad():
            for weight, dense_w in plan['copies']:
                weight.data.copy_(dense_w)
            for dsts, srcs in plan['rest']:          # everything else: restore the mean, as load_state_dict would
                torch._foreach_copy_(dsts, srcs)

    def _plan_sample(self, tier, key):
        names = self._param_names()
        current = self.model.state_dict(keep_vars=True)
        offs, total = {}, 0
        for layer, (first, second) in self.inv_state.items():
            offs[layer] = total
            total += (first.size(0) * second.size(0) + 3) // 4 * 4
        like = next(iter(self.inv_state.values()))[0]
        flat = torch.empty(total, device=like.device, dtype=like.dtype)
        entries, copies, written = [], [], set()
        for name, layer in self._selected():
            if name not in ['Linear', 'Conv2d']:
                continue
            weight, bias = layer.weight, layer.bias
            first, second = self.inv_state[layer]
            K, M = first.size(0), second.size(0)
            z = flat[offs[layer]:offs[layer] + K * M].view(K, M)
            _, la, lg, _ = self._gemm_operands(layer, first, second, None, True)
            mean_w = self.model_state[names[id(weight)]]
            mean_b = self.model_state[names[id(bias)]] if bias is not None else None
            if weight.is_contiguous():
                w_out = weight.data
            else:                      # e.g. a channels-last convolution weight: sample into a dense buffer, copy_ applies
                w_out = torch.empty(weight.shape, dtype=weight.dtype, device=weight.device)      # the parameter's strides
                mean_w = mean_w.contiguous()
                copies.append((weight, w_out))
            entries.append(dict(LG=lg, LA=la, z=z, has_bias=bias is not None, mu_w=mean_w, mu_b=mean_b, w_out=w_out,
                                b_out=None if bias is None else bias.data))
            written.add(names[id(weight)])
            if bias is not None:
                written.add(names[id(bias)])
        by_type = {}
        for k, v in current.items():
            if k not in written:
                src = self.model_state[k]
                by_type.setdefault((v.dtype, src.dtype), ([], []))
                by_type[(v.dtype, src.dtype)][0].append(v.data)
                by_type[(v.dtype, src.dtype)][1].append(src)
        arr, ws_bytes = nat.prepare_sample_batch(entries)
        return {'key': key, 'noise': flat, 'arr': arr, 'n': len(entries), 'ws_bytes': ws_bytes, 'copies': copies,
                'rest': list(by_type.values()), 'keep': entries}

    def sample_many(self,
                    samples: int,
                    noise: Optional[Dict] = None) -> Dict[Module, Tensor]:
        """`samples` posterior draws of EVERY layer at once: {layer: (samples, M, K) tensor}, draw s of a layer being what
        `sample(layer)` returns (reference: curvatures.py:387-392) for the s-th noise matrix.  The S draws of a layer are
        one pair of GEMMs and all layers share one kernel launch (crv_sample_matrix_normal_multi) -- the sampling half of
        the BNN evaluation loop (scripts/evaluate.py:121-152).  `noise` optionally maps layer -> (samples, K, M)."""
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        tier = _gemm_tier(self.precision)
        entries, outs = [], {}
        for layer, (first, second) in self.inv_state.items():
            K, M = first.size(0), second.size(0)
            if noise is None:
                z = torch.randn(samples * K, M, device=first.device, dtype=first.dtype)
                callers = False
            else:
                z = noise[layer].reshape(samples * K, M)
                z = z if z.is_contiguous() else z.contiguous()
                callers = True
            _, la, lg, z = self._gemm_operands(layer, first, second, z, callers)
            out = torch.empty(M, samples, K, device=first.device, dtype=first.dtype)
            entries.append((lg, la, z, out))
            outs[layer] = out.permute(1, 0, 2)          # (samples, M, K) view
        nat.sample_matrix_normal_multi(entries, samples, tier)
        return outs

    def replace_with(self, draws: Dict[Module, Tensor], index: int):
        """Install draw `index` of `sample_many`: selected layers get mean + sample, every other parameter and buffer
        its mean -- the state `sample_and_replace` leaves behind (reference: curvatures.py:117-129)."""
        names = self._param_names()
        current = self.model.state_dict(keep_vars=True)
        written = set()
        with torch.no_grad():
            for layer, d in draws.items():
                s = d[index]
                weight, bias = layer.weight, layer.bias
                mean_w = self.model_state[names[id(weight)]]
                if bias is not None:
                    bias.data.copy_(self.model_state[names[id(bias)]])
                    bias.data.add_(s[:, -1])
                    s = s[:, :-1]
                    written.add(names[id(bias)])
                weight.data.copy_(mean_w)
                weight.data.add_(s.reshape(weight.shape))
                written.add(names[id(weight)])
            rest = [(v.data, self.model_state[k]) for k, v in current.items() if k not in written]
            by_type = {}
            for dst, src in rest:
                by_type.setdefault((dst.dtype, src.dtype), ([], []))
                by_type[(dst.dtype, src.dtype)][0].append(dst)
                by_type[(dst.dtype, src.dtype)][1].append(src)
            for dsts, srcs in by_type.values():
                torch._foreach_copy_(dsts, srcs)

    def _gemm_operands(self, key, first, second, z, z_is_callers):
        """Operands of the two-GEMM draw for this estimator's tier: tensor-core tiers get the inverse factors rounded
        to TF32 once per `invert` (cached) and the noise rounded (a copy if the caller owns it)."""
        tier = _gemm_tier(self.precision)
        if tier != nat.PREC_FP32:
            first, second = _rounded(self.__dict__.setdefault('_inv_tf32', {}), key, (first, second))
            if z is not None:
                z = nat.round_tf32(z, out=None if z_is_callers else z)
        return tier, first, second, z

    def _sample_into(self, key, weight, bias, mean_w, mean_b, noise):
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        first, second = self.inv_state[key]
        z, callers = self._pooled_noise(key, first, second, noise)
        if callers is None:        # pre-drawn and already rounded
            tier, first, second, _ = self._gemm_operands(key, first, second, None, True)
        else:
            tier, first, second, z = self._gemm_operands(key, first, second, z, callers)
        if getattr(self, '_deferred', None) is not None:
            self._deferred.append(dict(LG=second, LA=first, z=z, has_bias=bias is not None, mu_w=mean_w, mu_b=mean_b,
                                       w_out=weight.data, b_out=None if bias is None else bias.data))
            return
        nat.sample_matrix_normal(second, first, z, bias is not None, mu_w=mean_w, mu_b=mean_b,
                                 w_out=weight.data, b_out=None if bias is None else bias.data,
                                 precision=tier)


class EFB(Curvature):
    """Eigenvalue-corrected Kronecker factorisation (reference: curvatures.py:395-460):
    `state[layer] += (QG^T [wgrad|bgrad] QA)**2`, `diags[layer] += batch_size * [wgrad|bgrad]**2`."""

    def __init__(self,
                 model: Union[Module, Sequential],
                 factors: Dict[Module, Tensor],
                 layer_types: Union[List[str], str] = None,
                 precision: Union[str, int, None] = None,
                 eigvecs: Optional[Dict] = None):
        """`eigvecs` optionally supplies precomputed eigenbases {layer: (QA, QG)} (eigenvectors are only defined
        up to sign / rotation inside degenerate eigenspaces, so parity tests feed the oracle's)."""
        super().__init__(model, layer_types, precision)
        self.eigvecs = get_eigenvectors(factors) if eigvecs is None else eigvecs
        self.diags = dict()

    def _ensure_arena(self):
        if self.arena is None:
            layers = [l for n, l in self._selected() if n in ['Linear', 'Conv2d']]
            shapes = []
            for layer in layers:
                M, K0, hb = self._layer_dims(layer)
                shapes += [(M, K0 + hb), (M, K0 + hb)]
            self.arena = FactorArena(shapes, layers[0].weight.device)
            self._views = {l: (self.arena.views[2 * i], self.arena.views[2 * i + 1]) for i, l in enumerate(layers)}

    def update(self,
               batch_size: int):
        self._ensure_arena()
        tier = _gemm_tier(self.precision)
        plan = self.__dict__.get('_update_plan')
        if plan is None or plan['tier'] != tier:
            plan = self._plan_update(tier)
            self._update_plan = plan
        # per step only the gradient pointers change: refresh them in the prebuilt C-ABI item arrays
        keep = []
        darr = plan['diag']
        for i, (weight, bias) in enumerate(plan['params']):
            wg = _grad_of(weight, 'weight')
            if not (wg.is_cuda and wg.dtype == torch.float32):
                nat._dev(wg, 'weight.grad')
            darr[i].wgrad = wg.data_ptr()
            keep.append(wg)
            if bias is not None:
                bg = _grad_of(bias, 'bias')
                if not (bg.is_cuda and bg.dtype == torch.float32):
                    nat._dev(bg, 'bias.grad')
                darr[i].bgrad = bg.data_ptr()
                keep.append(bg)
        # diags += batch_size * g^2 and the concatenated gradient copies, one launch for the whole model
        nat.run_diag_batch(darr, len(plan['params']), batch_size, plan['grads'])
        if tier != nat.PREC_FP32:      # the gradient copies live in one flat buffer: one rounding launch for all layers
            nat.round_tf32(plan['grads'], out=plan['grads'])
        # lambdas += (QG^T g QA)^2 for every layer: one C-ABI call (crv_efb_project_batch)
        nat.run_efb_batch(plan['efb'], len(plan['params']), plan['ws_bytes'], tier, plan['grads'])

    def _plan_update(self, tier):
        layers = []
        for name, layer in self._selected():
            if name in ['Linear', 'Conv2d']:
                layers.append(layer)
            elif name == 'MultiheadAttention':
                raise NotImplementedError
        sizes = []
        for layer in layers:
            if layer not in self.state:
                self.state[layer], self.diags[layer] = self._views[layer]
            sizes.append((self.state[layer].numel() + 63) // 64 * 64)
        dev = self.state[layers[0]].device
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)       # gradient copies [wgrad | bgrad] of all layers
        diag_entries, efb_entries, params, off = [], [], [], 0
        for layer, sz in zip(layers, sizes):
            lam = self.state[layer]
            grads = flat[off:off + lam.numel()].view_as(lam)
            off += sz
            qa, qg = self.eigvecs[layer]
            if tier != nat.PREC_FP32:      # eigenbases rounded to TF32 once
                qa, qg = _rounded(self.__dict__.setdefault('_eig_tf32', {}), layer, (qa, qg))
            diag_entries.append((flat[:layer.weight.numel()].view(layer.weight.shape[0], -1), flat[:lam.shape[0]] if layer.bias is not None
                                 else None, self.diags[layer], grads))
            efb_entries.append((qg, qa, grads, lam))
            params.append((layer.weight, layer.bias))
        darr = nat.prepare_diag_batch(diag_entries)       # (gradient pointers are placeholders: refreshed every step)
        earr, ws_bytes = nat.prepare_efb_batch(efb_entries, round_g=False)
        return {'tier': tier, 'diag': darr, 'efb': earr, 'ws_bytes': ws_bytes, 'params': params, 'grads': flat,
                'keep': (diag_entries, efb_entries)}

    def invert(self,
               add: Union[float, list, tuple] = 0.,
               multiply: Union[float, list, tuple] = 1.):
        assert self.state, "State dict is empty. Did you call 'update' prior to this?"
        if self.inv_state:
            Warning("State has already been inverted. Is this expected?")
        for index, (layer, value) in enumerate(self.state.items()):
            n, s = _damping(add, multiply, index, len(self.state))
            out = torch.empty_like(value)
            nat.elementwise_inv_sqrt(value, n, s, out)
            self.inv_state[layer] = out

    def _draw_noise_pool(self):
        # (EFB's draw scales the noise by the inverse state before the GEMMs: no rounding here)
        return _flat_noise({k: (qa.size(0), qg.size(0)) for k, (qa, qg) in self.eigvecs.items()},
                           next(iter(self.eigvecs.values()))[0], nat.PREC_FP32) if self.eigvecs else None

    def _noise(self, first, second, noise):
        if noise is None:
            return torch.randn(first.size(0), second.size(0), device=first.device, dtype=first.dtype)
        return noise if noise.is_contiguous() else noise.contiguous()

    def sample(self,
               layer: Module,
               noise: Optional[Tensor] = None) -> Tensor:
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        first, second = self.eigvecs[layer]
        z = self._noise(first, second, noise)
        out = torch.empty(second.size(0), first.size(0), device=first.device, dtype=first.dtype)
        tier = _gemm_tier(self.precision)
        if tier != nat.PREC_FP32:
            first, second = _rounded(self.__dict__.setdefault('_eig_tf32', {}), layer, (first, second))
        nat.sample_matrix_normal(second, first, z, False, row_scale=self.inv_state[layer], s_out=out, precision=tier)
        return out

    def _sample_into(self, key, weight, bias, mean_w, mean_b, noise):
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        first, second = self.eigvecs[key]
        z, _ = self._pooled_noise(key, first, second, noise)
        tier = _gemm_tier(self.precision)
        if tier != nat.PREC_FP32:
            first, second = _rounded(self.__dict__.setdefault('_eig_tf32', {}), key, (first, second))
        if getattr(self, '_deferred', None) is not None:
            self._deferred.append(dict(LG=second, LA=first, z=z, has_bias=bias is not None, row_scale=self.inv_state[key],
                                       mu_w=mean_w, mu_b=mean_b, w_out=weight.data,
                                       b_out=None if bias is None else bias.data))
            return
        nat.sample_matrix_normal(second, first, z, bias is not None, row_scale=self.inv_state[key],
                                 mu_w=mean_w, mu_b=mean_b, w_out=weight.data,
                                 b_out=None if bias is None else bias.data, precision=tier)


class INF(Curvature):
    """Low-rank "information form" estimator (reference: curvatures.py:463-672).

    No data pass and no north-star kernel of its own: the rank selection is host-side index logic exactly as in
    the reference; the dense algebra runs on the device through the library's GEMM (`crv_gemm`), with the
    Kronecker products of the reference replaced by the equivalent small GEMMs
    (`_diagonal_accumulator(QA,QG,l) == ((QA*QA) L (QG*QG)^T).flatten()`); the general (non-symmetric) matrix
    inverses of `pre_sampler` use `torch.linalg` (cuSOLVER), as they are one-shot, LeNet-sized operations."""

    def __init__(self,
                 model: Union[Module, Sequential],
                 diags: Dict[Module, Tensor],
                 factors: Dict[Module, Tensor],
                 lambdas: Dict[Module, Tensor],
                 layer_types: Union[List[str], str] = None,
                 precision: Union[str, int, None] = None,
                 eigvecs: Optional[Dict] = None):
        super().__init__(model, layer_types, precision)
        assert diags.keys() == factors.keys() == lambdas.keys()
        self.eigvecs = get_eigenvectors(factors) if eigvecs is None else eigvecs
        self.lambdas = lambdas
        self.diags = diags

    def update(self,
               rank: int = 100):
        values = zip(list(self.diags.keys()),
                     list(self.eigvecs.values()),
                     list(self.lambdas.values()),
                     list(self.diags.values()))
        for layer, eigvecs, lambdas, diags in values:
            xxt_eigvecs, ggt_eigvecs = eigvecs
            lambda_vec = lambdas.t().contiguous().view(-1)
            diag_vec = diags.t().contiguous().view(-1)
            lr_xxt_eigvecs, lr_ggt_eigvecs, lr_lambda = self._dim_reduction(xxt_eigvecs, ggt_eigvecs, lambda_vec, rank)
            sif_diag = self._diagonal_accumulator(lr_xxt_eigvecs, lr_ggt_eigvecs, lr_lambda)
            self.state[layer] = (lr_xxt_eigvecs, lr_ggt_eigvecs, lr_lambda, diag_vec - sif_diag)

    def invert(self,
               add: Union[float, list, tuple] = 0.,
               multiply: Union[float, list, tuple] = 1.):
        assert self.state, "State dict is empty. Did you call 'update' prior to this?"
        if self.inv_state:
            Warning("State has already been inverted. Is this expected?")
        for index, (layer, value) in enumerate(self.state.items()):
            n, s = _damping(add, multiply, index, len(self.state))
            lr_frst_eigvecs, lr_scnd_eigvecs, lr_lambda, correction = value
            correction[correction < 0] = 0     # in place, like the reference (curvatures.py:523)
            reg_lr_lambda = (s * lr_lambda).sqrt()
            reg_inv_correction = torch.empty_like(correction)
            nat.elementwise_inv_sqrt(correction, n, s, reg_inv_correction)
            pre_sample = self.pre_sampler(lr_frst_eigvecs, lr_scnd_eigvecs, reg_lr_lambda, reg_inv_correction)
            self.inv_state[layer] = (lr_frst_eigvecs, lr_scnd_eigvecs, reg_inv_correction, pre_sample)

    def sample(self,
               layer: Module,
               noise: Optional[Tensor] = None) -> Tensor:
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        a, b, c, d = self.inv_state[layer]
        return self.sampler(a, b, c, d, noise).reshape(a.shape[0], b.shape[0]).t()

    @staticmethod
    def pre_sampler(frst_eigvecs: Tensor,
                    scnd_eigvecs: Tensor,
                    reg_lambda: Tensor,
                    reg_inv_correction: Tensor) -> Tensor:
        """reference: curvatures.py:538-572.  V = diag(c) kron(QA,QG) diag(l) is never formed:
        V^T V = (QA^T diag-weighted QA) combined with QG through the per-row weights c^2 reshaped (K, M):
        (V^T V)[(a,b),(a',b')] = l_ab l_a'b' sum_{k,m} c_km^2 QA[k,a] QA[k,a'] QG[m,b] QG[m,b']."""
        K, ra = frst_eigvecs.shape
        M, rg = scnd_eigvecs.shape
        c2 = (reg_inv_correction ** 2).view(K, M).contiguous()
        # pair products of eigenvector columns: PA[k,(a,a')] = QA[k,a] QA[k,a'];  PG[m,(b,b')] likewise
        PA = (frst_eigvecs[:, :, None] * frst_eigvecs[:, None, :]).reshape(K, ra * ra).contiguous()
        PG = (scnd_eigvecs[:, :, None] * scnd_eigvecs[:, None, :]).reshape(M, rg * rg).contiguous()
        T = nat.gemm(c2, PG)                                  # (K, rg*rg)
        W = nat.gemm(PA, T, transA=True)                      # (ra*ra, rg*rg): W[(a,a'),(b,b')]
        vtv = W.view(ra, ra, rg, rg).permute(0, 2, 1, 3).reshape(ra * rg, ra * rg)
        vtv = reg_lambda[:, None] * vtv * reg_lambda[None, :]
        vtv = (vtv + vtv.t()) / 2.
        # The r x r chain below (Cholesky + three explicit inverses, curvatures.py:564-570) has condition numbers
        # of 1e6..1e8 on LeNet-5: evaluated in fp32 it is O(0.1..1) away from the exact result (the reference's own
        # fp32 output is, too).  It is a one-shot, r <= ~1.5k operation, so it runs in fp64 on the device
        # (torch.linalg -> cuSOLVER, the library call the north star allows for one-shot factorizations); measured
        # against the fp64 oracle this is 1e-7 instead of 1e-1.
        v64 = vtv.double()
        eye = torch.eye(v64.shape[0], device=v64.device, dtype=v64.dtype)
        A_c_inv = torch.linalg.inv(torch.linalg.cholesky(v64))
        B_c = torch.linalg.cholesky(v64 + eye)
        C = A_c_inv.t() @ (B_c - eye) @ A_c_inv
        L_c = torch.linalg.inv(torch.linalg.inv(C) + v64).to(vtv.dtype)
        return reg_lambda[:, None] * L_c * reg_lambda[None, :]

    @staticmethod
    def sampler(frst_eigvecs: Tensor,
                scnd_eigvecs: Tensor,
                reg_inv_correction: Tensor,
                pre_sample: Tensor,
                noise: Optional[Tensor] = None) -> Tensor:
        """reference: curvatures.py:574-600."""
        K, M = frst_eigvecs.shape[0], scnd_eigvecs.shape[0]
        X = torch.randn(K * M, device=frst_eigvecs.device, dtype=frst_eigvecs.dtype) if noise is None else noise
        Y_l = reg_inv_correction * X
        unvec_Y_l = Y_l.reshape(M, K)
        Xq = nat.gemm(nat.gemm(scnd_eigvecs, unvec_Y_l, transA=True), frst_eigvecs)          # (rg, ra)
        Qx = nat.gemm(pre_sample, Xq.t().contiguous().view(-1, 1)).view(-1)
        unvec_Qx = Qx.reshape(scnd_eigvecs.shape[1], frst_eigvecs.shape[1])
        X_p_s = nat.gemm(nat.gemm(scnd_eigvecs, unvec_Qx), frst_eigvecs, transB=True)        # (M, K)
        Y_r = reg_inv_correction ** 2 * X_p_s.t().contiguous().view(-1)
        return Y_l - Y_r

    @staticmethod
    def _dim_reduction(frst_eigvecs: Tensor,
                       scnd_eigvecs: Tensor,
                       lambda_vec: Tensor,
                       rank: int):
        """reference: curvatures.py:602-647 (host-side index selection, 1-based like the reference)."""
        if rank >= lambda_vec.shape[0]:
            return frst_eigvecs, scnd_eigvecs, lambda_vec
        m = scnd_eigvecs.shape[1]
        order = (torch.argsort(-torch.abs(lambda_vec))[:rank] + 1).tolist()
        idx_left = sorted({int((t - 1.) / m + 1.) for t in order})
        idx_right = sorted({t - m * (int((t - 1.) / m + 1.) - 1) for t in order})
        dev = lambda_vec.device
        picks = torch.tensor([m * (i - 1) + j - 1 for i in idx_left for j in idx_right], device=dev)
        left = torch.tensor([i - 1 for i in idx_left], device=dev)
        right = torch.tensor([j - 1 for j in idx_right], device=dev)
        return (frst_eigvecs.index_select(1, left).contiguous(),
                scnd_eigvecs.index_select(1, right).contiguous(),
                lambda_vec.index_select(0, picks))

    @staticmethod
    def _diagonal_accumulator(xxt_eigvecs: Tensor,
                              ggt_eigvecs: Tensor,
                              lambda_vec: Tensor):
        """reference: curvatures.py:649-672, kron-free: out[i*m + p] = sum_{a,b} QA[i,a]^2 QG[p,b]^2 l[a*rg+b]."""
        ra, rg = xxt_eigvecs.shape[1], ggt_eigvecs.shape[1]
        lam = lambda_vec.view(ra, rg).contiguous()
        qa2 = (xxt_eigvecs ** 2).contiguous()
        qg2 = (ggt_eigvecs ** 2).contiguous()
        return nat.gemm(nat.gemm(qa2, lam), qg2, transB=True).view(-1)
