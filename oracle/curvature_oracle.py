"""CPU oracle for the Fisher-estimation hot path of DLR-RM/curvature.

TEST INFRASTRUCTURE ONLY.  Nothing under ``curvature_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker (or as the timed CPU baseline), never as the product path.

It is a *restatement* of the reference algorithm (``curvature/curvatures.py``
and ``curvature/utils.py`` of the reference tree), written from the behaviour
documented in SURVEY.md section 8(a), as plain torch-CPU / numpy arithmetic.
The arithmetic of the reference lives in PyTorch ATen (pin: ``torch>=1.6.0``,
reference ``setup.py:25``; executed here with torch 2.11.0), which is not part
of the reference tree.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the *real*
reference (imported from ``/root/reference`` in the build container) on seeded
inputs, checks this restatement against it, and commits the reference's outputs
as fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks
the restatement against those fixtures wherever the suite runs.  The only
known-answer test the reference itself ships for this path is the ``kron``
doctest (``curvature/utils.py:301-308``); it is reproduced in the tests too.

Every function cites the reference lines it follows.  ``dtype`` can be raised
to float64 to obtain a high-precision truth for tolerance studies; float32
reproduces the reference bit-for-bit on CPU.
"""
from __future__ import annotations

import copy
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Tensor
from torch.nn import Module

SUPPORTED = ("Linear", "Conv2d", "MultiheadAttention")


# ----------------------------------------------------------------------------
# index maps and elementary pieces
# ----------------------------------------------------------------------------
def _pair(v) -> Tuple[int, int]:
    if isinstance(v, (tuple, list)):
        return int(v[0]), int(v[1])
    return int(v), int(v)


def conv_output_hw(H: int, W: int, kernel_size, padding, stride) -> Tuple[int, int]:
    kh, kw = _pair(kernel_size)
    ph, pw = _pair(padding)
    sh, sw = _pair(stride)
    return (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1


def im2col_index_map(C: int, H: int, W: int, kernel_size, padding, stride) -> np.ndarray:
    """Integer index map of the reference's ``F.unfold`` call (curvatures.py:329).

    Returns an int64 array ``idx[K0, L]`` with ``K0 = C*kh*kw`` and ``L = OH*OW``:
    ``idx[k, l]`` is the flat offset into one image ``x[n].reshape(-1)`` that
    patch row ``k = c*kh*kw + i*kw + j`` reads at output position
    ``l = oh*OW + ow``, or ``-1`` where the read falls into the zero padding.
    Dilation 1 and groups 1 only -- the reference passes neither to unfold.
    """
    kh, kw = _pair(kernel_size)
    ph, pw = _pair(padding)
    sh, sw = _pair(stride)
    OH, OW = conv_output_hw(H, W, kernel_size, padding, stride)
    c = np.arange(C)[:, None, None, None, None]
    i = np.arange(kh)[None, :, None, None, None]
    j = np.arange(kw)[None, None, :, None, None]
    oh = np.arange(OH)[None, None, None, :, None]
    ow = np.arange(OW)[None, None, None, None, :]
    ih = oh * sh - ph + i
    iw = ow * sw - pw + j
    inside = (ih >= 0) & (ih < H) & (iw >= 0) & (iw < W)
    flat = (c * H + ih) * W + iw
    idx = np.where(inside, flat, -1)
    idx = np.broadcast_to(idx, (C, kh, kw, OH, OW))
    return idx.reshape(C * kh * kw, OH * OW).astype(np.int64)


def unfold_patches(x: Tensor, kernel_size, padding, stride) -> Tensor:
    """(N,C,H,W) -> (N, C*kh*kw, OH*OW); pure gather through ``im2col_index_map``
    (restates ``F.unfold`` as the reference calls it, curvatures.py:329)."""
    N, C, H, W = x.shape
    idx = torch.from_numpy(im2col_index_map(C, H, W, kernel_size, padding, stride))
    flat = torch.cat([x.reshape(N, -1), x.new_zeros(N, 1)], dim=1)  # slot -1 -> 0
    gather = torch.where(idx < 0, torch.full_like(idx, C * H * W), idx)
    return flat[:, gather.reshape(-1)].reshape(N, idx.shape[0], idx.shape[1])


def kfac_input_matrix(x: Tensor, layer: Module) -> Tensor:
    """The (K, R) matrix whose Gram matrix is the first Kronecker factor
    (curvatures.py:328-335): unfolded activations with rows ordered
    ``c*kh*kw + i*kw + j``, columns ``n*L + oh*OW + ow``, and a trailing row of
    ones iff the layer has a bias."""
    if layer.__class__.__name__ == "Conv2d":
        cols = unfold_patches(x.detach(), layer.kernel_size, layer.padding, layer.stride)
        mat = cols.permute(1, 0, 2).contiguous().view(cols.shape[1], -1)
    else:
        mat = x.detach().t()
    if layer.bias is not None:
        mat = torch.cat([mat, torch.ones_like(mat[:1])], dim=0)
    return mat


def kfac_grad_matrix(g: Tensor, layer: Module) -> Tensor:
    """The (M, R) matrix behind the second factor (curvatures.py:339-342).
    ``g`` is already scaled by the batch size (curvatures.py:310)."""
    if layer.__class__.__name__ == "Conv2d":
        return g.detach().permute(1, 0, 2, 3).contiguous().view(g.shape[1], -1)
    return g.detach().t()


def gram_over_columns(mat: Tensor) -> Tensor:
    """``mat @ mat.T / R`` (curvatures.py:336, 343)."""
    return torch.mm(mat, mat.t()) / float(mat.shape[1])


def kfac_factors(x: Tensor, g_scaled: Tensor, layer: Module) -> Tuple[Tensor, Tensor]:
    return (gram_over_columns(kfac_input_matrix(x, layer)),
            gram_over_columns(kfac_grad_matrix(g_scaled, layer)))


def layer_grad_matrix(layer: Module) -> Tensor:
    """``[weight.grad.view(M,-1) | bias.grad]`` (curvatures.py:151-153, 424-426)."""
    g = layer.weight.grad.contiguous().view(layer.weight.grad.shape[0], -1)
    if layer.bias is not None:
        g = torch.cat([g, layer.bias.grad.unsqueeze(dim=1)], dim=1)
    return g


def damped_inverse_cholesky(factor: Tensor, add: float, multiply: float) -> Tensor:
    """Lower Cholesky factor of the inverse of the damped, symmetrised factor
    (curvatures.py:368-379): ``reg = sqrt(s)*F + sqrt(n)*I``; ``reg=(reg+reg^T)/2``;
    ``chol(inv(reg))``."""
    eye = torch.diag(factor.new_full((factor.shape[0],), add ** 0.5))
    reg = multiply ** 0.5 * factor + eye
    reg = (reg + reg.t()) / 2.0
    return torch.linalg.cholesky(torch.linalg.inv(reg))


def inv_sqrt_damped(value: Tensor, add, multiply) -> Tensor:
    """``sqrt(1/(s*v+n))`` (curvatures.py:188, 450, 526)."""
    return torch.reciprocal(multiply * value + add).sqrt()


def kron(a: Tensor, b: Tensor) -> Tensor:
    """Kronecker product (utils.py:288-310): out[(a0,c0),(a1,c1)] = a[a0,a1]*b[c0,c1]."""
    return (a[:, None, :, None] * b[None, :, None, :]).reshape(a.shape[0] * b.shape[0],
                                                              a.shape[1] * b.shape[1])


def eigenvectors_of_factors(factors: Dict) -> Dict:
    """utils.py:45-60: eigenvectors (columns, ascending eigenvalues) of F+F^T for
    both factors.  ``torch.symeig`` is gone from torch>=2; ``linalg.eigh`` on the
    upper triangle is the same LAPACK syevd call."""
    out = dict()
    for layer, (xxt, ggt) in factors.items():
        _, qa = torch.linalg.eigh(xxt + xxt.t(), UPLO="U")
        _, qg = torch.linalg.eigh(ggt + ggt.t(), UPLO="U")
        out[layer] = (qa, qg)
    return out


def eigenvalues_of_factors(factors: Iterable) -> Tensor:
    """utils.py:21-42: for a 2-element entry the outer product of the two factor
    spectra (of F itself, not F+F^T), else the entry flattened."""
    chunks = [torch.empty(0)]
    for factor in factors:
        if len(factor) == 2:
            ea = torch.linalg.eigvalsh(factor[0], UPLO="U")
            eg = torch.linalg.eigvalsh(factor[1], UPLO="U")
            chunks.append(torch.outer(ea, eg).reshape(-1))
        else:
            chunks.append(factor.contiguous().view(-1))
    return torch.cat(chunks)


# ----------------------------------------------------------------------------
# estimators (same surface as the reference classes; ``noise`` is injectable)
# ----------------------------------------------------------------------------
def _damping(add, multiply, index: int, count: int, scalar_types) -> Tuple[float, float]:
    if not isinstance(add, scalar_types) and not isinstance(multiply, scalar_types):
        assert len(add) == len(multiply) == count
        return add[index], multiply[index]
    return add, multiply


class Curvature:
    """curvatures.py:17-129."""

    def __init__(self, model: Module, layer_types: Union[List[str], str, None] = None):
        self.model = model
        self.model_state = copy.deepcopy(model.state_dict())
        if isinstance(layer_types, str):
            self.layer_types = [layer_types]
        elif isinstance(layer_types, list):
            self.layer_types = list(layer_types) if layer_types else list(SUPPORTED)
        elif layer_types is None:
            self.layer_types = list(SUPPORTED)
        else:
            raise TypeError
        for t in self.layer_types:
            assert t in SUPPORTED
        self.state = dict()
        self.inv_state = dict()

    def _layers(self):
        for layer in self.model.modules():
            name = layer.__class__.__name__
            if name in self.layer_types:
                yield name, layer

    @staticmethod
    def _replace(sample: Tensor, weight: Tensor, bias: Optional[Tensor] = None):
        """curvatures.py:67-82: last column -> bias, the rest -> weight, added in place."""
        if bias is not None:
            bias.data.add_(sample[:, -1].contiguous().view(*bias.shape))
            sample = sample[:, :-1]
        weight.data.add_(sample.contiguous().view(*weight.shape))

    def sample(self, layer, noise: Optional[Tensor] = None) -> Tensor:
        raise NotImplementedError

    def sample_and_replace(self, noise: Optional[Dict] = None):
        """curvatures.py:117-129.  ``noise`` optionally maps layer -> the Gaussian
        tensor the reference would have drawn for that layer."""
        self.model.load_state_dict(self.model_state)
        for name, layer in self._layers():
            if name in ("Linear", "Conv2d"):
                z = None if noise is None else noise[layer]
                self._replace(self.sample(layer, z), layer.weight, layer.bias)
            else:
                for w, b, key in ((layer.in_proj_weight, layer.in_proj_bias, "attn_in"),
                                  (layer.out_proj.weight, layer.out_proj.bias, "attn_out")):
                    z = None if noise is None else noise[key]
                    self._replace(self.sample(key, z), w, b)


class Diagonal(Curvature):
    """curvatures.py:132-193."""

    def update(self, batch_size: int):
        for name, layer in self._layers():
            if name in ("Linear", "Conv2d"):
                items = [(layer, layer_grad_matrix(layer))]
            else:
                items = []
                for key, w, b in (("attn_in", layer.in_proj_weight, layer.in_proj_bias),
                                  ("attn_out", layer.out_proj.weight, layer.out_proj.bias)):
                    g = w.grad.contiguous().view(w.grad.shape[0], -1)
                    items.append((key, torch.cat([g, b.grad.unsqueeze(dim=1)], dim=1)))
            for key, g in items:
                sq = g ** 2 * batch_size
                if key in self.state:
                    self.state[key] += sq
                else:
                    self.state[key] = sq

    def invert(self, add=0., multiply=1.):
        assert self.state, "State dict is empty. Did you call 'update' prior to this?"
        for index, (layer, value) in enumerate(self.state.items()):
            if isinstance(add, (list, tuple)) and isinstance(multiply, (list, tuple)):
                assert len(add) == len(multiply) == len(self.state)
                n, s = add[index], multiply[index]
            else:
                n, s = add, multiply
            self.inv_state[layer] = inv_sqrt_damped(value, n, s)

    def sample(self, layer, noise: Optional[Tensor] = None) -> Tensor:
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        inv = self.inv_state[layer]
        z = torch.randn_like(inv) if noise is None else noise
        return z * inv


class KFAC(Curvature):
    """curvatures.py:264-392."""

    def __init__(self, model: Module, layer_types=None):
        super().__init__(model, layer_types)
        self.hooks = list()
        self.record = dict()
        for name, layer in self._layers():
            if name in ("Linear", "Conv2d"):
                self.record[layer] = [None, None]
                self.hooks.append(layer.register_forward_pre_hook(self._save_input))
                self.hooks.append(layer.register_backward_hook(self._save_output))
            else:
                raise NotImplementedError

    def _save_input(self, module, input):
        self.record[module][0] = input[0]

    def _save_output(self, module, grad_input, grad_output):
        self.record[module][1] = grad_output[0] * grad_output[0].size(0)

    def update(self, batch_size: int):
        for name, layer in self._layers():
            if name not in ("Linear", "Conv2d"):
                raise NotImplementedError
            x, g = self.record[layer]
            first, second = kfac_factors(x, g, layer)
            if layer in self.state:
                self.state[layer][0] += first
                self.state[layer][1] += second
            else:
                self.state[layer] = [first, second]

    def invert(self, add=0., multiply=1.):
        assert self.state, "State dict is empty. Did you call 'update' prior to this?"
        for index, (layer, (first, second)) in enumerate(self.state.items()):
            n, s = _damping(add, multiply, index, len(self.state), (float, int))
            n, s = float(n), float(s)
            self.inv_state[layer] = (damped_inverse_cholesky(first, n, s),
                                     damped_inverse_cholesky(second, n, s))

    def sample(self, layer, noise: Optional[Tensor] = None) -> Tensor:
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        first, second = self.inv_state[layer]
        z = torch.randn(first.size(0), second.size(0), dtype=first.dtype) if noise is None else noise
        return (first @ z @ second.t()).t()


class EFB(Curvature):
    """curvatures.py:395-460."""

    def __init__(self, model: Module, factors: Dict, layer_types=None, eigvecs: Optional[Dict] = None):
        super().__init__(model, layer_types)
        self.eigvecs = eigenvectors_of_factors(factors) if eigvecs is None else eigvecs
        self.diags = dict()

    def update(self, batch_size: int):
        for name, layer in self._layers():
            if name not in ("Linear", "Conv2d"):
                raise NotImplementedError
            g = layer_grad_matrix(layer)
            qa, qg = self.eigvecs[layer]
            lambdas = (qg.t() @ g @ qa) ** 2
            if layer in self.state:
                self.state[layer] += lambdas
                self.diags[layer] += g ** 2 * batch_size
            else:
                self.state[layer] = lambdas
                self.diags[layer] = g ** 2 * batch_size

    def invert(self, add=0., multiply=1.):
        assert self.state, "State dict is empty. Did you call 'update' prior to this?"
        for index, (layer, value) in enumerate(self.state.items()):
            n, s = _damping(add, multiply, index, len(self.state), (float, int))
            self.inv_state[layer] = inv_sqrt_damped(value, n, s)

    def sample(self, layer, noise: Optional[Tensor] = None) -> Tensor:
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        qa, qg = self.eigvecs[layer]
        z = torch.randn(qa.size(0), qg.size(0), dtype=qa.dtype) if noise is None else noise.clone()
        z = z * self.inv_state[layer].t()
        return (qa @ z @ qg.t()).t()


class INF(Curvature):
    """curvatures.py:463-672 (literal conventions of SURVEY.md section 3.5)."""

    def __init__(self, model, diags, factors, lambdas, layer_types=None, eigvecs: Optional[Dict] = None):
        super().__init__(model, layer_types)
        assert diags.keys() == factors.keys() == lambdas.keys()
        self.eigvecs = eigenvectors_of_factors(factors) if eigvecs is None else eigvecs
        self.lambdas = lambdas
        self.diags = diags

    def update(self, rank: int = 100):
        for layer in list(self.diags.keys()):
            qa, qg = self.eigvecs[layer]
            lambda_vec = self.lambdas[layer].t().contiguous().view(-1)
            diag_vec = self.diags[layer].t().contiguous().view(-1)
            lr_qa, lr_qg, lr_lambda = self._dim_reduction(qa, qg, lambda_vec, rank)
            sif = self._diagonal_accumulator(lr_qa, lr_qg, lr_lambda)
            self.state[layer] = (lr_qa, lr_qg, lr_lambda, diag_vec - sif)

    def invert(self, add=0., multiply=1.):
        assert self.state, "State dict is empty. Did you call 'update' prior to this?"
        for index, (layer, value) in enumerate(self.state.items()):
            n, s = _damping(add, multiply, index, len(self.state), (float, int))
            lr_qa, lr_qg, lr_lambda, correction = value
            correction[correction < 0] = 0          # in place, as curvatures.py:523
            reg_lambda = (s * lr_lambda).sqrt()
            reg_inv_corr = inv_sqrt_damped(correction, n, s)
            pre = self.pre_sampler(lr_qa, lr_qg, reg_lambda, reg_inv_corr)
            self.inv_state[layer] = (lr_qa, lr_qg, reg_inv_corr, pre)

    def sample(self, layer, noise: Optional[Tensor] = None) -> Tensor:
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        a, b, c, d = self.inv_state[layer]
        return self.sampler(a, b, c, d, noise).reshape(a.shape[0], b.shape[0]).t()

    @staticmethod
    def pre_sampler(qa: Tensor, qg: Tensor, reg_lambda: Tensor, reg_inv_corr: Tensor) -> Tensor:
        """curvatures.py:538-572."""
        S = torch.diag(reg_lambda)
        V = reg_inv_corr.contiguous().view(-1, 1) * kron(qa, qg) @ S
        vtv = V.t() @ V
        vtv = (vtv + vtv.t()) / 2.
        eye = torch.eye(S.shape[0], dtype=S.dtype)
        A = torch.linalg.inv(torch.linalg.cholesky(vtv))
        B = torch.linalg.cholesky(vtv + eye)
        C = A.t() @ (B - eye) @ A
        Lc = torch.linalg.inv(torch.linalg.inv(C) + vtv)
        return S @ Lc @ S

    @staticmethod
    def sampler(qa: Tensor, qg: Tensor, reg_inv_corr: Tensor, pre_sample: Tensor,
                noise: Optional[Tensor] = None) -> Tensor:
        """curvatures.py:574-600."""
        X = torch.randn(qa.shape[0] * qg.shape[0], dtype=qa.dtype) if noise is None else noise
        Yl = reg_inv_corr * X
        U = Yl.reshape(qg.shape[0], qa.shape[0])
        Xq = qg.t() @ U @ qa
        Qx = pre_sample @ Xq.t().contiguous().view(-1)
        Wm = Qx.reshape(qg.shape[1], qa.shape[1])
        Xps = qg @ Wm @ qa.t()
        Yr = reg_inv_corr ** 2 * Xps.t().contiguous().view(-1)
        return Yl - Yr

    @staticmethod
    def _dim_reduction(qa: Tensor, qg: Tensor, lambda_vec: Tensor, rank: int):
        """curvatures.py:602-647."""
        if rank >= lambda_vec.shape[0]:
            return qa, qg, lambda_vec
        m = qg.shape[1]
        order = torch.argsort(-torch.abs(lambda_vec))[:rank] + 1      # 1-based, as the reference
        left, right = [], []
        for z in range(rank):
            i = int((order[z] - 1.) / m + 1.)
            j = int(order[z]) - m * (i - 1)
            left.append(i)
            right.append(j)
        left = sorted(set(left))
        right = sorted(set(right))
        picks = [m * (i - 1) + j - 1 for i in left for j in right]
        return (qa[:, [i - 1 for i in left]], qg[:, [j - 1 for j in right]], lambda_vec[picks])

    @staticmethod
    def _diagonal_accumulator(qa: Tensor, qg: Tensor, lambda_vec: Tensor) -> Tensor:
        """curvatures.py:649-672: for every row i of qa, (kron(qa[i], qg)**2) @ lambda."""
        n, m = qa.shape[0], qg.shape[0]
        out = torch.zeros(n * m, dtype=lambda_vec.dtype)
        for i in range(n):
            out[i * m:(i + 1) * m] = (kron(qa[i, :].unsqueeze(0), qg) ** 2) @ lambda_vec
        return out


# ----------------------------------------------------------------------------
# model used by BASELINE configs 1-2 (architecture of curvature/lenet5.py:11-24;
# weights are seeded random-init here -- the bundled .pth is reference data and
# is not copied into this repository)
# ----------------------------------------------------------------------------
class Flatten(torch.nn.Module):
    def forward(self, x):
        return x.view(x.size(0), -1)


def lenet5() -> torch.nn.Sequential:
    return torch.nn.Sequential(
        torch.nn.Conv2d(1, 6, 5, padding=2), torch.nn.ReLU(), torch.nn.MaxPool2d(2, 2),
        torch.nn.Conv2d(6, 16, 5), torch.nn.ReLU(), torch.nn.MaxPool2d(2, 2),
        Flatten(),
        torch.nn.Linear(16 * 5 * 5, 120), torch.nn.ReLU(),
        torch.nn.Linear(120, 84), torch.nn.ReLU(),
        torch.nn.Linear(84, 10))


def fisher_step(model: Module, x: Tensor, labels: Optional[Tensor] = None,
                generator: Optional[torch.Generator] = None, retain_graph: bool = False):
    """One pass of the estimation loop body (scripts/factors.py:51-59,
    scripts/test.py:33-44): forward, sample labels from the model's own
    categorical output (unless given), mean cross-entropy, zero_grad, backward.
    Returns (loss, labels, logits)."""
    logits = model(x)
    if labels is None:
        probs = torch.softmax(logits.detach().float().cpu(), dim=1)
        labels = torch.multinomial(probs, 1, generator=generator).squeeze(1).to(logits.device)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    model.zero_grad()
    loss.backward(retain_graph=retain_graph)
    return loss, labels, logits
