"""ctypes binding of libcurvature_b200.so (the C ABI declared in include/curvature_b200.h).

There is no fallback of any kind: if the shared library is missing this module raises at
import, and every compute call raises ``RuntimeError`` with the library's own message when the
kernel launch fails (e.g. no CUDA device).  torch is used only for device memory and streams.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcurvature_b200.so")

PREC_FP32, PREC_TF32, PREC_BF16X3, PREC_BF16, PREC_TF32_TMA = 0, 1, 2, 3, 4
PRECISION_NAMES = {"fp32": PREC_FP32, "tf32": PREC_TF32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16,
                   "tf32_tma": PREC_TF32_TMA}
# parity tier of each arithmetic tier (relative Frobenius error of a factor against the fp32 reference)
PRECISION_TOLERANCE = {PREC_FP32: 1e-5, PREC_BF16X3: 1e-5, PREC_TF32: 1e-3, PREC_BF16: 1e-3, PREC_TF32_TMA: 1e-3}
DEFAULT_PRECISION = "bf16x3"
OP_SYRK_CONV, OP_SYRK_ROWS, OP_EFB_PROJECT, OP_CHOL_INV, OP_SAMPLE_MN, OP_SYRK_CONV_NHWC, OP_SYRK_ROWS_NHWC = range(7)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m curvature_b200.build` "
        "(nvcc, sm_100a).  curvature_b200 has no CPU or PyTorch fallback.")

_lib = ctypes.CDLL(LIB_PATH)

_f32p = c_void_p  # device pointers travel as integers


def _sig(name, restype, *argtypes):
    fn = getattr(_lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_abi_version = _sig("crv_abi_version", c_int)
_last_error = _sig("crv_last_error", c_char_p)
_sm_count = _sig("crv_device_sm_count", c_int)
_profile_enable = _sig("crv_profile_enable", c_int, c_int)
_debug_timeline = _sig("crv_debug_timeline", c_int, ctypes.c_void_p)
_debug_trace = _sig("crv_debug_trace", c_int, ctypes.c_void_p, c_int)
_debug_trace_count = _sig("crv_debug_trace_count", c_int)
_profile_collect = _sig("crv_profile_collect", c_int, POINTER(ctypes.c_double), POINTER(ctypes.c_double),
                       POINTER(ctypes.c_double), POINTER(ctypes.c_longlong), c_int)
_workspace_bytes = _sig("crv_workspace_bytes", c_size_t, c_int, POINTER(c_int64), c_int)
_syrk_conv = _sig("crv_syrk_conv_accum", c_int, _f32p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                  c_int, c_int, c_int, c_float, _f32p, c_void_p, c_size_t, c_int, c_void_p)
_syrk_rows = _sig("crv_syrk_rows_accum", c_int, _f32p, c_int, c_int, c_int, c_int, c_float, _f32p, c_void_p,
                  c_size_t, c_int, c_void_p)
_syrk_conv_nhwc = _sig("crv_syrk_conv_accum_nhwc", c_int, _f32p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                       c_int, c_int, c_int, c_int, c_float, _f32p, c_void_p, c_size_t, c_int, c_void_p)
_syrk_rows_nhwc = _sig("crv_syrk_rows_accum_nhwc", c_int, _f32p, c_int, c_int, c_int, c_int, c_float, _f32p,
                       c_void_p, c_size_t, c_int, c_void_p)
_stream_join = _sig("crv_stream_join", c_int, c_void_p)
_stream_fork = _sig("crv_stream_fork", c_int, c_void_p)
_diag_accum = _sig("crv_diag_accum", c_int, _f32p, _f32p, c_int, c_int, c_float, _f32p, _f32p, c_void_p)
_efb_project = _sig("crv_efb_project_accum", c_int, _f32p, _f32p, _f32p, c_int, c_int, _f32p, c_void_p,
                    c_size_t, c_int, c_void_p)
_chol_inv = _sig("crv_chol_inv_batched", c_int, POINTER(c_void_p), POINTER(c_int), c_int, POINTER(c_float),
                 POINTER(c_float), POINTER(c_void_p), c_void_p, c_void_p, c_size_t, c_void_p)
_sample_mn = _sig("crv_sample_matrix_normal", c_int, _f32p, _f32p, _f32p, _f32p, c_int, c_int, c_int, _f32p,
                  _f32p, _f32p, _f32p, _f32p, c_void_p, c_size_t, c_int, c_void_p)
_round_tf32 = _sig("crv_round_tf32", c_int, _f32p, _f32p, c_size_t, c_void_p)
_inv_sqrt = _sig("crv_elementwise_inv_sqrt", c_int, _f32p, c_float, c_float, _f32p, c_size_t, c_void_p)
_diag_sample = _sig("crv_diag_sample", c_int, _f32p, _f32p, c_int, c_int, c_int, _f32p, _f32p, _f32p, _f32p,
                    _f32p, c_void_p)
_gemm = _sig("crv_gemm", c_int, _f32p, c_int, c_int, _f32p, c_int, c_int, _f32p, c_int, c_int, c_int, c_int,
             c_float, c_float, c_int, c_void_p)



class SyrkDenseItem(ctypes.Structure):
    """crv_syrk_dense_item (include/curvature_b200.h, K1f)."""
    _fields_ = [("x", c_void_p), ("N", c_int), ("C", c_int), ("H", c_int), ("W", c_int), ("kh", c_int), ("kw", c_int),
                ("sh", c_int), ("sw", c_int), ("ph", c_int), ("pw", c_int), ("has_bias", c_int), ("alpha", c_float),
                ("F", c_void_p)]


class SyrkItem(ctypes.Structure):
    """crv_syrk_item (include/curvature_b200.h)"""
    _fields_ = [("x", c_void_p), ("N", c_int), ("C", c_int), ("H", c_int), ("W", c_int), ("kh", c_int), ("kw", c_int),
                ("sh", c_int), ("sw", c_int), ("ph", c_int), ("pw", c_int), ("alpha", c_float), ("F", c_void_p),
                ("nchw", c_int), ("zero_mean", c_int)]


class EfbItem(ctypes.Structure):
    """crv_efb_item (include/curvature_b200.h)"""
    _fields_ = [("QG", c_void_p), ("QA", c_void_p), ("G", c_void_p), ("M", c_int), ("K", c_int), ("lambdas", c_void_p),
                ("round_g", c_int)]


class SampleItem(ctypes.Structure):
    """crv_sample_item (include/curvature_b200.h)"""
    _fields_ = [("LG", c_void_p), ("LA", c_void_p), ("z", c_void_p), ("row_scale", c_void_p), ("M", c_int), ("K0", c_int),
                ("has_bias", c_int), ("mu_w", c_void_p), ("mu_b", c_void_p), ("w_out", c_void_p), ("b_out", c_void_p),
                ("s_out", c_void_p)]


_efb_batch_ws = _sig("crv_efb_project_batch_workspace", c_size_t, POINTER(EfbItem), c_int)
_efb_batch = _sig("crv_efb_project_batch", c_int, POINTER(EfbItem), c_int, c_void_p, c_size_t, c_int, c_void_p)
_sample_batch_ws = _sig("crv_sample_matrix_normal_batch_workspace", c_size_t, POINTER(SampleItem), c_int)
_sample_batch = _sig("crv_sample_matrix_normal_batch", c_int, POINTER(SampleItem), c_int, c_void_p, c_size_t, c_int, c_void_p)


class SampleMultiItem(ctypes.Structure):
    """crv_sample_multi_item (include/curvature_b200.h)"""
    _fields_ = [("LG", c_void_p), ("LA", c_void_p), ("z", c_void_p), ("M", c_int), ("K", c_int), ("s_out", c_void_p)]


_sample_multi_ws = _sig("crv_sample_matrix_normal_multi_workspace", c_size_t, POINTER(SampleMultiItem), c_int, c_int)
_sample_multi = _sig("crv_sample_matrix_normal_multi", c_int, POINTER(SampleMultiItem), c_int, c_int, c_void_p, c_size_t, c_int,
                     c_void_p)


class DiagItem(ctypes.Structure):
    """crv_diag_item (include/curvature_b200.h)"""
    _fields_ = [("wgrad", c_void_p), ("bgrad", c_void_p), ("M", c_int), ("K0", c_int), ("state", c_void_p),
                ("grads_out", c_void_p)]


_diag_batch = _sig("crv_diag_accum_batch", c_int, POINTER(DiagItem), c_int, c_float, c_void_p)
_debug_partition = _sig("crv_debug_partition", c_int, POINTER(SyrkItem), c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int),
                       POINTER(c_int), POINTER(c_int), POINTER(ctypes.c_uint), c_int, POINTER(c_int), POINTER(c_int),
                       POINTER(c_int), c_int)
_syrk_batch_ws = _sig("crv_syrk_batch_nhwc_workspace", c_size_t, POINTER(SyrkItem), c_int, c_int)
_syrk_batch = _sig("crv_syrk_batch_nhwc", c_int, POINTER(SyrkItem), c_int, c_void_p, c_size_t, c_int, c_void_p)

_syrk_batch_dense = _sig("crv_syrk_batch_dense", c_int, POINTER(SyrkDenseItem), c_int, c_void_p)

ABI_VERSION = _abi_version()
EXPORTED_SYMBOLS = (
    "crv_abi_version", "crv_last_error", "crv_device_sm_count", "crv_workspace_bytes", "crv_profile_enable",
    "crv_profile_collect", "crv_debug_timeline", "crv_debug_trace", "crv_debug_trace_count",
    "crv_syrk_conv_accum", "crv_syrk_rows_accum", "crv_syrk_conv_accum_nhwc", "crv_syrk_rows_accum_nhwc", "crv_syrk_batch_nhwc", "crv_syrk_batch_nhwc_workspace", "crv_syrk_batch_dense", "crv_debug_partition", "crv_stream_join", "crv_stream_fork",
    "crv_diag_accum", "crv_diag_accum_batch", "crv_efb_project_accum", "crv_efb_project_batch",
    "crv_efb_project_batch_workspace", "crv_sample_matrix_normal_batch", "crv_sample_matrix_normal_batch_workspace",
    "crv_sample_matrix_normal_multi", "crv_sample_matrix_normal_multi_workspace",
    "crv_chol_inv_batched", "crv_sample_matrix_normal", "crv_round_tf32", "crv_elementwise_inv_sqrt", "crv_diag_sample",
    "crv_gemm")

# counts kernel-launching C-ABI calls (bench.py reports launches from it)
launch_calls = 0


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"curvature_b200::{what} failed: {_last_error().decode(errors='replace')}")


def _dev(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"curvature_b200: {what} must live on a CUDA device "
                           "(there is no CPU fallback; the CUDA kernels are the only implementation)")
    if t.dtype != torch.float32:
        raise TypeError(f"curvature_b200: {what} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"curvature_b200: {what} must be contiguous")
    return t.data_ptr()


def _opt(t, what):
    return None if t is None else _dev(t, what)


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def default_precision():
    name = os.environ.get("CURVATURE_B200_PRECISION", DEFAULT_PRECISION).lower()
    if name not in PRECISION_NAMES:
        raise ValueError(f"CURVATURE_B200_PRECISION={name!r}; expected one of {sorted(PRECISION_NAMES)}")
    return PRECISION_NAMES[name]


def resolve_precision(p):
    if p is None:
        return default_precision()
    if isinstance(p, str):
        if p.lower() not in PRECISION_NAMES:
            raise ValueError(f"precision={p!r}; expected one of {sorted(PRECISION_NAMES)}")
        return PRECISION_NAMES[p.lower()]
    if int(p) not in PRECISION_NAMES.values():
        raise ValueError(f"precision={p!r}; expected one of {sorted(PRECISION_NAMES.values())}")
    return int(p)


_workspaces = {}


def workspace(nbytes, device):
    """Grow-only scratch buffer per device (kernels never allocate).  Growing it drains the device first: a split
    reduction on the library's side stream may still be reading the old buffer."""
    nbytes = max(int(nbytes), 256)
    buf = _workspaces.get(device)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            torch.cuda.synchronize(device)
        buf = torch.empty(nbytes + nbytes // 4, dtype=torch.uint8, device=device)
        _workspaces[device] = buf
    return buf


def stream_fork(device=None):
    """Tell the library that every tensor the following SYRK calls read is complete on torch's current stream now."""
    if device is not None and torch.device(device).type != "cuda":
        raise RuntimeError("curvature_b200: the model must live on a CUDA device "
                           "(there is no CPU fallback; the CUDA kernels are the only implementation)")
    with torch.cuda.device(device):
        _check(_stream_fork(torch.cuda.current_stream().cuda_stream), "crv_stream_fork")


def stream_join(device=None):
    """Make torch's current stream wait for the split reductions still running on the library's side stream."""
    if device is not None and torch.device(device).type != "cuda":
        return          # (nothing can have been enqueued for a CPU model: every launch refuses CPU tensors)
    with torch.cuda.device(device):
        _check(_stream_join(torch.cuda.current_stream().cuda_stream), "crv_stream_join")


def workspace_bytes(op, dims):
    arr = (c_int64 * len(dims))(*[int(d) for d in dims])
    return _workspace_bytes(op, arr, len(dims))


def sm_count():
    return _sm_count()


KERNEL_CLASSES = ("syrk_nhwc_bf16", "syrk_nhwc_tf32", "syrk_staged_nchw", "syrk_split_reduce", "cast_prepass", "syrk_simt_fp32")


def profile_enable(on=True):
    """Bracket every SYRK-family kernel launch with a CUDA event pair on its stream (bench.py's roofline)."""
    _check(_profile_enable(int(bool(on))), "crv_profile_enable")


def debug_timeline(buf=None):
    """Profiling aid: per-CTA time stamps of the channels-last SYRK kernel into `buf` (int64 CUDA tensor, >= 1280)."""
    _check(_debug_timeline(None if buf is None else buf.data_ptr()), "crv_debug_timeline")


def debug_trace(slots=0, device=None):
    """Profiling aid: start (slots > 0; returns the int64 buffer, shape (slots, 8)) or stop (slots = 0) the launch-level
    trace of the channels-last SYRK path -- device-clock first-start / last-end of every contraction, reduction and pre-pass
    kernel with the stream overlap left on.  debug_trace_count() = launches recorded so far."""
    if not slots:
        _check(_debug_trace(None, 0), "crv_debug_trace")
        return None
    buf = torch.zeros(slots, 8, dtype=torch.int64, device=device)
    buf[:, 0:6:2] = -1
    _check(_debug_trace(buf.data_ptr(), slots), "crv_debug_trace")
    return buf


def debug_trace_count():
    return _debug_trace_count()


def profile_collect():
    """{kernel class: {ms, flops, bytes, launches}} of the launches recorded since the last collect."""
    n = len(KERNEL_CLASSES)
    ms = (ctypes.c_double * n)()
    fl = (ctypes.c_double * n)()
    by = (ctypes.c_double * n)()
    la = (ctypes.c_longlong * n)()
    _check(_profile_collect(ms, fl, by, la, n), "crv_profile_collect")
    return {k: {"ms": ms[i], "flops": fl[i], "bytes": by[i], "launches": int(la[i])} for i, k in enumerate(KERNEL_CLASSES)}


def _dense(t, what):
    """Device pointer of a dense fp32 CUDA tensor in ANY dense layout (the caller states the layout)."""
    if not t.is_cuda:
        raise RuntimeError(f"curvature_b200: {what} must live on a CUDA device "
                           "(there is no CPU fallback; the CUDA kernels are the only implementation)")
    if t.dtype != torch.float32:
        raise TypeError(f"curvature_b200: {what} must be float32, got {t.dtype}")
    return t.data_ptr()


def _is_channels_last(t):
    """True if the 4-D tensor is dense in [N][H][W][C] order and NOT also dense in [N][C][H][W] order."""
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous()


def _tensor_core(precision):
    return precision in (PREC_TF32, PREC_TF32_TMA, PREC_BF16, PREC_BF16X3)


def _copy_tier(precision):
    """Tiers whose pre-pass writes a bf16 copy of the operand anyway and can therefore also take NCHW-dense sources
    (the pre-pass then transposes to channels-last)."""
    return precision in (PREC_BF16, PREC_BF16X3)


def syrk_conv_accum(x, kernel_size, stride, padding, has_bias, alpha, out, precision=PREC_FP32, join=True):
    """out (K,K) += alpha * unfold(x) unfold(x)^T  with the optional ones row (K1a / K1c).

    x is the logical (N,C,H,W) activation.  A channels-last tensor goes to the TMA-fed MN-major kernel
    (crv_syrk_conv_accum_nhwc) when the tier is a tensor-core one and the geometry qualifies; every other case
    uses the NCHW kernel (crv_syrk_conv_accum), on a contiguous copy if the tensor is not NCHW-dense."""
    global launch_calls
    N, C, H, W = x.shape
    kh, kw = kernel_size
    sh, sw = stride
    ph, pw = padding
    K = C * kh * kw + int(bool(has_bias))
    if tuple(out.shape) != (K, K):
        raise ValueError(f"factor has shape {tuple(out.shape)}, expected {(K, K)}")
    dims = [N, C, H, W, kh, kw, sh, sw, ph, pw, int(bool(has_bias)), precision]
    item = nhwc_item(x, kernel_size, stride, padding, has_bias, alpha, out, precision)
    if item is not None:
        syrk_batch_nhwc([item], precision, x.device, join=join)
        return
    if not x.is_contiguous():
        x = x.contiguous()
    # operands the channels-last kernel cannot take: thread-staged TF32 kernel for the 1e-3 tiers, fp32 CUDA cores for the
    # 1e-5 tiers (the library maps bf16 -> tf32 and bf16x3 -> fp32 itself)
    nb = workspace_bytes(OP_SYRK_CONV, dims)
    ws = workspace(nb, x.device)
    launch_calls += 1 if precision in (PREC_FP32, PREC_BF16X3) else 2
    _check(_syrk_conv(_dev(x, "activation"), N, C, H, W, kh, kw, sh, sw, ph, pw, int(bool(has_bias)),
                      float(alpha), _dev(out, "factor"), ws.data_ptr(), ws.numel(), precision, _stream(x)),
           "crv_syrk_conv_accum")


def dense_item(x, kernel_size, stride, padding, has_bias, alpha, out):
    """crv_syrk_dense_item of a DENSE operand (K1f), or None: x is the (N,C,H,W) activation of a convolution
    (`kernel_size` given) or an (N, M, ...) rows operand (`kernel_size` None: output gradient / Linear input)."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        return None
    if kernel_size is None:
        N, M = x.shape[0], x.shape[1]
        L = 1
        for d in x.shape[2:]:
            L *= d
        dims = (N, M, L, 1, 1, 1, 1, 1, 0, 0)
    else:
        if x.dim() != 4:
            return None
        dims = (*x.shape, *kernel_size, *stride, *padding)
    D = dims[1] * dims[4] * dims[5] + int(bool(has_bias))
    if tuple(out.shape) != (D, D) or not out.is_contiguous():
        return None
    return SyrkDenseItem(x.data_ptr(), *[int(d) for d in dims], int(bool(has_bias)), float(alpha), out.data_ptr())


def syrk_dense_array(items):
    return (SyrkDenseItem * len(items))(*items)


def syrk_batch_dense(arr, n, device):
    """One launch for every factor of a small model (K1f, exact fp32 products on the CUDA cores)."""
    global launch_calls
    launch_calls += 1
    with torch.cuda.device(device):
        _check(_syrk_batch_dense(arr, n, torch.cuda.current_stream(device).cuda_stream), "crv_syrk_batch_dense")


def syrk_rows_accum(g, has_bias, alpha, out, precision=PREC_FP32, join=True):
    """out (D,D) += alpha * sum_{n,l} g[n,:,l] g[n,:,l]^T for g viewed as (N, M, L) (K1b / K1d).

    Channels-last 4-D operands and 2-D (Linear) operands are [R][M] matrices in memory: they go to the TMA-fed
    MN-major kernel when the tier is a tensor-core one and the geometry qualifies."""
    global launch_calls
    N, M = g.shape[0], g.shape[1]
    L = 1
    for s in g.shape[2:]:
        L *= s
    D = M + int(bool(has_bias))
    if tuple(out.shape) != (D, D):
        raise ValueError(f"factor has shape {tuple(out.shape)}, expected {(D, D)}")
    item = nhwc_item(g, None, None, None, has_bias, alpha, out, precision)
    if item is not None:
        syrk_batch_nhwc([item], precision, g.device, join=join)
        return
    if not g.is_contiguous():
        g = g.contiguous()
    nb = workspace_bytes(OP_SYRK_ROWS, [N, M, L, int(bool(has_bias)), precision])
    ws = workspace(nb, g.device)
    launch_calls += 1 if precision in (PREC_FP32, PREC_BF16X3) else 2
    _check(_syrk_rows(_dev(g, "operand"), N, M, L, int(bool(has_bias)), float(alpha), _dev(out, "factor"),
                      ws.data_ptr(), ws.numel(), precision, _stream(g)), "crv_syrk_rows_accum")


def nhwc_item(t, kernel_size, stride, padding, has_bias, alpha, out, precision, zero_mean=None):
    """The crv_syrk_item of `out += alpha * X X^T` if the channels-last kernel can take the operand, else None.
    `t` is a conv input (N,C,H,W) with the layer's geometry, or -- kernel_size None -- a rows operand (N,M,...).
    `zero_mean` (default: True for rows operands, False for convolution inputs): see crv_syrk_item::zero_mean."""
    if not _tensor_core(precision) or has_bias or t.data_ptr() % 16:
        return None
    if kernel_size is None:
        N, M = t.shape[0], t.shape[1]
        L = 1
        for s in t.shape[2:]:
            L *= s
        rows_major = _is_channels_last(t) or (t.dim() == 2 and t.is_contiguous()) or \
            (t.dim() >= 3 and L == 1 and t.is_contiguous())
        if tuple(out.shape) != (M, M):
            return None
        if rows_major:
            nchw = 0
        elif t.is_contiguous() and _copy_tier(precision):     # (N, M, L) dense: the pre-pass transposes it
            nchw = 1
        else:
            return None
        item = SyrkItem(_dense(t, "operand"), N, M, 1, L, 1, 1, 1, 1, 0, 0, float(alpha), _dev(out, "factor"), nchw,
                        1 if zero_mean is None else int(bool(zero_mean)))
        if not _syrk_batch_ws(ctypes.byref(item), 1, precision):
            return None
        return item
    N, C, H, W = t.shape
    kh, kw = kernel_size
    sh, sw = stride
    ph, pw = padding
    K = C * kh * kw
    if tuple(out.shape) != (K, K):
        return None
    if _is_channels_last(t):
        nchw = 0
    elif t.is_contiguous() and (C <= 4 or _copy_tier(precision)):
        nchw = 1          # NCHW-dense: the packed small-C path and the transposing pre-pass of the copy tiers take it
    else:
        return None
    item = SyrkItem(_dense(t, "activation"), N, C, H, W, kh, kw, sh, sw, ph, pw, float(alpha), _dev(out, "factor"), nchw,
                    int(bool(zero_mean)))
    if not _syrk_batch_ws(ctypes.byref(item), 1, precision):
        return None
    return item


def debug_partition(geoms, precision, sms=148, which=0):
    """Host-only: how the scheduler cuts a batch of (N, C, H, W, kh, kw, sh, sw, ph, pw) items into launches and CTA ranges.
    Returns dict(launch_of_item, n_launches, boundaries=[(pair, box)...] of launch `which`, nbox, nb)."""
    items = [SyrkItem(4096, *[int(v) for v in g], 1.0, 4096, 0, 0) for g in geoms]
    arr = (SyrkItem * len(items))(*items)
    loi = (c_int * len(items))()
    nl, G, pairs = c_int(0), c_int(0), c_int(0)
    cap, pcap = 256, 1 << 16
    q, b = (c_int * cap)(), (ctypes.c_uint * cap)()
    nbox, nb = (c_int * pcap)(), (c_int * pcap)()
    _check(_debug_partition(arr, len(items), precision, sms, which, loi, ctypes.byref(nl), ctypes.byref(G), q, b, cap,
                            ctypes.byref(pairs), nbox, nb, pcap), "crv_debug_partition")
    return {"launch_of_item": list(loi), "n_launches": nl.value, "boundaries": [(q[c], b[c]) for c in range(G.value + 1)],
            "nbox": list(nbox[:pairs.value]), "nb": list(nb[:pairs.value])}


def syrk_batch_nhwc(items, precision, device, join=True):
    """One C-ABI call for a list of crv_syrk_items (K1e): read-once factors share stream-K launches."""
    global launch_calls
    if not items:
        return
    arr = (SyrkItem * len(items))(*items)
    nb = _syrk_batch_ws(arr, len(items), precision)
    if not nb:
        raise RuntimeError("crv_syrk_batch_nhwc_workspace: " + (_last_error() or b"unsupported item").decode())
    ws = workspace(nb, device)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        launch_calls += 2 * len(items)
        _check(_syrk_batch(arr, len(items), ws.data_ptr(), ws.numel(), precision, stream), "crv_syrk_batch_nhwc")
        if join:
            _check(_stream_join(stream), "crv_stream_join")


def syrk_item_array(items, precision):
    """(ctypes array of the items, workspace bytes of the batch call) -- built once per geometry by KFAC.update."""
    if not items:
        return None, 0
    arr = (SyrkItem * len(items))(*items)
    nb = _syrk_batch_ws(arr, len(items), precision)
    if not nb:
        raise RuntimeError("crv_syrk_batch_nhwc_workspace: " + (_last_error() or b"unsupported item").decode())
    return arr, nb


def syrk_batch_arr(arr, n, ws_bytes, precision, device, join=True):
    """crv_syrk_batch_nhwc on a prebuilt item array (see syrk_item_array)."""
    global launch_calls
    ws = workspace(ws_bytes, device)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        launch_calls += 2 * n
        _check(_syrk_batch(arr, n, ws.data_ptr(), ws.numel(), precision, stream), "crv_syrk_batch_nhwc")
        if join:
            _check(_stream_join(stream), "crv_stream_join")


def diag_accum(wgrad, bgrad, scale, state=None, grads_out=None):
    """state (M,K) += scale * [wgrad | bgrad]^2; optionally also emit the concatenated grads (K2)."""
    global launch_calls
    M = wgrad.shape[0]
    K0 = wgrad.numel() // M
    launch_calls += 1
    _check(_diag_accum(_dev(wgrad, "weight.grad"), _opt(bgrad, "bias.grad"), M, K0, float(scale),
                       _opt(state, "state"), _opt(grads_out, "grads_out"), _stream(wgrad)), "crv_diag_accum")


def prepare_diag_batch(entries):
    """ctypes item array of crv_diag_accum_batch for (wgrad, bgrad or None, state or None, grads_out or None) entries;
    callers that repeat the call every step keep the array and only refresh the gradient pointers."""
    items = []
    for wgrad, bgrad, state, grads_out in entries:
        M = wgrad.shape[0]
        items.append(DiagItem(_dev(wgrad, "weight.grad"), _opt(bgrad, "bias.grad"), M, wgrad.numel() // M,
                              _opt(state, "state"), _opt(grads_out, "grads_out")))
    return (DiagItem * len(items))(*items)


def run_diag_batch(arr, n, scale, like):
    global launch_calls
    launch_calls += 1
    _check(_diag_batch(arr, n, float(scale), _stream(like)), "crv_diag_accum_batch")


def diag_accum_batch(entries, scale):
    """One launch for a list of (wgrad, bgrad or None, state or None, grads_out or None) (K2b)."""
    if not entries:
        return
    run_diag_batch(prepare_diag_batch(entries), len(entries), scale, entries[0][0])


def efb_project_accum(QG, QA, G, lambdas, precision=PREC_FP32):
    """lambdas (M,K) += (QG^T G QA)^2 (K3)."""
    global launch_calls
    M, K = G.shape
    ws = workspace(workspace_bytes(OP_EFB_PROJECT, [M, K]), G.device)
    launch_calls += 2
    _check(_efb_project(_dev(QG, "QG"), _dev(QA, "QA"), _dev(G, "grads"), M, K, _dev(lambdas, "lambdas"),
                        ws.data_ptr(), ws.numel(), precision, _stream(G)), "crv_efb_project_accum")


def prepare_efb_batch(entries, round_g):
    """(item array, workspace bytes) of crv_efb_project_batch for (QG, QA, G, lambdas) entries."""
    items = [EfbItem(_dev(QG, "QG"), _dev(QA, "QA"), _dev(G, "grads"), G.shape[0], G.shape[1], _dev(lam, "lambdas"),
                     int(bool(round_g))) for QG, QA, G, lam in entries]
    arr = (EfbItem * len(items))(*items)
    return arr, _efb_batch_ws(arr, len(items))


def run_efb_batch(arr, n, ws_bytes, precision, like):
    global launch_calls
    ws = workspace(ws_bytes, like.device)
    launch_calls += 1
    _check(_efb_batch(arr, n, ws.data_ptr(), ws.numel(), precision, _stream(like)), "crv_efb_project_batch")


def efb_project_batch(entries, precision, round_g):
    """One call for a list of (QG, QA, G, lambdas) (K3b): lambdas += (QG^T G QA)^2 per entry."""
    if not entries:
        return
    arr, nb = prepare_efb_batch(entries, round_g)
    run_efb_batch(arr, len(entries), nb, precision, entries[0][2])


def prepare_sample_batch(entries):
    """(item array, workspace bytes) of crv_sample_matrix_normal_batch for dicts with the arguments of K5."""
    items = []
    for e in entries:
        LG, LA, z = e["LG"], e["LA"], e["z"]
        M, K = LG.shape[0], LA.shape[0]
        has_bias = int(bool(e["has_bias"]))
        if tuple(z.shape) != (K, M):
            raise ValueError(f"noise has shape {tuple(z.shape)}, expected {(K, M)}")
        items.append(SampleItem(_dev(LG, "LG"), _dev(LA, "LA"), _dev(z, "noise"), _opt(e.get("row_scale"), "row_scale"), M,
                                K - has_bias, has_bias, _opt(e.get("mu_w"), "mu_w"), _opt(e.get("mu_b"), "mu_b"),
                                _opt(e.get("w_out"), "weight"), _opt(e.get("b_out"), "bias"), _opt(e.get("s_out"), "sample")))
    arr = (SampleItem * len(items))(*items)
    return arr, _sample_batch_ws(arr, len(items))


def run_sample_batch(arr, n, ws_bytes, precision, like):
    global launch_calls
    ws = workspace(ws_bytes, like.device)
    launch_calls += 1
    _check(_sample_batch(arr, n, ws.data_ptr(), ws.numel(), precision, _stream(like)), "crv_sample_matrix_normal_batch")


def sample_matrix_normal_batch(entries, precision):
    """One call for a list of dicts with the arguments of sample_matrix_normal (K5b)."""
    if not entries:
        return
    arr, nb = prepare_sample_batch(entries)
    run_sample_batch(arr, len(entries), nb, precision, entries[0]["LG"])


def sample_matrix_normal_multi(entries, S, precision):
    """S stacked draws per layer in one call (K5c): entries = [(LG (M,M), LA (K,K), z (S*K, M), out (M, S, K))]."""
    global launch_calls
    if not entries or S <= 0:
        return
    items = []
    for LG, LA, z, out in entries:
        M, K = LG.shape[0], LA.shape[0]
        if tuple(z.shape) != (S * K, M) or tuple(out.shape) != (M, S, K):
            raise ValueError(f"noise {tuple(z.shape)} / output {tuple(out.shape)}, expected {(S * K, M)} / {(M, S, K)}")
        items.append(SampleMultiItem(_dev(LG, "LG"), _dev(LA, "LA"), _dev(z, "noise"), M, K, _dev(out, "samples")))
    arr = (SampleMultiItem * len(items))(*items)
    dev = entries[0][0].device
    ws = workspace(_sample_multi_ws(arr, len(items), S), dev)
    launch_calls += 1
    _check(_sample_multi(arr, len(items), S, ws.data_ptr(), ws.numel(), precision, _stream(entries[0][0])),
           "crv_sample_matrix_normal_multi")


def chol_inv_workspace_bytes(dims):
    return workspace_bytes(OP_CHOL_INV, [len(dims)] + [int(d) for d in dims])


def chol_inv_batched(factors, adds, muls, outs, ws=None):
    """outs[i] = chol_lower(inv(sym(sqrt(mul_i) F_i + sqrt(add_i) I)))  (K4).  Returns the device
    int32 info tensor (0 = ok).  `ws`: a caller-owned uint8 workspace of chol_inv_workspace_bytes(dims) bytes (for a call
    that runs on another stream beside one that uses the binding's shared workspace); default: the shared one."""
    global launch_calls
    count = len(factors)
    dev = factors[0].device
    dims = [int(f.shape[0]) for f in factors]
    for f, o in zip(factors, outs):
        if f.shape[0] != f.shape[1] or tuple(o.shape) != tuple(f.shape):
            raise ValueError("factors must be square and outputs must match them")
    Fp = (c_void_p * count)(*[_dev(f, "factor") for f in factors])
    Lp = (c_void_p * count)(*[_dev(o, "output") for o in outs])
    dm = (c_int * count)(*dims)
    ad = (c_float * count)(*[float(a) for a in adds])
    mu = (c_float * count)(*[float(m) for m in muls])
    info = torch.empty(count, dtype=torch.int32, device=dev)
    nb = workspace_bytes(OP_CHOL_INV, [count] + dims)
    if ws is None:
        ws = workspace(nb, dev)
    elif ws.numel() * ws.element_size() < nb or ws.device != dev:
        raise ValueError("chol_inv_batched: workspace too small or on another device")
    launch_calls += 4 + 3 * ((max(dims) + 31) // 32)
    _check(_chol_inv(Fp, dm, count, ad, mu, Lp, info.data_ptr(), ws.data_ptr(), ws.numel() * ws.element_size(),
                     torch.cuda.current_stream(dev).cuda_stream), "crv_chol_inv_batched")
    return info


def sample_matrix_normal(LG, LA, z, has_bias, row_scale=None, mu_w=None, mu_b=None, w_out=None, b_out=None,
                         s_out=None, precision=PREC_FP32):
    """S = LG z^T LA^T, optionally written as mean + S into weight / bias (K5)."""
    global launch_calls
    M = LG.shape[0]
    K = LA.shape[0]
    K0 = K - int(bool(has_bias))
    if tuple(z.shape) != (K, M):
        raise ValueError(f"noise has shape {tuple(z.shape)}, expected {(K, M)}")
    ws = workspace(workspace_bytes(OP_SAMPLE_MN, [M, K]), LG.device)
    launch_calls += 2 + int(row_scale is not None)
    _check(_sample_mn(_dev(LG, "LG"), _dev(LA, "LA"), _dev(z, "noise"), _opt(row_scale, "row_scale"), M, K0,
                      int(bool(has_bias)), _opt(mu_w, "mu_w"), _opt(mu_b, "mu_b"), _opt(w_out, "weight"),
                      _opt(b_out, "bias"), _opt(s_out, "sample"), ws.data_ptr(), ws.numel(), precision,
                      _stream(LG)), "crv_sample_matrix_normal")


def round_tf32(t, out=None):
    """Round a dense fp32 tensor to the nearest TF32 values (into `out`, default a new tensor; `out=t` rounds in place)."""
    global launch_calls
    if out is None:
        out = torch.empty_like(t)
    launch_calls += 1
    _check(_round_tf32(_dev(t, "tensor"), _dev(out, "out"), t.numel(), _stream(t)), "crv_round_tf32")
    return out


def elementwise_inv_sqrt(v, add, mul, out):
    global launch_calls
    launch_calls += 1
    _check(_inv_sqrt(_dev(v, "value"), float(add), float(mul), _dev(out, "out"), v.numel(), _stream(v)),
           "crv_elementwise_inv_sqrt")


def diag_sample(z, inv, has_bias, mu_w=None, mu_b=None, w_out=None, b_out=None, s_out=None):
    global launch_calls
    M, K = inv.shape
    K0 = K - int(bool(has_bias))
    launch_calls += 1
    _check(_diag_sample(_dev(z, "noise"), _dev(inv, "inv_state"), M, K0, int(bool(has_bias)),
                        _opt(mu_w, "mu_w"), _opt(mu_b, "mu_b"), _opt(w_out, "weight"), _opt(b_out, "bias"),
                        _opt(s_out, "sample"), _stream(z)), "crv_diag_sample")


def gemm(A, B, transA=False, transB=False, alpha=1.0, beta=0.0, out=None, precision=PREC_FP32):
    """Row-major C = alpha op(A) op(B) + beta C through the library's own GEMM kernel."""
    global launch_calls
    A = A if A.is_contiguous() else A.contiguous()
    B = B if B.is_contiguous() else B.contiguous()
    m = A.shape[1] if transA else A.shape[0]
    k = A.shape[0] if transA else A.shape[1]
    kb = B.shape[1] if transB else B.shape[0]
    n = B.shape[0] if transB else B.shape[1]
    if k != kb:
        raise ValueError(f"inner dimensions differ: {k} vs {kb}")
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=A.device)
    if m == 0 or n == 0:
        return out
    if k == 0:
        return out.zero_() if beta == 0.0 else out.mul_(beta)
    launch_calls += 1
    _check(_gemm(_dev(A, "A"), A.shape[1], int(transA), _dev(B, "B"), B.shape[1], int(transB), _dev(out, "C"),
                 out.shape[1], m, n, k, float(alpha), float(beta), precision, _stream(A)), "crv_gemm")
    return out
