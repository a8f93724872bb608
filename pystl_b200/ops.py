"""``torch.library`` custom ops around the C ABI.

Three entry points of ``include/pystl_b200.h`` become dispatcher ops

    pystl_b200::robustness(signal[B, T, D], handle)                    -> out[B, T, C]
    pystl_b200::robustness_grad(signal, grad_out[B, T, C], handle)     -> grad_signal[B, T, D]
    pystl_b200::value_and_grad(signal, scale, handle)                  -> (out, acc[2] float64, grad_signal)

(plus ``open_window / open_window_grad`` for the suffix-scan kernels of csrc/scan.cu: open-ended windows on
traces longer than one tile) with fake (meta) kernels and batching rules; ``evaluate`` -- an ``autograd.Function`` over them with
``setup_context`` and a ``vmap`` rule -- is the differentiable entry point the semantics objects call,
so that ``torch.compile`` traces through a robustness evaluation without a graph break and
``torch.func.vmap / grad`` compose with it (the precedent is ``jax.vmap`` over the reference's JAX
backend, tests/test_jax_semantics.py:310-368 of the reference).  A compiled formula
(``compiler.CompiledFormula``: ctypes arrays, not a tensor) is passed as an integer handle.

PyTorch is plumbing here: the ops hand raw device pointers and the current stream to
``libpystl_b200.so``; there is no CPU implementation to fall back to.
"""

from __future__ import annotations

import itertools
from typing import Dict, Optional, Tuple

import torch

from . import capi

_HANDLES: Dict[int, Tuple[object, int]] = {}  # handle -> (CompiledFormula, out_channels)
_next = itertools.count(1)


def handle_of(cf, out_channels: int) -> int:
    h = getattr(cf, "_op_handle", None)
    if h is None:
        h = next(_next)
        cf._op_handle = h
        if len(_HANDLES) > 4096:  # formulas compiled long ago: their semantics object still caches them
            for k in list(_HANDLES)[:2048]:
                _HANDLES.pop(k, None)
    _HANDLES[h] = (cf, int(out_channels))
    return h


def _lookup(handle: int):
    try:
        return _HANDLES[int(handle)]
    except KeyError:
        raise RuntimeError(f"pystl_b200: unknown formula handle {handle} (compile the formula in this process first)") from None


def _check(signal: torch.Tensor) -> None:
    if not signal.is_cuda:
        raise RuntimeError("pystl_b200 ops need a CUDA tensor; there is no CPU fallback")
    if signal.dim() != 3 or signal.dtype != torch.float32:
        raise ValueError("signal must be float32 with shape (batch, time, state_dim)")


_lib = torch.library.Library("pystl_b200", "DEF")
_lib.define("robustness(Tensor signal, int handle) -> Tensor")
_lib.define("robustness_grad(Tensor signal, Tensor grad_out, int handle) -> Tensor")
_lib.define("value_and_grad(Tensor signal, float scale, int handle) -> (Tensor, Tensor, Tensor)")
# (the ops are defined with the low-level Library API on purpose: torch.library.custom_op installs a
# generated autograd.Function on every op, which torch.func.grad refuses to run through in torch 2.11)


def _robustness_cuda(signal: torch.Tensor, handle: int) -> torch.Tensor:
    _check(signal)
    cf, C = _lookup(handle)
    sig = signal.contiguous()
    B, T, D = sig.shape
    out = torch.empty((B, T, C), dtype=torch.float32, device=sig.device)
    with torch.cuda.device(sig.device):
        rc = capi.lib().pystl_b200_eval_forward(cf.ptr(), sig.data_ptr(), B, T, D, out.data_ptr(), None,
                                                torch.cuda.current_stream(sig.device).cuda_stream)
    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
    return out


def _robustness_grad_cuda(signal: torch.Tensor, grad_out: torch.Tensor, handle: int) -> torch.Tensor:
    _check(signal)
    cf, C = _lookup(handle)
    sig = signal.contiguous()
    B, T, D = sig.shape
    g = grad_out.to(dtype=torch.float32).contiguous()
    if tuple(g.shape) != (B, T, C):
        raise ValueError(f"grad_out must have shape {(B, T, C)}, got {tuple(g.shape)}")
    grad = torch.empty_like(sig)
    with torch.cuda.device(sig.device):
        rc = capi.lib().pystl_b200_eval_backward(cf.ptr(), sig.data_ptr(), B, T, D, g.data_ptr(), 1.0, grad.data_ptr(),
                                                 torch.cuda.current_stream(sig.device).cuda_stream)
    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
    return grad


def _value_and_grad_cuda(signal: torch.Tensor, scale: float, handle: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Robustness trace, its (sum, 0) accumulator and the gradient of ``scale * sum(valid outputs)``
    in ONE pass over the signal (pystl_b200_eval_value_and_grad)."""
    _check(signal)
    cf, C = _lookup(handle)
    sig = signal.contiguous()
    B, T, D = sig.shape
    out = torch.empty((B, T, C), dtype=torch.float32, device=sig.device)
    acc = torch.zeros(2, dtype=torch.float64, device=sig.device)
    grad = torch.empty_like(sig)
    with torch.cuda.device(sig.device):
        rc = capi.lib().pystl_b200_eval_value_and_grad(cf.ptr(), sig.data_ptr(), B, T, D, None, float(scale), out.data_ptr(),
                                                       acc.data_ptr(), grad.data_ptr(),
                                                       torch.cuda.current_stream(sig.device).cuda_stream)
    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
    return out, acc, grad


_lib.impl("robustness", _robustness_cuda, "CUDA")
_lib.impl("robustness_grad", _robustness_grad_cuda, "CUDA")
_lib.impl("value_and_grad", _value_and_grad_cuda, "CUDA")
robustness = torch.ops.pystl_b200.robustness
robustness_grad = torch.ops.pystl_b200.robustness_grad
value_and_grad = torch.ops.pystl_b200.value_and_grad


@torch.library.register_fake("pystl_b200::robustness")
def _(signal, handle):
    _, C = _lookup(handle)
    return signal.new_empty((signal.shape[0], signal.shape[1], C))


@torch.library.register_fake("pystl_b200::robustness_grad")
def _(signal, grad_out, handle):
    return torch.empty_like(signal)


@torch.library.register_fake("pystl_b200::value_and_grad")
def _(signal, scale, handle):
    _, C = _lookup(handle)
    return (signal.new_empty((signal.shape[0], signal.shape[1], C)), signal.new_empty((2,), dtype=torch.float64),
            torch.empty_like(signal))


# ---- batching rules: traces are independent, so a vmapped dimension folds into the batch
def _fold(x: torch.Tensor, dim):
    x = x.movedim(dim, 0)
    return x.reshape((x.shape[0] * x.shape[1],) + tuple(x.shape[2:])), x.shape[0]


@torch.library.register_vmap("pystl_b200::robustness")
def _(info, in_dims, signal, handle):
    flat, n = _fold(signal, in_dims[0])
    out = robustness(flat, handle)
    return out.reshape((n, -1) + tuple(out.shape[1:])), 0


@torch.library.register_vmap("pystl_b200::robustness_grad")
def _(info, in_dims, signal, grad_out, handle):
    n = info.batch_size
    sig = signal.movedim(in_dims[0], 0) if in_dims[0] is not None else signal.unsqueeze(0).expand((n,) + tuple(signal.shape))
    g = grad_out.movedim(in_dims[1], 0) if in_dims[1] is not None else grad_out.unsqueeze(0).expand((n,) + tuple(grad_out.shape))
    fs = sig.reshape((-1,) + tuple(sig.shape[2:]))
    fg = g.reshape((-1,) + tuple(g.shape[2:]))
    out = robustness_grad(fs, fg, handle)
    return out.reshape((n, -1) + tuple(out.shape[1:])), 0


# ---- the differentiable entry point: eager autograd, torch.func.grad / vmap(grad(...)), torch.compile
class Robustness(torch.autograd.Function):
    """d robustness / d signal is the backward entry point of the C ABI (pystl_b200_eval_backward)."""

    generate_vmap_rule = False

    @staticmethod
    def forward(signal, handle):
        return robustness(signal, handle)

    @staticmethod
    def setup_context(ctx, inputs, output):
        signal, handle = inputs
        ctx.save_for_backward(signal)
        ctx.handle = handle

    @staticmethod
    def backward(ctx, grad_out):
        (signal,) = ctx.saved_tensors
        return robustness_grad(signal, grad_out, ctx.handle), None

    @staticmethod
    def vmap(info, in_dims, signal, handle):
        flat, n = _fold(signal, in_dims[0])
        out = Robustness.apply(flat, handle)
        return out.reshape((n, -1) + tuple(out.shape[1:])), 0


def evaluate(signal: torch.Tensor, handle: int) -> torch.Tensor:
    """Differentiable robustness evaluation of ``signal[B, T, D]`` -> ``out[B, T, C]``."""
    return Robustness.apply(signal, handle)


# ---- open-ended windows beyond one tile: suffix scan of a value trace (csrc/scan.cu) -------------------------------
_lib.define("open_window_fwd(Tensor y, int semantics, int mode, int start, float temperature, float eps, int p) -> (Tensor, Tensor)")
_lib.define("open_window_grad(Tensor y, Tensor out, Tensor arg, Tensor grad_out, int semantics, int mode, int start, "
            "float temperature, float eps, int p) -> Tensor")


def _check_trace(y: torch.Tensor) -> None:
    if not y.is_cuda:
        raise RuntimeError("pystl_b200 ops need a CUDA tensor; there is no CPU fallback")
    if y.dim() != 2 or y.dtype != torch.float32:
        raise ValueError("the value trace must be float32 with shape (batch, time)")


def _needs_arg(semantics: int, mode: int) -> bool:
    """Minima / maxima record where they came from: classical, and Always on a channel of cumulative pairs."""
    return semantics == capi.SEM_CLASSIC or (semantics == capi.SEM_CUMULATIVE and (mode & 1) == 0)


def _open_window_fwd_cuda(y, semantics, mode, start, temperature, eps, p):
    _check_trace(y)
    y = y.contiguous()
    B, T = y.shape
    out = torch.empty_like(y)
    arg = torch.empty((B, T) if _needs_arg(semantics, mode) else (0,), dtype=torch.int32, device=y.device)
    with torch.cuda.device(y.device):
        rc = capi.lib().pystl_b200_open_window_forward(int(semantics), int(mode), int(start), float(temperature), float(eps), int(p),
                                                       y.data_ptr(), B, T, out.data_ptr(), arg.data_ptr() if arg.numel() else None,
                                                       torch.cuda.current_stream(y.device).cuda_stream)
    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
    return out, arg


def _open_window_grad_cuda(y, out, arg, grad_out, semantics, mode, start, temperature, eps, p):
    _check_trace(y)
    y, out = y.contiguous(), out.contiguous()
    g = grad_out.to(dtype=torch.float32).contiguous()
    B, T = y.shape
    if tuple(g.shape) != (B, T) or tuple(out.shape) != (B, T):
        raise ValueError(f"out / grad_out must have shape {(B, T)}")
    grad = torch.empty_like(y)
    with torch.cuda.device(y.device):
        rc = capi.lib().pystl_b200_open_window_backward(int(semantics), int(mode), int(start), float(temperature), float(eps), int(p),
                                                        y.data_ptr(), out.data_ptr(), arg.data_ptr() if arg.numel() else None,
                                                        g.data_ptr(), B, T, grad.data_ptr(),
                                                        torch.cuda.current_stream(y.device).cuda_stream)
    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
    return grad


_lib.impl("open_window_fwd", _open_window_fwd_cuda, "CUDA")
_lib.impl("open_window_grad", _open_window_grad_cuda, "CUDA")
open_window_fwd = torch.ops.pystl_b200.open_window_fwd
open_window_grad = torch.ops.pystl_b200.open_window_grad


@torch.library.register_fake("pystl_b200::open_window_fwd")
def _(y, semantics, mode, start, temperature, eps, p):
    return torch.empty_like(y), y.new_empty(tuple(y.shape) if _needs_arg(semantics, mode) else (0,), dtype=torch.int32)


@torch.library.register_fake("pystl_b200::open_window_grad")
def _(y, out, arg, grad_out, semantics, mode, start, temperature, eps, p):
    return torch.empty_like(y)


class OpenWindow(torch.autograd.Function):
    """``out[b, t]`` = Always / Eventually of ``y[b, t + start .. T - 1]``; the adjoint is the mirrored prefix scan."""

    @staticmethod
    def forward(y, semantics, mode, start, temperature, eps, p):
        out, arg = open_window_fwd(y, semantics, mode, start, temperature, eps, p)
        return out, arg

    @staticmethod
    def setup_context(ctx, inputs, output):
        y, *ctx.params = inputs
        out, arg = output
        ctx.mark_non_differentiable(arg)
        ctx.save_for_backward(y, out, arg)

    @staticmethod
    def backward(ctx, grad_out, _grad_arg):
        y, out, arg = ctx.saved_tensors
        return (open_window_grad(y, out, arg, grad_out, *ctx.params),) + (None,) * 6

    generate_vmap_rule = False

    @staticmethod
    def vmap(info, in_dims, y, *params):  # traces are independent: a vmapped dimension folds into the batch
        flat, n = _fold(y, in_dims[0])
        out, arg = OpenWindow.apply(flat, *params)
        arg = arg.reshape((n, -1) + tuple(arg.shape[1:])) if arg.numel() else arg.new_empty((n, 0))
        return (out.reshape((n, -1) + tuple(out.shape[1:])), arg), (0, 0)


def open_window(y: torch.Tensor, semantics: int, mode: int, start: int, temperature: float, eps: float, p: int,
                nan_tail: bool = False) -> torch.Tensor:
    """Differentiable ``G[start, inf)`` (mode 0) / ``F[start, inf)`` (mode 1) of the value trace ``y[B, T]``.
    ``nan_tail``: the last ``start`` outputs (empty windows) are NaN instead of a finite filler."""
    return OpenWindow.apply(y, int(semantics), int(mode) | (2 if nan_tail else 0), int(start), float(temperature), float(eps), int(p))[0]


# ---- classical Until without an end beyond one tile: clamp-composition suffix scan (csrc/scan.cu) ------------------
_lib.define("open_until_fwd(Tensor left, Tensor right, int semantics, int start) -> (Tensor, Tensor)")
_lib.define("open_until_grad(Tensor key, Tensor grad_out, int semantics, int start) -> (Tensor, Tensor)")


def _open_until_fwd_cuda(left, right, semantics, start):
    _check_trace(left)
    _check_trace(right)
    if left.shape != right.shape:
        raise ValueError("UNTIL traces must have the same shape")
    left, right = left.contiguous(), right.contiguous()
    B, T = left.shape
    out = torch.empty_like(left)
    key = torch.empty((B, T), dtype=torch.int32, device=left.device)
    with torch.cuda.device(left.device):
        rc = capi.lib().pystl_b200_open_until_forward(int(semantics), int(start), left.data_ptr(), right.data_ptr(), B, T,
                                                      out.data_ptr(), key.data_ptr(), torch.cuda.current_stream(left.device).cuda_stream)
    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
    return out, key


def _open_until_grad_cuda(key, grad_out, semantics, start):
    g = grad_out.to(dtype=torch.float32).contiguous()
    _check_trace(g)
    key = key.contiguous()
    B, T = g.shape
    gl, gr = torch.empty_like(g), torch.empty_like(g)
    with torch.cuda.device(g.device):
        rc = capi.lib().pystl_b200_open_until_backward(int(semantics), int(start), key.data_ptr(), g.data_ptr(), B, T,
                                                       gl.data_ptr(), gr.data_ptr(), torch.cuda.current_stream(g.device).cuda_stream)
    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
    return gl, gr


_lib.impl("open_until_fwd", _open_until_fwd_cuda, "CUDA")
_lib.impl("open_until_grad", _open_until_grad_cuda, "CUDA")
open_until_fwd = torch.ops.pystl_b200.open_until_fwd
open_until_grad = torch.ops.pystl_b200.open_until_grad


@torch.library.register_fake("pystl_b200::open_until_fwd")
def _(left, right, semantics, start):
    return torch.empty_like(left), left.new_empty(tuple(left.shape), dtype=torch.int32)


@torch.library.register_fake("pystl_b200::open_until_grad")
def _(key, grad_out, semantics, start):
    return torch.empty_like(grad_out), torch.empty_like(grad_out)


class OpenUntil(torch.autograd.Function):
    """``out[b, t] = W(t + start)``, the unbounded classical Until of two value traces; cotangents go to the recorded sources."""

    @staticmethod
    def forward(left, right, semantics, start):
        return open_until_fwd(left, right, semantics, start)

    @staticmethod
    def setup_context(ctx, inputs, output):
        _, _, ctx.semantics, ctx.start = inputs
        out, key = output
        ctx.mark_non_differentiable(key)
        ctx.save_for_backward(key)

    @staticmethod
    def backward(ctx, grad_out, _grad_key):
        (key,) = ctx.saved_tensors
        gl, gr = open_until_grad(key, grad_out, ctx.semantics, ctx.start)
        return gl, gr, None, None

    generate_vmap_rule = False

    @staticmethod
    def vmap(info, in_dims, left, right, semantics, start):
        n = info.batch_size
        lf = left.movedim(in_dims[0], 0) if in_dims[0] is not None else left.unsqueeze(0).expand((n,) + tuple(left.shape))
        rt = right.movedim(in_dims[1], 0) if in_dims[1] is not None else right.unsqueeze(0).expand((n,) + tuple(right.shape))
        out, key = OpenUntil.apply(lf.reshape((-1,) + tuple(lf.shape[2:])), rt.reshape((-1,) + tuple(rt.shape[2:])), semantics, start)
        return (out.reshape((n, -1) + tuple(out.shape[1:])), key.reshape((n, -1) + tuple(key.shape[1:]))), (0, 0)


def open_until(left: torch.Tensor, right: torch.Tensor, semantics: int, start: int) -> torch.Tensor:
    """Differentiable unbounded Until of the value traces ``left, right [B, T]`` read at ``t + start`` (classical semantics)."""
    return OpenUntil.apply(left, right, int(semantics), int(start))[0]


# ---- the tensor the formula kernels read when some predicates are leaf channels -------------------------------------
import ctypes as _C


def _ptr_array(tensors):
    return (_C.c_void_p * max(len(tensors), 1))(*[None if t is None else t.data_ptr() for t in tensors])


class Assemble(torch.autograd.Function):
    """``[signal[..., channels] | leaf columns]`` in one pass (``pystl_b200_assemble_channels``; a ``torch.cat`` of an
    advanced-indexed signal and stacked columns is three, and its backward an accumulating ``index_put_``: 130 ms at
    16384 x 16384 samples).  The adjoint (``pystl_b200_scatter_channels``) writes the signal's gradient and one column per leaf."""

    @staticmethod
    def forward(signal, channels, *cols):
        _check(signal)
        sig = signal.contiguous()
        B, T, D = sig.shape
        cols = [c.to(torch.float32).contiguous() for c in cols]
        for c in cols:
            if tuple(c.shape) != (B, T) or not c.is_cuda:
                raise ValueError(f"leaf columns must be CUDA tensors of shape {(B, T)}")
        out = sig.new_empty((B, T, len(channels) + len(cols)))
        ch = (_C.c_int32 * max(len(channels), 1))(*channels)
        with torch.cuda.device(sig.device):
            rc = capi.lib().pystl_b200_assemble_channels(sig.data_ptr(), B * T, D, ch, len(channels), _ptr_array(cols), len(cols),
                                                         out.data_ptr(), torch.cuda.current_stream(sig.device).cuda_stream)
        capi.raise_for(rc, capi.lib().pystl_b200_last_error())
        return out

    @staticmethod
    def setup_context(ctx, inputs, output):
        signal, channels = inputs[0], inputs[1]
        ctx.channels = tuple(channels)
        ctx.D = signal.shape[-1]
        ctx.n_cols = len(inputs) - 2

    @staticmethod
    def backward(ctx, grad):
        if torch._C._functorch.is_functorch_wrapped_tensor(grad):  # vmap(grad(...)): the same adjoint by (batchable) torch ops on the device
            n = len(ctx.channels)
            gs = None
            if ctx.needs_input_grad[0]:
                gs = grad.new_zeros(tuple(grad.shape[:-1]) + (ctx.D,))
                if n:
                    gs = gs.index_copy(-1, torch.tensor(ctx.channels, device=grad.device), grad[..., :n])
            return (gs, None) + tuple(grad[..., n + j] for j in range(ctx.n_cols))
        g = grad.to(torch.float32).contiguous()
        B, T, _ = g.shape
        gs = g.new_empty((B, T, ctx.D)) if ctx.needs_input_grad[0] else None
        gcols = [g.new_empty((B, T)) if ctx.needs_input_grad[2 + j] else None for j in range(ctx.n_cols)]
        ch = (_C.c_int32 * max(len(ctx.channels), 1))(*ctx.channels)
        with torch.cuda.device(g.device):
            rc = capi.lib().pystl_b200_scatter_channels(g.data_ptr(), B * T, ctx.D, ch, len(ctx.channels),
                                                        None if gs is None else gs.data_ptr(), _ptr_array(gcols), ctx.n_cols,
                                                        torch.cuda.current_stream(g.device).cuda_stream)
        capi.raise_for(rc, capi.lib().pystl_b200_last_error())
        return (gs, None) + tuple(gcols)

    generate_vmap_rule = False

    @staticmethod
    def vmap(info, in_dims, signal, channels, *cols):
        n = info.batch_size

        def lead(x, d):
            x = x.movedim(d, 0) if d is not None else x.unsqueeze(0).expand((n,) + tuple(x.shape))
            return x.reshape((-1,) + tuple(x.shape[2:]))

        out = Assemble.apply(lead(signal, in_dims[0]), channels, *[lead(c, d) for c, d in zip(cols, in_dims[2:])])
        return out.reshape((n, -1) + tuple(out.shape[1:])), 0


def assemble(signal: torch.Tensor, channels, cols) -> torch.Tensor:
    if not cols and tuple(channels) == tuple(range(signal.shape[-1])):
        return signal
    if len(channels) > 32 or len(cols) > 16 or signal.shape[-1] > 64:  # beyond the kernel's argument block: the same tensor by torch ops (device side too)
        idx = torch.tensor([int(d) for d in channels], dtype=torch.long, device=signal.device)
        return torch.cat([signal.index_select(-1, idx)] + [c.unsqueeze(-1) for c in cols], dim=-1).contiguous()
    return Assemble.apply(signal, tuple(int(d) for d in channels), *cols)


# ---- batching rules of the adjoint ops (vmap over a backward pass, e.g. vmap(grad(loss)))
def _lead(x, d, n):
    x = x.movedim(d, 0) if d is not None else x.unsqueeze(0).expand((n,) + tuple(x.shape))
    return x.reshape((-1,) + tuple(x.shape[2:]))


@torch.library.register_vmap("pystl_b200::open_window_grad")
def _(info, in_dims, y, out, arg, grad_out, semantics, mode, start, temperature, eps, p):
    n = info.batch_size
    a = _lead(arg, in_dims[2], n) if arg.numel() else arg
    g = open_window_grad(_lead(y, in_dims[0], n), _lead(out, in_dims[1], n), a, _lead(grad_out, in_dims[3], n),
                         semantics, mode, start, temperature, eps, p)
    return g.reshape((n, -1) + tuple(g.shape[1:])), 0


@torch.library.register_vmap("pystl_b200::open_until_grad")
def _(info, in_dims, key, grad_out, semantics, start):
    n = info.batch_size
    gl, gr = open_until_grad(_lead(key, in_dims[0], n), _lead(grad_out, in_dims[1], n), semantics, start)
    return (gl.reshape((n, -1) + tuple(gl.shape[1:])), gr.reshape((n, -1) + tuple(gr.shape[1:]))), (0, 0)


# ---- CT-STL on value traces: the cumulative-time operator and the Until with an exclusive prefix (csrc/select.cu) ---
_lib.define("ctstl_cumulative_fwd(Tensor y, int start, int end, int k) -> (Tensor, Tensor)")
_lib.define("ctstl_cumulative_grad(Tensor arg, Tensor grad_out, int start, int end) -> Tensor")
_lib.define("ctstl_until_fwd(Tensor left, Tensor right, int start, int end) -> (Tensor, Tensor)")
_lib.define("ctstl_until_grad(Tensor key, Tensor grad_out, int start, int end) -> (Tensor, Tensor)")


def _ctstl_cumulative_fwd_cuda(y, start, end, k):
    _check_trace(y)
    y = y.contiguous()
    B, T = y.shape
    out = torch.empty_like(y)
    arg = torch.empty((B, T), dtype=torch.int32, device=y.device)
    with torch.cuda.device(y.device):
        rc = capi.lib().pystl_b200_ctstl_cumulative_forward(y.data_ptr(), B, T, int(start), int(end), int(k), out.data_ptr(),
                                                            arg.data_ptr(), torch.cuda.current_stream(y.device).cuda_stream)
    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
    return out, arg


def _ctstl_cumulative_grad_cuda(arg, grad_out, start, end):
    g = grad_out.to(dtype=torch.float32).contiguous()
    _check_trace(g)
    arg = arg.contiguous()
    B, T = g.shape
    gy = torch.empty_like(g)
    with torch.cuda.device(g.device):
        rc = capi.lib().pystl_b200_ctstl_cumulative_backward(arg.data_ptr(), g.data_ptr(), B, T, int(start), int(end), gy.data_ptr(),
                                                             torch.cuda.current_stream(g.device).cuda_stream)
    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
    return gy


def _ctstl_until_fwd_cuda(left, right, start, end):
    _check_trace(left)
    _check_trace(right)
    if left.shape != right.shape:
        raise ValueError("UNTIL traces must have the same shape")
    left, right = left.contiguous(), right.contiguous()
    B, T = left.shape
    out = torch.empty_like(left)
    key = torch.empty((B, T), dtype=torch.int32, device=left.device)
    with torch.cuda.device(left.device):
        rc = capi.lib().pystl_b200_ctstl_until_forward(left.data_ptr(), right.data_ptr(), B, T, int(start), int(end), out.data_ptr(),
                                                       key.data_ptr(), torch.cuda.current_stream(left.device).cuda_stream)
    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
    return out, key


def _ctstl_until_grad_cuda(key, grad_out, start, end):
    g = grad_out.to(dtype=torch.float32).contiguous()
    _check_trace(g)
    key = key.contiguous()
    B, T = g.shape
    gl, gr = torch.empty_like(g), torch.empty_like(g)
    with torch.cuda.device(g.device):
        rc = capi.lib().pystl_b200_ctstl_until_backward(key.data_ptr(), g.data_ptr(), B, T, int(start), int(end), gl.data_ptr(),
                                                        gr.data_ptr(), torch.cuda.current_stream(g.device).cuda_stream)
    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
    return gl, gr


_lib.impl("ctstl_cumulative_fwd", _ctstl_cumulative_fwd_cuda, "CUDA")
_lib.impl("ctstl_cumulative_grad", _ctstl_cumulative_grad_cuda, "CUDA")
_lib.impl("ctstl_until_fwd", _ctstl_until_fwd_cuda, "CUDA")
_lib.impl("ctstl_until_grad", _ctstl_until_grad_cuda, "CUDA")


@torch.library.register_fake("pystl_b200::ctstl_cumulative_fwd")
def _(y, start, end, k):
    return torch.empty_like(y), y.new_empty(tuple(y.shape), dtype=torch.int32)


@torch.library.register_fake("pystl_b200::ctstl_cumulative_grad")
def _(arg, grad_out, start, end):
    return torch.empty_like(grad_out)


@torch.library.register_fake("pystl_b200::ctstl_until_fwd")
def _(left, right, start, end):
    return torch.empty_like(left), left.new_empty(tuple(left.shape), dtype=torch.int32)


@torch.library.register_fake("pystl_b200::ctstl_until_grad")
def _(key, grad_out, start, end):
    return torch.empty_like(grad_out), torch.empty_like(grad_out)


class CtstlCumulative(torch.autograd.Function):
    """``out[b, t]`` = k-th largest of the window ``[t + start, min(t + end, T - 1)]`` of ``y[b]``; the cotangent of an output goes
    to the sample it selected."""

    @staticmethod
    def forward(y, start, end, k):
        return torch.ops.pystl_b200.ctstl_cumulative_fwd(y, start, end, k)

    @staticmethod
    def setup_context(ctx, inputs, output):
        _, ctx.start, ctx.end, _ = inputs
        out, arg = output
        ctx.mark_non_differentiable(arg)
        ctx.save_for_backward(arg)

    @staticmethod
    def backward(ctx, grad_out, _grad_arg):
        (arg,) = ctx.saved_tensors
        return torch.ops.pystl_b200.ctstl_cumulative_grad(arg, grad_out, ctx.start, ctx.end), None, None, None

    generate_vmap_rule = False

    @staticmethod
    def vmap(info, in_dims, y, start, end, k):
        yy = y.movedim(in_dims[0], 0)
        out, arg = CtstlCumulative.apply(yy.reshape((-1,) + tuple(yy.shape[2:])), start, end, k)
        n = info.batch_size
        return (out.reshape((n, -1) + tuple(out.shape[1:])), arg.reshape((n, -1) + tuple(arg.shape[1:]))), (0, 0)


class CtstlUntil(torch.autograd.Function):
    """CT-STL Until of two value traces (exclusive left prefix); cotangents go to the recorded sources."""

    @staticmethod
    def forward(left, right, start, end):
        return torch.ops.pystl_b200.ctstl_until_fwd(left, right, start, end)

    @staticmethod
    def setup_context(ctx, inputs, output):
        _, _, ctx.start, ctx.end = inputs
        out, key = output
        ctx.mark_non_differentiable(key)
        ctx.save_for_backward(key)

    @staticmethod
    def backward(ctx, grad_out, _grad_key):
        (key,) = ctx.saved_tensors
        gl, gr = torch.ops.pystl_b200.ctstl_until_grad(key, grad_out, ctx.start, ctx.end)
        return gl, gr, None, None

    generate_vmap_rule = False

    @staticmethod
    def vmap(info, in_dims, left, right, start, end):
        n = info.batch_size
        lf = left.movedim(in_dims[0], 0) if in_dims[0] is not None else left.unsqueeze(0).expand((n,) + tuple(left.shape))
        rt = right.movedim(in_dims[1], 0) if in_dims[1] is not None else right.unsqueeze(0).expand((n,) + tuple(right.shape))
        out, key = CtstlUntil.apply(lf.reshape((-1,) + tuple(lf.shape[2:])), rt.reshape((-1,) + tuple(rt.shape[2:])), start, end)
        return (out.reshape((n, -1) + tuple(out.shape[1:])), key.reshape((n, -1) + tuple(key.shape[1:]))), (0, 0)


@torch.library.register_vmap("pystl_b200::ctstl_cumulative_grad")
def _(info, in_dims, arg, grad_out, start, end):
    n = info.batch_size
    a, g = _lead(arg, in_dims[0], n), _lead(grad_out, in_dims[1], n)
    gy = torch.ops.pystl_b200.ctstl_cumulative_grad(a, g, start, end)
    return gy.reshape((n, -1) + tuple(gy.shape[1:])), 0


@torch.library.register_vmap("pystl_b200::ctstl_until_grad")
def _(info, in_dims, key, grad_out, start, end):
    n = info.batch_size
    kk, g = _lead(key, in_dims[0], n), _lead(grad_out, in_dims[1], n)
    gl, gr = torch.ops.pystl_b200.ctstl_until_grad(kk, g, start, end)
    return (gl.reshape((n, -1) + tuple(gl.shape[1:])), gr.reshape((n, -1) + tuple(gr.shape[1:]))), (0, 0)


def ctstl_cumulative(y: torch.Tensor, start: int, end: Optional[int], k: int) -> torch.Tensor:
    """Differentiable ``C^tau_I`` of the value trace ``y [B, T]``: k-th largest of the clipped window at every t (-inf where the window
    holds fewer than k samples); ``end=None``: no end."""
    return CtstlCumulative.apply(y, int(start), -1 if end is None else int(end), int(k))[0]


def ctstl_until(left: torch.Tensor, right: torch.Tensor, start: int, end: Optional[int]) -> torch.Tensor:
    """Differentiable CT-STL Until of the value traces ``left, right [B, T]`` (NaN where the window is empty)."""
    return CtstlUntil.apply(left, right, int(start), -1 if end is None else int(end))[0]
