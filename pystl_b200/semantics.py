"""The five semantics of the ``b200`` backend.

Each class implements the reference's seven-method operator interface
(``pystl/semantics/base.py:18-75``: ``predicate, boolean_not, boolean_and,
boolean_or, temporal_always, temporal_eventually, temporal_until``) so that it
can be handed to ``Formula.evaluate(signal, semantics, t)`` like any other
backend, and mirrors the constructor kwargs of ``pystl/torch/semantics.py``
(``temperature`` for smooth, ``eps`` / ``p`` for dgmsr; the others take none and
raise ``TypeError`` on unknown kwargs, like the reference classes).

The real work happens one level up, where the reference already lets a
semantics object take the whole formula (``Formula.evaluate_with_grad`` ->
``semantics.evaluate_with_grad``, api.py:64-78):

* ``robustness_trace(formula, signal[B, T, D])`` -- every ``t`` of every trace in
  one fused kernel launch; differentiable (``torch.autograd.Function``);
* ``evaluate_formula(formula, signal, t)`` -- value at one ``t`` (what
  ``phi.evaluate(signal, sem, t)`` returns), raising where the reference raises;
* ``evaluate_with_grad(formula, signal, t=...)`` -- ``(value, d value / d signal)``.

Tensors go to the CUDA library through the C ABI (``capi``) as raw pointers on
the current stream.  There is no CPU path: a CPU tensor is copied to the
device, evaluated there and copied back; without the CUDA library or a GPU
every call raises.
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Any, Dict, Optional, Sequence, Tuple

import torch

from . import capi
from .compiler import CompiledFormula, compile_formula, leaf_column, leaf_values
from .formula import Predicate, affine_predicate


@dataclass(frozen=True)
class B200CumulativeRobustness:
    """``(pos, neg)`` pair, like ``TorchCumulativeRobustness`` (torch/semantics.py:218-223)."""

    pos: Any
    neg: Any


def _device_of(signal: torch.Tensor) -> torch.device:
    if signal.is_cuda:
        return signal.device
    if not torch.cuda.is_available():
        raise RuntimeError("the b200 backend needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


try:  # plug into the reference's class hierarchy when it is installed (semantics/base.py:18-75)
    from pystl.semantics.base import Semantics as _ReferenceSemantics  # type: ignore
except Exception:  # the GPU box has no `pystl`: same interface, no common base
    _ReferenceSemantics = object


class _B200Semantics(_ReferenceSemantics):
    syntax = ""
    out_channels = 1
    force_open_leaf = False  # True: open-ended Always / Eventually always go through the suffix-scan kernels

    def __init__(self) -> None:
        self._cache: Dict[Tuple, Tuple[Any, CompiledFormula]] = {}
        self._verified: Dict[Tuple, Tuple[Any, Tuple[int, ...]]] = {}
        self._list_formulas: Dict[Tuple, Any] = {}
        self._open: Dict[Tuple, Tuple[Any, bool]] = {}

    # ---- kwargs forwarded to the C ABI -------------------------------------------------
    def _params(self) -> dict:
        return {}

    # ---- formula cache (the only mutable state; keyed by object identity + D) ----------
    def compile(self, formula: Any, n_channels: int, force_leaf: Tuple[int, ...] = (), open_leaf: bool = False) -> CompiledFormula:
        # the semantics' own parameters are part of the key: changing sem.temperature / eps / p
        # after the first call recompiles instead of being silently ignored
        key = (id(formula), int(n_channels), tuple(sorted(self._params().items())), tuple(force_leaf), bool(open_leaf))
        hit = self._cache.get(key)
        if hit is not None and hit[0] is formula:
            return hit[1]
        cf = compile_formula(formula, self.syntax, n_channels, force_leaf=force_leaf, open_leaf=open_leaf, **self._params())
        if len(self._cache) > 256:
            self._cache.clear()
        self._cache[key] = (formula, cf)
        return cf

    def _checked_compile(self, formula: Any, sig: torch.Tensor) -> CompiledFormula:
        """``compile`` + the check the probe alone cannot give: predicates recognised as affine by
        probing ``fn`` are compared with ``fn`` on rows of the REAL signal (first and last trace, a
        sample of ``t`` including the last one) once per (formula, D, T); a predicate that disagrees is
        recompiled as a leaf channel (``compiler.verify_probed``)."""
        B, T, D = sig.shape
        cf = self.compile(formula, D)
        open_leaf = cf.has_open and self._needs_open_leaf(formula, cf, T)
        if open_leaf:
            cf = self.compile(formula, D, open_leaf=True)
        if not cf.probed_preds or B == 0 or T == 0:
            return cf
        vkey = (id(formula), D, T)
        hit = self._verified.get(vkey)
        if hit is None or hit[0] is not formula:
            from .compiler import verify_probed

            rows = sig[[0, B - 1]] if B > 1 else sig[:1]
            force = verify_probed(cf, rows.detach().to(torch.float64).cpu().numpy())
            if len(self._verified) > 256:
                self._verified.clear()
            hit = self._verified[vkey] = (formula, force)
        return self.compile(formula, D, hit[1], open_leaf) if hit[1] else cf

    def _needs_open_leaf(self, formula: Any, cf: CompiledFormula, T: int) -> bool:
        """An ``Always / Eventually`` without an end is evaluated inside the formula kernel while the whole trace
        fits one tile (the plan says so); beyond that it goes through the suffix-scan kernels as a leaf channel
        (``compiler.OpenLeaf``).  Decided once per (formula, D, T, parameters)."""
        if self.force_open_leaf or os.environ.get("PYSTL_B200_OPEN_LEAF") == "1":
            return True  # (tests: the scan path against the in-kernel path at any T)
        key = (id(formula), cf.n_signal_channels, int(T), tuple(sorted(self._params().items())))
        hit = self._open.get(key)
        if hit is None or hit[0] is not formula:
            fails = []
            for backward in (0, 1):  # the gradient kernel keeps more planes per tile: its limit on T is lower
                p = capi.Plan()
                rc = capi.lib().pystl_b200_plan(cf.ptr(), int(T), cf.n_channels, backward, C.byref(p))
                # (lower.cpp: plan_geometry / codegen.cpp: fused_geometry; any other status is reported by the evaluation)
                fails.append(rc == capi.E_UNSUPPORTED and b"open-ended" in (capi.lib().pystl_b200_last_error() or b""))
            need = any(fails)
            if need and not fails[0]:  # values still fit one tile: keep them there when the scan path cannot take over
                try:
                    self.compile(formula, cf.n_signal_channels, open_leaf=True)
                except NotImplementedError:
                    need = False
            if len(self._open) > 256:
                self._open.clear()
            hit = self._open[key] = (formula, need)
        return hit[1]

    def invalidate(self, formula: Any = None) -> None:
        """Forget what was compiled for ``formula`` (all formulas when None): a predicate's ``fn`` is
        sampled once, so rebinding a value its closure reads needs this call to take effect."""
        if formula is None:
            self._cache.clear()
            self._verified.clear()
            self._open.clear()
            return
        for d in (self._cache, self._verified, self._open):
            for k in [k for k in d if k[0] == id(formula)]:
                d.pop(k, None)

    def precompile(self, formula: Any, horizon: int, n_channels: int, *, backward: bool = True) -> int:
        """Compile the fused kernels ``formula`` needs on traces of ``horizon`` samples into the on-disk cubin cache (no GPU
        needed), following the host-side split at open-ended operators on long traces.  Returns the
        number of kernels compiled or found; formulas that run on the interpreter tier count as 0."""
        from .compiler import OpenLeaf, OpenUntilLeaf

        cf = self.compile(formula, n_channels)
        if self.force_open_leaf or (cf.has_open and self._needs_open_leaf(formula, cf, int(horizon))):
            cf = self.compile(formula, n_channels, open_leaf=True)
        done = 0
        if cf.root_leaf is None:
            for bwd in ((0, 1) if backward else (0,)):
                rc = capi.lib().pystl_b200_fused_precompile(cf.ptr(), int(horizon), cf.n_channels, bwd)
                if rc == capi.OK:
                    done += 1
                elif rc != capi.E_UNSUPPORTED:
                    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
        for leaf in cf.leaf_preds:
            if isinstance(leaf, OpenLeaf):
                done += self.precompile(leaf.child, horizon, n_channels, backward=backward)
            elif isinstance(leaf, OpenUntilLeaf):
                done += self.precompile(leaf.left, horizon, n_channels, backward=backward)
                done += self.precompile(leaf.right, horizon, n_channels, backward=backward)
        return done

    def plan(self, formula: Any, horizon: int, n_channels: int, cf: Optional[CompiledFormula] = None) -> dict:
        """Tiling / validity report for ``formula`` on traces of ``horizon`` samples (host only)."""
        if cf is None:
            cf = self.compile(formula, n_channels)
            if cf.has_open and self._needs_open_leaf(formula, cf, int(horizon)):
                cf = self.compile(formula, n_channels, open_leaf=True)
        p = capi.Plan()
        rc = capi.lib().pystl_b200_plan(cf.ptr(), int(horizon), cf.n_channels, 0, C.byref(p))
        capi.raise_for(rc, capi.lib().pystl_b200_last_error())
        return p.as_dict()

    # ---- whole-formula evaluation --------------------------------------------------------
    def _forward_raw(self, cf: CompiledFormula, sig: torch.Tensor, mean_acc: Optional[torch.Tensor]) -> torch.Tensor:
        B, T, D = sig.shape
        out = torch.empty((B, T, self.out_channels), dtype=torch.float32, device=sig.device)
        stream = torch.cuda.current_stream(sig.device).cuda_stream
        with torch.cuda.device(sig.device):
            rc = capi.lib().pystl_b200_eval_forward(
                cf.ptr(), sig.data_ptr(), B, T, D, out.data_ptr(),
                None if mean_acc is None else mean_acc.data_ptr(), stream,
            )
        capi.raise_for(rc, capi.lib().pystl_b200_last_error())
        return out

    def _prepare(self, signal: Any) -> Tuple[torch.Tensor, bool, torch.device]:
        sig = torch.as_tensor(signal)
        squeeze = sig.dim() == 2
        if squeeze:
            sig = sig.unsqueeze(0)
        if sig.dim() != 3:
            raise ValueError("signal must have shape (time, state_dim) or (batch, time, state_dim).")
        home = sig.device
        dev = _device_of(sig)
        sig = sig.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
        return sig, squeeze, home

    def robustness_trace(self, formula: Any, signal: Any) -> Any:
        """Robustness at every ``t``: ``[B, T]`` for ``signal[B, T, D]`` (``[T]`` for ``[T, D]``).

        Cumulative returns ``B200CumulativeRobustness(pos, neg)`` of such tensors.  Entries the
        reference would raise ``ValueError`` on (empty window, api.py:218-221) are NaN.
        Differentiable: the evaluation is the dispatcher op ``pystl_b200::robustness`` (ops.py), whose
        autograd formula is the backward entry point of the C ABI; it also carries a fake kernel and a
        batching rule, so ``torch.compile`` and ``torch.func.vmap / grad`` work through it."""
        from . import ops

        sig, squeeze, home = self._prepare(signal)  # (stays connected to `signal` in the autograd graph)
        cf = self._checked_compile(formula, sig)
        if cf.root_leaf is not None:  # the formula IS one open-ended window / until: the scan kernels write its robustness
            out = self._root_leaf_trace(cf, sig, nan_tail=True)
        else:
            if cf.leaf_preds:  # non-affine predicates: leaf channels computed by torch, inside the graph
                sig = leaf_values(cf, sig, self)
            out = ops.evaluate(sig, ops.handle_of(cf, self.out_channels))
        if home != out.device:
            out = out.to(home)
        if squeeze:
            out = out[0]
        if self.out_channels == 2:
            return B200CumulativeRobustness(pos=out[..., 0], neg=out[..., 1])
        return out[..., 0]

    def _root_leaf_trace(self, cf: CompiledFormula, sig: torch.Tensor, nan_tail: bool) -> torch.Tensor:
        """``[B, T, C]`` robustness of a formula whose root is an open-ended window / until leaf (``cf.root_leaf``): the scan
        kernels' output, no formula kernel and no assembled input tensor.  ``slack = -1``: the reference raises at every t."""
        if cf.preds[0].slack < 0:
            nan = torch.full(sig.shape[:2], float("nan"), dtype=torch.float32, device=sig.device) + 0.0 * sig[..., 0]
            return torch.stack([nan] * self.out_channels, dim=-1)
        col = leaf_column(cf, cf.leaf_preds[cf.root_leaf], sig, self, nan_tail=nan_tail)
        return torch.stack(col, dim=-1) if isinstance(col, tuple) else col.unsqueeze(-1)  # [B, T, C]

    def _root_leaf_valid(self, cf: CompiledFormula, T: int) -> int:
        slack = cf.preds[0].slack
        return 0 if slack < 0 else max(T - slack, 0)

    def mean_robustness(self, formula: Any, signal: Any, *, group=None) -> torch.Tensor:
        """Mean of the valid outputs (cumulative: of ``pos + neg``), reduced inside the kernel.

        With ``torch.distributed`` initialised the ``(sum, count)`` pair is all-reduced (one
        16-byte NCCL all-reduce) so that every rank gets the mean over the whole sharded batch."""
        from .distributed import all_reduce_mean

        sig, _, _ = self._prepare(signal)
        cf = self._checked_compile(formula, sig)
        acc = torch.zeros(2, dtype=torch.float64, device=sig.device)
        if cf.root_leaf is not None:
            valid = self._root_leaf_valid(cf, sig.shape[1])
            with torch.no_grad():
                if valid > 0:
                    acc[0] = self._root_leaf_trace(cf, sig, nan_tail=False)[:, :valid].sum(dim=(1, 2)).sum(dtype=torch.float64)
            acc[1] = float(sig.shape[0] * valid)
            return all_reduce_mean(acc, group=group)
        plan = self.plan(formula, sig.shape[1], sig.shape[2], cf)
        with torch.no_grad():
            sig = leaf_values(cf, sig, self)
        self._forward_raw(cf, sig, acc)
        valid = 0 if plan["raises_always"] else max(sig.shape[1] - plan["start_slack"], 0)
        acc[1] = float(sig.shape[0] * valid)
        return all_reduce_mean(acc, group=group)

    def mean_robustness_and_grad(self, formula: Any, signal: Any, *, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """``(mean robustness, d mean / d signal)`` in ONE pass over the signal.

        The batched, all-``t`` counterpart of ``Formula.evaluate_with_grad`` (api.py:64-78): for
        formulas on the fused tier a single kernel reads the signal once and writes the robustness
        trace, its fused mean and the gradient (no forward recomputation); the cotangent of the
        mean, ``1 / count``, is a constant and needs no tensor.  Sharded batches: ``count`` and the
        mean are all-reduced over ``group``, the gradient needs no collective."""
        from .distributed import all_reduce_mean, all_reduce_count

        sig, squeeze, home = self._prepare(signal)
        sig = sig.detach()
        B, T, D = sig.shape
        cf = self._checked_compile(formula, sig)
        if cf.root_leaf is not None:  # the scan kernels and their adjoints are the whole evaluation
            valid = self._root_leaf_valid(cf, T)
            n_local = float(B * valid)
            n_total = all_reduce_count(n_local, sig.device, group=group)
            acc = torch.zeros(2, dtype=torch.float64, device=sig.device)
            if valid > 0:
                base = sig.detach().requires_grad_(True)
                with torch.enable_grad():
                    # (row sums in fp32, their sum in float64: summing with dtype=float64 materialises a float64 copy of the trace)
                    total = self._root_leaf_trace(cf, base, nan_tail=False)[:, :valid].sum(dim=(1, 2)).sum(dtype=torch.float64)
                # (torch.full: a fill kernel; torch.tensor(scalar, device=...) is a pageable host-to-device copy that makes the host
                # wait for everything queued before it)
                (grad,) = torch.autograd.grad(total, base, torch.full((), (1.0 / n_total) if n_total > 0 else 0.0, dtype=torch.float64,
                                                                      device=sig.device))
                acc[0] = total.detach()
            else:
                grad = torch.zeros_like(sig)
            acc[1] = n_local
            mean = all_reduce_mean(acc, group=group)
            if home != grad.device:
                grad = grad.to(home)
            return mean, (grad[0] if squeeze else grad)
        plan = self.plan(formula, T, D, cf)
        valid = 0 if plan["raises_always"] else max(T - plan["start_slack"], 0)
        n_local = float(B * valid)
        n_total = all_reduce_count(n_local, sig.device, group=group)
        out = torch.empty((B, T, self.out_channels), dtype=torch.float32, device=sig.device)
        acc = torch.zeros(2, dtype=torch.float64, device=sig.device)
        base = sig
        if cf.leaf_preds:  # the kernel differentiates up to the leaf channels, autograd through fn
            base = sig.detach().requires_grad_(True)
            with torch.enable_grad():
                sig = leaf_values(cf, base, self)
        kin = sig.detach()
        grad = torch.empty_like(kin)
        stream = torch.cuda.current_stream(sig.device).cuda_stream
        with torch.cuda.device(sig.device):
            rc = capi.lib().pystl_b200_eval_value_and_grad(
                cf.ptr(), kin.data_ptr(), B, T, cf.n_channels, None, (1.0 / n_total) if n_total > 0 else 0.0,
                out.data_ptr(), acc.data_ptr(), grad.data_ptr(), stream,
            )
        capi.raise_for(rc, capi.lib().pystl_b200_last_error())
        if cf.leaf_preds:
            (grad,) = torch.autograd.grad(sig, base, grad)
        acc[1] = n_local
        mean = all_reduce_mean(acc, group=group)
        if home != grad.device:
            grad = grad.to(home)
        return mean, (grad[0] if squeeze else grad)

    def mean_robustness_and_grad_host(self, formula: Any, signal_host: torch.Tensor, grad_host: Optional[torch.Tensor] = None, *,
                                      chunk_traces: int = 0, device: Optional[torch.device] = None, group=None,
                                      with_grad: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """``mean_robustness_and_grad`` for a batch that lives in HOST memory (pinned for full speed).

        The batch is streamed through the GPU in chunks of traces on three streams: the H2D copy of
        chunk i+1 and the D2H copy of the gradient of chunk i-1 overlap the kernel of chunk i (two
        device buffers each), so a step costs max(H2D, kernel, D2H) instead of their sum and the device
        never holds more than two chunks.  Returns ``(mean, grad_host)``; ``grad_host`` (allocated
        pinned when not given) holds ``d mean / d signal`` and is complete when the call returns.
        ``with_grad=False`` streams the forward kernel only (mean robustness)."""
        from .distributed import all_reduce_count, all_reduce_mean

        sig_h = torch.as_tensor(signal_host)
        if sig_h.dim() != 3 or sig_h.dtype != torch.float32 or sig_h.is_cuda:
            raise ValueError("signal_host must be a float32 CPU tensor of shape (batch, time, state_dim)")
        sig_h = sig_h.contiguous()
        B, T, D = sig_h.shape
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        if not torch.cuda.is_available():
            raise RuntimeError("the b200 backend needs a CUDA device; there is no CPU fallback")
        probe = sig_h[:1].to(dev)
        cf = self._checked_compile(formula, probe)
        if cf.leaf_preds:
            raise NotImplementedError("streaming evaluation takes affine predicates only (leaf channels need autograd on the whole batch)")
        plan = self.plan(formula, T, D, cf)
        valid = 0 if plan["raises_always"] else max(T - plan["start_slack"], 0)
        n_local = float(B * valid)
        n_total = all_reduce_count(n_local, dev, group=group)
        scale = (1.0 / n_total) if n_total > 0 else 0.0
        if chunk_traces <= 0:  # about 128 MiB of signal per chunk: long enough to run PCIe at full rate
            chunk_traces = max(1, min(B, (1 << 27) // max(T * D * 4, 1)))
        n_chunks = (B + chunk_traces - 1) // chunk_traces if B else 0
        if with_grad and grad_host is None:
            grad_host = torch.empty_like(sig_h, pin_memory=True)
        C_out = self.out_channels
        cur = torch.cuda.current_stream(dev)
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        acc = torch.zeros(2, dtype=torch.float64, device=dev)
        bufs = [
            dict(sig=torch.empty((chunk_traces, T, D), dtype=torch.float32, device=dev),
                 grad=torch.empty((chunk_traces, T, D), dtype=torch.float32, device=dev) if with_grad else None,
                 out=torch.empty((chunk_traces, T, C_out), dtype=torch.float32, device=dev),
                 loaded=torch.cuda.Event(), computed=torch.cuda.Event(), drained=torch.cuda.Event())
            for _ in range(min(2, max(n_chunks, 1)))
        ]
        lib = capi.lib()
        s_in.wait_stream(cur)
        with torch.cuda.device(dev):
            for i in range(n_chunks):
                b = bufs[i % len(bufs)]
                lo, hi = i * chunk_traces, min((i + 1) * chunk_traces, B)
                n = hi - lo
                with torch.cuda.stream(s_in):
                    if i >= len(bufs):
                        s_in.wait_event(b["computed"])  # the kernel that read this buffer has finished
                    b["sig"][:n].copy_(sig_h[lo:hi], non_blocking=True)
                    b["loaded"].record(s_in)
                cur.wait_event(b["loaded"])
                if with_grad and i >= len(bufs):
                    cur.wait_event(b["drained"])  # its previous gradient has left for the host
                if with_grad:
                    rc = lib.pystl_b200_eval_value_and_grad(cf.ptr(), b["sig"].data_ptr(), n, T, D, None, scale, b["out"].data_ptr(),
                                                            acc.data_ptr(), b["grad"].data_ptr(), cur.cuda_stream)
                else:
                    rc = lib.pystl_b200_eval_forward(cf.ptr(), b["sig"].data_ptr(), n, T, D, b["out"].data_ptr(), acc.data_ptr(),
                                                     cur.cuda_stream)
                capi.raise_for(rc, lib.pystl_b200_last_error())
                b["computed"].record(cur)
                if with_grad:
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(b["computed"])
                        grad_host[lo:hi].copy_(b["grad"][:n], non_blocking=True)
                        b["drained"].record(s_out)
        cur.wait_stream(s_out)
        cur.wait_stream(s_in)
        acc[1] = n_local
        mean = all_reduce_mean(acc, group=group)
        cur.synchronize()  # grad_host is complete
        return mean, grad_host

    def evaluate_formula(self, formula: Any, signal: Any, t: int = 0) -> Any:
        """``phi.evaluate(signal, sem, t)`` of the reference, for ``signal[T, D]`` or ``[B, T, D]``."""
        sig = torch.as_tensor(signal)
        T = sig.shape[-2]
        if t < 0 or t >= T:
            raise IndexError(f"t={t} is out of bounds for horizon={T}.")
        plan = self.plan(formula, T, sig.shape[-1])
        if plan["raises_always"] or t > T - 1 - plan["start_slack"]:
            raise ValueError(f"temporal window is empty at t={t} for a horizon of {T} samples.")
        out = self.robustness_trace(formula, signal)
        if self.out_channels == 2:
            return B200CumulativeRobustness(pos=out.pos[..., t], neg=out.neg[..., t])
        return out[..., t]

    def evaluate_with_grad(self, formula: Any, signal: Any, *, t: int = 0, **_unused: Any):
        """``(rho_t, d rho_t / d signal)`` like ``DgmsrSemantics.evaluate_with_grad`` (dgmsr.py:120-274)."""
        import numpy as np

        x = torch.as_tensor(signal, dtype=torch.float32)
        if x.dim() != 2:
            raise ValueError("signal must be a 2D array with shape (time, state_dim).")
        if t < 0 or t >= x.shape[0]:
            raise IndexError(f"t={t} is out of bounds for signal with horizon={x.shape[0]}.")
        x = x.to(_device_of(x)).requires_grad_(True)
        val = self.evaluate_formula(formula, x, t)
        scalar = (val.pos + val.neg) if self.out_channels == 2 else val
        (grad,) = torch.autograd.grad(scalar, x)
        if self.out_channels == 2:
            value = B200CumulativeRobustness(pos=float(val.pos.detach()), neg=float(val.neg.detach()))
        else:
            value = float(val.detach())
        return value, np.asarray(grad.detach().cpu().numpy(), dtype=float)

    # ---- the reference's operator interface (base.py:21-75) -------------------------------
    # This is what the reference's own recursion (api.py:142-302) calls, once per node and t, with
    # LISTS of backend-native values.  Values here are torch tensors of any (broadcastable) shape on
    # a CUDA device.  A list of n values is laid out along the TIME axis of a one-channel signal and
    # reduced by the fused window kernel -- Always / Eventually over [0, n-1] read at t = 0 is the
    # same reduction (base.py:47-61), and a window has no limit on n where an n-ary And / Or
    # instruction takes at most 8 operands.
    def predicate(self, predicate: Any, signal: Any, t: int) -> Any:
        if predicate.fn is None:
            raise ValueError(f"Predicate '{predicate.name}' does not define `fn(signal, t)`.")
        sig = torch.as_tensor(signal)
        value = torch.as_tensor(predicate.fn(sig, t), dtype=torch.float32, device=_device_of(sig))
        return self._leaf(value)

    def _leaf(self, value: torch.Tensor) -> Any:
        return value

    def boolean_not(self, value: Any) -> Any:
        return -torch.as_tensor(value)

    def _stack(self, values: Sequence[Any], where: str) -> torch.Tensor:
        if len(values) == 0:
            raise ValueError(f"{where} requires at least one value.")
        vals = [torch.as_tensor(v, dtype=torch.float32) for v in values]
        dev = next((v.device for v in vals if v.is_cuda), None)
        if dev is None:
            dev = _device_of(vals[0])
        vals = torch.broadcast_tensors(*[v.to(dev) for v in vals])
        return torch.stack(vals, dim=-1)  # [..., n]

    def _list_formula(self, kind: str, n: int, weights) -> Any:
        """``Always`` / ``Eventually`` over ``[0, n-1]`` of channel 0, cached per (kind, n, weights)."""
        from .formula import Always, Eventually

        w = None if weights is None else tuple(float(x) for x in weights)
        key = (kind, n, w)
        phi = self._list_formulas.get(key)
        if phi is None:
            if len(self._list_formulas) > 128:
                self._list_formulas.clear()
            leaf = affine_predicate("v", {0: 1.0}, 0.0)
            phi = (Always if kind == "and" else Eventually)(leaf, (0, n - 1), weights=w)
            self._list_formulas[key] = phi
        return phi

    def _reduce(self, values: Sequence[Any], kind: str, weights, where: str) -> torch.Tensor:
        x = self._stack(values, where)  # [..., n]
        n = x.shape[-1]
        out = self.robustness_trace(self._list_formula(kind, n, weights), x.reshape(-1, n, 1))
        return out[:, 0].reshape(x.shape[:-1])

    def boolean_and(self, values: Sequence[Any], *, weights: Optional[Sequence[float]] = None) -> Any:
        return self._reduce(values, "and", weights, "boolean_and")

    def boolean_or(self, values: Sequence[Any], *, weights: Optional[Sequence[float]] = None) -> Any:
        return self._reduce(values, "or", weights, "boolean_or")

    def temporal_always(self, values: Sequence[Any], *, weights: Optional[Sequence[float]] = None) -> Any:
        return self._reduce(values, "and", weights, "temporal_always")

    def temporal_eventually(self, values: Sequence[Any], *, weights: Optional[Sequence[float]] = None) -> Any:
        return self._reduce(values, "or", weights, "temporal_eventually")

    @staticmethod
    def _until_window(left_trace, right_trace, start: int, end: Optional[int]) -> Tuple[int, int]:
        # same checks, same messages as classic.py:50-61 / cumulative.py:85-94
        if len(left_trace) == 0 or len(right_trace) == 0:
            raise ValueError("UNTIL traces must be non-empty.")
        if len(left_trace) != len(right_trace):
            raise ValueError("UNTIL traces must have the same length.")
        if start < 0:
            raise ValueError("UNTIL start must be >= 0.")
        n = len(left_trace)
        last = n - 1 if end is None else min(end, n - 1)
        if start > last:
            raise ValueError("UNTIL temporal window is empty.")
        return n, last

    def _until_value(self, l: torch.Tensor, r: torch.Tensor, start: int, last: int, weights_left, weights_right, weights_pair):
        """``l, r``: ``[..., n]``; Until over ``[start, last]`` (already clipped to the traces) at t = 0."""
        from .formula import Until

        l, r = torch.broadcast_tensors(l, r)
        n = l.shape[-1]
        sig = torch.stack((l, r), dim=-1).reshape(-1, n, 2)  # [batch, time, (l, r)]
        key = ("until", start, last, None if weights_left is None else tuple(weights_left),
               None if weights_right is None else tuple(weights_right), tuple(weights_pair))
        phi = self._list_formulas.get(key)
        if phi is None:
            if len(self._list_formulas) > 128:
                self._list_formulas.clear()
            phi = Until(
                affine_predicate("l", {0: 1.0}), affine_predicate("r", {1: 1.0}), (start, last),
                weights_left=weights_left, weights_right=weights_right, weights_pair=weights_pair,
            )
            self._list_formulas[key] = phi
        return self.robustness_trace(phi, sig), l.shape[:-1]

    def temporal_until(
        self,
        *,
        left_trace: Sequence[Any],
        right_trace: Sequence[Any],
        start: int,
        end: Optional[int],
        weights_left: Optional[Sequence[float]] = None,
        weights_right: Optional[Sequence[float]] = None,
        weights_pair: Sequence[float] = (1.0, 1.0),
    ) -> Any:
        _, last = self._until_window(left_trace, right_trace, start, end)
        l = self._stack(left_trace, "temporal_until")  # [..., n]
        r = self._stack(right_trace, "temporal_until")
        out, shape = self._until_value(l, r, start, last, weights_left, weights_right, weights_pair)
        return out[:, 0].reshape(shape)


class B200ClassicRobustSemantics(_B200Semantics):
    """Hard min / max (``TorchClassicRobustSemantics``, torch/semantics.py:87-146)."""

    syntax = "classical"


class B200SmoothRobustSemantics(_B200Semantics):
    """Log-sum-exp soft min / max (``TorchSmoothRobustSemantics``, torch/semantics.py:149-215)."""

    syntax = "smooth"

    def __init__(self, *, temperature: float = 1.0) -> None:
        super().__init__()
        if temperature <= 0.0:  # torch/semantics.py:13-16
            raise ValueError("temperature must be > 0.")
        self.temperature = float(temperature)

    def _params(self) -> dict:
        return {"temperature": self.temperature}


class B200CumulativeSemantics(_B200Semantics):
    """``(pos, neg)`` cumulative robustness (``TorchCumulativeSemantics``, torch/semantics.py:226-345)."""

    syntax = "cumulative"
    out_channels = 2

    def _leaf(self, value: torch.Tensor) -> Any:
        zero = torch.zeros_like(value)
        return B200CumulativeRobustness(pos=torch.maximum(zero, value), neg=torch.minimum(zero, value))

    def boolean_not(self, value: Any) -> Any:
        return B200CumulativeRobustness(pos=-value.neg, neg=-value.pos)

    # (pos, neg) lists (cumulative.py:35-103).  The two channels never mix, pos >= 0 and neg <= 0
    # throughout (predicates clip, Not swaps and negates, min / max / sum keep the sign), and a
    # cumulative predicate maps a non-negative scalar p to (p, 0) and a non-positive n to (0, n): so the
    # list of pos values goes through the cumulative kernel as a one-channel signal and the answer is the
    # pos channel of the result, likewise neg.  min / max are the Always / Eventually windows of the
    # classic kernel (cumulative Eventually is a SUM, cumulative.py:61-72, so max cannot use it).
    def _channels(self, values, where: str):
        if len(values) == 0:
            raise ValueError(f"{where} requires at least one value.")
        return [v.pos for v in values], [v.neg for v in values]

    def _minmax(self, values, kind: str, where: str):
        pos, neg = self._channels(values, where)
        helper = self._helper()
        red = helper.temporal_always if kind == "min" else helper.temporal_eventually
        return B200CumulativeRobustness(pos=red(pos), neg=red(neg))

    def _helper(self):
        h = getattr(self, "_classic_helper", None)
        if h is None:
            h = self._classic_helper = B200ClassicRobustSemantics()
        return h

    def boolean_and(self, values, *, weights=None):
        return self._minmax(values, "min", "boolean_and")

    def boolean_or(self, values, *, weights=None):
        return self._minmax(values, "max", "boolean_or")

    def temporal_always(self, values, *, weights=None):
        return self._minmax(values, "min", "temporal_always")

    def temporal_eventually(self, values, *, weights=None):
        pos, neg = self._channels(values, "temporal_eventually")
        xp, xn = self._stack(pos, "temporal_eventually"), self._stack(neg, "temporal_eventually")
        n = xp.shape[-1]
        phi = self._list_formula("or", n, None)  # Eventually[0, n-1]: the sum window under this semantics
        sp = self.robustness_trace(phi, xp.reshape(-1, n, 1))
        sn = self.robustness_trace(phi, xn.reshape(-1, n, 1))
        return B200CumulativeRobustness(pos=sp.pos[:, 0].reshape(xp.shape[:-1]), neg=sn.neg[:, 0].reshape(xn.shape[:-1]))

    def temporal_until(self, *, left_trace, right_trace, start, end, weights_left=None, weights_right=None,
                       weights_pair=(1.0, 1.0)):
        """``sum_i min(r_i, min_{k<=i} l_k)`` per channel (cumulative.py:74-103) on lists of pairs."""
        _, last = self._until_window(left_trace, right_trace, start, end)
        lp, ln = self._channels(left_trace, "temporal_until")
        rp, rn = self._channels(right_trace, "temporal_until")
        up, shape = self._until_value(self._stack(lp, "temporal_until"), self._stack(rp, "temporal_until"), start, last, None, None, (1.0, 1.0))
        un, _ = self._until_value(self._stack(ln, "temporal_until"), self._stack(rn, "temporal_until"), start, last, None, None, (1.0, 1.0))
        return B200CumulativeRobustness(pos=up.pos[:, 0].reshape(shape), neg=un.neg[:, 0].reshape(shape))


class B200AgmRobustSemantics(_B200Semantics):
    """Arithmetic-geometric mean robustness (``TorchAgmRobustSemantics``, torch/semantics.py:348-463)."""

    syntax = "agm"


class B200DgmsrSemantics(_B200Semantics):
    """D-GMSR (``TorchDgmsrSemantics``, torch/semantics.py:509-587)."""

    syntax = "dgmsr"

    def __init__(self, *, eps: float = 1e-8, p: int = 1) -> None:
        super().__init__()
        self.eps = float(eps)
        self.p = int(p)

    def _params(self) -> dict:
        return {"eps": self.eps, "p": self.p}
