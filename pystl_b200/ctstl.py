"""CT-STL on the ``b200`` backend (``CtstlSemantics``, /root/reference/pystl/semantics/ctstl.py).

The reference exports the class but does not register it under ``create_semantics`` (semantics/__init__.py:47-74,
tests/test_api.py:255-257) -- neither does this backend: ``create_semantics("ctstl", backend="b200")`` raises ``KeyError``
like the reference, and the class is constructed directly, ``B200CtstlSemantics(delta=...)``.

CT-STL is the classical lattice (predicate, Not, And = min, Or = max, Always, Eventually: ctstl.py:50-76 and base.py:47-61)
plus two operators of its own, both evaluated by the kernels of ``csrc/select.cu`` on VALUE TRACES:

* ``temporal_until`` with an EXCLUSIVE left prefix (ctstl.py:78-105) -- ``pystl_b200_ctstl_until_forward / _backward``;
* ``temporal_cumulative`` (ctstl.py:107-127), the robustness of the cumulative-time operator ``C^tau_I``: the k-th largest
  value of the window, ``k = ceil(tau / delta)`` -- ``pystl_b200_ctstl_cumulative_forward / _backward``.  The reference has no
  formula node for it (its helper takes the window's values); ``CumulativeTime(child, interval, tau)`` is that node here.

A formula is cut at its Until / CumulativeTime nodes: every maximal subtree without them is ONE classical fused kernel
launch, the two CT-STL operators run on the value traces of their operands, and what lies above them reads those traces as
channels of a synthetic signal through the same classical kernels.  Everything is a differentiable torch op on the device.
"""

from __future__ import annotations

from math import ceil
from typing import Any, Optional, Sequence

import torch

from .formula import Always, And, Eventually, Formula, Interval, IntervalLike, Not, Or, affine_predicate, as_interval
from .semantics import B200ClassicRobustSemantics, _device_of


def tau_to_k(tau: float, delta: float) -> int:
    """ctstl.py:28-35, same checks and messages."""
    if delta <= 0:
        raise ValueError("delta must be > 0.")
    if tau <= 0:
        raise ValueError("tau must be > 0.")
    return int(ceil(tau / delta))


class CumulativeTime(Formula):
    """``C^tau_I child``: the child holds for at least ``tau`` time units inside the window; robustness = k-th largest of the
    child's values over ``[t + a, min(t + b, T - 1)]``, ``k = ceil(tau / delta)`` (ctstl.py:107-127)."""

    def __init__(self, child: Formula, interval: IntervalLike = Interval(), *, tau: float, delta: Optional[float] = None) -> None:
        self.child = child
        self.interval = as_interval(interval)
        self.tau = float(tau)
        self.delta = None if delta is None else float(delta)

    def __repr__(self) -> str:
        return f"CumulativeTime({self.child!r}, {self.interval}, tau={self.tau})"


def _has_ct_node(node: Any) -> bool:
    kind = type(node).__name__
    if kind in ("Until", "CumulativeTime"):
        return True
    if kind == "Predicate":
        return False
    if kind in ("And", "Or"):
        return any(_has_ct_node(c) for c in node.children)
    return _has_ct_node(node.child)


def _slack(node: Any) -> int:
    """Samples at the end of a trace where ``node`` raises in the reference (``compiler.start_slack`` with the CT-STL nodes):
    what a parent evaluated by the classical kernels must know about a value trace it reads as a channel.  An Until follows the
    classical rule (api.py:284-302).  ``C^tau_I`` never raises by itself (an empty window is -inf, ctstl.py:118-119); over an
    operand that raises near the horizon it raises wherever its window reaches that tail -- reported to a parent as the
    tail ``t > T - 1 - (slack + b)``, which for ``a > 0`` also covers the last ``a`` outputs (empty windows, -inf in the reference)."""
    from .compiler import NEVER, start_slack

    kind = type(node).__name__
    if not _has_ct_node(node):
        return start_slack(node)
    if kind == "Not":
        return _slack(node.child)
    if kind in ("And", "Or"):
        return max(_slack(c) for c in node.children)
    kids = (node.left, node.right) if kind == "Until" else (node.child,)
    cs = max(_slack(c) for c in kids)
    a, b = int(node.interval.start), node.interval.end
    if cs == 0:
        return 0 if kind == "CumulativeTime" else a
    if kind == "Until" or b is None or cs >= NEVER:
        return NEVER
    return min(NEVER, cs + int(b))


def _channel(i: int, slack: int):
    """The value trace in channel ``i`` of a synthetic signal, read as it stands; its last ``slack`` samples do not exist."""
    from .compiler import NEVER

    pred = affine_predicate(f"v{i}", {i: 1.0})
    if slack:
        pred.metadata["slack"] = -1 if slack >= NEVER else int(slack)
    return pred


class B200CtstlSemantics(B200ClassicRobustSemantics):
    """``CtstlSemantics(delta=...)`` on the device: operator interface (lists of values, ctstl.py:50-127) and whole traces."""

    syntax = "classical"  # the kernels every shared operator runs on

    def __init__(self, *, delta: float = 1.0) -> None:
        super().__init__()
        if delta <= 0:  # ctstl.py:47-49
            raise ValueError("delta must be > 0.")
        self.delta = float(delta)
        self._ct_helpers: dict = {}

    # ---- value traces ---------------------------------------------------------------------------------------------
    def cumulative_time_trace(self, values: Any, interval: IntervalLike = Interval(), *, tau: float, delta: Optional[float] = None) -> torch.Tensor:
        """``C^tau_I`` of a value trace ``[B, T]`` (or ``[T]``) at every t; -inf where the clipped window holds fewer than k samples."""
        from . import ops

        iv = as_interval(interval)
        k = tau_to_k(tau, self.delta if delta is None else float(delta))
        y = torch.as_tensor(values)
        home = y.device
        y = y.to(device=_device_of(y), dtype=torch.float32)
        squeeze = y.dim() == 1
        out = ops.ctstl_cumulative(y.unsqueeze(0) if squeeze else y, iv.start, iv.end, k)
        out = out[0] if squeeze else out
        return out.to(home) if home != out.device else out

    def until_trace(self, left: Any, right: Any, interval: IntervalLike = Interval()) -> torch.Tensor:
        """CT-STL Until of two value traces ``[B, T]`` (or ``[T]``) at every t; NaN where the window is empty."""
        from . import ops

        iv = as_interval(interval)
        l, r = torch.as_tensor(left), torch.as_tensor(right)
        home = l.device
        dev = _device_of(l if l.is_cuda else r)
        l, r = l.to(device=dev, dtype=torch.float32), r.to(device=dev, dtype=torch.float32)
        squeeze = l.dim() == 1
        out = ops.ctstl_until(l.unsqueeze(0) if squeeze else l, r.unsqueeze(0) if squeeze else r, iv.start, iv.end)
        out = out[0] if squeeze else out
        return out.to(home) if home != out.device else out

    # ---- whole formulas -------------------------------------------------------------------------------------------
    def _helper_formula(self, key: tuple, make) -> Any:
        phi = self._ct_helpers.get(key)
        if phi is None:
            if len(self._ct_helpers) > 128:
                self._ct_helpers.clear()
            phi = self._ct_helpers[key] = make()
        return phi

    def _trace(self, node: Any, sig: torch.Tensor) -> torch.Tensor:
        """``[B, T]`` robustness of ``node`` on ``sig [B, T, D]`` (NaN where the reference raises)."""
        if not _has_ct_node(node):
            return B200ClassicRobustSemantics.robustness_trace(self, node, sig)
        kind = type(node).__name__
        if kind == "Not":
            return -self._trace(node.child, sig)
        if kind in ("And", "Or"):
            cols = torch.stack([self._trace(c, sig) for c in node.children], dim=-1)  # [B, T, n]
            sl = tuple(_slack(c) for c in node.children)
            phi = self._helper_formula((kind, sl), lambda: (And if kind == "And" else Or)(*[_channel(i, s) for i, s in enumerate(sl)]))
            return B200ClassicRobustSemantics.robustness_trace(self, phi, cols)
        if kind in ("Always", "Eventually"):
            y = self._trace(node.child, sig).unsqueeze(-1)
            iv, sl = node.interval, _slack(node.child)
            phi = self._helper_formula((kind, iv.start, iv.end, sl), lambda: (Always if kind == "Always" else Eventually)(_channel(0, sl), iv))
            return B200ClassicRobustSemantics.robustness_trace(self, phi, y)
        if kind == "Until":
            l, r = self._trace(node.left, sig), self._trace(node.right, sig)
            out = self.until_trace(l, r, node.interval)
            # api.py:288-293 evaluates both operands on every step from t to the horizon: an operand that raises anywhere
            # (its NaN tail) makes the Until raise at every t
            bad = torch.isnan(l[:, -1]) | torch.isnan(r[:, -1])
            return torch.where(bad.unsqueeze(-1), torch.full_like(out, float("nan")), out)
        if kind == "CumulativeTime":
            return self.cumulative_time_trace(self._trace(node.child, sig), node.interval, tau=node.tau, delta=node.delta)
        raise TypeError(f"Unsupported formula node type: {type(node)!r}")

    def robustness_trace(self, formula: Any, signal: Any) -> Any:
        if not _has_ct_node(formula):
            return super().robustness_trace(formula, signal)
        sig, squeeze, home = self._prepare(signal)
        if sig.shape[1] == 0:
            raise ValueError("signal must have at least one sample.")
        out = self._trace(formula, sig)
        if home != out.device:
            out = out.to(home)
        return out[0] if squeeze else out

    def evaluate_formula(self, formula: Any, signal: Any, t: int = 0) -> Any:
        if not _has_ct_node(formula):
            return super().evaluate_formula(formula, signal, t)
        sig = torch.as_tensor(signal)
        T = sig.shape[-2]
        if t < 0 or t >= T:
            raise IndexError(f"t={t} is out of bounds for horizon={T}.")
        out = self.robustness_trace(formula, signal)[..., t]
        if bool(torch.isnan(out).any()):
            raise ValueError(f"temporal window is empty at t={t} for a horizon of {T} samples.")
        return out

    def _mean_parts(self, formula: Any, sig: torch.Tensor):
        """(sum of the outputs that exist, their number, the trace) -- all on the device, nothing synchronises."""
        out = self._trace(formula, sig)
        ok = ~torch.isnan(out)
        total = torch.where(ok, out, torch.zeros_like(out)).sum(dim=1).sum(dtype=torch.float64)
        return total, ok.sum().to(torch.float64), out

    def mean_robustness(self, formula: Any, signal: Any, *, group=None):
        """Mean of the outputs the reference does not raise on (an output of -inf -- a ``C^tau_I`` window shorter than k -- makes
        it -inf); with ``torch.distributed`` initialised over the whole sharded batch (one 16-byte all-reduce)."""
        if not _has_ct_node(formula):
            return super().mean_robustness(formula, signal, group=group)
        from .distributed import all_reduce_mean

        sig, _, _ = self._prepare(signal)
        with torch.no_grad():
            total, count, _ = self._mean_parts(formula, sig)
        return all_reduce_mean(torch.stack((total, count)), group=group)

    def mean_robustness_and_grad(self, formula: Any, signal: Any, *, group=None):
        """``(mean robustness, d mean / d signal)``: the trace of ``robustness_trace`` and autograd through its kernels (the formula
        is cut at its CT-STL nodes, so there is no single fused kernel); the count of valid outputs is reduced over ``group``
        before the backward pass, the gradient itself needs no collective."""
        if not _has_ct_node(formula):
            return super().mean_robustness_and_grad(formula, signal, group=group)
        import torch.distributed as dist

        from .distributed import all_reduce_mean

        sig, squeeze, home = self._prepare(signal)
        base = sig.detach().requires_grad_(True)
        with torch.enable_grad():
            total, count, _ = self._mean_parts(formula, base)
        n_total = count.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(n_total, op=dist.ReduceOp.SUM, group=group)
        # (-inf outputs carry no gradient -- their selection index is -1 -- and make the mean -inf, which is what it is)
        (grad,) = torch.autograd.grad(total, base, torch.where(n_total > 0, 1.0 / n_total, torch.zeros_like(n_total)))
        mean = all_reduce_mean(torch.stack((total.detach(), count)), group=group)
        if home != grad.device:
            grad = grad.to(home)
        return mean, (grad[0] if squeeze else grad)

    # ---- operator interface (lists of values at one t) ----------------------------------------------------------------
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
        """ctstl.py:78-105: the lists are the operands' values from t on; weights are ignored, like there."""
        self._until_window(left_trace, right_trace, start, end)  # same checks, same messages
        l = self._stack(left_trace, "temporal_until")  # [..., n]
        r = self._stack(right_trace, "temporal_until")
        l, r = torch.broadcast_tensors(l, r)
        n = l.shape[-1]
        out = self.until_trace(l.reshape(-1, n), r.reshape(-1, n), Interval(int(start), None if end is None else int(end)))
        return out[:, 0].reshape(l.shape[:-1])

    def temporal_cumulative(self, values: Sequence[Any], *, tau: float, delta: Optional[float] = None) -> Any:
        """ctstl.py:107-127: k-th largest of the values, -inf for an empty list or one shorter than k."""
        if len(values) == 0:
            return torch.tensor(float("-inf"))
        x = self._stack(values, "temporal_cumulative")  # [..., n]
        n = x.shape[-1]
        out = self.cumulative_time_trace(x.reshape(-1, n), Interval(0, n - 1), tau=tau, delta=delta)
        return out[:, 0].reshape(x.shape[:-1])
