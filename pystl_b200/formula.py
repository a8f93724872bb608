"""Formula syntax tree accepted by the B200 backend.

The backend consumes *either* these classes or the reference's own
``pystl.api`` nodes (``/root/reference/pystl/api.py:129-302``) -- the compiler
in :mod:`pystl_b200.compiler` duck-types on the class name and the field names,
which are kept identical to the reference:

========== =======================================================
node       fields (same names as ``pystl/api.py``)
========== =======================================================
Predicate  ``name, fn, metadata, grad``                (api.py:137-140)
Not        ``child``                                   (api.py:150)
And / Or   ``children, weights``                       (api.py:160-161,180-181)
Always /   ``child, interval, weights``                (api.py:200-202,230-232)
Eventually
Until      ``left, right, interval, weights_left,      (api.py:260-265)
           weights_right, weights_pair``
Interval   ``start, end`` (closed, ``end=None`` = open)  (api.py:29-40)
========== =======================================================

The module exists because the reference package is not installed on the GPU
box; when ``pystl`` *is* importable its nodes work unchanged.  On top of the
reference surface a predicate may carry a structured affine description
(``affine_predicate``), which is what the fused kernels evaluate; an opaque
``fn`` is probed once and recognised as affine when it is (see
``compiler.affine_of``).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Optional, Sequence, Union

IntervalLike = Union["Interval", tuple, list]


@dataclass(frozen=True)
class Interval:
    """Closed integer interval ``[start, end]``; ``end=None`` means "to the horizon"."""

    start: int = 0
    end: Optional[int] = None

    def __post_init__(self) -> None:
        # same two checks, same exception type as api.py:36-40
        if self.start < 0:
            raise ValueError("Interval start must be >= 0.")
        if self.end is not None and self.end < self.start:
            raise ValueError("Interval end must be >= start.")


def as_interval(interval: IntervalLike) -> Interval:
    if hasattr(interval, "start") and hasattr(interval, "end"):
        return Interval(int(interval.start), None if interval.end is None else int(interval.end))
    lo, hi = interval
    return Interval(int(lo), None if hi is None else int(hi))


class Formula:
    """Base node. Builder sugar mirrors api.py:85-126 (``& | ~ .always .eventually .until``)."""

    def __and__(self, other: "Formula") -> "And":
        return And(self, other)

    def __or__(self, other: "Formula") -> "Or":
        return Or(self, other)

    def __invert__(self) -> "Not":
        return Not(self)

    def always(self, interval: IntervalLike = Interval(), *, weights=None) -> "Always":
        return Always(self, interval, weights=weights)

    def eventually(self, interval: IntervalLike = Interval(), *, weights=None) -> "Eventually":
        return Eventually(self, interval, weights=weights)

    def until(
        self,
        other: "Formula",
        interval: IntervalLike = Interval(),
        *,
        weights_left=None,
        weights_right=None,
        weights_pair: Sequence[float] = (1.0, 1.0),
    ) -> "Until":
        return Until(
            self,
            other,
            interval,
            weights_left=weights_left,
            weights_right=weights_right,
            weights_pair=weights_pair,
        )

    # Reference-compatible entry points (api.py:57-78).  They dispatch to the
    # semantics object, which is where the B200 backend does whole-formula work.
    def evaluate(self, signal: Any, semantics: Any, t: int = 0) -> Any:
        whole = getattr(semantics, "evaluate_formula", None)
        if whole is not None:
            return whole(self, signal, t)
        return _evaluate_recursive(self, signal, semantics, t)

    def evaluate_with_grad(self, signal: Any, semantics: Any, t: int = 0, **kwargs: Any):
        evaluator = getattr(semantics, "evaluate_with_grad", None)
        if evaluator is None:
            raise NotImplementedError(f"{type(semantics).__name__} does not support gradients.")
        return evaluator(self, signal, t=t, **kwargs)


@dataclass(frozen=True, eq=False)
class Predicate(Formula):
    name: str
    fn: Optional[Callable[[Any, int], Any]] = None
    metadata: Optional[dict] = None
    grad: Optional[Callable[[Any, int], Any]] = None


@dataclass(frozen=True, eq=False)
class Not(Formula):
    child: Formula


class _Nary(Formula):
    children: tuple
    weights: Optional[Sequence[float]]

    def __init__(self, *children: Formula, weights: Optional[Sequence[float]] = None) -> None:
        if len(children) == 0:
            raise ValueError(f"{type(self).__name__} requires at least one child.")
        self.children = tuple(children)
        self.weights = None if weights is None else tuple(float(w) for w in weights)

    def __repr__(self) -> str:
        return f"{type(self).__name__}({', '.join(map(repr, self.children))}, weights={self.weights})"


class And(_Nary):
    pass


class Or(_Nary):
    pass


class _Unary(Formula):
    def __init__(self, child: Formula, interval: IntervalLike = Interval(), *, weights=None) -> None:
        self.child = child
        self.interval = as_interval(interval)
        self.weights = None if weights is None else tuple(float(w) for w in weights)

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self.child!r}, {self.interval}, weights={self.weights})"


class Always(_Unary):
    pass


class Eventually(_Unary):
    pass


class Until(Formula):
    def __init__(
        self,
        left: Formula,
        right: Formula,
        interval: IntervalLike = Interval(),
        *,
        weights_left=None,
        weights_right=None,
        weights_pair: Sequence[float] = (1.0, 1.0),
    ) -> None:
        self.left = left
        self.right = right
        self.interval = as_interval(interval)
        self.weights_left = None if weights_left is None else tuple(float(w) for w in weights_left)
        self.weights_right = None if weights_right is None else tuple(float(w) for w in weights_right)
        self.weights_pair = tuple(float(w) for w in weights_pair)

    def __repr__(self) -> str:
        return f"Until({self.left!r}, {self.right!r}, {self.interval})"


def affine_predicate(
    name: str,
    coeffs: Union[dict, Sequence[float]],
    bias: float = 0.0,
) -> Predicate:
    """Predicate ``rho(t) = bias + sum_d coeffs[d] * signal[t, d]``.

    ``coeffs`` is either a dense sequence over channels or ``{channel: coeff}``.
    The node carries both a reference-style ``fn(signal, t)`` (so it can be fed
    to the reference's numpy/torch backends unchanged) and
    ``metadata["affine"]``, the structured form the fused kernels use.  This is
    the convention SURVEY.md section 7 proposes; ``metadata`` is the reference's own
    free-form dict (api.py:139), already consulted for ``"grad"`` at
    semantics/dgmsr.py:174-177.
    """

    if isinstance(coeffs, dict):
        terms = tuple(sorted((int(d), float(c)) for d, c in coeffs.items() if float(c) != 0.0))
    else:
        terms = tuple((d, float(c)) for d, c in enumerate(coeffs) if float(c) != 0.0)
    bias = float(bias)

    def fn(signal: Any, t: int) -> Any:
        acc = None
        for d, c in terms:
            term = signal[t, ..., d] * c if c != 1.0 else signal[t, ..., d]
            acc = term if acc is None else acc + term
        if acc is None:
            return bias + 0.0 * signal[t, ..., 0]
        return acc + bias if bias != 0.0 else acc

    def grad(signal: Any, t: int):
        import numpy as np

        g = np.zeros(signal.shape[-1], dtype=float)
        for d, c in terms:
            g[d] = c
        return g

    return Predicate(name, fn=fn, metadata={"affine": (terms, bias)}, grad=grad)


def gt(channel: int, threshold: float, name: Optional[str] = None) -> Predicate:
    """``signal[channel] > threshold``  ->  rho = x - c."""
    return affine_predicate(name or f"s{channel}>{threshold:g}", {channel: 1.0}, -float(threshold))


def lt(channel: int, threshold: float, name: Optional[str] = None) -> Predicate:
    """``signal[channel] < threshold``  ->  rho = c - x."""
    return affine_predicate(name or f"s{channel}<{threshold:g}", {channel: -1.0}, float(threshold))


# --- reference-style recursive driver (api.py:142-302), kept for semantics objects
# --- that only implement the seven-method operator interface.


def window_indices(t: int, horizon: int, interval: Interval) -> range:
    lo = t + interval.start
    hi = horizon - 1 if interval.end is None else min(t + interval.end, horizon - 1)
    return range(lo, hi + 1)


def _evaluate_recursive(node: Any, signal: Any, sem: Any, t: int) -> Any:
    kind = type(node).__name__
    if kind == "Predicate":
        return sem.predicate(node, signal, t)
    if kind == "Not":
        return sem.boolean_not(_evaluate_recursive(node.child, signal, sem, t))
    if kind in ("And", "Or"):
        vals = [_evaluate_recursive(c, signal, sem, t) for c in node.children]
        op = sem.boolean_and if kind == "And" else sem.boolean_or
        return op(vals, weights=node.weights)
    horizon = signal.shape[0]
    if kind in ("Always", "Eventually"):
        idx = window_indices(t, horizon, node.interval)
        if len(idx) == 0:
            raise ValueError(f"{kind.upper()} window is empty at t={t} for interval={node.interval}.")
        vals = [_evaluate_recursive(node.child, signal, sem, tau) for tau in idx]
        op = sem.temporal_always if kind == "Always" else sem.temporal_eventually
        return op(vals, weights=node.weights)
    if kind == "CumulativeTime":  # (pystl_b200.ctstl: the node of CtstlSemantics.temporal_cumulative, ctstl.py:107-127)
        vals = [_evaluate_recursive(node.child, signal, sem, tau) for tau in window_indices(t, horizon, node.interval)]
        return sem.temporal_cumulative(vals, tau=node.tau, delta=node.delta)
    if kind == "Until":
        if t >= horizon:
            raise IndexError(f"t={t} is out of bounds for horizon={horizon}.")
        left = [_evaluate_recursive(node.left, signal, sem, tau) for tau in range(t, horizon)]
        right = [_evaluate_recursive(node.right, signal, sem, tau) for tau in range(t, horizon)]
        return sem.temporal_until(
            left_trace=left,
            right_trace=right,
            start=node.interval.start,
            end=node.interval.end,
            weights_left=node.weights_left,
            weights_right=node.weights_right,
            weights_pair=node.weights_pair,
        )
    raise TypeError(f"Unsupported formula node type: {type(node)!r}")


# --------------------------------------------------------------------------- specs
# A formula "spec" is a nested tuple, independent of any class hierarchy, so the
# same formula can be instantiated with these classes or with the reference's
# (``from_spec(spec, ns=pystl)``) -- that is how the golden vectors are made.
#
#   ("pred", {channel: coeff, ...}, bias)            affine predicate
#   ("not", child)
#   ("and", [children...], weights|None) / ("or", ...)
#   ("always", (a, b|None), child, weights|None) / ("eventually", ...)
#   ("until", (a, b|None), left, right, weights_left|None, weights_right|None, weights_pair|None)


def from_spec(spec: Any, ns: Any = None) -> Any:
    """Build a formula from a nested-tuple spec.  ``ns`` is any namespace exposing
    ``Predicate, Not, And, Or, Always, Eventually, Until`` (default: this module)."""
    import sys

    mod = sys.modules[__name__] if ns is None else ns
    op = spec[0]
    if op == "pred":
        _, coeffs, bias = spec
        p = affine_predicate(_pred_name(coeffs, bias), coeffs, bias)
        if ns is None:
            return p
        return mod.Predicate(p.name, fn=p.fn, metadata=p.metadata, grad=p.grad)
    if op == "not":
        return mod.Not(from_spec(spec[1], ns))
    if op in ("and", "or"):
        kids = [from_spec(c, ns) for c in spec[1]]
        weights = spec[2] if len(spec) > 2 else None
        return (mod.And if op == "and" else mod.Or)(*kids, weights=weights)
    if op in ("always", "eventually"):
        (a, b), child = spec[1], from_spec(spec[2], ns)
        weights = spec[3] if len(spec) > 3 else None
        cls = mod.Always if op == "always" else mod.Eventually
        return cls(child, (a, b), weights=weights)
    if op == "until":
        (a, b) = spec[1]
        left, right = from_spec(spec[2], ns), from_spec(spec[3], ns)
        wl = spec[4] if len(spec) > 4 else None
        wr = spec[5] if len(spec) > 5 else None
        wp = spec[6] if len(spec) > 6 and spec[6] is not None else (1.0, 1.0)
        return mod.Until(left, right, (a, b), weights_left=wl, weights_right=wr, weights_pair=wp)
    raise ValueError(f"unknown spec op {op!r}")


def _pred_name(coeffs: dict, bias: float) -> str:
    parts = [f"{c:+g}*s{d}" for d, c in sorted(coeffs.items())]
    return "".join(parts) + f"{bias:+g}"


def _gt(d: int, c: float):
    return ("pred", {d: 1.0}, -float(c))


def _lt(d: int, c: float):
    return ("pred", {d: -1.0}, float(c))


# The five BASELINE.json configurations; channel x=0, y=1.
BASELINE_SPECS = {
    # G[0,10](x>0) & F[0,5](y<1)
    "cfg1": ("and", [("always", (0, 10), _gt(0, 0.0)), ("eventually", (0, 5), _lt(1, 1.0))]),
    # G[0,50](x>0) | F[0,20](y<0.5)
    "cfg2": ("or", [("always", (0, 50), _gt(0, 0.0)), ("eventually", (0, 20), _lt(1, 0.5))]),
    # G[0,100]( F[0,10](x>0) & !(y>2) )
    "cfg3": ("always", (0, 100), ("and", [("eventually", (0, 10), _gt(0, 0.0)), ("not", _gt(1, 2.0))])),
    # G[0,200]( (x>0) U[0,64] (y>1) )
    "cfg4": ("always", (0, 200), ("until", (0, 64), _gt(0, 0.0), _gt(1, 1.0))),
    # (G[0,32]Pa & F[0,16]Pb) | !(Pc U[0,24] Pd), 8 channels, 10 nodes
    "cfg5": (
        "or",
        [
            (
                "and",
                [
                    ("always", (0, 32), ("pred", {0: 1.0, 4: -1.0}, 0.0)),
                    ("eventually", (0, 16), ("pred", {1: -1.0, 5: -0.5}, 1.0)),
                ],
            ),
            ("not", ("until", (0, 24), ("pred", {2: 1.0, 6: 1.0}, 0.5), ("pred", {3: 1.0, 7: -1.0}, -0.5))),
        ],
    ),
}

BASELINE_SHAPES = {  # (B, T, D)
    "cfg1": (1024, 1000, 2),
    "cfg2": (65536, 4096, 2),
    "cfg3": (32768, 8192, 2),
    "cfg4": (16384, 16384, 4),
    "cfg5": (1 << 20, 2048, 8),
}
