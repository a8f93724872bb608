"""Formula tree -> flat post-order node list (the ``pystl_formula_t`` of the C ABI).

Accepts ``pystl_b200.formula`` nodes and the reference's ``pystl.api`` nodes alike
(duck-typed on the class name; field names per api.py:129-302).

Predicates: the fused kernels evaluate affine predicates
``bias + sum_d coeff[d] * signal[t, d]``.  ``metadata["affine"]`` is used when
present; otherwise ``Predicate.fn`` -- an opaque callable in the reference
(api.py:24,133-143) -- is probed with float64 rows and accepted if it is an
affine function of ``signal[t]`` alone.  Any other callable becomes a *leaf channel*
(the precedent is the leaf-tensor path of _dgmsr.py:281-305): the
semantics object evaluates ``fn`` for every ``(trace, t)`` ON THE DEVICE with torch
(``leaf_values``), appends the result to the signal as an extra channel and the kernels read it as
the predicate ``1.0 * channel``; autograd carries the gradient back through ``fn``.  There is no
host evaluation path: a callable torch cannot trace raises ``NotImplementedError``.

Open-ended windows beyond one tile (``open_leaf=True``): the formula kernels keep the whole trace of an
``Always / Eventually`` without an end in ONE tile, which bounds T (8192 samples at most, fewer for large
formulas).  For longer traces the node becomes a leaf channel too (``OpenLeaf``): its operand's robustness trace
is evaluated by the formula kernels, reduced by the suffix-scan kernels of ``csrc/scan.cu``
(``pystl_b200_open_window_forward / _backward``) and read back as the predicate ``1.0 * channel`` whose ``slack``
tells the lowering where the values end (api.py:218-221: the reference raises at ``t + a > T - 1``).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import capi

_SEM_IDS = {
    "classical": capi.SEM_CLASSIC,
    "smooth": capi.SEM_SMOOTH,
    "cumulative": capi.SEM_CUMULATIVE,
    "agm": capi.SEM_AGM,
    "dgmsr": capi.SEM_DGMSR,
}


NEVER = 1 << 29  # start slack of a node the reference raises on at every t (lower.cpp: kNever)
OPEN_LEAF_SEMANTICS = ("classical", "smooth", "agm", "dgmsr", "cumulative")


@dataclass(frozen=True)
class OpenLeaf:
    """``Always / Eventually`` of ``child`` over ``[t + start, T - 1]``, computed outside the formula kernel."""

    child: Any
    mode: int  # 0 always, 1 eventually
    start: int
    name: str = "open-ended window"


@dataclass(frozen=True)
class OpenUntilLeaf:
    """``W(t + start)``: the classical Until of ``left, right`` without an end, computed outside the formula kernel."""

    left: Any
    right: Any
    start: int
    name: str = "open-ended until"


def start_slack(node: Any) -> int:
    """Samples at the end of the trace where ``node`` raises (empty window); mirrors lower.cpp."""
    kind = type(node).__name__
    if kind == "Predicate":
        return pred_slack_of(node)
    if kind == "Not":
        return start_slack(node.child)
    if kind in ("And", "Or"):
        return max(start_slack(c) for c in node.children)
    kids = (node.left, node.right) if kind == "Until" else (node.child,)
    cs = max(start_slack(c) for c in kids)
    a = int(node.interval.start)
    b = node.interval.end
    if cs == 0:
        return a
    if kind == "Until" or b is None or cs >= NEVER:
        return NEVER
    return min(NEVER, cs + int(b))


def pred_slack_of(pred: Any) -> int:
    """``metadata["slack"] = s``: the predicate reads a value trace computed outside the kernel whose last ``s`` samples do
    not exist (the reference raises there); ``-1``: it raises at every t.  Goes into ``pystl_pred_t.slack``."""
    s = int((getattr(pred, "metadata", None) or {}).get("slack", 0))
    return NEVER if s < 0 or s >= NEVER else s


def wants_leaf(pred: Any) -> bool:
    """A predicate can opt out of the affine fast path: ``metadata={"affine": None}`` or
    ``metadata={"leaf": True}`` routes it to a leaf channel whatever its ``fn`` looks like."""
    meta = getattr(pred, "metadata", None) or {}
    return bool(meta.get("leaf")) or ("affine" in meta and meta["affine"] is None)


_PROBE_ROWS = 24


def affine_of(pred: Any, n_channels: int) -> Tuple[Tuple[Tuple[int, float], ...], float]:
    """``((channel, coeff), ...), bias`` of a predicate, or NotImplementedError.

    Without ``metadata["affine"]`` the callable is PROBED: a fast reject only.  A predicate that
    passes is checked again against ``fn`` on the first real signal of every new shape
    (``verify_probed``) and demoted to a leaf channel when it disagrees there -- e.g. a function of
    ``t`` beyond the probe's rows, or one that is only affine close to the origin."""
    meta = getattr(pred, "metadata", None) or {}
    if wants_leaf(pred):
        raise NotImplementedError(f"Predicate '{pred.name}' asks for the leaf path")
    if "affine" in meta:
        terms, bias = meta["affine"]
        if isinstance(terms, dict):
            terms = tuple(sorted(terms.items()))
        return tuple((int(d), float(c)) for d, c in terms), float(bias)
    fn = getattr(pred, "fn", None)
    if fn is None:  # classic.py:16-19
        raise ValueError(f"Predicate '{pred.name}' does not define `fn(signal, t)`.")
    D = int(n_channels)
    R = _PROBE_ROWS

    def call(sig: np.ndarray, t: int) -> float:
        return float(fn(sig, t))

    try:
        zero = np.zeros((R, D))
        bias = call(zero, 1)
        coeffs = np.zeros(D)
        for d in range(D):
            probe = np.zeros((R, D))
            probe[1, d] = 1.0
            coeffs[d] = call(probe, 1) - bias
        rng = np.random.default_rng(12345)
        ok = True
        for scale in (3.0, 1e3, 1e-3, 50.0):  # a function that is affine near the origin only fails the wide rows
            sig = rng.normal(size=(R, D)) * scale
            for t in (0, 1, 2, 3, 7, R // 2, R - 2, R - 1):
                want = bias + float(coeffs @ sig[t])
                got = call(sig, t)
                if not np.isfinite(got) or abs(got - want) > 1e-9 * (abs(bias) + float(np.abs(coeffs) @ np.abs(sig[t])) + 1e-300):
                    ok = False
    except Exception as exc:  # the callable needs something the probe cannot give it
        raise NotImplementedError(
            f"Predicate '{pred.name}': fn(signal, t) could not be probed ({exc!r}); "
            "give it metadata={'affine': ({channel: coeff}, bias)}"
        ) from exc
    if not ok:
        raise NotImplementedError(
            f"Predicate '{pred.name}' is not an affine function of signal[t]; the b200 backend "
            "evaluates affine predicates only (metadata={'affine': ({channel: coeff}, bias)})."
        )
    terms = tuple((d, float(c)) for d, c in enumerate(coeffs) if c != 0.0)
    return terms, float(bias)


def verify_probed(cf: "CompiledFormula", traces: np.ndarray) -> Tuple[int, ...]:
    """Check the predicates whose affine form came from probing against ``fn`` on real traces
    (``traces[n, T, D]`` float64, a few rows of the caller's batch).  Returns ``id(fn)`` of those
    that disagree anywhere in the sample: the caller recompiles with them forced onto leaf channels."""
    bad = []
    n, T, _ = traces.shape
    rng = np.random.default_rng(T)
    ts = sorted({0, 1, 2, 3, T // 3, T // 2, max(T - 2, 0), T - 1, *rng.integers(0, T, size=8).tolist()})
    for pred, terms, bias in cf.probed_preds:
        ok = True
        try:
            for b in range(n):
                sig = traces[b]
                for t in ts:
                    if t < 0 or t >= T:
                        continue
                    want = bias + sum(c * sig[t, d] for d, c in terms)
                    got = float(pred.fn(sig, t))
                    mag = abs(bias) + sum(abs(c * sig[t, d]) for d, c in terms) + 1e-300
                    if not np.isfinite(got) or abs(got - want) > 1e-9 * mag:
                        ok = False
                        break
                if not ok:
                    break
        except Exception:
            ok = False
        if not ok:
            bad.append(id(pred.fn))
    return tuple(sorted(bad))


@dataclass
class CompiledFormula:
    """Owns the ctypes arrays a ``pystl_formula_t`` points into."""

    struct: capi.Formula
    nodes: Any
    children: Any
    preds: Any
    weights: Any
    n_channels_min: int  # highest channel referenced + 1
    channels_used: Tuple[int, ...]
    pred_terms: List[Tuple[Tuple[Tuple[int, float], ...], float]]
    n_signal_channels: int = 0   # channels of the user's signal
    leaf_preds: Tuple[Any, ...] = ()  # non-affine predicates / OpenLeaf / OpenUntilLeaf, in column order
    n_leaf_columns: int = 0  # one per leaf; two (pos, neg) for an open-ended window of the cumulative semantics
    has_open: bool = False  # an Always / Eventually without an end is evaluated inside the kernel (one tile)
    signal_channels: Tuple[int, ...] = ()  # channels of the user's signal the kernels read, in kernel order (leaves follow)
    root_leaf: Optional[int] = None  # the whole formula is leaf j (an open-ended window / until): no formula kernel runs
    probed_preds: Tuple[Any, ...] = ()  # (predicate, terms, bias) recognised as affine by probing only

    @property
    def n_channels(self) -> int:
        """Channels of the tensor the kernels read (signal + leaf channels)."""
        return len(self.signal_channels) + self.n_leaf_columns

    def ptr(self):
        return C.byref(self.struct)


def compile_formula(
    formula: Any,
    semantics: str,
    n_channels: int,
    *,
    temperature: float = 1.0,
    eps: float = 1e-8,
    p: int = 1,
    force_leaf: Tuple[int, ...] = (),
    open_leaf: bool = False,
) -> CompiledFormula:
    sem_id = _SEM_IDS[semantics]
    pred_slack: Dict[int, int] = {}
    pred_pair: set = set()
    has_open = [False]
    nodes: List[capi.Node] = []
    children: List[int] = []
    weights: List[float] = []
    preds: List[Tuple[Tuple[Tuple[int, float], ...], float]] = []
    pred_index: Dict[Any, int] = {}
    memo: Dict[int, int] = {}
    leaves: List[Any] = []
    leaf_index: Dict[int, int] = {}  # id(fn) / id(node) -> first leaf COLUMN (a cumulative open-ended window has two: pos, neg)
    leaf_first: List[int] = []       # first column of leaf j
    n_cols = [0]

    def add_leaf(key_: int, leaf: Any, width: int = 1) -> int:
        leaf_index[key_] = n_cols[0]
        leaf_first.append(n_cols[0])
        leaves.append(leaf)
        n_cols[0] += width
        return leaf_index[key_]
    probed: List[Any] = []
    probed_seen: set = set()

    def add_weights(w) -> Tuple[int, int]:
        if w is None:
            return -1, 0
        w = [float(x) for x in w]
        off = len(weights)
        weights.extend(w)
        return off, len(w)

    def interval_of(node) -> Tuple[int, int]:
        iv = node.interval
        a = int(iv.start)
        b = -1 if iv.end is None else int(iv.end)
        if a < 0:
            raise ValueError("Interval start must be >= 0.")
        if b != -1 and b < a:
            raise ValueError("Interval end must be >= start.")
        return a, b

    def visit(node: Any) -> int:
        key = id(node)
        if key in memo:
            return memo[key]
        kind = type(node).__name__
        n = capi.Node()
        n.w_off, n.w_len, n.w2_off, n.w2_len = -1, 0, -1, 0
        n.wp0 = n.wp1 = 1.0
        n.pred = -1
        kids: List[int] = []
        if kind == "Predicate" and "pair" in (getattr(node, "metadata", None) or {}):
            # a cumulative value that already exists as two channels (pos, neg) of the signal: read as it stands
            if semantics != "cumulative":
                raise ValueError(f"Predicate '{node.name}': metadata['pair'] names a cumulative (pos, neg) value")
            c0, c1 = (int(c) for c in node.metadata["pair"])
            for d in (c0, c1):
                if d < 0 or d >= n_channels:
                    raise ValueError(f"Predicate '{node.name}' reads channel {d} of a {n_channels}-channel signal")
            n.kind = capi.NODE_PRED
            pkey = (((c0, 1.0), (c1, 1.0)), 0.0, "pair")
            if pkey not in pred_index:
                pred_index[pkey] = len(preds)
                preds.append(pkey[:2])
                pred_pair.add(pred_index[pkey])
            n.pred = pred_index[pkey]
        elif kind == "Predicate":
            n.kind = capi.NODE_PRED
            try:
                if id(getattr(node, "fn", None)) in force_leaf:
                    raise NotImplementedError("demoted: fn disagreed with its probed affine form on a real signal")
                terms, bias = affine_of(node, n_channels)
                meta = getattr(node, "metadata", None) or {}
                if "affine" not in meta and id(node.fn) not in probed_seen:
                    probed_seen.add(id(node.fn))
                    probed.append((node, terms, bias))
            except NotImplementedError:
                # not affine: evaluated with torch on the device into leaf channel n_channels + j
                fkey = id(node.fn)
                if fkey not in leaf_index:
                    add_leaf(fkey, node)
                terms, bias = ((n_channels + leaf_index[fkey], 1.0),), 0.0
            if len(terms) > capi.MAX_TERMS:
                raise NotImplementedError(f"Predicate '{node.name}' has more than {capi.MAX_TERMS} terms")
            for d, _ in terms:
                if d < 0 or d >= n_channels + n_cols[0]:
                    raise ValueError(f"Predicate '{node.name}' reads channel {d} of a {n_channels}-channel signal")
            sl = pred_slack_of(node)
            pkey = (terms, bias) if sl == 0 else (terms, bias, sl)
            if pkey not in pred_index:
                pred_index[pkey] = len(preds)
                preds.append(pkey[:2])
                if sl:
                    pred_slack[pred_index[pkey]] = -1 if sl >= NEVER else sl
            n.pred = pred_index[pkey]
        elif kind == "Not":
            n.kind = capi.NODE_NOT
            kids = [visit(node.child)]
        elif kind in ("And", "Or"):
            n.kind = capi.NODE_AND if kind == "And" else capi.NODE_OR
            if len(node.children) == 0:
                raise ValueError(f"{kind} requires at least one child.")
            kids = [visit(c) for c in node.children]
            n.w_off, n.w_len = add_weights(node.weights)
        elif kind in ("Always", "Eventually") and node.interval.end is None and open_leaf:
            # beyond one tile: the operand's trace goes through the suffix-scan kernel and comes back as a channel
            a, _ = interval_of(node)
            if semantics not in OPEN_LEAF_SEMANTICS:
                raise NotImplementedError(
                    f"{kind} without an interval end on a trace longer than one tile: not offered for the {semantics} semantics"
                )
            if node.weights is not None:
                raise NotImplementedError(f"{kind} without an interval end on a trace longer than one tile takes no weights")
            n.kind = capi.NODE_PRED
            if semantics == "cumulative":  # the window of a (pos, neg) trace is two columns, read back as a pair predicate
                col = add_leaf(key, OpenLeaf(node.child, 0 if kind == "Always" else 1, a), width=2)
                pkey = (((n_channels + col, 1.0), (n_channels + col + 1, 1.0)), 0.0)
            else:
                col = add_leaf(key, OpenLeaf(node.child, 0 if kind == "Always" else 1, a))
                pkey = (((n_channels + col, 1.0),), 0.0)
            pred_index[pkey + ("leaf",)] = len(preds)
            preds.append(pkey)
            n.pred = len(preds) - 1
            if semantics == "cumulative":
                pred_pair.add(n.pred)
            pred_slack[n.pred] = a if start_slack(node.child) == 0 else -1
        elif kind in ("Always", "Eventually"):
            n.kind = capi.NODE_ALWAYS if kind == "Always" else capi.NODE_EVENTUALLY
            kids = [visit(node.child)]
            n.a, n.b = interval_of(node)
            n.w_off, n.w_len = add_weights(node.weights)
            has_open[0] |= n.b < 0
        elif kind == "Until" and node.interval.end is None and open_leaf:
            # U[a, inf) = min( G[0, a] left, W(t + a) ) with W the clamp-composition suffix scan (lower.cpp does the same in-tile)
            a, _ = interval_of(node)
            if semantics != "classical":
                raise NotImplementedError(
                    f"Until without an interval end on a trace longer than one tile: classical semantics only ({semantics}: O(T^2) per trace)"
                )
            col = add_leaf(key, OpenUntilLeaf(node.left, node.right, a))
            pkey = (((n_channels + col, 1.0),), 0.0)
            pred_index[pkey] = len(preds)
            preds.append(pkey)
            pred_slack[pred_index[pkey]] = a if max(start_slack(node.left), start_slack(node.right)) == 0 else -1
            if a == 0:
                n.kind = capi.NODE_PRED
                n.pred = pred_index[pkey]
            else:
                w = capi.Node()
                w.kind, w.pred, w.n_children, w.child_off = capi.NODE_PRED, pred_index[pkey], 0, len(children)
                w.w_off, w.w_len, w.w2_off, w.w2_len, w.wp0, w.wp1 = -1, 0, -1, 0, 1.0, 1.0
                nodes.append(w)
                w_idx = len(nodes) - 1
                left_idx = visit(node.left)
                g = capi.Node()
                g.kind, g.pred, g.n_children, g.child_off, g.a, g.b = capi.NODE_ALWAYS, -1, 1, len(children), 0, a
                g.w_off, g.w_len, g.w2_off, g.w2_len, g.wp0, g.wp1 = -1, 0, -1, 0, 1.0, 1.0
                children.append(left_idx)
                nodes.append(g)
                n.kind = capi.NODE_AND
                kids = [w_idx, len(nodes) - 1]
        elif kind == "Until":
            n.kind = capi.NODE_UNTIL
            kids = [visit(node.left), visit(node.right)]
            n.a, n.b = interval_of(node)
            has_open[0] |= n.b < 0
            n.w_off, n.w_len = add_weights(node.weights_left)
            n.w2_off, n.w2_len = add_weights(node.weights_right)
            wp = tuple(float(x) for x in node.weights_pair)
            if len(wp) < 2:
                raise ValueError("weights_pair requires at least 2 entries.")
            n.wp0, n.wp1 = wp[0], wp[1]
        else:
            raise TypeError(f"Unsupported formula node type: {type(node)!r}")
        n.n_children = len(kids)
        n.child_off = len(children)
        children.extend(kids)
        nodes.append(n)
        memo[key] = len(nodes) - 1
        return memo[key]

    visit(formula)

    # With leaf channels the kernels read a tensor the host assembles anyway: signal channels no predicate reads are left out
    # of it (an open-ended window over the whole formula needs none at all).
    signal_channels = tuple(range(n_channels))
    if leaves:
        signal_channels = tuple(sorted({d for terms, _ in preds for d, _ in terms if d < n_channels}))
        remap = {d: i for i, d in enumerate(signal_channels)}
        remap.update({n_channels + j: len(signal_channels) + j for j in range(n_cols[0])})
        preds = [(tuple((remap[d], c) for d, c in terms), bias) for terms, bias in preds]
    root = nodes[-1]
    root_leaf = None
    if leaves and root.kind == capi.NODE_PRED and len(nodes) == 1:
        terms, bias = preds[root.pred]
        col = terms[0][0] - len(signal_channels)
        if col in leaf_first and isinstance(leaves[leaf_first.index(col)], (OpenLeaf, OpenUntilLeaf)):
            root_leaf = leaf_first.index(col)

    node_arr = (capi.Node * len(nodes))(*nodes)
    child_arr = (C.c_int32 * max(len(children), 1))(*children)
    weight_arr = (C.c_float * max(len(weights), 1))(*weights)
    pred_arr = (capi.Pred * max(len(preds), 1))()
    used = set()
    for i, (terms, bias) in enumerate(preds):
        pred_arr[i].n_terms = len(terms)
        for k, (d, c) in enumerate(terms):
            pred_arr[i].channel[k] = d
            pred_arr[i].coeff[k] = c
            used.add(d)
        pred_arr[i].bias = bias
        pred_arr[i].slack = pred_slack.get(i, 0)
        pred_arr[i].pair = 1 if i in pred_pair else 0
    st = capi.Formula()
    st.semantics = sem_id
    st.temperature = float(temperature)
    st.eps = float(eps)
    st.p = int(p)
    st.n_nodes = len(nodes)
    st.nodes = C.cast(node_arr, C.POINTER(capi.Node))
    st.n_children = len(children)
    st.children = C.cast(child_arr, C.POINTER(C.c_int32))
    st.n_preds = len(preds)
    st.preds = C.cast(pred_arr, C.POINTER(capi.Pred))
    st.n_weights = len(weights)
    st.weights = C.cast(weight_arr, C.POINTER(C.c_float))
    return CompiledFormula(
        struct=st,
        nodes=node_arr,
        children=child_arr,
        preds=pred_arr,
        weights=weight_arr,
        n_channels_min=(max(used) + 1) if used else 1,
        channels_used=tuple(sorted(used)),
        pred_terms=preds,
        n_signal_channels=int(n_channels),
        leaf_preds=tuple(leaves),
        n_leaf_columns=n_cols[0],
        probed_preds=tuple(probed),
        has_open=has_open[0],
        signal_channels=signal_channels,
        root_leaf=root_leaf,
    )


def leaf_column(cf: CompiledFormula, pred: Any, sig, sem=None, nan_tail: bool = False):
    """One leaf channel ``[B, T]`` of ``sig[B, T, D]`` (torch, on the device; differentiable).

    An ``OpenLeaf`` / ``OpenUntilLeaf`` channel is ``sem.robustness_trace`` of the operand(s) through ``ops.open_window`` /
    ``ops.open_until``.  ``fn(signal, t)`` is the reference's per-trace callable (api.py:24,133-143): it is mapped over
    the batch with ``torch.func.vmap`` and over ``t`` either with a second vmap (callables that only
    index with ``t``) or with a Python loop (callables that compute with ``t``)."""
    import torch

    B, T, _ = sig.shape
    if isinstance(pred, (OpenLeaf, OpenUntilLeaf)):
        from . import ops

        if sem is None:
            raise RuntimeError("an open-ended leaf needs the semantics object that compiled it")
        if isinstance(pred, OpenUntilLeaf):
            yl, yr = sem.robustness_trace(pred.left, sig), sem.robustness_trace(pred.right, sig)
            return ops.open_until(yl.contiguous(), yr.contiguous(), cf.struct.semantics, pred.start)
        y = sem.robustness_trace(pred.child, sig)  # [B, T]; nested open-ended windows recurse here
        if cf.struct.semantics == capi.SEM_CUMULATIVE:  # each channel of the pairs on its own: minimum (Always) / sum (Eventually)
            return tuple(ops.open_window(ch.contiguous(), cf.struct.semantics, pred.mode, pred.start, 1.0, 1e-8, 1, nan_tail=nan_tail)
                         for ch in (y.pos, y.neg))
        return ops.open_window(y.contiguous(), cf.struct.semantics, pred.mode, pred.start,
                               cf.struct.temperature, cf.struct.eps, cf.struct.p, nan_tail=nan_tail)
    fn = pred.fn
    val = None
    try:
        t_idx = torch.arange(T, device=sig.device)
        per_trace = torch.func.vmap(lambda s, t: torch.as_tensor(fn(s, t), dtype=s.dtype), in_dims=(None, 0))
        val = torch.func.vmap(per_trace, in_dims=(0, None))(sig, t_idx)  # [B, T]
    except Exception:
        val = None
    if val is None or tuple(val.shape) != (B, T):
        try:
            rows = [torch.func.vmap(lambda s, tt=tt: torch.as_tensor(fn(s, tt), dtype=s.dtype))(sig) for tt in range(T)]
            val = torch.stack(rows, dim=1)
        except Exception as exc:
            raise NotImplementedError(
                f"Predicate '{pred.name}': fn(signal, t) is neither affine nor traceable by torch on the device ({exc!r})"
            ) from exc
    if tuple(val.shape) != (B, T):
        raise NotImplementedError(f"Predicate '{pred.name}': fn(signal, t) must return a scalar")
    return val.to(sig.dtype)


def leaf_values(cf: CompiledFormula, sig, sem=None):
    """``sig[B, T, D]`` (torch, on the device) -> the tensor the kernels read: the signal channels the predicates use
    (``cf.signal_channels``) followed by one channel per leaf predicate (``leaf_column``)."""
    import torch

    if not cf.leaf_preds:
        return sig
    from . import ops

    cols = []
    for pred in cf.leaf_preds:
        c = leaf_column(cf, pred, sig, sem)
        cols.extend(c if isinstance(c, tuple) else (c,))
    return ops.assemble(sig, cf.signal_channels, cols)
