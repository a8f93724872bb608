"""ORACLE (test infrastructure, never shipped, never on the product path).

Full-trace, batched, float64 restatement of the reference's semantics written
with torch CPU ops so that ``autograd`` yields the reference's gradients.  Where
``oracle/scalar.py`` mirrors the reference's per-``(trace, t)`` recursion, this
module evaluates every ``t`` of every trace at once (what the B200 kernels
produce), using the same reductions as ``pystl/torch/semantics.py`` so that the
gradient routing is the reference's:

* Classic  -- ``torch.min/max(dim)`` (first extremal index wins), torch/semantics.py:101-146
* Smooth   -- ``torch.logsumexp``, torch/semantics.py:155-215
* Cumulative -- per-channel min / max / sum, torch/semantics.py:229-345
* AGM      -- both branches + ``torch.where``, ``clamp(min=1e-12)`` in the log, torch/semantics.py:71-84,362-463
* D-GMSR   -- ``_gmsr_and_torch`` / ``_gmsr_or_torch``, torch/semantics.py:466-587

Windows clip at the horizon exactly like ``api.py:49-54``; a ``t`` whose window
is empty (the reference raises ``ValueError``) is returned as NaN.

Parity status: PINNED -- values and gradients are checked against the
unmodified reference in ``tests/test_oracle_golden.py`` (golden vectors) and,
when ``/root/reference`` is present, live in ``tests/test_oracle_vs_reference.py``.
"""

from __future__ import annotations

from typing import Any, Optional

import torch

F64 = torch.float64


def _w(weights, n: int, like: torch.Tensor) -> torch.Tensor:
    if weights is None:
        return torch.ones(n, dtype=F64)
    w = torch.as_tensor(list(weights), dtype=F64).reshape(-1)
    if w.numel() < n:
        raise ValueError(f"weights need at least {n} entries")
    return w[:n]


def _check_agm_weights(w: torch.Tensor) -> None:
    if bool((w < 0).any()):
        raise ValueError("weights must be >= 0")
    if float(w.sum()) <= 0.0:
        raise ValueError("weights need a positive entry")


class _Dense:
    """Reductions take ``x[..., n]`` with a boolean ``mask[..., n]`` of live entries and
    per-entry weights ``w[n]``; they reduce the last axis."""

    channels = 1

    def leaf(self, v):
        return v

    def neg(self, v):
        return -v

    def conj(self, x, mask, w):
        raise NotImplementedError

    def disj(self, x, mask, w):
        raise NotImplementedError

    def always(self, x, mask, w):
        return self.conj(x, mask, w)

    def eventually(self, x, mask, w):
        return self.disj(x, mask, w)

    # Until: lw, rw are [B, T, n] gathers of l[t+k], r[t+k]; mask likewise; cand[n] marks i >= a.
    def until(self, lw, rw, mask, cand, wl, wr, wp):
        raise NotImplementedError


class DClassic(_Dense):
    def conj(self, x, mask, w):
        return torch.min(torch.where(mask, x, torch.full_like(x, float("inf"))), dim=-1).values

    def disj(self, x, mask, w):
        return torch.max(torch.where(mask, x, torch.full_like(x, float("-inf"))), dim=-1).values

    def until(self, lw, rw, mask, cand, wl, wr, wp):
        inf = torch.full_like(lw, float("inf"))
        pre = torch.cummin(torch.where(mask, lw, inf), dim=-1).values
        c = torch.minimum(pre, rw)
        c = torch.where(mask & cand, c, torch.full_like(c, float("-inf")))
        return torch.max(c, dim=-1).values


class DSmooth(_Dense):
    def __init__(self, temperature: float = 1.0):
        if temperature <= 0.0:
            raise ValueError("temperature must be > 0.")
        self.tau = float(temperature)

    def _lse(self, z, mask):
        return torch.logsumexp(torch.where(mask, z, torch.full_like(z, float("-inf"))), dim=-1)

    def conj(self, x, mask, w):
        return -self.tau * self._lse(-x / self.tau, mask)

    def disj(self, x, mask, w):
        return self.tau * self._lse(x / self.tau, mask)

    def until(self, lw, rw, mask, cand, wl, wr, wp):
        tau = self.tau
        z = torch.where(mask, -lw / tau, torch.full_like(lw, float("-inf")))
        pre = torch.logcumsumexp(z, dim=-1)  # = -softmin(l[0..i]) / tau
        c = -tau * torch.logaddexp(pre, -rw / tau)  # softmin(prefix, r_i)
        return tau * self._lse(c / tau, mask & cand)


class DCumulative(_Dense):
    channels = 2

    def leaf(self, v):
        z = torch.zeros_like(v)
        return torch.stack((torch.maximum(z, v), torch.minimum(z, v)), dim=0)

    def neg(self, v):
        return torch.stack((-v[1], -v[0]), dim=0)

    def conj(self, x, mask, w):
        return torch.min(torch.where(mask, x, torch.full_like(x, float("inf"))), dim=-1).values

    def disj(self, x, mask, w):
        return torch.max(torch.where(mask, x, torch.full_like(x, float("-inf"))), dim=-1).values

    def eventually(self, x, mask, w):
        return torch.where(mask, x, torch.zeros_like(x)).sum(dim=-1)

    def until(self, lw, rw, mask, cand, wl, wr, wp):
        inf = torch.full_like(lw, float("inf"))
        pre = torch.cummin(torch.where(mask, lw, inf), dim=-1).values
        c = torch.minimum(rw, pre)
        return torch.where(mask & cand, c, torch.zeros_like(c)).sum(dim=-1)


class DAgm(_Dense):
    @staticmethod
    def _and(x, mask, w):
        wm = torch.where(mask, w.expand_as(x), torch.zeros_like(x))
        sw = wm.sum(dim=-1)
        xs = torch.where(mask, x, torch.ones_like(x))  # dead entries: positive, weight 0
        neg_part = (wm * torch.minimum(xs, torch.zeros_like(xs))).sum(dim=-1) / sw
        logs = torch.log(torch.clamp(1.0 + xs, min=1e-12))
        pos_part = torch.exp((wm * logs).sum(dim=-1) / sw) - 1.0
        return torch.where((xs > 0.0).all(dim=-1), pos_part, neg_part)

    @staticmethod
    def _or(x, mask, w):
        wm = torch.where(mask, w.expand_as(x), torch.zeros_like(x))
        sw = wm.sum(dim=-1)
        xs = torch.where(mask, x, -torch.ones_like(x))  # dead entries: negative, weight 0
        pos_part = (wm * torch.maximum(xs, torch.zeros_like(xs))).sum(dim=-1) / sw
        logs = torch.log(torch.clamp(1.0 - xs, min=1e-12))
        neg_part = 1.0 - torch.exp((wm * logs).sum(dim=-1) / sw)
        return torch.where((xs > 0.0).any(dim=-1), pos_part, neg_part)

    def conj(self, x, mask, w):
        _check_agm_weights(w)
        return self._and(x, mask, w)

    def disj(self, x, mask, w):
        _check_agm_weights(w)
        return self._or(x, mask, w)

    def until(self, lw, rw, mask, cand, wl, wr, wp):
        for ww in (wl, wr, wp):
            _check_agm_weights(ww)
        n = lw.shape[-1]
        wlm = torch.where(mask, wl.expand_as(lw), torch.zeros_like(lw))
        ls = torch.where(mask, lw, torch.ones_like(lw))
        csw = wlm.cumsum(dim=-1).clamp(min=1e-300)
        neg_part = (wlm * torch.minimum(ls, torch.zeros_like(ls))).cumsum(dim=-1) / csw
        pos_part = torch.exp((wlm * torch.log(torch.clamp(1.0 + ls, min=1e-12))).cumsum(dim=-1) / csw) - 1.0
        all_pos = torch.cummin((ls > 0.0).to(F64), dim=-1).values > 0.5
        pre = torch.where(all_pos, pos_part, neg_part)  # AND(l[0..i])
        pair = torch.stack((pre, torch.where(mask, rw, torch.ones_like(rw))), dim=-1)
        c = self._and(pair, torch.ones_like(pair, dtype=torch.bool), wp)  # [B,T,n]
        # candidate weights are indexed from the first candidate (agm.py:117-119)
        first = int(cand.to(torch.int64).argmax())
        wr_full = torch.zeros(n, dtype=F64)
        m = min(n - first, wr.numel())
        wr_full[first : first + m] = wr[:m]
        return self._or(c, mask & cand, wr_full)


class DDgmsr(_Dense):
    def __init__(self, eps: float = 1e-8, p: int = 1):
        self.eps = float(eps)
        self.p = int(p)

    def _and(self, x, mask, w):
        """torch/semantics.py:466-496 with dead entries given weight 0."""
        eps, p = self.eps, self.p
        eps_t = torch.as_tensor(eps, dtype=F64)
        wm = torch.where(mask, w.expand_as(x), torch.zeros_like(x))
        sw = wm.sum(dim=-1)
        xs = torch.where(mask, x, torch.ones_like(x))
        neg = xs <= 0.0
        sums = torch.where(neg, wm * xs ** (2 * p), torch.zeros_like(xs)).sum(dim=-1)
        mp = torch.pow(torch.pow(eps_t, p) + sums / sw, 1.0 / p)
        h_neg = torch.pow(eps_t, 0.5) - torch.pow(mp, 0.5)
        base = torch.where(neg, torch.ones_like(xs), xs)
        expo = torch.where(neg, torch.zeros_like(xs), 2.0 * wm)
        mult = torch.prod(torch.pow(base, expo), dim=-1)
        m0 = torch.pow(torch.pow(eps_t, sw) + mult, 1.0 / sw)
        h_pos = torch.pow(m0, 0.5) - torch.pow(eps_t, 0.5)
        return torch.where(neg.any(dim=-1), h_neg, h_pos)

    def conj(self, x, mask, w):
        return self._and(x, mask, w)

    def disj(self, x, mask, w):
        return -self._and(-x, mask, w)

    def until(self, lw, rw, mask, cand, wl, wr, wp):
        n = lw.shape[-1]
        # prefix AND for every i: O(n^2) gathers, fine at oracle sizes.  Column k=0
        # (the sample at t itself) is always live, so no prefix row is empty.
        tri = torch.tril(torch.ones(n, n, dtype=torch.bool))  # [i, k]: k <= i
        lpre = lw.unsqueeze(-2).expand(*lw.shape[:-1], n, n)
        mpre = mask.unsqueeze(-2) & tri
        y = self._and(lpre, mpre, wl)  # [..., n]
        pair = torch.stack((y, torch.where(mask, rw, torch.ones_like(rw))), dim=-1)
        s = self._and(pair, torch.ones_like(pair, dtype=torch.bool), wp)
        first = int(cand.to(torch.int64).argmax())
        wr_full = torch.zeros(n, dtype=F64)
        m = min(n - first, wr.numel())
        wr_full[first : first + m] = wr[:m]
        return -self._and(-s, mask & cand, wr_full)


def make(name: str, **kw) -> _Dense:
    key = name.strip().lower().replace("-", "_")
    key = {"arithmetic_geometric_mean": "agm"}.get(key, key)
    table = {"classical": DClassic, "smooth": DSmooth, "cumulative": DCumulative, "agm": DAgm, "dgmsr": DDgmsr}
    if key not in table:
        raise KeyError(f"Unknown syntax '{name}'")
    return table[key](**kw)


# --------------------------------------------------------------------------- driver
def _predicate_trace(node: Any, sig: torch.Tensor) -> torch.Tensor:
    """``sig`` is [B, T, D]; returns [B, T]."""
    meta = getattr(node, "metadata", None) or {}
    if "affine" in meta:
        terms, bias = meta["affine"]
        out = torch.full(sig.shape[:2], float(bias), dtype=F64)
        acc = None
        for d, c in terms:  # same association as formula.affine_predicate.fn
            term = sig[..., d] * c if c != 1.0 else sig[..., d]
            acc = term if acc is None else acc + term
        if acc is None:
            return out + 0.0 * sig[..., 0]
        return acc + bias if bias != 0.0 else acc
    if node.fn is None:
        raise ValueError(f"Predicate '{node.name}' does not define `fn(signal, t)`.")
    B, T, _ = sig.shape
    rows = []
    for b in range(B):  # reference contract: fn(signal[T, D], t) -> scalar
        rows.append(torch.stack([torch.as_tensor(node.fn(sig[b], t), dtype=F64) for t in range(T)]))
    return torch.stack(rows)


def _gather(x: torch.Tensor, a: int, n: int):
    """x[..., T] -> (x[..., t, k] = x[..., t+a+k], mask[t, k] = t+a+k <= T-1)."""
    T = x.shape[-1]
    idx = torch.arange(T).unsqueeze(1) + a + torch.arange(n).unsqueeze(0)
    mask = idx <= T - 1
    return x[..., idx.clamp(max=T - 1)], mask


def _finite(x: torch.Tensor) -> torch.Tensor:
    """NaN (= "the reference raises here") -> 1.0 with zero gradient, so autograd stays finite."""
    return torch.where(torch.isnan(x), torch.ones_like(x), x)


def _dummy_live(mask: torch.Tensor):
    live = mask.any(dim=-1)
    first = torch.zeros(mask.shape[-1], dtype=torch.bool)
    first[0] = True
    return mask | (~live.unsqueeze(-1) & first), live


def robustness_trace(node: Any, signal: torch.Tensor, sem: _Dense) -> torch.Tensor:
    """All ``t`` of all traces.  ``signal``: [B, T, D] float64 (may require grad).

    Returns [B, T] (Cumulative: [2, B, T] = (pos, neg)); NaN where the reference raises."""
    sig = signal.to(F64)
    T = sig.shape[1]
    kind = type(node).__name__
    if kind == "Predicate":
        return sem.leaf(_predicate_trace(node, sig))
    if kind == "Not":
        return sem.neg(robustness_trace(node.child, sig, sem))
    if kind in ("And", "Or"):
        vals = torch.stack([robustness_trace(c, sig, sem) for c in node.children], dim=-1)
        w = _w(node.weights, vals.shape[-1], vals)
        mask = torch.ones(vals.shape, dtype=torch.bool)
        bad = torch.isnan(vals)
        out = (sem.conj if kind == "And" else sem.disj)(_finite(vals), mask, w)
        # a child the reference raises on makes the parent raise too
        return torch.where(bad.any(dim=-1), torch.full_like(out, float("nan")), out)
    if kind in ("Always", "Eventually"):
        x = robustness_trace(node.child, sig, sem)
        a, b = node.interval.start, node.interval.end
        n = (T if b is None else b - a + 1)
        xw, mask = _gather(x, a, n)
        mask, live = _dummy_live(mask)
        # window-relative weights, truncated at the horizon (agm.py:28-35)
        w = _w(node.weights, n, xw) if node.weights is not None else torch.ones(n, dtype=F64)
        out = (sem.always if kind == "Always" else sem.eventually)(_finite(xw), mask, w)
        live = live & ~(torch.isnan(xw) & mask).any(dim=-1)
        return torch.where(live, out, torch.full_like(out, float("nan")))
    if kind == "Until":
        l = robustness_trace(node.left, sig, sem)
        r = robustness_trace(node.right, sig, sem)
        a, b = node.interval.start, node.interval.end
        n = T if b is None else b + 1
        lw, mask = _gather(l, 0, n)
        rw, _ = _gather(r, 0, n)
        cand = torch.arange(n) >= a
        live = (mask & cand).any(dim=-1)
        if max(required_start_slack(node.left), required_start_slack(node.right)) > 0:
            # api.py:288-293 evaluates both children up to the horizon, so a child
            # with a non-zero start raises for EVERY t
            live = torch.zeros_like(live)
        mask, _ = _dummy_live(mask)
        cand_safe = cand.clone()
        if not bool(cand.any()):
            cand_safe[0] = True
        wl = _w(node.weights_left, n, lw) if node.weights_left is not None else torch.ones(n, dtype=F64)
        nr = max(n - a, 1)
        wr = _w(node.weights_right, nr, lw) if node.weights_right is not None else torch.ones(nr, dtype=F64)
        wp = _w(node.weights_pair, 2, lw)
        # rows with no live candidate: evaluate on a dummy candidate, then NaN
        mc = mask & cand_safe
        dead = ~mc.any(dim=-1)
        if bool(dead.any()):
            first_c = torch.zeros(n, dtype=torch.bool)
            first_c[int(cand_safe.to(torch.int64).argmax())] = True
            mask = mask | (dead.unsqueeze(-1) & first_c)
        out = sem.until(_finite(lw), _finite(rw), mask, cand_safe, wl, wr, wp)
        live = live & ~((torch.isnan(lw) | torch.isnan(rw)) & mask).any(dim=-1)
        return torch.where(live, out, torch.full_like(out, float("nan")))
    raise TypeError(f"Unsupported formula node type: {type(node)!r}")


def required_start_slack(node: Any) -> int:
    """Largest sum of interval starts along a root-to-leaf path: outputs at
    ``t > T-1-slack`` have an empty window somewhere (the reference raises)."""
    kind = type(node).__name__
    if kind == "Predicate":
        return 0
    if kind == "Not":
        return required_start_slack(node.child)
    if kind in ("And", "Or"):
        return max(required_start_slack(c) for c in node.children)
    if kind in ("Always", "Eventually"):
        return node.interval.start + required_start_slack(node.child)
    return node.interval.start + max(required_start_slack(node.left), required_start_slack(node.right))
