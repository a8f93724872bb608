"""Shared helpers of the parity tests: tolerances, random formulas, comparisons."""

from __future__ import annotations

import numpy as np
import torch

from oracle import dense
from pystl_b200.formula import from_spec

# Tolerance of the fp32 kernels against the float64 reference:
# bit-exact for classic; |got - want| <= ATOL + RTOL * |want| for the smooth semantics.
RTOL = 1e-5
ATOL = 5e-6


def oracle_trace(spec, sem: str, kw: dict, sig64: np.ndarray) -> np.ndarray:
    out = dense.robustness_trace(from_spec(spec), torch.tensor(sig64), dense.make(sem, **kw)).numpy()
    return np.moveaxis(out, 0, -1) if sem == "cumulative" else out


def parity_log(kind: str, what: str, sem: str, ratio: float, scale: float) -> None:
    """PYSTL_PARITY_LOG=<file>: append the worst error / stated-bar ratio of every comparison (the
    measured table under profiles/ is made from it: tests/error_table.py --from-log)."""
    import os

    path = os.environ.get("PYSTL_PARITY_LOG")
    if path:
        with open(path, "a") as fh:
            fh.write(f"{kind}\t{sem}\t{ratio:.4g}\t{scale:g}\t{what}\n")


def scale_cap(scale: float) -> float:
    """PYSTL_PARITY_SCALE_CAP=<x>: tighten every tolerance multiplier to at most x (experiments)."""
    import os

    cap = os.environ.get("PYSTL_PARITY_SCALE_CAP")
    return min(scale, float(cap)) if cap else scale


def conditioning_slack(spec, sem: str, kw: dict, sig64: np.ndarray, b: int, t: int, c: int = 0) -> float:
    """First-order bound on how far the REFERENCE's own value at (b, t) moves when every input sample
    is rounded to fp32: sum_j |d out / d x_j| * ulp32(x_j) / 2, by autograd through the float64 oracle."""
    x = torch.tensor(sig64[b : b + 1], requires_grad=True)
    out = dense.robustness_trace(from_spec(spec), x, dense.make(sem, **kw))
    o = out[c, 0, t] if sem == "cumulative" else out[0, t]
    (g,) = torch.autograd.grad(o, x)
    ulp = np.spacing(np.abs(sig64[b]).astype(np.float32)).astype(np.float64)
    return float((g[0].abs().numpy() * ulp * 0.5).sum())


def assert_parity(got: np.ndarray, want64: np.ndarray, sem: str, what: str = "", scale: float = 1.0, ill=None) -> None:
    """``ill = (spec, kw, sig64)`` admits ILL-CONDITIONED outputs: an entry outside the bar passes when the
    reference's own value there moves by more than 1 / 64 of the miss under fp32 rounding of the inputs
    (``conditioning_slack``: D-GMSR composed four levels deep is that steep where an operand is ~ 1e-5,
    SURVEY.md appendix A10); at most 4 such entries per comparison."""
    scale = scale_cap(scale)
    assert got.shape == want64.shape, (what, got.shape, want64.shape)
    assert np.array_equal(np.isnan(got), np.isnan(want64)), f"{what}: NaN pattern differs"
    if sem == "classical":
        w32 = want64.astype(np.float32)
        bad = ~((got == w32) | (np.isnan(got) & np.isnan(w32)))
        assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} classic values are not bit-exact"
        return
    err = np.abs(got.astype(np.float64) - want64)
    tol = scale * (ATOL + RTOL * np.abs(want64))
    with np.errstate(invalid="ignore"):
        parity_log("value", what, sem, float(np.nanmax(np.where(np.isnan(want64), 0.0, err / (ATOL + RTOL * np.abs(want64))), initial=0.0)), scale)
    bad = err > tol
    bad &= ~np.isnan(want64)
    if bad.any() and ill is not None and int(bad.sum()) <= 4:
        spec, kw, sig64 = ill
        for idx in np.argwhere(bad):
            b, t = int(idx[0]), int(idx[1])
            slack = conditioning_slack(spec, sem, kw, sig64, b, t, int(idx[2]) if len(idx) > 2 else 0)
            parity_log("ill-conditioned", f"{what}@{b},{t}", sem, float(err[tuple(idx)] / max(slack, 1e-300)), scale)
            if err[tuple(idx)] <= tol[tuple(idx)] + 64.0 * slack:
                bad[tuple(idx)] = False
    assert not bad.any(), (
        f"{what}: {int(bad.sum())} of {bad.size} values off; worst abs err {np.nanmax(err):.3e} "
        f"at want={want64.flat[np.nanargmax(err)]:.6g}"
    )


def mixed_signal(rng: np.random.Generator, B: int, T: int, D: int) -> np.ndarray:
    """mu_b + 0.5 N(0,1), mu ~ U(-1, 3): both AGM / D-GMSR branches occur."""
    mu = rng.uniform(-1.0, 3.0, size=(B, 1, D))
    x = mu + 0.5 * rng.standard_normal((B, T, D))
    x[0] = rng.standard_normal((T, D))
    return x.astype(np.float32).astype(np.float64)


def random_spec(rng: np.random.Generator, D: int, depth: int, max_b: int, allow_open: bool = False):
    """Random formula spec over D channels with fp32-representable constants."""
    def pred():
        d = int(rng.integers(D))
        c = float(rng.choice([-0.5, 0.0, 0.25, 0.5, 1.0, 1.5, 2.0]))
        sign = float(rng.choice([-1.0, 1.0]))
        if D > 1 and rng.random() < 0.25:
            d2 = int((d + 1 + rng.integers(D - 1)) % D)
            return ("pred", {d: 1.0, d2: -1.0}, 0.0)
        return ("pred", {d: sign}, -sign * c)

    def interval():
        a = int(rng.choice([0, 0, 0, 1, 3]))
        if allow_open and rng.random() < 0.2:
            return (a, None)
        return (a, a + int(rng.integers(0, max_b + 1)))

    def node(d):
        if d == 0 or rng.random() < 0.15:
            return pred()
        r = rng.random()
        if r < 0.15:
            return ("not", node(d - 1))
        if r < 0.40:
            k = int(rng.integers(2, 4))
            return (str(rng.choice(["and", "or"])), [node(d - 1) for _ in range(k)])
        if r < 0.85:
            return (str(rng.choice(["always", "eventually"])), interval(), node(d - 1))
        a = int(rng.choice([0, 0, 1]))
        b = None if (allow_open and rng.random() < 0.2) else a + int(rng.integers(0, min(max_b, 24) + 1))
        return ("until", (a, b), node(d - 1), node(d - 1))

    return node(depth)


SEM_KW = [
    ("classical", {}),
    ("smooth", {"temperature": 0.5}),
    ("cumulative", {}),
    ("agm", {}),
    ("dgmsr", {"eps": 1e-8, "p": 1}),
    ("dgmsr", {"eps": 1e-8, "p": 2}),
]
