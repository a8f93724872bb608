"""CPU tier: backward pass of the kernel (tile_bwd.h, run by the host emulator) against the
reference's gradients (golden vectors: autograd through pystl/torch/semantics.py) and against
autograd through the dense oracle on larger / multi-tile cases."""

from __future__ import annotations

import numpy as np
import pytest
import torch

import emu_lib
from golden_cases import CASES, case_cotangent, case_signal
from oracle import dense
from parity_util import SEM_KW, mixed_signal, random_spec
from pystl_b200.formula import BASELINE_SPECS, from_spec

GOLD = np.load(__file__.replace("test_emu_backward.py", "golden/reference_vectors.npz"))

# gradient tolerance: fp32 kernel vs float64 autograd, cotangents of O(1)
G_RTOL, G_ATOL = 2e-5, 2e-5


def has_until(spec) -> bool:
    if spec[0] == "until":
        return True
    if spec[0] == "pred":
        return False
    if spec[0] == "not":
        return has_until(spec[1])
    if spec[0] in ("and", "or"):
        return any(has_until(c) for c in spec[1])
    return has_until(spec[2])


def supported(case) -> bool:
    if case["sem"] == "cumulative":
        return False
    if case["spec_name"] == "weighted" and case["sem"] in ("agm", "dgmsr"):
        return False
    return case["sem"] == "classical" or not has_until(case["spec"])


GRAD_CASES = [c for c in CASES if c["grad"] and supported(c)]


def oracle_grad(spec, sem, kw, sig64, g64):
    """d sum(g * out) / d signal by autograd through the dense oracle (cumulative: g is [B, T, 2])."""
    x = torch.tensor(sig64, requires_grad=True)
    out = dense.robustness_trace(from_spec(spec), x, dense.make(sem, **kw))
    live = ~torch.isnan(out)
    g = torch.tensor(np.moveaxis(g64, -1, 0)) if sem == "cumulative" else torch.tensor(g64)
    (torch.where(live, out, torch.zeros_like(out)) * g * live).sum().backward()
    return x.grad.numpy()


def grad_tol(sem: str, kw: dict):
    """(rtol, atol) of fp32 gradients against float64 autograd.  Smooth weights are
    exp((out - x) / tau): an fp32 rounding of out or x (6e-8 relative, plus the 2-ulp hardware
    exp2 / log2) is amplified by 1 / tau, so the tolerance scales with it below tau = 0.25."""
    scale = 1.0
    if sem == "smooth":
        scale = max(1.0, 0.25 / kw.get("temperature", 1.0))
    return G_RTOL * scale, G_ATOL * scale


def check(got, want, sem, what, rtol=G_RTOL, atol=G_ATOL):
    assert got.shape == want.shape
    assert np.isfinite(got).all(), f"{what}: non-finite gradient"
    if sem == "classical":
        # one-hot routing: the same entries are hit, with sums of the same cotangents
        assert np.array_equal(got != 0, want != 0), f"{what}: classic gradient routing differs"
    err = np.abs(got - want)
    from parity_util import parity_log, scale_cap

    parity_log("grad", what, sem, float((err / (G_ATOL + G_RTOL * np.abs(want))).max(initial=0.0)), max(rtol / G_RTOL, atol / G_ATOL))
    k = scale_cap(max(rtol / G_RTOL, atol / G_ATOL)) / max(rtol / G_RTOL, atol / G_ATOL)
    tol = k * (atol + rtol * np.abs(want))
    assert (err <= tol).all(), f"{what}: worst gradient error {err.max():.3e} (|want| up to {np.abs(want).max():.3e})"


@pytest.mark.parametrize("case", GRAD_CASES, ids=[c["name"] for c in GRAD_CASES])
@pytest.mark.parametrize("threads", [0, 8, 16])
def test_golden_gradients(case, threads):
    sig = case_signal(case)
    vals = GOLD[f"{case['name']}/values"]
    g = case_cotangent(case, vals.shape)
    try:
        got = emu_lib.eval_backward(from_spec(case["spec"]), case["sem"], sig, g, force_threads=threads, **case["kw"])
    except NotImplementedError:
        # tiles too small for the two-sided halo of this formula
        assert threads
        return
    check(got, GOLD[f"{case['name']}/grad"], case["sem"], case["name"], *grad_tol(case["sem"], case["kw"]))


BWD_SEMS = [SEM_KW[0], SEM_KW[1], SEM_KW[3], SEM_KW[4], SEM_KW[5]]


@pytest.mark.parametrize("sem,kw", BWD_SEMS, ids=["classical", "smooth", "agm", "dgmsr", "dgmsr_p2"])
@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg3"])
def test_baseline_formulas_multi_tile(cfg, sem, kw):
    rng = np.random.default_rng(17)
    sig = mixed_signal(rng, 2, 900, 2)
    g = rng.uniform(0.25, 1.0, size=(2, 900)).astype(np.float32).astype(np.float64)
    want = oracle_grad(BASELINE_SPECS[cfg], sem, kw, sig, g)
    for threads in (0, 64):
        got = emu_lib.eval_backward(from_spec(BASELINE_SPECS[cfg]), sem, sig, g, force_threads=threads, **kw)
        check(got, want, sem, f"{cfg}/{sem}/nt{threads}")


def test_constant_cotangent_is_mean_gradient():
    rng = np.random.default_rng(23)
    sig = mixed_signal(rng, 2, 300, 2)
    scale = 1.0 / (2 * 300)
    for sem, kw in BWD_SEMS[:3]:
        want = oracle_grad(BASELINE_SPECS["cfg2"], sem, kw, sig, np.full((2, 300), scale))
        got = emu_lib.eval_backward(from_spec(BASELINE_SPECS["cfg2"]), sem, sig, None, scale=scale, **kw)
        assert np.abs(got - want).max() <= 1e-7 + 2e-5 * np.abs(want).max()


@pytest.mark.parametrize("seed", range(16))
def test_random_formulas(seed):
    rng = np.random.default_rng(3000 + seed)
    D = int(rng.integers(1, 4))
    T = int(rng.integers(30, 200))
    spec = random_spec(rng, D, depth=3, max_b=24, allow_open=(seed % 5 == 0))
    sig = mixed_signal(rng, 2, T, D)
    g = rng.uniform(0.25, 1.0, size=(2, T)).astype(np.float32).astype(np.float64)
    for sem, kw in BWD_SEMS:
        if sem != "classical" and has_until(spec):
            continue
        want = oracle_grad(spec, sem, kw, sig, g)
        for threads in (0, 32):
            try:
                got = emu_lib.eval_backward(from_spec(spec), sem, sig, g, force_threads=threads, **kw)
            except NotImplementedError:
                assert threads
                continue
            check(got, want, sem, f"seed{seed}/{sem}/nt{threads}/{spec}")


def test_smooth_extreme_range_gradient():
    rng = np.random.default_rng(31)
    sig = (rng.standard_normal((2, 200, 2)) * 5.0).astype(np.float32).astype(np.float64)
    g = np.ones((2, 200))
    kw = {"temperature": 0.01}
    want = oracle_grad(BASELINE_SPECS["cfg1"], "smooth", kw, sig, g)
    got = emu_lib.eval_backward(from_spec(BASELINE_SPECS["cfg1"]), "smooth", sig, g, **kw)
    check(got, want, "smooth", "sharp cfg1", *grad_tol("smooth", kw))


def test_unsupported_backward_is_loud():
    sig = np.zeros((1, 20, 2), np.float32)
    with pytest.raises(NotImplementedError):
        emu_lib.eval_backward(from_spec(BASELINE_SPECS["cfg4"]), "cumulative", np.zeros((1, 20, 4), np.float32))
    with pytest.raises(NotImplementedError):
        emu_lib.eval_backward(from_spec(("until", (0, 3), ("pred", {0: 1.0}, 0.0), ("pred", {1: 1.0}, 0.0))), "smooth", sig)


# ---------------------------------------------------------------------------- classic Until under ties
def _pred_specs(spec, out):
    if spec[0] == "pred":
        out.append(spec)
    elif spec[0] == "not":
        _pred_specs(spec[1], out)
    elif spec[0] in ("and", "or"):
        for c in spec[1]:
            _pred_specs(c, out)
    elif spec[0] == "until":
        _pred_specs(spec[2], out)
        _pred_specs(spec[3], out)
    else:
        _pred_specs(spec[2], out)
    return out


def check_until_tie_routing(grad_fn, spec, seed: int = 0, T: int = 60, D: int = 2, n_out: int = 16):
    """Documents the one deviation of the classic gradient from the reference.

    ``torch/semantics.py:141-145`` folds the Until candidates with ``torch.minimum / maximum``, which
    SPLIT the cotangent 0.5 / 0.5 whenever two operands are equal, so a chain of ties between
    *different* leaves yields weights 1/2, 1/4, ... over them.  The kernels evaluate the Until through
    ``min(G[0,a] l, W(t+a), F[a,b] r)`` and route every output to ONE leaf (first
    index on ties).  Values are bit-exact either way.  On a signal quantised to 0.5 (ties everywhere) this
    checks, output by output (one-hot cotangents, one trace replicated over the batch), what both
    routings have in common:
      * the kernel's gradient is one-hot with the magnitude of the cotangent, the reference's weights
        sum to it;
      * both only touch ACTIVE leaves: a predicate sample whose value equals the output's value.
    Returns (outputs where the reference splits, outputs where the kernel's leaf is outside the
    reference's support): the deviation, which needs exact ties between different leaves and has measure
    zero on real-valued signals (the un-quantised corpora assert identical routing).
    ``grad_fn(spec, sig[B, T, D], g[B, T]) -> grad[B, T, D]``."""
    rng = np.random.default_rng(seed)
    base = np.round(mixed_signal(rng, 1, T, D)[0] * 2.0) / 2.0
    outs = rng.choice(T - 1, size=n_out, replace=False)
    sig = np.repeat(base[None], n_out, axis=0)
    g = np.zeros((n_out, T))
    g[np.arange(n_out), outs] = 1.0
    want = oracle_grad(spec, "classical", {}, sig, g)
    got = np.asarray(grad_fn(spec, sig, g), dtype=np.float64)
    values = dense.robustness_trace(from_spec(spec), torch.tensor(base[None]), dense.make("classical")).numpy()[0]
    preds = _pred_specs(spec, [])

    def active(i, d, v):  # some predicate reading channel d has the output's value (up to a Not) at sample i
        for _, coeffs, bias in preds:
            if d in coeffs and abs(abs(bias + sum(c * base[i, ch] for ch, c in coeffs.items())) - abs(v)) < 1e-12:
                return True
        return False

    n_split = n_outside = 0
    for k in range(n_out):
        w, m, v = want[k], got[k], values[outs[k]]
        if not np.any(w):  # the reference raises at this t (empty window): no gradient on either side
            assert not np.any(m)
            continue
        hit = np.argwhere(m != 0)
        assert len(hit) == 1, f"output {outs[k]}: the kernel's gradient is not one-hot"
        i, d = hit[0]
        assert abs(abs(m[i, d]) - 1.0) < 1e-6 and active(i, d, v), f"output {outs[k]}: routed to an inactive leaf"
        # the reference: weights 1/2^k over active leaves (several paths can reach one leaf with opposite
        # signs only through a Not, which these formulas do not put between tied operands)
        for (i2, d2) in np.argwhere(w != 0):
            assert active(i2, d2, v)
        n_split += int(np.count_nonzero(w) > 1)
        n_outside += int(w[i, d] == 0)
    return n_split, n_outside


TIE_SPECS = [
    ("until", (0, 6), ("pred", {0: 1.0}, 0.0), ("pred", {1: 1.0}, -0.5)),
    ("until", (2, 9), ("pred", {0: 1.0}, -0.5), ("not", ("pred", {1: 1.0}, -1.0))),
    ("always", (0, 4), ("until", (0, 5), ("pred", {0: 1.0}, 0.0), ("pred", {0: -1.0}, 1.0))),
]


def tie_routing_report(grad_fn):
    res = [check_until_tie_routing(grad_fn, spec, seed=s) for spec in TIE_SPECS for s in range(3)]
    splits, outside = sum(r[0] for r in res), sum(r[1] for r in res)
    assert splits > 0, "the quantised signals must exercise ties between different leaves"
    return splits, outside


def test_classic_until_tie_routing_touches_active_leaves_only():
    tie_routing_report(lambda spec, sig, g: emu_lib.eval_backward(from_spec(spec), "classical", sig, g))
