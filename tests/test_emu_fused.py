"""CPU tier: the FUSED kernels (generated source + fused_rt.cuh) compiled with g++ against the
fiber-based SIMT stand-in (tests/emu/simt_host.h), against the oracle."""

from __future__ import annotations

import pytest

import fused_emu_lib
from fused_cases import case_signal, forward_cases
from parity_util import assert_parity, oracle_trace
from pystl_b200.formula import BASELINE_SPECS, from_spec

FWD = forward_cases(light=True)


@pytest.mark.parametrize("case", FWD, ids=[c["name"] for c in FWD])
def test_forward(case):
    sig = case_signal(case)
    want = oracle_trace(case["spec"], case["sem"], case["kw"], sig)
    got = fused_emu_lib.eval_forward(from_spec(case["spec"]), case["sem"], sig, force_threads=case["threads"], **case["kw"])
    ill = (case["spec"], case["kw"], sig) if case["sem"] == "dgmsr" else None
    assert_parity(got, want, case["sem"], case["name"], scale=case["scale"], ill=ill)


# ---------------------------------------------------------------------------- backward
import re  # noqa: E402

import numpy as np  # noqa: E402

from fused_cases import backward_cases  # noqa: E402
from parity_util import mixed_signal, random_spec  # noqa: E402
from test_emu_backward import check, grad_tol, has_until, oracle_grad  # noqa: E402

BWD = backward_cases(light=True)


@pytest.mark.parametrize("case", BWD, ids=[c["name"] for c in BWD])
@pytest.mark.parametrize("cotangent", ["explicit", "constant"])
def test_backward(case, cotangent):
    sig = case_signal(case)
    rng = np.random.default_rng(case["seed"] + 1)
    B, T = sig.shape[:2]
    if cotangent == "explicit":
        g = rng.standard_normal((B, T)).astype(np.float32).astype(np.float64)
        got = fused_emu_lib.eval_backward(from_spec(case["spec"]), case["sem"], sig, g, force_threads=case["threads"], **case["kw"])
    else:
        g = np.full((B, T), 0.25)
        got = fused_emu_lib.eval_backward(from_spec(case["spec"]), case["sem"], sig, None, scale=0.25, force_threads=case["threads"], **case["kw"])
    want = oracle_grad(case["spec"], case["sem"], case["kw"], sig, g)
    rtol, atol = grad_tol(case["sem"], case["kw"])
    check(got.astype(np.float64), want, case["sem"], case["name"], rtol, atol)


def test_smooth_extreme_dynamic_range_is_exact():
    """tau = 0.01 on a signal of range ~30: the block-shifted sums underflow; the forward pass
    recomputes those windows exactly, the backward pass scatters them term by term."""
    rng = np.random.default_rng(3)
    sig = (rng.standard_normal((2, 300, 2)) * 5.0).astype(np.float32).astype(np.float64)
    kw = {"temperature": 0.01}
    rtol, atol = grad_tol("smooth", kw)
    spec = BASELINE_SPECS["cfg2"]
    got = fused_emu_lib.eval_forward(from_spec(spec), "smooth", sig, force_threads=64, **kw)
    assert_parity(got, oracle_trace(spec, "smooth", kw, sig), "smooth", "cfg2/tau0.01")
    g = rng.standard_normal((2, 300)).astype(np.float32).astype(np.float64)
    gb = fused_emu_lib.eval_backward(from_spec(spec), "smooth", sig, g, force_threads=64, **kw)
    check(gb.astype(np.float64), oracle_grad(spec, "smooth", kw, sig, g), "smooth", "cfg2/tau0.01", rtol, atol)


@pytest.mark.parametrize("cfg, T, kw", [("cfg2", 4096, {}), ("cfg2", 2304, {"temperature": 0.05}), ("cfg3", 4096, {})],
                         ids=["cfg2-T4096", "cfg2-T2304-tau0.05", "cfg3-T4096"])
def test_smooth_wide_tile_default(cfg, T, kw):
    """Smooth value + gradient of window formulas at BASELINE horizons with the geometry the product picks by itself: tiles of 288
    threads whose lifted operands are lifted again from the stash for the adjoint (codegen.cpp: kWideSmoothThreads, Gen::relift)."""
    spec = BASELINE_SPECS[cfg]
    rng = np.random.default_rng(T)
    sig = rng.standard_normal((2, T, 2)).astype(np.float32).astype(np.float64)
    cf = fused_emu_lib.compile_formula(from_spec(spec), "smooth", 2, **kw)
    src, _, _, _ = fused_emu_lib.prepare(cf, 2, T, 2, True, 0)
    assert re.search(r"constexpr int NT = 288\b", src), "the wide tile is not the default at this shape any more: move the test"
    g = rng.standard_normal((2, T)).astype(np.float32).astype(np.float64)
    got = fused_emu_lib.eval_backward(from_spec(spec), "smooth", sig, g, **kw)
    rtol, atol = grad_tol("smooth", kw)
    check(got.astype(np.float64), oracle_grad(spec, "smooth", kw, sig, g), "smooth", f"{cfg}/T{T}/wide", rtol, atol)


@pytest.mark.parametrize("seed", [1, 3, 8, 20, 25, 34])
def test_smooth_wide_tile_random_formulas(seed):
    """Random nested window formulas (windows up to 150 samples, 1 - 3 channels, tau 1 / 0.25) at horizons where the product picks
    the 288-thread tile by itself: halos on both sides of a tile, n-ary operators and negations with the re-lifted operands."""
    rng = np.random.default_rng(9000 + seed)
    D = int(rng.integers(1, 4))
    spec = random_spec(rng, D, depth=3, max_b=int(rng.choice([24, 60, 150])), allow_open=False)
    T = int(rng.choice([2100, 2304, 2500, 4096, 4608, 5000]))
    kw = {"temperature": float(rng.choice([1.0, 0.25]))}
    sig = mixed_signal(rng, 1, T, D)
    cf = fused_emu_lib.compile_formula(from_spec(spec), "smooth", D, **kw)
    src, _, _, _ = fused_emu_lib.prepare(cf, 1, T, D, True, 0)
    assert not has_until(spec) and re.search(r"constexpr int NT = 288\b", src), "seed list out of date: pick seeds that land on the wide tile"
    g = rng.uniform(0.25, 1.0, size=(1, T)).astype(np.float32).astype(np.float64)
    got = fused_emu_lib.eval_backward(from_spec(spec), "smooth", sig, g, **kw)
    rtol, atol = grad_tol("smooth", kw)
    check(got.astype(np.float64), oracle_grad(spec, "smooth", kw, sig, g), "smooth", f"wide/seed{seed}/{spec}", rtol, atol)


from fused_cases import cumulative_backward_cases  # noqa: E402

CBWD = cumulative_backward_cases(light=True)


@pytest.mark.parametrize("case", CBWD, ids=[c["name"] for c in CBWD])
def test_cumulative_backward(case):
    """Sum windows, arg-min / arg-max routing, until: against autograd through the dense oracle
    (torch.min / max(dim) and torch.minimum routing of the reference, torch/semantics.py:253-345)."""
    sig = case_signal(case)
    rng = np.random.default_rng(case["seed"] + 1)
    B, T = sig.shape[:2]
    g = rng.standard_normal((B, T, 2)).astype(np.float32).astype(np.float64)
    got = fused_emu_lib.eval_backward(from_spec(case["spec"]), "cumulative", sig, g, force_threads=case["threads"])
    check(got.astype(np.float64), oracle_grad(case["spec"], "cumulative", {}, sig, g), "cumulative", case["name"], 4e-5, 4e-5)  # (twice the bar: scatter-sums, see test_gpu_fused.CUM_GTOL)


from fused_cases import smooth_until_backward_cases  # noqa: E402

UBWD = smooth_until_backward_cases(light=True)


@pytest.mark.parametrize("case", UBWD, ids=[c["name"] for c in UBWD])
def test_smooth_until_backward(case):
    """Gradient of the smooth Until (closed form of SURVEY.md appendix A4, log-domain sweeps)."""
    sig = case_signal(case)
    rng = np.random.default_rng(case["seed"] + 1)
    g = rng.standard_normal(sig.shape[:2]).astype(np.float32).astype(np.float64)
    got = fused_emu_lib.eval_backward(from_spec(case["spec"]), "smooth", sig, g, force_threads=case["threads"], **case["kw"])
    rtol, atol = grad_tol("smooth", case["kw"])
    check(got.astype(np.float64), oracle_grad(case["spec"], "smooth", case["kw"], sig, g), "smooth", case["name"], rtol, atol)


from fused_cases import random_until_backward_cases  # noqa: E402

RUB = random_until_backward_cases(light=True)


@pytest.mark.parametrize("case", RUB, ids=[c["name"] for c in RUB])
def test_random_until_backward(case):
    """Randomised Until adjoints (see tests/test_gpu_fused.py: the same generator, four cases here)."""
    sig = case_signal(case)
    rng = np.random.default_rng(case["seed"] + 1)
    g = rng.standard_normal(sig.shape[:2]).astype(np.float32).astype(np.float64)
    try:
        got = fused_emu_lib.eval_backward(from_spec(case["spec"]), case["sem"], sig, g, force_threads=case["threads"], **case["kw"])
    except NotImplementedError as e:
        pytest.skip(str(e))
    rtol, atol = grad_tol(case["sem"], case["kw"])
    check(got.astype(np.float64), oracle_grad(case["spec"], case["sem"], case["kw"], sig, g), case["sem"], case["name"], rtol, atol)


from fused_cases import nan_flow_backward_cases  # noqa: E402

NFB = nan_flow_backward_cases(light=True)


@pytest.mark.parametrize("case", NFB, ids=[c["name"] for c in NFB])
def test_nan_values_do_not_leak_into_gradients(case):
    """Values that are NaN where the reference raises (an Until with a lower bound, near the end of the trace) sit under
    windows and And / Or: the gradient of the valid outputs must be finite and equal to the oracle's."""
    sig = case_signal(case)
    rng = np.random.default_rng(case["seed"] + 1)
    C = 2 if case["sem"] == "cumulative" else 1
    g = rng.standard_normal(sig.shape[:2] + ((2,) if C == 2 else ())).astype(np.float32).astype(np.float64)
    got = fused_emu_lib.eval_backward(from_spec(case["spec"]), case["sem"], sig, g, force_threads=case["threads"], **case["kw"]).astype(np.float64)
    rtol, atol = grad_tol(case["sem"], case["kw"])
    k = 2.0 if case["sem"] == "cumulative" else 1.0
    check(got, oracle_grad(case["spec"], case["sem"], case["kw"], sig, g), case["sem"], case["name"], k * rtol, k * atol)


from fused_cases import gm_backward_cases  # noqa: E402

GMB = gm_backward_cases(light=True)


@pytest.mark.parametrize("case", GMB, ids=[c["name"] for c in GMB])
def test_agm_dgmsr_until_and_weighted_backward(case):
    """Until (agm.py:90-130, dgmsr.py:88-118; with weights_left / right / pair) and per-position
    window weights: gradients against autograd through the dense oracle."""
    sig = case_signal(case)
    rng = np.random.default_rng(case["seed"] + 1)
    g = rng.standard_normal(sig.shape[:2]).astype(np.float32).astype(np.float64)
    got = fused_emu_lib.eval_backward(from_spec(case["spec"]), case["sem"], sig, g, force_threads=case["threads"], **case["kw"]).astype(np.float64)
    check(got, oracle_grad(case["spec"], case["sem"], case["kw"], sig, g), case["sem"], case["name"])


# ---------------------------------------------------------------------------- classic gradient (fused tier)


@pytest.mark.parametrize("seed", range(6))
def test_classic_gradient_random_formulas(seed):
    """One-hot routing of torch.min / max(dim) (first index on ties) through nested windows, n-ary
    operators and the decomposed Until, on signals quantised to 0.5 so that ties are frequent."""
    rng = np.random.default_rng(4100 + seed)
    D = int(rng.integers(1, 4))
    T = int(rng.integers(60, 260))
    spec = random_spec(rng, D, depth=3, max_b=24, allow_open=False)
    sig = mixed_signal(rng, 2, T, D)
    if seed % 2 and not has_until(spec):
        # (the reference's Until goes through torch.minimum / maximum, which SPLIT ties between
        # different leaves 0.5 / 0.5; no tier reproduces that, ties there have measure zero on real data)
        sig = np.round(sig * 2.0) / 2.0
    g = rng.uniform(0.25, 1.0, size=(2, T)).astype(np.float32).astype(np.float64)
    want = oracle_grad(spec, "classical", {}, sig, g)
    for threads in (0, 32):
        try:
            got = fused_emu_lib.eval_backward(from_spec(spec), "classical", sig, g, force_threads=threads)
        except NotImplementedError:
            assert threads  # the two-sided halo does not fit the forced tile
            continue
        check(got.astype(np.float64), want, "classical", f"seed{seed}/nt{threads}/{spec}")


def test_classic_until_tie_routing_touches_active_leaves_only():
    from test_emu_backward import tie_routing_report

    tie_routing_report(lambda spec, sig, g: fused_emu_lib.eval_backward(from_spec(spec), "classical", sig, g))


# ---------------------------------------------------------------------------- pair predicates (host-computed cumulative values)
@pytest.mark.parametrize("threads", [0, 64])
def test_cumulative_pair_predicates_read_the_channels_as_they_stand(threads):
    """``pystl_pred_t.pair``: a cumulative value that already exists as two channels (pos, neg) -- how an open-ended window
    computed by the scan kernels enters a formula.  Feeding (max(0, v), min(0, v)) of an affine v through pair predicates must
    reproduce the formula over the affine predicates (cumulative.py:24-30), values and gradients (chain rule through the
    max / min by hand)."""
    from pystl_b200.formula import Always, And, Eventually, Interval, Not, Or, Predicate, Until, affine_predicate

    rng = np.random.default_rng(4)
    B, T = 3, 700
    v = (0.6 * rng.standard_normal((B, T, 2))).astype(np.float32)
    sig4 = np.stack([np.maximum(v[..., 0], 0), np.minimum(v[..., 0], 0), np.maximum(v[..., 1], 0), np.minimum(v[..., 1], 0)], -1)
    x, y = affine_predicate("x", {0: 1.0}, 0.0), affine_predicate("y", {1: 1.0}, 0.0)
    px, py = Predicate("px", metadata={"pair": (0, 1)}), Predicate("py", metadata={"pair": (2, 3)})

    def build(a, b):
        return Or(Always(And(a, Not(b)), Interval(0, 7)), Eventually(Until(a, b, Interval(0, 5)), Interval(2, 9)))

    ref = fused_emu_lib.eval_forward(build(x, y), "cumulative", v, force_threads=threads)
    got = fused_emu_lib.eval_forward(build(px, py), "cumulative", sig4.astype(np.float32), force_threads=threads)
    ok = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), ok) and np.array_equal(got[ok], ref[ok])
    go = rng.uniform(0.25, 1.0, size=ref.shape).astype(np.float32)
    g_ref = fused_emu_lib.eval_backward(build(x, y), "cumulative", v, go, force_threads=threads)
    g = fused_emu_lib.eval_backward(build(px, py), "cumulative", sig4.astype(np.float32), go, force_threads=threads)
    for d in (0, 1):
        chained = g[..., 2 * d] * (v[..., d] > 0) + g[..., 2 * d + 1] * (v[..., d] < 0)
        assert np.allclose(chained, g_ref[..., d], rtol=1e-6, atol=1e-6)  # (one ulp: the affine path contracts coeff * (g+ d+ + g- d-) differently)
