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
    tol = atol + rtol * np.abs(want)
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
