"""GPU tier: the backward kernels through the public API (torch.autograd.Function over the C ABI)
against the reference's gradients (golden vectors) and autograd through the dense oracle."""

from __future__ import annotations

import os

import numpy as np
import pytest
import torch

from golden_cases import case_cotangent, case_signal
from parity_util import SEM_KW, mixed_signal, random_spec
from pystl_b200 import create_semantics
from pystl_b200.formula import BASELINE_SPECS, from_spec
from test_emu_backward import BWD_SEMS, GOLD, GRAD_CASES, check, grad_tol, has_until, oracle_grad

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _interpreter_tier(monkeypatch):
    """This file pins the INTERPRETER tier (the universal fallback); tests/test_gpu_fused.py runs
    the same kind of corpus through the fused tier."""
    monkeypatch.setenv("PYSTL_B200_TIER", "interp")


def gpu_grad(spec, sem, kw, sig64, g64, threads: int = 0):
    s = create_semantics(sem, backend="b200", **kw)
    x = torch.tensor(sig64, dtype=torch.float32, device="cuda", requires_grad=True)
    if threads:
        os.environ["PYSTL_B200_FORCE_THREADS"] = str(threads)
    try:
        out = s.robustness_trace(from_spec(spec), x)
        g = torch.tensor(g64, dtype=torch.float32, device="cuda")
        live = ~torch.isnan(out)
        (torch.where(live, out, torch.zeros_like(out)) * g).sum().backward()
    finally:
        os.environ.pop("PYSTL_B200_FORCE_THREADS", None)
    torch.cuda.synchronize()
    return x.grad.cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("case", GRAD_CASES, ids=[c["name"] for c in GRAD_CASES])
def test_golden_gradients(case):
    sig = case_signal(case)
    vals = GOLD[f"{case['name']}/values"]
    got = gpu_grad(case["spec"], case["sem"], case["kw"], sig, case_cotangent(case, vals.shape))
    check(got, GOLD[f"{case['name']}/grad"], case["sem"], case["name"], *grad_tol(case["sem"], case["kw"]))


@pytest.mark.parametrize("sem,kw", BWD_SEMS, ids=["classical", "smooth", "agm", "dgmsr", "dgmsr_p2"])
@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg3"])
def test_baseline_formulas_multi_tile(cfg, sem, kw):
    rng = np.random.default_rng(17)
    sig = mixed_signal(rng, 2, 900, 2)
    g = rng.uniform(0.25, 1.0, size=(2, 900)).astype(np.float32).astype(np.float64)
    want = oracle_grad(BASELINE_SPECS[cfg], sem, kw, sig, g)
    for threads in (0, 64):
        check(gpu_grad(BASELINE_SPECS[cfg], sem, kw, sig, g, threads), want, sem, f"{cfg}/{sem}/nt{threads}")


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
                got = gpu_grad(spec, sem, kw, sig, g, threads)
            except NotImplementedError:
                assert threads
                continue
            check(got, want, sem, f"seed{seed}/{sem}/nt{threads}/{spec}")


def test_mean_loss_gradient_full_batch_properties():
    """cfg2-shaped batch (scaled to 2048 traces): gradient of the mean robustness through the fused
    constant-cotangent path equals the explicit-cotangent path; sums of softmax weights are 1."""
    from pystl_b200 import capi

    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((2048, 4096, 2), generator=g, device="cuda")
    sem = create_semantics("smooth", backend="b200", temperature=1.0)
    phi = from_spec(BASELINE_SPECS["cfg2"])
    xg = x.clone().requires_grad_(True)
    sem.robustness_trace(phi, xg).mean().backward()
    cf = sem.compile(phi, 2)
    fused = torch.empty_like(x)
    rc = capi.lib().pystl_b200_eval_backward(
        cf.ptr(), x.data_ptr(), 2048, 4096, 2, None, 1.0 / (2048 * 4096), fused.data_ptr(),
        torch.cuda.current_stream().cuda_stream,
    )
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.allclose(fused, xg.grad, rtol=1e-5, atol=1e-12)
    # every output distributes exactly its cotangent over the inputs (the softmax weights of each
    # operator sum to one and |d pred / d x| = 1): the gradient entries sum to +-... in magnitude 1
    assert float(fused.abs().sum()) == pytest.approx(1.0, rel=1e-3)


def test_evaluate_with_grad_matches_reference_contract():
    """Formula.evaluate_with_grad(signal, sem, t) -> (value, grad[T, D]) like dgmsr.py:120-274."""
    rng = np.random.default_rng(5)
    sig = mixed_signal(rng, 1, 40, 2)[0]
    spec = ("eventually", (0, 2), ("always", (0, 2), ("and", [("pred", {0: 1.0}, -0.25), ("pred", {1: -1.0}, 0.55)])))
    phi = from_spec(spec)
    for sem_name, kw in (("dgmsr", {"eps": 1e-8, "p": 2}), ("smooth", {"temperature": 0.5}), ("classical", {})):
        sem = create_semantics(sem_name, backend="b200", **kw)
        value, grad = phi.evaluate_with_grad(sig, sem, t=3)
        g = np.zeros((1, 40))
        g[0, 3] = 1.0
        want = oracle_grad(spec, sem_name, kw, sig[None], g)[0]
        assert grad.shape == sig.shape and grad.dtype == np.float64
        check(grad, want, sem_name, f"evaluate_with_grad/{sem_name}")
