"""GPU tier: the CUDA library, called through the C ABI (pystl_b200.capi), against the oracle
and the reference's golden vectors.  Same corpus as the CPU emulator tier."""

from __future__ import annotations

import os

import numpy as np
import pytest
import torch

from golden_cases import CASES, case_signal
from parity_util import SEM_KW, assert_parity, mixed_signal, oracle_trace, random_spec
from pystl_b200 import create_semantics
from pystl_b200.formula import BASELINE_SPECS, from_spec

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _interpreter_tier(monkeypatch):
    """This file pins the INTERPRETER tier (the universal fallback); tests/test_gpu_fused.py runs
    the same kind of corpus through the fused tier."""
    monkeypatch.setenv("PYSTL_B200_TIER", "interp")

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
SUPPORTED = [c for c in CASES if not (c["spec_name"] == "weighted" and c["sem"] in ("agm", "dgmsr"))]


def gpu_trace(spec, sem: str, kw: dict, sig64: np.ndarray, threads: int = 0) -> np.ndarray:
    s = create_semantics(sem, backend="b200", **kw)
    x = torch.tensor(sig64, dtype=torch.float32, device="cuda")
    old = os.environ.pop("PYSTL_B200_FORCE_THREADS", None)
    if threads:
        os.environ["PYSTL_B200_FORCE_THREADS"] = str(threads)
    try:
        out = s.robustness_trace(from_spec(spec), x)
    finally:
        os.environ.pop("PYSTL_B200_FORCE_THREADS", None)
        if old is not None:
            os.environ["PYSTL_B200_FORCE_THREADS"] = old
    if sem == "cumulative":
        out = torch.stack((out.pos, out.neg), dim=-1)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("case", SUPPORTED, ids=[c["name"] for c in SUPPORTED])
def test_golden_cases(case):
    got = gpu_trace(case["spec"], case["sem"], case["kw"], case_signal(case))
    assert_parity(got, GOLD[f"{case['name']}/values"], case["sem"], case["name"])


@pytest.mark.parametrize("sem,kw", SEM_KW, ids=[f"{s}{''.join(f'_{v}' for v in k.values())}" for s, k in SEM_KW])
@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_baseline_formulas(cfg, sem, kw):
    rng = np.random.default_rng(7)
    D = {"cfg5": 8, "cfg4": 4}.get(cfg, 2)
    sig = mixed_signal(rng, 3, 700, D)
    want = oracle_trace(BASELINE_SPECS[cfg], sem, kw, sig)
    for threads in (0, 64):
        got = gpu_trace(BASELINE_SPECS[cfg], sem, kw, sig, threads)
        assert_parity(got, want, sem, f"{cfg}/{sem}/nt{threads}", scale=1.0)


@pytest.mark.parametrize("seed", range(24))
def test_random_formulas(seed):
    rng = np.random.default_rng(1000 + seed)
    D = int(rng.integers(1, 5))
    T = int(rng.integers(20, 260))
    spec = random_spec(rng, D, depth=3, max_b=40, allow_open=(seed % 4 == 0))
    sig = mixed_signal(rng, 2, T, D)
    for sem, kw in SEM_KW:
        want = oracle_trace(spec, sem, kw, sig)
        for threads in (0, 32):
            try:
                got = gpu_trace(spec, sem, kw, sig, threads)
            except NotImplementedError:
                assert threads and 8 * threads < T
                continue
            assert_parity(got, want, sem, f"seed{seed}/{sem}/nt{threads}/{spec}", scale=1.0)


def test_long_windows_van_herk():
    rng = np.random.default_rng(5)
    sig = mixed_signal(rng, 2, 900, 1)
    for a, b in [(0, 8), (1, 9), (3, 20), (7, 64), (8, 72), (5, 100), (0, 201), (13, 300), (2, 17), (6, 15)]:
        spec = ("always", (a, b), ("pred", {0: 1.0}, -0.25))
        for sem, kw in SEM_KW[:3]:
            want = oracle_trace(spec, sem, kw, sig)
            for threads in (0, 64):
                assert_parity(gpu_trace(spec, sem, kw, sig, threads), want, sem, f"G[{a},{b}]/{sem}/nt{threads}")


def test_smooth_extreme_dynamic_range_is_exact():
    rng = np.random.default_rng(3)
    sig = (rng.standard_normal((2, 300, 2)) * 5.0).astype(np.float32).astype(np.float64)
    kw = {"temperature": 0.01}
    for cfg in ("cfg1", "cfg2"):
        assert_parity(gpu_trace(BASELINE_SPECS[cfg], "smooth", kw, sig), oracle_trace(BASELINE_SPECS[cfg], "smooth", kw, sig), "smooth", cfg)


def test_every_cta_shape_and_unaligned_traces():
    """T chosen so that traces start at every 16-byte phase (TMA vs scalar staging), all CTA sizes."""
    rng = np.random.default_rng(21)
    spec = BASELINE_SPECS["cfg1"]
    for T in (999, 1000, 1001, 1002, 1003):
        sig = mixed_signal(rng, 5, T, 2)[..., :1].repeat(2, axis=-1)
        sig[..., 1] = rng.standard_normal((5, T)).astype(np.float32)
        for sem, kw in (SEM_KW[0], SEM_KW[1]):
            want = oracle_trace(spec, sem, kw, sig)
            for threads in (0, 32, 128, 256, 512, 1024):
                assert_parity(gpu_trace(spec, sem, kw, sig, threads), want, sem, f"T{T}/{sem}/nt{threads}")


def test_full_size_properties_cfg1():
    """BASELINE cfg1 at full size (1024 x 1000 x 2): a seeded sub-sample against the oracle, plus
    size-independent properties: G/F duality under negation, monotonicity in the window, shift
    invariance of the trace."""
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn((1024, 1000, 2), generator=g, device="cuda")
    sem = create_semantics("classical", backend="b200")
    phi = from_spec(BASELINE_SPECS["cfg1"])
    out = sem.robustness_trace(phi, x)
    assert out.shape == (1024, 1000) and bool(torch.isfinite(out).all())
    idx = [0, 1, 511, 1023]
    want = oracle_trace(BASELINE_SPECS["cfg1"], "classical", {}, x[idx].double().cpu().numpy())
    assert_parity(out[idx].cpu().numpy(), want, "classical", "cfg1 sample")
    # duality: G[a,b] p == -F[a,b] (-p), exactly
    G = from_spec(("always", (0, 10), ("pred", {0: 1.0}, 0.0)))
    F = from_spec(("not", ("eventually", (0, 10), ("not", ("pred", {0: 1.0}, 0.0)))))
    assert torch.equal(sem.robustness_trace(G, x), sem.robustness_trace(F, x))
    # a longer ALWAYS window can only lower the value
    G20 = from_spec(("always", (0, 20), ("pred", {0: 1.0}, 0.0)))
    assert bool((sem.robustness_trace(G20, x) <= sem.robustness_trace(G, x)).all())
    # time invariance: evaluating a suffix of the trace gives the suffix of the result (slice identity A1)
    assert torch.equal(sem.robustness_trace(phi, x[:, 300:].contiguous()), out[:, 300:])


def test_mean_reduction_and_host_tensors():
    rng = np.random.default_rng(2)
    sig = torch.tensor(mixed_signal(rng, 64, 333, 2), dtype=torch.float32)
    sem = create_semantics("smooth", backend="b200", temperature=0.7)
    phi = from_spec(("eventually", (2, 9), ("pred", {1: 1.0}, -1.0)))
    out = sem.robustness_trace(phi, sig)  # CPU tensor in -> CPU tensor out, computed on the GPU
    assert out.device.type == "cpu" and bool(torch.isnan(out[:, 331:]).all())
    mean = sem.mean_robustness(phi, sig.cuda())
    assert float(mean) == pytest.approx(float(out[:, :331].double().mean()), rel=1e-6)
    with pytest.raises(ValueError):
        sem.evaluate_formula(phi, sig[0], 332)
    assert float(sem.evaluate_formula(phi, sig[0], 5)) == pytest.approx(float(out[0, 5]))


def test_operator_interface_on_cuda_tensors():
    """The seven-method interface of semantics/base.py on CUDA tensors, driven by the reference-style
    recursive evaluator, equals the fused whole-formula kernel."""
    from pystl_b200.formula import _evaluate_recursive

    rng = np.random.default_rng(4)
    sig64 = mixed_signal(rng, 1, 12, 2)[0]
    x = torch.tensor(sig64, dtype=torch.float32, device="cuda")
    spec = ("or", [("always", (0, 3), ("pred", {0: 1.0}, 0.0)), ("until", (0, 4), ("pred", {0: 1.0}, 0.5), ("pred", {1: -1.0}, 1.0))])
    for sem_name, kw in (SEM_KW[0], SEM_KW[1], SEM_KW[3], SEM_KW[4]):
        sem = create_semantics(sem_name, backend="b200", **kw)
        phi = from_spec(spec)
        fused = sem.robustness_trace(phi, x)
        for t in (0, 3, 7):
            step = _evaluate_recursive(phi, x, sem, t)
            assert float(step) == pytest.approx(float(fused[t]), rel=2e-5, abs=1e-5)
