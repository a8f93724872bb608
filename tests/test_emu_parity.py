"""CPU tier: the kernel's CTA body (tile_eval.h) + lowering, run by the host emulator
(tests/emu/emu.cpp), against the oracle.  Checks tiling, halos, van Herk chunk
arithmetic, clipping and NaN placement without a GPU.  The GPU tier
(test_gpu_parity.py) runs the same cases through the CUDA library."""

from __future__ import annotations

import numpy as np
import pytest

import emu_lib
from golden_cases import CASES, case_signal
from parity_util import SEM_KW, assert_parity, mixed_signal, oracle_trace, random_spec
from pystl_b200.formula import BASELINE_SPECS, from_spec

GOLD = np.load(__file__.replace("test_emu_parity.py", "golden/reference_vectors.npz"))
SUPPORTED = [c for c in CASES if not (c["spec_name"] == "weighted" and c["sem"] in ("agm", "dgmsr"))]


@pytest.mark.parametrize("case", SUPPORTED, ids=[c["name"] for c in SUPPORTED])
@pytest.mark.parametrize("threads", [0, 8, 16])
def test_golden_cases(case, threads):
    """Emulated kernel vs the reference's own outputs (golden vectors), single- and multi-tile."""
    sig = case_signal(case)
    got = emu_lib.eval_forward(from_spec(case["spec"]), case["sem"], sig, force_threads=threads, **case["kw"])
    assert_parity(got, GOLD[f"{case['name']}/values"], case["sem"], case["name"])


@pytest.mark.parametrize("sem,kw", SEM_KW, ids=[f"{s}{''.join(f'_{v}' for v in k.values())}" for s, k in SEM_KW])
@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg3", "cfg5"])
def test_baseline_formulas_multi_tile(cfg, sem, kw):
    """The BASELINE formulas with their real windows on traces long enough to be unclipped,
    cut into several tiles (forced 64-thread CTAs = 512-sample tiles)."""
    rng = np.random.default_rng(7)
    D = {"cfg5": 8}.get(cfg, 2)
    sig = mixed_signal(rng, 3, 700, D)
    want = oracle_trace(BASELINE_SPECS[cfg], sem, kw, sig)
    got = emu_lib.eval_forward(from_spec(BASELINE_SPECS[cfg]), sem, sig, force_threads=64, **kw)
    assert_parity(got, want, sem, f"{cfg}/{sem}")


@pytest.mark.parametrize("sem,kw", SEM_KW[:3] + SEM_KW[3:5], ids=["classical", "smooth", "cumulative", "agm", "dgmsr"])
def test_cfg4_nested_until(sem, kw):
    rng = np.random.default_rng(11)
    sig = mixed_signal(rng, 2, 600, 4)
    want = oracle_trace(BASELINE_SPECS["cfg4"], sem, kw, sig)
    got = emu_lib.eval_forward(from_spec(BASELINE_SPECS["cfg4"]), sem, sig, force_threads=64, **kw)
    assert_parity(got, want, sem, f"cfg4/{sem}", scale=1.0)


@pytest.mark.parametrize("seed", range(24))
def test_random_formulas(seed):
    """Random nested formulas, every semantics, random tile size; windows up to 40 samples."""
    rng = np.random.default_rng(1000 + seed)
    D = int(rng.integers(1, 5))
    T = int(rng.integers(20, 260))
    spec = random_spec(rng, D, depth=3, max_b=40, allow_open=(seed % 4 == 0))
    sig = mixed_signal(rng, 2, T, D)
    for sem, kw in SEM_KW:
        want = oracle_trace(spec, sem, kw, sig)
        for threads in (0, int(rng.choice([16, 32, 64]))):
            try:
                got = emu_lib.eval_forward(from_spec(spec), sem, sig, force_threads=threads, **kw)
            except NotImplementedError:
                # only legal reason: an open-ended interval on a trace longer than the forced tile
                assert threads and 8 * threads < T
                continue
            # sums of up to ~40 cumulative terms: absolute error grows with the magnitude summed
            assert_parity(got, want, sem, f"seed{seed}/{sem}/nt{threads}/{spec}", scale=1.0)


def test_long_windows_van_herk():
    """Windows of 9 .. 300 samples, all alignments of a and b modulo the 8-sample chunk."""
    rng = np.random.default_rng(5)
    sig = mixed_signal(rng, 2, 900, 1)
    for a, b in [(0, 8), (1, 9), (3, 20), (7, 64), (8, 72), (5, 100), (0, 201), (13, 300), (2, 17), (6, 15)]:
        spec = ("always", (a, b), ("pred", {0: 1.0}, -0.25))
        for sem, kw in SEM_KW[:3]:
            want = oracle_trace(spec, sem, kw, sig)
            for threads in (0, 64):
                got = emu_lib.eval_forward(from_spec(spec), sem, sig, force_threads=threads, **kw)
                assert_parity(got, want, sem, f"G[{a},{b}]/{sem}/nt{threads}")


def test_smooth_extreme_dynamic_range_is_exact():
    """temperature so small that exp(-v/tau) spans hundreds of orders of magnitude inside one tile:
    the block-shifted fast path must hand over to the exact one."""
    rng = np.random.default_rng(3)
    sig = (rng.standard_normal((2, 300, 2)) * 5.0).astype(np.float32).astype(np.float64)
    kw = {"temperature": 0.01}
    for cfg in ("cfg1", "cfg2"):
        want = oracle_trace(BASELINE_SPECS[cfg], "smooth", kw, sig)
        got = emu_lib.eval_forward(from_spec(BASELINE_SPECS[cfg]), "smooth", sig, **kw)
        assert_parity(got, want, "smooth", cfg)


def test_mean_accumulator_and_shapes():
    rng = np.random.default_rng(9)
    sig = mixed_signal(rng, 3, 100, 2)
    spec = ("eventually", (2, 9), ("pred", {1: 1.0}, -1.0))
    got, acc = emu_lib.eval_forward(from_spec(spec), "classical", sig, mean=True)
    assert np.isnan(got[:, 98:]).all() and not np.isnan(got[:, :98]).any()
    assert acc[0] == pytest.approx(float(got[:, :98].astype(np.float64).sum()), rel=1e-6)
    # empty batch / empty trace are no-ops
    assert emu_lib.eval_forward(from_spec(spec), "classical", np.zeros((0, 10, 2), np.float32)).shape == (0, 10)


def test_error_classes():
    from pystl_b200.formula import Always, And, Predicate, gt

    with pytest.raises(ValueError):  # channel out of range
        emu_lib.eval_forward(gt(3, 0.0), "classical", np.zeros((1, 4, 2), np.float32))
    with pytest.raises(ValueError):  # agm.py:31-34
        emu_lib.eval_forward(And(gt(0, 0.0), gt(1, 0.0), weights=[1.0, -1.0]), "agm", np.zeros((1, 4, 2), np.float32))
    with pytest.raises(ValueError):  # agm.py:29-30
        emu_lib.eval_forward(And(gt(0, 0.0), gt(1, 0.0), weights=[1.0]), "dgmsr", np.zeros((1, 4, 2), np.float32))
    with pytest.raises(NotImplementedError):  # opaque, non-affine predicate
        emu_lib.eval_forward(Predicate("sq", fn=lambda s, t: s[t, 0] ** 2), "classical", np.zeros((1, 4, 2), np.float32))
    # an opaque but affine lambda (the reference's own fixtures) is recognised
    p1 = Predicate("speed_ok", fn=lambda s, t: 0.6 - s[t, 0])
    out = emu_lib.eval_forward(Always(p1, (0, 2)), "classical", np.array([[[0.2, 0.8], [0.3, 0.6], [0.5, 0.4], [0.7, 0.3]]]))
    assert out[0, 0] == pytest.approx(0.1, abs=1e-6)
