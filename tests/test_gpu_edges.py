"""GPU tier: the edges of the input space -- empty batches, horizons shorter than the windows,
outputs the reference raises on, one very long trace, and offsets beyond 2^31 elements."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import dense
from parity_util import SEM_KW, assert_parity
from pystl_b200 import create_semantics
from pystl_b200.formula import BASELINE_SPECS, from_spec

pytestmark = pytest.mark.gpu


def oracle(spec, sem, kw, sig64):
    out = dense.robustness_trace(from_spec(spec), torch.tensor(sig64, dtype=torch.float64), dense.make(sem, **kw))
    return out.numpy()


def got_trace(s, phi, x):
    out = s.robustness_trace(phi, x)
    if hasattr(out, "pos"):
        out = torch.stack((out.pos, out.neg), 0)
    return out.double().cpu().numpy()


@pytest.mark.parametrize("sem,kw", SEM_KW, ids=[f"{s}{k.get('p', '')}" for s, k in SEM_KW])
def test_empty_batch(sem, kw):
    s = create_semantics(sem, backend="b200", **kw)
    phi = from_spec(BASELINE_SPECS["cfg2"])
    x = torch.zeros(0, 64, 2, device="cuda")
    out = s.robustness_trace(phi, x)
    shape = tuple(out.pos.shape) if hasattr(out, "pos") else tuple(out.shape)
    assert shape == (0, 64)
    mean, grad = s.mean_robustness_and_grad(phi, x)
    assert tuple(grad.shape) == (0, 64, 2)


@pytest.mark.parametrize("sem,kw", SEM_KW, ids=[f"{s}{k.get('p', '')}" for s, k in SEM_KW])
@pytest.mark.parametrize("T", [1, 2, 7, 9, 23])
def test_horizon_shorter_than_the_windows(sem, kw, T):
    """Windows are clipped to the trace (api.py:49-54): a 1-sample trace still has a robustness."""
    rng = np.random.default_rng(T)
    sig = (rng.uniform(-1, 3, size=(3, 1, 2)) + 0.5 * rng.standard_normal((3, T, 2))).astype(np.float32).astype(np.float64)
    for cfg in ("cfg1", "cfg3"):
        s = create_semantics(sem, backend="b200", **kw)
        got = got_trace(s, from_spec(BASELINE_SPECS[cfg]), torch.tensor(sig, dtype=torch.float32, device="cuda"))
        assert_parity(got, oracle(BASELINE_SPECS[cfg], sem, kw, sig), sem, f"{cfg}/T{T}")


def test_outputs_the_reference_raises_on_are_nan_and_raise():
    """G[5,9] on a 7-sample trace: t = 0, 1 have a window, t >= 2 raise ValueError (api.py:218-221)."""
    spec = ("always", (5, 9), ("pred", {0: 1.0}, 0.0))
    s = create_semantics("classical", backend="b200")
    x = torch.randn(2, 7, 1, device="cuda")
    out = s.robustness_trace(from_spec(spec), x)
    assert torch.isfinite(out[:, :2]).all() and torch.isnan(out[:, 2:]).all()
    assert torch.equal(out[:, 0], x[:, 5:7, 0].min(dim=1).values)
    s.evaluate_formula(from_spec(spec), x[0], 1)
    with pytest.raises(ValueError):
        s.evaluate_formula(from_spec(spec), x[0], 2)
    with pytest.raises(IndexError):
        s.evaluate_formula(from_spec(spec), x[0], 7)
    # a window that never exists: every output is NaN, the mean is over zero outputs
    never = ("eventually", (9, 12), ("pred", {0: 1.0}, 0.0))
    assert torch.isnan(s.robustness_trace(from_spec(never), x)).all()
    # the gradient of the valid outputs only
    xg = x.clone().requires_grad_(True)
    torch.nan_to_num(s.robustness_trace(from_spec(spec), xg)).sum().backward()
    assert torch.isfinite(xg.grad).all() and (xg.grad[:, :5] == 0).all()


@pytest.mark.parametrize("sem,kw", [SEM_KW[0], SEM_KW[1], SEM_KW[2]], ids=["classical", "smooth", "cumulative"])
def test_one_long_trace(sem, kw):
    """B = 1, T = 2^18: hundreds of tiles of one trace; spot-checked through the slice identity
    (SURVEY.md appendix A1): out[t] only depends on signal[t : t + H + 1]."""
    T, H = 1 << 18, 264
    rng = np.random.default_rng(11)
    sig = rng.standard_normal((1, T, 4)).astype(np.float32)
    s = create_semantics(sem, backend="b200", **kw)
    got = got_trace(s, from_spec(BASELINE_SPECS["cfg4"]), torch.tensor(sig, device="cuda"))
    for t in (0, 1535, 1536, 100_003, T - 300, T - 2 * H):
        ref = oracle(BASELINE_SPECS["cfg4"], sem, kw, sig[:, t : t + H + 1].astype(np.float64))
        assert_parity(got[..., t : t + 1], ref[..., :1], sem, f"long/t{t}", scale=1.0)
    tail = oracle(BASELINE_SPECS["cfg4"], sem, kw, sig[:, T - 600 :].astype(np.float64))
    assert_parity(got[..., T - 600 :], tail, sem, "long/tail", scale=1.0)


def test_offsets_beyond_2_31_elements():
    """cfg5's shard: 131072 x 2048 x 8 = 2^31 floats (8.6 GB); first and last traces against the oracle."""
    free, _ = torch.cuda.mem_get_info()
    if free < 12 << 30:
        pytest.skip("needs 12 GB of device memory")
    B, T, D = 131072, 2048, 8
    x = torch.empty(B, T, D, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(3)
    for i in range(0, B, 16384):
        x[i : i + 16384].normal_(generator=gen)
    s = create_semantics("classical", backend="b200")
    out = s.robustness_trace(from_spec(BASELINE_SPECS["cfg5"]), x)
    for b in (0, 1, B // 2, B - 2, B - 1):
        ref = oracle(BASELINE_SPECS["cfg5"], "classical", {}, x[b : b + 1].double().cpu().numpy())
        assert_parity(out[b : b + 1].double().cpu().numpy(), ref, "classical", f"huge/b{b}")
