"""GPU tier: the reference's OWN test scenarios, run with ``backend="b200"``.

Ports of ``/root/reference/tests/test_api.py:129-338`` and
``tests/test_torch_semantics.py:41-151`` (the 4 x 2 fixture signal, predicates
``0.6 - s[t, 0]`` and ``s[t, 1] - 0.2``, ``phi.evaluate(signal, sem, t)``): same formulas, same
call, the backend keyword changed.  Expected values come from the pinned oracle
(``oracle/scalar.py``, the NumPy restatement checked against the reference's golden vectors),
compared at fp32 accuracy (abs 2e-6 where the reference uses 1e-12 on float64)."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import scalar
from pystl_b200 import And, Interval, Or, Predicate, create_semantics

pytestmark = pytest.mark.gpu

SIGNAL = np.array([[0.2, 0.8], [0.3, 0.6], [0.5, 0.4], [0.7, 0.3]], dtype=float)
ABS = 2e-6


def predicates():
    p1 = Predicate("speed_ok", fn=lambda s, t: 0.6 - s[t, 0])
    p2 = Predicate("alt_ok", fn=lambda s, t: s[t, 1] - 0.2)
    return p1, p2


def both(phi, sem_name, t, **kw):
    """(b200 value, oracle value) of phi.evaluate(signal, sem, t) -- the reference's call."""
    sem = create_semantics(sem_name, backend="b200", **kw)
    got = phi.evaluate(torch.tensor(SIGNAL, dtype=torch.float32, device="cuda"), sem, t=t)
    want = scalar.evaluate(phi, SIGNAL, scalar.make(sem_name, **kw), t)
    if sem_name == "cumulative":
        return (float(got.pos), float(got.neg)), (float(want[0]), float(want[1]))
    return float(got), float(want)


def test_boolean_and_temporal_always_with_classic_semantics():  # test_api.py:129-139
    p1, p2 = predicates()
    phi = (p1 & p2).always(Interval(0, 2))
    got, want = both(phi, "classical", 0)
    expected = min(min(0.6 - SIGNAL[t, 0], SIGNAL[t, 1] - 0.2) for t in [0, 1, 2])
    assert want == pytest.approx(expected, abs=1e-12)
    # (the fixture's 0.2, 0.6, ... are not float32 numbers: bit-exactness is tested on fp32 inputs elsewhere)
    assert got == pytest.approx(expected, abs=ABS)


def test_eventually_and_until_with_classic_semantics():  # test_api.py:142-167
    p1, p2 = predicates()
    got, want = both(p1.eventually((1, 3)), "classical", 0)
    assert want == pytest.approx(max(0.6 - SIGNAL[t, 0] for t in [1, 2, 3]), abs=1e-12)
    assert got == pytest.approx(want, abs=ABS)
    got, want = both(p1.until(p2, interval=(0, 3)), "classical", 0)
    l = 0.6 - SIGNAL[:, 0]
    r = SIGNAL[:, 1] - 0.2
    best = max(min(float(np.min(l[: i + 1])), float(r[i])) for i in range(4))
    assert want == pytest.approx(best, abs=1e-12)
    assert got == pytest.approx(best, abs=ABS)


def test_smooth_and_or_are_soft_versions():  # test_api.py:170-189
    p1, p2 = predicates()
    tau = 0.25
    for phi, sign in ((p1 & p2, -1.0), (p1 | p2, 1.0)):
        got, want = both(phi, "smooth", 0, temperature=tau)
        vals = np.array([0.4, 0.6])
        expected = sign * tau * np.log(np.sum(np.exp(sign * vals / tau)))
        assert want == pytest.approx(expected, abs=1e-12)
        assert got == pytest.approx(expected, abs=ABS)


def test_empty_window_raises_value_error():  # test_api.py:208-213
    p1, _ = predicates()
    sem = create_semantics("classical", backend="b200")
    with pytest.raises(ValueError):
        p1.always((5, 6)).evaluate(torch.tensor(SIGNAL, dtype=torch.float32, device="cuda"), sem, t=0)
    with pytest.raises(ValueError):
        Interval(3, 2)


def test_cumulative_until_and_eventually_is_sum():  # test_api.py:216-252
    p1, p2 = predicates()
    got, want = both(p1.until(p2, interval=(0, 3)), "cumulative", 0)
    assert got == pytest.approx(want, abs=ABS)
    got, want = both(p1.eventually((0, 3)), "cumulative", 0)
    pos = sum(max(0.0, 0.6 - SIGNAL[t, 0]) for t in range(4))
    neg = sum(min(0.0, 0.6 - SIGNAL[t, 0]) for t in range(4))
    assert want == pytest.approx((pos, neg), abs=1e-12)
    assert got == pytest.approx((pos, neg), abs=ABS)


@pytest.mark.parametrize("sem_name,kw", [("agm", {}), ("dgmsr", {"eps": 1e-8, "p": 1}), ("dgmsr", {"eps": 1e-8, "p": 2})])
def test_weighted_boolean_temporal_and_until(sem_name, kw):  # test_api.py:284-338, test_torch_semantics.py:104-151
    p1, p2 = predicates()
    cases = [
        (And(p1, p2, weights=[1.0, 3.0]), 0),
        (Or(p1, p2, weights=[1.0, 3.0]), 3),
        (p1.always((0, 2)), 0),
        (p1.eventually((0, 3)), 0),
        ((~p2).eventually((0, 3)), 0),
        (p1.always((0, 3), weights=[1.0, 2.0, 0.5, 1.5]), 0),
        (p1.until(p2, interval=(0, 3), weights_left=[1.0, 2.0, 1.5, 1.0], weights_right=[1.0, 1.2, 0.8, 1.4], weights_pair=(1.0, 1.3)), 0),
        (p1.until(p2, interval=(1, 3), weights_left=[1.0, 2.0, 1.5, 1.0], weights_right=[1.0, 1.2, 0.8], weights_pair=(2.0, 1.0)), 1),
    ]
    for phi, t in cases:
        got, want = both(phi, sem_name, t, **kw)
        assert got == pytest.approx(want, abs=ABS), (sem_name, t, phi)


def test_torch_backend_matrix_classic_and_smooth():  # test_torch_semantics.py:41-101
    p1, p2 = predicates()
    formulas = [p1 & p2, p1 | p2, (p1 & p2).always((0, 2)), p1.eventually((1, 3)), p1.until(p2, interval=(0, 3))]
    for sem_name, kw in (("classical", {}), ("smooth", {"temperature": 0.25})):
        for phi in formulas:
            for t in (0, 1):
                got, want = both(phi, sem_name, t, **kw)
                assert got == pytest.approx(want, abs=ABS), (sem_name, t)


def test_smooth_autograd_gives_finite_signal_shaped_gradient():  # test_torch_semantics.py:154-167
    p1, p2 = predicates()
    sem = create_semantics("smooth", backend="b200", temperature=0.25)
    sig = torch.tensor(SIGNAL, dtype=torch.float32, device="cuda", requires_grad=True)
    out = ((p1 & p2).always((0, 2))).evaluate(sig, sem, t=0)
    out.backward()
    assert sig.grad is not None and sig.grad.shape == sig.shape and bool(torch.isfinite(sig.grad).all())
