"""CPU tier of CT-STL: the oracle restatement (oracle/ctstl.py) against outputs of the unmodified reference class
``pystl.CtstlSemantics`` (tests/golden/ctstl_vectors.npz, made by tests/golden/make_ctstl_golden.py) and against the closed
forms the definitions give; the host logic of pystl_b200/ctstl.py (no GPU, no compute calls through the C ABI)."""

from __future__ import annotations

import os

import numpy as np
import pytest

from oracle import ctstl as oc

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ctstl_vectors.npz"))


def test_oracle_cumulative_matches_reference_vectors():
    y = GOLD["y"]
    for c, (a, b, tau, delta) in enumerate(GOLD["cum_cases"]):
        a, b = int(a), int(b)
        k = oc.tau_to_k(tau, delta)
        out, arg = oc.cumulative_trace(y, a, None if b < 0 else b, k)
        assert np.array_equal(out, GOLD["cum"][c]), (a, b, tau, delta)
        sel = arg >= 0
        assert np.array_equal(np.isfinite(out), sel)
        rows = np.nonzero(sel)
        assert np.array_equal(y[rows[0], arg[sel]], out[sel])  # the routed sample holds the value
        # ... and lies inside the output's window
        assert np.all(arg[sel] >= rows[1] + a)
        if b >= 0:
            assert np.all(arg[sel] <= rows[1] + b)


def test_oracle_until_matches_reference_vectors():
    y, l = GOLD["y"], GOLD["l"]
    for c, (a, b) in enumerate(GOLD["until_cases"]):
        out = oc.until_trace(l, y, int(a), None if b < 0 else int(b))
        assert np.array_equal(out, GOLD["until"][c], equal_nan=True), (a, b)


def test_oracle_closed_forms():
    rng = np.random.default_rng(5)
    v = rng.standard_normal(23)
    assert oc.temporal_cumulative(v, tau=1.0) == v.max()  # k = 1: Eventually
    assert oc.temporal_cumulative(v, tau=23.0) == v.min()  # k = n: Always
    assert oc.temporal_cumulative(v, tau=24.0) == -np.inf
    assert oc.temporal_cumulative([], tau=1.0) == -np.inf
    assert oc.temporal_cumulative(v, tau=2.5, delta=0.5) == np.sort(v)[-5]
    assert oc.tau_to_k(0.3, 0.1) == 3 and oc.tau_to_k(2.5, 1.0) == 3
    with pytest.raises(ValueError):
        oc.tau_to_k(0.0, 1.0)
    with pytest.raises(ValueError):
        oc.tau_to_k(1.0, 0.0)
    # exclusive prefix: candidate 0 is r[0] alone, candidate i ignores l[i]
    assert oc.temporal_until([-5.0, 9.0], [1.0, 2.0], 0, 0) == 1.0
    assert oc.temporal_until([3.0, -9.0], [0.0, 7.0], 1, 1) == 3.0
    with pytest.raises(ValueError):
        oc.temporal_until([1.0], [1.0], 1, 4)
    # stable-sort routing among equal samples
    assert oc.kth_arg([1.0, 2.0, 2.0, 2.0, 0.0], 1) == 3 and oc.kth_arg([1.0, 2.0, 2.0, 2.0, 0.0], 3) == 1


def test_host_logic_without_gpu():
    import pystl_b200
    from pystl_b200 import B200CtstlSemantics, CumulativeTime, create_semantics
    from pystl_b200.ctstl import _has_ct_node, tau_to_k
    from pystl_b200.formula import Always, Until, _evaluate_recursive, gt, lt

    with pytest.raises(KeyError):  # disabled in the reference too (tests/test_api.py:255-257)
        create_semantics("ctstl", backend="b200")
    with pytest.raises(ValueError):
        B200CtstlSemantics(delta=0.0)
    assert tau_to_k(2.5, 0.5) == 5
    x, y = gt(0, 0.0), lt(1, 0.5)
    assert not _has_ct_node(Always(x & ~y, (0, 3)))
    assert _has_ct_node(Always(Until(x, y, (0, 3)), (0, 3)))
    assert _has_ct_node(~CumulativeTime(x, (0, 5), tau=2.0))

    class OracleSem:  # the seven-method interface + temporal_cumulative on python floats (the reference's own recursion shape)
        def predicate(self, p, s, t): return float(p.fn(s, t))
        def boolean_not(self, v): return -v
        def boolean_and(self, vs, *, weights=None): return min(vs)
        def boolean_or(self, vs, *, weights=None): return max(vs)
        temporal_always, temporal_eventually = boolean_and, boolean_or
        def temporal_until(self, *, left_trace, right_trace, start, end, **_): return oc.temporal_until(left_trace, right_trace, start, end)
        def temporal_cumulative(self, vs, *, tau, delta=None): return oc.temporal_cumulative(vs, tau=tau, delta=1.0 if delta is None else delta)

    sig = np.random.default_rng(0).standard_normal((12, 2))
    phi = CumulativeTime(x, (1, 4), tau=2.0)
    for t in range(12):
        w = sig[t + 1 : t + 5, 0]
        want = np.sort(w)[-2] if w.size >= 2 else -np.inf
        assert _evaluate_recursive(phi, sig, OracleSem(), t) == want
    assert set(pystl_b200.__all__) >= {"B200CtstlSemantics", "CumulativeTime"}
