"""GPU tier: non-affine predicates (opaque ``fn(signal, t)`` callables, api.py:24,133-143).

The kernels read them as leaf channels that torch evaluates on the device (compiler.leaf_values);
values and gradients are checked against the float64 oracle, which calls ``fn`` per (trace, t)
like the reference does."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import dense
from pystl_b200 import create_semantics
from pystl_b200.formula import Always, And, Eventually, Not, Or, Predicate, Until, Interval

pytestmark = pytest.mark.gpu

SEMS = [("classical", {}), ("smooth", {"temperature": 0.5}), ("agm", {}), ("dgmsr", {"eps": 1e-8, "p": 1}), ("cumulative", {})]


def formula():
    dist = Predicate("inside", fn=lambda s, t: 1.5 - (s[t, 0] ** 2 + s[t, 1] ** 2) ** 0.5)      # indexes with t only
    wave = Predicate("wave", fn=lambda s, t: torch.sin(s[t, 1]) * s[t, 0] + 0.01 * t)               # computes with t
    lin = Predicate("x>0", fn=lambda s, t: s[t, 0])                                                  # affine: probed
    return Or(Always(dist, Interval(0, 9)), And(Eventually(wave, Interval(2, 7)), Not(lin)))


@pytest.mark.parametrize("sem,kw", SEMS, ids=[s for s, _ in SEMS])
def test_values_and_gradients(sem, kw):
    rng = np.random.default_rng(5)
    sig = (rng.uniform(-1, 1, size=(3, 1, 2)) + 0.6 * rng.standard_normal((3, 70, 2))).astype(np.float32)
    phi = formula()
    s = create_semantics(sem, backend="b200", **kw)
    cf = s.compile(phi, 2)
    assert len(cf.leaf_preds) == 2 and cf.signal_channels == (0,) and cf.n_channels == 3

    x = torch.tensor(sig, device="cuda", requires_grad=True)
    got = s.robustness_trace(phi, x)
    x64 = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
    want = dense.robustness_trace(phi, x64, dense.make(sem, **kw))
    if sem == "cumulative":
        got = torch.stack((got.pos, got.neg), 0)
    ok = torch.isfinite(want)
    assert torch.equal(torch.isfinite(got.detach().cpu()), ok)
    assert torch.allclose(got.detach().cpu().double()[ok], want.detach()[ok], rtol=1e-5, atol=5e-6)

    g = torch.tensor(rng.uniform(0.25, 1.0, size=tuple(want.shape)))
    (torch.nan_to_num(got) * g.to("cuda", torch.float32)).sum().backward()
    (torch.nan_to_num(want) * g).sum().backward()
    assert torch.allclose(x.grad.cpu().double(), x64.grad, rtol=2e-5, atol=2e-5)

    # the one-pass entry point chains through the leaves as well
    mean, grad = s.mean_robustness_and_grad(phi, torch.tensor(sig, device="cuda"))
    x2 = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
    w2 = dense.robustness_trace(phi, x2, dense.make(sem, **kw))
    w2 = w2.sum(0) if sem == "cumulative" else w2
    m2 = w2[torch.isfinite(w2)].mean()
    m2.backward()
    assert abs(float(mean) - float(m2)) < 2e-5 * (1 + abs(float(m2)))
    assert torch.allclose(grad.cpu().double(), x2.grad, rtol=2e-5, atol=2e-7)


def test_untraceable_predicate_is_loud():
    import math

    bad = Predicate("bad", fn=lambda s, t: math.sin(float(s[t, 0])))  # forces a host float: torch cannot map it
    s = create_semantics("classical", backend="b200")
    with pytest.raises(NotImplementedError):
        s.robustness_trace(Always(bad, Interval(0, 3)), torch.zeros(2, 20, 2, device="cuda"))


def test_assemble_kernel_matches_torch_cat():
    """``ops.assemble`` (pystl_b200_assemble_channels / _scatter_channels): the tensor the kernels read next to leaf columns."""
    from pystl_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(3)
    for (B, T, D), ch, L in (((3, 70, 4), (2, 0), 2), ((2, 513, 2), (1,), 1), ((1, 33, 8), (), 3), ((4, 257, 3), (0, 1, 2), 1),
                             ((2, 40, 40), tuple(range(0, 40)), 2), ((2, 40, 2), (0,), 17), ((2, 9, 70), (69, 3), 1), ((5, 1000, 5), (4, 1, 3), 4)):
        sig = torch.randn(B, T, D, device="cuda", generator=g, requires_grad=True)
        cols = [torch.randn(B, T, device="cuda", generator=g, requires_grad=True) for _ in range(L)]
        out = ops.assemble(sig, ch, cols)
        ref = torch.cat([sig.detach()[..., list(ch)]] + [c.detach()[..., None] for c in cols], -1)
        assert torch.equal(out.detach(), ref)
        w = torch.randn(out.shape, device="cuda", generator=g)
        (out * w).sum().backward()
        want = torch.zeros_like(sig)
        for i, d in enumerate(ch):
            want[..., d] = w[..., i]
        assert torch.equal(sig.grad, want)
        for j, c in enumerate(cols):
            assert torch.equal(c.grad, w[..., len(ch) + j])
    same = torch.randn(2, 5, 3, device="cuda")
    assert ops.assemble(same, (0, 1, 2), []) is same
