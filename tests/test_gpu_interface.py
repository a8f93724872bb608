"""GPU tier: the drop-in boundary beyond whole-formula evaluation.

* the reference's seven-method operator interface (semantics/base.py:18-75) driven by the reference's
  own recursion (api.py:142-302, restated in ``pystl_b200.formula._evaluate_recursive``) -- windows
  longer than 8 samples, weights, Until under all five semantics including Cumulative's
  ``temporal_until`` on ``(pos, neg)`` lists (cumulative.py:74-103);
* ``torch.library`` registration: ``torch.func.vmap / grad`` and ``torch.compile`` through the
  ``pystl_b200::robustness`` op (the reference's precedent is
  ``jax.vmap`` in tests/test_jax_semantics.py:310-368);
* predicates whose probed affine form is wrong on the real signal fall back to leaf channels.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import scalar
from parity_util import ATOL, RTOL, SEM_KW, mixed_signal
from pystl_b200 import create_semantics
from pystl_b200.formula import BASELINE_SPECS, Always, Interval, Predicate, _evaluate_recursive, from_spec

pytestmark = pytest.mark.gpu

W12 = [0.5, 1.0, 2.0, 1.5, 1.0, 0.25, 3.0, 1.0, 0.0, 2.0, 1.0, 0.75]
SPECS = {
    "long_windows": ("and", [("always", (0, 50), ("pred", {0: 1.0}, 0.25)), ("eventually", (2, 21), ("pred", {1: -1.0}, 0.5))]),
    "nested": ("always", (0, 12), ("or", [("eventually", (0, 9), ("pred", {0: 1.0}, 0.0)), ("not", ("pred", {1: 1.0}, -2.0))])),
    "until": ("until", (1, 14), ("pred", {0: 1.0}, 0.5), ("pred", {1: 1.0}, -0.5)),
    "until_open": ("until", (0, None), ("pred", {0: 1.0}, 0.5), ("pred", {1: 1.0}, -0.5)),
    "weighted": ("or", [("always", (0, 11), ("pred", {0: 1.0}, -0.25), W12),
                        ("until", (1, 6), ("pred", {0: 1.0}, 0.0), ("pred", {1: 1.0}, -0.5), [1.0, 2.0, 1.5, 1.0, 0.5, 2.5, 1.0],
                         [1.0, 1.2, 0.8, 1.4, 2.0, 0.6], [1.5, 0.5])], [1.0, 2.0]),
    "cfg4": ("always", (0, 20), ("until", (0, 9), ("pred", {0: 1.0}, 0.0), ("pred", {1: 1.0}, -1.0))),
}


def _close(got: float, want: float, what: str, scale: float = 1.0) -> None:
    assert abs(got - want) <= scale * (ATOL + RTOL * abs(want)), (what, got, want)


@pytest.mark.parametrize("name", sorted(SPECS))
@pytest.mark.parametrize("sem_name,kw", SEM_KW[:5], ids=[s for s, _ in SEM_KW[:5]])
def test_reference_recursion_through_the_operator_interface(name, sem_name, kw):
    rng = np.random.default_rng(11)
    sig64 = mixed_signal(rng, 2, 64, 2)[1]
    x = torch.tensor(sig64, dtype=torch.float32, device="cuda")
    phi = from_spec(SPECS[name])
    sem = create_semantics(sem_name, backend="b200", **kw)
    ref = scalar.make(sem_name, **kw)
    for t in (0, 7, 40):
        want = scalar.evaluate(phi, sig64, ref, t)
        got = _evaluate_recursive(phi, x, sem, t)
        if sem_name == "cumulative":
            _close(float(got.pos), want[0], f"{name}/{sem_name}/t{t}/pos")
            _close(float(got.neg), want[1], f"{name}/{sem_name}/t{t}/neg")
        elif sem_name == "classical":
            assert np.float32(float(got)) == np.float32(want), (name, t)
        else:
            _close(float(got), float(want), f"{name}/{sem_name}/t{t}")


def test_operator_interface_is_differentiable():
    rng = np.random.default_rng(12)
    sig64 = mixed_signal(rng, 1, 40, 2)[0]
    phi = from_spec(SPECS["long_windows"])
    for sem_name, kw in (SEM_KW[1], SEM_KW[2], SEM_KW[3]):
        sem = create_semantics(sem_name, backend="b200", **kw)
        x = torch.tensor(sig64, dtype=torch.float32, device="cuda", requires_grad=True)
        v = _evaluate_recursive(phi, x, sem, 0)
        v = v.pos + v.neg if sem_name == "cumulative" else v
        (g1,) = torch.autograd.grad(v, x)
        x2 = torch.tensor(sig64, dtype=torch.float32, device="cuda", requires_grad=True)
        w = sem.evaluate_formula(phi, x2, 0)
        w = w.pos + w.neg if sem_name == "cumulative" else w
        (g2,) = torch.autograd.grad(w, x2)
        assert torch.allclose(g1, g2, rtol=2e-5, atol=2e-5), (sem_name, float((g1 - g2).abs().max()))


# ---------------------------------------------------------------------------- torch.library registration
@pytest.mark.parametrize("sem_name,kw", SEM_KW[:5], ids=[s for s, _ in SEM_KW[:5]])
def test_vmap_and_functional_grad(sem_name, kw):
    """vmap over traces == the batched call (the reference has no batch axis: jax.vmap / torch.func.vmap
    over ``phi.evaluate`` is how its users batch)."""
    rng = np.random.default_rng(13)
    x = torch.tensor(mixed_signal(rng, 6, 300, 2), dtype=torch.float32, device="cuda")
    phi = from_spec(BASELINE_SPECS["cfg2"])
    sem = create_semantics(sem_name, backend="b200", **kw)

    def value(sig):  # one trace [T, D] -> scalar at t = 0
        r = sem.robustness_trace(phi, sig)
        r = r.pos + r.neg if sem_name == "cumulative" else r
        return r[0]

    batched = sem.robustness_trace(phi, x)
    batched = batched.pos + batched.neg if sem_name == "cumulative" else batched
    got = torch.func.vmap(value)(x)
    assert torch.equal(got, batched[:, 0])
    # per-trace gradients: vmap(grad(f)) == grad of the sum over the batch
    gv = torch.func.vmap(torch.func.grad(value))(x)
    xr = x.clone().requires_grad_(True)
    r = sem.robustness_trace(phi, xr)
    r = r.pos + r.neg if sem_name == "cumulative" else r
    r[:, 0].sum().backward()
    assert torch.allclose(gv, xr.grad, rtol=1e-6, atol=1e-7)


def test_torch_compile_traces_through_the_op():
    rng = np.random.default_rng(14)
    x = torch.tensor(mixed_signal(rng, 4, 500, 2), dtype=torch.float32, device="cuda")
    phi = from_spec(BASELINE_SPECS["cfg2"])
    sem = create_semantics("smooth", backend="b200", temperature=0.5)

    def loss(sig):
        return (sem.robustness_trace(phi, sig * 1.5) ** 2).mean()

    xe = x.clone().requires_grad_(True)
    le = loss(xe)
    le.backward()
    # aot_eager: dynamo + AOT autograd with the fake kernels and the registered autograd formula, no code generation
    compiled = torch.compile(loss, backend="aot_eager", fullgraph=True)
    xc = x.clone().requires_grad_(True)
    lc = compiled(xc)
    lc.backward()
    assert torch.allclose(lc, le) and torch.allclose(xc.grad, xe.grad, rtol=1e-6, atol=1e-8)


def test_one_pass_op_matches_the_two_pass_path():
    from pystl_b200 import ops

    rng = np.random.default_rng(15)
    x = torch.tensor(mixed_signal(rng, 8, 700, 2), dtype=torch.float32, device="cuda")
    phi = from_spec(BASELINE_SPECS["cfg3"])
    sem = create_semantics("agm", backend="b200")
    cf = sem.compile(phi, 2)
    out, acc, grad = torch.ops.pystl_b200.value_and_grad(x, 0.125, ops.handle_of(cf, 1))
    xr = x.clone().requires_grad_(True)
    r = sem.robustness_trace(phi, xr)
    (0.125 * r.sum()).backward()
    assert torch.allclose(out[..., 0], r.detach(), rtol=1e-6, atol=1e-7)
    assert abs(float(acc[0]) - float(r.double().sum())) < 1e-3
    assert torch.allclose(grad, xr.grad, rtol=2e-5, atol=2e-6)


# ---------------------------------------------------------------------------- probed predicates on real signals
def test_predicate_that_fools_the_probe_is_demoted_to_a_leaf_channel():
    late = torch.tensor([0.0] * 100 + [1.0] * 100)
    latenp = late.numpy()

    def fn(s, t):  # depends on t far beyond the probe's rows; works on numpy rows and on torch traces
        if isinstance(s, np.ndarray):
            return 0.5 - latenp[t] + s[t, 0]
        return 0.5 - late.to(s.device)[t] + s[t, 0]

    p = Predicate("late", fn=fn)
    phi = Always(p, Interval(0, 3))
    rng = np.random.default_rng(16)
    sig = mixed_signal(rng, 3, 200, 2)
    x = torch.tensor(sig, dtype=torch.float32, device="cuda")
    for sem_name, kw in (SEM_KW[0], SEM_KW[1]):
        sem = create_semantics(sem_name, backend="b200", **kw)
        got = sem.robustness_trace(phi, x).cpu().numpy()
        ref = scalar.make(sem_name, **kw)
        for b, t in ((0, 0), (1, 96), (1, 97), (2, 150), (2, 196)):
            want = scalar.evaluate(phi, sig[b], ref, t)
            assert abs(got[b, t] - want) <= 2e-5 + 2e-5 * abs(want), (sem_name, b, t, got[b, t], want)
