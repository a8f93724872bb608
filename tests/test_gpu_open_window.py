"""GPU tier: ``Always / Eventually`` without an interval end (api.py:49-54: the window runs to the horizon) on traces
LONGER than one tile -- the suffix-scan kernels of csrc/scan.cu behind ``compiler.OpenLeaf``.

1. against the float64 oracle (values and gradients) with the scan path forced at sizes the dense oracle finishes;
2. the scan path against the in-kernel path (one tile) on the same inputs;
3. at T = 20000 (beyond any tile) against float64 suffix scans written with torch cumulative ops -- the closed forms of
   classic.py:39-57 (min / max), smooth.py:47-65 (+-tau logsumexp(+-x / tau)) -- and through the properties the
   domain offers (a constant trace, monotone suffixes, the last valid sample equals the operand).
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import dense
from pystl_b200 import create_semantics
from pystl_b200.compiler import OpenLeaf, OpenUntilLeaf
from pystl_b200.formula import Always, And, Eventually, Interval, Not, Or, Until, affine_predicate

pytestmark = pytest.mark.gpu

SEMS = [("classical", {}), ("smooth", {"temperature": 0.5}), ("smooth", {"temperature": 0.05}), ("agm", {}), ("dgmsr", {"eps": 1e-8, "p": 1}),
        ("dgmsr", {"eps": 1e-8, "p": 2})]
IDS = ["classical", "smooth0.5", "smooth0.05", "agm", "dgmsr1", "dgmsr2"]


def preds():
    x = affine_predicate("x", {0: 1.0}, 0.1)
    y = affine_predicate("y", {1: -1.0, 0: 0.25}, 0.3)
    return x, y


def preds32():
    # coefficients that are float32 numbers: the classical kernels are bit-exact against float32(float64 reference)
    return affine_predicate("x", {0: 1.0}, 0.125), affine_predicate("y", {1: -1.0, 0: 0.25}, 0.375)


def formulas():
    x, y = preds()
    return {
        "G_open": Always(x, Interval(0, None)),
        "F_open_a": Eventually(y, Interval(7, None)),
        "mixed": And(Always(x, Interval(3, None)), Eventually(y, Interval(0, 10))),
        "under_window": Eventually(Always(Or(x, Not(y)), Interval(2, None)), Interval(0, 5)),
        "nested": Always(Eventually(x, Interval(0, None)), Interval(2, None)),
        "over_nary": Eventually(And(x, Not(y)), Interval(5, None)),
    }


def signal(B, T, seed):
    rng = np.random.default_rng(seed)
    base = rng.uniform(-0.5, 0.5, size=(B, 1, 2))
    return (base + 0.4 * rng.standard_normal((B, T, 2))).astype(np.float32)


def tol(sem):
    # classical values are exact; the others accumulate a few hundred terms per output in fp32
    return dict(rtol=1e-5, atol=1e-6) if sem == "classical" else dict(rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("name", list(formulas()))
@pytest.mark.parametrize("sem,kw", SEMS, ids=IDS)
def test_scan_path_against_the_oracle(sem, kw, name):
    T = 300
    phi = formulas()[name]
    sig = signal(3, T, 11)
    s = create_semantics(sem, backend="b200", **kw)
    s.force_open_leaf = True
    x = torch.tensor(sig, device="cuda", requires_grad=True)
    cf = s._checked_compile(phi, x.detach())
    assert any(isinstance(p, OpenLeaf) for p in cf.leaf_preds)
    got = s.robustness_trace(phi, x)
    x64 = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
    want = dense.robustness_trace(phi, x64, dense.make(sem, **kw))
    ok = torch.isfinite(want)
    assert torch.equal(torch.isfinite(got.detach().cpu()), ok)
    assert torch.allclose(got.detach().cpu().double()[ok], want.detach()[ok], **tol(sem))

    g = torch.tensor(np.random.default_rng(3).uniform(0.25, 1.0, size=tuple(want.shape)))
    (torch.nan_to_num(got) * g.to("cuda", torch.float32)).sum().backward()
    (torch.nan_to_num(want) * g).sum().backward()
    # (dgmsr: the reference's formulation multiplies the window's entries, _dgmsr.py:96-102 -- over a few hundred samples
    # the product underflows in float64 and autograd returns NaN for the whole trace; the kernels work in the log domain
    # and are compared on the traces the oracle can differentiate)
    fin = torch.isfinite(x64.grad).all(dim=(1, 2))
    assert bool(fin.any()) and (sem == "dgmsr" or bool(fin.all()))
    scale = float(x64.grad[fin].abs().max()) + 1e-12
    assert bool(torch.isfinite(x.grad).all())
    assert torch.allclose(x.grad.cpu().double()[fin], x64.grad[fin], rtol=5e-4, atol=2e-5 * scale)

    mean, grad = s.mean_robustness_and_grad(phi, torch.tensor(sig, device="cuda"))
    x2 = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
    w2 = dense.robustness_trace(phi, x2, dense.make(sem, **kw))
    m2 = w2[torch.isfinite(w2)].mean()
    m2.backward()
    assert abs(float(mean) - float(m2.detach())) < 1e-4 * (1 + abs(float(m2.detach())))
    fin = torch.isfinite(x2.grad).all(dim=(1, 2))
    assert torch.allclose(grad.cpu().double()[fin], x2.grad[fin], rtol=5e-4, atol=2e-5 * (float(x2.grad[fin].abs().max()) + 1e-12))


@pytest.mark.parametrize("sem,kw", SEMS, ids=IDS)
def test_scan_path_equals_the_in_kernel_path(sem, kw):
    T = 2048  # one tile for the value AND the gradient kernel of every semantics: both paths exist
    sig = torch.tensor(signal(4, T, 5), device="cuda")
    for name, phi in formulas().items():
        a = create_semantics(sem, backend="b200", **kw)
        b = create_semantics(sem, backend="b200", **kw)
        b.force_open_leaf = True
        xa = sig.clone().requires_grad_(True)
        xb = sig.clone().requires_grad_(True)
        ra, rb = a.robustness_trace(phi, xa), b.robustness_trace(phi, xb)
        ok = torch.isfinite(ra)
        assert torch.equal(ok, torch.isfinite(rb)), name
        assert torch.allclose(ra[ok], rb[ok], **tol(sem)), (name, float((ra[ok] - rb[ok]).abs().max()))
        torch.nan_to_num(ra).sum().backward()
        torch.nan_to_num(rb).sum().backward()
        scale = float(xa.grad.abs().max()) + 1e-12
        assert torch.allclose(xa.grad, xb.grad, rtol=2e-3, atol=1e-4 * scale), (name, float((xa.grad - xb.grad).abs().max()), scale)


def _suffix(y64, sem, kw, mode):
    """float64 closed forms of the open-ended window of a value trace y[B, T] (a = 0)."""
    r = torch.flip(y64, dims=(1,))
    if sem == "classical":
        c = torch.cummin(r, 1).values if mode == 0 else torch.cummax(r, 1).values
    else:
        tau = kw["temperature"]
        c = -tau * torch.logcumsumexp(-r / tau, 1) if mode == 0 else tau * torch.logcumsumexp(r / tau, 1)
    return torch.flip(c, dims=(1,))


@pytest.mark.parametrize("sem,kw", [SEMS[0], SEMS[1], SEMS[2]], ids=IDS[:3])
@pytest.mark.parametrize("a", [0, 37])
def test_long_traces_against_closed_forms(sem, kw, a):
    B, T = 5, 20000
    x, _ = preds()
    sig = signal(B, T, 9)
    s = create_semantics(sem, backend="b200", **kw)
    for mode, phi in ((0, Always(x, Interval(a, None))), (1, Eventually(x, Interval(a, None)))):
        xs = torch.tensor(sig, device="cuda", requires_grad=True)
        got = s.robustness_trace(phi, xs)
        assert s.plan(phi, T, 2)["start_slack"] == a
        x64 = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
        y64 = x64[..., 0] + 0.1
        full = _suffix(y64, sem, kw, mode)  # value of the window starting at u
        want = torch.full_like(full, float("nan"))
        want[:, : T - a] = full[:, a:]
        ok = torch.isfinite(want)
        assert torch.equal(torch.isfinite(got.detach().cpu()), ok)
        assert torch.allclose(got.detach().cpu().double()[ok], want.detach()[ok], rtol=1e-5, atol=2e-5)
        g = torch.tensor(np.random.default_rng(1).uniform(0.25, 1.0, size=(B, T)))
        (torch.nan_to_num(got) * g.to("cuda", torch.float32)).sum().backward()
        (torch.nan_to_num(want) * g).sum().backward()
        scale = float(x64.grad.abs().max())
        assert torch.allclose(xs.grad.cpu().double(), x64.grad, rtol=1e-3, atol=1e-5 * scale)
        # samples before t + a of the first output belong to no window
        if a:
            assert float(xs.grad[:, :a].abs().max()) == 0.0


@pytest.mark.parametrize("sem,kw", SEMS, ids=IDS)
def test_long_trace_properties(sem, kw):
    B, T = 3, 16385  # one sample past the largest tile
    x, y = preds()
    s = create_semantics(sem, backend="b200", **kw)
    sig = torch.tensor(signal(B, T, 21), device="cuda")
    g = s.robustness_trace(Always(x, Interval(0, None)), sig)
    f = s.robustness_trace(Eventually(x, Interval(0, None)), sig)
    leaf = sig[..., 0] + 0.1
    # the window [T-1, T-1] holds one sample: every semantics returns it (agm: the mean of one entry; dgmsr: up to
    # sqrt(eps) = 1e-4, _dgmsr.py:83-102)
    one = dict(rtol=1e-5, atol=1e-6) if sem != "dgmsr" else dict(rtol=1e-5, atol=2.5e-4)
    assert torch.allclose(g[:, -1], leaf[:, -1], **one)
    assert torch.allclose(f[:, -1], leaf[:, -1], **one)
    assert bool((g <= f + one["atol"] + 1e-5).all())
    if sem in ("classical", "smooth"):
        # a longer window can only lower Always / raise Eventually
        assert bool((g[:, :-1] <= g[:, 1:] + 1e-6).all()) and bool((f[:, :-1] >= f[:, 1:] - 1e-6).all())
    # a constant positive trace: Always == Eventually == the constant for classical / agm / dgmsr
    const = torch.full((2, T, 2), 0.7, device="cuda")
    gc = s.robustness_trace(Always(x, Interval(0, None)), const)
    if sem != "smooth":
        assert torch.allclose(gc, torch.full_like(gc, 0.8), rtol=2e-4, atol=1e-5)
    # mixed with bounded operators, one-pass mean + gradient agrees with autograd through the trace
    phi = And(Always(x, Interval(3, None)), Eventually(y, Interval(0, 10)))
    xs = sig.clone().requires_grad_(True)
    r = s.robustness_trace(phi, xs)
    assert int(torch.isnan(r).sum()) == B * 3 and bool(torch.isnan(r[:, -3:]).all())
    m = r[torch.isfinite(r)].mean()
    m.backward()
    mean, grad = s.mean_robustness_and_grad(phi, sig)
    assert abs(float(mean) - float(m)) < 1e-5 * (1 + abs(float(m)))
    assert torch.allclose(grad, xs.grad, rtol=1e-4, atol=1e-6 * float(xs.grad.abs().max()) + 1e-12)


# ---------------------------------------------------------------------------- cumulative: (pos, neg) pairs
def cumulative_formulas():
    x, y = preds32()
    return {
        "G_open": Always(x, Interval(0, None)),
        "F_open_a": Eventually(y, Interval(7, None)),
        "mixed": And(Always(x, Interval(3, None)), Eventually(y, Interval(0, 10))),
        "under_window": Eventually(Always(Or(x, Not(y)), Interval(2, None)), Interval(0, 5)),
        "negated": Not(Eventually(And(x, y), Interval(1, None))),
        "nested": Always(Eventually(x, Interval(0, None)), Interval(2, None)),
        "over_until": Eventually(Until(x, y, Interval(0, 6)), Interval(5, None)),
    }


@pytest.mark.parametrize("name", list(cumulative_formulas()))
def test_cumulative_scan_path_against_the_oracle(name):
    """Cumulative values are (pos, neg) pairs (cumulative.py:13-18): Always is the minimum of each channel, Eventually its SUM
    (cumulative.py:35-46, 62-73); the window's two columns come back into the formula through a pair predicate."""
    T = 300
    phi = cumulative_formulas()[name]
    sig = signal(3, T, 13)
    s = create_semantics("cumulative", backend="b200")
    s.force_open_leaf = True
    x = torch.tensor(sig, device="cuda", requires_grad=True)
    cf = s._checked_compile(phi, x.detach())
    assert any(isinstance(p, OpenLeaf) for p in cf.leaf_preds) and cf.n_leaf_columns == 2 * len(cf.leaf_preds)
    r = s.robustness_trace(phi, x)
    got = torch.stack((r.pos, r.neg), 0)
    x64 = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
    want = dense.robustness_trace(phi, x64, dense.make("cumulative"))
    ok = torch.isfinite(want)
    assert torch.equal(torch.isfinite(got.detach().cpu()), ok)
    assert torch.allclose(got.detach().cpu().double()[ok], want.detach()[ok], rtol=1e-5, atol=5e-6)
    g = torch.tensor(np.random.default_rng(3).uniform(0.25, 1.0, size=tuple(want.shape)))
    (torch.nan_to_num(got) * g.to("cuda", torch.float32)).sum().backward()
    (torch.nan_to_num(want) * g).sum().backward()
    scale = float(x64.grad.abs().max()) + 1e-12
    assert torch.allclose(x.grad.cpu().double(), x64.grad, rtol=4e-5, atol=4e-5 * scale)
    mean, grad = s.mean_robustness_and_grad(phi, torch.tensor(sig, device="cuda"))
    x2 = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
    w2 = dense.robustness_trace(phi, x2, dense.make("cumulative")).sum(0)
    m2 = w2[torch.isfinite(w2)].mean()
    m2.backward()
    assert abs(float(mean) - float(m2.detach())) < 1e-5 * (1 + abs(float(m2.detach())))
    assert torch.allclose(grad.cpu().double(), x2.grad, rtol=4e-5, atol=4e-5 * (float(x2.grad.abs().max()) + 1e-12))
    assert abs(float(s.mean_robustness(phi, torch.tensor(sig, device="cuda"))) - float(m2.detach())) < 1e-5 * (1 + abs(float(m2.detach())))


def test_cumulative_long_traces_against_closed_forms():
    B, T, a = 4, 20000, 11
    x, _ = preds32()
    sig = signal(B, T, 9)
    s = create_semantics("cumulative", backend="b200")
    x64 = torch.tensor(sig, dtype=torch.float64)
    v = x64[..., 0] + 0.125
    pos, neg = torch.clamp(v, min=0.0), torch.clamp(v, max=0.0)
    for kind, phi in (("G", Always(x, Interval(a, None))), ("F", Eventually(x, Interval(a, None)))):
        r = s.robustness_trace(phi, torch.tensor(sig, device="cuda"))
        assert s.plan(phi, T, 2)["start_slack"] == a
        for got, ch in ((r.pos, pos), (r.neg, neg)):
            fl = torch.flip(ch, dims=(1,))
            full = torch.flip(torch.cummin(fl, 1).values if kind == "G" else torch.cumsum(fl, 1), dims=(1,))
            want = torch.full_like(full, float("nan"))
            want[:, : T - a] = full[:, a:]
            ok = torch.isfinite(want)
            assert torch.equal(torch.isfinite(got.cpu()), ok)
            assert torch.allclose(got.cpu().double()[ok], want[ok], rtol=1e-5, atol=5e-6), (kind, float((got.cpu().double()[ok] - want[ok]).abs().max()))
    with pytest.raises(NotImplementedError):  # the cumulative Until without an end stays inside one tile
        s.robustness_trace(Until(x, x, Interval(0, None)), torch.tensor(sig, device="cuda"))


# ---------------------------------------------------------------------------- classical Until without an end
def until_formulas():
    x, y = preds32()
    return {
        "U_open": Until(x, y, Interval(0, None)),
        "U_open_a": Until(x, y, Interval(6, None)),
        "U_nested_operands": Until(Or(x, Not(y)), Eventually(y, Interval(0, 3)), Interval(4, None)),
        "U_mixed": And(Until(x, y, Interval(2, None)), Always(Not(y), Interval(1, None))),
        "U_of_U": Until(x, Until(y, x, Interval(0, None)), Interval(1, None)),
    }


@pytest.mark.parametrize("name", list(until_formulas()))
def test_open_until_against_the_oracle(name):
    T = 300
    phi = until_formulas()[name]
    sig = signal(3, T, 17)
    s = create_semantics("classical", backend="b200")
    s.force_open_leaf = True
    x = torch.tensor(sig, device="cuda", requires_grad=True)
    cf = s._checked_compile(phi, x.detach())
    assert any(isinstance(p, OpenUntilLeaf) for p in cf.leaf_preds)
    got = s.robustness_trace(phi, x)
    x64 = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
    want = dense.robustness_trace(phi, x64, dense.make("classical"))
    ok = torch.isfinite(want)
    assert torch.equal(torch.isfinite(got.detach().cpu()), ok)
    assert torch.equal(got.detach().cpu()[ok], want.detach()[ok].to(torch.float32))  # classical: bit-exact
    g = torch.tensor(np.random.default_rng(3).uniform(0.25, 1.0, size=tuple(want.shape)))
    (torch.nan_to_num(got) * g.to("cuda", torch.float32)).sum().backward()
    (torch.nan_to_num(want) * g).sum().backward()
    assert torch.allclose(x.grad.cpu().double(), x64.grad, rtol=1e-5, atol=1e-5)
    mean, grad = s.mean_robustness_and_grad(phi, torch.tensor(sig, device="cuda"))
    x2 = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
    w2 = dense.robustness_trace(phi, x2, dense.make("classical"))
    m2 = w2[torch.isfinite(w2)].mean()
    m2.backward()
    assert abs(float(mean) - float(m2.detach())) < 1e-5 * (1 + abs(float(m2.detach())))
    assert torch.allclose(grad.cpu().double(), x2.grad, rtol=1e-5, atol=1e-7)


def test_open_until_equals_the_in_kernel_path_and_scales():
    sig = torch.tensor(signal(4, 2048, 5), device="cuda")
    for name, phi in until_formulas().items():
        a = create_semantics("classical", backend="b200")
        b = create_semantics("classical", backend="b200")
        b.force_open_leaf = True
        xa, xb = sig.clone().requires_grad_(True), sig.clone().requires_grad_(True)
        ra, rb = a.robustness_trace(phi, xa), b.robustness_trace(phi, xb)
        ok = torch.isfinite(ra)
        assert torch.equal(ok, torch.isfinite(rb)) and torch.equal(ra[ok], rb[ok]), name
        torch.nan_to_num(ra).sum().backward()
        torch.nan_to_num(rb).sum().backward()
        assert torch.allclose(xa.grad, xb.grad, rtol=1e-5, atol=1e-4), (name, float((xa.grad - xb.grad).abs().max()))
    # T = 20000: the recurrence W(u) = min(l_u, max(r_u, W(u + 1))) in float64, one step per sample
    B, T, a0 = 3, 20000, 9
    x, y = preds32()
    phi = Until(x, y, Interval(a0, None))
    sg = signal(B, T, 23)
    s = create_semantics("classical", backend="b200")
    xs = torch.tensor(sg, device="cuda", requires_grad=True)
    got = s.robustness_trace(phi, xs)
    assert s.plan(phi, T, 2)["start_slack"] == a0
    x64 = torch.tensor(sg, dtype=torch.float64, requires_grad=True)
    l = x64[..., 0] + 0.125
    r = -x64[..., 1] + 0.25 * x64[..., 0] + 0.375
    w = [None] * T
    run = torch.full((B,), float("-inf"), dtype=torch.float64)
    for u in range(T - 1, -1, -1):
        run = torch.minimum(l[:, u], torch.maximum(r[:, u], run))
        w[u] = run
    W = torch.stack(w, 1)
    gl = l.unfold(1, a0 + 1, 1).amin(-1)  # G[0, a] l
    want = torch.full((B, T), float("nan"), dtype=torch.float64)
    want[:, : T - a0] = torch.minimum(gl[:, : T - a0], W[:, a0:])
    ok = torch.isfinite(want)
    assert torch.equal(torch.isfinite(got.detach().cpu()), ok)
    assert torch.equal(got.detach().cpu()[ok], want.detach()[ok].to(torch.float32))
    gsel = torch.tensor(np.random.default_rng(2).uniform(0.25, 1.0, size=(B, T)))
    (torch.nan_to_num(got) * gsel.to("cuda", torch.float32)).sum().backward()
    (torch.nan_to_num(want) * gsel).sum().backward()
    assert torch.allclose(xs.grad.cpu().double(), x64.grad, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("sem,kw", [SEMS[0], SEMS[2], SEMS[3], SEMS[4]], ids=["classical", "smooth0.05", "agm", "dgmsr1"])
def test_chunk_boundaries_and_degenerate_shapes(sem, kw):
    """The scan kernels walk 256-sample chunks, 8 samples per lane: every trace length around those boundaries, starts
    up to and past the horizon (a >= T: the window is empty at every t, the reference raises, api.py:218-221).  Called
    through the torch ops on raw value traces (no formula kernel, nothing to compile)."""
    from pystl_b200 import capi, ops

    sem_id = {"classical": capi.SEM_CLASSIC, "smooth": capi.SEM_SMOOTH, "agm": capi.SEM_AGM, "dgmsr": capi.SEM_DGMSR}[sem]
    tau, eps, p = kw.get("temperature", 1.0), kw.get("eps", 1e-8), kw.get("p", 1)
    v = affine_predicate("v", {0: 1.0}, 0.0)
    w = affine_predicate("w", {1: 1.0}, 0.0)
    rng = np.random.default_rng(77)
    for T in (1, 2, 8, 9, 32, 33, 255, 256, 257, 513, 777):
        sig = (0.5 * rng.standard_normal((2, T, 2))).astype(np.float32)
        for a in sorted({0, min(5, T - 1), T // 2, T - 1, T, T + 3}):
            for mode, phi in ((0, Always(v, Interval(a, None))), (1, Eventually(v, Interval(a, None)))):
                ys = torch.tensor(sig[..., 0], device="cuda", requires_grad=True)
                got = ops.open_window(ys, sem_id, mode, a, tau, eps, p)
                x64 = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
                want = dense.robustness_trace(phi, x64, dense.make(sem, **kw))
                ok = torch.isfinite(want)
                tag = (T, a, mode)
                assert int(ok.sum()) == 2 * max(T - a, 0), tag
                assert bool(torch.isfinite(got).all()), tag  # (positions the reference raises on hold a finite filler)
                if not bool(ok.any()):
                    continue
                if sem == "classical":
                    assert torch.equal(got.detach().cpu()[ok], want.detach()[ok].to(torch.float32)), tag
                else:
                    assert torch.allclose(got.detach().cpu().double()[ok], want.detach()[ok], **tol(sem)), tag
                gsel = torch.tensor(rng.uniform(0.25, 1.0, size=(2, T)))
                (got * torch.where(ok, gsel, 0.0).to("cuda", torch.float32)).sum().backward()
                (torch.nan_to_num(want) * gsel).sum().backward()
                fin = torch.isfinite(x64.grad).all(dim=(1, 2))  # (dgmsr: see test_scan_path_against_the_oracle)
                assert bool(torch.isfinite(ys.grad).all()), tag
                scale = float(x64.grad[fin].abs().max()) + 1e-12 if bool(fin.any()) else 1.0
                assert torch.allclose(ys.grad.cpu().double()[fin], x64.grad[fin][..., 0], rtol=5e-4, atol=2e-5 * scale), tag
            if sem != "classical":
                continue
            # Until: W(t + a); at a = 0 it is the reference's Until, at a > 0 the same trace read a samples later
            l = torch.tensor(sig[..., 0], device="cuda", requires_grad=True)
            r = torch.tensor(sig[..., 1], device="cuda", requires_grad=True)
            w0 = ops.open_until(l, r, sem_id, 0)
            x64 = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
            want = dense.robustness_trace(Until(v, w, Interval(0, None)), x64, dense.make("classical"))
            assert torch.equal(w0.detach().cpu(), want.detach().to(torch.float32)), (T, "until")
            gsel = torch.tensor(rng.uniform(0.25, 1.0, size=(2, T)))
            (w0 * gsel.to("cuda", torch.float32)).sum().backward()
            (want * gsel).sum().backward()
            assert torch.allclose(torch.stack((l.grad, r.grad), -1).cpu().double(), x64.grad, rtol=1e-5, atol=1e-6), (T, "until grad")
            wa = ops.open_until(l.detach(), r.detach(), sem_id, a)
            if a < T:
                assert torch.equal(wa[:, : T - a], w0.detach()[:, a:]), (T, a, "until shift")
    # an empty batch is a no-op
    assert tuple(ops.open_window(torch.zeros(0, 300, device="cuda"), sem_id, 0, 2, tau, eps, p).shape) == (0, 300)


# ---------------------------------------------------------------------------- random formulas: both routes must agree
def random_open_formula(seed: int, sem: str):
    """A random tree of depth <= 3 over two affine predicates with at least one open-ended operator."""
    rng = np.random.default_rng(1000 + seed)
    x, y = preds32()

    def interval(open_ok: bool):
        a = int(rng.integers(0, 6))
        if open_ok and rng.random() < 0.6:
            return Interval(a, None), True
        return Interval(a, a + int(rng.integers(0, 12))), False

    def build(depth: int, slack_ok: bool):
        """(formula, has an open-ended operator, start slack is zero)"""
        r = rng.random()
        if depth == 0 or r < 0.15:
            return (x if rng.random() < 0.5 else y), False, True
        if r < 0.3:
            f, o, z = build(depth - 1, slack_ok)
            return Not(f), o, z
        if r < 0.55:
            kids = [build(depth - 1, slack_ok) for _ in range(int(rng.integers(2, 4)))]
            return (And if rng.random() < 0.5 else Or)(*[k[0] for k in kids]), any(k[1] for k in kids), all(k[2] for k in kids)
        if r < 0.9 or sem != "classical":
            f, o, z = build(depth - 1, True)
            iv, is_open = interval(z)  # an operand with a start slack under an open-ended window raises everywhere: avoided
            if not slack_ok and iv.start > 0:
                iv = Interval(0, iv.end)
            return (Always if rng.random() < 0.5 else Eventually)(f, iv), o or is_open, z and iv.start == 0
        lf, lo, lz = build(depth - 1, False)
        rf, ro, rz = build(depth - 1, False)
        iv, is_open = interval(lz and rz)
        if not slack_ok and iv.start > 0:
            iv = Interval(0, iv.end)
        return Until(lf, rf, iv), lo or ro or is_open, lz and rz and iv.start == 0

    for _ in range(200):
        f, has_open, _ = build(3, True)
        if has_open:
            return f
    return Always(x, Interval(1, None))


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("sem,kw", [SEMS[0], SEMS[1], SEMS[3], SEMS[4], ("cumulative", {})], ids=["classical", "smooth0.5", "agm", "dgmsr1", "cumulative"])
def test_random_formulas_scan_route_equals_in_kernel_route(sem, kw, seed):
    """One tile holds T = 1000 for values and gradients of every semantics, so both routes exist: the open-ended operators
    inside the formula kernel (interpreter tier) and behind leaf channels (scan kernels + a fused parent)."""
    phi = random_open_formula(seed, sem)
    sig = torch.tensor(signal(3, 1000, 40 + seed), device="cuda")
    a = create_semantics(sem, backend="b200", **kw)
    b = create_semantics(sem, backend="b200", **kw)
    b.force_open_leaf = True
    xa, xb = sig.clone().requires_grad_(True), sig.clone().requires_grad_(True)
    rb = b.robustness_trace(phi, xb)
    if sem == "cumulative":
        # (its in-kernel gradient does not exist -- the interpreter tier has no cumulative adjoints -- so the automatic choice is
        # the scan route already: values from the in-kernel forward, gradients from the oracle)
        rb = torch.stack((rb.pos, rb.neg))
        inside = a.compile(phi, 2)
        assert inside.has_open and not inside.leaf_preds
        ri = a._forward_raw(inside, sig, None).permute(2, 0, 1)
        ok = torch.isfinite(ri)
        assert torch.equal(ok, torch.isfinite(rb)) and torch.allclose(ri[ok], rb[ok], rtol=1e-5, atol=5e-6), repr(phi)
        x64 = sig.detach().cpu().double().requires_grad_(True)
        want = dense.robustness_trace(phi, x64, dense.make("cumulative"))
        assert torch.allclose(rb.detach().cpu().double()[ok.cpu()], want.detach()[ok.cpu()], rtol=1e-5, atol=5e-6), repr(phi)
        g = torch.rand(tuple(want.shape), dtype=torch.float64, generator=torch.Generator().manual_seed(seed)) + 0.25
        (torch.nan_to_num(rb) * g.to("cuda", torch.float32)).sum().backward()
        (torch.nan_to_num(want) * g).sum().backward()
        scale = float(x64.grad.abs().max()) + 1e-12
        assert torch.allclose(xb.grad.cpu().double(), x64.grad, rtol=4e-5, atol=4e-5 * scale), (repr(phi), scale)
        return
    ra = a.robustness_trace(phi, xa)
    ok = torch.isfinite(ra)
    assert torch.equal(ok, torch.isfinite(rb)), repr(phi)
    if sem == "classical":
        assert torch.equal(ra[ok], rb[ok]), repr(phi)
    else:
        assert torch.allclose(ra[ok], rb[ok], rtol=2e-4, atol=2e-5), (repr(phi), float((ra[ok] - rb[ok]).abs().max()))
    if not bool(ok.any()):
        return
    g = torch.rand(ra.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(seed)) + 0.25
    (torch.nan_to_num(ra) * g).sum().backward()
    (torch.nan_to_num(rb) * g).sum().backward()
    scale = float(xa.grad.abs().max()) + 1e-12
    assert torch.allclose(xa.grad, xb.grad, rtol=2e-3, atol=2e-4 * scale), (repr(phi), float((xa.grad - xb.grad).abs().max()), scale)


def test_vmap_and_functional_grad_through_the_scan_route():
    """``torch.func.vmap`` over traces and ``torch.func.grad`` compose with the open-ended route like with the formula
    kernels (the precedent: jax.vmap over the reference's JAX backend, tests/test_jax_semantics.py:310-368)."""
    x, y = preds32()
    sig = torch.tensor(signal(5, 400, 31), device="cuda")
    for sem, kw, phi in (("smooth", {"temperature": 0.5}, And(Always(x, Interval(3, None)), Eventually(y, Interval(0, 10)))),
                         ("agm", {}, Eventually(Or(x, Not(y)), Interval(2, None))),
                         ("classical", {}, And(Until(x, y, Interval(2, None)), Always(Not(y), Interval(1, None))))):
        s = create_semantics(sem, backend="b200", **kw)
        s.force_open_leaf = True
        want = s.robustness_trace(phi, sig)
        got = torch.func.vmap(lambda tr: s.robustness_trace(phi, tr))(sig)
        assert torch.equal(torch.isfinite(got), torch.isfinite(want))
        ok = torch.isfinite(want)
        assert torch.equal(got[ok], want[ok]), sem

        def loss(tr):
            return torch.nan_to_num(s.robustness_trace(phi, tr)).sum()

        xs = sig.clone().requires_grad_(True)
        loss(xs).backward()
        g_f = torch.func.vmap(torch.func.grad(loss))(sig)
        assert torch.allclose(g_f, xs.grad, rtol=1e-6, atol=1e-7), sem


def test_a_window_that_always_raises_is_nan_everywhere():
    # an operand that is itself undefined near the horizon makes the open-ended window raise at every t (api.py:218-221)
    x, _ = preds()
    phi = Eventually(Always(x, Interval(1, 4)), Interval(5, None))
    sig = torch.tensor(signal(2, 20000, 2), device="cuda")
    for sem, kw in SEMS[:2]:
        s = create_semantics(sem, backend="b200", **kw)
        assert s.plan(phi, 20000, 2)["raises_always"] == 1
        assert bool(torch.isnan(s.robustness_trace(phi, sig)).all())
        want = dense.robustness_trace(phi, torch.tensor(signal(1, 40, 2), dtype=torch.float64), dense.make(sem, **kw))
        assert bool(torch.isnan(want).all())


def test_unsupported_cases_are_loud():
    x, y = preds()
    sig = torch.zeros(1, 20000, 2, device="cuda")
    with pytest.raises(NotImplementedError):
        create_semantics("agm", backend="b200").robustness_trace(Always(x, Interval(0, None), weights=[1.0] * 20000), sig)
    with pytest.raises(NotImplementedError):  # the non-classical Until without an end is O(T^2) per trace: one tile only
        create_semantics("smooth", backend="b200").robustness_trace(Until(x, y, Interval(0, None)), sig)
