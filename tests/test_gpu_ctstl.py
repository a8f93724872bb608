"""GPU tier of CT-STL (csrc/select.cu behind pystl_b200/ctstl.py): the cumulative-time operator C^tau_I and the Until with an
exclusive prefix, through the C ABI and through ``B200CtstlSemantics``.

Values are SELECTIONS of input samples, so they are compared bit for bit (``==`` against float32 of the float64 reference):
1. against the outputs of the unmodified reference class (tests/golden/ctstl_vectors.npz) and the oracle at random shapes,
   windows longer than a tile, traces of one sample, windows that are always clipped, no interval end;
2. at BASELINE-like sizes through properties: k = 1 is the classical Eventually and k = window length the classical Always
   (fused classical kernels on the same trace), the value is a sample of the window, rank counts bracket k;
3. gradients: the recorded index is the one a stable sort selects (oracle), backward is the scatter of the cotangent along it;
4. whole formulas cut at their CT-STL nodes against the reference-shaped recursion over the oracle, values and autograd.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import ctstl as oc
from pystl_b200 import B200CtstlSemantics, CumulativeTime, capi
from pystl_b200.formula import Always, And, Eventually, Interval, Not, Or, Until, _evaluate_recursive, affine_predicate

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ctstl_vectors.npz"))
DEV = "cuda:0"


def traces(seed, B, T, quant=0.0):
    rng = np.random.default_rng(seed)
    y = rng.uniform(-1, 1, size=(B, 1)) + rng.standard_normal((B, T))
    if quant:
        y = np.round(y / quant) * quant
    return y.astype(np.float32)


def cum_capi(y, a, b, k, want_arg=True):
    yt = torch.tensor(y, device=DEV)
    B, T = yt.shape
    out = torch.full_like(yt, 123.0)
    arg = torch.full((B, T), -7, dtype=torch.int32, device=DEV)
    rc = capi.lib().pystl_b200_ctstl_cumulative_forward(yt.data_ptr(), B, T, a, -1 if b is None else b, k, out.data_ptr(),
                                                        arg.data_ptr() if want_arg else None, torch.cuda.current_stream().cuda_stream)
    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
    torch.cuda.synchronize()
    return out.cpu().numpy(), arg.cpu().numpy()


def test_cumulative_matches_reference_vectors():
    y = GOLD["y"].astype(np.float32)
    for c, (a, b, tau, delta) in enumerate(GOLD["cum_cases"]):
        a, b = int(a), int(b)
        k = oc.tau_to_k(tau, delta)
        out, arg = cum_capi(y, a, None if b < 0 else b, k)
        assert np.array_equal(out, GOLD["cum"][c].astype(np.float32)), (a, b, tau, delta)
        _, want_arg = oc.cumulative_trace(y.astype(np.float64), a, None if b < 0 else b, k)
        assert np.array_equal(arg, want_arg), (a, b, tau, delta)


@pytest.mark.parametrize("B,T,a,b,k,quant", [
    (3, 1, 0, 0, 1, 0.0), (3, 1, 0, 5, 1, 0.0), (2, 1, 1, 5, 1, 0.0), (5, 513, 0, 7, 3, 0.0), (4, 1030, 2, 65, 17, 0.25),
    (3, 700, 0, 699, 350, 0.0), (2, 1500, 5, 900, 1, 0.5), (2, 1500, 5, 900, 896, 0.0), (3, 300, 0, None, 4, 0.5),
    (2, 2100, 100, None, 77, 0.0), (7, 64, 0, 2, 3, 1.0), (2, 40, 0, 3, 2, 0.0), (2, 40, 0, 4, 5, 0.0), (1, 4099, 0, 11, 12, 0.125),
    (3, 50, 60, 70, 1, 0.0),
])
def test_cumulative_against_oracle(B, T, a, b, k, quant):
    y = traces(B * 1000 + T, B, T, quant)
    out, arg = cum_capi(y, a, b, k)
    want, want_arg = oc.cumulative_trace(y.astype(np.float64), a, b, k)
    assert np.array_equal(out, want.astype(np.float32))
    assert np.array_equal(arg, want_arg)
    out2, _ = cum_capi(y, a, b, k, want_arg=False)  # the arg buffer is optional
    assert np.array_equal(out2, out)


def test_cumulative_backward_is_the_scatter_along_the_selection():
    B, T, a, b, k = 4, 777, 1, 40, 9
    y = traces(11, B, T, 0.25)
    yt = torch.tensor(y, device=DEV, requires_grad=True)
    sem = B200CtstlSemantics()
    out = sem.cumulative_time_trace(yt, (a, b), tau=float(k))
    g = torch.tensor(np.random.default_rng(3).standard_normal((B, T)).astype(np.float32), device=DEV)
    fin = torch.isfinite(out)
    (out.masked_fill(~fin, 0.0) * g).sum().backward()
    want, want_arg = oc.cumulative_trace(y.astype(np.float64), a, b, k)
    assert np.array_equal(out.detach().cpu().numpy(), want.astype(np.float32))
    gn = np.where(np.isfinite(want), g.cpu().numpy().astype(np.float64), 0.0)
    np.testing.assert_allclose(yt.grad.cpu().numpy(), oc.cumulative_grad(want_arg, gn), rtol=1e-5, atol=1e-5)


def test_cumulative_nan_in_window_and_errors():
    y = traces(1, 2, 60)
    y[0, 30] = np.nan
    out, arg = cum_capi(y, 0, 4, 2)
    assert np.all(np.isnan(out[0, 26:31])) and not np.isnan(out[0, :26]).any() and not np.isnan(out[0, 31:]).any()
    assert np.all(arg[0, 26:31] == -1) and not np.isnan(out[1]).any()
    lib = capi.lib()
    yt = torch.tensor(y, device=DEV)
    o = torch.empty_like(yt)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.pystl_b200_ctstl_cumulative_forward(yt.data_ptr(), 2, 60, 0, 4, 0, o.data_ptr(), None, st) == capi.E_INVALID  # tau must be > 0
    assert lib.pystl_b200_ctstl_cumulative_forward(yt.data_ptr(), 2, 60, -1, 4, 1, o.data_ptr(), None, st) == capi.E_INVALID
    assert lib.pystl_b200_ctstl_cumulative_forward(yt.data_ptr(), 2, 60, 5, 4, 1, o.data_ptr(), None, st) == capi.E_INVALID
    assert lib.pystl_b200_ctstl_cumulative_forward(None, 2, 60, 0, 4, 1, o.data_ptr(), None, st) == capi.E_INVALID
    assert lib.pystl_b200_ctstl_cumulative_forward(None, 0, 60, 0, 4, 1, None, None, st) == capi.OK  # empty batch
    sem = B200CtstlSemantics()
    with pytest.raises(ValueError):
        sem.cumulative_time_trace(yt, (0, 4), tau=0.0)
    with pytest.raises(NotImplementedError):  # a window beyond the shared-memory tile
        sem.cumulative_time_trace(torch.zeros(1, 60000, device=DEV), (0, None), tau=1.0)


def test_cumulative_at_scale_against_the_classical_windows():
    """k = 1 / k = window length are the classical Eventually / Always of the same trace (fused classical kernels); any k
    returns a sample of its window whose rank brackets k.  2^22 outputs, windows across tile boundaries."""
    B, T, a, b = 1024, 4096, 3, 66
    W = b - a + 1
    gen = torch.Generator(device=DEV).manual_seed(7)
    y = torch.randn(B, T, device=DEV, generator=gen)
    y = torch.round(y * 8) / 8  # ties
    sem = B200CtstlSemantics()
    v = affine_predicate("v", {0: 1.0})
    top = sem.cumulative_time_trace(y, (a, b), tau=1.0)
    ev = sem.robustness_trace(Eventually(v, (a, b)), y.unsqueeze(-1))
    valid = ~torch.isnan(ev)
    assert torch.equal(top[valid], ev[valid]) and bool(torch.isinf(top[~valid]).all())
    bot = sem.cumulative_time_trace(y, (a, b), tau=float(W))
    al = sem.robustness_trace(Always(v, (a, b)), y.unsqueeze(-1))
    full = torch.arange(T, device=DEV) + b <= T - 1
    assert torch.equal(bot[:, full], al[:, full]) and bool((bot[:, ~full] == float("-inf")).all())
    k = 20
    mid = sem.cumulative_time_trace(y, (a, b), tau=float(k))
    win = y.unfold(1, W, 1)[:, a:, :]  # [B, T - b, W]: the windows that are not clipped
    m = mid[:, : win.shape[1]].unsqueeze(-1)
    assert bool(((win > m).sum(-1) < k).all()) and bool(((win >= m).sum(-1) >= k).all())
    assert bool((top[:, : win.shape[1]] >= mid[:, : win.shape[1]]).all()) and bool((mid[:, : win.shape[1]] >= bot[:, : win.shape[1]]).all())


def until_capi(l, r, a, b):
    lt, rt = torch.tensor(l, device=DEV), torch.tensor(r, device=DEV)
    B, T = lt.shape
    out = torch.full_like(lt, 123.0)
    key = torch.full((B, T), -7, dtype=torch.int32, device=DEV)
    rc = capi.lib().pystl_b200_ctstl_until_forward(lt.data_ptr(), rt.data_ptr(), B, T, a, -1 if b is None else b, out.data_ptr(), key.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream)
    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
    torch.cuda.synchronize()
    return out.cpu().numpy(), key.cpu().numpy()


def test_until_matches_reference_vectors_and_routes_to_the_value():
    y, l = GOLD["y"].astype(np.float32), GOLD["l"].astype(np.float32)
    for c, (a, b) in enumerate(GOLD["until_cases"]):
        a, b = int(a), int(b)
        out, key = until_capi(l, y, a, None if b < 0 else b)
        assert np.array_equal(out, GOLD["until"][c].astype(np.float32), equal_nan=True), (a, b)
        ok = ~np.isnan(out)
        assert np.array_equal(key >= 0, ok)
        rows, ts = np.nonzero(ok)
        kk = key[ok]
        src = np.where(kk & 1, y[rows, kk >> 1], l[rows, kk >> 1])
        assert np.array_equal(src, out[ok])
        # an r comes from the window [t + a, t + b], an l from [t, t + b - 1]
        assert np.all((kk >> 1) >= ts + np.where(kk & 1, a, 0))


@pytest.mark.parametrize("B,T,a,b", [(3, 1, 0, 0), (3, 1, 0, None), (4, 300, 0, 24), (3, 1000, 5, 64), (2, 600, 0, None), (2, 90, 7, 7), (2, 50, 60, 80)])
def test_until_against_oracle_values_and_gradient(B, T, a, b):
    l, r = traces(B + T, B, T, 0.25), traces(B + T + 1, B, T, 0.0)
    sem = B200CtstlSemantics()
    lt = torch.tensor(l, device=DEV, requires_grad=True)
    rt = torch.tensor(r, device=DEV, requires_grad=True)
    out = sem.until_trace(lt, rt, (a, b))
    want = oc.until_trace(l.astype(np.float64), r.astype(np.float64), a, b)
    assert np.array_equal(out.detach().cpu().numpy(), want.astype(np.float32), equal_nan=True)
    g = torch.tensor(np.random.default_rng(9).standard_normal((B, T)).astype(np.float32), device=DEV)
    ok = ~torch.isnan(out)
    if not bool(ok.any()):
        return
    (out.masked_fill(~ok, 0.0) * g).sum().backward()
    _, key = until_capi(l, r, a, b)
    gl, gr = np.zeros((B, T)), np.zeros((B, T))
    gn = g.cpu().numpy()
    for i, t in zip(*np.nonzero(key >= 0)):
        (gr if key[i, t] & 1 else gl)[i, key[i, t] >> 1] += gn[i, t]
    np.testing.assert_allclose(lt.grad.cpu().numpy(), gl, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(rt.grad.cpu().numpy(), gr, rtol=1e-5, atol=1e-5)


class OracleSem:
    """The reference's recursion shape over python floats: the classical lattice + oracle/ctstl.py."""

    def predicate(self, p, s, t): return float(p.fn(s, t))
    def boolean_not(self, v): return -v
    def boolean_and(self, vs, *, weights=None): return min(vs)
    def boolean_or(self, vs, *, weights=None): return max(vs)
    temporal_always, temporal_eventually = boolean_and, boolean_or
    def temporal_until(self, *, left_trace, right_trace, start, end, **_): return oc.temporal_until(left_trace, right_trace, start, end)
    def temporal_cumulative(self, vs, *, tau, delta=None): return oc.temporal_cumulative(vs, tau=tau, delta=1.0 if delta is None else delta)


def ct_formulas():
    x = affine_predicate("x", {0: 1.0}, 0.125)
    y = affine_predicate("y", {1: -1.0, 0: 0.25}, 0.375)
    return {
        "until": Until(x, y, (0, 6)),
        "cum": CumulativeTime(x, (0, 9), tau=3.0),
        "cum_under_always": Always(CumulativeTime(Or(x, Not(y)), (1, 6), tau=2.0), (0, 4)),
        "until_in_nary": And(Eventually(x, (0, 3)), Not(Until(x, y, (1, 8))), CumulativeTime(y, (0, 5), tau=6.0)),
        "nested": Until(CumulativeTime(x, (0, 4), tau=2.0), Eventually(y, (0, 2)), (0, 10)),
        "open": Or(Until(x, y, (2, None)), CumulativeTime(y, (3, None), tau=2.0)),
        "plain": Always(x, (0, 5)) & Eventually(y, (2, 4)),
    }


@pytest.mark.parametrize("name", list(ct_formulas()))
def test_formulas_cut_at_ct_nodes_against_the_recursion(name):
    phi = ct_formulas()[name]
    B, T = 3, 41
    rng = np.random.default_rng(17)
    sig = (rng.uniform(-0.5, 0.5, size=(B, 1, 2)) + 0.5 * rng.standard_normal((B, T, 2))).astype(np.float32)
    sem = B200CtstlSemantics(delta=1.0)
    st = torch.tensor(sig, device=DEV, requires_grad=True)
    out = sem.robustness_trace(phi, st)
    want = np.full((B, T), np.nan)
    for i in range(B):
        for t in range(T):
            try:
                want[i, t] = _evaluate_recursive(phi, sig[i].astype(np.float64), OracleSem(), t)
            except ValueError:
                pass
    got = out.detach().cpu().numpy()
    assert np.array_equal(got, want.astype(np.float32), equal_nan=True), name
    # one t through the reference-shaped entry point, raising where the recursion raises
    t_ok = int(np.nonzero(~np.isnan(want[0]))[0][0]) if (~np.isnan(want[0])).any() else None
    if t_ok is not None:
        assert float(phi.evaluate(st[0].detach(), sem, t_ok)) == np.float32(want[0, t_ok])
    if np.isnan(want[0]).any():
        with pytest.raises(ValueError):
            phi.evaluate(st[0].detach(), sem, int(np.nonzero(np.isnan(want[0]))[0][0]))
    # gradient: a selection network -- d out[i, t] / d signal is the predicate's coefficients at ONE sample; check it by moving
    # the signal along the gradient of a random functional and comparing the first-order change (piecewise linear: exact for a small step)
    ok = torch.isfinite(out)
    if bool(ok.any()):
        g = torch.tensor(rng.standard_normal((B, T)).astype(np.float32), device=DEV)
        loss = (out.masked_fill(~ok, 0.0) * g).sum()
        (grad,) = torch.autograd.grad(loss, st)
        assert bool(torch.isfinite(grad).all())
        step = 1e-4 * torch.sign(grad)
        out2 = sem.robustness_trace(phi, (st + step).detach())
        ok2 = ok & torch.isfinite(out2)
        d_pred = float((grad * step).sum())
        d_true = float(((out2 - out.detach()).masked_fill(~ok2, 0.0) * g).sum())
        assert abs(d_pred - d_true) <= 0.05 * abs(d_pred) + 1e-4, (d_pred, d_true)


def test_operator_interface_lists():
    """What the reference's own recursion (api.py:142-302) calls: lists of per-step values, one t (ctstl.py:50-127)."""
    sem = B200CtstlSemantics(delta=0.5)
    y, l = GOLD["y"], GOLD["l"]
    dev = torch.device(DEV)
    vals = [torch.tensor(v, dtype=torch.float32, device=dev) for v in y[0, :9]]
    lv = [torch.tensor(v, dtype=torch.float32, device=dev) for v in l[0, :9]]
    assert float(sem.temporal_cumulative(vals, tau=1.5)) == np.float32(oc.temporal_cumulative(y[0, :9], tau=1.5, delta=0.5))
    assert float(sem.temporal_cumulative(vals, tau=1.0, delta=1.0)) == np.float32(y[0, :9].max())
    assert float(sem.temporal_cumulative(vals, tau=5.0)) == float("-inf")  # k = 10 > 9 values
    assert float(sem.temporal_cumulative([], tau=1.0)) == float("-inf")
    for a, b in ((0, 0), (0, 4), (2, None), (1, 30)):
        got = float(sem.temporal_until(left_trace=lv, right_trace=vals, start=a, end=b))
        assert got == np.float32(oc.temporal_until(l[0, :9], y[0, :9], a, b)), (a, b)
    with pytest.raises(ValueError):
        sem.temporal_until(left_trace=lv, right_trace=vals, start=9, end=None)
    with pytest.raises(ValueError):
        sem.temporal_until(left_trace=lv, right_trace=vals[:3], start=0, end=None)
    # batched values (a tensor per step), min / max of the classical lattice
    bv = [torch.tensor(y[:, i], dtype=torch.float32, device=dev) for i in range(9)]
    assert np.array_equal(sem.temporal_cumulative(bv, tau=1.5).cpu().numpy(), np.sort(y[:, :9], axis=1)[:, -3].astype(np.float32))
    assert np.array_equal(sem.boolean_and(bv).cpu().numpy(), y[:, :9].min(1).astype(np.float32))
    assert np.array_equal(sem.temporal_eventually(bv).cpu().numpy(), y[:, :9].max(1).astype(np.float32))


def test_mean_robustness_and_grad_of_a_formula_with_ct_nodes():
    x = affine_predicate("x", {0: 1.0}, 0.125)
    y = affine_predicate("y", {1: -1.0, 0: 0.25}, 0.375)
    phi = Or(Until(x, y, (1, 6)), CumulativeTime(And(x, Not(y)), (0, 5), tau=2.0))  # NaN at t = T - 1 (Until), no -inf (Or with the Until)
    B, T = 5, 64
    rng = np.random.default_rng(23)
    sig = (rng.uniform(-0.5, 0.5, size=(B, 1, 2)) + 0.5 * rng.standard_normal((B, T, 2))).astype(np.float32)
    sem = B200CtstlSemantics()
    st = torch.tensor(sig, device=DEV)
    mean, grad = sem.mean_robustness_and_grad(phi, st)
    xr = st.clone().requires_grad_(True)
    out = sem.robustness_trace(phi, xr)
    ok = ~torch.isnan(out)
    assert int(ok.sum()) == B * (T - 1)
    ref = out[ok].double().mean()
    (gref,) = torch.autograd.grad(ref, xr)
    ref = ref.detach()
    assert abs(float(mean) - float(ref)) <= 1e-6 * (1 + abs(float(ref)))
    assert torch.allclose(grad, gref.float(), rtol=1e-5, atol=1e-7)
    assert abs(float(sem.mean_robustness(phi, st)) - float(ref)) <= 1e-6 * (1 + abs(float(ref)))
    # a window shorter than k at the end of every trace: the mean is -inf, the gradient stays finite
    m2, g2 = sem.mean_robustness_and_grad(CumulativeTime(x, (0, 5), tau=3.0), st)
    assert float(m2) == float("-inf") and bool(torch.isfinite(g2).all())


def test_vmap_and_functional_grad_through_the_ct_nodes():
    """``torch.func.vmap`` over traces and ``vmap(grad)`` through the CT-STL ops (their batching rules, ops.py), like the formula kernels."""
    x = affine_predicate("x", {0: 1.0}, 0.125)
    y = affine_predicate("y", {1: -1.0, 0: 0.25}, 0.375)
    phi = And(Until(x, y, (0, 7)), Not(CumulativeTime(Or(x, y), (1, 9), tau=3.0)), Eventually(CumulativeTime(y, (0, 4), tau=2.0), (0, 3)))
    rng = np.random.default_rng(5)
    sig = torch.tensor((rng.uniform(-0.5, 0.5, size=(6, 1, 2)) + 0.5 * rng.standard_normal((6, 80, 2))).astype(np.float32), device=DEV)
    sem = B200CtstlSemantics()
    want = sem.robustness_trace(phi, sig)
    got = torch.func.vmap(lambda tr: sem.robustness_trace(phi, tr))(sig)
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    ok = ~torch.isnan(want)
    assert torch.equal(got[ok], want[ok])

    def loss(tr):
        out = sem.robustness_trace(phi, tr)
        return torch.where(torch.isfinite(out), out, torch.zeros_like(out)).sum()

    xs = sig.clone().requires_grad_(True)
    loss(xs).backward()
    g_f = torch.func.vmap(torch.func.grad(loss))(sig)
    assert torch.allclose(g_f, xs.grad, rtol=1e-6, atol=1e-7)
    # the value-trace ops directly, batched over a leading axis
    yy = torch.randn(3, 4, 50, device=DEV)
    ll = torch.randn(3, 4, 50, device=DEV)
    f = lambda v: sem.cumulative_time_trace(v, (0, 5), tau=2.0)
    assert torch.equal(torch.func.vmap(f)(yy), torch.stack([f(v) for v in yy]))
    u = lambda a, b: sem.until_trace(a, b, (1, 6))
    got_u, want_u = torch.func.vmap(u)(ll, yy), torch.stack([u(a, b) for a, b in zip(ll, yy)])
    assert torch.equal(torch.isnan(got_u), torch.isnan(want_u)) and torch.equal(got_u[~torch.isnan(want_u)], want_u[~torch.isnan(want_u)])
    assert torch.allclose(torch.func.vmap(u, in_dims=(None, 0))(ll[0], yy)[1], u(ll[0], yy[1]), rtol=0.0, atol=0.0, equal_nan=True)
