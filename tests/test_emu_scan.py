"""CPU tier: the open-ended scan kernels of ``pystl_b200/csrc/scan.cu`` themselves, compiled for the host (``tests/scan_emu_lib.py``:
every CUDA thread a fiber, warp shuffles through the fiber scheduler), against the oracle -- the route ``G[a, inf)`` / ``F[a, inf)``
and the classical ``U[a, inf)`` take on traces longer than one tile (``api.py:49-54``: an interval without an end is clipped at the
horizon).  The chunked suffix / prefix scans with their carries across 256-sample chunks, the lane-segment hand-off, the padded
staging buffers, the arg-min / arg-max provenance and the segmented-run adjoint are thereby checked without a GPU; the GPU tier
runs the same kernels through the C ABI (``tests/test_gpu_open_window.py``)."""

from __future__ import annotations

import numpy as np
import pytest

import scan_emu_lib as emu
from parity_util import SEM_KW, assert_parity, oracle_trace
from test_emu_backward import check, grad_tol, oracle_grad

X = ("pred", {0: 1.0}, 0.0)  # the operand's value trace IS the signal: predicate x_0 - 0
Y = ("pred", {1: 1.0}, 0.0)
KW = {name: kw for name, kw in SEM_KW}


def traces(seed, B, T):
    """Per-trace offsets so that both branches of AGM / D-GMSR occur (all-positive suffixes and suffixes with a violation)."""
    rng = np.random.default_rng(seed)
    mu = rng.uniform(-0.5, 2.5, size=(B, 1))
    return (mu + 0.6 * rng.standard_normal((B, T))).astype(np.float32).astype(np.float64)


# T around the 8-sample lane segments, the 256-sample chunks and several chunks; a = 0, inside a chunk, beyond a chunk, beyond T
SHAPES = [(3, 1, 0), (5, 7, 2), (2, 255, 0), (3, 256, 5), (6, 257, 1), (2, 777, 300), (5, 1300, 11), (2, 40, 64)]


@pytest.mark.parametrize("B,T,a", SHAPES)
@pytest.mark.parametrize("op", ["always", "eventually"])
@pytest.mark.parametrize("sem", ["classical", "smooth", "agm", "dgmsr"])
def test_open_window_values_and_gradients(sem, op, B, T, a):
    kw = KW[sem]
    y = traces(T + a, B, T)
    spec = (op, (a, None), X)
    want = oracle_trace(spec, sem, kw, y[..., None])
    out, arg = emu.window_forward(sem, op == "eventually", a, y, **_params(sem, kw))
    n = max(T - a, 0)  # outputs t < n exist; the others raise in the reference (api.py:218-221)
    assert np.isnan(want[:, n:]).all()
    if n:
        assert_parity(out[:, :n], want[:, :n], sem, f"scan/{sem}/{op}/T{T}/a{a}")
    if sem == "classical":
        bad = np.inf if op == "always" else -np.inf
        w = np.where(np.arange(T)[None, :, None] + a <= np.arange(T)[None, None, :], y[:, None, :], bad)  # [b, t, k]: y_k inside the window of t
        first = (np.argmin if op == "always" else np.argmax)(w, axis=2)
        assert np.array_equal(arg[:, :n], first[:, :n]), "arg-min / arg-max: first index on ties, like torch.min / max(dim)"
        assert (arg[:, n:] == -1).all()
    # outputs the reference raises on: NaN on request (the values are the formula's robustness), else finite (a parent kernel reads them)
    out_nan, _ = emu.window_forward(sem, op == "eventually", a, y, nan_tail=True, **_params(sem, kw))
    assert np.isnan(out_nan[:, n:]).all() and np.array_equal(out_nan[:, :n], out[:, :n])
    assert np.isfinite(out).all()
    if not n:
        return
    rng = np.random.default_rng(T)
    g = rng.standard_normal((B, T)).astype(np.float32).astype(np.float64)
    g[:, n:] = 0.0
    gy = emu.window_backward(sem, op == "eventually", a, y, out, arg, g, **_params(sem, kw))
    rtol, atol = grad_tol(sem, kw)
    want_g = oracle_grad(spec, sem, kw, y[..., None], g)[..., 0]
    # (D-GMSR over hundreds of all-positive samples: the reference's product of v^(2w), _dgmsr.py:97-99, overflows float64 and its
    # autograd returns NaN for that trace; the kernels work in the log domain and stay finite -- nothing to compare against there)
    ok = np.isfinite(want_g).all(axis=1)
    assert ok.any() and (sem == "dgmsr" or ok.all())
    assert np.isfinite(gy).all()
    check(gy[ok].astype(np.float64), want_g[ok], sem, f"scan/{sem}/{op}/T{T}/a{a}", rtol, atol)


def _params(sem, kw):
    if sem == "smooth":
        return {"temperature": kw.get("temperature", 1.0)}
    if sem == "dgmsr":
        return {"eps": kw.get("eps", 1e-8), "p": kw.get("p", 1)}
    return {}


@pytest.mark.parametrize("B,T,a", SHAPES)
def test_cumulative_eventually_is_the_suffix_sum_of_one_channel(B, T, a):
    """cumulative.py:62-73: Eventually SUMS the window, per channel of the (pos, neg) pair; the adjoint is the prefix sum of the cotangents."""
    y = np.maximum(traces(T, B, T), 0.0)
    out, _ = emu.window_forward("cumulative", True, a, y)
    n = max(T - a, 0)
    want = np.cumsum(y[:, ::-1], axis=1)[:, ::-1]
    if n:
        assert np.allclose(out[:, :n], want[:, a:], rtol=1e-5, atol=5e-6)
        g = np.random.default_rng(a).standard_normal((B, T)).astype(np.float32)
        gy = emu.window_backward("cumulative", True, a, y, out, np.zeros((B, T), np.int32), g)
        want_g = np.zeros((B, T))
        want_g[:, a:] = np.cumsum(g[:, :n].astype(np.float64), axis=1)
        assert np.allclose(gy, want_g, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("B,T,a", SHAPES)
def test_classical_until_without_an_end(B, T, a):
    """W(u) = min(l_u, max(r_u, W(u + 1))) (classic.py:59-80 with Interval.end None), out[t] = W(t + a), bit-exact;
    key[t] = 2 * sample + side names the ONE leaf sample the value comes from, and the adjoint scatters the cotangents there."""
    rng = np.random.default_rng(100 + T)
    l = traces(T, B, T) + 0.3
    r = traces(T + 1, B, T) - 0.8
    if T > 20:
        l[0] = np.round(l[0] * 2) / 2  # exact ties between samples
        r[0] = np.round(r[0] * 2) / 2
    out, key = emu.until_forward(a, l, r)
    W = np.empty((B, T))
    acc = np.full(B, -np.inf)
    for u in range(T - 1, -1, -1):
        acc = np.minimum(l[:, u], np.maximum(r[:, u], acc))
        W[:, u] = acc
    n = max(T - a, 0)
    assert np.array_equal(out[:, :n], W[:, a:].astype(np.float32))
    assert (key[:, n:] == -1).all()
    if not n:
        return
    if a == 0:  # the whole operator: against the oracle's Until
        want = oracle_trace(("until", (0, None), X, Y), "classical", {}, np.stack([l, r], -1))
        assert np.array_equal(out, want.astype(np.float32))
    # the key is a sample inside the suffix whose value IS the output
    k = key[:, :n]
    src = np.where(k & 1, np.take_along_axis(r, k >> 1, 1), np.take_along_axis(l, k >> 1, 1))
    assert np.array_equal(src.astype(np.float32), out[:, :n])
    assert ((k >> 1) >= np.arange(n)[None, :] + a).all() and ((k >> 1) < T).all()
    g = rng.standard_normal((B, T)).astype(np.float32)
    g[:, n:] = 0.0
    gl, gr = emu.until_backward(a, key, g)
    want_l, want_r = np.zeros((B, T)), np.zeros((B, T))
    rows = np.repeat(np.arange(B)[:, None], n, 1)
    np.add.at(want_l, (rows[(k & 1) == 0], (k >> 1)[(k & 1) == 0]), g[:, :n][(k & 1) == 0].astype(np.float64))
    np.add.at(want_r, (rows[(k & 1) == 1], (k >> 1)[(k & 1) == 1]), g[:, :n][(k & 1) == 1].astype(np.float64))
    assert np.allclose(gl, want_l, rtol=1e-5, atol=1e-5) and np.allclose(gr, want_r, rtol=1e-5, atol=1e-5)
    assert np.array_equal(gl != 0, want_l != 0) and np.array_equal(gr != 0, want_r != 0)
    if a == 0 and T > 1:
        # un-tied traces: the routing equals autograd through the reference's fold (torch/semantics.py:115-146)
        sig = np.stack([l, r], -1)[1:]
        want_sig = oracle_grad(("until", (0, None), X, Y), "classical", {}, sig, g[1:].astype(np.float64))
        if B > 1:
            check(np.stack([gl[1:], gr[1:]], -1).astype(np.float64), want_sig, "classical", f"scan/until/T{T}")


@pytest.mark.parametrize("templated", [True, False])
@pytest.mark.parametrize("rows,D,channels,L", [(1, 2, (0,), 1), (255, 4, (2, 0), 2), (1000, 3, (0, 1, 2), 1), (777, 8, (), 3), (2049, 5, (4, 1), 0),
                                               (300, 9, (8, 0, 3), 7)])
def test_assemble_and_scatter_rows(rows, D, channels, L, templated):
    """The input of a parent formula kernel: the signal channels its predicates read next to one column per open-ended leaf, and the
    adjoint (gradient columns back to the leaves, the read channels back into a zero-filled grad_signal); tile-stride loop on 3 CTAs."""
    rng = np.random.default_rng(rows + D)
    sig = rng.standard_normal((rows, D)).astype(np.float32)
    cols = [rng.standard_normal(rows).astype(np.float32) for _ in range(L)]
    out = emu.assemble(sig, channels, cols, templated)
    want = np.concatenate([sig[:, list(channels)].reshape(rows, len(channels))] + [c[:, None] for c in cols], axis=1)
    assert np.array_equal(out, want)
    g = rng.standard_normal(out.shape).astype(np.float32)
    skip = (1,) if L > 1 else ()
    gs, gcols = emu.scatter(g, D, channels, L, want_signal=True, skip_cols=skip, templated=templated)
    want_gs = np.zeros((rows, D), np.float32)
    for i, d in enumerate(channels):
        want_gs[:, d] = g[:, i]
    assert np.array_equal(gs, want_gs)
    for j in range(L):
        if j in skip:
            assert gcols[j] is None
        else:
            assert np.array_equal(gcols[j], g[:, len(channels) + j])
    if L:  # no signal gradient wanted (the signal does not require grad)
        _, only = emu.scatter(g, D, channels, L, want_signal=False, templated=templated)
        assert all(np.array_equal(only[j], g[:, len(channels) + j]) for j in range(L))
