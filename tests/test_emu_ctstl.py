"""CPU tier: the CT-STL kernels of ``pystl_b200/csrc/select.cu`` themselves, compiled for the host (``tests/select_emu_lib.py``:
every CUDA thread a fiber, real barriers), against the outputs of the unmodified reference class and the oracle -- the same
cases the GPU tier runs through the C ABI (``tests/test_gpu_ctstl.py``).  Index arithmetic of the 4-outputs-per-thread windows,
tile boundaries, staging, the key bisection with float compares, routing among equal samples and the gather adjoints are thereby
checked without a GPU."""

from __future__ import annotations

import os

import numpy as np
import pytest

import select_emu_lib as emu
from oracle import ctstl as oc

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ctstl_vectors.npz"))


def traces(seed, B, T, quant=0.0):
    rng = np.random.default_rng(seed)
    y = rng.uniform(-1, 1, size=(B, 1)) + rng.standard_normal((B, T))
    if quant:
        y = np.round(y / quant) * quant
    return y.astype(np.float32)


def test_cumulative_matches_reference_vectors():
    y = GOLD["y"].astype(np.float32)
    for c, (a, b, tau, delta) in enumerate(GOLD["cum_cases"]):
        a, b = int(a), None if b < 0 else int(b)
        k = oc.tau_to_k(tau, delta)
        out, arg = emu.cumulative_forward(y, a, b, k)
        assert np.array_equal(out, GOLD["cum"][c].astype(np.float32)), (a, b, tau, delta)
        _, want_arg = oc.cumulative_trace(y.astype(np.float64), a, b, k)
        assert np.array_equal(arg, want_arg), (a, b, tau, delta)


@pytest.mark.parametrize("B,T,a,b,k,quant", [
    (3, 1, 0, 0, 1, 0.0), (3, 1, 0, 5, 1, 0.0), (2, 1, 1, 5, 1, 0.0), (3, 513, 0, 7, 3, 0.0), (2, 1030, 2, 65, 17, 0.25),
    (2, 300, 0, 299, 150, 0.0), (1, 700, 5, 600, 1, 0.5), (1, 700, 5, 600, 596, 0.0), (2, 120, 0, None, 4, 0.5),
    (1, 640, 100, None, 77, 0.0), (5, 64, 0, 2, 3, 1.0), (2, 40, 0, 3, 2, 0.0), (2, 40, 0, 4, 5, 0.0), (3, 50, 60, 70, 1, 0.0),
])
def test_cumulative_against_oracle_with_gradient(B, T, a, b, k, quant):
    y = traces(B * 1000 + T, B, T, quant)
    out, arg = emu.cumulative_forward(y, a, b, k)
    want, want_arg = oc.cumulative_trace(y.astype(np.float64), a, b, k)
    assert np.array_equal(out, want.astype(np.float32))
    assert np.array_equal(arg, want_arg)
    out2, _ = emu.cumulative_forward(y, a, b, k, want_arg=False)
    assert np.array_equal(out2, out)
    g = np.random.default_rng(3).standard_normal((B, T)).astype(np.float32)
    gy = emu.cumulative_backward(arg, g, a, b)
    np.testing.assert_allclose(gy, oc.cumulative_grad(want_arg, g.astype(np.float64)), rtol=1e-5, atol=1e-5)


def test_cumulative_nan_and_special_values():
    y = traces(1, 2, 60)
    y[0, 30] = np.nan
    y[1, 10], y[1, 11], y[1, 12], y[1, 13] = np.inf, -np.inf, -0.0, 0.0
    out, arg = emu.cumulative_forward(y, 0, 4, 2)
    assert np.all(np.isnan(out[0, 26:31])) and not np.isnan(out[0, :26]).any() and not np.isnan(out[0, 31:]).any()
    assert np.all(arg[0, 26:31] == -1)
    want, want_arg = oc.cumulative_trace(y[1:].astype(np.float64), 0, 4, 2)
    assert np.array_equal(out[1:], want.astype(np.float32)) and np.array_equal(arg[1:], want_arg)
    top, _ = emu.cumulative_forward(y[1:], 0, 4, 1)
    assert np.all(np.isinf(top[0, 6:11]) & (top[0, 6:11] > 0))
    low, _ = emu.cumulative_forward(y[1:], 0, 4, 5)
    assert np.all(np.isinf(low[0, 7:12]) & (low[0, 7:12] < 0))


def test_until_matches_reference_vectors_and_routes_to_the_value():
    y, l = GOLD["y"].astype(np.float32), GOLD["l"].astype(np.float32)
    for c, (a, b) in enumerate(GOLD["until_cases"]):
        a, b = int(a), None if b < 0 else int(b)
        out, key = emu.until_forward(l, y, a, b)
        assert np.array_equal(out, GOLD["until"][c].astype(np.float32), equal_nan=True), (a, b)
        ok = ~np.isnan(out)
        assert np.array_equal(key >= 0, ok)
        rows, _ = np.nonzero(ok)
        kk = key[ok]
        assert np.array_equal(np.where(kk & 1, y[rows, kk >> 1], l[rows, kk >> 1]), out[ok])
        g = np.random.default_rng(c).standard_normal(out.shape).astype(np.float32)
        gl, gr = emu.until_backward(key, g, a, b)
        wl, wr = np.zeros(out.shape), np.zeros(out.shape)
        for i, t in zip(*np.nonzero(key >= 0)):
            (wr if key[i, t] & 1 else wl)[i, key[i, t] >> 1] += g[i, t]
        np.testing.assert_allclose(gl, wl, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(gr, wr, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,T,a,b", [(3, 1, 0, 0), (3, 1, 0, None), (4, 300, 0, 24), (2, 1000, 5, 64), (2, 400, 0, None), (2, 90, 7, 7), (2, 50, 60, 80)])
def test_until_against_oracle(B, T, a, b):
    l, r = traces(B + T, B, T, 0.25), traces(B + T + 1, B, T, 0.0)
    out, _ = emu.until_forward(l, r, a, b)
    want = oc.until_trace(l.astype(np.float64), r.astype(np.float64), a, b)
    assert np.array_equal(out, want.astype(np.float32), equal_nan=True)
