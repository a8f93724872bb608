"""Build / load the host emulation of the open-ended scan kernels (TEST INFRASTRUCTURE): ``pystl_b200/csrc/scan.cu`` compiled with
g++ against ``tests/emu/simt_host.h`` (every CUDA thread a fiber, one warp per trace as on the device).  The ``pystl_b200`` package
never imports this."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "pystl_b200", "csrc")
SRC = os.path.join(EMU, "scan_emu.cpp")
DEPS = [SRC, os.path.join(EMU, "scan_emu_shim.h"), os.path.join(EMU, "simt_host.h"), os.path.join(ROOT, "include", "pystl_b200.h")] + [
    os.path.join(CSRC, f) for f in ("scan.cu", "tile_bwd.h", "tile_eval.h", "program.h")
]
MUFU = os.environ.get("PYSTL_EMU_MUFU", "")  # perturb the transcendentals like the hardware approximations (tile_eval.h)
OUT = os.path.join(ROOT, "tests", "_emu", f"libscan_emu{'_mufu' + MUFU if MUFU else ''}.so")
SEM = {"classical": 0, "smooth": 1, "cumulative": 2, "agm": 3, "dgmsr": 4}  # include/pystl_b200.h: pystl_semantics_t
_lib = None


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    stale = force or not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS)
    if stale:
        tmp = f"{OUT}.{os.getpid()}.tmp"  # atomic publish: pytest-xdist workers may build concurrently
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-Wno-unused-function", "-I", EMU,
                        "-I", CSRC, SRC, "-o", tmp] + ([f"-DPYSTL_EMU_MUFU={int(MUFU)}"] if MUFU else []), check=True)
        os.replace(tmp, OUT)
    return OUT


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        h = C.CDLL(build())
        P, LL, I, F = C.c_void_p, C.c_longlong, C.c_int, C.c_float
        h.pystl_scan_emu_window_forward.argtypes = [I, I, I, F, F, I, P, LL, LL, P, P]
        h.pystl_scan_emu_window_backward.argtypes = [I, I, I, F, F, I, P, P, P, P, LL, LL, P]
        h.pystl_scan_emu_until_forward.argtypes = [I, P, P, LL, LL, P, P]
        h.pystl_scan_emu_until_backward.argtypes = [I, P, P, LL, LL, P, P]
        _lib = h
    return _lib


def _f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def window_forward(sem: str, eventually: bool, a: int, y, *, temperature=1.0, eps=1e-8, p=1, nan_tail=False):
    """out[B, T], arg[B, T] (arg is written for minima / maxima only) -- pystl_b200_open_window_forward without the stream."""
    y = _f32(y)
    B, T = y.shape
    out = np.full((B, T), 123.0, dtype=np.float32)
    arg = np.full((B, T), -7, dtype=np.int32)
    mode = int(eventually) + (2 if nan_tail else 0)
    lib().pystl_scan_emu_window_forward(SEM[sem], mode, int(a), temperature, eps, int(p), y.ctypes.data, B, T, out.ctypes.data, arg.ctypes.data)
    return out, arg


def window_backward(sem: str, eventually: bool, a: int, y, out, arg, g, *, temperature=1.0, eps=1e-8, p=1):
    y, out, g = _f32(y), _f32(out), _f32(g)
    arg = np.ascontiguousarray(arg, dtype=np.int32)
    B, T = y.shape
    gy = np.full((B, T), 123.0, dtype=np.float32)
    lib().pystl_scan_emu_window_backward(SEM[sem], int(eventually), int(a), temperature, eps, int(p), y.ctypes.data, out.ctypes.data,
                                         arg.ctypes.data, g.ctypes.data, B, T, gy.ctypes.data)
    return gy


def until_forward(a: int, l, r):
    l, r = _f32(l), _f32(r)
    B, T = l.shape
    out = np.full((B, T), 123.0, dtype=np.float32)
    key = np.full((B, T), -7, dtype=np.int32)
    lib().pystl_scan_emu_until_forward(int(a), l.ctypes.data, r.ctypes.data, B, T, out.ctypes.data, key.ctypes.data)
    return out, key


def until_backward(a: int, key, g):
    key = np.ascontiguousarray(key, dtype=np.int32)
    g = _f32(g)
    B, T = g.shape
    gl = np.full((B, T), 123.0, dtype=np.float32)
    gr = np.full((B, T), 123.0, dtype=np.float32)
    lib().pystl_scan_emu_until_backward(int(a), key.ctypes.data, g.ctypes.data, B, T, gl.ctypes.data, gr.ctypes.data)
    return gl, gr


def _ptrs(arrays):
    return (C.c_void_p * max(len(arrays), 1))(*[a.ctypes.data if a is not None else None for a in arrays])


def assemble(signal, channels, cols, templated=True):
    """out[r, :] = [signal[r, channels] | cols...] for rows r = (trace, t) -- pystl_b200_assemble_channels without the stream."""
    signal = _f32(signal)
    rows, D = signal.shape
    cols = [_f32(c) for c in cols]
    ch = np.ascontiguousarray(channels, dtype=np.int32)
    out = np.full((rows, len(ch) + len(cols)), 123.0, dtype=np.float32)
    h = lib()
    h.pystl_scan_emu_assemble.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    h.pystl_scan_emu_assemble(signal.ctypes.data, rows, D, ch.ctypes.data, len(ch), _ptrs(cols), len(cols), out.ctypes.data, int(templated))
    return out


def scatter(grad, D, channels, n_cols, want_signal=True, skip_cols=(), templated=True):
    """The adjoint: grad_signal[rows, D] (zeros in the channels nobody read) and one gradient column per leaf (None: not wanted)."""
    grad = _f32(grad)
    rows = grad.shape[0]
    ch = np.ascontiguousarray(channels, dtype=np.int32)
    gs = np.full((rows, D), 123.0, dtype=np.float32) if want_signal else None
    gcols = [None if j in skip_cols else np.full(rows, 123.0, dtype=np.float32) for j in range(n_cols)]
    h = lib()
    h.pystl_scan_emu_scatter.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    h.pystl_scan_emu_scatter(grad.ctypes.data, rows, D, ch.ctypes.data, len(ch), None if gs is None else gs.ctypes.data, _ptrs(gcols), n_cols,
                             int(templated))
    return gs, gcols
