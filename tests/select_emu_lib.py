"""Build / load the host emulation of the CT-STL kernels (TEST INFRASTRUCTURE): ``pystl_b200/csrc/select.cu`` compiled with g++
against ``tests/emu/simt_host.h`` (every CUDA thread a fiber).  The ``pystl_b200`` package never imports this."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "select_emu.cpp")
DEPS = [SRC, os.path.join(ROOT, "tests", "emu", "select_emu_shim.h"), os.path.join(ROOT, "tests", "emu", "simt_host.h"),
        os.path.join(ROOT, "pystl_b200", "csrc", "select.cu"), os.path.join(ROOT, "include", "pystl_b200.h")]
OUT = os.path.join(ROOT, "tests", "_emu", "libselect_emu.so")
_lib = None


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    stale = force or not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS)
    if stale:
        tmp = f"{OUT}.{os.getpid()}.tmp"  # atomic publish: pytest-xdist workers may build concurrently
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-Wno-unused-function", "-I",
                        os.path.join(ROOT, "tests", "emu"), SRC, "-o", tmp], check=True)
        os.replace(tmp, OUT)
    return OUT


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        h = C.CDLL(build())
        h.pystl_select_emu_cumulative_forward.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        h.pystl_select_emu_cumulative_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_void_p]
        h.pystl_select_emu_until_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        h.pystl_select_emu_until_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib = h
    return _lib


def _end(b):
    return -1 if b is None else int(b)


def cumulative_forward(y: np.ndarray, a: int, b, k: int, want_arg: bool = True):
    y = np.ascontiguousarray(y, dtype=np.float32)
    B, T = y.shape
    out = np.full((B, T), 123.0, dtype=np.float32)
    arg = np.full((B, T), -7, dtype=np.int32)
    lib().pystl_select_emu_cumulative_forward(y.ctypes.data, B, T, int(a), _end(b), int(k), out.ctypes.data, arg.ctypes.data if want_arg else None)
    return out, arg


def cumulative_backward(arg: np.ndarray, g: np.ndarray, a: int, b):
    arg = np.ascontiguousarray(arg, dtype=np.int32)
    g = np.ascontiguousarray(g, dtype=np.float32)
    B, T = g.shape
    gy = np.full((B, T), 123.0, dtype=np.float32)
    lib().pystl_select_emu_cumulative_backward(arg.ctypes.data, g.ctypes.data, B, T, int(a), _end(b), gy.ctypes.data)
    return gy


def until_forward(l: np.ndarray, r: np.ndarray, a: int, b):
    l = np.ascontiguousarray(l, dtype=np.float32)
    r = np.ascontiguousarray(r, dtype=np.float32)
    B, T = l.shape
    out = np.full((B, T), 123.0, dtype=np.float32)
    key = np.full((B, T), -7, dtype=np.int32)
    lib().pystl_select_emu_until_forward(l.ctypes.data, r.ctypes.data, B, T, int(a), _end(b), out.ctypes.data, key.ctypes.data)
    return out, key


def until_backward(key: np.ndarray, g: np.ndarray, a: int, b):
    key = np.ascontiguousarray(key, dtype=np.int32)
    g = np.ascontiguousarray(g, dtype=np.float32)
    B, T = g.shape
    gl = np.full((B, T), 123.0, dtype=np.float32)
    gr = np.full((B, T), 123.0, dtype=np.float32)
    lib().pystl_select_emu_until_backward(key.ctypes.data, g.ctypes.data, B, T, int(a), _end(b), gl.ctypes.data, gr.ctypes.data)
    return gl, gr
