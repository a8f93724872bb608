"""Build / load the host emulator of the CUDA tile kernel (TEST INFRASTRUCTURE).

``tests/emu/emu.cpp`` compiles the single-source CTA body
(``pystl_b200/csrc/tile_eval.h``) and the lowering (``lower.cpp``) with g++; the
CPU test tier uses it to check the kernel logic against the oracle.  The
``pystl_b200`` package never imports this.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pystl_b200 import capi
from pystl_b200.compiler import compile_formula

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "tests", "emu", "emu.cpp"), os.path.join(ROOT, "pystl_b200", "csrc", "lower.cpp"),
       os.path.join(ROOT, "pystl_b200", "csrc", "codegen.cpp")]
DEPS = SRC + [
    os.path.join(ROOT, "pystl_b200", "csrc", "tile_eval.h"),
    os.path.join(ROOT, "pystl_b200", "csrc", "tile_bwd.h"),
    os.path.join(ROOT, "pystl_b200", "csrc", "program.h"),
    os.path.join(ROOT, "pystl_b200", "csrc", "fused_host.h"),
    os.path.join(ROOT, "include", "pystl_b200.h"),
]
OUT = os.path.join(ROOT, "tests", "_emu", "libpystl_emu.so")
_lib = None


def _affine_only(cf):
    """The host emulators run the kernels only: leaf channels (non-affine predicates) are computed
    by torch on the device in the product and have no place here."""
    if cf.leaf_preds:
        raise NotImplementedError("the host emulator does not evaluate leaf (non-affine) predicates")
    return cf


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    stale = force or not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS)
    if stale:
        cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "pystl_b200", "csrc"), *SRC, "-o"]
        tmp = f"{OUT}.{os.getpid()}.tmp"  # atomic publish: pytest-xdist workers may build concurrently
        subprocess.run(cmd + [tmp], check=True)
        os.replace(tmp, OUT)
    return OUT


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        handle = C.CDLL(build())
        capi.declare(handle, prefix="pystl_emu")
        handle.pystl_emu_eval_forward.restype = C.c_int
        handle.pystl_emu_eval_forward.argtypes = [
            C.POINTER(capi.Formula), C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
        ]
        handle.pystl_emu_eval_backward.restype = C.c_int
        handle.pystl_emu_eval_backward.argtypes = [
            C.POINTER(capi.Formula), C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_float, C.c_void_p,
            C.c_int32,
        ]
        _lib = handle
    return _lib


def plan(formula, semantics: str, T: int, D: int, **kw) -> dict:
    cf = _affine_only(compile_formula(formula, semantics, D, **kw))
    p = capi.Plan()
    rc = lib().pystl_emu_plan(cf.ptr(), T, D, 0, C.byref(p))
    capi.raise_for(rc, lib().pystl_emu_last_error())
    return p.as_dict()


def eval_forward(formula, semantics: str, signal: np.ndarray, *, force_threads: int = 0, mean: bool = False, **kw):
    """signal [B, T, D] float32 -> out [B, T] (cumulative [B, T, 2]) float32."""
    sig = np.ascontiguousarray(signal, dtype=np.float32)
    B, T, D = sig.shape
    cf = _affine_only(compile_formula(formula, semantics, D, **kw))
    ch = 2 if semantics == "cumulative" else 1
    out = np.full((B, T, ch), np.float32(-12345.0), dtype=np.float32)
    acc = np.zeros(2, dtype=np.float64)
    rc = lib().pystl_emu_eval_forward(
        cf.ptr(), sig.ctypes.data, B, T, D, out.ctypes.data, acc.ctypes.data if mean else None, force_threads
    )
    capi.raise_for(rc, lib().pystl_emu_last_error())
    res = out if ch == 2 else out[..., 0]
    return (res, acc) if mean else res


def eval_backward(formula, semantics: str, signal: np.ndarray, grad_out=None, *, scale: float = 1.0,
                  force_threads: int = 0, **kw) -> np.ndarray:
    """d sum(grad_out * out) / d signal, [B, T, D] float32 (grad_out None: constant cotangent `scale`)."""
    sig = np.ascontiguousarray(signal, dtype=np.float32)
    B, T, D = sig.shape
    cf = _affine_only(compile_formula(formula, semantics, D, **kw))
    go = None if grad_out is None else np.ascontiguousarray(grad_out, dtype=np.float32)
    gs = np.full((B, T, D), np.float32(-12345.0), dtype=np.float32)
    rc = lib().pystl_emu_eval_backward(
        cf.ptr(), sig.ctypes.data, B, T, D, None if go is None else go.ctypes.data, scale, gs.ctypes.data, force_threads
    )
    capi.raise_for(rc, lib().pystl_emu_last_error())
    return gs
