"""Host emulation of the FUSED tier (TEST INFRASTRUCTURE).

The fused tier compiles one CUDA kernel per formula at run time (pystl_b200/csrc/codegen.cpp +
fused_rt.cuh).  For the CPU test tier the very same generated source is compiled with g++ against
``tests/emu/simt_host.h`` (one fiber per CUDA thread) and run tile by tile, so that the window
index arithmetic, halos and barrier placement are checked against the oracle without a GPU.
The ``pystl_b200`` package never imports this.
"""

from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

import emu_lib
from pystl_b200 import capi
from pystl_b200.compiler import compile_formula

ROOT = emu_lib.ROOT
CSRC = os.path.join(ROOT, "pystl_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")
OUT = os.path.join(ROOT, "tests", "_emu", "fused")
_HDRS = [os.path.join(CSRC, h) for h in ("fused_rt.cuh", "tile_eval.h", "tile_bwd.h", "program.h")] + [
    os.path.join(EMU, h) for h in ("simt_host.h", "fused_runner.h")
]
_kernels: dict = {}
_sizes = None


def _affine_only(cf):
    """The host emulators run the kernels only: leaf channels (non-affine predicates) are computed
    by torch on the device in the product and have no place here."""
    if cf.leaf_preds:
        raise NotImplementedError("the host emulator does not evaluate leaf (non-affine) predicates")
    return cf


def _lib():
    h = emu_lib.lib()
    if not hasattr(h, "_fused_declared"):
        h.pystl_emu_fused_prepare.restype = C.c_int
        h.pystl_emu_fused_prepare.argtypes = [
            C.POINTER(capi.Formula), C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_char_p, C.c_int32,
            C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
        ]
        h._fused_declared = True
    return h


def sizes():
    global _sizes
    if _sizes is None:
        a, b = C.c_int32(), C.c_int32()
        _lib().pystl_emu_fused_sizes(C.byref(a), C.byref(b))
        _sizes = (a.value, b.value)
    return _sizes


def prepare(cf, B: int, T: int, D: int, backward: bool, force_threads: int):
    pb, gb = sizes()
    src = C.create_string_buffer(1 << 18)
    prog = C.create_string_buffer(pb)
    geom = C.create_string_buffer(gb)
    smem = C.c_int32()
    rc = _lib().pystl_emu_fused_prepare(cf.ptr(), B, T, D, int(backward), force_threads, src, len(src), prog, geom, C.byref(smem))
    capi.raise_for(rc, _lib().pystl_emu_last_error())
    return src.value.decode(), prog, geom, smem.value


def kernel(source: str, backward: bool):
    hdr = hashlib.sha1()
    for h in _HDRS:
        hdr.update(open(h, "rb").read())
    mufu = os.environ.get("PYSTL_EMU_MUFU", "")  # perturb the transcendentals like the hardware approximations (tile_eval.h)
    key = hashlib.sha1((source + hdr.hexdigest() + str(backward) + mufu).encode()).hexdigest()[:20]
    if key in _kernels:
        return _kernels[key]
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, key + ".so")
    if not os.path.exists(so):
        cu = os.path.join(OUT, key + ".cpp")
        with open(cu, "w") as fh:
            fh.write(source)
            fh.write('\n#include "fused_runner.h"\n')
        import re

        name = re.search(r"PF_KERNEL\(NT, \d+\) (\w+)\(", source).group(1)
        cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-I", CSRC, "-I", EMU, f"-DPF_EMU_KERNEL={name}"]
        if backward:
            cmd.append("-DPF_EMU_BACKWARD")
        if mufu:
            cmd.append(f"-DPYSTL_EMU_MUFU={int(mufu)}")
        tmp = so + f".{os.getpid()}.tmp"
        subprocess.run(cmd + [cu, "-o", tmp], check=True)
        os.replace(tmp, so)
    h = C.CDLL(so)
    h.pystl_fused_emu_run.restype = C.c_int
    h.pystl_fused_emu_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                      C.c_void_p, C.c_int32]
    _kernels[key] = h
    return h


def eval_forward(formula, semantics: str, signal: np.ndarray, *, force_threads: int = 0, mean: bool = False, **kw):
    sig = np.ascontiguousarray(signal, dtype=np.float32)
    B, T, D = sig.shape
    cf = _affine_only(compile_formula(formula, semantics, D, **kw))
    src, prog, geom, smem = prepare(cf, B, T, D, False, force_threads)
    k = kernel(src, False)
    ch = 2 if semantics == "cumulative" else 1
    out = np.full((B, T, ch), np.float32(-12345.0), dtype=np.float32)
    acc = np.zeros(2, dtype=np.float64)
    k.pystl_fused_emu_run(prog, geom, sig.ctypes.data, out.ctypes.data, acc.ctypes.data if mean else None, None, 0.0, None, smem)
    res = out if ch == 2 else out[..., 0]
    return (res, acc) if mean else res


def eval_backward(formula, semantics: str, signal: np.ndarray, grad_out=None, *, scale: float = 1.0,
                  force_threads: int = 0, **kw) -> np.ndarray:
    sig = np.ascontiguousarray(signal, dtype=np.float32)
    B, T, D = sig.shape
    cf = _affine_only(compile_formula(formula, semantics, D, **kw))
    src, prog, geom, smem = prepare(cf, B, T, D, True, force_threads)
    k = kernel(src, True)
    go = None if grad_out is None else np.ascontiguousarray(grad_out, dtype=np.float32)
    gs = np.full((B, T, D), np.float32(-12345.0), dtype=np.float32)
    k.pystl_fused_emu_run(prog, geom, sig.ctypes.data, None, None, None if go is None else go.ctypes.data, scale,
                          gs.ctypes.data, smem)
    return gs
