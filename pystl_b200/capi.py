"""ctypes binding of ``include/pystl_b200.h`` (the C ABI of the CUDA library).

This is the stub INTEGRATION.md shows: plain pointers and sizes, torch only
supplies ``data_ptr()`` and the current stream.  There is no fallback: if
``libpystl_b200.so`` has not been built, :func:`lib` raises and so does every
backend call.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

ABI_VERSION = 2
MAX_TERMS = 8

OK, E_INVALID, E_UNSUPPORTED, E_CUDA = 0, -1, -2, -3

SEM_CLASSIC, SEM_SMOOTH, SEM_CUMULATIVE, SEM_AGM, SEM_DGMSR = range(5)
NODE_PRED, NODE_NOT, NODE_AND, NODE_OR, NODE_ALWAYS, NODE_EVENTUALLY, NODE_UNTIL = range(7)


class Node(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("n_children", C.c_int32),
        ("child_off", C.c_int32),
        ("a", C.c_int32),
        ("b", C.c_int32),
        ("pred", C.c_int32),
        ("w_off", C.c_int32),
        ("w_len", C.c_int32),
        ("w2_off", C.c_int32),
        ("w2_len", C.c_int32),
        ("wp0", C.c_float),
        ("wp1", C.c_float),
    ]


class Pred(C.Structure):
    _fields_ = [
        ("n_terms", C.c_int32),
        ("channel", C.c_int32 * MAX_TERMS),
        ("coeff", C.c_float * MAX_TERMS),
        ("bias", C.c_float),
        ("slack", C.c_int32),  # ABI 2: > 0 for a host-computed channel that ends before the horizon, -1 = never valid
        ("pair", C.c_int32),   # ABI 2, cumulative: the value is (signal[t, channel[0]], signal[t, channel[1]]) as it stands
    ]


class Formula(C.Structure):
    _fields_ = [
        ("semantics", C.c_int32),
        ("temperature", C.c_float),
        ("eps", C.c_float),
        ("p", C.c_int32),
        ("n_nodes", C.c_int32),
        ("nodes", C.POINTER(Node)),
        ("n_children", C.c_int32),
        ("children", C.POINTER(C.c_int32)),
        ("n_preds", C.c_int32),
        ("preds", C.POINTER(Pred)),
        ("n_weights", C.c_int32),
        ("weights", C.POINTER(C.c_float)),
    ]


class Plan(C.Structure):
    _fields_ = [
        ("out_channels", C.c_int32),
        ("lookahead", C.c_int32),
        ("start_slack", C.c_int32),
        ("raises_always", C.c_int32),
        ("threads", C.c_int32),
        ("tile_ext", C.c_int32),
        ("tile_out", C.c_int32),
        ("n_tiles", C.c_int32),
        ("smem_bytes", C.c_int32),
        ("n_instr", C.c_int32),
        ("n_planes", C.c_int32),
    ]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


EXPORTS = (
    "pystl_b200_abi_version",
    "pystl_b200_last_error",
    "pystl_b200_plan",
    "pystl_b200_eval_forward",
    "pystl_b200_eval_backward",
    "pystl_b200_eval_value_and_grad",
    "pystl_b200_fused_precompile",
    "pystl_b200_fused_plan",
    "pystl_b200_fused_source",
    "pystl_b200_open_window_forward",
    "pystl_b200_open_window_backward",
    "pystl_b200_open_until_forward",
    "pystl_b200_open_until_backward",
    "pystl_b200_assemble_channels",
    "pystl_b200_scatter_channels",
    "pystl_b200_ctstl_cumulative_forward",
    "pystl_b200_ctstl_cumulative_backward",
    "pystl_b200_ctstl_until_forward",
    "pystl_b200_ctstl_until_backward",
)

# PYSTL_B200_LIB points at an alternative build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("PYSTL_B200_LIB") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "csrc", "libpystl_b200.so"
)
_lib: Optional[C.CDLL] = None


def declare(lib: C.CDLL, prefix: str = "pystl_b200") -> None:
    """Attach argtypes / restypes (shared with the test-only host emulator)."""
    f = getattr(lib, f"{prefix}_last_error")
    f.restype = C.c_char_p
    f.argtypes = []
    f = getattr(lib, f"{prefix}_plan")
    f.restype = C.c_int
    f.argtypes = [C.POINTER(Formula), C.c_int64, C.c_int32, C.c_int32, C.POINTER(Plan)]
    if prefix == "pystl_b200":
        lib.pystl_b200_abi_version.restype = C.c_int
        lib.pystl_b200_abi_version.argtypes = []
        f = lib.pystl_b200_eval_forward
        f.restype = C.c_int
        f.argtypes = [
            C.POINTER(Formula), C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
            C.c_void_p, C.c_void_p, C.c_void_p,
        ]
        f = lib.pystl_b200_eval_backward
        f.restype = C.c_int
        f.argtypes = [
            C.POINTER(Formula), C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
            C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
        ]
        f = lib.pystl_b200_eval_value_and_grad
        f.restype = C.c_int
        f.argtypes = [
            C.POINTER(Formula), C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
            C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        ]
        f = lib.pystl_b200_fused_precompile
        f.restype = C.c_int
        f.argtypes = [C.POINTER(Formula), C.c_int64, C.c_int32, C.c_int32]
        f = lib.pystl_b200_fused_plan
        f.restype = C.c_int
        f.argtypes = [C.POINTER(Formula), C.c_int64, C.c_int32, C.c_int32, C.POINTER(Plan)]
        f = lib.pystl_b200_fused_source
        f.restype = C.c_int
        f.argtypes = [C.POINTER(Formula), C.c_int64, C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
        f = lib.pystl_b200_open_window_forward
        f.restype = C.c_int
        f.argtypes = [
            C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32,
            C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
        ]
        f = lib.pystl_b200_open_window_backward
        f.restype = C.c_int
        f.argtypes = [
            C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
        ]
        f = lib.pystl_b200_open_until_forward
        f.restype = C.c_int
        f.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        f = lib.pystl_b200_open_until_backward
        f.restype = C.c_int
        f.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        f = lib.pystl_b200_assemble_channels
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_void_p]
        f = lib.pystl_b200_scatter_channels
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]
        f = lib.pystl_b200_ctstl_cumulative_forward
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        f = lib.pystl_b200_ctstl_cumulative_backward
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        f = lib.pystl_b200_ctstl_until_forward
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        f = lib.pystl_b200_ctstl_until_backward
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]


def lib() -> C.CDLL:
    """The CUDA library.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C pystl_b200/csrc`). The b200 backend has no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        declare(handle)
        if handle.pystl_b200_abi_version() != ABI_VERSION:
            raise RuntimeError("libpystl_b200.so was built for a different ABI version; rebuild it")
        _lib = handle
    return _lib


def raise_for(rc: int, message: bytes | str) -> None:
    """Map a pystl_status to the exception class the reference raises."""
    if rc == OK:
        return
    msg = message.decode() if isinstance(message, bytes) else str(message)
    if rc == E_INVALID:
        raise ValueError(msg)
    if rc == E_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f"pystl_b200: {msg} (status {rc})")
