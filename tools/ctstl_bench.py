"""Timing of the CT-STL kernels alone (csrc/select.cu) on value traces resident in HBM: CUDA events around 10 launches
after 3 warm-ups, one JSON line per case.  Algorithmic bytes: 4 read + 4 written (+ 4 for the recorded index) per sample."""

from __future__ import annotations

import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pystl_b200 import capi  # noqa: E402

PEAK = 6537.6  # GB/s, MEASURED_PEAKS.json


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main() -> None:
    lib = capi.lib()
    st = torch.cuda.current_stream().cuda_stream
    B, T = 16384, 4096
    gen = torch.Generator(device="cuda").manual_seed(1)
    y = torch.randn(B, T, device="cuda", generator=gen)
    l = torch.randn(B, T, device="cuda", generator=gen)
    out, arg = torch.empty_like(y), torch.empty((B, T), dtype=torch.int32, device="cuda")
    g, gy, gl = torch.randn_like(y), torch.empty_like(y), torch.empty_like(y)
    for a, b, k in ((0, 15, 4), (0, 63, 16), (0, 63, 1), (0, 255, 64)):
        for with_arg in (False, True):
            ms = timed(lambda: lib.pystl_b200_ctstl_cumulative_forward(y.data_ptr(), B, T, a, b, k, out.data_ptr(), arg.data_ptr() if with_arg else None, st))
            nbytes = B * T * (12 if with_arg else 8)
            print(json.dumps({"kernel": "kth_window_fwd", "B": B, "T": T, "interval": [a, b], "k": k, "arg": with_arg, "ms": round(ms, 4),
                              "samples_per_s": B * T / ms * 1e3, "hbm_frac": nbytes / ms / 1e6 / PEAK,
                              "compares_per_s": 32.0 * (b - a + 4) * B * T / ms * 1e3}))
        ms = timed(lambda: lib.pystl_b200_ctstl_cumulative_backward(arg.data_ptr(), g.data_ptr(), B, T, a, b, gy.data_ptr(), st))
        print(json.dumps({"kernel": "kth_window_bwd", "B": B, "T": T, "interval": [a, b], "ms": round(ms, 4), "hbm_frac": B * T * 12 / ms / 1e6 / PEAK}))
    for a, b in ((0, 24), (0, 64), (5, 200)):
        ms = timed(lambda: lib.pystl_b200_ctstl_until_forward(l.data_ptr(), y.data_ptr(), B, T, a, b, out.data_ptr(), arg.data_ptr(), st))
        print(json.dumps({"kernel": "ctstl_until_fwd", "B": B, "T": T, "interval": [a, b], "ms": round(ms, 4), "hbm_frac": B * T * 16 / ms / 1e6 / PEAK}))
        ms = timed(lambda: lib.pystl_b200_ctstl_until_backward(arg.data_ptr(), g.data_ptr(), B, T, a, b, gl.data_ptr(), gy.data_ptr(), st))
        print(json.dumps({"kernel": "ctstl_until_bwd", "B": B, "T": T, "interval": [a, b], "ms": round(ms, 4), "hbm_frac": B * T * 16 / ms / 1e6 / PEAK}))


if __name__ == "__main__":
    main()
