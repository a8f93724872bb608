"""Time the open-ended window kernels (csrc/scan.cu) on their own: GB/s of algorithmic traffic.

forward  reads y, writes out (+ arg for classical): 8 (12) bytes per sample
backward reads y, out, grad_out, writes grad_y:     16 bytes per sample (classical: arg + grad_out read, grad_y zeroed: 12)
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pystl_b200 import capi, ops  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--traces", type=int, default=16384)
    ap.add_argument("--horizon", type=int, default=16384)
    ap.add_argument("--start", type=int, default=5)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warm", type=int, default=20)
    args = ap.parse_args()
    B, T = args.traces, args.horizon
    y = torch.randn(B, T, device="cuda") * 0.5 + 0.1
    g = torch.rand(B, T, device="cuda") + 0.25
    rows = []
    for _ in range(args.warm):  # clocks up before the first timed loop
        ops.open_window_fwd(y, capi.SEM_SMOOTH, 0, args.start, 0.5, 1e-8, 1)
    for name, sem, tau in (("classical", capi.SEM_CLASSIC, 1.0), ("smooth", capi.SEM_SMOOTH, 0.5), ("agm", capi.SEM_AGM, 1.0),
                           ("dgmsr", capi.SEM_DGMSR, 1.0)):
        for mode in (0, 1):
            out, arg = ops.open_window_fwd(y, sem, mode, args.start, tau, 1e-8, 1)
            gy = ops.open_window_grad(y, out, arg, g, sem, mode, args.start, tau, 1e-8, 1)
            for _ in range(2):  # the caching allocator gets its blocks before the timed loops
                ops.open_window_fwd(y, sem, mode, args.start, tau, 1e-8, 1)
                ops.open_window_grad(y, out, arg, g, sem, mode, args.start, tau, 1e-8, 1)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            for _ in range(args.iters):
                ops.open_window_fwd(y, sem, mode, args.start, tau, 1e-8, 1)
            ev[1].record()
            for _ in range(args.iters):
                ops.open_window_grad(y, out, arg, g, sem, mode, args.start, tau, 1e-8, 1)
            ev[2].record()
            torch.cuda.synchronize()
            f_ms = ev[0].elapsed_time(ev[1]) / args.iters
            b_ms = ev[1].elapsed_time(ev[2]) / args.iters
            fb = (12 if sem == capi.SEM_CLASSIC else 8) * B * T
            bb = (12 if sem == capi.SEM_CLASSIC else 16) * B * T
            rows.append(dict(semantics=name, mode="always" if mode == 0 else "eventually", traces=B, horizon=T,
                             fwd_ms=round(f_ms, 4), fwd_gbs=round(fb / f_ms / 1e6, 1), bwd_ms=round(b_ms, 4), bwd_gbs=round(bb / b_ms / 1e6, 1),
                             finite=bool(torch.isfinite(gy).all())))
            print(json.dumps(rows[-1]), flush=True)
    # classical Until without an end: reads l, r, writes out, key (16 B / sample); backward reads key, grad_out, zero-fills grad_l, grad_r
    r2 = torch.randn(B, T, device="cuda") * 0.5
    out, key = ops.open_until_fwd(y, r2, capi.SEM_CLASSIC, args.start)
    for _ in range(2):
        ops.open_until_fwd(y, r2, capi.SEM_CLASSIC, args.start)
        ops.open_until_grad(key, g, capi.SEM_CLASSIC, args.start)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(args.iters):
        ops.open_until_fwd(y, r2, capi.SEM_CLASSIC, args.start)
    ev[1].record()
    for _ in range(args.iters):
        ops.open_until_grad(key, g, capi.SEM_CLASSIC, args.start)
    ev[2].record()
    torch.cuda.synchronize()
    f_ms, b_ms = ev[0].elapsed_time(ev[1]) / args.iters, ev[1].elapsed_time(ev[2]) / args.iters
    print(json.dumps(dict(semantics="classical", mode="until", traces=B, horizon=T, fwd_ms=round(f_ms, 4),
                          fwd_gbs=round(16 * B * T / f_ms / 1e6, 1), bwd_ms=round(b_ms, 4), bwd_gbs=round(16 * B * T / b_ms / 1e6, 1))), flush=True)


if __name__ == "__main__":
    main()
