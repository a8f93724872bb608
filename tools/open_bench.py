"""Open-ended formulas through the PUBLIC API on long traces (T beyond one tile): CUDA-event time per call.

    python tools/open_bench.py [--traces 16384 --horizon 16384 --channels 2]

Each formula runs `sem.robustness_trace` (values) and `sem.mean_robustness_and_grad` (mean + gradient in one call); the
figure next to each is the traffic of an ideal single pass (signal read once, robustness written once; + gradient
written once), i.e. what a fully fused kernel would move.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pystl_b200 import create_semantics  # noqa: E402
from pystl_b200.formula import Always, And, Eventually, Interval, Or, Until, affine_predicate  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--traces", type=int, default=16384)
    ap.add_argument("--horizon", type=int, default=16384)
    ap.add_argument("--channels", type=int, default=2)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", type=int, default=-1, help="run one case (for a profiler)")
    ap.add_argument("--warm", type=int, default=2)
    args = ap.parse_args()
    B, T, D = args.traces, args.horizon, args.channels
    x = affine_predicate("x>0", {0: 1.0}, 0.0)
    y = affine_predicate("y<0.5", {1: -1.0}, 0.5)
    cases = [
        ("classical", {}, "G[0,inf)(x>0)", Always(x, Interval(0, None))),
        ("classical", {}, "(x>0) U[0,inf) (y<0.5)", Until(x, y, Interval(0, None))),
        ("smooth", {"temperature": 1.0}, "G[0,inf)(x>0) | F[5,inf)(y<0.5)", Or(Always(x, Interval(0, None)), Eventually(y, Interval(5, None)))),
        ("smooth", {"temperature": 1.0}, "G[0,inf)(F[0,20](y<0.5))", Always(Eventually(y, Interval(0, 20)), Interval(0, None))),
        ("agm", {}, "G[0,inf)(x>0) & F[0,10](y<0.5)", And(Always(x, Interval(0, None)), Eventually(y, Interval(0, 10)))),
        ("dgmsr", {"eps": 1e-8, "p": 1}, "F[3,inf)(x>0 & y<0.5)", Eventually(And(x, y), Interval(3, None))),
    ]
    sig = torch.randn(B, T, D, device="cuda") * 0.5
    for idx, (sem_name, kw, text, phi) in enumerate(cases):
        if args.only >= 0 and idx != args.only:
            continue
        sem = create_semantics(sem_name, backend="b200", **kw)
        for _ in range(args.warm):
            sem.robustness_trace(phi, sig)
            sem.mean_robustness_and_grad(phi, sig)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        for _ in range(args.iters):
            sem.robustness_trace(phi, sig)
        ev[1].record()
        for _ in range(args.iters):
            mean, grad = sem.mean_robustness_and_grad(phi, sig)
        ev[2].record()
        torch.cuda.synchronize()
        f_ms, g_ms = ev[0].elapsed_time(ev[1]) / args.iters, ev[1].elapsed_time(ev[2]) / args.iters
        n = B * T
        print(json.dumps(dict(semantics=sem_name, formula=text, traces=B, horizon=T, channels=D,
                              values_ms=round(f_ms, 3), values_samples_per_s=round(n / f_ms * 1e3, 0),
                              values_ideal_gbs=round(4 * (D + 1) * n / f_ms / 1e6, 1),
                              mean_grad_ms=round(g_ms, 3), mean_grad_samples_per_s=round(n / g_ms * 1e3, 0),
                              mean_grad_ideal_gbs=round(4 * (2 * D + 1) * n / g_ms / 1e6, 1),
                              mean=float(mean), grad_finite=bool(torch.isfinite(grad).all()),
                              peak_mem_gb=round(torch.cuda.max_memory_allocated() / 1e9, 2))), flush=True)
        del mean, grad
        torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
