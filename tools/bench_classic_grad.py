#!/usr/bin/env python
"""Time the Classic value-and-gradient path (cfg1 / cfg5 formulas) on one GPU.  Not part of bench.py:
the BASELINE configs ask for the Classic forward only; this reports what the gradient costs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pystl_b200
from pystl_b200.formula import BASELINE_SPECS, BASELINE_SHAPES, from_spec

for cfg, B in (("cfg1", 1 << 18), ("cfg5", 1 << 15)):
    _, T, D = BASELINE_SHAPES[cfg]
    phi = from_spec(BASELINE_SPECS[cfg])
    sem = pystl_b200.create_semantics("classical", backend="b200")
    x = torch.randn(B, T, D, device="cuda")
    for _ in range(3):
        m, g = sem.mean_robustness_and_grad(phi, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        m, g = sem.mean_robustness_and_grad(phi, x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    bytes_ = (4 * (2 if cfg == "cfg1" else 8) * 2 + 4 * D) * B * T  # read signal twice (fwd, bwd) + write grad
    print(f"{cfg} classical value+grad B={B}: {ms:.3f} ms, {B*T/ms*1e3:.3e} samples/s, {bytes_/ms/1e6:.0f} GB/s")
