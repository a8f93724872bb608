#!/usr/bin/env python
"""What the fused mean (one double atomic per CTA) and the accumulator memset cost per step (cfg2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pystl_b200
from pystl_b200 import capi
from pystl_b200.formula import BASELINE_SPECS, BASELINE_SHAPES, from_spec

B, T, D = BASELINE_SHAPES["cfg2"]
phi = from_spec(BASELINE_SPECS["cfg2"])
sem = pystl_b200.create_semantics("smooth", backend="b200", temperature=1.0)
x = torch.randn(B, T, D, device="cuda")
cf = sem.compile(phi, D)
out = torch.empty(B, T, 1, device="cuda"); grad = torch.empty_like(x); acc = torch.zeros(2, dtype=torch.float64, device="cuda")
lib = capi.lib(); st = torch.cuda.current_stream().cuda_stream
def run(use_acc, zero):
    if zero: acc.zero_()
    rc = lib.pystl_b200_eval_value_and_grad(cf.ptr(), x.data_ptr(), B, T, D, None, 1e-9, out.data_ptr(), acc.data_ptr() if use_acc else None, grad.data_ptr(), st)
    assert rc == 0
for name, ua, z in (("no mean", False, False), ("mean", True, False), ("memset + mean", True, True)):
    for _ in range(5): run(ua, z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): run(ua, z)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:14s} {e0.elapsed_time(e1)/100:.4f} ms/step")
