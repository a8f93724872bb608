"""GPU tier: the FUSED kernels (run-time specialised, register resident) through the C ABI,
pinned with PYSTL_B200_TIER=fused so that a silent fall-back to the interpreter tier fails the
test, against the oracle; plus fused == interpreter on the BASELINE shapes."""

from __future__ import annotations

import os

import numpy as np
import pytest
import torch

from fused_cases import case_signal, forward_cases
from parity_util import assert_parity, oracle_trace
from pystl_b200 import create_semantics
from pystl_b200.formula import BASELINE_SPECS, from_spec

pytestmark = pytest.mark.gpu

FWD = forward_cases()


class tier:
    def __init__(self, name, threads=0):
        self.env = {"PYSTL_B200_TIER": name, "PYSTL_B200_FORCE_THREADS": str(threads) if threads else None}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.env}
        for k, v in self.env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def run(spec, sem, kw, x, which, threads=0):
    s = create_semantics(sem, backend="b200", **kw)
    with tier(which, threads):
        out = s.robustness_trace(from_spec(spec), x)
    if sem == "cumulative":
        out = torch.stack((out.pos, out.neg), dim=-1)
    return out


@pytest.mark.parametrize("case", FWD, ids=[c["name"] for c in FWD])
def test_forward(case):
    sig = case_signal(case)
    want = oracle_trace(case["spec"], case["sem"], case["kw"], sig)
    x = torch.tensor(sig, dtype=torch.float32, device="cuda")
    got = run(case["spec"], case["sem"], case["kw"], x, "fused", case["threads"]).cpu().numpy()
    assert_parity(got, want, case["sem"], case["name"], scale=case["scale"])


@pytest.mark.parametrize("cfg,sem,kw,shape", [
    ("cfg1", "classical", {}, (1024, 1000, 2)),
    ("cfg2", "smooth", {"temperature": 1.0}, (2048, 4096, 2)),
    ("cfg3", "agm", {}, (512, 8192, 2)),
    ("cfg3", "dgmsr", {"eps": 1e-8, "p": 1}, (512, 8192, 2)),
    ("cfg2", "cumulative", {}, (512, 4096, 2)),
], ids=["cfg1", "cfg2", "cfg3-agm", "cfg3-dgmsr", "cfg2-cumulative"])
def test_fused_equals_interpreter_at_baseline_shapes(cfg, sem, kw, shape):
    """Full trace lengths of the BASELINE configs: the two tiers are independent implementations
    of the same lowering, so agreement at sizes the oracle cannot reach pins both."""
    g = torch.Generator(device="cuda").manual_seed(4321)
    x = torch.randn(shape, generator=g, device="cuda")
    x[: shape[0] // 2] += 1.5  # both branches of agm / dgmsr
    a = run(BASELINE_SPECS[cfg], sem, kw, x, "fused").double()
    b = run(BASELINE_SPECS[cfg], sem, kw, x, "interp").double()
    assert a.shape == b.shape and bool(torch.isfinite(a).all())
    if sem == "classical":
        assert torch.equal(a, b)
    else:
        assert torch.allclose(a, b, rtol=2e-5, atol=1e-5), float((a - b).abs().max())
    # seeded sub-sample against the oracle
    idx = [0, shape[0] // 2, shape[0] - 1]
    want = oracle_trace(BASELINE_SPECS[cfg], sem, kw, x[idx].double().cpu().numpy())
    assert_parity(a[idx].float().cpu().numpy(), want, sem, f"{cfg}/{sem}/full", scale=4.0 if sem == "cumulative" else 1.0)


def test_mean_accumulator_matches():
    g = torch.Generator(device="cuda").manual_seed(99)
    x = torch.randn((64, 1000, 2), generator=g, device="cuda")
    sem = create_semantics("smooth", backend="b200", temperature=0.7)
    phi = from_spec(BASELINE_SPECS["cfg2"])
    with tier("fused"):
        m = sem.mean_robustness(phi, x)
        out = sem.robustness_trace(phi, x)
    assert abs(float(m) - float(out.double().mean())) < 1e-6
