"""CPU tier: the FUSED kernels (generated source + fused_rt.cuh) compiled with g++ against the
fiber-based SIMT stand-in (tests/emu/simt_host.h), against the oracle."""

from __future__ import annotations

import pytest

import fused_emu_lib
from fused_cases import case_signal, forward_cases
from parity_util import assert_parity, oracle_trace
from pystl_b200.formula import from_spec

FWD = forward_cases(light=True)


@pytest.mark.parametrize("case", FWD, ids=[c["name"] for c in FWD])
def test_forward(case):
    sig = case_signal(case)
    want = oracle_trace(case["spec"], case["sem"], case["kw"], sig)
    got = fused_emu_lib.eval_forward(from_spec(case["spec"]), case["sem"], sig, force_threads=case["threads"], **case["kw"])
    assert_parity(got, want, case["sem"], case["name"], scale=case["scale"])
