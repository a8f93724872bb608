"""Parity corpus shared by the golden-vector generator and the tests.

Every case is small enough for the reference's per-``(trace, t)`` recursion:
formula spec (see ``pystl_b200.formula.from_spec``), semantics + kwargs, a
seeded signal and -- for gradient cases -- a seeded cotangent.
"""

from __future__ import annotations

import zlib

import numpy as np


def _gt(d, c):
    return ("pred", {d: 1.0}, -float(c))


def _lt(d, c):
    return ("pred", {d: -1.0}, float(c))


SPECS = {
    # BASELINE cfg1 verbatim
    "cfg1": (("and", [("always", (0, 10), _gt(0, 0.0)), ("eventually", (0, 5), _lt(1, 1.0))]), 40, 2),
    # BASELINE cfg2 verbatim (windows partly clipped at T=80)
    "cfg2": (("or", [("always", (0, 50), _gt(0, 0.0)), ("eventually", (0, 20), _lt(1, 0.5))]), 80, 2),
    # cfg3 / cfg4 / cfg5 with scaled intervals (the reference is O(prod W) per sample)
    "cfg3s": (("always", (0, 20), ("and", [("eventually", (0, 5), _gt(0, 0.0)), ("not", _gt(1, 2.0))])), 48, 2),
    "cfg4s": (("always", (0, 12), ("until", (0, 6), _gt(0, 0.0), _gt(1, 1.0))), 36, 4),
    "cfg5s": (
        (
            "or",
            [
                (
                    "and",
                    [
                        ("always", (0, 8), ("pred", {0: 1.0, 4: -1.0}, 0.0)),
                        ("eventually", (0, 5), ("pred", {1: -1.0, 5: -0.5}, 1.0)),
                    ],
                ),
                ("not", ("until", (0, 6), ("pred", {2: 1.0, 6: 1.0}, 0.5), ("pred", {3: 1.0, 7: -1.0}, -0.5))),
            ],
        ),
        32,
        8,
    ),
    # interval starts > 0: the tail of the trace raises in the reference (NaN here)
    "shifted": (("or", [("eventually", (2, 7), _gt(0, 0.0)), ("always", (1, 3), _lt(1, 1.0))]), 30, 2),
    # open-ended intervals
    "unbounded": (("and", [("always", (0, None), _gt(0, -0.5)), ("eventually", (3, None), _lt(1, 1.0))]), 24, 2),
    "until_open": (("until", (2, None), _gt(0, 0.0), _gt(1, 1.0)), 24, 2),
    "until_shift": (("until", (1, 9), _gt(0, 0.25), _lt(1, 1.5)), 28, 2),
    # negation above temporal operators, n-ary with a repeated leaf
    "negtemp": (("and", [("not", ("eventually", (0, 4), _gt(0, 0.0))), ("always", (0, 3), ("not", _gt(1, 1.0)))]), 24, 2),
    "nary3": (("or", [_gt(0, 0.0), _lt(1, 1.0), ("and", [_gt(0, 0.0), _gt(1, 0.5), _lt(0, 2.0)])]), 16, 2),
    # window of one sample, window longer than the trace
    "degenerate": (("and", [("always", (0, 0), _gt(0, 0.0)), ("eventually", (0, 100), _gt(1, 1.0))]), 12, 2),
    # weights (only AGM / D-GMSR use them; the others must ignore them)
    "weighted": (
        (
            "and",
            [
                ("always", (0, 3), _gt(0, 0.0), [1.0, 2.0, 0.5, 1.5]),
                ("or", [_lt(1, 1.0), _gt(1, 0.5)], [1.0, 3.0]),
                (
                    "until",
                    (0, 3),
                    _gt(0, 0.0),
                    _gt(1, 1.0),
                    [1.0, 2.0, 1.5, 1.0],
                    [1.0, 1.2, 0.8, 1.4],
                    (1.0, 1.3),
                ),
            ],
            [2.0, 1.0, 1.0],
        ),
        14,
        2,
    ),
}

SEMS = [
    ("classical", {}),
    ("smooth", {"temperature": 0.5}),
    ("smooth_sharp", {"temperature": 0.02}),
    ("cumulative", {}),
    ("agm", {}),
    ("dgmsr", {"eps": 1e-8, "p": 1}),
    ("dgmsr_p2", {"eps": 1e-8, "p": 2}),
]

# gradient golden vectors only where no t raises (cotangent spans all t)
GRAD_SPECS = {"cfg1", "cfg2", "cfg3s", "cfg4s", "cfg5s", "negtemp", "nary3", "degenerate", "weighted", "unbounded"}

B = 3


def _sem_name(tag: str) -> str:
    return tag.split("_")[0]


CASES = []
for _spec_name, (_spec, _T, _D) in SPECS.items():
    for _tag, _kw in SEMS:
        CASES.append(
            {
                "name": f"{_spec_name}/{_tag}",
                "spec_name": _spec_name,
                "spec": _spec,
                "T": _T,
                "D": _D,
                "sem": _sem_name(_tag),
                "kw": _kw,
                "grad": _spec_name in GRAD_SPECS,
            }
        )


def _seed(case) -> int:
    return zlib.crc32(case["spec_name"].encode()) & 0x7FFFFFFF


def case_signal(case) -> np.ndarray:
    """[B, T, D] float64 holding float32-representable values.
    mu_b + 0.5 N(0,1), mu ~ U(-1, 3) per trace and channel: exercises both AGM / D-GMSR branches."""
    rng = np.random.default_rng(_seed(case))
    mu = rng.uniform(-1.0, 3.0, size=(B, 1, case["D"]))
    x = mu + 0.5 * rng.standard_normal((B, case["T"], case["D"]))
    x[0] = rng.standard_normal((case["T"], case["D"]))  # one zero-mean trace: sign changes everywhere
    return x.astype(np.float32).astype(np.float64)


def case_cotangent(case, shape) -> np.ndarray:
    rng = np.random.default_rng(_seed(case) + 1)
    return rng.uniform(0.25, 1.0, size=shape).astype(np.float32).astype(np.float64)
