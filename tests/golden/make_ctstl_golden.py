"""Generate ``ctstl_vectors.npz`` by running the UNMODIFIED reference class ``pystl.CtstlSemantics``.

Run in the build container (where ``/root/reference`` is mounted):

    python tests/golden/make_ctstl_golden.py

Nothing from this repo is on that path.  Cases: seeded value traces with ties (quantised) and without; for every (a, b, tau,
delta) the reference's ``temporal_cumulative`` on the window ``[t + a, min(t + b, T - 1)]`` (api.py:49-54) and for every (a, b)
its ``temporal_until`` on the traces from t on (api.py:284-302); NaN where the reference raises ValueError.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("PYSTL_REFERENCE", "/root/reference"))

from pystl import CtstlSemantics  # noqa: E402  (the reference)

CUM_CASES = [  # (a, b, tau, delta); b = -1: no end
    (0, 0, 1.0, 1.0), (0, 4, 1.0, 1.0), (0, 4, 5.0, 1.0), (0, 4, 6.0, 1.0), (2, 9, 3.0, 1.0), (0, 16, 2.5, 0.5), (3, 3, 1.0, 1.0),
    (0, 40, 7.0, 1.0), (5, 30, 0.3, 0.1), (0, -1, 4.0, 1.0), (6, -1, 1.0, 1.0), (0, 63, 64.0, 1.0),
]
UNTIL_CASES = [(0, 0), (0, 1), (0, 5), (1, 5), (3, 3), (2, 12), (0, 40), (0, -1), (4, -1), (0, 200)]


def traces(seed: int, B: int, T: int, quant: float):
    rng = np.random.default_rng(seed)
    y = rng.uniform(-1, 1, size=(B, 1)) + rng.standard_normal((B, T))
    if quant:
        y = np.round(y / quant) * quant
    return y.astype(np.float32).astype(np.float64)


def main() -> None:
    sem = CtstlSemantics(delta=1.0)
    out = {}
    y = np.concatenate([traces(1, 3, 37, 0.0), traces(2, 3, 37, 0.5)])
    l = np.concatenate([traces(3, 3, 37, 0.0), traces(4, 3, 37, 0.5)])
    out["y"], out["l"] = y, l
    B, T = y.shape
    cum = np.zeros((len(CUM_CASES), B, T))
    for c, (a, b, tau, delta) in enumerate(CUM_CASES):
        for i in range(B):
            for t in range(T):
                hi = T - 1 if b < 0 else min(t + b, T - 1)
                cum[c, i, t] = sem.temporal_cumulative(list(y[i, t + a : hi + 1]), tau=tau, delta=delta)
    unt = np.full((len(UNTIL_CASES), B, T), np.nan)
    for c, (a, b) in enumerate(UNTIL_CASES):
        for i in range(B):
            for t in range(T):
                try:
                    unt[c, i, t] = sem.temporal_until(left_trace=list(l[i, t:]), right_trace=list(y[i, t:]), start=a, end=None if b < 0 else b)
                except ValueError:
                    pass
    out["cum_cases"] = np.asarray(CUM_CASES, dtype=np.float64)
    out["until_cases"] = np.asarray(UNTIL_CASES, dtype=np.int64)
    out["cum"], out["until"] = cum, unt
    np.savez_compressed(os.path.join(HERE, "ctstl_vectors.npz"), **out)
    print("ctstl_vectors.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
