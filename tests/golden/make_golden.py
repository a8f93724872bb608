"""Generate ``reference_vectors.npz`` by running the UNMODIFIED reference.

Run in the build container (where ``/root/reference`` is mounted):

    PYTHONPATH=/root/repo python tests/golden/make_golden.py

For every case in ``tests/golden_cases.py`` it evaluates
``phi.evaluate(signal[b], create_semantics(name, backend="numpy"), t)`` for every
``(b, t)`` (NaN where the reference raises ValueError) and, for the gradient
cases, back-propagates ``sum_t g[t] * phi.evaluate(signal_torch[b], sem_torch, t)``
through the reference's torch backend.  Nothing from this repo's product or
oracle code is on that path; ``pystl_b200.formula.from_spec`` is only used to
instantiate the reference's own node classes from a class-agnostic spec.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.environ.get("PYSTL_REFERENCE", "/root/reference"))

import torch  # noqa: E402

import pystl  # noqa: E402  (the reference)
from golden_cases import CASES, case_signal, case_cotangent  # noqa: E402
from pystl_b200.formula import from_spec  # noqa: E402


def reference_values(phi, sig: np.ndarray, sem_name: str, kw: dict) -> np.ndarray:
    sem = pystl.create_semantics(sem_name, backend="numpy", **kw)
    B, T, _ = sig.shape
    cum = sem_name == "cumulative"
    out = np.full((B, T, 2) if cum else (B, T), np.nan)
    for b in range(B):
        for t in range(T):
            try:
                v = phi.evaluate(sig[b], sem, t)
            except ValueError:
                continue
            out[b, t] = (v.pos, v.neg) if cum else float(v)
    return out


def reference_grad(phi, sig: np.ndarray, g: np.ndarray, vals: np.ndarray, sem_name: str, kw: dict) -> np.ndarray:
    sem = pystl.create_semantics(sem_name, backend="torch", **kw)
    B, T, _ = sig.shape
    grads = np.zeros_like(sig)
    for b in range(B):
        x = torch.tensor(sig[b], dtype=torch.float64, requires_grad=True)
        loss = torch.zeros((), dtype=torch.float64)
        for t in range(T):
            if np.isnan(vals[b, t]).any():
                continue  # the reference raises here: no contribution
            v = phi.evaluate(x, sem, t)
            if sem_name == "cumulative":
                loss = loss + g[b, t, 0] * v.pos + g[b, t, 1] * v.neg
            else:
                loss = loss + g[b, t] * v
        loss.backward()
        grads[b] = x.grad.numpy()
    return grads


def main() -> None:
    out = {}
    for case in CASES:
        name = case["name"]
        sig = case_signal(case)
        phi = from_spec(case["spec"], pystl)
        vals = reference_values(phi, sig, case["sem"], case["kw"])
        out[f"{name}/values"] = vals
        if case.get("grad"):
            g = case_cotangent(case, vals.shape)
            out[f"{name}/grad"] = reference_grad(phi, sig, g, vals, case["sem"], case["kw"])
        print(name, "ok", vals.shape, "nan:", int(np.isnan(vals).sum()))
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)


if __name__ == "__main__":
    main()
