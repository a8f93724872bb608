#!/usr/bin/env python
"""Fill the fused tier's cubin cache (pystl_b200/csrc/build/fused) ahead of time.

NVRTC needs no GPU, so the kernels of the BASELINE workloads (bench.py) and of the parity
corpus (tests/fused_cases.py) are compiled here, in parallel, and travel with the tree; on the
GPU box the library then only loads cubins.  A formula that is not in the cache is compiled
on first use (a second or two), so this is an optimisation, not a requirement.

    python tools/precompile.py [--tests] [--jobs 8]
"""

from __future__ import annotations

import argparse
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def _open(job):
    """The open-ended corpus (tests/test_gpu_open_window.py): formulas the host splits at open-ended operators."""
    _, which, name, sem, kw, T, force = job
    import test_gpu_open_window as ow
    from pystl_b200 import create_semantics

    os.environ.pop("PYSTL_B200_FORCE_THREADS", None)
    if which == "bench":
        import bench

        sem, kw, _, phi = bench.open_ended_cases()[name]
        T = bench.OPEN_ENDED_SHAPE[1]
    elif which == "random":
        phi = ow.random_open_formula(name, sem)
    else:
        phi = {"window": ow.formulas, "until": ow.until_formulas, "cumulative": ow.cumulative_formulas}[which]()[name]
    s = create_semantics(sem, backend="b200", **kw)
    s.force_open_leaf = force
    try:
        return (0 if s.precompile(phi, T, 2) else -2), ""
    except NotImplementedError as exc:
        return -2, str(exc)


def _one(job):
    if job[0] == "open":
        return _open(job)
    spec, sem, kw, T, D, threads, backward = job
    from pystl_b200 import capi
    from pystl_b200.compiler import compile_formula
    from pystl_b200.formula import from_spec

    if threads:
        os.environ["PYSTL_B200_FORCE_THREADS"] = str(threads)
    else:
        os.environ.pop("PYSTL_B200_FORCE_THREADS", None)
    lib = capi.lib()
    cf = compile_formula(from_spec(spec), sem, D, **kw)
    rc = lib.pystl_b200_fused_precompile(cf.ptr(), T, D, int(backward))
    return rc, (lib.pystl_b200_last_error().decode() if rc else "")


def jobs(with_tests: bool):
    from pystl_b200.formula import BASELINE_SHAPES, BASELINE_SPECS

    sys.path.insert(0, ROOT)
    import bench

    out = []
    for cfg, (sem, kw, bwd) in bench.WORKLOADS.items():
        _, T, D = BASELINE_SHAPES[cfg[:4]]
        out.append((BASELINE_SPECS[cfg[:4]], sem, kw, T, D, 0, False))
        if bwd:
            out.append((BASELINE_SPECS[cfg[:4]], sem, kw, T, D, 0, True))
    for i in range(len(bench.open_ended_cases())):  # the open-ended lines of the bench
        out.append(("open", "bench", i, "", {}, 0, False))
    if with_tests:
        import fused_cases

        for c in fused_cases.forward_cases():
            out.append((c["spec"], c["sem"], c["kw"], c["T"], c["D"], c["threads"], False))
        for c in fused_cases.backward_cases() + fused_cases.cumulative_backward_cases() + fused_cases.smooth_until_backward_cases() + fused_cases.gm_backward_cases() + fused_cases.random_until_backward_cases() + fused_cases.nan_flow_backward_cases():
            out.append((c["spec"], c["sem"], c["kw"], c["T"], c["D"], c["threads"], True))
        import test_gpu_open_window as ow

        for sem, kw in ow.SEMS:
            for name in ow.formulas():
                out.append(("open", "window", name, sem, kw, 300, True))
                out.append(("open", "window", name, sem, kw, 2048, True))
                out.append(("open", "window", name, sem, kw, 2048, False))
            for seed in range(8):
                if kw in ({}, {"temperature": 0.5}, {"eps": 1e-8, "p": 1}):
                    out.append(("open", "random", seed, sem, kw, 1000, True))
        for seed in range(8):
            out.append(("open", "random", seed, "cumulative", {}, 1000, True))
        for name in ow.cumulative_formulas():
            out.append(("open", "cumulative", name, "cumulative", {}, 300, True))
        for name in ow.until_formulas():
            for T, force in ((300, True), (2048, True), (2048, False)):
                out.append(("open", "until", name, "classical", {}, T, force))
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--tests", action="store_true", help="also the kernels of the parity corpus")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    args = ap.parse_args(argv)
    todo = jobs(args.tests)
    t0 = time.time()
    with mp.get_context("spawn").Pool(args.jobs) as pool:
        res = pool.map(_one, todo, chunksize=4)
    ok = sum(1 for rc, _ in res if rc == 0)
    unsup = sum(1 for rc, _ in res if rc == -2)
    bad = [(j, m) for j, (rc, m) in zip(todo, res) if rc not in (0, -2)]
    print(f"precompile: {ok} kernels cached, {unsup} on the interpreter tier, {len(bad)} failed, {time.time() - t0:.1f}s")
    for j, m in bad[:5]:
        print("  FAILED", j[1], j[0] if j[0] != "open" else j[2:], m[:300])
    return 1 if bad else 0


if __name__ == "__main__":
    raise SystemExit(main())
