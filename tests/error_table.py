#!/usr/bin/env python
"""Measured error table of the fused kernels against the float64 oracle (TEST INFRASTRUCTURE).

    python tests/error_table.py --backend gpu  [--out profiles/round2_error_table.md]
    python tests/error_table.py --backend emu  (host emulator: libm transcendentals, no GPU)

For every semantics x BASELINE formula (cfg1 .. cfg5) it evaluates the robustness trace and the
gradient of ``sum(g * out)`` on the parity signals of the test corpus (``mixed_signal``: both
branches of AGM / D-GMSR occur) and prints the worst

    value error     / (5e-6 + 1e-5 |ref|)        (the bar of DESIGN.md section 2)
    gradient error  / (2e-5 + 2e-5 |ref|)

over all entries, for two tilings (library's choice, and 64-thread tiles so that every trace
spans several tiles and ends inside one).  Classic rows report the number of entries that are
not bit-exact instead.  The table committed under profiles/ is the output of the ``gpu`` backend
on a B200; ``tests/test_gpu_*`` assert the same bars.
"""

from __future__ import annotations

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), HERE]

from parity_util import ATOL, RTOL, SEM_KW, mixed_signal, oracle_trace  # noqa: E402
from pystl_b200.formula import BASELINE_SPECS, from_spec  # noqa: E402
from test_emu_backward import G_ATOL, G_RTOL, oracle_grad  # noqa: E402

SHAPES = {"cfg1": (3, 700, 2), "cfg2": (3, 700, 2), "cfg3": (3, 700, 2), "cfg4": (3, 700, 4), "cfg5": (3, 700, 8)}
EXTRA = {"cfg5/T512": ("cfg5", (2, 512, 8)), "cfg5/T2048": ("cfg5", (2, 2048, 8)), "cfg4/T1024": ("cfg4", (2, 1024, 4))}


def gpu_eval(spec, sem, kw, sig64, g64, threads):
    import torch

    from pystl_b200 import capi, create_semantics

    s = create_semantics(sem, backend="b200", **kw)
    x = torch.tensor(sig64, dtype=torch.float32, device="cuda")
    B, T, D = x.shape
    cf = s.compile(from_spec(spec), D)
    C = s.out_channels
    out = torch.empty((B, T, C), dtype=torch.float32, device="cuda")
    grad = torch.full_like(x, -12345.0)
    go = torch.tensor(g64, dtype=torch.float32, device="cuda").contiguous()
    os.environ["PYSTL_B200_TIER"] = "fused"
    if threads:
        os.environ["PYSTL_B200_FORCE_THREADS"] = str(threads)
    else:
        os.environ.pop("PYSTL_B200_FORCE_THREADS", None)
    rc = capi.lib().pystl_b200_eval_value_and_grad(cf.ptr(), x.data_ptr(), B, T, D, go.data_ptr(), 1.0, out.data_ptr(), None,
                                                   grad.data_ptr(), torch.cuda.current_stream().cuda_stream)
    capi.raise_for(rc, capi.lib().pystl_b200_last_error())
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    return (o if C == 2 else o[..., 0]), grad.cpu().numpy().astype(np.float64)


def emu_eval(spec, sem, kw, sig64, g64, threads):
    import fused_emu_lib

    out = fused_emu_lib.eval_forward(from_spec(spec), sem, sig64, force_threads=threads, **kw)
    grad = fused_emu_lib.eval_backward(from_spec(spec), sem, sig64, g64, force_threads=threads, **kw)
    return out, grad.astype(np.float64)


def one(backend, name, cfg, shape, sem, kw, threads, seed=7):
    rng = np.random.default_rng(seed)
    sig = mixed_signal(rng, *shape)
    B, T, _ = shape
    gshape = (B, T, 2) if sem == "cumulative" else (B, T)
    g = rng.standard_normal(gshape).astype(np.float32).astype(np.float64)
    spec = BASELINE_SPECS[cfg]
    want = oracle_trace(spec, sem, kw, sig)
    wantg = oracle_grad(spec, sem, kw, sig, g)
    try:
        got, gotg = (gpu_eval if backend == "gpu" else emu_eval)(spec, sem, kw, sig, g, threads)
    except NotImplementedError as exc:
        return dict(name=name, sem=sem + str(kw.get("p", "")), threads=threads, error=str(exc)[:60])
    nanpat = bool(np.array_equal(np.isnan(got), np.isnan(want)))
    live = ~np.isnan(want)
    if sem == "classical":
        v = float(((got != want.astype(np.float32)) & live).sum())
    else:
        err = np.abs(got.astype(np.float64) - want)
        v = float(np.nanmax(np.where(live, err / (ATOL + RTOL * np.abs(want)), 0.0)))
    gerr = np.abs(gotg - wantg)
    gq = float((gerr / (G_ATOL + G_RTOL * np.abs(wantg))).max())
    route = bool(np.array_equal(gotg != 0, wantg != 0)) if sem == "classical" else True
    return dict(name=name, sem=sem + str(kw.get("p", "")), threads=threads, value=v, grad=gq, nan_ok=nanpat, route_ok=route,
                gmax=float(np.abs(wantg).max()))


def _pc(job):
    cfg, sem, kw, T, D, threads = job
    from pystl_b200 import capi
    from pystl_b200.compiler import compile_formula

    if threads:
        os.environ["PYSTL_B200_FORCE_THREADS"] = str(threads)
    else:
        os.environ.pop("PYSTL_B200_FORCE_THREADS", None)
    cf = compile_formula(from_spec(BASELINE_SPECS[cfg]), sem, D, **kw)
    rc = capi.lib().pystl_b200_fused_precompile(cf.ptr(), T, D, 1)
    return rc, capi.lib().pystl_b200_last_error().decode() if rc else ""


def precompile(work) -> int:
    import multiprocessing as mp

    jobs = []
    for name, cfg, shape in work:
        for sem, kw in SEM_KW:
            for threads in (0, 64):
                if name == "cfg5/T2048" and threads:
                    threads = 256
                jobs.append((cfg, sem, kw, shape[1], shape[2], threads))
    with mp.get_context("spawn").Pool(os.cpu_count() or 4) as pool:
        res = pool.map(_pc, jobs)
    bad = [(j, m) for j, (rc, m) in zip(jobs, res) if rc]
    print(f"{len(jobs) - len(bad)} kernels cached, {len(bad)} not", bad[:3])
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gpu", choices=["gpu", "emu"])
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--precompile", action="store_true", help="only fill the cubin cache (no GPU needed)")
    args = ap.parse_args()
    rows = []
    work = [(c, c, SHAPES[c]) for c in SHAPES] + ([] if args.quick else [(n, c, s) for n, (c, s) in EXTRA.items()])
    if args.precompile:
        return precompile(work)
    for name, cfg, shape in work:
        for sem, kw in SEM_KW:
            if args.only and args.only not in f"{name}/{sem}":
                continue
            for threads in (0, 64):
                if name in EXTRA and threads == 0 and shape[1] <= 512:
                    continue
                if name == "cfg5/T2048":
                    threads = 0 if threads == 0 else 256  # exactly one 256-thread tile
                rows.append(one(args.backend, name, cfg, shape, sem, kw, threads))
                r = rows[-1]
                print(r, flush=True)
    lines = [
        f"# Worst error / tolerance of the fused kernels vs the float64 oracle ({args.backend})",
        "",
        f"Values: |err| / ({ATOL:g} + {RTOL:g} |ref|); gradients: |err| / ({G_ATOL:g} + {G_RTOL:g} |ref|); cotangent N(0,1).",
        "Classic: number of entries that are not bit-exact (values) -- its gradient column is the same ratio, routing must match.",
        "",
        "| formula | semantics | threads | value | gradient | NaN pattern | routing |",
        "|---|---|---|---|---|---|---|",
    ]
    for r in rows:
        if "error" in r:
            lines.append(f"| {r['name']} | {r['sem']} | {r['threads']} | n/a: {r['error']} | | | |")
        else:
            lines.append(f"| {r['name']} | {r['sem']} | {r['threads']} | {r['value']:.3g} | {r['grad']:.3g} | {'ok' if r['nan_ok'] else 'DIFFERS'} | "
                         f"{'ok' if r['route_ok'] else 'DIFFERS'} |")
    text = "\n".join(lines) + "\n"
    if args.out:
        with open(args.out, "w") as fh:
            fh.write(text)
    print(text)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
