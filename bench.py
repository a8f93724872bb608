#!/usr/bin/env python
"""Benchmark of the robustness-evaluation hot path (BASELINE.json metric: trace-samples/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl b200|reference]

A *step* is one pass of the hot path over one batch of synthetic traces: the fused
formula kernel over ``signal[B, T, D]`` (plus the gradient kernel for the forward+backward
workloads).  ``value`` is whole-job trace-samples/sec with the inputs resident in HBM;
``e2e`` is the same metric through the public API with pinned HOST buffers (H2D copy of
the batch and D2H read of the mean robustness inside the timed region).

Under torchrun (N > 1) every rank evaluates its own shard of the batch (weak scaling: the
per-GPU batch is fixed) and the mean robustness of every step is finished by one 16-byte NCCL
all-reduce issued asynchronously (it overlaps the next step's kernel; all of them are waited for
inside the timed region); the time is the max over ranks of the device-timed region.

Besides the headline workload the line carries, driver-measured in the same run,
``other_kernels`` -- kernel time and roofline fraction of every BASELINE.json configuration
(cfg1 .. cfg5, all five semantics for cfg5, forward and value-and-gradient) at its named size --
and ``cfg5_sharded``: BASELINE config 5 as named, 2^20 traces x 2048 x 8 split over the N ranks
(strong scaling) with the NCCL mean-robustness reduction.  At N = 1 two more blocks time what is on the
path but in no BASELINE configuration, through the public API: ``open_ended`` (operators without an
interval end on traces longer than one tile, csrc/scan.cu) and ``ctstl`` (CT-STL's selection kernel
and exclusive-prefix Until, csrc/select.cu).

``--impl reference`` times the CPU implementation of the same path on all host cores on a bounded
sample of the same workload: for forward workloads the oracle port of the reference's
per-(trace, t) recursion (oracle/scalar.py); for forward+backward workloads the batched float64
restatement of the reference's torch backend under autograd (oracle/dense.py) -- the Python
reference itself cannot travel to the GPU box.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line

# stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner) write to
# file descriptor 1 directly, so fd 1 is pointed at stderr for the whole run and the JSON line goes
# to a private duplicate of the original stdout.
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())

import numpy as np  # noqa: E402
import torch  # noqa: E402

from pystl_b200.formula import BASELINE_SHAPES, BASELINE_SPECS, from_spec  # noqa: E402

METRIC = "trace-samples/sec"
UNIT = "trace-samples/s"

# semantics, kwargs, does the workload include the backward pass
WORKLOADS = {
    "cfg1": ("classical", {}, False),
    "cfg1x": ("classical", {}, False),  # cfg1's formula on 2^20 traces: the roofline figure (cfg1 itself is 8 MB, L2 resident)
    "cfg1xb": ("classical", {}, True),  # ... with the one-hot gradient (value-and-gradient kernel)
    "cfg2": ("smooth", {"temperature": 1.0}, True),
    "cfg3": ("agm", {}, True),
    "cfg3d": ("dgmsr", {"eps": 1e-8, "p": 1}, True),
    "cfg4": ("cumulative", {}, False),
    "cfg4b": ("cumulative", {}, True),  # ... with the gradient of mean(pos + neg)
    "cfg5": ("classical", {}, False),
    # cfg5's formula under the other four semantics ("all five semantics")
    "cfg5s": ("smooth", {"temperature": 1.0}, False),
    "cfg5a": ("agm", {}, False),
    "cfg5d": ("dgmsr", {"eps": 1e-8, "p": 1}, False),
    "cfg5c": ("cumulative", {}, False),
    # ... and with the gradient (value-and-gradient kernels; the Until adjoints are O(b) scatters)
    "cfg5b": ("classical", {}, True),
    "cfg5sb": ("smooth", {"temperature": 1.0}, True),
    "cfg5ab": ("agm", {}, True),
    "cfg5db": ("dgmsr", {"eps": 1e-8, "p": 1}, True),
    "cfg5cb": ("cumulative", {}, True),
}
DESCR = {
    "cfg1": "Classic G[0,10](x>0) & F[0,5](y<1), 1024x1000x2, forward",
    "cfg1x": "Classic G[0,10](x>0) & F[0,5](y<1), scaled to 1048576x1000x2, forward",
    "cfg1xb": "Classic G[0,10](x>0) & F[0,5](y<1), scaled to 524288x1000x2, forward+backward",
    "cfg2": "Smooth tau=1 G[0,50](x>0) | F[0,20](y<0.5), 65536x4096x2, forward+backward",
    "cfg3": "AGM G[0,100](F[0,10](x>0) & !(y>2)), 32768x8192x2, forward+backward",
    "cfg3d": "D-GMSR G[0,100](F[0,10](x>0) & !(y>2)), 32768x8192x2, forward+backward",
    "cfg4": "Cumulative G[0,200]((x>0) U[0,64] (y>1)), 16384x16384x4, forward",
    "cfg4b": "Cumulative G[0,200]((x>0) U[0,64] (y>1)), 16384x16384x4, forward+backward",
    "cfg5": "Classic 10-node mixed formula, 8 channels, 2^20 x 2048 x 8 (sharded), forward",
    "cfg5s": "Smooth tau=1 10-node mixed formula, 8 channels, 2^20 x 2048 x 8 (sharded), forward",
    "cfg5a": "AGM 10-node mixed formula, 8 channels, 2^20 x 2048 x 8 (sharded), forward",
    "cfg5d": "D-GMSR 10-node mixed formula, 8 channels, 2^20 x 2048 x 8 (sharded), forward",
    "cfg5c": "Cumulative 10-node mixed formula, 8 channels, 2^20 x 2048 x 8 (sharded), forward",
    "cfg5b": "Classic 10-node mixed formula, 8 channels, 2^20 x 2048 x 8 (sharded), forward+backward",
    "cfg5sb": "Smooth tau=1 10-node mixed formula, 8 channels, 2^20 x 2048 x 8 (sharded), forward+backward",
    "cfg5ab": "AGM 10-node mixed formula, 8 channels, 2^20 x 2048 x 8 (sharded), forward+backward",
    "cfg5db": "D-GMSR 10-node mixed formula, 8 channels, 2^20 x 2048 x 8 (sharded), forward+backward",
    "cfg5cb": "Cumulative 10-node mixed formula, 8 channels, 2^20 x 2048 x 8 (sharded), forward+backward",
}


def algorithmic_bytes_per_sample(cfg: str, backward: bool) -> float:
    """forward 4*D_ref + 4*C_out; backward 4*C_out(grad_out, 0 when the
    constant cotangent is fused) + 4*D_ref (re-read) + 4*D (grad write)."""
    _, _, D = BASELINE_SHAPES[cfg[:4]]
    d_ref = {"cfg1": 2, "cfg1x": 2, "cfg2": 2, "cfg3": 2, "cfg3d": 2, "cfg4": 2, "cfg5": 8}["cfg1x" if cfg.startswith("cfg1x") else ("cfg3d" if cfg == "cfg3d" else cfg[:4])]
    c_out = 2 if WORKLOADS[cfg][0] == "cumulative" else 1
    fwd = 4 * d_ref + 4 * c_out
    if not backward:
        return fwd
    return fwd + 4 * d_ref + 4 * D  # constant cotangent fused: no grad_out read


# Secondary ceiling ("when it binds"): instructions per trace-sample on the pipe that binds the
# O(b) Until walks of the non-lattice semantics, derived in DESIGN.md section 4 ("Pipe floors").  B200 SM: FMA pipe 128,
# ALU pipe 64 (FMNMX, compares, selects), XU pipe 16 (MUFU: ex2 / lg2 / rcp / sqrt) lanes per clock.
PIPE_LANES = {"xu": 16, "alu": 64}
try:  # the clock the floors are computed at: the measured maximum SM clock of this pool's B200s
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as _fh:
        SM_MAX_MHZ = float(json.load(_fh).get("sm_max_mhz", 1965.0))
except (OSError, ValueError):
    SM_MAX_MHZ = 1965.0
PIPE_OPS = {
    # cumulative Until [0, 64]: per candidate and channel one FMNMX (+ one FADD on the FMA pipe); G[0, 200] adds ~6
    ("cfg4", "forward"): ("alu", 65 * 2 + 2 * 2 + 6),
    # cfg5's Until [0, 24]: 25 candidates per sample and channel
    ("cfg5c", "forward"): ("alu", 25 * 2 + 2 * 2 + 12),
    # MUFU per sample as ncu counts them (smsp__inst_executed_pipe_xu x 32 / samples, profiles/round2/results.md):
    ("cfg5s", "forward"): ("xu", 32.4),   # one rcp per TWO candidates (rcp2), lifts shared by 8 outputs, window lifts / finishes
    ("cfg5a", "forward"): ("xu", 76.6),   # ex2 + lg2 per candidate, 1 / cnt from a table
    ("cfg5d", "forward"): ("xu", 205.7),  # 3 (flagged groups) to 8 (general groups) per candidate
    ("cfg5sb", "value_and_grad"): ("xu", 56.4),
    ("cfg5ab", "value_and_grad"): ("xu", 145.9),
    ("cfg5db", "value_and_grad"): ("xu", 514.7),
}


def pipe_floor(cfg: str, kind: str, B: int, T: int, ms_: float, sm_mhz: float) -> dict | None:
    """{"pipe", "ops_per_sample", "floor_ms", "frac"}: time the binding instruction pipe needs at 100 % utilisation
    over the measured kernel time, or None where HBM is the only ceiling stated."""
    ent = PIPE_OPS.get((cfg, kind))
    if ent is None:
        return None
    pipe, ops = ent
    floor_ms = ops * B * T / (148.0 * PIPE_LANES[pipe] * sm_mhz * 1e6) * 1e3
    return {"pipe": pipe, "ops_per_sample": ops, "lanes_per_clk_per_sm": PIPE_LANES[pipe], "sm_mhz": sm_mhz, "floor_ms": floor_ms,
            "frac": floor_ms / ms_}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: NVML in-process (one sample
    every ~2 ms, so that even a 100 ms region yields a median under load), nvidia-smi as fallback."""

    QUERY = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
        "clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self.source = "nvml"
        self._halt = threading.Event()
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[index]) if index < len(ids) else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)))
            self._nvml = pynvml
        except Exception:
            self.source = "nvidia-smi"

    def _sample_nvml(self):
        nv = self._nvml
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(
            ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.index)],
            capture_output=True, text=True, timeout=5,
        ).stdout.strip()
        if not out:
            return
        r = [c.strip() for c in out.split(",")]
        if r[0].replace(".", "").isdigit():
            self.sm.append(float(r[0]))
        if len(r) > 1 and r[1].replace(".", "").isdigit():
            self.mx.append(float(r[1]))
        for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
            if v.lower().startswith("active"):
                self.reasons.add(n)

    def run(self):
        while not self._halt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._halt.wait(0.002 if self._nvml is not None else 0.02)

    def stop(self) -> dict:
        self._halt.set()
        self.join(timeout=6)
        return {
            "sm_mhz": float(np.median(self.sm)) if self.sm else None,
            "sm_max_mhz": max(self.mx) if self.mx else None,
            "reasons": sorted(self.reasons),
            "samples": len(self.sm),
            "source": self.source,
        }


def tier_of(lib, cf, T: int, D: int, backward: bool) -> str:
    """Which tier evaluates this formula: the run-time specialised fused kernels or the interpreter."""
    if os.environ.get("PYSTL_B200_TIER") == "interp":
        return "interpreter"
    rc = lib.pystl_b200_fused_precompile(cf.ptr(), T, D, int(backward))
    return "fused" if rc == 0 else "interpreter"


def measured_peak_gbs() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    cfg, lo, hi, T_eval, n_t, seed = args
    from oracle import scalar

    sem_name, kw, _ = WORKLOADS[cfg]
    _, T, D = BASELINE_SHAPES[cfg[:4]]
    phi = from_spec(BASELINE_SPECS[cfg[:4]])
    sem = scalar.make(sem_name, **kw)
    rng = np.random.default_rng(seed)
    done = 0
    t0 = time.perf_counter()
    for b in range(lo, hi):
        sig = rng.standard_normal((T_eval, D)).astype(np.float32).astype(np.float64)
        ts = rng.integers(0, T_eval, size=n_t)
        for t in ts:
            scalar.evaluate_sliced(phi, sig, sem, int(t))
            done += 1
    return done, time.perf_counter() - t0


def cpu_port_throughput(cfg: str, budget_s: float, cores: int, pool=None, per: float | None = None) -> dict:
    """Oracle port of the reference recursion on a bounded sample: (trace, t) pairs per second.
    One evaluation = one trace-sample of the forward pass (the reference has no batched path).
    `pool` / `per` (seconds per evaluation on one core) let a caller reuse the workers and the
    calibration across steps."""
    import multiprocessing as mp

    _, T, _ = BASELINE_SHAPES[cfg[:4]]
    T_eval = min(T, 1024)
    if per is None:  # calibrate on one core
        n, dt = _cpu_worker((cfg, 0, 1, T_eval, 4, 0))
        per = dt / max(n, 1)
    total = max(int(budget_s / per), 8) * cores
    n_t = 8
    traces = max(total // n_t, cores)
    chunks = [(cfg, i * traces // cores, (i + 1) * traces // cores, T_eval, n_t, 100 + i) for i in range(cores)]
    t0 = time.perf_counter()
    if pool is not None:
        res = pool.map(_cpu_worker, chunks)
    elif cores > 1:
        with mp.get_context("fork").Pool(cores) as own:
            res = own.map(_cpu_worker, chunks)
    else:
        res = [_cpu_worker(chunks[0])]
    wall = time.perf_counter() - t0
    done = sum(r[0] for r in res)
    return {
        "value": done / wall,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"{done} (trace,t) pairs: {traces} traces x {n_t} random t, horizon {T_eval} of {T}, "
        f"slice identity (SURVEY A1), forward only, float64",
        "sec_per_eval_one_core": per,
    }


def open_ended_cases():
    from pystl_b200.formula import Always, Eventually, Interval, Or, Until, affine_predicate

    x = affine_predicate("x>0", {0: 1.0}, 0.0)
    y = affine_predicate("y<0.5", {1: -1.0}, 0.5)
    return (
        ("classical", {}, "G[0,inf)(x>0)", Always(x, Interval(0, None))),
        ("classical", {}, "(x>0) U[0,inf) (y<0.5)", Until(x, y, Interval(0, None))),
        ("smooth", {"temperature": 1.0}, "G[0,inf)(x>0) | F[5,inf)(y<0.5)", Or(Always(x, Interval(0, None)), Eventually(y, Interval(5, None)))),
    )


OPEN_ENDED_SHAPE = (8192, 16384, 2)


def open_ended_lines(dev) -> list:
    """Open-ended operators on traces longer than one tile, through the public API: the operand's formula
    kernel, the suffix-scan kernels of csrc/scan.cu and -- when the operator is not the root -- the parent formula kernel.  Not a
    BASELINE configuration (those are all bounded); reported next to them because `Interval.end = None` is part of the path
    (api.py:49-54).  CUDA-event time per call on torch's current stream, 3 calls after 2 warm-up calls."""
    import pystl_b200

    B, T, D = OPEN_ENDED_SHAPE
    cases = open_ended_cases()
    g = torch.Generator(device=dev).manual_seed(7)
    sig = torch.randn(B, T, D, device=dev, generator=g) * 0.5
    out = []
    for sem_name, kw, text, phi in cases:
        try:
            sem = pystl_b200.create_semantics(sem_name, backend="b200", **kw)
            for _ in range(2):
                sem.robustness_trace(phi, sig)
                sem.mean_robustness_and_grad(phi, sig)
            torch.cuda.synchronize(dev)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            for _ in range(3):
                sem.robustness_trace(phi, sig)
            ev[1].record()
            for _ in range(3):
                mean, grad = sem.mean_robustness_and_grad(phi, sig)
            ev[2].record()
            torch.cuda.synchronize(dev)
            f_ms, g_ms = ev[0].elapsed_time(ev[1]) / 3, ev[1].elapsed_time(ev[2]) / 3
            out.append({"semantics": sem_name, "formula": text, "traces": B, "horizon": T, "channels": D,
                        "values_ms": f_ms, "values_samples_per_s": B * T / (f_ms * 1e-3),
                        "mean_and_grad_ms": g_ms, "mean_and_grad_samples_per_s": B * T / (g_ms * 1e-3), "mean_robustness": float(mean)})
            del mean, grad
        except Exception as exc:  # a secondary line never fails the bench
            out.append({"semantics": sem_name, "formula": text, "error": f"{type(exc).__name__}: {exc}"[:200]})
    del sig
    torch.cuda.empty_cache()
    return out


def ctstl_lines(dev) -> list:
    """CT-STL's own operators on value traces through the public API (`B200CtstlSemantics`; not a BASELINE
    configuration): the cumulative-time operator C^tau_[0,63] with k = 16 (a selection, bound by the compare rate: 32 (W + 3)
    compares per output) and the Until with an exclusive prefix, values and gradients.  CUDA-event time per call, 3 calls after 2."""
    import pystl_b200

    B, T = 16384, 4096
    g = torch.Generator(device=dev).manual_seed(11)
    y = torch.randn(B, T, device=dev, generator=g).requires_grad_(True)
    l = torch.randn(B, T, device=dev, generator=g).requires_grad_(True)
    sem = pystl_b200.B200CtstlSemantics(delta=1.0)
    cases = (("C^16_[0,63] y", 64, lambda: sem.cumulative_time_trace(y, (0, 63), tau=16.0)),
             ("l U[0,24] y (exclusive prefix)", 0, lambda: sem.until_trace(l, y, (0, 24))))
    out = []
    for text, window, fn in cases:
        try:
            def both():
                o = fn()
                (gy,) = torch.autograd.grad(o.sum(), y)
                return gy
            for _ in range(2):
                both()
            torch.cuda.synchronize(dev)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            with torch.no_grad():
                for _ in range(3):
                    fn()
            ev[1].record()
            for _ in range(3):
                both()
            ev[2].record()
            torch.cuda.synchronize(dev)
            f_ms, g_ms = ev[0].elapsed_time(ev[1]) / 3, ev[1].elapsed_time(ev[2]) / 3
            line = {"formula": text, "traces": B, "horizon": T, "values_ms": f_ms, "values_samples_per_s": B * T / (f_ms * 1e-3),
                    "values_and_grad_ms": g_ms}
            if window:
                line["compares_per_s"] = 32.0 * (window + 3) * B * T / (f_ms * 1e-3)
                line["bound"] = "ALU pipe, one FSET per compare (= the issue rate of the FSET + FADD pairs): 148 SMs x 64 lanes x clock; ncu: ALU 85 %, issue 85 %"
            out.append(line)
        except Exception as exc:  # a secondary line never fails the bench
            out.append({"formula": text, "error": f"{type(exc).__name__}: {exc}"[:200]})
    del y, l
    torch.cuda.empty_cache()
    return out


def cpu_vectorised_throughput(cfg: str, budget_s: float, with_bwd: bool) -> dict:
    """Second reported CPU figure (baseline 2): the all-t, batched float64
    restatement with torch CPU ops on every host thread (oracle/dense.py mirrors the reference's
    torch backend operator by operator, so autograd gives the reference's gradients).  This is the
    strongest CPU formulation of the path, far above the reference's own per-(trace, t) recursion."""
    from oracle import dense

    sem_name, kw, _ = WORKLOADS[cfg]
    _, T, D = BASELINE_SHAPES[cfg[:4]]
    T_eval = min(T, 1024)
    phi = from_spec(BASELINE_SPECS[cfg[:4]])
    sem = dense.make(sem_name, **kw)
    gen = torch.Generator().manual_seed(7)
    Bp, done, t_used = 8, 0, 0.0
    while t_used < budget_s:
        sig = torch.randn((Bp, T_eval, D), generator=gen, dtype=torch.float64, requires_grad=with_bwd)
        t0 = time.perf_counter()
        out = dense.robustness_trace(phi, sig, sem)
        if with_bwd:
            torch.nan_to_num(out).mean().backward()
        dt = time.perf_counter() - t0
        if done or dt > 0.2 * budget_s:  # the first small batches only calibrate
            done += Bp * T_eval
            t_used += dt
        elif dt < 0.05 * budget_s:
            Bp *= 2
        else:
            done += Bp * T_eval
            t_used += dt
    return {
        "value": done / t_used,
        "unit": UNIT,
        "cores": torch.get_num_threads(),
        "kind": "port (vectorised torch float64, oracle/dense.py)",
        "sample": f"{done // T_eval} traces x horizon {T_eval} of {T}, all t, {'forward+backward' if with_bwd else 'forward'}, float64",
    }


def run_reference(args) -> None:
    """CPU arm.  Forward workloads: the oracle port of the reference's per-(trace, t) recursion
    (oracle/scalar.py) on every host core.  Forward+backward workloads: the reference differentiates
    by torch autograd through pystl/torch/semantics.py; its restatement with a backward leg is
    oracle/dense.py (same operators, batched over traces and t, float64, every torch thread), so a
    step is value AND gradient like the GPU arm's; the forward-only scalar figure rides along."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.workload
    _, _, with_bwd = WORKLOADS[cfg]
    if args.forward_only:
        with_bwd = False
    cores = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm is to use every host thread it can, at any N
    torch.set_num_threads(cores)
    steps = max(args.steps, 1)
    import multiprocessing as mp

    vals = []
    # every step is a bounded sample of the workload; the whole run (warm-up included) is sized
    # for about a minute of wall clock whatever --steps / --warmup are
    budget = max(60.0 / (steps + args.warmup), 0.25)
    scalar_fwd = None
    if with_bwd:
        with mp.get_context("fork").Pool(cores) as pool:  # (before torch spins up its own threads)
            scalar_fwd = cpu_port_throughput(cfg, 2.0, cores, pool)
        for _ in range(args.warmup):
            cpu_vectorised_throughput(cfg, budget, True)
        t0 = time.perf_counter()
        base = None
        for _ in range(steps):
            base = cpu_vectorised_throughput(cfg, budget, True)
            vals.append(base["value"])
        wall = time.perf_counter() - t0
        base["kind"] = "port"
        base["port"] = "oracle/dense.py: batched float64 restatement of pystl/torch/semantics.py, value + torch-autograd gradient"
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            per = cpu_port_throughput(cfg, 0.25, cores, pool)["sec_per_eval_one_core"]
            for _ in range(args.warmup):
                cpu_port_throughput(cfg, budget, cores, pool, per)
            t0 = time.perf_counter()
            base = None
            for _ in range(steps):
                base = cpu_port_throughput(cfg, budget, cores, pool, per)
                vals.append(base["value"])
            wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    base["value"] = value
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{cfg}: {DESCR[cfg]}", "pass": "forward+backward" if with_bwd else "forward",
                   "note": "CPU restatement of the reference path on the host cores, bounded sample"},
        "cpu_baseline": base,
        "cpu_baseline_scalar_forward": scalar_fwd,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- GPU arm
def default_batch(cfg: str, world: int = 1, strong: bool = False) -> int:
    B, _, _ = BASELINE_SHAPES[cfg[:4]]
    if cfg.startswith("cfg1xb"):
        return 1 << 19
    if cfg.startswith("cfg1x"):
        return 1 << 20
    if cfg[:4] == "cfg5":
        # BASELINE config 5: 2^20 traces sharded over the GPUs of one box; weak-scaling lines use the
        # shard of an 8-GPU run, the strong-scaling `cfg5_sharded` object splits 2^20 over the ranks
        return (1 << 20) // (world if strong else 8)
    return B


def measured_traffic() -> dict:
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu captures
    (profiles/traffic.json, written by tools/ncu_summary.py from `ncu --set full` reports)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


class Workload:
    """One BASELINE configuration resident on this rank's GPU: signal, outputs, compiled formula."""

    def __init__(self, cfg: str, B: int, dev, rank: int, with_bwd=None):
        import pystl_b200
        from pystl_b200 import capi

        self.cfg = cfg
        self.sem_name, kw, bwd = WORKLOADS[cfg]
        self.with_bwd = bwd if with_bwd is None else with_bwd
        _, self.T, self.D = BASELINE_SHAPES[cfg[:4]]
        self.B = B
        self.dev = dev
        self.phi = from_spec(BASELINE_SPECS[cfg[:4]])
        self.sem = pystl_b200.create_semantics(self.sem_name, backend="b200", **kw)
        self.lib = capi.lib()
        self.capi = capi
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        self.signal = torch.randn((B, self.T, self.D), generator=gen, device=dev, dtype=torch.float32)
        self.cf = self.sem.compile(self.phi, self.D)
        self.C_out = self.sem.out_channels
        self.out = torch.empty((B, self.T, self.C_out), dtype=torch.float32, device=dev)
        self.grad = torch.empty_like(self.signal) if self.with_bwd else None
        self.stream = torch.cuda.current_stream(dev)
        self.plan = self.sem.plan(self.phi, self.T, self.D)
        self.n_valid = B * max(self.T - self.plan["start_slack"], 0)
        self.tier = tier_of(self.lib, self.cf, self.T, self.D, self.with_bwd)
        if self.tier == "fused":  # report the geometry of the kernel that actually runs
            import ctypes

            fp = capi.Plan()
            if self.lib.pystl_b200_fused_plan(self.cf.ptr(), self.T, self.D, int(self.with_bwd), ctypes.byref(fp)) == 0:
                self.plan = fp.as_dict()
        # the interpreter tier has no fused value-and-gradient kernel: forward + backward launches
        self.launches_per_step = 2 if (self.with_bwd and self.tier == "interpreter") else 1

    def check(self, rc):
        self.capi.raise_for(rc, self.lib.pystl_b200_last_error())

    def forward(self, acc_ptr=None):
        self.check(self.lib.pystl_b200_eval_forward(self.cf.ptr(), self.signal.data_ptr(), self.B, self.T, self.D, self.out.data_ptr(),
                                                    acc_ptr, self.stream.cuda_stream))

    def backward(self, scale):
        self.check(self.lib.pystl_b200_eval_backward(self.cf.ptr(), self.signal.data_ptr(), self.B, self.T, self.D, None, scale,
                                                     self.grad.data_ptr(), self.stream.cuda_stream))

    def value_and_grad(self, scale, acc_ptr=None):
        # value AND gradient of the mean-robustness loss through the C ABI: one fused kernel reads the
        # signal once, writes out[B,T], its mean accumulator and grad_signal[B,T,D]
        self.check(self.lib.pystl_b200_eval_value_and_grad(self.cf.ptr(), self.signal.data_ptr(), self.B, self.T, self.D, None, scale,
                                                           self.out.data_ptr(), acc_ptr, self.grad.data_ptr(), self.stream.cuda_stream))

    def time_kernel(self, fn, n: int) -> float:
        """mean CUDA-event time of n launches (ms), events on the launching stream, after one warm-up"""
        fn()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b_ in evs:
            a.record(self.stream)
            fn()
            b_.record(self.stream)
        torch.cuda.synchronize(self.dev)
        return float(np.mean([a.elapsed_time(b_) for a, b_ in evs]))

    def roof(self, kind: str, ms_: float, bytes_per_sample: float, peak: float, traffic: dict) -> dict:
        gbs = bytes_per_sample * self.B * self.T / (ms_ * 1e-3) / 1e9
        name = f"{self.tier}:{kind}<{self.sem_name}>"
        line = {"workload": self.cfg, "kernel": name, "traces": self.B, "kernel_ms": ms_, "achieved": gbs, "frac": gbs / peak,
                "algorithmic_bytes_per_sample": bytes_per_sample, "samples_per_s": self.B * self.T / (ms_ * 1e-3),
                "traffic": traffic.get(f"{self.cfg}:{kind}<{self.sem_name}>")}
        src = traffic.get("_source", {}).get(f"{self.cfg}:{kind}<{self.sem_name}>")
        if src:
            line["traffic_source"] = src  # which committed ncu capture the DRAM bytes come from
        floor = pipe_floor(self.cfg, kind, self.B, self.T, ms_, SM_MAX_MHZ)
        if floor:
            line["pipe_floor"] = floor  # the ceiling that binds this kernel; `frac` above stays the HBM fraction
        return line

    def free(self):
        self.signal = self.out = self.grad = None
        torch.cuda.empty_cache()


# every BASELINE configuration at its named size: (workload key, passes to time)
OTHER_KERNELS = [
    ("cfg1", ("forward",)), ("cfg1x", ("forward",)), ("cfg1xb", ("value_and_grad",)),
    ("cfg2", ("forward",)), ("cfg3", ("forward", "value_and_grad")), ("cfg3d", ("forward", "value_and_grad")),
    ("cfg4", ("forward",)), ("cfg4b", ("value_and_grad",)),
    ("cfg5", ("forward",)), ("cfg5s", ("forward",)), ("cfg5c", ("forward",)), ("cfg5a", ("forward",)), ("cfg5d", ("forward",)),
    ("cfg5b", ("value_and_grad",)), ("cfg5sb", ("value_and_grad",)), ("cfg5cb", ("value_and_grad",)), ("cfg5ab", ("value_and_grad",)),
    ("cfg5db", ("value_and_grad",)),
]


def run_b200(args) -> None:
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = args.workload
    B = args.batch or default_batch(cfg)
    w = Workload(cfg, B, dev, rank, with_bwd=False if args.forward_only else None)
    with_bwd, T, D = w.with_bwd, w.T, w.D
    steps = args.steps
    warm = max(args.warmup, 3)
    scale = 1.0 / (world * w.n_valid)  # constant cotangent: the loss is the mean over the sharded batch
    # one (sum, count) accumulator per step: no memset inside the loop, and the all-reduce of step i
    # (NCCL, asynchronous) overlaps the kernel of step i + 1
    accs = torch.zeros((steps + warm, 2), dtype=torch.float64, device=dev)
    pending = []

    def step(i: int) -> int:
        a = accs[i]
        if with_bwd and w.launches_per_step == 1:
            w.value_and_grad(scale, a.data_ptr())
        else:
            w.forward(a.data_ptr())
            if with_bwd:
                w.backward(scale)
        if world > 1:
            pending.append(dist.all_reduce(a, async_op=True))
        return w.launches_per_step

    def drain():
        for work in pending:
            work.wait()
        pending.clear()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(warm):
        step(i)
    drain()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    launches = 0
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(w.stream)
    for i in range(steps):
        launches += step(warm + i)
    drain()  # every step's mean is final inside the timed region
    t_end.record(w.stream)
    barrier()
    clocks = sampler.stop()
    ms = t_start.elapsed_time(t_end)
    if world > 1:
        tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    samples_per_step = world * B * T
    value = samples_per_step * steps / (ms * 1e-3)
    mean_last = float(accs[warm + steps - 1][0].item()) / (world * w.n_valid) if w.n_valid else float("nan")

    # kernels alone, CUDA events on the launching stream
    peak, peak_src = measured_peak_gbs()
    traffic = measured_traffic()
    k_ms = w.time_kernel(lambda: w.forward(None), steps)
    bytes_fwd = algorithmic_bytes_per_sample(cfg, False)
    fwd_roof = w.roof("forward", k_ms, bytes_fwd, peak, traffic)
    roofline = {"bound": "hbm", **fwd_roof, "peak": peak, "unit": "GB/s", "peak_source": peak_src}
    if with_bwd:
        b_ms = w.time_kernel(lambda: w.backward(scale), steps)
        bwd_roof = w.roof("backward", b_ms, algorithmic_bytes_per_sample(cfg, True) - bytes_fwd, peak, traffic)
        if w.launches_per_step == 1:
            # the step's only kernel: reads the signal once, writes the trace and the gradient
            vg_ms = w.time_kernel(lambda: w.value_and_grad(scale, None), steps)
            vg = w.roof("value_and_grad", vg_ms, bytes_fwd + 4 * D, peak, traffic)
            roofline = {**roofline, **vg, "other_kernels": [fwd_roof, bwd_roof]}
        else:
            dom, oth = (bwd_roof, fwd_roof) if b_ms > k_ms else (fwd_roof, bwd_roof)
            roofline = {**roofline, **dom, "other_kernels": [oth]}
    roofline.setdefault("other_kernels", [])

    # end to end through the public API: the WHOLE batch in pinned host memory, streamed in chunks
    # (H2D of chunk i+1 and D2H of the gradient of chunk i-1 overlap the kernel of chunk i)
    e2e = None
    if not args.no_e2e:
        try:
            host = torch.empty((B, T, D), dtype=torch.float32, pin_memory=True)
            host.copy_(w.signal)
            ghost = torch.empty((B, T, D), dtype=torch.float32, pin_memory=True) if with_bwd else None
            e_steps = max(2, min(steps, 5))

            def e2e_step():
                mean, _ = w.sem.mean_robustness_and_grad_host(w.phi, host, ghost, device=dev, with_grad=with_bwd)
                return float(mean.cpu())

            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                e2e_step()
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            if world > 1:
                tm = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                dt = float(tm.item())
            h2d = B * T * D * 4
            d2h = (B * T * D * 4 if with_bwd else 0) + 8
            e2e = {
                "value": world * B * T * e_steps / dt,
                "unit": UNIT,
                "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "traces_per_step": B,
                "ms_per_step": 1e3 * dt / e_steps,
                "h2d_gbs": h2d * e_steps / dt / 1e9,
                "d2h_gbs": d2h * e_steps / dt / 1e9,
                "what": "sem.mean_robustness_and_grad_host: pinned host batch in, mean robustness and (pinned) host gradient out; "
                        "chunked, copies overlapped with the kernel" if with_bwd else
                        "sem.mean_robustness_and_grad_host(with_grad=False): pinned host batch in, mean robustness out",
            }
            del host, ghost
        except NotImplementedError as exc:
            e2e = {"value": None, "unit": UNIT, "error": str(exc)}

    head_plan, head_tier = w.plan, w.tier
    w.free()

    # every other BASELINE configuration at its named size, same run (kernel-only, a few launches each)
    if not args.no_others:
        n_l = 5
        for key, passes in OTHER_KERNELS:
            if key == cfg and not args.forward_only and all((p == "value_and_grad") == with_bwd for p in passes):
                continue  # the headline kernel itself
            try:
                ww = Workload(key, default_batch(key), dev, rank, with_bwd=("value_and_grad" in passes))
                sc = 1.0 / (world * max(ww.n_valid, 1))
                for ps in passes:
                    if ps == "forward":
                        ms_k = ww.time_kernel(lambda: ww.forward(None), n_l)
                        r = ww.roof("forward", ms_k, algorithmic_bytes_per_sample(key, False), peak, traffic)
                    else:
                        if ww.launches_per_step != 1:
                            ms_k = ww.time_kernel(lambda: (ww.forward(None), ww.backward(sc)), n_l)
                        else:
                            ms_k = ww.time_kernel(lambda: ww.value_and_grad(sc, None), n_l)
                        r = ww.roof("value_and_grad", ms_k, algorithmic_bytes_per_sample(key, False) + 4 * ww.D, peak, traffic)
                    launches += n_l + 1
                    roofline["other_kernels"].append(r)
                ww.free()
            except Exception as exc:  # a secondary line never fails the bench
                roofline["other_kernels"].append({"workload": key, "error": f"{type(exc).__name__}: {exc}"[:200]})
        if world > 1:  # max over ranks of every secondary kernel time
            t = torch.tensor([r.get("kernel_ms", 0.0) for r in roofline["other_kernels"]], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            for r, v in zip(roofline["other_kernels"], t.tolist()):
                if "kernel_ms" in r and v > 0:
                    k = r["kernel_ms"] / v
                    r["kernel_ms"] = v
                    for f in ("achieved", "frac", "samples_per_s"):
                        r[f] *= k

    # BASELINE config 5 as named: 2^20 traces x 2048 x 8 split over the ranks, NCCL mean reduction
    sharded = None
    if not args.no_sharded:
        sharded = {"scaling": "strong", "total_traces": 1 << 20, "horizon": 2048, "channels": 8, "n_gpus": world,
                   "note": "forward of every trace + fused mean + one 16-byte NCCL all-reduce per step; device-timed, max over ranks",
                   "semantics": {}}
        Bs = (1 << 20) // world
        n_s = 5
        for key in ("cfg5", "cfg5s", "cfg5c", "cfg5a", "cfg5d"):
            try:
                ww = Workload(key, Bs, dev, rank, with_bwd=False)
                acs = torch.zeros((n_s + 1, 2), dtype=torch.float64, device=dev)

                def sstep(i):
                    ww.forward(acs[i].data_ptr())
                    if world > 1:
                        pending.append(dist.all_reduce(acs[i], async_op=True))

                sstep(n_s)
                drain()
                barrier()
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(ww.stream)
                for i in range(n_s):
                    sstep(i)
                drain()
                b_.record(ww.stream)
                barrier()
                t_ms = a.elapsed_time(b_)
                if world > 1:
                    tm = torch.tensor([t_ms], dtype=torch.float64, device=dev)
                    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                    t_ms = float(tm.item())
                launches += n_s + 1
                bytes_ = algorithmic_bytes_per_sample(key, False)
                tot = (1 << 20) * 2048
                sharded["semantics"][ww.sem_name] = {
                    "value": tot * n_s / (t_ms * 1e-3), "unit": UNIT, "ms_per_step": t_ms / n_s,
                    "hbm_frac_per_gpu": bytes_ * tot / world / (t_ms / n_s * 1e-3) / 1e9 / peak,
                    "mean_robustness": float(acs[0][0].item()) / (world * max(ww.n_valid, 1)),
                }
                ww.free()
            except Exception as exc:
                sharded["semantics"][key] = {"error": f"{type(exc).__name__}: {exc}"[:200]}

    open_ended = open_ended_lines(dev) if (world == 1 and not args.no_others) else None
    ctstl = ctstl_lines(dev) if (world == 1 and not args.no_others) else None

    cpu = None
    cpu_vec = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_port_throughput(cfg, args.cpu_seconds, 1)
        try:
            cpu_vec = cpu_vectorised_throughput(cfg, min(args.cpu_seconds, 10.0), with_bwd)
        except Exception as e:  # reported baseline only: never fail the bench on it
            cpu_vec = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": steps,
            "warmup": warm,
            "ms_per_step": ms / steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": f"{cfg}: {DESCR[cfg]}",
                "traces_per_gpu": B,
                "horizon": T,
                "channels": D,
                "pass": "forward+backward" if with_bwd else "forward",
                "l2": "inputs larger than L2 (no flush needed)" if B * T * D * 4 > (1 << 28) else "input fits L2 (CPU-comparison config)",
                "plan": head_plan,
                "tier": head_tier,
                "mean_robustness_last_step": mean_last,
            },
            "roofline": roofline,
            "cfg5_sharded": sharded,
            "open_ended": open_ended,
            "ctstl": ctstl,
            "cpu_baseline": cpu,
            "cpu_baseline_vectorised": cpu_vec,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override traces per GPU")
    ap.add_argument("--forward-only", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the kernel lines of the other BASELINE configurations")
    ap.add_argument("--no-sharded", action="store_true", help="skip BASELINE config 5 at its full sharded size")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
