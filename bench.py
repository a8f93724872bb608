#!/usr/bin/env python
"""Benchmark of the robustness-evaluation hot path (BASELINE.json metric: trace-samples/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl b200|reference]

A *step* is one pass of the hot path over one batch of synthetic traces: the fused
formula kernel over ``signal[B, T, D]`` (plus the gradient kernel for the forward+backward
workloads).  ``value`` is whole-job trace-samples/sec with the inputs resident in HBM;
``e2e`` is the same metric through the public API with pinned HOST buffers (H2D copy of
the batch and D2H read of the mean robustness inside the timed region).

Under torchrun (N > 1) every rank evaluates its own shard of the batch (weak scaling: the
per-GPU batch is fixed) and the mean robustness is finished by one 16-byte NCCL
all-reduce; the time is the max over ranks of the device-timed region.

``--impl reference`` times the CPU implementation of the same path: the oracle port of the
reference's per-(trace, t) recursion (oracle/scalar.py; the Python reference itself cannot
travel to the GPU box), on all host cores, on a bounded sample of the same workload.
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.workload
    cores = os.cpu_count() or 1
    steps = max(args.steps, 1)
    import multiprocessing as mp

    vals = []
    # every step is a bounded sample of the workload; the whole run (warm-up included) is sized
    # for about a minute of wall clock whatever --steps / --warmup are
    budget = max(60.0 / (steps + args.warmup), 0.25)
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
        "config": {"workload": f"{cfg}: {DESCR[cfg]}", "note": "CPU oracle port of the reference recursion, bounded sample"},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- GPU arm
def run_b200(args) -> None:
    import torch.distributed as dist

    import pystl_b200
    from pystl_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = args.workload
    sem_name, kw, with_bwd = WORKLOADS[cfg]
    if args.forward_only:
        with_bwd = False
    B, T, D = BASELINE_SHAPES[cfg[:4]]
    if args.batch:
        B = args.batch
    if cfg == "cfg1x" and not args.batch:
        B = 1 << 20
    if cfg == "cfg1xb" and not args.batch:
        B = 1 << 19
    if cfg[:4] == "cfg5" and not args.batch:
        # 2^20 traces are sharded over 8 GPUs in the BASELINE config: 2^17 per GPU
        B = (1 << 20) // 8
    phi = from_spec(BASELINE_SPECS[cfg[:4]])
    sem = pystl_b200.create_semantics(sem_name, backend="b200", **kw)
    capi.lib()  # fail loudly if the CUDA library is missing

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    signal = torch.randn((B, T, D), generator=gen, device=dev, dtype=torch.float32)
    cf = sem.compile(phi, D)
    C_out = sem.out_channels
    out = torch.empty((B, T, C_out), dtype=torch.float32, device=dev)
    acc = torch.zeros(2, dtype=torch.float64, device=dev)
    grad = torch.empty_like(signal) if with_bwd else None
    stream = torch.cuda.current_stream(dev)
    lib = capi.lib()
    plan = sem.plan(phi, T, D)
    n_valid = B * max(T - plan["start_slack"], 0)
    launches = 0
    tier = tier_of(lib, cf, T, D, with_bwd)
    if tier == "fused":  # report the geometry of the kernel that actually runs
        import ctypes

        fp = capi.Plan()
        if lib.pystl_b200_fused_plan(cf.ptr(), T, D, int(with_bwd), ctypes.byref(fp)) == 0:
            plan = fp.as_dict()
    # the interpreter tier has no fused value-and-gradient kernel: forward + backward launches
    launches_per_step = 2 if (with_bwd and tier == "interpreter") else 1

    def step() -> int:
        acc.zero_()
        if with_bwd:
            # value AND gradient of the mean-robustness loss through the C ABI: one fused kernel reads
            # the signal once, writes out[B,T], its mean accumulator and grad_signal[B,T,D]
            # (constant cotangent 1 / (world * n_valid): the loss is the mean over the sharded batch)
            rc = lib.pystl_b200_eval_value_and_grad(
                cf.ptr(), signal.data_ptr(), B, T, D, None, 1.0 / (world * n_valid), out.data_ptr(), acc.data_ptr(),
                grad.data_ptr(), stream.cuda_stream,
            )
        else:
            rc = lib.pystl_b200_eval_forward(cf.ptr(), signal.data_ptr(), B, T, D, out.data_ptr(), acc.data_ptr(), stream.cuda_stream)
        capi.raise_for(rc, lib.pystl_b200_last_error())
        if world > 1:
            dist.all_reduce(acc)
        return launches_per_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # kernel-only time of the forward kernel (the dominant one) for the roofline
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for i in range(args.steps):
        launches += step()
    t_end.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = t_start.elapsed_time(t_end)
    if world > 1:
        tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    samples_per_step = world * B * T
    value = samples_per_step * args.steps / (ms * 1e-3)

    # kernels alone, CUDA events on the launching stream
    def time_kernel(fn) -> float:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        capi.raise_for(fn(), lib.pystl_b200_last_error())  # warm-up (first use may compile the kernel)
        for a, b_ in evs:
            a.record(stream)
            capi.raise_for(fn(), lib.pystl_b200_last_error())
            b_.record(stream)
        torch.cuda.synchronize(dev)
        return float(np.mean([a.elapsed_time(b_) for a, b_ in evs]))

    peak, peak_src = measured_peak_gbs()

    def roof(name, ms_, bytes_per_sample):
        gbs = bytes_per_sample * B * T / (ms_ * 1e-3) / 1e9
        return {"kernel": name, "achieved": gbs, "frac": gbs / peak, "algorithmic_bytes_per_sample": bytes_per_sample, "kernel_ms": ms_}

    k_ms = time_kernel(lambda: lib.pystl_b200_eval_forward(cf.ptr(), signal.data_ptr(), B, T, D, out.data_ptr(), None, stream.cuda_stream))
    bytes_fwd = algorithmic_bytes_per_sample(cfg, False)
    fwd_roof = roof(f"{tier}:forward<{sem_name}>", k_ms, bytes_fwd)
    roofline = {"bound": "hbm", **fwd_roof, "peak": peak, "unit": "GB/s", "traffic": None, "peak_source": peak_src}
    if with_bwd:
        scale = 1.0 / (world * n_valid)
        b_ms = time_kernel(lambda: lib.pystl_b200_eval_backward(cf.ptr(), signal.data_ptr(), B, T, D, None, scale, grad.data_ptr(), stream.cuda_stream))
        bwd_roof = roof(f"{tier}:backward<{sem_name}>", b_ms, algorithmic_bytes_per_sample(cfg, True) - bytes_fwd)
        if launches_per_step == 1:
            # the step's only kernel: reads the signal once, writes the trace and the gradient
            vg_ms = time_kernel(lambda: lib.pystl_b200_eval_value_and_grad(
                cf.ptr(), signal.data_ptr(), B, T, D, None, scale, out.data_ptr(), None, grad.data_ptr(), stream.cuda_stream))
            vg = roof(f"{tier}:value_and_grad<{sem_name}>", vg_ms, bytes_fwd + 4 * D)
            roofline = {**roofline, **vg, "other_kernels": [fwd_roof, bwd_roof]}
        else:
            dom, oth = (bwd_roof, fwd_roof) if b_ms > k_ms else (fwd_roof, bwd_roof)
            roofline = {**roofline, **dom, "other_kernels": [oth]}

    # end to end through the public API with pinned host buffers
    e2e = None
    if not args.no_e2e:
        Bh = min(B, max(1, (1 << 28) // (T * D)))  # <= 1 GiB of pinned input per step
        host = torch.empty((Bh, T, D), dtype=torch.float32).pin_memory()
        host.copy_(signal[:Bh].cpu())
        e_steps = max(2, min(args.steps, 5))

        def e2e_step():
            x = host.to(dev, non_blocking=True)
            if with_bwd:
                loss, g_ = sem.mean_robustness_and_grad(phi, x)
                return float(loss.cpu()), g_
            return float(sem.mean_robustness(phi, x).cpu()), None

        try:
            e2e_step()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(e_steps):
                e2e_step()
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            e2e = {
                "value": world * Bh * T * e_steps / dt,
                "unit": UNIT,
                "h2d_bytes_per_step": Bh * T * D * 4,
                "d2h_bytes_per_step": 8,
                "traces_per_step": Bh,
            }
        except NotImplementedError as exc:
            e2e = {"value": None, "unit": UNIT, "error": str(exc)}

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
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps,
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
                "plan": plan,
                "tier": tier,
            },
            "roofline": roofline,
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
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
