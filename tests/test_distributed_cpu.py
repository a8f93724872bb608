"""CPU tier: the N > 1 path (batch sharding + the 16-byte mean-robustness all-reduce) with two gloo
ranks.  Per-rank partial sums come from the host emulator of the kernel; the product's
`shard_bounds` / `all_reduce_mean` are what is under test."""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pystl_b200.distributed import all_reduce_mean, shard_bounds


def test_shard_bounds_cover_the_batch():
    for n, w in [(10, 3), (8, 8), (3, 8), (1 << 20, 8), (0, 2)]:
        spans = [shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 3, 3)


def _worker(rank: int, world: int, port: int, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emu_lib
    from parity_util import mixed_signal
    from pystl_b200.formula import BASELINE_SPECS, from_spec

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sig = mixed_signal(np.random.default_rng(99), 11, 150, 2)  # same on every rank
        lo, hi = shard_bounds(sig.shape[0], rank, world)
        phi = from_spec(BASELINE_SPECS["cfg1"])
        out, acc = emu_lib.eval_forward(phi, "smooth", sig[lo:hi], mean=True, temperature=0.5)
        acc[1] = out.size
        mean = all_reduce_mean(torch.tensor(acc))
        q.put((rank, float(mean), float(out.astype(np.float64).sum()), int(out.size)))
    finally:
        dist.destroy_process_group()


def test_two_rank_mean_reduction_matches_single_process():
    import emu_lib
    from parity_util import mixed_signal
    from pystl_b200.formula import BASELINE_SPECS, from_spec

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sig = mixed_signal(np.random.default_rng(99), 11, 150, 2)
    full = emu_lib.eval_forward(from_spec(BASELINE_SPECS["cfg1"]), "smooth", sig, temperature=0.5)
    want = float(full.astype(np.float64).mean())
    assert res[0][1] == pytest.approx(want, rel=1e-6) and res[1][1] == pytest.approx(want, rel=1e-6)
    assert res[0][3] + res[1][3] == full.size  # the shards partition the batch
