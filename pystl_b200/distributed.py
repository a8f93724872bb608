"""Batch sharding across ranks (one process per GPU) and the mean-robustness reduction.

The path shards by trace: every trace is independent (the reference has no batch
axis at all, api.py:215-302), so each rank evaluates a contiguous slice of the
batch with no data-path collective.  The only exchange is the optional mean
robustness: each rank's kernel accumulates ``(sum, count)`` and ONE 16-byte
all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests) finishes it.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch


def shard_bounds(n_traces: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous ``[lo, hi)`` slice of the batch owned by ``rank`` (sizes differ by at most 1)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    base, rem = divmod(int(n_traces), world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_mean(acc: torch.Tensor, *, group=None) -> torch.Tensor:
    """``acc = (sum, count)`` float64 on this rank -> global mean (NaN if no valid output anywhere)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc[0] / acc[1]


def all_reduce_count(n_local: float, device, *, group=None) -> float:
    """Number of valid outputs over all ranks (the denominator of the sharded mean; its reciprocal
    is the constant cotangent of the fused value-and-gradient kernel)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.tensor([n_local], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return float(t.item())
    return float(n_local)
