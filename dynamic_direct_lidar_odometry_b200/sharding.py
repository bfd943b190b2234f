"""Partitioning of independent registrations over ranks (BASELINE config C5).

Scan pairs are independent, so the only multi-GPU structure is a contiguous split of the unit range
and a max-over-ranks of the per-rank device time; there is no data-path collective.  The reduction
uses whatever torch.distributed backend the launcher initialised (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import Sequence, Tuple


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [begin, end) of `n_units` for `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, extra = divmod(n_units, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(values: Sequence[float], dist=None, device=None):
    """Element-wise maximum of `values` over all ranks (identity without a process group)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch

    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def sum_over_ranks(values: Sequence[float], dist=None, device=None):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch

    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.cpu()]
