"""Partition helpers for the multi-GPU paths.  One process per GPU (torchrun); batched 1-D lines
are independent, so sharding them needs NO collective on the data path -- torch.distributed is only used for
barriers and for the max-over-ranks of device timings.  Row sharding of 2-D grids (sharded2d.py, csrc/gs_sharded2d.cu) uses the
same contiguous-block rule."""
from __future__ import annotations

from typing import Tuple


def block_shard(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition: (start, count) of `rank`'s block; the first total % world ranks get one
    extra item.  Covers [0, total) exactly once for every world >= 1."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


line_shard = block_shard   # batched 1-D: lines
row_shard = block_shard    # 2-D: rows of grid / predict / result


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (identity when not distributed)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def whole_job_rate(units_per_rank: float, world: int, max_seconds: float) -> float:
    """Aggregate throughput of a weak-scaling run: all ranks' units over the slowest rank's time."""
    return units_per_rank * world / max_seconds
