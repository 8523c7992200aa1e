"""Partitioning of independent units (query points, optimiser restarts) over ranks: one process per
GPU, no data-path collective.  torch.distributed is used only by callers for the
rendezvous, barriers and the max-over-ranks timing reduction."""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slab [lo, hi) of rank `rank`; slabs differ in size by at most one item."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def round_robin(n_items: int, rank: int, world: int) -> list[int]:
    """Indices of the restarts owned by `rank` (restart r -> rank r mod world)."""
    return list(range(rank, n_items, world))
