"""Host-side row partitioning for the row-sharded ONLYUNITDIAG path.

Rank r of `world` owns the contiguous rows [r*rpr, min(n, (r+1)*rpr)) with rpr = ceil(n / world); this is the layout
csrc/dist.h assumes (equal-count NCCL all-gather with zero padding on the last rank).  A shard of C is the CSC slice
of the owned COLUMNS: for the symmetric C of a MaxCut instance these are the owned rows, and in general they are
exactly the lists the kernel needs ((Y*C)(:, j) = sum_i C(i, j) Y(:, i), ManiSDP_onlyunitdiag.m:118)."""
from __future__ import annotations


def rows_per_rank(n: int, world: int) -> int:
    return (n + world - 1) // world if world > 1 else n


def row_range(n: int, world: int, rank: int):
    rpr = rows_per_rank(n, world)
    return min(n, rank * rpr), min(n, (rank + 1) * rpr)


def shard_C(C_csc, world: int, rank: int):
    """Columns [row_begin, row_end) of C (CSC), global row indices kept."""
    n = C_csc.shape[0]
    r0, r1 = row_range(n, world, rank)
    return C_csc[:, r0:r1], r0, r1
