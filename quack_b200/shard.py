"""Sharding of reads over ranks/devices: every read is an independent unit and every
output is a sum of per-read contributions, so shards need no data-path collective -- only the final
integer sum of the count arrays (one ncclReduce inside qb_finish)."""
from __future__ import annotations

import numpy as np


def shard_range(rank: int, world: int, n_units: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) of `n_units` for `rank` of `world` (sizes differ by at most 1)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(n_units, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def merge_rows(parts, adapters_enabled: bool):
    """Host restatement of what the reduce produces from per-shard RAW results [(rows, max_length, n_reads)]:
    element-wise sum over the longest shard, max_length = longest, n_reads = sum.  Without -a every shard
    carries kmer_count[10] = its number of reads longer than 10 (reference quack.c:210-217), which sums
    correctly as well."""
    ml = max(p[1] for p in parts)
    rows = np.zeros((ml, 97), dtype=np.uint64)
    for r, m, _ in parts:
        rows[:m] += np.asarray(r, dtype=np.uint64)[:m]
    return rows, ml, sum(p[2] for p in parts)
