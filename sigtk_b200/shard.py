"""Read-range sharding across GPUs: reads are independent, so a batch is split into contiguous
read ranges with (nearly) equal sample counts, one range per rank; there is no collective on the data path, only
a host-side gather of per-read results in read order."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_ranges(read_lens: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous read ranges [lo, hi) per rank, balanced by the number of samples; a read is never split."""
    n = len(read_lens)
    if world <= 1:
        return [(0, n)]
    csum = np.concatenate([[0], np.cumsum(np.asarray(read_lens, dtype=np.int64))])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        k = int(np.searchsorted(csum, target, side="left"))
        # choose the read boundary closest to the target, never moving backwards
        if k > 0 and k <= n and abs(int(csum[k - 1]) - target) < abs(int(csum[min(k, n)]) - target):
            k -= 1
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_in_read_order(per_rank_results: List[list]) -> list:
    """host-side gather: ranks hold consecutive read ranges, so concatenation restores the input order"""
    out = []
    for part in per_rank_results:
        out.extend(part)
    return out
