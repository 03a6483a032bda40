"""Multi-GPU partitioning of the path.  Every (CPI, channel) is independent through all
stages, so CPIs are dealt to ranks in contiguous blocks, each rank runs the
whole chain on its GPU, and per-rank detection lists are concatenated: NO collective on the
data path.  One process per GPU; torch.distributed is used by the callers only for the
barrier / max-over-ranks timing and to gather the (tiny) detection lists."""
from __future__ import annotations

import numpy as np


def shard_cpis(n_cpi: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block [start, stop) of CPIs owned by `rank`; blocks differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_cpi, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def merge_detections(per_rank: list, starts: list) -> np.ndarray:
    """Concatenate per-rank detection lists (CPI indices local to each rank) into one list in
    canonical (cpi, channel, range, doppler) order with global CPI indices."""
    out = []
    for dets, start in zip(per_rank, starts):
        d = np.array(dets, copy=True)
        d["cpi"] += start
        out.append(d)
    if not out:
        return np.zeros(0, dtype=[("cpi", "<i4"), ("channel", "<i4"), ("range", "<i4"), ("doppler", "<i4"), ("mag", "<f4")])
    allr = np.concatenate(out)
    order = np.lexsort((allr["doppler"], allr["range"], allr["channel"], allr["cpi"]))
    return allr[order]
