"""Read sharding for multi-GPU runs: the index is replicated on every GPU, reads are split into
contiguous ranges of (almost) equal total bases, one range per rank, and every rank writes its
results into its own slice of the output -- no collective on the data path.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(offsets: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Split reads [0, n) into `world` contiguous ranges balancing the number of bases.
    offsets: int64[n+1]. Returns [(r0, r1)] per rank; ranges may be empty."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n = offsets.size - 1
    base0, total = int(offsets[0]), int(offsets[-1] - offsets[0])
    cuts = [0]
    for r in range(1, world):
        target = base0 + (total * r) // world
        cuts.append(int(np.searchsorted(offsets, target, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.minimum(cuts, n))
    return [(int(cuts[i]), int(cuts[i + 1])) for i in range(world)]


def output_bounds(offsets: np.ndarray, k: int, bounds: list[tuple[int, int]]) -> list[tuple[int, int]]:
    """Result slices matching shard_bounds: read i produces max(0, len_i - k + 1) values."""
    lens = np.diff(np.asarray(offsets, dtype=np.int64))
    cnt = np.concatenate([[0], np.cumsum(np.maximum(lens - k + 1, 0))])
    return [(int(cnt[a]), int(cnt[b])) for a, b in bounds]


def shard_batch(ascii_: np.ndarray, offsets: np.ndarray, rank: int, world: int) -> tuple[np.ndarray, np.ndarray, tuple[int, int]]:
    """The (ascii, offsets) batch of one rank, offsets rebased to 0, plus its read range."""
    r0, r1 = shard_bounds(offsets, world)[rank]
    off = np.ascontiguousarray(offsets[r0:r1 + 1] - offsets[r0])
    return np.ascontiguousarray(ascii_[offsets[r0]:offsets[r1]]), off, (r0, r1)
