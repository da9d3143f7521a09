"""Host-side mirror of the image partition used for multi-GPU rendering.

The frame is cut into tw x th tiles; tile k (row-major over tiles) belongs to rank k % count - the same rule
csrc/core.cu MakeTiling + k_raygen apply on the device.  Used by bench.py / tests to reason about ownership
without a GPU (e.g. the 2-rank gloo test).
"""
from __future__ import annotations

import numpy as np


def tile_owner(nx: int, ny: int, count: int, tw: int = 16, th: int = 16) -> np.ndarray:
    """(ny, nx) int32 array: rank that renders each pixel."""
    tiles_x = (nx + tw - 1) // tw
    ty, tx = np.mgrid[0:ny, 0:nx]
    k = (ty // th) * tiles_x + (tx // tw)
    return (k % max(1, count)).astype(np.int32)


def tile_mask(nx: int, ny: int, rank: int, count: int, tw: int = 16, th: int = 16) -> np.ndarray:
    return tile_owner(nx, ny, count, tw, th) == rank


def local_tile_count(nx: int, ny: int, rank: int, count: int, tw: int = 16, th: int = 16) -> int:
    ntiles = ((nx + tw - 1) // tw) * ((ny + th - 1) // th)
    return (ntiles - rank + count - 1) // count if ntiles > rank else 0
