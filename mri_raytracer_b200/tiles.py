"""Integer pixel <-> tile <-> rank maps — pure Python ints.

The dispatch geometry of the reference (``[numthreads(8,8,1)]``, ``thread_count=[W,H,1]``,
guard ``any(tid.xy >= imageSize)``: inr/viewer/brats_rt.slang:86-89,
inr/viewer/brats_viewer.py:431-432) as an integer contract that must be bit-exact between
this file, csrc/tiles.h and the kernels (tests/test_tiles.py, tests/test_gpu_tiles.py).
"""
from __future__ import annotations

from typing import List, Tuple

TILE = 8
SHIFT = 3
MASK = 7


def tiles_x(W: int) -> int:
    return (W + MASK) >> SHIFT


def tiles_y(H: int) -> int:
    return (H + MASK) >> SHIFT


def tile_count(W: int, H: int) -> int:
    return tiles_x(W) * tiles_y(H)


def tile_of_pixel(x: int, y: int, W: int) -> int:
    return (y >> SHIFT) * tiles_x(W) + (x >> SHIFT)


def lane_of_pixel(x: int, y: int) -> int:
    return ((y & MASK) << SHIFT) + (x & MASK)


def linear_pixel(x: int, y: int, W: int) -> int:
    return y * W + x


def pixel_of_tile_lane(tile: int, lane: int, W: int) -> Tuple[int, int]:
    tx, ty = tile % tiles_x(W), tile // tiles_x(W)
    return (tx << SHIFT) + (lane & MASK), (ty << SHIFT) + (lane >> SHIFT)


def rank_tile_range(ntiles: int, rank: int, nranks: int) -> Tuple[int, int]:
    """Contiguous split: [floor(r*T/R), floor((r+1)*T/R))."""
    return (rank * ntiles) // nranks, ((rank + 1) * ntiles) // nranks


def rank_row_range(H: int, W: int, rank: int, nranks: int) -> Tuple[int, int]:
    """Whole tile-rows per rank (so each rank's pixels are one contiguous [rows, W, 4] slab that
    all_gather can concatenate): rank r owns tile rows [floor(r*Ty/R), floor((r+1)*Ty/R))."""
    ty = tiles_y(H)
    r0, r1 = (rank * ty) // nranks, ((rank + 1) * ty) // nranks
    return min(r0 * TILE, H), min(r1 * TILE, H)


def interleaved_row_owner(tile_row: int, nranks: int) -> int:
    """Load-balanced alternative: tile row -> rank, round robin."""
    return tile_row % nranks


def rank_of_tile(tile: int, ntiles: int, nranks: int) -> int:
    """Inverse of :func:`rank_tile_range`."""
    r = min(nranks - 1, (tile * nranks) // max(ntiles, 1))
    while rank_tile_range(ntiles, r, nranks)[0] > tile:
        r -= 1
    while rank_tile_range(ntiles, r, nranks)[1] <= tile:
        r += 1
    return r


def all_rank_ranges(ntiles: int, nranks: int) -> List[Tuple[int, int]]:
    return [rank_tile_range(ntiles, r, nranks) for r in range(nranks)]


def interleaved_rows(n_tile_rows: int, rank: int, nranks: int):
    """Tile rows of ``rank`` in the interleaved image-space partition (``ty % nranks == rank``): the
    map `mrt_render_forward_batch_scatter(row_mod=nranks, row_rem=rank)` renders.  Interleaved
    because the object sits in the middle of the image: contiguous bands would leave the outer ranks
    with background only."""
    return range(rank, n_tile_rows, nranks)
