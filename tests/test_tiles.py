"""Integer tile map: Python (tiles.py) vs C (csrc/tiles.h through the C ABI), bit-exact,
exhaustive over W,H in 1..70 and 1/2/4/8 ranks (SURVEY.md §8(c) test 7)."""
import ctypes as C

from mri_raytracer_b200 import tiles


def test_tile_map_python_vs_c_exhaustive(built_lib):
    L = built_lib
    for W in range(1, 71):
        assert L.mrt_tiles_x(W) == tiles.tiles_x(W)
        for H in (1, 2, 7, 8, 9, 31, 64, 70):
            assert L.mrt_tile_count(W, H) == tiles.tile_count(W, H)
            seen = set()
            for y in range(H):
                for x in range(W):
                    t, l = tiles.tile_of_pixel(x, y, W), tiles.lane_of_pixel(x, y)
                    assert L.mrt_tile_of_pixel(x, y, W) == t
                    assert L.mrt_lane_of_pixel(x, y) == l
                    assert tiles.pixel_of_tile_lane(t, l, W) == (x, y)
                    assert 0 <= t < tiles.tile_count(W, H) and 0 <= l < 64
                    seen.add((t, l))
            assert len(seen) == W * H


def test_rank_ranges_partition(built_lib):
    L = built_lib
    b, e = C.c_int32(), C.c_int32()
    for nt in list(range(0, 70)) + [16384, 65536, 262144]:
        for R in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(R):
                L.mrt_rank_tile_range(nt, r, R, C.byref(b), C.byref(e))
                assert (b.value, e.value) == tiles.rank_tile_range(nt, r, R)
                assert b.value == prev and e.value >= b.value
                prev = e.value
            assert prev == nt
            for t in range(0, nt, max(1, nt // 50)):
                r = tiles.rank_of_tile(t, nt, R)
                lo, hi = tiles.rank_tile_range(nt, r, R)
                assert lo <= t < hi


def test_row_ranges_cover_image():
    for H in (1, 8, 9, 100, 1024, 2048):
        for R in (1, 2, 4, 8):
            rows = [tiles.rank_row_range(H, 64, r, R) for r in range(R)]
            assert rows[0][0] == 0 and rows[-1][1] == H
            for a, b in zip(rows, rows[1:]):
                assert a[1] == b[0]


def test_fast_tile_division_constant_is_exact():
    """The march kernel divides tile ids by tiles_x with a multiply-high when the host has proven it
    exact (c_api.cu derive(): tdiv_mul); same rule restated here and checked exhaustively for small
    images and on boundaries for large ones."""
    for W in list(range(1, 200)) + [512, 1000, 1024, 2048, 4096, 4099, 8192, 65536]:
        d = (W + 7) >> 3
        for H in (1, 7, 64, 1024, 4096, 65536):
            nt = d * ((H + 7) >> 3)
            if d > 1 and nt * d < 2 ** 32:
                m = 2 ** 32 // d + 1
                assert m < 2 ** 32
                probe = range(nt) if nt <= 4096 else [0, 1, d - 1, d, d + 1, nt // 2, nt - d, nt - 1]
                for n in probe:
                    assert (n * m) >> 32 == n // d, (W, H, n)
