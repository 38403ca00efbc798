// tiles.h — the integer dispatch-geometry contract, shared by host and device code.
//
// Reference: `[numthreads(8,8,1)]` + `thread_count=[W,H,1]` + guard
// `any(tid.xy >= imageSize)` (inr/viewer/brats_rt.slang:86-89,
// inr/viewer/brats_viewer.py:431-432).  Everything here is integer-only and must stay
// bit-exact with mri_raytracer_b200/tiles.py (tests/test_tiles.py checks it exhaustively).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MRT_HD __host__ __device__ __forceinline__
#else
#define MRT_HD static inline
#endif

#define MRT_TILE_SHIFT 3
#define MRT_TILE_EDGE 8
#define MRT_TILE_MASK 7

MRT_HD int32_t mrt_tiles_x_(int32_t W) { return (W + MRT_TILE_MASK) >> MRT_TILE_SHIFT; }
MRT_HD int32_t mrt_tiles_y_(int32_t H) { return (H + MRT_TILE_MASK) >> MRT_TILE_SHIFT; }
MRT_HD int32_t mrt_tile_of_pixel_(int32_t x, int32_t y, int32_t W) {
  return (y >> MRT_TILE_SHIFT) * mrt_tiles_x_(W) + (x >> MRT_TILE_SHIFT);
}
MRT_HD int32_t mrt_lane_of_pixel_(int32_t x, int32_t y) {
  return ((y & MRT_TILE_MASK) << MRT_TILE_SHIFT) + (x & MRT_TILE_MASK);
}
// tile id + lane-in-tile (0..63) -> pixel
MRT_HD void mrt_pixel_of_tile_lane_(int32_t tile, int32_t lane, int32_t W, int32_t* x, int32_t* y) {
  const int32_t tx = tile % mrt_tiles_x_(W);
  const int32_t ty = tile / mrt_tiles_x_(W);
  *x = (tx << MRT_TILE_SHIFT) + (lane & MRT_TILE_MASK);
  *y = (ty << MRT_TILE_SHIFT) + (lane >> MRT_TILE_SHIFT);
}
// contiguous split of ntiles over nranks: [floor(r*T/R), floor((r+1)*T/R))
MRT_HD int32_t mrt_rank_tile_begin_(int32_t ntiles, int32_t rank, int32_t nranks) {
  return (int32_t)(((int64_t)rank * (int64_t)ntiles) / (int64_t)nranks);
}
