// march.cuh — device-side building blocks shared by the forward and backward kernels.
//
// Follows inr/viewer/brats_rt.slang (reference) row by row; see DESIGN.md §kernels for
// the arithmetic contract.  Ray set-up uses explicit round-to-nearest intrinsics
// (no FMA contraction) so that o, d, t0, t1 and the per-ray sample count n are
// bit-identical to a one-IEEE-op-at-a-time CPU evaluation; the per-sample math is
// allowed to contract (tolerance 1e-4, see tests/).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "tiles.h"

#define MRT_BRICK_SHIFT 3

// Kernel-side constant block, derived on the host from MrtParams (c_api.cu).
struct KParams {
  int W, H;
  float eye[3], U[3], V[3], Wv[3];
  float focal;            // 1/tan(fovY/2), evaluated in double on the host
  int ortho; float halfH;
  float bmin[3], vs[3];
  int dims[3];            // X, Y, Z
  unsigned pitchY, pitchZ; // packed-layout pitches in voxels (kernels.h mrt_layout)
  float dt, nearT, farT;
  float inv_dt, inv_vs[3];  // 1/dt, 1/vs (rounded once on the host; only the non-bit-exact index-space math uses them)
  float bg[3];
  float wgt[4];           // volWeight if enabled else 0
  float inv_wsum;         // 1/wSum if wSum > 0 else 1
  float lo, inv_ww;       // window: val = saturate((v - lo) * inv_ww)
  float wq[4];            // folded per-channel weight wgt[c]*inv_wsum*inv_ww
  float wbias;            // -lo*inv_ww
  float neg_dt_log2e;     // -dt*log2(e)
  float ia, gamma;
  int showSeg, showPred;
  float lut[8][4];
  float thr;
  int maxSteps, tMode, alphaMode, tfMode, tfN;
  int skip;               // occupancy skipping enabled
  int nbx, nby, nbz;      // brick grid
  int tile_begin, tile_end;
  // sort-last shard (all zero when off): owned cell range [slo, shi) in GLOBAL voxel indices;
  // the packed buffer / brick grid / pitches are those of the sub-volume starting at slo.
  int half;               // voxel storage: 0 fp32, 1 fp16, 2 u8 (narrow types: single channel, widened to fp32 on load)
  int shard;
  int slo[3], shi[3];
  unsigned base_off;      // slo.x + slo.y*pitchY + slo.z*pitchZ, subtracted from global sample indices
  unsigned tdiv_mul;      // floor(2^32/tiles_x)+1 when tile/tiles_x == umulhi(tile, tdiv_mul) for every tile id, else 0
  unsigned idx_bias;      // base_off + 0x4b000000*(1 + pitchY + pitchZ) mod 2^32 (see mrt_sample_raw)
  // soft occupancy (docs/DifferentiableRendering.md section 11: "hard empty-space skipping -> continuous
  // occupancy o(x) in [0,1] learned and used multiplicatively"): one value per 8^3 brick, sigma' = o * sigma.
  // GENERIC variants only; nullptr = off.  docc: where the backward accumulates dL/do.
  const float* occ; float* docc;
};
__device__ __forceinline__ int mrt_brick_id(const KParams& P, int ix, int iy, int iz) {
  return ((iz >> MRT_BRICK_SHIFT) * P.nby + (iy >> MRT_BRICK_SHIFT)) * P.nbx + (ix >> MRT_BRICK_SHIFT);
}

// Cameras of a batch of views rendered by ONE launch (blockIdx.y = view): everything else in
// KParams is shared.  12 floats per view: eye, U, V, W.
#define MRT_MAX_VIEWS 64
struct CamBatch { float cam[MRT_MAX_VIEWS][12]; };

// Sort-last exchange fused into the march: image row y belongs to strip y / rows, whose pixels go
// to base[strip] (a peer-mapped buffer of the strip's owner rank) instead of the local image.
#define MRT_MAX_STRIPS 16
// spans (optional, sparse framebuffer gather): per view of the launch and per tile row an int2
// (x0, x1), the inclusive pixel span outside which every ray of that row band certainly misses every
// active brick (mrt_view_spans_kernel); tiles_y entries per view.  Tiles outside are NOT stored by the
// march unless store_outside is set: the owner of the image fills them with the background itself
// (mrt_fill_outside_spans), from the same spans.  With store_outside the spans only serve as the
// (much cheaper) replacement of the per-ray box test: one LDG + two compares per warp.
// view_base (optional, image-space scatter): a DEVICE array with one image pointer per view of the
// launch — each view's [H][W] frame may live on a different GPU (peer-mapped) — replacing the
// contiguous out_rgba.  row_mod > 1 (image-space tile partition): the launch renders only the tile
// rows ty with ty % row_mod == row_rem, interleaved over the ranks because the object sits mid-image.
struct StripTargets {
  float4* base[MRT_MAX_STRIPS]; int n, rows; const int2* spans; int store_outside;
  float4* const* view_base; int row_mod, row_rem;
};

// Outputs of the checkpointing (training) forward, consumed by the segment-parallel backward:
// ck[(c-1)][view][H][W] = (C, T) of the ray before slot c*S (1 <= c < nseg), k_end[view][H][W] = the
// ray's end slot (n_taken), warp_kmax[view][2*tiles] = the largest k_end of each half tile.
struct CkptOut { float4* ck; int S, nseg; int32_t* k_end; int32_t* warp_kmax; };

struct Ray {
  float ox, oy, oz, dx, dy, dz;
  float t0, t1;
  int n;        // # of sample slots k with t0 + k*dt < t1 (0 if miss)
};

__device__ __forceinline__ float mrt_norm3(float x, float y, float z) {
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// makePrimary (brats_rt.slang:36-46) / orthographic, then de-zero + aabbHit +
// near/far clamp (:95-99, :48-57, :107-109) and the indexed sample count.
// `cam` = 12 floats (eye, U, V, W): KParams::eye for a single frame, one row of CamBatch for a
// batch of views.
__device__ __forceinline__ Ray mrt_setup_ray(const KParams& P, const float* __restrict__ cam, int px, int py) {
  Ray r;
  const float* eye = cam; const float* U = cam + 3; const float* V = cam + 6; const float* Wv = cam + 9;
  const float dimx = (float)P.W, dimy = (float)P.H;
  const float ndcx = __fdiv_rn(__fadd_rn((float)px, 0.5f), dimx);
  const float ndcy = __fdiv_rn(__fadd_rn((float)py, 0.5f), dimy);
  const float uvx = __fsub_rn(__fmul_rn(ndcx, 2.0f), 1.0f);
  const float uvy = __fsub_rn(__fmul_rn(ndcy, 2.0f), 1.0f);
  const float aspect = __fdiv_rn(dimx, fmaxf(1.0f, dimy));
  if (P.ortho) {
    const float halfW = __fmul_rn(aspect, P.halfH);
    const float ax = __fmul_rn(uvx, halfW);
    const float ay = -__fmul_rn(uvy, P.halfH);
    r.ox = __fadd_rn(__fadd_rn(eye[0], __fmul_rn(ax, U[0])), __fmul_rn(ay, V[0]));
    r.oy = __fadd_rn(__fadd_rn(eye[1], __fmul_rn(ax, U[1])), __fmul_rn(ay, V[1]));
    r.oz = __fadd_rn(__fadd_rn(eye[2], __fmul_rn(ax, U[2])), __fmul_rn(ay, V[2]));
    r.dx = Wv[0]; r.dy = Wv[1]; r.dz = Wv[2];
  } else {
    float cx = __fdiv_rn(__fmul_rn(uvx, aspect), P.focal);
    float cy = __fdiv_rn(-uvy, P.focal);
    float cz = 1.0f;
    const float inv = mrt_norm3(cx, cy, cz);
    cx = __fdiv_rn(cx, inv); cy = __fdiv_rn(cy, inv); cz = __fdiv_rn(cz, inv);
    const float rx = __fadd_rn(__fadd_rn(__fmul_rn(cx, U[0]), __fmul_rn(cy, V[0])), __fmul_rn(cz, Wv[0]));
    const float ry = __fadd_rn(__fadd_rn(__fmul_rn(cx, U[1]), __fmul_rn(cy, V[1])), __fmul_rn(cz, Wv[1]));
    const float rz = __fadd_rn(__fadd_rn(__fmul_rn(cx, U[2]), __fmul_rn(cy, V[2])), __fmul_rn(cz, Wv[2]));
    const float n = mrt_norm3(rx, ry, rz);
    r.dx = __fdiv_rn(rx, n); r.dy = __fdiv_rn(ry, n); r.dz = __fdiv_rn(rz, n);
    r.ox = eye[0]; r.oy = eye[1]; r.oz = eye[2];
  }
  // de-zero: sign dropped on purpose (:96-98); rcpDir uses the patched copy only
  const float ex = fabsf(r.dx) < 1e-6f ? 1e-6f : r.dx;
  const float ey = fabsf(r.dy) < 1e-6f ? 1e-6f : r.dy;
  const float ez = fabsf(r.dz) < 1e-6f ? 1e-6f : r.dz;
  const float rx_ = __fdiv_rn(1.0f, ex), ry_ = __fdiv_rn(1.0f, ey), rz_ = __fdiv_rn(1.0f, ez);
  const float bxm = __fadd_rn(P.bmin[0], __fmul_rn(P.vs[0], (float)P.dims[0]));
  const float bym = __fadd_rn(P.bmin[1], __fmul_rn(P.vs[1], (float)P.dims[1]));
  const float bzm = __fadd_rn(P.bmin[2], __fmul_rn(P.vs[2], (float)P.dims[2]));
  const float ax0 = __fmul_rn(__fsub_rn(P.bmin[0], r.ox), rx_), ax1 = __fmul_rn(__fsub_rn(bxm, r.ox), rx_);
  const float ay0 = __fmul_rn(__fsub_rn(P.bmin[1], r.oy), ry_), ay1 = __fmul_rn(__fsub_rn(bym, r.oy), ry_);
  const float az0 = __fmul_rn(__fsub_rn(P.bmin[2], r.oz), rz_), az1 = __fmul_rn(__fsub_rn(bzm, r.oz), rz_);
  const float tmin = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fminf(az0, az1));
  const float tmax = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fmaxf(az0, az1));
  bool hit = tmax >= fmaxf(tmin, 0.0f);
  r.t0 = fmaxf(tmin, fmaxf(0.0f, P.nearT));
  r.t1 = (P.farT > 0.0f) ? fminf(tmax, P.farT) : tmax;
  hit = hit && !(r.t1 <= r.t0);
  int n = 0;
  if (hit) {
    n = (int)ceilf(__fdiv_rn(__fsub_rn(r.t1, r.t0), P.dt));
    n = max(n, 0);
#pragma unroll 1
    for (int it = 0; it < 3; ++it) {   // exact fix-up: n = #{k : t0 + k*dt < t1}
      if (__fadd_rn(r.t0, __fmul_rn((float)n, P.dt)) < r.t1) ++n;
      if (n > 0 && !(__fadd_rn(r.t0, __fmul_rn((float)(n - 1), P.dt)) < r.t1)) --n;
    }
    if (P.maxSteps > 0) n = min(n, P.maxSteps);
  }
  r.n = n;
  return r;
}

// ---- active-brick box ------------------------------------------------------------------
// Bounding box of the active bricks (tail of the skip-level buffer, written by the classify
// kernel) in GLOBAL index space, widened by MRT_BOX_MARGIN voxels.  Every sample slot that can
// contribute has its position inside it, so (i) a ray whose line never enters the box is
// background, whatever its exact set-up says, and (ii) slots outside the ray's box interval are
// no-ops.  The margin (1/4 voxel) is ~300x the worst error of the approximate ray below.
#define MRT_BOX_MARGIN 0.25f
struct ActiveBox { float lo[3], hi[3]; };
__device__ __forceinline__ float mrt_rcp(float x) {        // MUFU.RCP, 1 ulp; rcp(0) = +inf
  float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ ActiveBox mrt_active_box(const KParams& P, const uint8_t* __restrict__ levels) {
  const size_t nb = (size_t)P.nbx * P.nby * P.nbz;
  const int* b = reinterpret_cast<const int*>(levels + ((nb + 15) & ~(size_t)15));
  ActiveBox A;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int mlo = __ldg(b + a), mhi = __ldg(b + 3 + a);       // (max(-lo), max(hi)) in bricks; INT_MIN-ish = none
    const bool none = mhi < 0;
    A.lo[a] = none ? 3.0e38f : (float)(((-mlo) << MRT_BRICK_SHIFT) + P.slo[a]) - MRT_BOX_MARGIN;
    A.hi[a] = none ? -3.0e38f : (float)(((mhi + 1) << MRT_BRICK_SHIFT) + P.slo[a]) + MRT_BOX_MARGIN;
  }
  return A;
}
// [tin, tout] of the line o + t*d (index space) inside the box; empty when tout < max(tin, 0)
__device__ __forceinline__ void mrt_box_interval(const ActiveBox& A, float ox, float oy, float oz,
                                                 float dx, float dy, float dz, float* tin, float* tout) {
  const float ix = mrt_rcp(dx), iy = mrt_rcp(dy), iz = mrt_rcp(dz);   // +-inf for a zero component: standard slab rule
  const float ax = (A.lo[0] - ox) * ix, bx = (A.hi[0] - ox) * ix;
  const float ay = (A.lo[1] - oy) * iy, by = (A.hi[1] - oy) * iy;
  const float az = (A.lo[2] - oz) * iz, bz = (A.hi[2] - oz) * iz;
  *tin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));        // NaN (0*inf) drops out of fmin/fmax: conservative
  *tout = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
}
// Cheap conservative test with an APPROXIMATE, un-normalised ray (no IEEE divisions, no sqrt):
// false only if the pixel's ray certainly misses the active box.
__device__ __forceinline__ bool mrt_ray_may_hit(const KParams& P, const float* __restrict__ cam, int px, int py,
                                                const ActiveBox& A) {
  const float* eye = cam; const float* U = cam + 3; const float* V = cam + 6; const float* Wv = cam + 9;
  const float uvx = ((float)px + 0.5f) * (2.0f * mrt_rcp((float)P.W)) - 1.0f;
  const float uvy = ((float)py + 0.5f) * (2.0f * mrt_rcp((float)P.H)) - 1.0f;
  const float aspect = (float)P.W * mrt_rcp(fmaxf(1.0f, (float)P.H));
  float ox, oy, oz, dx, dy, dz;
  if (P.ortho) {
    const float ax = uvx * aspect * P.halfH, ay = -uvy * P.halfH;
    ox = eye[0] + ax * U[0] + ay * V[0]; oy = eye[1] + ax * U[1] + ay * V[1]; oz = eye[2] + ax * U[2] + ay * V[2];
    dx = Wv[0]; dy = Wv[1]; dz = Wv[2];
  } else {
    const float inv_f = mrt_rcp(P.focal);
    const float cx = uvx * aspect * inv_f, cy = -uvy * inv_f;
    dx = cx * U[0] + cy * V[0] + Wv[0]; dy = cx * U[1] + cy * V[1] + Wv[1]; dz = cx * U[2] + cy * V[2] + Wv[2];
    ox = eye[0]; oy = eye[1]; oz = eye[2];
  }
  const float sx = mrt_rcp(P.vs[0]), sy = mrt_rcp(P.vs[1]), sz = mrt_rcp(P.vs[2]);
  float tin, tout;
  mrt_box_interval(A, (ox - P.bmin[0]) * sx, (oy - P.bmin[1]) * sy, (oz - P.bmin[2]) * sz, dx * sx, dy * sy, dz * sz,
                   &tin, &tout);
  return tout >= fmaxf(tin, 0.0f);
}

// Screen footprint of a box of voxels for one camera: its 8 corners, widened by MRT_SPAN_MARGIN voxels
// (> MRT_BOX_MARGIN, so a ray the slab test lets through always lies inside), in pixel coordinates.  A view's
// spans — one x-extent per 8-pixel row band (tile row), empty: x0 > x1 — are the union over its active bricks of
// these footprints' bounding rectangles, rounded outward by one pixel (forward.cu: mrt_view_spans_kernel).
#define MRT_SPAN_MARGIN 0.75f
// Projects the 8 corners of box A (widened to MRT_SPAN_MARGIN) to pixel coordinates.  Returns 0 = done,
// 1 = no culling possible (degenerate basis, or the box reaches behind the eye), 2 = empty box.
__device__ __forceinline__ int mrt_project_box(const KParams& P, const float* __restrict__ cam, const ActiveBox& A,
                                               float cx[8], float cy[8]) {
  const float* eye = cam; const float* U = cam + 3; const float* V = cam + 6; const float* Wv = cam + 9;
  if (A.hi[0] < A.lo[0]) return 2;                                         // no active brick
  // camera coordinates of a world offset w: (xc, yc, zc) = M^-1 w with M = [U V W] — rows (VxW, WxU,
  // UxV)/det, which is M^T for the orthonormal bases the cameras produce but stays correct for any
  // basis the C ABI lets through (the exact ray set-up never assumes orthonormality either)
  const float r0x = V[1] * Wv[2] - V[2] * Wv[1], r0y = V[2] * Wv[0] - V[0] * Wv[2], r0z = V[0] * Wv[1] - V[1] * Wv[0];
  const float r1x = Wv[1] * U[2] - Wv[2] * U[1], r1y = Wv[2] * U[0] - Wv[0] * U[2], r1z = Wv[0] * U[1] - Wv[1] * U[0];
  const float r2x = U[1] * V[2] - U[2] * V[1], r2y = U[2] * V[0] - U[0] * V[2], r2z = U[0] * V[1] - U[1] * V[0];
  const float det = U[0] * r0x + U[1] * r0y + U[2] * r0z;
  if (!(fabsf(det) > 1e-12f)) return 1;                                    // degenerate basis: no culling
  const float idet = 1.0f / det;
  const float m = MRT_SPAN_MARGIN - MRT_BOX_MARGIN;
  const float aspect = (float)P.W / fmaxf(1.0f, (float)P.H);
  bool behind = false;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float ix = (c & 1) ? A.hi[0] + m : A.lo[0] - m;
    const float iy = (c & 2) ? A.hi[1] + m : A.lo[1] - m;
    const float iz = (c & 4) ? A.hi[2] + m : A.lo[2] - m;
    const float wx = P.bmin[0] + ix * P.vs[0] - eye[0], wy = P.bmin[1] + iy * P.vs[1] - eye[1], wz = P.bmin[2] + iz * P.vs[2] - eye[2];
    const float xc = (wx * r0x + wy * r0y + wz * r0z) * idet;
    const float yc = (wx * r1x + wy * r1y + wz * r1z) * idet;
    const float zc = (wx * r2x + wy * r2y + wz * r2z) * idet;
    float uvx, uvy;
    if (P.ortho) {
      uvx = xc / (aspect * P.halfH); uvy = -yc / P.halfH;
    } else {
      if (!(zc > 1e-4f)) { behind = true; uvx = uvy = 0.0f; }
      else { uvx = (xc / zc) * P.focal / aspect; uvy = -(yc / zc) * P.focal; }
    }
    cx[c] = (uvx + 1.0f) * 0.5f * (float)P.W - 0.5f; cy[c] = (uvy + 1.0f) * 0.5f * (float)P.H - 0.5f;
  }
  return behind ? 1 : 0;                                                   // the box reaches behind the eye: everything
}
// does the 8x8 tile with left pixel column tx0 intersect its row band's span?
__device__ __forceinline__ bool mrt_tile_in_span(int2 sp, int tx0) {
  return tx0 <= sp.y && tx0 + (MRT_TILE_EDGE - 1) >= sp.x;
}

// ---- voxel vector types -------------------------------------------------------------
template <int NCH> struct Vox;
template <> struct Vox<1> { typedef float T; };
template <> struct Vox<2> { typedef float2 T; };
template <> struct Vox<4> { typedef float4 T; };
// element type the sampler loads (HALF = the storage type: 0 fp32 Vox<NCH>, 1 __half, 2 uint8_t; the
// narrow types are single-channel).  u8 voxels are widened to their INTEGER value: the /255 of
// volume_render.slang:38 is linear, so it is folded into KParams::wq on the host.
template <int NCH, int HALF> struct VoxT { typedef typename Vox<NCH>::T T; };
template <> struct VoxT<1, 1> { typedef __half T; };
template <> struct VoxT<1, 2> { typedef uint8_t T; };
// HALF = 3, the "quad" layout of a single-channel fp32 volume: element (x,y,z) holds the four voxels
// (x,y) (x+1,y) (x,y+1) (x+1,y+1) of slice z, so a trilinear footprint is TWO 16-byte loads (slices z
// and z+1) instead of eight 4-byte ones: a quarter of the load instructions and about half the L1
// data-stage wavefronts of the march, for 4x the bytes (mrt_pack_volume_quad).  Forward only.
template <> struct VoxT<1, 3> { typedef float4 T; };
// HALF = 4: the same quad layout over fp16 voxels (4 x __half = 8 B per element; mrt_pack_volume_quad_f16).
template <> struct VoxT<1, 4> { typedef uint2 T; };

__device__ __forceinline__ float lerpf(float a, float b, float t) { return fmaf(t, b - a, a); }

// Folded modality blend + window/level of ONE voxel (all four are linear maps, so they commute
// with the trilinear interpolation; doing them per corner turns 7 vector lerps into 7 scalar
// ones):  sum_c wq[c]*s_c  with  wq[c] = volWeight[c]/wSum/ww  (0 for disabled channels).
// The constant -(wl - ww/2)/ww is added once after interpolation (weights sum to 1).
__device__ __forceinline__ float foldv(float s, const KParams& P) { return s * P.wq[0]; }
__device__ __forceinline__ float mrt_f32(float v) { return v; }
__device__ __forceinline__ float2 mrt_f32(float2 v) { return v; }
__device__ __forceinline__ float4 mrt_f32(float4 v) { return v; }
__device__ __forceinline__ float mrt_f32(__half v) { return __half2float(v); }
__device__ __forceinline__ float mrt_f32(uint8_t v) { return (float)v; }
__device__ __forceinline__ float mrt_scalar(float v) { return v; }
__device__ __forceinline__ float mrt_scalar(__half v) { return __half2float(v); }
__device__ __forceinline__ float mrt_scalar(uint8_t v) { return (float)v; }
__device__ __forceinline__ float mrt_scalar(float2 v) { return v.x; }   // never used (NCH == 1 only)
__device__ __forceinline__ float mrt_scalar(float4 v) { return v.x; }
__device__ __forceinline__ float foldv(float2 s, const KParams& P) { return fmaf(s.y, P.wq[1], s.x * P.wq[0]); }
__device__ __forceinline__ float foldv(float4 s, const KParams& P) {
  return fmaf(s.w, P.wq[3], fmaf(s.z, P.wq[2], fmaf(s.y, P.wq[1], s.x * P.wq[0])));
}

// Index-space ray: pIdx(t) = oi + t*di  (== (o + t*d - bmin)/voxelSize up to rounding).
struct IdxRay { float ox, oy, oz, dx, dy, dz; };
__device__ __forceinline__ IdxRay mrt_index_ray(const KParams& P, const Ray& r) {
  IdxRay q;
  q.ox = (r.ox - P.bmin[0]) * P.inv_vs[0]; q.oy = (r.oy - P.bmin[1]) * P.inv_vs[1]; q.oz = (r.oz - P.bmin[2]) * P.inv_vs[2];
  q.dx = r.dx * P.inv_vs[0]; q.dy = r.dy * P.inv_vs[1]; q.dz = r.dz * P.inv_vs[2];
  return q;
}

// sampleLinear (brats_rt.slang:60-76) on the packed multi-channel layout, then the
// modality blend (:123-130).  Returns v; also the integer base + fractions if wanted.
// mx/my/mz hold the raw bits of (q + 2^23) rounded toward -inf, i.e. 0x4b000000 + floor(q).
#define MRT_MAGIC_BITS 0x4b000000u
struct Cell {
  uint32_t mx, my, mz; float fx, fy, fz;
  __device__ __forceinline__ int ix() const { return (int)(mx - MRT_MAGIC_BITS); }
  __device__ __forceinline__ int iy() const { return (int)(my - MRT_MAGIC_BITS); }
  __device__ __forceinline__ int iz() const { return (int)(mz - MRT_MAGIC_BITS); }
};

// CLAMP = false: the caller has proven 0 <= p <= hi on every axis (the clamps are identities there)
template <bool CLAMP>
__device__ __forceinline__ Cell mrt_cell_t(float px, float py, float pz, float hix, float hiy, float hiz) {
  Cell c;
  const float qx = CLAMP ? fminf(fmaxf(px, 0.0f), hix) : px;
  const float qy = CLAMP ? fminf(fmaxf(py, 0.0f), hiy) : py;
  const float qz = CLAMP ? fminf(fmaxf(pz, 0.0f), hiz) : pz;
  // floor of a value in [0, 2^22) without the conversion (XU) pipe: q + 2^23 rounded toward
  // -inf is exactly 2^23 + floor(q); its low mantissa bits are the integer.  Same result as
  // floorf / (int), three FADD/IADD instead of F2I + FRND per axis.
  const float mx = __fadd_rd(qx, 8388608.0f), my = __fadd_rd(qy, 8388608.0f), mz = __fadd_rd(qz, 8388608.0f);
  c.mx = __float_as_uint(mx); c.my = __float_as_uint(my); c.mz = __float_as_uint(mz);
  c.fx = qx - (mx - 8388608.0f); c.fy = qy - (my - 8388608.0f); c.fz = qz - (mz - 8388608.0f);
  return c;
}
__device__ __forceinline__ Cell mrt_cell(const KParams& P, float px, float py, float pz,
                                         float hix, float hiy, float hiz) {
  return mrt_cell_t<true>(px, py, pz, hix, hiy, hiz);
}

// sampleLinear (brats_rt.slang:60-76; lerp order x, y, z) of the folded scalar field, i.e.
// raw = ((blend of the <=4 modalities, :123-130) - (wl - ww/2)) / ww  (:132) before saturate.
// The 8 corners of a cell, fetched (mrt_fetch) separately from their interpolation (mrt_interp) so
// that a kernel can issue the loads of the NEXT slot before the dependent arithmetic of this one.
template <int NCH, int HALF> struct CornerT { typedef typename VoxT<NCH, HALF>::T T; };
template <> struct CornerT<1, 3> { typedef float T; };
template <> struct CornerT<1, 4> { typedef float T; };
template <int NCH, int HALF> struct Corners { typename CornerT<NCH, HALF>::T v[8]; };

template <int NCH, int HALF> struct FetchImpl {
  static __device__ __forceinline__ Corners<NCH, HALF> run(const KParams& P, const typename VoxT<NCH, HALF>::T* __restrict__ vol,
                                                           const Cell& c) {
    typedef typename VoxT<NCH, HALF>::T VT;
    // element index straight from the magic-number bits: the three -0x4b000000 corrections and the
    // shard offset are one precomputed constant (uint32 wrap-around is exact)
    const uint32_t b = c.mx + c.my * P.pitchY + c.mz * P.pitchZ - P.idx_bias;
    const char* q0 = reinterpret_cast<const char*>(vol) + (size_t)b * sizeof(VT);
    const VT* p0 = reinterpret_cast<const VT*>(q0);
    const VT* p1 = reinterpret_cast<const VT*>(q0 + (size_t)P.pitchY * sizeof(VT));
    const VT* p2 = reinterpret_cast<const VT*>(q0 + (size_t)P.pitchZ * sizeof(VT));
    const VT* p3 = reinterpret_cast<const VT*>(q0 + ((size_t)P.pitchY + (size_t)P.pitchZ) * sizeof(VT));
    Corners<NCH, HALF> k;
    k.v[0] = __ldg(p0); k.v[1] = __ldg(p0 + 1);
    k.v[2] = __ldg(p1); k.v[3] = __ldg(p1 + 1);
    k.v[4] = __ldg(p2); k.v[5] = __ldg(p2 + 1);
    k.v[6] = __ldg(p3); k.v[7] = __ldg(p3 + 1);
    return k;
  }
};
template <> struct FetchImpl<1, 3> {
  static __device__ __forceinline__ Corners<1, 3> run(const KParams& P, const float4* __restrict__ vol, const Cell& c) {
    const uint32_t b = c.mx + c.my * P.pitchY + c.mz * P.pitchZ - P.idx_bias;
    const float4* p0 = vol + b;
    const float4 lo = __ldg(p0), hi = __ldg(p0 + P.pitchZ);
    Corners<1, 3> k;
    k.v[0] = lo.x; k.v[1] = lo.y; k.v[2] = lo.z; k.v[3] = lo.w;
    k.v[4] = hi.x; k.v[5] = hi.y; k.v[6] = hi.z; k.v[7] = hi.w;
    return k;
  }
};
template <> struct FetchImpl<1, 4> {
  static __device__ __forceinline__ Corners<1, 4> run(const KParams& P, const uint2* __restrict__ vol, const Cell& c) {
    const uint32_t b = c.mx + c.my * P.pitchY + c.mz * P.pitchZ - P.idx_bias;
    const uint2* p0 = vol + b;
    const uint2 lo = __ldg(p0), hi = __ldg(p0 + P.pitchZ);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo.x)), bb = __half22float2(*reinterpret_cast<const __half2*>(&lo.y));
    const float2 cc = __half22float2(*reinterpret_cast<const __half2*>(&hi.x)), d = __half22float2(*reinterpret_cast<const __half2*>(&hi.y));
    Corners<1, 4> k;
    k.v[0] = a.x; k.v[1] = a.y; k.v[2] = bb.x; k.v[3] = bb.y;
    k.v[4] = cc.x; k.v[5] = cc.y; k.v[6] = d.x; k.v[7] = d.y;
    return k;
  }
};
template <int NCH, int HALF = 0>
__device__ __forceinline__ Corners<NCH, HALF> mrt_fetch(const KParams& P, const typename VoxT<NCH, HALF>::T* __restrict__ vol,
                                                       const Cell& c) {
  return FetchImpl<NCH, HALF>::run(P, vol, c);
}

template <int NCH, int HALF = 0>
__device__ __forceinline__ float mrt_interp(const KParams& P, const Corners<NCH, HALF>& k, const Cell& c) {
  if (NCH == 1) {
    // one modality: the (linear) window scale is applied once, after the interpolation
    const float f0 = mrt_scalar(k.v[0]), f1 = mrt_scalar(k.v[1]), f2 = mrt_scalar(k.v[2]), f3 = mrt_scalar(k.v[3]);
    const float f4 = mrt_scalar(k.v[4]), f5 = mrt_scalar(k.v[5]), f6 = mrt_scalar(k.v[6]), f7 = mrt_scalar(k.v[7]);
    const float s = lerpf(lerpf(lerpf(f0, f1, c.fx), lerpf(f2, f3, c.fx), c.fy),
                          lerpf(lerpf(f4, f5, c.fx), lerpf(f6, f7, c.fx), c.fy), c.fz);
    return fmaf(s, P.wq[0], P.wbias);
  } else {
    const float c000 = foldv(mrt_f32(k.v[0]), P), c100 = foldv(mrt_f32(k.v[1]), P), c010 = foldv(mrt_f32(k.v[2]), P), c110 = foldv(mrt_f32(k.v[3]), P);
    const float c001 = foldv(mrt_f32(k.v[4]), P), c101 = foldv(mrt_f32(k.v[5]), P), c011 = foldv(mrt_f32(k.v[6]), P), c111 = foldv(mrt_f32(k.v[7]), P);
    const float s = lerpf(lerpf(lerpf(c000, c100, c.fx), lerpf(c010, c110, c.fx), c.fy),
                          lerpf(lerpf(c001, c101, c.fx), lerpf(c011, c111, c.fx), c.fy), c.fz);
    return s + P.wbias;
  }
}

// d(raw)/d(index-space position) of the same interpolant (docs/DifferentiableRendering.md section 6,
// :116-127: ds/dx = sum_n v_n dw_n/dx), per axis the lerp of the four corner differences.
template <int NCH, int HALF = 0>
__device__ __forceinline__ void mrt_interp_grad(const KParams& P, const Corners<NCH, HALF>& k, const Cell& c,
                                                float* gx, float* gy, float* gz) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (NCH == 1) ? mrt_scalar(k.v[i]) * P.wq[0] : foldv(mrt_f32(k.v[i]), P);
  // corner order: v[0..7] = c000, c100, c010, c110, c001, c101, c011, c111
  *gx = lerpf(lerpf(v[1] - v[0], v[3] - v[2], c.fy), lerpf(v[5] - v[4], v[7] - v[6], c.fy), c.fz);
  *gy = lerpf(lerpf(v[2] - v[0], v[3] - v[1], c.fx), lerpf(v[6] - v[4], v[7] - v[5], c.fx), c.fz);
  *gz = lerpf(lerpf(v[4] - v[0], v[5] - v[1], c.fx), lerpf(v[6] - v[2], v[7] - v[3], c.fx), c.fy);
}

// sampleLinear (brats_rt.slang:60-76; lerp order x, y, z) of the folded scalar field, i.e.
// raw = ((blend of the <=4 modalities, :123-130) - (wl - ww/2)) / ww  (:132) before saturate.
template <int NCH, int HALF = 0>
__device__ __forceinline__ float mrt_sample_raw(const KParams& P, const typename VoxT<NCH, HALF>::T* __restrict__ vol,
                                                const Cell& c) {
  return mrt_interp<NCH, HALF>(P, mrt_fetch<NCH, HALF>(P, vol, c), c);
}

// sampleLabel (brats_rt.slang:78-83): round half away from zero.
__device__ __forceinline__ int mrt_sample_label(const KParams& P, const int32_t* __restrict__ lab,
                                                float px, float py, float pz) {
  const float qx = fminf(fmaxf(px, 0.0f), (float)P.dims[0] - 1.0f);
  const float qy = fminf(fmaxf(py, 0.0f), (float)P.dims[1] - 1.0f);
  const float qz = fminf(fmaxf(pz, 0.0f), (float)P.dims[2] - 1.0f);
  const uint32_t ix = (uint32_t)roundf(qx), iy = (uint32_t)roundf(qy), iz = (uint32_t)roundf(qz);
  const uint32_t idx = ix + iy * (uint32_t)P.dims[0] + iz * (uint32_t)P.dims[0] * (uint32_t)P.dims[1];
  return __ldg(lab + idx);
}

// saturate (+gamma) (brats_rt.slang:132-133)
template <bool GENERIC>
__device__ __forceinline__ float mrt_window(const KParams& P, float raw) {
  float val = __saturatef(raw);
  if (GENERIC) { if (P.gamma != 1.0f) val = powf(val, P.gamma); }
  return val;
}

// alpha = 1 - exp(-sigma*dt)  (:137) as 1 - 2^(sigma * (-dt*log2 e)): one FMUL + MUFU.EX2.
// ex2.approx is accurate to 2 ulp of a result in (0,1], i.e. ~1.2e-7 absolute on alpha.
__device__ __forceinline__ float mrt_ex2(float x) {
  float y;                                  // x <= 0 here; a denormal result flushes to 0: alpha = 1 either way
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mrt_alpha(const KParams& P, float sigma) {
  return 1.0f - mrt_ex2(sigma * P.neg_dt_log2e);
}

// Shared-memory LUT: entry j holds tf[j] and the forward difference tf[min(j+1,N-1)] - tf[j],
// so the lerp of SURVEY §8 A7  (u = val*(N-1); lerp(tf[floor u], tf[min(floor u+1,N-1)], frac))
// is one fma per component with the SAME rounding as a + t*(b-a).
struct TfEntry { float4 base, delta; };

__device__ __forceinline__ void mrt_tf_stage(TfEntry* __restrict__ s_tf, const float4* __restrict__ tf, int N) {
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float4 a = __ldg(tf + i), b = __ldg(tf + min(i + 1, N - 1));
    s_tf[i].base = a;
    s_tf[i].delta = make_float4(b.x - a.x, b.y - a.y, b.z - a.z, b.w - a.w);
  }
}

__device__ __forceinline__ float4 mrt_lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
// `s_tf` as a shared-window address (__cvta_generic_to_shared, once per thread): keeps the
// generic->shared conversion out of the sample loop.
__device__ __forceinline__ float4 mrt_tf_lookup(uint32_t s_tf, float nm1, float val,
                                                int* j0o = nullptr, float* fro = nullptr) {
  const float u = val * nm1;                 // val in [0,1] => 0 <= floor(u) <= N-1
  const float mu = __fadd_rd(u, 8388608.0f);  // floor without the conversion pipe (see mrt_cell)
  const int j0 = __float_as_int(mu) - 0x4b000000;
  const float fr = u - (mu - 8388608.0f);
  const uint32_t ea = s_tf + (uint32_t)j0 * (uint32_t)sizeof(TfEntry);
  const float4 a = mrt_lds128(ea), d = mrt_lds128(ea + 16);
  if (j0o) { *j0o = j0; *fro = fr; }
  return make_float4(fmaf(fr, d.x, a.x), fmaf(fr, d.y, a.y), fmaf(fr, d.z, a.z), fmaf(fr, d.w, a.w));
}

// The march's copy of the look-up: `s_tf_adj` = shared-window address of the LUT minus
// (0x4b000000 << 5) mod 2^32, so that the entry address is ONE shift-add of the magic-number bits of
// floor(u), and both 16-byte halves of the entry come from one address register.
#define MRT_TF_ADJ 0x60000000u
__device__ __forceinline__ float4 mrt_tf_lookup_adj(uint32_t s_tf_adj, float nm1, float val) {
  static_assert(sizeof(TfEntry) == 32, "entry address = index << 5");
  const float u = val * nm1;
  const float mu = __fadd_rd(u, 8388608.0f);
  const float fr = u - (mu - 8388608.0f);
  const uint32_t ea = s_tf_adj + (__float_as_uint(mu) << 5);
  float4 a, d;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%8];\n\tld.shared.v4.f32 {%4,%5,%6,%7}, [%8+16];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "r"(ea));
  return make_float4(fmaf(fr, d.x, a.x), fmaf(fr, d.y, a.y), fmaf(fr, d.z, a.z), fmaf(fr, d.w, a.w));
}

// The march's LUT in shared memory with a 48-byte entry stride (base at +0, delta at +16, 16 B pad):
// entry j starts in 16-byte bank group 3j mod 8, so eight consecutive entries fall into eight
// different groups (32-byte entries only ever use every other group: ncu, shared wavefronts per
// look-up) — the bank spread of a structure-of-arrays layout while the delta stays at an immediate
// offset of the base.  Falls back to 32 bytes for LUTs whose padded copy would cost occupancy.
// `adj` = base address - 0x4b000000 * stride (mod 2^32), so the entry address is ONE multiply-add of
// the magic-number bits of floor(u).
__device__ __forceinline__ uint32_t mrt_tf_stride(int N) { return N <= 512 ? 48u : 32u; }
__device__ __forceinline__ void mrt_tf_stage_strided(unsigned char* __restrict__ s, uint32_t stride,
                                                     const float4* __restrict__ tf, int N) {
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float4 a = __ldg(tf + i), b = __ldg(tf + min(i + 1, N - 1));
    float4* e = reinterpret_cast<float4*>(s + (size_t)i * stride);
    e[0] = a;
    e[1] = make_float4(b.x - a.x, b.y - a.y, b.z - a.z, b.w - a.w);
  }
}
__device__ __forceinline__ float4 mrt_tf_lookup_strided(uint32_t adj, uint32_t stride, float nm1, float val) {
  const float u = val * nm1;
  const float mu = __fadd_rd(u, 8388608.0f);
  const float fr = u - (mu - 8388608.0f);
  const uint32_t ea = __float_as_uint(mu) * stride + adj;
  float4 a, d;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%8];\n\tld.shared.v4.f32 {%4,%5,%6,%7}, [%8+16];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "r"(ea));
  return make_float4(fmaf(fr, d.x, a.x), fmaf(fr, d.y, a.y), fmaf(fr, d.z, a.z), fmaf(fr, d.w, a.w));
}

// CTA index -> position in the launch's tile range, "middle-out": CTA 0 takes the middle of the
// range and successive CTAs alternate outward.  The tile ids of a frame are row-major, so the
// rows through the image centre — where the camera frames the volume and rays are longest —
// are scheduled first and the cheap border rows last: the tail of the grid is filled with
// short CTAs instead of leaving SMs idle behind a few long ones (ncu: SM-active 61 % -> see
// profiles/).  Pure integer map, a bijection of [0, n).
__device__ __forceinline__ int mrt_middle_out(int b, int n) {
  const int mid = n >> 1;
  return (b & 1) ? mid - 1 - (b >> 1) : mid + (b >> 1);
}

// tiles.h mrt_pixel_of_tile_lane_ with the division by tiles_x done as a multiply-high when the
// host has proven it exact for this image size (KParams::tdiv_mul != 0).
__device__ __forceinline__ void mrt_pixel_of_tile_lane_fast(const KParams& P, int tile, int lane, int* x, int* y) {
  const int txn = mrt_tiles_x_(P.W);
  const int ty = P.tdiv_mul ? (int)__umulhi((unsigned)tile, P.tdiv_mul) : tile / txn;
  const int tx = tile - ty * txn;
  *x = (tx << MRT_TILE_SHIFT) + (lane & MRT_TILE_MASK);
  *y = (ty << MRT_TILE_SHIFT) + (lane >> MRT_TILE_SHIFT);
}

// Physical lane (0..31) of a warp -> logical lane (0..63) inside the 8x8 tile.  A warp owns an
// 8-wide x 4-tall half tile; each group of 8 consecutive lanes (the unit a 128-bit load is
// processed in) is a compact 4x2 pixel block, so its eight 2x2x2 footprints overlap as much
// as possible.  The LOGICAL lane is the reference's (y&7)*8+(x&7) (tiles.h).
__device__ __forceinline__ int mrt_logical_lane(int half, int lane) {
  const int qd = lane >> 3, j = lane & 7;
  const int x = ((qd & 1) << 2) + (j & 3);
  const int y = (half << 2) + ((qd >> 1) << 1) + (j >> 2);
  return (y << 3) + x;
}

// Exit of the ray from the aligned cell (cx,cy,cz) of edge 2^sh voxels, in index space, with
// the exit planes pulled inward by 1/64 voxel (>> any fp32 rounding of the positions):
// returns the number of consecutive sample slots, starting at slot time t, that are
// guaranteed to lie inside the cell; always >= 1 so the march progresses.
#define MRT_PLANE_EPS 0.015625f
__device__ __forceinline__ int mrt_cell_slots(const IdxRay& q, float ivx, float ivy, float ivz,
                                              int cx, int cy, int cz, int sh, float t, float inv_dt,
                                              int ox = 0, int oy = 0, int oz = 0) {
  // (cx,cy,cz) index cells of the (sub-)volume whose origin sits at global voxel (ox,oy,oz)
  const float plx = (q.dx > 0.0f) ? (float)(ox + ((cx + 1) << sh)) - MRT_PLANE_EPS : (float)(ox + (cx << sh)) + MRT_PLANE_EPS;
  const float ply = (q.dy > 0.0f) ? (float)(oy + ((cy + 1) << sh)) - MRT_PLANE_EPS : (float)(oy + (cy << sh)) + MRT_PLANE_EPS;
  const float plz = (q.dz > 0.0f) ? (float)(oz + ((cz + 1) << sh)) - MRT_PLANE_EPS : (float)(oz + (cz << sh)) + MRT_PLANE_EPS;
  // an axis the ray does not move along never bounds the exit
  const float tx = (q.dx != 0.0f) ? (plx - q.ox) * ivx : 3.0e38f;
  const float ty = (q.dy != 0.0f) ? (ply - q.oy) * ivy : 3.0e38f;
  const float tz = (q.dz != 0.0f) ? (plz - q.oz) * ivz : 3.0e38f;
  const float te = fminf(fminf(tx, ty), tz);
  const float ns = floorf((te - t) * inv_dt) + 1.0f;     // slots j with t + j*dt <= te
  return (ns >= 1.0f) ? (int)fminf(ns, 1.0e9f) : 1;
}

// The same question in SLOT units for a ray whose slot k sits at so + k*sd (index space): per axis
// isd = 1/sd and ec = -so*isd (isd = 0, ec = 3e38 for an axis the ray does not move along), so the
// slot coordinate of a plane pl is one fma.  (jx,jy,jz) = integer position inside the (sub-)volume
// whose origin sits at global voxel (ox,oy,oz); the cell is the aligned 2^sh cube around it.
struct SlotRay { float isdx, isdy, isdz, ecx, ecy, ecz; };
__device__ __forceinline__ SlotRay mrt_slot_ray(float sox, float soy, float soz, float sdx, float sdy, float sdz) {
  SlotRay r;
  r.isdx = (sdx != 0.0f) ? mrt_rcp(sdx) : 0.0f; r.ecx = (sdx != 0.0f) ? -sox * r.isdx : 3.0e38f;
  r.isdy = (sdy != 0.0f) ? mrt_rcp(sdy) : 0.0f; r.ecy = (sdy != 0.0f) ? -soy * r.isdy : 3.0e38f;
  r.isdz = (sdz != 0.0f) ? mrt_rcp(sdz) : 0.0f; r.ecz = (sdz != 0.0f) ? -soz * r.isdz : 3.0e38f;
  return r;
}
__device__ __forceinline__ int mrt_cell_slots_k(const SlotRay& r, int jx, int jy, int jz, int sh, float kf,
                                                int ox, int oy, int oz) {
  const int mask = (int)(0xffffffffu << sh);
  const float Sf = __int_as_float((127 + sh) << 23);           // 2^sh
  const float far_adj = Sf - MRT_PLANE_EPS;
  const float plx = (float)((jx & mask) + ox) + ((r.isdx > 0.0f) ? far_adj : MRT_PLANE_EPS);
  const float ply = (float)((jy & mask) + oy) + ((r.isdy > 0.0f) ? far_adj : MRT_PLANE_EPS);
  const float plz = (float)((jz & mask) + oz) + ((r.isdz > 0.0f) ? far_adj : MRT_PLANE_EPS);
  const float ke = fminf(fminf(fmaf(plx, r.isdx, r.ecx), fmaf(ply, r.isdy, r.ecy)), fmaf(plz, r.isdz, r.ecz));
  const float ns = floorf(ke - kf) + 1.0f;                     // slots j with k + j <= ke
  return (ns >= 1.0f) ? (int)fminf(ns, 1.0e9f) : 1;
}

// Sort-last: candidate slot range [ks, ke) of a ray inside a shard's sub-box (expanded by one
// voxel so that the global dims-1.001 clamp and fp32 rounding can never hide an owned slot);
// ownership itself is decided per slot, exactly, from the integer base index.
__device__ __forceinline__ void mrt_shard_range(const KParams& P, const IdxRay& q, float t0, float inv_dt, int n,
                                                int* ks, int* ke) {
  float tin = -3.0e38f, tout = 3.0e38f;
  const float o[3] = {q.ox, q.oy, q.oz}, d[3] = {q.dx, q.dy, q.dz};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float lo = (float)P.slo[a] - 1.0f, hi = (float)P.shi[a] + 1.0f;
    if (d[a] != 0.0f) {
      const float inv = 1.0f / d[a];
      const float ta = (lo - o[a]) * inv, tb = (hi - o[a]) * inv;
      tin = fmaxf(tin, fminf(ta, tb)); tout = fminf(tout, fmaxf(ta, tb));
    } else if (o[a] < lo || o[a] > hi) { tin = 3.0e38f; tout = -3.0e38f; }
  }
  if (!(tout >= tin)) { *ks = 0; *ke = 0; return; }
  const float a = floorf((tin - t0) * inv_dt) - 1.0f, b = ceilf((tout - t0) * inv_dt) + 2.0f;
  *ks = (int)fminf(fmaxf(a, 0.0f), (float)n);
  *ke = (int)fminf(fmaxf(b, 0.0f), (float)n);
}
// slots, starting at slot time t, guaranteed to stay inside the shard's owned box [slo, shi)
__device__ __forceinline__ int mrt_shard_slots(const KParams& P, const IdxRay& q, float ivx, float ivy, float ivz,
                                               float t, float inv_dt) {
  const float plx = (q.dx > 0.0f) ? (float)P.shi[0] - MRT_PLANE_EPS : (float)P.slo[0] + MRT_PLANE_EPS;
  const float ply = (q.dy > 0.0f) ? (float)P.shi[1] - MRT_PLANE_EPS : (float)P.slo[1] + MRT_PLANE_EPS;
  const float plz = (q.dz > 0.0f) ? (float)P.shi[2] - MRT_PLANE_EPS : (float)P.slo[2] + MRT_PLANE_EPS;
  const float tx = (q.dx != 0.0f) ? (plx - q.ox) * ivx : 3.0e38f;
  const float ty = (q.dy != 0.0f) ? (ply - q.oy) * ivy : 3.0e38f;
  const float tz = (q.dz != 0.0f) ? (plz - q.oz) * ivz : 3.0e38f;
  const float ns = floorf((fminf(fminf(tx, ty), tz) - t) * inv_dt) + 1.0f;
  return (ns >= 1.0f) ? (int)fminf(ns, 1.0e9f) : 1;
}
__device__ __forceinline__ bool mrt_shard_owns(const KParams& P, int ix, int iy, int iz) {
  return ix >= P.slo[0] && ix < P.shi[0] && iy >= P.slo[1] && iy < P.shi[1] && iz >= P.slo[2] && iz < P.shi[2];
}
