// occupancy.cu — min/max occupancy brick grid and the per-frame skip-level map.
//
// No reference code (empty-space skipping is only mentioned in the reference's docs:
// docs/Methodology-ROI-Neural-Volumetric-Rendering.md:34, docs/showcase-plan.md:20).
// Contract: skipping must never change the image.  A brick is marked empty only when every
// sample slot whose trilinear base index lies in it is provably a no-op (sigma == 0 for the
// whole reachable TF range, and no overlay label present).
#include "march.cuh"
#include "kernels.h"
#include <float.h>
#include <limits.h>

// ---------------------------------------------------------------- build (once per volume)
// brick b covers voxel indices [8b, 8b+8] per axis (clipped): the union of the 2x2x2
// footprints of all samples with base index in [8b, 8b+7], and of the nearest-label
// footprint round(p) of the same samples.
template <int NCH, int HALF = 0>
__global__ void __launch_bounds__(128)
mrt_build_minmax_kernel(const typename VoxT<NCH, HALF>::T* __restrict__ vol, int X, int Y, int Z,
                        size_t pitchY, size_t pitchZ, int nbx, int nby, float2* __restrict__ minmax) {
  const int b = blockIdx.x;
  const int bx = b % nbx, by = (b / nbx) % nby, bz = b / (nbx * nby);
  const int x0 = bx << 3, y0 = by << 3, z0 = bz << 3;
  const int ex = min(9, X - x0), ey = min(9, Y - y0), ez = min(9, Z - z0);
  const int nvox = ex * ey * ez;
  float mn[NCH], mx[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) { mn[c] = FLT_MAX; mx[c] = -FLT_MAX; }
  for (int i = threadIdx.x; i < nvox; i += blockDim.x) {
    const int lx = i % ex, ly = (i / ex) % ey, lz = i / (ex * ey);
    const size_t idx = (size_t)(x0 + lx) + pitchY * (size_t)(y0 + ly) + pitchZ * (size_t)(z0 + lz);
    auto v = mrt_f32(__ldg(vol + idx));
    if (HALF == 2) *reinterpret_cast<float*>(&v) = __fdiv_rn(*reinterpret_cast<float*>(&v), 255.0f);   // u8: the value is v/255 (volume_render.slang:38)
    const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
    for (int c = 0; c < NCH; ++c) { mn[c] = fminf(mn[c], f[c]); mx[c] = fmaxf(mx[c], f[c]); }
  }
  __shared__ float s_mn[4][NCH], s_mx[4][NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) { s_mn[warp][c] = mn[c]; s_mx[warp][c] = mx[c]; }
  }
  __syncthreads();
  if (threadIdx.x < NCH) {
    const int c = threadIdx.x;
    float a = s_mn[0][c], z = s_mx[0][c];
    for (int w = 1; w < 4; ++w) { a = fminf(a, s_mn[w][c]); z = fmaxf(z, s_mx[w][c]); }
    minmax[(size_t)b * NCH + c] = make_float2(a, z);
  }
}

cudaError_t mrt_launch_build_occupancy(const void* packed, int pc, int X, int Y, int Z, float* minmax,
                                       cudaStream_t st) {
  const int nbx = (X + 7) >> 3, nby = (Y + 7) >> 3, nbz = (Z + 7) >> 3;
  const int nb = nbx * nby * nbz;
  int64_t pY, pZ;
  mrt_layout(pc, X, Y, Z, &pY, &pZ);
  switch (pc) {
    case 1: mrt_build_minmax_kernel<1><<<nb, 128, 0, st>>>((const float*)packed, X, Y, Z, pY, pZ, nbx, nby, (float2*)minmax); break;
    case 2: mrt_build_minmax_kernel<2><<<nb, 128, 0, st>>>((const float2*)packed, X, Y, Z, pY, pZ, nbx, nby, (float2*)minmax); break;
    case 4: mrt_build_minmax_kernel<4><<<nb, 128, 0, st>>>((const float4*)packed, X, Y, Z, pY, pZ, nbx, nby, (float2*)minmax); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t mrt_launch_build_occupancy_f16(const void* packed, int X, int Y, int Z, float* minmax, cudaStream_t st) {
  const int nbx = (X + 7) >> 3, nby = (Y + 7) >> 3, nbz = (Z + 7) >> 3;
  int64_t pY, pZ;
  mrt_layout_e(1, 2, X, Y, Z, &pY, &pZ);
  mrt_build_minmax_kernel<1, 1><<<nbx * nby * nbz, 128, 0, st>>>((const __half*)packed, X, Y, Z, pY, pZ, nbx, nby,
                                                                (float2*)minmax);
  return cudaGetLastError();
}
cudaError_t mrt_launch_build_occupancy_u8(const void* packed, int X, int Y, int Z, float* minmax, cudaStream_t st) {
  const int nbx = (X + 7) >> 3, nby = (Y + 7) >> 3, nbz = (Z + 7) >> 3;
  int64_t pY, pZ;
  mrt_layout_e(1, 1, X, Y, Z, &pY, &pZ);
  mrt_build_minmax_kernel<1, 2><<<nbx * nby * nbz, 128, 0, st>>>((const uint8_t*)packed, X, Y, Z, pY, pZ, nbx, nby,
                                                                (float2*)minmax);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(128)
mrt_label_any_kernel(const int32_t* __restrict__ lab, int X, int Y, int Z, int nbx, int nby,
                     uint8_t* __restrict__ any) {
  const int b = blockIdx.x;
  const int bx = b % nbx, by = (b / nbx) % nby, bz = b / (nbx * nby);
  const int x0 = bx << 3, y0 = by << 3, z0 = bz << 3;
  const int ex = min(9, X - x0), ey = min(9, Y - y0), ez = min(9, Z - z0);
  const int nvox = ex * ey * ez;
  int found = 0;
  for (int i = threadIdx.x; i < nvox; i += blockDim.x) {
    const int lx = i % ex, ly = (i / ex) % ey, lz = i / (ex * ey);
    const size_t idx = (size_t)(x0 + lx) + (size_t)X * ((size_t)(y0 + ly) + (size_t)Y * (size_t)(z0 + lz));
    const int l = __ldg(lab + idx);
    found |= (l > 0 && l < 8);                                  // brats_rt.slang:145
  }
  const int r = __syncthreads_or(found);
  if (threadIdx.x == 0) any[b] = (uint8_t)(r != 0);
}

cudaError_t mrt_launch_label_occupancy(const int32_t* labels, int X, int Y, int Z, uint8_t* any, cudaStream_t st) {
  const int nbx = (X + 7) >> 3, nby = (Y + 7) >> 3, nbz = (Z + 7) >> 3;
  mrt_label_any_kernel<<<nbx * nby * nbz, 128, 0, st>>>(labels, X, Y, Z, nbx, nby, any);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- classify (per frame)
// Conservative interval arithmetic: blended value range -> window/level range -> the TF
// entries a sample in this brick can touch.  Margins (1e-5 on v, 1e-3 of a LUT bin) are far
// wider than any fp32 rounding in the sampler and far narrower than a LUT bin.
template <int NCH>
__device__ __forceinline__ bool mrt_brick_active(const KParams& P, const float2* __restrict__ minmax,
                                                 const float4* __restrict__ tf, const uint8_t* __restrict__ seg_any,
                                                 const uint8_t* __restrict__ pred_any, int b) {
  bool act = false;
  float lo = 0.0f, hi = 0.0f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const float2 mm = __ldg(minmax + (size_t)b * NCH + c);
    const float w = P.wgt[c];
    lo += (w >= 0.0f) ? w * mm.x : w * mm.y;
    hi += (w >= 0.0f) ? w * mm.y : w * mm.x;
  }
  float a = lo * P.inv_wsum, z = hi * P.inv_wsum;
  if (a > z) { const float t = a; a = z; z = t; }
  const float mv = 1e-5f * (1.0f + fmaxf(fabsf(a), fabsf(z)));
  a -= mv; z += mv;
  float r0 = (a - P.lo) * P.inv_ww, r1 = (z - P.lo) * P.inv_ww;
  if (r0 > r1) { const float t = r0; r0 = r1; r1 = t; }
  const float mr = 1e-5f * (1.0f + fmaxf(fabsf(r0), fabsf(r1)));
  r0 -= mr; r1 += mr;
  float v0 = __saturatef(r0), v1 = __saturatef(r1);
  if (P.gamma != 1.0f) {                       // pow is monotone on [0,1] for gamma > 0
    const float p0 = powf(v0, P.gamma), p1 = powf(v1, P.gamma);
    v0 = fmaxf(fminf(p0, p1) - 1e-5f, 0.0f); v1 = fminf(fmaxf(p0, p1) + 1e-5f, 1.0f);
    if (!(P.gamma > 0.0f)) { v0 = 0.0f; v1 = 1.0f; }
  }
  if (P.tfMode == 0) {
    act = (v1 > 0.0f) && (P.ia != 0.0f);        // sigma = val*intensityAlpha, gated on val > 0 (:135)
  } else {
    const float s = (float)(P.tfN - 1);
    int j0 = (int)floorf(v0 * s - 1e-3f), j1 = (int)floorf(v1 * s + 1e-3f) + 1;
    j0 = max(j0, 0); j1 = min(j1, P.tfN - 1);
    for (int j = j0; j <= j1; ++j) {
      if (__ldg(&tf[j].w) != 0.0f) { act = true; break; }
    }
  }
  if (P.showSeg && (seg_any == nullptr || seg_any[b])) act = true;     // no label grid supplied: stay exact
  if (P.showPred && (pred_any == nullptr || pred_any[b])) act = true;
  return act;
}

// One CTA per 8x8x8-brick super-cell (64^3 voxels): classify each brick, then reduce the
// flags over the aligned 2^3, 4^3 and 8^3 brick cells and store, per brick, the largest
// aligned empty cell that contains it (skip level 1..4; 0 = active).  Bricks outside the
// grid count as empty (no sample slot ever lands there).  Non-FLAT levels additionally carry
// 0x80 | 2..4 for an active brick inside an aligned all-active 2^3 / 4^3 / 8^3 brick cell.
// FLAT variant (for the backward): a brick only counts as skippable if it is empty AND flat
// (min == max: every voxel holds the same value), and a coarse cell only if all its bricks are
// flat-empty with the SAME value — then every slot inside has identical TF bin, colour and
// dL/dsigma, which the backward adds in closed form.
template <int NCH, bool FLAT>
__global__ void __launch_bounds__(512)
mrt_classify_kernel(const __grid_constant__ KParams P, const float2* __restrict__ minmax,
                    const float4* __restrict__ tf, const uint8_t* __restrict__ seg_any,
                    const uint8_t* __restrict__ pred_any, uint8_t* __restrict__ levels, int* __restrict__ box,
                    int sbx, int sby) {
  __shared__ uint8_t s_act[512];
  __shared__ uint8_t s_or2[64];
  __shared__ uint8_t s_or4[8];
  __shared__ uint8_t s_or8;
  __shared__ uint8_t s_all[512], s_and2[64], s_and4[8], s_and8;   // all-active cells (bricks outside the grid: don't care)
  __shared__ int s_box[6];
  __shared__ float s_lo[512], s_hi[512], s_lo2[64], s_hi2[64], s_lo4[8], s_hi4[8];
  const int t = threadIdx.x;
  const int lx = t & 7, ly = (t >> 3) & 7, lz = t >> 6;
  const int sx = blockIdx.x % sbx, sy = (blockIdx.x / sbx) % sby, sz = blockIdx.x / (sbx * sby);
  const int bx = (sx << 3) + lx, by = (sy << 3) + ly, bz = (sz << 3) + lz;
  const bool inside = bx < P.nbx && by < P.nby && bz < P.nbz;
  const int b = (bz * P.nby + by) * P.nbx + bx;
  if (t < 6) s_box[t] = INT_MIN;
  bool act = inside && mrt_brick_active<NCH>(P, minmax, tf, seg_any, pred_any, b);
  if (FLAT) {
    float lo = 3.0e38f, hi = -3.0e38f;          // out-of-grid bricks are compatible with any value
    if (inside) {
      bool flat = true;
#pragma unroll
      for (int c = 0; c < NCH; ++c) { const float2 mm = __ldg(minmax + (size_t)b * NCH + c); flat = flat && (mm.x == mm.y); }
      const float2 m0 = __ldg(minmax + (size_t)b * NCH);
      lo = m0.x; hi = m0.y;
      if (!flat || NCH != 1) act = true;
    }
    s_lo[t] = lo; s_hi[t] = hi;
  }
  s_act[t] = act;
  if (!FLAT) s_all[t] = act || !inside;
  __syncthreads();
  // bounding box of the active bricks (brick units), all six as maxima: (-lo, hi).  The march
  // culls rays / whole CTAs against it before the exact ray set-up and clips every ray's slot
  // range to it (mrt_active_box).
  if (act) {
    atomicMax(&s_box[0], -bx); atomicMax(&s_box[1], -by); atomicMax(&s_box[2], -bz);
    atomicMax(&s_box[3], bx);  atomicMax(&s_box[4], by);  atomicMax(&s_box[5], bz);
  }
  if (t < 64) {            // 2x2x2 groups: group (gx,gy,gz) in 4x4x4
    const int gx = t & 3, gy = (t >> 2) & 3, gz = t >> 4;
    int o = 0, a = 1;
    float lo = 3.0e38f, hi = -3.0e38f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = ((gz * 2 + (i >> 2)) << 6) + ((gy * 2 + ((i >> 1) & 1)) << 3) + gx * 2 + (i & 1);
      o |= s_act[j];
      if (!FLAT) a &= s_all[j];
      if (FLAT) { lo = fminf(lo, s_lo[j]); hi = fmaxf(hi, s_hi[j]); }
    }
    if (FLAT) { s_lo2[t] = lo; s_hi2[t] = hi; if (lo < hi) o = 1; }
    s_or2[t] = (uint8_t)o;
    if (!FLAT) s_and2[t] = (uint8_t)a;
  }
  __syncthreads();
  if (t < 8) {             // 4x4x4 groups: (hx,hy,hz) in 2x2x2, each = 2x2x2 of the or2 groups
    const int hx = t & 1, hy = (t >> 1) & 1, hz = t >> 2;
    int o = 0, a = 1;
    float lo = 3.0e38f, hi = -3.0e38f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = ((hz * 2 + (i >> 2)) << 4) + ((hy * 2 + ((i >> 1) & 1)) << 2) + hx * 2 + (i & 1);
      o |= s_or2[j];
      if (!FLAT) a &= s_and2[j];
      if (FLAT) { lo = fminf(lo, s_lo2[j]); hi = fmaxf(hi, s_hi2[j]); }
    }
    if (FLAT) { s_lo4[t] = lo; s_hi4[t] = hi; if (lo < hi) o = 1; }
    s_or4[t] = (uint8_t)o;
    if (!FLAT) s_and4[t] = (uint8_t)a;
  }
  __syncthreads();
  if (t == 0) {
    int o = 0, a = 1;
    float lo = 3.0e38f, hi = -3.0e38f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o |= s_or4[i];
      if (!FLAT) a &= s_and4[i];
      if (FLAT) { lo = fminf(lo, s_lo4[i]); hi = fmaxf(hi, s_hi4[i]); }
    }
    if (FLAT && lo < hi) o = 1;
    s_or8 = (uint8_t)o;
    if (!FLAT) s_and8 = (uint8_t)a;
  }
  __syncthreads();
  if (t < 6 && s_box[t] != INT_MIN) atomicMax(box + t, s_box[t]);     // s_box complete: 3 barriers ago
  if (inside) {
    int lvl = 0;
    if (!act) {
      lvl = 1;
      if (!s_or2[((lz >> 1) << 4) + ((ly >> 1) << 2) + (lx >> 1)]) {
        lvl = 2;
        if (!s_or4[((lz >> 2) << 2) + ((ly >> 2) << 1) + (lx >> 2)]) {
          lvl = 3;
          if (!s_or8) lvl = 4;
        }
      }
    }
    // the march's levels also mark the largest aligned ALL-ACTIVE cell around an active brick
    // (0x80 | 2..4 = 16^3, 32^3, 64^3 voxels; cell edge 2^((lvl & 7) + 2) like the empty levels), so a
    // ray inside a solid region asks once per cell instead of once per brick
    if (!FLAT && act && s_and2[((lz >> 1) << 4) + ((ly >> 1) << 2) + (lx >> 1)]) {
      lvl = 0x80 | 2;
      if (s_and4[((lz >> 2) << 2) + ((ly >> 2) << 1) + (lx >> 2)]) {
        lvl = 0x80 | 3;
        if (s_and8) lvl = 0x80 | 4;
      }
    }
    levels[b] = (uint8_t)lvl;
  }
}

cudaError_t mrt_launch_classify(const KParams& P, const float* minmax, int pc, const float* tf,
                                const uint8_t* seg_any, const uint8_t* pred_any, uint8_t* levels,
                                bool require_flat, cudaStream_t st) {
  const int sbx = (P.nbx + 7) >> 3, sby = (P.nby + 7) >> 3, sbz = (P.nbz + 7) >> 3;
  const int grid = sbx * sby * sbz;
  // tail of the levels buffer (mrt_skip_levels_bytes): the active-brick box, reset to "nothing"
  int* box = reinterpret_cast<int*>(levels + mrt_levels_box_offset(P.nbx * P.nby * P.nbz));
  cudaError_t e0 = cudaMemsetAsync(box, 0x80, 8 * sizeof(int), st);
  if (e0 != cudaSuccess) return e0;
#define MRT_CL(N, F) mrt_classify_kernel<N, F><<<grid, 512, 0, st>>>(P, (const float2*)minmax, (const float4*)tf, seg_any, pred_any, levels, box, sbx, sby)
  switch (pc) {
    case 1: if (require_flat) MRT_CL(1, true); else MRT_CL(1, false); break;
    case 2: if (require_flat) MRT_CL(2, true); else MRT_CL(2, false); break;
    case 4: if (require_flat) MRT_CL(4, true); else MRT_CL(4, false); break;
    default: return cudaErrorInvalidValue;
  }
#undef MRT_CL
  return cudaGetLastError();
}
