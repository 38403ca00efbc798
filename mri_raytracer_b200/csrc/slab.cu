// slab.cu — the single-volume near/far-slab renderer
// (volume_cs, scripts/volumeRendering/volume_render.slang:104-148) over a packed u8 volume.
//
// The ray positions are a running sum (rayPos += stepVec, :143) and gate a hard inside
// test (:136), so the whole position path uses explicit round-to-nearest ops: it is
// bit-identical to a one-op-at-a-time CPU evaluation.  One byte per voxel in HBM (the
// reference spends a u32 lane per voxel, app.py:150-158).
#include "march.cuh"
#include "kernels.h"
#include "../../include/mrt.h"

__device__ __forceinline__ float slab_u8(const uint8_t* __restrict__ vol, int x, int y, int z, int X, int Y) {
  const uint32_t idx = (uint32_t)x + (uint32_t)y * (uint32_t)X + (uint32_t)z * (uint32_t)X * (uint32_t)Y;   // :33
  return (float)__ldg(vol + idx) / 255.0f;                                                                  // :38
}

// sampleTrilinear, volume_render.slang:41-65
__device__ __forceinline__ float slab_trilinear(const uint8_t* __restrict__ vol, float u, float v, float w,
                                                int X, int Y, int Z) {
  const float x = __saturatef(u) * ((float)X - 1.0f);
  const float y = __saturatef(v) * ((float)Y - 1.0f);
  const float z = __saturatef(w) * ((float)Z - 1.0f);
  const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
  const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0;
  const int x1 = min(x0 + 1, X - 1), y1 = min(y0 + 1, Y - 1), z1 = min(z0 + 1, Z - 1);
  const float tx = x - fx0, ty = y - fy0, tz = z - fz0;
  const float c000 = slab_u8(vol, x0, y0, z0, X, Y), c100 = slab_u8(vol, x1, y0, z0, X, Y);
  const float c010 = slab_u8(vol, x0, y1, z0, X, Y), c110 = slab_u8(vol, x1, y1, z0, X, Y);
  const float c001 = slab_u8(vol, x0, y0, z1, X, Y), c101 = slab_u8(vol, x1, y0, z1, X, Y);
  const float c011 = slab_u8(vol, x0, y1, z1, X, Y), c111 = slab_u8(vol, x1, y1, z1, X, Y);
  const float c00 = lerpf(c000, c100, tx), c01 = lerpf(c001, c101, tx);
  const float c10 = lerpf(c010, c110, tx), c11 = lerpf(c011, c111, tx);
  const float c0 = lerpf(c00, c10, ty), c1 = lerpf(c01, c11, ty);
  return lerpf(c0, c1, tz);
}

struct SlabK {
  int W, H; float tan_half, steps; float n, f;
  float eye[3], U[3], V[3], Wv[3];
  int X, Y, Z; int tile_begin, tile_end;
};

__global__ void __launch_bounds__(128)
mrt_slab_kernel(const __grid_constant__ SlabK P, const uint8_t* __restrict__ vol, float4* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = P.tile_begin + mrt_middle_out(blockIdx.x, gridDim.x) * 2 + (warp >> 1);
  if (tile >= P.tile_end) return;
  int px, py;
  mrt_pixel_of_tile_lane_(tile, mrt_logical_lane(warp & 1, lane), P.W, &px, &py);
  if (px >= P.W || py >= P.H) return;                                                  // :109
  const float invx = __fdiv_rn(1.0f, (float)P.W), invy = __fdiv_rn(1.0f, (float)P.H);   // :111
  const float uvx = __fmul_rn(__fadd_rn((float)px, 0.5f), invx);                       // :115
  const float uvy = __fmul_rn(__fadd_rn((float)py, 0.5f), invy);
  const float ndcx = __fsub_rn(__fmul_rn(uvx, 2.0f), 1.0f);                            // :116
  const float ndcy = __fsub_rn(1.0f, __fmul_rn(uvy, 2.0f));
  const float aspect = __fdiv_rn((float)P.W, fmaxf(1.0f, (float)P.H));                 // :118
  const float vx = __fmul_rn(__fmul_rn(ndcx, aspect), P.tan_half);                     // :119
  const float vy = __fmul_rn(ndcy, P.tan_half);
  float pos[3], stp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {                                                        // :122-124
    const float wn = __fadd_rn(__fadd_rn(__fadd_rn(P.eye[i], __fmul_rn(P.U[i], __fmul_rn(vx, P.n))),
                                         __fmul_rn(P.V[i], __fmul_rn(vy, P.n))), __fmul_rn(P.Wv[i], __fmul_rn(1.0f, P.n)));
    const float wf = __fadd_rn(__fadd_rn(__fadd_rn(P.eye[i], __fmul_rn(P.U[i], __fmul_rn(vx, P.f))),
                                         __fmul_rn(P.V[i], __fmul_rn(vy, P.f))), __fmul_rn(P.Wv[i], __fmul_rn(1.0f, P.f)));
    pos[i] = wn;
    stp[i] = __fdiv_rn(__fsub_rn(wf, wn), P.steps);
  }
  const float scale = __fdiv_rn(4.0f, P.steps);                                        // :140
  const unsigned nsteps = (unsigned)P.steps;                                           // :134
  float accum = 0.0f;
  for (unsigned i = 0; i < nsteps; ++i) {
    const bool inside = pos[0] < 1.0f && pos[1] < 1.0f && pos[2] < 1.0f &&
                        pos[0] > -1.0f && pos[1] > -1.0f && pos[2] > -1.0f;            // :136
    if (inside && accum < 1.0f) {                                                      // :137
      const float u = __fmul_rn(0.5f, __fadd_rn(pos[0], 1.0f));                        // :139
      const float v = __fmul_rn(0.5f, __fadd_rn(pos[1], 1.0f));
      const float w = __fmul_rn(0.5f, __fadd_rn(pos[2], 1.0f));
      const float s = slab_trilinear(vol, u, v, w, P.X, P.Y, P.Z) * scale;             // :140
      accum += (1.0f - accum) * s;                                                     // :141
    }
    pos[0] = __fadd_rn(pos[0], stp[0]); pos[1] = __fadd_rn(pos[1], stp[1]); pos[2] = __fadd_rn(pos[2], stp[2]);  // :143
    if (accum > 0.995f) break;                                                         // :144
  }
  out[(size_t)py * P.W + px] = make_float4(accum, accum, accum, 1.0f);                 // :147
}

cudaError_t mrt_launch_slab(const MrtSlabParams& M, float tan_half, const uint8_t* vol, float* out,
                            int tile_begin, int tile_end, cudaStream_t st) {
  SlabK K;
  K.W = (int)M.imageSize[0]; K.H = (int)M.imageSize[1];
  K.tan_half = tan_half;
  K.steps = fmaxf(1.0f, M.stepCount);                                                  // :124,:130
  K.n = fmaxf(0.0f, M.nearPlane);                                                      // :120
  K.f = fmaxf(K.n, M.farPlane);                                                        // :121
  for (int i = 0; i < 3; ++i) { K.eye[i] = M.eye[i]; K.U[i] = M.U[i]; K.V[i] = M.V[i]; K.Wv[i] = M.W[i]; }
  K.X = (int)M.volDim[0]; K.Y = (int)M.volDim[1]; K.Z = (int)M.volDim[2];
  K.tile_begin = tile_begin; K.tile_end = tile_end;
  const int ntiles = tile_end - tile_begin;
  if (ntiles <= 0) return cudaSuccess;
  mrt_slab_kernel<<<(ntiles + 1) / 2, 128, 0, st>>>(K, vol, (float4*)out);
  return cudaGetLastError();
}
