// c_api.cu — the extern "C" boundary declared in include/mrt.h.
//
// Validates arguments, derives the kernel constant block (KParams) from the reference's
// `struct Params` layout (MrtParams), launches on the caller's stream, and reports errors
// through a thread-local string.  No allocation, no synchronisation (except *_host).
#include "march.cuh"
#include "kernels.h"
#include <stdlib.h>
#include "../../include/mrt.h"
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  return fail(MRT_ERR_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
}
#define MRT_REQUIRE(cond, ...) do { if (!(cond)) return fail(MRT_ERR_BAD_ARG, __VA_ARGS__); } while (0)

extern "C" {

int mrt_version(void) { return MRT_VERSION; }
const char* mrt_last_error(void) { return g_err; }
size_t mrt_sizeof_params(void) { return sizeof(MrtParams); }
size_t mrt_sizeof_slab_params(void) { return sizeof(MrtSlabParams); }
size_t mrt_sizeof_camera(void) { return sizeof(MrtCamera); }
int32_t mrt_max_views_per_launch(void) { return MRT_MAX_VIEWS; }

// ---------------------------------------------------------------- tiles (host)
int32_t mrt_tiles_x(int32_t W) { return mrt_tiles_x_(W); }
int32_t mrt_tiles_y(int32_t H) { return mrt_tiles_y_(H); }
int32_t mrt_tile_count(int32_t W, int32_t H) { return mrt_tiles_x_(W) * mrt_tiles_y_(H); }
int32_t mrt_tile_of_pixel(int32_t x, int32_t y, int32_t W) { return mrt_tile_of_pixel_(x, y, W); }
int32_t mrt_lane_of_pixel(int32_t x, int32_t y) { return mrt_lane_of_pixel_(x, y); }
void mrt_rank_tile_range(int32_t ntiles, int32_t rank, int32_t nranks, int32_t* begin, int32_t* end) {
  *begin = mrt_rank_tile_begin_(ntiles, rank, nranks);
  *end = mrt_rank_tile_begin_(ntiles, rank + 1, nranks);
}
int mrt_tile_index_map(int32_t W, int32_t H, int32_t* out_tile, int32_t* out_lane, void* stream) {
  MRT_REQUIRE(W > 0 && H > 0 && out_tile && out_lane, "tile_index_map: bad arguments");
  cudaError_t e = mrt_launch_tile_map(W, H, out_tile, out_lane, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "tile_index_map");
}

// ---------------------------------------------------------------- layout
void mrt_packed_layout(int32_t C, int32_t X, int32_t Y, int32_t Z, int64_t* pitchY, int64_t* pitchZ) {
  mrt_layout(mrt_packed_channels(C), X, Y, Z, pitchY, pitchZ);
}
size_t mrt_packed_volume_bytes(int32_t C, int32_t X, int32_t Y, int32_t Z) {
  if (C < 1 || C > 4 || X < 1 || Y < 1 || Z < 1) return 0;
  int64_t pY, pZ;
  mrt_layout(mrt_packed_channels(C), X, Y, Z, &pY, &pZ);
  return (size_t)pZ * Z * sizeof(float) * mrt_packed_channels(C);
}
static int check_dims(const char* who, int C, int X, int Y, int Z) {
  MRT_REQUIRE(C >= 1 && C <= 4, "%s: C=%d outside 1..4", who, C);
  // dims-1.001 clamp (brats_rt.slang:62) needs >= 2 voxels per axis
  MRT_REQUIRE(X >= 2 && Y >= 2 && Z >= 2, "%s: dims (%d,%d,%d) must be >= 2 per axis", who, X, Y, Z);
  int64_t pY, pZ;
  mrt_layout(mrt_packed_channels(C), X, Y, Z, &pY, &pZ);
  MRT_REQUIRE((uint64_t)pZ * Z < (1ull << 32), "%s: more than 2^32 voxels per shard (SURVEY Q14)", who);
  return MRT_OK;
}
int mrt_pack_volume_f32(const float* planar, int32_t C, int32_t X, int32_t Y, int32_t Z, void* packed, void* stream) {
  MRT_REQUIRE(planar && packed, "pack_volume: null pointer");
  if (int r = check_dims("pack_volume", C, X, Y, Z)) return r;
  cudaError_t e = mrt_launch_pack(planar, C, X, Y, Z, packed, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "pack_volume");
}
size_t mrt_packed_volume_bytes_f16(int32_t X, int32_t Y, int32_t Z) {
  if (X < 1 || Y < 1 || Z < 1) return 0;
  int64_t pY, pZ;
  mrt_layout_e(1, 2, X, Y, Z, &pY, &pZ);
  return ((size_t)pZ * Z * 2 + 15) & ~(size_t)15;
}
void mrt_packed_layout_f16(int32_t X, int32_t Y, int32_t Z, int64_t* pitchY, int64_t* pitchZ) {
  mrt_layout_e(1, 2, X, Y, Z, pitchY, pitchZ);
}
static int check_dims_f16(const char* who, int X, int Y, int Z) {
  MRT_REQUIRE(X >= 2 && Y >= 2 && Z >= 2, "%s: dims (%d,%d,%d) must be >= 2 per axis", who, X, Y, Z);
  int64_t pY, pZ;
  mrt_layout_e(1, 2, X, Y, Z, &pY, &pZ);
  MRT_REQUIRE((uint64_t)pZ * Z < (1ull << 32), "%s: more than 2^32 voxels per shard (SURVEY Q14)", who);
  return MRT_OK;
}
int mrt_pack_volume_f16(const void* planar_f16, int32_t X, int32_t Y, int32_t Z, void* packed, void* stream) {
  MRT_REQUIRE(planar_f16 && packed, "pack_volume_f16: null pointer");
  if (int r = check_dims_f16("pack_volume_f16", X, Y, Z)) return r;
  cudaError_t e = mrt_launch_pack_f16(planar_f16, X, Y, Z, packed, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "pack_volume_f16");
}
int mrt_unpack_volume_f16(const void* packed, int32_t X, int32_t Y, int32_t Z, void* planar_f16, void* stream) {
  MRT_REQUIRE(planar_f16 && packed, "unpack_volume_f16: null pointer");
  if (int r = check_dims_f16("unpack_volume_f16", X, Y, Z)) return r;
  cudaError_t e = mrt_launch_unpack_f16(packed, X, Y, Z, planar_f16, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "unpack_volume_f16");
}
int mrt_build_occupancy_f16(const void* packed, int32_t X, int32_t Y, int32_t Z, float* minmax, void* stream) {
  MRT_REQUIRE(packed && minmax, "build_occupancy_f16: null pointer");
  if (int r = check_dims_f16("build_occupancy_f16", X, Y, Z)) return r;
  cudaError_t e = mrt_launch_build_occupancy_f16(packed, X, Y, Z, minmax, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "build_occupancy_f16");
}
// u8 storage (scripts/volumeRendering/app.py:145-158 uploads the volume as bytes; volume_render.slang:33-38
// reads value = byte/255): 1 byte per voxel, 128 voxels per 128-byte line
size_t mrt_packed_volume_bytes_u8(int32_t X, int32_t Y, int32_t Z) {
  if (X < 1 || Y < 1 || Z < 1) return 0;
  int64_t pY, pZ;
  mrt_layout_e(1, 1, X, Y, Z, &pY, &pZ);
  return ((size_t)pZ * Z + 15) & ~(size_t)15;
}
// quad layout of a single-channel fp32 volume (march.cuh VoxT<1,3>)
size_t mrt_packed_volume_bytes_quad(int32_t X, int32_t Y, int32_t Z) {
  if (X < 2 || Y < 2 || Z < 2) return 0;
  int64_t pY, pZ;
  mrt_layout_e(1, 16, X, Y, Z, &pY, &pZ);
  return (size_t)pZ * (size_t)Z * 16;
}
int mrt_pack_volume_quad(const float* packed1, int32_t X, int32_t Y, int32_t Z, void* quad, void* stream) {
  MRT_REQUIRE(packed1 && quad, "pack_volume_quad: null pointer");
  if (int r = check_dims("pack_volume_quad", 1, X, Y, Z)) return r;
  int64_t pY, pZ;
  mrt_layout_e(1, 16, X, Y, Z, &pY, &pZ);
  MRT_REQUIRE((uint64_t)pZ * (uint64_t)Z < (1ull << 32), "pack_volume_quad: more than 2^32 elements");
  cudaError_t e = mrt_launch_pack_quad(packed1, X, Y, Z, quad, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "pack_volume_quad");
}
size_t mrt_packed_volume_bytes_quad_f16(int32_t X, int32_t Y, int32_t Z) {
  if (X < 2 || Y < 2 || Z < 2) return 0;
  int64_t pY, pZ;
  mrt_layout_e(1, 8, X, Y, Z, &pY, &pZ);
  return (size_t)pZ * (size_t)Z * 8;
}
int mrt_pack_volume_quad_f16(const void* packed_f16, int32_t X, int32_t Y, int32_t Z, void* quad, void* stream) {
  MRT_REQUIRE(packed_f16 && quad, "pack_volume_quad_f16: null pointer");
  if (int r = check_dims_f16("pack_volume_quad_f16", X, Y, Z)) return r;
  int64_t pY, pZ;
  mrt_layout_e(1, 8, X, Y, Z, &pY, &pZ);
  MRT_REQUIRE((uint64_t)pZ * (uint64_t)Z < (1ull << 32), "pack_volume_quad_f16: more than 2^32 elements");
  cudaError_t e = mrt_launch_pack_quad_f16(packed_f16, X, Y, Z, quad, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "pack_volume_quad_f16");
}
static int check_dims_u8(const char* who, int X, int Y, int Z) {
  MRT_REQUIRE(X >= 2 && Y >= 2 && Z >= 2, "%s: dims (%d,%d,%d) must be >= 2 per axis", who, X, Y, Z);
  int64_t pY, pZ;
  mrt_layout_e(1, 1, X, Y, Z, &pY, &pZ);
  MRT_REQUIRE((uint64_t)pZ * Z < (1ull << 32), "%s: more than 2^32 voxels per shard (SURVEY Q14)", who);
  return MRT_OK;
}
int mrt_pack_volume_u8(const uint8_t* planar_u8, int32_t X, int32_t Y, int32_t Z, void* packed, void* stream) {
  MRT_REQUIRE(planar_u8 && packed, "pack_volume_u8: null pointer");
  if (int r = check_dims_u8("pack_volume_u8", X, Y, Z)) return r;
  cudaError_t e = mrt_launch_pack_u8(planar_u8, X, Y, Z, packed, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "pack_volume_u8");
}
int mrt_build_occupancy_u8(const void* packed, int32_t X, int32_t Y, int32_t Z, float* minmax, void* stream) {
  MRT_REQUIRE(packed && minmax, "build_occupancy_u8: null pointer");
  if (int r = check_dims_u8("build_occupancy_u8", X, Y, Z)) return r;
  cudaError_t e = mrt_launch_build_occupancy_u8(packed, X, Y, Z, minmax, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "build_occupancy_u8");
}
int mrt_unpack_volume_f32(const void* packed, int32_t C, int32_t X, int32_t Y, int32_t Z, float* planar, void* stream) {
  MRT_REQUIRE(planar && packed, "unpack_volume: null pointer");
  if (int r = check_dims("unpack_volume", C, X, Y, Z)) return r;
  cudaError_t e = mrt_launch_unpack(packed, C, X, Y, Z, planar, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "unpack_volume");
}

// ---------------------------------------------------------------- params
static int derive(const MrtParams* M, int C, int tfN, bool have_bits, int tile_begin, int tile_end, KParams* K) {
  MRT_REQUIRE(M != nullptr, "params is null");
  memset(K, 0, sizeof(*K));
  K->W = (int)M->imageSize[0]; K->H = (int)M->imageSize[1];
  MRT_REQUIRE(K->W > 0 && K->H > 0 && K->W <= 65536 && K->H <= 65536, "imageSize (%d,%d) invalid", K->W, K->H);
  for (int i = 0; i < 3; ++i) {
    K->eye[i] = M->eye[i]; K->U[i] = M->U[i]; K->V[i] = M->V[i]; K->Wv[i] = M->W[i];
    K->bmin[i] = M->volMin[i]; K->vs[i] = M->voxelSize[i]; K->dims[i] = (int)M->dims[i];
    K->bg[i] = M->bgColor[i];
    MRT_REQUIRE(M->voxelSize[i] > 0.0f, "voxelSize[%d] must be > 0", i);
  }
  MRT_REQUIRE(C >= 1 && C <= 4, "render: C=%d outside 1..4", C);
  MRT_REQUIRE(K->dims[0] >= 2 && K->dims[1] >= 2 && K->dims[2] >= 2, "render: dims (%d,%d,%d) must be >= 2 per axis",
              K->dims[0], K->dims[1], K->dims[2]);
  int ldim[3] = {K->dims[0], K->dims[1], K->dims[2]};       // dims of the buffer actually sampled
  if (M->shardEnabled) {
    K->shard = 1;
    for (int i = 0; i < 3; ++i) {
      MRT_REQUIRE(M->shardLo[i] < M->shardHi[i] && (int)M->shardHi[i] <= K->dims[i] - 1,
                  "shard range [%u,%u) on axis %d outside the cell range [0,%d)", M->shardLo[i], M->shardHi[i], i,
                  K->dims[i] - 1);
      K->slo[i] = (int)M->shardLo[i]; K->shi[i] = (int)M->shardHi[i];
      ldim[i] = K->shi[i] - K->slo[i] + 1;
    }
    MRT_REQUIRE(!M->showSeg && !M->showPred, "label overlays are not supported on sharded volumes");
    MRT_REQUIRE(M->tMode == 0, "sharded volumes need indexed stepping (tMode 0)");
  }
  MRT_REQUIRE(M->volDtype <= 4, "volDtype %u unknown (0 fp32, 1 fp16, 2 u8, 3 fp32 quad, 4 fp16 quad)", M->volDtype);
  K->half = (int)M->volDtype;
  if (K->half) {
    MRT_REQUIRE(C == 1, "fp16 / u8 / quad volumes are single-channel (C=%d)", C);
    MRT_REQUIRE(!M->showSeg && !M->showPred, "label overlays are not supported on fp16 / u8 / quad volumes");
  }
  {
    int64_t pY, pZ;
    mrt_layout_e(mrt_packed_channels(C), K->half == 1 ? 2 : (K->half == 2 ? 1 : (K->half == 3 ? 16 : (K->half == 4 ? 8 : 4))), ldim[0], ldim[1], ldim[2], &pY, &pZ);
    MRT_REQUIRE((uint64_t)pZ * ldim[2] < (1ull << 32), "more than 2^32 voxels per shard (SURVEY Q14)");
    K->pitchY = (unsigned)pY; K->pitchZ = (unsigned)pZ;
    K->base_off = K->shard ? (unsigned)(K->slo[0] + K->slo[1] * pY + K->slo[2] * pZ) : 0u;
    K->idx_bias = K->base_off + 0x4b000000u * (1u + K->pitchY + K->pitchZ);
  }
  // tan evaluated once in double, rounded to float (documented deviation from the per-thread fp32 tan)
  K->ortho = M->ortho ? 1 : 0;
  K->halfH = M->orthoHalfHeight;
  if (!K->ortho) {
    MRT_REQUIRE(M->fovY > 0.0f && M->fovY < 3.14159f, "fovY %g outside (0, pi)", (double)M->fovY);
    K->focal = (float)(1.0 / tan(0.5 * (double)M->fovY));
  } else {
    MRT_REQUIRE(M->orthoHalfHeight > 0.0f, "orthoHalfHeight must be > 0");
    K->focal = 1.0f;
  }
  MRT_REQUIRE(M->stepSize > 0.0f, "stepSize must be > 0");
  K->dt = M->stepSize; K->nearT = M->nearT; K->farT = M->farT;
  K->inv_dt = 1.0f / K->dt;
  for (int i = 0; i < 3; ++i) K->inv_vs[i] = 1.0f / K->vs[i];
  MRT_REQUIRE(M->ww > 0.0f, "ww must be > 0 (SURVEY Q10)");
  float wsum = 0.0f;
  for (int c = 0; c < 4; ++c) {
    const bool en = (c < C) && (M->volEnabled[c] != 0);
    K->wgt[c] = en ? M->volWeight[c] : 0.0f;
    if (en) wsum += M->volWeight[c];                       // brats_rt.slang:125-128
  }
  K->inv_wsum = (wsum > 0.0f) ? 1.0f / wsum : 1.0f;        // :130
  K->lo = M->wl - M->ww * 0.5f;                            // :132
  K->inv_ww = 1.0f / M->ww;
  for (int c = 0; c < 4; ++c) K->wq[c] = K->wgt[c] * K->inv_wsum * K->inv_ww;
  if (K->half == 2) K->wq[0] = K->wq[0] / 255.0f;         // u8 voxels are sampled as integers: value = byte/255 (volume_render.slang:38)
  K->wbias = -K->lo * K->inv_ww;
  K->neg_dt_log2e = -(float)((double)M->stepSize * 1.4426950408889634);
  K->ia = M->intensityAlpha;
  K->gamma = M->gamma;
  K->showSeg = M->showSeg ? 1 : 0; K->showPred = M->showPred ? 1 : 0;
  memcpy(K->lut, M->lutColorAlpha, sizeof(K->lut));
  K->thr = (M->ertThreshold != 0.0f) ? M->ertThreshold : 0.01f;   // :117
  K->maxSteps = (int)M->maxSteps;
  MRT_REQUIRE(M->tMode <= 1, "tMode %u unknown", M->tMode);
  K->tMode = (int)M->tMode; K->alphaMode = M->alphaMode ? 1 : 0;
  K->tfMode = M->tfMode ? 1 : 0;
  K->tfN = K->tfMode ? tfN : 0;
  if (K->tfMode) MRT_REQUIRE(tfN >= 2 && tfN <= MRT_MAX_TF, "tfN=%d outside 2..%d", tfN, MRT_MAX_TF);
  K->skip = (M->skipEmpty && have_bits && K->tMode == 0) ? 1 : 0;
  K->nbx = (ldim[0] + 7) >> 3; K->nby = (ldim[1] + 7) >> 3; K->nbz = (ldim[2] + 7) >> 3;
  K->occ = nullptr; K->docc = nullptr;
  const int nt = mrt_tiles_x_(K->W) * mrt_tiles_y_(K->H);
  MRT_REQUIRE(tile_begin >= 0 && tile_begin <= tile_end && tile_end <= nt,
              "tile range [%d,%d) outside [0,%d]", tile_begin, tile_end, nt);
  K->tile_begin = tile_begin; K->tile_end = tile_end;
  {  // n*m >> 32 == n/d for m = floor(2^32/d)+1 whenever n*d < 2^32 (d = tiles_x, n = any tile id)
    const uint64_t d = (uint64_t)mrt_tiles_x_(K->W);
    K->tdiv_mul = (d > 1 && (uint64_t)nt * d < (1ull << 32)) ? (unsigned)((1ull << 32) / d + 1) : 0u;
  }
  return MRT_OK;
}

// ---------------------------------------------------------------- modality fold
static void blend_weights(const MrtParams* M, int C, float wgt[4], float* inv_wsum) {
  float wsum = 0.0f;
  for (int c = 0; c < 4; ++c) {
    const bool en = (c < C) && (M->volEnabled[c] != 0);
    wgt[c] = en ? M->volWeight[c] : 0.0f;
    if (en) wsum += M->volWeight[c];                       // brats_rt.slang:125-128
  }
  *inv_wsum = (wsum > 0.0f) ? 1.0f / wsum : 1.0f;          // :130
}
int mrt_fold_volume_f32(const MrtParams* params, const float* planar, int32_t C, float* folded, void* stream) {
  MRT_REQUIRE(params && planar && folded, "fold_volume: null pointer");
  const int X = (int)params->dims[0], Y = (int)params->dims[1], Z = (int)params->dims[2];
  if (int r = check_dims("fold_volume", C, X, Y, Z)) return r;
  float wgt[4], inv_wsum;
  blend_weights(params, C, wgt, &inv_wsum);
  cudaError_t e = mrt_launch_fold(planar, C, X, Y, Z, wgt, inv_wsum, folded, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "fold_volume");
}
int mrt_fold_volume_occupancy_quad_f32(const MrtParams* params, const float* planar, int32_t C, float* folded,
                                       void* quad, float* minmax, void* stream) {
  MRT_REQUIRE(params && planar && quad && minmax, "fold_volume_occupancy_quad: null pointer");
  const int X = (int)params->dims[0], Y = (int)params->dims[1], Z = (int)params->dims[2];
  if (int r = check_dims("fold_volume_occupancy_quad", C, X, Y, Z)) return r;
  int64_t qY, qZ;
  mrt_layout_e(1, 16, X, Y, Z, &qY, &qZ);
  MRT_REQUIRE((uint64_t)qZ * (uint64_t)Z < (1ull << 32), "fold_volume_occupancy_quad: more than 2^32 elements");
  float wgt[4], inv_wsum;
  blend_weights(params, C, wgt, &inv_wsum);
  cudaError_t e = mrt_launch_fold_occ_quad(planar, C, X, Y, Z, wgt, inv_wsum, folded, quad, minmax, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "fold_volume_occupancy_quad");
}
int mrt_fold_volume_occupancy_f32(const MrtParams* params, const float* planar, int32_t C, float* folded,
                                  float* minmax, void* stream) {
  MRT_REQUIRE(params && planar && folded && minmax, "fold_volume_occupancy: null pointer");
  const int X = (int)params->dims[0], Y = (int)params->dims[1], Z = (int)params->dims[2];
  if (int r = check_dims("fold_volume_occupancy", C, X, Y, Z)) return r;
  float wgt[4], inv_wsum;
  blend_weights(params, C, wgt, &inv_wsum);
  cudaError_t e = mrt_launch_fold_occ(planar, C, X, Y, Z, wgt, inv_wsum, folded, minmax, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "fold_volume_occupancy");
}
int mrt_unfold_grad_f32(const MrtParams* params, const float* dfolded, int32_t C, float* dplanar, void* stream) {
  MRT_REQUIRE(params && dfolded && dplanar, "unfold_grad: null pointer");
  const int X = (int)params->dims[0], Y = (int)params->dims[1], Z = (int)params->dims[2];
  if (int r = check_dims("unfold_grad", C, X, Y, Z)) return r;
  float wgt[4], inv_wsum;
  blend_weights(params, C, wgt, &inv_wsum);
  cudaError_t e = mrt_launch_unfold_grad(dfolded, C, X, Y, Z, wgt, inv_wsum, dplanar, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "unfold_grad");
}

// ---------------------------------------------------------------- occupancy
int32_t mrt_brick_count(int32_t X, int32_t Y, int32_t Z) {
  if (X < 1 || Y < 1 || Z < 1) return 0;
  return ((X + 7) >> 3) * ((Y + 7) >> 3) * ((Z + 7) >> 3);
}
size_t mrt_skip_levels_bytes(int32_t X, int32_t Y, int32_t Z) {
  const int32_t nb = mrt_brick_count(X, Y, Z);
  return nb > 0 ? mrt_levels_box_offset((size_t)nb) + 8 * sizeof(int32_t) : 0;
}
int mrt_build_occupancy(const void* packed, int32_t C, int32_t X, int32_t Y, int32_t Z, float* minmax, void* stream) {
  MRT_REQUIRE(packed && minmax, "build_occupancy: null pointer");
  if (int r = check_dims("build_occupancy", C, X, Y, Z)) return r;
  cudaError_t e = mrt_launch_build_occupancy(packed, mrt_packed_channels(C), X, Y, Z, minmax, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "build_occupancy");
}
int mrt_build_label_occupancy(const int32_t* labels, int32_t X, int32_t Y, int32_t Z, uint8_t* any, void* stream) {
  MRT_REQUIRE(labels && any, "build_label_occupancy: null pointer");
  if (int r = check_dims("build_label_occupancy", 1, X, Y, Z)) return r;
  cudaError_t e = mrt_launch_label_occupancy(labels, X, Y, Z, any, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "build_label_occupancy");
}
int mrt_classify_bricks(const MrtParams* params, const float* minmax, int32_t C, const float* tf, int32_t tfN,
                        const uint8_t* seg_any, const uint8_t* pred_any, uint8_t* skip_levels, int32_t flat,
                        void* stream) {
  MRT_REQUIRE(minmax && skip_levels, "classify_bricks: null pointer");
  KParams K;
  if (int r = derive(params, C, tfN, true, 0, 0, &K)) return r;
  MRT_REQUIRE(!K.tfMode || tf != nullptr, "classify_bricks: tfMode=1 needs tf");
  cudaError_t e = mrt_launch_classify(K, minmax, mrt_packed_channels(C), tf, seg_any, pred_any, skip_levels,
                                      flat != 0, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "classify_bricks");
}

// ---------------------------------------------------------------- forward / backward
int mrt_render_forward(const MrtParams* params, const void* packed, int32_t C, const float* tf, int32_t tfN,
                       const uint8_t* skip_levels, const int32_t* labels, const int32_t* preds,
                       float* out_rgba, float* out_T, int32_t* out_counts,
                       int32_t tile_begin, int32_t tile_end, void* stream) {
  MRT_REQUIRE(packed && out_rgba, "render_forward: null volume or output");
  KParams K;
  if (int r = derive(params, C, tfN, skip_levels != nullptr, tile_begin, tile_end, &K)) return r;
  MRT_REQUIRE(!K.tfMode || tf != nullptr, "render_forward: tfMode=1 needs tf");
  if (K.showSeg && !labels) K.showSeg = 0;                  // brats_viewer.py:423 (showSeg only with a buffer)
  if (K.showPred && !preds) K.showPred = 0;                 // brats_viewer.py:424
  cudaError_t e = mrt_launch_forward(K, nullptr, 1, mrt_packed_channels(C), packed, tf, skip_levels, labels, preds,
                                     out_rgba, out_T, out_counts, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "render_forward");
}

// ---------------------------------------------------------------- soft (learnable) occupancy
static int soft_occ_common(const char* who, const KParams& K) {
  if (K.half || K.shard || K.showSeg || K.showPred)
    return fail(MRT_ERR_UNSUPPORTED, "%s: fp32 unsharded volumes without overlays", who);
  return MRT_OK;
}
int mrt_render_forward_soft_occ(const MrtParams* params, const void* packed, int32_t C, const float* tf, int32_t tfN,
                                const uint8_t* skip_levels, const float* soft_occ, float* out_rgba,
                                int32_t tile_begin, int32_t tile_end, void* stream) {
  MRT_REQUIRE(packed && out_rgba && soft_occ, "render_forward_soft_occ: null pointer");
  KParams K;
  if (int r = derive(params, C, tfN, skip_levels != nullptr, tile_begin, tile_end, &K)) return r;
  MRT_REQUIRE(!K.tfMode || tf != nullptr, "render_forward_soft_occ: tfMode=1 needs tf");
  K.showSeg = K.showPred = 0;
  if (int r = soft_occ_common("render_forward_soft_occ", K)) return r;
  K.occ = soft_occ;
  cudaError_t e = mrt_launch_forward(K, nullptr, 1, mrt_packed_channels(C), packed, tf, skip_levels, nullptr, nullptr,
                                     out_rgba, nullptr, nullptr, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "render_forward_soft_occ");
}
int mrt_render_backward_soft_occ(const MrtParams* params, const void* packed, int32_t C, const float* tf, int32_t tfN,
                                 const float* soft_occ, const float* out_rgba, const float* dL_dout,
                                 void* dL_dvol, float* dL_dtf, float* dL_dsoft_occ, void* scratch,
                                 int32_t tile_begin, int32_t tile_end, void* stream) {
  MRT_REQUIRE(packed && out_rgba && dL_dout && scratch && soft_occ, "render_backward_soft_occ: null pointer");
  MRT_REQUIRE(dL_dvol || dL_dtf || dL_dsoft_occ, "render_backward_soft_occ: nothing to differentiate");
  KParams K;
  if (int r = derive(params, C, tfN, false, tile_begin, tile_end, &K)) return r;
  MRT_REQUIRE(!K.tfMode || tf != nullptr, "render_backward_soft_occ: tfMode=1 needs tf");
  K.showSeg = K.showPred = 0;
  if (int r = soft_occ_common("render_backward_soft_occ", K)) return r;
  K.occ = soft_occ; K.docc = dL_dsoft_occ;
  MrtBwdArgs A = {};
  A.tf = tf; A.out_rgba = out_rgba; A.dL_dout = dL_dout;
  A.dvol = dL_dvol; A.dtf = dL_dtf; A.scratch = scratch;
  cudaError_t e = mrt_launch_backward(K, nullptr, 1, mrt_packed_channels(C), packed, A, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "render_backward_soft_occ");
}

int mrt_render_forward_tma(const MrtParams* params, const void* packed, const float* tf, int32_t tfN,
                           const uint8_t* skip_levels, float* out_rgba, int32_t box_edge, int32_t tile,
                           uint64_t* stats, void* stream) {
  MRT_REQUIRE(params && packed && out_rgba && skip_levels, "render_forward_tma: null pointer");
  MRT_REQUIRE(!params->tfMode || tf, "render_forward_tma: tfMode=1 needs a LUT");
  KParams K;
  if (int r = derive(params, 1, tfN, true, 0, mrt_tile_count(params->imageSize[0], params->imageSize[1]), &K)) return r;
  MRT_REQUIRE(!K.half && !K.shard && K.tMode == 0 && K.gamma == 1.0f && !K.showSeg && !K.showPred && K.skip,
              "render_forward_tma: scalar fp32 single-channel volumes, indexed stepping, skipping on, no shards / overlays / gamma");
  MRT_REQUIRE((box_edge == 8 || box_edge == 16) && (tile == 8 || tile == 16), "render_forward_tma: box_edge and tile must be 8 or 16");
  cudaError_t e = mrt_launch_forward_tma(K, box_edge, tile, packed, tf, skip_levels, out_rgba, stats, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "render_forward_tma");
}
int mrt_render_forward_batch(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                             const void* packed, int32_t C, const float* tf, int32_t tfN,
                             const uint8_t* skip_levels, const int32_t* labels, const int32_t* preds,
                             float* out_rgba, float* out_T, int32_t* out_counts,
                             int32_t tile_begin, int32_t tile_end, void* stream) {
  MRT_REQUIRE(packed && out_rgba, "render_forward_batch: null volume or output");
  MRT_REQUIRE(cams != nullptr && nviews >= 1, "render_forward_batch: needs >= 1 camera");
  KParams K;
  if (int r = derive(params, C, tfN, skip_levels != nullptr, tile_begin, tile_end, &K)) return r;
  MRT_REQUIRE(!K.tfMode || tf != nullptr, "render_forward_batch: tfMode=1 needs tf");
  if (K.showSeg && !labels) K.showSeg = 0;
  if (K.showPred && !preds) K.showPred = 0;
  static_assert(sizeof(MrtCamera) == 16 * sizeof(float), "MrtCamera layout");
  cudaError_t e = cudaSuccess;
  float chunk[MRT_MAX_VIEWS * 12];
  const size_t npix = (size_t)K.W * K.H;
  for (int v0 = 0; v0 < nviews && e == cudaSuccess; v0 += MRT_MAX_VIEWS) {
    const int nv = (nviews - v0 < MRT_MAX_VIEWS) ? nviews - v0 : MRT_MAX_VIEWS;
    for (int v = 0; v < nv; ++v)
      for (int i = 0; i < 3; ++i) {
        const MrtCamera& c = cams[v0 + v];
        chunk[v * 12 + i] = c.eye[i]; chunk[v * 12 + 3 + i] = c.U[i];
        chunk[v * 12 + 6 + i] = c.V[i]; chunk[v * 12 + 9 + i] = c.W[i];
      }
    e = mrt_launch_forward(K, chunk, nv, mrt_packed_channels(C), packed, tf, skip_levels, labels, preds,
                           out_rgba + (size_t)v0 * npix * 4, out_T ? out_T + (size_t)v0 * npix : nullptr,
                           out_counts ? out_counts + (size_t)v0 * npix * 4 : nullptr, (cudaStream_t)stream);
  }
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "render_forward_batch");
}

static void pack_cams(const MrtCamera* cams, int n, float* out12) {
  for (int v = 0; v < n; ++v)
    for (int i = 0; i < 3; ++i) {
      const MrtCamera& c = cams[v];
      out12[v * 12 + i] = c.eye[i]; out12[v * 12 + 3 + i] = c.U[i];
      out12[v * 12 + 6 + i] = c.V[i]; out12[v * 12 + 9 + i] = c.W[i];
    }
}
int mrt_view_spans(const MrtParams* params, const MrtCamera* cams, int32_t nviews, int32_t C, const uint8_t* skip_levels,
                   int32_t* spans, void* stream) {
  MRT_REQUIRE(params && cams && skip_levels && spans && nviews >= 1, "view_spans: bad arguments");
  KParams K;
  MrtParams Pg = *params;
  Pg.tfMode = 0;                              // geometry only: the transfer function plays no role here
  if (int r = derive(&Pg, C, 0, true, 0, 0, &K)) return r;
  cudaError_t e = cudaSuccess;
  float chunk[MRT_MAX_VIEWS * 12];
  for (int v0 = 0; v0 < nviews && e == cudaSuccess; v0 += MRT_MAX_VIEWS) {
    const int nv = (nviews - v0 < MRT_MAX_VIEWS) ? nviews - v0 : MRT_MAX_VIEWS;
    pack_cams(cams + v0, nv, chunk);
    e = mrt_launch_view_spans(K, chunk, nv, skip_levels, spans + (size_t)v0 * 2 * mrt_tiles_y_(K.H), (cudaStream_t)stream);
  }
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "view_spans");
}
int mrt_render_forward_batch_sparse(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                                    const void* packed, int32_t C, const float* tf, int32_t tfN,
                                    const uint8_t* skip_levels, float* out_rgba, int32_t* spans,
                                    int32_t store_outside, void* stream) {
  MRT_REQUIRE(packed && out_rgba && spans && skip_levels, "render_forward_batch_sparse: null pointer");
  MRT_REQUIRE(cams != nullptr && nviews >= 1, "render_forward_batch_sparse: needs >= 1 camera");
  KParams K;
  const int W = params ? (int)params->imageSize[0] : 0, H = params ? (int)params->imageSize[1] : 0;
  if (int r = derive(params, C, tfN, true, 0, mrt_tile_count(W > 0 ? W : 1, H > 0 ? H : 1), &K)) return r;
  MRT_REQUIRE(K.skip, "render_forward_batch_sparse: needs skipEmpty=1 with indexed stepping");
  MRT_REQUIRE(K.gamma == 1.0f, "render_forward_batch_sparse: gamma != 1 takes the generic kernel, which does not cull");
  MRT_REQUIRE(!K.tfMode || tf != nullptr, "render_forward_batch_sparse: tfMode=1 needs tf");
  K.showSeg = K.showPred = 0;
  cudaError_t e = cudaSuccess;
  float chunk[MRT_MAX_VIEWS * 12];
  const size_t npix = (size_t)K.W * K.H;
  for (int v0 = 0; v0 < nviews && e == cudaSuccess; v0 += MRT_MAX_VIEWS) {
    const int nv = (nviews - v0 < MRT_MAX_VIEWS) ? nviews - v0 : MRT_MAX_VIEWS;
    pack_cams(cams + v0, nv, chunk);
    if (store_outside == 1) // single-GPU fast path: the spans are this call's own scratch, computed here
      e = mrt_launch_view_spans(K, chunk, nv, skip_levels, const_cast<int32_t*>(spans) + (size_t)v0 * 2 * mrt_tiles_y_(K.H),
                                (cudaStream_t)stream);
    if (e != cudaSuccess) break;
    e = mrt_launch_forward_sparse(K, chunk, nv, mrt_packed_channels(C), packed, tf, skip_levels,
                                  out_rgba + (size_t)v0 * npix * 4, spans + (size_t)v0 * 2 * mrt_tiles_y_(K.H), store_outside ? 1 : 0,
                                  (cudaStream_t)stream);
  }
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "render_forward_batch_sparse");
}
// One call = one frame-loop step of a device-resident renderer whose modality weights (or voxels) may
// have changed: blend + occupancy + quad layout, classify, spans, ONE batched march.  The four
// launches are queued back to back from C, so the host costs one call per step instead of three
// (measured: 0.72 -> 0.66 ms per cfg2 step would be the gain of removing the host path entirely).
static int refold_impl(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                       const float* planar, int32_t C, void* quad, float* minmax, uint8_t* skip_levels,
                       int32_t* spans, const float* tf, int32_t tfN, float* out_rgba,
                       float* const* view_out_dev, int32_t row_mod, int32_t row_rem,
                       void* ev_march_begin, void* ev_march_end, int32_t stage, void* stream) {
  MRT_REQUIRE(stage >= 0 && stage <= 2, "render_views_refold: stage %d unknown (0 all, 1 fold only, 2 after the fold)", stage);
  MRT_REQUIRE(params && planar && quad && minmax, "render_views_refold: null pointer");
  MRT_REQUIRE(stage == 1 || (cams && skip_levels && spans && (out_rgba || view_out_dev)), "render_views_refold: null pointer");
  MRT_REQUIRE(stage == 1 || nviews >= 1, "render_views_refold: needs >= 1 camera");
  if (!params->skipEmpty || params->tMode != 0 || params->gamma != 1.0f || params->volDtype != 0 || params->shardEnabled)
    return fail(MRT_ERR_UNSUPPORTED, "render_views_refold: needs skipEmpty=1, indexed stepping, gamma 1, an unsharded fp32 planar volume "
                                     "(use the separate calls otherwise)");
  if (stage != 2) {
    if (int r = mrt_fold_volume_occupancy_quad_f32(params, planar, C, nullptr, quad, minmax, stream)) return r;
    if (stage == 1) return MRT_OK;
  }
  MrtParams P = *params;                       // the folded field is rendered as ONE modality of weight 1, no overlays
  P.volEnabled[0] = 1; P.volEnabled[1] = P.volEnabled[2] = P.volEnabled[3] = 0;
  P.volWeight[0] = P.volWeight[1] = P.volWeight[2] = P.volWeight[3] = 1.0f;
  P.showSeg = P.showPred = 0;
  if (int r = mrt_classify_bricks(&P, minmax, 1, tf, tfN, nullptr, nullptr, skip_levels, 0, stream)) return r;
  P.volDtype = 3;
  if (ev_march_begin) {
    cudaError_t e = cudaEventRecord((cudaEvent_t)ev_march_begin, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "render_views_refold");
  }
  if (view_out_dev) {
    // distributed framebuffer: spans of every view (owners fill outside them), then this rank's tile rows of every
    // view stored straight into the owners' frames; tiles outside the spans are not stored
    MrtParams Ps = P; Ps.volDtype = 0;
    if (int r = mrt_view_spans(&Ps, cams, nviews, 1, skip_levels, spans, stream)) return r;
    if (int r = mrt_render_forward_batch_scatter(&P, cams, nviews, quad, 1, tf, tfN, skip_levels, view_out_dev, spans, 0,
                                                 row_mod, row_rem, stream)) return r;
  } else {
    // store_outside = 1: the call computes the spans into `spans` itself, then marches into dense frames
    if (int r = mrt_render_forward_batch_sparse(&P, cams, nviews, quad, 1, tf, tfN, skip_levels, out_rgba, spans, 1, stream)) return r;
  }
  if (ev_march_end) {
    cudaError_t e = cudaEventRecord((cudaEvent_t)ev_march_end, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "render_views_refold");
  }
  return MRT_OK;
}
int mrt_render_views_refold(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                            const float* planar, int32_t C, void* quad, float* minmax, uint8_t* skip_levels,
                            int32_t* spans, const float* tf, int32_t tfN, float* out_rgba,
                            void* ev_march_begin, void* ev_march_end, int32_t stage, void* stream) {
  return refold_impl(params, cams, nviews, planar, C, quad, minmax, skip_levels, spans, tf, tfN, out_rgba, nullptr, 0, 0,
                     ev_march_begin, ev_march_end, stage, stream);
}
int mrt_render_views_refold_scatter(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                                    const float* planar, int32_t C, void* quad, float* minmax, uint8_t* skip_levels,
                                    int32_t* spans, const float* tf, int32_t tfN, float* const* view_out_dev,
                                    int32_t row_mod, int32_t row_rem, int32_t stage, void* stream) {
  MRT_REQUIRE(stage == 1 || view_out_dev, "render_views_refold_scatter: null pointer");
  return refold_impl(params, cams, nviews, planar, C, quad, minmax, skip_levels, spans, tf, tfN, nullptr, view_out_dev,
                     row_mod, row_rem, nullptr, nullptr, stage, stream);
}
int mrt_render_forward_batch_scatter(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                                     const void* packed, int32_t C, const float* tf, int32_t tfN,
                                     const uint8_t* skip_levels, float* const* view_out_dev, const int32_t* spans,
                                     int32_t store_outside, int32_t row_mod, int32_t row_rem, void* stream) {
  MRT_REQUIRE(packed && view_out_dev && spans && skip_levels, "render_forward_batch_scatter: null pointer");
  MRT_REQUIRE(cams != nullptr && nviews >= 1, "render_forward_batch_scatter: needs >= 1 camera");
  MRT_REQUIRE(row_mod >= 0 && (row_mod <= 1 || (row_rem >= 0 && row_rem < row_mod)), "render_forward_batch_scatter: bad row partition %d/%d",
              row_rem, row_mod);
  KParams K;
  const int W = params ? (int)params->imageSize[0] : 0, H = params ? (int)params->imageSize[1] : 0;
  if (int r = derive(params, C, tfN, true, 0, mrt_tile_count(W > 0 ? W : 1, H > 0 ? H : 1), &K)) return r;
  MRT_REQUIRE(K.skip, "render_forward_batch_scatter: needs skipEmpty=1 with indexed stepping");
  MRT_REQUIRE(K.gamma == 1.0f, "render_forward_batch_scatter: gamma != 1 takes the generic kernel, which does not cull");
  MRT_REQUIRE(!K.tfMode || tf != nullptr, "render_forward_batch_scatter: tfMode=1 needs tf");
  K.showSeg = K.showPred = 0;
  cudaError_t e = cudaSuccess;
  float chunk[MRT_MAX_VIEWS * 12];
  for (int v0 = 0; v0 < nviews && e == cudaSuccess; v0 += MRT_MAX_VIEWS) {
    const int nv = (nviews - v0 < MRT_MAX_VIEWS) ? nviews - v0 : MRT_MAX_VIEWS;
    pack_cams(cams + v0, nv, chunk);
    e = mrt_launch_forward_sparse(K, chunk, nv, mrt_packed_channels(C), packed, tf, skip_levels, nullptr,
                                  spans + (size_t)v0 * 2 * mrt_tiles_y_(K.H), store_outside ? 1 : 0, (cudaStream_t)stream,
                                  view_out_dev + v0, row_mod, row_rem);
  }
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "render_forward_batch_scatter");
}
int mrt_fill_outside_spans(const MrtParams* params, const int32_t* spans, int32_t nviews, float* out_rgba, void* stream) {
  return mrt_fill_outside_spans_delta(params, spans, nullptr, nviews, out_rgba, stream);
}
int mrt_fill_outside_spans_delta(const MrtParams* params, const int32_t* spans, const int32_t* prev_spans, int32_t nviews,
                                 float* out_rgba, void* stream) {
  MRT_REQUIRE(params && spans && out_rgba && nviews >= 1, "fill_outside_spans: bad arguments");
  const int W = (int)params->imageSize[0], H = (int)params->imageSize[1];
  MRT_REQUIRE(W > 0 && H > 0 && W <= 65536 && H <= 65536, "fill_outside_spans: imageSize invalid");
  KParams K;
  memset(&K, 0, sizeof(K));
  K.W = W; K.H = H; K.tile_begin = 0; K.tile_end = mrt_tile_count(W, H);
  for (int i = 0; i < 3; ++i) K.bg[i] = params->bgColor[i];
  K.alphaMode = params->alphaMode ? 1 : 0; K.shard = params->shardEnabled ? 1 : 0;
  {
    const uint64_t d = (uint64_t)mrt_tiles_x_(W);
    K.tdiv_mul = (d > 1 && (uint64_t)K.tile_end * d < (1ull << 32)) ? (unsigned)((1ull << 32) / d + 1) : 0u;
  }
  cudaError_t e = mrt_launch_fill_outside(K, nviews, spans, prev_spans, out_rgba, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "fill_outside_spans");
}

// ---------------------------------------------------------------- checkpointed forward + backward
// Upper bound of any ray's slot count: the box diagonal (and the near/far window) over the step.
static int max_ray_slots(const KParams& K) {
  double d2 = 0.0;
  for (int i = 0; i < 3; ++i) { const double e = (double)K.vs[i] * K.dims[i]; d2 += e * e; }
  double len = sqrt(d2);
  if (K.farT > 0.0f) { const double w = (double)K.farT - fmax(0.0, (double)K.nearT); if (w < len) len = w > 0.0 ? w : 0.0; }
  double n = ceil(len / (double)K.dt) + 2.0;
  if (K.maxSteps > 0 && n > K.maxSteps) n = K.maxSteps;
  if (n > 2.0e9) n = 2.0e9;
  return (int)n;
}
int mrt_checkpoint_plan(const MrtParams* params, int32_t seg_slots_hint, int32_t* seg_slots, int32_t* nseg) {
  MRT_REQUIRE(params && seg_slots && nseg, "checkpoint_plan: null pointer");
  KParams K;
  MrtParams Pg = *params;
  Pg.tfMode = 0;
  if (int r = derive(&Pg, 1, 0, false, 0, 0, &K)) return r;
  const int nmax = max_ray_slots(K);
  int S = seg_slots_hint > 0 ? seg_slots_hint : 32;
  const int max_seg = 64;
  if ((nmax + S - 1) / S > max_seg) S = (((nmax + max_seg - 1) / max_seg) + 7) & ~7;
  *seg_slots = S;
  *nseg = (nmax + S - 1) / S < 1 ? 1 : (nmax + S - 1) / S;
  return MRT_OK;
}
size_t mrt_checkpoint_bytes(int32_t W, int32_t H, int32_t nviews, int32_t nseg) {
  if (W < 1 || H < 1 || nviews < 1 || nseg < 1) return 0;
  return (size_t)(nseg - 1) * nviews * W * H * 4 * sizeof(float);
}
int32_t mrt_half_tile_count(int32_t W, int32_t H) { return 2 * mrt_tiles_x_(W) * mrt_tiles_y_(H); }

int mrt_render_forward_ckpt(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                            const void* packed, int32_t C, const float* tf, int32_t tfN,
                            const uint8_t* skip_levels, const int32_t* labels, const int32_t* preds,
                            float* out_rgba, float* ckpt, int32_t seg_slots, int32_t nseg,
                            int32_t* k_end, int32_t* warp_kmax,
                            int32_t tile_begin, int32_t tile_end, void* stream) {
  MRT_REQUIRE(packed && out_rgba && k_end && warp_kmax, "render_forward_ckpt: null pointer");
  MRT_REQUIRE(seg_slots >= 1 && nseg >= 1 && (nseg == 1 || ckpt), "render_forward_ckpt: bad checkpoint plan");
  MRT_REQUIRE(cams == nullptr || (nviews >= 1 && nviews <= MRT_MAX_VIEWS), "render_forward_ckpt: nviews outside 1..%d", MRT_MAX_VIEWS);
  KParams K;
  if (int r = derive(params, C, tfN, skip_levels != nullptr, tile_begin, tile_end, &K)) return r;
  MRT_REQUIRE(!K.tfMode || tf != nullptr, "render_forward_ckpt: tfMode=1 needs tf");
  if (K.half || K.shard || K.tMode != 0 || K.gamma != 1.0f)
    return fail(MRT_ERR_UNSUPPORTED, "render_forward_ckpt: fp32 unsharded volumes, indexed stepping, gamma 1 only");
  MRT_REQUIRE((long long)nseg * seg_slots >= max_ray_slots(K), "render_forward_ckpt: %d segments of %d slots cannot hold %d-slot rays",
              nseg, seg_slots, max_ray_slots(K));
  if (K.showSeg && !labels) K.showSeg = 0;
  if (K.showPred && !preds) K.showPred = 0;
  float chunk[MRT_MAX_VIEWS * 12];
  if (cams) pack_cams(cams, nviews, chunk);
  cudaError_t e = mrt_launch_forward_ckpt(K, cams ? chunk : nullptr, nviews, mrt_packed_channels(C), packed, tf, skip_levels,
                                          labels, preds, out_rgba, ckpt, seg_slots, nseg, k_end, warp_kmax,
                                          (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "render_forward_ckpt");
}

size_t mrt_backward_scratch_bytes(int32_t W, int32_t H, int32_t nviews, int32_t tfN, int32_t nseg) {
  if (W < 1 || H < 1 || nviews < 1) return 0;
  return mrt_bwd_scratch_bytes(W, H, nviews, tfN < 2 ? 2 : tfN, nseg < 1 ? 1 : nseg);
}

int mrt_render_backward(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                        const void* packed, int32_t C, const float* tf, int32_t tfN,
                        const uint8_t* flat_levels, const float* minmax,
                        const int32_t* labels, const int32_t* preds, const float* out_rgba, const float* dL_dout,
                        const float* ckpt, int32_t seg_slots, int32_t nseg, const int32_t* k_end, const int32_t* warp_kmax,
                        void* dL_dvol, float* dL_dtf, void* scratch, float* dL_dray, uint64_t* stats,
                        int32_t tile_begin, int32_t tile_end, void* stream) {
  MRT_REQUIRE(packed && out_rgba && dL_dout && scratch, "render_backward: null pointer");
  MRT_REQUIRE(dL_dvol || dL_dtf || dL_dray, "render_backward: nothing to differentiate");
  MRT_REQUIRE(cams == nullptr || (nviews >= 1 && nviews <= MRT_MAX_VIEWS), "render_backward: nviews outside 1..%d", MRT_MAX_VIEWS);
  KParams K;
  if (int r = derive(params, C, tfN, flat_levels != nullptr && minmax != nullptr, tile_begin, tile_end, &K)) return r;
  if (K.half > 1) return fail(MRT_ERR_UNSUPPORTED, "render_backward: u8 / quad volumes are forward-only (fp32 and fp16 storage differentiate)");
  if (K.half && (K.tMode != 0 || K.gamma != 1.0f || K.showSeg || K.showPred))
    return fail(MRT_ERR_UNSUPPORTED, "render_backward: fp16 volumes need indexed stepping, gamma 1, no overlays");
  MRT_REQUIRE(!K.tfMode || tf != nullptr, "render_backward: tfMode=1 needs tf");
  const bool seg = k_end != nullptr || warp_kmax != nullptr || ckpt != nullptr;
  if (K.shard && (seg || dL_dray))
    return fail(MRT_ERR_UNSUPPORTED, "render_backward: sharded volumes take the whole-ray path (no checkpoints, no ray gradients)");
  if (seg) {
    MRT_REQUIRE(k_end && warp_kmax && seg_slots >= 1 && nseg >= 1 && (nseg == 1 || ckpt),
                "render_backward: the segmented path needs ckpt, k_end and warp_kmax of mrt_render_forward_ckpt");
    MRT_REQUIRE(K.tMode == 0 && K.gamma == 1.0f, "render_backward: checkpoints need indexed stepping and gamma 1");
  }
  if (K.showSeg && !labels) K.showSeg = 0;
  if (K.showPred && !preds) K.showPred = 0;
  MrtBwdArgs A = {};
  A.tf = tf; A.flat_levels = flat_levels; A.minmax = minmax; A.labels = labels; A.preds = preds;
  A.out_rgba = out_rgba; A.dL_dout = dL_dout;
  A.ck = seg ? (ckpt ? ckpt : out_rgba) : nullptr;   // nseg == 1: no checkpoint is ever read
  A.seg_slots = seg_slots; A.nseg = nseg; A.k_end = k_end; A.warp_kmax = warp_kmax;
  A.dvol = dL_dvol; A.dtf = dL_dtf; A.scratch = scratch; A.dray = dL_dray; A.stats = stats;
  float chunk[MRT_MAX_VIEWS * 12];
  if (cams) pack_cams(cams, nviews, chunk);
  cudaError_t e = mrt_launch_backward(K, cams ? chunk : nullptr, cams ? nviews : 1, mrt_packed_channels(C), packed, A,
                                      (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "render_backward");
}

// ---------------------------------------------------------------- one call per training step
#define MRT_TRAIN_MAX_PARTS 8
// The workspace is carved into 256-byte aligned pieces, in this order.
struct TrainCarve {
  size_t folded, minmax, skip, flat, ckpt, k_end, kmax, dfolded, priv, scratch, scratch_stride, mse, total;
  int seg_slots, nseg;
};
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static int train_carve(const MrtParams* params, int32_t nviews, int32_t tfN, TrainCarve* T) {
  const int X = (int)params->dims[0], Y = (int)params->dims[1], Z = (int)params->dims[2];
  if (int r = check_dims("train_step_mse", 1, X, Y, Z)) return r;
  const int W = (int)params->imageSize[0], H = (int)params->imageSize[1];
  MRT_REQUIRE(W >= 1 && H >= 1 && nviews >= 1 && nviews <= MRT_MAX_VIEWS, "train_step_mse: image %dx%d, %d views (1..%d)", W, H,
              nviews, MRT_MAX_VIEWS);
  int S = 0, nseg = 0;
  if (int r = mrt_checkpoint_plan(params, 0, &S, &nseg)) return r;
  T->seg_slots = S; T->nseg = nseg;
  const size_t vol_bytes = mrt_packed_volume_bytes(1, X, Y, Z);
  const size_t nb = (size_t)mrt_brick_count(X, Y, Z);
  const int ntf = (params->tfMode && tfN >= 2) ? tfN : 2;
  size_t o = 0;
  T->folded = o;  o += align256(vol_bytes);
  T->minmax = o;  o += align256(nb * 2 * sizeof(float));
  T->skip = o;    o += align256(mrt_skip_levels_bytes(X, Y, Z));
  T->flat = o;    o += align256(mrt_skip_levels_bytes(X, Y, Z));
  T->ckpt = o;    o += align256(mrt_checkpoint_bytes(W, H, nviews, nseg));
  T->k_end = o;   o += align256((size_t)nviews * W * H * sizeof(int32_t));
  T->kmax = o;    o += align256((size_t)nviews * mrt_half_tile_count(W, H) * sizeof(int32_t));
  T->dfolded = o; o += align256(vol_bytes);
  // the privatised dL/dtf copies are shared by all parts; every part has its own task list + counters
  T->priv = o;    o += align256(mrt_bwd_zeroed_scratch_bytes(ntf));
  T->scratch_stride = align256(mrt_bwd_scratch_bytes(W, H, nviews, ntf, nseg));
  T->scratch = o; o += T->scratch_stride * MRT_TRAIN_MAX_PARTS;
  T->mse = o;     o += align256(MRT_MSE_WORK_BYTES);
  T->total = o;
  return MRT_OK;
}
size_t mrt_train_step_workspace_bytes(const MrtParams* params, int32_t nviews, int32_t tfN) {
  TrainCarve T;
  if (!params || train_carve(params, nviews, tfN, &T) != MRT_OK) return 0;
  return T.total;
}

// A second stream per device carries what does not depend on the march — the flat-brick classification, the
// gradient-buffer clears — beside it, and the adjoint (so that the loss reduction can run beside THAT on the
// caller's stream).  It forks from and joins the caller's stream through events: the caller sees one
// stream-ordered call (and may capture it in a graph).  With the loss gradient formed per pixel inside the
// adjoint, the backward of an image part depends on no other part's forward, so the image can also be split
// into tile-range parts, part i differentiated while part i+1 marches (MRT_TRAIN_PARTS; measured, see below).
struct TrainSide { cudaStream_t s2; cudaEvent_t ev[MRT_TRAIN_MAX_PARTS + 2]; bool ok; };
static TrainSide* train_side() {
  static TrainSide side[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  TrainSide* S = &side[dev];
  if (!S->ok) {
    if (cudaStreamCreateWithFlags(&S->s2, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (int i = 0; i < MRT_TRAIN_MAX_PARTS + 2; ++i)
      if (cudaEventCreateWithFlags(&S->ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    S->ok = true;
  }
  return S;
}
#define MRT_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, "train_step_mse"); } while (0)
// Stage trace for tools/ (mrt.h, diagnostics): when armed, the NEXT mrt_train_step_mse records a timing
// event on the caller's stream after every stage; mrt_debug_train_trace(ms[6]) synchronises and returns the
// device time of fold | classify | march (all parts) | wait for the last adjoint + dL/dtf reduce | fold adjoint |
// loss + final join, as seen by the caller's stream.
static cudaEvent_t g_trace_ev[7];
static int g_trace_state = 0;            // 0 off, 1 armed, 2 recorded
#define MRT_TRACE(i) do { if (g_trace_state == 1) cudaEventRecord(g_trace_ev[i], s); } while (0)
int mrt_debug_train_trace(float* ms6) {
  if (ms6 == nullptr) {                  // arm
    if (!g_trace_ev[0]) for (int i = 0; i < 7; ++i) if (cudaEventCreate(&g_trace_ev[i]) != cudaSuccess) return MRT_ERR_CUDA;
    g_trace_state = 1;
    return MRT_OK;
  }
  if (g_trace_state != 2) return MRT_ERR_BAD_ARG;
  if (cudaEventSynchronize(g_trace_ev[6]) != cudaSuccess) return MRT_ERR_CUDA;
  for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&ms6[i], g_trace_ev[i], g_trace_ev[i + 1]);
  g_trace_state = 0;
  return MRT_OK;
}
static int train_env_int(const char* name, int dflt, int lo, int hi) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  const int x = atoi(v);
  return x < lo ? lo : (x > hi ? hi : x);
}
// Measured at cfg3 (gpurun_out/call_t4.txt, profiles/r02_train_step_parts.txt): 1 part 0.494 ms per step, 2 parts
// 0.627, 4 parts 0.644, 8 parts 0.953 — the persistent adjoint CTAs fill the register file (5 x 128 threads x 102
// registers per SM), so the next part's march finds no room beside them and the parts serialise with one tail each.
// The split stays as a knob (MRT_TRAIN_PARTS) for the measurement; the default is ONE part.
#ifndef MRT_TRAIN_PARTS
#define MRT_TRAIN_PARTS 1
#endif

int mrt_train_step_mse(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                       const float* planar, int32_t C, const float* tf, int32_t tfN, const float* target_rgba,
                       void* workspace, float* out_rgba, float* loss, float* dL_dplanar, float* dL_dtf, void* stream) {
  MRT_REQUIRE(params && planar && target_rgba && workspace && out_rgba, "train_step_mse: null pointer");
  MRT_REQUIRE(dL_dplanar || dL_dtf, "train_step_mse: nothing to differentiate");
  if (!params->skipEmpty || params->tMode != 0 || params->gamma != 1.0f || params->volDtype != 0 || params->shardEnabled ||
      params->showSeg || params->showPred)
    return fail(MRT_ERR_UNSUPPORTED, "train_step_mse: needs skipEmpty=1, indexed stepping, gamma 1, no overlays, an unsharded fp32 "
                                     "planar volume (use mrt_render_forward_ckpt + mrt_render_backward otherwise)");
  MRT_REQUIRE(!dL_dtf || (params->tfMode && tf), "train_step_mse: dL_dtf needs tfMode=1 and a LUT");
  const int V = cams ? nviews : 1;
  TrainCarve T;
  if (int r = train_carve(params, V, tfN, &T)) return r;
  TrainSide* S = train_side();
  if (!S) return fail(MRT_ERR_CUDA, "train_step_mse: cannot create the side stream");
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  float* folded = reinterpret_cast<float*>(ws + T.folded);
  float* minmax = reinterpret_cast<float*>(ws + T.minmax);
  uint8_t* skip = ws + T.skip; uint8_t* flat = ws + T.flat;
  float* ckpt = reinterpret_cast<float*>(ws + T.ckpt);
  int32_t* k_end = reinterpret_cast<int32_t*>(ws + T.k_end);
  int32_t* kmax = reinterpret_cast<int32_t*>(ws + T.kmax);
  float* dfolded = reinterpret_cast<float*>(ws + T.dfolded);
  const int X = (int)params->dims[0], Y = (int)params->dims[1], Z = (int)params->dims[2];
  const int W = (int)params->imageSize[0], H = (int)params->imageSize[1];
  const int ntf = (params->tfMode && tfN >= 2) ? tfN : 2;
  const int ntiles = mrt_tile_count(W, H);
  // measurement knobs (tools/time_train_step.py): no side stream at all; number of image parts; unfold always
  static const bool no_side = getenv("MRT_TRAIN_NO_SIDE") != nullptr;
  static const int parts_cfg = train_env_int("MRT_TRAIN_PARTS", MRT_TRAIN_PARTS, 1, MRT_TRAIN_MAX_PARTS);
  static const bool no_direct = getenv("MRT_TRAIN_NO_DIRECT") != nullptr;
  cudaStream_t s = (cudaStream_t)stream, s2 = no_side ? s : S->s2;
  int parts = parts_cfg;
  if (parts > ntiles) parts = ntiles;

  MrtParams P = *params;                       // the folded field is rendered as ONE modality of weight 1
  P.volEnabled[0] = 1; P.volEnabled[1] = P.volEnabled[2] = P.volEnabled[3] = 0;
  P.volWeight[0] = P.volWeight[1] = P.volWeight[2] = P.volWeight[3] = 1.0f;
  // A single modality whose fold is the identity: the adjoint reduces straight into the caller's planar
  // gradient ([Z][Y][X] pitches; reductions go to L2, which has no use for the sampler layout's bank skew) and the
  // fold's adjoint (a 2 x 67 MB layout copy at 256^3) disappears.
  float wgt[4], inv_wsum;
  blend_weights(params, C, wgt, &inv_wsum);
  const bool direct = dL_dplanar && C == 1 && wgt[0] * inv_wsum == 1.0f && !no_direct;

  // caller's stream: fold + occupancy -> classify -> checkpointing march, part by part
  MRT_TRACE(0);
  if (int r = mrt_fold_volume_occupancy_f32(params, planar, C, folded, minmax, s)) return r;
  MRT_CU(cudaEventRecord(S->ev[0], s));
  MRT_TRACE(1);
  if (int r = mrt_classify_bricks(&P, minmax, 1, tf, tfN, nullptr, nullptr, skip, 0, s)) return r;
  MRT_TRACE(2);
  // side stream: flat-brick levels, cleared gradient buffers
  MRT_CU(cudaStreamWaitEvent(s2, S->ev[0], 0));
  if (int r = mrt_classify_bricks(&P, minmax, 1, tf, tfN, nullptr, nullptr, flat, 1, s2)) return r;
  MRT_CU(cudaMemsetAsync(ws + T.priv, 0, T.scratch - T.priv, s2));                       // shared privatised dL/dtf copies
  MRT_CU(cudaMemset2DAsync(ws + T.scratch, T.scratch_stride, 0, 256, (size_t)parts, s2));   // every part's task counters
  if (dL_dplanar)
    MRT_CU(cudaMemsetAsync(direct ? dL_dplanar : dfolded, 0,
                           direct ? (size_t)X * Y * Z * sizeof(float) : mrt_packed_volume_bytes(1, X, Y, Z), s2));
  if (dL_dtf) MRT_CU(cudaMemsetAsync(dL_dtf, 0, (size_t)ntf * 4 * sizeof(float), s2));

  float chunk[MRT_MAX_VIEWS * 12];
  if (cams) pack_cams(cams, V, chunk);
  MrtBwdArgs A = {};
  A.tf = tf; A.flat_levels = flat; A.minmax = minmax;
  A.out_rgba = out_rgba; A.dL_dout = nullptr; A.target = target_rgba;
  A.gscale = 2.0f / (float)((size_t)V * W * H * 4);
  A.ck = T.nseg > 1 ? ckpt : out_rgba; A.seg_slots = T.seg_slots; A.nseg = T.nseg; A.k_end = k_end; A.warp_kmax = kmax;
  A.dvol = dL_dplanar ? (direct ? (void*)dL_dplanar : (void*)dfolded) : nullptr;
  if (direct) { A.grad_pitchY = (uint32_t)X; A.grad_pitchZ = (uint32_t)X * (uint32_t)Y; }
  A.dtf = dL_dtf; A.scratch_zeroed = 1; A.no_dtf_reduce = 1; A.shared_priv = ws + T.priv + 256;
  for (int i = 0; i < parts; ++i) {
    const int t0 = (int)((long long)ntiles * i / parts), t1 = (int)((long long)ntiles * (i + 1) / parts);
    KParams K;
    if (int r = derive(&P, 1, tfN, true, t0, t1, &K)) return r;
    MRT_REQUIRE(!K.tfMode || tf != nullptr, "train_step_mse: tfMode=1 needs tf");
    MRT_CU(mrt_launch_forward_ckpt(K, cams ? chunk : nullptr, V, 1, folded, tf, skip, nullptr, nullptr, out_rgba,
                                   T.nseg > 1 ? ckpt : nullptr, T.seg_slots, T.nseg, k_end, kmax, s, /*clear_aux=*/false));
    MRT_CU(cudaEventRecord(S->ev[1 + i], s));
    MRT_CU(cudaStreamWaitEvent(s2, S->ev[1 + i], 0));
    A.scratch = ws + T.scratch + T.scratch_stride * (size_t)i;
    MRT_CU(mrt_launch_backward(K, cams ? chunk : nullptr, V, 1, folded, A, s2));
  }
  MRT_TRACE(3);
  MRT_CU(cudaEventRecord(S->ev[MRT_TRAIN_MAX_PARTS + 1], s2));
  // caller's stream again: the loss beside the last adjoint, then the dL/dtf reduction and the fold's adjoint
  if (loss) MRT_CU(mrt_launch_mse(out_rgba, target_rgba, (size_t)V * W * H * 4, ws + T.mse, loss, s));
  MRT_CU(cudaStreamWaitEvent(s, S->ev[MRT_TRAIN_MAX_PARTS + 1], 0));
  if (dL_dtf) MRT_CU(mrt_launch_dtf_reduce(ws + T.priv + 256, 64, ntf, dL_dtf, s));
  MRT_TRACE(4);
  if (dL_dplanar && !direct)
    if (int r = mrt_unfold_grad_f32(params, dfolded, C, dL_dplanar, s)) return r;
  MRT_TRACE(5);
  MRT_TRACE(6);
  if (g_trace_state == 1) g_trace_state = 2;
  return MRT_OK;
}

// ---------------------------------------------------------------- adaptive (inverse-CDF) sampling
size_t mrt_adaptive_scratch_bytes(int32_t tfN) { return mrt_adaptive_scratch(tfN < 2 ? 2 : tfN); }

static int adaptive_common(const char* who, const MrtParams* params, int32_t C, int32_t tfN, int32_t K, int32_t J, float eps_w,
                           int32_t tile_begin, int32_t tile_end, KParams* Kp) {
  if (int r = derive(params, C, tfN, false, tile_begin, tile_end, Kp)) return r;
  MRT_REQUIRE(K >= 1 && K <= 64, "%s: n_coarse=%d outside 1..64", who, K);
  MRT_REQUIRE(J >= 1 && J <= 65536, "%s: n_fine=%d outside 1..65536", who, J);
  MRT_REQUIRE(eps_w > 0.0f, "%s: eps_w must be > 0 (the importance needs a floor for the CDF to be invertible)", who);
  if (Kp->half || Kp->shard || Kp->showSeg || Kp->showPred || Kp->gamma != 1.0f)
    return fail(MRT_ERR_UNSUPPORTED, "%s: fp32 unsharded volumes without overlays, gamma 1", who);
  return MRT_OK;
}
int mrt_render_adaptive_forward(const MrtParams* params, const void* packed, int32_t C, const float* tf, int32_t tfN,
                                int32_t n_coarse, int32_t n_fine, float eps_w, float* out_rgba,
                                int32_t tile_begin, int32_t tile_end, void* stream) {
  MRT_REQUIRE(packed && out_rgba, "render_adaptive_forward: null pointer");
  KParams K;
  if (int r = adaptive_common("render_adaptive_forward", params, C, tfN, n_coarse, n_fine, eps_w, tile_begin, tile_end, &K)) return r;
  MRT_REQUIRE(!K.tfMode || tf != nullptr, "render_adaptive_forward: tfMode=1 needs tf");
  cudaError_t e = mrt_launch_adaptive(K, n_coarse, n_fine, eps_w, mrt_packed_channels(C), packed, tf, out_rgba, nullptr, nullptr,
                                      nullptr, nullptr, false, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "render_adaptive_forward");
}
int mrt_render_adaptive_backward(const MrtParams* params, const void* packed, int32_t C, const float* tf, int32_t tfN,
                                 int32_t n_coarse, int32_t n_fine, float eps_w, const float* out_rgba, const float* dL_dout,
                                 void* dL_dvol, float* dL_dtf, void* scratch, int32_t tile_begin, int32_t tile_end, void* stream) {
  MRT_REQUIRE(packed && out_rgba && dL_dout, "render_adaptive_backward: null pointer");
  MRT_REQUIRE(dL_dvol || dL_dtf, "render_adaptive_backward: nothing to differentiate");
  MRT_REQUIRE(!dL_dtf || scratch, "render_adaptive_backward: dL_dtf needs the scratch buffer");
  KParams K;
  if (int r = adaptive_common("render_adaptive_backward", params, C, tfN, n_coarse, n_fine, eps_w, tile_begin, tile_end, &K)) return r;
  MRT_REQUIRE(!K.tfMode || tf != nullptr, "render_adaptive_backward: tfMode=1 needs tf");
  cudaError_t e = mrt_launch_adaptive(K, n_coarse, n_fine, eps_w, mrt_packed_channels(C), packed, tf, const_cast<float*>(out_rgba),
                                      dL_dout, dL_dvol, dL_dtf, scratch, true, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "render_adaptive_backward");
}

// ---------------------------------------------------------------- slab
int mrt_render_slab_u8(const MrtSlabParams* P, const uint8_t* vol_u8, float* out_rgba,
                       int32_t tile_begin, int32_t tile_end, void* stream) {
  MRT_REQUIRE(P && vol_u8 && out_rgba, "render_slab: null pointer");
  const int W = (int)P->imageSize[0], H = (int)P->imageSize[1];
  MRT_REQUIRE(W > 0 && H > 0, "render_slab: imageSize invalid");
  MRT_REQUIRE(P->volDim[0] >= 1 && P->volDim[1] >= 1 && P->volDim[2] >= 1, "render_slab: volDim invalid");
  MRT_REQUIRE((uint64_t)P->volDim[0] * P->volDim[1] * P->volDim[2] < (1ull << 32), "render_slab: volume too large");
  const int nt = mrt_tiles_x_(W) * mrt_tiles_y_(H);
  MRT_REQUIRE(tile_begin >= 0 && tile_begin <= tile_end && tile_end <= nt, "render_slab: tile range invalid");
  const float tan_half = (float)tan(0.5 * (double)P->fovY);
  cudaError_t e = mrt_launch_slab(*P, tan_half, vol_u8, out_rgba, tile_begin, tile_end, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "render_slab");
}

// ---------------------------------------------------------------- ingest
int mrt_decode_bc4(const uint8_t* blocks, int32_t W, int32_t H, int32_t D, uint8_t* out, void* stream) {
  MRT_REQUIRE(blocks && out && W > 0 && H > 0 && D > 0, "decode_bc4: bad arguments");
  MRT_REQUIRE(((uintptr_t)blocks & 7) == 0, "decode_bc4: blocks must be 8-byte aligned");
  cudaError_t e = mrt_launch_bc4(blocks, W, H, D, out, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "decode_bc4");
}
int mrt_u8_to_f32(const uint8_t* in, size_t n, float* out, void* stream) {
  MRT_REQUIRE(in && out, "u8_to_f32: null pointer");
  cudaError_t e = mrt_launch_u8_to_f32(in, n, out, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "u8_to_f32");
}
int mrt_normalize_f32(const float* in, size_t n, float vmin, float rng, float* out, void* stream) {
  MRT_REQUIRE(in && out, "normalize_f32: null pointer");
  MRT_REQUIRE(rng > 0.0f, "normalize_f32: rng must be > 0");
  cudaError_t e = mrt_launch_normalize(in, n, vmin, rng, out, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "normalize_f32");
}

// ---------------------------------------------------------------- INR inference (producer of gPreds)
int mrt_inr_predict(const float* mods_planar, int32_t M, int32_t X, int32_t Y, int32_t Z, const float* weights,
                    const int32_t* layer_dims, int32_t n_layers, int32_t fourier_freqs, int32_t* out_labels,
                    float* out_logits, int32_t impl, void* stream) {
  MRT_REQUIRE(mods_planar && weights && layer_dims && out_labels, "inr_predict: null pointer");
  MRT_REQUIRE(M >= 1 && M <= 8 && X >= 2 && Y >= 2 && Z >= 2, "inr_predict: bad volume shape");
  MRT_REQUIRE(n_layers >= 1 && n_layers <= 8, "inr_predict: n_layers=%d outside 1..8", n_layers);
  MRT_REQUIRE(fourier_freqs >= 0, "inr_predict: fourier_freqs must be >= 0");
  MRT_REQUIRE(layer_dims[0] == 3 + 6 * fourier_freqs + M, "inr_predict: input width %d != 3 + 6*%d + %d (model.py:21-23)",
              layer_dims[0], fourier_freqs, M);
  MRT_REQUIRE(layer_dims[0] <= 64, "inr_predict: input width %d > 64 is not supported", layer_dims[0]);
  for (int l = 1; l < n_layers; ++l)
    MRT_REQUIRE(layer_dims[l] >= 1 && layer_dims[l] <= 64, "inr_predict: hidden width %d outside 1..64", layer_dims[l]);
  MRT_REQUIRE(layer_dims[n_layers] >= 1 && layer_dims[n_layers] <= 8, "inr_predict: %d classes outside 1..8",
              layer_dims[n_layers]);
  MRT_REQUIRE(impl >= 0 && impl <= 2, "inr_predict: impl %d unknown (0 auto, 1 fp32 FFMA, 2 tensor cores)", impl);
  cudaError_t e = mrt_launch_inr(mods_planar, M, X, Y, Z, weights, layer_dims, n_layers, fourier_freqs, out_labels,
                                 out_logits, impl, (cudaStream_t)stream);
  if (e == cudaErrorNotSupported) return fail(MRT_ERR_UNSUPPORTED, "inr_predict: the network does not fit the tensor-core kernel");
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "inr_predict");
}

// ---------------------------------------------------------------- compositing / probe
int mrt_composite_over(const float* partials, int32_t K, const int32_t* order, size_t npix, const float* bg3,
                       int32_t alphaMode, float* out_rgba, void* stream) {
  MRT_REQUIRE(partials && order && out_rgba && bg3 && K >= 1, "composite_over: bad arguments");
  float* outs[1] = {out_rgba};
  cudaError_t e = mrt_launch_composite(partials, K, order, npix, bg3[0], bg3[1], bg3[2], alphaMode, outs, 1,
                                       (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "composite_over");
}
int mrt_composite_over_multi(const float* partials, int32_t K, const int32_t* order, size_t npix, const float* bg3,
                             int32_t alphaMode, float* const* outs, int32_t nouts, void* stream) {
  MRT_REQUIRE(partials && order && outs && bg3 && K >= 1, "composite_over_multi: bad arguments");
  MRT_REQUIRE(nouts >= 1 && nouts <= MRT_MAX_STRIPS, "composite_over_multi: nouts=%d outside 1..%d", nouts, MRT_MAX_STRIPS);
  for (int j = 0; j < nouts; ++j) MRT_REQUIRE(outs[j] != nullptr, "composite_over_multi: outs[%d] is null", j);
  cudaError_t e = mrt_launch_composite(partials, K, order, npix, bg3[0], bg3[1], bg3[2], alphaMode, outs, nouts,
                                       (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "composite_over_multi");
}
int mrt_render_forward_strips(const MrtParams* params, const void* packed, int32_t C, const float* tf, int32_t tfN,
                              const uint8_t* skip_levels, float* const* strip_out, int32_t nstrips, int32_t strip_rows,
                              int32_t tile_begin, int32_t tile_end, void* stream) {
  MRT_REQUIRE(packed && strip_out, "render_forward_strips: null volume or strip table");
  MRT_REQUIRE(nstrips >= 1 && nstrips <= MRT_MAX_STRIPS, "render_forward_strips: nstrips=%d outside 1..%d", nstrips,
              MRT_MAX_STRIPS);
  KParams K;
  if (int r = derive(params, C, tfN, skip_levels != nullptr, tile_begin, tile_end, &K)) return r;
  MRT_REQUIRE(!K.tfMode || tf != nullptr, "render_forward_strips: tfMode=1 needs tf");
  MRT_REQUIRE(strip_rows >= 1 && (int64_t)nstrips * strip_rows >= K.H, "render_forward_strips: %d strips of %d rows do not cover H=%d",
              nstrips, strip_rows, K.H);
  for (int j = 0; j < nstrips; ++j) MRT_REQUIRE(strip_out[j] != nullptr, "render_forward_strips: strip_out[%d] is null", j);
  K.showSeg = K.showPred = 0;
  cudaError_t e = mrt_launch_forward_strips(K, mrt_packed_channels(C), packed, tf, skip_levels, strip_out, nstrips,
                                            strip_rows, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "render_forward_strips");
}
int mrt_gather_probe(const void* buf, size_t bytes, size_t n_gathers, uint32_t seed, float* out, void* stream) {
  MRT_REQUIRE(buf && out, "gather_probe: null pointer");
  cudaError_t e = mrt_launch_gather_probe(buf, bytes, n_gathers, seed, out, (cudaStream_t)stream);
  return e == cudaSuccess ? MRT_OK : cuda_fail(e, "gather_probe");
}

// ---------------------------------------------------------------- host-buffer entry
#define MRT_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(MRT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); goto done; } } while (0)
#define MRT_CALL(call) do { rc = (call); if (rc != MRT_OK) goto done; } while (0)

int mrt_render_host(const MrtParams* params, const float* planar_host, int32_t C, const float* tf_host, int32_t tfN,
                    const int32_t* labels_host, const int32_t* preds_host, float* out_rgba_host) {
  MRT_REQUIRE(params && planar_host && out_rgba_host, "render_host: null pointer");
  const int X = (int)params->dims[0], Y = (int)params->dims[1], Z = (int)params->dims[2];
  if (int r = check_dims("render_host", C, X, Y, Z)) return r;
  const int W = (int)params->imageSize[0], H = (int)params->imageSize[1];
  MRT_REQUIRE(W > 0 && H > 0, "render_host: imageSize invalid");
  MRT_REQUIRE(!params->tfMode || (tf_host && tfN >= 2 && tfN <= MRT_MAX_TF), "render_host: tf invalid");
  const size_t nvox = (size_t)X * Y * Z, npix = (size_t)W * H;
  const int pc = mrt_packed_channels(C);
  const int nb = mrt_brick_count(X, Y, Z);
  int rc = MRT_OK;
  cudaStream_t st = nullptr;
  float *d_planar = nullptr, *d_tf = nullptr, *d_minmax = nullptr, *d_out = nullptr;
  void* d_packed = nullptr;
  int32_t *d_lab = nullptr, *d_pred = nullptr;
  uint8_t *d_seg_any = nullptr, *d_pred_any = nullptr;
  uint8_t* d_bits = nullptr;
  const size_t packed_bytes = mrt_packed_volume_bytes(C, X, Y, Z);
  MrtParams P = *params;
  const bool useSeg = P.showSeg && labels_host, usePred = P.showPred && preds_host;
  P.showSeg = useSeg; P.showPred = usePred;
  MRT_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  MRT_CUDA(cudaMallocAsync(&d_planar, nvox * C * sizeof(float), st));
  MRT_CUDA(cudaMemcpyAsync(d_planar, planar_host, nvox * C * sizeof(float), cudaMemcpyHostToDevice, st));
  MRT_CUDA(cudaMallocAsync(&d_packed, packed_bytes, st));
  MRT_CUDA(cudaMemsetAsync(d_packed, 0, packed_bytes, st));
  MRT_CALL(mrt_pack_volume_f32(d_planar, C, X, Y, Z, d_packed, st));
  if (P.tfMode) {
    MRT_CUDA(cudaMallocAsync(&d_tf, (size_t)tfN * 4 * sizeof(float), st));
    MRT_CUDA(cudaMemcpyAsync(d_tf, tf_host, (size_t)tfN * 4 * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  if (useSeg) {
    MRT_CUDA(cudaMallocAsync(&d_lab, nvox * sizeof(int32_t), st));
    MRT_CUDA(cudaMemcpyAsync(d_lab, labels_host, nvox * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  }
  if (usePred) {
    MRT_CUDA(cudaMallocAsync(&d_pred, nvox * sizeof(int32_t), st));
    MRT_CUDA(cudaMemcpyAsync(d_pred, preds_host, nvox * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  }
  MRT_CUDA(cudaMallocAsync(&d_out, npix * 4 * sizeof(float), st));
  if (P.skipEmpty && P.tMode == 0) {
    MRT_CUDA(cudaMallocAsync(&d_minmax, (size_t)nb * pc * 2 * sizeof(float), st));
    MRT_CUDA(cudaMallocAsync(&d_bits, mrt_skip_levels_bytes(X, Y, Z), st));
    MRT_CALL(mrt_build_occupancy(d_packed, C, X, Y, Z, d_minmax, st));
    if (useSeg) {
      MRT_CUDA(cudaMallocAsync(&d_seg_any, nb, st));
      MRT_CALL(mrt_build_label_occupancy(d_lab, X, Y, Z, d_seg_any, st));
    }
    if (usePred) {
      MRT_CUDA(cudaMallocAsync(&d_pred_any, nb, st));
      MRT_CALL(mrt_build_label_occupancy(d_pred, X, Y, Z, d_pred_any, st));
    }
    MRT_CALL(mrt_classify_bricks(&P, d_minmax, C, d_tf, tfN, d_seg_any, d_pred_any, d_bits, 0, st));
  }
  MRT_CALL(mrt_render_forward(&P, d_packed, C, d_tf, tfN, d_bits, d_lab, d_pred, d_out, nullptr, nullptr,
                              0, mrt_tile_count(W, H), st));
  MRT_CUDA(cudaMemcpyAsync(out_rgba_host, d_out, npix * 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
  MRT_CUDA(cudaStreamSynchronize(st));
done:
  if (st) {
    if (d_packed) cudaFreeAsync(d_packed, st);
    if (d_planar) cudaFreeAsync(d_planar, st);
    if (d_tf) cudaFreeAsync(d_tf, st);
    if (d_lab) cudaFreeAsync(d_lab, st);
    if (d_pred) cudaFreeAsync(d_pred, st);
    if (d_out) cudaFreeAsync(d_out, st);
    if (d_minmax) cudaFreeAsync(d_minmax, st);
    if (d_bits) cudaFreeAsync(d_bits, st);
    if (d_seg_any) cudaFreeAsync(d_seg_any, st);
    if (d_pred_any) cudaFreeAsync(d_pred_any, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
  }
  return rc;
}

}  // extern "C"
