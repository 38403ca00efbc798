/*
 * mrt.h — C ABI of libmrt.so, the B200-native volume ray-marcher.
 *
 * This is the drop-in boundary for the reference's compute-dispatch seam
 * (klukaszek/MRI-RayTracer).  Each entry point names the reference interface it
 * replaces (paths relative to the reference checkout):
 *
 *   kernel.dispatch(thread_count=[W,H,1], vars={gOutput, gIntensity0..3, gLabels,
 *                   gPreds, gParams})            inr/viewer/brats_viewer.py:431-442
 *     -> mrt_render_forward()                    (shader: inr/viewer/brats_rt.slang:85-168)
 *   kernel.dispatch(vars={gOutput, gParams, gVolumeU8})
 *                                                scripts/volumeRendering/app.py:350-358
 *     -> mrt_render_slab_u8()                    (shader: volume_render.slang:104-148)
 *   device.create_buffer + Buffer.copy_from_numpy per modality
 *                                                inr/viewer/brats_viewer.py:182-186,219-230
 *     -> mrt_pack_volume_f32()  (+ mrt_build_occupancy, which the reference lacks)
 *   docs/DifferentiableRendering.md:88-127  (maths only, no reference code)
 *     -> mrt_render_backward()
 *
 * Rules of the boundary:
 *   - plain pointers and sizes only; no torch / C++ types;
 *   - the CALLER owns every buffer (inputs, outputs, scratch); the library never
 *     allocates persistent device memory and never frees caller memory;
 *   - device entry points are stream-ordered on the `stream` argument
 *     (a cudaStream_t passed as void*) and never synchronise it;
 *     only the *_host entry points (host buffers in, host buffers out) synchronise;
 *   - return 0 on success, a negative MrtStatus on failure; the message is in
 *     mrt_last_error() (thread-local);
 *   - re-entrant, no global mutable state apart from the thread-local error string.
 */
#ifndef MRT_H_
#define MRT_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRT_VERSION 200  /* 0.2.0 */

typedef enum MrtStatus {
  MRT_OK = 0,
  MRT_ERR_BAD_ARG = -1,
  MRT_ERR_LAUNCH = -2,
  MRT_ERR_UNSUPPORTED = -3,
  MRT_ERR_CUDA = -4
} MrtStatus;

/* Brick edge (voxels) of the min/max occupancy grid. */
#define MRT_BRICK 8
/* Screen-space tile edge (pixels): [numthreads(8,8,1)], brats_rt.slang:86. */
#define MRT_TILE 8
/* Largest transfer-function LUT held in shared memory. */
#define MRT_MAX_TF 1024

/*
 * MrtParams — the operator's config block.
 *
 * The first 368 bytes are byte-for-byte the reference's `struct Params` constant
 * buffer (inr/viewer/brats_rt.slang:12-31; filled at brats_viewer.py:405-426), so a
 * host that already builds that block can copy it in unchanged.  The tail holds the
 * extensions SURVEY.md §8 adds; all-zero tail == reference behaviour.
 */
typedef struct MrtParams {
  uint32_t imageSize[2]; float fovY; float pad0;
  float eye[3]; float pad1;
  float U[3]; float pad2;
  float V[3]; float pad3;
  float W[3]; float pad4;
  float volMin[3]; float pad5;
  float voxelSize[3]; float pad6;
  uint32_t dims[3]; uint32_t pad7;          /* (X, Y, Z) */
  float stepSize; float nearT; float farT; float pad8;
  float bgColor[3]; float pad9;
  uint32_t volEnabled[4];
  float volWeight[4];
  float ww; float wl; float intensityAlpha; float padInt;
  float gamma; float gradBoost; float gradScale; float padTone;   /* gradBoost/gradScale: declared, never read */
  uint32_t showSeg; uint32_t showPred; uint32_t padFlags[2];
  float lutColorAlpha[8][4];
  /* ---- extensions (offset 368) ---- */
  uint32_t ortho;           /* 0 pinhole (makePrimary), 1 orthographic */
  float orthoHalfHeight;    /* world half-height of the ortho window */
  float ertThreshold;       /* 0 -> 0.01 (brats_rt.slang:117) */
  uint32_t maxSteps;        /* 0 -> unlimited ([MaxIters(1024)] is a hint) */
  uint32_t tMode;           /* 0 indexed t_k = t0 + k*dt ; 1 accumulate t += dt (reference, Q4) */
  uint32_t alphaMode;       /* 0 alpha = 1 (reference :167) ; 1 alpha = 1 - T */
  uint32_t skipEmpty;       /* 1 use the occupancy brick grid (needs `occupancy` != NULL) */
  uint32_t tfMode;          /* 0 reference window/level intensity TF (:132-140) ; 1 1D LUT tf[N][4] */
  /* ---- brick-sharded (sort-last) rendering, offset 400; all zero = off ----
   * `dims`, `volMin`, `voxelSize` keep describing the WHOLE volume (rays, slots and the global
   * dims-1.001 clamp are unchanged); the `packed` buffer holds only voxels
   * [shardLo, shardHi] inclusive (dims shardHi-shardLo+1, +1-voxel halo), and this call shades
   * exactly the slots whose trilinear base index lies in the cell range [shardLo, shardHi).
   * Output is the partial (r,g,b premultiplied WITHOUT background, a = T_local) for
   * mrt_composite_over.  Early termination acts on the shard-local transmittance. */
  uint32_t shardEnabled; uint32_t shardLo[3]; uint32_t shardHi[3];
  uint32_t volDtype;        /* 0: `packed` holds fp32 voxels; 1: fp16, 2: u8, 3: fp32 quads, 4: fp16 quads — single-channel (mrt_pack_volume_f16 / _u8 / _quad / _quad_f16), forward only */
} MrtParams;

/* One camera of a batch of views: the four camera rows of `struct Params`
 * (brats_rt.slang:15-18; filled per frame from OrbitalCamera.get_basis(), brats_viewer.py:400-410),
 * with the same 16-byte row padding. */
typedef struct MrtCamera {
  float eye[3]; float pad0;
  float U[3]; float pad1;
  float V[3]; float pad2;
  float W[3]; float pad3;
} MrtCamera;

/* Slab renderer params: `struct Params` of scripts/volumeRendering/volume_render.slang:9-21
 * (std140-like packing of that cbuffer is backend-defined; this is our own plain layout). */
typedef struct MrtSlabParams {
  uint32_t imageSize[2]; float fovY; float stepCount;
  float nearPlane; float farPlane; float pad0[2];
  float eye[3]; float padEye;
  float U[3]; float padU;
  float V[3]; float padV;
  float W[3]; float padW;
  uint32_t volDim[3]; uint32_t padDim;
} MrtSlabParams;

/* -------------------------------------------------------------------- misc */
int mrt_version(void);
const char* mrt_last_error(void);
/* sizeof(MrtParams) as compiled into the library (ABI self-check for bindings). */
size_t mrt_sizeof_params(void);
size_t mrt_sizeof_slab_params(void);
size_t mrt_sizeof_camera(void);
/* Views rendered by one kernel launch of mrt_render_forward_batch (larger batches are chunked). */
int32_t mrt_max_views_per_launch(void);

/* ------------------------------------------------ integer tile map (host)
 * The bit-exact integer contract of the dispatch geometry
 * (`numthreads(8,8,1)`, `thread_count=[W,H,1]`, brats_rt.slang:86-89;
 * brats_viewer.py:431-432): pixel (x,y) -> tile (x>>3, y>>3), tile id
 * ty*ceil(W/8)+tx, lane (y&7)*8+(x&7), linear pixel y*W+x. */
int32_t mrt_tiles_x(int32_t W);
int32_t mrt_tiles_y(int32_t H);
int32_t mrt_tile_count(int32_t W, int32_t H);
int32_t mrt_tile_of_pixel(int32_t x, int32_t y, int32_t W);
int32_t mrt_lane_of_pixel(int32_t x, int32_t y);
/* Contiguous tile range of `rank` out of `nranks`: [floor(r*T/R), floor((r+1)*T/R)). */
void mrt_rank_tile_range(int32_t ntiles, int32_t rank, int32_t nranks, int32_t* begin, int32_t* end);
/* Same map evaluated on the device for every pixel (bit-exactness check of the kernels'
 * own indexing): out_tile[y*W+x], out_lane[y*W+x]. Device pointers. */
int mrt_tile_index_map(int32_t W, int32_t H, int32_t* out_tile, int32_t* out_lane, void* stream);

/* ------------------------------------------------ volume layout
 * Planar [C][Z][Y][X] fp32 (the reference's one-buffer-per-modality flatten,
 * brats_viewer.py:64) -> the packed layout the sampler reads: channel-interleaved voxels
 *   C == 1 : float,  C == 2 : float2,  C == 3,4 : float4 (missing channel = 0)
 * so ONE vector load per trilinear corner fetches all modalities, at element index
 *   x + pitchY*y + pitchZ*z      (pitchY >= X, pitchZ >= pitchY*Y, in voxels).
 * The pitches are skewed (pitchY = S/4, pitchZ = S/2 modulo S, S = voxels per 128-byte line)
 * so that the 2x2x2 neighbourhoods a warp gathers fall into distinct L1 data banks; the
 * padding voxels are never read.  mrt_packed_layout is the single source of the pitches. */
void mrt_packed_layout(int32_t C, int32_t X, int32_t Y, int32_t Z, int64_t* pitchY, int64_t* pitchZ);
size_t mrt_packed_volume_bytes(int32_t C, int32_t X, int32_t Y, int32_t Z);
int mrt_pack_volume_f32(const float* planar, int32_t C, int32_t X, int32_t Y, int32_t Z,
                        void* packed, void* stream);
/* Inverse (used for dL/dvolume): packed -> planar [C][Z][Y][X]. */
int mrt_unpack_volume_f32(const void* packed, int32_t C, int32_t X, int32_t Y, int32_t Z,
                          float* planar, void* stream);

/* fp16 storage (BASELINE config 5: 2048^3 fp16, brick-sharded): single-channel planar
 * [Z][Y][X] half <-> packed half, same skew rule with 64 voxels per 128-byte line.  Render it
 * with C = 1 and params->volDtype = 1; the sampler widens each corner to fp32 and interpolates
 * in fp32 (results equal the fp32 path on the fp16-rounded values).  Forward only. */
size_t mrt_packed_volume_bytes_f16(int32_t X, int32_t Y, int32_t Z);
void mrt_packed_layout_f16(int32_t X, int32_t Y, int32_t Z, int64_t* pitchY, int64_t* pitchZ);
int mrt_pack_volume_f16(const void* planar_f16, int32_t X, int32_t Y, int32_t Z, void* packed, void* stream);
int mrt_unpack_volume_f16(const void* packed, int32_t X, int32_t Y, int32_t Z, void* planar_f16, void* stream);
int mrt_build_occupancy_f16(const void* packed, int32_t X, int32_t Y, int32_t Z, float* minmax, void* stream);

/* u8 storage, 1 byte per voxel (the reference's single-volume app uploads bytes — one per u32 lane,
 * scripts/volumeRendering/app.py:145-158 — and reads value = byte/255, volume_render.slang:33-38):
 * planar [Z][Y][X] uint8 -> packed uint8 with 128 voxels per 128-byte line.  Render it with C = 1 and
 * params->volDtype = 2 through mrt_render_forward(_batch): the same marcher (occupancy skipping,
 * TF, ERT) gathering 8 B per sample; the image equals the fp32 path on byte/255 to fp32 rounding.
 * mrt_build_occupancy_u8 returns min/max already divided by 255.  Forward only. */
size_t mrt_packed_volume_bytes_u8(int32_t X, int32_t Y, int32_t Z);
int mrt_pack_volume_u8(const uint8_t* planar_u8, int32_t X, int32_t Y, int32_t Z, void* packed, void* stream);
int mrt_build_occupancy_u8(const void* packed, int32_t X, int32_t Y, int32_t Z, float* minmax, void* stream);

/* "quad" sampler layout of a single-channel fp32 volume: element (x,y,z) = the four voxels (x,y)
 * (x+1,y) (x,y+1) (x+1,y+1) of slice z (neighbours clamped at the far faces), 16 B per voxel.  A
 * trilinear footprint (brats_rt.slang:60-76) is then two 16-byte loads instead of eight 4-byte ones:
 * the march issues a quarter of the load instructions for 4x the bytes.  Built from the packed C = 1
 * layout (mrt_pack_volume_f32 with C = 1, or the folded volume); the occupancy grid is the one of that
 * source.  Render it with C = 1 and params->volDtype = 3; the image is bit-identical to volDtype 0.
 * Forward only. */
size_t mrt_packed_volume_bytes_quad(int32_t X, int32_t Y, int32_t Z);
int mrt_pack_volume_quad(const float* packed1, int32_t X, int32_t Y, int32_t Z, void* quad, void* stream);
/* The same over fp16 voxels (8 B per element), from the packed fp16 layout of mrt_pack_volume_f16;
 * params->volDtype = 4.  Bit-identical to volDtype 1. */
size_t mrt_packed_volume_bytes_quad_f16(int32_t X, int32_t Y, int32_t Z);
int mrt_pack_volume_quad_f16(const void* packed_f16, int32_t X, int32_t Y, int32_t Z, void* quad, void* stream);

/* ------------------------------------------------ modality fold
 * The modality blend v = sum_c w_c s_c / wSum (brats_rt.slang:123-130) is linear and commutes
 * with trilinear interpolation.  mrt_fold_volume_f32 evaluates it ONCE per voxel for the
 * (volEnabled, volWeight) in `params`, from planar [C][Z][Y][X] fp32 into a single-channel
 * volume in the packed C=1 layout (size mrt_packed_volume_bytes(1,X,Y,Z)); rendering that
 * volume with C=1, volEnabled=(1,0,0,0), volWeight=(1,..) gives the same image (to fp32
 * rounding) with 8 scalar loads per sample instead of 8 float4 loads.  Re-fold when the
 * weights change.  mrt_unfold_grad_f32 is the adjoint: dL/dplanar[c] = w_c/wSum * dL/dfolded. */
int mrt_fold_volume_f32(const MrtParams* params, const float* planar, int32_t C, float* folded, void* stream);
int mrt_unfold_grad_f32(const MrtParams* params, const float* dfolded, int32_t C, float* dplanar, void* stream);
/* mrt_fold_volume_f32 + mrt_build_occupancy(folded, C=1) fused into one pass over the planar
 * volume: `minmax` is float2[mrt_brick_count] of the FOLDED field.  Same outputs, bit for bit. */
int mrt_fold_volume_occupancy_f32(const MrtParams* params, const float* planar, int32_t C, float* folded,
                                  float* minmax, void* stream);
/* The same pass writing the march's quad layout (mrt_pack_volume_quad) directly: `quad`
 * (mrt_packed_volume_bytes_quad bytes) and `minmax` are filled; the scalar folded volume only when
 * `folded` is non-NULL.  One launch instead of fold+occupancy followed by the quad pack. */
int mrt_fold_volume_occupancy_quad_f32(const MrtParams* params, const float* planar, int32_t C, float* folded,
                                       void* quad, float* minmax, void* stream);

/* ------------------------------------------------ occupancy brick grid
 * (new relative to the reference; must never change the image.)
 * Grid of ceil(dim/8)^3 bricks; brick b holds per-channel (min,max) over voxels
 * [8b, 8b+8] (inclusive: the whole trilinear/nearest footprint of any sample whose
 * base index lies in the brick).  minmax: float2[nbricks][Cp] with Cp = packed channel
 * count (1,2,4).  label_any: optional uint8[nbricks], 1 if any label in 1..7. */
int32_t mrt_brick_count(int32_t X, int32_t Y, int32_t Z);
int mrt_build_occupancy(const void* packed, int32_t C, int32_t X, int32_t Y, int32_t Z,
                        float* minmax, void* stream);
int mrt_build_label_occupancy(const int32_t* labels, int32_t X, int32_t Y, int32_t Z,
                              uint8_t* label_any, void* stream);
/* Per-frame classification into a skip-level byte per brick:
 *   0 : active (some sample in the brick may contribute);
 *   l in 1..4 : the aligned cell of 2^(l-1) bricks per axis (8,16,32,64 voxels) that contains
 *               this brick is provably empty under (params, tf, labels): every sample slot
 *               whose trilinear base index lies in it is a no-op, so the march may leap to
 *               the cell's exit.
 *   0x80 | l, l in 2..4 (flat == 0 only) : active, and so is every brick of the aligned cell of
 *               2^(l-1) bricks per axis around it: the march shades through the whole cell on one
 *               look-up.  Consumers that only ask "empty?" test (level & 0x80) == 0 && level != 0.
 *   flat != 0 (levels for the BACKWARD): a cell additionally has to be flat — every voxel in it
 *               holds one and the same value — so that all slots inside share one TF bin and
 *               one dL/dsigma, which mrt_render_backward adds in closed form.
 *
 * `skip_levels` is a buffer of mrt_skip_levels_bytes(X,Y,Z) bytes: uint8[nbricks] followed, 16-byte
 * aligned, by int32[8] holding the bounding box of the active bricks, which the march uses to cull
 * rays / CTAs that cannot touch an active brick before any exact ray set-up and to clip every
 * ray's slot range. */
size_t mrt_skip_levels_bytes(int32_t X, int32_t Y, int32_t Z);
int mrt_classify_bricks(const MrtParams* params, const float* minmax, int32_t C,
                        const float* tf, int32_t tfN,
                        const uint8_t* seg_any, const uint8_t* pred_any,
                        uint8_t* skip_levels, int32_t flat, void* stream);

/* ------------------------------------------------ forward
 * Renders tiles [tile_begin, tile_end) of the [H][W] image (tile ids as above).
 *   packed      : packed volume (see mrt_pack_volume_f32), C = logical channel count (1..4)
 *   tf, tfN     : LUT [tfN][4] (r,g,b,sigma) fp32, used when params->tfMode == 1
 *   skip_levels : the mrt_skip_levels_bytes() buffer filled by mrt_classify_bricks, or NULL (then
 *                 skipEmpty is ignored)
 *   labels/preds: optional int32 [Z][Y][X] (gLabels / gPreds), used when showSeg / showPred
 *   out_rgba    : float4 [H][W]  (row 0 = top)
 *   out_T       : optional float [H][W], final transmittance
 *   out_counts  : optional int32 [H][W][4] = (n_clip, n_taken, n_evaluated, n_segments)
 */
int mrt_render_forward(const MrtParams* params, const void* packed, int32_t C,
                       const float* tf, int32_t tfN, const uint8_t* skip_levels,
                       const int32_t* labels, const int32_t* preds,
                       float* out_rgba, float* out_T, int32_t* out_counts,
                       int32_t tile_begin, int32_t tile_end, void* stream);

/* The staged-brick variant of the single-view march (forward_tma.cu): one CTA per `tile`x`tile`
 * pixel block (8 or 16) streams the boxes of `box_edge`^3 voxels (8 or 16, +1 halo) its ray bundle
 * crosses through a double-buffered shared-memory stage with 3-D TMA box loads and samples them from
 * shared memory; same image as mrt_render_forward (same sampler arithmetic, brats_rt.slang:60-76).
 * Scalar fp32 single-channel `packed` (C = 1 layout, e.g. the folded volume), params->skipEmpty set and
 * `skip_levels` from mrt_classify_bricks(flat = 0), indexed stepping, no shards / overlays / gamma.
 * `stats` (optional, device, 4 x uint64, zeroed by the caller): slots shaded from the stage, slots
 * finished by direct gathers, boxes staged, lists that overflowed.  MEASURED SLOWER than the direct
 * gathers of mrt_render_forward at every configuration: kept selectable, not the default. */
int mrt_render_forward_tma(const MrtParams* params, const void* packed, const float* tf, int32_t tfN,
                           const uint8_t* skip_levels, float* out_rgba, int32_t box_edge, int32_t tile,
                           uint64_t* stats, void* stream);

/* A batch of views of ONE volume under ONE parameter block: what the reference's frame loop
 * does with `nviews` successive dispatches whose params differ only in (eye, U, V, W)
 * (brats_viewer.py:400-442).  Here it is one launch per <= mrt_max_views_per_launch() views
 * (grid.y = view), so the long central rays of one view overlap the short border rays of the
 * next.  `cams` is a HOST array; params->eye/U/V/W are ignored.  Outputs are contiguous
 * [nviews][H][W](...) device buffers; view v of the batch is bit-identical to
 * mrt_render_forward with cams[v] copied into params. */
int mrt_render_forward_batch(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                             const void* packed, int32_t C,
                             const float* tf, int32_t tfN, const uint8_t* skip_levels,
                             const int32_t* labels, const int32_t* preds,
                             float* out_rgba, float* out_T, int32_t* out_counts,
                             int32_t tile_begin, int32_t tile_end, void* stream);

/* Sparse variant for a framebuffer that lives on ANOTHER GPU (image-space gather through peer
 * memory).  mrt_view_spans computes, per view and per tile row (8 pixel rows), the inclusive pixel
 * span (x0,x1) outside which every ray certainly misses every ACTIVE BRICK: the union over the
 * bricks mrt_classify_bricks left active of the bounding rectangles of their projected boxes,
 * rounded outward (empty: x0 > x1) — much tighter than the hull of the bricks' bounding box (the bench's
 * head: 3 200 tiles per 1024^2 view instead of 5 500).  It is
 * deterministic in its inputs, so the sender and the owner of the image compute identical spans
 * independently.  mrt_render_forward_batch_sparse does NOT store the tiles outside its views'
 * spans; the owner of the image calls mrt_fill_outside_spans on its local copy (any time: the two
 * write disjoint pixels).  Together == mrt_render_forward_batch, bit for bit, with the background
 * (typically > half of the frame) never crossing NVLink and its fill off the critical path.
 * Whole image only, skipping required.  `spans`: device int32[nviews][mrt_tiles_y(H)][2].
 * store_outside != 0: the march DOES store the background of the outside tiles itself — the
 * single-GPU fast path: the spans then only replace the per-ray box test of mrt_render_forward_batch
 * by one load and two compares per warp (same image, bit for bit).  store_outside == 1: `spans` is
 * pure scratch, the call computes it itself (mrt_view_spans) before the march; == 2: `spans` already
 * holds mrt_view_spans of the same arguments.
 * The camera basis (U, V, W) may be any non-degenerate basis; a degenerate one disables the cull. */
int mrt_view_spans(const MrtParams* params, const MrtCamera* cams, int32_t nviews, int32_t C,
                   const uint8_t* skip_levels, int32_t* spans, void* stream);
int mrt_render_forward_batch_sparse(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                                    const void* packed, int32_t C, const float* tf, int32_t tfN,
                                    const uint8_t* skip_levels, float* out_rgba, int32_t* spans,
                                    int32_t store_outside, void* stream);
int mrt_fill_outside_spans(const MrtParams* params, const int32_t* spans, int32_t nviews,
                           float* out_rgba, void* stream);
/* Delta form for a framebuffer that is reused batch after batch: `prev_spans` = the spans of the batch
 * that last wrote `out_rgba` (after which the image held the background everywhere outside them, same
 * background colour): only tiles inside the previous spans and outside the new ones are written — a few
 * MB instead of the > half frame per view when the camera moves in small steps.  prev_spans == NULL:
 * mrt_fill_outside_spans. */
int mrt_fill_outside_spans_delta(const MrtParams* params, const int32_t* spans, const int32_t* prev_spans,
                                 int32_t nviews, float* out_rgba, void* stream);
/* Image-space partition across GPUs with the framebuffer gather done by the march's own stores
 * (north_star: "partitioned by image-space tiles with the volume replicated, framebuffer gathered
 * over NCCL/NVLink"): like mrt_render_forward_batch_sparse, but
 *   view_out_dev : DEVICE array of nviews pointers, the [H][W] float4 frame of each view — local or
 *                  peer-mapped memory of whichever GPU owns that view (frames striped over the ranks
 *                  spread the ingress instead of funnelling it into one root);
 *   row_mod, row_rem : render only tile rows ty with ty % row_mod == row_rem (row_mod <= 1: all) —
 *                  rank r of R passes (R, r); rows are interleaved because the object sits mid-image.
 * `spans` from mrt_view_spans of the same cameras (every rank computes the same integers).  The
 * union over row_rem = 0..row_mod-1 equals mrt_render_forward_batch, bit for bit. */
int mrt_render_forward_batch_scatter(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                                     const void* packed, int32_t C, const float* tf, int32_t tfN,
                                     const uint8_t* skip_levels, float* const* view_out_dev, const int32_t* spans,
                                     int32_t store_outside, int32_t row_mod, int32_t row_rem, void* stream);

/* ------------------------------------------------ training forward (checkpoints)
 * The forward of differentiable rendering.  docs/DifferentiableRendering.md:213 ("checkpointing")
 * is the only hint the reference gives on storage; here the march records, per ray, its colour and
 * transmittance (C, T) before every slot c*seg_slots and its end slot, so that the backward can
 * differentiate a ray as independent segments (one warp task each) instead of one serial chain.
 *   mrt_checkpoint_plan : slots per segment (hint 0 -> 32) and segment count for params' geometry
 *                         (segment count is capped at 64 by growing the segment)
 *   ckpt      : device float4 [nseg-1][nviews][H][W], mrt_checkpoint_bytes() bytes (NULL if nseg == 1)
 *   k_end     : device int32 [nviews][H][W], every ray's end slot (the oracle's n_taken)
 *   warp_kmax : device int32 [nviews][mrt_half_tile_count(W,H)], largest k_end per 8x4 half tile
 * `cams` NULL: one view with the camera in params.  nviews <= mrt_max_views_per_launch().
 * Image identical to mrt_render_forward(_batch), bit for bit.  fp32 unsharded volumes, tMode 0,
 * gamma 1 (otherwise use mrt_render_forward and the unsegmented backward). */
int mrt_checkpoint_plan(const MrtParams* params, int32_t seg_slots_hint, int32_t* seg_slots, int32_t* nseg);
size_t mrt_checkpoint_bytes(int32_t W, int32_t H, int32_t nviews, int32_t nseg);
int32_t mrt_half_tile_count(int32_t W, int32_t H);
int mrt_render_forward_ckpt(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                            const void* packed, int32_t C, const float* tf, int32_t tfN,
                            const uint8_t* skip_levels, const int32_t* labels, const int32_t* preds,
                            float* out_rgba, float* ckpt, int32_t seg_slots, int32_t nseg,
                            int32_t* k_end, int32_t* warp_kmax,
                            int32_t tile_begin, int32_t tile_end, void* stream);

/* ------------------------------------------------ one call per frame-loop step
 * The reference's frame loop re-reads its UI state every frame (inr/viewer/brats_viewer.py:400-442: weights,
 * window, cameras) and dispatches.  mrt_render_views_refold is that step for a device-resident planar
 * volume in ONE call: mrt_fold_volume_occupancy_quad_f32 (blend + occupancy + quad layout) ->
 * mrt_classify_bricks -> spans + mrt_render_forward_batch_sparse (dense frames [nviews][H][W][4]), all queued
 * on `stream` back to back.  Frames are bit-identical to the separate calls.  Caller-owned scratch:
 *   quad        mrt_packed_volume_bytes_quad(X,Y,Z) bytes     minmax  mrt_brick_count * 2 floats
 *   skip_levels mrt_skip_levels_bytes(X,Y,Z) bytes            spans   nviews * mrt_tiles_y(H) * 2 int32
 * ev_march_begin / ev_march_end: optional cudaEvent_t recorded around the spans + march launches (a host
 * that wants the march's own device time); NULL to skip.  stage: 0 = the whole step; 1 = only the fold pass
 * (needs no cameras: a host can queue it first and prepare its camera array while the GPU folds), 2 = the
 * rest of a step whose stage 1 was queued on the same stream.  Needs skipEmpty = 1, tMode = 0, gamma = 1, no
 * overlays, volDtype = 0 (MRT_ERR_UNSUPPORTED otherwise: use the separate calls). */
int mrt_render_views_refold(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                            const float* planar, int32_t C, void* quad, float* minmax, uint8_t* skip_levels,
                            int32_t* spans, const float* tf, int32_t tfN, float* out_rgba,
                            void* ev_march_begin, void* ev_march_end, int32_t stage, void* stream);

/* The same step for the distributed framebuffer (mrt_render_forward_batch_scatter's arguments): spans of all
 * views into `spans`, this rank's tile rows (ty % row_mod == row_rem) of every view stored to view_out_dev[v];
 * tiles outside the spans are not stored (the frame's owner fills them, mrt_fill_outside_spans). */
int mrt_render_views_refold_scatter(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                                    const float* planar, int32_t C, void* quad, float* minmax, uint8_t* skip_levels,
                                    int32_t* spans, const float* tf, int32_t tfN, float* const* view_out_dev,
                                    int32_t row_mod, int32_t row_rem, int32_t stage, void* stream);

/* ------------------------------------------------ backward
 * Adjoint of the forward w.r.t. the volume and the transfer function
 * (docs/DifferentiableRendering.md:88-127).  Recomputes the forward per ray (segment).
 *   cams, nviews: like mrt_render_forward_ckpt; all per-view arrays are [nviews][H][W](...)
 *   flat_levels: optional uint8[nbricks] from mrt_classify_bricks(..., flat=1) together with
 *                `minmax` (from mrt_build_occupancy): flat-empty cells are leapt with their exact
 *                closed-form contribution to dL/dtf (single-channel layouts only; else ignored)
 *   out_rgba   : the forward's output (needed for the suffix sums)
 *   dL_dout    : float4 [nviews][H][W]
 *   ckpt, seg_slots, nseg, k_end, warp_kmax : the outputs of mrt_render_forward_ckpt for the SAME
 *                arguments -> segment-parallel backward; all NULL / 0 -> one task per half tile that
 *                walks whole rays (any tMode / gamma)
 *   dL_dvol    : packed layout, same shape as `packed`, ACCUMULATED into (caller zeroes).  fp16 storage
 *                (volDtype 1): the gradient is fp32 with the fp16 layout's element pitches
 *                (mrt_packed_layout_f16; pitchZ*Z floats)
 *   sharded volumes (params->shardEnabled; sort-last sub-boxes): whole-ray path only (ckpt... NULL,
 *                dL_dray NULL).  `out_rgba` / `dL_dout` are the shard's PARTIAL (premultiplied rgb
 *                without background, T_local) and its gradient: .w of dL_dout is dL/dT_local.
 *                A slot is differentiated iff its base cell is owned, as in the forward, so the
 *                shards' gradients add up to the unsharded gradient.  u8 / quad layouts: forward only.
 *   dL_dtf     : float [tfN][4] (tfMode 1) or float[2][4] (tfMode 0: entry [1][3] is
 *                dL/d intensityAlpha), ACCUMULATED into (caller zeroes)
 *   scratch    : device scratch of mrt_backward_scratch_bytes(W,H,nviews,tfN,nseg) bytes (task list,
 *                counters, privatised dL/dtf accumulators; initialised by the call)
 *   dL_dray    : optional float [nviews][H][W][6] = (dL/do, dL/dd) per ray, world units, with the sample
 *                times t_i held fixed (docs/DifferentiableRendering.md section 9, :172-188:
 *                dL/do = sum_i dL/dx_i, dL/dd = sum_i t_i dL/dx_i; dL/dx_i through the trilinear
 *                gradient of section 6).  ACCUMULATED into (caller zeroes).
 *   stats      : optional device uint64[2], ACCUMULATED: sample slots shaded, warp tasks executed
 */
size_t mrt_backward_scratch_bytes(int32_t W, int32_t H, int32_t nviews, int32_t tfN, int32_t nseg);
int mrt_render_backward(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                        const void* packed, int32_t C,
                        const float* tf, int32_t tfN,
                        const uint8_t* flat_levels, const float* minmax,
                        const int32_t* labels, const int32_t* preds,
                        const float* out_rgba, const float* dL_dout,
                        const float* ckpt, int32_t seg_slots, int32_t nseg,
                        const int32_t* k_end, const int32_t* warp_kmax,
                        void* dL_dvol, float* dL_dtf, void* scratch, float* dL_dray, uint64_t* stats,
                        int32_t tile_begin, int32_t tile_end, void* stream);

/* ------------------------------------------------ one call per training step
 * BASELINE cfg3 / docs/DifferentiableRendering.md:88-127 + :213 as ONE call: the optimisation loop the doc
 * describes (render, L = mean((C - C_target)^2), dL/dvolume, dL/dTF) with the volume a learnable planar
 * tensor that changes every step.  Queued from C: mrt_fold_volume_occupancy_f32 -> mrt_classify_bricks
 * (skip levels; flat levels on a side stream) -> mrt_render_forward_ckpt -> the segment-parallel adjoint,
 * which forms G = (2/n)(out - target) per pixel itself (no dL/dout tensor is ever materialised) ->
 * mrt_unfold_grad_f32; the gradient buffers are cleared and the loss is reduced on a library-owned side
 * stream that forks from and joins `stream`, beside the march and the adjoint.  Results equal
 * mrt_render_forward_ckpt + mse + mrt_render_backward (image bit for bit, gradients to rounding).
 *   planar     : device float [C][Z][Y][X], the learnable volume (read only)
 *   tf         : device float [tfN][4] (tfMode 1) or NULL (tfMode 0)
 *   target_rgba: device float [nviews][H][W][4]
 *   workspace  : device, mrt_train_step_workspace_bytes(params, nviews, tfN) bytes, caller-owned, ZERO IT ONCE
 *                before the first call; it may be reused by every later step of the same geometry
 *   out_rgba   : device float [nviews][H][W][4], this step's image (result)
 *   loss       : optional device float[1] = mean((out - target)^2) over all nviews*H*W*4 values
 *   dL_dplanar : optional device float [C][Z][Y][X], overwritten      dL_dtf : optional float [tfN][4], overwritten
 * `cams` NULL: one view with the camera in params.  Needs skipEmpty = 1, tMode = 0, gamma = 1, no overlays,
 * volDtype = 0, unsharded (MRT_ERR_UNSUPPORTED otherwise: use the separate calls).
 * The side stream and its events are one set per device: calls on the same device must be issued from one host
 * thread (their side work then runs in call order; each call joins the side stream before it returns control of
 * `stream`).  Two workspaces may be in flight on two caller streams; one workspace belongs to one stream. */
size_t mrt_train_step_workspace_bytes(const MrtParams* params, int32_t nviews, int32_t tfN);
int mrt_train_step_mse(const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                       const float* planar, int32_t C, const float* tf, int32_t tfN, const float* target_rgba,
                       void* workspace, float* out_rgba, float* loss, float* dL_dplanar, float* dL_dtf, void* stream);
/* Diagnostics (tools/time_train_step.py; not thread-safe): mrt_debug_train_trace(NULL) arms a stage trace — the NEXT
 * mrt_train_step_mse records a timing event on the caller's stream after every stage; mrt_debug_train_trace(ms)
 * then synchronises and fills ms[6] with the device time of fold | classify | march | loss + wait for the adjoint
 * (side stream) + dL/dtf reduce | fold adjoint | final join, as the caller's stream saw them.  Measurement knobs read
 * once from the environment: MRT_TRAIN_NO_SIDE (everything on the caller's stream), MRT_TRAIN_NO_DIRECT (always
 * through the fold adjoint), MRT_TRAIN_PARTS=n (split the image into n tile-range parts, part i differentiated
 * while part i+1 marches: the measured loser). */
int mrt_debug_train_trace(float* ms6);

/* ------------------------------------------------ soft (learnable) occupancy
 * docs/DifferentiableRendering.md section 11 (:202-206; maths only, no reference code): "hard empty-space
 * skipping -> continuous occupancy o(x) in [0,1] learned and used multiplicatively".  `soft_occ` holds one
 * float per 8^3 brick of the occupancy grid ([nbz][nby][nbx], mrt_brick_count entries); a sample whose
 * trilinear base cell lies in brick b uses sigma' = soft_occ[b] * sigma in the compositing of
 * brats_rt.slang:135-139.  soft_occ == 1 everywhere reproduces mrt_render_forward bit for bit; hard
 * skipping (skip_levels) stays exact next to it, because a brick the classification proves empty has
 * sigma == 0 whatever its occupancy.  The backward returns dL/dsoft_occ[b] = sum of dL/dsigma'_i * sigma_i
 * over the samples of brick b (ACCUMULATED into; caller zeroes) next to dL/dvolume and dL/dtf (both now
 * through sigma = o * sigma).  fp32 unsharded volumes, no overlays; whole-ray backward. */
int mrt_render_forward_soft_occ(const MrtParams* params, const void* packed, int32_t C, const float* tf, int32_t tfN,
                                const uint8_t* skip_levels, const float* soft_occ, float* out_rgba,
                                int32_t tile_begin, int32_t tile_end, void* stream);
int mrt_render_backward_soft_occ(const MrtParams* params, const void* packed, int32_t C, const float* tf, int32_t tfN,
                                 const float* soft_occ, const float* out_rgba, const float* dL_dout,
                                 void* dL_dvol, float* dL_dtf, float* dL_dsoft_occ, void* scratch,
                                 int32_t tile_begin, int32_t tile_end, void* stream);

/* ------------------------------------------------ differentiable adaptive sampling
 * docs/DifferentiableRendering.md section 7 (:131-148; maths only, no reference code): per ray a
 * coarse pass of n_coarse uniform samples gives importance weights w_k = sigma_k + eps_w, their
 * piecewise-linear CDF is inverted at the fixed quantiles (j+1/2)/n_fine, and the n_fine samples
 * placed there are composited front to back with alpha_j = 1 - exp(-sigma_j * Delta_j), Delta_j the
 * length of sample j's quantile interval.  The backward includes the dependence of the sample
 * times and interval lengths on the weights (the implicit differentiation of :142-146).
 * fp32 unsharded volumes, no overlays, no skipping (the sampler is the acceleration).
 *   n_coarse <= 64;  dL_dvol / dL_dtf ACCUMULATED into (caller zeroes), layouts as mrt_render_backward;
 *   scratch: mrt_adaptive_scratch_bytes(tfN) bytes when dL_dtf != NULL. */
size_t mrt_adaptive_scratch_bytes(int32_t tfN);
int mrt_render_adaptive_forward(const MrtParams* params, const void* packed, int32_t C, const float* tf, int32_t tfN,
                                int32_t n_coarse, int32_t n_fine, float eps_w, float* out_rgba,
                                int32_t tile_begin, int32_t tile_end, void* stream);
int mrt_render_adaptive_backward(const MrtParams* params, const void* packed, int32_t C, const float* tf, int32_t tfN,
                                 int32_t n_coarse, int32_t n_fine, float eps_w, const float* out_rgba, const float* dL_dout,
                                 void* dL_dvol, float* dL_dtf, void* scratch, int32_t tile_begin, int32_t tile_end, void* stream);

/* ------------------------------------------------ slab renderer (u8 volume)
 * volume_cs, scripts/volumeRendering/volume_render.slang:104-148.
 * vol_u8: uint8 [Z][Y][X], one byte per voxel (the reference stores one voxel per u32
 * lane, app.py:150-158; values are identical). */
int mrt_render_slab_u8(const MrtSlabParams* params, const uint8_t* vol_u8,
                       float* out_rgba, int32_t tile_begin, int32_t tile_end, void* stream);

/* ------------------------------------------------ ingest formats
 * BC4 (single-channel block compression) decode, scripts/volumeRendering/app.py:200-250:
 * blocks: uint8 [D][ceil(H/4)*ceil(W/4)][8]  ->  out: uint8 [D][H][W]. */
int mrt_decode_bc4(const uint8_t* blocks, int32_t W, int32_t H, int32_t D, uint8_t* out, void* stream);
/* u8 -> fp32 /255 (volume_render.slang:38). */
int mrt_u8_to_f32(const uint8_t* in, size_t n, float* out, void* stream);
/* Affine normalise + clip to [0,1]: out = clip((in - vmin)/rng, 0, 1)
 * (inr/viewer/brats_viewer.py:50-56; percentiles are computed by the host). */
int mrt_normalize_f32(const float* in, size_t n, float vmin, float rng, float* out, void* stream);

/* ------------------------------------------------ INR inference (the producer of gPreds)
 * predict_volume of the reference's implicit-neural-representation segmenter (inr/inr/model.py:119-141,
 * called at inr/viewer/brats_viewer.py:250-310): per voxel, normalised coordinates -> Fourier
 * features (:11-18) -> [coords | features | M intensities] (:21-23) -> dense/ReLU chain (:43-50) ->
 * argmax.  One fused kernel.  impl 0 (default): tcgen05 tensor cores — 128-voxel tiles, fp32
 * accumulators in TMEM, every product a 3-term TF32 split so the logits match fp32 to ~1e-6 — when
 * the network fits (widths <= 64, weights within shared memory), else the fp32 FFMA kernel;
 * impl 1: fp32 FFMA on the CUDA cores; impl 2: tensor cores or MRT_ERR_UNSUPPORTED.
 *   mods_planar : device fp32 [M][Z][Y][X] (the renderer's planar layout), z-scored by the caller
 *                 like brats_viewer.py:279-287
 *   weights     : device fp32, per layer W[in][out] row-major followed by b[out]
 *   layer_dims  : HOST int32[n_layers+1]; layer_dims[0] must equal 3 + 6*fourier_freqs + M;
 *                 widths <= 64, classes <= 8, n_layers <= 8
 *   out_labels  : device int32 [Z][Y][X] — directly usable as `preds` of mrt_render_forward
 *   out_logits  : optional device fp32 [Z][Y][X][classes] */
int mrt_inr_predict(const float* mods_planar, int32_t M, int32_t X, int32_t Y, int32_t Z,
                    const float* weights, const int32_t* layer_dims, int32_t n_layers, int32_t fourier_freqs,
                    int32_t* out_labels, float* out_logits, int32_t impl, void* stream);

/* ------------------------------------------------ sort-last compositing
 * Ordered front-to-back `over` of K partial images (premultiplied colour + transmittance):
 *   (C,T) <- (C_a + T_a*C_b, T_a*T_b), front first; bg added once at the end, like the
 *   reference's `C = bgColor` start (brats_rt.slang:111): out = (bg + C, alpha).
 * partials: float4 [K][npix] = (r,g,b,T); order: int32[K] front-to-back indices. */
int mrt_composite_over(const float* partials, int32_t K, const int32_t* order, size_t npix,
                       const float* bg3, int32_t alphaMode, float* out_rgba, void* stream);

/* Same composite, stored to `nouts` (<= 16) images at once: the local one and/or peer-mapped copies
 * on other GPUs, so that the final all-gather of sort-last rendering is done by the composite
 * kernel's own stores over NVLink.  `outs` is a HOST array of device pointers. */
int mrt_composite_over_multi(const float* partials, int32_t K, const int32_t* order, size_t npix,
                             const float* bg3, int32_t alphaMode, float* const* outs, int32_t nouts, void* stream);

/* Sort-last exchange fused into the march (one frame): image rows [s*strip_rows, (s+1)*strip_rows)
 * are written to strip_out[s] (float4 [strip_rows][W], row 0 = the strip's first image row) instead
 * of one [H][W] image — with strip_out[s] a peer-mapped buffer of strip s's owner GPU, the
 * all-to-all of partial images happens through the kernel's stores.  `strip_out` is a HOST array of
 * `nstrips` (<= 16) device pointers; nstrips*strip_rows >= H. */
int mrt_render_forward_strips(const MrtParams* params, const void* packed, int32_t C,
                              const float* tf, int32_t tfN, const uint8_t* skip_levels,
                              float* const* strip_out, int32_t nstrips, int32_t strip_rows,
                              int32_t tile_begin, int32_t tile_end, void* stream);

/* ------------------------------------------------ roofline probe
 * Random 32-byte-sector gather over a buffer of `bytes` (power of two), `n_gathers`
 * sectors per launch; writes a checksum so the loads cannot be elided.  Used by bench.py
 * to measure the L2-resident and HBM-resident gather ceilings. */
int mrt_gather_probe(const void* buf, size_t bytes, size_t n_gathers, uint32_t seed,
                     float* out_checksum, void* stream);

/* ------------------------------------------------ host-buffer entry point
 * The call a non-CUDA host makes: host volume in, host image out.  Allocates and frees
 * its own temporary device memory, copies H2D, packs, builds + classifies occupancy,
 * renders, copies D2H, and synchronises.  planar_host: [C][Z][Y][X] fp32; out_rgba_host:
 * float4 [H][W].  labels/preds may be NULL. */
int mrt_render_host(const MrtParams* params, const float* planar_host, int32_t C,
                    const float* tf_host, int32_t tfN,
                    const int32_t* labels_host, const int32_t* preds_host,
                    float* out_rgba_host);

/* ------------------------------------------------ host-buffer pipeline
 * mrt_render_host for a host that renders step after step (each step: a [C][Z][Y][X] volume, a TF
 * and `nviews` cameras in, `nviews` frames out — the reference's load_dir + frame loop,
 * inr/viewer/brats_viewer.py:188-248,369-450, without a window): a multi-buffered object that
 * owns its device buffers and four streams (upload / prepare / march / download), so that the
 * download of step i, the compute of step i+1 and the upload of step i+2 overlap.  Host buffers
 * should be page-locked for the copies to be asynchronous.  submit() returns after queueing;
 * wait(ticket) blocks until that step's frames are in out_rgba_host.  Fixed geometry per object
 * (C, dims, image size); fp32 volumes without label overlays.  One thread at a time per object. */
typedef struct MrtHostPipeline MrtHostPipeline;
int mrt_host_pipeline_create(MrtHostPipeline** out, int32_t C, int32_t X, int32_t Y, int32_t Z,
                             int32_t W, int32_t H, int32_t max_views, int32_t max_tfN, int32_t depth);
/* Make a planar [C][Z][Y][X] fp32 host volume RESIDENT on the device (queued; returns at once; the
 * host buffer must stay valid until the next wait()).  This is the reference's own split: it uploads
 * the modality buffers once in load_dir (inr/viewer/brats_viewer.py:219-230) and per frame only
 * refills `gParams` (:405-426).  Steps submitted with planar_host == NULL render the resident volume. */
int mrt_host_pipeline_set_volume(MrtHostPipeline* p, const float* planar_host);
/* out_flags */
#define MRT_OUT_FRESH 1   /* out_rgba_host holds unknown data: write the whole background, not just the damage */
/* Queue one step.  planar_host: this step's own volume (uploaded now), or NULL for the resident one.
 * Frames come back sparse when skipping is on (skipEmpty, tMode 0, gamma 1): per view only the
 * bounding rectangle of the tiles that can differ from the background crosses PCIe; the rest of the
 * host frame is kept at the background by the pipeline, which remembers what it last wrote into
 * each output buffer and clears only the damage.  Contract: an out_rgba_host buffer that was handed
 * to submit() is written by nobody else until mrt_host_pipeline_forget() or the pipeline's
 * destruction (or pass MRT_OUT_FRESH), and is not submitted again before its ticket was waited for.
 * The frames in host memory equal mrt_render_forward_batch + a dense download, bit for bit. */
int mrt_host_pipeline_submit(MrtHostPipeline* p, const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                             const float* planar_host, const float* tf_host, int32_t tfN,
                             float* out_rgba_host, int32_t out_flags, int64_t* ticket);
int mrt_host_pipeline_wait(MrtHostPipeline* p, int64_t ticket);
/* Bytes moved by the last submitted step: [0] host->device, [1] device->host, [2] host-side background fill. */
void mrt_host_pipeline_last_bytes(const MrtHostPipeline* p, uint64_t out3[3]);
void mrt_host_pipeline_forget(MrtHostPipeline* p, const float* out_rgba_host);
const char* mrt_host_pipeline_error(const MrtHostPipeline* p);
void mrt_host_pipeline_destroy(MrtHostPipeline* p);

#ifdef __cplusplus
}
#endif
#endif /* MRT_H_ */
