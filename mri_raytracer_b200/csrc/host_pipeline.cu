// host_pipeline.cu — host buffers in, host frames out, as a pipelined service.
//
// mrt_render_host (c_api.cu) is the one-shot call a non-CUDA host makes; it serialises
// H2D -> prepare -> march -> D2H.  A host that renders step after step keeps PCIe and the GPU busy
// instead: this object owns `depth` slots of device buffers and four streams, and for every
// submitted step queues
//     h2d stream     : the step's inputs — TF (+ a planar volume when the caller passes one)
//     prep stream    : fold + occupancy, classify, per-view screen spans (+ the spans to the host)
//     compute stream : ONE batched march of all views (span cull)
//     d2h stream     : the finished frames
// chained by events, so the steps overlap each other.
//
// Round 2 (the round-1 pipeline moved 143 MB up and 134 MB down per step and was PCIe-bound at
// 3.1 ms against 0.85 ms of compute):
//   * the volume can be RESIDENT (mrt_host_pipeline_set_volume), as in the reference, which uploads
//     its buffers once at load time and only refills `gParams` per frame
//     (inr/viewer/brats_viewer.py:219-230 vs :405-426); a step's inputs are then cameras, params,
//     modality weights and the TF;
//   * frames come down SPARSE: only the tiles inside each view's spans — the screen footprint of
//     the active bricks (mrt_view_spans), outside which every pixel is the background — cross PCIe.  When the
//     output buffer is page-locked (hence device-mapped under UVA) the march kernel stores those
//     tiles STRAIGHT into host memory, exactly as the multi-GPU path stores into a peer GPU: the
//     transfer overlaps the march and there is no copy at all; otherwise each view's bounding
//     rectangle is copied (one strided copy per view).  The rest of the host frame is kept at the
//     background by damage tracking: the pipeline remembers, per tile row, the range it last wrote
//     into each output buffer and clears only what the new range no longer covers (everything, the
//     first time it sees a buffer).  The frames in host memory are bit-identical to a dense download.
// It is the only part of the library that owns device memory (documented exception to "the
// caller owns every buffer": the caller here has no device pointers at all).
#include "march.cuh"
#include "kernels.h"
#include "../../include/mrt.h"
#include <new>
#include <unordered_map>
#include <vector>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct HpOutState {
  std::vector<int32_t> ranges;      // [view][tile row][2]: pixel columns (inclusive) that may hold non-background; x0 > x1: none
  int nviews;
  float bg[4];
  int W, H;
  int64_t last_ticket;
};

struct MrtHostPipeline {
  int C, X, Y, Z, W, H, max_views, max_tf, depth;
  size_t planar_bytes, packed_bytes, frame_bytes, levels_bytes;
  int nb, tiles_y;
  cudaStream_t s_h2d, s_prep, s_cmp, s_d2h;
  float* d_resident;          // planar volume uploaded by set_volume (nullptr: none)
  cudaEvent_t e_resident;
  struct Slot {
    float* d_planar; void* d_packed; void* d_quad; float* d_minmax; uint8_t* d_levels; float* d_tf; float* d_frames;
    int32_t* d_spans; int32_t* h_spans;
    cudaEvent_t e_h2d, e_prep, e_cmp, e_done;
    int64_t ticket;           // last ticket submitted into this slot (-1: none)
  } slot[4];
  int64_t next_ticket;
  uint64_t d2h_bytes_last, h2d_bytes_last, host_fill_bytes_last;
  std::unordered_map<const float*, HpOutState>* outs;
  char err[256];
};

extern "C" {

void mrt_host_pipeline_destroy(MrtHostPipeline* p) {
  if (!p) return;
  if (p->s_h2d) cudaStreamSynchronize(p->s_h2d);
  if (p->s_prep) cudaStreamSynchronize(p->s_prep);
  if (p->s_cmp) cudaStreamSynchronize(p->s_cmp);
  if (p->s_d2h) cudaStreamSynchronize(p->s_d2h);
  for (int i = 0; i < p->depth; ++i) {
    MrtHostPipeline::Slot& s = p->slot[i];
    cudaFree(s.d_planar); cudaFree(s.d_packed); cudaFree(s.d_quad); cudaFree(s.d_minmax); cudaFree(s.d_levels); cudaFree(s.d_tf);
    cudaFree(s.d_frames); cudaFree(s.d_spans);
    if (s.h_spans) cudaFreeHost(s.h_spans);
    if (s.e_h2d) cudaEventDestroy(s.e_h2d);
    if (s.e_prep) cudaEventDestroy(s.e_prep);
    if (s.e_cmp) cudaEventDestroy(s.e_cmp);
    if (s.e_done) cudaEventDestroy(s.e_done);
  }
  cudaFree(p->d_resident);
  if (p->e_resident) cudaEventDestroy(p->e_resident);
  if (p->s_h2d) cudaStreamDestroy(p->s_h2d);
  if (p->s_prep) cudaStreamDestroy(p->s_prep);
  if (p->s_cmp) cudaStreamDestroy(p->s_cmp);
  if (p->s_d2h) cudaStreamDestroy(p->s_d2h);
  delete p->outs;
  delete p;
}

const char* mrt_host_pipeline_error(const MrtHostPipeline* p) { return p ? p->err : "null pipeline"; }

#define HP_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf(p->err, sizeof(p->err), "%s: %s", #call, cudaGetErrorString(e_)); rc = MRT_ERR_CUDA; goto fail; } } while (0)

// The march samples from the 16 B/voxel quad layout (mrt_pack_volume_quad) unless MRT_HP_QUAD=0.  While the
// frames' PCIe writes bounded this pipeline (spans = hull of the active box, 45 MB per step) the faster sampler
// bought nothing and building the 4x larger layout cost a little (0.945 vs 0.918 ms per 8-view step), so it was
// off; with the brick-union spans (26 MB per step) the step is bound by the march itself and the quad layout wins,
// 0.640 vs 0.664 ms (gpurun_out/bench_t8_*.json).
static bool hp_use_quad() {
  static const char* env = getenv("MRT_HP_QUAD");
  return !(env && env[0] == '0');
}

int mrt_host_pipeline_create(MrtHostPipeline** out, int32_t C, int32_t X, int32_t Y, int32_t Z, int32_t W, int32_t H,
                             int32_t max_views, int32_t max_tfN, int32_t depth) {
  if (!out) return MRT_ERR_BAD_ARG;
  *out = nullptr;
  if (C < 1 || C > 4 || X < 2 || Y < 2 || Z < 2 || W < 1 || H < 1 || max_views < 1 || max_tfN < 2 ||
      max_tfN > MRT_MAX_TF || depth < 1 || depth > 4)
    return MRT_ERR_BAD_ARG;
  MrtHostPipeline* p = new (std::nothrow) MrtHostPipeline;
  if (!p) return MRT_ERR_CUDA;
  memset(p, 0, sizeof(*p));
  p->outs = new (std::nothrow) std::unordered_map<const float*, HpOutState>();
  if (!p->outs) { delete p; return MRT_ERR_CUDA; }
  int rc = MRT_OK;
  p->C = C; p->X = X; p->Y = Y; p->Z = Z; p->W = W; p->H = H; p->max_views = max_views; p->max_tf = max_tfN;
  p->depth = depth;
  p->planar_bytes = (size_t)X * Y * Z * C * sizeof(float);
  p->packed_bytes = mrt_packed_volume_bytes(1, X, Y, Z);          // folded (or single-channel) layout
  p->frame_bytes = (size_t)W * H * 4 * sizeof(float);
  p->levels_bytes = mrt_skip_levels_bytes(X, Y, Z);
  p->nb = mrt_brick_count(X, Y, Z);
  p->tiles_y = mrt_tiles_y_(H);
  for (int i = 0; i < depth; ++i) p->slot[i].ticket = -1;
  HP_CUDA(cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking));
  {
    // the prepare stage of step i+1 (fold, occupancy, classify, spans: short kernels) must not queue
    // behind the 65k-CTA march of step i: highest priority, so its CTAs take the SM slots the march
    // frees as it drains (MRT_HP_PRIO=0 switches this off)
    int lo_p = 0, hi_p = 0;
    static const char* env = getenv("MRT_HP_PRIO");
    const bool prio = !(env && env[0] == '0');
    HP_CUDA(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
    HP_CUDA(cudaStreamCreateWithPriority(&p->s_prep, cudaStreamNonBlocking, prio ? hi_p : lo_p));
  }
  HP_CUDA(cudaStreamCreateWithFlags(&p->s_cmp, cudaStreamNonBlocking));
  HP_CUDA(cudaStreamCreateWithFlags(&p->s_d2h, cudaStreamNonBlocking));
  HP_CUDA(cudaEventCreateWithFlags(&p->e_resident, cudaEventDisableTiming));
  for (int i = 0; i < depth; ++i) {
    MrtHostPipeline::Slot& s = p->slot[i];
    HP_CUDA(cudaMalloc(&s.d_packed, p->packed_bytes));
    if (hp_use_quad()) HP_CUDA(cudaMalloc(&s.d_quad, mrt_packed_volume_bytes_quad(X, Y, Z)));
    HP_CUDA(cudaMalloc(&s.d_minmax, (size_t)p->nb * 2 * sizeof(float)));
    HP_CUDA(cudaMalloc(&s.d_levels, p->levels_bytes));
    HP_CUDA(cudaMalloc(&s.d_tf, (size_t)max_tfN * 4 * sizeof(float)));
    HP_CUDA(cudaMalloc(&s.d_frames, p->frame_bytes * max_views));
    HP_CUDA(cudaMalloc(&s.d_spans, (size_t)max_views * p->tiles_y * 2 * sizeof(int32_t)));
    HP_CUDA(cudaMallocHost(&s.h_spans, (size_t)max_views * p->tiles_y * 2 * sizeof(int32_t)));
    HP_CUDA(cudaEventCreateWithFlags(&s.e_h2d, cudaEventDisableTiming));
    HP_CUDA(cudaEventCreateWithFlags(&s.e_prep, cudaEventDisableTiming));
    HP_CUDA(cudaEventCreateWithFlags(&s.e_cmp, cudaEventDisableTiming));
    HP_CUDA(cudaEventCreateWithFlags(&s.e_done, cudaEventDisableTiming));
  }
  *out = p;
  return MRT_OK;
fail:
  mrt_host_pipeline_destroy(p);
  return rc;
}

int mrt_host_pipeline_set_volume(MrtHostPipeline* p, const float* planar_host) {
  if (!p) return MRT_ERR_BAD_ARG;
  int rc = MRT_OK;
  if (!planar_host) { snprintf(p->err, sizeof(p->err), "set_volume: null pointer"); return MRT_ERR_BAD_ARG; }
  // steps in flight may still read the previous resident volume
  HP_CUDA(cudaStreamSynchronize(p->s_prep));
  if (!p->d_resident) HP_CUDA(cudaMalloc(&p->d_resident, p->planar_bytes));
  HP_CUDA(cudaMemcpyAsync(p->d_resident, planar_host, p->planar_bytes, cudaMemcpyHostToDevice, p->s_h2d));
  HP_CUDA(cudaEventRecord(p->e_resident, p->s_h2d));
  return MRT_OK;
fail:
  return rc;
}

// background into columns [ox0, ox1] minus [nx0, nx1] of image rows [y0, y1] (row-major float4, width W)
static size_t hp_fill_damage(float* img, int W, int y0, int y1, int ox0, int ox1, int nx0, int nx1, const float bg[4]) {
  if (ox1 < ox0) return 0;
  int seg[2][2] = {{ox0, ox1}, {1, 0}};
  if (nx1 >= nx0) {
    seg[0][1] = (nx0 - 1 < ox1) ? nx0 - 1 : ox1;
    seg[1][0] = (nx1 + 1 > ox0) ? nx1 + 1 : ox0; seg[1][1] = ox1;
  }
  size_t n = 0;
  for (int y = y0; y <= y1; ++y) {
    float* row = img + (size_t)y * W * 4;
    for (int k = 0; k < 2; ++k)
      for (int x = seg[k][0]; x <= seg[k][1]; ++x) {
        float* px = row + (size_t)x * 4;
        px[0] = bg[0]; px[1] = bg[1]; px[2] = bg[2]; px[3] = bg[3];
        ++n;
      }
  }
  return n * 4 * sizeof(float);
}

// can the device address this host buffer directly (page-locked => mapped under UVA)?
static float* hp_device_view(const float* host, size_t bytes) {
  static const char* env = getenv("MRT_HP_ZEROCOPY");
  if (env && env[0] == '0') return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) return nullptr;
  cudaPointerAttributes b;                                 // the whole range must belong to the allocation
  if (cudaPointerGetAttributes(&b, reinterpret_cast<const char*>(host) + bytes - 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (b.type != cudaMemoryTypeHost) return nullptr;
  return reinterpret_cast<float*>(a.devicePointer);
}

int mrt_host_pipeline_submit(MrtHostPipeline* p, const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                             const float* planar_host, const float* tf_host, int32_t tfN, float* out_rgba_host,
                             int32_t out_flags, int64_t* ticket) {
  if (!p) return MRT_ERR_BAD_ARG;
  int rc = MRT_OK;
#define HP_REQ(cond, msg) do { if (!(cond)) { snprintf(p->err, sizeof(p->err), "submit: %s", msg); return MRT_ERR_BAD_ARG; } } while (0)
  HP_REQ(params && cams && out_rgba_host, "null pointer");
  HP_REQ(planar_host || p->d_resident, "no volume: pass planar_host or call mrt_host_pipeline_set_volume first");
  HP_REQ(nviews >= 1 && nviews <= p->max_views, "nviews outside 1..max_views");
  HP_REQ((int)params->dims[0] == p->X && (int)params->dims[1] == p->Y && (int)params->dims[2] == p->Z, "params->dims differ from the pipeline's");
  HP_REQ((int)params->imageSize[0] == p->W && (int)params->imageSize[1] == p->H, "params->imageSize differs from the pipeline's");
  HP_REQ(!params->tfMode || (tf_host && tfN >= 2 && tfN <= p->max_tf), "tf invalid");
  HP_REQ(!params->shardEnabled && !params->volDtype && !params->showSeg && !params->showPred, "shards / fp16 / overlays are not supported here");
  {
    const int64_t t = p->next_ticket;
    MrtHostPipeline::Slot& s = p->slot[t % p->depth];
    // the slot's previous occupant must have left the device (its frames are on the host)
    if (s.ticket >= 0) HP_CUDA(cudaEventSynchronize(s.e_done));
    // an output buffer still being written by an earlier step must not be touched yet
    HpOutState& os = (*p->outs)[out_rgba_host];
    const bool known = !os.ranges.empty() && !(out_flags & MRT_OUT_FRESH) && os.W == p->W && os.H == p->H;
    if (!os.ranges.empty() && os.last_ticket >= 0) {
      MrtHostPipeline::Slot& prev = p->slot[os.last_ticket % p->depth];
      if (prev.ticket == os.last_ticket) HP_CUDA(cudaEventSynchronize(prev.e_done));
    }
    uint64_t h2d = sizeof(MrtParams) + (uint64_t)nviews * sizeof(MrtCamera);
    // ---- upload of the step's inputs
    const float* d_planar = p->d_resident;
    if (planar_host) {
      if (!s.d_planar) HP_CUDA(cudaMalloc(&s.d_planar, p->planar_bytes));
      HP_CUDA(cudaMemcpyAsync(s.d_planar, planar_host, p->planar_bytes, cudaMemcpyHostToDevice, p->s_h2d));
      d_planar = s.d_planar;
      h2d += p->planar_bytes;
    } else {
      HP_CUDA(cudaStreamWaitEvent(p->s_prep, p->e_resident, 0));
    }
    if (params->tfMode) {
      HP_CUDA(cudaMemcpyAsync(s.d_tf, tf_host, (size_t)tfN * 4 * sizeof(float), cudaMemcpyHostToDevice, p->s_h2d));
      h2d += (uint64_t)tfN * 4 * sizeof(float);
    }
    HP_CUDA(cudaEventRecord(s.e_h2d, p->s_h2d));
    // ---- prepare: fold + occupancy, classify, spans
    HP_CUDA(cudaStreamWaitEvent(p->s_prep, s.e_h2d, 0));
    MrtParams P = *params;
    const bool skip = P.skipEmpty && P.tMode == 0;
    const bool sparse = skip && P.gamma == 1.0f;            // the span path (cull + sparse download)
    int Ce = p->C;
    bool quad_done = false;
    if (p->C > 1) {                    // modality fold (+ occupancy of the folded field, + the quad layout, in the same pass)
      if (skip && s.d_quad) {
        rc = mrt_fold_volume_occupancy_quad_f32(&P, d_planar, p->C, nullptr, s.d_quad, s.d_minmax, p->s_prep);
        quad_done = true;
      } else if (skip) rc = mrt_fold_volume_occupancy_f32(&P, d_planar, p->C, (float*)s.d_packed, s.d_minmax, p->s_prep);
      else rc = mrt_fold_volume_f32(&P, d_planar, p->C, (float*)s.d_packed, p->s_prep);
      P.volEnabled[0] = 1; P.volEnabled[1] = P.volEnabled[2] = P.volEnabled[3] = 0;
      P.volWeight[0] = 1.0f;
      Ce = 1;
    } else {
      rc = mrt_pack_volume_f32(d_planar, 1, p->X, p->Y, p->Z, s.d_packed, p->s_prep);
      if (rc == MRT_OK && skip) rc = mrt_build_occupancy(s.d_packed, 1, p->X, p->Y, p->Z, s.d_minmax, p->s_prep);
    }
    if (rc == MRT_OK && skip)
      rc = mrt_classify_bricks(&P, s.d_minmax, Ce, s.d_tf, tfN, nullptr, nullptr, s.d_levels, 0, p->s_prep);
    const void* d_sampler = s.d_packed;
    if (rc == MRT_OK && s.d_quad) {      // two 16-byte loads per sample instead of eight scalar ones; same image
      if (!quad_done) rc = mrt_pack_volume_quad((const float*)s.d_packed, p->X, p->Y, p->Z, s.d_quad, p->s_prep);
      d_sampler = s.d_quad;
    }
    const size_t span_bytes = (size_t)nviews * p->tiles_y * 2 * sizeof(int32_t);
    if (rc == MRT_OK && sparse) {
      rc = mrt_view_spans(&P, cams, nviews, Ce, s.d_levels, s.d_spans, p->s_prep);
      if (rc == MRT_OK) HP_CUDA(cudaMemcpyAsync(s.h_spans, s.d_spans, span_bytes, cudaMemcpyDeviceToHost, p->s_prep));
    }
    if (rc != MRT_OK) { snprintf(p->err, sizeof(p->err), "submit: %s", mrt_last_error()); return rc; }
    HP_CUDA(cudaEventRecord(s.e_prep, p->s_prep));
    if (s.d_quad) P.volDtype = 3;                            // (classify / spans above took the scalar description)
    // ---- march + download
    HP_CUDA(cudaStreamWaitEvent(p->s_cmp, s.e_prep, 0));
    uint64_t d2h = 0, filled = 0;
    const float bgp[4] = {P.bgColor[0], P.bgColor[1], P.bgColor[2], P.alphaMode ? 0.0f : 1.0f};
    const int ty = p->tiles_y;
    float* out_dev = sparse ? hp_device_view(out_rgba_host, p->frame_bytes * nviews) : nullptr;
    if (sparse) {
      // spans precomputed above: one load + two compares per warp instead of a per-ray box test.
      // Zero-copy: the in-span tiles go straight to the (device-mapped) host frames and nothing else is
      // stored; staged: complete frames into the slot, bounding rectangles copied below.
      if (out_dev) rc = mrt_render_forward_batch_sparse(&P, cams, nviews, d_sampler, Ce, s.d_tf, tfN, s.d_levels, out_dev, s.d_spans, 0, p->s_cmp);
      else rc = mrt_render_forward_batch_sparse(&P, cams, nviews, d_sampler, Ce, s.d_tf, tfN, s.d_levels, s.d_frames, s.d_spans, 2, p->s_cmp);
    } else {
      rc = mrt_render_forward_batch(&P, cams, nviews, d_sampler, Ce, s.d_tf, tfN, skip ? s.d_levels : nullptr,
                                    nullptr, nullptr, s.d_frames, nullptr, nullptr, 0,
                                    mrt_tile_count(p->W, p->H), p->s_cmp);
    }
    if (rc != MRT_OK) { snprintf(p->err, sizeof(p->err), "submit: %s", mrt_last_error()); return rc; }
    HP_CUDA(cudaEventRecord(s.e_cmp, p->s_cmp));
    HP_CUDA(cudaStreamWaitEvent(p->s_d2h, s.e_cmp, 0));
    std::vector<int32_t> now((size_t)nviews * ty * 2);
    if (sparse) {
      // the spans of THIS step are on the host as soon as the (short) prepare stage is done; the
      // march runs meanwhile
      HP_CUDA(cudaEventSynchronize(s.e_prep));
      d2h += span_bytes;
      const bool same_bg = known && memcmp(os.bg, bgp, sizeof(bgp)) == 0;
      for (int v = 0; v < nviews; ++v) {
        const int32_t* sp = s.h_spans + (size_t)v * ty * 2;
        int32_t* nr = now.data() + (size_t)v * ty * 2;
        int rx0 = p->W, rx1 = -1, ry0 = p->H, ry1 = -1;
        for (int b = 0; b < ty; ++b) {
          const int x0 = sp[2 * b], x1 = sp[2 * b + 1];
          nr[2 * b] = 1; nr[2 * b + 1] = 0;
          if (x0 > x1) continue;
          // whole tiles are stored by the march: round the span outward to tile columns
          const int tx0 = x0 & ~MRT_TILE_MASK, tx1 = ((x1 | MRT_TILE_MASK) < p->W - 1) ? (x1 | MRT_TILE_MASK) : p->W - 1;
          nr[2 * b] = tx0; nr[2 * b + 1] = tx1;
          const int y0 = b << MRT_TILE_SHIFT, y1 = (y0 + MRT_TILE_EDGE - 1 < p->H - 1) ? y0 + MRT_TILE_EDGE - 1 : p->H - 1;
          if (tx0 < rx0) rx0 = tx0;
          if (tx1 > rx1) rx1 = tx1;
          if (y0 < ry0) ry0 = y0;
          if (y1 > ry1) ry1 = y1;
          if (out_dev) d2h += (uint64_t)(tx1 - tx0 + 1) * (y1 - y0 + 1) * 4 * sizeof(float);
        }
        if (!out_dev && rx1 >= rx0) {
          // staged: the bounding rectangle of the view's spans in one strided copy; every band inside it
          // then holds device data (background included), so the tracked range is the rectangle's
          for (int b = ry0 >> MRT_TILE_SHIFT; b <= (ry1 >> MRT_TILE_SHIFT); ++b) { nr[2 * b] = rx0; nr[2 * b + 1] = rx1; }
        }
        float* himg = out_rgba_host + (size_t)v * p->W * p->H * 4;
        // the host frame outside the new ranges must hold the background: clear what the previous
        // ranges covered and the new ones do not (everything else when the contents are unknown)
        const bool have_old = same_bg && v < os.nviews;
        for (int b = 0; b < ty; ++b) {
          const int y0 = b << MRT_TILE_SHIFT, y1 = (y0 + MRT_TILE_EDGE - 1 < p->H - 1) ? y0 + MRT_TILE_EDGE - 1 : p->H - 1;
          const int ox0 = have_old ? os.ranges[((size_t)v * ty + b) * 2] : 0;
          const int ox1 = have_old ? os.ranges[((size_t)v * ty + b) * 2 + 1] : p->W - 1;
          filled += hp_fill_damage(himg, p->W, y0, y1, ox0, ox1, nr[2 * b], nr[2 * b + 1], bgp);
        }
        if (!out_dev && rx1 >= rx0) {
          const size_t off = ((size_t)ry0 * p->W + rx0) * 4;
          const size_t wbytes = (size_t)(rx1 - rx0 + 1) * 4 * sizeof(float);
          HP_CUDA(cudaMemcpy2DAsync(himg + off, (size_t)p->W * 4 * sizeof(float),
                                    s.d_frames + (size_t)v * p->W * p->H * 4 + off, (size_t)p->W * 4 * sizeof(float),
                                    wbytes, (size_t)(ry1 - ry0 + 1), cudaMemcpyDeviceToHost, p->s_d2h));
          d2h += wbytes * (size_t)(ry1 - ry0 + 1);
        }
      }
    } else {
      HP_CUDA(cudaMemcpyAsync(out_rgba_host, s.d_frames, p->frame_bytes * nviews, cudaMemcpyDeviceToHost, p->s_d2h));
      d2h += p->frame_bytes * nviews;
      for (int v = 0; v < nviews; ++v)
        for (int b = 0; b < ty; ++b) { now[((size_t)v * ty + b) * 2] = 0; now[((size_t)v * ty + b) * 2 + 1] = p->W - 1; }
    }
    os.ranges.swap(now);
    os.nviews = nviews;               // view slots beyond nviews: contents unknown to us from now on
    memcpy(os.bg, bgp, sizeof(bgp));
    os.W = p->W; os.H = p->H; os.last_ticket = t;
    HP_CUDA(cudaEventRecord(s.e_done, p->s_d2h));
    p->d2h_bytes_last = d2h; p->h2d_bytes_last = h2d; p->host_fill_bytes_last = filled;
    s.ticket = t;
    p->next_ticket = t + 1;
    if (ticket) *ticket = t;
  }
  return MRT_OK;
fail:
  return rc;
#undef HP_REQ
}

int mrt_host_pipeline_wait(MrtHostPipeline* p, int64_t ticket) {
  if (!p) return MRT_ERR_BAD_ARG;
  int rc = MRT_OK;
  if (ticket < 0 || ticket >= p->next_ticket) { snprintf(p->err, sizeof(p->err), "wait: unknown ticket"); return MRT_ERR_BAD_ARG; }
  {
    MrtHostPipeline::Slot& s = p->slot[ticket % p->depth];
    if (s.ticket != ticket) return MRT_OK;          // the slot has been reused: that step completed long ago
    HP_CUDA(cudaEventSynchronize(s.e_done));
  }
  return MRT_OK;
fail:
  return rc;
}

// bytes moved by the LAST submitted step: [0] host -> device, [1] device -> host, [2] host-side background fill
void mrt_host_pipeline_last_bytes(const MrtHostPipeline* p, uint64_t out3[3]) {
  if (!p || !out3) return;
  out3[0] = p->h2d_bytes_last; out3[1] = p->d2h_bytes_last; out3[2] = p->host_fill_bytes_last;
}

// forget what the pipeline knows about an output buffer (the caller wrote to it, or freed it)
void mrt_host_pipeline_forget(MrtHostPipeline* p, const float* out_rgba_host) {
  if (p && p->outs) p->outs->erase(out_rgba_host);
}

}  // extern "C"
