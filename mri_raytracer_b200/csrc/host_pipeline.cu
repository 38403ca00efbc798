// host_pipeline.cu — host buffers in, host frames out, as a double-buffered pipeline.
//
// mrt_render_host (c_api.cu) is the one-shot call a non-CUDA host makes; it serialises
// H2D -> prepare -> march -> D2H.  A host that renders step after step (a new volume / TF / orbit
// batch each step) can keep PCIe busy in both directions instead: this object owns `depth` slots
// of device buffers and three streams, and for every submitted step queues
//     h2d stream     : planar volume + TF                      (host -> device)
//     compute stream : fold + occupancy, classify, ONE batched march of all views
//     d2h stream     : the finished frames                     (device -> host)
// chained by events, so step i's download and step i+1's upload overlap each other and the
// compute in between.  It is the only part of the library that owns device memory (documented
// exception to "the caller owns every buffer": the caller here has no device pointers at all).
#include "march.cuh"
#include "kernels.h"
#include "../../include/mrt.h"
#include <new>
#include <stdio.h>
#include <string.h>

struct MrtHostPipeline {
  int C, X, Y, Z, W, H, max_views, max_tf, depth;
  size_t planar_bytes, packed_bytes, frame_bytes, levels_bytes;
  int nb;
  cudaStream_t s_h2d, s_cmp, s_d2h;
  struct Slot {
    float* d_planar; void* d_packed; float* d_minmax; uint8_t* d_levels; float* d_tf; float* d_frames;
    cudaEvent_t e_h2d, e_cmp, e_done;
    int64_t ticket;           // last ticket submitted into this slot (-1: none)
  } slot[4];
  int64_t next_ticket;
  char err[256];
};

extern "C" {

void mrt_host_pipeline_destroy(MrtHostPipeline* p) {
  if (!p) return;
  if (p->s_h2d) cudaStreamSynchronize(p->s_h2d);
  if (p->s_cmp) cudaStreamSynchronize(p->s_cmp);
  if (p->s_d2h) cudaStreamSynchronize(p->s_d2h);
  for (int i = 0; i < p->depth; ++i) {
    MrtHostPipeline::Slot& s = p->slot[i];
    cudaFree(s.d_planar); cudaFree(s.d_packed); cudaFree(s.d_minmax); cudaFree(s.d_levels); cudaFree(s.d_tf);
    cudaFree(s.d_frames);
    if (s.e_h2d) cudaEventDestroy(s.e_h2d);
    if (s.e_cmp) cudaEventDestroy(s.e_cmp);
    if (s.e_done) cudaEventDestroy(s.e_done);
  }
  if (p->s_h2d) cudaStreamDestroy(p->s_h2d);
  if (p->s_cmp) cudaStreamDestroy(p->s_cmp);
  if (p->s_d2h) cudaStreamDestroy(p->s_d2h);
  delete p;
}

const char* mrt_host_pipeline_error(const MrtHostPipeline* p) { return p ? p->err : "null pipeline"; }

#define HP_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf(p->err, sizeof(p->err), "%s: %s", #call, cudaGetErrorString(e_)); rc = MRT_ERR_CUDA; goto fail; } } while (0)

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
  int rc = MRT_OK;
  p->C = C; p->X = X; p->Y = Y; p->Z = Z; p->W = W; p->H = H; p->max_views = max_views; p->max_tf = max_tfN;
  p->depth = depth;
  p->planar_bytes = (size_t)X * Y * Z * C * sizeof(float);
  p->packed_bytes = mrt_packed_volume_bytes(1, X, Y, Z);          // folded (or single-channel) layout
  p->frame_bytes = (size_t)W * H * 4 * sizeof(float);
  p->levels_bytes = mrt_skip_levels_bytes(X, Y, Z);
  p->nb = mrt_brick_count(X, Y, Z);
  for (int i = 0; i < depth; ++i) p->slot[i].ticket = -1;
  HP_CUDA(cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking));
  HP_CUDA(cudaStreamCreateWithFlags(&p->s_cmp, cudaStreamNonBlocking));
  HP_CUDA(cudaStreamCreateWithFlags(&p->s_d2h, cudaStreamNonBlocking));
  for (int i = 0; i < depth; ++i) {
    MrtHostPipeline::Slot& s = p->slot[i];
    HP_CUDA(cudaMalloc(&s.d_planar, p->planar_bytes));
    HP_CUDA(cudaMalloc(&s.d_packed, p->packed_bytes));
    HP_CUDA(cudaMalloc(&s.d_minmax, (size_t)p->nb * 2 * sizeof(float)));
    HP_CUDA(cudaMalloc(&s.d_levels, p->levels_bytes));
    HP_CUDA(cudaMalloc(&s.d_tf, (size_t)max_tfN * 4 * sizeof(float)));
    HP_CUDA(cudaMalloc(&s.d_frames, p->frame_bytes * max_views));
    HP_CUDA(cudaEventCreateWithFlags(&s.e_h2d, cudaEventDisableTiming));
    HP_CUDA(cudaEventCreateWithFlags(&s.e_cmp, cudaEventDisableTiming));
    HP_CUDA(cudaEventCreateWithFlags(&s.e_done, cudaEventDisableTiming));
  }
  *out = p;
  return MRT_OK;
fail:
  mrt_host_pipeline_destroy(p);
  return rc;
}

int mrt_host_pipeline_submit(MrtHostPipeline* p, const MrtParams* params, const MrtCamera* cams, int32_t nviews,
                             const float* planar_host, const float* tf_host, int32_t tfN, float* out_rgba_host,
                             int64_t* ticket) {
  if (!p) return MRT_ERR_BAD_ARG;
  int rc = MRT_OK;
#define HP_REQ(cond, msg) do { if (!(cond)) { snprintf(p->err, sizeof(p->err), "submit: %s", msg); return MRT_ERR_BAD_ARG; } } while (0)
  HP_REQ(params && cams && planar_host && out_rgba_host, "null pointer");
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
    // ---- upload
    HP_CUDA(cudaMemcpyAsync(s.d_planar, planar_host, p->planar_bytes, cudaMemcpyHostToDevice, p->s_h2d));
    if (params->tfMode)
      HP_CUDA(cudaMemcpyAsync(s.d_tf, tf_host, (size_t)tfN * 4 * sizeof(float), cudaMemcpyHostToDevice, p->s_h2d));
    HP_CUDA(cudaEventRecord(s.e_h2d, p->s_h2d));
    // ---- prepare + march
    HP_CUDA(cudaStreamWaitEvent(p->s_cmp, s.e_h2d, 0));
    MrtParams P = *params;
    const bool skip = P.skipEmpty && P.tMode == 0;
    int Ce = p->C;
    if (p->C > 1) {                    // modality fold (+ occupancy of the folded field in the same pass)
      if (skip) rc = mrt_fold_volume_occupancy_f32(&P, s.d_planar, p->C, (float*)s.d_packed, s.d_minmax, p->s_cmp);
      else rc = mrt_fold_volume_f32(&P, s.d_planar, p->C, (float*)s.d_packed, p->s_cmp);
      P.volEnabled[0] = 1; P.volEnabled[1] = P.volEnabled[2] = P.volEnabled[3] = 0;
      P.volWeight[0] = 1.0f;
      Ce = 1;
    } else {
      rc = mrt_pack_volume_f32(s.d_planar, 1, p->X, p->Y, p->Z, s.d_packed, p->s_cmp);
      if (rc == MRT_OK && skip) rc = mrt_build_occupancy(s.d_packed, 1, p->X, p->Y, p->Z, s.d_minmax, p->s_cmp);
    }
    if (rc == MRT_OK && skip)
      rc = mrt_classify_bricks(&P, s.d_minmax, Ce, s.d_tf, tfN, nullptr, nullptr, s.d_levels, 0, p->s_cmp);
    if (rc == MRT_OK)
      rc = mrt_render_forward_batch(&P, cams, nviews, s.d_packed, Ce, s.d_tf, tfN, skip ? s.d_levels : nullptr,
                                    nullptr, nullptr, s.d_frames, nullptr, nullptr, 0,
                                    mrt_tile_count(p->W, p->H), p->s_cmp);
    if (rc != MRT_OK) { snprintf(p->err, sizeof(p->err), "submit: %s", mrt_last_error()); return rc; }
    HP_CUDA(cudaEventRecord(s.e_cmp, p->s_cmp));
    // ---- download
    HP_CUDA(cudaStreamWaitEvent(p->s_d2h, s.e_cmp, 0));
    HP_CUDA(cudaMemcpyAsync(out_rgba_host, s.d_frames, p->frame_bytes * nviews, cudaMemcpyDeviceToHost, p->s_d2h));
    HP_CUDA(cudaEventRecord(s.e_done, p->s_d2h));
    // the next upload into this slot must not overtake this step's compute (it reads d_planar):
    // guaranteed by the cudaEventSynchronize(e_done) above, e_done being recorded after e_cmp
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

}  // extern "C"
