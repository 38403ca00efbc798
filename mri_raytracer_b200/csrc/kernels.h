// kernels.h — internal launcher prototypes (one per .cu translation unit).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

struct KParams;

cudaError_t mrt_launch_forward(const KParams& P, const float* cams, int nviews, int packed_ch, const void* vol,
                               const float* tf, const uint8_t* levels, const int32_t* labels, const int32_t* preds,
                               float* out_rgba, float* out_T, int32_t* out_counts, cudaStream_t st);

cudaError_t mrt_launch_forward_strips(const KParams& P, int packed_ch, const void* vol, const float* tf,
                                      const uint8_t* levels, float* const* strip_out, int nstrips, int strip_rows,
                                      cudaStream_t st);

cudaError_t mrt_launch_forward_sparse(const KParams& P, const float* cams, int nviews, int packed_ch, const void* vol,
                                      const float* tf, const uint8_t* levels, float* out_rgba, const int32_t* spans,
                                      int store_outside, cudaStream_t st, float* const* view_base = nullptr,
                                      int row_mod = 0, int row_rem = 0);
cudaError_t mrt_launch_view_spans(const KParams& P, const float* cams, int nviews, const uint8_t* levels, int32_t* spans,
                                  cudaStream_t st);
cudaError_t mrt_launch_fill_outside(const KParams& P, int nviews, const int32_t* spans, const int32_t* prev_spans, float* out_rgba,
                                    cudaStream_t st);

// staged-brick (TMA + shared memory) variant of the single-view march, forward_tma.cu
cudaError_t mrt_launch_forward_tma(const KParams& P, int box_edge, int tile, const void* vol, const float* tf,
                                   const uint8_t* levels, float* out_rgba, void* stats, cudaStream_t st);

cudaError_t mrt_launch_forward_ckpt(const KParams& P, const float* cams, int nviews, int packed_ch, const void* vol,
                                    const float* tf, const uint8_t* levels, const int32_t* labels, const int32_t* preds,
                                    float* out_rgba, float* ck, int seg_slots, int nseg, int32_t* k_end, int32_t* warp_kmax,
                                    cudaStream_t st, bool clear_aux = true);

// Everything mrt_launch_backward takes besides geometry and the volume.  ck / k_end / warp_kmax
// (all three, from mrt_launch_forward_ckpt) switch on the segment-parallel path.
struct MrtBwdArgs {
  const float* tf; const uint8_t* flat_levels; const float* minmax;
  const int32_t* labels; const int32_t* preds;
  const float* out_rgba; const float* dL_dout;
  const float* ck; int seg_slots, nseg; const int32_t* k_end; const int32_t* warp_kmax;
  void* dvol; float* dtf; void* scratch; float* dray; void* stats;
  // fused MSE loss (mrt_train_step_mse): dL_dout == nullptr and the kernel forms G = gscale * (out_rgba - target)
  // per pixel itself; scratch_zeroed: the caller already cleared the counters and the privatised dL/dtf block
  const float* target; float gscale; int scratch_zeroed;
  // gradient buffer with its own element pitches (0 = the volume's packed pitches): the one-call training step lets
  // the adjoint reduce a single-modality gradient straight into the caller's planar [Z][Y][X] tensor
  uint32_t grad_pitchY, grad_pitchZ;
  int no_dtf_reduce;           // leave the privatised dL/dtf copies where they are (the caller reduces several launches at once)
  void* shared_priv;           // privatised dL/dtf block shared by several launches (mrt_bwd_zeroed_scratch_bytes - 256 bytes, zeroed)
};
// loss[0] = mean((a - b)^2) over n floats (n % 4 == 0), deterministic (fixed-order two-stage sum).
// `work` = MRT_MSE_WORK_BYTES bytes, zero before the FIRST use (the kernel leaves it ready for the next).
#define MRT_MSE_WORK_BYTES 4096
cudaError_t mrt_launch_mse(const float* a, const float* b, size_t n, void* work, float* loss, cudaStream_t st);
size_t mrt_bwd_zeroed_scratch_bytes(int ntf);
cudaError_t mrt_launch_backward(const KParams& P, const float* cams, int nviews, int packed_ch, const void* vol,
                                const MrtBwdArgs& A, cudaStream_t st);
size_t mrt_bwd_scratch_bytes(int W, int H, int nviews, int ntf, int nseg);
// dtf[j] += sum over `ncopies` privatised [ntf][2] float4 accumulators (lo -> entry j, hi -> entry j+1)
cudaError_t mrt_launch_dtf_reduce(const void* priv, int ncopies, int ntf, float* dtf, cudaStream_t st);

// adaptive (inverse-CDF) sampling, forward (bwd = false: writes out_rgba) and backward
cudaError_t mrt_launch_adaptive(const KParams& P, int K, int J, float eps_w, int packed_ch, const void* vol, const float* tf,
                                float* out_rgba, const float* dL_dout, void* dvol, float* dtf, void* scratch, bool bwd,
                                cudaStream_t st);
size_t mrt_adaptive_scratch(int ntf);

cudaError_t mrt_launch_pack_f16(const void* planar_f16, int X, int Y, int Z, void* packed, cudaStream_t st);
cudaError_t mrt_launch_unpack_f16(const void* packed, int X, int Y, int Z, void* planar_f16, cudaStream_t st);
cudaError_t mrt_launch_build_occupancy_f16(const void* packed, int X, int Y, int Z, float* minmax, cudaStream_t st);
cudaError_t mrt_launch_pack_u8(const void* planar_u8, int X, int Y, int Z, void* packed, cudaStream_t st);
cudaError_t mrt_launch_build_occupancy_u8(const void* packed, int X, int Y, int Z, float* minmax, cudaStream_t st);
cudaError_t mrt_launch_pack_quad(const float* packed1, int X, int Y, int Z, void* quad, cudaStream_t st);
cudaError_t mrt_launch_pack_quad_f16(const void* packed_f16, int X, int Y, int Z, void* quad, cudaStream_t st);
cudaError_t mrt_launch_pack(const float* planar, int C, int X, int Y, int Z, void* packed, cudaStream_t st);
cudaError_t mrt_launch_unpack(const void* packed, int C, int X, int Y, int Z, float* planar, cudaStream_t st);

cudaError_t mrt_launch_fold(const float* planar, int C, int X, int Y, int Z, const float* wgt, float inv_wsum,
                            float* folded, cudaStream_t st);
cudaError_t mrt_launch_fold_occ(const float* planar, int C, int X, int Y, int Z, const float* wgt, float inv_wsum,
                                float* folded, float* minmax, cudaStream_t st);
cudaError_t mrt_launch_fold_occ_quad(const float* planar, int C, int X, int Y, int Z, const float* wgt, float inv_wsum,
                                     float* folded, void* quad, float* minmax, cudaStream_t st);
cudaError_t mrt_launch_unfold_grad(const float* dfolded, int C, int X, int Y, int Z, const float* wgt, float inv_wsum,
                                   float* dplanar, cudaStream_t st);

cudaError_t mrt_launch_build_occupancy(const void* packed, int packed_ch, int X, int Y, int Z,
                                       float* minmax, cudaStream_t st);
cudaError_t mrt_launch_label_occupancy(const int32_t* labels, int X, int Y, int Z, uint8_t* any, cudaStream_t st);
cudaError_t mrt_launch_classify(const KParams& P, const float* minmax, int packed_ch, const float* tf,
                                const uint8_t* seg_any, const uint8_t* pred_any, uint8_t* levels,
                                bool require_flat, cudaStream_t st);

cudaError_t mrt_launch_tile_map(int W, int H, int32_t* out_tile, int32_t* out_lane, cudaStream_t st);
cudaError_t mrt_launch_gather_probe(const void* buf, size_t bytes, size_t n, uint32_t seed, float* out,
                                    cudaStream_t st);
cudaError_t mrt_launch_composite(const float* partials, int K, const int32_t* order, size_t npix,
                                 float bgr, float bgg, float bgb, int alphaMode, float* const* outs, int nouts,
                                 cudaStream_t st);
cudaError_t mrt_launch_bc4(const uint8_t* blocks, int W, int H, int D, uint8_t* out, cudaStream_t st);
cudaError_t mrt_launch_u8_to_f32(const uint8_t* in, size_t n, float* out, cudaStream_t st);
cudaError_t mrt_launch_normalize(const float* in, size_t n, float vmin, float rng, float* out, cudaStream_t st);

cudaError_t mrt_launch_inr(const float* mods, int M, int X, int Y, int Z, const float* weights, const int32_t* layer_dims,
                           int n_layers, int fourier_freqs, int32_t* labels, float* logits, int impl, cudaStream_t st);

struct MrtSlabParams;
cudaError_t mrt_launch_slab(const MrtSlabParams& P, float tan_half, const uint8_t* vol, float* out,
                            int tile_begin, int tile_end, cudaStream_t st);

// The skip-level buffer is uint8[nbricks] followed (16-byte aligned) by int32[8]: the bounding
// box of the active bricks in brick units, stored as maxima (-lox, -loy, -loz, hix, hiy, hiz, -, -).
static inline size_t mrt_levels_box_offset(size_t nbricks) { return (nbricks + 15) & ~(size_t)15; }

static inline int mrt_packed_channels(int C) { return C <= 1 ? 1 : (C == 2 ? 2 : 4); }

// Skewed pitches of the packed layout (see include/mrt.h "volume layout"): with S voxels per
// 128-byte line, pitchY = S/4 and pitchZ = S/2 (mod S) put the cells of a small 3-D
// neighbourhood into distinct L1 data banks, so a warp's gather is not serialised by bank
// conflicts between rows/slices (row pitches that are multiples of 128 B alias every row).
// `elem_bytes` = 4 (fp32 voxels), 2 (fp16), 1 (u8) or 16 (quad of fp32), the last three single channel.
static inline void mrt_layout_e(int packed_ch, int elem_bytes, int X, int Y, int Z, int64_t* pitchY, int64_t* pitchZ) {
  const int64_t S = 128 / (packed_ch * elem_bytes);
  int64_t py = X;
  while (py % S != S / 4) ++py;
  int64_t pz = py * Y;
  while (pz % S != S / 2) ++pz;
  (void)Z;
  *pitchY = py; *pitchZ = pz;
}
static inline void mrt_layout(int packed_ch, int X, int Y, int Z, int64_t* pitchY, int64_t* pitchZ) {
  mrt_layout_e(packed_ch, 4, X, Y, Z, pitchY, pitchZ);
}
