// inr.cu — INR inference over a whole volume: the producer of the
// prediction label volume `gPreds` that the renderer overlays.
//
// Reference: inr/inr/model.py — fourier_features :11-18, build_input :21-23, apply_mlp :43-50,
// predict_volume :119-141 — called by the viewer at inr/viewer/brats_viewer.py:250-310 (JAX, 200 k
// voxels per chunk).  Here: ONE fused kernel, one thread per voxel: coordinates -> Fourier
// features -> dense/ReLU chain -> argmax, with every layer's weights resident in shared memory
// (read as warp-wide broadcasts) and the activations in registers; input voxels are read in the
// renderer's own planar [M][Z][Y][X] layout and labels are written as int32 [Z][Y][X], the layout
// Volume(preds=...) takes — so the transpose of brats_viewer.py:297 disappears.
// fp32 FFMA on the CUDA cores: the reference network is 31 -> 64 x 4 -> 4 (29 kFLOP per voxel,
// 0.26 TFLOP per BraTS case).  A tcgen05 version (128-voxel tiles, TMEM accumulators) is the
// planned follow-up; argmax parity wants fp32 accumulation either way.
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.h"

#define MRT_INR_MAX_LAYERS 8
#define MRT_INR_MAX_CLASSES 8

struct InrNet {
  int n_layers;                               // dense layers (hidden ones have ReLU, the last has none)
  int dims[MRT_INR_MAX_LAYERS + 1];           // dims[0] = input width, dims[n_layers] = classes
  int src_off[MRT_INR_MAX_LAYERS];            // offset of layer l's W (then b) in the caller's packed weights
  int k;                                      // Fourier frequencies
  int M;                                      // modalities
};

// shared-memory image of the weights: hidden layers padded to [HID][HID] (+[HID] bias), the last
// layer to [HID][MRT_INR_MAX_CLASSES] (+[MRT_INR_MAX_CLASSES]); zero padding makes every loop a
// compile-time HID x HID (or HID x 8) nest with no predicates
template <int HID>
__host__ __device__ inline int inr_smem_floats(int n_layers) {
  return (n_layers - 1) * (HID * HID + HID) + HID * MRT_INR_MAX_CLASSES + MRT_INR_MAX_CLASSES;
}

template <int HID>
__global__ void __launch_bounds__(128)
mrt_inr_kernel(const __grid_constant__ InrNet N, const float* __restrict__ mods, int X, int Y, int Z,
               const float* __restrict__ wts, int32_t* __restrict__ labels, float* __restrict__ logits) {
  extern __shared__ __align__(16) float s_w[];
  const int total = inr_smem_floats<HID>(N.n_layers);
  for (int i = threadIdx.x; i < total; i += blockDim.x) s_w[i] = 0.0f;
  __syncthreads();
  for (int l = 0; l < N.n_layers; ++l) {
    const int in = N.dims[l], out = N.dims[l + 1];
    const bool last = (l == N.n_layers - 1);
    const int ld = last ? MRT_INR_MAX_CLASSES : HID;
    float* W = s_w + l * (HID * HID + HID);
    float* b = W + HID * ld;
    const float* src = wts + N.src_off[l];
    for (int i = threadIdx.x; i < in * out; i += blockDim.x) W[(i / out) * ld + (i % out)] = __ldg(src + i);
    for (int i = threadIdx.x; i < out; i += blockDim.x) b[i] = __ldg(src + in * out + i);
  }
  __syncthreads();

  const size_t nvox = (size_t)X * Y * Z;
  const int ncls = N.dims[N.n_layers];
  for (size_t vox = (size_t)blockIdx.x * blockDim.x + threadIdx.x; vox < nvox; vox += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(vox % X), y = (int)((vox / X) % Y), z = (int)(vox / ((size_t)X * Y));
    float h[HID], g[HID];
#pragma unroll
    for (int i = 0; i < HID; ++i) h[i] = 0.0f;
    // model.py:128 normalises in float64 ((grid / (n-1)) * 2 - 1) and then casts to float32
    float c[3];
    c[0] = (float)(((double)x / (double)(X - 1)) * 2.0 - 1.0);
    c[1] = (float)(((double)y / (double)(Y - 1)) * 2.0 - 1.0);
    c[2] = (float)(((double)z / (double)(Z - 1)) * 2.0 - 1.0);
    // build_input (:21-23): [coords | per coordinate: sin(f pi x) f=1..k, then cos | intensities]; the
    // writes below use compile-time indices only (the runtime bounds are predicates), so h[] stays in registers
    const float pi = 3.14159265358979323846f;
#pragma unroll
    for (int i = 0; i < HID; ++i) {
      float v = 0.0f;
      const int j = i - 3;                                   // index into the Fourier block
      if (i < 3) {
        v = (i == 0) ? c[0] : ((i == 1) ? c[1] : c[2]);
      } else if (j < 6 * N.k) {
        const int d = j / (2 * N.k), r = j - d * 2 * N.k;    // coordinate, position inside its [sin.. | cos..] group
        const int f = (r < N.k ? r : r - N.k) + 1;
        const float cd = (d == 0) ? c[0] : ((d == 1) ? c[1] : c[2]);
        const float ang = __fmul_rn(__fmul_rn(cd, (float)f), pi);            // :14 (coords * freqs) * pi, in fp32
        v = (r < N.k) ? sinf(ang) : cosf(ang);
      } else if (j - 6 * N.k < N.M) {
        v = __ldg(mods + (size_t)(j - 6 * N.k) * nvox + vox);
      }
      h[i] = v;
    }
    // apply_mlp (:43-50)
    for (int l = 0; l < N.n_layers - 1; ++l) {
      const float* W = s_w + l * (HID * HID + HID);
      const float* b = W + HID * HID;
#pragma unroll
      for (int j = 0; j < HID; ++j) g[j] = b[j];
#pragma unroll
      for (int i = 0; i < HID; ++i) {
        const float hi = h[i];
        const float4* row = reinterpret_cast<const float4*>(W + i * HID);
#pragma unroll
        for (int j4 = 0; j4 < HID / 4; ++j4) {
          const float4 w = row[j4];                           // warp-wide broadcast
          g[4 * j4 + 0] = fmaf(hi, w.x, g[4 * j4 + 0]); g[4 * j4 + 1] = fmaf(hi, w.y, g[4 * j4 + 1]);
          g[4 * j4 + 2] = fmaf(hi, w.z, g[4 * j4 + 2]); g[4 * j4 + 3] = fmaf(hi, w.w, g[4 * j4 + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < HID; ++j) h[j] = fmaxf(g[j], 0.0f);
    }
    {
      const float* W = s_w + (N.n_layers - 1) * (HID * HID + HID);
      const float* b = W + HID * MRT_INR_MAX_CLASSES;
      float o[MRT_INR_MAX_CLASSES];
#pragma unroll
      for (int j = 0; j < MRT_INR_MAX_CLASSES; ++j) o[j] = b[j];
#pragma unroll
      for (int i = 0; i < HID; ++i) {
        const float hi = h[i];
        const float4* row = reinterpret_cast<const float4*>(W + i * MRT_INR_MAX_CLASSES);
        const float4 w0 = row[0], w1 = row[1];
        o[0] = fmaf(hi, w0.x, o[0]); o[1] = fmaf(hi, w0.y, o[1]); o[2] = fmaf(hi, w0.z, o[2]); o[3] = fmaf(hi, w0.w, o[3]);
        o[4] = fmaf(hi, w1.x, o[4]); o[5] = fmaf(hi, w1.y, o[5]); o[6] = fmaf(hi, w1.z, o[6]); o[7] = fmaf(hi, w1.w, o[7]);
      }
      int best = 0; float bv = o[0];
#pragma unroll
      for (int j = 1; j < MRT_INR_MAX_CLASSES; ++j) if (j < ncls && o[j] > bv) { bv = o[j]; best = j; }   // first maximum, like argmax
      labels[vox] = best;
      if (logits != nullptr) {
#pragma unroll
        for (int j = 0; j < MRT_INR_MAX_CLASSES; ++j) if (j < ncls) logits[vox * ncls + j] = o[j];
      }
    }
  }
}

// Second version: the activations live in shared memory, one column per thread ([HID][block]:
// conflict-free, no barrier needed — a thread only ever touches its own column), so the loop over the
// input index can stay ROLLED (a rolled loop cannot index a register array): ~330 instructions of
// loop body instead of ~5000 fully unrolled ones (instruction-cache friendly), ~100 registers instead
// of 180, 16 warps per SM instead of 8 sharing one copy of the weights.
#define MRT_INR_BLOCK 512
template <int HID>
__global__ void __launch_bounds__(MRT_INR_BLOCK)
mrt_inr_kernel2(const __grid_constant__ InrNet N, const float* __restrict__ mods, int X, int Y, int Z,
                const float* __restrict__ wts, int32_t* __restrict__ labels, float* __restrict__ logits) {
  extern __shared__ __align__(16) float s_w[];
  const int total = inr_smem_floats<HID>(N.n_layers);
  float* s_h = s_w + ((total + 3) & ~3);                 // [HID][MRT_INR_BLOCK]
  for (int i = threadIdx.x; i < total; i += blockDim.x) s_w[i] = 0.0f;
  __syncthreads();
  for (int l = 0; l < N.n_layers; ++l) {
    const int in = N.dims[l], out = N.dims[l + 1];
    const bool last = (l == N.n_layers - 1);
    const int ld = last ? MRT_INR_MAX_CLASSES : HID;
    float* W = s_w + l * (HID * HID + HID);
    float* b = W + HID * ld;
    const float* src = wts + N.src_off[l];
    for (int i = threadIdx.x; i < in * out; i += blockDim.x) W[(i / out) * ld + (i % out)] = __ldg(src + i);
    for (int i = threadIdx.x; i < out; i += blockDim.x) b[i] = __ldg(src + in * out + i);
  }
  __syncthreads();

  float* hcol = s_h + threadIdx.x;                        // my column: element i at hcol[i * MRT_INR_BLOCK]
  const size_t nvox = (size_t)X * Y * Z;
  const int ncls = N.dims[N.n_layers];
  for (size_t vox = (size_t)blockIdx.x * blockDim.x + threadIdx.x; vox < nvox; vox += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(vox % X), y = (int)((vox / X) % Y), z = (int)(vox / ((size_t)X * Y));
    float c[3];
    c[0] = (float)(((double)x / (double)(X - 1)) * 2.0 - 1.0);      // model.py:128 (float64, then cast)
    c[1] = (float)(((double)y / (double)(Y - 1)) * 2.0 - 1.0);
    c[2] = (float)(((double)z / (double)(Z - 1)) * 2.0 - 1.0);
    const float pi = 3.14159265358979323846f;
    for (int i = 0; i < HID; ++i) {                       // build_input (:21-23)
      float v = 0.0f;
      const int j = i - 3;
      if (i < 3) {
        v = (i == 0) ? c[0] : ((i == 1) ? c[1] : c[2]);
      } else if (j < 6 * N.k) {
        const int d = j / (2 * N.k), r = j - d * 2 * N.k;
        const int f = (r < N.k ? r : r - N.k) + 1;
        const float cd = (d == 0) ? c[0] : ((d == 1) ? c[1] : c[2]);
        const float ang = __fmul_rn(__fmul_rn(cd, (float)f), pi);    // :14 (coords * freqs) * pi, in fp32
        v = (r < N.k) ? sinf(ang) : cosf(ang);
      } else if (j - 6 * N.k < N.M) {
        v = __ldg(mods + (size_t)(j - 6 * N.k) * nvox + vox);
      }
      hcol[i * MRT_INR_BLOCK] = v;
    }
    float g[HID];
    for (int l = 0; l < N.n_layers - 1; ++l) {            // apply_mlp (:43-50), hidden layers
      const float* W = s_w + l * (HID * HID + HID);
      const float* b = W + HID * HID;
#pragma unroll
      for (int j = 0; j < HID; ++j) g[j] = b[j];
#pragma unroll 4
      for (int i = 0; i < HID; ++i) {
        const float hi = hcol[i * MRT_INR_BLOCK];
        const float4* row = reinterpret_cast<const float4*>(W + i * HID);
#pragma unroll
        for (int j4 = 0; j4 < HID / 4; ++j4) {
          const float4 w = row[j4];                       // warp-wide broadcast
          g[4 * j4 + 0] = fmaf(hi, w.x, g[4 * j4 + 0]); g[4 * j4 + 1] = fmaf(hi, w.y, g[4 * j4 + 1]);
          g[4 * j4 + 2] = fmaf(hi, w.z, g[4 * j4 + 2]); g[4 * j4 + 3] = fmaf(hi, w.w, g[4 * j4 + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < HID; ++j) hcol[j * MRT_INR_BLOCK] = fmaxf(g[j], 0.0f);
    }
    {
      const float* W = s_w + (N.n_layers - 1) * (HID * HID + HID);
      const float* b = W + HID * MRT_INR_MAX_CLASSES;
      float o[MRT_INR_MAX_CLASSES];
#pragma unroll
      for (int j = 0; j < MRT_INR_MAX_CLASSES; ++j) o[j] = b[j];
#pragma unroll 4
      for (int i = 0; i < HID; ++i) {
        const float hi = hcol[i * MRT_INR_BLOCK];
        const float4* row = reinterpret_cast<const float4*>(W + i * MRT_INR_MAX_CLASSES);
        const float4 w0 = row[0], w1 = row[1];
        o[0] = fmaf(hi, w0.x, o[0]); o[1] = fmaf(hi, w0.y, o[1]); o[2] = fmaf(hi, w0.z, o[2]); o[3] = fmaf(hi, w0.w, o[3]);
        o[4] = fmaf(hi, w1.x, o[4]); o[5] = fmaf(hi, w1.y, o[5]); o[6] = fmaf(hi, w1.z, o[6]); o[7] = fmaf(hi, w1.w, o[7]);
      }
      int best = 0; float bv = o[0];
#pragma unroll
      for (int j = 1; j < MRT_INR_MAX_CLASSES; ++j) if (j < ncls && o[j] > bv) { bv = o[j]; best = j; }   // first maximum, like argmax
      labels[vox] = best;
      if (logits != nullptr) {
#pragma unroll
        for (int j = 0; j < MRT_INR_MAX_CLASSES; ++j) if (j < ncls) logits[vox * ncls + j] = o[j];
      }
    }
  }
}

template <int HID>
static cudaError_t launch_inr(const InrNet& N, const float* mods, int X, int Y, int Z, const float* wts,
                              int32_t* labels, float* logits, cudaStream_t st) {
  const size_t nvox = (size_t)X * Y * Z;
#ifndef MRT_INR_V1
  const size_t wfl = ((size_t)inr_smem_floats<HID>(N.n_layers) + 3) & ~(size_t)3;
  const size_t smem = (wfl + (size_t)HID * MRT_INR_BLOCK) * sizeof(float);
  if (smem <= 227 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mrt_inr_kernel2<HID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    size_t grid = (nvox + MRT_INR_BLOCK - 1) / MRT_INR_BLOCK;
    if (grid > 148 * 2) grid = 148 * 2;          // the weights are staged once per CTA
    mrt_inr_kernel2<HID><<<(int)grid, MRT_INR_BLOCK, smem, st>>>(N, mods, X, Y, Z, wts, labels, logits);
    return cudaGetLastError();
  }
#endif
  // first version (activations in registers, fully unrolled): also the fall-back for very deep networks
  const size_t smem1 = (size_t)inr_smem_floats<HID>(N.n_layers) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(mrt_inr_kernel<HID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
  if (e != cudaSuccess) return e;
  size_t grid = (nvox + 127) / 128;
  if (grid > 148 * 8) grid = 148 * 8;            // persistent-ish: the weights are staged once per CTA
  mrt_inr_kernel<HID><<<(int)grid, 128, smem1, st>>>(N, mods, X, Y, Z, wts, labels, logits);
  return cudaGetLastError();
}

// layer_dims: n_layers + 1 widths; weights: per layer W[in][out] row-major followed by b[out]
cudaError_t mrt_launch_inr(const float* mods, int M, int X, int Y, int Z, const float* weights, const int32_t* layer_dims,
                           int n_layers, int fourier_freqs, int32_t* labels, float* logits, cudaStream_t st) {
  InrNet N = {};
  N.n_layers = n_layers; N.k = fourier_freqs; N.M = M;
  int off = 0, hid = 0;
  for (int l = 0; l <= n_layers; ++l) N.dims[l] = layer_dims[l];
  for (int l = 0; l < n_layers; ++l) {
    N.src_off[l] = off;
    off += N.dims[l] * N.dims[l + 1] + N.dims[l + 1];
    if (l < n_layers - 1 && N.dims[l + 1] > hid) hid = N.dims[l + 1];
  }
  if (N.dims[0] > hid) hid = N.dims[0];
  if (hid <= 32) return launch_inr<32>(N, mods, X, Y, Z, weights, labels, logits, st);
  if (hid <= 64) return launch_inr<64>(N, mods, X, Y, Z, weights, labels, logits, st);
  return cudaErrorInvalidValue;
}
