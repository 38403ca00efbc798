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
// Two implementations of the same network:
//   * mrt_inr_tc_kernel (default): the dense chain on the 5th-generation tensor cores — tcgen05.mma
//     kind::tf32 with 128-voxel M tiles, fp32 accumulators in TMEM, the activations fed back as the A
//     operand FROM TMEM (bias + ReLU applied by the epilogue warps between tcgen05.ld and
//     tcgen05.st), the weights resident in shared memory as K-major UMMA operands.  Every product
//     is evaluated as a 3-term TF32 split (hi*hi + lo*hi + hi*lo, fp32 accumulate), so the logits
//     agree with fp32 FFMA to ~1e-6 and argmax parity holds — a single-pass bf16/tf32 product
//     would not keep labels stable where the top-2 logits are close;
//   * mrt_inr_kernel2: fp32 FFMA on the CUDA cores, one thread per voxel (the parity reference for
//     the tensor-core kernel and the fallback for networks whose weights do not fit shared memory).
// The reference network is 31 -> 64 x 4 -> 4 (29 kFLOP per voxel, 0.26 TFLOP per BraTS case).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "kernels.h"

#define MRT_INR_MAX_LAYERS 8
#define MRT_INR_MAX_CLASSES 8

struct InrNet {
  int n_layers;                               // dense layers (hidden ones have ReLU, the last has none)
  int dims[MRT_INR_MAX_LAYERS + 1];           // dims[0] = input width, dims[n_layers] = classes
  int src_off[MRT_INR_MAX_LAYERS];            // offset of layer l's W (then b) in the caller's packed weights
  int k;                                      // Fourier frequencies
  int M;                                      // modalities
  int dbg;                                    // dev only (MRT_INR_DBG): 1 = issue no MMA, 2 = skip the epilogue arithmetic, 4 = skip the layer-0 operand
};

// shared-memory image of the weights: hidden layers padded to [HID][HID] (+[HID] bias), the last
// layer to [HID][MRT_INR_MAX_CLASSES] (+[MRT_INR_MAX_CLASSES]); zero padding makes every loop a
// compile-time HID x HID (or HID x 8) nest with no predicates
template <int HID>
__host__ __device__ inline int inr_smem_floats(int n_layers) {
  return (n_layers - 1) * (HID * HID + HID) + HID * MRT_INR_MAX_CLASSES + MRT_INR_MAX_CLASSES;
}

// Second version: the activations live in shared memory, one column per thread ([HID][block]:
// conflict-free, no barrier needed — a thread only ever touches its own column), so the loop over the
// input index can stay ROLLED (a rolled loop cannot index a register array): ~330 instructions of
// loop body instead of ~5000 fully unrolled ones (instruction-cache friendly), ~100 registers instead
// of 180, 16 warps per SM instead of 8 sharing one copy of the weights.
template <int HID, int MRT_INR_BLOCK>
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

// ------------------------------------------------------------------------------------------------
// Tensor-core version (tcgen05 / TMEM).  One CTA per SM, persistent over 128-voxel tiles:
//   warps 0-7, 8-15 : the epilogue warps of tile "slot" 0 and 1 (TMEM lane = voxel row); within a slot
//                     warps 0-3 own activation columns 0-31, warps 4-7 columns 32-63
//   warps 16, 17    : one elected thread per slot issues that slot's tcgen05.mma
// Per slot the TMEM holds D [64 columns] | A_hi [64] | A_lo [64] | ones [8].  A layer is
//   D = ones*B_bias + A_lo*W_hi + A_hi*W_lo + A_hi*W_hi      (tf32 operands, fp32 accumulate in TMEM)
// with A read from TMEM and W (K-major, no swizzle: 8x16-byte core matrices, K chunks LBO apart,
// 8-row groups SBO = 128 B apart) from shared memory; the bias rides on a constant A block (two
// columns of ones against b_hi, b_lo), so the epilogue has nothing to add.  While the tensor core
// works on one slot the other slot's warpgroup runs its epilogue: tcgen05.ld D -> ReLU -> split into
// tf32 hi/lo -> tcgen05.st A (3 instructions per activation: FMNMX, LOP, FADD — cvt.rna.tf32 costs
// five SASS instructions, so hi is the TRUNCATED value: the tensor core ignores the low 13 mantissa
// bits anyway and lo = v - hi is exact).  The two hand-offs per slot are
// mbarriers: a_ready (one arrival per epilogue warp, after its lanes' tcgen05.st have completed) and
// d_ready (tcgen05.commit; one lane per warp waits and releases its warp).
// The layer-0 operand comes from per-axis tables built once per CTA in shared memory: for every x,
// y and z index its normalised coordinate and the 2k Fourier features (model.py:11-18 depend on one
// coordinate each), so a voxel costs three table rows and M loads instead of 6k sin/cos evaluations.
#define INR_TC_EPI_WARPS 16        // 2 slots x 2 column halves x 4 warps (TMEM lanes 32*(warp%4)..+31)
#define INR_TC_THREADS (32 * (INR_TC_EPI_WARPS + 2))
#define INR_TC_SLOT_COLS 256
#define INR_TC_HID 64
#define INR_TC_NLAST 16

__device__ __forceinline__ uint32_t inr_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void inr_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(inr_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void inr_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(inr_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void inr_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "INR_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra INR_DONE_%=;\n"
      "bra INR_WAIT_%=;\n"
      "INR_DONE_%=:\n"
      "}\n" :: "r"(inr_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool inr_elect_one() {             // one lane of the (converged) warp
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t inr_tf32(float x) {      // round to nearest tf32 (low 13 mantissa bits zero)
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
// shared-memory matrix descriptor: K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t inr_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, both K-major, M = 128
__device__ __forceinline__ uint32_t inr_idesc(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void inr_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void inr_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(inr_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void inr_tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
__device__ __forceinline__ void inr_tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void inr_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void inr_tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void inr_tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// byte offsets of the shared-memory image
struct InrTcLayout {
  int col_off;                                  // int32[64]: where column i of the layer-0 operand comes from
  int tab_off, tab_row;                         // float[X+Y+Z][tab_row]: coordinate + 2k Fourier features per axis index
  int b_off[MRT_INR_MAX_LAYERS];                // bias images: [np][8] K-major, k=0 -> b_hi, k=1 -> b_lo
  int w_off[MRT_INR_MAX_LAYERS], kp[MRT_INR_MAX_LAYERS], np[MRT_INR_MAX_LAYERS];
  int total;
};
static inline InrTcLayout inr_tc_layout(const InrNet& N, int X, int Y, int Z) {
  InrTcLayout L = {};
  int off = 64;                                               // [0,4) TMEM base, [16,48) four mbarriers
  L.col_off = off; off += 64 * (int)sizeof(int32_t);
  L.tab_row = 1 + 2 * N.k;
  L.tab_off = off; off += (X + Y + Z) * L.tab_row * (int)sizeof(float);
  off = (off + 127) & ~127;
  for (int l = 0; l < N.n_layers; ++l) {
    L.kp[l] = (l == 0) ? ((N.dims[0] + 7) & ~7) : INR_TC_HID;
    L.np[l] = (l == N.n_layers - 1) ? INR_TC_NLAST : INR_TC_HID;
    L.b_off[l] = off; off += L.np[l] * 8 * (int)sizeof(float);
    L.w_off[l] = off; off += 2 * L.kp[l] * L.np[l] * (int)sizeof(float);      // hi image, then lo image
  }
  L.total = off;
  return L;
}

__global__ void __launch_bounds__(INR_TC_THREADS, 1)
mrt_inr_tc_kernel(const __grid_constant__ InrNet N, const __grid_constant__ InrTcLayout L, const float* __restrict__ mods,
                  int X, int Y, int Z, const float* __restrict__ wts, int32_t* __restrict__ labels, float* __restrict__ logits) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_raw);
  uint64_t* a_ready = reinterpret_cast<uint64_t*>(s_raw + 16);     // [2]
  uint64_t* d_ready = reinterpret_cast<uint64_t*>(s_raw + 32);     // [2]
  int32_t* s_col = reinterpret_cast<int32_t*>(s_raw + L.col_off);
  float* s_tab = reinterpret_cast<float*>(s_raw + L.tab_off);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // (warp-uniform for the compiler)
  const int nl = N.n_layers;

  // ---- one-time set-up: weight / bias images (tf32 hi / lo), axis tables, barriers, TMEM
  for (int i = threadIdx.x; i < (L.total - L.col_off) / 4; i += blockDim.x) reinterpret_cast<float*>(s_raw + L.col_off)[i] = 0.0f;
  __syncthreads();
  for (int l = 0; l < nl; ++l) {
    const int in = N.dims[l], out = N.dims[l + 1], np = L.np[l], kp = L.kp[l];
    const float* src = wts + N.src_off[l];
    uint32_t* hi = reinterpret_cast<uint32_t*>(s_raw + L.w_off[l]);
    uint32_t* lo = hi + kp * np;
    for (int i = threadIdx.x; i < in * out; i += blockDim.x) {
      const int k = i / out, n = i - k * out;                  // W[k][n]  ->  B[n][k] (K-major)
      const float w = __ldg(src + i);
      const uint32_t h = inr_tf32(w);
      const int e = (k >> 2) * (np * 4) + n * 4 + (k & 3);     // 16-byte chunk (k/4) of row n
      hi[e] = h;
      lo[e] = inr_tf32(w - __uint_as_float(h));
    }
    uint32_t* bi = reinterpret_cast<uint32_t*>(s_raw + L.b_off[l]);            // [chunk 0: k 0..3][np rows], [chunk 1]
    for (int i = threadIdx.x; i < out; i += blockDim.x) {
      const float bv = __ldg(src + in * out + i);
      const uint32_t h = inr_tf32(bv);
      bi[i * 4 + 0] = h;
      bi[i * 4 + 1] = inr_tf32(bv - __uint_as_float(h));
    }
  }
  // layer-0 column sources: 0 = zero padding; 1 + (axis << 8 | entry) = table; 0x10000 + m = modality m
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    int code = 0;
    const int jf = i - 3;
    if (i < 3) code = 1 + (i << 8);
    else if (jf < 6 * N.k) { const int dd = jf / (2 * N.k), r = jf - dd * 2 * N.k; code = 1 + ((dd << 8) | (1 + r)); }
    else if (jf - 6 * N.k < N.M) code = 0x10000 + (jf - 6 * N.k);
    s_col[i] = code;
  }
  for (int i = threadIdx.x; i < X + Y + Z; i += blockDim.x) {
    const int n = (i < X) ? X : ((i < X + Y) ? Y : Z), idx = (i < X) ? i : ((i < X + Y) ? i - X : i - X - Y);
    const float c = (float)(((double)idx / (double)(n - 1)) * 2.0 - 1.0);       // model.py:128 (float64, then cast)
    float* row = s_tab + (size_t)i * L.tab_row;
    row[0] = c;
    const float pi = 3.14159265358979323846f;
    for (int f = 1; f <= N.k; ++f) {
      const float ang = __fmul_rn(__fmul_rn(c, (float)f), pi);                  // :14 (coords * freqs) * pi, in fp32
      row[f] = sinf(ang);                                                        // :15-17 sines, then cosines
      row[N.k + f] = cosf(ang);
    }
  }
  if (threadIdx.x == 0) {
    inr_mbar_init(&a_ready[0], INR_TC_EPI_WARPS / 2); inr_mbar_init(&a_ready[1], INR_TC_EPI_WARPS / 2);   // one arrival per warp
    inr_mbar_init(&d_ready[0], 1);   inr_mbar_init(&d_ready[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(inr_smem_u32(s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the weight images are read by the tensor core (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;

  const size_t nvox = (size_t)X * Y * Z;
  const int ntiles = (int)((nvox + 127) / 128);
  const int mine = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;   // tiles blockIdx.x + i*gridDim.x

  if (warp >= INR_TC_EPI_WARPS) {
    // ---------------------------------------------------------------- MMA issuers (one warp per slot)
    // The WHOLE warp runs this loop with warp-uniform values and one elected lane issues: with a
    // divergent `if (lane == 0)` around it the compiler cannot keep the descriptors in uniform
    // registers and wraps every UTCHMMA in R2UR.BROADCAST + ELECT loops — 15 instructions and ~100
    // cycles of one thread per MMA, which made the issue loop (not the tensor core) the bottleneck.
    const int j = warp - INR_TC_EPI_WARPS;
    const int cnt = (mine + 1 - j) >> 1;
    const uint32_t d = tmem + (uint32_t)(j * INR_TC_SLOT_COLS), a_hi = d + 64, a_lo = d + 128, a_one = d + 192;
    const uint32_t desc_hi = 8u | (1u << 14);                   // SBO = 128 B, descriptor version 1 (bits 32..47 of the descriptor)
    uint32_t ph = 0;
    for (int q = 0; q < cnt; ++q) {
      for (int l = 0; l < nl; ++l) {
        const int ksteps = L.kp[l] >> 3, np = L.np[l];
        const uint32_t idesc = inr_idesc(np);
        const uint32_t lbo16 = (uint32_t)np;                    // LBO = np * 16 bytes, in 16-byte units
        const uint32_t w_hi = ((inr_smem_u32(s_raw + L.w_off[l]) & 0x3FFFFu) >> 4) | (lbo16 << 16);
        const uint32_t w_lo = w_hi + (uint32_t)((L.kp[l] * np * 4) >> 4);
        const uint32_t b_b = ((inr_smem_u32(s_raw + L.b_off[l]) & 0x3FFFFu) >> 4) | (lbo16 << 16);
        const uint32_t step = 2u * lbo16;                       // one MMA consumes two 16-byte K chunks
        inr_mbar_wait(&a_ready[j], ph);
        ph ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (inr_elect_one()) {
          if (!(N.dbg & 1)) inr_mma_ts(d, a_one, ((uint64_t)desc_hi << 32) | b_b, idesc, 0u);                  // D = bias
#pragma unroll
          for (int s8 = 0; s8 < INR_TC_HID / 8; ++s8) {
            if (s8 < ksteps && !(N.dbg & 1)) {
              const uint64_t bh = ((uint64_t)desc_hi << 32) | (w_hi + (uint32_t)s8 * step);
              const uint64_t bl = ((uint64_t)desc_hi << 32) | (w_lo + (uint32_t)s8 * step);
              inr_mma_ts(d, a_lo + 8 * s8, bh, idesc, 1u);
              inr_mma_ts(d, a_hi + 8 * s8, bl, idesc, 1u);
              inr_mma_ts(d, a_hi + 8 * s8, bh, idesc, 1u);
            }
          }
          inr_commit(&d_ready[j]);                              // arrives when every MMA above has completed
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warpgroups
    const int j = warp >> 3;                                     // slot
    const int half = (warp >> 2) & 1;                            // which 32 activation columns this warp owns
    const int row = (warp & 3) * 32 + lane;                      // TMEM lane == voxel row of the tile
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(j * INR_TC_SLOT_COLS);
    const uint32_t t_d = lane_base, t_hi = lane_base + 64, t_lo = lane_base + 128;
    const int ncls = N.dims[nl];
    const int cnt = (mine + 1 - j) >> 1;
    if (half == 0) {                                             // the constant A block of the bias MMA: ones in columns 0, 1
      uint32_t one[8] = {0x3f800000u, 0x3f800000u, 0u, 0u, 0u, 0u, 0u, 0u};
      inr_tmem_st8(lane_base + 192, one);
    }
    for (int q = 0; q < cnt; ++q) {
      const int tile = (int)blockIdx.x + (2 * q + j) * (int)gridDim.x;
      const size_t vox = (size_t)tile * 128 + row;
      const bool valid = vox < nvox;
      // ---- layer-0 operand: [coords | Fourier features | intensities] (model.py:11-23), zero padded to kp[0];
      // the two halves of the slot take alternate 16-column chunks
      {
        const unsigned vv = valid ? (unsigned)vox : 0u;         // (nvox < 2^32: checked by the launcher)
        const unsigned yz = vv / (unsigned)X, x = vv - yz * (unsigned)X, z = yz / (unsigned)Y, y = yz - z * (unsigned)Y;
        const float* rx = s_tab + (size_t)x * L.tab_row;
        const float* ry = s_tab + (size_t)(X + y) * L.tab_row;
        const float* rz = s_tab + (size_t)(X + Y + z) * L.tab_row;
        const int kp0 = L.kp[0];
        for (int c0 = 16 * half; c0 < kp0 && !(N.dbg & 4); c0 += 32) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int code = s_col[c0 + u];                      // warp-uniform
            float v = 0.0f;
            if (code >= 0x10000) v = valid ? __ldg(mods + (size_t)(code - 0x10000) * nvox + vox) : 0.0f;
            else if (code > 0) {
              const int ax = (code - 1) >> 8, e = (code - 1) & 255;
              v = (ax == 0 ? rx : (ax == 1 ? ry : rz))[e];
            }
            hi[u] = __float_as_uint(v) & 0xFFFFE000u;
            lo[u] = __float_as_uint(v - __uint_as_float(hi[u]));
          }
          inr_tmem_st16(t_hi + c0, hi);
          inr_tmem_st16(t_lo + c0, lo);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) inr_mbar_arrive(&a_ready[j]);
      for (int l = 0; l < nl; ++l) {
        if (lane == 0) inr_mbar_wait(&d_ready[j], (uint32_t)((q * nl + l) & 1));     // one d_ready phase per (tile, layer)
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (l < nl - 1) {                                        // ReLU (model.py:47; the bias came with the MMA), next layer's A operand
          uint32_t dreg[32];
          const int cb = 32 * half;
          if (!(N.dbg & 2)) {
          inr_tmem_ld32_nowait(t_d + cb, dreg);
          inr_tmem_ld_wait();
          }
#pragma unroll
          for (int c0 = 0; c0 < 32 && !(N.dbg & 2); c0 += 16) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float v = fmaxf(__uint_as_float(dreg[c0 + u]), 0.0f);
              hi[u] = __float_as_uint(v) & 0xFFFFE000u;          // tf32 by truncation: what the tensor core would read anyway
              lo[u] = __float_as_uint(v - __uint_as_float(hi[u]));
            }
            inr_tmem_st16(t_hi + cb + c0, hi);
            inr_tmem_st16(t_lo + cb + c0, lo);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) inr_mbar_arrive(&a_ready[j]);
        } else if (half == 0) {                                  // logits (:49), argmax (:135-137)
          uint32_t dreg[16];
          inr_tmem_ld16_nowait(t_d, dreg);
          inr_tmem_ld_wait();
          float o[MRT_INR_MAX_CLASSES];
#pragma unroll
          for (int u = 0; u < MRT_INR_MAX_CLASSES; ++u) o[u] = __uint_as_float(dreg[u]);
          int best = 0; float bv = o[0];
#pragma unroll
          for (int u = 1; u < MRT_INR_MAX_CLASSES; ++u) if (u < ncls && o[u] > bv) { bv = o[u]; best = u; }   // first maximum, like argmax
          if (valid) {
            labels[vox] = best;
            if (logits != nullptr) {
#pragma unroll
              for (int u = 0; u < MRT_INR_MAX_CLASSES; ++u) if (u < ncls) logits[vox * ncls + u] = o[u];
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // D is free for the next tile's first MMA
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

static int inr_num_sms() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  return n;
}

// -> cudaErrorNotSupported when the network or the volume does not fit the tensor-core kernel (the caller falls back)
static cudaError_t launch_inr_tc(const InrNet& N, const float* mods, int X, int Y, int Z, const float* wts,
                                 int32_t* labels, float* logits, cudaStream_t st) {
  if (N.n_layers < 2) return cudaErrorNotSupported;
  if (N.dims[0] > INR_TC_HID || N.dims[N.n_layers] > MRT_INR_MAX_CLASSES || 1 + 2 * N.k > 255) return cudaErrorNotSupported;
  for (int l = 1; l < N.n_layers; ++l) if (N.dims[l] > INR_TC_HID) return cudaErrorNotSupported;
  const size_t nvox = (size_t)X * Y * Z;
  if (nvox >= (1ull << 32) - 128) return cudaErrorNotSupported;
  if ((long long)(X + Y + Z) * (1 + 2 * N.k) * 4 > 200 * 1024) return cudaErrorNotSupported;
  const InrTcLayout L = inr_tc_layout(N, X, Y, Z);
  if (L.total > 227 * 1024) return cudaErrorNotSupported;     // weights + axis tables must fit shared memory
  const long long ntiles = (long long)((nvox + 127) / 128);
  long long grid = inr_num_sms();
  if (grid > ntiles) grid = ntiles;
  // one CTA per SM: a CTA allocates all 512 TMEM columns, so a second resident CTA would wait for them
  // (ask for more than half of the shared memory to be sure)
  int smem = L.total;
  if (smem < 120 * 1024) smem = 120 * 1024;
  cudaError_t e = cudaFuncSetAttribute(mrt_inr_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  mrt_inr_tc_kernel<<<(int)grid, INR_TC_THREADS, smem, st>>>(N, L, mods, X, Y, Z, wts, labels, logits);
  return cudaGetLastError();
}

template <int HID, int BLOCK>
static cudaError_t launch_inr_b(const InrNet& N, const float* mods, int X, int Y, int Z, const float* wts,
                                int32_t* labels, float* logits, cudaStream_t st) {
  const size_t nvox = (size_t)X * Y * Z;
  const size_t wfl = ((size_t)inr_smem_floats<HID>(N.n_layers) + 3) & ~(size_t)3;
  const size_t smem = (wfl + (size_t)HID * BLOCK) * sizeof(float);
  if (smem > 227 * 1024) return cudaErrorNotSupported;
  cudaError_t e = cudaFuncSetAttribute(mrt_inr_kernel2<HID, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  size_t grid = (nvox + BLOCK - 1) / BLOCK;
  const size_t cap = 148 * (size_t)(1024 / BLOCK);          // the weights are staged once per CTA
  if (grid > cap) grid = cap;
  mrt_inr_kernel2<HID, BLOCK><<<(int)grid, BLOCK, smem, st>>>(N, mods, X, Y, Z, wts, labels, logits);
  return cudaGetLastError();
}
// the widest block whose activation columns still fit next to the weights (deep networks take a narrower one)
template <int HID>
static cudaError_t launch_inr(const InrNet& N, const float* mods, int X, int Y, int Z, const float* wts,
                              int32_t* labels, float* logits, cudaStream_t st) {
  cudaError_t e = launch_inr_b<HID, 512>(N, mods, X, Y, Z, wts, labels, logits, st);
  if (e == cudaErrorNotSupported) e = launch_inr_b<HID, 256>(N, mods, X, Y, Z, wts, labels, logits, st);
  if (e == cudaErrorNotSupported) e = launch_inr_b<HID, 128>(N, mods, X, Y, Z, wts, labels, logits, st);
  return e == cudaErrorNotSupported ? cudaErrorInvalidValue : e;
}

// layer_dims: n_layers + 1 widths; weights: per layer W[in][out] row-major followed by b[out]
// impl: 0 = tensor cores when the network fits (else FFMA), 1 = fp32 FFMA, 2 = tensor cores or fail
cudaError_t mrt_launch_inr(const float* mods, int M, int X, int Y, int Z, const float* weights, const int32_t* layer_dims,
                           int n_layers, int fourier_freqs, int32_t* labels, float* logits, int impl, cudaStream_t st) {
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
  { static const char* dbg = getenv("MRT_INR_DBG"); N.dbg = dbg ? atoi(dbg) : 0; }
  if (impl != 1) {
    cudaError_t e = launch_inr_tc(N, mods, X, Y, Z, weights, labels, logits, st);
    if (e != cudaErrorNotSupported || impl == 2) return e;
  }
  if (hid <= 32) return launch_inr<32>(N, mods, X, Y, Z, weights, labels, logits, st);
  if (hid <= 64) return launch_inr<64>(N, mods, X, Y, Z, weights, labels, logits, st);
  return cudaErrorInvalidValue;
}
