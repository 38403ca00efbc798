// backward.cu — adjoint of the forward march w.r.t. the volume and the transfer function.
//
// Spec: docs/DifferentiableRendering.md §5-§6 (:88-127).  No reference code exists; the
// ground truth is the oracle's autograd.  Per ray the forward is re-marched in the SAME
// order with the SAME arithmetic (so every early-termination decision repeats), and the
// doc's O(N) T-adjoint recurrence is evaluated front-to-back through the identity
//     sum_{j>i} G.c_j alpha_j T_{j-1}  =  G.(C_out - bg)  -  sum_{j<=i} G.c_j alpha_j T_{j-1}
// so nothing but the forward image has to be stored:
//     dL/dsigma_i = dt * ( (1-alpha_i) T_{i-1} (G.c_i)  -  suffix_i  -  T_N * dL/dT_N )
//     dL/dc_i     = G * alpha_i T_{i-1}
// then through the LUT lerp (-> dL/dtf[j0], dL/dtf[j1], dL/dval), window/level, the
// modality blend and the trilinear weights (-> 8 scatter-adds per sample).
// dL/dtf is reduced in shared memory per CTA and flushed once; dL/dvolume goes to L2 with
// vector reductions (one red.v4.f32 per corner for the 4-modality layout).
#include "march.cuh"
#include "kernels.h"

#ifndef MRT_BWD_TPB
#define MRT_BWD_TPB 2
#endif

__device__ __forceinline__ void vox_atomic_add(float* p, float w, const KParams& P) {
  atomicAdd(p, w * P.wq[0]);
}
__device__ __forceinline__ void vox_atomic_add(float2* p, float w, const KParams& P) {
  atomicAdd(p, make_float2(w * P.wq[0], w * P.wq[1]));
}
__device__ __forceinline__ void vox_atomic_add(float4* p, float w, const KParams& P) {
  atomicAdd(p, make_float4(w * P.wq[0], w * P.wq[1], w * P.wq[2], w * P.wq[3]));
}

template <int NCH, bool LABELS, bool GENERIC>
__global__ void __launch_bounds__(64 * MRT_BWD_TPB)
mrt_bwd_kernel(const __grid_constant__ KParams P,
               const typename Vox<NCH>::T* __restrict__ vol,
               const float4* __restrict__ tf,
               const int32_t* __restrict__ labels,
               const int32_t* __restrict__ preds,
               const float4* __restrict__ out_rgba,
               const float4* __restrict__ dL_dout,
               typename Vox<NCH>::T* __restrict__ dvol,
               float* __restrict__ dtf) {
  typedef typename Vox<NCH>::T VT;
  extern __shared__ __align__(16) unsigned char s_raw[];   // [ntf] LUT | [16] labels | [ntf*4] dtf accum
  const int ntf = P.tfMode ? P.tfN : 2;
  TfEntry* s_tf = reinterpret_cast<TfEntry*>(s_raw);
  float4* s_lab = reinterpret_cast<float4*>(s_tf + ntf);
  float* s_dtf = reinterpret_cast<float*>(s_lab + 16);

  if (P.tfMode) {
    mrt_tf_stage(s_tf, tf, ntf);
  } else if (threadIdx.x == 0) {
    // the reference intensity TF (:135-138) is the 2-entry LUT [(0,0,0,0), (1,1,1,intensityAlpha)]
    s_tf[0].base = make_float4(0.f, 0.f, 0.f, 0.f); s_tf[0].delta = make_float4(1.f, 1.f, 1.f, P.ia);
    s_tf[1].base = make_float4(1.f, 1.f, 1.f, P.ia); s_tf[1].delta = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = threadIdx.x; i < ntf * 4; i += blockDim.x) s_dtf[i] = 0.0f;
  if (LABELS) {
    if (threadIdx.x < 16) {
      const int l = threadIdx.x & 7;
      const float boost = threadIdx.x < 8 ? 1.0f : 1.5f;
      const float a = mrt_alpha(P, P.lut[l][3] * boost);
      s_lab[threadIdx.x] = make_float4(P.lut[l][0], P.lut[l][1], P.lut[l][2], (l > 0) ? a : 0.0f);
    }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = P.tile_begin + mrt_middle_out(blockIdx.x, gridDim.x) * MRT_BWD_TPB + (warp >> 1);
  int px = -1, py = -1;
  bool live = tile < P.tile_end;
  if (live) {
    mrt_pixel_of_tile_lane_(tile, mrt_logical_lane(warp & 1, lane), P.W, &px, &py);
    live = (px < P.W && py < P.H);
  }

  if (live) {
    const size_t pix = (size_t)py * P.W + px;
    const float4 G = __ldg(dL_dout + pix);
    const Ray ray = mrt_setup_ray(P, px, py);
    if (ray.n > 0 && (G.x != 0.0f || G.y != 0.0f || G.z != 0.0f || (P.alphaMode && G.w != 0.0f))) {
      const float4 Cout = __ldg(out_rgba + pix);
      const float S_tot = G.x * (Cout.x - P.bg[0]) + G.y * (Cout.y - P.bg[1]) + G.z * (Cout.z - P.bg[2]);
      // alphaMode 1: a = 1 - T_N  =>  dL/dT_N = -G.w ;  dsigma_i += -dt*T_N*dL/dT_N
      const float tn_term = P.alphaMode ? -(1.0f - Cout.w) * G.w : 0.0f;   // = T_N * dL/dT_N
      const IdxRay q = mrt_index_ray(P, ray);
      const float hix = (float)P.dims[0] - 1.001f, hiy = (float)P.dims[1] - 1.001f, hiz = (float)P.dims[2] - 1.001f;
      const float dt = P.dt, thr = P.thr;
      const float nm1 = (float)(ntf - 1);
      const uint32_t sY = P.pitchY, sZ = P.pitchZ;
      float T = 1.0f, prefix = 0.0f;
      int k = 0;
      float t_run = ray.t0;
      while (true) {
        float t;
        if (GENERIC && P.tMode == 1) {
          if (!(t_run < ray.t1 && T > thr && (P.maxSteps == 0 || k < P.maxSteps))) break;
          t = t_run;
        } else {
          if (!(k < ray.n && T > thr)) break;
          t = fmaf((float)k, dt, ray.t0);
        }
        const float ppx = fmaf(t, q.dx, q.ox), ppy = fmaf(t, q.dy, q.oy), ppz = fmaf(t, q.dz, q.oz);
        const Cell c = mrt_cell(P, ppx, ppy, ppz, hix, hiy, hiz);
        const float raw = mrt_sample_raw<NCH>(P, vol, c);
        const float val = mrt_window<GENERIC>(P, raw);
        if (P.tfMode || val > 0.0f) {
          int j0; float fr;
          const float4 rgba = mrt_tf_lookup(s_tf, nm1, val, &j0, &fr);
          const float alpha = mrt_alpha(P, rgba.w);
          const float aT = alpha * T;
          const float gc = G.x * rgba.x + G.y * rgba.y + G.z * rgba.z;
          prefix = fmaf(aT, gc, prefix);
          const float suffix = S_tot - prefix;
          const float dsig = dt * ((1.0f - alpha) * T * gc - suffix - tn_term);
          const float dr = aT * G.x, dg = aT * G.y, db = aT * G.z;
          if (dtf != nullptr) {
            const float w0 = 1.0f - fr;
            atomicAdd(s_dtf + j0 * 4 + 0, w0 * dr); atomicAdd(s_dtf + j0 * 4 + 1, w0 * dg);
            atomicAdd(s_dtf + j0 * 4 + 2, w0 * db); atomicAdd(s_dtf + j0 * 4 + 3, w0 * dsig);
            if (fr != 0.0f) {
              const int j1 = min(j0 + 1, ntf - 1);
              atomicAdd(s_dtf + j1 * 4 + 0, fr * dr); atomicAdd(s_dtf + j1 * 4 + 1, fr * dg);
              atomicAdd(s_dtf + j1 * 4 + 2, fr * db); atomicAdd(s_dtf + j1 * 4 + 3, fr * dsig);
            }
          }
          if (dvol != nullptr) {
            const float4 d4 = s_tf[j0].delta;
            float dval = nm1 * (dr * d4.x + dg * d4.y + db * d4.z + dsig * d4.w);
            if (GENERIC) {
              if (P.gamma != 1.0f) dval *= P.gamma * powf(__saturatef(raw), P.gamma - 1.0f);
            }
            // saturate: torch.clamp passes the gradient on the closed interval [0,1]
            const float dv = (raw >= 0.0f && raw <= 1.0f) ? dval : 0.0f;
            if (dv != 0.0f) {
              const uint32_t b = (uint32_t)c.ix + (uint32_t)c.iy * sY + (uint32_t)c.iz * sZ;
              VT* p0 = dvol + b; VT* p1 = p0 + sY; VT* p2 = p0 + sZ; VT* p3 = p2 + sY;
              const float gx0 = 1.0f - c.fx, gy0 = 1.0f - c.fy, gz0 = 1.0f - c.fz;
              const float w00 = dv * gy0 * gz0, w10 = dv * c.fy * gz0, w01 = dv * gy0 * c.fz, w11 = dv * c.fy * c.fz;
              vox_atomic_add(p0, w00 * gx0, P); vox_atomic_add(p0 + 1, w00 * c.fx, P);
              vox_atomic_add(p1, w10 * gx0, P); vox_atomic_add(p1 + 1, w10 * c.fx, P);
              vox_atomic_add(p2, w01 * gx0, P); vox_atomic_add(p2 + 1, w01 * c.fx, P);
              vox_atomic_add(p3, w11 * gx0, P); vox_atomic_add(p3 + 1, w11 * c.fx, P);
            }
          }
          T *= (1.0f - alpha);
        }
        if (LABELS) {          // overlays carry no gradient but attenuate what lies behind
          if (P.showSeg) {
            const int l = mrt_sample_label(P, labels, ppx, ppy, ppz);
            if (l > 0 && l < 8) {
              const float4 col = s_lab[l];
              prefix = fmaf(col.w * T, G.x * col.x + G.y * col.y + G.z * col.z, prefix);
              T *= (1.0f - col.w);
            }
          }
          if (P.showPred) {
            const int l = mrt_sample_label(P, preds, ppx, ppy, ppz);
            if (l > 0 && l < 8) {
              const float4 col = s_lab[8 + l];
              prefix = fmaf(col.w * T, G.x * col.x + G.y * col.y + G.z * col.z, prefix);
              T *= (1.0f - col.w);
            }
          }
        }
        ++k;
        if (GENERIC) t_run += dt;
      }
    }
  }
  __syncthreads();
  if (dtf != nullptr) {
    for (int i = threadIdx.x; i < ntf * 4; i += blockDim.x) {
      const float x = s_dtf[i];
      if (x != 0.0f) atomicAdd(dtf + i, x);
    }
  }
}

template <int NCH, bool LABELS, bool GENERIC>
static cudaError_t launch_bwd(const KParams& P, const void* vol, const float* tf, const int32_t* labels,
                              const int32_t* preds, const float* out_rgba, const float* dL_dout, void* dvol,
                              float* dtf, cudaStream_t st) {
  typedef typename Vox<NCH>::T VT;
  const int ntiles = P.tile_end - P.tile_begin;
  if (ntiles <= 0) return cudaSuccess;
  const int grid = (ntiles + MRT_BWD_TPB - 1) / MRT_BWD_TPB;
  const int ntf = P.tfMode ? P.tfN : 2;
  const size_t smem = (size_t)ntf * sizeof(TfEntry) + 16 * sizeof(float4) + (size_t)ntf * 4 * sizeof(float);
  mrt_bwd_kernel<NCH, LABELS, GENERIC><<<grid, 64 * MRT_BWD_TPB, smem, st>>>(
      P, (const VT*)vol, (const float4*)tf, labels, preds, (const float4*)out_rgba, (const float4*)dL_dout,
      (VT*)dvol, dtf);
  return cudaGetLastError();
}

template <int NCH>
static cudaError_t dispatch_bwd(const KParams& P, bool lab, bool gen, const void* vol, const float* tf,
                                const int32_t* labels, const int32_t* preds, const float* o, const float* g,
                                void* dvol, float* dtf, cudaStream_t st) {
  if (lab) return gen ? launch_bwd<NCH, true, true>(P, vol, tf, labels, preds, o, g, dvol, dtf, st)
                      : launch_bwd<NCH, true, false>(P, vol, tf, labels, preds, o, g, dvol, dtf, st);
  return gen ? launch_bwd<NCH, false, true>(P, vol, tf, labels, preds, o, g, dvol, dtf, st)
             : launch_bwd<NCH, false, false>(P, vol, tf, labels, preds, o, g, dvol, dtf, st);
}

cudaError_t mrt_launch_backward(const KParams& P, int packed_ch, const void* vol, const float* tf,
                                const int32_t* labels, const int32_t* preds, const float* out_rgba,
                                const float* dL_dout, void* dvol, float* dtf, cudaStream_t st) {
  const bool lab = (P.showSeg || P.showPred);
  const bool gen = (P.tMode != 0) || (P.gamma != 1.0f);
  switch (packed_ch) {
    case 1: return dispatch_bwd<1>(P, lab, gen, vol, tf, labels, preds, out_rgba, dL_dout, dvol, dtf, st);
    case 2: return dispatch_bwd<2>(P, lab, gen, vol, tf, labels, preds, out_rgba, dL_dout, dvol, dtf, st);
    case 4: return dispatch_bwd<4>(P, lab, gen, vol, tf, labels, preds, out_rgba, dL_dout, dvol, dtf, st);
  }
  return cudaErrorInvalidValue;
}
