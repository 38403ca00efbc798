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
// dL/dtf goes to one of 64 privatised L2-resident copies with one 16-byte vector reduction per
// touched LUT entry (reduced by a tiny second kernel); dL/dvolume goes to L2 with native
// reductions (scalar for the folded / single-modality layout, red.v2/.v4 for interleaved).
// Cells that are flat (one value) and empty are leapt with their closed-form dL/dtf term.
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

#define MRT_DTF_COPIES 64     // privatised dL/dtf accumulators in L2 (CTA b uses copy b % 64)

// dL/dtf accumulation: one 16-byte vector reduction per touched LUT entry into this CTA's
// privatised copy (global fp32 atomics are native REDG.F32x4; shared-memory fp32 atomicAdd
// compiles to a CAS spin loop that serialises badly when a warp's lanes share a bin).
__device__ __forceinline__ void dtf_add(float4* __restrict__ dtfp, int j, float w, float dr, float dg, float db,
                                        float ds) {
  atomicAdd(dtfp + j, make_float4(w * dr, w * dg, w * db, w * ds));
}

template <int NCH, bool LABELS, bool SKIP, bool GENERIC>
__global__ void __launch_bounds__(64 * MRT_BWD_TPB)
mrt_bwd_kernel(const __grid_constant__ KParams P,
               const typename Vox<NCH>::T* __restrict__ vol,
               const float4* __restrict__ tf,
               const uint8_t* __restrict__ flat_levels,
               const float2* __restrict__ minmax,
               const int32_t* __restrict__ labels,
               const int32_t* __restrict__ preds,
               const float4* __restrict__ out_rgba,
               const float4* __restrict__ dL_dout,
               typename Vox<NCH>::T* __restrict__ dvol,
               float4* __restrict__ dtf_priv,
               float* __restrict__ dray) {
  typedef typename Vox<NCH>::T VT;
  extern __shared__ __align__(16) unsigned char s_raw[];   // [ntf] LUT | [16] labels
  const int ntf = P.tfMode ? P.tfN : 2;
  TfEntry* s_tf = reinterpret_cast<TfEntry*>(s_raw);
  float4* s_lab = reinterpret_cast<float4*>(s_tf + ntf);

  if (P.tfMode) {
    mrt_tf_stage(s_tf, tf, ntf);
  } else if (threadIdx.x == 0) {
    // the reference intensity TF (:135-138) is the 2-entry LUT [(0,0,0,0), (1,1,1,intensityAlpha)]
    s_tf[0].base = make_float4(0.f, 0.f, 0.f, 0.f); s_tf[0].delta = make_float4(1.f, 1.f, 1.f, P.ia);
    s_tf[1].base = make_float4(1.f, 1.f, 1.f, P.ia); s_tf[1].delta = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (LABELS) {
    if (threadIdx.x < 16) {
      const int l = threadIdx.x & 7;
      const float boost = threadIdx.x < 8 ? 1.0f : 1.5f;
      const float a = mrt_alpha(P, P.lut[l][3] * boost);
      s_lab[threadIdx.x] = make_float4(P.lut[l][0], P.lut[l][1], P.lut[l][2], (l > 0) ? a : 0.0f);
    }
  }
  __syncthreads();
  float4* const dtfp = dtf_priv ? dtf_priv + (size_t)(blockIdx.x & (MRT_DTF_COPIES - 1)) * ntf : nullptr;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = P.tile_begin + mrt_middle_out(blockIdx.x, gridDim.x) * MRT_BWD_TPB + (warp >> 1);
  if (tile >= P.tile_end) return;
  int px, py;
  mrt_pixel_of_tile_lane_(tile, mrt_logical_lane(warp & 1, lane), P.W, &px, &py);
  if (px >= P.W || py >= P.H) return;

  const size_t pix = (size_t)py * P.W + px;
  const float4 G = __ldg(dL_dout + pix);
  const Ray ray = mrt_setup_ray(P, P.eye, px, py);
  if (!(ray.n > 0 && (G.x != 0.0f || G.y != 0.0f || G.z != 0.0f || (P.alphaMode && G.w != 0.0f)))) return;

  const float4 Cout = __ldg(out_rgba + pix);
  const float S_tot = G.x * (Cout.x - P.bg[0]) + G.y * (Cout.y - P.bg[1]) + G.z * (Cout.z - P.bg[2]);
  // alphaMode 1: a = 1 - T_N  =>  dL/dT_N = -G.w ;  dsigma_i += -dt*T_N*dL/dT_N
  const float tn_term = P.alphaMode ? -(1.0f - Cout.w) * G.w : 0.0f;   // = T_N * dL/dT_N
  const IdxRay q = mrt_index_ray(P, ray);
  const float hix = (float)P.dims[0] - 1.001f, hiy = (float)P.dims[1] - 1.001f, hiz = (float)P.dims[2] - 1.001f;
  const float dt = P.dt, thr = P.thr;
  const float nm1 = (float)(ntf - 1);
  uint32_t s_tf_addr = (uint32_t)__cvta_generic_to_shared(s_tf);
    asm volatile("" : "+r"(s_tf_addr));        // opaque: keep the address in a register, do not re-derive it per sample
  const uint32_t sY = P.pitchY, sZ = P.pitchZ;
  float T = 1.0f, prefix = 0.0f;
  int k = 0;
  float gox = 0.0f, goy = 0.0f, goz = 0.0f, gdx = 0.0f, gdy = 0.0f, gdz = 0.0f;   // dL/do, dL/dd of this ray

  // one sample slot with its adjoint
  auto shade = [&](float t) {
    const float ppx = fmaf(t, q.dx, q.ox), ppy = fmaf(t, q.dy, q.oy), ppz = fmaf(t, q.dz, q.oz);
    const Cell c = mrt_cell(P, ppx, ppy, ppz, hix, hiy, hiz);
    const Corners<NCH, false> cor = mrt_fetch<NCH, false>(P, vol, c);
    const float raw = mrt_interp<NCH, false>(P, cor, c);
    const float val = mrt_window<GENERIC>(P, raw);
    if (P.tfMode || val > 0.0f) {
      int j0; float fr;
      const float4 rgba = mrt_tf_lookup(s_tf_addr, nm1, val, &j0, &fr);
      const float alpha = mrt_alpha(P, rgba.w);
      const float aT = alpha * T;
      const float gc = G.x * rgba.x + G.y * rgba.y + G.z * rgba.z;
      prefix = fmaf(aT, gc, prefix);
      const float suffix = S_tot - prefix;
      const float dsig = dt * ((1.0f - alpha) * T * gc - suffix - tn_term);
      const float dr = aT * G.x, dg = aT * G.y, db = aT * G.z;
      if (dtfp != nullptr) {
        dtf_add(dtfp, j0, 1.0f - fr, dr, dg, db, dsig);
        if (fr != 0.0f) dtf_add(dtfp, min(j0 + 1, ntf - 1), fr, dr, dg, db, dsig);
      }
      if (dvol != nullptr || dray != nullptr) {
        const float4 d4 = s_tf[j0].delta;
        float dval = nm1 * (dr * d4.x + dg * d4.y + db * d4.z + dsig * d4.w);
        if (GENERIC) {
          if (P.gamma != 1.0f) dval *= P.gamma * powf(__saturatef(raw), P.gamma - 1.0f);
        }
        // saturate: torch.clamp passes the gradient on the closed interval [0,1]
        const float dv = (raw >= 0.0f && raw <= 1.0f) ? dval : 0.0f;
        if (dray != nullptr && dv != 0.0f) {
          // docs/DifferentiableRendering.md section 9 (:172-188): x_i = o + t_i d with fixed t_i, so
          // dL/do += dL/dx_i and dL/dd += t_i dL/dx_i, with dL/dx_i = dL/ds * ds/dx (section 6); an
          // axis on which the position was clamped (:62) carries no gradient
          float sx, sy, sz;
          mrt_interp_grad<NCH, false>(P, cor, c, &sx, &sy, &sz);
          const float wx = (ppx >= 0.0f && ppx <= hix) ? dv * sx / P.vs[0] : 0.0f;
          const float wy = (ppy >= 0.0f && ppy <= hiy) ? dv * sy / P.vs[1] : 0.0f;
          const float wz = (ppz >= 0.0f && ppz <= hiz) ? dv * sz / P.vs[2] : 0.0f;
          gox += wx; goy += wy; goz += wz;
          gdx = fmaf(t, wx, gdx); gdy = fmaf(t, wy, gdy); gdz = fmaf(t, wz, gdz);
        }
        if (dvol != nullptr && dv != 0.0f) {
          const uint32_t b = (uint32_t)c.ix() + (uint32_t)c.iy() * sY + (uint32_t)c.iz() * sZ;
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
  };

  if (GENERIC && P.tMode == 1) {
    float t = ray.t0;
    while (t < ray.t1 && T > thr && (P.maxSteps == 0 || k < P.maxSteps)) { shade(t); t += dt; ++k; }
  } else if (SKIP) {
    // Flat-empty cells (every voxel of the cell holds the same value c AND sigma == 0 over the TF
    // bins c maps to AND no overlay label): all ns slots inside have the same (bin, frac, colour),
    // alpha == 0, so T, prefix and hence dL/dsigma are identical for every slot; the volume
    // gradient is exactly 0 (the LUT slope of sigma is 0 there and dL/dc = alpha*T*G = 0).
    // Their whole contribution is ns * dL/dsigma onto two LUT entries: one reduction per cell.
    const float ivx = 1.0f / q.dx, ivy = 1.0f / q.dy, ivz = 1.0f / q.dz;
    const float inv_dt = 1.0f / dt;
    const int n = ray.n;
    int kact = 0;
    for (;;) {
      while (k >= kact && k < n && T > thr) {
        const float t = fmaf((float)k, dt, ray.t0);
        const float ppx = fmaf(t, q.dx, q.ox), ppy = fmaf(t, q.dy, q.oy), ppz = fmaf(t, q.dz, q.oz);
        const int ix = (int)fminf(fmaxf(ppx, 0.0f), hix);
        const int iy = (int)fminf(fmaxf(ppy, 0.0f), hiy);
        const int iz = (int)fminf(fmaxf(ppz, 0.0f), hiz);
        const int bid = ((iz >> MRT_BRICK_SHIFT) * P.nby + (iy >> MRT_BRICK_SHIFT)) * P.nbx + (ix >> MRT_BRICK_SHIFT);
        const int lvl = __ldg(flat_levels + bid);
        const int sh = lvl ? lvl + (MRT_BRICK_SHIFT - 1) : MRT_BRICK_SHIFT;
        const int kend = min(n, k + mrt_cell_slots(q, ivx, ivy, ivz, ix >> sh, iy >> sh, iz >> sh, sh, t, inv_dt));
        if (lvl) {
          if (dtfp != nullptr) {
            const float cval = __ldg(&minmax[(size_t)bid * NCH].x);
            const float val = mrt_window<GENERIC>(P, cval * P.wq[0] + P.wbias);
            if (P.tfMode || val > 0.0f) {
              int j0; float fr;
              const float4 rgba = mrt_tf_lookup(s_tf_addr, nm1, val, &j0, &fr);
              const float gc = G.x * rgba.x + G.y * rgba.y + G.z * rgba.z;
              const float dsig = (float)(kend - k) * dt * (T * gc - (S_tot - prefix) - tn_term);
              dtf_add(dtfp, j0, 1.0f - fr, 0.f, 0.f, 0.f, dsig);
              if (fr != 0.0f) dtf_add(dtfp, min(j0 + 1, ntf - 1), fr, 0.f, 0.f, 0.f, dsig);
            }
          }
          k = kend;
        } else {
          kact = kend;
        }
      }
      if (!(k < n && T > thr)) break;
      shade(fmaf((float)k, dt, ray.t0));
      ++k;
    }
  } else {
    while (k < ray.n && T > thr) { shade(fmaf((float)k, dt, ray.t0)); ++k; }
  }
  if (dray != nullptr) {                      // [H][W][6] = (dL/do, dL/dd); rays that returned early keep the caller's zeros
    float* r = dray + pix * 6;
    r[0] = gox; r[1] = goy; r[2] = goz; r[3] = gdx; r[4] = gdy; r[5] = gdz;
  }
}

// dtf[i] += sum over the privatised copies
__global__ void mrt_dtf_reduce_kernel(const float* __restrict__ priv, int ncopies, int nfloats, float* __restrict__ dtf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nfloats) return;
  float s = 0.0f;
  for (int c = 0; c < ncopies; ++c) s += priv[(size_t)c * nfloats + i];
  dtf[i] += s;
}

template <int NCH, bool LABELS, bool SKIP, bool GENERIC>
static cudaError_t launch_bwd(const KParams& P, const void* vol, const float* tf, const uint8_t* flat_levels,
                              const float* minmax, const int32_t* labels, const int32_t* preds,
                              const float* out_rgba, const float* dL_dout, void* dvol, float* dtf, void* scratch,
                              float* dray, cudaStream_t st) {
  typedef typename Vox<NCH>::T VT;
  const int ntiles = P.tile_end - P.tile_begin;
  if (ntiles <= 0) return cudaSuccess;
  const int grid = (ntiles + MRT_BWD_TPB - 1) / MRT_BWD_TPB;
  const int ntf = P.tfMode ? P.tfN : 2;
  const size_t smem = (size_t)ntf * sizeof(TfEntry) + 16 * sizeof(float4);
  if (dtf) {
    cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)MRT_DTF_COPIES * ntf * sizeof(float4), st);
    if (e != cudaSuccess) return e;
  }
  mrt_bwd_kernel<NCH, LABELS, SKIP, GENERIC><<<grid, 64 * MRT_BWD_TPB, smem, st>>>(
      P, (const VT*)vol, (const float4*)tf, flat_levels, (const float2*)minmax, labels, preds,
      (const float4*)out_rgba, (const float4*)dL_dout, (VT*)dvol, dtf ? (float4*)scratch : nullptr, dray);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (dtf) {
    mrt_dtf_reduce_kernel<<<(ntf * 4 + 255) / 256, 256, 0, st>>>((const float*)scratch, MRT_DTF_COPIES, ntf * 4, dtf);
    e = cudaGetLastError();
  }
  return e;
}

size_t mrt_bwd_scratch_bytes(int ntf) { return (size_t)MRT_DTF_COPIES * ntf * sizeof(float4); }

template <int NCH, bool SKIP>
static cudaError_t dispatch_bwd(const KParams& P, bool lab, bool gen, const void* vol, const float* tf,
                                const uint8_t* fl, const float* mm, const int32_t* labels, const int32_t* preds,
                                const float* o, const float* g, void* dvol, float* dtf, void* scr, float* dray,
                                cudaStream_t st) {
  if (lab) return gen ? launch_bwd<NCH, true, SKIP, true>(P, vol, tf, fl, mm, labels, preds, o, g, dvol, dtf, scr, dray, st)
                      : launch_bwd<NCH, true, SKIP, false>(P, vol, tf, fl, mm, labels, preds, o, g, dvol, dtf, scr, dray, st);
  return gen ? launch_bwd<NCH, false, SKIP, true>(P, vol, tf, fl, mm, labels, preds, o, g, dvol, dtf, scr, dray, st)
             : launch_bwd<NCH, false, SKIP, false>(P, vol, tf, fl, mm, labels, preds, o, g, dvol, dtf, scr, dray, st);
}

cudaError_t mrt_launch_backward(const KParams& P, int packed_ch, const void* vol, const float* tf,
                                const uint8_t* flat_levels, const float* minmax,
                                const int32_t* labels, const int32_t* preds, const float* out_rgba,
                                const float* dL_dout, void* dvol, float* dtf, void* scratch, float* dray,
                                cudaStream_t st) {
  const bool lab = (P.showSeg || P.showPred);
  const bool gen = (P.tMode != 0) || (P.gamma != 1.0f);
  const bool skip = P.skip && flat_levels != nullptr && minmax != nullptr && P.tMode == 0 && packed_ch == 1;
  switch (packed_ch) {
    case 1: return skip ? dispatch_bwd<1, true>(P, lab, gen, vol, tf, flat_levels, minmax, labels, preds, out_rgba, dL_dout, dvol, dtf, scratch, dray, st)
                        : dispatch_bwd<1, false>(P, lab, gen, vol, tf, flat_levels, minmax, labels, preds, out_rgba, dL_dout, dvol, dtf, scratch, dray, st);
    case 2: return dispatch_bwd<2, false>(P, lab, gen, vol, tf, flat_levels, minmax, labels, preds, out_rgba, dL_dout, dvol, dtf, scratch, dray, st);
    case 4: return dispatch_bwd<4, false>(P, lab, gen, vol, tf, flat_levels, minmax, labels, preds, out_rgba, dL_dout, dvol, dtf, scratch, dray, st);
  }
  return cudaErrorInvalidValue;
}
