// adaptive.cu — differentiable adaptive sampling: coarse pass -> piecewise-linear CDF of the
// extinction -> fine samples at fixed quantiles (inverse CDF), forward and backward.
//
// Spec: docs/DifferentiableRendering.md section 7 (:131-148) — maths only, no reference code; the
// ground truth is oracle/oracle_adaptive.py and its autograd.  Per ray with clip interval [t0,t1):
//   coarse   : K bins of width h = (t1-t0)/K sampled at their centres, w_k = sigma_k + eps_w (:133)
//   CDF      : W_k = sum_{l<k} w_l, F = W/W_K piecewise linear (:134-136)
//   quantile : Q(u) = t0 + h (k + (u W_K - W_k)/w_k),  W_k <= u W_K < W_{k+1} (:138-140)
//   fine     : sample j at Q((j+1/2)/J) stands for [Q(j/J), Q((j+1)/J)), alpha_j = 1-exp(-sigma_j Delta_j)
// then the front-to-back compositing with early termination of brats_rt.slang:117,135-139.
// Backward: besides the usual adjoints of every fine sample (-> dL/dtf, dL/dvolume), the sample
// TIME and interval length depend on the importance weights: with g = dL/dQ(u) for a quantile in
// bin k at fraction f,
//   dQ/dw_l = h/w_k (u - [l<k]) - [l=k] h f/w_k          (the doc's implicit formula, :142-146, spelled out)
// so dL/dw_l = A - sum_{k>l} S_k - D_l with three per-ray accumulators, and dL/dw_l flows through
// the coarse samples' sigma into the LUT and the volume.  dL/dQ itself is the position gradient of
// section 6 (ds/dx . d) for the sample time and +-sigma_j (dL/dalpha)(1-alpha) for the interval ends.
// One thread per ray; the coarse weights live in local memory (K <= 64).
#include "march.cuh"
#include "kernels.h"

#define MRT_ADP_MAXK 64
#define MRT_ADP_COPIES 64

struct AdpArgs { int K, J; float eps_w; };

__device__ __forceinline__ void adp_vox_add(float* p, float w, const KParams& P) { atomicAdd(p, w * P.wq[0]); }
__device__ __forceinline__ void adp_vox_add(float2* p, float w, const KParams& P) {
  atomicAdd(p, make_float2(w * P.wq[0], w * P.wq[1]));
}
__device__ __forceinline__ void adp_vox_add(float4* p, float w, const KParams& P) {
  atomicAdd(p, make_float4(w * P.wq[0], w * P.wq[1], w * P.wq[2], w * P.wq[3]));
}

template <int NCH, bool BWD>
__global__ void __launch_bounds__(128)
mrt_adaptive_kernel(const __grid_constant__ KParams P, const __grid_constant__ AdpArgs A,
                    const typename Vox<NCH>::T* __restrict__ vol, const float4* __restrict__ tf,
                    float4* __restrict__ out_rgba, const float4* __restrict__ dL_dout,
                    typename Vox<NCH>::T* __restrict__ dvol, float4* __restrict__ dtf_priv) {
  typedef typename Vox<NCH>::T VT;
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int ntf = P.tfMode ? P.tfN : 2;
  TfEntry* s_tf = reinterpret_cast<TfEntry*>(s_raw);
  if (P.tfMode) {
    mrt_tf_stage(s_tf, tf, ntf);
  } else if (threadIdx.x == 0) {
    // the reference intensity TF (:135-138) is the 2-entry LUT [(0,0,0,0), (1,1,1,intensityAlpha)]
    s_tf[0].base = make_float4(0.f, 0.f, 0.f, 0.f); s_tf[0].delta = make_float4(1.f, 1.f, 1.f, P.ia);
    s_tf[1].base = make_float4(1.f, 1.f, 1.f, P.ia); s_tf[1].delta = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = P.tile_begin + blockIdx.x * 2 + (warp >> 1);
  if (tile >= P.tile_end) return;
  int px, py;
  mrt_pixel_of_tile_lane_(tile, mrt_logical_lane(warp & 1, lane), P.W, &px, &py);
  if (px >= P.W || py >= P.H) return;
  const size_t pix = (size_t)py * P.W + px;
  const Ray ray = mrt_setup_ray(P, P.eye, px, py);
  const bool hit = ray.n > 0;
  float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
  if (BWD) {
    G = __ldg(dL_dout + pix);
    if (!hit || !(G.x != 0.0f || G.y != 0.0f || G.z != 0.0f || (P.alphaMode && G.w != 0.0f))) return;
  } else if (!hit) {
    out_rgba[pix] = make_float4(P.bg[0], P.bg[1], P.bg[2], P.alphaMode ? 0.0f : 1.0f);
    return;
  }
  const IdxRay q = mrt_index_ray(P, ray);
  const float hix = (float)P.dims[0] - 1.001f, hiy = (float)P.dims[1] - 1.001f, hiz = (float)P.dims[2] - 1.001f;
  const float nm1 = (float)(ntf - 1), thr = P.thr;
  const float log2e = 1.4426950408889634f;
  uint32_t s_tf_addr = (uint32_t)__cvta_generic_to_shared(s_tf);
  const int K = A.K, J = A.J;
  const float h = (ray.t1 - ray.t0) / (float)K;

  // the field at ray parameter t: window/level value, LUT entry, and (optionally) what the adjoint needs
  auto field = [&](float t, Cell* c, Corners<NCH, false>* cor, float* raw, int* j0, float* fr) -> float4 {
    const float ppx = fmaf(t, q.dx, q.ox), ppy = fmaf(t, q.dy, q.oy), ppz = fmaf(t, q.dz, q.oz);
    *c = mrt_cell(P, ppx, ppy, ppz, hix, hiy, hiz);
    *cor = mrt_fetch<NCH, false>(P, vol, *c);
    *raw = mrt_interp<NCH, false>(P, *cor, *c);
    const float val = __saturatef(*raw);
    float4 rgba = mrt_tf_lookup(s_tf_addr, nm1, val, j0, fr);
    if (!P.tfMode && !(val > 0.0f)) rgba.w = 0.0f;          // :135 (val > 0); the 2-entry LUT already gives sigma = 0 there
    return rgba;
  };

  // ---- coarse stage
  float w[MRT_ADP_MAXK], Wc[MRT_ADP_MAXK + 1];
  Wc[0] = 0.0f;
  for (int k = 0; k < K; ++k) {
    Cell c; Corners<NCH, false> cor; float raw, fr; int j0;
    const float4 rgba = field(ray.t0 + ((float)k + 0.5f) * h, &c, &cor, &raw, &j0, &fr);
    w[k] = rgba.w + A.eps_w;
    Wc[k + 1] = Wc[k] + w[k];
  }
  const float Wt = Wc[K];
  int kq = 0;                                             // the quantiles are visited in increasing u: one monotone pointer
  auto quantile = [&](float u, float* frac) -> float {
    const float target = u * Wt;
    while (kq < K - 1 && Wc[kq + 1] <= target) ++kq;
    *frac = (target - Wc[kq]) / w[kq];
    return ray.t0 + h * ((float)kq + *frac);
  };

  float Cr = P.bg[0], Cg = P.bg[1], Cb = P.bg[2], T = 1.0f;
  // backward state
  float S_tot = 0.0f, tn_term = 0.0f, prefix = 0.0f, Aall = 0.0f;
  float Sk[BWD ? MRT_ADP_MAXK : 1], Dk[BWD ? MRT_ADP_MAXK : 1];
  float4* gpriv = nullptr;
  if (BWD) {
    const float4 Cout = __ldg(out_rgba + pix);
    S_tot = G.x * (Cout.x - P.bg[0]) + G.y * (Cout.y - P.bg[1]) + G.z * (Cout.z - P.bg[2]);
    tn_term = P.alphaMode ? -(1.0f - Cout.w) * G.w : 0.0f;
    for (int k = 0; k < K; ++k) { Sk[k] = 0.0f; Dk[k] = 0.0f; }
    if (dtf_priv) gpriv = dtf_priv + (size_t)(blockIdx.x & (MRT_ADP_COPIES - 1)) * ntf * 2;
  }
  auto dq = [&](float g, float u, int k, float frac) {      // dL/dQ(u) = g  ->  the three accumulators
    const float gh = g * h / w[k];
    Aall = fmaf(gh, u, Aall);
    Sk[k] += gh;
    Dk[k] = fmaf(gh, frac, Dk[k]);
  };

  float flo; int klo = 0;
  float qlo = ray.t0;                                      // Q(0) = t0: independent of the weights
  flo = 0.0f;
  for (int j = 0; j < J && T > thr; ++j) {                 // :117
    float fm, fh;
    const float um = ((float)j + 0.5f) / (float)J;
    const float tm = quantile(um, &fm);
    const int km = kq;
    const float uh = (float)(j + 1) / (float)J;
    float qhi = ray.t1;                                    // Q(1) = t1
    int kh = K - 1;
    fh = 0.0f;
    if (j + 1 < J) { qhi = quantile(uh, &fh); kh = kq; }
    const float delta = qhi - qlo;
    Cell c; Corners<NCH, false> cor; float raw, fr; int j0;
    const float4 rgba = field(tm, &c, &cor, &raw, &j0, &fr);
    const float alpha = 1.0f - mrt_ex2(rgba.w * delta * (-log2e));
    const float aT = alpha * T;
    if (!BWD) {
      Cr = fmaf(aT, rgba.x, Cr); Cg = fmaf(aT, rgba.y, Cg); Cb = fmaf(aT, rgba.z, Cb);
    } else if (P.tfMode || __saturatef(raw) > 0.0f) {
      const float gc = G.x * rgba.x + G.y * rgba.y + G.z * rgba.z;
      prefix = fmaf(aT, gc, prefix);
      const float base = (1.0f - alpha) * T * gc - (S_tot - prefix) - tn_term;   // dL/dalpha / ... see backward.cu
      const float dsig = delta * base, ddelta = rgba.w * base;
      const float dr = aT * G.x, dg = aT * G.y, db = aT * G.z;
      if (gpriv) {
        const float f0 = 1.0f - fr;
        atomicAdd(gpriv + 2 * j0, make_float4(f0 * dr, f0 * dg, f0 * db, f0 * dsig));
        if (fr != 0.0f) atomicAdd(gpriv + 2 * j0 + 1, make_float4(fr * dr, fr * dg, fr * db, fr * dsig));
      }
      const float4 d4 = s_tf[j0].delta;
      const float dval = nm1 * (dr * d4.x + dg * d4.y + db * d4.z + dsig * d4.w);
      const float dv = (raw >= 0.0f && raw <= 1.0f) ? dval : 0.0f;     // saturate: torch.clamp's closed interval
      if (dv != 0.0f) {
        if (dvol) {
          const uint32_t b = (uint32_t)c.ix() + (uint32_t)c.iy() * P.pitchY + (uint32_t)c.iz() * P.pitchZ;
          VT* p0 = dvol + b; VT* p1 = p0 + P.pitchY; VT* p2 = p0 + P.pitchZ; VT* p3 = p2 + P.pitchY;
          const float gx0 = 1.0f - c.fx, gy0 = 1.0f - c.fy, gz0 = 1.0f - c.fz;
          const float w00 = dv * gy0 * gz0, w10 = dv * c.fy * gz0, w01 = dv * gy0 * c.fz, w11 = dv * c.fy * c.fz;
          adp_vox_add(p0, w00 * gx0, P); adp_vox_add(p0 + 1, w00 * c.fx, P);
          adp_vox_add(p1, w10 * gx0, P); adp_vox_add(p1 + 1, w10 * c.fx, P);
          adp_vox_add(p2, w01 * gx0, P); adp_vox_add(p2 + 1, w01 * c.fx, P);
          adp_vox_add(p3, w11 * gx0, P); adp_vox_add(p3 + 1, w11 * c.fx, P);
        }
        // the sample moves with its quantile: dL/dt = dL/ds * (ds/dx . dx/dt), x in index space, clamped axes carry nothing (:62)
        float sx, sy, sz;
        mrt_interp_grad<NCH, false>(P, cor, c, &sx, &sy, &sz);
        const float ppx = fmaf(tm, q.dx, q.ox), ppy = fmaf(tm, q.dy, q.oy), ppz = fmaf(tm, q.dz, q.oz);
        float gt = 0.0f;
        if (ppx >= 0.0f && ppx <= hix) gt = fmaf(sx, q.dx, gt);
        if (ppy >= 0.0f && ppy <= hiy) gt = fmaf(sy, q.dy, gt);
        if (ppz >= 0.0f && ppz <= hiz) gt = fmaf(sz, q.dz, gt);
        dq(dv * gt, um, km, fm);
      }
      if (j > 0) dq(-ddelta, (float)j / (float)J, klo, flo);            // Delta_j = Q(u_{j+1}) - Q(u_j); the ends Q(0), Q(1) are fixed
      if (j + 1 < J) dq(ddelta, uh, kh, fh);
    }
    T *= (1.0f - alpha);
    qlo = qhi; klo = kh; flo = fh;
  }
  if (!BWD) {
    out_rgba[pix] = make_float4(Cr, Cg, Cb, P.alphaMode ? 1.0f - T : 1.0f);
    return;
  }
  // ---- dL/dw_l = A - sum_{k>l} S_k - D_l, back through the coarse samples' sigma
  float suffix = 0.0f;
  for (int l = K - 1; l >= 0; --l) {
    const float gs = Aall - suffix - Dk[l];
    suffix += Sk[l];
    if (gs == 0.0f) continue;
    Cell c; Corners<NCH, false> cor; float raw, fr; int j0;
    field(ray.t0 + ((float)l + 0.5f) * h, &c, &cor, &raw, &j0, &fr);
    // no (val > 0) gate here: :135 gates the COMPOSITING of a sample; the importance weight is
    // sigma = ia * clamp(raw), whose derivative at raw == 0 (air) follows torch.clamp's closed interval
    if (gpriv) {
      atomicAdd(&gpriv[2 * j0].w, (1.0f - fr) * gs);
      if (fr != 0.0f) atomicAdd(&gpriv[2 * j0 + 1].w, fr * gs);
    }
    const float dval = nm1 * gs * s_tf[j0].delta.w;
    const float dv = (raw >= 0.0f && raw <= 1.0f) ? dval : 0.0f;
    if (dv != 0.0f && dvol) {
      const uint32_t b = (uint32_t)c.ix() + (uint32_t)c.iy() * P.pitchY + (uint32_t)c.iz() * P.pitchZ;
      VT* p0 = dvol + b; VT* p1 = p0 + P.pitchY; VT* p2 = p0 + P.pitchZ; VT* p3 = p2 + P.pitchY;
      const float gx0 = 1.0f - c.fx, gy0 = 1.0f - c.fy, gz0 = 1.0f - c.fz;
      const float w00 = dv * gy0 * gz0, w10 = dv * c.fy * gz0, w01 = dv * gy0 * c.fz, w11 = dv * c.fy * c.fz;
      adp_vox_add(p0, w00 * gx0, P); adp_vox_add(p0 + 1, w00 * c.fx, P);
      adp_vox_add(p1, w10 * gx0, P); adp_vox_add(p1 + 1, w10 * c.fx, P);
      adp_vox_add(p2, w01 * gx0, P); adp_vox_add(p2 + 1, w01 * c.fx, P);
      adp_vox_add(p3, w11 * gx0, P); adp_vox_add(p3 + 1, w11 * c.fx, P);
    }
  }
}

size_t mrt_adaptive_scratch(int ntf) { return (size_t)MRT_ADP_COPIES * ntf * 2 * sizeof(float4); }

template <int NCH>
static cudaError_t launch_adp(const KParams& P, const AdpArgs& A, const void* vol, const float* tf, float* out_rgba,
                              const float* dL_dout, void* dvol, float* dtf, void* scratch, bool bwd, cudaStream_t st) {
  typedef typename Vox<NCH>::T VT;
  const int ntiles = P.tile_end - P.tile_begin;
  if (ntiles <= 0) return cudaSuccess;
  const int ntf = P.tfMode ? P.tfN : 2;
  const size_t smem = (size_t)ntf * sizeof(TfEntry);
  const int grid = (ntiles + 1) / 2;
  if (!bwd) {
    mrt_adaptive_kernel<NCH, false><<<grid, 128, smem, st>>>(P, A, (const VT*)vol, (const float4*)tf, (float4*)out_rgba, nullptr,
                                                             nullptr, nullptr);
    return cudaGetLastError();
  }
  if (dtf) {
    cudaError_t e = cudaMemsetAsync(scratch, 0, mrt_adaptive_scratch(ntf), st);
    if (e != cudaSuccess) return e;
  }
  mrt_adaptive_kernel<NCH, true><<<grid, 128, smem, st>>>(P, A, (const VT*)vol, (const float4*)tf, (float4*)out_rgba,
                                                          (const float4*)dL_dout, (VT*)dvol, dtf ? (float4*)scratch : nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (dtf) e = mrt_launch_dtf_reduce(scratch, MRT_ADP_COPIES, ntf, dtf, st);
  return e;
}

cudaError_t mrt_launch_adaptive(const KParams& P, int K, int J, float eps_w, int packed_ch, const void* vol, const float* tf,
                                float* out_rgba, const float* dL_dout, void* dvol, float* dtf, void* scratch, bool bwd,
                                cudaStream_t st) {
  if (K < 1 || K > MRT_ADP_MAXK || J < 1 || !(eps_w > 0.0f)) return cudaErrorInvalidValue;
  AdpArgs A; A.K = K; A.J = J; A.eps_w = eps_w;
  switch (packed_ch) {
    case 1: return launch_adp<1>(P, A, vol, tf, out_rgba, dL_dout, dvol, dtf, scratch, bwd, st);
    case 2: return launch_adp<2>(P, A, vol, tf, out_rgba, dL_dout, dvol, dtf, scratch, bwd, st);
    case 4: return launch_adp<4>(P, A, vol, tf, out_rgba, dL_dout, dvol, dtf, scratch, bwd, st);
  }
  return cudaErrorInvalidValue;
}
