// backward.cu — adjoint of the forward march w.r.t. the volume and the transfer function.
//
// Spec: docs/DifferentiableRendering.md §5-§6 (:88-127).  No reference code exists; the
// ground truth is the oracle's autograd.  Per ray the forward is re-marched in the SAME
// order with the SAME arithmetic, and the doc's O(N) T-adjoint recurrence is evaluated
// front-to-back through the identity
//     sum_{j>i} G.c_j alpha_j T_{j-1}  =  G.(C_out - bg)  -  sum_{j<=i} G.c_j alpha_j T_{j-1}
// so nothing but the forward image (and, optionally, its checkpoints) has to be stored:
//     dL/dsigma_i = dt * ( (1-alpha_i) T_{i-1} (G.c_i)  -  suffix_i  -  T_N * dL/dT_N )
//     dL/dc_i     = G * alpha_i T_{i-1}
// then through the LUT lerp (-> dL/dtf[j0], dL/dtf[j1], dL/dval), window/level, the
// modality blend and the trilinear weights (-> 8 scatter-adds per sample).
//
// Round-2 design (the round-1 kernel ran one warp per half tile over WHOLE rays: a 512^2 frame is
// less than one wave and a few 500-slot rays were the whole tail — 14 % warps active):
//   * segment-parallel: the checkpointing forward (forward.cu, CKPT) stores (C, T) every S slots
//     (docs/DifferentiableRendering.md:213 suggests exactly this) plus every ray's end slot, so a
//     ray is differentiated by ceil(k_end/S) independent warp tasks of <= S slots each;
//   * persistent CTAs pull (view, half tile, segment) tasks from a compacted list through one atomic
//     counter: no tail, any number of views per launch;
//   * dL/dtf goes to one of 64 privatised L2-resident copies with one 16-byte vector reduction per
//     touched LUT entry (reduced by a tiny second kernel).  A per-warp shared-memory histogram
//     (MATCH.ANY-ranked read-modify-writes, no atomics) was built and measured: it removes those
//     reductions but its 8 KB per warp leave the SM with 12-28 KB of L1 for the gathers; at cfg3 it
//     lost, 0.342 vs 0.295 ms (profiles/r02_time_bwd_cornercache_hist{0,1}.json), and was removed;
//   * dL/dvolume: each lane keeps the eight weights of its current cell in registers (corner cache)
//     and reduces into L2 only the corners it leaves behind when the ray moves to a neighbouring
//     cell: ~3.5 instead of 8 native reductions per slot (scalar for the folded / single-modality
//     layout, red.v2/.v4 for interleaved);
//   * brick look-ups are warp-wide: when one lane runs out of known-active slots EVERY lane extends
//     its knowledge by one cell at its own frontier (the forward's phase 1).
// Cells that are flat (one value) and empty are leapt with their closed-form dL/dtf term.
#include "march.cuh"
#include "kernels.h"
#include <limits.h>

#ifndef MRT_BWD_WARPS
#define MRT_BWD_WARPS 4
#endif
#ifndef MRT_BWD_MINB
#define MRT_BWD_MINB 5           // resident CTAs per SM the register allocation aims for (102 registers)
#endif
#define MRT_DTF_COPIES 64     // privatised dL/dtf accumulators in L2 (CTA b uses copy b % 64)

__device__ __forceinline__ void vox_atomic_add(float* p, float w, const KParams& P) {
  atomicAdd(p, w * P.wq[0]);
}
__device__ __forceinline__ void vox_atomic_add(float2* p, float w, const KParams& P) {
  atomicAdd(p, make_float2(w * P.wq[0], w * P.wq[1]));
}
__device__ __forceinline__ void vox_atomic_add(float4* p, float w, const KParams& P) {
  atomicAdd(p, make_float4(w * P.wq[0], w * P.wq[1], w * P.wq[2], w * P.wq[3]));
}

// Corner cache flushes (acc[i], i = x + 2y + 4z, belongs to voxel pb + x + y*gY + z*gZ of the gradient buffer).
#define MRT_CORNER(pb, i) ((pb) + ((i) & 1) + (((i) >> 1) & 1) * IO.gY + ((i) >> 2) * IO.gZ)
#define MRT_FLUSH1(pb, i) do { if (acc[i] != 0.0f) vox_atomic_add(MRT_CORNER(pb, i), acc[i], P); } while (0)
#define MRT_FLUSH4(pb, a, b, c_, d) do { MRT_FLUSH1(pb, a); MRT_FLUSH1(pb, b); MRT_FLUSH1(pb, c_); MRT_FLUSH1(pb, d); } while (0)
#define MRT_FLUSH_ALL(pb) do { MRT_FLUSH4(pb, 0, 1, 2, 3); MRT_FLUSH4(pb, 4, 5, 6, 7); \
    acc[0] = acc[1] = acc[2] = acc[3] = acc[4] = acc[5] = acc[6] = acc[7] = 0.0f; } while (0)

// Everything the kernel reads or writes besides the volume, in one constant block.
struct BwdIO {
  const float4* tf; const uint8_t* flat_levels; const float2* minmax;
  const int32_t* labels; const int32_t* preds;
  const float4* out_rgba; const float4* dL_dout;
  const float4* target; float gscale;   // dL_dout == nullptr: G = gscale * (out_rgba - target) (fused MSE loss)
  const float4* ck;            // checkpoints [(c-1)][view][H][W] = (C, T) before slot c*S; nullptr = unsegmented
  const int32_t* k_end;        // [view][H][W] end slot of every ray (forward's n_taken)
  float4* dtf_priv;            // [MRT_DTF_COPIES][ntf][2]: (lo, hi) = contributions to entry j0 and j0+1
  float* dray;                 // [view][H][W][6]
  unsigned long long* stats;   // [0] lane-slots shaded, [1] warp tasks that did work
  const uint2* tasks; const unsigned* ntasks; unsigned* next;
  int S, nviews;
  uint32_t g_base_off, gY, gZ;  // element pitches of the GRADIENT buffer (= the volume's unless the caller asks for another layout)
};

// HALF = 1: fp16 voxel storage (single channel); the gradient buffer stays fp32 with the SAME element
// pitches.  Sort-last shards (P.shard; whole-ray tasks only, SKIP off): the slot range is clipped to
// the sub-box, a slot is differentiated iff its base cell is owned (forward.cu's rule), the state
// starts at (0,0,0,1) and the upstream gradient of the partial's fourth channel is dL/dT_local.
template <int NCH, bool LABELS, bool SKIP, bool GENERIC, int HALF = 0>
__global__ void __launch_bounds__(32 * MRT_BWD_WARPS, MRT_BWD_MINB)
mrt_bwd_kernel(const __grid_constant__ KParams P, const __grid_constant__ CamBatch B,
               const __grid_constant__ BwdIO IO,
               const typename VoxT<NCH, HALF>::T* __restrict__ vol, typename Vox<NCH>::T* __restrict__ dvol) {
  typedef typename Vox<NCH>::T VT;
  extern __shared__ __align__(16) unsigned char s_raw[];   // [ntf] LUT | [16] labels
  const int ntf = P.tfMode ? P.tfN : 2;
  TfEntry* s_tf = reinterpret_cast<TfEntry*>(s_raw);
  float4* s_lab = reinterpret_cast<float4*>(s_tf + ntf);
  const unsigned full = 0xffffffffu;

  if (P.tfMode) {
    mrt_tf_stage(s_tf, IO.tf, ntf);
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
  const bool want_tf = IO.dtf_priv != nullptr;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  float4* const gpriv = want_tf ? IO.dtf_priv + (size_t)(blockIdx.x & (MRT_DTF_COPIES - 1)) * ntf * 2 : nullptr;
  const bool seg = IO.k_end != nullptr;
  const int nht = 2 * mrt_tiles_x_(P.W) * mrt_tiles_y_(P.H);
  const size_t npix = (size_t)P.W * P.H;
  const unsigned ntasks = __ldg(IO.ntasks);
  const float hix = (float)P.dims[0] - 1.001f, hiy = (float)P.dims[1] - 1.001f, hiz = (float)P.dims[2] - 1.001f;
  const float dt = P.dt, thr = P.thr;
  const float nm1 = (float)(ntf - 1);
  uint32_t s_tf_addr = (uint32_t)__cvta_generic_to_shared(s_tf);
  asm volatile("" : "+r"(s_tf_addr));        // opaque: keep the address in a register, do not re-derive it per sample
  const bool acc_mode = GENERIC && P.tMode == 1;           // reference-faithful running sum t += dt (unsegmented only)
  const float bgx = P.shard ? 0.0f : P.bg[0], bgy = P.shard ? 0.0f : P.bg[1], bgz = P.shard ? 0.0f : P.bg[2];
  unsigned n_shaded = 0, n_tasks = 0;

  // the queue is read one task ahead: the atomic of the NEXT fetch is in flight while this task runs
  unsigned t_raw = 0;
  if (lane == 0) t_raw = atomicAdd(IO.next, 1u);
  for (;;) {
    const unsigned t_id = __shfl_sync(full, t_raw, 0);
    if (t_id >= ntasks) break;
    const uint2 task = __ldg(IO.tasks + t_id);
    if (lane == 0) t_raw = atomicAdd(IO.next, 1u);
    const int view = (int)(task.x / (unsigned)nht), ht = (int)(task.x - (unsigned)view * (unsigned)nht), sg = (int)task.y;
    int px, py;
    mrt_pixel_of_tile_lane_fast(P, ht >> 1, mrt_logical_lane(ht & 1, lane), &px, &py);
    const bool inside = px < P.W && py < P.H;
    const size_t pixl = inside ? (size_t)py * P.W + px : 0, pix = (size_t)view * npix + pixl;
    // everything the task needs from memory, requested together (one round trip, not four)
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 Cout = inside ? __ldg(IO.out_rgba + pix) : zero4;
    float4 G = zero4;
    if (IO.dL_dout) {
      if (inside) G = __ldg(IO.dL_dout + pix);
    } else if (inside) {                        // d mean((out - target)^2) / d out
      const float4 tg = __ldg(IO.target + pix);
      G = make_float4(IO.gscale * (Cout.x - tg.x), IO.gscale * (Cout.y - tg.y), IO.gscale * (Cout.z - tg.z),
                      IO.gscale * (Cout.w - tg.w));
    }
    const int ke = (seg && inside) ? __ldg(IO.k_end + pix) : 0;
    float4 c0 = make_float4(bgx, bgy, bgz, 1.0f);                        // state before the segment's first slot
    if (seg && sg > 0 && inside) c0 = __ldg(IO.ck + ((size_t)(sg - 1) * IO.nviews + view) * npix + pixl);
    bool live = inside && (G.x != 0.0f || G.y != 0.0f || G.z != 0.0f || ((P.alphaMode || P.shard) && G.w != 0.0f));
    int k0 = 0, k1 = INT_MAX;
    if (seg) {
      k0 = sg * IO.S;
      k1 = min(ke, k0 + IO.S);
      live = live && k1 > k0;
    }
    if (!__any_sync(full, live)) continue;
    const Ray ray = mrt_setup_ray(P, B.cam[view], px, py);
    k1 = min(k1, ray.n);
    const IdxRay q = mrt_index_ray(P, ray);
    if (P.shard) {                       // candidate slots inside this shard's sub-box (ownership is decided per slot)
      int ks, ke2;
      mrt_shard_range(P, q, ray.t0, 1.0f / dt, k1, &ks, &ke2);
      k0 = max(k0, ks); k1 = min(k1, ke2);
    }
    live = live && k1 > k0;
    if (!__any_sync(full, live)) continue;
    if (!live) k1 = k0;
    ++n_tasks;

    const float S_tot = G.x * (Cout.x - bgx) + G.y * (Cout.y - bgy) + G.z * (Cout.z - bgz);
    // alphaMode 1: a = 1 - T_N  =>  dL/dT_N = -G.w ;  dsigma_i += -dt*T_N*dL/dT_N.  A shard's partial
    // carries T_local itself in .w: dL/dT_N = G.w
    const float tn_term = P.shard ? Cout.w * G.w : (P.alphaMode ? -(1.0f - Cout.w) * G.w : 0.0f);   // = T_N * dL/dT_N
    float T = c0.w;
    float prefix = G.x * (c0.x - bgx) + G.y * (c0.y - bgy) + G.z * (c0.z - bgz);
    const float ivx = 1.0f / q.dx, ivy = 1.0f / q.dy, ivz = 1.0f / q.dz;
    const float inv_dt = 1.0f / dt;
    int k = k0, kact = k0;
    float tacc = ray.t0;
    int lj = -1; float lfr = 0.0f, lacc = 0.0f;                  // dL/dsigma of the flat-empty cells leapt so far (one LUT entry)
    int ccx = -0x40000000, ccy = 0, ccz = 0;                     // corner cache: cell (none yet) and its eight weights
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

    int kblk = -1;                                             // == kact: the cell at the frontier is known flat-empty (parked)
    for (;;) {
      const bool alive = live && (seg || T > thr);
      if (SKIP) {
        // Flat-empty cells (every voxel of the cell holds the same value c AND sigma == 0 over the TF
        // bins c maps to AND no overlay label): all ns slots inside have the same (bin, frac, colour),
        // alpha == 0, so T, prefix and hence dL/dsigma are identical for every slot; the volume
        // gradient is exactly 0 (the LUT slope of sigma is 0 there and dL/dc = alpha*T*G = 0).
        // Their whole contribution is ns * dL/dsigma onto two LUT entries.
        // Look-ups are warp-wide (an instruction costs the same for 1 lane or 32): as soon as ONE lane
        // has no known-active slot left, EVERY lane extends its knowledge [k, kact) by one cell at
        // its own frontier; a lane AT a flat-empty cell leaps it, a lane that finds one ahead parks.
        const bool need = alive && k < k1 && k >= kact;
        if (__any_sync(full, need)) {
          if (alive && kact < k1 && (k >= kact || kblk != kact)) {
            const float t = fmaf((float)kact, dt, ray.t0);
            const float ppx = fmaf(t, q.dx, q.ox), ppy = fmaf(t, q.dy, q.oy), ppz = fmaf(t, q.dz, q.oz);
            const int ix = (int)fminf(fmaxf(ppx, 0.0f), hix);
            const int iy = (int)fminf(fmaxf(ppy, 0.0f), hiy);
            const int iz = (int)fminf(fmaxf(ppz, 0.0f), hiz);
            const int bid = ((iz >> MRT_BRICK_SHIFT) * P.nby + (iy >> MRT_BRICK_SHIFT)) * P.nbx + (ix >> MRT_BRICK_SHIFT);
            const int lvl = __ldg(IO.flat_levels + bid);
            const int sh = lvl ? lvl + (MRT_BRICK_SHIFT - 1) : MRT_BRICK_SHIFT;
            const int kend = min(k1, kact + mrt_cell_slots(q, ivx, ivy, ivz, ix >> sh, iy >> sh, iz >> sh, sh, t, inv_dt));
            if (!lvl) {
              kact = kend;
            } else if (k < kact) {
              kblk = kact;
            } else {
              if (want_tf) {
                const float cval = __ldg(&IO.minmax[(size_t)bid * NCH].x);
                const float val = mrt_window<GENERIC>(P, cval * P.wq[0] + P.wbias);
                if (P.tfMode || val > 0.0f) {
                  int j0; float fr;
                  const float4 rgba = mrt_tf_lookup(s_tf_addr, nm1, val, &j0, &fr);
                  const float gc = G.x * rgba.x + G.y * rgba.y + G.z * rgba.z;
                  const float dsig = (float)(kend - k) * dt * (T * gc - (S_tot - prefix) - tn_term);
                  if (j0 != lj || fr != lfr) {
                    if (lj >= 0 && lacc != 0.0f) {
                      atomicAdd(&gpriv[2 * lj].w, (1.0f - lfr) * lacc);
                      if (lfr != 0.0f) atomicAdd(&gpriv[2 * lj + 1].w, lfr * lacc);
                    }
                    lj = j0; lfr = fr; lacc = 0.0f;
                  }
                  lacc += dsig;
                }
              }
              k = kact = kend;
            }
          }
          continue;
        }
      }
      const bool on = alive && (acc_mode ? (tacc < ray.t1 && (P.maxSteps == 0 || k < P.maxSteps)) : (k < k1));
      if (!__any_sync(full, on)) break;
      bool addtf = false;
      int j0 = 0; float fr = 0.0f;
      float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (on) {
        const float t = acc_mode ? tacc : fmaf((float)k, dt, ray.t0);
        const float ppx = fmaf(t, q.dx, q.ox), ppy = fmaf(t, q.dy, q.oy), ppz = fmaf(t, q.dz, q.oz);
        const Cell c = mrt_cell(P, ppx, ppy, ppz, hix, hiy, hiz);
        const bool mine = !P.shard || mrt_shard_owns(P, c.ix(), c.iy(), c.iz());   // another shard's slot: a no-op here
        Corners<NCH, HALF> cor;
        float raw = 0.0f;
        if (mine) { cor = mrt_fetch<NCH, HALF>(P, vol, c); raw = mrt_interp<NCH, HALF>(P, cor, c); }
        const float val = mrt_window<GENERIC>(P, raw);
        if (mine && (P.tfMode || val > 0.0f)) {
          const float4 rgba = mrt_tf_lookup(s_tf_addr, nm1, val, &j0, &fr);
          float so = 1.0f; int obid = 0;           // soft occupancy (section 11): sigma' = o(brick) * sigma
          if (GENERIC) {
            if (P.occ != nullptr) { obid = mrt_brick_id(P, c.ix(), c.iy(), c.iz()); so = __ldg(P.occ + obid); }
          }
          const float alpha = mrt_alpha(P, rgba.w * so);
          const float aT = alpha * T;
          const float gc = G.x * rgba.x + G.y * rgba.y + G.z * rgba.z;
          prefix = fmaf(aT, gc, prefix);
          const float suffix = S_tot - prefix;
          float dsig = dt * ((1.0f - alpha) * T * gc - suffix - tn_term);        // dL/dsigma'
          if (GENERIC) {
            if (P.occ != nullptr) {
              if (P.docc != nullptr && dsig * rgba.w != 0.0f) atomicAdd(P.docc + obid, dsig * rgba.w);   // dL/do = dL/dsigma' * sigma
              dsig *= so;                                                            // dL/dsigma
            }
          }
          const float dr = aT * G.x, dg = aT * G.y, db = aT * G.z;
          addtf = want_tf;
          g4 = make_float4(dr, dg, db, dsig);
          if (dvol != nullptr || IO.dray != nullptr) {
            const float4 d4 = s_tf[j0].delta;
            float dval = nm1 * (dr * d4.x + dg * d4.y + db * d4.z + dsig * d4.w);
            if (GENERIC) {
              if (P.gamma != 1.0f) dval *= P.gamma * powf(__saturatef(raw), P.gamma - 1.0f);
            }
            // saturate: torch.clamp passes the gradient on the closed interval [0,1]
            const float dv = (raw >= 0.0f && raw <= 1.0f) ? dval : 0.0f;
            if (IO.dray != nullptr && dv != 0.0f) {
              // docs/DifferentiableRendering.md section 9 (:172-188): x_i = o + t_i d with fixed t_i, so
              // dL/do += dL/dx_i and dL/dd += t_i dL/dx_i, with dL/dx_i = dL/ds * ds/dx (section 6); an
              // axis on which the position was clamped (:62) carries no gradient
              float sx, sy, sz;
              mrt_interp_grad<NCH, HALF>(P, cor, c, &sx, &sy, &sz);
              const float wx = (ppx >= 0.0f && ppx <= hix) ? dv * sx / P.vs[0] : 0.0f;
              const float wy = (ppy >= 0.0f && ppy <= hiy) ? dv * sy / P.vs[1] : 0.0f;
              const float wz = (ppz >= 0.0f && ppz <= hiz) ? dv * sz / P.vs[2] : 0.0f;
              // (straight to memory: per-ray register accumulators would cost the common
              // dL/dvolume + dL/dtf case six registers at the 80-register budget of 24 warps per SM)
              float* r = IO.dray + pix * 6;
              if (wx != 0.0f) { atomicAdd(r + 0, wx); atomicAdd(r + 3, t * wx); }
              if (wy != 0.0f) { atomicAdd(r + 1, wy); atomicAdd(r + 4, t * wy); }
              if (wz != 0.0f) { atomicAdd(r + 2, wz); atomicAdd(r + 5, t * wz); }
            }
            if (dvol != nullptr && dv != 0.0f) {
              // dL/dvolume through the eight trilinear weights, accumulated in the per-lane corner
              // cache: consecutive slots of a ray share their cell or move to a face neighbour, so
              // only the corners LEFT BEHIND are reduced into L2 (about 3.5 instead of 8 per slot;
              // REDG costs about one LSU cycle per active lane, which is what bounded this kernel)
              const int ix = c.ix(), iy = c.iy(), iz = c.iz();
              const int ddx = ix - ccx, ddy = iy - ccy, ddz = iz - ccz;
              if ((ddx | ddy | ddz) != 0) {
                VT* pb = dvol + ((uint32_t)ccx + (uint32_t)ccy * IO.gY + (uint32_t)ccz * IO.gZ - IO.g_base_off);
                if (max(max(abs(ddx), abs(ddy)), abs(ddz)) > 1) {
                  MRT_FLUSH_ALL(pb);
                } else {
                  if (ddx > 0) { MRT_FLUSH4(pb, 0, 2, 4, 6); acc[0] = acc[1]; acc[2] = acc[3]; acc[4] = acc[5]; acc[6] = acc[7]; acc[1] = acc[3] = acc[5] = acc[7] = 0.0f; pb += 1; }
                  if (ddx < 0) { MRT_FLUSH4(pb, 1, 3, 5, 7); acc[1] = acc[0]; acc[3] = acc[2]; acc[5] = acc[4]; acc[7] = acc[6]; acc[0] = acc[2] = acc[4] = acc[6] = 0.0f; pb -= 1; }
                  if (ddy > 0) { MRT_FLUSH4(pb, 0, 1, 4, 5); acc[0] = acc[2]; acc[1] = acc[3]; acc[4] = acc[6]; acc[5] = acc[7]; acc[2] = acc[3] = acc[6] = acc[7] = 0.0f; pb += IO.gY; }
                  if (ddy < 0) { MRT_FLUSH4(pb, 2, 3, 6, 7); acc[2] = acc[0]; acc[3] = acc[1]; acc[6] = acc[4]; acc[7] = acc[5]; acc[0] = acc[1] = acc[4] = acc[5] = 0.0f; pb -= IO.gY; }
                  if (ddz > 0) { MRT_FLUSH4(pb, 0, 1, 2, 3); acc[0] = acc[4]; acc[1] = acc[5]; acc[2] = acc[6]; acc[3] = acc[7]; acc[4] = acc[5] = acc[6] = acc[7] = 0.0f; }
                  if (ddz < 0) { MRT_FLUSH4(pb, 4, 5, 6, 7); acc[4] = acc[0]; acc[5] = acc[1]; acc[6] = acc[2]; acc[7] = acc[3]; acc[0] = acc[1] = acc[2] = acc[3] = 0.0f; }
                }
                ccx = ix; ccy = iy; ccz = iz;
              }
              const float gx0 = 1.0f - c.fx, gy0 = 1.0f - c.fy, gz0 = 1.0f - c.fz;
              const float w00 = dv * gy0 * gz0, w10 = dv * c.fy * gz0, w01 = dv * gy0 * c.fz, w11 = dv * c.fy * c.fz;
              acc[0] = fmaf(w00, gx0, acc[0]); acc[1] = fmaf(w00, c.fx, acc[1]);
              acc[2] = fmaf(w10, gx0, acc[2]); acc[3] = fmaf(w10, c.fx, acc[3]);
              acc[4] = fmaf(w01, gx0, acc[4]); acc[5] = fmaf(w01, c.fx, acc[5]);
              acc[6] = fmaf(w11, gx0, acc[6]); acc[7] = fmaf(w11, c.fx, acc[7]);
            }
          }
          T *= (1.0f - alpha);
        }
        if (LABELS) {          // overlays carry no gradient but attenuate what lies behind
          if (P.showSeg) {
            const int l = mrt_sample_label(P, IO.labels, ppx, ppy, ppz);
            if (l > 0 && l < 8) {
              const float4 col = s_lab[l];
              prefix = fmaf(col.w * T, G.x * col.x + G.y * col.y + G.z * col.z, prefix);
              T *= (1.0f - col.w);
            }
          }
          if (P.showPred) {
            const int l = mrt_sample_label(P, IO.preds, ppx, ppy, ppz);
            if (l > 0 && l < 8) {
              const float4 col = s_lab[8 + l];
              prefix = fmaf(col.w * T, G.x * col.x + G.y * col.y + G.z * col.z, prefix);
              T *= (1.0f - col.w);
            }
          }
        }
        ++k; ++n_shaded;
        if (GENERIC) tacc += dt;
      }
      if (addtf) {
        const float f0 = 1.0f - fr;
        atomicAdd(gpriv + 2 * j0, make_float4(f0 * g4.x, f0 * g4.y, f0 * g4.z, f0 * g4.w));
        if (fr != 0.0f) atomicAdd(gpriv + 2 * j0 + 1, make_float4(fr * g4.x, fr * g4.y, fr * g4.z, fr * g4.w));
      }
    }
    if (lj >= 0 && lacc != 0.0f) {
      atomicAdd(&gpriv[2 * lj].w, (1.0f - lfr) * lacc);
      if (lfr != 0.0f) atomicAdd(&gpriv[2 * lj + 1].w, lfr * lacc);
    }
    if (dvol != nullptr && ccx >= 0) {
      VT* pb = dvol + ((uint32_t)ccx + (uint32_t)ccy * IO.gY + (uint32_t)ccz * IO.gZ - IO.g_base_off);
      MRT_FLUSH_ALL(pb);
    }
  }

  if (IO.stats != nullptr) {
    const unsigned s = __reduce_add_sync(full, n_shaded);
    if (lane == 0) { atomicAdd(IO.stats, (unsigned long long)s); atomicAdd(IO.stats + 1, (unsigned long long)n_tasks); }
  }
}

// Task list of one backward launch: every (view, half tile) of the tile range contributes
// ceil(kmax/S) segment tasks (kmax = the longest end slot among its 32 rays, recorded by the
// checkpointing forward), or exactly one when the launch is unsegmented.
__global__ void __launch_bounds__(256)
mrt_bwd_tasks_kernel(int W, int H, int tile_begin, int tile_end, int nviews, int S, const int32_t* __restrict__ warp_kmax,
                     uint2* __restrict__ tasks, unsigned* __restrict__ counters) {
  const int nht = 2 * mrt_tiles_x_(W) * mrt_tiles_y_(H);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)nviews * nht) return;
  const int ht = (int)(i % nht), tile = ht >> 1;
  if (tile < tile_begin || tile >= tile_end) return;
  int nl = 1;
  if (warp_kmax != nullptr) nl = (__ldg(warp_kmax + i) + S - 1) / S;
  if (nl <= 0) return;
  const unsigned base = atomicAdd(counters, (unsigned)nl);
  for (int s = 0; s < nl; ++s) tasks[base + s] = make_uint2((unsigned)i, (unsigned)s);
}

// dtf[j] += sum over the privatised copies of lo[j] + hi[j-1]  (hi of the last entry belongs to itself).
// One CTA per LUT entry, one thread per copy: the copies are read in parallel (a serial loop over them
// was 64 dependent L2 round trips, 23 us for a 256-entry LUT), then a warp + shared-memory reduction.
__global__ void __launch_bounds__(64)
mrt_dtf_reduce_kernel(const float4* __restrict__ priv, int ncopies, int ntf, float4* __restrict__ dtf) {
  __shared__ float4 s_part[2];
  const int j = blockIdx.x;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = threadIdx.x; c < ncopies; c += blockDim.x) {
    const float4* p = priv + (size_t)c * ntf * 2;
    const float4 a = p[2 * j];
    s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    if (j > 0) { const float4 b = p[2 * (j - 1) + 1]; s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w; }
    if (j == ntf - 1) { const float4 b = p[2 * j + 1]; s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w; }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
    s.z += __shfl_xor_sync(0xffffffffu, s.z, o); s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
  }
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float4 a = s_part[0], b = s_part[1];
    float4 d = dtf[j];
    d.x += a.x + b.x; d.y += a.y + b.y; d.z += a.z + b.z; d.w += a.w + b.w;
    dtf[j] = d;
  }
}

cudaError_t mrt_launch_dtf_reduce(const void* priv, int ncopies, int ntf, float* dtf, cudaStream_t st) {
  mrt_dtf_reduce_kernel<<<ntf, 64, 0, st>>>((const float4*)priv, ncopies, ntf, (float4*)dtf);
  return cudaGetLastError();
}

// scratch layout: [0,256) counters (ntasks, next) | privatised dL/dtf | task list
static inline size_t bwd_priv_bytes(int ntf) { return (size_t)MRT_DTF_COPIES * ntf * 2 * sizeof(float4); }
size_t mrt_bwd_scratch_bytes(int W, int H, int nviews, int ntf, int nseg) {
  const size_t nht = 2 * (size_t)mrt_tiles_x_(W) * mrt_tiles_y_(H);
  return 256 + bwd_priv_bytes(ntf) + nht * (size_t)nviews * (size_t)(nseg < 1 ? 1 : nseg) * sizeof(uint2);
}

size_t mrt_bwd_zeroed_scratch_bytes(int ntf) { return 256 + bwd_priv_bytes(ntf); }

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_num_sms = n;
    else
      return 148;
  }
  return g_num_sms;
}

template <int NCH, bool LABELS, bool SKIP, bool GENERIC, int HALF = 0>
static cudaError_t launch_bwd(const KParams& P, const CamBatch& B, int nviews, const void* vol, const MrtBwdArgs& A,
                              cudaStream_t st) {
  typedef typename Vox<NCH>::T VT;
  typedef typename VoxT<NCH, HALF>::T ST;
  const int ntiles = P.tile_end - P.tile_begin;
  if (ntiles <= 0) return cudaSuccess;
  const int ntf = P.tfMode ? P.tfN : 2;
  const int nht = 2 * mrt_tiles_x_(P.W) * mrt_tiles_y_(P.H);
  const bool seg = A.ck != nullptr && A.k_end != nullptr && A.warp_kmax != nullptr && A.seg_slots > 0;
  unsigned char* scr = reinterpret_cast<unsigned char*>(A.scratch);
  unsigned* counters = reinterpret_cast<unsigned*>(scr);
  float4* priv = A.shared_priv ? reinterpret_cast<float4*>(A.shared_priv) : reinterpret_cast<float4*>(scr + 256);
  uint2* tasks = reinterpret_cast<uint2*>(scr + 256 + bwd_priv_bytes(ntf));
  cudaError_t e = cudaSuccess;
  if (!A.scratch_zeroed) e = cudaMemsetAsync(scr, 0, 256 + (A.dtf ? bwd_priv_bytes(ntf) : 0), st);
  if (e != cudaSuccess) return e;
  const long long nthreads = (long long)nviews * nht;
  mrt_bwd_tasks_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(P.W, P.H, P.tile_begin, P.tile_end, nviews,
                                                                         seg ? A.seg_slots : 1, seg ? A.warp_kmax : nullptr,
                                                                         tasks, counters);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;

  const size_t smem = (size_t)ntf * sizeof(TfEntry) + 16 * sizeof(float4);
  auto kern = mrt_bwd_kernel<NCH, LABELS, SKIP, GENERIC, HALF>;
  // (attribute + occupancy query memoised per instantiation, device and LUT size: both are host round trips)
  static int c_dev = -1, c_occ = 0; static size_t c_smem = 0;
  int dev = 0, occ = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev == c_dev && smem == c_smem) {
    occ = c_occ;
  } else {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * MRT_BWD_WARPS, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    c_dev = dev; c_smem = smem; c_occ = occ;
  }
  const long long max_tasks = (long long)nviews * 2 * ntiles * (seg ? A.nseg : 1);
  long long grid = (max_tasks + MRT_BWD_WARPS - 1) / MRT_BWD_WARPS;
  if (grid > (long long)num_sms() * occ) grid = (long long)num_sms() * occ;
  BwdIO IO = {};
  IO.tf = (const float4*)A.tf; IO.flat_levels = A.flat_levels; IO.minmax = (const float2*)A.minmax;
  IO.labels = A.labels; IO.preds = A.preds;
  IO.out_rgba = (const float4*)A.out_rgba; IO.dL_dout = (const float4*)A.dL_dout;
  IO.target = (const float4*)A.target; IO.gscale = A.gscale;
  IO.ck = seg ? (const float4*)A.ck : nullptr; IO.k_end = seg ? A.k_end : nullptr;
  IO.dtf_priv = A.dtf ? priv : nullptr;
  IO.dray = A.dray; IO.stats = (unsigned long long*)A.stats;
  IO.tasks = tasks; IO.ntasks = counters; IO.next = counters + 1;
  IO.S = seg ? A.seg_slots : 0; IO.nviews = nviews;
  IO.gY = A.grad_pitchY ? A.grad_pitchY : P.pitchY; IO.gZ = A.grad_pitchZ ? A.grad_pitchZ : P.pitchZ;
  IO.g_base_off = A.grad_pitchY ? 0u : P.base_off;
  kern<<<(unsigned)grid, 32 * MRT_BWD_WARPS, smem, st>>>(P, B, IO, (const ST*)vol, (VT*)A.dvol);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (A.dtf && !A.no_dtf_reduce) {
    mrt_dtf_reduce_kernel<<<ntf, 64, 0, st>>>(priv, MRT_DTF_COPIES, ntf, (float4*)A.dtf);
    e = cudaGetLastError();
  }
  return e;
}

template <int NCH, bool SKIP>
static cudaError_t dispatch_bwd(const KParams& P, const CamBatch& B, int nviews, bool lab, bool gen, const void* vol,
                                const MrtBwdArgs& A, cudaStream_t st) {
  if (lab) return gen ? launch_bwd<NCH, true, SKIP, true>(P, B, nviews, vol, A, st)
                      : launch_bwd<NCH, true, SKIP, false>(P, B, nviews, vol, A, st);
  return gen ? launch_bwd<NCH, false, SKIP, true>(P, B, nviews, vol, A, st)
             : launch_bwd<NCH, false, SKIP, false>(P, B, nviews, vol, A, st);
}

// `cams` = nviews x 12 floats (eye, U, V, W per view), nullptr = the single camera in P; at most
// MRT_MAX_VIEWS views per call (the C entry chunks).
cudaError_t mrt_launch_backward(const KParams& P, const float* cams, int nviews, int packed_ch, const void* vol,
                                const MrtBwdArgs& A, cudaStream_t st) {
  CamBatch B;
  if (cams == nullptr) {
    nviews = 1;
    for (int i = 0; i < 3; ++i) { B.cam[0][i] = P.eye[i]; B.cam[0][3 + i] = P.U[i]; B.cam[0][6 + i] = P.V[i]; B.cam[0][9 + i] = P.Wv[i]; }
  } else {
    if (nviews < 1 || nviews > MRT_MAX_VIEWS) return cudaErrorInvalidValue;
    for (int v = 0; v < nviews; ++v) for (int i = 0; i < 12; ++i) B.cam[v][i] = cams[(size_t)v * 12 + i];
  }
  const bool lab = (P.showSeg || P.showPred);
  const bool gen = (P.tMode != 0) || (P.gamma != 1.0f) || (P.occ != nullptr);
  const bool skip = P.skip && A.flat_levels != nullptr && A.minmax != nullptr && P.tMode == 0 && packed_ch == 1 && !P.shard && P.occ == nullptr;
  if (P.half) {                        // fp16 storage: single channel, no overlays, indexed stepping (c_api.cu checks)
    if (P.half != 1 || packed_ch != 1 || lab || gen) return cudaErrorInvalidValue;
    return skip ? launch_bwd<1, false, true, false, 1>(P, B, nviews, vol, A, st)
                : launch_bwd<1, false, false, false, 1>(P, B, nviews, vol, A, st);
  }
  switch (packed_ch) {
    case 1: return skip ? dispatch_bwd<1, true>(P, B, nviews, lab, gen, vol, A, st)
                        : dispatch_bwd<1, false>(P, B, nviews, lab, gen, vol, A, st);
    case 2: return dispatch_bwd<2, false>(P, B, nviews, lab, gen, vol, A, st);
    case 4: return dispatch_bwd<4, false>(P, B, nviews, lab, gen, vol, A, st);
  }
  return cudaErrorInvalidValue;
}
