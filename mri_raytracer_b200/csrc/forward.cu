// forward.cu — the ray-march forward kernel (brats_main, inr/viewer/brats_rt.slang:85-168).
//
// One warp = one 8x4 half of an 8x8 screen tile (rays of a warp stay in one brick
// neighbourhood); MRT_FWD_TPB tiles per CTA.  Per ray: exact set-up (march.cuh), then the
// march as a two-phase loop that keeps the warp converged on the expensive part:
//   phase 1 (cheap, per-lane loop): locate the 8^3 brick of the current sample slot, read its
//           per-frame skip level, and leap over the largest empty aligned cell around it
//           (8..64 voxels; exact: those slots are provably no-ops) until the slot lies in an
//           active brick;
//   phase 2 (all live lanes together): shade ONE slot — fp32 trilinear of the folded scalar
//           field from the packed multi-channel layout (8 vector loads), saturate, TF
//           (shared-memory LUT), front-to-back compositing, early ray termination.
#include "march.cuh"
#include "kernels.h"

#ifndef MRT_RUN_MIN
#define MRT_RUN_MIN 1           // slots of known-active run a lane likes to keep ahead of itself
#endif
#ifndef MRT_FWD_MINB
#define MRT_FWD_MINB 10         // resident CTAs per SM the register allocation aims for (single-channel variants; measured 8: 0.634, 9: 0.643, 10: 0.618 ms per cfg2 batch)
#endif
#ifndef MRT_FWD_TPB
#define MRT_FWD_TPB 2           // 8x8 tiles per CTA  (CTA = 64*TPB threads)
#endif

template <bool B> struct BoolTag { static constexpr bool value = B; };

// CKPT (training forward): additionally stores (C, T) of every ray before each slot c*CK.S, the
// ray's end slot and every warp's longest end slot, for the segment-parallel backward (backward.cu).
template <int NCH, bool LABELS, bool SKIP, bool GENERIC, int HALF, bool CKPT = false>
__global__ void __launch_bounds__(64 * MRT_FWD_TPB, (NCH == 4 ? 768 : 128 * ((LABELS || GENERIC || CKPT) ? 8 : MRT_FWD_MINB)) / (64 * MRT_FWD_TPB))
mrt_fwd_kernel(const __grid_constant__ KParams P,
               const __grid_constant__ CamBatch B,
               const __grid_constant__ StripTargets S,
               const __grid_constant__ CkptOut CK,
               const typename VoxT<NCH, HALF>::T* __restrict__ vol,
               const float4* __restrict__ tf,
               const uint8_t* __restrict__ levels,
               const int32_t* __restrict__ labels,
               const int32_t* __restrict__ preds,
               float4* __restrict__ out_rgba,
               float* __restrict__ out_T,
               int4* __restrict__ out_counts) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  unsigned char* s_tf = s_raw;                                              // [tfN] entries of mrt_tf_stride bytes
  const uint32_t tf_stride = mrt_tf_stride(P.tfN);
  float4* s_lab = reinterpret_cast<float4*>(s_raw + (P.tfMode ? (size_t)P.tfN * tf_stride : 0));   // [0..7] seg, [8..15] pred

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tile = P.tile_begin + mrt_middle_out(blockIdx.x, gridDim.x) * MRT_FWD_TPB + (warp >> 1);
  bool tile_ok = tile < P.tile_end;
  if (S.row_mod > 1) {          // interleaved tile rows: local tile index -> (local row, column) -> global tile id
    const int txn = mrt_tiles_x_(P.W), li = tile - P.tile_begin;
    const int lrow = li / txn, ty = S.row_rem + S.row_mod * lrow;
    tile_ok = tile_ok && ty < mrt_tiles_y_(P.H);
    tile = ty * txn + (li - lrow * txn);
  }
  const int view = blockIdx.y;                                             // batch of views: one camera each
  int px = 0, py = 0;
  if (tile_ok) mrt_pixel_of_tile_lane_fast(P, tile, mrt_logical_lane(warp & 1, lane), &px, &py);
  // :89 — pixels outside the image keep their lane alive (the skip loop uses warp votes) with an
  // empty ray, and never store
  const bool inside = tile_ok && (px < P.W) && (py < P.H);
  const size_t pix = ((size_t)view * P.H + py) * P.W + px;
  float4* dst = out_rgba + pix;
  if (S.view_base != nullptr) dst = S.view_base[view] + ((size_t)py * P.W + px);
  if (S.n && inside) { const int strip = py / S.rows; dst = S.base[strip] + ((size_t)(py - strip * S.rows) * P.W + px); }

  // Cull against the active-brick box BEFORE any expensive work: a ray that cannot enter it is
  // pure background.  Whole CTAs of such rays (most of the frame outside the head) skip the LUT
  // staging and the exact ray set-up; whole warps skip the set-up.
  bool maybe = inside;
  ActiveBox abox;
  bool stored = true;                 // sparse gather: tiles outside the view's spans are filled by the image's owner
  if (SKIP) {
    if (!GENERIC && !CKPT) {                                               // the counting / checkpointing variants need every exact n
      if (S.spans != nullptr) {
        // the view's precomputed spans (projected hull of the active box, mrt_view_spans) answer the
        // question for the whole tile: one load and two compares instead of a per-ray slab test
        bool in_span = false;
        if (tile_ok) {
          const int2 sp = __ldg(S.spans + (size_t)view * mrt_tiles_y_(P.H) + (py >> MRT_TILE_SHIFT));
          in_span = mrt_tile_in_span(sp, px & ~MRT_TILE_MASK);              // warp-uniform (one tile per warp pair)
        }
        maybe = inside && in_span;
        stored = in_span || (S.store_outside != 0);
      } else {
        abox = mrt_active_box(P, levels);
        maybe = inside && mrt_ray_may_hit(P, B.cam[view], px, py, abox);
      }
      const bool cta_any = __syncthreads_or(maybe);
      if (!cta_any || !__any_sync(0xffffffffu, maybe)) {
        if (inside && stored) {
          *dst = P.shard ? make_float4(0.0f, 0.0f, 0.0f, 1.0f)
                         : make_float4(P.bg[0], P.bg[1], P.bg[2], P.alphaMode ? 0.0f : 1.0f);
          if (out_T) out_T[pix] = 1.0f;
        }
        if (!cta_any) return;                                              // nobody needs the LUT
        maybe = false;                                                     // this warp only keeps the barrier company
      }
    }
    if (GENERIC || CKPT || S.spans != nullptr) abox = mrt_active_box(P, levels);   // (still needed to clip the slot ranges)
  }

  if (P.tfMode) mrt_tf_stage_strided(s_tf, tf_stride, tf, P.tfN);
  if (LABELS) {
    if (threadIdx.x < 16) {
      const int l = threadIdx.x & 7;
      const float boost = threadIdx.x < 8 ? 1.0f : 1.5f;                   // :158
      const float a = mrt_alpha(P, P.lut[l][3] * boost);                   // :147
      s_lab[threadIdx.x] = make_float4(P.lut[l][0], P.lut[l][1], P.lut[l][2], (l > 0) ? a : 0.0f);
    }
  }
  __syncthreads();
  if (!tile_ok) return;
  if (SKIP && !GENERIC && !CKPT) { if (!__any_sync(0xffffffffu, maybe)) return; }   // culled warp (already stored)

  Ray ray = mrt_setup_ray(P, B.cam[view], px, py);
  if (!maybe) ray.n = 0;
  // a sort-last shard renders a partial: premultiplied colour WITHOUT background, alpha = T_local
  float Cr = P.shard ? 0.0f : P.bg[0], Cg = P.shard ? 0.0f : P.bg[1], Cb = P.shard ? 0.0f : P.bg[2];   // :111
  float T = 1.0f;                                                          // :112
  int k = 0, n_eval = 0, n_seg = 0;
  // checkpoint c (1 <= c < CK.nseg) = state before slot c*CK.S, layer c-1 of CK.ck ([layer][view][H][W])
  int next_ck = CKPT ? CK.S : 0x7fffffff;
  float4* ckp = CKPT ? CK.ck + pix : nullptr;
  const size_t ck_layer = CKPT ? (size_t)gridDim.y * P.H * P.W : 0;
  const int ck_end = CKPT ? CK.nseg * CK.S : 0;
  auto ck_flush = [&](int kk) {          // the state is final for every slot boundary <= kk
    while (next_ck <= kk && next_ck < ck_end) {
      if (inside) *ckp = make_float4(Cr, Cg, Cb, T);
      ckp += ck_layer; next_ck += CK.S;
    }
  };

  {                                           // rays with n == 0 (miss / outside) run zero iterations below
    const IdxRay q = mrt_index_ray(P, ray);
    const float hix = (float)P.dims[0] - 1.001f, hiy = (float)P.dims[1] - 1.001f, hiz = (float)P.dims[2] - 1.001f;
    const float dt = P.dt, thr = P.thr;
    float nm1 = (float)(P.tfN - 1);
    asm volatile("" : "+f"(nm1));              // opaque: one register, not an I2F per sample
    uint32_t s_tf_adj = (uint32_t)__cvta_generic_to_shared(s_tf) - 0x4b000000u * tf_stride;
    asm volatile("" : "+r"(s_tf_adj));         // opaque: keep the address in a register, do not re-derive it per sample
    uint32_t s_tf_stride = tf_stride;
    asm volatile("" : "+r"(s_tf_stride));
    // sample slot k sits at index-space position so + k*sd  (t_k = t0 + k*dt folded into the ray: one
    // fma per axis; the look-ups of phase 1 use the same expression, so skipping stays exact)
    float sox = fmaf(ray.t0, q.dx, q.ox), soy = fmaf(ray.t0, q.dy, q.oy), soz = fmaf(ray.t0, q.dz, q.oz);
    float sdx = q.dx * dt, sdy = q.dy * dt, sdz = q.dz * dt;

    // one sample slot at index-space position pp (:119-162).  `on` = false turns the slot into an
    // exact no-op WITHOUT a branch (alpha := 0), so a warp's run of slots is straight-line code;
    // INT = the caller has proven the position inside [0, dims-1.001] (no clamps needed).
    auto shade_p = [&](float ppx, float ppy, float ppz, bool on, auto tfm, auto inter) {
      constexpr bool TFM = decltype(tfm)::value, INT = decltype(inter)::value;
      const Cell c = mrt_cell_t<!INT>(ppx, ppy, ppz, hix, hiy, hiz);
      const float val = mrt_window<GENERIC>(P, mrt_sample_raw<NCH, HALF>(P, vol, c));
      float so = P.neg_dt_log2e;               // soft occupancy (section 11): sigma' = o(brick) * sigma
      if (GENERIC) { if (P.occ != nullptr) so *= __ldg(P.occ + mrt_brick_id(P, c.ix(), c.iy(), c.iz())); }
      // alpha = 1 - e, e = exp(-sigma dt) (:137); C += alpha T rgb, T *= 1 - alpha (:138-139) as
      // T' = T e, alpha T = T - T'  (same quantities, two instructions less)
      if (TFM) {
        const float4 rgba = mrt_tf_lookup_strided(s_tf_adj, s_tf_stride, nm1, val);
        const float e = mrt_ex2(rgba.w * so);
        const float Tn = T * (on ? e : 1.0f);
        const float aT = T - Tn;
        Cr = fmaf(aT, rgba.x, Cr); Cg = fmaf(aT, rgba.y, Cg); Cb = fmaf(aT, rgba.z, Cb);
        T = Tn;
      } else {
        // :135 — val == 0 gives e == 1 exactly: the (val > 0) gate is implicit
        const float e = mrt_ex2(val * P.ia * so);                          // :136-137
        const float Tn = T * (on ? e : 1.0f);
        const float c1 = (T - Tn) * val;                                   // :138
        Cr += c1; Cg += c1; Cb += c1;
        T = Tn;                                                            // :139
      }
      if (LABELS) {
        if (P.showSeg) {                                                   // :143-151
          const int l = mrt_sample_label(P, labels, ppx, ppy, ppz);
          if (on && l > 0 && l < 8) {
            const float4 col = s_lab[l];
            const float aT = col.w * T;
            Cr = fmaf(aT, col.x, Cr); Cg = fmaf(aT, col.y, Cg); Cb = fmaf(aT, col.z, Cb);
            T *= (1.0f - col.w);
          }
        }
        if (P.showPred) {                                                  // :154-162
          const int l = mrt_sample_label(P, preds, ppx, ppy, ppz);
          if (on && l > 0 && l < 8) {
            const float4 col = s_lab[8 + l];
            const float aT = col.w * T;
            Cr = fmaf(aT, col.x, Cr); Cg = fmaf(aT, col.y, Cg); Cb = fmaf(aT, col.z, Cb);
            T *= (1.0f - col.w);
          }
        }
      }
    };
    const BoolTag<true> yes; const BoolTag<false> no;

    if (GENERIC && P.tMode == 1) {
      // reference-faithful running sum t += stepSize (:113,:164); no skipping possible
      auto loop = [&](auto tfm) {
        float t = ray.t0;
        while (ray.n > 0 && t < ray.t1 && T > thr && (P.maxSteps == 0 || k < P.maxSteps)) {
          shade_p(fmaf(t, q.dx, q.ox), fmaf(t, q.dy, q.oy), fmaf(t, q.dz, q.oz), true, tfm, no);
          t += dt; ++k; ++n_eval;
        }
      };
      if (P.tfMode) loop(yes); else loop(no);
    } else if (SKIP) {
      const float inv_dt = P.inv_dt;
      int n = ray.n;
      if (P.shard) { int ks; mrt_shard_range(P, q, ray.t0, inv_dt, ray.n, &ks, &n); k = ks; }
      const int n_full = n;
      {   // slots outside the ray's interval in the active-brick box are no-ops: clip [k, n) to it
        float tin, tout;
        mrt_box_interval(abox, q.ox, q.oy, q.oz, q.dx, q.dy, q.dz, &tin, &tout);
        if (tout >= fmaxf(tin, 0.0f)) {
          const float a = floorf((tin - ray.t0) * inv_dt) - 1.0f, b = ceilf((tout - ray.t0) * inv_dt) + 1.0f;
          k = max(k, (int)fminf(fmaxf(a, 0.0f), (float)n));
          n = min(n, (int)fminf(fmaxf(b, 0.0f), (float)n));
        } else {
          n = k;
        }
      }
      // Clamp-free march: every slot of [k, n] lies within 2 steps (+ margins) of the active box, so
      // when the box widened by 2.5 steps stays inside [0, dims-1.001] the sampler's clamps are
      // identities for the whole warp.  A lane without slots never shades for real, but it rides
      // along masked: park it on a harmless position (the box centre).
      bool interior = false;
      if (!GENERIC && !LABELS) {
        bool lane_int = true;
        if (n > k) {
          const float mx = 2.5f * fabsf(sdx), my = 2.5f * fabsf(sdy), mz = 2.5f * fabsf(sdz);
          lane_int = (abox.lo[0] - mx >= 0.0f) && (abox.hi[0] + mx <= hix) && (abox.lo[1] - my >= 0.0f) &&
                     (abox.hi[1] + my <= hiy) && (abox.lo[2] - mz >= 0.0f) && (abox.hi[2] + mz <= hiz);
        } else if (abox.hi[0] >= abox.lo[0]) {
          sox = 0.5f * (abox.lo[0] + abox.hi[0]); soy = 0.5f * (abox.lo[1] + abox.hi[1]); soz = 0.5f * (abox.lo[2] + abox.hi[2]);
          sdx = sdy = sdz = 0.0f;
        } else {
          lane_int = false;                                                // no active brick at all: nothing to shade anyway
        }
        interior = !P.shard && __all_sync(0xffffffffu, lane_int);
      }
      const SlotRay sr = mrt_slot_ray(sox, soy, soz, sdx, sdy, sdz);
      // a warp's run of m slots: straight-line, masked by the lane's own liveness
      // (a masked lane still FETCHES at its frozen slot: always inside the buffer for a whole volume —
      // the sampler clamps to it — but not for a shard's sub-volume, which therefore branches instead)
      auto run = [&](int m, bool live, auto tfm, auto inter) {
        bool on = live;
        float kf = (float)k;
        if (P.shard) {
          for (int i = 0; i < m; ++i) {
            if (on) {
              shade_p(fmaf(kf, sdx, sox), fmaf(kf, sdy, soy), fmaf(kf, sdz, soz), true, tfm, no);
              kf += 1.0f; if (GENERIC) ++n_eval;
              if (CKPT) { if ((int)kf == next_ck) ck_flush(next_ck); }
              on = T > thr;
            }
          }
        } else {
          for (int i = 0; i < m; ++i) {
            shade_p(fmaf(kf, sdx, sox), fmaf(kf, sdy, soy), fmaf(kf, sdz, soz), on, tfm, inter);
            if (on) { kf += 1.0f; if (GENERIC) ++n_eval; }
            if (CKPT) { if (on && (int)kf == next_ck) ck_flush(next_ck); }
            on = on && (T > thr);
          }
        }
        k = (int)kf;
      };
      // Per-lane knowledge of the ray, in slot indices (k <= kact <= kf <= kl):
      //   [k, kact)   current run, inside active bricks: to be shaded
      //   [kact, kf)  known empty (leapt cells / another shard's slots)
      //   [kf, kl)    look-ahead run, known active (empty when kl == kf)
      //   [kl, n)     unknown
      int kact = k, kf = k, kl = k;
      for (;;) {
        if (k >= kact && T > thr) {                       // current run exhausted: take the look-ahead
          k = kf; kact = kl; kf = kl;                     // (not after ERT: k stays the oracle's n_taken)
          if (CKPT) ck_flush(k);                          // the slots leapt over are no-ops
        }
        const bool live = (k < n) && (T > thr);            // :117
        // phase 1 (warp-wide): as soon as ONE lane has nothing to shade, EVERY lane extends its
        // knowledge by one brick look-up at its own frontier kl — an instruction costs the same
        // for 1 lane or 32, so the other lanes' look-ups ride along for free and the (divergent)
        // look-up code runs once per ~brick length instead of once per shaded slot.
        // one warp reduction answers both questions: bit 0 = some lane is live, bit 1 = some live
        // lane has nothing to shade
        // (a lane also asks for a look-up when its run is short and directly extensible, so that
        // the warp-uniform runs of phase 2 stay long)
        const bool want = live && kl < n && (k >= kact || (kl == kact && kact - k < MRT_RUN_MIN));
        const unsigned wst = __reduce_or_sync(0xffffffffu, (live ? 1u : 0u) | (want ? 2u : 0u));
        if (wst & 2u) {
          if (live && kl < n) {
            const float klf = (float)kl;
            const float ppx = fmaf(klf, sdx, sox), ppy = fmaf(klf, sdy, soy), ppz = fmaf(klf, sdz, soz);
            int ix, iy, iz;                                                // == floor of the clamped coord
            if (interior) { ix = (int)ppx; iy = (int)ppy; iz = (int)ppz; }
            else {
              ix = (int)fminf(fmaxf(ppx, 0.0f), hix); iy = (int)fminf(fmaxf(ppy, 0.0f), hiy); iz = (int)fminf(fmaxf(ppz, 0.0f), hiz);
            }
            int lvl = 1, kend = kl + 1;                                   // another rank's slot: a 1-slot gap
            if (!P.shard || mrt_shard_owns(P, ix, iy, iz)) {
              const int jx = ix - P.slo[0], jy = iy - P.slo[1], jz = iz - P.slo[2];   // brick grid is shard-local
              lvl = __ldg(levels + (((jz >> MRT_BRICK_SHIFT) * P.nby + (jy >> MRT_BRICK_SHIFT)) * P.nbx +
                                    (jx >> MRT_BRICK_SHIFT)));
              // 0 = active brick, 1..4 = empty cell, 0x80 | 2..4 = all-active cell; edge 2^((lvl & 7) + 2)
              const int sh = lvl ? (lvl & 7) + (MRT_BRICK_SHIFT - 1) : MRT_BRICK_SHIFT;
              lvl = (lvl & 0x80) ? 0 : lvl;
              kend = min(n, kl + mrt_cell_slots_k(sr, jx, jy, jz, sh, klf, P.slo[0], P.slo[1], P.slo[2]));
              // an ACTIVE brick may straddle the shard's far faces: stop at the owned box's exit
              if (P.shard && !lvl)
                kend = min(kend, kl + mrt_shard_slots(P, q, mrt_rcp(q.dx), mrt_rcp(q.dy), mrt_rcp(q.dz),
                                                      fmaf(klf, dt, ray.t0), inv_dt));
              if (GENERIC) ++n_seg;
            }
            if (!lvl) {
              if (kl == kact) kact = kf = kl = kend;       // active, directly behind the current run: extend it
              else kl = kend;                              // active behind a gap: start / extend the look-ahead run
            } else if (kl == kf) kf = kl = kend;           // empty, directly behind the gap: widen the gap
            // (an empty cell behind a look-ahead run cannot be recorded yet: looked up again later)
          }
          continue;
        }
        if (!wst) break;
        // phase 2: every live lane owns a run; the warp shades m = the shortest remaining run
        // slots back to back (no lane can run dry before that, so nothing has to be re-checked
        // but the lane's own early termination)
        const int m = __reduce_min_sync(0xffffffffu, live ? kact - k : 0x7fffffff);
        if (P.tfMode) { if (interior) run(m, live, yes, yes); else run(m, live, yes, no); }
        else          { if (interior) run(m, live, no, yes);  else run(m, live, no, no); }
      }
      if (T > thr) {                       // ran to the end: the oracle's n_taken counts the clipped no-op slots too
        k = n_full;
        if (CKPT) ck_flush(n_full - 1);
      }
    } else {
      int n = ray.n;
      if (P.shard) { int ks; mrt_shard_range(P, q, ray.t0, P.inv_dt, ray.n, &ks, &n); k = ks; }
      auto loop = [&](auto tfm) {
        while (k < n && T > thr) {                                         // :117
          const float kf = (float)k;
          const float ppx = fmaf(kf, sdx, sox), ppy = fmaf(kf, sdy, soy), ppz = fmaf(kf, sdz, soz);
          if (P.shard) {
            const int ix = (int)fminf(fmaxf(ppx, 0.0f), hix), iy = (int)fminf(fmaxf(ppy, 0.0f), hiy),
                      iz = (int)fminf(fmaxf(ppz, 0.0f), hiz);
            if (!mrt_shard_owns(P, ix, iy, iz)) { ++k; continue; }
          }
          shade_p(ppx, ppy, ppz, true, tfm, no);
          ++k; if (GENERIC) ++n_eval;
          if (CKPT) { if (k == next_ck) ck_flush(k); }
        }
      };
      if (P.tfMode) loop(yes); else loop(no);
      if (CKPT) { if (T > thr) ck_flush(n - 1); }
    }
  }
  if (CKPT) {                           // (culled warps / CTAs returned above: the launcher zero-fills both arrays)
    const int km = __reduce_max_sync(0xffffffffu, inside ? k : 0);
    const size_t nht = 2 * (size_t)mrt_tiles_x_(P.W) * mrt_tiles_y_(P.H);
    if (lane == 0) CK.warp_kmax[(size_t)view * nht + 2 * (size_t)tile + (warp & 1)] = km;
    if (inside) CK.k_end[pix] = k;
  }
  if (!inside) return;
  *dst = make_float4(Cr, Cg, Cb, P.shard ? T : (P.alphaMode ? 1.0f - T : 1.0f));  // :167
  if (out_T) out_T[pix] = T;
  if (GENERIC) { if (out_counts) out_counts[pix] = make_int4(ray.n, k, n_eval, n_seg); }
}

// ------------------------------------------------------------------------- dispatch
static const StripTargets g_no_targets = {};      // plain [views][H][W] image, every tile stored
static cudaError_t mrt_launch_forward_to(const KParams& P, const StripTargets& S, const float* cams, int nviews,
                                         int packed_ch, const void* vol, const float* tf, const uint8_t* levels,
                                         const int32_t* labels, const int32_t* preds, float* out_rgba, float* out_T,
                                         int32_t* out_counts, cudaStream_t st);
static const CkptOut g_no_ckpt = {};
template <int NCH, bool LABELS, bool SKIP, bool GENERIC, int HALF = 0, bool CKPT = false>
static cudaError_t launch_fwd(const KParams& P, const CamBatch& B, const StripTargets& S, int nviews, const void* vol, const float* tf, const uint8_t* levels,
                              const int32_t* labels, const int32_t* preds, float* out_rgba, float* out_T,
                              int32_t* out_counts, cudaStream_t st, const CkptOut& CK = g_no_ckpt) {
  const int ntiles = P.tile_end - P.tile_begin;
  if (ntiles <= 0) return cudaSuccess;
  const int grid = (ntiles + MRT_FWD_TPB - 1) / MRT_FWD_TPB;
  const size_t smem = (size_t)(P.tfMode ? P.tfN : 0) * (P.tfN <= 512 ? 48 : 32) + 16 * sizeof(float4);
  mrt_fwd_kernel<NCH, LABELS, SKIP, GENERIC, HALF, CKPT><<<dim3(grid, nviews), 64 * MRT_FWD_TPB, smem, st>>>(
      P, B, S, CK, (const typename VoxT<NCH, HALF>::T*)vol, (const float4*)tf, levels, labels, preds,
      (float4*)out_rgba, out_T, (int4*)out_counts);
  return cudaGetLastError();
}

// Training forward: one launch of <= MRT_MAX_VIEWS views that also records the checkpoints.
template <int NCH>
static cudaError_t dispatch_fwd_ckpt(const KParams& P, const CamBatch& B, int nviews, bool lab, bool skip, const void* vol,
                                     const float* tf, const uint8_t* levels, const int32_t* labels, const int32_t* preds,
                                     float* o, const CkptOut& CK, cudaStream_t st) {
  if (lab) return skip ? launch_fwd<NCH, true, true, false, 0, true>(P, B, g_no_targets, nviews, vol, tf, levels, labels, preds, o, nullptr, nullptr, st, CK)
                       : launch_fwd<NCH, true, false, false, 0, true>(P, B, g_no_targets, nviews, vol, tf, levels, labels, preds, o, nullptr, nullptr, st, CK);
  return skip ? launch_fwd<NCH, false, true, false, 0, true>(P, B, g_no_targets, nviews, vol, tf, levels, labels, preds, o, nullptr, nullptr, st, CK)
              : launch_fwd<NCH, false, false, false, 0, true>(P, B, g_no_targets, nviews, vol, tf, levels, labels, preds, o, nullptr, nullptr, st, CK);
}
cudaError_t mrt_launch_forward_ckpt(const KParams& P, const float* cams, int nviews, int packed_ch, const void* vol,
                                    const float* tf, const uint8_t* levels, const int32_t* labels, const int32_t* preds,
                                    float* out_rgba, float* ck, int seg_slots, int nseg, int32_t* k_end, int32_t* warp_kmax,
                                    cudaStream_t st, bool clear_aux) {
  if (P.half || P.shard || P.tMode != 0 || P.gamma != 1.0f || seg_slots < 1 || nseg < 1 || !k_end || !warp_kmax ||
      (nseg > 1 && !ck))
    return cudaErrorInvalidValue;
  CamBatch B;
  if (cams == nullptr) {
    nviews = 1;
    for (int i = 0; i < 3; ++i) { B.cam[0][i] = P.eye[i]; B.cam[0][3 + i] = P.U[i]; B.cam[0][6 + i] = P.V[i]; B.cam[0][9 + i] = P.Wv[i]; }
  } else {
    if (nviews < 1 || nviews > MRT_MAX_VIEWS) return cudaErrorInvalidValue;
    for (int v = 0; v < nviews; ++v) for (int i = 0; i < 12; ++i) B.cam[v][i] = cams[(size_t)v * 12 + i];
  }
  const size_t npix = (size_t)P.W * P.H, nht = 2 * (size_t)mrt_tiles_x_(P.W) * mrt_tiles_y_(P.H);
  // every tile of [tile_begin, tile_end) writes its k_end / warp_kmax entries (the checkpointing variant culls nothing);
  // the entries of the OTHER tiles are defined as zero: clear unless the range is the whole image or the caller
  // covers the image with several launches (clear_aux = false)
  cudaError_t e = cudaSuccess;
  if (clear_aux && !(P.tile_begin == 0 && P.tile_end == mrt_tiles_x_(P.W) * mrt_tiles_y_(P.H))) {
    e = cudaMemsetAsync(k_end, 0, npix * nviews * sizeof(int32_t), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(warp_kmax, 0, nht * nviews * sizeof(int32_t), st);
  }
  if (e != cudaSuccess) return e;
  CkptOut CK;
  CK.ck = reinterpret_cast<float4*>(ck); CK.S = seg_slots; CK.nseg = nseg; CK.k_end = k_end; CK.warp_kmax = warp_kmax;
  const bool lab = (P.showSeg || P.showPred);
  const bool skip = P.skip && levels != nullptr;
  switch (packed_ch) {
    case 1: return dispatch_fwd_ckpt<1>(P, B, nviews, lab, skip, vol, tf, levels, labels, preds, out_rgba, CK, st);
    case 2: return dispatch_fwd_ckpt<2>(P, B, nviews, lab, skip, vol, tf, levels, labels, preds, out_rgba, CK, st);
    case 4: return dispatch_fwd_ckpt<4>(P, B, nviews, lab, skip, vol, tf, levels, labels, preds, out_rgba, CK, st);
  }
  return cudaErrorInvalidValue;
}

template <int NCH>
static cudaError_t dispatch_fwd(const KParams& P, const CamBatch& B, const StripTargets& T, int nviews, bool lab, bool skip, bool gen, const void* vol, const float* tf,
                                const uint8_t* levels, const int32_t* labels, const int32_t* preds,
                                float* o, float* oT, int32_t* oc, cudaStream_t st) {
#define MRT_CASE(L, S, G) if (lab == L && skip == S && gen == G) \
    return launch_fwd<NCH, L, S, G>(P, B, T, nviews, vol, tf, levels, labels, preds, o, oT, oc, st);
  MRT_CASE(false, false, false) MRT_CASE(false, true, false)
  MRT_CASE(true, false, false)  MRT_CASE(true, true, false)
  MRT_CASE(false, false, true)  MRT_CASE(false, true, true)
  MRT_CASE(true, false, true)   MRT_CASE(true, true, true)
#undef MRT_CASE
  return cudaErrorInvalidValue;
}

// `cams` = nviews x 12 floats (eye, U, V, W per view), nullptr = the single camera in P.  The
// outputs are [nviews][H][W](...) contiguous.  Views are rendered by ONE launch per chunk of
// MRT_MAX_VIEWS (blockIdx.y = view): the short CTAs of one view fill the SMs that the long
// central rays of the previous one leave idle, so the per-launch tail is paid once per batch.
// Sparse framebuffer gather.  The span of a (view, tile row) is the x-extent of everything that can be
// non-background there: the UNION, over the ACTIVE BRICKS, of the bounding rectangles of their projected boxes
// (mrt_project_box: margins and outward rounding as before, now per brick).  Round 2 first projected only the
// bounding box of all active bricks: for the bench's head that hull holds 5 500 tiles per 1024^2 view, the union
// of the brick footprints 3 200 — everything in between was marched, stored and sent over PCIe / NVLink although
// it is pure background.  Every sample slot a ray can evaluate lies in an active brick, i.e. the pixel lies inside
// that brick's projection: pixels outside the union are exactly the background.  Integer min / max reductions:
// deterministic in (P, cam, levels), so the sender of a sparse gather and the owner of the image still compute the
// same spans independently.  mrt_spans_init_kernel empties the spans first.
// mrt_fill_outside_kernel (receiving side): background into every tile outside its row's span — exactly the
// tiles the senders skip.  Same tile geometry as the march.
__global__ void __launch_bounds__(128)
mrt_spans_init_kernel(int n, int2* __restrict__ spans) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) spans[i] = make_int2(0x7fffffff, -1);
}
// One thread per (brick, view).  The CTA's 128 consecutive bricks (four x-rows of the bench grid) touch nearly the
// same bands, so their rectangles are first merged in shared memory (SMEM: one int2 per band) and every touched
// band then costs the CTA one look and at most two reductions in global memory — the per-thread version sent a
// hundred reductions at each of a view's 8 cache lines.  SMEM = false (more bands than fit): straight to global.
template <bool SMEM>
__global__ void __launch_bounds__(128)
mrt_view_spans_kernel(const __grid_constant__ KParams P, const __grid_constant__ CamBatch B, int nviews,
                      const uint8_t* __restrict__ levels, int2* __restrict__ spans) {
  extern __shared__ int2 s_sp[];                                           // [ty] when SMEM
  const int b = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  const int nb = P.nbx * P.nby * P.nbz;
  const int ty = mrt_tiles_y_(P.H);
  if (v >= nviews) return;                                                 // (whole CTA)
  int2* sp = spans + (size_t)v * ty;
  if (SMEM) {
    for (int i = threadIdx.x; i < ty; i += blockDim.x) s_sp[i] = make_int2(0x7fffffff, -1);
    __syncthreads();
  }
  int2* dst = SMEM ? s_sp : sp;
  bool work = b < nb;
  if (work) {
    const int lvl = __ldg(levels + b);
    work = lvl == 0 || (lvl & 0x80);                                       // an empty brick: no slot in it is ever evaluated
  }
  if (work) {
    const int bx = b % P.nbx, by = (b / P.nbx) % P.nby, bz = b / (P.nbx * P.nby);
    // A brick whose six face neighbours are all active adds nothing to the union: a ray through it enters and leaves
    // through a face (or an edge / corner, which the neighbours' boxes, widened by the margin, contain as well), i.e.
    // through an active neighbour whose own footprint covers the pixel.  Only the surface bricks project — a third
    // of the bench head's active bricks, a seventh of a solid 512^3 interior's.
    if (bx > 0 && bx < P.nbx - 1 && by > 0 && by < P.nby - 1 && bz > 0 && bz < P.nbz - 1) {
      const int sxy = P.nbx * P.nby;
      const int l0 = __ldg(levels + b - 1), l1 = __ldg(levels + b + 1), l2 = __ldg(levels + b - P.nbx),
                l3 = __ldg(levels + b + P.nbx), l4 = __ldg(levels + b - sxy), l5 = __ldg(levels + b + sxy);
      const auto act = [](int l) { return l == 0 || (l & 0x80) != 0; };
      if (act(l0) && act(l1) && act(l2) && act(l3) && act(l4) && act(l5)) work = false;
    }
    if (work) {
      ActiveBox A;                                                         // as mrt_active_box, for this one brick
      A.lo[0] = (float)((bx << MRT_BRICK_SHIFT) + P.slo[0]) - MRT_BOX_MARGIN; A.hi[0] = (float)(((bx + 1) << MRT_BRICK_SHIFT) + P.slo[0]) + MRT_BOX_MARGIN;
      A.lo[1] = (float)((by << MRT_BRICK_SHIFT) + P.slo[1]) - MRT_BOX_MARGIN; A.hi[1] = (float)(((by + 1) << MRT_BRICK_SHIFT) + P.slo[1]) + MRT_BOX_MARGIN;
      A.lo[2] = (float)((bz << MRT_BRICK_SHIFT) + P.slo[2]) - MRT_BOX_MARGIN; A.hi[2] = (float)(((bz + 1) << MRT_BRICK_SHIFT) + P.slo[2]) + MRT_BOX_MARGIN;
      float cx[8], cy[8];
      int rx0 = 0, rx1 = P.W - 1, b0 = 0, b1 = ty - 1;                     // behind the eye / degenerate basis: no culling
      if (mrt_project_box(P, B.cam[v], A, cx, cy) == 0) {
        float ymin = cy[0], ymax = cy[0], xmin = cx[0], xmax = cx[0];
#pragma unroll
        for (int c = 1; c < 8; ++c) {
          ymin = fminf(ymin, cy[c]); ymax = fmaxf(ymax, cy[c]); xmin = fminf(xmin, cx[c]); xmax = fmaxf(xmax, cx[c]);
        }
        // The brick contributes its bounding RECTANGLE to every band it touches: at the usual zoom (a brick = 3-5
        // tiles) that is at most a tile looser than the band's cut through the footprint's hull, for a tenth of the
        // arithmetic and, above all, bounded work per thread: cutting hulls band by band — and merging all-active
        // cells into one box, whose thread then walked 30+ bands alone — took 31-34 us per 8-view batch under ncu.
        const float lim = 1.0e8f;
        rx0 = max(0, (int)floorf(fmaxf(xmin, -lim)) - 1); rx1 = min(P.W - 1, (int)ceilf(fminf(xmax, lim)) + 1);
        // bands whose pixel rows, widened by one pixel like the x-extent, meet [ymin, ymax]
        b0 = max(0, (int)floorf((fmaxf(ymin, -16.0f) - (float)MRT_TILE_EDGE) * (1.0f / MRT_TILE_EDGE)));
        b1 = min(ty - 1, (int)floorf((fminf(ymax, (float)P.H + 16.0f) + 1.0f) * (1.0f / MRT_TILE_EDGE)));
        if (!(ymax >= -2.0f) || !(ymin <= (float)P.H + 1.0f)) b1 = b0 - 1;  // off screen
      }
      if (rx0 <= rx1) {
        for (int band = b0; band <= b1; ++band) {
          if (SMEM) {
            atomicMin(&dst[band].x, rx0); atomicMax(&dst[band].y, rx1);
          } else {
            // (a plain look first: a stale value is only ever LOOSER than the current one, so a reduction that
            // could matter is never skipped, and interior bricks stop hammering the same few words)
            const int2 cur = __ldcg(sp + band);
            if (rx0 < cur.x) atomicMin(&sp[band].x, rx0);
            if (rx1 > cur.y) atomicMax(&sp[band].y, rx1);
          }
        }
      }
    }
  }
  if (SMEM) {
    __syncthreads();
    for (int i = threadIdx.x; i < ty; i += blockDim.x) {
      const int2 m = s_sp[i];
      if (m.x > m.y) continue;
      const int2 cur = __ldcg(sp + i);
      if (m.x < cur.x) atomicMin(&sp[i].x, m.x);
      if (m.y > cur.y) atomicMax(&sp[i].y, m.y);
    }
  }
}
__global__ void __launch_bounds__(256)
mrt_fill_outside_kernel(const __grid_constant__ KParams P, const int2* __restrict__ spans,
                        const int2* __restrict__ prev, int nviews, float4* __restrict__ out) {
  // one WARP per tile (two 512-byte warp stores), grid-stride over (view, tile)
  const float4 bgp = P.shard ? make_float4(0.0f, 0.0f, 0.0f, 1.0f)
                             : make_float4(P.bg[0], P.bg[1], P.bg[2], P.alphaMode ? 0.0f : 1.0f);
  const int lane = threadIdx.x & 31;
  const int ntiles = P.tile_end - P.tile_begin, ty = mrt_tiles_y_(P.H);
  const size_t total = (size_t)ntiles * nviews;
  const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t m = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < total; m += nwarps) {
    const int view = (int)(m / (size_t)ntiles), tile = P.tile_begin + (int)(m - (size_t)view * ntiles);
    int px, py;
    mrt_pixel_of_tile_lane_fast(P, tile, lane, &px, &py);                 // lanes 0..31 = the tile's upper half
    if (mrt_tile_in_span(__ldg(spans + (size_t)view * ty + (py >> MRT_TILE_SHIFT)), px & ~MRT_TILE_MASK)) continue;
    // delta fill: the image already holds the background outside `prev` (the spans of the batch that
    // last wrote this buffer), so only tiles that were inside those and are outside the new ones change
    if (prev != nullptr && !mrt_tile_in_span(__ldg(prev + (size_t)view * ty + (py >> MRT_TILE_SHIFT)), px & ~MRT_TILE_MASK)) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int y = py + 4 * h;
      if (px < P.W && y < P.H) out[((size_t)view * P.H + y) * P.W + px] = bgp;
    }
  }
}
cudaError_t mrt_launch_view_spans(const KParams& P, const float* cams, int nviews, const uint8_t* levels, int32_t* spans,
                                  cudaStream_t st) {
  const int ty = mrt_tiles_y_(P.H);
  for (int v0 = 0; v0 < nviews; v0 += MRT_MAX_VIEWS) {
    const int nv = (nviews - v0 < MRT_MAX_VIEWS) ? nviews - v0 : MRT_MAX_VIEWS;
    CamBatch B;
    for (int v = 0; v < nv; ++v) for (int i = 0; i < 12; ++i) B.cam[v][i] = cams[(size_t)(v0 + v) * 12 + i];
    int2* sp = reinterpret_cast<int2*>(spans) + (size_t)v0 * ty;
    const int nb = P.nbx * P.nby * P.nbz;
    mrt_spans_init_kernel<<<(nv * ty + 127) / 128, 128, 0, st>>>(nv * ty, sp);
    const size_t smem = (size_t)ty * sizeof(int2);
    if (smem <= 32 * 1024) mrt_view_spans_kernel<true><<<dim3((nb + 127) / 128, nv), 128, smem, st>>>(P, B, nv, levels, sp);
    else mrt_view_spans_kernel<false><<<dim3((nb + 127) / 128, nv), 128, 0, st>>>(P, B, nv, levels, sp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
cudaError_t mrt_launch_fill_outside(const KParams& P, int nviews, const int32_t* spans, const int32_t* prev_spans, float* out_rgba,
                                    cudaStream_t st) {
  const int ntiles = P.tile_end - P.tile_begin;
  if (ntiles <= 0 || nviews <= 0) return cudaSuccess;
  size_t grid = ((size_t)ntiles * nviews + 7) / 8;
  if (grid > 148 * 16) grid = 148 * 16;
  mrt_fill_outside_kernel<<<(int)grid, 256, 0, st>>>(P, reinterpret_cast<const int2*>(spans),
                                                     reinterpret_cast<const int2*>(prev_spans), nviews, (float4*)out_rgba);
  return cudaGetLastError();
}

cudaError_t mrt_launch_forward_sparse(const KParams& P, const float* cams, int nviews, int packed_ch, const void* vol,
                                      const float* tf, const uint8_t* levels, float* out_rgba, const int32_t* spans,
                                      int store_outside, cudaStream_t st, float* const* view_base, int row_mod,
                                      int row_rem) {
  if (nviews > MRT_MAX_VIEWS) return cudaErrorInvalidValue;     // the caller chunks (span offsets go with it)
  StripTargets S = {};
  S.spans = reinterpret_cast<const int2*>(spans);
  S.store_outside = store_outside;
  S.view_base = reinterpret_cast<float4* const*>(view_base);
  if (row_mod > 1) {
    if (row_rem < 0 || row_rem >= row_mod) return cudaErrorInvalidValue;
    S.row_mod = row_mod; S.row_rem = row_rem;
    // the launch enumerates the rank's LOCAL tiles: its rows x all columns
    KParams Q = P;
    const int ty = mrt_tiles_y_(P.H), rows = ty > row_rem ? (ty - row_rem + row_mod - 1) / row_mod : 0;
    Q.tile_begin = 0; Q.tile_end = rows * mrt_tiles_x_(P.W);
    return mrt_launch_forward_to(Q, S, cams, nviews, packed_ch, vol, tf, levels, nullptr, nullptr, out_rgba, nullptr,
                                 nullptr, st);
  }
  return mrt_launch_forward_to(P, S, cams, nviews, packed_ch, vol, tf, levels, nullptr, nullptr, out_rgba, nullptr,
                               nullptr, st);
}

cudaError_t mrt_launch_forward_strips(const KParams& P, int packed_ch, const void* vol, const float* tf,
                                      const uint8_t* levels, float* const* strip_out, int nstrips, int strip_rows,
                                      cudaStream_t st) {
  if (nstrips < 1 || nstrips > MRT_MAX_STRIPS || strip_rows < 1 || (long long)nstrips * strip_rows < P.H)
    return cudaErrorInvalidValue;
  StripTargets S = {};
  for (int i = 0; i < nstrips; ++i) S.base[i] = reinterpret_cast<float4*>(strip_out[i]);
  S.n = nstrips; S.rows = strip_rows;
  return mrt_launch_forward_to(P, S, nullptr, 1, packed_ch, vol, tf, levels, nullptr, nullptr, strip_out[0], nullptr,
                               nullptr, st);
}

cudaError_t mrt_launch_forward(const KParams& P, const float* cams, int nviews, int packed_ch, const void* vol,
                               const float* tf, const uint8_t* levels, const int32_t* labels, const int32_t* preds,
                               float* out_rgba, float* out_T, int32_t* out_counts, cudaStream_t st) {
  if (cams != nullptr && nviews > MRT_MAX_VIEWS) {
    const size_t npix = (size_t)P.W * P.H;
    for (int v0 = 0; v0 < nviews; v0 += MRT_MAX_VIEWS) {
      const int nv = (nviews - v0 < MRT_MAX_VIEWS) ? nviews - v0 : MRT_MAX_VIEWS;
      cudaError_t e = mrt_launch_forward(P, cams + (size_t)v0 * 12, nv, packed_ch, vol, tf, levels, labels, preds,
                                         out_rgba + (size_t)v0 * npix * 4, out_T ? out_T + (size_t)v0 * npix : nullptr,
                                         out_counts ? out_counts + (size_t)v0 * npix * 4 : nullptr, st);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  return mrt_launch_forward_to(P, g_no_targets, cams, nviews, packed_ch, vol, tf, levels, labels, preds, out_rgba, out_T,
                               out_counts, st);
}

// one launch (<= MRT_MAX_VIEWS views) with explicit output targets (strips / sparse spans)
static cudaError_t mrt_launch_forward_to(const KParams& P, const StripTargets& S, const float* cams, int nviews,
                                         int packed_ch, const void* vol, const float* tf, const uint8_t* levels,
                                         const int32_t* labels, const int32_t* preds, float* out_rgba, float* out_T,
                                         int32_t* out_counts, cudaStream_t st) {
  CamBatch B;
  if (cams == nullptr) {
    nviews = 1;
    for (int i = 0; i < 3; ++i) { B.cam[0][i] = P.eye[i]; B.cam[0][3 + i] = P.U[i]; B.cam[0][6 + i] = P.V[i]; B.cam[0][9 + i] = P.Wv[i]; }
  } else {
    if (nviews < 1) return cudaSuccess;
    for (int v = 0; v < nviews; ++v) for (int i = 0; i < 12; ++i) B.cam[v][i] = cams[(size_t)v * 12 + i];
  }
  const bool lab = (P.showSeg || P.showPred);
  const bool skip = P.skip && levels != nullptr && P.tMode == 0;
  const bool gen = (P.tMode != 0) || (P.gamma != 1.0f) || (out_counts != nullptr) || (P.occ != nullptr);
  if (P.half) {            // fp16 / u8 / quad voxels: single channel, no label overlays (c_api.cu checks)
    if (packed_ch != 1 || lab) return cudaErrorInvalidValue;
#define MRT_NARROW(H) \
    if (skip) return gen ? launch_fwd<1, false, true, true, H>(P, B, S, nviews, vol, tf, levels, labels, preds, out_rgba, out_T, out_counts, st) \
                         : launch_fwd<1, false, true, false, H>(P, B, S, nviews, vol, tf, levels, labels, preds, out_rgba, out_T, out_counts, st); \
    return gen ? launch_fwd<1, false, false, true, H>(P, B, S, nviews, vol, tf, levels, labels, preds, out_rgba, out_T, out_counts, st) \
               : launch_fwd<1, false, false, false, H>(P, B, S, nviews, vol, tf, levels, labels, preds, out_rgba, out_T, out_counts, st);
    if (P.half == 1) { MRT_NARROW(1) }
    if (P.half == 3) { MRT_NARROW(3) }
    if (P.half == 4) { MRT_NARROW(4) }
    MRT_NARROW(2)
#undef MRT_NARROW
  }
  switch (packed_ch) {
    case 1: return dispatch_fwd<1>(P, B, S, nviews, lab, skip, gen, vol, tf, levels, labels, preds, out_rgba, out_T, out_counts, st);
    case 2: return dispatch_fwd<2>(P, B, S, nviews, lab, skip, gen, vol, tf, levels, labels, preds, out_rgba, out_T, out_counts, st);
    case 4: return dispatch_fwd<4>(P, B, S, nviews, lab, skip, gen, vol, tf, levels, labels, preds, out_rgba, out_T, out_counts, st);
  }
  return cudaErrorInvalidValue;
}
