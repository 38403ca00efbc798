// forward_tma.cu — the staged-brick march: the MEASURED ALTERNATIVE to forward.cu's direct gathers.
//
// north_star asks for "bricked 3D tiles staged via TMA/shared memory".  This kernel does exactly
// that for the single-channel fp32 sampler: one CTA per screen tile (8x8 or 16x16 pixels); the CTA
// lists the boxes (1^3 or 2^3 bricks = 8^3 / 16^3 voxels, +1 halo) its ray bundle crosses, front to
// back, skips the ones whose bricks are all empty, and streams them through a double-buffered
// shared-memory stage with 3-D TMA box loads (cp.async.bulk.tensor.3d, tensor map over the packed
// volume; completion on an mbarrier).  Every ray shades the slots whose trilinear base cell lies in
// the staged box FROM SHARED MEMORY with the arithmetic of forward.cu (same sampler, same TF, same
// compositing: brats_rt.slang:60-76,117-139), so the image is that of mrt_fwd_kernel.
//
// Exactness does not depend on the box order being right: a ray only ever advances its own slot
// counter, and whatever the staged traversal leaves unshaded (a box missing from the list, a list
// that overflowed, an order violation) is finished by the direct-gather loop at the end; `stats`
// counts both kinds of slots so that a measurement can tell how much really went through the stage.
//
// Measured result (DESIGN.md "TMA staged bricks"): slower than the direct gathers at every
// configuration — the shared-memory gathers cost the same L1 data-stage wavefronts as L1 hits, the
// rays of a tile idle while boxes they do not cross are staged, and the per-box barrier serialises
// the CTA.  Kept as a selectable variant (mrt_render_forward_tma) and as the regression test of
// that claim, not as the default path.
#include <cuda.h>
#include "march.cuh"
#include "kernels.h"

#define MRT_TMA_MAXLIST 1024

template <int BOXB> struct TmaBox {
  static constexpr int V = 8 * BOXB;                 // voxels per box edge
  static constexpr int SH = 3 + (BOXB == 2 ? 1 : 0); // log2(V)
  static constexpr int NX = ((V + 1 + 3) / 4) * 4;   // staged x extent (inner TMA box dimension: multiple of 16 B)
  static constexpr int NY = V + 1, NZ = V + 1;
  static constexpr int FLOATS = NX * NY * NZ;
  static constexpr int BYTES = FLOATS * 4;
  static constexpr int STRIDE = (BYTES + 127) & ~127; // 128-byte aligned stage
};

__device__ __forceinline__ void tma_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tma_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "LAB_WAIT: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra LAB_WAIT;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// stats: [0] slots shaded from the shared-memory stage, [1] slots finished by direct gathers,
//        [2] boxes staged (TMA loads), [3] lists that overflowed MRT_TMA_MAXLIST
template <int BOXB, int TILE>
__global__ void __launch_bounds__(TILE * TILE)
mrt_fwd_tma_kernel(const __grid_constant__ KParams P, const __grid_constant__ CUtensorMap tmap,
                   const float* __restrict__ vol, const float4* __restrict__ tf, const uint8_t* __restrict__ levels,
                   float4* __restrict__ out_rgba, unsigned long long* __restrict__ stats) {
  typedef TmaBox<BOXB> BX;
  constexpr int NT = TILE * TILE, NW = NT / 32;
  extern __shared__ __align__(128) unsigned char s_raw[];
  float* s_box = reinterpret_cast<float*>(s_raw);                                   // 2 stages
  TfEntry* s_tf = reinterpret_cast<TfEntry*>(s_raw + 2 * BX::STRIDE);               // [tfN] (2 entries for tfMode 0)
  __shared__ __align__(8) unsigned long long s_bar[2];
  __shared__ int s_list[MRT_TMA_MAXLIST];
  __shared__ float s_corner[4][6];                   // so, sd of the bundle's four corner rays
  __shared__ float s_red[NW][6];
  __shared__ int s_cnt[NT], s_n;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int txn = mrt_tiles_x_(P.W), tyn = mrt_tiles_y_(P.H);
  // CTA -> 8x8 tile (TILE == 8) or a 2x2 block of them (TILE == 16); warp = one 8x4 half tile as in forward.cu
  int tx, ty;
  if (TILE == 8) { tx = blockIdx.x % txn; ty = blockIdx.x / txn; }
  else {
    const int bxn = (txn + 1) >> 1;
    const int st = warp >> 1;
    tx = 2 * (blockIdx.x % bxn) + (st & 1); ty = 2 * (blockIdx.x / bxn) + (st >> 1);
  }
  const int ll = mrt_logical_lane(warp & 1, lane);
  const int px = (tx << MRT_TILE_SHIFT) + (ll & MRT_TILE_MASK), py = (ty << MRT_TILE_SHIFT) + (ll >> MRT_TILE_SHIFT);
  const bool inside = tx < txn && ty < tyn && px < P.W && py < P.H;

  const int ntf = P.tfMode ? P.tfN : 2;
  if (P.tfMode) mrt_tf_stage(s_tf, tf, ntf);
  else if (t == 0) {      // the reference intensity TF (:135-138) as the 2-entry LUT [(0,0,0,0), (1,1,1,intensityAlpha)]
    s_tf[0].base = make_float4(0.f, 0.f, 0.f, 0.f); s_tf[0].delta = make_float4(1.f, 1.f, 1.f, P.ia);
    s_tf[1].base = make_float4(1.f, 1.f, 1.f, P.ia); s_tf[1].delta = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&s_bar[0]);
  if (t == 0) {
    tma_mbar_init(bar0, 1); tma_mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_n = 0;
  }

  // ---- per-ray set-up, exactly as forward.cu
  Ray ray = mrt_setup_ray(P, P.eye, inside ? px : 0, inside ? py : 0);
  if (!inside) ray.n = 0;
  const IdxRay q = mrt_index_ray(P, ray);
  const float hix = (float)P.dims[0] - 1.001f, hiy = (float)P.dims[1] - 1.001f, hiz = (float)P.dims[2] - 1.001f;
  const float dt = P.dt, thr = P.thr;
  const float nm1 = (float)(ntf - 1);
  const uint32_t s_tf_adj = (uint32_t)__cvta_generic_to_shared(s_tf) - MRT_TF_ADJ;
  const float sox = fmaf(ray.t0, q.dx, q.ox), soy = fmaf(ray.t0, q.dy, q.oy), soz = fmaf(ray.t0, q.dz, q.oz);
  const float sdx = q.dx * dt, sdy = q.dy * dt, sdz = q.dz * dt;
  const ActiveBox abox = mrt_active_box(P, levels);
  int k = 0, n = ray.n;
  {
    float tin, tout;
    mrt_box_interval(abox, q.ox, q.oy, q.oz, q.dx, q.dy, q.dz, &tin, &tout);
    if (tout >= fmaxf(tin, 0.0f)) {
      const float a = floorf((tin - ray.t0) * P.inv_dt) - 1.0f, b = ceilf((tout - ray.t0) * P.inv_dt) + 1.0f;
      k = max(k, (int)fminf(fmaxf(a, 0.0f), (float)n));
      n = min(n, (int)fminf(fmaxf(b, 0.0f), (float)n));
    } else {
      n = k;
    }
  }
  const SlotRay sr = mrt_slot_ray(sox, soy, soz, sdx, sdy, sdz);
  float Cr = P.bg[0], Cg = P.bg[1], Cb = P.bg[2], T = 1.0f;

  // ---- the bundle: AABB of all ray segments [slot k, slot n-1] (clamped like the sampler), corner rays
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  if (n > k) {
    const float ka = (float)k, kb = (float)(n - 1);
    const float a[3] = {fmaf(ka, sdx, sox), fmaf(ka, sdy, soy), fmaf(ka, sdz, soz)};
    const float b[3] = {fmaf(kb, sdx, sox), fmaf(kb, sdy, soy), fmaf(kb, sdz, soz)};
    const float hh[3] = {hix, hiy, hiz};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      lo[i] = fminf(fmaxf(fminf(a[i], b[i]), 0.0f), hh[i]); hi[i] = fminf(fmaxf(fmaxf(a[i], b[i]), 0.0f), hh[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      lo[i] = fminf(lo[i], __shfl_xor_sync(0xffffffffu, lo[i], o)); hi[i] = fmaxf(hi[i], __shfl_xor_sync(0xffffffffu, hi[i], o));
    }
    if (lane == 0) { s_red[warp][i] = lo[i]; s_red[warp][3 + i] = hi[i]; }
  }
  {   // corner rays of the pixel block (pixel centres: every ray of the CTA lies in their convex hull)
    const int cx0 = (TILE == 8 ? tx : (tx & ~1)) << MRT_TILE_SHIFT, cy0 = (TILE == 8 ? ty : (ty & ~1)) << MRT_TILE_SHIFT;
    if (t < 4) {
      const int cx = min(cx0 + ((t & 1) ? TILE - 1 : 0), P.W - 1), cy = min(cy0 + ((t & 2) ? TILE - 1 : 0), P.H - 1);
      const Ray cr = mrt_setup_ray(P, P.eye, cx, cy);
      const IdxRay cq = mrt_index_ray(P, cr);
      s_corner[t][0] = cq.ox; s_corner[t][1] = cq.oy; s_corner[t][2] = cq.oz;
      s_corner[t][3] = cq.dx; s_corner[t][4] = cq.dy; s_corner[t][5] = cq.dz;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    lo[i] = s_red[0][i]; hi[i] = s_red[0][3 + i];
    for (int w = 1; w < NW; ++w) { lo[i] = fminf(lo[i], s_red[w][i]); hi[i] = fmaxf(hi[i], s_red[w][3 + i]); }
  }
  const bool any = hi[0] >= lo[0];
  // major axis of the bundle = largest direction component of corner ray 0 (index space)
  int ax = 0;
  {
    const float ddx = fabsf(s_corner[0][3]), ddy = fabsf(s_corner[0][4]), ddz = fabsf(s_corner[0][5]);
    ax = (ddy > ddx && ddy >= ddz) ? 1 : ((ddz > ddx && ddz > ddy) ? 2 : 0);
  }
  const int au = (ax + 1) % 3, av = (ax + 2) % 3;
  int cb0[3], cb1[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { cb0[i] = any ? ((int)lo[i]) >> BX::SH : 0; cb1[i] = any ? ((int)hi[i]) >> BX::SH : -1; }
  const int nslab = cb1[ax] - cb0[ax] + 1;
  const bool fwd_a = s_corner[0][3 + ax] > 0.0f;
  // outward-from-the-eye order inside a slab (a valid visibility order for every ray of a pinhole
  // camera; for an orthographic one the "eye" sits at infinity against the view direction)
  int eu, ev;
  if (P.ortho) {
    eu = s_corner[0][3 + au] > 0.0f ? -(1 << 28) : (1 << 28); ev = s_corner[0][3 + av] > 0.0f ? -(1 << 28) : (1 << 28);
  } else {
    eu = (int)floorf(fminf(fmaxf(s_corner[0][au], -1.0e6f), 1.0e6f)) >> BX::SH; ev = (int)floorf(fminf(fmaxf(s_corner[0][av], -1.0e6f), 1.0e6f)) >> BX::SH;
  }
  const int nb[3] = {P.nbx, P.nby, P.nbz};
  auto box_active = [&](int bx_, int by_, int bz_) -> bool {       // any of the box's bricks non-empty?
    bool act = false;
#pragma unroll
    for (int i = 0; i < BOXB * BOXB * BOXB; ++i) {
      const int jx = bx_ * BOXB + (i % BOXB), jy = by_ * BOXB + ((i / BOXB) % BOXB), jz = bz_ * BOXB + (i / (BOXB * BOXB));
      if (jx < nb[0] && jy < nb[1] && jz < nb[2]) {
        const int lvl = __ldg(levels + ((size_t)jz * nb[1] + jy) * nb[0] + jx);
        act = act || lvl == 0 || (lvl & 0x80);
      }
    }
    return act;
  };
  // slab s (one per thread and pass): rectangle of boxes the bundle's hull crosses, then the active ones, in order
  auto slab_boxes = [&](int si, bool emit, int at) -> int {
    const int s = fwd_a ? cb0[ax] + si : cb1[ax] - si;
    const float p0 = (float)(s << BX::SH) - 1.0f, p1 = (float)((s + 1) << BX::SH) + 1.0f;
    float u0 = 3.0e38f, u1 = -3.0e38f, v0 = 3.0e38f, v1 = -3.0e38f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float oa = s_corner[c][ax], da = s_corner[c][3 + ax];
      const float inv = (da != 0.0f) ? 1.0f / da : 0.0f;
      const float ta = (p0 - oa) * inv, tb = (p1 - oa) * inv;
      const float ua = fmaf(ta, s_corner[c][3 + au], s_corner[c][au]), ub = fmaf(tb, s_corner[c][3 + au], s_corner[c][au]);
      const float va = fmaf(ta, s_corner[c][3 + av], s_corner[c][av]), vb = fmaf(tb, s_corner[c][3 + av], s_corner[c][av]);
      u0 = fminf(u0, fminf(ua, ub)); u1 = fmaxf(u1, fmaxf(ua, ub)); v0 = fminf(v0, fminf(va, vb)); v1 = fmaxf(v1, fmaxf(va, vb));
    }
    const float hh[3] = {hix, hiy, hiz};
    const int bu0 = max(cb0[au], ((int)fminf(fmaxf(u0 - 1.0f, 0.0f), hh[au])) >> BX::SH);
    const int bu1 = min(cb1[au], ((int)fminf(fmaxf(u1 + 1.0f, 0.0f), hh[au])) >> BX::SH);
    const int bv0 = max(cb0[av], ((int)fminf(fmaxf(v0 - 1.0f, 0.0f), hh[av])) >> BX::SH);
    const int bv1 = min(cb1[av], ((int)fminf(fmaxf(v1 + 1.0f, 0.0f), hh[av])) >> BX::SH);
    int cnt = 0;
    const int nu = bu1 - bu0 + 1, nv = bv1 - bv0 + 1;
    for (int iu = 0; iu < nu; ++iu) {
      // outward from eu: first eu.. upward, then below eu downward
      const int us = min(max(eu, bu0), bu1 + 1), nup = bu1 - us + 1;
      const int u = iu < nup ? us + iu : us - 1 - (iu - nup);
      for (int iv = 0; iv < nv; ++iv) {
        const int vs = min(max(ev, bv0), bv1 + 1), nvp = bv1 - vs + 1;
        const int v = iv < nvp ? vs + iv : vs - 1 - (iv - nvp);
        int b3[3]; b3[ax] = s; b3[au] = u; b3[av] = v;
        if (box_active(b3[0], b3[1], b3[2])) {
          if (emit && at + cnt < MRT_TMA_MAXLIST) s_list[at + cnt] = b3[0] | (b3[1] << 10) | (b3[2] << 20);
          ++cnt;
        }
      }
    }
    return cnt;
  };
  int total = 0;
  for (int base = 0; base < nslab; base += NT) {          // NT slabs per pass, in front-to-back order
    const int si = base + t;
    const int c = (si < nslab) ? slab_boxes(si, false, 0) : 0;
    s_cnt[t] = c;
    __syncthreads();
    int off = total;
    for (int j = 0; j < t; ++j) off += s_cnt[j];         // (NT <= 256 additions; list building is a small part of the kernel)
    int sum = 0;
    for (int j = 0; j < NT; ++j) sum += s_cnt[j];
    if (si < nslab && c) slab_boxes(si, true, off);
    total += sum;
    __syncthreads();
  }
  const int nlist = min(total, MRT_TMA_MAXLIST);
  if (t == 0 && stats) {
    atomicAdd(stats + 2, (unsigned long long)nlist);
    if (total > MRT_TMA_MAXLIST) atomicAdd(stats + 3, 1ull);
  }

  // ---- one slot, from the staged box (`sm` != nullptr) or from global memory; same arithmetic as forward.cu
  auto composite = [&](float raw) {
    const float val = __saturatef(raw);
    if (P.tfMode) {
      const float4 rgba = mrt_tf_lookup_adj(s_tf_adj, nm1, val);
      const float e = mrt_ex2(rgba.w * P.neg_dt_log2e);
      const float Tn = T * e;
      const float aT = T - Tn;
      Cr = fmaf(aT, rgba.x, Cr); Cg = fmaf(aT, rgba.y, Cg); Cb = fmaf(aT, rgba.z, Cb);
      T = Tn;
    } else {
      const float e = mrt_ex2(val * P.ia * P.neg_dt_log2e);
      const float Tn = T * e;
      const float c1 = (T - Tn) * val;
      Cr += c1; Cg += c1; Cb += c1;
      T = Tn;
    }
  };
  auto level_at = [&](int ix, int iy, int iz) -> int {
    return __ldg(levels + (((iz >> MRT_BRICK_SHIFT) * P.nby + (iy >> MRT_BRICK_SHIFT)) * P.nbx + (ix >> MRT_BRICK_SHIFT)));
  };
  unsigned long long n_stage = 0, n_direct = 0;

  // ---- the staged traversal
  const uint32_t box_s = (uint32_t)__cvta_generic_to_shared(s_box);
  auto issue = [&](int c) {
    const int e = s_list[c];
    const int bx_ = e & 1023, by_ = (e >> 10) & 1023, bz_ = (e >> 20) & 1023;
    const uint32_t bar = bar0 + 8 * (c & 1);
    tma_mbar_expect(bar, BX::BYTES);
    tma_load_3d(box_s + (c & 1) * BX::STRIDE, &tmap, bx_ << BX::SH, by_ << BX::SH, bz_ << BX::SH, bar);
  };
  if (t == 0) {
    if (nlist > 0) issue(0);
    if (nlist > 1) issue(1);
  }
  for (int c = 0; c < nlist; ++c) {
    tma_mbar_wait(bar0 + 8 * (c & 1), (c >> 1) & 1);
    const int e = s_list[c];
    const int bx_ = e & 1023, by_ = (e >> 10) & 1023, bz_ = (e >> 20) & 1023;
    const float* sm = s_box + (c & 1) * (BX::STRIDE / 4);
    // advance over empty cells; shade while the slot's base cell lies in THIS box
    while (k < n && T > thr) {
      const float kf = (float)k;
      const float ppx = fmaf(kf, sdx, sox), ppy = fmaf(kf, sdy, soy), ppz = fmaf(kf, sdz, soz);
      const int ix = (int)fminf(fmaxf(ppx, 0.0f), hix), iy = (int)fminf(fmaxf(ppy, 0.0f), hiy), iz = (int)fminf(fmaxf(ppz, 0.0f), hiz);
      const int lvl = level_at(ix, iy, iz);
      if (lvl != 0 && !(lvl & 0x80)) {                     // empty cell: leap
        k = min(n, k + mrt_cell_slots_k(sr, ix, iy, iz, (lvl & 7) + (MRT_BRICK_SHIFT - 1), kf, 0, 0, 0));
        continue;
      }
      if ((ix >> BX::SH) != bx_ || (iy >> BX::SH) != by_ || (iz >> BX::SH) != bz_) break;   // another box: wait for it
      const int kend = min(n, k + mrt_cell_slots_k(sr, ix, iy, iz, BX::SH, kf, 0, 0, 0));
      for (; k < kend && T > thr; ++k) {
        const float kk = (float)k;
        const Cell cc = mrt_cell_t<true>(fmaf(kk, sdx, sox), fmaf(kk, sdy, soy), fmaf(kk, sdz, soz), hix, hiy, hiz);
        const int lx = cc.ix() - (bx_ << BX::SH), ly = cc.iy() - (by_ << BX::SH), lz = cc.iz() - (bz_ << BX::SH);
        // (the conservative cell exit keeps the slot inside the box; should rounding ever disagree, fall through to a direct gather)
        if ((unsigned)lx >= (unsigned)BX::V || (unsigned)ly >= (unsigned)BX::V || (unsigned)lz >= (unsigned)BX::V) break;
        const float* p0 = sm + (lz * BX::NY + ly) * BX::NX + lx;
        Corners<1, 0> cr;
        cr.v[0] = p0[0]; cr.v[1] = p0[1]; cr.v[2] = p0[BX::NX]; cr.v[3] = p0[BX::NX + 1];
        cr.v[4] = p0[BX::NX * BX::NY]; cr.v[5] = p0[BX::NX * BX::NY + 1];
        cr.v[6] = p0[BX::NX * BX::NY + BX::NX]; cr.v[7] = p0[BX::NX * BX::NY + BX::NX + 1];
        composite(mrt_interp<1, 0>(P, cr, cc));
        ++n_stage;
      }
      if (k < kend) break;                                 // early termination or the guard above
    }
    __syncthreads();                                       // everyone is done with stage c & 1
    if (t == 0 && c + 2 < nlist) issue(c + 2);
  }

  // ---- whatever is left: direct gathers (exactness never depends on the list)
  while (k < n && T > thr) {
    const float kf = (float)k;
    const float ppx = fmaf(kf, sdx, sox), ppy = fmaf(kf, sdy, soy), ppz = fmaf(kf, sdz, soz);
    const int ix = (int)fminf(fmaxf(ppx, 0.0f), hix), iy = (int)fminf(fmaxf(ppy, 0.0f), hiy), iz = (int)fminf(fmaxf(ppz, 0.0f), hiz);
    const int lvl = level_at(ix, iy, iz);
    const int sh = lvl ? (lvl & 7) + (MRT_BRICK_SHIFT - 1) : MRT_BRICK_SHIFT;
    const int kend = min(n, k + mrt_cell_slots_k(sr, ix, iy, iz, sh, kf, 0, 0, 0));
    if (lvl != 0 && !(lvl & 0x80)) { k = kend; continue; }
    for (; k < kend && T > thr; ++k) {
      const float kk = (float)k;
      const Cell cc = mrt_cell_t<true>(fmaf(kk, sdx, sox), fmaf(kk, sdy, soy), fmaf(kk, sdz, soz), hix, hiy, hiz);
      composite(mrt_sample_raw<1, 0>(P, vol, cc));
      ++n_direct;
    }
  }
  if (stats) {
#pragma unroll
    for (int o = 16; o; o >>= 1) { n_stage += __shfl_xor_sync(0xffffffffu, n_stage, o); n_direct += __shfl_xor_sync(0xffffffffu, n_direct, o); }
    if (lane == 0) { if (n_stage) atomicAdd(stats, n_stage); if (n_direct) atomicAdd(stats + 1, n_direct); }
  }
  if (inside) out_rgba[(size_t)py * P.W + px] = make_float4(Cr, Cg, Cb, P.alphaMode ? 1.0f - T : 1.0f);
}

// ------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled tma_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

template <int BOXB, int TILE>
static cudaError_t launch_tma(const KParams& P, const void* vol, const float* tf, const uint8_t* levels, float* out_rgba,
                              unsigned long long* stats, cudaStream_t st) {
  typedef TmaBox<BOXB> BX;
  PFN_encodeTiled enc = tma_encode_fn();
  if (!enc) return cudaErrorNotSupported;
  CUtensorMap map;
  const cuuint64_t gdim[3] = {(cuuint64_t)P.dims[0], (cuuint64_t)P.dims[1], (cuuint64_t)P.dims[2]};
  const cuuint64_t gstr[2] = {(cuuint64_t)P.pitchY * 4, (cuuint64_t)P.pitchZ * 4};
  const cuuint32_t box[3] = {BX::NX, BX::NY, BX::NZ};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(vol), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorInvalidValue;
  const int txn = mrt_tiles_x_(P.W), tyn = mrt_tiles_y_(P.H);
  const int grid = TILE == 8 ? txn * tyn : ((txn + 1) >> 1) * ((tyn + 1) >> 1);
  const size_t smem = 2 * (size_t)BX::STRIDE + (size_t)(P.tfMode ? P.tfN : 2) * sizeof(TfEntry);
  cudaError_t e = cudaFuncSetAttribute(mrt_fwd_tma_kernel<BOXB, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  mrt_fwd_tma_kernel<BOXB, TILE><<<grid, TILE * TILE, smem, st>>>(P, map, (const float*)vol, (const float4*)tf, levels,
                                                                  (float4*)out_rgba, stats);
  return cudaGetLastError();
}

// box_edge: 8 or 16 voxels; tile: 8 or 16 pixels.  Single-channel fp32 scalar layout, one view (P's camera),
// skip levels required, no shards / overlays / gamma.
cudaError_t mrt_launch_forward_tma(const KParams& P, int box_edge, int tile, const void* vol, const float* tf,
                                   const uint8_t* levels, float* out_rgba, void* stats, cudaStream_t st) {
  if (P.half || P.shard || P.tMode != 0 || P.gamma != 1.0f || P.showSeg || P.showPred || !levels) return cudaErrorInvalidValue;
  if (P.nbx > 1023 || P.nby > 1023 || P.nbz > 1023) return cudaErrorInvalidValue;
  unsigned long long* s = (unsigned long long*)stats;
  if (box_edge == 8 && tile == 8) return launch_tma<1, 8>(P, vol, tf, levels, out_rgba, s, st);
  if (box_edge == 8 && tile == 16) return launch_tma<1, 16>(P, vol, tf, levels, out_rgba, s, st);
  if (box_edge == 16 && tile == 8) return launch_tma<2, 8>(P, vol, tf, levels, out_rgba, s, st);
  if (box_edge == 16 && tile == 16) return launch_tma<2, 16>(P, vol, tf, levels, out_rgba, s, st);
  return cudaErrorInvalidValue;
}
