"""Multi-GPU rendering: one process per GPU (torch.distributed), SURVEY.md §8(e).

The reference is single-GPU; these two modes are what BASELINE.json's north_star adds:

* image space: the volume, TF and occupancy grid are replicated; a batch of views is split over
  ranks by interleaved screen-space TILE ROWS of every view (or by whole views), and the framebuffer
  gather is fused into the march: :class:`PeerFramebuffer` owns symmetric (peer-mapped) frames and
  every rank's ONE batched launch stores its pixels straight into the owner GPU of each view over
  NVLink.  Owners are striped over the ranks (or one root).  ``render_views`` is the plain NCCL
  ``all_gather`` version of the same partitions.  Rays are independent: no other exchange.
* sort-last (``render_sort_last`` / :class:`PeerSortLast`): the volume is split into axis-aligned
  sub-boxes (+1 voxel halo so trilinear sampling at the cut faces is exact); each rank marches every
  ray through its own sub-box only, producing premultiplied colour + transmittance; image strips are
  exchanged and composited front-to-back in visibility order (``mrt_composite_over``).

Everything here drives CUDA tensors; the host-side logic (partitions, orders, collectives) is
exercised on CPU with the gloo backend by tests that substitute duck-typed stand-ins for the
renderer objects (tests/test_dist_gloo.py) — there is no oracle hook and no CPU path in this module.
"""
from __future__ import annotations

import warnings
from dataclasses import replace
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import tiles
from .params import RenderParams


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _render_batch(volume, tf, P: RenderParams, cams: Sequence, tile_range: Tuple[int, int], out: torch.Tensor):
    """``len(cams)`` views into the contiguous ``out [V,H,W,4]``: ONE batched launch
    (api.Volume.forward_batch)."""
    if not len(cams):
        return
    volume.forward_batch(replace(P, tfMode=1 if tf is not None else 0), list(cams), tf, out=out, tile_range=tile_range)


# ----------------------------------------------------------------------------- image space
def view_partition(n_views: int, rank: int, nranks: int) -> Tuple[int, int]:
    """Whole-view split (needs n_views % nranks == 0 for the single fused gather)."""
    return (rank * n_views) // nranks, ((rank + 1) * n_views) // nranks


def padded_rows(H: int, nranks: int) -> int:
    """Rows per rank in the gather buffer: whole tile rows, padded so every rank is equal."""
    ty = tiles.tiles_y(H)
    return ((ty + nranks - 1) // nranks) * tiles.TILE


def render_views(volume, cams: Sequence, tf, P: RenderParams, mode: str = "views", group=None,
                 device=None, gather: bool = True) -> torch.Tensor:
    """Render ``len(cams)`` views of one volume across all ranks -> ``[V,H,W,4]`` on every rank
    through ONE NCCL ``all_gather_into_tensor`` (``gather=False``: only this rank's part is valid;
    used to time compute alone).  The peer-memory version is :class:`PeerFramebuffer`."""
    rank, R = _world(group)
    W, H = P.imageSize
    V = len(cams)
    device = device if device is not None else getattr(volume, "device", "cpu")
    nt = tiles.tile_count(W, H)
    if mode == "views":
        if V % R != 0:
            raise ValueError(f"mode='views' needs len(cams) ({V}) divisible by world size ({R})")
        out = torch.empty((V, H, W, 4), dtype=torch.float32, device=device)
        v0, v1 = view_partition(V, rank, R)
        _render_batch(volume, tf, P, cams[v0:v1], (0, nt), out[v0:v1])
        if R > 1 and gather:
            dist.all_gather_into_tensor(out.view(-1), out[v0:v1].reshape(-1), group=group)
        return out
    if mode == "tiles":
        rows = padded_rows(H, R)
        tx = tiles.tiles_x(W)
        # rank r owns tile rows [r*rows/8, (r+1)*rows/8) of every view (clipped to the image)
        buf = torch.zeros((R, V, rows, W, 4), dtype=torch.float32, device=device)
        tr0 = min(rank * (rows // tiles.TILE), tiles.tiles_y(H))
        tr1 = min((rank + 1) * (rows // tiles.TILE), tiles.tiles_y(H))
        y0, y1 = tr0 * tiles.TILE, min(tr1 * tiles.TILE, H)
        if tr1 > tr0:
            full = torch.empty((V, H, W, 4), dtype=torch.float32, device=device)
            _render_batch(volume, tf, P, cams, (tr0 * tx, tr1 * tx), full)
            buf[rank, :, : y1 - y0] = full[:, y0:y1]
        if R > 1 and gather:
            dist.all_gather_into_tensor(buf.view(-1), buf[rank].reshape(-1), group=group)
        # [R,V,rows,W,4] -> [V, R*rows, W, 4] -> crop
        img = buf.permute(1, 0, 2, 3, 4).reshape(V, R * rows, W, 4)[:, :H]
        return img.contiguous()
    raise ValueError(f"unknown mode {mode!r}")


# ----------------------------------------------------------------------------- fused gather
class PeerFramebuffer:
    """Framebuffer of ``n_views`` frames distributed over the ranks, written by the render kernels
    themselves: every rank holds symmetric (peer-mapped) memory for the frames it OWNS, and each
    rank's ONE batched march stores the pixels it renders straight into the owner GPU of their view
    through the NVLink peer mapping (16-byte coalesced stores, fire-and-forget), so the transfer
    overlaps the march of the following rays and no collective sits on the data path — one
    stream-ordered barrier per batch.

    owners="striped" (default): view ``v`` lives on rank ``v // ceil(n_views/R)`` — every GPU
    receives 1/R of the traffic (a single root's NVLink ingress is what bounded the round-1 gather
    at 8 GPUs).  owners="root": all frames on ``root``.

    partition="tiles" (default): rank r renders tile rows ``ty % R == r`` of EVERY view
    (north_star's image-space tile partition: a fixed batch scales strongly, and the load is even
    because every rank gets a slice of every view).  partition="views": rank r renders whole views
    ``[r*V/R, (r+1)*V/R)``.

    Tiles outside a view's *spans* — per tile row the x-extent of the union of the active bricks' projected footprints, a
    deterministic function of (camera, params, occupancy) that every rank computes for itself
    (``mrt_view_spans``) — are not sent; the owner fills them with the background on a side stream
    while everybody marches (``mrt_fill_outside_spans``; disjoint pixels, no ordering needed).

    Double-buffered: batch b is written into buffer ``b % 2``.  ``finish()`` returns this rank's
    owned frames ``[n_owned,H,W,4]``; they stay valid until the ``finish()`` of the NEXT batch is
    called on this rank's stream (peers cannot start batch b+2, which reuses the buffer, before
    every rank has passed the barrier of batch b+1).

    Symmetric memory is required: construction raises when it is unavailable unless
    ``allow_nccl_fallback=True`` (then a warning is issued, every rank renders its part into a full
    local ``[V,H,W,4]`` buffer and ``finish()`` completes it with ONE NCCL ``all_reduce``-free
    ``all_gather``; ``self.p2p`` is False and every rank owns — sees — all frames)."""

    def __init__(self, n_views: int, H: int, W: int, device, group=None, owners: str = "striped", root: int = 0,
                 partition: str = "tiles", allow_nccl_fallback: bool = False):
        self.rank, self.R = _world(group)
        self.group = group if group is not None else (dist.group.WORLD if self.R > 1 else None)
        if owners not in ("striped", "root") or partition not in ("tiles", "views"):
            raise ValueError("owners must be 'striped' or 'root', partition 'tiles' or 'views'")
        self.V, self.H, self.W, self.root, self.owners, self.partition = int(n_views), int(H), int(W), int(root), owners, partition
        R = self.R
        if partition == "views" and self.V % R != 0:
            raise ValueError(f"partition='views' needs n_views ({self.V}) divisible by world size ({R})")
        self.per_owner = (self.V + R - 1) // R if owners == "striped" else self.V
        self.device = torch.device(device)
        self.p2p = False
        self.why = None
        self.batch = 0
        self._fill_done = None
        shape = (2, self.per_owner, H, W, 4)
        if R > 1 and self.device.type == "cuda":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self.local = symm_mem.empty(shape, dtype=torch.float32, device=self.device)
                self.hdl = symm_mem.rendezvous(self.local, self.group)
                peers = [self.hdl.get_buffer(j, shape, torch.float32) for j in range(R)]
                frame = H * W * 4 * 4
                ptrs = torch.empty((2, self.V), dtype=torch.int64)
                for b in range(2):
                    for v in range(self.V):
                        o, slot = self.owner_of(v)
                        ptrs[b, v] = peers[o][b].data_ptr() + slot * frame
                self.view_ptrs = ptrs.to(self.device)
                self.p2p = True
            except Exception as e:                      # depends on the platform (driver, topology, torch build)
                self.why = f"{type(e).__name__}: {e}"
        elif R > 1:
            self.why = f"symmetric memory needs CUDA tensors (device={self.device})"
        if R == 1:
            self.local = torch.empty(shape, dtype=torch.float32, device=self.device)
            if self.device.type == "cuda":
                frame = H * W * 4 * 4
                self.view_ptrs = torch.tensor([[self.local[b].data_ptr() + v * frame for v in range(self.V)] for b in range(2)],
                                              dtype=torch.int64, device=self.device)
                self.p2p = True
        if not self.p2p:
            if not allow_nccl_fallback:
                raise RuntimeError(f"PeerFramebuffer: symmetric (peer-mapped) memory is unavailable — {self.why}; "
                                   "pass allow_nccl_fallback=True to gather with NCCL instead")
            warnings.warn(f"PeerFramebuffer: falling back to an NCCL all_gather ({self.why})", RuntimeWarning, stacklevel=2)
            self.full = torch.zeros((self.V, H, W, 4), dtype=torch.float32, device=self.device)
        ty = tiles.tiles_y(H)
        if self.p2p:
            self.spans = torch.empty((self.V, ty, 2), dtype=torch.int32, device=self.device)
            self.side = torch.cuda.Stream(device=self.device)
            n_own = max(1, len(self.owned_views()))
            self._prev_spans = [torch.empty((n_own, ty, 2), dtype=torch.int32, device=self.device) for _ in range(2)]
            self._prev_key = [None, None]            # (background, alphaMode) the buffer was last filled with

    # ---- integer maps (replicated on every rank)
    def owner_of(self, v: int) -> Tuple[int, int]:
        """(owner rank, slot in the owner's buffer) of view ``v``."""
        if self.owners == "root":
            return self.root, v
        return v // self.per_owner, v % self.per_owner

    def owned_views(self, rank: Optional[int] = None) -> range:
        r = self.rank if rank is None else rank
        if not self.p2p:
            return range(self.V)
        if self.owners == "root":
            return range(self.V) if r == self.root else range(0)
        return range(min(r * self.per_owner, self.V), min((r + 1) * self.per_owner, self.V))

    # ---- one batch
    def render(self, volume, cams: Sequence, tf, P: RenderParams):
        """Queue this rank's share of the batch (``cams`` = the cameras of ALL ``n_views`` views, in
        frame order, identical on every rank).  Call :meth:`finish` afterwards."""
        from . import api
        if len(cams) != self.V:
            raise ValueError(f"the framebuffer holds {self.V} views, got {len(cams)} cameras")
        cams = list(cams)
        W, H = P.imageSize
        if (W, H) != (self.W, self.H):
            raise ValueError("imageSize does not match the framebuffer")
        tfm = 1 if tf is not None else 0
        Pm = P if P.tfMode == tfm else P.derived(("tfmode", tfm), lambda p: replace(p, tfMode=tfm))
        R, r = self.R, self.rank
        if not self.p2p:
            self.full.zero_()
            if self.partition == "views":
                v0, v1 = view_partition(self.V, r, R)
                _render_batch(volume, tf, Pm, cams[v0:v1], (0, tiles.tile_count(W, H)), self.full[v0:v1])
            else:
                tr0, tr1 = tiles.rank_tile_range(tiles.tiles_y(H), r, R)          # contiguous tile rows
                tx = tiles.tiles_x(W)
                _render_batch(volume, tf, Pm, cams, (tr0 * tx, tr1 * tx), self.full)
            return
        Pq = Pm.with_projection_of(cams[0])
        if self.partition == "tiles" and hasattr(volume, "stale_and_fusable") and volume.stale_and_fusable(Pq):
            # stale folded volume (weights changed): fold + occupancy + layout, classify, spans of every view and this
            # rank's tile rows of every view in ONE staged library call (mrt_render_views_refold_scatter)
            buf = self.batch & 1
            volume.refold_and_scatter(Pq, cams, tf, self.view_ptrs[buf], self.spans, R, r)
            self._fill_owned(api, Pq, buf)
            return
        plan = volume.sparse_plan(Pm, cams, tf)
        if plan is None:
            raise RuntimeError("PeerFramebuffer.render needs the span path: an occupancy grid, skipEmpty=1, indexed stepping, "
                               "gamma 1 and no label overlays")
        packed, Cn, Pe, bits = plan
        buf = self.batch & 1
        arr = api._camera_array(cams)                                 # one host-side packing for both launches
        api.view_spans(Pe, arr, Cn, bits, out=self.spans)             # all views: senders and owners need them
        if self.partition == "tiles":
            api.render_forward_batch_scatter(Pe, arr, packed, Cn, tf, bits, self.view_ptrs[buf], self.spans,
                                             store_outside=False, row_mod=R, row_rem=r)
        else:
            v0, v1 = view_partition(self.V, r, R)
            if v1 > v0:
                api.render_forward_batch_scatter(Pe, arr[v0:v1], packed, Cn, tf, bits, self.view_ptrs[buf, v0:v1].contiguous(),
                                                 self.spans[v0:v1], store_outside=False)
        self._fill_owned(api, Pe, buf)

    def _fill_owned(self, api, Pe, buf):
        own = self.owned_views()
        if len(own):
            # background of the owned views outside their spans, on a side stream, while everybody
            # marches (queued after this rank's own march so that its launch is not delayed)
            # Delta fill: this buffer already holds the background outside the spans of the batch that last
            # wrote it (two batches ago), so only the tiles those spans covered and the new ones do not are
            # written — unless the background changed or the buffer is new.
            key = (tuple(float(v) for v in Pe.bgColor), int(Pe.alphaMode))
            prev = self._prev_spans[buf] if self._prev_key[buf] == key else None
            ev = torch.cuda.Event(); ev.record()
            with torch.cuda.stream(self.side):
                self.side.wait_event(ev)
                now = self.spans[own.start:own.stop]
                api.fill_outside_spans(Pe, now, self.local[buf, :len(own)], prev_spans=prev)
                self._prev_spans[buf].copy_(now)
                self._prev_key[buf] = key
                self._fill_done = torch.cuda.Event(); self._fill_done.record()

    def finish(self) -> torch.Tensor:
        """Complete the batch: one barrier (the owner also joins its background fill), or the NCCL
        gather of the fallback.  -> this rank's owned frames ``[n_owned,H,W,4]`` (see class doc for
        how long they stay valid)."""
        if not self.p2p:
            if self.R > 1:
                if self.partition == "views":
                    v0, v1 = view_partition(self.V, self.rank, self.R)
                    dist.all_gather_into_tensor(self.full.view(-1), self.full[v0:v1].reshape(-1).clone(), group=self.group)
                else:
                    dist.all_reduce(self.full, op=dist.ReduceOp.SUM, group=self.group)   # disjoint rows, zeros elsewhere
            return self.full
        buf = self.batch & 1
        if self.R > 1:
            self.hdl.barrier()          # stream-ordered: after this rank's march kernel; all peers' stores have landed
        if self._fill_done is not None:
            torch.cuda.current_stream().wait_event(self._fill_done)
            self._fill_done = None
        self.batch += 1
        return self.local[buf, :len(self.owned_views())]


# ----------------------------------------------------------------------------- differentiable, image space
def render_differentiable(volume: torch.Tensor, cam, tf: Optional[torch.Tensor], P: RenderParams, group=None,
                          **kw) -> torch.Tensor:
    """Differentiable rendering, data-parallel over screen tiles: the
    volume and TF are replicated, rank r renders — and differentiates — tile range
    ``tiles.rank_tile_range(ntiles, r, R)``, and the frame is assembled with ONE differentiable
    ``all_reduce(SUM)`` (the ranks' images are disjoint and zero elsewhere).  Every rank gets the
    whole ``[H,W,4]`` image and may compute any loss on it; after ``loss.backward()`` each rank holds
    the gradient contribution of ITS tiles — finish with :func:`allreduce_gradients` (the volume-
    sized ``all_reduce`` of dL/dvolume and the tiny one of dL/dtf), exactly like data-parallel
    training."""
    from . import api
    rank, R = _world(group)
    Pc = P.with_camera(cam) if cam is not None else P
    W, H = Pc.imageSize
    tr = tiles.rank_tile_range(tiles.tile_count(W, H), rank, R)
    part = api.render(volume, None, tf, Pc, tile_range=tr, **kw)
    if R == 1:
        return part
    import torch.distributed.nn.functional as dfn
    # every rank evaluates the same loss on the whole image, so the upstream gradient of the summed
    # frame arrives R times through the differentiable all_reduce: pre-divide the local part's path
    return _ScaleGrad.apply(dfn.all_reduce(part, op=dist.ReduceOp.SUM, group=group if group is not None else dist.group.WORLD),
                            1.0 / R)


class _ScaleGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s):
        ctx.s = s
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.s, None


def allreduce_gradients(tensors: Sequence[torch.Tensor], group=None):
    """``all_reduce(SUM)`` of the ``.grad`` of every tensor that has one (dL/dvolume, dL/dtf)."""
    _, R = _world(group)
    if R == 1:
        return
    for t in tensors:
        if t is not None and t.grad is not None:
            dist.all_reduce(t.grad, op=dist.ReduceOp.SUM, group=group)


# ----------------------------------------------------------------------------- sort-last
def shard_grid(nranks: int) -> Tuple[int, int, int]:
    """Sub-box grid (gx,gy,gz) with gx*gy*gz == nranks, as cubic as possible (8 -> 2x2x2)."""
    g = [1, 1, 1]
    n = nranks
    ax = 2
    while n > 1:
        p = next(q for q in (2, 3, 5, 7, 11, 13) if n % q == 0) if any(n % q == 0 for q in (2, 3, 5, 7, 11, 13)) else n
        g[ax] *= p
        n //= p
        ax = (ax - 1) % 3
    return g[0], g[1], g[2]


def shard_box(dims: Tuple[int, int, int], grid: Tuple[int, int, int], rank: int):
    """Voxel range of ``rank``'s sub-box: cells [lo, hi) per axis, i.e. voxels [lo, hi] inclusive
    (the +1 halo voxel makes trilinear sampling inside the cell range exact).  Returns
    (lo[3], hi[3]) in (x,y,z) order and the shard's grid coordinate."""
    gx, gy, gz = grid
    cx, cy, cz = rank % gx, (rank // gx) % gy, rank // (gx * gy)
    lo, hi = [], []
    for c, g, d in ((cx, gx, dims[0]), (cy, gy, dims[1]), (cz, gz, dims[2])):
        cells = d - 1                                   # trilinear cells along this axis
        a, b = (c * cells) // g, ((c + 1) * cells) // g
        lo.append(a)
        hi.append(b)
    return tuple(lo), tuple(hi), (cx, cy, cz)


def visibility_order(eye: np.ndarray, P: RenderParams, grid: Tuple[int, int, int]) -> List[int]:
    """Front-to-back order of the sub-boxes for a pinhole eye (or an ortho direction): along
    each axis the slab containing the eye comes first, then slabs by increasing distance; the
    lexicographic combination is a valid visibility order for an axis-aligned grid."""
    dims = P.dims
    vs = np.asarray(P.voxelSize, dtype=np.float64)
    vmin = np.asarray(P.volMin, dtype=np.float64)
    keys = []
    for r in range(grid[0] * grid[1] * grid[2]):
        lo, hi, _ = shard_box(dims, grid, r)
        d = 0.0
        k = []
        for a in range(3):
            if P.ortho:
                # parallel rays along W: order by the slab centre projected on the view direction
                c = vmin[a] + 0.5 * (lo[a] + hi[a]) * vs[a]
                k.append(c * float(np.asarray(P.W, dtype=np.float64)[a]))
            else:
                a0, a1 = vmin[a] + lo[a] * vs[a], vmin[a] + hi[a] * vs[a]
                e = float(eye[a])
                k.append(0.0 if a0 <= e <= a1 else min(abs(e - a0), abs(e - a1)))
        keys.append((sum(k) if P.ortho else 0.0, tuple(k) if not P.ortho else (), r))
    if P.ortho:
        return [r for _, _, r in sorted(keys)]
    # per-axis rank of each slab, then sort boxes by the tuple of per-axis ranks
    return [r for _, _, r in sorted(keys, key=lambda t: (t[1], t[2]))]


def slice_shard(planar: torch.Tensor, lo, hi) -> torch.Tensor:
    """Voxels [lo, hi] inclusive of a ``[C,Z,Y,X]`` volume (cells [lo,hi) + the 1-voxel halo)."""
    return planar[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1].contiguous()


def composite_over(partials: torch.Tensor, order: Sequence[int], bg, alpha_mode: int = 0) -> torch.Tensor:
    """Ordered front-to-back `over` of ``[K,npix,4]`` CUDA partials -> ``[npix,4]`` (``mrt_composite_over``)."""
    if not partials.is_cuda:
        raise RuntimeError("composite_over needs CUDA tensors: the render path has no CPU fallback")
    import ctypes as C
    from ._lib import check, lib
    K, npix = partials.shape[0], partials.shape[1]
    o = torch.tensor(list(order), dtype=torch.int32, device=partials.device)
    bga = np.asarray(bg, dtype=np.float32)
    out = torch.empty((npix, 4), dtype=torch.float32, device=partials.device)
    check(lib().mrt_composite_over(partials.contiguous().data_ptr(), K, o.data_ptr(), npix, bga.ctypes.data,
                                   int(alpha_mode), out.data_ptr(), torch.cuda.current_stream().cuda_stream),
          "composite_over")
    return out


def composite_over_differentiable(partials: torch.Tensor, order: Sequence[int], bg, alpha_mode: int = 0) -> torch.Tensor:
    """The ordered `over` of :func:`composite_over` written in differentiable tensor operations:
    ``[K,npix,4]`` partials (premultiplied rgb, T) -> ``[npix,4]``; C = bg + sum_k (prod_{j<k} T_j) C_k.
    Used by the differentiable sort-last path (the forward-only paths use the fused kernel)."""
    p = partials[list(order)]
    T = p[..., 3]
    Tcum = torch.cumprod(T, dim=0)
    Tprev = torch.cat([torch.ones_like(T[:1]), Tcum[:-1]], dim=0)
    rgb = (Tprev.unsqueeze(-1) * p[..., :3]).sum(dim=0) + torch.as_tensor(bg, dtype=p.dtype, device=p.device)[:3]
    a = (1.0 - Tcum[-1]) if alpha_mode else torch.ones_like(Tcum[-1])
    return torch.cat([rgb, a.unsqueeze(-1)], dim=-1)


class _AllToAllStrips(torch.autograd.Function):
    """``all_to_all_single`` of equal image strips; its adjoint is the same exchange of the gradients."""

    @staticmethod
    def forward(ctx, send, group):
        ctx.group = group
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv.view(-1), send.contiguous().view(-1), group=group)
        return recv

    @staticmethod
    def backward(ctx, g):
        back = torch.empty_like(g)
        dist.all_to_all_single(back.view(-1), g.contiguous().view(-1), group=ctx.group)
        return back, None


def render_sort_last_differentiable(sub: torch.Tensor, cam, tf, P: RenderParams, grid: Tuple[int, int, int],
                                    group=None, storage: Optional[torch.dtype] = None):
    """Differentiable cfg5: this rank renders the partial of ITS sub-box (``sub`` = voxels
    ``shard_box(P.dims, grid, rank)`` inclusive, fp32 or fp16, ``requires_grad`` as wanted) with
    :func:`api.render_shard`, image strips are exchanged by ONE differentiable ``all_to_all_single``
    and this rank composites ITS strip front to back.  Returns ``(strip [rows,W,4], row0)``: rows
    ``row0 .. row0+rows`` of the frame (rows beyond ``H`` are padding).  Build the loss from the
    strip (each pixel belongs to exactly one rank, so the ranks' losses add up to the frame's);
    ``loss.backward()`` then leaves dL/d(sub) on the rank that stores those voxels — halo voxels are
    stored by two shards, each holding its own cells' share — and dL/dtf as this rank's share:
    finish with :func:`allreduce_gradients` on tf."""
    from . import api
    rank, R = _world(group)
    if grid[0] * grid[1] * grid[2] != R:
        raise ValueError(f"grid {grid} does not match world size {R}")
    W, H = P.imageSize
    Pc = P.with_camera(cam) if cam is not None else P
    lo, hi, _ = shard_box(Pc.dims, grid, rank)
    partial = api.render_shard(sub, (lo, hi), Pc.dims, None, tf, Pc, storage=storage)
    order = visibility_order(np.asarray(Pc.eye, dtype=np.float64), Pc, grid)
    rows = padded_rows(H, R)
    if R == 1:
        return composite_over_differentiable(partial.reshape(1, H * W, 4), order, Pc.bgColor, Pc.alphaMode).reshape(H, W, 4), 0
    pad = torch.zeros((R * rows - H, W, 4), dtype=torch.float32, device=partial.device)
    pad[..., 3] = 1.0                                    # padding rows: empty partial (T = 1)
    send = torch.cat([partial, pad], dim=0).reshape(R, rows, W, 4)
    recv = _AllToAllStrips.apply(send, group)            # recv[j] = rank j's partial of MY strip
    mine = composite_over_differentiable(recv.reshape(R, rows * W, 4), order, Pc.bgColor, Pc.alphaMode)
    return mine.reshape(rows, W, 4), rank * rows


def render_sort_last(shard_volume, cam, tf, P: RenderParams, grid: Tuple[int, int, int], group=None) -> torch.Tensor:
    """One frame of a brick-sharded volume (BASELINE config 5) through NCCL: this rank marches every
    ray through its own sub-box only (``shard_volume`` = Volume(..., shard=shard_box(dims, grid,
    rank))) -> partial (premultiplied rgb, T); image strips are exchanged with ONE
    ``all_to_all_single``; each rank composites its strip front-to-back in visibility order and the
    finished strips are all-gathered.  Early termination acts per shard (use a small
    ``ertThreshold``: the result equals the unsharded render to within it).  The peer-memory version
    is :class:`PeerSortLast`."""
    rank, R = _world(group)
    if grid[0] * grid[1] * grid[2] != R:
        raise ValueError(f"grid {grid} does not match world size {R}")
    W, H = P.imageSize
    Pc = P.with_camera(cam) if cam is not None else P
    partial = shard_volume.forward(replace(Pc, tfMode=1 if tf is not None else 0), tf)
    device = partial.device
    order = visibility_order(np.asarray(Pc.eye, dtype=np.float64), Pc, grid)
    if R == 1:
        return composite_over(partial.reshape(1, H * W, 4), order, Pc.bgColor, Pc.alphaMode).reshape(H, W, 4)
    rows = padded_rows(H, R)
    send = torch.zeros((R, rows, W, 4), dtype=torch.float32, device=device)
    send[..., 3] = 1.0                                  # padding rows: empty partial (T = 1)
    send.view(R * rows, W, 4)[:H] = partial
    recv = torch.empty_like(send)                       # recv[j] = rank j's partial of MY strip
    dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
    mine = composite_over(recv.reshape(R, rows * W, 4), order, Pc.bgColor, Pc.alphaMode)
    full = torch.empty((R, rows * W, 4), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(full.view(-1), mine.reshape(-1).contiguous(), group=group)
    return full.reshape(R * rows, W, 4)[:H].contiguous()


class PeerSortLast:
    """Sort-last rendering with BOTH exchanges done by kernel stores over NVLink (symmetric memory):

    1. the march of rank i scatters the image rows of its partial straight into slot i of the
       strip owners' receive buffers (``mrt_render_forward_strips``) — the all-to-all;
    2. one stream-ordered barrier;
    3. rank j composites the R partials of its strip front to back and stores the finished strip
       into EVERY rank's final image (``mrt_composite_over_multi``) — the all-gather;
    4. one barrier (also protects the receive buffers from the next frame's stores).

    No NCCL collective and no staging copy is on the data path.  Symmetric memory is required:
    construction raises when it is unavailable unless ``allow_nccl_fallback=True`` (then, with a
    warning, ``render`` goes through :func:`render_sort_last`; ``self.p2p`` is False).  ``emulate=R`` builds the single-process equivalent (all "ranks" on this
    GPU, plain device buffers) used by the tests."""

    def __init__(self, H: int, W: int, device, group=None, emulate: int = 0, allow_nccl_fallback: bool = False):
        self.rank, self.R = (0, int(emulate)) if emulate else _world(group)
        self.emulate = bool(emulate)
        self.group = group if group is not None else (dist.group.WORLD if (self.R > 1 and not emulate) else None)
        self.H, self.W = H, W
        self.rows = padded_rows(H, self.R)
        R, rows = self.R, self.rows
        self.p2p = False
        self.hdl_recv = self.hdl_final = None
        rshape, fshape = (R, rows, W, 4), (R * rows, W, 4)
        if self.emulate:
            self.recv_all = torch.zeros((R,) + rshape, dtype=torch.float32, device=device)     # [owner][src]
            self.final_all = torch.zeros((R,) + fshape, dtype=torch.float32, device=device)
            self.p2p = True
        elif R > 1 and torch.device(device).type == "cuda":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self.recv = symm_mem.empty(rshape, dtype=torch.float32, device=device)
                self.final = symm_mem.empty(fshape, dtype=torch.float32, device=device)
                self.hdl_recv = symm_mem.rendezvous(self.recv, self.group)
                self.hdl_final = symm_mem.rendezvous(self.final, self.group)
                self.recv_peer = [self.hdl_recv.get_buffer(j, rshape, torch.float32) for j in range(R)]
                self.final_peer = [self.hdl_final.get_buffer(j, fshape, torch.float32) for j in range(R)]
                self.p2p = True
            except Exception as e:                      # depends on the platform (driver, topology, torch build)
                self.why = f"{type(e).__name__}: {e}"
        if not self.p2p and not self.emulate and R > 1:
            if not allow_nccl_fallback:
                raise RuntimeError(f"PeerSortLast: symmetric (peer-mapped) memory is unavailable — {getattr(self, 'why', 'no CUDA device')}; "
                                   "pass allow_nccl_fallback=True to exchange with NCCL (render_sort_last) instead")
            warnings.warn(f"PeerSortLast: falling back to NCCL all_to_all + all_gather ({getattr(self, 'why', 'no CUDA device')})",
                          RuntimeWarning, stacklevel=2)

    # the two halves, rank-parametrised so that the emulation can run them for every "rank"
    def _march(self, rank, shard_volume, tf, Pc):
        from . import api
        Pm = replace(Pc, tfMode=1 if tf is not None else 0)
        packed, Cn, Pe = shard_volume.prepared(Pm)
        bits = shard_volume.skip_levels(Pm, tf)
        if self.emulate:
            ptrs = [self.recv_all[j, rank].data_ptr() for j in range(self.R)]
        else:
            ptrs = [self.recv_peer[j][rank].data_ptr() for j in range(self.R)]
        api.render_forward_strips(Pe, packed, Cn, tf, bits, ptrs, self.rows)

    def _composite(self, rank, Pc, order):
        from . import api
        R, rows, W = self.R, self.rows, self.W
        if self.emulate:
            parts = self.recv_all[rank].view(R, rows * W, 4)
            outs = [self.final_all[j, rank * rows:(rank + 1) * rows].data_ptr() for j in range(R)]
        else:
            parts = self.recv.view(R, rows * W, 4)
            outs = [self.final_peer[j][rank * rows:(rank + 1) * rows].data_ptr() for j in range(R)]
        api.composite_over_multi(parts, order, Pc.bgColor, Pc.alphaMode, outs)

    def render(self, shard_volume, cam, tf, P: RenderParams, grid: Tuple[int, int, int]) -> torch.Tensor:
        """One frame -> ``[H,W,4]`` (complete on every rank).  ``shard_volume`` = this rank's
        Volume(..., shard=shard_box(dims, grid, rank)); in emulation, a list of all ranks' volumes."""
        if grid[0] * grid[1] * grid[2] != self.R:
            raise ValueError(f"grid {grid} does not match world size {self.R}")
        Pc = P.with_camera(cam) if cam is not None else P
        if (Pc.imageSize[0], Pc.imageSize[1]) != (self.W, self.H):
            raise ValueError("imageSize does not match the exchange buffers")
        order = visibility_order(np.asarray(Pc.eye, dtype=np.float64), Pc, grid)
        if self.emulate:
            for r in range(self.R):
                self._march(r, shard_volume[r], tf, Pc)
            for r in range(self.R):
                self._composite(r, Pc, order)
            return self.final_all[0, :self.H]
        if not self.p2p:
            return render_sort_last(shard_volume, None, tf, Pc, grid, group=self.group)
        self._march(self.rank, shard_volume, tf, Pc)
        self.hdl_recv.barrier()
        self._composite(self.rank, Pc, order)
        self.hdl_final.barrier()
        return self.final[:self.H]


def render_sort_last_emulated(planar: torch.Tensor, cam, tf, P: RenderParams, grid: Tuple[int, int, int],
                              fold: bool = True) -> torch.Tensor:
    """All shards of ``grid`` rendered one after the other on ONE GPU, then composited: the
    single-process equivalent of :func:`render_sort_last` (tests, and hosts with fewer GPUs
    than shards)."""
    from . import api
    R = grid[0] * grid[1] * grid[2]
    Z, Y, X = (int(v) for v in planar.shape[1:])
    W, H = P.imageSize
    Pc = P.with_camera(cam) if cam is not None else P
    parts = []
    for r in range(R):
        lo, hi, _ = shard_box((X, Y, Z), grid, r)
        V = api.Volume(slice_shard(planar, lo, hi), shard=(lo, hi), global_dims=(X, Y, Z), fold=fold)
        parts.append(V.forward(replace(Pc, tfMode=1 if tf is not None else 0), tf).reshape(H * W, 4))
    order = visibility_order(np.asarray(Pc.eye, dtype=np.float64), Pc, grid)
    return composite_over(torch.stack(parts), order, Pc.bgColor, Pc.alphaMode).reshape(H, W, 4)
