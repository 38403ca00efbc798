"""Public Python API: volume tensor + camera + transfer function in, RGBA image out.

This is the host-side mirror of the reference's dispatch seam
(``kernel.dispatch(thread_count=[W,H,1], vars={gOutput, gIntensity0..3, gLabels, gPreds,
gParams})``, inr/viewer/brats_viewer.py:431-442) and of the functional form its authors
planned (``module.render(pixel=..., volume=..., params=..., _result=out)``,
docs/Methodology-ROI-Neural-Volumetric-Rendering.md:88-96).  PyTorch only supplies device
memory, streams and autograd plumbing; every kernel is in libmrt.so (include/mrt.h).
There is no CPU fallback: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import replace
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .camera import Camera
from .params import RenderParams, SlabParams
from . import tiles as _tiles


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    # (torch.cuda.current_stream() builds a Stream object through three Python layers: ~3 us a call,
    # several calls per frame)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _need_cuda(t: torch.Tensor, name: str, dtype=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the render path has no CPU fallback")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def packed_channels(Cn: int) -> int:
    return 1 if Cn <= 1 else (2 if Cn == 2 else 4)


# ----------------------------------------------------------------------------- low level
def packed_layout(Cn: int, dims) -> Tuple[int, int]:
    """(pitchY, pitchZ) in voxels of the packed layout (``mrt_packed_layout``)."""
    X, Y, Z = dims
    py, pz = C.c_int64(), C.c_int64()
    lib().mrt_packed_layout(Cn, X, Y, Z, C.byref(py), C.byref(pz))
    return int(py.value), int(pz.value)


def pack_volume(planar: torch.Tensor) -> torch.Tensor:
    """[C,Z,Y,X] fp32 -> packed, channel-interleaved, bank-skewed layout (flat fp32 buffer;
    voxel (x,y,z) at element ``(x + pitchY*y + pitchZ*z) * Cp``, Cp = 1,2,4)."""
    _need_cuda(planar, "volume", torch.float32)
    Cn, Z, Y, X = planar.shape
    nbytes = lib().mrt_packed_volume_bytes(Cn, X, Y, Z)
    if nbytes == 0:
        raise ValueError(f"volume shape {tuple(planar.shape)} unsupported")
    packed = torch.zeros((nbytes // 4,), dtype=torch.float32, device=planar.device)
    check(lib().mrt_pack_volume_f32(planar.data_ptr(), Cn, X, Y, Z, packed.data_ptr(), _stream()), "pack_volume")
    return packed


def pack_volume_f16(planar: torch.Tensor) -> torch.Tensor:
    """[1,Z,Y,X] (or [Z,Y,X]) fp16 -> packed fp16 buffer (``mrt_pack_volume_f16``)."""
    _need_cuda(planar, "volume", torch.float16)
    Z, Y, X = planar.shape[-3:]
    nbytes = lib().mrt_packed_volume_bytes_f16(X, Y, Z)
    packed = torch.empty((nbytes // 2,), dtype=torch.float16, device=planar.device)
    check(lib().mrt_pack_volume_f16(planar.data_ptr(), X, Y, Z, packed.data_ptr(), _stream()), "pack_volume_f16")
    return packed


def pack_volume_u8(planar: torch.Tensor) -> torch.Tensor:
    """[1,Z,Y,X] (or [Z,Y,X]) uint8 -> packed uint8 buffer (``mrt_pack_volume_u8``)."""
    _need_cuda(planar, "volume", torch.uint8)
    Z, Y, X = planar.shape[-3:]
    nbytes = lib().mrt_packed_volume_bytes_u8(X, Y, Z)
    packed = torch.empty((nbytes,), dtype=torch.uint8, device=planar.device)
    check(lib().mrt_pack_volume_u8(planar.data_ptr(), X, Y, Z, packed.data_ptr(), _stream()), "pack_volume_u8")
    return packed


def pack_volume_quad(packed1: torch.Tensor, dims, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """packed single-channel fp32 volume -> the "quad" sampler layout (``mrt_pack_volume_quad``:
    16 B per voxel holding its 2x2 (x,y) neighbourhood; render with ``volDtype=3``)."""
    X, Y, Z = (int(v) for v in dims)
    nbytes = lib().mrt_packed_volume_bytes_quad(X, Y, Z)
    if out is None or out.numel() * out.element_size() != nbytes:
        out = torch.empty((nbytes // 4,), dtype=torch.float32, device=packed1.device)
    check(lib().mrt_pack_volume_quad(packed1.data_ptr(), X, Y, Z, out.data_ptr(), _stream()), "pack_volume_quad")
    return out


def pack_volume_quad_f16(packed_f16: torch.Tensor, dims, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """packed fp16 volume -> the quad layout over fp16 voxels (``mrt_pack_volume_quad_f16``; ``volDtype=4``)."""
    X, Y, Z = (int(v) for v in dims)
    nbytes = lib().mrt_packed_volume_bytes_quad_f16(X, Y, Z)
    if out is None or out.numel() * out.element_size() != nbytes:
        out = torch.empty((nbytes // 2,), dtype=torch.float16, device=packed_f16.device)
    check(lib().mrt_pack_volume_quad_f16(packed_f16.data_ptr(), X, Y, Z, out.data_ptr(), _stream()), "pack_volume_quad_f16")
    return out


def build_occupancy_u8(packed: torch.Tensor, dims) -> torch.Tensor:
    X, Y, Z = dims
    mm = torch.empty((lib().mrt_brick_count(X, Y, Z), 1, 2), dtype=torch.float32, device=packed.device)
    check(lib().mrt_build_occupancy_u8(packed.data_ptr(), X, Y, Z, mm.data_ptr(), _stream()), "build_occupancy_u8")
    return mm


def unpack_volume_f16(packed: torch.Tensor, dims) -> torch.Tensor:
    X, Y, Z = dims
    planar = torch.empty((1, Z, Y, X), dtype=torch.float16, device=packed.device)
    check(lib().mrt_unpack_volume_f16(packed.data_ptr(), X, Y, Z, planar.data_ptr(), _stream()), "unpack_volume_f16")
    return planar


def build_occupancy_f16(packed: torch.Tensor, dims) -> torch.Tensor:
    X, Y, Z = dims
    mm = torch.empty((lib().mrt_brick_count(X, Y, Z), 1, 2), dtype=torch.float32, device=packed.device)
    check(lib().mrt_build_occupancy_f16(packed.data_ptr(), X, Y, Z, mm.data_ptr(), _stream()), "build_occupancy_f16")
    return mm


def unpack_volume(packed: torch.Tensor, Cn: int, dims) -> torch.Tensor:
    X, Y, Z = dims
    planar = torch.empty((Cn, Z, Y, X), dtype=torch.float32, device=packed.device)
    check(lib().mrt_unpack_volume_f32(packed.data_ptr(), Cn, X, Y, Z, planar.data_ptr(), _stream()), "unpack_volume")
    return planar


def build_occupancy(packed: torch.Tensor, Cn: int, dims) -> torch.Tensor:
    """Per-brick (min,max) per packed channel: float32 [nbricks, Cp, 2]."""
    X, Y, Z = dims
    nb = lib().mrt_brick_count(X, Y, Z)
    mm = torch.empty((nb, packed_channels(Cn), 2), dtype=torch.float32, device=packed.device)
    check(lib().mrt_build_occupancy(packed.data_ptr(), Cn, X, Y, Z, mm.data_ptr(), _stream()), "build_occupancy")
    return mm


def build_label_occupancy(labels: torch.Tensor) -> torch.Tensor:
    _need_cuda(labels, "labels", torch.int32)
    Z, Y, X = labels.shape
    nb = lib().mrt_brick_count(X, Y, Z)
    out = torch.empty((nb,), dtype=torch.uint8, device=labels.device)
    check(lib().mrt_build_label_occupancy(labels.data_ptr(), X, Y, Z, out.data_ptr(), _stream()),
          "build_label_occupancy")
    return out


def skip_levels_buffer(P: RenderParams, device) -> torch.Tensor:
    """uint8 buffer for ``mrt_classify_bricks``: one level byte per brick + the active-brick box tail
    (``mrt_skip_levels_bytes``); sized for the (sub-)volume the params describe."""
    if P.shard is not None:
        X, Y, Z = (int(h) - int(l) + 1 for l, h in zip(P.shard[0], P.shard[1]))
    else:
        X, Y, Z = P.dims
    return torch.empty((lib().mrt_skip_levels_bytes(X, Y, Z),), dtype=torch.uint8, device=device)


def classify_bricks(P: RenderParams, minmax: torch.Tensor, Cn: int, tf: Optional[torch.Tensor],
                    seg_any: Optional[torch.Tensor] = None, pred_any: Optional[torch.Tensor] = None,
                    out: Optional[torch.Tensor] = None, flat: bool = False) -> torch.Tensor:
    if out is None:
        out = skip_levels_buffer(P, minmax.device)
    s = P.to_struct()
    check(lib().mrt_classify_bricks(C.byref(s), minmax.data_ptr(), Cn, _ptr(tf), 0 if tf is None else tf.shape[0],
                                    _ptr(seg_any), _ptr(pred_any), out.data_ptr(), int(flat), _stream()),
          "classify_bricks")
    return out


def render_forward(P: RenderParams, packed: torch.Tensor, Cn: int, tf: Optional[torch.Tensor] = None,
                   skip_levels: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                   preds: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                   out_T: Optional[torch.Tensor] = None, out_counts: Optional[torch.Tensor] = None,
                   tile_range: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """Thin wrapper over ``mrt_render_forward`` (no allocation beyond the output image)."""
    W, H = P.imageSize
    if out is None:
        out = torch.empty((H, W, 4), dtype=torch.float32, device=packed.device)
    t0, t1 = tile_range if tile_range is not None else (0, _tiles.tile_count(W, H))
    s = P.to_struct()
    check(lib().mrt_render_forward(C.byref(s), packed.data_ptr(), Cn, _ptr(tf), 0 if tf is None else tf.shape[0],
                                   _ptr(skip_levels), _ptr(labels), _ptr(preds), out.data_ptr(), _ptr(out_T),
                                   _ptr(out_counts), t0, t1, _stream()), "render_forward")
    return out


def render_forward_tma(P: RenderParams, packed: torch.Tensor, tf: Optional[torch.Tensor], skip_levels: torch.Tensor,
                       box_edge: int = 8, tile: int = 8, out: Optional[torch.Tensor] = None,
                       stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The staged-brick variant of the march (``mrt_render_forward_tma``, forward_tma.cu): boxes of
    ``box_edge``^3 voxels streamed through shared memory with 3-D TMA loads, one CTA per
    ``tile``x``tile`` pixels.  Same image as :func:`render_forward` on a scalar single-channel fp32
    ``packed`` volume; measured slower — a selectable variant, not the default.
    ``stats``: optional zeroed int64[4] CUDA tensor (staged slots, direct slots, boxes, overflows)."""
    W, H = P.imageSize
    if out is None:
        out = torch.empty((H, W, 4), dtype=torch.float32, device=packed.device)
    s = P.to_struct()
    check(lib().mrt_render_forward_tma(C.byref(s), packed.data_ptr(), _ptr(tf), 0 if tf is None else tf.shape[0],
                                       skip_levels.data_ptr(), out.data_ptr(), int(box_edge), int(tile), _ptr(stats),
                                       _stream()), "render_forward_tma")
    return out


def _camera_array(cams) -> np.ndarray:
    """``MrtCamera[len(cams)]`` as a float32 ``[V,16]`` array (rows eye|pad, U|pad, V|pad, W|pad); an
    array built by an earlier call passes through (callers that launch several kernels per batch)."""
    if isinstance(cams, np.ndarray):
        return cams
    a = np.zeros((len(cams), 4, 4), dtype=np.float32)
    a[:, :, :3] = np.asarray([(c.eye, c.U, c.V, c.W) for c in cams], dtype=np.float32)
    return a.reshape(len(cams), 16)


def render_forward_batch(P: RenderParams, cams: Sequence, packed: torch.Tensor, Cn: int,
                         tf: Optional[torch.Tensor] = None, skip_levels: Optional[torch.Tensor] = None,
                         labels: Optional[torch.Tensor] = None, preds: Optional[torch.Tensor] = None,
                         out: Optional[torch.Tensor] = None, out_T: Optional[torch.Tensor] = None,
                         out_counts: Optional[torch.Tensor] = None,
                         tile_range: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """Thin wrapper over ``mrt_render_forward_batch``: ``len(cams)`` views of one volume under one
    parameter block in one launch (per 64 views) -> ``[V,H,W,4]``.  ``cams`` are
    :class:`camera.Camera` objects; they must share fov / projection with ``P`` (only
    eye/U/V/W vary inside a batch, like successive frames of the reference's loop)."""
    W, H = P.imageSize
    V = len(cams)
    if V < 1:
        raise ValueError("render_forward_batch needs at least one camera")
    if out is None:
        out = torch.empty((V, H, W, 4), dtype=torch.float32, device=packed.device)
    elif tuple(out.shape) != (V, H, W, 4) or not out.is_contiguous():
        raise ValueError(f"out must be contiguous [V,H,W,4]={(V, H, W, 4)}, got {tuple(out.shape)}")
    t0, t1 = tile_range if tile_range is not None else (0, _tiles.tile_count(W, H))
    s = P.to_struct()
    arr = _camera_array(cams)
    check(lib().mrt_render_forward_batch(C.byref(s), arr.ctypes.data, V, packed.data_ptr(), Cn, _ptr(tf),
                                         0 if tf is None else tf.shape[0], _ptr(skip_levels), _ptr(labels),
                                         _ptr(preds), out.data_ptr(), _ptr(out_T), _ptr(out_counts), t0, t1,
                                         _stream()), "render_forward_batch")
    return out


def view_spans(P: RenderParams, cams: Sequence, Cn: int, skip_levels: torch.Tensor,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``mrt_view_spans``: per view and tile row the pixel span (x0,x1) that contains every ray able
    to reach an active brick -> int32 ``[V, tiles_y, 2]`` on the device."""
    if out is None:
        out = torch.empty((len(cams), _tiles.tiles_y(P.imageSize[1]), 2), dtype=torch.int32, device=skip_levels.device)
    s = P.to_struct()
    arr = _camera_array(cams)
    check(lib().mrt_view_spans(C.byref(s), arr.ctypes.data, len(cams), Cn, skip_levels.data_ptr(), out.data_ptr(),
                               _stream()), "view_spans")
    return out


def render_forward_batch_sparse(P: RenderParams, cams: Sequence, packed: torch.Tensor, Cn: int,
                                tf: Optional[torch.Tensor], skip_levels: torch.Tensor, out_ptr: int,
                                spans: torch.Tensor, store_outside: bool = False):
    """``mrt_render_forward_batch_sparse``: like :func:`render_forward_batch` into the (peer) image at
    ``out_ptr``, except that tiles outside the views' spans are not stored (unless ``store_outside``:
    then the spans only replace the per-ray box test)."""
    s = P.to_struct()
    arr = _camera_array(cams)
    check(lib().mrt_render_forward_batch_sparse(C.byref(s), arr.ctypes.data, len(cams), packed.data_ptr(), Cn,
                                                _ptr(tf), 0 if tf is None else tf.shape[0], skip_levels.data_ptr(),
                                                int(out_ptr), spans.data_ptr(), int(bool(store_outside)), _stream()),
          "render_forward_batch_sparse")


def render_forward_batch_scatter(P: RenderParams, cams: Sequence, packed: torch.Tensor, Cn: int,
                                 tf: Optional[torch.Tensor], skip_levels: torch.Tensor, view_ptrs: torch.Tensor,
                                 spans: torch.Tensor, store_outside: bool = False, row_mod: int = 0, row_rem: int = 0):
    """``mrt_render_forward_batch_scatter``: view ``v`` of the batch is stored to the frame at device
    address ``view_ptrs[v]`` (int64 CUDA tensor; local or peer-mapped), only tile rows
    ``ty % row_mod == row_rem`` are rendered — the image-space tile partition with the framebuffer
    gather fused into the march."""
    if view_ptrs.dtype != torch.int64 or not view_ptrs.is_cuda or view_ptrs.numel() < len(cams):
        raise ValueError("view_ptrs must be an int64 CUDA tensor with one address per view")
    s = P.to_struct()
    arr = _camera_array(cams)
    check(lib().mrt_render_forward_batch_scatter(C.byref(s), arr.ctypes.data, len(cams), packed.data_ptr(), Cn,
                                                 _ptr(tf), 0 if tf is None else tf.shape[0], skip_levels.data_ptr(),
                                                 view_ptrs.data_ptr(), spans.data_ptr(), int(bool(store_outside)),
                                                 int(row_mod), int(row_rem), _stream()),
          "render_forward_batch_scatter")


def fill_outside_spans(P: RenderParams, spans: torch.Tensor, out: torch.Tensor, prev_spans: Optional[torch.Tensor] = None):
    """``mrt_fill_outside_spans(_delta)``: background into every tile outside its row's span; with
    ``prev_spans`` (the spans of the batch that last wrote ``out``) only into the tiles those covered."""
    s = P.to_struct()
    check(lib().mrt_fill_outside_spans_delta(C.byref(s), spans.data_ptr(), _ptr(prev_spans), int(spans.shape[0]), out.data_ptr(),
                                             _stream()), "fill_outside_spans")


def render_forward_strips(P: RenderParams, packed: torch.Tensor, Cn: int, tf: Optional[torch.Tensor],
                          skip_levels: Optional[torch.Tensor], strip_ptrs: Sequence[int], strip_rows: int,
                          tile_range: Optional[Tuple[int, int]] = None):
    """``mrt_render_forward_strips``: one frame whose image rows are scattered to per-strip buffers
    (``strip_ptrs`` = device addresses, e.g. peer-mapped memory of the strips' owner GPUs)."""
    W, H = P.imageSize
    t0, t1 = tile_range if tile_range is not None else (0, _tiles.tile_count(W, H))
    arr = (C.c_void_p * len(strip_ptrs))(*[int(p) for p in strip_ptrs])
    s = P.to_struct()
    check(lib().mrt_render_forward_strips(C.byref(s), packed.data_ptr(), Cn, _ptr(tf), 0 if tf is None else tf.shape[0],
                                          _ptr(skip_levels), C.cast(arr, C.c_void_p), len(strip_ptrs), int(strip_rows),
                                          t0, t1, _stream()), "render_forward_strips")


def composite_over_multi(partials: torch.Tensor, order: Sequence[int], bg, alpha_mode: int, out_ptrs: Sequence[int]):
    """``mrt_composite_over_multi``: ordered `over` of ``[K,npix,4]`` partials stored to every address
    in ``out_ptrs`` (local and/or peer-mapped ``[npix,4]`` images)."""
    _need_cuda(partials, "partials", torch.float32)
    K, npix = int(partials.shape[0]), int(partials.shape[1])
    o = torch.tensor(list(order), dtype=torch.int32, device=partials.device)
    bga = np.asarray(bg, dtype=np.float32)
    arr = (C.c_void_p * len(out_ptrs))(*[int(p) for p in out_ptrs])
    check(lib().mrt_composite_over_multi(partials.data_ptr(), K, o.data_ptr(), npix, bga.ctypes.data, int(alpha_mode),
                                         C.cast(arr, C.c_void_p), len(out_ptrs), _stream()), "composite_over_multi")


class Checkpoints:
    """What the training forward (``mrt_render_forward_ckpt``) records for the segment-parallel
    backward: ``ck [nseg-1,V,H,W,4]`` (colour, transmittance before every slot ``c*seg_slots``),
    ``k_end [V,H,W]`` (every ray's end slot) and ``warp_kmax [V, half tiles]``."""
    __slots__ = ("ck", "seg_slots", "nseg", "k_end", "warp_kmax")

    def __init__(self, ck, seg_slots, nseg, k_end, warp_kmax):
        self.ck, self.seg_slots, self.nseg, self.k_end, self.warp_kmax = ck, seg_slots, nseg, k_end, warp_kmax


def checkpoint_plan(P: RenderParams, seg_slots: int = 0) -> Tuple[int, int]:
    """(slots per segment, segments) for the geometry in ``P`` (``mrt_checkpoint_plan``)."""
    s = P.to_struct()
    S, n = C.c_int32(), C.c_int32()
    check(lib().mrt_checkpoint_plan(C.byref(s), int(seg_slots), C.byref(S), C.byref(n)), "checkpoint_plan")
    return int(S.value), int(n.value)


def render_forward_ckpt(P: RenderParams, cams: Optional[Sequence], packed: torch.Tensor, Cn: int,
                        tf: Optional[torch.Tensor] = None, skip_levels: Optional[torch.Tensor] = None,
                        labels: Optional[torch.Tensor] = None, preds: Optional[torch.Tensor] = None,
                        out: Optional[torch.Tensor] = None, tile_range: Optional[Tuple[int, int]] = None,
                        seg_slots: int = 0):
    """``mrt_render_forward_ckpt``: the forward of differentiable rendering -> (image ``[V,H,W,4]``
    — ``[H,W,4]`` when ``cams`` is None — and its :class:`Checkpoints`)."""
    W, H = P.imageSize
    V = 1 if cams is None else len(cams)
    dev = packed.device
    S, nseg = checkpoint_plan(P, seg_slots)
    if out is None:
        out = torch.empty((V, H, W, 4) if cams is not None else (H, W, 4), dtype=torch.float32, device=dev)
    ck = torch.empty((max(nseg - 1, 0), V, H, W, 4), dtype=torch.float32, device=dev)
    k_end = torch.empty((V, H, W), dtype=torch.int32, device=dev)
    kmax = torch.empty((V, lib().mrt_half_tile_count(W, H)), dtype=torch.int32, device=dev)
    t0, t1 = tile_range if tile_range is not None else (0, _tiles.tile_count(W, H))
    s = P.to_struct()
    arr = None if cams is None else _camera_array(cams)
    check(lib().mrt_render_forward_ckpt(C.byref(s), None if arr is None else arr.ctypes.data, V, packed.data_ptr(), Cn,
                                        _ptr(tf), 0 if tf is None else tf.shape[0], _ptr(skip_levels), _ptr(labels),
                                        _ptr(preds), out.data_ptr(), ck.data_ptr() if nseg > 1 else None, S, nseg,
                                        k_end.data_ptr(), kmax.data_ptr(), t0, t1, _stream()), "render_forward_ckpt")
    return out, Checkpoints(ck, S, nseg, k_end, kmax)


def render_backward(P: RenderParams, packed: torch.Tensor, Cn: int, tf: Optional[torch.Tensor],
                    labels: Optional[torch.Tensor], preds: Optional[torch.Tensor], out_rgba: torch.Tensor,
                    dL_dout: torch.Tensor, want_dvol: bool = True, want_dtf: bool = True,
                    tile_range: Optional[Tuple[int, int]] = None, flat_levels: Optional[torch.Tensor] = None,
                    minmax: Optional[torch.Tensor] = None, want_dray: bool = False,
                    cams: Optional[Sequence] = None, ckpt: Optional[Checkpoints] = None,
                    stats: Optional[torch.Tensor] = None):
    """``mrt_render_backward`` -> (dL/dvolume packed or None, dL/dtf [N,4] or None), plus dL/d(o,d)
    ``[H,W,6]`` (``[V,H,W,6]`` with ``cams``) as a third element when ``want_dray``.  ``ckpt`` (from
    :func:`render_forward_ckpt` with the same arguments) selects the segment-parallel path;
    ``stats``: optional int64[2] device tensor accumulating (sample slots shaded, warp tasks)."""
    W, H = P.imageSize
    V = 1 if cams is None else len(cams)
    t0, t1 = tile_range if tile_range is not None else (0, _tiles.tile_count(W, H))
    # (fp16 storage: the gradient is fp32 with the fp16 layout's element pitches)
    dvol = torch.zeros(packed.shape, dtype=torch.float32, device=packed.device) if want_dvol else None
    ntf = tf.shape[0] if (tf is not None and P.tfMode) else 2
    dtf = torch.zeros((ntf, 4), dtype=torch.float32, device=packed.device) if want_dtf else None
    nseg = ckpt.nseg if ckpt is not None else 1
    scratch = torch.empty((lib().mrt_backward_scratch_bytes(W, H, V, ntf, nseg) // 4,), dtype=torch.float32,
                          device=packed.device)
    dray = None
    if want_dray:
        dray = torch.zeros((V, H, W, 6) if cams is not None else (H, W, 6), dtype=torch.float32, device=packed.device)
    s = P.to_struct()
    arr = None if cams is None else _camera_array(cams)
    ck = ckpt
    check(lib().mrt_render_backward(C.byref(s), None if arr is None else arr.ctypes.data, V, packed.data_ptr(), Cn,
                                    _ptr(tf), 0 if tf is None else tf.shape[0],
                                    _ptr(flat_levels), _ptr(minmax), _ptr(labels), _ptr(preds),
                                    out_rgba.data_ptr(), dL_dout.data_ptr(),
                                    None if (ck is None or ck.nseg <= 1) else ck.ck.data_ptr(),
                                    0 if ck is None else ck.seg_slots, 0 if ck is None else ck.nseg,
                                    None if ck is None else ck.k_end.data_ptr(),
                                    None if ck is None else ck.warp_kmax.data_ptr(),
                                    _ptr(dvol), _ptr(dtf), scratch.data_ptr(), _ptr(dray), _ptr(stats), t0, t1, _stream()),
          "render_backward")
    return (dvol, dtf, dray) if want_dray else (dvol, dtf)


def ray_gradients(volume: "Volume", camera: Optional[Camera], tf: Optional[torch.Tensor], params: RenderParams,
                  dL_dout: torch.Tensor) -> torch.Tensor:
    """dL/d(ray origin), dL/d(ray direction) per pixel, ``[H,W,6]`` in world units, for an upstream
    image gradient ``dL_dout [H,W,4]``: docs/DifferentiableRendering.md section 9 (:172-188) with the
    sample times held fixed — dL/do = sum_i dL/dx_i, dL/dd = sum_i t_i dL/dx_i — the entry point of
    camera / pose optimisation (for the pinhole camera dL/d eye = the sum of dL/do over the rays)."""
    if not isinstance(volume, Volume):
        raise TypeError("ray_gradients needs a prepared Volume")
    P = params if camera is None else params.with_camera(camera)
    P = replace(P, tfMode=1 if tf is not None else 0)
    _need_cuda(dL_dout, "dL_dout", torch.float32)
    packed, Cn, Pe = volume.prepared(P, quad=False)             # the backward reads the scalar layout
    out = render_forward(Pe, packed, Cn, tf, volume._classify(P, Pe, Cn, tf), volume.labels, volume.preds)
    _, _, dray = render_backward(Pe, packed, Cn, tf, volume.labels, volume.preds, out, dL_dout.contiguous(),
                                 want_dvol=False, want_dtf=False, want_dray=True)
    return dray


# ----------------------------------------------------------------------------- fold
def fold_volume(planar: torch.Tensor, P: RenderParams) -> torch.Tensor:
    """Blend the modalities once per voxel (``mrt_fold_volume_f32``): [C,Z,Y,X] -> packed C=1."""
    _need_cuda(planar, "volume", torch.float32)
    Cn, Z, Y, X = planar.shape
    nbytes = lib().mrt_packed_volume_bytes(1, X, Y, Z)
    folded = torch.zeros((nbytes // 4,), dtype=torch.float32, device=planar.device)
    s = replace(P, dims=(X, Y, Z), shard=None).to_struct()
    check(lib().mrt_fold_volume_f32(C.byref(s), planar.data_ptr(), Cn, folded.data_ptr(), _stream()), "fold_volume")
    return folded


def fold_volume_occupancy(planar: torch.Tensor, P: RenderParams) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fold + occupancy grid of the folded field in one pass (``mrt_fold_volume_occupancy_f32``):
    -> (packed C=1 volume, minmax float32 [nbricks,1,2])."""
    _need_cuda(planar, "volume", torch.float32)
    Cn, Z, Y, X = planar.shape
    nbytes = lib().mrt_packed_volume_bytes(1, X, Y, Z)
    folded = torch.empty((nbytes // 4,), dtype=torch.float32, device=planar.device)   # padding is never read
    mm = torch.empty((lib().mrt_brick_count(X, Y, Z), 1, 2), dtype=torch.float32, device=planar.device)
    s = (P if (tuple(P.dims) == (X, Y, Z) and P.shard is None) else replace(P, dims=(X, Y, Z), shard=None)).to_struct()
    check(lib().mrt_fold_volume_occupancy_f32(C.byref(s), planar.data_ptr(), Cn, folded.data_ptr(), mm.data_ptr(),
                                              _stream()), "fold_volume_occupancy")
    return folded, mm


def fold_volume_occupancy_quad(planar: torch.Tensor, P: RenderParams, quad: Optional[torch.Tensor] = None,
                               minmax: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fold + occupancy + the march's quad layout in ONE pass (``mrt_fold_volume_occupancy_quad_f32``):
    -> (quad buffer for ``volDtype=3``, minmax float32 [nbricks,1,2]); no scalar folded volume is written.
    ``quad`` / ``minmax``: buffers of an earlier call to reuse."""
    _need_cuda(planar, "volume", torch.float32)
    Cn, Z, Y, X = planar.shape
    nbytes = lib().mrt_packed_volume_bytes_quad(X, Y, Z)
    if quad is None or quad.numel() * quad.element_size() != nbytes:
        quad = torch.empty((nbytes // 4,), dtype=torch.float32, device=planar.device)
    if minmax is None:
        minmax = torch.empty((lib().mrt_brick_count(X, Y, Z), 1, 2), dtype=torch.float32, device=planar.device)
    s = (P if (tuple(P.dims) == (X, Y, Z) and P.shard is None) else replace(P, dims=(X, Y, Z), shard=None)).to_struct()
    check(lib().mrt_fold_volume_occupancy_quad_f32(C.byref(s), planar.data_ptr(), Cn, None, quad.data_ptr(), minmax.data_ptr(),
                                                   _stream()), "fold_volume_occupancy_quad")
    return quad, minmax


def unfold_grad(dfolded: torch.Tensor, P: RenderParams, Cn: int) -> torch.Tensor:
    X, Y, Z = P.dims
    out = torch.empty((Cn, Z, Y, X), dtype=torch.float32, device=dfolded.device)
    s = P.to_struct()
    check(lib().mrt_unfold_grad_f32(C.byref(s), dfolded.data_ptr(), Cn, out.data_ptr(), _stream()), "unfold_grad")
    return out


def folded_params(P: RenderParams) -> RenderParams:
    """Params for rendering a folded (pre-blended) volume as a single modality."""
    return replace(P, volEnabled=(1, 0, 0, 0), volWeight=(1.0, 1.0, 1.0, 1.0))


def _fold_key(P: RenderParams, Cn: int):
    return (tuple(int(bool(e)) for e in P.volEnabled[:Cn]), struct.pack(f"<{Cn}f", *P.volWeight[:Cn]))


# ----------------------------------------------------------------------------- Volume
class Volume:
    """A device-resident volume prepared for rendering: sampler layout, occupancy brick grid,
    optional label volumes, and the reference's world scaling
    (inr/viewer/brats_viewer.py:204-210: voxelSize = zooms*1.8/max_dim, volMin = -extent/2).

    ``fold=True`` (default for C > 1): the modality blend is evaluated once per voxel for the
    current (volEnabled, volWeight) and cached, so the march gathers 8 scalars per sample; the
    cache is rebuilt when the weights change.  ``fold=False`` keeps the channel-interleaved
    layout (one float4 gather per corner) and blends per sample like the reference shader."""

    def __init__(self, planar: torch.Tensor, labels: Optional[torch.Tensor] = None,
                 preds: Optional[torch.Tensor] = None, zooms=(1.0, 1.0, 1.0), occupancy: bool = True,
                 fold: bool = True, shard=None, global_dims=None, quad: Optional[bool] = None):
        """``shard=((lox,loy,loz),(hix,hiy,hiz))`` + ``global_dims``: ``planar`` holds only voxels
        [lo, hi] (inclusive) of a larger volume — a sort-last sub-box (dist.render_sort_last).
        ``quad``: sample from the 16 B/voxel quad layout (``pack_volume_quad``; two 16-byte loads per
        sample instead of eight scalar ones, bit-identical images).  Default: on whenever the sampler
        is single-channel fp32 without label overlays; the layout is rebuilt with the fold."""
        self.half = isinstance(planar, torch.Tensor) and planar.dtype == torch.float16
        self.u8 = isinstance(planar, torch.Tensor) and planar.dtype == torch.uint8
        _need_cuda(planar, "volume", torch.float16 if self.half else (torch.uint8 if self.u8 else torch.float32))
        if planar.dim() != 4 or not (1 <= planar.shape[0] <= 4):
            raise ValueError(f"volume must be [C,Z,Y,X] with C in 1..4, got {tuple(planar.shape)}")
        self.C = int(planar.shape[0])
        if (self.half or self.u8) and (self.C != 1 or labels is not None or preds is not None):
            raise ValueError("fp16 / u8 volumes are single-channel and take no label overlays")
        Z, Y, X = (int(v) for v in planar.shape[1:])
        self.dims = (X, Y, Z)
        self.shard = None
        self.global_dims = self.dims
        if shard is not None:
            lo, hi = tuple(int(v) for v in shard[0]), tuple(int(v) for v in shard[1])
            if tuple(h - l + 1 for l, h in zip(lo, hi)) != self.dims:
                raise ValueError(f"shard {shard} does not match the sub-volume dims {self.dims}")
            if labels is not None or preds is not None:
                raise ValueError("label overlays are not supported on sharded volumes")
            self.shard, self.global_dims = (lo, hi), tuple(int(v) for v in global_dims)
        self.device = planar.device
        self.occupancy = occupancy
        self.fold = bool(fold) and self.C > 1
        self.planar = planar if self.fold else None
        self._key = None
        if self.fold:
            self.packed = self.minmax = None
        elif self.half:
            self.packed = pack_volume_f16(planar)
            self.minmax = build_occupancy_f16(self.packed, self.dims) if occupancy else None
        elif self.u8:
            self.packed = pack_volume_u8(planar)
            self.minmax = build_occupancy_u8(self.packed, self.dims) if occupancy else None
        else:
            self.packed = pack_volume(planar)
            self.minmax = build_occupancy(self.packed, self.C, self.dims) if occupancy else None
        self.labels = self.preds = self.seg_any = self.pred_any = None
        self.set_labels(labels)
        self.set_preds(preds)
        single = (self.fold or self.C == 1) and not self.u8
        if quad and not single:
            raise ValueError("the quad layout needs a single-channel fp32 / fp16 sampler (C == 1 or fold=True)")
        self.quad = single if quad is None else bool(quad)
        self._quad_buf = None
        self._quad_ok = False
        from .synth import world_box
        self.voxel_size, self.vol_min = world_box(self.global_dims, zooms)
        self._bits = None
        self._spans = None

    def prepared(self, P: RenderParams, quad: Optional[bool] = None):
        """-> (sampler buffer, channel count, params) to hand to ``render_forward``.  ``quad=False``
        forces the scalar layout (the backward kernels read that one)."""
        if self.shard is not None:
            shard = self.shard
            P = P.derived(("shard", shard), lambda p: replace(p, shard=shard))
        if self.u8 or (self.half and not self.quad):
            vd = 1 if self.half else 2
            P = P.derived(("voldtype", vd), lambda p: replace(p, volDtype=vd))
        overlays = (self.labels is not None and P.showSeg) or (self.preds is not None and P.showPred)
        want_quad = (self.quad if quad is None else (quad and self.quad)) and not overlays
        if self.fold:
            key = _fold_key(P, self.C)
            if key != self._key:                       # weights changed: every cached layout is stale
                self.packed, self._quad_ok, self._key = None, False, key
            Pf = P.derived("folded", folded_params)    # how the folded field is rendered; P still holds the blend weights
            if want_quad and self.occupancy:           # one pass: fold + occupancy + quad layout, no scalar copy
                if not self._quad_ok:
                    self._quad_buf, self.minmax = fold_volume_occupancy_quad(self.planar, P, self._quad_buf, self.minmax)
                    self._quad_ok = True
                return self._quad_buf, 1, Pf.derived("quad", lambda p: replace(p, volDtype=3))
            if self.packed is None:
                if self.occupancy:
                    self.packed, self.minmax = fold_volume_occupancy(self.planar, P)
                else:
                    self.packed, self.minmax = fold_volume(self.planar, P), None
            P = Pf
        if want_quad:
            if not self._quad_ok:
                self._quad_buf = (pack_volume_quad_f16 if self.half else pack_volume_quad)(self.packed, self.dims, out=self._quad_buf)
                self._quad_ok = True
            vq = 4 if self.half else 3
            return self._quad_buf, 1, P.derived(("quad", vq), lambda p: replace(p, volDtype=vq))
        if self.half and P.volDtype != 1:
            P = P.derived(("voldtype", 1), lambda p: replace(p, volDtype=1))
        return self.packed, (1 if self.fold else self.C), P

    def invalidate(self):
        """Drop the folded-volume cache (the next frame re-folds and rebuilds the occupancy grid)."""
        if self.fold:
            self._key = None

    def set_labels(self, labels: Optional[torch.Tensor]):
        """gLabels (inr/viewer/brats_viewer.py:233-237)."""
        self.labels = self._check_labels(labels)
        self.seg_any = build_label_occupancy(self.labels) if (self.labels is not None and self.occupancy) else None

    def set_preds(self, preds: Optional[torch.Tensor]):
        """gPreds (inr/viewer/brats_viewer.py:293-299)."""
        self.preds = self._check_labels(preds)
        self.pred_any = build_label_occupancy(self.preds) if (self.preds is not None and self.occupancy) else None

    def _check_labels(self, lab):
        if lab is None:
            return None
        _need_cuda(lab, "labels", torch.int32)
        X, Y, Z = self.dims
        if tuple(lab.shape) != (Z, Y, X):
            raise ValueError(f"labels must be [Z,Y,X]={Z, Y, X}, got {tuple(lab.shape)}")
        return lab

    def frame_params(self, P: RenderParams) -> RenderParams:
        """Fill dims / voxelSize / volMin from the volume."""
        return replace(P, dims=self.global_dims, voxelSize=tuple(float(v) for v in self.voxel_size),
                       volMin=tuple(float(v) for v in self.vol_min))

    def frame_camera(self, cam):
        """``frame_volume`` (inr/viewer/brats_viewer.py:320-324): target = centre, radius = 0.8*|extent|."""
        ext = self.voxel_size * np.asarray(self.global_dims, dtype=np.float32)
        cam.target = (self.vol_min + 0.5 * ext).astype(np.float32)
        cam.radius = float(np.linalg.norm(ext) * 0.8)
        return cam

    def skip_levels(self, P: RenderParams, tf: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """Per-frame skip-level byte per brick (``mrt_classify_bricks``), or None if skipping is off."""
        packed, Cn, Pe = self.prepared(P)
        return self._classify(P, Pe, Cn, tf)

    def _classify(self, P: RenderParams, Pe: RenderParams, Cn: int, tf: Optional[torch.Tensor],
                  own_labels: bool = True, own_preds: bool = True):
        """``own_labels`` / ``own_preds`` = False: the frame overlays a label volume OTHER than the
        Volume's own, so its per-brick label occupancy does not apply — classification then keeps
        every brick active for that overlay (exact, just slower) instead of skipping by stale flags."""
        if self.minmax is None or not P.skipEmpty or P.tMode != "indexed":
            return None
        if self._bits is None:
            self._bits = skip_levels_buffer(Pe, self.device)
        return classify_bricks(Pe, self.minmax, Cn, tf, self.seg_any if own_labels else None,
                               self.pred_any if own_preds else None, out=self._bits)

    def forward(self, P: RenderParams, tf: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
                out_T: Optional[torch.Tensor] = None, out_counts: Optional[torch.Tensor] = None,
                tile_range: Optional[Tuple[int, int]] = None, labels=None, preds=None) -> torch.Tensor:
        """classify + march for one frame (two launches); ``P.tfMode`` must already be set."""
        packed, Cn, Pe = self.prepared(P)
        bits = self._classify(P, Pe, Cn, tf, own_labels=labels is None or labels is self.labels,
                              own_preds=preds is None or preds is self.preds)
        return render_forward(Pe, packed, Cn, tf, bits, labels if labels is not None else self.labels,
                              preds if preds is not None else self.preds, out=out, out_T=out_T,
                              out_counts=out_counts, tile_range=tile_range)


    def forward_batch(self, P: RenderParams, cams: Sequence, tf: Optional[torch.Tensor],
                      out: Optional[torch.Tensor] = None, out_T: Optional[torch.Tensor] = None,
                      out_counts: Optional[torch.Tensor] = None,
                      tile_range: Optional[Tuple[int, int]] = None, march_events=None) -> torch.Tensor:
        """classify ONCE (the skip levels do not depend on the camera) + one batched march.  When the
        folded volume is stale (weights changed / :meth:`invalidate`) and the frame is a plain one, the
        whole step — fold + occupancy + layout, classify, spans, march — is ONE library call
        (``mrt_render_views_refold``).  ``march_events``: optional pair of recorded-once
        ``torch.cuda.Event(enable_timing=True)`` that bracket the march on the device."""
        P = P.with_projection_of(cams[0])              # projection (fov / ortho window) of the batch
        if tile_range is None and out_T is None and out_counts is None and self.stale_and_fusable(P):
            return self._refold_and_march(P, cams, tf, out, march_events)
        packed, Cn, Pe = self.prepared(P)
        bits = self._classify(P, Pe, Cn, tf)
        return self.march_batch(Pe, cams, packed, Cn, tf, bits, out=out, out_T=out_T, out_counts=out_counts,
                                tile_range=tile_range)

    def stale_and_fusable(self, P: RenderParams) -> bool:
        """True when the folded volume is stale for ``P`` (weights changed / :meth:`invalidate`) and the
        frame is a plain one: the step can then be queued by ``mrt_render_views_refold(_scatter)``."""
        return (self.fold and self.quad and self.occupancy and self.shard is None and bool(P.skipEmpty)
                and P.tMode == "indexed" and P.gamma == 1.0 and _fold_key(P, self.C) != self._key
                and not (self.labels is not None and P.showSeg) and not (self.preds is not None and P.showPred))

    def _refold_stage1(self, P: RenderParams):
        """Queue the fold pass (it needs no cameras) -> (params struct, stream); the caller prepares its
        camera array while the GPU folds, then queues stage 2."""
        X, Y, Z = self.dims
        dev = self.device
        nbytes = lib().mrt_packed_volume_bytes_quad(X, Y, Z)
        if self._quad_buf is None or self._quad_buf.numel() * self._quad_buf.element_size() != nbytes:
            self._quad_buf = torch.empty((nbytes // 4,), dtype=torch.float32, device=dev)
        if self.minmax is None:
            self.minmax = torch.empty((lib().mrt_brick_count(X, Y, Z), 1, 2), dtype=torch.float32, device=dev)
        s = P.to_struct()
        stream = _stream()
        check(lib().mrt_render_views_refold(C.byref(s), None, 0, self.planar.data_ptr(), self.C, self._quad_buf.data_ptr(),
                                            self.minmax.data_ptr(), None, None, None, 0, None, None, None, 1, stream),
              "render_views_refold")
        self.packed, self._quad_ok, self._key = None, True, _fold_key(P, self.C)
        if self._bits is None:
            self._bits = skip_levels_buffer(P, dev)
        return s, stream

    def refold_and_scatter(self, P: RenderParams, cams: Sequence, tf: Optional[torch.Tensor], view_ptrs: torch.Tensor,
                           spans: torch.Tensor, row_mod: int, row_rem: int):
        """``mrt_render_views_refold_scatter``: the stale-volume step of the distributed framebuffer
        (``dist.PeerFramebuffer``) in one staged library call.  ``P`` must satisfy :meth:`stale_and_fusable`."""
        s, stream = self._refold_stage1(P)
        arr = _camera_array(cams)
        check(lib().mrt_render_views_refold_scatter(C.byref(s), arr.ctypes.data, len(cams), self.planar.data_ptr(), self.C,
                                                    self._quad_buf.data_ptr(), self.minmax.data_ptr(), self._bits.data_ptr(),
                                                    spans.data_ptr(), _ptr(tf), 0 if tf is None else tf.shape[0],
                                                    view_ptrs.data_ptr(), int(row_mod), int(row_rem), 2, stream),
              "render_views_refold_scatter")

    def _refold_and_march(self, P: RenderParams, cams: Sequence, tf: Optional[torch.Tensor],
                          out: Optional[torch.Tensor], march_events=None) -> torch.Tensor:
        """``mrt_render_views_refold``: the stale-volume step in one (staged) call; leaves every cache of the
        Volume (quad layout, min/max, skip levels) as the separate calls would."""
        W, H = P.imageSize
        V = len(cams)
        dev = self.device
        # stage 1 is queued first: the camera array and the output checks below are prepared while the GPU folds
        s, stream = self._refold_stage1(P)
        ty = _tiles.tiles_y(H)
        if self._spans is None or self._spans.shape[0] < V or self._spans.shape[1] != ty:
            self._spans = torch.empty((V, ty, 2), dtype=torch.int32, device=dev)
        if out is None:
            out = torch.empty((V, H, W, 4), dtype=torch.float32, device=dev)
        elif tuple(out.shape) != (V, H, W, 4) or not out.is_contiguous():
            raise ValueError(f"out must be contiguous [V,H,W,4]={(V, H, W, 4)}, got {tuple(out.shape)}")
        if V > lib().mrt_max_views_per_launch() and march_events is not None:
            raise ValueError("march_events bracket a single launch: at most mrt_max_views_per_launch views")
        arr = _camera_array(cams)
        e0 = e1 = None
        if march_events is not None:
            e0, e1 = (C.c_void_p(int(e.cuda_event)) for e in march_events)
        check(lib().mrt_render_views_refold(C.byref(s), arr.ctypes.data, V, self.planar.data_ptr(), self.C,
                                            self._quad_buf.data_ptr(), self.minmax.data_ptr(), self._bits.data_ptr(),
                                            self._spans.data_ptr(), _ptr(tf), 0 if tf is None else tf.shape[0], out.data_ptr(),
                                            e0, e1, 2, stream), "render_views_refold")
        return out

    def march_batch(self, Pe: RenderParams, cams: Sequence, packed: torch.Tensor, Cn: int,
                    tf: Optional[torch.Tensor], bits: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
                    out_T: Optional[torch.Tensor] = None, out_counts: Optional[torch.Tensor] = None,
                    tile_range: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        """The march of a prepared batch.  Whole frames without overlays / counters take the span path:
        ``mrt_view_spans`` (a few microseconds) turns the per-ray box test into one load per warp."""
        W, H = Pe.imageSize
        V = len(cams)
        plain = (bits is not None and tile_range is None and out_T is None and out_counts is None and Pe.gamma == 1.0
                 and not (self.labels is not None and Pe.showSeg) and not (self.preds is not None and Pe.showPred))
        if not plain:
            return render_forward_batch(Pe, cams, packed, Cn, tf, bits, self.labels, self.preds, out=out,
                                        out_T=out_T, out_counts=out_counts, tile_range=tile_range)
        if out is None:
            out = torch.empty((V, H, W, 4), dtype=torch.float32, device=packed.device)
        elif tuple(out.shape) != (V, H, W, 4) or not out.is_contiguous():
            raise ValueError(f"out must be contiguous [V,H,W,4]={(V, H, W, 4)}, got {tuple(out.shape)}")
        ty = _tiles.tiles_y(H)
        if self._spans is None or self._spans.shape[0] < V or self._spans.shape[1] != ty:
            self._spans = torch.empty((V, ty, 2), dtype=torch.int32, device=packed.device)
        # store_outside: the call computes the spans into the scratch itself, then marches
        render_forward_batch_sparse(Pe, cams, packed, Cn, tf, bits, out.data_ptr(), self._spans[:V], store_outside=True)
        return out

    def sparse_plan(self, P: RenderParams, cams: Sequence, tf: Optional[torch.Tensor]):
        """Prepare a sparse batched march: -> (packed, Cn, Pe, skip_levels) or None if this
        configuration cannot use it (no occupancy grid / skipping off / gamma != 1 / overlays)."""
        P = P.with_projection_of(cams[0])
        if (self.labels is not None and P.showSeg) or (self.preds is not None and P.showPred) or P.gamma != 1.0:
            return None
        packed, Cn, Pe = self.prepared(P)
        bits = self._classify(P, Pe, Cn, tf)
        if bits is None:
            return None
        return packed, Cn, Pe, bits


# ----------------------------------------------------------------------------- autograd
def _segmentable(P: RenderParams) -> bool:
    """The checkpointing forward covers indexed stepping with gamma 1 (mrt_render_forward_ckpt)."""
    return P.tMode == "indexed" and P.gamma == 1.0


class _RenderFn(torch.autograd.Function):
    """Differentiable rendering of one frame or (``cams`` given) a batch of views of one volume."""

    @staticmethod
    def forward(ctx, planar, tf, P: RenderParams, labels, preds, fold, tile_range=None, cams=None):
        Cn = planar.shape[0]
        want_occ = bool(P.skipEmpty) and P.tMode == "indexed"
        # (a single modality takes the fold path too when the occupancy grid is wanted: its fused
        # kernel lays the volume out and builds the brick min/max in one pass over the voxels)
        fold = bool(fold) and (Cn > 1 or want_occ)
        mm = None
        if fold and want_occ:
            (packed, mm), Ce, Pe = fold_volume_occupancy(planar.detach(), P), 1, folded_params(P)
        elif fold:
            packed, Ce, Pe = fold_volume(planar.detach(), P), 1, folded_params(P)
        else:
            packed, Ce, Pe = pack_volume(planar.detach()), Cn, P
        bits = flat = None
        if want_occ:
            if mm is None:
                mm = build_occupancy(packed, Ce, P.dims)
            seg_any = build_label_occupancy(labels) if (labels is not None and P.showSeg) else None
            pred_any = build_label_occupancy(preds) if (preds is not None and P.showPred) else None
            bits = classify_bricks(Pe, mm, Ce, tf, seg_any, pred_any)
            if Ce == 1:
                flat = classify_bricks(Pe, mm, Ce, tf, seg_any, pred_any, flat=True)
        W, H = P.imageSize
        V = None if cams is None else len(cams)
        out = None
        if tile_range is not None:      # a rank's share of the frame: the other pixels are exact zeros
            out = torch.zeros((H, W, 4) if V is None else (V, H, W, 4), dtype=torch.float32, device=planar.device)
        ck = None
        if _segmentable(Pe):
            out, ck = render_forward_ckpt(Pe, cams, packed, Ce, tf, bits, labels, preds, out=out, tile_range=tile_range)
        elif cams is None:
            out = render_forward(Pe, packed, Ce, tf, bits, labels, preds, out=out, tile_range=tile_range)
        else:
            out = render_forward_batch(Pe, cams, packed, Ce, tf, bits, labels, preds, out=out, tile_range=tile_range)
        ctx.tile_range, ctx.cams, ctx.ck = tile_range, cams, ck
        ctx.P, ctx.Pe, ctx.Cn, ctx.Ce, ctx.fold = P, Pe, Cn, Ce, fold
        ctx.mm, ctx.flat = mm, flat
        ctx.labels, ctx.preds = labels, preds
        ctx.save_for_backward(packed, tf if tf is not None else torch.empty(0, device=planar.device), out)
        ctx.has_tf = tf is not None
        return out

    @staticmethod
    def backward(ctx, g):
        packed, tf, out = ctx.saved_tensors
        tf = tf if ctx.has_tf else None
        want_vol, want_tf = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and ctx.has_tf
        dvol, dtf = render_backward(ctx.Pe, packed, ctx.Ce, tf, ctx.labels, ctx.preds, out,
                                    g.contiguous(), want_dvol=want_vol, want_dtf=want_tf,
                                    flat_levels=ctx.flat, minmax=ctx.mm, tile_range=ctx.tile_range,
                                    cams=ctx.cams, ckpt=ctx.ck)
        gvol = None
        if want_vol:
            gvol = unfold_grad(dvol, ctx.P, ctx.Cn) if ctx.fold else unpack_volume(dvol, ctx.Cn, ctx.P.dims)
        return gvol, (dtf if want_tf else None), None, None, None, None, None, None


def packed_layout_f16(dims) -> Tuple[int, int]:
    """(pitchY, pitchZ) in voxels of the packed fp16 layout (``mrt_packed_layout_f16``)."""
    X, Y, Z = dims
    py, pz = C.c_int64(), C.c_int64()
    lib().mrt_packed_layout_f16(X, Y, Z, C.byref(py), C.byref(pz))
    return int(py.value), int(pz.value)


class _ShardRenderFn(torch.autograd.Function):
    """Differentiable PARTIAL render of one sort-last sub-box (premultiplied rgb without background,
    T_local), fp32 or fp16 voxel storage: the building block of a differentiable cfg5."""

    @staticmethod
    def forward(ctx, sub, tf, P: RenderParams, shard, half):
        st = sub.detach()
        if half and st.dtype != torch.float16:
            st = st.half()                       # storage rounding; the gradient passes straight through it
        vol = Volume(st.contiguous(), shard=shard, global_dims=P.dims, quad=False)
        out = vol.forward(P, tf)
        packed, Ce, Pe = vol.prepared(P, quad=False)
        ctx.Pe, ctx.dims, ctx.half, ctx.in_dtype = Pe, vol.dims, bool(vol.half), sub.dtype
        ctx.has_tf = tf is not None
        ctx.save_for_backward(packed, tf if tf is not None else torch.empty(0, device=sub.device), out)
        return out

    @staticmethod
    def backward(ctx, g):
        packed, tf, out = ctx.saved_tensors
        tf = tf if ctx.has_tf else None
        want_vol, want_tf = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and ctx.has_tf
        dvol, dtf = render_backward(ctx.Pe, packed, 1, tf, None, None, out, g.contiguous(),
                                    want_dvol=want_vol, want_dtf=want_tf)
        gvol = None
        if want_vol:
            X, Y, Z = ctx.dims
            if ctx.half:
                pY, pZ = packed_layout_f16(ctx.dims)
                gvol = dvol.as_strided((1, Z, Y, X), (0, pZ, pY, 1)).contiguous()
            else:
                gvol = unpack_volume(dvol, 1, ctx.dims)
            gvol = gvol.to(ctx.in_dtype)
        return gvol, (dtf if want_tf else None), None, None, None


def render_shard(sub: torch.Tensor, shard, global_dims, camera: Optional[Camera], tf: Optional[torch.Tensor],
                 params: RenderParams, storage: Optional[torch.dtype] = None) -> torch.Tensor:
    """Partial frame of ONE sort-last sub-box, differentiable w.r.t. ``sub`` and ``tf``.

    sub     : ``[1,z,y,x]`` CUDA fp32 or fp16 tensor holding voxels ``[lo, hi]`` (inclusive) of the
              global volume, ``shard = (lo, hi)`` in (x,y,z) order (``dist.shard_box``).
    storage : ``torch.float16`` samples an fp16 copy of an fp32 ``sub`` (cfg5's storage; the gradient
              comes back in fp32, straight through the rounding).  An fp16 ``sub`` gets an fp16 gradient.
    Returns float32 ``[H,W,4]`` = (premultiplied rgb WITHOUT background, T_local): composite the
    shards' partials front to back (``dist.composite_over_differentiable``) for the image.  The
    gradients of all shards, added in global coordinates, equal the unsharded gradient; early
    termination acts per shard (use a small ``ertThreshold``)."""
    if sub.dim() != 4 or sub.shape[0] != 1:
        raise ValueError(f"sub must be [1,z,y,x], got {tuple(sub.shape)}")
    _need_cuda(sub, "sub", sub.dtype if sub.dtype in (torch.float16, torch.float32) else torch.float32)
    P = params if camera is None else params.with_camera(camera)
    if tuple(P.dims) != tuple(int(v) for v in global_dims):
        raise ValueError(f"params.dims {P.dims} != global_dims {tuple(global_dims)}")
    if tf is not None:
        _need_cuda(tf, "tf", torch.float32)
    P = replace(P, tfMode=1 if tf is not None else 0)
    P.validate()
    lo, hi = tuple(int(v) for v in shard[0]), tuple(int(v) for v in shard[1])
    half = sub.dtype == torch.float16 or storage == torch.float16
    return _ShardRenderFn.apply(sub, tf, P, (lo, hi), half)


def render(volume: Union[torch.Tensor, Volume], camera: Optional[Camera], tf: Optional[torch.Tensor],
           params: RenderParams, labels: Optional[torch.Tensor] = None,
           preds: Optional[torch.Tensor] = None, fold: bool = True,
           tile_range: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """Render one frame: -> float32 ``[H, W, 4]`` (row 0 = top of the image).

    volume : ``[C,Z,Y,X]`` CUDA fp32 tensor (differentiable) or a prepared :class:`Volume`.
    camera : :class:`camera.Camera` (eye/U/V/W/fov/ortho); ``None`` keeps the ones in ``params``.
    tf     : ``[N,4]`` fp32 (r,g,b,sigma) 1D LUT, or ``None`` for the reference's window/level
             intensity transfer function (brats_rt.slang:132-140).
    params : :class:`RenderParams` (the reference's ``struct Params`` + extensions).
    labels, preds : optional int32 ``[Z,Y,X]`` overlays (gLabels / gPreds).
    fold   : tensor input only — blend the modalities once per voxel before marching
             (see :class:`Volume`); ``False`` blends per sample like the reference shader.
    tile_range : render (and differentiate) only tiles ``[begin, end)``; the other pixels of the
             returned image are zero (image-space data parallelism, ``dist.render_differentiable``).
    Differentiable w.r.t. ``volume`` and ``tf`` when ``volume`` is a tensor.
    """
    P = params if camera is None else params.with_camera(camera)
    if tf is not None:
        _need_cuda(tf, "tf", torch.float32)
        if tf.dim() != 2 or tf.shape[1] != 4 or not (2 <= tf.shape[0] <= _lib.MRT_MAX_TF):
            raise ValueError(f"tf must be [N,4] with 2 <= N <= {_lib.MRT_MAX_TF}, got {tuple(tf.shape)}")
    P = replace(P, tfMode=1 if tf is not None else 0)
    if isinstance(volume, Volume):
        V = volume
        if tuple(P.dims) != tuple(V.global_dims):
            raise ValueError(f"params.dims {P.dims} != volume dims {V.global_dims}")
        P.validate()
        if tile_range is not None:
            W, H = P.imageSize
            out = torch.zeros((H, W, 4), dtype=torch.float32, device=V.device)
            return V.forward(P, tf, out=out, tile_range=tile_range, labels=labels, preds=preds)
        return V.forward(P, tf, labels=labels, preds=preds)
    _need_cuda(volume, "volume", torch.float32)
    if volume.dim() != 4 or not (1 <= volume.shape[0] <= 4):
        raise ValueError(f"volume must be [C,Z,Y,X] with C in 1..4, got {tuple(volume.shape)}")
    Z, Y, X = (int(v) for v in volume.shape[1:])
    if tuple(P.dims) != (X, Y, Z):
        raise ValueError(f"params.dims {P.dims} != volume dims {(X, Y, Z)}")
    for name, lab in (("labels", labels), ("preds", preds)):
        if lab is not None:
            _need_cuda(lab, name, torch.int32)
            if tuple(lab.shape) != (Z, Y, X):
                raise ValueError(f"{name} must be [Z,Y,X]")
    return _RenderFn.apply(volume, tf, P, labels, preds, fold, tile_range)


class TrainStep:
    """One optimisation step of differentiable rendering with an MSE image loss in ONE library call
    (``mrt_train_step_mse``; BASELINE cfg3, docs/DifferentiableRendering.md:88-127 + :213): fold +
    occupancy of the CURRENT volume, classification, checkpointing march, loss, segment-parallel
    adjoint (which forms ``dL/dC = 2 (C - target) / n`` per pixel itself) and the fold's adjoint,
    queued back to back from C with the buffer clears and the loss reduction on a side stream.

    The results equal ``mse_loss(render(volume, ...), target).backward()`` — image bit for bit, loss
    and gradients to rounding — without the autograd graph, the dL/dC tensor and ~20 Python-level
    launches.  Buffers are owned by the object and reused: the tensors a call returns are valid
    until the next call.

        step = TrainStep(params, n_views=1, tf_entries=256)
        loss, image, dvol, dtf = step(volume, tf, target)          # all device tensors
    """

    def __init__(self, params: RenderParams, n_views: int = 1, tf_entries: int = 0, device=None):
        self.P = replace(params, tfMode=1 if tf_entries else 0)
        self.P.validate()
        self.V, self.ntf = int(n_views), int(tf_entries)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        s = self.P.to_struct()
        nbytes = lib().mrt_train_step_workspace_bytes(C.byref(s), self.V, self.ntf)
        if nbytes == 0:
            raise ValueError("TrainStep: " + lib().mrt_last_error().decode("utf-8", "replace"))
        self.workspace = torch.zeros((nbytes,), dtype=torch.uint8, device=self.device)   # zeroed ONCE (header contract)
        W, H = self.P.imageSize
        self.image = torch.empty((self.V, H, W, 4), dtype=torch.float32, device=self.device)
        self.loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self.dvol = None
        self.dtf = torch.empty((self.ntf, 4), dtype=torch.float32, device=self.device) if self.ntf else None

    def __call__(self, volume: torch.Tensor, tf: Optional[torch.Tensor], target: torch.Tensor,
                 cams: Optional[Sequence] = None, want_dvol: bool = True, want_dtf: bool = True):
        """-> (loss 0-dim, image ``[V,H,W,4]`` (``[H,W,4]`` when ``cams`` is None), dL/dvolume
        ``[C,Z,Y,X]`` or None, dL/dtf ``[N,4]`` or None)."""
        P = self.P
        _need_cuda(volume, "volume", torch.float32)
        _need_cuda(target, "target", torch.float32)
        if volume.dim() != 4 or not (1 <= volume.shape[0] <= 4):
            raise ValueError(f"volume must be [C,Z,Y,X] with C in 1..4, got {tuple(volume.shape)}")
        Z, Y, X = (int(v) for v in volume.shape[1:])
        if tuple(P.dims) != (X, Y, Z):
            raise ValueError(f"params.dims {P.dims} != volume dims {(X, Y, Z)}")
        nv = 1 if cams is None else len(cams)
        if nv != self.V:
            raise ValueError(f"TrainStep was built for {self.V} views, got {nv}")
        if target.numel() != self.image.numel():
            raise ValueError(f"target must hold {tuple(self.image.shape)} values, got {tuple(target.shape)}")
        if (tf is None) != (self.ntf == 0) or (tf is not None and tuple(tf.shape) != (self.ntf, 4)):
            raise ValueError(f"TrainStep was built for tf_entries={self.ntf}")
        if tf is not None:
            _need_cuda(tf, "tf", torch.float32)
        want_dtf = bool(want_dtf) and tf is not None
        if not (want_dvol or want_dtf):
            raise ValueError("nothing to differentiate")
        if want_dvol and (self.dvol is None or self.dvol.shape != volume.shape):
            self.dvol = torch.empty_like(volume)
        s = P.to_struct()
        arr = None if cams is None else _camera_array(cams)
        check(lib().mrt_train_step_mse(C.byref(s), None if arr is None else arr.ctypes.data, nv, volume.data_ptr(),
                                       int(volume.shape[0]), _ptr(tf), self.ntf, target.data_ptr(),
                                       self.workspace.data_ptr(), self.image.data_ptr(), self.loss.data_ptr(),
                                       self.dvol.data_ptr() if want_dvol else None,
                                       self.dtf.data_ptr() if want_dtf else None, _stream()), "train_step_mse")
        img = self.image if cams is not None else self.image[0]
        return self.loss, img, (self.dvol if want_dvol else None), (self.dtf if want_dtf else None)


class _AdaptiveFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, planar, tf, P: RenderParams, K, J, eps_w):
        Cn = planar.shape[0]
        packed = pack_volume(planar.detach())
        W, H = P.imageSize
        out = torch.empty((H, W, 4), dtype=torch.float32, device=planar.device)
        s = P.to_struct()
        check(lib().mrt_render_adaptive_forward(C.byref(s), packed.data_ptr(), Cn, _ptr(tf), 0 if tf is None else tf.shape[0],
                                                int(K), int(J), float(eps_w), out.data_ptr(), 0, _tiles.tile_count(W, H), _stream()),
              "render_adaptive_forward")
        ctx.P, ctx.Cn, ctx.K, ctx.J, ctx.eps_w, ctx.has_tf = P, Cn, int(K), int(J), float(eps_w), tf is not None
        ctx.save_for_backward(packed, tf if tf is not None else torch.empty(0, device=planar.device), out)
        return out

    @staticmethod
    def backward(ctx, g):
        packed, tf, out = ctx.saved_tensors
        tf = tf if ctx.has_tf else None
        want_vol, want_tf = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and ctx.has_tf
        P = ctx.P
        W, H = P.imageSize
        dvol = torch.zeros_like(packed) if want_vol else None
        ntf = tf.shape[0] if tf is not None else 2
        dtf = torch.zeros((ntf, 4), dtype=torch.float32, device=packed.device) if want_tf else None
        scratch = torch.empty((lib().mrt_adaptive_scratch_bytes(ntf) // 4,), dtype=torch.float32, device=packed.device) if want_tf else None
        s = P.to_struct()
        check(lib().mrt_render_adaptive_backward(C.byref(s), packed.data_ptr(), ctx.Cn, _ptr(tf), 0 if tf is None else tf.shape[0],
                                                 ctx.K, ctx.J, ctx.eps_w, out.data_ptr(), g.contiguous().data_ptr(), _ptr(dvol),
                                                 _ptr(dtf), _ptr(scratch), 0, _tiles.tile_count(W, H), _stream()),
              "render_adaptive_backward")
        gvol = unpack_volume(dvol, ctx.Cn, P.dims) if want_vol else None
        return gvol, dtf, None, None, None, None


class _SoftOccFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, planar, tf, occ, P: RenderParams):
        Cn = planar.shape[0]
        packed = pack_volume(planar.detach())
        occ_c = occ.detach().contiguous()
        W, H = P.imageSize
        bits = None
        if P.skipEmpty and P.tMode == "indexed":
            bits = classify_bricks(P, build_occupancy(packed, Cn, P.dims), Cn, tf)
        out = torch.empty((H, W, 4), dtype=torch.float32, device=planar.device)
        s = P.to_struct()
        check(lib().mrt_render_forward_soft_occ(C.byref(s), packed.data_ptr(), Cn, _ptr(tf), 0 if tf is None else tf.shape[0],
                                                _ptr(bits), occ_c.data_ptr(), out.data_ptr(), 0, _tiles.tile_count(W, H),
                                                _stream()), "render_forward_soft_occ")
        ctx.P, ctx.Cn, ctx.has_tf = P, Cn, tf is not None
        ctx.save_for_backward(packed, tf if tf is not None else torch.empty(0, device=planar.device), occ_c, out)
        return out

    @staticmethod
    def backward(ctx, g):
        packed, tf, occ, out = ctx.saved_tensors
        tf = tf if ctx.has_tf else None
        P, Cn = ctx.P, ctx.Cn
        W, H = P.imageSize
        want_vol, want_tf, want_occ = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and ctx.has_tf, ctx.needs_input_grad[2]
        dvol = torch.zeros_like(packed) if want_vol else None
        ntf = tf.shape[0] if tf is not None else 2
        dtf = torch.zeros((ntf, 4), dtype=torch.float32, device=packed.device) if want_tf else None
        docc = torch.zeros_like(occ) if want_occ else None
        scratch = torch.empty((lib().mrt_backward_scratch_bytes(W, H, 1, ntf, 1) // 4,), dtype=torch.float32, device=packed.device)
        s = P.to_struct()
        check(lib().mrt_render_backward_soft_occ(C.byref(s), packed.data_ptr(), Cn, _ptr(tf), 0 if tf is None else tf.shape[0],
                                                 occ.data_ptr(), out.data_ptr(), g.contiguous().data_ptr(), _ptr(dvol), _ptr(dtf),
                                                 _ptr(docc), scratch.data_ptr(), 0, _tiles.tile_count(W, H), _stream()),
              "render_backward_soft_occ")
        gvol = unpack_volume(dvol, Cn, P.dims) if want_vol else None
        return gvol, dtf, docc, None


def render_soft_occupancy(volume: torch.Tensor, camera: Optional[Camera], tf: Optional[torch.Tensor], params: RenderParams,
                          occupancy: torch.Tensor) -> torch.Tensor:
    """Differentiable rendering with a learnable, continuous occupancy (docs/DifferentiableRendering.md
    section 11, :202-206: the smooth replacement of hard empty-space skipping): ``occupancy`` is a
    float32 ``[nbz, nby, nbx]`` tensor over the 8^3-voxel brick grid, values in [0,1]; a sample in brick
    b composites with ``sigma' = occupancy[b] * sigma``.  Differentiable w.r.t. ``volume`` ``[C,Z,Y,X]``,
    ``tf`` and ``occupancy``; ``occupancy == 1`` reproduces :func:`render`.  -> ``[H,W,4]``."""
    _need_cuda(volume, "volume", torch.float32)
    _need_cuda(occupancy, "occupancy", torch.float32)
    P = params if camera is None else params.with_camera(camera)
    Z, Y, X = (int(v) for v in volume.shape[1:])
    if tuple(P.dims) != (X, Y, Z):
        raise ValueError(f"params.dims {P.dims} != volume dims {(X, Y, Z)}")
    nb = ((Z + 7) // 8, (Y + 7) // 8, (X + 7) // 8)
    if tuple(occupancy.shape) != nb:
        raise ValueError(f"occupancy must be [nbz,nby,nbx] = {nb}, got {tuple(occupancy.shape)}")
    if tf is not None:
        _need_cuda(tf, "tf", torch.float32)
    P = replace(P, tfMode=1 if tf is not None else 0, showSeg=0, showPred=0)
    P.validate()
    return _SoftOccFn.apply(volume, tf, occupancy, P)


def render_adaptive(volume: torch.Tensor, camera: Optional[Camera], tf: Optional[torch.Tensor], params: RenderParams,
                    n_coarse: int = 16, n_fine: int = 64, eps_w: float = 1e-3) -> torch.Tensor:
    """Differentiable adaptive sampling (docs/DifferentiableRendering.md section 7, :131-148): a coarse
    pass of ``n_coarse`` uniform samples per ray -> piecewise-linear CDF of the extinction (+ ``eps_w``)
    -> ``n_fine`` samples at the fixed quantiles ``(j+1/2)/n_fine`` of the inverse CDF, composited
    front to back with the length of each quantile interval.  ``volume``: ``[C,Z,Y,X]`` CUDA fp32
    tensor; differentiable w.r.t. it and ``tf`` (the gradient includes the motion of the samples with
    the weights).  -> float32 ``[H,W,4]``."""
    P = params if camera is None else params.with_camera(camera)
    _need_cuda(volume, "volume", torch.float32)
    if volume.dim() != 4 or not (1 <= volume.shape[0] <= 4):
        raise ValueError(f"volume must be [C,Z,Y,X] with C in 1..4, got {tuple(volume.shape)}")
    Z, Y, X = (int(v) for v in volume.shape[1:])
    if tuple(P.dims) != (X, Y, Z):
        raise ValueError(f"params.dims {P.dims} != volume dims {(X, Y, Z)}")
    if tf is not None:
        _need_cuda(tf, "tf", torch.float32)
    P = replace(P, tfMode=1 if tf is not None else 0)
    P.validate()
    return _AdaptiveFn.apply(volume, tf, P, n_coarse, n_fine, eps_w)


def render_views(volume: Union[torch.Tensor, Volume], cams: Sequence, tf: Optional[torch.Tensor], params: RenderParams,
                 out: Optional[torch.Tensor] = None, tile_range: Optional[Tuple[int, int]] = None,
                 fold: bool = True, march_events=None) -> torch.Tensor:
    """Render a batch of views of a prepared :class:`Volume` -> float32 ``[V,H,W,4]``: the
    reference's frame loop over successive camera poses (inr/viewer/brats_viewer.py:400-442) as
    one classify + one march launch per 64 views.  View ``v`` is bit-identical to
    ``render(volume, cams[v], tf, params)``.  The cameras must share the projection
    (fov / ortho window) — only eye/U/V/W vary within a batch.  With a ``[C,Z,Y,X]`` tensor instead of
    a Volume the batch is differentiable w.r.t. the tensor and ``tf`` (<= 64 views)."""
    cams = list(cams)
    if not cams:
        raise ValueError("render_views needs at least one camera")
    c0 = cams[0]
    for c in cams[1:]:
        if (bool(c.ortho), float(c.fovY)) != (bool(c0.ortho), float(c0.fovY)) or (
                c.ortho and float(c.ortho_half_height) != float(c0.ortho_half_height)):
            raise ValueError("all cameras of a batch must share fovY / ortho / orthoHalfHeight")
    if tf is not None:
        _need_cuda(tf, "tf", torch.float32)
        if tf.dim() != 2 or tf.shape[1] != 4 or not (2 <= tf.shape[0] <= _lib.MRT_MAX_TF):
            raise ValueError(f"tf must be [N,4] with 2 <= N <= {_lib.MRT_MAX_TF}, got {tuple(tf.shape)}")
    tfm = 1 if tf is not None else 0
    P = params if params.tfMode == tfm else params.derived(("tfmode", tfm), lambda p: replace(p, tfMode=tfm))
    if isinstance(volume, torch.Tensor):
        # differentiable w.r.t. the volume tensor and the TF: one checkpointing march + one
        # segment-parallel backward launch for the whole batch (a multi-view training step)
        _need_cuda(volume, "volume", torch.float32)
        if volume.dim() != 4 or not (1 <= volume.shape[0] <= 4):
            raise ValueError(f"volume must be [C,Z,Y,X] with C in 1..4, got {tuple(volume.shape)}")
        Z, Y, X = (int(v) for v in volume.shape[1:])
        if tuple(P.dims) != (X, Y, Z):
            raise ValueError(f"params.dims {P.dims} != volume dims {(X, Y, Z)}")
        if len(cams) > lib().mrt_max_views_per_launch():
            raise ValueError(f"a differentiable batch holds at most {lib().mrt_max_views_per_launch()} views")
        if out is not None:
            raise ValueError("out= is not supported for a differentiable batch")
        P = P.with_camera(c0)
        P.validate()
        return _RenderFn.apply(volume, tf, P, None, None, fold, tile_range, cams)
    if not isinstance(volume, Volume):
        raise TypeError("render_views needs a prepared Volume or a [C,Z,Y,X] tensor")
    if tuple(P.dims) != tuple(volume.global_dims):
        raise ValueError(f"params.dims {P.dims} != volume dims {volume.global_dims}")
    P.validate()
    if march_events is not None:
        return volume.forward_batch(P, cams, tf, out=out, tile_range=tile_range, march_events=march_events)
    return volume.forward_batch(P, cams, tf, out=out, tile_range=tile_range)


def render_aux(volume: Volume, camera: Optional[Camera], tf: Optional[torch.Tensor], params: RenderParams):
    """Like :func:`render` on a prepared Volume, additionally returning the final transmittance
    ``T [H,W]`` and per-ray counters ``[H,W,4]`` = (n_clip, n_taken, n_evaluated, n_segments)."""
    P = params if camera is None else params.with_camera(camera)
    P = replace(P, tfMode=1 if tf is not None else 0)
    W, H = P.imageSize
    out_T = torch.empty((H, W), dtype=torch.float32, device=volume.device)
    counts = torch.zeros((H, W, 4), dtype=torch.int32, device=volume.device)
    img = volume.forward(P, tf, out_T=out_T, out_counts=counts)
    return img, out_T, counts


def render_slab(vol_u8: torch.Tensor, camera: Optional[Camera], params: SlabParams,
                tile_range: Optional[Tuple[int, int]] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The single-volume slab renderer (``volume_cs``, volume_render.slang:104-148) over a
    uint8 ``[Z,Y,X]`` CUDA volume -> float32 ``[H,W,4]``."""
    _need_cuda(vol_u8, "vol_u8", torch.uint8)
    P = params if camera is None else params.with_camera(camera)
    Z, Y, X = (int(v) for v in vol_u8.shape)
    if tuple(P.volDim) != (X, Y, Z):
        raise ValueError(f"params.volDim {P.volDim} != volume dims {(X, Y, Z)}")
    W, H = P.imageSize
    if out is None:
        out = torch.empty((H, W, 4), dtype=torch.float32, device=vol_u8.device)
    t0, t1 = tile_range if tile_range is not None else (0, _tiles.tile_count(W, H))
    s = P.to_struct()
    check(lib().mrt_render_slab_u8(C.byref(s), vol_u8.data_ptr(), out.data_ptr(), t0, t1, _stream()), "render_slab")
    return out


def render_host(volume: np.ndarray, params: RenderParams, tf: Optional[np.ndarray] = None,
                labels: Optional[np.ndarray] = None, preds: Optional[np.ndarray] = None,
                out: Optional[np.ndarray] = None) -> np.ndarray:
    """Host buffers in, host image out (``mrt_render_host``): the call a non-CUDA host makes.
    Copies H2D, packs, builds the occupancy grid, renders, copies D2H, synchronises."""
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    Cn, Z, Y, X = vol.shape
    P = replace(params, tfMode=1 if tf is not None else 0, dims=(X, Y, Z))
    W, H = P.imageSize
    if out is None:
        out = np.empty((H, W, 4), dtype=np.float32)
    tfa = None if tf is None else np.ascontiguousarray(tf, dtype=np.float32)
    la = None if labels is None else np.ascontiguousarray(labels, dtype=np.int32)
    pa = None if preds is None else np.ascontiguousarray(preds, dtype=np.int32)
    hp = lambda a: None if a is None else a.ctypes.data
    s = P.to_struct()
    check(lib().mrt_render_host(C.byref(s), hp(vol), Cn, hp(tfa), 0 if tfa is None else tfa.shape[0],
                                hp(la), hp(pa), hp(out)), "render_host")
    return out


class InrWeights:
    """The reference's parameter list uploaded once (``inr_upload_params``): flat device tensor + layer widths."""
    __slots__ = ("wts", "dims", "n_layers")

    def __init__(self, wts: torch.Tensor, dims: Sequence[int], n_layers: int):
        self.wts, self.dims, self.n_layers = wts, list(dims), int(n_layers)


def inr_upload_params(params: Sequence[dict], device) -> InrWeights:
    """``[{"W": [in,out], "b": [out]}, ...]`` (numpy or torch) -> :class:`InrWeights` on ``device``; pass it to
    :func:`inr_predict` instead of the list when predicting many volumes with one network (saves the
    host-side flatten + upload per call)."""
    dims = [int(np.asarray(params[0]["W"]).shape[0])] + [int(np.asarray(p["W"]).shape[1]) for p in params]
    flat = []
    for p in params:
        W = torch.as_tensor(np.asarray(p["W"], dtype=np.float32)) if not isinstance(p["W"], torch.Tensor) else p["W"].float().cpu()
        b = torch.as_tensor(np.asarray(p["b"], dtype=np.float32)) if not isinstance(p["b"], torch.Tensor) else p["b"].float().cpu()
        flat += [W.reshape(-1), b.reshape(-1)]
    return InrWeights(torch.cat(flat).contiguous().to(device), dims, len(params))


def inr_predict(mods: torch.Tensor, params, fourier_freqs: int, return_logits: bool = False,
                impl: str = "auto"):
    """INR segmentation of a whole volume on the GPU (``mrt_inr_predict``): the reference's
    ``predict_volume`` (inr/inr/model.py:119-141) as one fused kernel.

    mods   : ``[M,Z,Y,X]`` CUDA float32, z-scored per modality (:func:`volume.zscore_modalities`)
    params : the reference's parameter list ``[{"W": [in,out], "b": [out]}, ...]`` (numpy or torch), or
             the :class:`InrWeights` of :func:`inr_upload_params`
    impl   : "auto" (tcgen05 tensor cores when the network fits, else FFMA), "ffma" (fp32 on the CUDA
             cores: the parity reference), "tensor" (tensor cores or an error)
    -> int32 labels ``[Z,Y,X]`` — directly usable as ``preds`` of :class:`Volume` / :func:`render` —
    and, with ``return_logits``, float32 ``[Z,Y,X,classes]``."""
    _need_cuda(mods, "mods", torch.float32)
    M, Z, Y, X = (int(v) for v in mods.shape)
    w = params if isinstance(params, InrWeights) else inr_upload_params(params, mods.device)
    dims = w.dims
    labels = torch.empty((Z, Y, X), dtype=torch.int32, device=mods.device)
    logits = torch.empty((Z, Y, X, dims[-1]), dtype=torch.float32, device=mods.device) if return_logits else None
    ld = (C.c_int32 * len(dims))(*dims)
    check(lib().mrt_inr_predict(mods.data_ptr(), M, X, Y, Z, w.wts.data_ptr(), C.cast(ld, C.c_void_p), w.n_layers,
                                int(fourier_freqs), labels.data_ptr(), _ptr(logits),
                                {"auto": 0, "ffma": 1, "tensor": 2}[impl], _stream()), "inr_predict")
    return (labels, logits) if return_logits else labels


class HostPipeline:
    """Host buffers in, host frames out, pipelined (``mrt_host_pipeline_*``): per step a list of
    cameras, params and an optional ``[N,4]`` TF go in — plus a ``[C,Z,Y,X]`` float32 host volume, or
    ``None`` to render the RESIDENT volume uploaded once with :meth:`set_volume`, the way the
    reference uploads its buffers at load time and only refills ``gParams`` per frame — and
    ``[V,H,W,4]`` frames land in a host array.  Upload, prepare, march and download of successive
    steps overlap (four streams); pass page-locked arrays (e.g. ``torch.Tensor.pin_memory()``
    ``.numpy()``) so the copies are asynchronous.  ``submit`` queues a step and returns a ticket;
    ``wait`` blocks until that step's frames are on the host.

    Frames are downloaded sparse (only each view's bounding rectangle of non-background tiles); the
    pipeline keeps the rest of an output array at the background colour by tracking what it last
    wrote there.  An array handed to ``submit`` therefore belongs to the pipeline until
    :meth:`forget` / :meth:`close`: do not write to it in between."""

    def __init__(self, C_: int, dims, image_size, max_views: int, max_tf: int = 256, depth: int = 2):
        X, Y, Z = (int(v) for v in dims)
        W, H = (int(v) for v in image_size)
        self._h = C.c_void_p()
        self.C, self.dims, self.image_size, self.max_views = int(C_), (X, Y, Z), (W, H), int(max_views)
        rc = lib().mrt_host_pipeline_create(C.byref(self._h), self.C, X, Y, Z, W, H, int(max_views), int(max_tf), int(depth))
        if rc != 0:
            raise _lib.MrtError(f"host_pipeline_create failed ({rc})")
        self._keep = {}
        self._outs = {}          # address -> array: the outputs the pipeline tracks (kept alive so addresses are not recycled)
        self._resident = None

    def _err(self, what, rc):
        return _lib.MrtError(f"{what} failed ({rc}): {lib().mrt_host_pipeline_error(self._h).decode('utf-8', 'replace')}")

    def _check_volume(self, volume):
        X, Y, Z = self.dims
        if volume.dtype != np.float32 or not volume.flags.c_contiguous or volume.shape != (self.C, Z, Y, X):
            raise ValueError(f"volume must be C-contiguous float32 {(self.C, Z, Y, X)}")

    def set_volume(self, volume: np.ndarray):
        """Upload ``volume`` once; later ``submit(None, ...)`` calls render it."""
        self._check_volume(volume)
        rc = lib().mrt_host_pipeline_set_volume(self._h, volume.ctypes.data)
        if rc != 0:
            raise self._err("host_pipeline_set_volume", rc)
        self._resident = volume

    def submit(self, volume: Optional[np.ndarray], cams: Sequence, params: RenderParams, tf: Optional[np.ndarray],
               out: np.ndarray, fresh: bool = False) -> int:
        W, H = self.image_size
        if volume is not None:
            self._check_volume(volume)
        elif self._resident is None:
            raise ValueError("no volume: pass one or call set_volume() first")
        V = len(cams)
        if out.dtype != np.float32 or not out.flags.c_contiguous or out.shape != (V, H, W, 4):
            raise ValueError(f"out must be C-contiguous float32 {(V, H, W, 4)}")
        if tf is not None and (tf.dtype != np.float32 or not tf.flags.c_contiguous or tf.ndim != 2 or tf.shape[1] != 4):
            raise ValueError("tf must be C-contiguous float32 [N,4]")
        P = replace(params.with_camera(cams[0]), tfMode=1 if tf is not None else 0, dims=self.dims, imageSize=self.image_size)
        s = P.to_struct()
        arr = _camera_array(cams)
        ticket = C.c_int64(-1)
        addr = out.ctypes.data
        flags = 1 if (fresh or addr not in self._outs) else 0
        rc = lib().mrt_host_pipeline_submit(self._h, C.byref(s), arr.ctypes.data, V,
                                            None if volume is None else volume.ctypes.data,
                                            None if tf is None else tf.ctypes.data, 0 if tf is None else tf.shape[0],
                                            addr, flags, C.byref(ticket))
        if rc != 0:
            raise self._err("host_pipeline_submit", rc)
        self._outs[addr] = out
        self._keep[ticket.value] = (volume, tf, out)      # the host buffers must outlive the queued copies
        return int(ticket.value)

    def last_bytes(self):
        """(host->device, device->host, host background fill) bytes of the last submitted step."""
        b = (C.c_uint64 * 3)()
        lib().mrt_host_pipeline_last_bytes(self._h, C.byref(b))
        return int(b[0]), int(b[1]), int(b[2])

    def forget(self, out: np.ndarray):
        """Stop tracking ``out`` (the caller wants to write to it, or to free it)."""
        lib().mrt_host_pipeline_forget(self._h, out.ctypes.data)
        self._outs.pop(out.ctypes.data, None)

    def wait(self, ticket: int):
        rc = lib().mrt_host_pipeline_wait(self._h, int(ticket))
        if rc != 0:
            raise self._err("host_pipeline_wait", rc)
        for t in [t for t in self._keep if t <= ticket]:
            del self._keep[t]

    def close(self):
        if self._h:
            lib().mrt_host_pipeline_destroy(self._h)
            self._h = C.c_void_p()
            self._outs.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def tile_index_map(W: int, H: int, device="cuda"):
    """Device evaluation of the integer tile map (bit-exactness check): (tile[H,W], lane[H,W])."""
    t = torch.full((H, W), -1, dtype=torch.int32, device=device)
    l = torch.full((H, W), -1, dtype=torch.int32, device=device)
    check(lib().mrt_tile_index_map(W, H, t.data_ptr(), l.data_ptr(), _stream()), "tile_index_map")
    return t, l
