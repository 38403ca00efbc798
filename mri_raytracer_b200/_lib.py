"""ctypes binding of libmrt.so — the C ABI in include/mrt.h.

The product path fails loudly when the CUDA library is missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmrt.so"
_lib = None

MRT_MAX_TF = 1024
MRT_BRICK = 8
MRT_TILE = 8


class MrtParams(C.Structure):
    """Mirror of ``struct MrtParams`` (include/mrt.h); the first 368 bytes are the reference's
    ``struct Params`` cbuffer (inr/viewer/brats_rt.slang:12-31)."""
    _fields_ = [
        ("imageSize", C.c_uint32 * 2), ("fovY", C.c_float), ("pad0", C.c_float),
        ("eye", C.c_float * 3), ("pad1", C.c_float),
        ("U", C.c_float * 3), ("pad2", C.c_float),
        ("V", C.c_float * 3), ("pad3", C.c_float),
        ("W", C.c_float * 3), ("pad4", C.c_float),
        ("volMin", C.c_float * 3), ("pad5", C.c_float),
        ("voxelSize", C.c_float * 3), ("pad6", C.c_float),
        ("dims", C.c_uint32 * 3), ("pad7", C.c_uint32),
        ("stepSize", C.c_float), ("nearT", C.c_float), ("farT", C.c_float), ("pad8", C.c_float),
        ("bgColor", C.c_float * 3), ("pad9", C.c_float),
        ("volEnabled", C.c_uint32 * 4),
        ("volWeight", C.c_float * 4),
        ("ww", C.c_float), ("wl", C.c_float), ("intensityAlpha", C.c_float), ("padInt", C.c_float),
        ("gamma", C.c_float), ("gradBoost", C.c_float), ("gradScale", C.c_float), ("padTone", C.c_float),
        ("showSeg", C.c_uint32), ("showPred", C.c_uint32), ("padFlags", C.c_uint32 * 2),
        ("lutColorAlpha", (C.c_float * 4) * 8),
        ("ortho", C.c_uint32), ("orthoHalfHeight", C.c_float), ("ertThreshold", C.c_float),
        ("maxSteps", C.c_uint32),
        ("tMode", C.c_uint32), ("alphaMode", C.c_uint32), ("skipEmpty", C.c_uint32), ("tfMode", C.c_uint32),
        ("shardEnabled", C.c_uint32), ("shardLo", C.c_uint32 * 3), ("shardHi", C.c_uint32 * 3), ("volDtype", C.c_uint32),
    ]


class MrtCamera(C.Structure):
    """Mirror of ``struct MrtCamera``: the four camera rows of the reference's ``struct Params``
    (inr/viewer/brats_rt.slang:15-18)."""
    _fields_ = [("eye", C.c_float * 3), ("pad0", C.c_float), ("U", C.c_float * 3), ("pad1", C.c_float),
                ("V", C.c_float * 3), ("pad2", C.c_float), ("W", C.c_float * 3), ("pad3", C.c_float)]


class MrtSlabParams(C.Structure):
    """Mirror of ``struct MrtSlabParams`` (scripts/volumeRendering/volume_render.slang:9-21)."""
    _fields_ = [
        ("imageSize", C.c_uint32 * 2), ("fovY", C.c_float), ("stepCount", C.c_float),
        ("nearPlane", C.c_float), ("farPlane", C.c_float), ("pad0", C.c_float * 2),
        ("eye", C.c_float * 3), ("padEye", C.c_float),
        ("U", C.c_float * 3), ("padU", C.c_float),
        ("V", C.c_float * 3), ("padV", C.c_float),
        ("W", C.c_float * 3), ("padW", C.c_float),
        ("volDim", C.c_uint32 * 3), ("padDim", C.c_uint32),
    ]


_vp, _i32, _u32, _sz, _f = C.c_void_p, C.c_int32, C.c_uint32, C.c_size_t, C.c_float
_PP = C.POINTER(MrtParams)
_SP = C.POINTER(MrtSlabParams)

# name -> (restype, argtypes): exactly the symbols include/mrt.h declares.
PROTOTYPES = {
    "mrt_version": (C.c_int, []),
    "mrt_last_error": (C.c_char_p, []),
    "mrt_sizeof_params": (_sz, []),
    "mrt_sizeof_slab_params": (_sz, []),
    "mrt_sizeof_camera": (_sz, []),
    "mrt_max_views_per_launch": (_i32, []),
    "mrt_tiles_x": (_i32, [_i32]),
    "mrt_tiles_y": (_i32, [_i32]),
    "mrt_tile_count": (_i32, [_i32, _i32]),
    "mrt_tile_of_pixel": (_i32, [_i32, _i32, _i32]),
    "mrt_lane_of_pixel": (_i32, [_i32, _i32]),
    "mrt_rank_tile_range": (None, [_i32, _i32, _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "mrt_tile_index_map": (C.c_int, [_i32, _i32, _vp, _vp, _vp]),
    "mrt_packed_layout": (None, [_i32, _i32, _i32, _i32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mrt_packed_volume_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "mrt_pack_volume_f32": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "mrt_unpack_volume_f32": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "mrt_packed_volume_bytes_f16": (_sz, [_i32, _i32, _i32]),
    "mrt_packed_layout_f16": (None, [_i32, _i32, _i32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mrt_pack_volume_f16": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "mrt_unpack_volume_f16": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "mrt_build_occupancy_f16": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "mrt_render_forward_tma": (C.c_int, [C.POINTER(MrtParams), _vp, _vp, _i32, _vp, _vp, _i32, _i32, _vp, _vp]),
    "mrt_packed_volume_bytes_quad_f16": (_sz, [_i32, _i32, _i32]),
    "mrt_pack_volume_quad_f16": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "mrt_packed_volume_bytes_quad": (_sz, [_i32, _i32, _i32]),
    "mrt_pack_volume_quad": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "mrt_packed_volume_bytes_u8": (_sz, [_i32, _i32, _i32]),
    "mrt_pack_volume_u8": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "mrt_build_occupancy_u8": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "mrt_fold_volume_f32": (C.c_int, [_PP, _vp, _i32, _vp, _vp]),
    "mrt_fold_volume_occupancy_f32": (C.c_int, [_PP, _vp, _i32, _vp, _vp, _vp]),
    "mrt_fold_volume_occupancy_quad_f32": (C.c_int, [_PP, _vp, _i32, _vp, _vp, _vp, _vp]),
    "mrt_unfold_grad_f32": (C.c_int, [_PP, _vp, _i32, _vp, _vp]),
    "mrt_brick_count": (_i32, [_i32, _i32, _i32]),
    "mrt_skip_levels_bytes": (_sz, [_i32, _i32, _i32]),
    "mrt_build_occupancy": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "mrt_build_label_occupancy": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "mrt_classify_bricks": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _i32, _vp]),
    "mrt_render_forward": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "mrt_render_forward_batch": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                           _i32, _i32, _vp]),
    "mrt_view_spans": (C.c_int, [_PP, _vp, _i32, _i32, _vp, _vp, _vp]),
    "mrt_render_forward_batch_sparse": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _i32, _vp]),
    "mrt_fill_outside_spans": (C.c_int, [_PP, _vp, _i32, _vp, _vp]),
    "mrt_fill_outside_spans_delta": (C.c_int, [_PP, _vp, _vp, _i32, _vp, _vp]),
    "mrt_render_forward_batch_scatter": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "mrt_checkpoint_plan": (C.c_int, [_PP, _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "mrt_checkpoint_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "mrt_half_tile_count": (_i32, [_i32, _i32]),
    "mrt_render_forward_ckpt": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                                          _vp, _vp, _i32, _i32, _vp]),
    "mrt_backward_scratch_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32]),
    "mrt_render_backward": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "mrt_render_views_refold": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp]),
    "mrt_render_views_refold_scatter": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _vp]),
    "mrt_train_step_workspace_bytes": (_sz, [_PP, _i32, _i32]),
    "mrt_debug_train_trace": (C.c_int, [_vp]),
    "mrt_train_step_mse": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mrt_render_forward_soft_occ": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp]),
    "mrt_render_backward_soft_occ": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "mrt_adaptive_scratch_bytes": (_sz, [_i32]),
    "mrt_render_adaptive_forward": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _i32, _i32, _f, _vp, _i32, _i32, _vp]),
    "mrt_render_adaptive_backward": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _i32, _i32, _f, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "mrt_render_slab_u8": (C.c_int, [_SP, _vp, _vp, _i32, _i32, _vp]),
    "mrt_decode_bc4": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "mrt_u8_to_f32": (C.c_int, [_vp, _sz, _vp, _vp]),
    "mrt_normalize_f32": (C.c_int, [_vp, _sz, _f, _f, _vp, _vp]),
    "mrt_inr_predict": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _i32, _vp]),
    "mrt_composite_over": (C.c_int, [_vp, _i32, _vp, _sz, _vp, _i32, _vp, _vp]),
    "mrt_composite_over_multi": (C.c_int, [_vp, _i32, _vp, _sz, _vp, _i32, _vp, _i32, _vp]),
    "mrt_render_forward_strips": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "mrt_gather_probe": (C.c_int, [_vp, _sz, _sz, _u32, _vp, _vp]),
    "mrt_host_pipeline_create": (C.c_int, [C.POINTER(_vp), _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    "mrt_host_pipeline_set_volume": (C.c_int, [_vp, _vp]),
    "mrt_host_pipeline_submit": (C.c_int, [_vp, _PP, _vp, _i32, _vp, _vp, _i32, _vp, _i32, C.POINTER(C.c_int64)]),
    "mrt_host_pipeline_last_bytes": (None, [_vp, C.POINTER(C.c_uint64 * 3)]),
    "mrt_host_pipeline_forget": (None, [_vp, _vp]),
    "mrt_host_pipeline_wait": (C.c_int, [_vp, C.c_int64]),
    "mrt_host_pipeline_error": (C.c_char_p, [_vp]),
    "mrt_host_pipeline_destroy": (None, [_vp]),
    "mrt_render_host": (C.c_int, [_PP, _vp, _i32, _vp, _i32, _vp, _vp, _vp]),
}


class MrtError(RuntimeError):
    pass


def lib():
    """Load libmrt.so (once).  Raises if it has not been built — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise MrtError(
            f"{LIB_PATH} is missing: build it with `python -m mri_raytracer_b200.build` "
            "(there is no CPU fallback for the render path)")
    L = C.CDLL(str(LIB_PATH))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(L, name)           # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if L.mrt_sizeof_params() != C.sizeof(MrtParams):
        raise MrtError(f"MrtParams ABI mismatch: lib {L.mrt_sizeof_params()} vs ctypes {C.sizeof(MrtParams)}")
    if L.mrt_sizeof_slab_params() != C.sizeof(MrtSlabParams):
        raise MrtError("MrtSlabParams ABI mismatch")
    if L.mrt_sizeof_camera() != C.sizeof(MrtCamera):
        raise MrtError("MrtCamera ABI mismatch")
    _lib = L
    return L


def check(rc: int, what: str = "libmrt"):
    if rc != 0:
        msg = lib().mrt_last_error().decode("utf-8", "replace")
        raise MrtError(f"{what} failed ({rc}): {msg}")
