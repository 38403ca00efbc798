"""Volume ingest: the step immediately before the hot path.

Host-side mirrors of the reference loaders, array in / array out (NIfTI file I/O itself —
``nibabel`` — is out of scope and absent here):

* :func:`normalize_percentile`  — ``load_nifti_float`` (inr/viewer/brats_viewer.py:46-65):
  1st / 99.5th percentile window -> clip to [0,1] -> (X,Y,Z) -> (Z,Y,X) x-fastest flatten.
* :func:`labels_from_float`     — ``load_seg_uint`` (inr/viewer/brats_viewer.py:68-74).
* :func:`world_scaling`         — ``load_dir`` (inr/viewer/brats_viewer.py:204-210).
* :func:`nifti_mask_to_u8`      — ``_load_nifti_mask`` (scripts/volumeRendering/app.py:167-198).
* :func:`decode_bc4_host` / :func:`decode_bc4` — ``_load_volume_bc4`` (app.py:200-250), on the
  host (numpy) and as a CUDA kernel (``mrt_decode_bc4``).
* :func:`normalize_on_device`, :func:`u8_to_f32` — the same arithmetic as CUDA kernels for
  volumes that are already resident.

All of these are checked against vectors produced by executing the reference functions
(tests/golden/make_golden.py -> tests/test_ingest_golden.py).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


# ----------------------------------------------------------------------------- fp32 modalities
def percentile_window(data: np.ndarray) -> Tuple[float, float]:
    """(vmin, rng) of the display window (brats_viewer.py:50-55)."""
    vmin = float(np.percentile(data, 1.0))
    vmax = float(np.percentile(data, 99.5))
    if vmax <= vmin:
        vmax = float(np.max(data))
        vmin = float(np.min(data))
    return vmin, max(1e-6, vmax - vmin)


def normalize_percentile(data_xyz: np.ndarray):
    """float32 (X,Y,Z) array -> (linear [Z*Y*X] x-fastest, norm (X,Y,Z), dims uint32[3]).

    Same values and layout as ``load_nifti_float``; ``linear.reshape(Z,Y,X)`` is one channel of
    the ``[C,Z,Y,X]`` tensor the renderer takes."""
    data = np.asarray(data_xyz, dtype=np.float32)
    vmin, rng = percentile_window(data)
    norm = np.clip((data - vmin) / rng, 0.0, 1.0).astype(np.float32)
    linear = np.ascontiguousarray(norm.transpose(2, 1, 0).reshape(-1))
    return linear, norm, np.array(norm.shape, dtype=np.uint32)


def labels_from_float(data_xyz: np.ndarray) -> np.ndarray:
    """Segmentation stored as floats -> uint32 labels, (Z,Y,X) flatten (brats_viewer.py:68-74)."""
    labels = np.rint(np.asarray(data_xyz, dtype=np.float32)).astype(np.uint32)
    return np.ascontiguousarray(labels.transpose(2, 1, 0).reshape(-1))


def world_scaling(dims_xyz, zooms) -> Tuple[np.ndarray, np.ndarray]:
    """voxelSize = zooms * float32(1.8 / max_dim); volMin = -0.5 * voxelSize * dims
    (brats_viewer.py:204-210), all float32."""
    d = np.asarray(dims_xyz, dtype=np.uint32)
    scale = np.float32(1.8 / float(max(d)))
    voxel_size = (np.asarray(zooms, dtype=np.float32) * scale).astype(np.float32)
    vol_min = -0.5 * (voxel_size * d.astype(np.float32))
    return voxel_size, vol_min


def stack_modalities(linears, dims_xyz) -> np.ndarray:
    """List of flattened modalities -> ``[C,Z,Y,X]`` float32 (the renderer's input)."""
    X, Y, Z = (int(v) for v in dims_xyz)
    return np.stack([np.asarray(l, dtype=np.float32).reshape(Z, Y, X) for l in linears], axis=0)


# ----------------------------------------------------------------------------- u8 volumes
def nifti_mask_to_u8(data_xyz: np.ndarray, mode: str = "occupancy") -> np.ndarray:
    """Mask -> uint8 [Z,Y,X] (app.py:179-195): occupancy = (data > 0.5)*255;
    labels: BraTS 1 -> 85, 2 -> 170, 4 -> 255."""
    data = np.asarray(data_xyz, dtype=np.float32)
    if mode == "occupancy":
        vol = (data > 0.5).astype(np.uint8) * 255
    elif mode == "labels":
        vol = np.zeros_like(data, dtype=np.uint8)
        vol[np.isclose(data, 1.0)] = 85
        vol[np.isclose(data, 2.0)] = 170
        vol[np.isclose(data, 4.0)] = 255
    else:
        raise ValueError(f"Unknown mask_mode '{mode}'. Use 'occupancy' or 'labels'.")
    return np.ascontiguousarray(np.transpose(vol, (2, 1, 0)))


def decode_bc4_host(blocks: np.ndarray, W: int, H: int, D: int) -> np.ndarray:
    """BC4 block stream uint8 [D, ceil(H/4)*ceil(W/4), 8] -> uint8 [D,H,W] on the host.
    Block = (r0, r1, 48 bits of sixteen 3-bit palette codes); 8-entry palette with six
    interpolants when r0 > r1, else four interpolants + 0 + 255 (app.py:216-236)."""
    bw, bh = (W + 3) // 4, (H + 3) // 4
    b = np.asarray(blocks, dtype=np.uint8).reshape(D, bw * bh, 8)
    r0 = b[..., 0].astype(np.int32)
    r1 = b[..., 1].astype(np.int32)
    bits = np.zeros(b.shape[:2], dtype=np.uint64)
    for i in range(6):
        bits |= b[..., 2 + i].astype(np.uint64) << np.uint64(8 * i)
    pal = np.zeros(b.shape[:2] + (8,), dtype=np.int32)
    pal[..., 0], pal[..., 1] = r0, r1
    big = r0 > r1
    for i in range(1, 7):
        six = ((7 - i) * r0 + i * r1 + 3) // 7
        four = ((5 - i) * r0 + i * r1 + 2) // 5 if i < 5 else (0 if i == 5 else 255)
        pal[..., i + 1] = np.where(big, six, four)
    out = np.zeros((D, bh * 4, bw * 4), dtype=np.uint8)
    for t in range(16):
        code = ((bits >> np.uint64(3 * t)) & np.uint64(7)).astype(np.int64)
        texel = np.take_along_axis(pal, code[..., None], axis=2)[..., 0].astype(np.uint8).reshape(D, bh, bw)
        out[:, (t >> 2)::4, (t & 3)::4] = texel
    return np.ascontiguousarray(out[:, :H, :W])


# ----------------------------------------------------------------------------- device kernels
def decode_bc4(blocks, W: int, H: int, D: int):
    """CUDA BC4 decode: uint8 CUDA tensor of blocks -> uint8 CUDA tensor [D,H,W]."""
    import torch
    from ._lib import check, lib
    from .api import _need_cuda, _stream
    _need_cuda(blocks, "blocks", torch.uint8)
    bw, bh = (W + 3) // 4, (H + 3) // 4
    if blocks.numel() != D * bw * bh * 8:
        raise ValueError(f"BC4 data size mismatch: {blocks.numel()} vs {D * bw * bh * 8}")
    out = torch.empty((D, H, W), dtype=torch.uint8, device=blocks.device)
    check(lib().mrt_decode_bc4(blocks.data_ptr(), W, H, D, out.data_ptr(), _stream()), "decode_bc4")
    return out


def u8_to_f32(vol_u8):
    """uint8 CUDA tensor -> float32 / 255 (volume_render.slang:38)."""
    import torch
    from ._lib import check, lib
    from .api import _need_cuda, _stream
    _need_cuda(vol_u8, "vol_u8", torch.uint8)
    out = torch.empty(vol_u8.shape, dtype=torch.float32, device=vol_u8.device)
    check(lib().mrt_u8_to_f32(vol_u8.data_ptr(), vol_u8.numel(), out.data_ptr(), _stream()), "u8_to_f32")
    return out


def normalize_on_device(data, vmin: float, rng: float):
    """clip((data - vmin)/rng, 0, 1) on a float32 CUDA tensor, fp32 op by op like numpy."""
    import torch
    from ._lib import check, lib
    from .api import _need_cuda, _stream
    _need_cuda(data, "data", torch.float32)
    out = torch.empty_like(data)
    check(lib().mrt_normalize_f32(data.data_ptr(), data.numel(), float(vmin), float(rng), out.data_ptr(),
                                  _stream()), "normalize_f32")
    return out


def zscore_modalities(planar):
    """Per-modality z-score over the non-zero voxels, zeros included in the output as (0-mu)/sigma —
    the viewer's preprocessing before INR inference (inr/viewer/brats_viewer.py:279-287:
    ``mask = arr != 0; mu = arr[mask].mean(); sigma = arr[mask].std() + 1e-6``).  ``planar``: torch
    ``[M,Z,Y,X]`` float32 on any device -> same shape."""
    import torch
    out = torch.empty_like(planar)
    for m in range(planar.shape[0]):
        arr = planar[m]
        mask = arr != 0
        if bool(mask.any()):
            vals = arr[mask]
            mu = vals.mean()
            sigma = vals.std(unbiased=False) + 1e-6
            out[m] = (arr - mu) / sigma
        else:
            out[m] = arr
    return out
