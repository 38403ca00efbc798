"""Seeded synthetic BraTS-shaped volumes: there is no network and the
reference ships no loadable volume, so tests/bench use this generator.

``make_brats_like(C, dims=(X,Y,Z), seed)`` -> float32 ``[C,Z,Y,X]`` in [0,1]:
skull-stripped look (exactly 0 outside a centred ellipsoid with semi-axes
(0.36X, 0.42Y, 0.40Z), ~25 % of voxels non-zero), smooth low-frequency tissue contrast,
per-voxel noise and three Gaussian "tumour" blobs.  Deterministic for a given
(seed, dims, C) on any device: all randomness is either a tiny CPU-generated table or an
integer hash of the voxel index.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def _hash_noise(n: int, seed: int, device) -> torch.Tensor:
    """U(-1,1) from an integer hash of the linear voxel index (device independent)."""
    i = torch.arange(n, dtype=torch.int64, device=device)
    x = (i + 0x9E3779B9 * (seed + 1)) & 0xFFFFFFFF
    x = ((x ^ (x >> 16)) * 0x7FEB352D) & 0xFFFFFFFF
    x = ((x ^ (x >> 15)) * 0x846CA68B) & 0xFFFFFFFF
    x = x ^ (x >> 16)
    return (x & 0xFFFFFF).to(torch.float32) * (2.0 / 16777216.0) - 1.0


def _coords(dims, device):
    X, Y, Z = dims
    z = torch.arange(Z, dtype=torch.float32, device=device)[:, None, None]
    y = torch.arange(Y, dtype=torch.float32, device=device)[None, :, None]
    x = torch.arange(X, dtype=torch.float32, device=device)[None, None, :]
    return x, y, z


def _blobs(dims, seed: int):
    g = torch.Generator().manual_seed(10_000 + seed)
    X, Y, Z = dims
    u = torch.rand(3, 5, generator=g)
    out = []
    for b in range(3):
        cx = (0.35 + 0.30 * float(u[b, 0])) * X
        cy = (0.35 + 0.30 * float(u[b, 1])) * Y
        cz = (0.35 + 0.30 * float(u[b, 2])) * Z
        sig = (6.0 + 8.0 * float(u[b, 3])) * (min(dims) / 155.0)
        out.append((cx, cy, cz, max(sig, 1.0)))
    return out


def make_brats_like(C: int = 1, dims: Tuple[int, int, int] = (240, 240, 155), seed: int = 0,
                    device: Optional[torch.device] = None, with_labels: bool = False):
    device = torch.device(device) if device is not None else torch.device("cpu")
    X, Y, Z = dims
    x, y, z = _coords(dims, device)
    ex = ((x - 0.5 * (X - 1)) / (0.36 * X)) ** 2 + ((y - 0.5 * (Y - 1)) / (0.42 * Y)) ** 2 \
        + ((z - 0.5 * (Z - 1)) / (0.40 * Z)) ** 2
    mask = ex <= 1.0
    blobs = _blobs(dims, seed)
    vols = []
    blob_field = torch.zeros((Z, Y, X), dtype=torch.float32, device=device)
    for (cx, cy, cz, sig) in blobs:
        blob_field = torch.maximum(blob_field, torch.exp(-((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2) / (2 * sig * sig)))
    for c in range(C):
        g = torch.Generator().manual_seed(seed * 131 + c)
        low = torch.rand(1, 1, 8, 8, 8, generator=g).to(device)
        smooth = F.interpolate(low, size=(Z, Y, X), mode="trilinear", align_corners=True)[0, 0]
        noise = _hash_noise(X * Y * Z, seed * 131 + c, device).reshape(Z, Y, X)
        v = 0.35 + 0.25 * smooth + 0.05 * noise
        sign = 1.0 if (c % 2 == 0) else -1.0
        v = v + 0.3 * sign * blob_field
        v = torch.where(mask, v.clamp(0.0, 1.0), torch.zeros_like(v))
        vols.append(v)
    vol = torch.stack(vols, dim=0).contiguous()
    if not with_labels:
        return vol
    lab = torch.zeros((Z, Y, X), dtype=torch.int32, device=device)
    lab = torch.where(blob_field > 0.35, torch.full_like(lab, 2), lab)      # edema
    lab = torch.where(blob_field > 0.60, torch.full_like(lab, 1), lab)      # necrotic core
    lab = torch.where(blob_field > 0.80, torch.full_like(lab, 3), lab)      # enhancing
    lab = torch.where(mask, lab, torch.zeros_like(lab))
    return vol, lab.contiguous()


def _interp_axis(table: torch.Tensor, n_global: int, lo: int, hi: int, dim: int) -> torch.Tensor:
    """Linear interpolation (align_corners) of ``table`` along ``dim`` from 8 knots to the global
    positions lo..hi of an axis with n_global samples."""
    pos = torch.arange(lo, hi + 1, dtype=torch.float64, device=table.device) * (7.0 / max(n_global - 1, 1))
    i0 = pos.floor().clamp(0, 6).to(torch.int64)
    f = (pos - i0.to(torch.float64)).to(torch.float32)
    a = table.index_select(dim, i0)
    b = table.index_select(dim, i0 + 1)
    shape = [1] * table.dim()
    shape[dim] = -1
    return a + (b - a) * f.view(shape)


def make_brats_like_box(dims: Tuple[int, int, int], lo, hi, seed: int = 0, device=None,
                        dtype: torch.dtype = torch.float16, zchunk: int = 32) -> torch.Tensor:
    """Voxels [lo, hi] (inclusive, (x,y,z) order) of a single-channel BraTS-like volume of GLOBAL
    size ``dims`` -> ``[1, nz, ny, nx]``.  Same recipe as :func:`make_brats_like` (ellipsoid mask,
    smooth 8^3 tissue table, hashed per-voxel noise, three blobs) but every voxel is a pure function
    of (seed, dims, global index), so a brick-sharded volume (BASELINE config 5: 2048^3 over 8
    GPUs) is generated shard by shard, slab by slab, never materialising the whole."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    X, Y, Z = dims
    nx, ny, nz = (int(h) - int(l) + 1 for l, h in zip(lo, hi))
    g = torch.Generator().manual_seed(seed * 131 + 7)
    table = torch.rand(8, 8, 8, generator=g).to(device)                    # [z][y][x] knots
    txy = _interp_axis(_interp_axis(table, X, lo[0], hi[0], 2), Y, lo[1], hi[1], 1)      # [8, ny, nx]
    blobs = _blobs(dims, seed)
    x = torch.arange(lo[0], hi[0] + 1, dtype=torch.float32, device=device)[None, None, :]
    y = torch.arange(lo[1], hi[1] + 1, dtype=torch.float32, device=device)[None, :, None]
    out = torch.empty((1, nz, ny, nx), dtype=dtype, device=device)
    xi = torch.arange(lo[0], hi[0] + 1, dtype=torch.int64, device=device)[None, None, :]
    yi = torch.arange(lo[1], hi[1] + 1, dtype=torch.int64, device=device)[None, :, None]
    for z0 in range(lo[2], hi[2] + 1, zchunk):
        z1 = min(z0 + zchunk - 1, hi[2])
        z = torch.arange(z0, z1 + 1, dtype=torch.float32, device=device)[:, None, None]
        zi = torch.arange(z0, z1 + 1, dtype=torch.int64, device=device)[:, None, None]
        ex = ((x - 0.5 * (X - 1)) / (0.36 * X)) ** 2 + ((y - 0.5 * (Y - 1)) / (0.42 * Y)) ** 2 \
            + ((z - 0.5 * (Z - 1)) / (0.40 * Z)) ** 2
        smooth = _interp_axis(txy, Z, z0, z1, 0)
        lin = (zi * Y + yi) * X + xi
        h = (lin + 0x9E3779B9 * (seed + 1)) & 0xFFFFFFFF
        h = ((h ^ (h >> 16)) * 0x7FEB352D) & 0xFFFFFFFF
        h = ((h ^ (h >> 15)) * 0x846CA68B) & 0xFFFFFFFF
        h = h ^ (h >> 16)
        noise = (h & 0xFFFFFF).to(torch.float32) * (2.0 / 16777216.0) - 1.0
        v = 0.35 + 0.25 * smooth + 0.05 * noise
        for (cx, cy, cz, sig) in blobs:
            v = v + 0.3 * torch.exp(-((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2) / (2 * sig * sig))
        v = torch.where(ex <= 1.0, v.clamp(0.0, 1.0), torch.zeros_like(v))
        out[0, z0 - lo[2]:z1 - lo[2] + 1] = v.to(dtype)
    return out


def ramp_tf(n: int = 256, sigma_scale: float = 40.0, cutoff: float = 0.08) -> torch.Tensor:
    """The bench transfer function: rgb = val, sigma = 40*val above 0.08 else 0."""
    val = torch.linspace(0.0, 1.0, n)
    sig = torch.where(val > cutoff, sigma_scale * val, torch.zeros_like(val))
    return torch.stack([val, val, val, sig], dim=1).contiguous()


def world_box(dims: Tuple[int, int, int], zooms=(1.0, 1.0, 1.0)):
    """World scaling of the reference loader (inr/viewer/brats_viewer.py:204-210):
    voxelSize = zooms * float32(1.8/max_dim), volMin = -0.5 * voxelSize * dims (all float32)."""
    import numpy as np
    d = np.asarray(dims, dtype=np.uint32)
    scale = np.float32(1.8 / float(max(d)))
    vs = (np.asarray(zooms, dtype=np.float32) * scale).astype(np.float32)
    vmin = -0.5 * (vs * d.astype(np.float32))
    return vs, vmin.astype(np.float32)
