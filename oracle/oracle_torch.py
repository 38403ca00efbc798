"""CPU-torch oracle for the volume ray-march hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement of the reference's Slang compute shaders; it is
the checker, never the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product path (``mri_raytracer_b200``) never imports anything from ``oracle/``.

PARITY STATUS: **parity unpinned** for the shader arithmetic.  The reference
(klukaszek/MRI-RayTracer) ships no tests, golden images or CPU renderer, and its
shaders need ``slangpy`` + a window to run (absent here, no network).  What IS
pinned by reference-executed golden vectors: the orbital cameras
(``tests/golden/camera_*.json``) and the NIfTI ingest normalisation/flatten
(``tests/golden/ingest_*.npz``); see ``tests/golden/make_golden.py``.  The shader
restatement below is validated by analytic known-answer tests
(``tests/test_oracle_known_answers.py``) and cross-checked against a second,
independently written scalar C restatement (``oracle/oracle_c.c``).

Every function cites the reference lines it follows (paths relative to
/root/reference).

Conventions
-----------
* ``P`` is any object exposing the reference's ``struct Params`` field names
  (``inr/viewer/brats_rt.slang:12-31``): imageSize, fovY, eye, U, V, W, volMin,
  voxelSize, dims, stepSize, nearT, farT, bgColor, volEnabled, volWeight, ww, wl,
  intensityAlpha, gamma, showSeg, showPred, lutColorAlpha — plus the extensions
  SURVEY.md §8 defines: ortho, orthoHalfHeight, ertThreshold, maxSteps, tMode,
  alphaMode.
* Volume is ``[C, Z, Y, X]`` (x fastest), exactly the reference's flatten
  (``inr/viewer/brats_viewer.py:64``).
* All arithmetic is done one IEEE operation at a time in ``dtype`` (fp32 by
  default, fp64 for gradcheck): torch's element-wise CPU ops never contract to FMA.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

ERT_DEFAULT = 0.01  # brats_rt.slang:117  `T > 0.01`


# --------------------------------------------------------------------------- params
def _get(P, name, default=None):
    if isinstance(P, dict):
        return P.get(name, default)
    return getattr(P, name, default)


def focal_from_fov(fovY: float) -> np.float32:
    """f = 1/tan(fovY/2)  (brats_rt.slang:41).

    Deliberate, documented deviation: the shader evaluates ``tan`` per thread in
    fp32 with backend-defined precision.  The oracle (and the product's host code)
    evaluate it once in float64 and round to float32, so every implementation sees
    the same 32 bits.
    """
    return np.float32(1.0 / math.tan(0.5 * float(fovY)))


def _vec(x, dtype):
    return torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=torch.float64).to(dtype)


def _f32in(x, dtype):
    """Scalars enter the arithmetic as float32 bit patterns (the cbuffer is fp32)."""
    return torch.tensor(float(np.float32(x)), dtype=dtype)


# --------------------------------------------------------------------------- A2 / A3
def make_rays(P, px: torch.Tensor, py: torch.Tensor, dtype=torch.float32):
    """Primary rays for integer pixel coords (px = column, py = row).

    Pinhole: ``makePrimary`` brats_rt.slang:36-46 (same code raymarch.slang:45-58).
    Orthographic: no reference code; SURVEY.md §8(a) row A3 definition.
    Returns (o[N,3], d[N,3]).
    """
    Wd, Hd = _get(P, "imageSize")
    dimx = torch.tensor(float(Wd), dtype=dtype)
    dimy = torch.tensor(float(Hd), dtype=dtype)
    half = torch.tensor(0.5, dtype=dtype)
    two = torch.tensor(2.0, dtype=dtype)
    one = torch.tensor(1.0, dtype=dtype)
    ndcx = (px.to(dtype) + half) / dimx                      # :39
    ndcy = (py.to(dtype) + half) / dimy
    uvx = ndcx * two - one                                    # :40
    uvy = ndcy * two - one
    aspect = dimx / torch.maximum(one, dimy)                  # :42
    eye = _vec(_get(P, "eye"), dtype)
    U = _vec(_get(P, "U"), dtype)
    V = _vec(_get(P, "V"), dtype)
    Wv = _vec(_get(P, "W"), dtype)
    if int(_get(P, "ortho", 0)):
        halfH = _f32in(_get(P, "orthoHalfHeight"), dtype)
        halfW = aspect * halfH
        ax = uvx * halfW
        ay = -(uvy * halfH)
        o = (eye[None, :] + ax[:, None] * U[None, :]) + ay[:, None] * V[None, :]
        d = Wv[None, :].expand(px.shape[0], 3).contiguous()
        return o, d
    f = torch.tensor(float(focal_from_fov(_get(P, "fovY"))), dtype=dtype)  # :41
    cx = uvx * aspect / f                                     # :43
    cy = -uvy / f
    cz = torch.ones_like(cx)
    inv = torch.sqrt((cx * cx + cy * cy) + cz * cz)           # normalize(): v / sqrt(dot)
    cx, cy, cz = cx / inv, cy / inv, cz / inv
    rd = (cx[:, None] * U[None, :] + cy[:, None] * V[None, :]) + cz[:, None] * Wv[None, :]  # :44
    n = torch.sqrt((rd[:, 0] * rd[:, 0] + rd[:, 1] * rd[:, 1]) + rd[:, 2] * rd[:, 2])
    d = rd / n[:, None]
    o = eye[None, :].expand(px.shape[0], 3).contiguous()
    return o, d


# --------------------------------------------------------------------------- A4
def box_bounds(P, dtype=torch.float32):
    """bmin / bmax (brats_rt.slang:92-93)."""
    bmin = _vec(_get(P, "volMin"), dtype)
    vs = _vec(_get(P, "voxelSize"), dtype)
    dims = _vec(_get(P, "dims"), dtype)
    bmax = bmin + vs * dims
    return bmin, bmax, vs


def clip_rays(P, o, d, dtype=torch.float32):
    """De-zero + slab test + near/far clamp (brats_rt.slang:95-99, 48-57, 107-109).

    Returns (t0, t1, hit) where hit already includes the ``t1 <= t0`` rejection.
    """
    bmin, bmax, _ = box_bounds(P, dtype)
    eps = torch.tensor(1e-6, dtype=dtype)
    dz = torch.where(d.abs() < eps, eps.expand_as(d), d)      # sign dropped on purpose (:96-98)
    rcp = torch.tensor(1.0, dtype=dtype) / dz                 # :99
    ta = (bmin[None, :] - o) * rcp                            # :50
    tb = (bmax[None, :] - o) * rcp                            # :51
    tsm = torch.minimum(ta, tb)
    tbg = torch.maximum(ta, tb)
    tmin = torch.maximum(torch.maximum(tsm[:, 0], tsm[:, 1]), tsm[:, 2])   # :54
    tmax = torch.minimum(torch.minimum(tbg[:, 0], tbg[:, 1]), tbg[:, 2])   # :55
    zero = torch.tensor(0.0, dtype=dtype)
    hit = tmax >= torch.maximum(tmin, zero)                   # :56
    nearT = _f32in(_get(P, "nearT", 0.0), dtype)
    farT = _f32in(_get(P, "farT", 0.0), dtype)
    t0 = torch.maximum(tmin, torch.maximum(zero, nearT))      # :107
    t1 = torch.where(farT > 0, torch.minimum(tmax, farT), tmax)  # :108
    hit = hit & ~(t1 <= t0)                                   # :109
    return t0, t1, hit


# --------------------------------------------------------------------------- A5 / A7
def _lerp(a, b, t):
    return a + t * (b - a)                                    # HLSL lerp


def sample_linear(vol: torch.Tensor, pIdx: torch.Tensor, dims_xyz):
    """``sampleLinear`` brats_rt.slang:60-76.  vol is [Z,Y,X]; pIdx is [N,3] (x,y,z)."""
    dtype = pIdx.dtype
    X, Y, Z = (int(v) for v in dims_xyz)
    if dtype == torch.float32:   # float3(dims) - 1.001 evaluated in fp32, like the shader
        hi = torch.tensor([float(np.float32(n) - np.float32(1.001)) for n in (X, Y, Z)], dtype=dtype)
    else:
        hi = torch.tensor([X - 1.001, Y - 1.001, Z - 1.001], dtype=dtype)
    q = torch.minimum(torch.maximum(pIdx, torch.zeros((), dtype=dtype)), hi[None, :])  # :62
    i = torch.floor(q)                                        # :63
    f = q - i                                                 # :64
    ii = i.to(torch.int64)
    sY, sZ = X, X * Y                                         # :66
    b = ii[:, 0] + ii[:, 1] * sY + ii[:, 2] * sZ              # :67
    flat = vol.reshape(-1)
    c000 = flat[b]; c100 = flat[b + 1]                        # :69
    c010 = flat[b + sY]; c110 = flat[b + sY + 1]              # :70
    c001 = flat[b + sZ]; c101 = flat[b + sZ + 1]              # :71
    c011 = flat[b + sZ + sY]; c111 = flat[b + sZ + sY + 1]    # :72
    fx, fy, fz = f[:, 0], f[:, 1], f[:, 2]
    return _lerp(_lerp(_lerp(c000, c100, fx), _lerp(c010, c110, fx), fy),
                 _lerp(_lerp(c001, c101, fx), _lerp(c011, c111, fx), fy), fz)  # :74-75


def _round_half_away(x):
    return torch.sign(x) * torch.floor(x.abs() + 0.5)


def sample_label(lab: torch.Tensor, pIdx: torch.Tensor, dims_xyz):
    """``sampleLabel`` brats_rt.slang:78-83 (round = half away from zero)."""
    dtype = pIdx.dtype
    X, Y, Z = (int(v) for v in dims_xyz)
    hi = torch.tensor([X - 1.0, Y - 1.0, Z - 1.0], dtype=dtype)
    q = torch.minimum(torch.maximum(pIdx, torch.zeros((), dtype=dtype)), hi[None, :])
    ii = _round_half_away(q).to(torch.int64)
    idx = ii[:, 0] + ii[:, 1] * X + ii[:, 2] * (X * Y)
    return lab.reshape(-1)[idx].to(torch.int64)


def tf_lookup(tf: torch.Tensor, val: torch.Tensor):
    """1D transfer-function LUT, linear interpolation.

    u = val*(N-1); j0 = floor(u); j1 = min(j0+1, N-1); rgba = lerp(tf[j0], tf[j1], u-j0).
    """
    N = tf.shape[0]
    u = val * torch.tensor(float(N - 1), dtype=val.dtype)
    j0f = torch.floor(u)
    fr = u - j0f
    j0 = j0f.to(torch.int64).clamp(0, N - 1)
    j1 = torch.clamp(j0 + 1, max=N - 1)
    a = tf[j0]
    b = tf[j1]
    return a + fr[:, None] * (b - a)


# --------------------------------------------------------------------------- march
def pixel_grid(P):
    Wd, Hd = _get(P, "imageSize")
    ys, xs = torch.meshgrid(torch.arange(Hd), torch.arange(Wd), indexing="ij")
    return xs.reshape(-1), ys.reshape(-1)


def sample_counts(P, t0, t1, dtype=torch.float32):
    """n = #{k >= 0 : t0 + k*dt < t1} for tMode='indexed' (monotone => a prefix)."""
    dt = _f32in(_get(P, "stepSize"), dtype)
    n = torch.ceil((t1 - t0) / dt).to(torch.int64).clamp(min=0)
    for _ in range(3):  # exact fix-up of the estimate under fp rounding
        tk = t0 + n.to(dtype) * dt
        n = torch.where(tk < t1, n + 1, n)
        tkm = t0 + (n - 1).to(dtype) * dt
        n = torch.where((n > 0) & ~(tkm < t1), n - 1, n)
    return n


def render(volume: torch.Tensor, P, tf: Optional[torch.Tensor] = None,
           labels: Optional[torch.Tensor] = None, preds: Optional[torch.Tensor] = None,
           pixels=None, dtype=torch.float32, force_steps: Optional[torch.Tensor] = None,
           chunk: int = 1 << 16, return_aux: bool = False, ray_delta=None, soft_occ: Optional[torch.Tensor] = None):
    """``brats_main`` (brats_rt.slang:85-168) restated on CPU.

    volume : [C,Z,Y,X] (C<=4) values; differentiable leaf allowed.
    tf     : None -> the reference's window/level intensity transfer function
             (:132-140: emission=val, sigma=val*intensityAlpha);
             [N,4] (r,g,b,sigma) -> the generalised 1D LUT (row A7).
    labels / preds : optional [Z,Y,X] integer volumes (gLabels / gPreds, :141-162).
    pixels : optional (px, py) int tensors to render a subset of rays; default all.
    force_steps : optional per-ray int tensor overriding the ERT decision (test aid:
             lets a test check that an ERT flip in the kernel was a justified tie).
    ray_delta : optional (do, dd), two [nray,3] tensors ADDED to the ray origins / directions after
             the clip and the sample times t_i are fixed — their autograd gradients are the
             dL/do, dL/dd of docs/DifferentiableRendering.md section 9 (:172-188, "x_i = o + t_i d
             with fixed t_i").
    soft_occ : optional [nbz,nby,nbx] continuous occupancy over 8^3-voxel bricks
             (docs/DifferentiableRendering.md section 11, :202-206: "continuous occupancy o(x) in [0,1]
             learned and used multiplicatively"): a sample whose trilinear base cell lies in brick b
             composites with sigma' = soft_occ[b] * sigma.  Maths-only in the reference: parity unpinned.
    Returns rgba [H,W,4] (or [N,4] with ``pixels``) and optionally aux dict with
    T, n_samples (clip count), n_taken (after ERT), ert_margin.
    """
    C = volume.shape[0]
    X, Y, Z = (int(v) for v in _get(P, "dims"))
    assert tuple(volume.shape[1:]) == (Z, Y, X), (volume.shape, (Z, Y, X))
    Wd, Hd = _get(P, "imageSize")
    if pixels is None:
        px, py = pixel_grid(P)
    else:
        px, py = pixels
    nray = px.shape[0]
    vol = volume.to(dtype)
    tfd = tf.to(dtype) if tf is not None else None

    bg = _vec(_get(P, "bgColor", (0, 0, 0)), dtype)
    dt = _f32in(_get(P, "stepSize"), dtype)
    ww = _f32in(_get(P, "ww", 1.0), dtype)
    wl = _f32in(_get(P, "wl", 0.5), dtype)
    ia = _f32in(_get(P, "intensityAlpha", 0.4), dtype)
    gamma = float(np.float32(_get(P, "gamma", 1.0)))
    en = [int(v) for v in _get(P, "volEnabled", (1, 1, 1, 1))]
    wt = [_f32in(v, dtype) for v in _get(P, "volWeight", (1, 1, 1, 1))]
    show_seg = int(_get(P, "showSeg", 0)) and labels is not None
    show_pred = int(_get(P, "showPred", 0)) and preds is not None
    lut8 = torch.as_tensor(np.asarray(_get(P, "lutColorAlpha", np.zeros((8, 4))), dtype=np.float32)).to(dtype)
    thr = _f32in(_get(P, "ertThreshold", ERT_DEFAULT), dtype)
    max_steps = int(_get(P, "maxSteps", 0) or 0)
    t_mode = _get(P, "tMode", "indexed")
    alpha_mode = int(_get(P, "alphaMode", 0))
    lo = wl - ww * torch.tensor(0.5, dtype=dtype)             # :132  (wl - ww*0.5)
    one = torch.tensor(1.0, dtype=dtype)
    pred_boost = torch.tensor(1.5, dtype=dtype)               # :158

    out = []
    aux_T, aux_n, aux_taken, aux_margin = [], [], [], []
    for s in range(0, nray, chunk):
        cpx, cpy = px[s:s + chunk], py[s:s + chunk]
        o, d = make_rays(P, cpx, cpy, dtype)
        t0, t1, hit = clip_rays(P, o, d, dtype)
        bmin, _, vs = box_bounds(P, dtype)
        n_all = torch.where(hit, sample_counts(P, t0, t1, dtype), torch.zeros_like(cpx))
        if max_steps > 0:
            n_all = n_all.clamp(max=max_steps)
        nr = cpx.shape[0]
        Ccol = bg[None, :].expand(nr, 3).clone()              # :111
        T = torch.ones(nr, dtype=dtype)                       # :112
        taken = torch.zeros(nr, dtype=torch.int64)
        margin = torch.full((nr,), float("inf"), dtype=torch.float64)
        hidx = torch.nonzero(hit).reshape(-1)
        if hidx.numel() > 0:
            ho, hd, ht0, ht1 = o[hidx], d[hidx], t0[hidx], t1[hidx]
            if ray_delta is not None:
                ho = ho + ray_delta[0][s:s + chunk][hidx].to(dtype)
                hd = hd + ray_delta[1][s:s + chunk][hidx].to(dtype)
            hn = n_all[hidx]
            hC = Ccol[hidx]
            hT = T[hidx]
            htaken = taken[hidx]
            hmargin = margin[hidx]
            hforce = force_steps.reshape(-1)[s:s + chunk][hidx] if force_steps is not None else None
            t_run = ht0.clone()
            kmax = int(hn.max().item())
            k = 0
            while True:
                if t_mode == "indexed":
                    if k >= kmax:
                        break
                    t = ht0 + torch.tensor(float(k), dtype=dtype) * dt
                    in_range = hn > k
                else:  # 'accumulate' : the reference's running sum  t += stepSize (:113,:164)
                    t = t_run
                    in_range = t < ht1
                    if max_steps > 0:
                        in_range = in_range & (k < max_steps)
                if hforce is not None:
                    active = in_range & (hforce > k)
                else:
                    active = in_range & (hT > thr)            # :117
                    m = ((hT.detach().to(torch.float64) / float(thr)) - 1.0).abs()
                    hmargin = torch.where(in_range, torch.minimum(hmargin, m), hmargin)
                if not bool(active.any()):                    # deactivation is permanent
                    break
                p = ho + t[:, None] * hd                      # :119
                pIdx = (p - bmin[None, :]) / vs[None, :]      # :120
                v = torch.zeros_like(t)
                wsum = torch.zeros((), dtype=dtype)
                for c in range(min(C, 4)):                    # :123-128
                    if en[c] != 0:
                        v = v + sample_linear(vol[c], pIdx, (X, Y, Z)) * wt[c]
                        wsum = wsum + wt[c]
                if float(wsum) > 0.0:                         # :130
                    v = v / wsum
                val = torch.clamp((v - lo) / ww, 0.0, 1.0)    # :132 saturate
                if gamma != 1.0:                              # :133 (pow(x,1)==x exactly)
                    val = torch.pow(val, torch.tensor(gamma, dtype=dtype))
                newC, newT = hC, hT
                so = None
                if soft_occ is not None:                      # brick of the sample's base cell (sampleLinear's floor, :62-63)
                    hi3 = torch.tensor([float(np.float32(n) - np.float32(1.001)) for n in (X, Y, Z)], dtype=dtype) \
                        if dtype == torch.float32 else torch.tensor([X - 1.001, Y - 1.001, Z - 1.001], dtype=dtype)
                    bi = torch.floor(torch.minimum(torch.maximum(pIdx.detach(), torch.zeros((), dtype=dtype)), hi3[None, :])).to(torch.int64) >> 3
                    so = soft_occ.to(dtype)[bi[:, 2], bi[:, 1], bi[:, 0]]
                if tfd is None:                               # :135-140
                    a = val * ia
                    if so is not None:
                        a = a * so
                    alpha = one - torch.exp(-a * dt)
                    gate = val > 0
                    alpha = torch.where(gate, alpha, torch.zeros_like(alpha))
                    newC = newC + ((alpha * newT) * val)[:, None]
                    newT = newT * (one - alpha)
                else:
                    rgba = tf_lookup(tfd, val)
                    sig = rgba[:, 3] if so is None else rgba[:, 3] * so
                    alpha = one - torch.exp(-sig * dt)
                    newC = newC + (alpha * newT)[:, None] * rgba[:, :3]
                    newT = newT * (one - alpha)
                if show_seg:                                  # :143-151
                    l = sample_label(labels, pIdx.detach(), (X, Y, Z))
                    ok = (l > 0) & (l < 8)
                    col = lut8[l.clamp(0, 7)]
                    alpha = one - torch.exp(-col[:, 3] * dt)
                    alpha = torch.where(ok, alpha, torch.zeros_like(alpha))
                    newC = newC + (alpha * newT)[:, None] * col[:, :3]
                    newT = newT * (one - alpha)
                if show_pred:                                 # :154-162
                    l = sample_label(preds, pIdx.detach(), (X, Y, Z))
                    ok = (l > 0) & (l < 8)
                    col = lut8[l.clamp(0, 7)]
                    alpha = one - torch.exp(-col[:, 3] * dt * pred_boost)
                    alpha = torch.where(ok, alpha, torch.zeros_like(alpha))
                    newC = newC + (alpha * newT)[:, None] * col[:, :3]
                    newT = newT * (one - alpha)
                hC = torch.where(active[:, None], newC, hC)
                hT = torch.where(active, newT, hT)
                htaken = htaken + active.to(torch.int64)
                if t_mode != "indexed":
                    t_run = torch.where(active, t_run + dt, t_run)   # :164
                k += 1
            Ccol = Ccol.index_copy(0, hidx, hC)
            T = T.index_copy(0, hidx, hT)
            taken = taken.index_copy(0, hidx, htaken)
            margin = margin.index_copy(0, hidx, hmargin)
        a_out = torch.ones(nr, dtype=dtype) if alpha_mode == 0 else (one - T)   # :167 alpha == 1
        out.append(torch.cat([Ccol, a_out[:, None]], dim=1))
        aux_T.append(T); aux_n.append(n_all); aux_taken.append(taken); aux_margin.append(margin)
    rgba = torch.cat(out, dim=0)
    if pixels is None:
        rgba = rgba.reshape(Hd, Wd, 4)
    if not return_aux:
        return rgba
    shp = (Hd, Wd) if pixels is None else (nray,)
    aux = dict(T=torch.cat(aux_T).reshape(shp), n_samples=torch.cat(aux_n).reshape(shp),
               n_taken=torch.cat(aux_taken).reshape(shp), ert_margin=torch.cat(aux_margin).reshape(shp))
    return rgba, aux


# --------------------------------------------------------------------------- slab variant
def sample_u8_trilinear(vol_u8: torch.Tensor, uvw: torch.Tensor, dims_xyz, dtype=torch.float32):
    """``sampleU8`` + ``sampleTrilinear`` (scripts/volumeRendering/volume_render.slang:28-65)."""
    X, Y, Z = (int(v) for v in dims_xyz)
    dm1 = torch.tensor([X - 1.0, Y - 1.0, Z - 1.0], dtype=dtype)
    xyz = torch.clamp(uvw, 0.0, 1.0) * dm1[None, :]           # :43-45
    p0f = torch.floor(xyz)
    p0 = p0f.to(torch.int64)
    d1 = torch.tensor([X - 1, Y - 1, Z - 1], dtype=torch.int64)
    p1 = torch.minimum(p0 + 1, d1[None, :])                   # :48
    t = xyz - p0f                                             # :49
    flat = vol_u8.reshape(-1)

    def S(ix, iy, iz):
        idx = ix + iy * X + iz * (X * Y)                      # :33
        return flat[idx].to(dtype) / torch.tensor(255.0, dtype=dtype)  # :38

    c000 = S(p0[:, 0], p0[:, 1], p0[:, 2]); c100 = S(p1[:, 0], p0[:, 1], p0[:, 2])
    c010 = S(p0[:, 0], p1[:, 1], p0[:, 2]); c110 = S(p1[:, 0], p1[:, 1], p0[:, 2])
    c001 = S(p0[:, 0], p0[:, 1], p1[:, 2]); c101 = S(p1[:, 0], p0[:, 1], p1[:, 2])
    c011 = S(p0[:, 0], p1[:, 1], p1[:, 2]); c111 = S(p1[:, 0], p1[:, 1], p1[:, 2])
    c00 = _lerp(c000, c100, t[:, 0]); c01 = _lerp(c001, c101, t[:, 0])   # :58-61
    c10 = _lerp(c010, c110, t[:, 0]); c11 = _lerp(c011, c111, t[:, 0])
    c0 = _lerp(c00, c10, t[:, 1]); c1 = _lerp(c01, c11, t[:, 1])          # :62-63
    return _lerp(c0, c1, t[:, 2])                                         # :64


def render_slab(vol_u8: torch.Tensor, P, pixels=None, dtype=torch.float32):
    """``volume_cs`` (scripts/volumeRendering/volume_render.slang:104-148).

    P fields: imageSize, fovY, stepCount, nearPlane, farPlane, eye, U, V, W, volDim.
    vol_u8 is [Z,Y,X] uint8 (one byte per voxel; the reference widens each to a u32
    lane, app.py:150-158 — a storage detail that does not change values).
    """
    Wd, Hd = _get(P, "imageSize")
    X, Y, Z = (int(v) for v in _get(P, "volDim"))
    if pixels is None:
        px, py = pixel_grid(P)
    else:
        px, py = pixels
    dimx = torch.tensor(float(Wd), dtype=dtype); dimy = torch.tensor(float(Hd), dtype=dtype)
    one = torch.tensor(1.0, dtype=dtype); two = torch.tensor(2.0, dtype=dtype)
    half = torch.tensor(0.5, dtype=dtype)
    invx = one / dimx; invy = one / dimy                      # :111
    uvx = (px.to(dtype) + half) * invx                        # :115
    uvy = (py.to(dtype) + half) * invy
    ndcx = uvx * two - one                                    # :116
    ndcy = one - uvy * two
    # :117 tan evaluated once in float64 then rounded (same deviation as focal_from_fov)
    th = torch.tensor(float(np.float32(math.tan(0.5 * float(_get(P, "fovY"))))), dtype=dtype)
    aspect = dimx / torch.maximum(one, dimy)                  # :118
    vx = ndcx * aspect * th                                   # :119
    vy = ndcy * th
    vz = torch.ones_like(vx)
    n = torch.maximum(torch.tensor(0.0, dtype=dtype), _f32in(_get(P, "nearPlane"), dtype))  # :120
    f = torch.maximum(n, _f32in(_get(P, "farPlane"), dtype))  # :121
    eye = _vec(_get(P, "eye"), dtype); U = _vec(_get(P, "U"), dtype)
    V = _vec(_get(P, "V"), dtype); Wv = _vec(_get(P, "W"), dtype)

    def world(dist):                                          # :122-123
        return ((eye[None, :] + U[None, :] * (vx * dist)[:, None]) + V[None, :] * (vy * dist)[:, None]) \
            + Wv[None, :] * (vz * dist)[:, None]

    wn, wf = world(n), world(f)
    sc = _f32in(_get(P, "stepCount"), dtype)
    steps = torch.maximum(one, sc)                            # :124,:130
    stepv = (wf - wn) / steps                                 # :124
    nsteps = int(np.uint32(float(steps)))                     # :134 (uint)steps
    accum = torch.zeros(px.shape[0], dtype=dtype)
    pos = wn.clone()
    alive = torch.ones(px.shape[0], dtype=torch.bool)
    scale = torch.tensor(4.0, dtype=dtype) / steps            # :140
    for _ in range(nsteps):
        inside = (pos < 1.0).all(dim=1) & (pos > -1.0).all(dim=1)          # :136
        do = alive & inside & (accum < 1.0)                                # :137
        if bool(do.any()):
            uvw = half * (pos + one)                                       # :139
            s = sample_u8_trilinear(vol_u8, uvw, (X, Y, Z), dtype) * scale # :140
            accum = torch.where(do, accum + (one - accum) * s, accum)      # :141
        pos = torch.where(alive[:, None], pos + stepv, pos)               # :143
        alive = alive & ~(accum > 0.995)                                   # :144
        if not bool(alive.any()):
            break
    rgba = torch.stack([accum, accum, accum, torch.ones_like(accum)], dim=1)  # :147
    if pixels is None:
        rgba = rgba.reshape(Hd, Wd, 4)
    return rgba


# --------------------------------------------------------------------------- helpers
def params(**kw):
    """Convenience: a namespace with the reference viewer's defaults
    (inr/viewer/brats_viewer.py:112,126-135)."""
    d = dict(imageSize=(64, 64), fovY=math.radians(70.0), eye=(0, 0, -3), U=(1, 0, 0), V=(0, 1, 0),
             W=(0, 0, 1), volMin=(-0.9, -0.9, -0.9), voxelSize=(0.1, 0.1, 0.1), dims=(18, 18, 18),
             stepSize=0.05, nearT=0.0, farT=0.0, bgColor=(0.0, 0.0, 0.0), volEnabled=(1, 1, 1, 1),
             volWeight=(1.0, 1.0, 1.0, 1.0), ww=1.0, wl=0.5, intensityAlpha=0.4, gamma=1.0,
             showSeg=0, showPred=0, lutColorAlpha=np.zeros((8, 4), np.float32),
             ortho=0, orthoHalfHeight=1.0, ertThreshold=ERT_DEFAULT, maxSteps=0, tMode="indexed",
             alphaMode=0)
    d.update(kw)
    return SimpleNamespace(**d)
