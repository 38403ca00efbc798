"""RenderParams — the operator's config block.

Field names, meaning and defaults are the reference's: ``struct Params``
(inr/viewer/brats_rt.slang:12-31) as filled every frame by the viewer
(inr/viewer/brats_viewer.py:405-426; defaults :112,126-135,138-144).  The tail fields are
the extensions SURVEY.md §8 defines (ortho camera, indexed stepping, LUT transfer function,
empty-space skipping); their defaults keep reference behaviour except ``tMode`` (indexed,
SURVEY Q4) and ``skipEmpty`` (exact, never changes the image).
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field, replace
from typing import Sequence, Tuple

import numpy as np

from ._lib import MrtParams, MrtSlabParams

T_INDEXED, T_ACCUMULATE = "indexed", "accumulate"

# byte layout of struct MrtParams (include/mrt.h), row by row; checked against ctypes in tests/test_abi.py
_PARAMS_PACK = struct.Struct("<2I2f" + "4f" * 6 + "4I" + "4f" + "4f" + "4I" + "4f" + "4f" + "4f" + "4I" + "32f"
                             + "I2fI" + "4I" + "8I")


def default_label_lut() -> np.ndarray:
    """The viewer's fixed 8-entry label LUT (inr/viewer/brats_viewer.py:138-144)."""
    lut = np.zeros((8, 4), dtype=np.float32)
    lut[1] = (0.0, 0.4, 1.0, 0.9)     # NCR/NET
    lut[2] = (0.0, 0.8, 0.0, 0.7)     # edema
    lut[3] = (1.0, 0.1, 0.1, 0.9)     # enhancing
    lut[4] = (1.0, 0.1, 0.1, 0.9)     # BraTS label 4 = enhancing (older numbering)
    return lut


def _v3(x) -> Tuple[float, float, float]:
    """3 floats rounded to float32 (what the kernel will see)."""
    if isinstance(x, np.ndarray) and x.dtype == np.float32 and x.shape == (3,):
        return (float(x[0]), float(x[1]), float(x[2]))
    a = np.asarray(x, dtype=np.float32).reshape(3)
    return (float(a[0]), float(a[1]), float(a[2]))


@dataclass
class RenderParams:
    imageSize: Tuple[int, int] = (512, 512)          # (W, H)
    fovY: float = math.radians(70.0)
    eye: Sequence[float] = (0.0, 0.0, -3.0)
    U: Sequence[float] = (1.0, 0.0, 0.0)
    V: Sequence[float] = (0.0, 1.0, 0.0)
    W: Sequence[float] = (0.0, 0.0, 1.0)
    volMin: Sequence[float] = (-0.9, -0.9, -0.9)
    voxelSize: Sequence[float] = (0.0075, 0.0075, 0.0075)
    dims: Tuple[int, int, int] = (240, 240, 155)     # (X, Y, Z)
    stepSize: float = 0.05
    nearT: float = 0.0
    farT: float = 0.0
    bgColor: Sequence[float] = (0.0, 0.0, 0.0)
    volEnabled: Sequence[int] = (1, 1, 1, 1)
    volWeight: Sequence[float] = (1.0, 1.0, 1.0, 1.0)
    ww: float = 1.0
    wl: float = 0.5
    intensityAlpha: float = 0.4
    gamma: float = 1.0
    gradBoost: float = 1.5      # declared by the shader, never read
    gradScale: float = 1.0      # declared by the shader, never read
    showSeg: int = 0
    showPred: int = 0
    lutColorAlpha: np.ndarray = field(default_factory=default_label_lut)
    # ---- extensions ----
    ortho: int = 0
    orthoHalfHeight: float = 1.0
    ertThreshold: float = 0.01
    maxSteps: int = 0
    tMode: str = T_INDEXED
    alphaMode: int = 0
    skipEmpty: int = 1
    tfMode: int = 0             # set by render(): 1 when a LUT tensor is passed
    shard: object = None        # ((lox,loy,loz), (hix,hiy,hiz)) voxel range of a sort-last sub-box, or None
    volDtype: int = 0           # 0 fp32 voxels, 1 fp16 single-channel (set by api.Volume)

    def __setattr__(self, k, v):
        d = self.__dict__
        d[k] = v
        if "_struct" in d or "_derived" in d:        # the packed C struct and derived copies are cached per instance
            d.pop("_struct", None)
            d.pop("_derived", None)

    def derived(self, key, make):
        """Memoised derived copy (``make(self)``), dropped when any field of this instance changes:
        a frame loop that re-submits the same parameter block does not pay ``dataclasses.replace``
        (34 fields, ~6 us each time) several times per frame."""
        cache = self.__dict__.get("_derived")
        if cache is None:
            cache = {}
            self.__dict__["_derived"] = cache
        out = cache.get(key)
        if out is None:
            out = make(self)
            cache[key] = out
        return out

    def validate(self):
        W, H = self.imageSize
        if W <= 0 or H <= 0:
            raise ValueError(f"imageSize {self.imageSize} must be positive")
        if not (self.ww > 0):
            raise ValueError("ww must be > 0 (the reference's slider minimum is 0.01)")
        if not (self.stepSize > 0):
            raise ValueError("stepSize must be > 0")
        if any(int(d) < 2 for d in self.dims):
            raise ValueError(f"dims {self.dims} must be >= 2 per axis")
        if self.tMode not in (T_INDEXED, T_ACCUMULATE):
            raise ValueError(f"tMode {self.tMode!r} unknown")
        if np.asarray(self.lutColorAlpha).shape != (8, 4):
            raise ValueError("lutColorAlpha must be [8,4]")

    def with_projection_of(self, cam) -> "RenderParams":
        """For batched launches (the cameras travel separately): ``self`` if its projection already is
        the camera's, else :meth:`with_camera`."""
        proj = (float(cam.fovY), int(cam.ortho), float(cam.ortho_half_height))
        if (float(self.fovY), int(self.ortho), float(self.orthoHalfHeight)) == proj:
            return self
        return self.derived(("projection", proj), lambda p: replace(p, fovY=proj[0], ortho=proj[1], orthoHalfHeight=proj[2]))

    def with_camera(self, cam) -> "RenderParams":
        """Copy with eye/U/V/W/fov/ortho taken from a :class:`camera.Camera`."""
        return replace(self, eye=_v3(cam.eye), U=_v3(cam.U), V=_v3(cam.V), W=_v3(cam.W), fovY=float(cam.fovY),
                       ortho=int(cam.ortho), orthoHalfHeight=float(cam.ortho_half_height))

    def to_struct(self) -> MrtParams:
        """The C struct, packed in one go (filling 432 bytes field by field through ctypes costs
        ~45 us — as much as launching a kernel; ``struct.pack`` + ``from_buffer_copy`` takes ~8)."""
        cached = self.__dict__.get("_struct")
        if cached is not None:
            return cached
        self.validate()
        f3 = lambda v: (float(v[0]), float(v[1]), float(v[2]))
        lut = np.asarray(self.lutColorAlpha, dtype=np.float32).reshape(32).tolist()
        slo = shi = (0, 0, 0)
        if self.shard is not None:
            slo, shi = tuple(int(v) for v in self.shard[0]), tuple(int(v) for v in self.shard[1])
            for i in range(3):
                if not (0 <= slo[i] < shi[i] <= int(self.dims[i]) - 1):
                    raise ValueError(f"shard {self.shard} outside the volume's cell range")
        raw = _PARAMS_PACK.pack(
            int(self.imageSize[0]), int(self.imageSize[1]), float(self.fovY), 0.0,
            *f3(self.eye), 0.0, *f3(self.U), 0.0, *f3(self.V), 0.0, *f3(self.W), 0.0,
            *f3(self.volMin), 0.0, *f3(self.voxelSize), 0.0,
            int(self.dims[0]), int(self.dims[1]), int(self.dims[2]), 0,
            float(self.stepSize), float(self.nearT), float(self.farT), 0.0,
            *f3(self.bgColor), 0.0,
            *(1 if int(e) else 0 for e in self.volEnabled[:4]),
            *(float(w) for w in self.volWeight[:4]),
            float(self.ww), float(self.wl), float(self.intensityAlpha), 0.0,
            float(self.gamma), float(self.gradBoost), float(self.gradScale), 0.0,
            int(bool(self.showSeg)), int(bool(self.showPred)), 0, 0,
            *lut,
            int(bool(self.ortho)), float(self.orthoHalfHeight), float(self.ertThreshold), int(self.maxSteps),
            0 if self.tMode == T_INDEXED else 1, int(bool(self.alphaMode)), int(bool(self.skipEmpty)),
            int(bool(self.tfMode)),
            1 if self.shard is not None else 0, *slo, *shi, int(self.volDtype))
        st = MrtParams.from_buffer_copy(raw)
        self.__dict__["_struct"] = st                # callers only read it (C.byref)
        return st



@dataclass
class SlabParams:
    """``struct Params`` of scripts/volumeRendering/volume_render.slang:9-21
    (filled at scripts/volumeRendering/app.py:336-347: fov 72 deg, 64 steps, near 4.3, far 4.4)."""
    imageSize: Tuple[int, int] = (512, 512)
    fovY: float = math.radians(72.0)
    stepCount: float = 64.0
    nearPlane: float = 4.3
    farPlane: float = 4.4
    eye: Sequence[float] = (0.0, 0.0, -4.2)
    U: Sequence[float] = (1.0, 0.0, 0.0)
    V: Sequence[float] = (0.0, 1.0, 0.0)
    W: Sequence[float] = (0.0, 0.0, 1.0)
    volDim: Tuple[int, int, int] = (180, 216, 180)

    def with_camera(self, cam) -> "SlabParams":
        return replace(self, eye=_v3(cam.eye), U=_v3(cam.U), V=_v3(cam.V), W=_v3(cam.W))

    def to_struct(self) -> MrtSlabParams:
        s = MrtSlabParams()
        s.imageSize[0], s.imageSize[1] = int(self.imageSize[0]), int(self.imageSize[1])
        s.fovY, s.stepCount = float(self.fovY), float(self.stepCount)
        s.nearPlane, s.farPlane = float(self.nearPlane), float(self.farPlane)
        for name in ("eye", "U", "V", "W"):
            arr = getattr(s, name)
            for i, v in enumerate(_v3(getattr(self, name))):
                arr[i] = v
        for i in range(3):
            s.volDim[i] = int(self.volDim[i])
        return s
