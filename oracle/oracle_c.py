"""ctypes wrapper of the scalar C oracle (oracle/oracle_c.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SO = HERE / "_build" / "liboracle_c.so"


class OParams(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("fovY", C.c_float),
                ("eye", C.c_float * 3), ("U", C.c_float * 3), ("V", C.c_float * 3), ("Wv", C.c_float * 3),
                ("volMin", C.c_float * 3), ("voxelSize", C.c_float * 3), ("dims", C.c_int32 * 3),
                ("stepSize", C.c_float), ("nearT", C.c_float), ("farT", C.c_float), ("bgColor", C.c_float * 3),
                ("volEnabled", C.c_int32 * 4), ("volWeight", C.c_float * 4),
                ("ww", C.c_float), ("wl", C.c_float), ("intensityAlpha", C.c_float), ("gamma", C.c_float),
                ("showSeg", C.c_int32), ("showPred", C.c_int32), ("lut", (C.c_float * 4) * 8),
                ("ortho", C.c_int32), ("orthoHalfHeight", C.c_float), ("ertThreshold", C.c_float),
                ("maxSteps", C.c_int32), ("tMode", C.c_int32), ("alphaMode", C.c_int32), ("useTf", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not SO.exists():
            subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
        _lib = C.CDLL(str(SO))
        _lib.oracle_c_sizeof_params.restype = C.c_int
        assert _lib.oracle_c_sizeof_params() == C.sizeof(OParams)
        _lib.oracle_c_render.restype = C.c_int
        _lib.oracle_c_render.argtypes = [C.POINTER(OParams), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int]
    return _lib


def _get(P, name, default=None):
    return P.get(name, default) if isinstance(P, dict) else getattr(P, name, default)


def to_struct(P, use_tf: bool) -> OParams:
    s = OParams()
    s.W, s.H = (int(v) for v in _get(P, "imageSize"))
    s.fovY = float(_get(P, "fovY"))
    for dst, src in (("eye", "eye"), ("U", "U"), ("V", "V"), ("Wv", "W"), ("volMin", "volMin"),
                     ("voxelSize", "voxelSize"), ("bgColor", "bgColor")):
        a = np.asarray(_get(P, src), dtype=np.float32).reshape(3)
        for i in range(3):
            getattr(s, dst)[i] = float(a[i])
    for i in range(3):
        s.dims[i] = int(_get(P, "dims")[i])
    s.stepSize, s.nearT, s.farT = float(_get(P, "stepSize")), float(_get(P, "nearT", 0.0)), float(_get(P, "farT", 0.0))
    for i in range(4):
        s.volEnabled[i] = int(bool(_get(P, "volEnabled")[i]))
        s.volWeight[i] = float(_get(P, "volWeight")[i])
    s.ww, s.wl = float(_get(P, "ww")), float(_get(P, "wl"))
    s.intensityAlpha, s.gamma = float(_get(P, "intensityAlpha")), float(_get(P, "gamma", 1.0))
    s.showSeg, s.showPred = int(bool(_get(P, "showSeg", 0))), int(bool(_get(P, "showPred", 0)))
    lut = np.asarray(_get(P, "lutColorAlpha", np.zeros((8, 4))), dtype=np.float32)
    for i in range(8):
        for j in range(4):
            s.lut[i][j] = float(lut[i, j])
    s.ortho = int(bool(_get(P, "ortho", 0)))
    s.orthoHalfHeight = float(_get(P, "orthoHalfHeight", 1.0))
    s.ertThreshold = float(_get(P, "ertThreshold", 0.01))
    s.maxSteps = int(_get(P, "maxSteps", 0) or 0)
    s.tMode = 0 if _get(P, "tMode", "indexed") == "indexed" else 1
    s.alphaMode = int(bool(_get(P, "alphaMode", 0)))
    s.useTf = int(use_tf)
    return s


def render(volume, P, tf=None, labels=None, preds=None, pixels=None, return_aux=False, threads=1):
    """volume [C,Z,Y,X] float32 array-like -> rgba float32 [H,W,4] (or [N,4] with pixels=(px,py))."""
    vol = np.ascontiguousarray(np.asarray(volume, dtype=np.float32))
    Cn = vol.shape[0]
    tfa = None if tf is None else np.ascontiguousarray(np.asarray(tf, dtype=np.float32))
    la = None if labels is None else np.ascontiguousarray(np.asarray(labels, dtype=np.int32))
    pa = None if preds is None else np.ascontiguousarray(np.asarray(preds, dtype=np.int32))
    s = to_struct(P, tfa is not None)
    if la is None:
        s.showSeg = 0
    if pa is None:
        s.showPred = 0
    if pixels is None:
        n, px, py = s.W * s.H, None, None
    else:
        px = np.ascontiguousarray(np.asarray(pixels[0], dtype=np.int32))
        py = np.ascontiguousarray(np.asarray(pixels[1], dtype=np.int32))
        n = px.size
    out = np.empty((n, 4), dtype=np.float32)
    counts = np.empty((n, 2), dtype=np.int32)
    T = np.empty((n,), dtype=np.float32)
    p = lambda a: None if a is None else a.ctypes.data
    rc = lib().oracle_c_render(C.byref(s), p(vol), Cn, p(tfa), 0 if tfa is None else tfa.shape[0], p(la), p(pa),
                               p(px), p(py), n, p(out), p(counts), p(T), int(threads))
    assert rc == 0
    if pixels is None:
        out, counts, T = out.reshape(s.H, s.W, 4), counts.reshape(s.H, s.W, 2), T.reshape(s.H, s.W)
    if return_aux:
        return out, dict(n_samples=counts[..., 0], n_taken=counts[..., 1], T=T)
    return out
