"""The C ABI: the library loads, exports every symbol include/mrt.h declares, and the
MrtParams layout begins with the reference's 368-byte `struct Params` cbuffer
(inr/viewer/brats_rt.slang:12-31).  No GPU needed (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

from mri_raytracer_b200 import RenderParams, _lib
from mri_raytracer_b200._lib import MrtParams, MrtSlabParams, PROTOTYPES

ROOT = Path(__file__).resolve().parent.parent


def _declared_functions():
    src = (ROOT / "include" / "mrt.h").read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrt_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(built_lib):
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(built_lib, n), f"libmrt.so lacks {n}"
        assert n in PROTOTYPES, f"_lib.py does not bind {n}"
    assert sorted(PROTOTYPES) == names
    assert built_lib.mrt_version() >= 200
    assert built_lib.mrt_last_error() is not None


def test_params_layout_matches_reference_cbuffer(built_lib):
    assert built_lib.mrt_sizeof_params() == C.sizeof(MrtParams) == 432
    assert built_lib.mrt_sizeof_slab_params() == C.sizeof(MrtSlabParams)
    off = {f[0]: getattr(MrtParams, f[0]).offset for f in MrtParams._fields_}
    # 16-byte rows of the cbuffer, in order
    want = dict(imageSize=0, fovY=8, eye=16, U=32, V=48, W=64, volMin=80, voxelSize=96, dims=112, stepSize=128,
                nearT=132, farT=136, bgColor=144, volEnabled=160, volWeight=176, ww=192, wl=196, intensityAlpha=200,
                gamma=208, gradBoost=212, gradScale=216, showSeg=224, showPred=228, lutColorAlpha=240, ortho=368,
                shardEnabled=400)
    for k, v in want.items():
        assert off[k] == v, (k, off[k], v)


def test_render_params_round_trip():
    P = RenderParams(imageSize=(33, 17), dims=(5, 6, 7), volEnabled=(1, 0, 1, 0), volWeight=(1, 2, 3, 4), tMode="accumulate")
    s = P.to_struct()
    assert tuple(s.imageSize) == (33, 17) and tuple(s.dims) == (5, 6, 7)
    assert tuple(s.volEnabled) == (1, 0, 1, 0) and s.tMode == 1
    assert abs(s.lutColorAlpha[2][1] - 0.8) < 1e-7 and s.lutColorAlpha[3][0] == 1.0


def test_host_side_helpers_need_no_gpu(built_lib):
    py, pz = C.c_int64(), C.c_int64()
    for Cn, S in ((1, 32), (2, 16), (4, 8)):
        built_lib.mrt_packed_layout(Cn, 240, 240, 155, C.byref(py), C.byref(pz))
        assert py.value >= 240 and pz.value >= py.value * 240
        assert py.value % S == S // 4 and pz.value % S == S // 2
        assert built_lib.mrt_packed_volume_bytes(Cn, 240, 240, 155) == pz.value * 155 * 4 * Cn
    assert built_lib.mrt_brick_count(240, 240, 155) == 30 * 30 * 20
    assert built_lib.mrt_packed_volume_bytes(5, 8, 8, 8) == 0
    # workspace of the one-call training step: host arithmetic only, grows with the views, 0 for an invalid request
    from mri_raytracer_b200 import RenderParams
    P = RenderParams(imageSize=(64, 48), dims=(40, 36, 28), tfMode=1)
    s = P.to_struct()
    w1 = built_lib.mrt_train_step_workspace_bytes(C.byref(s), 1, 64)
    w3 = built_lib.mrt_train_step_workspace_bytes(C.byref(s), 3, 64)
    assert w1 > 2 * built_lib.mrt_packed_volume_bytes(1, 40, 36, 28) and w3 > w1 and w1 % 256 == 0
    assert built_lib.mrt_train_step_workspace_bytes(C.byref(s), 0, 64) == 0


def test_packed_params_equal_field_by_field_ctypes():
    """RenderParams.to_struct packs the 432 bytes in one struct.pack call; this rebuilds the same
    struct field by field through ctypes and compares the bytes (layout, order, rounding)."""
    import numpy as np
    from dataclasses import replace
    rng = np.random.default_rng(0)
    base = RenderParams(imageSize=(321, 123), dims=(31, 17, 9))
    cases = [base,
             replace(base, eye=tuple(rng.normal(size=3)), U=tuple(rng.normal(size=3)), V=tuple(rng.normal(size=3)),
                     W=tuple(rng.normal(size=3)), volMin=(-0.3, -0.2, -0.1), voxelSize=(0.01, 0.02, 0.03),
                     stepSize=0.0123, nearT=0.4, farT=7.5, bgColor=(0.1, 0.2, 0.3), volEnabled=(1, 0, 1, 0),
                     volWeight=(0.5, 1.5, 2.5, 3.5), ww=0.7, wl=0.3, intensityAlpha=1.7, gamma=2.2, showSeg=1, showPred=1,
                     ortho=1, orthoHalfHeight=0.77, ertThreshold=0.02, maxSteps=99, tMode="accumulate", alphaMode=1,
                     skipEmpty=0, tfMode=1, volDtype=1),
             replace(base, shard=((1, 2, 3), (20, 10, 7)))]
    for P in cases:
        s = MrtParams()
        s.imageSize[:] = [int(v) for v in P.imageSize]; s.fovY = P.fovY
        for k in ("eye", "U", "V", "W", "volMin", "voxelSize", "bgColor"):
            getattr(s, k)[:] = [float(np.float32(v)) for v in getattr(P, k)]
        s.dims[:] = [int(v) for v in P.dims]
        s.stepSize, s.nearT, s.farT = P.stepSize, P.nearT, P.farT
        s.volEnabled[:] = [int(bool(v)) for v in P.volEnabled]; s.volWeight[:] = list(P.volWeight)
        s.ww, s.wl, s.intensityAlpha = P.ww, P.wl, P.intensityAlpha
        s.gamma, s.gradBoost, s.gradScale = P.gamma, P.gradBoost, P.gradScale
        s.showSeg, s.showPred = int(bool(P.showSeg)), int(bool(P.showPred))
        lut = np.asarray(P.lutColorAlpha, dtype=np.float32)
        for i in range(8):
            s.lutColorAlpha[i][:] = [float(v) for v in lut[i]]
        s.ortho, s.orthoHalfHeight, s.ertThreshold, s.maxSteps = int(bool(P.ortho)), P.orthoHalfHeight, P.ertThreshold, P.maxSteps
        s.tMode, s.alphaMode = (0 if P.tMode == "indexed" else 1), int(bool(P.alphaMode))
        s.skipEmpty, s.tfMode, s.volDtype = int(bool(P.skipEmpty)), int(bool(P.tfMode)), P.volDtype
        if P.shard is not None:
            s.shardEnabled = 1; s.shardLo[:] = list(P.shard[0]); s.shardHi[:] = list(P.shard[1])
        assert bytes(P.to_struct()) == bytes(s)
    assert C.sizeof(_lib.MrtCamera) == 64


def test_checkpoint_plan_is_host_only(built_lib):
    """mrt_checkpoint_plan: segments cover the longest possible ray (box diagonal / step), default 32
    slots per segment, at most 64 segments (the segment grows instead)."""
    from mri_raytracer_b200 import api
    from scenes import framed_params
    P = framed_params((256, 256, 256), 512, 512)
    S, n = api.checkpoint_plan(P)
    assert S == 32 and n * S >= 256 * 3 ** 0.5 / 0.5 and n <= 64
    S4, n4 = api.checkpoint_plan(P, 4)
    assert n4 <= 64 and S4 % 8 == 0 and n4 * S4 >= 887
    P5 = framed_params((1024, 1024, 1024), 64, 64)
    S5, n5 = api.checkpoint_plan(P5)
    assert n5 <= 64 and n5 * S5 >= 1024 * 3 ** 0.5 / 0.5
    assert built_lib.mrt_checkpoint_bytes(512, 512, 2, n) == (n - 1) * 2 * 512 * 512 * 16
    assert built_lib.mrt_half_tile_count(37, 29) == 2 * 5 * 4
    assert built_lib.mrt_backward_scratch_bytes(64, 64, 1, 256, 3) >= 64 * 256 * 32 + 128 * 3 * 8
