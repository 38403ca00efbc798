"""Generate golden vectors by EXECUTING the reference's own Python (numpy-only parts).

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The reference's renderers are Slang shaders that cannot run here (no slangpy/slangc), but
its host-side Python is importable once the two unavailable third-party modules are
stubbed: ``slangpy`` (never called by the functions we execute) and ``nibabel`` (replaced
by an in-memory fake that hands our synthetic arrays to the reference loader).  Everything
written below is the OUTPUT OF REFERENCE CODE:

* camera_arbitrary_up.json  — inr/viewer/camera.py  OrbitalCamera.get_basis/orbit/pan/zoom
* camera_yup.json           — scripts/raymarch/camera.py OrbitalCamera (same API, Y-up)
* ingest.npz                — inr/viewer/brats_viewer.py load_nifti_float / load_seg_uint /
                              BraTSViewer.load_dir (world scaling) / frame_volume
* bc4.npz                   — scripts/volumeRendering/app.py App._load_volume_bc4
* nifti_mask.npz            — scripts/volumeRendering/app.py App._load_nifti_mask

These files travel with the repo; the GPU box never reads /root/reference.
"""
from __future__ import annotations

import gzip
import importlib.util
import json
import math
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def _load(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------ cameras
def camera_cases():
    rng = np.random.default_rng(1234)
    cases = []
    ups = [None, (0, 1, 0), (0, 0, 1), (1, 0, 0), (0, -1, 0), (0, 0, -1), (-1, 0, 0)]
    for i in range(40):
        cases.append(dict(
            target=[float(np.float32(x)) for x in rng.uniform(-0.5, 0.5, 3)] if i % 3 else None,
            radius=float(rng.uniform(0.3, 6.0)), phi=float(rng.uniform(0.02, math.pi - 0.02)),
            theta=float(rng.uniform(-7.0, 7.0)), up=ups[i % len(ups)], fov_deg=float(rng.uniform(20, 110)),
            ops=[["orbit", float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1))],
                 ["zoom", float(rng.uniform(0.5, 1.5))],
                 ["pan", float(rng.uniform(-50, 50)), float(rng.uniform(-50, 50))]][: i % 4]))
    # degenerate / boundary states
    cases.append(dict(target=None, radius=2.0, phi=0.01, theta=0.0, up=(0, 1, 0), fov_deg=70.0, ops=[]))
    cases.append(dict(target=None, radius=2.0, phi=math.pi - 0.01, theta=1.0, up=(0, 0, 1), fov_deg=70.0, ops=[]))
    cases.append(dict(target=None, radius=3.0, phi=math.pi * 0.5, theta=0.0, up=None, fov_deg=55.0, ops=[]))
    cases.append(dict(target=None, radius=1e-7, phi=1.0, theta=0.3, up=None, fov_deg=55.0, ops=[]))   # fn < 1e-6
    return cases


def run_camera(cls, case, arbitrary_up: bool):
    kw = dict(initial_radius=case["radius"], initial_phi=case["phi"], initial_theta=case["theta"],
              min_radius=1e-9)
    if case["target"] is not None:
        kw["initial_target"] = np.array(case["target"], dtype=np.float32)
    if arbitrary_up and case["up"] is not None:
        kw["world_up"] = np.array(case["up"], dtype=np.float32)
    cam = cls(**kw)
    cam.set_fov_degrees(case["fov_deg"])
    for op in case["ops"]:
        if op[0] == "orbit":
            cam.orbit(op[1], op[2])
        elif op[0] == "zoom":
            cam.zoom(op[1])
        elif op[0] == "pan":
            if arbitrary_up:
                cam.pan(op[1], op[2], viewport_height=480.0)
            else:
                cam.pan(op[1], op[2])
    eye, right, up, fwd = cam.get_basis()
    f = lambda a: [float(np.float32(v)) for v in np.asarray(a).reshape(-1)]
    return dict(eye=f(eye), right=f(right), up=f(up), forward=f(fwd), eye_only=f(cam.get_eye_position()),
                target=f(cam.target), radius=float(cam.radius), phi=float(cam.phi), theta=float(cam.theta),
                dtypes=[str(np.asarray(a).dtype) for a in (eye, right, up, fwd)])


def make_cameras():
    cases = camera_cases()
    for fname, path, arb in (("camera_arbitrary_up.json", REF / "inr/viewer/camera.py", True),
                             ("camera_yup.json", REF / "scripts/raymarch/camera.py", False)):
        mod = _load("ref_camera_" + ("arb" if arb else "yup"), path)
        rows = [dict(case=c, out=run_camera(mod.OrbitalCamera, c, arb)) for c in cases]
        (OUT / fname).write_text(json.dumps(dict(source=str(path.relative_to(REF)), rows=rows), indent=1))
        print("wrote", fname, len(rows))


# ------------------------------------------------------------------ fake nibabel
class _FakeHeader:
    def __init__(self, zooms):
        self._z = tuple(float(z) for z in zooms)

    def get_zooms(self):
        return self._z


class _FakeImg:
    def __init__(self, data, zooms):
        self._d = np.asarray(data)
        self.header = _FakeHeader(zooms)

    def get_fdata(self, dtype=np.float64):
        return self._d.astype(dtype)


_REGISTRY: dict[str, _FakeImg] = {}


def _install_stubs():
    nib = types.ModuleType("nibabel")
    nib.load = lambda p: _REGISTRY[Path(str(p)).name]
    sys.modules["nibabel"] = nib
    sys.modules["slangpy"] = types.ModuleType("slangpy")


def synth_modality(rng, shape, scale):
    x = rng.gamma(2.0, 1.0, size=shape).astype(np.float32) * scale
    x[rng.random(shape) < 0.3] = 0.0            # skull-stripped background
    x.flat[rng.integers(0, x.size, 5)] = 50.0 * scale   # outliers above the 99.5th percentile
    return x


def make_ingest():
    _install_stubs()
    sys.path.insert(0, str(REF / "inr/viewer"))
    sys.modules.pop("camera", None)
    bv = _load("ref_brats_viewer", REF / "inr/viewer/brats_viewer.py")
    rng = np.random.default_rng(7)
    shape = (12, 10, 7)             # (X, Y, Z) as nibabel returns it
    zooms = (1.0, 1.0, 1.5)
    out = {}
    mods = {}
    for suf, scale in (("t1n", 300.0), ("t1c", 1.0), ("t2w", 0.01), ("t2f", 4000.0)):
        raw = synth_modality(rng, shape, scale)
        name = f"CASE-{suf}.nii.gz"
        _REGISTRY[name] = _FakeImg(raw, zooms)
        mods[suf] = raw
        linear, norm, dims, z = bv.load_nifti_float(Path("/nowhere") / name)
        out[f"raw_{suf}"] = raw
        out[f"linear_{suf}"] = linear
        out[f"norm_{suf}"] = norm
        out["dims"] = dims
        out["zooms"] = z
    const = np.full(shape, 3.0, dtype=np.float32)          # vmax <= vmin branch
    _REGISTRY["CONST-t1n.nii.gz"] = _FakeImg(const, zooms)
    lin_c, _, _, _ = bv.load_nifti_float(Path("/nowhere/CONST-t1n.nii.gz"))
    out["raw_const"] = const
    out["linear_const"] = lin_c
    seg = rng.integers(0, 4, size=shape).astype(np.float32) + rng.uniform(-0.2, 0.2, size=shape).astype(np.float32)
    _REGISTRY["CASE-seg.nii.gz"] = _FakeImg(seg, zooms)
    slin, sdims, _ = bv.load_seg_uint(Path("/nowhere/CASE-seg.nii.gz"))
    out["raw_seg"] = seg
    out["linear_seg"] = slin

    # world scaling + framing through the reference's own load_dir / frame_volume
    class _Buf:
        def copy_from_numpy(self, a):
            self.data = np.array(a)

    class _Txt:
        text = ""

    fake = types.SimpleNamespace()
    fake._create_float_buffer = lambda lin: _Buf()
    fake._create_uint_buffer = lambda lin: _Buf()
    fake.info = _Txt()
    fake.camera = bv.OrbitalCamera(initial_radius=3.0, world_up=np.array([0.0, 1.0, 0.0], dtype=np.float32))
    fake.pred_check = types.SimpleNamespace(value=False)
    fake.vol_dims = None
    fake.frame_volume = types.MethodType(bv.BraTSViewer.frame_volume, fake)
    with tempfile.TemporaryDirectory() as td:
        for suf in ("t1n", "t1c", "t2w", "t2f", "seg"):
            (Path(td) / f"CASE-{suf}.nii.gz").write_bytes(b"")
        bv.BraTSViewer.load_dir(fake, Path(td))
    out["voxel_size"] = np.asarray(fake.voxel_size)
    out["vol_min"] = np.asarray(fake.vol_min)
    out["vol_dims"] = np.asarray(fake.vol_dims)
    out["cam_target"] = np.asarray(fake.camera.target)
    out["cam_radius"] = np.asarray(fake.camera.radius)
    eye, right, up, fwd = fake.camera.get_basis()
    out["cam_eye"] = eye; out["cam_right"] = right; out["cam_up"] = up; out["cam_forward"] = fwd
    np.savez_compressed(OUT / "ingest.npz", **out)
    print("wrote ingest.npz", sorted(out))


# ------------------------------------------------------------------ volumeRendering app
def make_bc4_and_mask():
    _install_stubs()
    sys.modules.pop("camera", None)
    app = _load("ref_volume_app", REF / "scripts/volumeRendering/app.py")
    rng = np.random.default_rng(99)
    W, H, D = 13, 10, 3
    bw, bh = (W + 3) // 4, (H + 3) // 4
    blocks = rng.integers(0, 256, size=(D, bw * bh, 8), dtype=np.uint8)
    blocks[0, 0, 0], blocks[0, 0, 1] = 200, 17      # r0 > r1 branch
    blocks[0, 1, 0], blocks[0, 1, 1] = 17, 200      # r0 < r1 branch
    blocks[0, 2, 0], blocks[0, 2, 1] = 77, 77       # r0 == r1
    captured = {}
    fake = types.SimpleNamespace(volume_width=W, volume_height=H, volume_depth=D)
    fake._upload_u8_volume_from_array = lambda a: captured.__setitem__("vox", np.array(a))
    with tempfile.TemporaryDirectory() as td:
        p = Path(td) / "v.bin-gz"
        with gzip.open(p, "wb") as f:
            f.write(blocks.tobytes())
        app.App._load_volume_bc4(fake, p)
    np.savez_compressed(OUT / "bc4.npz", blocks=blocks, W=W, H=H, D=D, decoded=captured["vox"].reshape(D, H, W))
    print("wrote bc4.npz")

    shape = (9, 8, 6)
    lab = rng.choice(np.array([0.0, 1.0, 2.0, 4.0], dtype=np.float32), size=shape)
    _REGISTRY["mask.nii.gz"] = _FakeImg(lab, (1, 1, 1))
    res = {}
    for mode in ("occupancy", "labels"):
        fake = types.SimpleNamespace()
        fake._upload_u8_volume_from_array = lambda a, m=mode: res.__setitem__(m, np.array(a))
        with tempfile.TemporaryDirectory() as td:
            p = Path(td) / "mask.nii.gz"
            p.write_bytes(b"")
            app.App._load_nifti_mask(fake, p, mode=mode)
        res[mode + "_whd"] = np.array([fake.volume_width, fake.volume_height, fake.volume_depth])
    np.savez_compressed(OUT / "nifti_mask.npz", raw=lab, **res)
    print("wrote nifti_mask.npz")


if __name__ == "__main__":
    make_cameras()
    make_ingest()
    make_bc4_and_mask()
