"""Ingest path vs vectors produced by EXECUTING the reference loaders
(tests/golden/make_golden.py: inr/viewer/brats_viewer.py load_nifti_float / load_seg_uint /
load_dir / frame_volume; scripts/volumeRendering/app.py _load_volume_bc4 / _load_nifti_mask)."""
from pathlib import Path

import numpy as np
import pytest

from mri_raytracer_b200 import OrbitalCamera, volume as V

G = Path(__file__).parent / "golden"


def test_percentile_normalise_and_flatten_bitwise():
    g = np.load(G / "ingest.npz")
    for suf in ("t1n", "t1c", "t2w", "t2f"):
        linear, norm, dims = V.normalize_percentile(g[f"raw_{suf}"])
        assert linear.dtype == np.float32 and np.array_equal(linear.view(np.uint32), g[f"linear_{suf}"].view(np.uint32))
        assert np.array_equal(norm, g[f"norm_{suf}"]) and np.array_equal(dims, g["dims"])
    lin_c, _, _ = V.normalize_percentile(g["raw_const"])        # vmax <= vmin branch
    assert np.array_equal(lin_c, g["linear_const"])
    assert np.array_equal(V.labels_from_float(g["raw_seg"]), g["linear_seg"])


def test_world_scaling_and_framing_bitwise():
    g = np.load(G / "ingest.npz")
    vs, vmin = V.world_scaling(g["dims"], g["zooms"])
    assert np.array_equal(vs.view(np.uint32), g["voxel_size"].view(np.uint32))
    assert np.array_equal(vmin.view(np.uint32), g["vol_min"].view(np.uint32))
    # frame_volume (brats_viewer.py:320-324) through our camera
    cam = OrbitalCamera(initial_radius=3.0, world_up=np.array([0.0, 1.0, 0.0], dtype=np.float32))
    ext = vs * g["dims"].astype(np.float32)
    cam.target = vmin + 0.5 * ext
    cam.radius = np.linalg.norm(ext) * 0.8
    assert np.array_equal(np.asarray(cam.target), g["cam_target"]) and float(cam.radius) == float(g["cam_radius"])
    eye, right, up, fwd = cam.get_basis()
    for got, key in ((eye, "cam_eye"), (right, "cam_right"), (up, "cam_up"), (fwd, "cam_forward")):
        assert np.array_equal(np.asarray(got, dtype=np.float32).view(np.uint32), g[key].view(np.uint32)), key


def test_bc4_host_decoder_and_mask_bitwise():
    g = np.load(G / "bc4.npz")
    out = V.decode_bc4_host(g["blocks"], int(g["W"]), int(g["H"]), int(g["D"]))
    assert np.array_equal(out, g["decoded"])
    m = np.load(G / "nifti_mask.npz")
    for mode in ("occupancy", "labels"):
        u8 = V.nifti_mask_to_u8(m["raw"], mode)
        assert np.array_equal(u8.reshape(-1), m[mode])
        assert tuple(m[mode + "_whd"]) == m["raw"].shape
    with pytest.raises(ValueError):
        V.nifti_mask_to_u8(m["raw"], "nope")


@pytest.mark.gpu
def test_device_ingest_kernels_bitwise(cuda):
    import torch
    g = np.load(G / "bc4.npz")
    W, H, D = int(g["W"]), int(g["H"]), int(g["D"])
    out = V.decode_bc4(torch.from_numpy(g["blocks"].copy()).cuda().reshape(-1), W, H, D)
    assert np.array_equal(out.cpu().numpy(), g["decoded"])
    f = V.u8_to_f32(out)
    assert np.array_equal(f.cpu().numpy(), (g["decoded"].astype(np.float32) / np.float32(255.0)))
    gi = np.load(G / "ingest.npz")
    for suf in ("t1n", "t2w"):
        raw = gi[f"raw_{suf}"]
        vmin, rng = V.percentile_window(raw)
        norm = V.normalize_on_device(torch.from_numpy(raw.copy()).cuda(), vmin, rng)
        assert np.array_equal(norm.cpu().numpy().view(np.uint32), gi[f"norm_{suf}"].view(np.uint32))
    with pytest.raises(ValueError):
        V.decode_bc4(torch.zeros(7, dtype=torch.uint8, device="cuda"), W, H, D)
