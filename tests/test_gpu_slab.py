"""Slab renderer (volume_cs, scripts/volumeRendering/volume_render.slang:104-148) vs the oracle."""
import math
from dataclasses import replace

import numpy as np
import pytest
import torch

from mri_raytracer_b200 import Camera, OrbitalCameraYUp, SlabParams, api
from parity import O

pytestmark = pytest.mark.gpu


def _vol(dims, seed):
    g = torch.Generator().manual_seed(seed)
    X, Y, Z = dims
    v = torch.rand(Z, Y, X, generator=g)
    z, y, x = torch.meshgrid(torch.linspace(-1, 1, Z), torch.linspace(-1, 1, Y), torch.linspace(-1, 1, X), indexing="ij")
    v = v * ((x * x + y * y + z * z) < 0.7)
    return (v * 255).to(torch.uint8).contiguous()


@pytest.mark.parametrize("near,far,steps", [(4.3, 4.4, 64.0), (2.5, 6.0, 200.0), (3.0, 5.5, 16.0)])
def test_slab_matches_oracle(cuda, near, far, steps):
    dims = (30, 36, 30)
    vol = _vol(dims, 3)
    # the app's camera: Y-up orbit, radius 4.2, phi 80 deg, theta 25 deg, fov 72 deg (app.py:34,336)
    cam = OrbitalCameraYUp(initial_radius=4.2, initial_phi=math.radians(80.0), initial_theta=math.radians(25.0))
    c = Camera.from_orbital(cam)
    P = SlabParams(imageSize=(70, 45), stepCount=steps, nearPlane=near, farPlane=far, volDim=dims).with_camera(c)
    img = api.render_slab(vol.cuda(), None, P).cpu()
    ref = O.render_slab(vol, P)
    assert (img - ref).abs().max() <= 1e-4
    assert float(img[..., 3].min()) == 1.0
    if far - near > 1.0:
        assert float(img[..., 0].max()) > 0.05


def test_slab_tile_ranges_and_errors(cuda):
    dims = (16, 16, 16)
    vol = _vol(dims, 1).cuda()
    P = SlabParams(imageSize=(24, 24), stepCount=40.0, nearPlane=3.0, farPlane=5.5, volDim=dims)
    full = api.render_slab(vol, None, P)
    out = torch.full((24, 24, 4), -3.0, device="cuda")
    api.render_slab(vol, None, P, tile_range=(0, 4), out=out)
    assert torch.equal(out[:8], full[:8]) and torch.equal(out[8:16, :8], full[8:16, :8])
    assert float(out[8:16, 8:].max()) == -3.0 and float(out[16:].max()) == -3.0     # untouched tiles
    api.render_slab(vol, None, P, tile_range=(4, 9), out=out)
    assert torch.equal(out, full)
    with pytest.raises(ValueError):
        api.render_slab(vol, None, replace(P, volDim=(8, 8, 8)))
