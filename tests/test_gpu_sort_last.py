"""Sort-last (brick-sharded) rendering: every shard marches only its own sub-box, partials are
composited front to back in visibility order.  With early termination effectively off the
result must equal the unsharded render (and the oracle) to 1e-4."""
from dataclasses import replace

import numpy as np
import pytest
import torch

from mri_raytracer_b200 import api, dist as mdist
from mri_raytracer_b200.synth import ramp_tf
from scenes import small_scene
from parity import O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("grid", [(2, 1, 1), (2, 2, 2), (1, 3, 2)])
@pytest.mark.parametrize("C,ortho,fold", [(1, False, True), (4, False, True), (2, True, False)])
def test_sort_last_equals_unsharded(cuda, grid, C, ortho, fold):
    vol, _, P = small_scene(C=C, dims=(41, 35, 29), W=72, H=56, seed=30 + C, ortho=ortho, theta_deg=33.0, phi_deg=64.0)
    P = replace(P, tfMode=1, ertThreshold=1e-6, bgColor=(0.05, 0.1, 0.15))
    tf = ramp_tf(64, sigma_scale=12.0, cutoff=0.1)
    full = api.render(api.Volume(vol.cuda(), fold=fold), None, tf.cuda(), P)
    sl = mdist.render_sort_last_emulated(vol.cuda(), None, tf.cuda(), P, grid, fold=fold)
    assert (sl - full).abs().max() <= 2e-5
    ref = O.render(vol, P, tf=tf)
    assert (sl.cpu() - ref).abs().max() <= 1e-4


def test_sort_last_eye_inside_grid_and_skipping_exact(cuda):
    vol, _, P = small_scene(C=1, dims=(40, 40, 40), W=64, H=64, seed=12)
    P = replace(P, tfMode=1, ertThreshold=1e-6, eye=(0.05, -0.02, 0.1))          # camera inside the volume
    tf = ramp_tf(32, sigma_scale=6.0, cutoff=0.2)
    full = api.render(api.Volume(vol.cuda()), None, tf.cuda(), P)
    a = mdist.render_sort_last_emulated(vol.cuda(), None, tf.cuda(), P, (2, 2, 2))
    b = mdist.render_sort_last_emulated(vol.cuda(), None, tf.cuda(), replace(P, skipEmpty=0), (2, 2, 2))
    assert torch.equal(a, b), "skipping changed a sharded render"
    assert (a - full).abs().max() <= 2e-5


def test_sort_last_fp16_shards(cuda):
    """config 5's shape in miniature: fp16 voxels, 2x2x2 shards, premultiplied partials."""
    vol, _, P = small_scene(C=1, dims=(45, 38, 33), W=64, H=48, seed=21, theta_deg=41.0, phi_deg=70.0)
    P = replace(P, tfMode=1, ertThreshold=1e-6)
    tf = ramp_tf(64, sigma_scale=10.0, cutoff=0.1)
    volh = vol.half()
    full = api.render(api.Volume(volh.cuda()), None, tf.cuda(), P)
    sl = mdist.render_sort_last_emulated(volh.cuda(), None, tf.cuda(), P, (2, 2, 2))
    assert (sl - full).abs().max() <= 2e-5
    ref = O.render(volh.float(), P, tf=tf)
    assert (sl.cpu() - ref).abs().max() <= 1e-4


@pytest.mark.parametrize("grid,H", [((2, 2, 2), 56), ((1, 3, 1), 50), ((2, 1, 1), 8)])
def test_fused_strip_exchange_equals_staged(cuda, grid, H):
    """mrt_render_forward_strips + mrt_composite_over_multi (the peer-memory exchange, emulated in
    one process) give bit-for-bit the image of the staged path (partials, reorder, composite)."""
    vol, _, P = small_scene(C=1, dims=(41, 35, 29), W=72, H=H, seed=33, theta_deg=12.0, phi_deg=100.0)
    P = replace(P, tfMode=1, ertThreshold=1e-6, bgColor=(0.2, 0.1, 0.3), alphaMode=1)
    tf = ramp_tf(64, sigma_scale=12.0, cutoff=0.1).cuda()
    dims = (41, 35, 29)
    R = grid[0] * grid[1] * grid[2]
    vols = []
    for r in range(R):
        lo, hi, _ = mdist.shard_box(dims, grid, r)
        vols.append(api.Volume(mdist.slice_shard(vol.cuda(), lo, hi), shard=(lo, hi), global_dims=dims))
    ex = mdist.PeerSortLast(H, 72, "cuda", emulate=R)
    fused = ex.render(vols, None, tf, P, grid).clone()
    staged = mdist.render_sort_last_emulated(vol.cuda(), None, tf, P, grid)
    assert torch.equal(fused, staged)
    for j in range(1, R):            # every "rank" holds the same finished image
        assert torch.equal(ex.final_all[j, :H], fused)


def test_shard_argument_checks(cuda):
    vol, _, P = small_scene(C=1, dims=(16, 16, 16), W=16, H=16)
    with pytest.raises(ValueError):
        api.Volume(vol.cuda(), shard=((0, 0, 0), (8, 8, 8)), global_dims=(16, 16, 16))     # dims mismatch
    lo, hi, _ = mdist.shard_box((16, 16, 16), (2, 1, 1), 1)
    V = api.Volume(mdist.slice_shard(vol.cuda(), lo, hi), shard=(lo, hi), global_dims=(16, 16, 16))
    with pytest.raises(ValueError):
        api.render(V, None, None, replace(P, dims=(9, 16, 16)))
    with pytest.raises(api._lib.MrtError):
        V.forward(replace(P, tMode="accumulate"), None)
