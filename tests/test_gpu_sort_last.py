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


# ----------------------------------------------------------------------------- differentiable shards (cfg5 trains)
def _sharded_gradients(vol, tf, P, grid, storage):
    """loss = mean((composite of the shards' partials - target)^2); -> (image, dL/dvolume assembled in
    global coordinates from the shards' own gradients, dL/dtf)."""
    dims = tuple(P.dims)
    R = grid[0] * grid[1] * grid[2]
    subs, parts = [], []
    t = tf.clone().requires_grad_(True)
    for r in range(R):
        lo, hi, _ = mdist.shard_box(dims, grid, r)
        sub = mdist.slice_shard(vol, lo, hi).clone().requires_grad_(True)
        subs.append((sub, lo, hi))
        parts.append(api.render_shard(sub, (lo, hi), dims, None, t, P, storage=storage).reshape(-1, 4))
    order = mdist.visibility_order(np.asarray(P.eye, dtype=np.float64), P, grid)
    W, H = P.imageSize
    img = mdist.composite_over_differentiable(torch.stack(parts), order, P.bgColor, P.alphaMode).reshape(H, W, 4)
    target = torch.linspace(0, 1, img.numel(), device=img.device).reshape(img.shape)
    ((img - target) ** 2).mean().backward()
    g = torch.zeros_like(vol)
    for sub, lo, hi in subs:
        g[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] += sub.grad.float()
    return img.detach(), g, t.grad


@pytest.mark.parametrize("grid", [(1, 1, 1), (2, 2, 2), (1, 3, 2)])
@pytest.mark.parametrize("storage,ortho", [(None, False), (torch.float16, False), (None, True)])
def test_sharded_backward_adds_up_to_the_unsharded_gradient(cuda, grid, storage, ortho):
    """Missing-7 of VERDICT r1: backward for sharded and fp16 volumes.  Every shard differentiates exactly
    the slots it owns, so the shards' gradients (halo voxels summed) equal the unsharded render's, and the
    fp16-storage gradient equals the fp32 gradient at the fp16-rounded values."""
    vol, _, P = small_scene(C=1, dims=(41, 35, 29), W=56, H=40, seed=31, ortho=ortho, theta_deg=33.0, phi_deg=64.0)
    P = replace(P, tfMode=1, ertThreshold=1e-6, bgColor=(0.05, 0.1, 0.15))
    tf = ramp_tf(64, sigma_scale=12.0, cutoff=0.1).cuda()
    v = vol.cuda()
    if storage is torch.float16:
        v = v.half().float()                                  # the values both paths sample
    img, g, gt = _sharded_gradients(v, tf, P, grid, storage)
    a = v.clone().requires_grad_(True); b = tf.clone().requires_grad_(True)
    ref = api.render(a, None, b, P)
    target = torch.linspace(0, 1, ref.numel(), device=ref.device).reshape(ref.shape)
    ((ref - target) ** 2).mean().backward()
    assert (img - ref.detach()).abs().max() <= 2e-5
    assert (g - a.grad).abs().max() <= 1e-3 * a.grad.abs().max()
    assert (gt - b.grad).abs().max() <= 1e-3 * b.grad.abs().max()


def test_sharded_backward_matches_oracle_autograd(cuda):
    vol, _, P = small_scene(C=1, dims=(24, 22, 20), W=40, H=32, seed=8, theta_deg=20.0, phi_deg=75.0)
    P = replace(P, tfMode=1, ertThreshold=1e-6)
    tf = ramp_tf(32, sigma_scale=15.0, cutoff=0.1)
    _, g, gt = _sharded_gradients(vol.cuda(), tf.cuda(), P, (2, 2, 1), None)
    a = vol.clone().requires_grad_(True); b = tf.clone().requires_grad_(True)
    ref = O.render(a, P, tf=b)
    target = torch.linspace(0, 1, ref.numel()).reshape(ref.shape)
    ((ref - target) ** 2).mean().backward()
    assert (g.cpu() - a.grad).abs().max() <= 1e-3 * a.grad.abs().max()
    assert (gt.cpu() - b.grad).abs().max() <= 1e-3 * b.grad.abs().max()


def test_fp16_input_gets_an_fp16_gradient_and_T_local_carries_gradient(cuda):
    vol, _, P = small_scene(C=1, dims=(20, 20, 20), W=32, H=32, seed=2)
    P = replace(P, tfMode=1, ertThreshold=1e-6)
    tf = ramp_tf(32, sigma_scale=8.0, cutoff=0.1).cuda()
    lo, hi, _ = mdist.shard_box((20, 20, 20), (2, 1, 1), 0)
    sub = mdist.slice_shard(vol.cuda(), lo, hi).half().requires_grad_(True)
    part = api.render_shard(sub, (lo, hi), (20, 20, 20), None, tf, P)
    part[..., 3].sum().backward()                             # loss on T_local alone: only the .w channel carries gradient
    assert sub.grad is not None and sub.grad.dtype == torch.float16
    assert float(sub.grad.float().abs().max()) > 0.0
    # denser matter lowers the transmittance: the gradient is non-positive wherever the TF slope of sigma is >= 0
    assert float(sub.grad.float().max()) <= 1e-6
