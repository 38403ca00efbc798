"""Soft (learnable) occupancy on the GPU (docs/DifferentiableRendering.md section 11): sigma' = o(brick) * sigma,
forward 1e-4 and gradients 1e-3 against the oracle's autograd, through the C ABI."""
from dataclasses import replace

import pytest
import torch

from mri_raytracer_b200 import api
from mri_raytracer_b200.synth import ramp_tf
from scenes import small_scene
from parity import O

pytestmark = pytest.mark.gpu


def _occ(dims, seed):
    X, Y, Z = dims
    g = torch.Generator().manual_seed(seed)
    return 0.05 + 0.9 * torch.rand((Z + 7) // 8, (Y + 7) // 8, (X + 7) // 8, generator=g)


@pytest.mark.parametrize("C,use_tf,ortho", [(1, True, False), (2, True, True), (4, False, False)])
def test_soft_occupancy_forward_and_gradients_match_the_oracle(cuda, C, use_tf, ortho):
    dims = (28, 22, 19)
    vol, _, P = small_scene(C=C, dims=dims, W=40, H=32, seed=40 + C, ortho=ortho, theta_deg=31.0, phi_deg=66.0)
    P = replace(P, tfMode=1 if use_tf else 0, ertThreshold=1e-6, bgColor=(0.05, 0.1, 0.15), intensityAlpha=6.0)
    tf = ramp_tf(32, sigma_scale=12.0, cutoff=0.1) if use_tf else None
    occ = _occ(dims, 5)
    a = vol.clone().requires_grad_(True); o = occ.clone().requires_grad_(True)
    b = tf.clone().requires_grad_(True) if use_tf else None
    ref = O.render(a, P, tf=b, soft_occ=o)
    wgt = torch.rand(ref.shape, generator=torch.Generator().manual_seed(9))
    (ref * wgt).sum().backward()

    ga = vol.cuda().requires_grad_(True); go = occ.cuda().requires_grad_(True)
    gb = tf.cuda().requires_grad_(True) if use_tf else None
    img = api.render_soft_occupancy(ga, None, gb, P, go)
    (img * wgt.cuda()).sum().backward()
    assert (img.detach().cpu() - ref.detach()).abs().max() <= 1e-4
    assert (go.grad.cpu() - o.grad).abs().max() <= 1e-3 * o.grad.abs().max()
    assert (ga.grad.cpu() - a.grad).abs().max() <= 1e-3 * a.grad.abs().max()
    if use_tf:
        assert (gb.grad.cpu() - b.grad).abs().max() <= 1e-3 * b.grad.abs().max()


def test_unit_occupancy_reproduces_render_and_skipping_stays_exact(cuda):
    dims = (33, 30, 27)
    vol, _, P = small_scene(C=1, dims=dims, W=48, H=40, seed=3)
    P = replace(P, tfMode=1)
    tf = ramp_tf(64, sigma_scale=10.0, cutoff=0.15).cuda()
    ones = torch.ones((4, 4, 5), device="cuda")
    plain = api.render(api.Volume(vol.cuda(), quad=False), None, tf, P)
    soft = api.render_soft_occupancy(vol.cuda(), None, tf, P, ones)
    assert torch.equal(soft, plain)
    occ = _occ(dims, 7).cuda()
    assert torch.equal(api.render_soft_occupancy(vol.cuda(), None, tf, P, occ),
                       api.render_soft_occupancy(vol.cuda(), None, tf, replace(P, skipEmpty=0), occ))


def test_soft_occupancy_argument_checks(cuda):
    vol, _, P = small_scene(C=1, dims=(16, 16, 16), W=16, H=16)
    with pytest.raises(ValueError):
        api.render_soft_occupancy(vol.cuda(), None, None, P, torch.ones((2, 2, 3), device="cuda"))
