"""Known-answer tests that pin the adaptive-sampling oracle (oracle/oracle_adaptive.py; spec:
docs/DifferentiableRendering.md:131-148, maths only).  CPU only."""
import math
from dataclasses import replace

import numpy as np
import torch

from mri_raytracer_b200.synth import ramp_tf
from oracle import oracle_adaptive as A
from oracle import oracle_torch as O
from scenes import small_scene


def test_homogeneous_volume_closed_form_for_any_partition():
    """sigma constant along the ray: prod(1 - alpha_j) = exp(-sigma * sum(Delta_j)) = exp(-sigma L) whatever
    the quantile intervals are, so C = bg + c (1 - exp(-sigma L)) exactly (the intervals tile [t0,t1))."""
    _, _, P = small_scene(C=1, dims=(12, 10, 8), W=9, H=7, seed=0)
    P = replace(P, tfMode=1, ertThreshold=1e-12, bgColor=(0.1, 0.2, 0.3))
    vol = torch.full((1, 8, 10, 12), 0.6)
    tf = ramp_tf(16, sigma_scale=3.0, cutoff=0.0).double()
    img, aux = A.render_adaptive(vol, P, tf=tf, n_coarse=5, n_fine=7, dtype=torch.float64, return_aux=True)
    px, py = O.pixel_grid(P)
    o, d = O.make_rays(P, px, py, torch.float64)
    t0, t1, hit = O.clip_rays(P, o, d, torch.float64)
    rgba = O.tf_lookup(tf, torch.full((1,), 0.6, dtype=torch.float64))[0]
    L = torch.where(hit, t1 - t0, torch.zeros_like(t0))
    want = torch.tensor([0.1, 0.2, 0.3], dtype=torch.float64)[None, :] + rgba[None, :3] * (1 - torch.exp(-rgba[3] * L))[:, None]
    assert torch.allclose(img.reshape(-1, 4)[:, :3], want, atol=1e-12)
    assert bool((aux["n_taken"].reshape(-1)[hit] == 7).all()) and bool((aux["n_taken"].reshape(-1)[~hit] == 0).all())


def test_uniform_importance_is_the_midpoint_rule_and_importance_moves_samples():
    vol, _, P = small_scene(C=1, dims=(20, 18, 16), W=10, H=8, seed=3)
    P = replace(P, tfMode=1, ertThreshold=1e-12)
    tf = ramp_tf(32, sigma_scale=8.0, cutoff=0.2).double()
    J = 24
    # a huge floor makes the importance constant: quantiles are uniform, sample j sits at the centre of slot j
    a = A.render_adaptive(vol, P, tf=tf, n_coarse=6, n_fine=J, eps_w=1e9, dtype=torch.float64)
    px, py = O.pixel_grid(P)
    o, d = O.make_rays(P, px, py, torch.float64)
    t0, t1, hit = O.clip_rays(P, o, d, torch.float64)
    bmin, _, vs = O.box_bounds(P, torch.float64)
    Cc = torch.zeros(px.numel(), 3, dtype=torch.float64); T = torch.ones(px.numel(), dtype=torch.float64)
    for j in range(J):
        t = t0 + (j + 0.5) * (t1 - t0) / J
        pIdx = ((o + t[:, None] * d) - bmin[None, :]) / vs[None, :]
        val = torch.clamp((O.sample_linear(vol[0].double(), pIdx, (20, 18, 16)) - 0.0) / 1.0, 0.0, 1.0)
        rgba = O.tf_lookup(tf, val)
        al = torch.where(hit, 1 - torch.exp(-rgba[:, 3] * (t1 - t0) / J), torch.zeros_like(t))
        Cc = Cc + (al * T)[:, None] * rgba[:, :3]; T = T * (1 - al)
    assert torch.allclose(a.reshape(-1, 4)[:, :3], Cc, atol=1e-9)
    # where the extinction is concentrated (a sharp TF) and fine samples are few, placing them by
    # the importance lands much closer to the converged image than the same number of uniform ones
    tfs = ramp_tf(32, sigma_scale=40.0, cutoff=0.45).double()
    ref = A.render_adaptive(vol, P, tf=tfs, n_coarse=6, n_fine=4096, eps_w=1e9, dtype=torch.float64)
    uniform = A.render_adaptive(vol, P, tf=tfs, n_coarse=6, n_fine=8, eps_w=1e9, dtype=torch.float64)
    adaptive = A.render_adaptive(vol, P, tf=tfs, n_coarse=48, n_fine=8, eps_w=1e-3, dtype=torch.float64)
    assert float((adaptive - ref).abs().mean()) < 0.5 * float((uniform - ref).abs().mean())


def test_adaptive_oracle_gradcheck_fp64():
    """autograd of the explicit inverse CDF == the doc's implicit differentiation (:142-146); checked
    against finite differences in float64 w.r.t. the volume and the LUT."""
    vol, _, P = small_scene(C=2, dims=(6, 5, 5), W=4, H=3, seed=1)
    P = replace(P, tfMode=1, ertThreshold=1e-12, volWeight=(1.0, 0.5, 1.0, 1.0))
    tf = ramp_tf(8, sigma_scale=6.0, cutoff=0.1).double()
    tf[:, 1] = tf[:, 1] ** 2
    g = torch.Generator().manual_seed(0)
    wgt = torch.rand(3, 4, 4, generator=g, dtype=torch.float64)
    v = vol.double().requires_grad_(True)
    t = tf.clone().requires_grad_(True)

    def f(vv, tt):
        return (A.render_adaptive(vv, P, tf=tt, n_coarse=4, n_fine=6, eps_w=1e-2, dtype=torch.float64) * wgt).sum()
    assert torch.autograd.gradcheck(f, (v, t), eps=1e-7, atol=1e-6, rtol=1e-4, nondet_tol=0.0)
