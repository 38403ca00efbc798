"""CPU (torch) restatement of differentiable adaptive sampling.  TEST INFRASTRUCTURE ONLY.

Spec: /root/reference/docs/DifferentiableRendering.md section 7 (:131-148) — the reference has the
maths only, no code and no tests, so this restatement is **parity unpinned**; it is checked by
known-answer tests and an fp64 ``gradcheck`` (tests/test_oracle_adaptive.py), and its autograd is the
ground truth for the CUDA kernel's gradients (the doc's implicit-differentiation formula, :142-146,
is exactly the derivative of the explicit inverse used here).

Definition (per ray, clip interval [t0, t1) from the same ray set-up as the uniform march):
  coarse   : K uniform bins of width h = (t1-t0)/K, sampled at their centres tc_k = t0 + (k+1/2) h
             (:133); importance w_k = sigma(tc_k) + eps_w >= eps_w > 0 (the doc's m(r(t)) >= 0 with a
             floor so that F is strictly increasing);
  CDF      : W_0 = 0, W_k = sum_{l<k} w_l, F(t) = W(t)/W_K piecewise linear (:134-136);
  quantile : Q(u) = t0 + h (k + (u W_K - W_k)/w_k) for W_k <= u W_K < W_{k+1}  (:138-140);
  fine     : J samples at tm_j = Q((j+1/2)/J); sample j stands for the quantile interval
             [Q(j/J), Q((j+1)/J)) of length Delta_j (they tile [t0, t1) exactly);
  march    : alpha_j = 1 - exp(-sigma_j Delta_j), front-to-back compositing with early termination
             exactly as brats_rt.slang:117,135-139.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import oracle_torch as O


def render_adaptive(volume: torch.Tensor, P, tf: Optional[torch.Tensor] = None, n_coarse: int = 16, n_fine: int = 32,
                    eps_w: float = 1e-3, pixels=None, dtype=torch.float32, return_aux: bool = False):
    """-> rgba [H,W,4] (or [N,4] with ``pixels``).  Differentiable w.r.t. ``volume`` and ``tf``."""
    C = volume.shape[0]
    X, Y, Z = (int(v) for v in O._get(P, "dims"))
    Wd, Hd = O._get(P, "imageSize")
    px, py = O.pixel_grid(P) if pixels is None else pixels
    nray = px.shape[0]
    vol = volume.to(dtype)
    tfd = tf.to(dtype) if tf is not None else None
    K, J = int(n_coarse), int(n_fine)
    bg = O._vec(O._get(P, "bgColor", (0, 0, 0)), dtype)
    ww = O._f32in(O._get(P, "ww", 1.0), dtype)
    wl = O._f32in(O._get(P, "wl", 0.5), dtype)
    ia = O._f32in(O._get(P, "intensityAlpha", 0.4), dtype)
    en = [int(v) for v in O._get(P, "volEnabled", (1, 1, 1, 1))]
    wt = [O._f32in(v, dtype) for v in O._get(P, "volWeight", (1, 1, 1, 1))]
    thr = O._f32in(O._get(P, "ertThreshold", O.ERT_DEFAULT), dtype)
    alpha_mode = int(O._get(P, "alphaMode", 0))
    lo = wl - ww * torch.tensor(0.5, dtype=dtype)
    one = torch.tensor(1.0, dtype=dtype)
    epsw = torch.tensor(float(np.float32(eps_w)), dtype=dtype)

    o, d = O.make_rays(P, px, py, dtype)
    t0, t1, hit = O.clip_rays(P, o, d, dtype)
    bmin, _, vs = O.box_bounds(P, dtype)
    Ccol = bg[None, :].expand(nray, 3).clone()
    T = torch.ones(nray, dtype=dtype)
    taken = torch.zeros(nray, dtype=torch.int64)
    margin = torch.full((nray,), float("inf"), dtype=torch.float64)     # closest approach of T to the ERT threshold
    hidx = torch.nonzero(hit).reshape(-1)
    if hidx.numel() > 0:
        ho, hd, ht0, ht1 = o[hidx], d[hidx], t0[hidx], t1[hidx]

        def shade(t):
            """(rgb [n,3], sigma [n]) of the field at ray parameter t (brats_rt.slang:119-137 / the LUT)."""
            p = ho + t[:, None] * hd
            pIdx = (p - bmin[None, :]) / vs[None, :]
            v = torch.zeros_like(t)
            wsum = torch.zeros((), dtype=dtype)
            for c in range(min(C, 4)):
                if en[c] != 0:
                    v = v + O.sample_linear(vol[c], pIdx, (X, Y, Z)) * wt[c]
                    wsum = wsum + wt[c]
            if float(wsum) > 0.0:
                v = v / wsum
            val = torch.clamp((v - lo) / ww, 0.0, 1.0)
            if tfd is None:
                return val[:, None].expand(-1, 3), val * ia
            rgba = O.tf_lookup(tfd, val)
            return rgba[:, :3], rgba[:, 3]

        h = (ht1 - ht0) / torch.tensor(float(K), dtype=dtype)
        w = []
        for k in range(K):                                            # coarse stage (:133)
            tc = ht0 + torch.tensor(k + 0.5, dtype=dtype) * h
            w.append(shade(tc)[1] + epsw)
        w = torch.stack(w, dim=1)                                     # [n,K]
        Wc = [torch.zeros_like(ht0)]
        for k in range(K):                                            # sequential prefix sums, like the kernel
            Wc.append(Wc[-1] + w[:, k])
        Wc = torch.stack(Wc, dim=1)                                   # [n,K+1]
        Wt = Wc[:, K]

        def quantile(u: float):                                       # Q(u), :138-140
            target = torch.tensor(float(u), dtype=dtype) * Wt
            k = torch.searchsorted(Wc[:, 1:].detach().contiguous(), target.detach()[:, None], right=True).reshape(-1)
            k = k.clamp(max=K - 1)
            Wk = Wc.gather(1, k[:, None]).reshape(-1)
            wk = w.gather(1, k[:, None]).reshape(-1)
            return ht0 + h * (k.to(dtype) + (target - Wk) / wk)

        qb = [ht0] + [quantile(j / J) for j in range(1, J)] + [ht1]
        hC, hT = Ccol[hidx], T[hidx]
        htaken = taken[hidx]
        hmargin = torch.full((hidx.numel(),), float("inf"), dtype=torch.float64)
        for j in range(J):
            active = hT > thr                                         # :117
            hmargin = torch.minimum(hmargin, ((hT.detach().to(torch.float64) / float(thr)) - 1.0).abs())
            if not bool(active.any()):
                break
            tm = quantile((j + 0.5) / J)
            delta = qb[j + 1] - qb[j]
            rgb, sigma = shade(tm)
            alpha = one - torch.exp(-sigma * delta)
            if tfd is None:
                alpha = torch.where(sigma > 0, alpha, torch.zeros_like(alpha))     # :135 (val > 0)
            newC = hC + (alpha * hT)[:, None] * rgb
            newT = hT * (one - alpha)
            hC = torch.where(active[:, None], newC, hC)
            hT = torch.where(active, newT, hT)
            htaken = htaken + active.to(torch.int64)
        Ccol = Ccol.index_copy(0, hidx, hC)
        T = T.index_copy(0, hidx, hT)
        taken = taken.index_copy(0, hidx, htaken)
        margin = margin.index_copy(0, hidx, hmargin)
    a_out = torch.ones(nray, dtype=dtype) if alpha_mode == 0 else (one - T)
    rgba = torch.cat([Ccol, a_out[:, None]], dim=1)
    if pixels is None:
        rgba = rgba.reshape(Hd, Wd, 4)
    if not return_aux:
        return rgba
    shp = (Hd, Wd) if pixels is None else (nray,)
    return rgba, dict(T=T.reshape(shp), n_taken=taken.reshape(shp), hit=hit.reshape(shp), ert_margin=margin.reshape(shp))
