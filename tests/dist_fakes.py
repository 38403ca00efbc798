"""Test-only stand-ins that let the multi-rank HOST logic of mri_raytracer_b200.dist run on CPU with
the gloo backend: a duck-typed `Volume` whose renders come from the oracle, and the torch restatement
of the ordered `over` composite.  The product package has no such hooks and no CPU path."""
from __future__ import annotations

from dataclasses import replace

import torch


def tile_pixels(P, tile_range):
    from mri_raytracer_b200 import tiles
    W, H = P.imageSize
    xs, ys = [], []
    for t in range(*tile_range):
        for lane in range(64):
            x, y = tiles.pixel_of_tile_lane(t, lane, W)
            if x < W and y < H:
                xs.append(x); ys.append(y)
    return torch.tensor(xs, dtype=torch.long), torch.tensor(ys, dtype=torch.long)


class OracleVolume:
    """Quacks like api.Volume for dist.render_views / PeerFramebuffer's NCCL fallback / render_sort_last."""
    device = torch.device("cpu")

    def __init__(self, vol, tf=None, partial=None):
        self.vol, self.tf, self.partial = vol, tf, partial

    def forward_batch(self, P, cams, tf, out=None, tile_range=None):
        from oracle import oracle_torch as O
        W, H = P.imageSize
        tr = tile_range if tile_range is not None else (0, ((W + 7) // 8) * ((H + 7) // 8))
        px, py = tile_pixels(P, tr)
        for v, cam in enumerate(cams):
            if px.numel():
                out[v][py, px] = O.render(self.vol, replace(P.with_camera(cam), tfMode=1 if tf is not None else 0), tf=tf,
                                          pixels=(px, py))
        return out

    def forward(self, P, tf):
        return self.partial            # sort-last tests feed a fixed per-rank partial image


def composite_over_torch(partials, order, bg, alpha_mode=0):
    """Ordered front-to-back `over` in torch: partials [K,npix,4] = (premultiplied rgb, T)."""
    C = torch.zeros_like(partials[0, :, :3])
    T = torch.ones_like(partials[0, :, 3])
    for k in order:
        C = C + T[:, None] * partials[k, :, :3]
        T = T * partials[k, :, 3]
    bgv = torch.as_tensor(bg, dtype=C.dtype, device=C.device)
    a = (1.0 - T) if alpha_mode else torch.ones_like(T)
    return torch.cat([bgv[None, :] + C, a[:, None]], dim=1)


def oracle_render_part(volume, camera, tf, P, tile_range=None, **kw):
    """Replacement for api.render inside dist.render_differentiable: the differentiable oracle on the
    pixels of a tile range, zeros elsewhere."""
    from oracle import oracle_torch as O
    W, H = P.imageSize
    px, py = tile_pixels(P, tile_range)
    out = torch.zeros((H, W, 4), dtype=torch.float32)
    if px.numel():
        out = out.index_put((py, px), O.render(volume, replace(P, tfMode=1), tf=tf, pixels=(px, py)))
    return out
