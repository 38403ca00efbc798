"""Host-side check of the geometry behind mrt_view_spans (csrc/forward.cu: mrt_view_spans_kernel), restated in
numpy: a view's spans are the union, over the ACTIVE bricks, of the bounding rectangles of their projected boxes —
and the kernel lets only SURFACE bricks (one of the six face neighbours inactive or outside the grid) project.

1. the surface bricks alone give exactly the spans of all active bricks (the covering argument in the kernel's comment);
2. the spans are conservative: any point of an active brick projects into the span of its pixel row's band.

No GPU, no library call: the CUDA kernel itself is compared against dense renders bit for bit in the `-m gpu` tests
(sparse + fill == dense); this file pins the reasoning the kernel's pruning rests on."""
import math

import numpy as np
import pytest

from mri_raytracer_b200 import Camera, OrbitalCamera

BRICK, TILE = 8, 8
BOX_MARGIN, SPAN_MARGIN = 0.25, 0.75           # march.cuh


def project(cam, W, H, vs, vmin, pts, ortho_half=None):
    """index-space points [N,3] -> pixel coordinates, as mrt_project_box (pinhole / orthographic)."""
    w = vmin + pts * vs - np.asarray(cam.eye, dtype=np.float64)
    U, V, Wv = (np.asarray(a, dtype=np.float64) for a in (cam.U, cam.V, cam.W))
    xc, yc, zc = w @ U, w @ V, w @ Wv
    aspect = W / H
    if cam.ortho:
        uvx, uvy = xc / (aspect * cam.ortho_half_height), -yc / cam.ortho_half_height
    else:
        assert (zc > 1e-4).all()
        focal = 1.0 / math.tan(0.5 * cam.fovY)
        uvx, uvy = (xc / zc) * focal / aspect, -(yc / zc) * focal
    return (uvx + 1.0) * 0.5 * W - 0.5, (uvy + 1.0) * 0.5 * H - 0.5


def brick_rect(cam, W, H, vs, vmin, b):
    lo = np.asarray(b, dtype=np.float64) * BRICK - SPAN_MARGIN
    hi = (np.asarray(b, dtype=np.float64) + 1) * BRICK + SPAN_MARGIN
    corners = np.array([[(hi if (c >> a) & 1 else lo)[a] for a in range(3)] for c in range(8)])
    x, y = project(cam, W, H, vs, vmin, corners)
    rx0, rx1 = max(0, math.floor(x.min()) - 1), min(W - 1, math.ceil(x.max()) + 1)
    b0 = max(0, math.floor((y.min() - TILE) / TILE))
    b1 = min(H // TILE - 1, math.floor((y.max() + 1.0) / TILE))
    return rx0, rx1, b0, b1


def spans_of(cam, W, H, vs, vmin, bricks):
    ty = H // TILE
    sp = np.stack([np.full(ty, 2 ** 31 - 1), np.full(ty, -1)], axis=1)
    for b in bricks:
        rx0, rx1, b0, b1 = brick_rect(cam, W, H, vs, vmin, b)
        if rx0 > rx1:
            continue
        sp[b0:b1 + 1, 0] = np.minimum(sp[b0:b1 + 1, 0], rx0)
        sp[b0:b1 + 1, 1] = np.maximum(sp[b0:b1 + 1, 1], rx1)
    return sp


def blob(shape, rng):
    """A head-like occupancy: a bumpy ellipsoid with a few cavities, in bricks [nbz, nby, nbx]."""
    nbz, nby, nbx = shape
    z, y, x = np.meshgrid(np.arange(nbz), np.arange(nby), np.arange(nbx), indexing="ij")
    c = (np.array(shape) - 1) / 2.0
    r = np.sqrt(((z - c[0]) / (0.45 * nbz)) ** 2 + ((y - c[1]) / (0.42 * nby)) ** 2 + ((x - c[2]) / (0.40 * nbx)) ** 2)
    act = r + 0.15 * rng.standard_normal(shape) < 1.0
    act[tuple(rng.integers(1, s - 1, size=5) for s in shape)] = False     # cavities inside
    return act


@pytest.mark.parametrize("ortho", [False, True])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_surface_bricks_give_the_spans_of_all_active_bricks(seed, ortho):
    rng = np.random.default_rng(seed)
    shape = (9, 11, 12)                                                    # bricks (z, y, x)
    act = blob(shape, rng)
    nbz, nby, nbx = shape
    dims = np.array([nbx, nby, nbz]) * BRICK
    vs = np.full(3, 1.8 / dims.max()); vmin = -0.5 * vs * dims
    W, H = 320, 256
    idx = [(int(x), int(y), int(z)) for z, y, x in np.argwhere(act)]
    pad = np.pad(act, 1, constant_values=False)                           # outside the grid counts as inactive
    surf = []
    for (x, y, z) in idx:
        n6 = (pad[z + 1, y + 1, x] and pad[z + 1, y + 1, x + 2] and pad[z + 1, y, x + 1] and pad[z + 1, y + 2, x + 1]
              and pad[z, y + 1, x + 1] and pad[z + 2, y + 1, x + 1])
        if not n6:
            surf.append((x, y, z))
    assert 0 < len(surf) < len(idx), "the blob must have interior bricks for the test to mean anything"
    orb = OrbitalCamera(initial_radius=float(np.linalg.norm(vs * dims) * 0.8), initial_phi=1.1, initial_theta=0.3 + seed)
    orb.set_fov_degrees(60.0)
    for k in range(6):
        orb.theta = 0.3 + seed + 2.0 * math.pi * k / 6
        orb.phi = 0.6 + 0.3 * k
        cam = Camera.from_orbital(orb, ortho=ortho)
        full = spans_of(cam, W, H, vs, vmin, idx)
        pruned = spans_of(cam, W, H, vs, vmin, surf)
        assert np.array_equal(full, pruned), f"view {k}: pruning the interior bricks changed the spans"
        # conservative: random points of active bricks (the positions sample slots can take) land inside their band's span
        b = np.asarray(idx)[rng.integers(0, len(idx), size=400)]
        pts = (b + rng.random((400, 3))) * BRICK
        px, py = project(cam, W, H, vs, vmin, pts)
        for x_, y_ in zip(px, py):
            row = int(round(y_))
            if 0 <= row < H and -0.5 <= x_ <= W - 0.5:
                s0, s1 = full[row // TILE]
                assert s0 <= x_ <= s1
