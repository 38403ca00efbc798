"""Soft (learnable) occupancy in the oracle — docs/DifferentiableRendering.md section 11 (:202-206) exists in
the reference as one sentence of maths, so the restatement is pinned by known answers and gradcheck.  CPU only."""
import math

import numpy as np
import torch

from parity import O
from test_oracle_known_answers import _axis_params


def test_unit_occupancy_is_the_plain_render_and_zero_occupancy_is_the_background():
    P = _axis_params(dims=(20, 20, 20), voxelSize=(0.05, 0.05, 0.05), bgColor=(0.1, 0.2, 0.3), ertThreshold=1e-9)
    vol = 0.2 + 0.6 * torch.rand(1, 20, 20, 20, generator=torch.Generator().manual_seed(1))
    plain = O.render(vol, P)
    assert torch.equal(O.render(vol, P, soft_occ=torch.ones(3, 3, 3)), plain)
    zero = O.render(vol, P, soft_occ=torch.zeros(3, 3, 3))
    assert torch.equal(zero[..., :3], torch.tensor([0.1, 0.2, 0.3]).expand(5, 5, 3))


def test_homogeneous_volume_with_uniform_occupancy_closed_form():
    """sigma' = o * sigma  =>  C = bg + c (1 - exp(-o sigma dt))^... : the closed form of the plain case with o*sigma."""
    P = _axis_params(bgColor=(0.0, 0.0, 0.0), ertThreshold=1e-9)
    v0, o = 0.6, 0.35
    vol = torch.full((1, 10, 10, 10), v0)
    img, aux = O.render(vol, P, return_aux=True, dtype=torch.float64, soft_occ=torch.full((2, 2, 2), o, dtype=torch.float64))
    N = int(aux["n_samples"][2, 2])
    alpha = 1 - math.exp(-o * v0 * 0.4 * float(np.float32(0.05)))
    assert abs(float(img[2, 2, 0]) - v0 * (1 - (1 - alpha) ** N)) < 1e-6


def test_occupancy_is_looked_up_per_brick_of_the_base_cell():
    """Two bricks along the ray (z < 8 and z >= 8): switching one off removes exactly its slots."""
    P = _axis_params(dims=(10, 10, 16), voxelSize=(0.1, 0.1, 0.1), volMin=(-0.5, -0.5, -0.8), bgColor=(0, 0, 0), ertThreshold=1e-9)
    vol = torch.full((1, 16, 10, 10), 0.5)
    occ = torch.ones(2, 2, 2, dtype=torch.float64)
    occ[1] = 0.0                                                     # bricks with iz >= 8 are empty
    img, aux = O.render(vol, P, return_aux=True, dtype=torch.float64, soft_occ=occ)
    # slots k with floor(clamp(z_k)) < 8: z index = k * 0.5 (step 0.05 world = half a voxel) -> k = 0..15
    alpha = 1 - math.exp(-0.5 * 0.4 * float(np.float32(0.05)))
    assert abs(float(img[2, 2, 0]) - 0.5 * (1 - (1 - alpha) ** 16)) < 1e-6


def test_fp64_gradcheck_with_soft_occupancy():
    torch.manual_seed(0)
    dims = (10, 9, 12)
    P = O.params(imageSize=(4, 4), dims=dims, voxelSize=(0.2, 0.2, 0.2), volMin=(-1.0, -0.9, -1.2), stepSize=0.17,
                 eye=(0.4, 0.3, -3.0), tfMode=1, ertThreshold=1e-6)
    vol = (0.2 + 0.6 * torch.rand(1, 12, 9, 10, dtype=torch.float64)).requires_grad_(True)
    tf = torch.rand(5, 4, dtype=torch.float64).requires_grad_(True)
    occ = (0.1 + 0.8 * torch.rand(2, 2, 2, dtype=torch.float64)).requires_grad_(True)
    f = lambda v, t, o: O.render(v, P, tf=t * torch.tensor([1.0, 1.0, 1.0, 3.0], dtype=torch.float64), dtype=torch.float64, soft_occ=o)
    assert torch.autograd.gradcheck(f, (vol, tf, occ), eps=1e-6, atol=1e-6, rtol=1e-4, nondet_tol=0.0)
