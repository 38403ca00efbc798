"""Known-answer tests that pin the oracle itself (SURVEY.md §8(c) tests 1-5, 9, 10).
The reference ships no tests or golden images for the shader path, so closed-form answers
stand in for them.  CPU only."""
import math
from dataclasses import replace

import numpy as np
import pytest
import torch

from parity import O
from scenes import small_scene, framed_params


def _axis_params(dims=(10, 10, 10), W=5, H=5, **kw):
    """Ortho camera looking down +z through the box centre, world voxelSize 0.1."""
    d = dict(imageSize=(W, H), dims=dims, voxelSize=(0.1, 0.1, 0.1), volMin=(-0.5, -0.5, -0.5),
             eye=(0.0, 0.0, -2.0), U=(1, 0, 0), V=(0, 1, 0), W=(0, 0, 1), ortho=1, orthoHalfHeight=0.3,
             stepSize=0.05, intensityAlpha=0.4)
    d.update(kw)
    return O.params(**d)


def test_homogeneous_volume_closed_form():
    """C = bg + c*(1-(1-alpha)^N), alpha = 1-exp(-sigma*dt), N = #samples (test 1)."""
    P = _axis_params(bgColor=(0.1, 0.2, 0.3), ertThreshold=1e-9)
    v0 = 0.6
    vol = torch.full((1, 10, 10, 10), v0)
    img, aux = O.render(vol, P, return_aux=True, dtype=torch.float64)
    N = int(aux["n_samples"][2, 2])
    assert N == 20                                   # box depth 1.0 / dt 0.05, first sample on the face
    alpha = 1 - math.exp(-v0 * 0.4 * float(np.float32(0.05)))
    want = v0 * (1 - (1 - alpha) ** N)
    got = img[2, 2]
    assert abs(float(got[0]) - (0.1 + want)) < 1e-6 and abs(float(got[2]) - (0.3 + want)) < 1e-6
    assert float(got[3]) == 1.0                      # alpha channel is the constant 1 (brats_rt.slang:167)
    assert abs(float(aux["T"][2, 2]) - (1 - alpha) ** N) < 1e-7


def test_single_voxel_impulse_weights():
    """Trilinear weights, x-fastest layout, lerp order and the dims-1.001 clamp (test 2)."""
    dims = (6, 5, 4)
    vol = torch.zeros(4, 5, 6)
    vol[2, 3, 1] = 1.0                                # z=2, y=3, x=1
    p = torch.tensor([[1.25, 2.5, 1.75], [0.3, 3.0, 2.0], [1.0, 3.0, 2.0], [5.0, 4.0, 3.0], [-1.0, 9.0, 2.0]])
    s = O.sample_linear(vol, p, dims)
    assert abs(float(s[0]) - 0.75 * 0.5 * 0.75) < 1e-6
    assert abs(float(s[1]) - 0.3) < 1e-6
    assert float(s[2]) == 1.0
    assert float(s[3]) == 0.0
    vol2 = torch.zeros(4, 5, 6); vol2[3, 4, 5] = 1.0  # far corner: max weight 0.999^3 by the clamp
    s2 = O.sample_linear(vol2, torch.tensor([[5.0, 4.0, 3.0], [50.0, 40.0, 30.0]]), dims)
    assert abs(float(s2[0]) - 0.999 ** 3) < 1e-5 and float(s2[0]) == float(s2[1])


def test_linear_ramp_is_reproduced_exactly():
    """Trilinear interpolation is exact for linear fields (test 3)."""
    X, Y, Z = 9, 7, 5
    z, y, x = torch.meshgrid(torch.arange(Z), torch.arange(Y), torch.arange(X), indexing="ij")
    vol = (0.25 * x + 0.5 * y + 1.0 * z).double()
    p = torch.rand(200, 3, dtype=torch.float64) * torch.tensor([X - 1.01, Y - 1.01, Z - 1.01])
    s = O.sample_linear(vol, p, (X, Y, Z))
    assert torch.allclose(s, 0.25 * p[:, 0] + 0.5 * p[:, 1] + p[:, 2], atol=1e-12)


def test_miss_grazing_inside_and_zero_direction():
    """brats_rt.slang:95-99 (de-zero keeps +1e-6), :107-109 (near/far), misses write bg (test 4)."""
    P = _axis_params(bgColor=(0.5, 0.25, 0.125), orthoHalfHeight=2.0, W=9, H=9)
    vol = torch.ones(1, 10, 10, 10)
    img, aux = O.render(vol, P, return_aux=True)
    assert int(aux["n_samples"][0, 0]) == 0 and torch.equal(img[0, 0], torch.tensor([0.5, 0.25, 0.125, 1.0]))
    assert int(aux["n_samples"][4, 4]) == 20           # d = (0,0,1): two zero components, still hits
    Pin = _axis_params(eye=(0.0, 0.0, 0.0), ortho=0)
    _, a2 = O.render(vol, Pin, return_aux=True)
    assert int(a2["n_samples"][2, 2]) == 10            # eye inside: t0 clamps to 0, half the depth
    Pnf = _axis_params(nearT=1.7, farT=2.0)
    _, a3 = O.render(vol, Pnf, return_aux=True)
    assert int(a3["n_samples"][2, 2]) == 6             # t in [1.7, 2.0): 1.7,1.75,...,1.95
    Pnone = _axis_params(nearT=5.0)
    img4, a4 = O.render(vol, Pnone, return_aux=True)
    assert int(a4["n_samples"].sum()) == 0 and float(img4[..., :3].abs().sum()) == 0.0


def test_ert_straddle():
    """sigma chosen so T crosses 0.01 between samples k and k+1 (test 5)."""
    P = _axis_params(intensityAlpha=40.0)
    vol = torch.full((1, 10, 10, 10), 1.0)
    _, aux = O.render(vol, P, return_aux=True, dtype=torch.float64)
    e = math.exp(-40.0 * float(np.float32(0.05)))
    k = 0
    T = 1.0
    while T > 0.01:
        T *= e; k += 1
    assert int(aux["n_taken"][2, 2]) == k == 3
    assert abs(float(aux["T"][2, 2]) - T) < 1e-12


def test_reference_tf_equals_two_entry_lut():
    """brats_rt.slang:132-140 == LUT [(0,0,0,0),(1,1,1,intensityAlpha)] (test 9, row A7)."""
    vol, _, P = small_scene(C=2, dims=(20, 18, 16), W=24, H=20, seed=3)
    P = replace(P, intensityAlpha=9.0)
    a = O.render(vol, P)
    b = O.render(vol, P, tf=torch.tensor([[0.0, 0, 0, 0], [1.0, 1, 1, 9.0]]))
    assert torch.equal(a, b)


def test_indexed_vs_accumulated_stepping():
    """t_k = t0 + k*dt vs the reference's running sum differ far below the tolerance (test 10, Q4)."""
    vol, _, P = small_scene(C=1, dims=(40, 36, 28), W=32, H=32, seed=1)
    P = replace(P, intensityAlpha=6.0)
    a = O.render(vol, P)
    b = O.render(vol, replace(P, tMode="accumulate"))
    assert float((a - b).abs().max()) < 1e-4


def test_background_is_added_unattenuated_and_y_is_flipped():
    """Q1: C starts at bgColor; Q16: image row 0 is the TOP (camera-space y = -uv.y)."""
    P = _axis_params(W=1, H=8, orthoHalfHeight=0.45, bgColor=(0.0, 0.0, 0.0), intensityAlpha=5.0)
    vol = torch.zeros(1, 10, 10, 10)
    vol[0, :, 8:, :] = 1.0                              # bright slab at high y (top of the world)
    img = O.render(vol, P)
    assert float(img[0, 0, 0]) > 0.1 and float(img[7, 0, 0]) == 0.0


def test_fp64_gradcheck_small():
    """torch.autograd.gradcheck of the oracle in fp64 (test 8)."""
    torch.manual_seed(0)
    dims = (6, 6, 6)
    P = O.params(imageSize=(4, 4), dims=dims, voxelSize=(0.3, 0.3, 0.3), volMin=(-0.9, -0.9, -0.9), stepSize=0.21,
                 eye=(0.4, 0.3, -2.5), tfMode=1, ertThreshold=1e-6)
    vol = (0.2 + 0.6 * torch.rand(1, 6, 6, 6, dtype=torch.float64)).requires_grad_(True)
    tf = torch.rand(5, 4, dtype=torch.float64).requires_grad_(True)
    f = lambda v, t: O.render(v, P, tf=t * torch.tensor([1.0, 1.0, 1.0, 3.0], dtype=torch.float64), dtype=torch.float64)
    assert torch.autograd.gradcheck(f, (vol, tf), eps=1e-6, atol=1e-6, rtol=1e-4, nondet_tol=0.0)
