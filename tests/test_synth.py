"""The seeded synthetic volumes: determinism and the box-wise generator
used for the brick-sharded configuration."""
import torch

from mri_raytracer_b200.synth import make_brats_like, make_brats_like_box, ramp_tf
from mri_raytracer_b200 import dist as mdist


def test_make_brats_like_is_deterministic_and_brats_shaped():
    a = make_brats_like(2, (30, 28, 20), seed=3)
    b = make_brats_like(2, (30, 28, 20), seed=3)
    assert torch.equal(a, b) and a.shape == (2, 20, 28, 30)
    assert float(a.min()) == 0.0 and float(a.max()) <= 1.0
    frac = float((a[0] > 0).float().mean())
    assert 0.15 < frac < 0.35          # skull-stripped look: ~25 % of the voxels are non-zero
    assert not torch.equal(a, make_brats_like(2, (30, 28, 20), seed=4))


def test_box_generator_is_a_pure_function_of_the_global_index():
    dims = (37, 29, 23)
    full = make_brats_like_box(dims, (0, 0, 0), (36, 28, 22), seed=3, dtype=torch.float32, zchunk=5)
    assert full.shape == (1, 23, 29, 37)
    # any sub-box, any slab size: the same voxels
    box = make_brats_like_box(dims, (5, 7, 9), (20, 28, 15), seed=3, dtype=torch.float32, zchunk=4)
    assert torch.equal(box, full[:, 9:16, 7:29, 5:21])
    # the shards of a 2x2x2 grid tile the volume (cells once, +1 halo voxel shared)
    grid = (2, 2, 2)
    for r in range(8):
        lo, hi, _ = mdist.shard_box(dims, grid, r)
        sub = make_brats_like_box(dims, lo, hi, seed=3, dtype=torch.float16)
        assert torch.equal(sub, mdist.slice_shard(full, lo, hi).half())
    frac = float((full > 0).float().mean())
    assert 0.15 < frac < 0.35


def test_ramp_tf_shape():
    tf = ramp_tf(64)
    assert tf.shape == (64, 4) and float(tf[0, 3]) == 0.0 and float(tf[-1, 3]) == 40.0
