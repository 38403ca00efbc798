"""The kernels' own pixel <-> tile <-> lane indexing, evaluated on the device, is bit-exact
with tiles.py."""
import pytest
import torch

from mri_raytracer_b200 import api, tiles

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H", [(1, 1), (8, 8), (9, 7), (70, 33), (1024, 1024), (513, 259)])
def test_device_tile_map_bit_exact(cuda, W, H):
    t, l = api.tile_index_map(W, H)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    want_t = (ys >> 3) * tiles.tiles_x(W) + (xs >> 3)
    want_l = ((ys & 7) << 3) + (xs & 7)
    assert torch.equal(t.cpu().long(), want_t) and torch.equal(l.cpu().long(), want_l)
    assert tiles.tile_of_pixel(W - 1, H - 1, W) == int(t[-1, -1]) == tiles.tile_count(W, H) - 1
