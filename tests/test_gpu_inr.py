"""INR inference on the GPU (mrt_inr_predict) against the numpy
restatement of the reference's predict_volume (oracle/oracle_inr.py): logits max-abs 1e-4, labels
identical wherever the oracle's top-2 logits are more than 1e-4 apart; and the label volume feeds
the renderer's prediction overlay."""
from dataclasses import replace

import numpy as np
import pytest
import torch

from mri_raytracer_b200 import api, volume as mvol
from oracle import oracle_inr as I

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("impl", ["ffma", "auto"])
@pytest.mark.parametrize("dims,M,k,hidden", [((23, 19, 17), 4, 4, [64, 64, 64, 64]), ((9, 8, 7), 2, 2, [32, 16]),
                                              ((12, 11, 10), 4, 0, [])])
def test_inr_predict_matches_oracle(cuda, dims, M, k, hidden, impl):
    X, Y, Z = dims
    rng = np.random.default_rng(7)
    params = I.init_mlp(rng, I.input_dim(M, k), hidden, 4)
    for p in params:                                     # non-zero biases exercise the bias path
        p["b"] = rng.normal(scale=0.1, size=p["b"].shape).astype(np.float32)
    raw = torch.from_numpy(rng.gamma(2.0, 50.0, size=(M, Z, Y, X)).astype(np.float32))
    raw[:, :, :2] = 0.0                                  # background zeros, like skull-stripped MRI
    mods = mvol.zscore_modalities(raw)
    labels, logits = api.inr_predict(mods.cuda(), params, k, return_logits=True, impl=impl)
    # oracle works in the reference's [M,H,W,D] = [M,X,Y,Z] order
    pred, want = I.predict_volume(params, mods.numpy().transpose(0, 3, 2, 1), k, chunk=1000, return_logits=True)
    got = logits.cpu().numpy().transpose(2, 1, 0, 3)     # [Z,Y,X,c] -> [X,Y,Z,c]
    assert np.abs(got - want).max() <= 1e-4
    lab = labels.cpu().numpy()
    assert lab.shape == (Z, Y, X)
    assert np.array_equal(I.to_renderer_labels(pred).shape, lab.shape)
    top2 = np.sort(want, axis=-1)[..., -2:]
    clear = (top2[..., 1] - top2[..., 0]) > 1e-4
    same = lab.transpose(2, 1, 0) == pred
    assert bool(same[clear].all()) and float(same.mean()) > 0.999


def test_zscore_matches_viewer_preprocessing(cuda):
    rng = np.random.default_rng(1)
    raw = rng.gamma(2.0, 50.0, size=(2, 6, 5, 4)).astype(np.float32)
    raw[1] = 0.0                                         # an all-zero modality stays as it is
    raw[0, :2] = 0.0
    got = mvol.zscore_modalities(torch.from_numpy(raw).cuda()).cpu().numpy()
    arr = raw[0]; mask = arr != 0
    want0 = (arr - arr[mask].mean()) / (arr[mask].std() + 1e-6)      # inr/viewer/brats_viewer.py:279-287
    assert np.abs(got[0] - want0).max() <= 1e-5 and np.array_equal(got[1], raw[1])


def test_inr_labels_feed_the_prediction_overlay(cuda):
    import sys
    from scenes import small_scene
    from parity import O
    vol, _, P = small_scene(C=4, dims=(20, 18, 16), W=32, H=24, seed=2)
    rng = np.random.default_rng(3)
    params = I.init_mlp(rng, I.input_dim(4, 2), [32, 32], 4)
    preds = api.inr_predict(mvol.zscore_modalities(vol).cuda(), params, 2)
    assert int((preds > 0).sum()) > 0
    P = replace(P, showPred=1, intensityAlpha=5.0)
    V = api.Volume(vol.cuda(), preds=preds)
    img = api.render(V, None, None, P).cpu()
    ref = O.render(vol, P, preds=preds.cpu().long())
    assert (img - ref).abs().max() <= 1e-4


def test_tensor_core_inr_equals_the_fp32_kernel(cuda):
    """The tcgen05 kernel (3-term TF32 split, fp32 accumulation in TMEM) against the fp32 FFMA kernel
    on a volume of many 128-voxel tiles with a ragged tail: logits within 2e-5, labels identical
    wherever the top-2 logits are more than 1e-3 apart; the deepest network the API takes still fits;
    a single dense layer is refused by impl="tensor" and served by impl="auto"."""
    X, Y, Z, M, k = 131, 37, 29, 4, 4
    rng = np.random.default_rng(11)
    params = I.init_mlp(rng, I.input_dim(M, k), [64, 64, 64, 64], 4)
    for p in params:
        p["b"] = rng.normal(scale=0.2, size=p["b"].shape).astype(np.float32)
    raw = torch.from_numpy(rng.gamma(2.0, 50.0, size=(M, Z, Y, X)).astype(np.float32))
    mods = mvol.zscore_modalities(raw).cuda()
    lab_f, log_f = api.inr_predict(mods, params, k, return_logits=True, impl="ffma")
    lab_t, log_t = api.inr_predict(mods, params, k, return_logits=True, impl="tensor")
    assert float((log_t - log_f).abs().max()) <= 2e-5
    top2 = torch.sort(log_f, dim=-1).values[..., -2:]
    clear = (top2[..., 1] - top2[..., 0]) > 1e-3
    assert bool((lab_t == lab_f)[clear].all()) and float((lab_t == lab_f).float().mean()) > 0.9999
    deep = I.init_mlp(rng, I.input_dim(M, k), [64] * 6, 4)            # 7 layers: 190 KB of tf32 hi/lo weight images + tables still fit
    ld, gd = api.inr_predict(mods, deep, k, return_logits=True, impl="tensor")
    _, gdf = api.inr_predict(mods, deep, k, return_logits=True, impl="ffma")
    assert float((gd - gdf).abs().max()) <= 2e-5
    linear = I.init_mlp(rng, I.input_dim(M, k), [], 4)                # a single dense layer has no hidden activations to feed back
    with pytest.raises(api._lib.MrtError):
        api.inr_predict(mods, linear, k, impl="tensor")
    assert api.inr_predict(mods, linear, k, impl="auto").shape == (Z, Y, X)
