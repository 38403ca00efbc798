"""Known-answer tests of the INR inference restatement (oracle/oracle_inr.py; SURVEY.md section 8(f)
rank 3 groundwork).  The reference (JAX) cannot run here: parity unpinned, see the oracle's header."""
import math

import numpy as np

from oracle import oracle_inr as I


def test_fourier_feature_order_and_values():
    c = np.array([[0.5, -1.0, 0.25]], dtype=np.float32)
    ff = I.fourier_features(c, 2)
    assert ff.shape == (1, 12)
    # per coordinate: sin(1*pi*x), sin(2*pi*x), cos(1*pi*x), cos(2*pi*x)
    want = []
    for x in c[0]:
        want += [math.sin(math.pi * x), math.sin(2 * math.pi * x), math.cos(math.pi * x), math.cos(2 * math.pi * x)]
    assert np.allclose(ff[0], want, atol=1e-6)
    x = I.build_input(c, np.array([[1, 2, 3, 4]], dtype=np.float32), 2)
    assert x.shape == (1, I.input_dim(4, 2)) and np.array_equal(x[0, :3], c[0]) and np.array_equal(x[0, -4:], [1, 2, 3, 4])


def test_mlp_known_answer_and_relu():
    params = [{"W": np.array([[1.0, -1.0], [0.5, 2.0]], dtype=np.float32), "b": np.array([0.0, -1.0], dtype=np.float32)},
              {"W": np.array([[1.0], [3.0]], dtype=np.float32), "b": np.array([0.5], dtype=np.float32)}]
    x = np.array([[2.0, 1.0], [-1.0, 0.25]], dtype=np.float32)
    # layer 1: [2.5, -1] -> relu [2.5, 0];  [-0.875, 0.5] -> relu [0, 0.5]
    assert np.allclose(I.apply_mlp(params, x), [[3.0], [2.0]])


def test_predict_volume_selects_the_brightest_modality_and_layout():
    """With an MLP that forwards the 4 intensities to the 4 logits, the prediction is the argmax
    modality per voxel; coordinates follow the ij meshgrid; chunking does not matter."""
    M, H, W, D, k = 4, 5, 4, 3, 2
    nin = I.input_dim(M, k)
    Wm = np.zeros((nin, 4), dtype=np.float32)
    Wm[-4:, :] = np.eye(4, dtype=np.float32)
    params = [{"W": Wm, "b": np.zeros(4, dtype=np.float32)}]                  # a single (last) layer: no ReLU
    rng = np.random.default_rng(0)
    mods = rng.normal(size=(M, H, W, D)).astype(np.float32)
    pred = I.predict_volume(params, mods, k, chunk=7)
    assert pred.shape == (H, W, D) and pred.dtype == np.int16
    assert np.array_equal(pred, np.argmax(mods, axis=0))
    assert np.array_equal(pred, I.predict_volume(params, mods, k, chunk=10_000))
    lab = I.to_renderer_labels(pred)
    assert lab.shape == (D, W, H) and lab.dtype == np.int32 and lab[2, 1, 4] == pred[4, 1, 2]
    # an MLP that reads only the first normalised coordinate: sign splits the volume along H
    Wc = np.zeros((nin, 2), dtype=np.float32); Wc[0, 0] = 1.0; Wc[0, 1] = -1.0
    p2 = I.predict_volume([{"W": Wc, "b": np.zeros(2, dtype=np.float32)}], mods, k)
    assert np.array_equal(p2[:, 0, 0], [1, 1, 0, 0, 0])                       # x = -1, -.5 -> class 1; 0 ties to 0; .5, 1 -> 0


def test_reference_sized_model_runs():
    rng = np.random.default_rng(1)
    params = I.init_mlp(rng, I.input_dim(4, 4), [64, 64, 64, 64], 4)           # inr/interactive.ipynb cell 1 config
    mods = rng.normal(size=(4, 12, 10, 8)).astype(np.float32)
    pred = I.predict_volume(params, mods, 4)
    assert pred.shape == (12, 10, 8) and 0 <= int(pred.min()) and int(pred.max()) <= 3
