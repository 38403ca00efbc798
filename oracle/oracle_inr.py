"""CPU (numpy) restatement of the reference's INR inference path.  TEST INFRASTRUCTURE ONLY.

The checker of SURVEY.md section 8(f) rank 3: INR ``predict_volume`` on the GPU (``mrt_inr_predict``,
csrc/inr.cu — the producer of the ``gPreds`` label volume the renderer overlays,
inr/viewer/brats_viewer.py:250-310).  Only tests/, ``__graft_entry__.smoke()`` and the benchmark's CPU
baseline import this file; nothing under ``mri_raytracer_b200/`` does.

PARITY STATUS: **parity unpinned** — the reference implements this in JAX (inr/inr/model.py), which
is not installed in this image, so the restatement cannot be checked against the executed reference;
it is checked by known-answer tests (tests/test_oracle_inr.py).  All arithmetic is float32, JAX's
default.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

F32 = np.float32


def fourier_features(coords: np.ndarray, k: int) -> np.ndarray:
    """inr/inr/model.py:11-18.  coords [B,dim] -> [B, dim*2k]; per coordinate the k sines
    (frequencies 1..k times pi) followed by the k cosines."""
    coords = np.asarray(coords, dtype=F32)
    B, dim = coords.shape
    freqs = np.arange(1, k + 1).astype(F32)                                   # :13
    ang = coords[..., None] * freqs[None, None, :] * F32(math.pi)             # :14
    ff = np.concatenate([np.sin(ang), np.cos(ang)], axis=-1)                  # :15-17  [B,dim,2k]
    return ff.reshape(B, dim * 2 * k).astype(F32)


def build_input(coords: np.ndarray, intensities: np.ndarray, fourier_freqs: int) -> np.ndarray:
    """inr/inr/model.py:21-23: [coords | fourier features | intensities]."""
    return np.concatenate([np.asarray(coords, dtype=F32), fourier_features(coords, fourier_freqs),
                           np.asarray(intensities, dtype=F32)], axis=-1)


def apply_mlp(params: Sequence[dict], x: np.ndarray) -> np.ndarray:
    """inr/inr/model.py:43-50: dense + ReLU for every layer but the last, dense for the last."""
    *hidden, last = params
    h = np.asarray(x, dtype=F32)
    for layer in hidden:
        h = np.maximum(h @ np.asarray(layer["W"], dtype=F32) + np.asarray(layer["b"], dtype=F32), F32(0))
    return h @ np.asarray(last["W"], dtype=F32) + np.asarray(last["b"], dtype=F32)


def predict_volume(params: Sequence[dict], mods: np.ndarray, fourier_freqs: int, chunk: int = 200000,
                   return_logits: bool = False):
    """inr/inr/model.py:119-141.  mods [M,H,W,D] (z-scored per modality by the caller,
    inr/viewer/brats_viewer.py:279-287) -> int16 labels [H,W,D]: coordinates on the ``ij`` meshgrid
    normalised to [-1,1] by (n-1), argmax over the logits."""
    mods = np.asarray(mods, dtype=F32)
    M, H, W, D = mods.shape
    xs, ys, zs = np.arange(H), np.arange(W), np.arange(D)
    grid = np.stack(np.meshgrid(xs, ys, zs, indexing="ij"), axis=-1).reshape(-1, 3)          # :125
    intens = mods.transpose(1, 2, 3, 0).reshape(-1, M)                                       # :126
    norm = (grid / np.array([H - 1, W - 1, D - 1])) * 2.0 - 1.0                              # :128 (float64, then cast)
    preds, logs = [], []
    for i in range(0, len(grid), chunk):                                                     # :131
        x_in = build_input(norm[i:i + chunk].astype(F32), intens[i:i + chunk], fourier_freqs)
        logits = apply_mlp(params, x_in)
        preds.append(np.argmax(logits, axis=-1).astype(np.int16))                            # :135-137
        if return_logits:
            logs.append(logits)
    pred = np.concatenate(preds, axis=0).reshape(H, W, D)                                    # :139-140
    if return_logits:
        return pred, np.concatenate(logs, axis=0).reshape(H, W, D, -1)
    return pred


def to_renderer_labels(pred_hwd: np.ndarray) -> np.ndarray:
    """The viewer's hand-off to the renderer (inr/viewer/brats_viewer.py:293-299): [X,Y,Z] -> flat
    [Z][Y][X] uint32 — here int32 [Z,Y,X], the layout ``api.Volume(preds=...)`` takes."""
    return np.ascontiguousarray(np.transpose(np.asarray(pred_hwd), (2, 1, 0)).astype(np.int32))


def init_mlp(rng: np.random.Generator, in_dim: int, hidden_dims: Sequence[int], out_dim: int) -> List[dict]:
    """Glorot-uniform weights, zero biases (inr/inr/model.py:26-40); numpy RNG, so the values differ
    from a JAX-initialised model — only the distribution and shapes follow the reference."""
    dims = [in_dim] + list(hidden_dims) + [out_dim]
    out = []
    for a, b in zip(dims[:-1], dims[1:]):
        lim = math.sqrt(6.0 / (a + b))
        out.append({"W": rng.uniform(-lim, lim, size=(a, b)).astype(F32), "b": np.zeros((b,), dtype=F32)})
    return out


def input_dim(n_modalities: int, fourier_freqs: int) -> int:
    """3 coordinates + 3*2k Fourier features + M intensities (inr/inr/train.py: in_dim)."""
    return 3 + 3 * 2 * fourier_freqs + n_modalities
