"""Shared scene builders for the tests: the same RenderParams object is handed to the CUDA
path and (duck-typed, by field name) to the oracle."""
from __future__ import annotations

import math
from dataclasses import replace

import numpy as np
import torch

from mri_raytracer_b200 import Camera, OrbitalCamera, RenderParams
from mri_raytracer_b200.synth import make_brats_like, ramp_tf, world_box


def framed_params(dims, W, H, theta_deg=25.0, phi_deg=80.0, fov_deg=70.0, ortho=False, step_vox=0.5,
                  zooms=(1.0, 1.0, 1.0), radius_scale=0.8, **kw) -> RenderParams:
    """Reference framing (inr/viewer/brats_viewer.py:204-210,320-324) + the bench camera angles."""
    vs, vmin = world_box(dims, zooms)
    ext = vs * np.asarray(dims, dtype=np.float32)
    cam = OrbitalCamera(initial_radius=3.0, initial_theta=math.radians(theta_deg),
                        initial_phi=math.radians(phi_deg))
    cam.set_fov_degrees(fov_deg)
    cam.target = (vmin + 0.5 * ext).astype(np.float32)
    cam.radius = float(np.linalg.norm(ext) * radius_scale)
    c = Camera.from_orbital(cam, ortho=ortho)
    P = RenderParams(imageSize=(W, H), dims=tuple(dims), voxelSize=tuple(float(v) for v in vs),
                     volMin=tuple(float(v) for v in vmin), stepSize=float(np.float32(step_vox) * vs[0]), **kw)
    return P.with_camera(c)


def small_scene(C=1, dims=(40, 36, 28), W=48, H=40, seed=0, labels=False, **kw):
    out = make_brats_like(C, dims, seed=seed, with_labels=labels)
    vol, lab = (out if labels else (out, None))
    P = framed_params(dims, W, H, **kw)
    return vol, lab, P
