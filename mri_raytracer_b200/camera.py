"""Host-side orbital cameras and the per-frame ray spec.

Mirrors the reference's two camera classes so a viewer can switch over unchanged:

* :class:`OrbitalCamera`      — arbitrary world-up; reference ``inr/viewer/camera.py:8-130``
  (byte-identical copy at ``scripts/brats/camera.py``).
* :class:`OrbitalCameraYUp`   — the original Y-up camera; reference
  ``scripts/raymarch/camera.py:8-114`` (used by ``scripts/volumeRendering/app.py:14``).

Same constructor arguments, same methods (``get_eye_position``, ``get_basis``, ``orbit``,
``pan``, ``zoom``, ``reset``, ``set_fov_degrees``, ``set_aspect``), same float32/float64
rounding sequence: ``tests/test_camera_golden.py`` checks both against vectors produced by
importing the reference classes (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

_F = np.float32


def _f3(x, y, z) -> np.ndarray:
    return np.array([x, y, z], dtype=_F)


def _unit(v: np.ndarray, length: float) -> np.ndarray:
    return (v / length).astype(_F)


class _OrbitBase:
    """State + controls shared by both cameras (spherical orbit about ``target``)."""

    def __init__(self, initial_target=None, initial_radius=2.0, initial_phi=math.pi * 0.5,
                 initial_theta=0.0, min_radius=0.1, max_radius=100.0, min_phi=0.01,
                 max_phi=math.pi - 0.01, aspect=16.0 / 9.0, fovY_radians=math.radians(55.0),
                 near=0.1, far=1000.0):
        t0 = _f3(0.0, 0.0, 0.0) if initial_target is None else np.asarray(initial_target).astype(_F)
        self._home = dict(target=t0, radius=float(initial_radius), phi=float(initial_phi),
                          theta=float(initial_theta), min_radius=float(min_radius),
                          max_radius=float(max_radius), min_phi=float(min_phi), max_phi=float(max_phi))
        self.reset()
        self.fovY_radians = float(fovY_radians)
        self.aspect = float(aspect)
        self.near = float(near)
        self.far = float(far)

    def reset(self):
        h = self._home
        self.target = h["target"].copy()
        self.radius, self.phi, self.theta = h["radius"], h["phi"], h["theta"]
        self.min_radius, self.max_radius = h["min_radius"], h["max_radius"]
        self.min_phi, self.max_phi = h["min_phi"], h["max_phi"]

    # controls ------------------------------------------------------------------
    def orbit(self, d_theta: float, d_phi: float):
        self.theta += float(d_theta)
        self.phi = max(self.min_phi, min(self.max_phi, self.phi + float(d_phi)))

    def zoom(self, factor: float):
        self.radius = max(self.min_radius, min(self.max_radius, self.radius * float(factor)))

    def set_fov_degrees(self, fov_deg: float):
        self.fovY_radians = math.radians(float(fov_deg))

    def set_aspect(self, aspect: float):
        self.aspect = float(aspect)

    def _pan(self, dx: float, dy: float, pixels: float):
        _, right, up, _ = self.get_basis()
        world_h = 2.0 * self.radius * math.tan(max(1e-3, self.fovY_radians * 0.5))
        scale = world_h / pixels
        self.target = (self.target - right * (float(dx) * scale) + up * (float(dy) * scale)).astype(_F)

    def _forward(self, eye: np.ndarray) -> np.ndarray:
        fwd = self.target - eye
        length = float(np.linalg.norm(fwd))
        return _f3(0.0, 0.0, -1.0) if length < 1e-6 else _unit(fwd, length)


class OrbitalCamera(_OrbitBase):
    """Arbitrary-up orbital camera (reference ``inr/viewer/camera.py``)."""

    def __init__(self, *args, world_up: Optional[np.ndarray] = None, **kw):
        self.world_up = _f3(0.0, 1.0, 0.0) if world_up is None else np.asarray(world_up).astype(_F)
        super().__init__(*args, **kw)

    def _base_frame(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(right, front, up) frame the spherical angles are measured in (camera.py:62-77)."""
        up = self.world_up
        ref = _f3(0.0, 0.0, 1.0)
        if abs(float(np.dot(up, ref))) > 0.999:
            ref = _f3(1.0, 0.0, 0.0)
        right = np.cross(ref, up)
        rl = float(np.linalg.norm(right))
        if rl < 1e-6:
            right, rl = _f3(1.0, 0.0, 0.0), 1.0
        right = _unit(right, rl)
        front = np.cross(up, right).astype(_F)
        fl = float(np.linalg.norm(front))
        if fl > 0:
            front = _unit(front, fl)
        return right, front, up

    def get_eye_position(self) -> np.ndarray:
        right, front, up = self._base_frame()
        sp, cp = math.sin(self.phi), math.cos(self.phi)
        offset = (sp * math.cos(self.theta)) * right + (sp * math.sin(self.theta)) * front + cp * up
        return (self.target + self.radius * offset.astype(_F)).astype(_F)

    def get_basis(self):
        """-> (eye, right, up, forward), float32 (camera.py:87-107)."""
        eye = self.get_eye_position()
        fwd = self._forward(eye)
        right = np.cross(fwd, self.world_up)
        rl = float(np.linalg.norm(right))
        if rl < 1e-6:
            right = self._base_frame()[0]
            rl = float(np.linalg.norm(right))
        if rl > 0:
            right = _unit(right, rl)
        up = np.cross(right, fwd).astype(_F)
        if float(np.dot(up, self.world_up)) < 0.0:     # keep the image upright
            up, right = -up, -right
        return eye.astype(_F), right, up, fwd

    def pan(self, dx: float, dy: float, viewport_height: Optional[float] = None):
        ok = viewport_height is not None and viewport_height > 0
        self._pan(dx, dy, max(1.0, float(viewport_height) if ok else 720.0))


class OrbitalCameraYUp(_OrbitBase):
    """Y-up orbital camera (reference ``scripts/raymarch/camera.py``): theta = 0 sits on +X."""

    def get_eye_position(self) -> np.ndarray:
        sp, cp = math.sin(self.phi), math.cos(self.phi)
        return _f3(self.target[0] + self.radius * sp * math.cos(self.theta),
                   self.target[1] + self.radius * cp,
                   self.target[2] + self.radius * sp * math.sin(self.theta))

    def get_basis(self):
        """-> (eye, right, up, forward), float32 (camera.py:70-88)."""
        eye = self.get_eye_position()
        fwd = self._forward(eye)
        right = np.cross(fwd, _f3(0.0, 1.0, 0.0))
        rl = float(np.linalg.norm(right))
        if rl < 1e-6:
            right = np.cross(fwd, _f3(0.0, 0.0, 1.0))
            rl = float(np.linalg.norm(right))
        if rl > 0:
            right = _unit(right, rl)
        up = np.cross(right, fwd).astype(_F)
        return eye.astype(_F), right, up, fwd

    def pan(self, dx: float, dy: float):
        self._pan(dx, dy, 720.0)


@dataclass(frozen=True)
class Camera:
    """What one frame needs from a camera: eye + basis (+ projection)."""
    eye: np.ndarray
    U: np.ndarray
    V: np.ndarray
    W: np.ndarray
    fovY: float
    ortho: bool = False
    ortho_half_height: float = 1.0

    @staticmethod
    def from_orbital(cam: _OrbitBase, ortho: bool = False) -> "Camera":
        eye, right, up, fwd = cam.get_basis()
        # ortho frames the target plane exactly like the pinhole does
        half_h = float(cam.radius) * math.tan(0.5 * cam.fovY_radians)
        return Camera(eye=eye, U=right, V=up, W=fwd, fovY=cam.fovY_radians, ortho=ortho,
                      ortho_half_height=half_h)


def orbit_views(cam: _OrbitBase, n: int, ortho: bool = False):
    """theta_k = theta_0 + 2*pi*k/n (BASELINE config 4: 64 orbit views)."""
    th0 = cam.theta
    out = []
    for k in range(n):
        cam.theta = th0 + 2.0 * math.pi * k / n
        out.append(Camera.from_orbital(cam, ortho=ortho))
    cam.theta = th0
    return out
