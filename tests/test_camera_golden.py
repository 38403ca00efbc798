"""Orbital cameras vs golden vectors produced by EXECUTING the reference classes
(tests/golden/make_golden.py; reference inr/viewer/camera.py and scripts/raymarch/camera.py)."""
import json
from pathlib import Path

import numpy as np
import pytest

from mri_raytracer_b200.camera import OrbitalCamera, OrbitalCameraYUp

G = Path(__file__).parent / "golden"


def _run(cls, case, arbitrary_up):
    kw = dict(initial_radius=case["radius"], initial_phi=case["phi"], initial_theta=case["theta"], min_radius=1e-9)
    if case["target"] is not None:
        kw["initial_target"] = np.array(case["target"], dtype=np.float32)
    if arbitrary_up and case["up"] is not None:
        kw["world_up"] = np.array(case["up"], dtype=np.float32)
    cam = cls(**kw)
    cam.set_fov_degrees(case["fov_deg"])
    for op in case["ops"]:
        if op[0] == "orbit":
            cam.orbit(op[1], op[2])
        elif op[0] == "zoom":
            cam.zoom(op[1])
        elif arbitrary_up:
            cam.pan(op[1], op[2], viewport_height=480.0)
        else:
            cam.pan(op[1], op[2])
    return cam


@pytest.mark.parametrize("fname,cls,arb", [("camera_arbitrary_up.json", OrbitalCamera, True),
                                           ("camera_yup.json", OrbitalCameraYUp, False)])
def test_camera_matches_reference_bitwise(fname, cls, arb):
    rows = json.loads((G / fname).read_text())["rows"]
    assert len(rows) >= 40
    for row in rows:
        cam = _run(cls, row["case"], arb)
        eye, right, up, fwd = cam.get_basis()
        exp = row["out"]
        for got, key in ((eye, "eye"), (right, "right"), (up, "up"), (fwd, "forward"),
                         (cam.get_eye_position(), "eye_only"), (cam.target, "target")):
            got = np.asarray(got)
            assert got.dtype == np.float32
            want = np.array(exp[key], dtype=np.float32)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (fname, key, row["case"], got, want)
        assert float(cam.radius) == exp["radius"] and cam.phi == exp["phi"] and cam.theta == exp["theta"]


def test_reset_and_limits():
    cam = OrbitalCamera(initial_radius=3.0)
    cam.orbit(1.0, 10.0)
    assert cam.phi == cam.max_phi
    cam.zoom(1e9)
    assert cam.radius == cam.max_radius
    cam.reset()
    assert cam.radius == 3.0 and cam.theta == 0.0
