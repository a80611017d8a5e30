"""Pins the oracle camera (oracle/camera.h) against
 (1) vectors produced by the reference's own numpy port (tests/golden/camera_vectors.json), and
 (2) the known-answer cases of source/test/util/{FTheta,Rectilinear,Orthographic}Test.cpp.
"""
import json
import math
import os

import numpy as np
import pytest

from facebook360_dep_b200 import capi
from tests import oracle_hooks as oh

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "camera_vectors.json")))


def test_world_to_pixel_matches_reference_port(oracle):
    n = 0
    for rec in GOLD["cameras"]:
        d = capi.camera_desc_from_json(rec["json"])
        pts = np.array([p["p"] for p in rec["points"]])
        pix, sees = oh.camera_project(oracle, d, pts)
        ref = np.array([p["pixel"] for p in rec["points"]])
        ref_sees = np.array([p["sees"] for p in rec["points"]])
        # behind-camera rectilinear points project through tan(pi/2) ~ 1.6e16: compare on-sensor-ish ones
        ok = np.isfinite(ref).all(axis=1) & (np.abs(ref).max(axis=1) < 1e6)
        # the numpy port keeps the raw rig-JSON rotation, the C++ re-unitarises it (Camera.cpp:77-87):
        # agreement is bounded by the JSON's own orthonormality (res/test/cameras/ftheta.json is only
        # orthonormal to 4e-4, which is why FThetaTest.cpp:41 uses isApprox(right, 1e-3))
        cj = rec["json"]
        R = np.array([cj["right"], cj["up"], -np.array(cj["forward"])])
        orth = np.abs(R @ R.T - np.eye(3)).max()
        tol = 1e-6 + 4000.0 * orth
        assert np.abs(pix[ok] - ref[ok]).max() < tol, (cj["id"], orth)
        if orth < 1e-12:
            assert np.array_equal(sees, ref_sees), cj["id"]
        n += len(pts)
    assert n > 800


def test_pixel_to_world_matches_reference_port(oracle):
    for rec in GOLD["cameras"]:
        d = capi.camera_desc_from_json(rec["json"])
        cj = rec["json"]
        R = np.array([cj["right"], cj["up"], -np.array(cj["forward"])])
        orth = np.abs(R @ R.T - np.eye(3)).max()
        for p in rec["pixels"]:
            w, _ = oh.camera_unproject(oracle, d, np.array([p["pixel"]]), p["depth"])
            assert np.abs(w[0] - np.array(p["world"])).max() < (1e-9 + 2 * orth) * max(1.0, p["depth"])


def test_distort_undistort_matches_reference_port(oracle):
    for rec in GOLD["cameras"]:
        d = capi.camera_desc_from_json(rec["json"])
        r = np.array([x["r"] for x in rec["distort"]])
        a, b = oh.camera_distort(oracle, d, r)
        assert np.allclose(a, [x["distort"] for x in rec["distort"]], rtol=1e-13, atol=1e-15)
        assert np.allclose(b, [x["undistort"] for x in rec["distort"]], rtol=1e-12, atol=1e-12)
        _, dmax, _ = oh.camera_info(oracle, d)
        if rec["distortion_max"] is None:
            assert math.isinf(dmax)
        else:
            assert abs(dmax - rec["distortion_max"]) <= 1e-10 * rec["distortion_max"]


# ---- known-answer cases restated from source/test/util/FThetaTest.cpp:19-159 -------------------
FTHETA = {
    "version": 1, "type": "FTHETA",
    "origin": [-10.51814, 13.00734, -4.22656],
    "forward": [-0.6096207796429852, 0.7538922995778138, -0.24496715221587234],
    "up": [0.7686134846014325, 0.6376793279268061, 0.050974366338976666],
    "right": [0.19502945167097138, -0.15702371237098722, -0.9681462011153862],
    "resolution": [2448, 2048], "focal": [1240, -1240], "id": "cam1",
}


def _with(cam, **kw):
    c = dict(cam)
    c.update(kw)
    return c


@pytest.mark.parametrize("kind", ["FTHETA", "RECTILINEAR", "ORTHOGRAPHIC"])
def test_center_pixel_and_round_trip(oracle, kind):
    cam = _with(FTHETA, type=kind)
    d = capi.camera_desc_from_json(cam)
    rot, _, _ = oh.camera_info(oracle, d)
    # TestInitialization: right() ~ json right (1e-3), principal ray -> centre pixel (1e-10)
    assert np.allclose(rot[0], cam["right"], atol=1e-3)
    assert np.allclose(rot @ rot.T, np.eye(3), atol=1e-14)  # re-unitarised
    fwd = -rot[2]
    pix, _ = oh.camera_project(oracle, d, np.array([np.array(cam["origin"]) + fwd]))
    assert abs(pix[0, 0] - 2448 / 2) < 1e-10 and abs(pix[0, 1] - 2048 / 2) < 1e-10
    # TestUndoPixel (CameraTestUtil.cpp:12-20): rig(pixel(p)).pointAt(d) == p to 1e-10 relative
    dist = 3.1
    v = np.array([-2.0, 3.0, -1.0])
    target = np.array(cam["origin"]) + dist * v / np.linalg.norm(v)
    pix, _ = oh.camera_project(oracle, d, np.array([target]))
    back, _ = oh.camera_unproject(oracle, d, pix, dist)
    assert np.linalg.norm(back[0] - target) <= 1e-10 * np.linalg.norm(target)


def test_fov_truth_table(oracle):
    # FThetaTest.cpp TestFOV
    d0 = capi.camera_desc_from_json(FTHETA)
    p11, oc = oh.camera_unproject(oracle, d0, np.array([[1.0, 1.0]]), 1e4)
    _, sees = oh.camera_project(oracle, d0, p11)
    assert sees[0] and not oc[0]
    d1 = capi.camera_desc_from_json(_with(FTHETA, fov=0.1 * math.pi))
    _, _, cosfov = oh.camera_info(oracle, d1)
    assert abs(math.acos(cosfov) - 0.1 * math.pi) < 1e-10
    p11, oc = oh.camera_unproject(oracle, d1, np.array([[1.0, 1.0]]), 1e4)
    _, sees = oh.camera_project(oracle, d1, p11)
    assert not sees[0] and oc[0]
    pc, oc = oh.camera_unproject(oracle, d1, np.array([[1200.0, 1000.0]]), 1e4)
    _, sees = oh.camera_project(oracle, d1, pc)
    assert sees[0] and not oc[0]


def test_distortion_known_answers(oracle):
    # TestDistortUndistort: d = (0.20, 0.02): undistort(distort(2)) == 2 within 1/kNearInfinity
    d = capi.camera_desc_from_json(_with(FTHETA, distortion=[0.20, 0.02]))
    a, _ = oh.camera_distort(oracle, d, np.array([2.0]))
    _, b = oh.camera_distort(oracle, d, a)
    assert abs(b[0] - 2.0) < 1e-4
    # negative real roots -> distortionMax = inf ; imaginary roots -> inf
    for dist in ([2 / 3.0, 1 / 5.0], [1.0, 1.0]):
        dd = capi.camera_desc_from_json(_with(FTHETA, distortion=dist))
        assert math.isinf(oh.camera_info(oracle, dd)[1])
    # TestUndistortMonotonic
    dd = capi.camera_desc_from_json(_with(FTHETA, distortion=[-0.03658484692522479, -0.004515457470690702]))
    ys = np.arange(0, 3, 0.1)
    _, und = oh.camera_distort(oracle, dd, ys)
    assert np.all(np.diff(und) >= -1e-4)


def test_normalize_invariance(oracle):
    # TestNormalize / TestRescale: pixel coordinates scale with the resolution
    d = capi.camera_desc_from_json(FTHETA)
    pts = np.array(FTHETA["origin"]) + np.array([[-2.0, 3.0, -1.0], [-1.0, 2.0, 0.5]])
    pix, _ = oh.camera_project(oracle, d, pts)
    pixn, _ = oh.camera_project(oracle, d, pts, normalized=True)
    assert np.allclose(pixn * np.array([2448, 2048]), pix, rtol=1e-12)


def test_filter_destinations_order():
    # DerpTest.cpp:18-26: filterDestinations("cam4,cam15,cam0") keeps the requested order
    from facebook360_dep_b200 import rigio
    ids = ["cam%d" % i for i in range(16)]
    assert rigio.filter_destinations(ids, "cam4,cam15,cam0") == [4, 15, 0]
    assert rigio.filter_destinations(ids, "") == list(range(16))
