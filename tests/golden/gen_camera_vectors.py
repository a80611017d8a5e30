#!/usr/bin/env python3
"""Generates tests/golden/camera_vectors.json by IMPORTING THE REFERENCE'S OWN numpy camera port
(/root/reference/scripts/util/camera.py, rig.py) — run in the build container only; the GPU box has
no /root/reference, which is why the vectors are committed.

For every camera of res/test/rigs/rig.json (16 FTHETA cams with distortion and fov) and the three
single-camera fixtures res/test/cameras/{ftheta,rectilinear,orthographic}.json it records
  world_to_pixel(p), sees(p)             for seeded random rig-space points
  pixel_to_world(px, depth)              for seeded random pixels
  distort(r), undistort(r), distortion_max
Not recorded: is_outside_image_circle — the numpy port compares |sensor|^2 with |edge| (not
squared, camera.py:189-199), which differs from the C++ it was ported from (Camera.h:166-178).
"""
import json
import os
import sys

import numpy as np

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "scripts", "util"))
from camera import Camera  # noqa: E402


def main():
    out = {"generator": "tests/golden/gen_camera_vectors.py", "source": "scripts/util/camera.py", "cameras": []}
    cam_jsons = []
    rig = json.load(open(os.path.join(REF, "res/test/rigs/rig.json")))
    cam_jsons += rig["cameras"]
    for name in ("ftheta", "rectilinear", "orthographic"):
        p = os.path.join(REF, "res/test/cameras", name + ".json")
        if os.path.exists(p):
            cam_jsons.append(json.load(open(p)))
    rng = np.random.RandomState(20260924)
    for cj in cam_jsons:
        cam = Camera(json_string=json.dumps(cj))
        rec = {"json": cj, "points": [], "pixels": [], "distort": []}
        pos = np.asarray(cam.position, dtype=float)
        for _ in range(48):
            d = rng.uniform(0.4, 30.0)
            v = rng.normal(size=3)
            v /= np.linalg.norm(v)
            if rng.rand() < 0.6:  # bias towards the forward hemisphere
                v = v + 1.5 * np.asarray(cam.forward())
                v /= np.linalg.norm(v)
            p = pos + d * v
            pix = cam.world_to_pixel(p)
            sees, _ = cam.sees(p)
            rec["points"].append({"p": p.tolist(), "pixel": [float(pix[0]), float(pix[1])], "sees": bool(sees)})
        res = np.asarray(cam.resolution, dtype=float)
        for _ in range(32):
            px = np.array([rng.uniform(0.15, 0.85) * res[0], rng.uniform(0.15, 0.85) * res[1]])
            depth = float(rng.uniform(0.5, 50.0))
            w = cam.pixel_to_world(px, depth)
            rec["pixels"].append({"pixel": px.tolist(), "depth": depth, "world": np.asarray(w).tolist()})
        for r in np.linspace(0.0, 1.6, 17):
            rec["distort"].append({"r": float(r), "distort": float(cam.distort(r)), "undistort": float(cam.undistort(r))})
        dm = cam.get_distortion_max()
        rec["distortion_max"] = None if dm > 1e49 else float(dm)
        out["cameras"].append(rec)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "camera_vectors.json")
    json.dump(out, open(dst, "w"))
    print("wrote", dst, len(out["cameras"]), "cameras")


if __name__ == "__main__":
    main()
