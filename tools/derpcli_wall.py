#!/usr/bin/env python3
"""DerpCLI files-in -> files-out wall time per frame on the headline rig (16 cameras, 2048^2, 5 levels, 128 candidates):
PNG decode on host threads, uploads, the five levels handed over in device memory, PFM writing overlapped with the
next level.  Usage: python tools/derpcli_wall.py [--frames 2] [--size 2048]"""
import argparse, json, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
from facebook360_dep_b200 import capi, synth

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--size", type=int, default=2048)
ap.add_argument("--cams", type=int, default=16)
ap.add_argument("--levels", type=int, default=5)
a = ap.parse_args()
W = H = a.size
cuda = capi.load_cuda()
rig = synth.ring_rig(a.cams, W, H, kind="FTHETA")
root = tempfile.mkdtemp(prefix="derpcli_wall_")
inp, out = os.path.join(root, "in"), os.path.join(root, "out")
os.makedirs(os.path.join(inp, "rigs"))
json.dump(rig, open(os.path.join(inp, "rigs", "rig_calibrated.json"), "w"))
t0 = time.time()
for f in range(a.frames):
    colors, _ = synth.render_rig(rig, W, H, scene=synth.Scene(seed=42, shift=(0.01 * f, 0, 0)), device="cuda")
    for s, c in enumerate(colors):
        c = np.ascontiguousarray(c)
        for k in range(a.levels):
            img = c if k == 0 else cuda.downscale_area(c, W >> k, H >> k)  # scripts/render/resize.py: every level from the full size
            d = os.path.join(inp, "video", "color_levels", "level_%d" % k, rig["cameras"][s]["id"])
            os.makedirs(d, exist_ok=True)
            cv2.imwrite(os.path.join(d, "%06d.png" % f), img, [cv2.IMWRITE_PNG_COMPRESSION, 1])
print("dataset written in %.1fs" % (time.time() - t0), flush=True)
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "facebook360_dep_b200", "bin", "DerpCLI")
for label, extra in (("cold (geometry caches empty, first frames)", []),):
    t0 = time.time()
    p = subprocess.run([exe, "--input_root=" + inp, "--output_root=" + out, "--first=000000", "--last=%06d" % (a.frames - 1),
                        "--partial_coverage=true", "--num_depths=128"] + extra, capture_output=True, text=True)
    dt = time.time() - t0
    if p.returncode != 0:
        print(p.stderr[-2000:])
        raise SystemExit(1)
    n = sum(len(fs) for _, _, fs in os.walk(out))
    print("DerpCLI %s: %.2fs wall for %d frame(s) = %.2f s/frame, %d output files" % (label, dt, a.frames, dt / a.frames, n), flush=True)
    for line in p.stderr.splitlines():
        if "TOTAL" in line or "Elapsed" in line:
            print("  " + line.strip())
subprocess.run(["rm", "-rf", root])
