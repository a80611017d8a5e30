#!/usr/bin/env python3
"""Launches every kernel of the library once at 16 cameras x 2048^2 (level 0 of 5) so that an ncu launch list
(gpu__time_duration) gives each kernel's device time for the roofline table in profiles/README.md."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from facebook360_dep_b200 import capi, synth
W = 2048
rig = synth.ring_rig(16, W, W, kind="FTHETA")
colors, true = synth.render_rig(rig, W, W, device="cuda")
L = capi.load_cuda()
ctx = capi.Context(L, capi.rig_descs(rig))
for rep in range(2):  # second pass = warm geometry cache
    ctx.level_begin(W, W, level=0, num_levels=5, full_width=W, full_height=W)
    ctx.set_colors(colors)
    coarse = np.ascontiguousarray(true[0][::2, ::2])
    ctx.upsample_from(0, coarse)
    ctx.reproject(0)
    ctx.random_proposals(0, 2)
    ctx.ping_pong(0, 1)
    ctx.bilateral(0)
    ctx.median(0)
    ctx.mask_fov(0)
for d in range(16):
    ctx.set_disparity(d, true[d])
ctx.mismatches()
rng = np.random.RandomState(0)
T = 5
guides = [colors[0]] * T
disps = [true[0]] * T
masks = [np.ones((W, W), np.uint8)] * T
L.temporal_filter(guides, disps, masks, 2, 0.01, 1, 0.5, 1.0, 0.5)
L.camera_mesh(true[0], (float(W), float(W)), 651.9)  # K17: the mesh of one level-0 disparity map
L.downscale_area(colors[0], W // 2, W // 2)
print("done")
