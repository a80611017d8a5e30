#!/usr/bin/env python3
"""Coarse-to-fine run of BASELINE.json configs[1] (16 cameras, 2048^2, 128 candidates, 5 levels) through
derp_process_level, timed per level with CUDA events; prints evaluations and Mpix·cand/s per level."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from facebook360_dep_b200 import capi, synth

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=2048)
ap.add_argument("--cams", type=int, default=16)
ap.add_argument("--levels", type=int, default=5)
ap.add_argument("--depths", type=int, default=128)
a = ap.parse_args()
W = a.size
rig = synth.ring_rig(a.cams, W, W, kind="FTHETA")
colors, _ = synth.render_rig(rig, W, W, device="cuda")
pyr = [colors]
for L in range(1, a.levels):
    pyr.append([synth.downscale_area(c, 2) for c in pyr[-1]])
L_ = capi.load_cuda()
ctx = capi.Context(L_, capi.rig_descs(rig))
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
for rep in range(2):
    prev = None
    tot_ms = tot_e = 0
    for level in range(a.levels - 1, -1, -1):
        w = W >> level
        ctx.level_begin(w, w, level=level, num_levels=a.levels, full_width=W, full_height=W)
        ctx.set_colors(pyr[level])
        if prev is not None:
            for d in range(a.cams):
                ctx.upsample_from(d, prev[d])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.process_level(num_depths=a.depths)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ev, hits = ctx.get_counters()
        prev = [ctx.get_disparity(d, want_cost=False) for d in range(a.cams)]
        if rep == 1:
            print("level %d (%4d^2): %8.2f ms  %7.1f Mevals  %8.1f Mpix·cand/s  vbar %.2f" % (level, w, ms, ev / 1e6, ev / ms / 1e3, hits / max(ev, 1)))
        tot_ms += ms
        tot_e += ev
    if rep == 1:
        print("TOTAL: %.1f ms, %.1f Mevals, %.1f Mpix·cand/s" % (tot_ms, tot_e / 1e6, tot_e / tot_ms / 1e3))
