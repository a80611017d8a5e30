#!/usr/bin/env python3
"""Per-stage timing at one fine level: random proposals and ping-pong for one destination, with work counters."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from facebook360_dep_b200 import capi, synth
ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=2048)
ap.add_argument("--cams", type=int, default=16)
a = ap.parse_args()
W = a.size
rig = synth.ring_rig(a.cams, W, W, kind="FTHETA")
colors, true = synth.render_rig(rig, W, W, device="cuda")
L = capi.load_cuda()
ctx = capi.Context(L, capi.rig_descs(rig))
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
ctx.level_begin(W, W, level=0, num_levels=5, full_width=W, full_height=W)
ctx.set_colors(colors)
rng = np.random.RandomState(0)
d = 0
start = (true[d] * (1 + 0.05 * rng.standard_normal(true[d].shape))).astype(np.float32).clip(1e-4, 2.0)
ctx.reproject(d)
def timed(fn, reps=3):
    out = []
    for _ in range(reps):
        ctx.set_disparity(d, start, np.zeros_like(start), np.zeros_like(start))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    ev, hits = ctx.get_counters()
    return min(out), ev, hits
var = ctx.get_variance(d); fov = ctx.get_fov_mask(d)
print("fov px %.2fM  var>=floor %.2fM  var>=1e-4 %.2fM" % (fov.sum()/1e6, ((var >= ctx.get_var_noise_floor()) & (fov > 0)).sum()/1e6, ((var >= 1e-4) & (fov > 0)).sum()/1e6))
for name, fn in (("proposals(2)", lambda: ctx.random_proposals(d, 2)), ("pingpong(1)", lambda: ctx.ping_pong(d, 1)), ("bilateral", lambda: ctx.bilateral(d)), ("median", lambda: ctx.median(d)), ("reproject", lambda: ctx.reproject(d))):
    ms, ev, hits = timed(fn)
    if name in ("bilateral", "median", "reproject"): ev = hits = 0
    print("%-13s %8.3f ms  %7.2f Mevals  %7.2f Mtriples  %6.2f Gtriples/s" % (name, ms, ev/1e6, hits/1e6, hits/ms/1e6 if ms else 0))
