#!/usr/bin/env python3
"""Kernel-iteration helper: times the sweep (derp_brute_force) for a few destinations of the bf128_l0
rig at a chosen size and prints a hash of the winner-index maps so kernel variants can be compared."""
import argparse, hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from facebook360_dep_b200 import capi, synth

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--cams", type=int, default=16)
ap.add_argument("--depths", type=int, default=128)
ap.add_argument("--dsts", type=int, default=2)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--kind", default="FTHETA")
a = ap.parse_args()
W = H = a.size
rig = synth.ring_rig(a.cams, W, H, kind=a.kind, hfov_deg=120.0 if a.kind == "RECTILINEAR" else None)
colors, _ = synth.render_rig(rig, W, H, device="cuda")
L = capi.load_cuda()
ctx = capi.Context(L, capi.rig_descs(rig))
ctx.level_begin(W, H)
ctx.set_colors(colors)
h = hashlib.sha1()
for d in range(a.dsts):
    ctx.reproject(d)
    idx = ctx.brute_force(d, num_depths=a.depths)
    h.update(idx.tobytes())
    e, hits = ctx.get_counters()
    ctx.profile(True)
    for _ in range(a.reps):
        ctx.brute_force(d, num_depths=a.depths, want_index=False)
    ms, n = ctx.get_profile()
    ctx.profile(False)
    try:
        print("dst %d: sweep stats (refined, seeds) %s" % (d, ctx.sweep_stats()))
    except Exception as ex:
        print("no sweep stats", ex)
    print("dst %d: sweep %.2f ms/launch  %.2f Gpix·cand/s  %.2f Gtriples/s  vbar %.2f" % (
        d, ms / n, e / (ms / n) / 1e6, hits / (ms / n) / 1e6, hits / e))
print("idx sha1", h.hexdigest()[:16])
