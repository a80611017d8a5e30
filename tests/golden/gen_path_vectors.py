#!/usr/bin/env python3
"""Generates tests/golden/path_vectors.npz: a frozen known-answer case of the depth path itself.

The reference ships no golden vectors for computeCost / brute force / the fine-level stages (parity
unpinned by the reference), so the CPU oracle is the pin — and this fixture pins the ORACLE in time: inputs (rig JSON,
u16 frames of a 4-camera 48 x 40 rig, two pyramid levels) and the oracle's outputs at the commit that generated it.
tests/test_golden_path.py checks (CPU) that the oracle still reproduces them bit for bit and (GPU) that the CUDA
library reproduces the integer winner indices exactly and the float maps within the stated bars, without the oracle
in the loop.  Regenerate only when the oracle is deliberately changed:  python tests/golden/gen_path_vectors.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from facebook360_dep_b200 import capi, synth  # noqa: E402
from tests import oracle_libs  # noqa: E402

W, H, S, D = 48, 40, 4, 24


def run(lib, rig, fine, coarse):
    """Brute force on the single-level problem + a 2-level coarse-to-fine run; returns dict of arrays."""
    out = {}
    ctx = capi.Context(lib, capi.rig_descs(rig))
    ctx.level_begin(W, H)
    ctx.set_colors(fine)
    for d in range(S):
        ctx.reproject(d)
        out["bf_idx%d" % d] = ctx.brute_force(d, num_depths=D)
        disp, cost, conf = ctx.get_disparity(d)
        out["bf_disp%d" % d], out["bf_cost%d" % d], out["bf_conf%d" % d] = disp, cost, conf
    cw, ch = W // 2, H // 2
    ctx.level_begin(cw, ch, level=1, num_levels=2, full_width=W, full_height=H)
    ctx.set_colors(coarse)
    ctx.process_level(num_depths=D)
    low = [ctx.get_disparity(d, want_cost=False) for d in range(S)]
    ctx.level_begin(W, H, level=0, num_levels=2, full_width=W, full_height=H)
    ctx.set_colors(fine)
    for d in range(S):
        ctx.upsample_from(d, low[d])
    ctx.process_level(num_depths=D, mismatches_start_level=0)
    for d in range(S):
        out["c2f_low%d" % d] = low[d]
        out["c2f_disp%d" % d] = ctx.get_disparity(d, want_cost=False)
    return out


def main():
    rig = synth.ring_rig(S, W, H, kind="RECTILINEAR", hfov_deg=120.0)
    fine, _ = synth.render_rig(rig, W, H, scene=synth.Scene(seed=21))
    coarse = [synth.downscale_area(c, 2) for c in fine]
    oracle = oracle_libs.load_oracle()
    out = run(oracle, rig, fine, coarse)
    out["rig_json"] = np.frombuffer(json.dumps(rig).encode(), dtype=np.uint8)
    for s in range(S):
        out["fine%d" % s] = fine[s]
        out["coarse%d" % s] = coarse[s]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "path_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "covered px:",
          sum(int((out["bf_idx%d" % d] >= 0).sum()) for d in range(S)))


if __name__ == "__main__":
    main()
