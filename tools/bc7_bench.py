#!/usr/bin/env python3
"""derp_bc7_compress (K18) on a 2048^2 (and 4096^2) RGBA8 surface with device-resident input and output: ms per image (CUDA
events), blocks/s, algorithmic bytes (64 B in + 16 B out per block) against the measured HBM peak — the kernel is bound by
instruction issue, the byte figure only shows how far from memory-bound it is — and the reference's own encoder
(oracle/_ref: kernel.ispc built by the vendored ispc, one host thread like ConvertToBinary's per-camera task) on the same
surface, with the fraction of byte-identical blocks."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from facebook360_dep_b200 import capi
from tests import oracle_libs
from tests.test_bc7 import surface

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="2048,4096")
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
L = capi.load_cuda()
ref = oracle_libs.load_ref()
peak = 6573.5
try:
    peak = json.load(open(os.path.join(capi.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
for size in [int(s) for s in a.sizes.split(",")]:
    w = h = size
    rgba = surface(9, w, h, "smooth")
    src = torch.from_numpy(rgba).cuda()
    out = torch.empty(w * h, dtype=torch.uint8, device="cuda")

    def run():
        L.check(L.lib.derp_bc7_compress(0, src.data_ptr(), w, h, out.data_ptr()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    blocks = (w // 4) * (h // 4)
    alg = 80 * blocks
    line = "%d^2: %.3f ms per image (%d blocks, %.1f M blocks/s), algorithmic %.1f MB -> %.0f GB/s = %.3f of the measured HBM peak (%.0f GB/s)" % (
        size, ms, blocks, blocks / ms / 1e3, alg / 1e6, alg / ms / 1e6, alg / ms / 1e6 / peak, peak)
    if ref is not None and size <= 2048:
        want = np.empty(w * h, np.uint8)
        t0 = time.time()
        ref.check(ref.lib.derp_bc7_compress(0, rgba.ctypes.data, w, h, want.ctypes.data))
        cpu = time.time() - t0
        same = (want.reshape(-1, 16) == out.cpu().numpy().reshape(-1, 16)).all(1).mean()
        line += "; reference encoder on one host thread: %.2f s (%.0fx); %.2f %% of the blocks byte-identical" % (cpu, cpu * 1e3 / ms, 100 * same)
    print(line)
