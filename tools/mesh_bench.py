#!/usr/bin/env python3
"""derp_camera_mesh (K17) at 2048^2 / 4096^2 with device-resident input and outputs: ms per mesh (CUDA events), algorithmic
bytes (4 B disparity in, 12 B per vertex and 12 B per face out) against the measured HBM peak, and the reference's own
MeshUtil.h (oracle/_ref, one host thread like ConvertToBinary's per-camera task) on the same map."""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from facebook360_dep_b200 import capi
from tests import oracle_libs
from tests.test_mesh import disparity_case

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="2048,4096")
ap.add_argument("--reps", type=int, default=10)
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
    d = disparity_case(np.random.RandomState(3), w, h)
    dd = torch.from_numpy(d).cuda()
    vtx = torch.empty((w * h, 3), dtype=torch.float32, device="cuda")
    idx = torch.empty((2 * w * h, 3), dtype=torch.int32, device="cuda")
    nv, nf = C.c_uint64(), C.c_uint64()

    def run():
        L.check(L.lib.derp_camera_mesh(0, dd.data_ptr(), w, h, 1.0, float(w), float(h), 651.9, 0.95, None, 0, 0,
                                       vtx.data_ptr(), idx.data_ptr(), C.byref(nv), C.byref(nf)))
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
    alg = 4 * w * h + 12 * nv.value + 12 * nf.value
    line = "%d^2: %.3f ms per mesh (%d vertexes, %d faces), algorithmic %.1f MB -> %.0f GB/s = %.2f of the measured HBM peak (%.0f GB/s)" % (
        size, ms, nv.value, nf.value, alg / 1e6, alg / ms / 1e6, alg / ms / 1e6 / peak, peak)
    if ref is not None:
        t0 = time.time()
        rv, ri = ref.camera_mesh(d, (float(w), float(h)), 651.9)
        cpu = time.time() - t0
        same = len(rv) == nv.value and len(ri) == nf.value and np.array_equal(ri, idx[:nf.value].cpu().numpy().view(np.uint32))
        line += "; reference MeshUtil.h on one host thread: %.2f s (%.0fx), faces identical: %s" % (cpu, cpu * 1e3 / ms, same)
    print(line)
