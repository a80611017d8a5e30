"""The filtered brute-force sweep (derp_refine.cuh): lower bound of every (pixel, candidate), exact cost only where the
bound does not exclude the candidate.  Two things are checked on the GPU:
  1. the BOUND: for every (pixel, candidate) of a destination the bound is <= the exact cost (derp_debug_lower_bound
     evaluates both; zero violations is the correctness condition of the filter), and it is tight enough to be useful;
  2. the RESULT: index / disparity / cost / confidence / work counters of the filtered sweep equal the plain sweep's bit
     for bit (and, through tests/test_gpu_parity.py which now runs filtered by default, the oracle's)."""
import ctypes as C

import numpy as np
import pytest

from facebook360_dep_b200 import capi, synth
from tests.parity_util import same_float_bits, scene_inputs

pytestmark = pytest.mark.gpu

CASES = [
    ("rect4", dict(num_cams=4, width=96, height=80, kind="RECTILINEAR", hfov_deg=120.0), 150),
    ("ftheta8d", dict(num_cams=8, width=112, height=96, kind="FTHETA", distorted=True), 150),
    ("sphere16", dict(num_cams=16, width=112, height=72, kind="SPHERE"), 64),
    ("ring16_512", dict(num_cams=16, width=512, height=512, kind="FTHETA"), 128),
]


def lower_bound_stats(cuda, ctx, dst, D):
    f = cuda.lib.derp_debug_lower_bound
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_uint64)]
    st = (C.c_uint64 * 5)()
    cuda.check(f(ctx.h, dst, D, 0.5, 1e4, st))
    return [int(v) for v in st]


@pytest.mark.parametrize("name,cfg,D", CASES)
def test_lower_bound_never_exceeds_exact_cost(cuda, name, cfg, D):
    rig, colors, _ = scene_inputs(**cfg)
    W, H = cfg["width"], cfg["height"]
    ctx = capi.Context(cuda, capi.rig_descs(rig))
    ctx.level_begin(W, H)
    ctx.set_colors(colors)
    for d in (0, len(colors) // 2):
        ctx.reproject(d)
        n, bad, unknown, tight, keep = lower_bound_stats(cuda, ctx, d, D)
        assert n > 0
        assert bad == 0, "%d of %d lower bounds exceed the exact cost" % (bad, n)
        assert unknown <= 0.05 * n, (unknown, n)
        print("%s dst %d: %d evaluations, %.1f %% bounds within 5 %%, ideal survivors %.2f %%, unknown %.3f %%" % (
            name, d, n, 100.0 * tight / n, 100.0 * keep / n, 100.0 * unknown / n))
    ctx.close()


@pytest.mark.parametrize("name,cfg,D", CASES)
def test_filtered_sweep_equals_plain_sweep(cuda, name, cfg, D):
    rig, colors, _ = scene_inputs(**cfg)
    W, H = cfg["width"], cfg["height"]
    ctx = capi.Context(cuda, capi.rig_descs(rig))
    ctx.level_begin(W, H)
    ctx.set_colors(colors)
    for d in range(0, len(colors), max(1, len(colors) // 3)):
        ctx.reproject(d)
        out = []
        for mode in (1, 2):
            ctx.set_sweep_mode(mode)
            idx = ctx.brute_force(d, num_depths=D)
            out.append((idx,) + ctx.get_disparity(d) + (ctx.get_counters(), ctx.sweep_stats()))
        (i1, d1, c1, f1, n1, s1), (i2, d2, c2, f2, n2, s2) = out
        assert np.array_equal(i1, i2)
        assert same_float_bits(d1, d2).all() and same_float_bits(c1, c2).all() and same_float_bits(f1, f2).all()
        assert n1 == n2, "work counters must count the algorithmic work in both modes"
        assert s1 == (0, 0) and s2[1] > 0
        print("%s dst %d: refined %.2f %% of %d evaluations" % (name, d, 100.0 * s2[0] / max(1, n2[0]), n2[0]))
    ctx.close()


def test_filtered_sweep_with_foreground_masks(cuda):
    cfg = dict(num_cams=8, width=112, height=96, kind="FTHETA", distorted=True)
    rig, colors, _ = scene_inputs(**cfg)
    W, H, S = 112, 96, 8
    rng = np.random.RandomState(5)
    yy, xx = np.mgrid[0:H, 0:W]
    masks = [(((xx - W / 2 - 6 * s) ** 2 + (yy - H / 2) ** 2) < (0.38 * W) ** 2).astype(np.uint8) for s in range(S)]
    bgs = [np.full((H, W), 0.05, np.float32) + rng.uniform(0, 0.4, (H, W)).astype(np.float32) for _ in range(S)]
    ctx = capi.Context(cuda, capi.rig_descs(rig))
    ctx.level_begin(W, H, use_foreground_masks=True)
    ctx.set_colors(colors)
    ctx.set_foreground_masks(masks)
    ctx.set_background_disparity(bgs)
    for d in (0, 5):
        ctx.reproject(d)
        res = []
        for mode in (1, 2):
            ctx.set_sweep_mode(mode)
            res.append((ctx.brute_force(d, num_depths=40, partial_coverage=False),) + ctx.get_disparity(d) + (ctx.get_counters(),))
        assert np.array_equal(res[0][0], res[1][0])
        for a, b in zip(res[0][1:4], res[1][1:4]):
            assert same_float_bits(a, b).all()
        assert res[0][4] == res[1][4]
    ctx.close()
