"""GPU parity tests proper: the CUDA library vs the CPU oracle through the same C ABI, same seeded
inputs.  Bar (BASELINE.json north_star): integer disparity indices bit-exact; float depth within 1e-3
relative.  Most stages are in fact bit-exact; the tolerance classes are stated where they apply:
  * fp64 atan2 / sin / cos differ by <= 2 ulp between CUDA and glibc: a projected coordinate narrowed
    to fp32 can flip its last bit on ~1e-7 of pixels -> tables/costs allow a 1e-5 mismatching fraction;
  * expf differs in the last bits -> bilateral / temporal outputs compared at 2e-6 relative.
"""
import numpy as np
import pytest

from facebook360_dep_b200 import capi, synth
from tests.parity_util import both, make_pair, mismatch_fraction, same_float_bits, scene_inputs

pytestmark = pytest.mark.gpu

RIGS = [
    # cfg-1 of BASELINE.json: 4-camera synthetic rig (rectilinear, hfov 120 so neighbours overlap)
    ("rect4", dict(num_cams=4, width=96, height=80, kind="RECTILINEAR", hfov_deg=120.0)),
    ("ftheta8d", dict(num_cams=8, width=112, height=96, kind="FTHETA", distorted=True)),
    ("ftheta5", dict(num_cams=5, width=72, height=72, kind="FTHETA")),
    # layout of the reference's own 16-camera test rig: cameras on a sphere, rolled, per-camera intrinsics, 14:9 sensor
    ("sphere16", dict(num_cams=16, width=112, height=72, kind="SPHERE")),
]


def _begin(ctxs, colors, W, H, **kw):
    both(ctxs, "level_begin", W, H, **kw)
    both(ctxs, "set_colors", colors)


@pytest.mark.parametrize("name,cfg", RIGS)
def test_level_tables(cuda, oracle, name, cfg):
    rig, colors, _ = scene_inputs(**cfg)
    W, H = cfg["width"], cfg["height"]
    ctxs = make_pair(cuda, oracle, rig)
    _begin(ctxs, colors, W, H)
    S = len(colors)
    for d in range(S):
        g, o = both(ctxs, "get_fov_mask", d)
        assert np.array_equal(g, o)
    for s in range(S):
        g, o = both(ctxs, "get_variance", s)
        assert np.array_equal(g.view(np.uint32), o.view(np.uint32)), "variance must be bit-exact"
    for d in (0, S - 1):
        both(ctxs, "reproject", d)
        for s in range(S):
            gw, ow = both(ctxs, "get_proj_warp", s)
            assert mismatch_fraction(gw, ow) <= 1e-5
            assert np.array_equal(np.isnan(gw), np.isnan(ow))
            fin = ~np.isnan(ow)
            assert np.abs(gw[fin] - ow[fin]).max(initial=0) <= 1e-3
            gc, oc = both(ctxs, "get_proj_color", s)
            assert (gc != oc).mean() <= 1e-5, "projColor"
            gb, ob = both(ctxs, "get_proj_bias", s)
            assert (gb != ob).mean() <= 2e-5, "projBias"


@pytest.mark.parametrize("name,cfg", RIGS)
def test_eval_cost_bit_exact(cuda, oracle, name, cfg):
    rig, colors, true_disp = scene_inputs(**cfg)
    W, H = cfg["width"], cfg["height"]
    ctxs = make_pair(cuda, oracle, rig)
    _begin(ctxs, colors, W, H)
    rng = np.random.RandomState(1)
    for d in (0, len(colors) // 2):
        both(ctxs, "reproject", d)
        for disp in (np.full((H, W), 0.31, np.float32), true_disp[d],
                     rng.uniform(1e-4, 2.0, size=(H, W)).astype(np.float32)):
            (gc, gf), (oc, of) = both(ctxs, "eval_cost", d, disp)
            assert mismatch_fraction(gc, oc) <= 1e-5, name
            assert mismatch_fraction(gf, of) <= 1e-5, name
            assert ctxs[0].get_counters() == ctxs[1].get_counters()


@pytest.mark.parametrize("name,cfg", RIGS)
@pytest.mark.parametrize("num_depths", [32, 150])
def test_brute_force_indices_bit_exact(cuda, oracle, name, cfg, num_depths):
    rig, colors, _ = scene_inputs(**cfg)
    W, H = cfg["width"], cfg["height"]
    ctxs = make_pair(cuda, oracle, rig)
    _begin(ctxs, colors, W, H)
    total_bad = 0
    for d in range(len(colors)):
        both(ctxs, "reproject", d)
        gi, oi = both(ctxs, "brute_force", d, num_depths=num_depths)
        total_bad += int((gi != oi).sum())
        (gd, gc, gf), (od, oc, of) = both(ctxs, "get_disparity", d)
        ok = gi == oi
        assert same_float_bits(gd, od)[ok].all()
        assert same_float_bits(gc, oc)[ok].all()
        assert same_float_bits(gf, of)[ok].all()
        assert ctxs[0].get_counters() == ctxs[1].get_counters()
    assert total_bad == 0, "%d winner indices differ" % total_bad


def test_brute_force_cfg1_full_size(cuda, oracle):
    """BASELINE.json configs[0]: 4-camera synthetic rig, 512x512, 32 candidates, single level."""
    cfg = dict(num_cams=4, width=512, height=512, kind="RECTILINEAR", hfov_deg=120.0)
    rig, colors, true_disp = scene_inputs(**cfg)
    ctxs = make_pair(cuda, oracle, rig)
    _begin(ctxs, colors, 512, 512)
    for d in range(4):
        both(ctxs, "reproject", d)
        gi, oi = both(ctxs, "brute_force", d, num_depths=32)
        assert np.array_equal(gi, oi)
        gd, od = both(ctxs, "get_disparity", d, want_cost=False)
        assert same_float_bits(gd, od).all()
        # sanity only (not a parity check): with just two overlapping neighbours per camera and 32
        # candidates the sweep still lands within two candidate steps on a good share of covered pixels
        cov = oi >= 0
        assert (np.abs(od - true_disp[d])[cov] < 2 * 2.0 / 31).mean() > 0.3


def test_max_size_cfg4_one_destination(cuda, oracle):
    """BASELINE.json configs[3] (the largest): 24 cameras at 4096 x 4096.  One destination, a handful of candidates:
    the size is what is tested (16.8 M-pixel planes, 24-plane pair tables of 6.4 GB, 64-bit offsets)."""
    import torch
    W = H = 4096
    rig = synth.ring_rig(24, W, H, kind="FTHETA")
    colors, _ = synth.render_rig(rig, W, H, scene=synth.Scene(seed=9), device="cuda")
    torch.cuda.empty_cache()
    ctxs = make_pair(cuda, oracle, rig, dst_to_src=[5])
    _begin(ctxs, colors, W, H)
    both(ctxs, "reproject", 0)
    gi, oi = both(ctxs, "brute_force", 0, num_depths=6)
    assert (oi >= 0).mean() > 0.5
    assert np.array_equal(gi, oi)
    (gd, gc, gf), (od, oc, of) = both(ctxs, "get_disparity", 0)
    assert same_float_bits(gd, od).all()
    assert mismatch_fraction(gc, oc) <= 1e-5
    assert ctxs[0].get_counters() == ctxs[1].get_counters()


def test_coverage_check_matches_reference_abort(cuda, oracle):
    # Derp.cpp:334-339: CHECK(partialCoverage || useForegroundMasks) when no candidate is visible
    cfg = dict(num_cams=4, width=64, height=64, kind="RECTILINEAR", hfov_deg=60.0)
    rig, colors, _ = scene_inputs(**cfg)
    ctxs = make_pair(cuda, oracle, rig)
    _begin(ctxs, colors, 64, 64)
    both(ctxs, "reproject", 0)
    for c in ctxs:
        with pytest.raises(capi.DerpError) as e:
            c.brute_force(0, num_depths=16, partial_coverage=False)
        assert e.value.code == capi.ECOVERAGE


def _coarse_to_fine_start(ctxs, rig, colors, W, H, level=1, num_levels=3):
    """Brute force at half size with the ORACLE, upsample into both contexts at (W, H)."""
    cw, ch = W // 2, H // 2
    coarse_colors = [synth.downscale_area(c, 2) for c in colors]
    oc = ctxs[1]
    oc.level_begin(cw, ch, level=level + 1, num_levels=num_levels, full_width=W, full_height=H)
    oc.set_colors(coarse_colors)
    coarse = []
    for d in range(len(colors)):
        oc.reproject(d)
        oc.brute_force(d, num_depths=48, want_index=False)
        oc.mask_fov(d)
        coarse.append(oc.get_disparity(d, want_cost=False))
    both(ctxs, "level_begin", W, H, level=level, num_levels=num_levels, full_width=W, full_height=H)
    both(ctxs, "set_colors", colors)
    for d in range(len(colors)):
        both(ctxs, "upsample_from", d, coarse[d])
    return coarse


@pytest.mark.parametrize("name,cfg", RIGS[:2])
def test_fine_level_stages(cuda, oracle, name, cfg):
    rig, colors, _ = scene_inputs(**cfg)
    W, H = cfg["width"], cfg["height"]
    ctxs = make_pair(cuda, oracle, rig)
    _coarse_to_fine_start(ctxs, rig, colors, W, H)
    S = len(colors)
    for d in range(S):
        gd, od = both(ctxs, "get_disparity", d, want_cost=False)
        assert same_float_bits(gd, od).all(), "Lanczos upsample must be bit-exact (host-built tap tables)"
    for d in range(S):
        both(ctxs, "reproject", d)
        both(ctxs, "random_proposals", d, 2)
        (gd, gc, gf), (od, oc, of) = both(ctxs, "get_disparity", d)
        assert mismatch_fraction(gd, od) <= 2e-5, "random proposals disparity"
        assert mismatch_fraction(gc, oc) <= 2e-5 and mismatch_fraction(gf, of) <= 2e-5
        assert ctxs[0].get_counters() == ctxs[1].get_counters()
        ctxs[0].set_disparity(d, od, oc, of)  # re-synchronise before the next stage
        both(ctxs, "ping_pong", d, 2)
        (gd, gc, gf), (od, oc, of) = both(ctxs, "get_disparity", d)
        assert mismatch_fraction(gd, od) <= 2e-5, "ping-pong disparity"
        assert mismatch_fraction(gc, oc) <= 2e-5
        assert ctxs[0].get_counters() == ctxs[1].get_counters()
        ctxs[0].set_disparity(d, od, oc, of)
    # mismatch handling (all cameras' disparity)
    both(ctxs, "mismatches")
    for d in range(S):
        gd, od = both(ctxs, "get_disparity", d, want_cost=False)
        # same class as the other fp64-limited stages: a source coordinate whose last fp32 bit flips (atan2 ulp) moves the
        # bilinear read of that camera's disparity by ~1e-7 relative, which only matters within that distance of the
        # +-10 % agreement thresholds
        assert mismatch_fraction(gd, od) <= 2e-5
        gm, om = both(ctxs, "get_mismatch_mask", d)
        assert (gm != om).mean() <= 2e-5
        ctxs[0].set_disparity(d, od)
    for d in range(S):
        both(ctxs, "bilateral", d)
        gd, od = both(ctxs, "get_disparity", d, want_cost=False)
        fin = np.isfinite(od)
        assert np.array_equal(np.isfinite(gd), fin)
        assert (np.abs(gd - od)[fin] <= 2e-6 * np.abs(od)[fin] + 1e-12).all(), "bilateral"
        ctxs[0].set_disparity(d, od)
        both(ctxs, "median", d)
        gd, od = both(ctxs, "get_disparity", d, want_cost=False)
        assert same_float_bits(gd, od).all(), "median must be bit-exact"
        both(ctxs, "mask_fov", d)
        gd, od = both(ctxs, "get_disparity", d, want_cost=False)
        assert same_float_bits(gd, od).all()


def test_foreground_mask_paths(cuda, oracle):
    cfg = RIGS[1][1]
    rig, colors, true_disp = scene_inputs(**cfg)
    W, H = cfg["width"], cfg["height"]
    S = len(colors)
    rng = np.random.RandomState(5)
    yy, xx = np.mgrid[0:H, 0:W]
    masks = [(((xx - W / 2 - 6 * s) ** 2 + (yy - H / 2) ** 2) < (0.38 * W) ** 2).astype(np.uint8) for s in range(S)]
    bgs = [np.full((H, W), 0.05, np.float32) + rng.uniform(0, 0.01, (H, W)).astype(np.float32) for _ in range(S)]
    ctxs = make_pair(cuda, oracle, rig)
    both(ctxs, "level_begin", W, H, use_foreground_masks=True)
    both(ctxs, "set_colors", colors)
    both(ctxs, "set_foreground_masks", masks)
    both(ctxs, "set_background_disparity", bgs)
    for d in (0, 3):
        both(ctxs, "reproject", d)
        gi, oi = both(ctxs, "brute_force", d, num_depths=40, partial_coverage=False)
        assert np.array_equal(gi, oi)
        (gd, gc, gf), (od, oc, of) = both(ctxs, "get_disparity", d)
        assert same_float_bits(gd, od).all() and same_float_bits(gc, oc).all()
        both(ctxs, "random_proposals", d, 2)
        both(ctxs, "ping_pong", d, 1)
        gd, od = both(ctxs, "get_disparity", d, want_cost=False)
        assert mismatch_fraction(gd, od) <= 1e-4
        ctxs[0].set_disparity(d, od)
        both(ctxs, "bilateral", d)
        both(ctxs, "median", d)
        both(ctxs, "mask_fov", d)
        gd, od = both(ctxs, "get_disparity", d, want_cost=False)
        fin = np.isfinite(od)
        assert np.array_equal(np.isfinite(gd), fin)
        assert (np.abs(gd - od)[fin] <= 1e-5 * np.abs(od)[fin] + 1e-12).mean() > 0.999
    # masked upsampling (nearest + spiral fill + background)
    cw, ch = W // 2, H // 2
    coarse = rng.uniform(0.06, 1.5, (ch, cw)).astype(np.float32)
    coarse[rng.uniform(size=coarse.shape) < 0.1] = np.nan
    cmask = masks[2][::2, ::2].copy()
    both(ctxs, "upsample_from", 2, coarse, cmask, masks[2])
    gd, od = both(ctxs, "get_disparity", 2, want_cost=False)
    assert same_float_bits(gd, od).all(), "masked upsample must be bit-exact"


def test_process_level_coarse_to_fine(cuda, oracle):
    """3-level coarse-to-fine run, each library on its own (no re-synchronisation): end-to-end class (iii)
    of SURVEY.md §8(c): >= 99.9 % of pixels within 1e-3 relative."""
    cfg = dict(num_cams=6, width=128, height=128, kind="FTHETA", distorted=True)
    rig, colors, _ = scene_inputs(**cfg)
    W = H = 128
    pyr = [colors, [synth.downscale_area(c, 2) for c in colors], [synth.downscale_area(c, 4) for c in colors]]
    ctxs = make_pair(cuda, oracle, rig)
    prev = None
    for level in (2, 1, 0):
        w = W >> level
        both(ctxs, "level_begin", w, w, level=level, num_levels=3, full_width=W, full_height=H)
        both(ctxs, "set_colors", pyr[level])
        if prev is not None:
            for c, p in zip(ctxs, prev):
                for d in range(6):
                    c.upsample_from(d, p[d])
        both(ctxs, "process_level", num_depths=64)
        prev = [[c.get_disparity(d, want_cost=False) for d in range(6)] for c in ctxs]
        good = tot = 0
        for d in range(6):
            g, o = prev[0][d], prev[1][d]
            assert np.array_equal(np.isnan(g), np.isnan(o))
            fin = ~np.isnan(o)
            good += (np.abs(g - o)[fin] <= 1e-3 * np.abs(o)[fin]).sum()
            tot += fin.sum()
        assert good / tot >= 0.999, (level, good / tot)


def test_temporal_and_joint_bilateral(cuda, oracle):
    rng = np.random.RandomState(11)
    H, W, T = 60, 68, 5
    base = rng.randint(0, 65536, (H, W, 3))
    guides = [np.clip(base + rng.randint(-300, 300, (H, W, 3)), 0, 65535).astype(np.uint16) for _ in range(T)]
    disps = [rng.uniform(1e-3, 2, (H, W)).astype(np.float32) for _ in range(T)]
    masks = [(rng.uniform(size=(H, W)) > 0.15).astype(np.uint8) for _ in range(T)]
    for off, r in ((2, 1), (0, 2), (4, 0)):
        g = cuda.temporal_filter(guides, disps, masks, off, 0.01, r, 0.5, 1.0, 0.5)
        o = oracle.temporal_filter(guides, disps, masks, off, 0.01, r, 0.5, 1.0, 0.5)
        fin = np.isfinite(o)
        assert np.array_equal(np.isfinite(g), fin)
        assert (np.abs(g - o)[fin] <= 2e-6 * np.abs(o)[fin]).all()
    img = disps[0]
    guide = guides[0].astype(np.float32) * (np.float32(1.0) / np.float32(65535.0))  # loadImage<Vec3f>
    # radius 3 and 17 (= scale^2 + 1 of a 4x UpsampleDisparity) run from the shared-memory tile, 20 from global memory
    for radius in (3, 17, 20):
        g = cuda.joint_bilateral_f32(img, guide, masks[0], radius, 0.05, 0.5, 0.5, 1.0)
        o = oracle.joint_bilateral_f32(img, guide, masks[0], radius, 0.05, 0.5, 0.5, 1.0)
        assert (np.abs(g - o) <= 2e-6 * np.abs(o)).all(), radius


def test_constant_division_is_exact(cuda):
    """The filters divide per tap by loop-invariant constants with a 3-instruction sequence that the library
    validates exhaustively against IEEE division per constant; the constants the path uses must all qualify
    (otherwise the kernels silently take the slow division and the timing claims would not hold)."""
    import ctypes as C
    f = cuda.lib.derp_test_div_const
    f.restype, f.argtypes = C.c_int, [C.c_int, C.c_float]
    s1, s2 = np.float32(0.005), np.float32(0.01)
    for c in (3.0, 65535.0, float(np.float32(2.0) * (s1 * s1)), float(s2 * s2), float(np.float32(2.0) * np.float32(0.05) ** 2)):
        assert f(0, c) == 1, c
    assert f(0, 0.0) == 0  # degenerate divisors never take the fast path


def test_standalone_upsample(cuda, oracle):
    rig = synth.ring_rig(4, 96, 64, kind="FTHETA")
    d = capi.camera_desc_from_json(rig["cameras"][1])
    rng = np.random.RandomState(2)
    coarse = rng.uniform(1e-4, 2, (32, 48)).astype(np.float32)
    coarse[3:6, 7:9] = np.nan
    for (w, h) in ((96, 64), (100, 70), (48, 32)):
        g = cuda.upsample_disparity(d, coarse, w, h)
        o = oracle.upsample_disparity(d, coarse, w, h)
        assert same_float_bits(g, o).all()
    cm = (rng.uniform(size=(32, 48)) > 0.3).astype(np.uint8)
    fm = (rng.uniform(size=(64, 96)) > 0.2).astype(np.uint8)
    bg = rng.uniform(0.01, 0.02, (64, 96)).astype(np.float32)
    g = cuda.upsample_disparity(d, coarse, 96, 64, bg, cm, fm, True)
    o = oracle.upsample_disparity(d, coarse, 96, 64, bg, cm, fm, True)
    assert same_float_bits(g, o).all()


def test_dst_subset_and_edge_sizes(cuda, oracle):
    # --cameras subset (filterDestinations) and tiny / ragged sizes
    cfg = dict(num_cams=5, width=72, height=72, kind="FTHETA")
    rig, colors, _ = scene_inputs(**cfg)
    ctxs = make_pair(cuda, oracle, rig, dst_to_src=[3, 0])
    _begin(ctxs, colors, 72, 72)
    for d in (0, 1):
        both(ctxs, "reproject", d)
        gi, oi = both(ctxs, "brute_force", d, num_depths=20)
        assert np.array_equal(gi, oi)
    rig2 = synth.ring_rig(3, 33, 3, kind="FTHETA")
    col2, _ = synth.render_rig(rig2, 33, 3)
    ctxs = make_pair(cuda, oracle, rig2)
    _begin(ctxs, col2, 33, 3)
    both(ctxs, "reproject", 1)
    gi, oi = both(ctxs, "brute_force", 1, num_depths=7)
    assert np.array_equal(gi, oi)
    with pytest.raises(capi.DerpError):
        ctxs[0].level_begin(2, 2)
    with pytest.raises(capi.DerpError):
        ctxs[0].brute_force(0)  # stage before reproject of that dst -> DERP_ESTATE


def test_many_overlapping_sources(cuda, oracle):
    """12-camera planar array: up to 11 sources contribute to one cost (more than the 8 shared-memory selection
    slots -> local-memory overflow path, introselect on 9..11 elements)."""
    W, H = 80, 64
    rig = synth.wall_rig(12, W, H)
    colors, _ = synth.render_rig(rig, W, H, scene=synth.Scene(seed=7))
    ctxs = make_pair(cuda, oracle, rig)
    _begin(ctxs, colors, W, H)
    for d in (0, 5, 11):
        both(ctxs, "reproject", d)
        gi, oi = both(ctxs, "brute_force", d, num_depths=40, min_depth_m=1.0)
        assert np.array_equal(gi, oi)
        (gd, gc, gf), (od, oc, of) = both(ctxs, "get_disparity", d)
        assert same_float_bits(gc, oc).all()
        ev, hits = ctxs[0].get_counters()
        assert (ev, hits) == ctxs[1].get_counters()
    assert hits / ev > 8.5, hits / ev  # the overflow path really ran


def test_selection_ties_and_mixed_counts(cuda, oracle):
    """Robust-mean selection by table (derp_select.cuh::robustSumTable): duplicated source cameras produce EQUAL
    (biased SSD, unbiased SSD) pairs, i.e. ties among the first keys -> the table path must hand over to the general
    libstdc++-order algorithm; a ring with few cameras next to a many-camera wall gives warps whose lanes hold
    different source counts (K = 6 and K = 8 instances, n <= 3 shortcut, > 8 general path)."""
    import copy
    W, H = 96, 64
    rig = synth.ring_rig(6, W, H, kind="FTHETA")
    colors, _ = synth.render_rig(rig, W, H, scene=synth.Scene(seed=21), noise=False)
    for k in (1, 2, 4):  # exact clones: same pose, same image
        cam = copy.deepcopy(rig["cameras"][k])
        cam["id"] = "clone%d" % k
        rig["cameras"].append(cam)
        colors.append(colors[k].copy())
    ctxs = make_pair(cuda, oracle, rig)
    _begin(ctxs, colors, W, H)
    for d in (0, 3, 7):
        both(ctxs, "reproject", d)
        gi, oi = both(ctxs, "brute_force", d, num_depths=48)
        assert np.array_equal(gi, oi)
        (gd, gc, gf), (od, oc, of) = both(ctxs, "get_disparity", d)
        assert same_float_bits(gc, oc).all() and same_float_bits(gd, od).all()
        assert ctxs[0].get_counters() == ctxs[1].get_counters()


def test_camera_sharded_mismatches(cuda, oracle):
    """Destination cameras of one frame dealt to two contexts (what DerpCLI --gpus does with a single frame, here both
    on cuda:0): estimate per shard, exchange the disparity planes device-to-device by address, mismatch handling per
    shard.  Must equal the all-camera context of the same library bit for bit, and the oracle run of the sharded
    protocol within the float bar."""
    from tests.test_shard import _MM, _mm_inputs, _mm_run
    rig, colors, init = _mm_inputs()
    S = _MM["S"]
    descs = capi.rig_descs(rig)
    shards = [[0, 2], [1, 3]]

    def run_sharded(lib):
        ctxs = [capi.Context(lib, descs, own) for own in shards]
        res = {}
        W, H = _MM["W"], _MM["H"]
        kw = dict(random_proposals=1, ping_pong_iterations=1, mismatches_start_level=0)
        for ctx, own in zip(ctxs, shards):
            ctx.level_begin(W, H, level=0, num_levels=2, full_width=W, full_height=H, var_noise_floor=0.0,
                            var_high_thresh=1e9)
            ctx.set_colors(colors)
            for i, cam in enumerate(own):
                ctx.set_disparity(i, init[cam])
            ctx.level_estimate(**kw)
        planes = [None] * S
        for ctx, own in zip(ctxs, shards):
            for i, cam in enumerate(own):
                planes[cam] = ctx.disparity_ptr(i)
        for ctx in ctxs:  # every shard copies BEFORE any shard updates (the barrier of the multi-GPU protocol)
            ctx.gather_disparities(planes)
        for ctx, own in zip(ctxs, shards):
            ctx.mismatches_gathered()
            ctx.level_filter(**kw)
            for i, cam in enumerate(own):
                res[cam] = (ctx.get_disparity(i, want_cost=False), ctx.get_mismatch_mask(i))
        return res

    got = run_sharded(cuda)
    whole = capi.Context(cuda, descs)
    ref = _mm_run(whole, list(range(S)), colors, init, 1, lambda c: c.mismatches())
    for cam in range(S):
        assert np.array_equal(got[cam][1], ref[cam][1])
        assert same_float_bits(got[cam][0], ref[cam][0]).all()
    orc = run_sharded(oracle)
    for cam in range(S):
        assert (got[cam][1] != orc[cam][1]).mean() <= 1e-3
        fin = np.isfinite(orc[cam][0])
        assert np.array_equal(np.isfinite(got[cam][0]), fin)
        close = np.abs(got[cam][0] - orc[cam][0])[fin] <= 1e-3 * np.abs(orc[cam][0])[fin]
        assert close.mean() >= 1 - 1e-3


def test_downscale_area_pyramid_widths(cuda, oracle):
    """§8(f) pyramid pre-resize: cv::resize INTER_AREA exactly as scripts/render/resize.py:79 applies it (every level from
    the full-size image; widths scripts/render/config.py:46), bit for bit vs the oracle restatement, which is pinned to
    cv2 4.13 (tests/test_oracle_cv.py).  Integer ratios (2 x 2 integer mean, other areas) and general ratios."""
    rng = np.random.RandomState(4)
    full = rng.randint(0, 65536, (540, 840, 3)).astype(np.uint16)  # the reference rig's 3360 x 2160 at 1/4
    for width in (512, 256, 200, 128, 100, 80, 60, 50):  # 2048 .. 50 of the reference at 1/4 + the small levels as they are
        height = round(540 / 840 * width)
        height += height % 2
        g = cuda.downscale_area(full, width, height)
        o = oracle.downscale_area(full, width, height)
        assert np.array_equal(g, o), width
    sq = rng.randint(0, 65536, (512, 512, 3)).astype(np.uint16)
    for w in (256, 128, 64, 200, 100):
        assert np.array_equal(cuda.downscale_area(sq, w, w), oracle.downscale_area(sq, w, w)), w
    assert np.array_equal(cuda.downscale_area(sq, 512, 512), sq)
    with pytest.raises(capi.DerpError):
        cuda.downscale_area(sq, 600, 600)  # enlarging is not INTER_AREA's shrinking path


def test_level_handoff_in_memory(cuda, oracle):
    """§8(f) in-memory level hand-off: derp_level_keep + derp_upsample_from_kept must give the bytes of the file round
    trip they replace (derp_get_disparity -> derp_upsample_from), with and without foreground masks."""
    cfg = dict(num_cams=5, width=96, height=80, kind="FTHETA")
    rig, colors, true_disp = scene_inputs(**cfg)
    W, H = 96, 80
    coarse_colors = [cuda.downscale_area(c, W // 2, H // 2) for c in colors]
    for use_fg in (False, True):
        rng = np.random.RandomState(8)
        ctx = capi.Context(cuda, capi.rig_descs(rig))
        ctx.level_begin(W // 2, H // 2, level=1, num_levels=2, full_width=W, full_height=H, use_foreground_masks=use_fg)
        ctx.set_colors(coarse_colors)
        coarse = [rng.uniform(0.01, 1.5, (H // 2, W // 2)).astype(np.float32) for _ in range(5)]
        for d in range(5):
            coarse[d][rng.uniform(size=coarse[d].shape) < 0.05] = np.nan
            ctx.set_disparity(d, coarse[d])
        ctx.level_keep()
        masks = [(rng.uniform(size=(H, W)) > 0.2).astype(np.uint8) for _ in range(5)]
        cmasks = [m[::2, ::2].copy() for m in masks]
        ctx.level_begin(W, H, level=0, num_levels=2, full_width=W, full_height=H, use_foreground_masks=use_fg)
        ctx.set_colors(colors)
        if use_fg:
            ctx.set_foreground_masks(masks)
            ctx.set_background_disparity([np.full((H, W), 0.02, np.float32)] * 5)
        got = []
        for d in range(5):
            ctx.upsample_from_kept(d, cmasks[d] if use_fg else None, masks[d] if use_fg else None)
            got.append(ctx.get_disparity(d, want_cost=False))
        for d in range(5):
            ctx.upsample_from(d, coarse[d], cmasks[d] if use_fg else None, masks[d] if use_fg else None)
            assert same_float_bits(got[d], ctx.get_disparity(d, want_cost=False)).all(), (use_fg, d)
        ctx.close()


def test_foreground_mask(cuda, oracle):
    """§8(f) GenerateForegroundMasks body (BackgroundSubtractionUtil.h:20-59): CUDA vs the oracle restatement (itself
    pinned to the cv2 call sequence, tests/test_oracle_cv.py), bit for bit."""
    rng = np.random.RandomState(6)
    H, W = 150, 211
    bg = np.clip(rng.normal(30000, 9000, (H, W, 3)), 0, 65535).astype(np.uint16)
    fr = np.clip(bg.astype(np.int64) + rng.randint(-1500, 1500, bg.shape), 0, 65535).astype(np.uint16)
    fr[40:100, 60:150] = rng.randint(0, 65536, (60, 90, 3)).astype(np.uint16)
    for blur, close, thr in ((1, 4, 0.04), (0, 3, 0.04), (1, 0, 0.02), (1, 7, 0.06)):
        g = cuda.foreground_mask(bg, fr, blur, thr, close)
        o = oracle.foreground_mask(bg, fr, blur, thr, close)
        assert np.array_equal(g, o), (blur, close, thr)
        assert 0.05 < g.mean() < 0.95
    with pytest.raises(capi.DerpError):
        cuda.foreground_mask(bg, fr, 2, 0.04, 4)


def test_24_camera_rig_all_stages(cuda, oracle):
    """BASELINE.json configs[3]'s camera count (24) at a small size, every stage of a fine level incl. mismatch handling:
    the shared-memory footprints (S - 1 selection slots per thread, S camera structs) scale with the rig."""
    W, H, S = 72, 56, 24
    rig = synth.ring_rig(S, W, H, kind="FTHETA")
    colors, true_disp = synth.render_rig(rig, W, H, scene=synth.Scene(seed=12))
    ctxs = make_pair(cuda, oracle, rig)
    both(ctxs, "level_begin", W, H, level=0, num_levels=2, full_width=W, full_height=H)
    both(ctxs, "set_colors", colors)
    rng = np.random.RandomState(3)
    for d in range(S):
        start = np.clip(true_disp[d] * rng.uniform(0.85, 1.2, (H, W)).astype(np.float32), 1e-4, 2.0).astype(np.float32)
        both(ctxs, "set_disparity", d, start, np.zeros_like(start), np.zeros_like(start))
    both(ctxs, "process_level", num_depths=32, mismatches_start_level=0)
    good = tot = 0
    for d in range(S):
        g, o = both(ctxs, "get_disparity", d, want_cost=False)
        assert np.array_equal(np.isnan(g), np.isnan(o))
        fin = ~np.isnan(o)
        good += int((np.abs(g - o)[fin] <= 1e-3 * np.abs(o)[fin]).sum())
        tot += int(fin.sum())
    assert good / tot >= 0.999
    both(ctxs, "reproject", 7)
    gi, oi = both(ctxs, "brute_force", 7, num_depths=24)
    assert np.array_equal(gi, oi)
