"""The oracle pinned to REFERENCE-AUTHORED code: oracle/libderp_oracle.so (our CPU restatement) against
oracle/_ref/libderp_ref.so = the reference's own Derp.cpp / DerpUtil.cpp / UpsampleDisparityLib.cpp / Camera.cpp /
CvUtil.cpp / ImageUtil.cpp (+ PyramidLevel.h, TemporalBilateralFilter.h, CvUtil.h, Camera.h) compiled where they lie under
/root/reference against the stand-in headers of oracle/refshim (recipe: oracle/Makefile, glue: oracle/ref_bridge.cpp).

Both run on this CPU with the same libm / libstdc++, so EVERY comparison here is bit-for-bit (NaN == NaN):
SURVEY.md §8 rows a1-a19, a20 (PFM bytes), a22, a23's filter.  The candidate count of the reference is the compile-time
constant 150 (Derp.h:33), so brute force is compared at 150 candidates.

Skipped when oracle/_ref is not built (it needs /root/reference, which exists in the build container only; the built
library travels to the GPU box, where tests/test_gpu_parity.py compares the CUDA library with it directly)."""
import os

import numpy as np
import pytest

from facebook360_dep_b200 import capi, synth
from tests.parity_util import both, same_float_bits, scene_inputs

RIGS = [
    ("rect4", dict(num_cams=4, width=64, height=48, kind="RECTILINEAR", hfov_deg=120.0)),
    ("ftheta8d", dict(num_cams=8, width=72, height=60, kind="FTHETA", distorted=True)),
    ("sphere16", dict(num_cams=16, width=56, height=36, kind="SPHERE")),
]


def pair(oracle, ref, rig, dst_to_src=None):
    descs = capi.rig_descs(rig)
    return capi.Context(oracle, descs, dst_to_src), capi.Context(ref, descs, dst_to_src)


def eq_bits(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype == np.float32:
        return bool(same_float_bits(a, b).all())
    return bool(np.array_equal(a, b))


@pytest.mark.parametrize("name,cfg", RIGS)
def test_tables_cost_and_brute_force(oracle, ref, name, cfg):
    """a1-a8 (camera, FOV masks, warp tables, bicubic reprojection, bias, variance), a10/a11 (computeCost / computeSSD on
    arbitrary hypothesis maps) and a12 (brute force + WTA + border) against the reference's own functions."""
    rig, colors, true_disp = scene_inputs(**cfg)
    W, H, S = cfg["width"], cfg["height"], len(colors)
    ctxs = pair(oracle, ref, rig)
    both(ctxs, "level_begin", W, H)
    both(ctxs, "set_colors", colors)
    assert ctxs[0].get_var_noise_floor() == ctxs[1].get_var_noise_floor()
    for d in range(S):
        o, r = both(ctxs, "get_fov_mask", d)
        assert eq_bits(o, r), "generateFovMasks"
        o, r = both(ctxs, "get_variance", d)
        assert eq_bits(o, r), "computeImageVariance"
    rng = np.random.RandomState(3)
    for d in (0, S // 2, S - 1):
        both(ctxs, "reproject", d)
        for s in range(S):
            for getter in ("get_proj_warp", "get_proj_color", "get_proj_bias"):
                o, r = both(ctxs, getter, s)
                assert eq_bits(o, r), (getter, d, s)
        for disp in (np.full((H, W), 0.31, np.float32), true_disp[d], rng.uniform(1e-4, 2.0, (H, W)).astype(np.float32)):
            (oc, of), (rc, rf) = both(ctxs, "eval_cost", d, disp)
            assert eq_bits(oc, rc) and eq_bits(of, rf), "computeCost"
        oi, ri = both(ctxs, "brute_force", d, num_depths=150)
        assert eq_bits(oi, ri), "winner indices"
        for o, r in zip(*both(ctxs, "get_disparity", d)):
            assert eq_bits(o, r), "brute-force disparity / cost / confidence"


def test_reference_candidate_count_is_fixed(ref):
    rig, colors, _ = scene_inputs(**RIGS[0][1])
    ctx = capi.Context(ref, capi.rig_descs(rig))
    ctx.level_begin(64, 48)
    ctx.set_colors(colors)
    ctx.reproject(0)
    with pytest.raises(capi.DerpError):
        ctx.brute_force(0, num_depths=32)  # kNumDepths = 150 is compiled in (Derp.h:33)


def test_coverage_check(oracle, ref):
    """Derp.cpp:334-339: CHECK(partialCoverage || useForegroundMasks) — the reference's own CHECK fires."""
    cfg = dict(num_cams=4, width=40, height=40, kind="RECTILINEAR", hfov_deg=60.0)
    rig, colors, _ = scene_inputs(**cfg)
    ctxs = pair(oracle, ref, rig)
    both(ctxs, "level_begin", 40, 40)
    both(ctxs, "set_colors", colors)
    both(ctxs, "reproject", 0)
    for c in ctxs:
        with pytest.raises(capi.DerpError) as e:
            c.brute_force(0, num_depths=150, partial_coverage=False)
        assert e.value.code == capi.ECOVERAGE
    oi, ri = both(ctxs, "brute_force", 0, num_depths=150, partial_coverage=True)
    assert eq_bits(oi, ri) and (oi == -1).any()
    for o, r in zip(*both(ctxs, "get_disparity", 0)):
        assert eq_bits(o, r)


def _start_fine_level(ctxs, colors, W, H, level=1, num_levels=3):
    """Coarse brute force with the ORACLE, handed to both libraries through their own upsampleDisparities."""
    cw, ch = W // 2, H // 2
    coarse_colors = [synth.downscale_area(c, 2) for c in colors]
    oc = ctxs[0]
    oc.level_begin(cw, ch, level=level + 1, num_levels=num_levels, full_width=W, full_height=H)
    oc.set_colors(coarse_colors)
    coarse = []
    for d in range(len(colors)):
        oc.reproject(d)
        oc.brute_force(d, num_depths=150, want_index=False)
        oc.mask_fov(d)
        coarse.append(oc.get_disparity(d, want_cost=False))
    both(ctxs, "level_begin", W, H, level=level, num_levels=num_levels, full_width=W, full_height=H)
    both(ctxs, "set_colors", colors)
    for d in range(len(colors)):
        both(ctxs, "upsample_from", d, coarse[d])


@pytest.mark.parametrize("name,cfg", RIGS[:2])
def test_fine_level_stages(oracle, ref, name, cfg):
    """a13 random proposals (libstdc++ minstd stream), a14 ping-pong, a15 mismatches, a16 joint bilateral, a17 median,
    a18 maskFov, a19 Lanczos upsampling — stage by stage, no re-synchronisation needed: every stage is bit-identical."""
    rig, colors, _ = scene_inputs(**cfg)
    W, H, S = cfg["width"], cfg["height"], len(colors)
    ctxs = pair(oracle, ref, rig)
    _start_fine_level(ctxs, colors, W, H)

    def check(what):
        for d in range(S):
            for o, r in zip(*both(ctxs, "get_disparity", d)):
                assert eq_bits(o, r), (what, d)

    check("upsampleDisparities (Lanczos4)")
    for d in range(S):  # the oracle / CUDA libraries keep one destination's pair tables resident at a time
        both(ctxs, "reproject", d)
        both(ctxs, "random_proposals", d, 2)
        for o, r in zip(*both(ctxs, "get_disparity", d)):
            assert eq_bits(o, r), ("randomProposals", d)
        both(ctxs, "ping_pong", d, 2)
        for o, r in zip(*both(ctxs, "get_disparity", d)):
            assert eq_bits(o, r), ("pingPong", d)
    both(ctxs, "mismatches")
    check("handleDisparityMismatches")
    for d in range(S):
        o, r = both(ctxs, "get_mismatch_mask", d)
        assert eq_bits(o, r)
    for d in range(S):
        both(ctxs, "bilateral", d)
    check("bilateralFilter")
    for d in range(S):
        both(ctxs, "median", d)
    check("medianFilter")
    for d in range(S):
        both(ctxs, "mask_fov", d)
    check("maskFov")


def test_foreground_mask_paths(oracle, ref):
    cfg = RIGS[1][1]
    rig, colors, _ = scene_inputs(**cfg)
    W, H, S = cfg["width"], cfg["height"], len(colors)
    rng = np.random.RandomState(5)
    yy, xx = np.mgrid[0:H, 0:W]
    masks = [(((xx - W / 2 - 4 * s) ** 2 + (yy - H / 2) ** 2) < (0.38 * W) ** 2).astype(np.uint8) for s in range(S)]
    bgs = [np.full((H, W), 0.05, np.float32) + rng.uniform(0, 0.01, (H, W)).astype(np.float32) for _ in range(S)]
    ctxs = pair(oracle, ref, rig)
    both(ctxs, "level_begin", W, H, use_foreground_masks=True)
    both(ctxs, "set_foreground_masks", masks)
    both(ctxs, "set_background_disparity", bgs)
    both(ctxs, "set_colors", colors)
    for d in (0, 3):
        both(ctxs, "reproject", d)
        oi, ri = both(ctxs, "brute_force", d, num_depths=150, partial_coverage=False)
        assert eq_bits(oi, ri)
        both(ctxs, "random_proposals", d, 2)
        both(ctxs, "ping_pong", d, 1)
        both(ctxs, "bilateral", d)
        both(ctxs, "median", d)
        both(ctxs, "mask_fov", d)
        for o, r in zip(*both(ctxs, "get_disparity", d)):
            assert eq_bits(o, r), d
    cw, ch = W // 2, H // 2
    coarse = rng.uniform(0.06, 1.5, (ch, cw)).astype(np.float32)
    coarse[rng.uniform(size=coarse.shape) < 0.1] = np.nan
    cmask = masks[2][::2, ::2].copy()
    both(ctxs, "upsample_from", 2, coarse, cmask, masks[2])
    o, r = both(ctxs, "get_disparity", 2, want_cost=False)
    assert eq_bits(o, r), "masked upsample (nearest + spiral fill + background)"


def test_process_level_three_levels_and_pfm_bytes(oracle, ref, tmp_path):
    """processLevel itself (Derp.cpp:1005-1034), three levels coarse to fine with mismatch handling, each library on its
    own; the reference also runs its saveResults, whose PFM bytes (a20) must equal our writer's."""
    import ctypes as C
    cfg = dict(num_cams=5, width=64, height=64, kind="FTHETA", distorted=True)
    rig, colors, _ = scene_inputs(**cfg)
    W = H = 64
    pyr = [colors, [synth.downscale_area(c, 2) for c in colors], [synth.downscale_area(c, 4) for c in colors]]
    ctxs = pair(oracle, ref, rig)
    prev = None
    for level in (2, 1, 0):
        w = W >> level
        both(ctxs, "level_begin", w, w, level=level, num_levels=3, full_width=W, full_height=H)
        both(ctxs, "set_colors", pyr[level])
        if prev is not None:
            for c, p in zip(ctxs, prev):
                for d in range(5):
                    c.upsample_from(d, p[d])
        both(ctxs, "process_level", num_depths=150, mismatches_start_level=1)
        prev = [[c.get_disparity(d, want_cost=False) for d in range(5)] for c in ctxs]
        for d in range(5):
            assert eq_bits(prev[0][d], prev[1][d]), (level, d)
    f = ref.lib.derp_ref_output_dir
    f.restype, f.argtypes = C.c_char_p, [C.c_void_p]
    out_dir = f(ctxs[1].h).decode()
    pfm = os.path.join(out_dir, "disparity_levels", "level_0", "cam2", "000000.pfm")
    theirs = open(pfm, "rb").read()
    ours = b"Pf\n%d %d\n-1.0\n" % (W, H) + prev[0][2].tobytes()  # the layout io.h's writePfm produces (tests/test_apps.py)
    assert theirs == ours


def test_temporal_and_joint_bilateral(oracle, ref):
    """a22 temporalJointBilateralFilter and the Vec3f-guided generalizedJointBilateralFilter of UpsampleDisparity (a23)."""
    rng = np.random.RandomState(11)
    H, W, T = 40, 44, 5
    base = rng.randint(0, 65536, (H, W, 3))
    guides = [np.clip(base + rng.randint(-300, 300, (H, W, 3)), 0, 65535).astype(np.uint16) for _ in range(T)]
    disps = [rng.uniform(1e-3, 2, (H, W)).astype(np.float32) for _ in range(T)]
    masks = [(rng.uniform(size=(H, W)) > 0.15).astype(np.uint8) for _ in range(T)]
    for off, r in ((2, 1), (0, 2), (4, 0)):
        o = oracle.temporal_filter(guides, disps, masks, off, 0.01, r, 0.5, 1.0, 0.5)
        g = ref.temporal_filter(guides, disps, masks, off, 0.01, r, 0.5, 1.0, 0.5)
        assert eq_bits(o, g)
    guide = guides[0].astype(np.float32) * (np.float32(1.0) / np.float32(65535.0))
    for radius in (2, 5):
        o = oracle.joint_bilateral_f32(disps[0], guide, masks[0], radius, 0.05, 0.5, 0.5, 1.0)
        g = ref.joint_bilateral_f32(disps[0], guide, masks[0], radius, 0.05, 0.5, 0.5, 1.0)
        assert eq_bits(o, g)


def test_standalone_upsample(oracle, ref):
    rig = synth.ring_rig(4, 96, 64, kind="FTHETA")
    d = capi.camera_desc_from_json(rig["cameras"][1])
    rng = np.random.RandomState(2)
    coarse = rng.uniform(1e-4, 2, (32, 48)).astype(np.float32)
    coarse[3:6, 7:9] = np.nan
    for (w, h) in ((96, 64), (100, 70), (48, 32)):
        assert eq_bits(oracle.upsample_disparity(d, coarse, w, h), ref.upsample_disparity(d, coarse, w, h))
    cm = (rng.uniform(size=(32, 48)) > 0.3).astype(np.uint8)
    fm = (rng.uniform(size=(64, 96)) > 0.2).astype(np.uint8)
    bg = rng.uniform(0.01, 0.02, (64, 96)).astype(np.float32)
    assert eq_bits(oracle.upsample_disparity(d, coarse, 96, 64, bg, cm, fm, True),
                   ref.upsample_disparity(d, coarse, 96, 64, bg, cm, fm, True))


def test_reference_rig_file(oracle, ref):
    """The reference's own 16-camera rig (res/test/rigs/rig.json; its cameras are part of tests/golden/camera_vectors.json):
    FOV masks and one destination's warp tables through both libraries at a reduced size."""
    import json
    vec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "camera_vectors.json")))
    cams = [c["json"] for c in vec["cameras"] if c["json"].get("resolution") == [3360, 2160]]
    cams = sorted(cams, key=lambda c: int(c["id"][3:]))
    assert [c["id"] for c in cams] == ["cam%d" % i for i in range(16)]
    rig = {"cameras": cams}
    W, H = 84, 54  # 3360 x 2160 / 40
    S = len(rig["cameras"])
    rng = np.random.RandomState(0)
    colors = [rng.randint(0, 65536, (H, W, 3)).astype(np.uint16) for _ in range(S)]
    ctxs = pair(oracle, ref, rig, dst_to_src=[4, 15, 0])  # DerpTest.cpp:18-26's destination list
    both(ctxs, "level_begin", W, H, full_width=3360, full_height=2160)
    both(ctxs, "set_colors", colors)
    for d in range(3):
        o, r = both(ctxs, "get_fov_mask", d)
        assert eq_bits(o, r)
        both(ctxs, "reproject", d)
        for s in range(S):
            o, r = both(ctxs, "get_proj_warp", s)
            assert eq_bits(o, r), (d, s)
