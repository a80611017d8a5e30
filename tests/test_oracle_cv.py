"""Pins the oracle's OpenCV-primitive restatements to cv2 4.13.0 outputs (tests/golden/cv_vectors.npz,
generator tests/golden/gen_cv_vectors.py)."""
import os

import numpy as np
import pytest

from tests import oracle_hooks as oh

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv_vectors.npz"))


def test_remap_bicubic_bit_exact(oracle):
    for s, d in (("remap_src", "remap_dst"), ("remap_src2", "remap_dst2")):
        got = oh.remap_bicubic(oracle, G[s], G["remap_map"])
        assert np.array_equal(got, G[d]), "mismatches: %d" % (got != G[d]).sum()


def test_blur3_bit_exact(oracle):
    assert np.array_equal(oh.blur3(oracle, G["blur_src"]), G["blur_dst"])
    assert np.array_equal(oh.blur3(oracle, G["blur_src_small"]), G["blur_dst_small"])


@pytest.mark.parametrize("tag", ["var", "var_small", "var_smooth"])
def test_variance_close(oracle, tag):
    got = oh.variance(oracle, G[tag + "_src"])
    ref = G[tag + "_dst"]
    # tolerance class: cv2's float box filter uses SIMD/FMA-dispatched code; 1e-6 of the image scale
    scale = max(np.abs(ref).max(), 1e-12)
    assert np.abs(got - ref).max() <= 1e-6 * scale + 1e-12, np.abs(got - ref).max() / scale


@pytest.mark.parametrize("size", [(62, 46), (50, 37), (31, 23), (100, 80)])
def test_resize(oracle, size):
    W, H = size
    src = G["resize_src"]
    assert np.array_equal(oh.nearest(oracle, src, W, H), G["nearest_%dx%d" % size])
    got = oh.lanczos4(oracle, src, W, H)
    assert np.abs(got - G["lanczos_%dx%d" % size]).max() <= 2e-6


@pytest.mark.parametrize("size", [(42, 27), (21, 27), (28, 27), (51, 33), (84, 54), (13, 9), (50, 32)])
def test_resize_area_bit_exact(oracle, size):
    """cv::resize INTER_AREA u16x3 (scripts/render/resize.py:79): integer and general ratios, bit for bit vs cv2 4.13."""
    w, h = size
    got = oh.resize_area(oracle, G["area_src"], w, h)
    assert np.array_equal(got, G["area_%dx%d" % (w, h)])


@pytest.mark.parametrize("tag,blur,close", [("fg_mask_b1_c4", 1, 4), ("fg_mask_b0_c3", 0, 3), ("fg_mask_b1_c0", 1, 0)])
def test_foreground_mask_bit_exact(oracle, tag, blur, close):
    """generateForegroundMask (BackgroundSubtractionUtil.h:20-59) against the same sequence of cv2 4.13 calls."""
    got = oracle.foreground_mask(G["fg_template"], G["fg_frame"], blur, 0.04, close)
    assert np.array_equal(got, G[tag])
    assert 0 < got.mean() < 1
