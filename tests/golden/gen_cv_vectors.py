#!/usr/bin/env python3
"""Generates tests/golden/cv_vectors.npz with the cv2 wheel of this image (cv2 4.13.0, single
thread like source/test/DepUnitTest.cpp:14).  Pins the oracle's restatement of the OpenCV primitives
the reference calls (OpenCV itself is not under /root/reference and the reference pins no version):
  remap INTER_CUBIC BORDER_CONSTANT u16x3   (DerpUtil.cpp:199-205 project)
  blur 3x3 u16x3                            (DerpUtil.cpp:208-210 colorBias, CvUtil.h:314-323)
  computeImageVariance                      (DerpUtil.cpp:214-237) via the same cv2 calls
  resize INTER_LANCZOS4 / INTER_NEAREST f32 (UpsampleDisparityLib.cpp:125,145)
  resize INTER_AREA u16x3, shrinking          (scripts/render/resize.py:79, UpsampleDisparity.cpp:117)
"""
import os

import cv2
import numpy as np

cv2.setNumThreads(1)


def main():
    rng = np.random.RandomState(12345)
    out = {}
    # ---- remap
    sh, sw, dh, dw = 37, 53, 41, 47
    src = rng.randint(0, 65536, size=(sh, sw, 3)).astype(np.uint16)
    mp = np.empty((dh, dw, 2), np.float32)
    mp[..., 0] = rng.uniform(-6, sw + 5, size=(dh, dw))
    mp[..., 1] = rng.uniform(-6, sh + 5, size=(dh, dw))
    # exact-grid and half-grid coordinates, NaNs, far outside
    mp[0, :10, 0] = np.arange(10)
    mp[0, :10, 1] = 5
    mp[1, :10, 0] = np.arange(10) + 0.5
    mp[1, :10, 1] = 5.5
    mp[2, :8] = np.nan
    mp[3, :4, 0] = 1e9
    mp[3, 4:8, 1] = -1e9
    mp[4, :16, 0] = np.arange(16) / 64.0 + 3  # 1/64 steps: exercises the 1/32 rounding (half-even)
    mp[4, :16, 1] = 7.015625
    out["remap_src"] = src
    out["remap_map"] = mp
    out["remap_dst"] = cv2.remap(src, mp, None, cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)
    # smooth-ish second case (values near saturation)
    src2 = np.clip(rng.normal(60000, 6000, size=(sh, sw, 3)), 0, 65535).astype(np.uint16)
    out["remap_src2"] = src2
    out["remap_dst2"] = cv2.remap(src2, mp, None, cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)
    # ---- blur
    img = rng.randint(0, 65536, size=(29, 31, 3)).astype(np.uint16)
    out["blur_src"] = img
    out["blur_dst"] = cv2.blur(img, (3, 3))
    small = rng.randint(0, 12, size=(9, 7, 3)).astype(np.uint16)
    out["blur_src_small"] = small
    out["blur_dst_small"] = cv2.blur(small, (3, 3))
    # ---- variance (same call sequence as the reference)
    for tag, im in (("var", img), ("var_small", small),
                    ("var_smooth", np.clip(rng.normal(30000, 900, size=(33, 35, 3)), 0, 65535).astype(np.uint16))):
        # cv::Mat::convertTo(CV_32F, 1.0f/65535.0f) (CvUtil.h:171-183): float multiply
        f = im.astype(np.float32) * (np.float32(1.0) / np.float32(65535.0))
        mean = cv2.blur(f, (3, 3))
        msq = cv2.blur(f * f, (3, 3))
        v = msq - mean * mean
        w = np.array([0.3333, 0.3334, 0.3333], np.float32)
        var = v[..., 0] * w[2] + v[..., 1] * w[1] + v[..., 2] * w[0]
        out[tag + "_src"] = im
        out[tag + "_dst"] = var.astype(np.float32)
    # ---- resize
    d = rng.uniform(1e-4, 2.0, size=(23, 31)).astype(np.float32)
    out["resize_src"] = d
    for (W, H) in ((62, 46), (50, 37), (31, 23), (100, 80)):
        out["lanczos_%dx%d" % (W, H)] = cv2.resize(d, (W, H), interpolation=cv2.INTER_LANCZOS4)
        out["nearest_%dx%d" % (W, H)] = cv2.resize(d, (W, H), interpolation=cv2.INTER_NEAREST)
    # ---- INTER_AREA (pyramid pre-resize): integer ratios (2x2 fast path, 4x4, 3x2) and general ratios, including the
    # reference rig's 3360 x 2160 -> 2048 x 1318 ratio at a reduced size
    a = rng.randint(0, 65536, size=(54, 84, 3)).astype(np.uint16)
    out["area_src"] = a
    for (W, H) in ((42, 27), (21, 27), (28, 27), (51, 33), (84, 54), (13, 9), (50, 32)):
        out["area_%dx%d" % (W, H)] = cv2.resize(a, (W, H), interpolation=cv2.INTER_AREA)
    # ---- generateForegroundMask (BackgroundSubtractionUtil.h:20-59) with the same cv2 calls: GaussianBlur 3x3 sigma 0,
    # convertTo float, absdiff, per-pixel L2 norm (double) > 0.04, MORPH_CLOSE with a 4x4 rectangle
    bgimg = np.clip(rng.normal(30000, 9000, size=(40, 52, 3)), 0, 65535).astype(np.uint16)
    fr = bgimg.copy()
    fr[8:22, 10:30] = np.clip(fr[8:22, 10:30].astype(np.int64) + rng.randint(-9000, 9000, size=(14, 20, 3)), 0, 65535).astype(np.uint16)
    fr = np.clip(fr.astype(np.int64) + rng.randint(-600, 600, size=fr.shape), 0, 65535).astype(np.uint16)
    out["fg_template"] = bgimg
    out["fg_frame"] = fr
    out["fg_gauss"] = cv2.GaussianBlur(fr, (3, 3), 0)
    for tag, blur, close in (("fg_mask_b1_c4", 1, 4), ("fg_mask_b0_c3", 0, 3), ("fg_mask_b1_c0", 1, 0)):
        tb = cv2.GaussianBlur(bgimg, (3, 3), 0) if blur else bgimg
        fb = cv2.GaussianBlur(fr, (3, 3), 0) if blur else fr
        a32 = np.float32(1.0) / np.float32(65535.0)
        diff = cv2.absdiff(tb.astype(np.float32) * a32, fb.astype(np.float32) * a32)
        mask = (np.sqrt((diff.astype(np.float64) ** 2).sum(-1)) > np.float64(np.float32(0.04))).astype(np.uint8)
        if close:
            mask = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (close, close)))
        out[tag] = mask
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cv_vectors.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, "cv2", cv2.__version__)


if __name__ == "__main__":
    main()
