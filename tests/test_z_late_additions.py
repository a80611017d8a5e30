"""Late additions, GPU leg.  (1) GPU blocks == host instantiation on the hard BC7 surfaces (two-colour blocks, 0 / 255 extremes, near-flat blocks, one
varying channel, the ends of the value range).  The CPU suite already shows the encoder source byte-identical to the
reference on these (tests/test_bc7.py); this file adds the GPU leg.  It was written after the round's GPU minutes were spent
(the five surfaces of tests/test_bc7.py ARE GPU-verified), so it is named to run last under `pytest -x`.
(2) ConvertToBinary --output_formats=...,pfm: the rasterised mesh the app writes next to its .vtx / .idx equals the
reference's mesh_util::writePfm (oracle/_ref) of those very files; the rasteriser itself is host code and is checked
without a GPU in tests/test_mesh.py::test_raster_pfm_equals_reference_write_pfm.
(3) ConvertToBinary --color_scale=0.5 with the bc7 format: the blocks of the surface that
tests/test_apps.py::test_colour_streams_with_color_scale pins to the reference's sequence, through derp_bc7_compress.
(4) GenerateForegroundMasks --blur_radius=2 / 3: blurred on the host (pinned to cv2 without a GPU), the library's
verified difference + closing kernels after it.
(5) UpsampleDisparity with a colour guide larger / smaller than the output: the guide resized like cv_util::resizeImage
(cv2.resize INTER_AREA of the float image), then the oracle's upsampling and guided filter."""
import numpy as np
import pytest

from tests.test_bc7 import HARD_KINDS, host_blocks, surface


@pytest.mark.gpu
@pytest.mark.parametrize("kind", HARD_KINDS)
def test_gpu_blocks_equal_host_instantiation_hard_surfaces(cuda, kind):
    for (w, h) in ((256, 192), (70, 50)):
        rgba = surface(6, w, h, kind)
        assert np.array_equal(cuda.bc7_compress(rgba), host_blocks(rgba))


@pytest.mark.gpu
def test_convert_to_binary_raster_pfm(tmp_path, cuda):
    import ctypes as C
    import json
    import os
    from tests import oracle_libs
    from tests.test_apps import _mesh_dataset, run
    ref = oracle_libs.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rig, disps = _mesh_dataset(tmp_path, F=1)
    H, W = next(iter(disps.values())).shape
    run("ConvertToBinary", "--rig=" + str(tmp_path / "rigs" / "rig.json"), "--first=000000", "--last=000000",
        "--disparity=" + str(tmp_path / "disparity"), "--bin=" + str(tmp_path / "bin"), "--triangles=2000",
        "--output_formats=idx,vtx,pfm")
    f = ref.lib.derp_ref_write_raster_pfm
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_char_p]
    for cam in rig["cameras"]:
        stem = str(tmp_path / "bin" / cam["id"] / "000000")
        v = np.fromfile(stem + ".vtx", np.float32).reshape(-1, 3).astype(np.float64)
        i = np.fromfile(stem + ".idx", np.uint32).reshape(-1, 3)
        want = str(tmp_path / (cam["id"] + "_ref.pfm"))
        assert f(v.ctypes.data, len(v), i.ctypes.data, len(i), W, H, cam["resolution"][0], cam["resolution"][1], want.encode()) == 0
        assert open(stem + ".pfm", "rb").read() == open(want, "rb").read()


@pytest.mark.gpu
def test_convert_to_binary_bc7_with_color_scale(tmp_path, cuda):
    import cv2
    import json
    import os
    from facebook360_dep_b200 import synth
    from tests.test_apps import run
    W, H = 64, 48
    rig = synth.ring_rig(2, W, H, kind="FTHETA")
    os.makedirs(tmp_path / "rigs", exist_ok=True)
    json.dump(rig, open(tmp_path / "rigs" / "rig.json", "w"))
    rng = np.random.RandomState(11)
    yy, xx = np.mgrid[0:H, 0:W]
    for cam in rig["cameras"]:
        img = np.stack([0.5 + 0.4 * np.sin(xx / 9.), 0.5 + 0.4 * np.cos(yy / 7.), (xx + yy) / float(W + H)], -1)
        img = ((img + rng.normal(0, 0.01, img.shape)).clip(0, 1) * 65535).astype(np.uint16)
        os.makedirs(tmp_path / "color" / cam["id"], exist_ok=True)
        assert cv2.imwrite(str(tmp_path / "color" / cam["id"] / "000000.png"), img)
    run("ConvertToBinary", "--rig=" + str(tmp_path / "rigs" / "rig.json"), "--first=000000", "--last=000000",
        "--color=" + str(tmp_path / "color"), "--bin=" + str(tmp_path / "bin"), "--output_formats=bc7", "--color_scale=0.5")
    for cam in rig["cameras"]:
        surface = str(tmp_path / (cam["id"] + ".surface"))
        run("IoSelfTest", "--mode=bc7surface", "--in=" + str(tmp_path / "color" / cam["id"] / "000000.png"), "--scale=0.5",
            "--out=" + surface)
        rgba = np.fromfile(surface, np.uint8).reshape(H // 2, W // 2, 4)
        got = np.fromfile(str(tmp_path / "bin" / cam["id"] / "000000.bc7"), np.uint8)
        assert np.array_equal(got, cuda.bc7_compress(rgba))


@pytest.mark.gpu
@pytest.mark.parametrize("radius", [2, 3])
def test_generate_foreground_masks_larger_blur(tmp_path, cuda, radius):
    import cv2
    import json
    import os
    from facebook360_dep_b200 import synth
    from tests.test_apps import run
    S, Wf, Hf = 2, 120, 90
    rig = synth.ring_rig(S, Wf, Hf, kind="FTHETA")
    os.makedirs(tmp_path / "rigs", exist_ok=True)
    json.dump(rig, open(tmp_path / "rigs" / "rig.json", "w"))
    rng = np.random.RandomState(2)
    bgdir, fgdir, mdir = str(tmp_path / "bg"), str(tmp_path / "fg"), str(tmp_path / "masks")
    ids = [c["id"] for c in rig["cameras"]]
    for s, cid in enumerate(ids):
        bgimg = np.clip(rng.normal(30000, 9000, (Hf, Wf, 3)), 0, 65535).astype(np.uint16)
        fr = bgimg.copy()
        fr[20:60, 30 + 5 * s:80] = rng.randint(0, 65536, (40, 50 - 5 * s, 3)).astype(np.uint16)
        os.makedirs(os.path.join(bgdir, cid))
        os.makedirs(os.path.join(fgdir, cid))
        cv2.imwrite(os.path.join(bgdir, cid, "000000.png"), bgimg)
        cv2.imwrite(os.path.join(fgdir, cid, "000007.png"), fr)
    run("GenerateForegroundMasks", "--rig=" + str(tmp_path / "rigs" / "rig.json"), "--color=" + fgdir, "--background_color=" + bgdir,
        "--foreground_masks=" + mdir, "--first=000007", "--last=000007", "--width=80", "--blur_radius=%d" % radius)
    Wo, Ho = 80, int(np.rint(80 * Hf / np.float32(Wf)))
    a32 = np.float32(1.0) / np.float32(65535.0)
    k = 2 * radius + 1
    for cid in ids:
        b = cv2.resize(cv2.imread(os.path.join(bgdir, cid, "000000.png"), cv2.IMREAD_UNCHANGED), (Wo, Ho), interpolation=cv2.INTER_AREA)
        f = cv2.resize(cv2.imread(os.path.join(fgdir, cid, "000007.png"), cv2.IMREAD_UNCHANGED), (Wo, Ho), interpolation=cv2.INTER_AREA)
        diff = cv2.absdiff(cv2.GaussianBlur(b, (k, k), 0).astype(np.float32) * a32, cv2.GaussianBlur(f, (k, k), 0).astype(np.float32) * a32)
        m = (np.sqrt((diff.astype(np.float64) ** 2).sum(-1)) > np.float64(np.float32(0.04))).astype(np.uint8)
        m = cv2.morphologyEx(m, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (4, 4)))
        got = cv2.imread(os.path.join(mdir, cid, "000007.png"), cv2.IMREAD_UNCHANGED)
        assert got.dtype == np.uint8 and np.array_equal(got, m * 255), cid


@pytest.mark.gpu
@pytest.mark.parametrize("guide_size", [192, 64])  # larger than the 96-wide output (area sums), smaller (bilinear variant)
def test_upsample_disparity_with_resized_guide(tmp_path, cuda, oracle, guide_size):
    import cv2
    import json
    import os
    from facebook360_dep_b200 import capi, synth
    from tests.test_apps import read_pfm, run, write_pfm
    S, W = 2, 48
    rig = synth.ring_rig(S, W, W, kind="FTHETA")
    os.makedirs(tmp_path / "rigs", exist_ok=True)
    json.dump(rig, open(tmp_path / "rigs" / "rig.json", "w"))
    rng = np.random.RandomState(3)
    yy, xx = np.mgrid[0:W, 0:W].astype(np.float32)
    guides, coarse = {}, {}
    for cam in rig["cameras"]:
        g = rng.randint(0, 65536, (guide_size, guide_size, 3)).astype(np.uint16)
        os.makedirs(tmp_path / "color" / cam["id"], exist_ok=True)
        assert cv2.imwrite(str(tmp_path / "color" / cam["id"] / "000000.png"), g)
        guides[cam["id"]] = g
        d = (0.3 + 0.1 * np.sin(xx / 9.0) * np.cos(yy / 7.0)).astype(np.float32)
        write_pfm(str(tmp_path / "disparity" / cam["id"] / "000000.pfm"), d)
        coarse[cam["id"]] = d
    run("UpsampleDisparity", "--rig=" + str(tmp_path / "rigs" / "rig.json"), "--disparity=" + str(tmp_path / "disparity"),
        "--output=" + str(tmp_path / "up"), "--resolution=96", "--color=" + str(tmp_path / "color"), "--first=000000", "--last=000000")
    for cam in rig["cameras"]:
        up = read_pfm(os.path.join(str(tmp_path / "up"), cam["id"], "000000.pfm"))
        ref = oracle.upsample_disparity(capi.camera_desc_from_json(cam), coarse[cam["id"]], 96, 96)
        f = guides[cam["id"]].astype(np.float32) * (np.float32(1.0) / np.float32(65535.0))
        guide = cv2.resize(f, (96, 96), interpolation=cv2.INTER_AREA)
        radius = int((96.0 / W) ** 2 + 1)
        ref = oracle.joint_bilateral_f32(ref, guide, np.ones((96, 96), np.uint8), radius, 0.05, 0.5, 0.5, 1.0)
        assert (np.abs(up - ref) <= 2e-6 * np.abs(ref) + 2e-7).all()
