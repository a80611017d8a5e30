"""Late additions, GPU leg.  (1) GPU blocks == host instantiation on the hard BC7 surfaces (two-colour blocks, 0 / 255 extremes, near-flat blocks, one
varying channel, the ends of the value range).  The CPU suite already shows the encoder source byte-identical to the
reference on these (tests/test_bc7.py); this file adds the GPU leg.  It was written after the round's GPU minutes were spent
(the five surfaces of tests/test_bc7.py ARE GPU-verified), so it is named to run last under `pytest -x`.
(2) ConvertToBinary --output_formats=...,pfm: the rasterised mesh the app writes next to its .vtx / .idx equals the
reference's mesh_util::writePfm (oracle/_ref) of those very files; the rasteriser itself is host code and is checked
without a GPU in tests/test_mesh.py::test_raster_pfm_equals_reference_write_pfm.
(3) ConvertToBinary --color_scale=0.5 with the bc7 format: the blocks of the surface that
tests/test_apps.py::test_colour_streams_with_color_scale pins to the reference's sequence, through derp_bc7_compress."""
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
