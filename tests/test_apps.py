"""The drop-in executables (facebook360_dep_b200/bin): flag surface, file formats, and — on the GPU box —
end-to-end runs over a synthetic dataset laid out like the reference's input tree."""
import ctypes as C
import json
import os
import re
import subprocess

import cv2
import numpy as np
import pytest

from facebook360_dep_b200 import capi, synth

ROOT = capi.ROOT
BIN = os.path.join(ROOT, "facebook360_dep_b200", "bin")
HOST = os.path.join(ROOT, "facebook360_dep_b200", "csrc", "host")


@pytest.fixture(scope="module", autouse=True)
def built_apps():
    subprocess.check_call(["make", "-C", HOST], stdout=subprocess.DEVNULL)


def run(app, *args, check=True):
    p = subprocess.run([os.path.join(BIN, app)] + list(args), capture_output=True, text=True)
    if check and p.returncode != 0:
        raise AssertionError("%s failed (%d):\n%s" % (app, p.returncode, p.stderr[-2000:]))
    return p


# flag tables of the reference apps: name -> (type, default) (DerpCLI.cpp:40-67,
# TemporalBilateralFilter.cpp:40-59, UpsampleDisparity.cpp:37-55)
REF_FLAGS = {
    "DerpCLI": {
        "background_disp": ("string", ""), "background_frame": ("string", "000000"), "cameras": ("string", ""),
        "color": ("string", ""), "do_bilateral_filter": ("bool", "true"), "do_median_filter": ("bool", "true"),
        "first": ("string", "000000"), "foreground_masks": ("string", ""), "input_root": ("string", ""),
        "last": ("string", "000000"), "level_end": ("int32", "-1"), "level_start": ("int32", "-1"),
        "max_depth_m": ("double", "1e4"), "min_depth_m": ("double", ".50"), "mismatches_start_level": ("int32", "-1"),
        "num_levels": ("int32", "-1"), "output_formats": ("string", ""), "output_root": ("string", ""),
        "partial_coverage": ("bool", "false"), "ping_pong_iterations": ("int32", "1"),
        "random_proposals": ("int32", "2"), "resolution": ("int32", "2048"), "rig": ("string", ""),
        "save_debug_images": ("bool", "false"), "threads": ("int32", "-1"), "use_foreground_masks": ("bool", "false"),
        "var_high_thresh": ("double", "1e-3"), "var_noise_floor": ("double", "4e-5"),
    },
    "TemporalBilateralFilter": {
        "color": ("string", ""), "cameras": ("string", ""), "disparity": ("string", ""), "first": ("string", "000000"),
        "foreground_masks": ("string", ""), "input_root": ("string", ""), "last": ("string", "000000"),
        "level": ("int32", "0"), "output_formats": ("string", ""), "output_root": ("string", ""),
        "resolution": ("int32", "2048"), "rig": ("string", ""), "sigma": ("double", "0.01"),
        "space_radius": ("int32", "-1"), "threads": ("int32", "-1"), "time_radius": ("int32", "2"),
        "use_foreground_masks": ("bool", "false"), "weight_b": ("double", "0.5"), "weight_g": ("double", "1.0"),
        "weight_r": ("double", "1.0"),
    },
    "LayerDisparities": {  # LayerDisparities.cpp:36-44
        "background_disp": ("string", ""), "background_frame": ("string", "000000"), "cameras": ("string", ""),
        "first": ("string", "000000"), "foreground_disp": ("string", ""), "last": ("string", "000000"),
        "output": ("string", ""), "rig": ("string", ""), "threads": ("int32", "-1"),
    },
    "GenerateForegroundMasks": {  # source/render/GenerateForegroundMasks.cpp:43-55
        "background_color": ("string", ""), "background_frame": ("string", "000000"), "blur_radius": ("int32", "1"),
        "cameras": ("string", ""), "color": ("string", ""), "first": ("string", ""), "foreground_masks": ("string", ""),
        "last": ("string", ""), "morph_closing_size": ("int32", "4"), "rig": ("string", ""), "threads": ("int32", "-1"),
        "threshold": ("double", "0.04"), "width": ("int32", "2048"),
    },
    "ConvertToBinary": {  # source/mesh_stream/ConvertToBinary.cpp:62-87
        "bin": ("string", "bin"), "cameras": ("string", ""), "color": ("string", ""), "color_scale": ("double", "1"),
        "depth_scale": ("double", "1"), "disparity": ("string", ""), "first": ("string", ""),
        "foreground_masks": ("string", ""), "fuse_strip": ("int32", "1"), "fused": ("string", ""),
        "gamma_correction": ("expr", "2.2 / 1.8"), "last": ("string", ""), "output_formats": ("string", "idx,vtx,bc7"), "rig": ("string", ""),
        "run_conversion": ("bool", "true"), "tear_ratio": ("double", "0.95"), "threads": ("int32", "-1"),
        "triangles": ("int32", "150000"),
    },
    "UpsampleDisparity": {
        "background_disp": ("string", ""), "background_frame": ("string", "000000"), "cameras": ("string", ""),
        "color": ("string", ""), "disparity": ("string", ""), "first": ("string", "000000"),
        "foreground_masks_in": ("string", ""), "foreground_masks_out": ("string", ""), "height": ("int32", "-1"),
        "last": ("string", "000000"), "output": ("string", ""), "output_formats": ("string", ""),
        "resolution": ("int32", "-1"), "rig": ("string", ""), "sigma": ("double", "0.05"), "threads": ("int32", "-1"),
        "weight_b": ("double", "0.5"), "weight_g": ("double", "0.5"), "weight_r": ("double", "1.0"),
    },
}


@pytest.mark.parametrize("app", sorted(REF_FLAGS))
def test_flag_surface_matches_reference(app):
    """The render pipeline learns an app's flags by scraping DEFINE_ lines from its .cpp
    (scripts/util/system_util.py:123-176): the same scrape of our sources must yield the reference's flags
    with the reference's types and defaults (extra B200 flags allowed)."""
    src = open(os.path.join(HOST, app + ".cpp")).read()
    found = {}
    for m in re.finditer(r'DEFINE_(\w+)\(\s*(\w+)\s*,\s*("[^"]*"|[^,]*?)\s*,\s*"', src):
        found[m.group(2)] = (m.group(1), m.group(3).strip().strip('"'))
    for name, (typ, default) in REF_FLAGS[app].items():
        assert name in found, name
        if typ == "expr":  # a double whose default is written as an expression
            assert found[name] == ("double", default), name
            continue
        assert found[name][0] == typ, name
        if typ == "double":
            assert float(found[name][1]) == float(default), name
        else:
            assert found[name][1] == default, name
    extra = set(found) - set(REF_FLAGS[app])
    assert extra <= {"num_depths", "gpus", "gpu"}, extra
    h = run(app, "--help", check=False)
    for name in REF_FLAGS[app]:
        assert "-" + name + " " in h.stdout


def test_png_pfm_io_matches_cv2(tmp_path):
    rng = np.random.RandomState(0)
    cases = {
        "c16": rng.randint(0, 65536, (13, 17, 3)).astype(np.uint16),
        "c8": rng.randint(0, 256, (11, 9, 3)).astype(np.uint8),
        "g16": rng.randint(0, 65536, (7, 21)).astype(np.uint16),
        "a8": rng.randint(0, 256, (9, 10, 4)).astype(np.uint8),
    }
    for name, img in cases.items():
        p = str(tmp_path / (name + ".png"))
        assert cv2.imwrite(p, img)
        out = str(tmp_path / (name + ".raw"))
        r = run("IoSelfTest", "--in=" + p, "--mode=color", "--out=" + out)
        w, h = map(int, r.stdout.split()[-2:])
        got = np.fromfile(out, np.uint16).reshape(h, w, 3)
        # cv_util::loadImage<Vec3w>: depth -> 16U (x257 for 8 bit), gray -> BGR, alpha dropped (CvUtil.h:227-284)
        ref = img if img.dtype == np.uint16 else img.astype(np.uint16) * 257
        if ref.ndim == 2:
            ref = np.repeat(ref[..., None], 3, axis=2)
        assert np.array_equal(got, ref[..., :3]), name
    # masks: threshold 127 on the 8-bit value
    m = rng.randint(0, 256, (8, 12)).astype(np.uint8)
    p = str(tmp_path / "m.png")
    cv2.imwrite(p, m)
    out = str(tmp_path / "m.raw")
    run("IoSelfTest", "--in=" + p, "--mode=mask", "--out=" + out)
    assert np.array_equal(np.fromfile(out, np.uint8).reshape(8, 12), (m > 127).astype(np.uint8))
    # PFM: header + top-down little-endian rows (CvUtil.cpp:39-49); png = u16(clamp(d,0,1)*65535), NaN -> 0
    d = rng.uniform(-0.2, 1.3, (6, 5)).astype(np.float32)
    d[2, 3] = np.nan
    pf = tmp_path / "d.pfm"
    with open(pf, "wb") as f:
        f.write(b"Pf\n5 6\n-1.0\n")
        f.write(d.tobytes())
    run("IoSelfTest", "--in=" + str(pf), "--mode=float", "--out=" + str(tmp_path / "o.pfm"))
    assert open(tmp_path / "o.pfm", "rb").read() == open(pf, "rb").read()
    run("IoSelfTest", "--in=" + str(pf), "--mode=float", "--out=" + str(tmp_path / "o.png"))
    png = cv2.imread(str(tmp_path / "o.png"), cv2.IMREAD_UNCHANGED)
    ref = np.nan_to_num(np.clip(d, 0, 1) * 65535.0, nan=0.0)
    assert png.dtype == np.uint16 and np.array_equal(png, np.rint(ref).astype(np.uint16))


def test_rig_json_parser_matches_python_binding(tmp_path):
    rig = json.load(open(os.path.join(ROOT, "tests", "golden", "camera_vectors.json")))
    cams = [c["json"] for c in rig["cameras"][:16]]
    p = tmp_path / "rig.json"
    json.dump({"cameras": cams}, open(p, "w"))
    out = tmp_path / "rig.bin"
    r = run("IoSelfTest", "--in=" + str(p), "--mode=rig", "--out=" + str(out))
    ids = r.stdout.split()[:16]
    assert ids == [c["id"] for c in cams]
    raw = open(out, "rb").read()
    descs = capi.rig_descs({"cameras": cams})
    assert raw == bytes(descs)


def write_dataset(root, rig, frames, levels, masks=None):
    """<root>/video/color_levels/level_L/<cam>/<frame>.png + rigs/rig_calibrated.json (ImageTypes.h:16-47)."""
    os.makedirs(os.path.join(root, "rigs"), exist_ok=True)
    json.dump(rig, open(os.path.join(root, "rigs", "rig_calibrated.json"), "w"))
    for f, colors in enumerate(frames):
        for L in range(levels):
            for cam, img in zip(rig["cameras"], colors):
                d = os.path.join(root, "video", "color_levels", "level_%d" % L, cam["id"])
                os.makedirs(d, exist_ok=True)
                im = img if L == 0 else synth.downscale_area(img, 1 << L)
                assert cv2.imwrite(os.path.join(d, "%06d.png" % f), im)


def read_pfm(path):
    with open(path, "rb") as f:
        assert f.readline() == b"Pf\n"
        w, h = map(int, f.readline().split())
        f.readline()
        return np.frombuffer(f.read(), np.float32).reshape(h, w)


def test_gpu_worker_rendezvous():
    """The barrier + published-address table DerpCLI's per-GPU worker threads use around camera-sharded mismatch
    handling (exchange.h), hammered on CPU threads: 5 parties x 200 rounds, every party must see all values of the
    round between the two barriers."""
    r = run("IoSelfTest", "--mode=exchange")
    assert "exchange ok" in r.stdout


def write_pfm(path, a):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    a = np.ascontiguousarray(a, np.float32)
    with open(path, "wb") as f:
        f.write(b"Pf\n%d %d\n-1.0\n" % (a.shape[1], a.shape[0]))
        f.write(a.tobytes())


def test_layer_disparities_matches_opencv(tmp_path):
    """LayerDisparities is host code only (file IO + one select per pixel); its output must equal what the reference's
    OpenCV expression writes: threshold(fg, 0, 1, BINARY) mask, fg.mul(mask) + bg.mul(1 - mask), imwrite(x * 255) of a
    float matrix (= convertTo 8U).  Pinned with cv2 itself, including NaN / negative / > 1 / half-way values."""
    rig = synth.ring_rig(3, 40, 24)
    os.makedirs(tmp_path / "rigs")
    json.dump(rig, open(tmp_path / "rigs" / "rig.json", "w"))
    rng = np.random.RandomState(4)
    H, W = 24, 40
    want = {}
    for cam in ("cam0", "cam2"):
        bg = rng.uniform(0, 1.2, (H, W)).astype(np.float32)
        bg[3, 5] = np.nan
        write_pfm(str(tmp_path / "bg" / cam / "000000.pfm"), bg)
        for f in (7, 8):
            fg = rng.uniform(-0.2, 1.1, (H, W)).astype(np.float32)
            fg[rng.uniform(size=(H, W)) < 0.3] = np.nan
            fg[0, :8] = np.array([0.0, -0.0, 0.5 / 255, 1.5 / 255, 2.5 / 255, 254.5 / 255, 1.0, 300.0], np.float32)
            fg[1, 0] = np.inf
            write_pfm(str(tmp_path / "fg" / cam / ("%06d.pfm" % f)), fg)
            with np.errstate(invalid="ignore"):
                mask = cv2.threshold(fg, 0.0, 1.0, cv2.THRESH_BINARY)[1]
                layer = fg * mask + bg * (1 - mask)
                ref_path = str(tmp_path / ("ref_%s_%d.png" % (cam, f)))
                assert cv2.imwrite(ref_path, layer * 255)
            want[(cam, f)] = cv2.imread(ref_path, cv2.IMREAD_UNCHANGED)
    run("LayerDisparities", "--rig=" + str(tmp_path / "rigs" / "rig.json"), "--background_disp=" + str(tmp_path / "bg"),
        "--foreground_disp=" + str(tmp_path / "fg"), "--output=" + str(tmp_path / "out"), "--first=000007", "--last=000008",
        "--cameras=cam0,cam2")
    for (cam, f), ref in want.items():
        got = cv2.imread(str(tmp_path / "out" / "disparity" / cam / ("%06d.png" % f)), cv2.IMREAD_UNCHANGED)
        assert got is not None and got.dtype == np.uint8 and got.shape == (H, W)
        assert np.array_equal(got, ref), (cam, f, np.argwhere(got != ref)[:5])
    assert not os.path.exists(tmp_path / "out" / "disparity" / "cam1")


def test_derpcli_aborts_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rig = synth.ring_rig(3, 32, 32)
    colors, _ = synth.render_rig(rig, 32, 32)
    write_dataset(str(tmp_path / "in"), rig, [colors], 1)
    p = run("DerpCLI", "--input_root=" + str(tmp_path / "in"), "--output_root=" + str(tmp_path / "out"),
            "--partial_coverage", check=False)
    assert p.returncode != 0 and "CUDA" in p.stderr  # glog-style FATAL + abort: no CPU fallback


@pytest.mark.gpu
def test_apps_end_to_end(tmp_path, cuda, oracle):
    W = H = 96
    S, F = 5, 3
    rig = synth.ring_rig(S, W, H, kind="FTHETA")
    frames = []
    for f in range(F):
        colors, _ = synth.render_rig(rig, W, H, scene=synth.Scene(seed=42, shift=(0.01 * f, 0, 0)))
        frames.append(colors)
    inp, out = str(tmp_path / "in"), str(tmp_path / "out")
    write_dataset(inp, rig, frames, 2)
    run("DerpCLI", "--input_root=" + inp, "--output_root=" + out, "--first=000000", "--last=000002",
        "--partial_coverage=true", "--num_depths=48", "--output_formats=png", "--gpus=1")
    # same pipeline through the Python binding of the same library: PFMs must be byte-identical
    descs = capi.rig_descs(rig)
    for lib, exact in ((cuda, True), (oracle, False)):
        ctx = capi.Context(lib, descs)
        for f in range(F):
            coarse = [synth.downscale_area(c, 2) for c in frames[f]]
            ctx.level_begin(W // 2, H // 2, level=1, num_levels=2, full_width=W, full_height=H)
            ctx.set_colors(coarse)
            ctx.process_level(num_depths=48)
            c1 = [ctx.get_disparity(d, want_cost=False) for d in range(S)]
            ctx.level_begin(W, H, level=0, num_levels=2, full_width=W, full_height=H)
            ctx.set_colors(frames[f])
            for d in range(S):
                if exact:  # the app re-reads the coarser level from its PFM, which is lossless
                    ctx.upsample_from(d, c1[d])
                else:
                    ctx.upsample_from(d, c1[d])
            ctx.process_level(num_depths=48)
            for L, disps in ((1, c1), (0, [ctx.get_disparity(d, want_cost=False) for d in range(S)])):
                for d in range(S):
                    got = read_pfm(os.path.join(out, "disparity_levels", "level_%d" % L, "cam%d" % d, "%06d.pfm" % f))
                    ref = disps[d]
                    if exact:
                        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (L, d, f)
                    else:
                        assert np.array_equal(np.isnan(got), np.isnan(ref))
                        fin = ~np.isnan(ref)
                        assert (np.abs(got - ref)[fin] <= 1e-3 * np.abs(ref)[fin]).mean() >= 0.999
        ctx.close()
    png = cv2.imread(os.path.join(out, "disparity_levels", "level_0", "cam0", "000001.png"), cv2.IMREAD_UNCHANGED)
    assert png is not None and png.dtype == np.uint16 and png.shape == (H, W)
    # temporal filter over the 3 frames, level 0
    run("TemporalBilateralFilter", "--input_root=" + inp, "--output_root=" + out, "--rig=" + inp + "/rigs/rig_calibrated.json",
        "--first=000001", "--last=000001", "--level=0", "--time_radius=2")
    tf = read_pfm(os.path.join(out, "disparity_time_filtered_levels", "level_0", "cam2", "000001.pfm"))
    ctx = capi.Context(oracle, descs)
    ctx.level_begin(W, H)
    fov = ctx.get_fov_mask(2)
    guides = [frames[f][2] for f in range(F)]
    disps = [read_pfm(os.path.join(out, "disparity_levels", "level_0", "cam2", "%06d.pfm" % f)) for f in range(F)]
    ref = oracle.temporal_filter(guides, disps, [fov] * F, 1, 0.01, 1, 0.5, 1.0, 0.5)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(tf), fin)
    assert (np.abs(tf - ref)[fin] <= 2e-6 * np.abs(ref)[fin]).all()
    # upsample level-1 disparity to 96 wide with colour guidance
    run("UpsampleDisparity", "--rig=" + inp + "/rigs/rig_calibrated.json", "--disparity=" + out + "/disparity_levels/level_1",
        "--output=" + out + "/disparity_upsample", "--resolution=96", "--color=" + inp + "/video/color_levels/level_0",
        "--first=000000", "--last=000000", "--cameras=cam1,cam3")
    up = read_pfm(os.path.join(out, "disparity_upsample", "cam3", "000000.pfm"))
    coarse = read_pfm(os.path.join(out, "disparity_levels", "level_1", "cam3", "000000.pfm"))
    d3 = capi.camera_desc_from_json(rig["cameras"][3])
    ref = oracle.upsample_disparity(d3, coarse, 96, 96)
    guide = frames[0][3].astype(np.float32) * (np.float32(1.0) / np.float32(65535.0))
    ref = oracle.joint_bilateral_f32(ref, guide, np.ones((96, 96), np.uint8), 5, 0.05, 0.5, 0.5, 1.0)
    assert not os.path.exists(os.path.join(out, "disparity_upsample", "cam0"))
    # Lanczos ringing next to the NaN->1e-4 fill gives values near zero: absolute floor on the tolerance
    assert (np.abs(up - ref) <= 2e-6 * np.abs(ref) + 2e-7).all()


def _read_exr_y(path):
    """Minimal reader of the single-channel float scan-line EXR io.h writes (OpenEXR file layout)."""
    import struct
    b = open(path, "rb").read()
    assert struct.unpack("<I", b[:4])[0] == 20000630 and struct.unpack("<I", b[4:8])[0] == 2
    pos, attrs = 8, {}
    while b[pos] != 0:
        e = b.index(b"\0", pos)
        name = b[pos:e].decode()
        pos = e + 1
        e = b.index(b"\0", pos)
        typ = b[pos:e].decode()
        pos = e + 1
        size = struct.unpack("<I", b[pos:pos + 4])[0]
        attrs[name] = (typ, b[pos + 4:pos + 4 + size])
        pos += 4 + size
    pos += 1
    assert attrs["compression"][1] == b"\0" and attrs["channels"][1][:2] == b"Y\0"
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    offs = struct.unpack("<%dQ" % h, b[pos:pos + 8 * h])
    out = np.empty((h, w), np.float32)
    for y in range(h):
        yy, nbytes = struct.unpack("<iI", b[offs[y]:offs[y] + 8])
        assert yy == y and nbytes == w * 4
        out[y] = np.frombuffer(b[offs[y] + 8:offs[y] + 8 + nbytes], np.float32)
    return out


@pytest.mark.gpu
def test_debug_images_exr_and_foreground_masks(tmp_path, cuda):
    """a20: --save_debug_images writes what PyramidLevel::saveDebugImages writes (8-bit cost / confidence PNGs, the red
    mismatch overlay, the 16-bit disparity preview), --output_formats=exr writes a readable EXR; §8(f): the
    GenerateForegroundMasks app against the same sequence of cv2 calls the reference's code makes."""
    W = H = 64
    S = 4
    rig = synth.ring_rig(S, W, H, kind="FTHETA")
    colors, _ = synth.render_rig(rig, W, H, scene=synth.Scene(seed=3))
    inp, out = str(tmp_path / "in"), str(tmp_path / "out")
    write_dataset(inp, rig, [colors], 2)
    run("DerpCLI", "--input_root=" + inp, "--output_root=" + out, "--partial_coverage=true", "--num_depths=32",
        "--output_formats=png,exr", "--save_debug_images=true", "--mismatches_start_level=0")
    disp = read_pfm(os.path.join(out, "disparity_levels", "level_0", "cam1", "000000.pfm"))
    exr = _read_exr_y(os.path.join(out, "disparity_levels", "level_0", "cam1", "000000.exr"))
    assert np.array_equal(exr.view(np.uint32), disp.view(np.uint32))
    # the same level through the binding, for cost / confidence / mismatch mask
    ctx = capi.Context(cuda, capi.rig_descs(rig))
    ctx.level_begin(W // 2, H // 2, level=1, num_levels=2, full_width=W, full_height=H)
    ctx.set_colors([synth.downscale_area(c, 2) for c in colors])
    ctx.process_level(num_depths=32, mismatches_start_level=0)
    c1 = [ctx.get_disparity(d, want_cost=False) for d in range(S)]
    ctx.level_begin(W, H, level=0, num_levels=2, full_width=W, full_height=H)
    ctx.set_colors(colors)
    for d in range(S):
        ctx.upsample_from(d, c1[d])
    ctx.process_level(num_depths=32, mismatches_start_level=0)
    d1, cost, conf = ctx.get_disparity(1)
    assert np.array_equal(d1.view(np.uint32), disp.view(np.uint32))

    def u8(v):  # saturate_cast<uchar>(cvRound(v)); cvRound answers INT_MIN outside the int range (NaN, inf, huge) -> 0
        with np.errstate(invalid="ignore", over="ignore"):
            ok = np.isfinite(v) & (np.abs(v) < 2147483648.0)
            r = np.where(ok, np.clip(np.rint(np.where(ok, v, 0)), 0, 255), 0)
        return r.astype(np.uint8)

    png = cv2.imread(os.path.join(out, "cost", "level_0", "cam1", "000000.png"), cv2.IMREAD_UNCHANGED)
    assert png.dtype == np.uint8 and png.shape == (H, W)
    with np.errstate(over="ignore"):  # ping-pong leaves +inf / huge costs on skipped pixels
        scaled_cost = cost * np.float32(255.0 / 100.0)
    assert np.array_equal(png, u8(scaled_cost))
    png = cv2.imread(os.path.join(out, "confidence", "level_0", "cam1", "000000.png"), cv2.IMREAD_UNCHANGED)
    assert np.array_equal(png, u8(conf * np.float32(255.0 * 100.0)))
    ov = cv2.imread(os.path.join(out, "mismatches", "level_0", "cam1", "000000.png"), cv2.IMREAD_UNCHANGED)
    assert ov.dtype == np.uint8 and ov.shape == (H, W, 4)
    mm, fov = ctx.get_mismatch_mask(1).astype(bool), ctx.get_fov_mask(1).astype(bool)
    assert (ov[fov & mm] == np.array([0, 0, 255, 255], np.uint8)).all()
    assert (ov[~fov] == 0).all()
    keep = fov & ~mm
    assert np.array_equal(ov[keep][:, 0], u8(d1 * np.float32(255.0))[keep]) and (ov[keep][:, 3] == 255).all()
    prev = cv2.imread(os.path.join(out, "disparity_levels", "level_0", "cam1", "000000.png"), cv2.IMREAD_UNCHANGED)
    assert prev.dtype == np.uint16
    ctx.close()

    # ---- GenerateForegroundMasks
    rng = np.random.RandomState(2)
    bgdir, fgdir, mdir = str(tmp_path / "bg"), str(tmp_path / "fg"), str(tmp_path / "masks")
    Wf, Hf = 120, 90
    for s in range(S):
        bgimg = np.clip(rng.normal(30000, 9000, (Hf, Wf, 3)), 0, 65535).astype(np.uint16)
        fr = bgimg.copy()
        fr[20:60, 30 + 5 * s:80] = rng.randint(0, 65536, (40, 50 - 5 * s, 3)).astype(np.uint16)
        os.makedirs(os.path.join(bgdir, "cam%d" % s))
        os.makedirs(os.path.join(fgdir, "cam%d" % s))
        cv2.imwrite(os.path.join(bgdir, "cam%d" % s, "000000.png"), bgimg)
        cv2.imwrite(os.path.join(fgdir, "cam%d" % s, "000007.png"), fr)
    run("GenerateForegroundMasks", "--rig=" + inp + "/rigs/rig_calibrated.json", "--color=" + fgdir, "--background_color=" + bgdir,
        "--foreground_masks=" + mdir, "--first=000007", "--last=000007", "--width=80")
    Wo, Ho = 80, int(np.rint(80 * Hf / np.float32(Wf)))
    a32 = np.float32(1.0) / np.float32(65535.0)
    for s in range(S):
        b = cv2.resize(cv2.imread(os.path.join(bgdir, "cam%d" % s, "000000.png"), cv2.IMREAD_UNCHANGED), (Wo, Ho), interpolation=cv2.INTER_AREA)
        f = cv2.resize(cv2.imread(os.path.join(fgdir, "cam%d" % s, "000007.png"), cv2.IMREAD_UNCHANGED), (Wo, Ho), interpolation=cv2.INTER_AREA)
        diff = cv2.absdiff(cv2.GaussianBlur(b, (3, 3), 0).astype(np.float32) * a32, cv2.GaussianBlur(f, (3, 3), 0).astype(np.float32) * a32)
        m = (np.sqrt((diff.astype(np.float64) ** 2).sum(-1)) > np.float64(np.float32(0.04))).astype(np.uint8)
        m = cv2.morphologyEx(m, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (4, 4)))
        got = cv2.imread(os.path.join(mdir, "cam%d" % s, "000007.png"), cv2.IMREAD_UNCHANGED)
        assert got.dtype == np.uint8 and np.array_equal(got, m * 255), s


def _mesh_dataset(tmp_path, W=96, H=80, S=3, F=2):
    rig = synth.ring_rig(S, W, H, kind="FTHETA")
    os.makedirs(tmp_path / "rigs", exist_ok=True)
    json.dump(rig, open(tmp_path / "rigs" / "rig.json", "w"))
    rng = np.random.RandomState(4)
    disps = {}
    for cam in rig["cameras"]:
        for f in range(F):
            yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
            d = (0.3 + 0.1 * np.sin(xx / 9.0 + f) * np.cos(yy / 7.0)).astype(np.float32)
            d[(xx + yy) % 31 < 9] *= 1.5
            d[rng.uniform(size=d.shape) < 0.03] = np.nan
            write_pfm(str(tmp_path / "disparity" / cam["id"] / ("%06d.pfm" % f)), d)
            disps[cam["id"], f] = d
    return rig, disps


def test_convert_to_binary_refuses_unbuilt_parts(tmp_path):
    """Bad command lines stop like the reference's (no GPU needed to get that far): unknown formats, --color_scale > 1.  Without a GPU the BC7 colour
    format stops with the CUDA error — there is no host encoder behind the executable."""
    rig, _ = _mesh_dataset(tmp_path, F=1)
    base = ["--rig=" + str(tmp_path / "rigs" / "rig.json"), "--first=000000", "--last=000000",
            "--disparity=" + str(tmp_path / "disparity"), "--bin=" + str(tmp_path / "bin")]
    p = run("ConvertToBinary", *base, "--output_formats=idx,vtx,exr", check=False)
    assert p.returncode != 0 and "Invalid output format" in p.stderr
    H, W = 80, 96
    for cam in rig["cameras"]:
        os.makedirs(tmp_path / "color" / cam["id"], exist_ok=True)
        assert cv2.imwrite(str(tmp_path / "color" / cam["id"] / "000000.png"), np.full((H, W, 3), 1000, np.uint16))
    p = run("ConvertToBinary", *base, "--color=" + str(tmp_path / "color"), "--color_scale=1.5", check=False)
    assert p.returncode != 0  # CHECK_LE(FLAGS_color_scale, 1) like the reference (ConvertToBinary.cpp:347)
    import torch
    if not torch.cuda.is_available():
        p = run("ConvertToBinary", base[0], base[1], base[2], base[4], "--color=" + str(tmp_path / "color"),
                "--output_formats=bc7", check=False)
        assert p.returncode != 0 and "cuda" in p.stderr.lower() and not os.path.exists(tmp_path / "bin" / "cam0" / "000000.bc7")


@pytest.mark.gpu
def test_convert_to_binary_bc7(tmp_path, cuda):
    """The reference's DEFAULT --output_formats (idx, vtx, bc7) with a colour directory: the .bc7 files are the library's
    blocks of the image as cv2.imread(IMREAD_UNCHANGED) holds it (16-bit BGR, 8-bit BGRA and gray PNGs), the fused stream
    carries them; against the reference's own bc7_util::compressBC7 the files agree as far as IEEE vs estimate arithmetic
    allows (tests/test_bc7.py)."""
    from tests import bc7_decode, oracle_libs
    rig, disps = _mesh_dataset(tmp_path, F=1)
    H, W = next(iter(disps.values())).shape
    rng = np.random.RandomState(6)
    yy, xx = np.mgrid[0:H, 0:W]
    base = np.stack([0.5 + 0.4 * np.sin(xx / 9.), 0.5 + 0.4 * np.cos(yy / 7.), (xx + yy) / float(W + H)], -1)
    imgs = {}
    for n, cam in enumerate(rig["cameras"]):
        img = (base + rng.normal(0, 0.01, base.shape)).clip(0, 1)
        if n == 0:
            img = (img * 65535).astype(np.uint16)
        elif n == 1:
            img = np.concatenate([(img * 255).astype(np.uint8), np.full((H, W, 1), 200, np.uint8)], -1)
        else:
            img = (img[..., 0] * 65535).astype(np.uint16)
        os.makedirs(tmp_path / "color" / cam["id"], exist_ok=True)
        assert cv2.imwrite(str(tmp_path / "color" / cam["id"] / "000000.png"), img)
        imgs[cam["id"]] = img
    run("ConvertToBinary", "--rig=" + str(tmp_path / "rigs" / "rig.json"), "--first=000000", "--last=000000",
        "--disparity=" + str(tmp_path / "disparity"), "--color=" + str(tmp_path / "color"), "--bin=" + str(tmp_path / "bin"),
        "--fused=" + str(tmp_path / "fused"))
    catalog = json.load(open(tmp_path / "fused" / "fused.json"))
    disk = open(tmp_path / "fused" / "fused_0.bin", "rb").read()
    ref = oracle_libs.load_ref()
    for cam in rig["cameras"]:
        img = imgs[cam["id"]]
        stored = img if img.ndim == 3 else np.repeat(img[..., None], 3, -1)  # GRAY2BGRA replicates the channel
        got = np.fromfile(str(tmp_path / "bin" / cam["id"] / "000000.bc7"), np.uint8)
        assert got.size == W * H
        assert np.array_equal(got, cuda.bc7_compress_image(stored[..., :3].copy(), 2.2 / 1.8))
        e = catalog["frames"]["000000"][cam["id"]][".bc7"]
        assert disk[e["offset"]:e["offset"] + e["size"]] == got.tobytes()
        assert os.path.exists(tmp_path / "bin" / cam["id"] / "000000.vtx")
        if ref is not None:
            want = ref.bc7_compress_image(stored[..., :3].copy(), 2.2 / 1.8)
            assert (got.reshape(-1, 16) == want.reshape(-1, 16)).all(1).mean() > 0.8  # 96.9-100 % measured; CPU dependent
            assert {bc7_decode.block_mode(b) for b in got.reshape(-1, 16)} <= {1, 3, 6}


@pytest.mark.gpu
def test_convert_to_binary_meshes(tmp_path, cuda, oracle):
    """Files in -> files out: .vtx / .idx equal the checker's mesh of the same PFMs byte for byte (--depth_scale and a
    foreground mask included), the .obj lists the same mesh, the fused stream holds every file at the catalogued offset
    with 0x5A padding to the 512 KiB stripe, and the fused rig is written."""
    from tests import oracle_libs
    checker = oracle_libs.load_ref() or oracle
    rig, disps = _mesh_dataset(tmp_path)
    H, W = next(iter(disps.values())).shape
    rng = np.random.RandomState(1)
    masks = {}
    for cam in rig["cameras"]:
        for f in range(2):
            m = (rng.uniform(size=(H // 2, W // 2)) > 0.15).astype(np.uint8)
            d = tmp_path / "masks" / cam["id"]
            os.makedirs(d, exist_ok=True)
            assert cv2.imwrite(str(d / ("%06d.png" % f)), m * 255)
            masks[cam["id"], f] = m
    run("ConvertToBinary", "--rig=" + str(tmp_path / "rigs" / "rig.json"), "--first=000000", "--last=000001",
        "--disparity=" + str(tmp_path / "disparity"), "--foreground_masks=" + str(tmp_path / "masks"),
        "--bin=" + str(tmp_path / "bin"), "--fused=" + str(tmp_path / "fused"), "--fuse_strip=2", "--triangles=0",
        "--depth_scale=0.5", "--output_formats=idx,vtx,obj")
    catalog = json.load(open(tmp_path / "fused" / "fused.json"))
    assert catalog["metadata"]["isLittleEndian"] is True
    disks = [open(tmp_path / "fused" / ("fused_%d.bin" % i), "rb").read() for i in range(2)]
    stripe = 512 * 1024

    def fused_bytes(offset, size):  # StripedFile.h:96-101
        out = b""
        while size:
            s = offset // stripe
            local = (s // 2) * stripe + offset % stripe
            n = min(size, stripe - offset % stripe)
            out += disks[s % 2][local:local + n]
            offset += n
            size -= n
        return out

    for cam in rig["cameras"]:
        res, focal = cam["resolution"], cam["focal"][0]
        for f in range(2):
            v, i = checker.camera_mesh(disps[cam["id"], f], res, focal, depth_scale=0.5, tear_ratio=0.95,
                                       foreground_mask=masks[cam["id"], f])
            stem = tmp_path / "bin" / cam["id"] / ("%06d" % f)
            assert open(str(stem) + ".vtx", "rb").read() == v.tobytes()
            assert open(str(stem) + ".idx", "rb").read() == i.tobytes()
            obj = open(str(stem) + ".obj").read().splitlines()
            assert sum(l.startswith("v ") for l in obj) == len(v) and sum(l.startswith("f ") for l in obj) == len(i)
            entry = catalog["frames"]["%06d" % f][cam["id"]]
            assert entry["offset"] % stripe == 0
            for ext in (".idx", ".vtx", ".obj"):
                e = entry[ext]
                assert fused_bytes(e["offset"], e["size"]) == open(str(stem) + ext, "rb").read()
            end = entry["offset"] + entry["size"]
            pad = fused_bytes(end, -end % stripe)
            assert pad == b"\x5a" * len(pad)
    # default --triangles (150 000 > these meshes: untouched) and a real target: files equal the reference's sequence
    if checker.backend == "reference-cpu":
        run("ConvertToBinary", "--rig=" + str(tmp_path / "rigs" / "rig.json"), "--first=000000", "--last=000000",
            "--disparity=" + str(tmp_path / "disparity"), "--bin=" + str(tmp_path / "bin2"), "--triangles=3000",
            "--output_formats=idx,vtx")
        for cam in rig["cameras"]:
            v, i = checker.camera_mesh(disps[cam["id"], 0], cam["resolution"], cam["focal"][0], triangles=3000)
            stem = tmp_path / "bin2" / cam["id"] / "000000"
            assert len(i) <= 3000 or len(i) < 2 * (W - 1) * (H - 1)
            assert open(str(stem) + ".vtx", "rb").read() == v.tobytes()
            assert open(str(stem) + ".idx", "rb").read() == i.tobytes()
    fused_rig = json.load(open(tmp_path / "fused" / "rig_fused.json"))
    assert [c["id"] for c in fused_rig["cameras"]] == [c["id"] for c in rig["cameras"]]
    assert np.allclose(fused_rig["cameras"][0]["focal"], rig["cameras"][0]["focal"])


def test_png_inflate_equals_zlib(tmp_path):
    """The PNG reader's own inflate (csrc/host/inflate.h) against zlib: stored, fixed and dynamic blocks from every
    compression level and strategy, two window sizes, empty and one-byte inputs, incompressible and highly repetitive data,
    matches at the maximum distance, flush points inside the stream; malformed streams (truncated, corrupted, wrong size) are
    declined (the reader then lets zlib report the error) and nothing is written past the output."""
    import zlib
    rng = np.random.RandomState(0)
    yy, xx = np.mgrid[0:256, 0:256]
    data = {
        "empty": b"", "one": b"x", "zeros": bytes(200000), "random": rng.bytes(100000),
        "text": b"the quick brown fox jumps over the lazy dog " * 3000,
        "lowent": bytes(rng.randint(0, 4, 150000).astype(np.uint8)),
        "u16smooth": ((np.sin(xx / 30.) * np.cos(yy / 17.) * 20000 + 30000) + rng.normal(0, 50, xx.shape)).astype(">u2").tobytes(),
        "maxdist": rng.bytes(32768) * 4,
        "short_runs": b"".join(bytes([i % 251]) * ((i % 7) + 1) for i in range(30000)),
    }
    z_path = str(tmp_path / "z.bin")

    def verdict(z, n):
        open(z_path, "wb").write(z)
        return run("IoSelfTest", "--in=" + z_path, "--mode=inflate", "--size=%d" % n).stdout.strip().splitlines()[-1].split()[0]

    for name, d in data.items():
        for level in (0, 1, 9):
            for k, strategy in enumerate((zlib.Z_DEFAULT_STRATEGY, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED)):
                wbits = 15 if (k + level) % 2 else 9
                c = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
                assert verdict(c.compress(d) + c.flush(), len(d)) == "same", (name, level, strategy, wbits)
    c = zlib.compressobj(6)
    d = data["random"] + data["zeros"]
    z = c.compress(d[:30000]) + c.flush(zlib.Z_SYNC_FLUSH) + c.compress(d[30000:]) + c.flush(zlib.Z_FULL_FLUSH) + c.flush()
    assert verdict(z, len(d)) == "same"
    z = zlib.compress(b"hello world" * 1000)
    assert verdict(z[:-9], 11000) == "declined" and verdict(z, 10999) == "declined"
    corrupt = bytearray(z)
    corrupt[20] ^= 0x55
    assert verdict(bytes(corrupt), 11000) in ("declined", "same")  # a flipped bit may still decode, but then the checksum fails


def test_rgba_stream_matches_opencv(tmp_path):
    """The ".rgba" stream of ConvertToBinary (convertColor, ConvertToBinary.cpp:138-146): loadImage<Vec4b> = convertTo 8U
    with the float scale 255/65535, BGR -> BGRA (alpha 255), then BGRA -> RGBA; the host loader against the same cv2 calls
    on 16-bit and 8-bit, 3- and 4-channel and gray PNGs."""
    rng = np.random.RandomState(3)
    cases = {
        "c16": rng.randint(0, 65536, (9, 13, 3)).astype(np.uint16),
        "c8": rng.randint(0, 256, (9, 13, 3)).astype(np.uint8),
        "a16": rng.randint(0, 65536, (7, 5, 4)).astype(np.uint16),
        "g16": rng.randint(0, 65536, (6, 8)).astype(np.uint16),
    }
    cases["c16"][0, :4] = [[0, 128, 65535], [127, 129, 386], [65407, 65408, 65409], [32767, 32768, 32896]]  # rounding edges
    for name, img in cases.items():
        p = str(tmp_path / (name + ".png"))
        assert cv2.imwrite(p, img)
        out = str(tmp_path / (name + ".rgba"))
        r = run("IoSelfTest", "--in=" + p, "--mode=rgba", "--out=" + out)
        w, h = map(int, r.stdout.split()[-2:])
        got = np.fromfile(out, np.uint8).reshape(h, w, 4)
        src = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        if src.dtype == np.uint16:  # cv::Mat::convertTo(CV_8U, 255.0f / 65535.0f)
            src = cv2.convertScaleAbs(src, alpha=float(np.float32(255.0) / np.float32(65535.0)))
        code = {2: cv2.COLOR_GRAY2BGRA, 3: cv2.COLOR_BGR2BGRA}.get(src.ndim if src.ndim == 2 else src.shape[2])
        bgra = src if code is None else cv2.cvtColor(src, code)
        assert np.array_equal(got, cv2.cvtColor(bgra, cv2.COLOR_BGRA2RGBA)), name


def test_gaussian_blur_radii_match_opencv(tmp_path):
    """cv_util::gaussianBlur for GenerateForegroundMasks --blur_radius 2 and 3 (the host stage; radius 1 is the library's 3 x 3
    kernel, checked on the GPU): cv2.GaussianBlur((2 r + 1)^2, sigma 0) on 16-bit 3-channel images, bit for bit, down to images
    smaller than the kernel."""
    rng = np.random.RandomState(1)
    src, out = str(tmp_path / "g.raw"), str(tmp_path / "g.out")
    for (w, h) in ((53, 37), (8, 5), (3, 3), (64, 1), (1, 9)):
        img = rng.randint(0, 65536, (h, w, 3)).astype(np.uint16)
        img.tofile(src)
        for r in (1, 2, 3):
            run("IoSelfTest", "--mode=gauss", "--in=" + src, "--width=%d" % w, "--height=%d" % h, "--size=%d" % r, "--out=" + out)
            assert np.array_equal(np.fromfile(out, np.uint16).reshape(h, w, 3), cv2.GaussianBlur(img, (2 * r + 1, 2 * r + 1), 0)), (w, h, r)
    p = run("IoSelfTest", "--mode=gauss", "--in=" + src, "--width=1", "--height=9", "--size=4", "--out=" + out, check=False)
    assert p.returncode != 0 and "blur_radius" in p.stderr  # sizes above 7 use a computed kernel: not restated, refused


def test_area_resize_matches_opencv(tmp_path):
    """cv::resize INTER_AREA of the colour streams under --color_scale < 1 (csrc/host/area_resize.h) against cv2: float and
    8-bit images with 4 and 3 channels; ratios of exactly 2 (OpenCV's vector bodies), other integer ratios (row-order sums,
    in groups of four for float), mixed and fractional ratios (separable taps); bit for bit."""
    rng = np.random.RandomState(0)
    src, out = str(tmp_path / "in.raw"), str(tmp_path / "out.raw")

    def resize(img, dw, dh, simd4=False):
        h, w, c = img.shape
        img.tofile(src)
        run("IoSelfTest", "--mode=area", "--in=" + src, "--width=%d" % w, "--height=%d" % h, "--channels=%d" % c,
            "--dst_width=%d" % dw, "--dst_height=%d" % dh, "--type=" + ("f32" if img.dtype == np.float32 else "u8"), "--out=" + out,
            *(["--simd4"] if simd4 else []))
        return np.fromfile(out, img.dtype).reshape(dh, dw, c)

    for (w, h, dw, dh) in [(64, 48, 32, 24), (66, 50, 33, 25), (60, 45, 20, 15), (64, 48, 16, 12), (70, 50, 35, 10),
                           (101, 77, 40, 31), (96, 64, 95, 63), (50, 40, 7, 3), (30, 20, 30, 10)]:
        for dtype, c in ((np.float32, 4), (np.float32, 3), (np.uint8, 4), (np.uint8, 3)):
            img = rng.uniform(0, 1, (h, w, c)).astype(np.float32) if dtype == np.float32 else rng.randint(0, 256, (h, w, c)).astype(np.uint8)
            want = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA).reshape(dh, dw, c)
            assert resize(img, dw, dh).tobytes() == want.tobytes(), (w, h, dw, dh, dtype.__name__, c)
    # an axis that grows (UpsampleDisparity's colour guide smaller than the output): OpenCV's bilinear variant, float images
    for (w, h, dw, dh) in [(20, 15, 40, 30), (20, 15, 33, 21), (17, 13, 50, 29), (20, 15, 40, 10), (20, 15, 10, 30), (31, 7, 32, 8),
                           (16, 16, 17, 16), (5, 4, 64, 48)]:
        img = rng.uniform(0, 1, (h, w, 3)).astype(np.float32)
        want = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA)
        assert resize(img, dw, dh).tobytes() == want.tobytes(), (w, h, dw, dh)
    # the BC7 stream resizes B, G, R with the arithmetic of the 4-channel image the reference carries
    img4 = rng.uniform(0, 1, (48, 64, 4)).astype(np.float32)
    want = cv2.resize(img4, (32, 24), interpolation=cv2.INTER_AREA)[..., :3]
    assert resize(np.ascontiguousarray(img4[..., :3]), 32, 24, simd4=True).tobytes() == np.ascontiguousarray(want).tobytes()


@pytest.mark.parametrize("scale", [0.5, 0.375, 1.0 / 3.0, 0.73])
def test_colour_streams_with_color_scale(tmp_path, scale):
    """--color_scale < 1 (scripts/ui/export.py:311-318 sets it for exports below full width): the RGBA8 surface handed to the
    BC7 encoder = the reference's own sequence — cv_util::convertTo to [0, 1] floats, BGR -> BGRA, cv::resize INTER_AREA
    (cv2), bc7_util::gammaCorrect (the reference's function, oracle/_ref) — and the .rgba stream = convertTo 8 bit, BGRA,
    INTER_AREA, RGBA; 16-bit colour, 8-bit colour with alpha, 16-bit grey."""
    import ctypes as C
    from tests import oracle_libs
    ref = oracle_libs.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    g = ref.lib.derp_ref_gamma_correct
    g.restype, g.argtypes = C.c_uint8, [C.c_float, C.c_float]
    gamma = 2.2 / 1.8
    rng = np.random.RandomState(5)
    H, W = 48, 64
    cases = {"c16": rng.randint(0, 65536, (H, W, 3)).astype(np.uint16), "a8": rng.randint(0, 256, (H, W, 4)).astype(np.uint8),
             "g16": rng.randint(0, 65536, (H, W)).astype(np.uint16)}
    for name, img in cases.items():
        png = str(tmp_path / (name + ".png"))
        assert cv2.imwrite(png, img)
        stored = cv2.imread(png, cv2.IMREAD_UNCHANGED)
        dw, dh = int(np.floor(W * scale + 0.5)), int(np.floor(H * scale + 0.5))  # std::round
        # ---- BC7 surface
        top = np.float32(65535.0 if stored.dtype == np.uint16 else 255.0)
        f = stored.astype(np.float32) * (np.float32(1.0) / top)
        bgra = cv2.cvtColor(f, {2: cv2.COLOR_GRAY2BGRA, 3: cv2.COLOR_BGR2BGRA}.get(f.ndim if f.ndim == 2 else f.shape[2])) \
            if (f.ndim == 2 or f.shape[2] == 3) else f
        small = cv2.resize(bgra, (dw, dh), interpolation=cv2.INTER_AREA)
        want = np.zeros((dh, dw, 4), np.uint8)
        want[..., 3] = 255
        for c in range(3):
            want[..., c] = np.array([g(float(v), gamma) for v in small[..., 2 - c].ravel()], np.uint8).reshape(dh, dw)
        out = str(tmp_path / (name + ".surface"))
        r = run("IoSelfTest", "--mode=bc7surface", "--in=" + png, "--scale=%r" % scale, "--out=" + out)
        assert list(map(int, r.stdout.split()[-2:])) == [dw, dh]
        assert np.fromfile(out, np.uint8).tobytes() == want.tobytes(), name
        # ---- .rgba stream
        s8 = stored if stored.dtype == np.uint8 else cv2.convertScaleAbs(stored, alpha=float(np.float32(255.0) / np.float32(65535.0)))
        code = {2: cv2.COLOR_GRAY2BGRA, 3: cv2.COLOR_BGR2BGRA}.get(s8.ndim if s8.ndim == 2 else s8.shape[2])
        bgra8 = s8 if code is None else cv2.cvtColor(s8, code)
        want8 = cv2.cvtColor(cv2.resize(bgra8, (dw, dh), interpolation=cv2.INTER_AREA), cv2.COLOR_BGRA2RGBA)
        out = str(tmp_path / (name + ".rgba"))
        run("IoSelfTest", "--mode=rgba", "--in=" + png, "--scale=%r" % scale, "--out=" + out)
        assert np.fromfile(out, np.uint8).tobytes() == want8.tobytes(), name


def test_convert_to_binary_rgba_only(tmp_path):
    """ConvertToBinary with a colour directory and --output_formats=rgba touches no GPU stage: .rgba files = the cv2
    sequence, the rig is rescaled to the colour resolution (resizeRig, ConvertToBinary.cpp:322-343), the fused stream holds
    the files."""
    W, H = 48, 40
    rig = synth.ring_rig(2, 2 * W, 2 * H, kind="FTHETA")  # rig at twice the colour resolution -> rescale by 0.5
    os.makedirs(tmp_path / "rigs", exist_ok=True)
    json.dump(rig, open(tmp_path / "rigs" / "rig.json", "w"))
    rng = np.random.RandomState(8)
    imgs = {}
    for cam in rig["cameras"]:
        d = tmp_path / "color" / cam["id"]
        os.makedirs(d, exist_ok=True)
        imgs[cam["id"]] = rng.randint(0, 65536, (H, W, 3)).astype(np.uint16)
        assert cv2.imwrite(str(d / "000000.png"), imgs[cam["id"]])
    run("ConvertToBinary", "--rig=" + str(tmp_path / "rigs" / "rig.json"), "--first=000000", "--last=000000",
        "--color=" + str(tmp_path / "color"), "--bin=" + str(tmp_path / "bin"), "--fused=" + str(tmp_path / "fused"),
        "--output_formats=rgba")
    for cam in rig["cameras"]:
        got = np.fromfile(tmp_path / "bin" / cam["id"] / "000000.rgba", np.uint8).reshape(H, W, 4)
        src = cv2.convertScaleAbs(imgs[cam["id"]], alpha=float(np.float32(255.0) / np.float32(65535.0)))
        assert np.array_equal(got, cv2.cvtColor(cv2.cvtColor(src, cv2.COLOR_BGR2BGRA), cv2.COLOR_BGRA2RGBA))
    fused_rig = json.load(open(tmp_path / "fused" / "rig_fused.json"))
    c0, r0 = fused_rig["cameras"][0], rig["cameras"][0]
    assert c0["resolution"] == [W, H] and np.allclose(c0["focal"], np.array(r0["focal"]) * 0.5)
    assert np.allclose(c0.get("principal", [W / 2, H / 2]), [W / 2, H / 2])
    catalog = json.load(open(tmp_path / "fused" / "fused.json"))
    e = catalog["frames"]["000000"][rig["cameras"][1]["id"]][".rgba"]
    assert e["size"] == W * H * 4 and e["offset"] % (512 * 1024) == 0
    # --color_scale = 0.5 (scripts/ui/export.py:311-318): the stream is the INTER_AREA-shrunk 8-bit BGRA image, the rig follows
    run("ConvertToBinary", "--rig=" + str(tmp_path / "rigs" / "rig.json"), "--first=000000", "--last=000000",
        "--color=" + str(tmp_path / "color"), "--bin=" + str(tmp_path / "bin_half"), "--output_formats=rgba", "--color_scale=0.5")
    for cam in rig["cameras"]:
        got = np.fromfile(tmp_path / "bin_half" / cam["id"] / "000000.rgba", np.uint8).reshape(H // 2, W // 2, 4)
        src = cv2.cvtColor(cv2.convertScaleAbs(imgs[cam["id"]], alpha=float(np.float32(255.0) / np.float32(65535.0))), cv2.COLOR_BGR2BGRA)
        assert np.array_equal(got, cv2.cvtColor(cv2.resize(src, (W // 2, H // 2), interpolation=cv2.INTER_AREA), cv2.COLOR_BGRA2RGBA))
    half = json.load(open(tmp_path / "bin_half" / "rig_fused.json"))["cameras"][0]
    assert half["resolution"] == [W // 2, H // 2] and np.allclose(half["focal"], np.array(r0["focal"]) * 0.25)
