"""Dry run of the GPU-only late tests on the CPU: the apps' GPU stages are replaced by their CPU equivalents (oracle mesh, host
instantiation of the BC7 encoder, cv2 for the masks) so that everything else in the tests executes."""
import sys, os, json, pathlib, tempfile, numpy as np, cv2
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tests.test_apps as TA
import tests.test_z_late_additions as Z
from tests import oracle_libs
from tests.test_bc7 import host_blocks
oracle = oracle_libs.load_oracle()
real_run = TA.run

class FakeCuda:
    def bc7_compress(self, rgba):
        return host_blocks(np.ascontiguousarray(rgba))

def fake_run(app, *args, check=True):
    kv = dict(a[2:].split("=", 1) for a in args if a.startswith("--") and "=" in a)
    if app == "ConvertToBinary" and "disparity" in kv:  # raster test: CPU mesh (not simplified) + the apps' rasteriser
        rig = json.load(open(kv["rig"]))
        for cam in rig["cameras"]:
            d = TA.read_pfm(os.path.join(kv["disparity"], cam["id"], "000000.pfm"))
            h, w = d.shape
            v, i = oracle.camera_mesh(d, cam["resolution"], cam["focal"][0])
            stem = os.path.join(kv["bin"], cam["id"], "000000")
            os.makedirs(os.path.dirname(stem), exist_ok=True)
            v.astype(np.float32).tofile(stem + ".vtx"); i.astype(np.uint32).tofile(stem + ".idx")
            real_run("IoSelfTest", "--mode=raster", "--in=" + stem + ".vtx", "--faces=" + stem + ".idx", "--width=%d" % w, "--height=%d" % h,
                     "--resolution_x=%r" % float(cam["resolution"][0]), "--resolution_y=%r" % float(cam["resolution"][1]), "--out=" + stem + ".pfm")
        return None
    if app == "ConvertToBinary" and kv.get("output_formats") == "bc7":
        rig = json.load(open(kv["rig"]))
        for cam in rig["cameras"]:
            surf = os.path.join(kv["bin"], cam["id"] + ".tmp")
            os.makedirs(os.path.join(kv["bin"], cam["id"]), exist_ok=True)
            r = real_run("IoSelfTest", "--mode=bc7surface", "--in=" + os.path.join(kv["color"], cam["id"], "000000.png"), "--scale=" + kv["color_scale"], "--out=" + surf)
            w, h = map(int, r.stdout.split()[-2:])
            host_blocks(np.fromfile(surf, np.uint8).reshape(h, w, 4)).tofile(os.path.join(kv["bin"], cam["id"], "000000.bc7"))
        return None
    if app == "GenerateForegroundMasks":
        rig = json.load(open(kv["rig"])); r = int(kv["blur_radius"]); k = 2 * r + 1
        for cam in rig["cameras"]:
            b = cv2.imread(os.path.join(kv["background_color"], cam["id"], "000000.png"), cv2.IMREAD_UNCHANGED)
            f = cv2.imread(os.path.join(kv["color"], cam["id"], "000007.png"), cv2.IMREAD_UNCHANGED)
            Wo = int(kv["width"]); Ho = int(np.rint(Wo * b.shape[0] / np.float32(b.shape[1])))
            b = cv2.resize(b, (Wo, Ho), interpolation=cv2.INTER_AREA); f = cv2.resize(f, (Wo, Ho), interpolation=cv2.INTER_AREA)
            a32 = np.float32(1.0) / np.float32(65535.0)
            diff = cv2.absdiff(cv2.GaussianBlur(b, (k, k), 0).astype(np.float32) * a32, cv2.GaussianBlur(f, (k, k), 0).astype(np.float32) * a32)
            m = (np.sqrt((diff.astype(np.float64) ** 2).sum(-1)) > np.float64(np.float32(0.04))).astype(np.uint8)
            m = cv2.morphologyEx(m, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (4, 4)))
            os.makedirs(os.path.join(kv["foreground_masks"], cam["id"]), exist_ok=True)
            cv2.imwrite(os.path.join(kv["foreground_masks"], cam["id"], "000007.png"), m * 255)
        return None
    if app == "UpsampleDisparity":
        from facebook360_dep_b200 import capi
        rig = json.load(open(kv["rig"])); R = int(kv["resolution"])
        for cam in rig["cameras"]:
            d = TA.read_pfm(os.path.join(kv["disparity"], cam["id"], "000000.pfm"))
            up = oracle.upsample_disparity(capi.camera_desc_from_json(cam), d, R, R)
            g = cv2.imread(os.path.join(kv["color"], cam["id"], "000000.png"), cv2.IMREAD_UNCHANGED).astype(np.float32) * (np.float32(1.0) / np.float32(65535.0))
            g = cv2.resize(g, (R, R), interpolation=cv2.INTER_AREA)
            out = oracle.joint_bilateral_f32(up, g, np.ones((R, R), np.uint8), int((R / d.shape[1]) ** 2 + 1), 0.05, 0.5, 0.5, 1.0)
            TA.write_pfm(os.path.join(kv["output"], cam["id"], "000000.pfm"), out)
        return None
    return real_run(app, *args, check=check)

TA.run = fake_run
cuda = FakeCuda()
for kind in Z.HARD_KINDS:
    Z.test_gpu_blocks_equal_host_instantiation_hard_surfaces.__wrapped__(cuda, kind) if hasattr(Z.test_gpu_blocks_equal_host_instantiation_hard_surfaces, "__wrapped__") else Z.test_gpu_blocks_equal_host_instantiation_hard_surfaces(cuda, kind)
print("hard surfaces: test body runs")
for name, fn, extra in (("raster pfm", Z.test_convert_to_binary_raster_pfm, ()), ("bc7 color_scale", Z.test_convert_to_binary_bc7_with_color_scale, ()),
                        ("masks r=2", Z.test_generate_foreground_masks_larger_blur, (2,)), ("masks r=3", Z.test_generate_foreground_masks_larger_blur, (3,)),
                        ("upsample, larger guide", Z.test_upsample_disparity_with_resized_guide, (oracle, 192)),
                        ("upsample, smaller guide", Z.test_upsample_disparity_with_resized_guide, (oracle, 64))):
    with tempfile.TemporaryDirectory() as d:
        fn(pathlib.Path(d), cuda, *extra)
    print(name + ": test body runs and its assertions hold with the CPU stand-ins")
