"""Camera mesh of a disparity map (first slice: the geometry half of ConvertToBinary's convertDepth
before simplification — ConvertToBinary.cpp:150-183, MeshUtil.h).  CPU: the oracle restatement against the reference's own
MeshUtil.h (oracle/_ref) and hand-checked small cases.  GPU: the CUDA library against the checker, exactly (index work and
IEEE fp64 divisions only — no tolerance)."""
import os

import numpy as np
import pytest

from tests import oracle_libs


def disparity_case(rng, w, h, nan_frac=0.02, zero_frac=0.005, steps=True):
    """Smooth disparity with depth discontinuities (tears), NaN holes and a few zeros (infinite depth)."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    d = 0.4 + 0.2 * np.sin(xx / 17.0) * np.cos(yy / 11.0)
    if steps:
        d[(xx + 2 * yy) % 61 < 20] *= 1.6  # discontinuities well beyond the 0.95 tear ratio
        d += rng.uniform(-0.012, 0.012, d.shape)  # and many ratios close to it
    d = d.astype(np.float32)
    d[rng.uniform(size=d.shape) < nan_frac] = np.nan
    d[rng.uniform(size=d.shape) < zero_frac] = 0.0
    return d


def check_mesh_invariants(vtx, idx, w, h):
    assert vtx.dtype == np.float32 and idx.dtype == np.uint32
    if len(idx):
        assert idx.max() < len(vtx)
        # every vertex is referenced (applyMaskToVertexesAndFaces drops the others), in ascending first-use-free order
        assert np.array_equal(np.unique(idx), np.arange(len(vtx)))
    assert len(idx) <= 2 * (w - 1) * (h - 1)


CASES = [
    dict(w=37, h=23, scale=1.0, fg=False, tear=0.95),
    dict(w=64, h=48, scale=1.0, fg=True, tear=0.95),
    dict(w=96, h=80, scale=0.5, fg=False, tear=0.95),
    dict(w=101, h=67, scale=0.37, fg=True, tear=0.9),
    dict(w=40, h=40, scale=1.0, fg=False, tear=0.0),
    dict(w=2, h=2, scale=1.0, fg=False, tear=0.95),
    dict(w=1, h=5, scale=1.0, fg=False, tear=0.95),
]


def run_case(lib, case, seed=0):
    rng = np.random.RandomState(seed)
    d = disparity_case(rng, case["w"], case["h"])
    fg = None
    if case["fg"]:  # a mask at another resolution, values 0 / 1 / 255 / 254 (bit 0 decides, like Mat_<bool> & Mat_<bool>)
        fg = rng.choice(np.array([0, 1, 255, 254], np.uint8), size=(case["h"] // 2 + 3, case["w"] // 2 + 1), p=[0.1, 0.4, 0.4, 0.1])
    return lib.camera_mesh(d, (case["w"] * 4.0, case["h"] * 4.0), 317.25, depth_scale=case["scale"],
                           tear_ratio=case["tear"], foreground_mask=fg)


@pytest.mark.parametrize("case", CASES)
def test_oracle_equals_reference_mesh_util(oracle, case):
    ref = oracle_libs.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    for seed in range(3):
        ov, oi = run_case(oracle, case, seed)
        rv, ri = run_case(ref, case, seed)
        assert np.array_equal(oi, ri)
        assert np.array_equal(ov.view(np.uint32), rv.view(np.uint32))
        check_mesh_invariants(ov, oi, case["w"], case["h"])


def test_flat_quad_by_hand(oracle):
    """2 x 2 constant disparity: both triangles, split along the tl-br diagonal rule (|tl-br| < |tr-bl| is false ->
    triangles 0 and 3), vertexes = pixel centres scaled to the camera resolution, z = focal * disparity."""
    d = np.full((2, 2), 0.5, np.float32)
    v, f = oracle.camera_mesh(d, (8.0, 8.0), 100.0)
    assert np.array_equal(f, np.array([[2, 1, 0], [1, 2, 3]], np.uint32))
    assert np.allclose(v, [[2, 2, 50], [6, 2, 50], [2, 6, 50], [6, 6, 50]])
    # one NaN corner: only the triangle opposite to it survives the vertex mask, and the vertexes are re-indexed
    d[0, 0] = np.nan
    v, f = oracle.camera_mesh(d, (8.0, 8.0), 100.0)
    assert len(v) == 3 and len(f) <= 1
    # a tear: one corner much closer -> the triangle that avoids it
    d = np.array([[0.5, 0.5], [0.5, 5.0]], np.float32)
    v, f = oracle.camera_mesh(d, (8.0, 8.0), 100.0)
    assert np.array_equal(f, np.array([[2, 1, 0]], np.uint32)) and len(v) == 3


def test_bad_arguments(oracle):
    with pytest.raises(Exception):
        oracle.camera_mesh(np.ones((4, 4), np.float32), (4.0, 4.0), 1.0, depth_scale=1.5)  # CHECK_LE(depth_scale, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_mesh_equals_checker(cuda, oracle, case):
    ref = oracle_libs.load_ref()
    checker = ref if ref is not None else oracle
    for seed in range(3):
        gv, gi = run_case(cuda, case, seed)
        cv, ci = run_case(checker, case, seed)
        assert np.array_equal(gi, ci)
        assert np.array_equal(gv.view(np.uint32), cv.view(np.uint32))


@pytest.mark.gpu
def test_gpu_mesh_full_size(cuda, oracle):
    """2048 x 2048 (level-0 disparity of the headline rig) with holes, tears and a foreground mask: exact against the
    oracle, and the mesh invariants."""
    rng = np.random.RandomState(11)
    w = h = 2048
    d = disparity_case(rng, w, h)
    fg = (rng.uniform(size=(h // 4, w // 4)) > 0.2).astype(np.uint8)
    gv, gi = cuda.camera_mesh(d, (float(w), float(h)), 651.9, foreground_mask=fg)
    ov, oi = oracle.camera_mesh(d, (float(w), float(h)), 651.9, foreground_mask=fg)
    assert len(gi) > 1_000_000
    assert np.array_equal(gi, oi)
    assert np.array_equal(gv.view(np.uint32), ov.view(np.uint32))
    check_mesh_invariants(gv, gi, w, h)


# ---- mesh simplification (ConvertToBinary.cpp:186-203): product host code vs the reference's MeshSimplifier.cpp ------
def _simplify(lib, name, xyz, idx, triangles, strictness=0.2, remove_boundary=False):
    import ctypes as C
    f = getattr(lib.lib, name)
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p,
                  C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    xyz = np.ascontiguousarray(xyz, np.float64)
    idx = np.ascontiguousarray(idx, np.uint32)
    ov, oi = np.empty_like(xyz), np.empty_like(idx)
    nv, nf = C.c_uint64(), C.c_uint64()
    assert f(xyz.ctypes.data, len(xyz), idx.ctypes.data, len(idx), triangles, strictness, int(remove_boundary),
             ov.ctypes.data, oi.ctypes.data, C.byref(nv), C.byref(nf)) == 0
    return ov[:nv.value].copy(), oi[:nf.value].copy()


def _surface_mesh(oracle, seed, w, h, smooth):
    rng = np.random.RandomState(seed)
    if smooth:  # one sheet, interior edges contract freely
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        d = (0.4 + 0.05 * np.sin(xx / 13.0) * np.cos(yy / 9.0) + rng.uniform(-0.002, 0.002, (h, w))).astype(np.float32)
    else:  # tears and holes: many boundaries
        d = disparity_case(rng, w, h, nan_frac=0.01, zero_frac=0.0)
    v, i = oracle.camera_mesh(d, (w * 4.0, h * 4.0), 300.0)
    return v.astype(np.float64), i


SIMPLIFY_CASES = [(0, 64, 48, True, 2000, False), (1, 96, 72, True, 1500, True), (2, 80, 60, False, 3000, False),
                  (3, 120, 90, False, 100, True), (4, 33, 27, True, 10, False), (5, 150, 110, True, 6000, False),
                  # torn meshes that cannot reach the target: long runs of sweeps that contract nothing (the fast path of run())
                  (11, 160, 120, False, 50, False), (16, 250, 180, False, 10000, False), (15, 96, 200, False, 1, True)]


@pytest.mark.parametrize("seed,w,h,smooth,target,rb", SIMPLIFY_CASES)
def test_simplifier_equals_reference(oracle, seed, w, h, smooth, target, rb):
    """The library's host simplifier (derp_simplify.h, through its test hook — no GPU involved) against the reference's own
    MeshSimplifier.cpp compiled into oracle/_ref, on meshes with and without boundaries: same vertex bits, same faces."""
    from facebook360_dep_b200 import capi
    ref = oracle_libs.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    prod = capi.load_cuda()
    xyz, idx = _surface_mesh(oracle, seed, w, h, smooth)
    pv, pi = _simplify(prod, "derp_test_simplify", xyz, idx, target, remove_boundary=rb)
    rv, ri = _simplify(ref, "derp_ref_simplify", xyz, idx, target, remove_boundary=rb)
    assert len(ri) < len(idx)  # something was contracted
    assert np.array_equal(pi, ri)
    assert np.array_equal(pv.view(np.uint64), rv.view(np.uint64))


def test_simplifier_equals_reference_on_random_meshes(oracle):
    """96 more meshes against the reference's MeshSimplifier.cpp: random sizes, sheets and torn meshes, targets of 1 face, a
    few hundred, a few thousand and more than there are (nothing to do), boundary edges kept or removed — the combinations
    that decide how the sweep loop ends (target reached inside a sweep, exactly at a sweep's end, never)."""
    from facebook360_dep_b200 import capi
    ref = oracle_libs.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    prod = capi.load_cuda()
    for seed in range(20, 32):
        rng = np.random.RandomState(seed)
        w, h = int(rng.randint(20, 140)), int(rng.randint(20, 140))
        smooth = bool(rng.randint(0, 2))
        xyz, idx = _surface_mesh(oracle, seed, w, h, smooth)
        for target in (1, int(rng.randint(2, 400)), int(rng.randint(400, 20000)), 10 ** 7):
            for rb in (False, True):
                pv, pi = _simplify(prod, "derp_test_simplify", xyz, idx, target, remove_boundary=rb)
                rv, ri = _simplify(ref, "derp_ref_simplify", xyz, idx, target, remove_boundary=rb)
                assert np.array_equal(pi, ri), (seed, w, h, smooth, target, rb)
                assert np.array_equal(pv.view(np.uint64), rv.view(np.uint64)), (seed, w, h, smooth, target, rb)


def test_raster_pfm_equals_reference_write_pfm(oracle, tmp_path):
    """ConvertToBinary's "pfm" format (mesh_util::writePfm, MeshUtil.h:24-70): the host rasteriser of the apps (io.h
    rasterMesh, through IoSelfTest) against the reference's own function compiled into oracle/_ref, on a camera mesh as
    written to .vtx / .idx and on the same mesh simplified — bit for bit (both sides solve calcBarycentrics' 2 x 2 system by
    elimination: oracle/_ref has a stand-in for Eigen, see io.h) — and against the mesh itself: the vertexes of an
    unsimplified mesh are pixel centres, and those pixels show their z."""
    import ctypes as C
    import subprocess
    from facebook360_dep_b200 import capi
    ref = oracle_libs.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    w, h = 72, 56
    d = disparity_case(np.random.RandomState(9), w, h, nan_frac=0.02, zero_frac=0.0)
    res = (w * 4.0, h * 4.0)
    v32, idx = oracle.camera_mesh(d, res, 300.0)
    sv, si = _simplify(capi.load_cuda(), "derp_test_simplify", v32.astype(np.float64), idx, 1500)
    f = ref.lib.derp_ref_write_raster_pfm
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_char_p]
    exe = os.path.join(capi.ROOT, "facebook360_dep_b200", "bin", "IoSelfTest")
    for name, vv, ii in (("full", v32, idx), ("simplified", sv.astype(np.float32), si)):
        vv = np.ascontiguousarray(vv, np.float32)
        ii = np.ascontiguousarray(ii, np.uint32)
        vv.tofile(str(tmp_path / "m.vtx"))
        ii.tofile(str(tmp_path / "m.idx"))
        out = str(tmp_path / (name + ".pfm"))
        subprocess.run([exe, "--mode=raster", "--in=" + str(tmp_path / "m.vtx"), "--faces=" + str(tmp_path / "m.idx"),
                        "--width=%d" % w, "--height=%d" % h, "--resolution_x=%r" % res[0], "--resolution_y=%r" % res[1],
                        "--out=" + out], check=True, capture_output=True)
        v64 = vv.astype(np.float64)
        want = str(tmp_path / (name + "_ref.pfm"))
        assert f(v64.ctypes.data, len(v64), ii.ctypes.data, len(ii), w, h, res[0], res[1], want.encode()) == 0
        got_bytes, want_bytes = open(out, "rb").read(), open(want, "rb").read()
        assert got_bytes == want_bytes, name
        raster = np.frombuffer(got_bytes[-w * h * 4:], np.float32).reshape(h, w)
        covered = raster > -3e38
        assert covered.mean() > 0.5
        if name == "full":  # every vertex of the unsimplified mesh sits on a pixel centre: that pixel shows the vertex's z
            px = np.rint(v32[:, 0] / 4.0 - 0.5).astype(int)
            py = np.rint(v32[:, 1] / 4.0 - 0.5).astype(int)
            assert np.allclose(v32[:, 0] / 4.0 - 0.5, px, atol=1e-4) and np.allclose(v32[:, 1] / 4.0 - 0.5, py, atol=1e-4)
            seen = covered[py, px]
            assert seen.mean() > 0.95 and np.allclose(raster[py, px][seen], v32[seen, 2], rtol=2e-6)


@pytest.mark.gpu
def test_gpu_mesh_simplified_equals_reference(cuda):
    """derp_camera_mesh_simplified end to end (GPU mesh in double precision -> host contraction sweeps -> float32 / uint32)
    against convertDepth's sequence on the reference's own code."""
    ref = oracle_libs.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.RandomState(2)
    yy, xx = np.mgrid[0:150, 0:200].astype(np.float32)
    d = (0.4 + 0.05 * np.sin(xx / 13.0) * np.cos(yy / 9.0) + rng.uniform(-0.002, 0.002, xx.shape)).astype(np.float32)
    d[(xx > 90) & (xx < 100)] *= 1.4  # a tear
    d[rng.uniform(size=d.shape) < 0.002] = np.nan
    d[40:44, 50:60] = -0.3  # invalid negative disparities: z < 0 -> FLT_MIN whenever --triangles > 0
    for tri in (10 ** 6, 20000, 5000):  # the first target is above the mesh size: nothing to contract
        gv, gi = cuda.camera_mesh(d, (800.0, 600.0), 300.0, triangles=tri)
        rv, ri = ref.camera_mesh(d, (800.0, 600.0), 300.0, triangles=tri)
        assert np.array_equal(gi, ri)
        assert np.array_equal(gv.view(np.uint32), rv.view(np.uint32))


def _flat_grid(w, h):
    """An exactly planar w x h vertex grid (z = 1) split into 2 (w-1)(h-1) triangles: every quadric error is exactly 0."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    xyz = np.stack([xx.ravel(), yy.ravel(), np.ones(w * h)], 1)
    q = (np.arange(h - 1)[:, None] * w + np.arange(w - 1)[None, :]).ravel().astype(np.uint32)
    idx = np.concatenate([np.stack([q, q + 1, q + w], 1), np.stack([q + 1, q + w + 1, q + w], 1)])
    return xyz, idx


_PLANAR_CHILD = """
import sys, numpy as np
sys.path.insert(0, %r)
from tests import oracle_libs, test_mesh
from facebook360_dep_b200 import capi
lib, name = (capi.load_cuda(), "derp_test_simplify") if sys.argv[1] == "product" else (oracle_libs.load_ref(), "derp_ref_simplify")
xyz, idx = test_mesh._flat_grid(12, 10)
v, i = test_mesh._simplify(lib, name, xyz, idx, 4)
np.save(sys.argv[2], i)
"""


def test_simplifier_terminates_on_planar_mesh(tmp_path):
    """A constant-disparity region is exactly planar: every edge cost is 0, the interior contracts, and once only boundary
    edges are left the percentile threshold is exactly 0 and `threshold *= 2 * ++count` never grows -- the reference's loop
    (MeshSimplifier.cpp:483-493) never ends (shown here with its own code, in a child process with a time limit).  The
    library stops with the mesh it has (derp_simplify.h, documented deviation), and that mesh is a valid one."""
    import subprocess
    import sys
    root = oracle_libs.ROOT
    out = str(tmp_path / "faces.npy")
    subprocess.run([sys.executable, "-c", _PLANAR_CHILD % root, "product", out], check=True, timeout=60)
    faces = np.load(out)
    xyz, idx = _flat_grid(12, 10)
    assert 0 < len(faces) < len(idx)
    # boundary vertices are never contracted (removeBoundaryEdges = false): all 40 rim vertices survive
    assert len(np.unique(faces)) >= 2 * (12 + 10) - 4
    assert all(len(set(f)) == 3 for f in faces.tolist())
    if oracle_libs.load_ref() is not None:
        with pytest.raises(subprocess.TimeoutExpired):
            subprocess.run([sys.executable, "-c", _PLANAR_CHILD % root, "reference", out], check=True, timeout=8)
