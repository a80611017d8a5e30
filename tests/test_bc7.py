"""BC7 colour of ConvertToBinary (colour half): the product's block encoder (derp_bc7.cuh) against the
reference's own encoder — kernel.ispc compiled by the ispc binary the reference vendors, with the reference's flags, into
oracle/_ref (tests infrastructure).

Three statements, from strongest to weakest:
  1. the encoder source restates the reference's algorithm exactly: built as plain C++ with the x86 estimate instructions the
     reference build uses for a / b and rsqrt (oracle/libbc7_x86.so), its blocks are BYTE-IDENTICAL to the reference's;
  2. the CUDA kernel runs the same source in IEEE arithmetic; the host instantiation of that source (test hook) is what the
     GPU must reproduce byte for byte (-m gpu);
  3. IEEE vs estimate arithmetic only moves decisions that hang on the last bit: most blocks are identical, the decoded error
     of the two encodings is the same within a few hundredths of a dB (decoder: tests/bc7_decode.py, from the format rules).
"""
import ctypes as C
import os

import numpy as np
import pytest

from tests import bc7_decode, oracle_libs

KINDS = ["smooth", "noise", "flat", "edges", "ramp"]
HARD_KINDS = ["twocolour", "extremes", "nearflat", "channel", "dark"]


def surface(seed, w, h, kind):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    if kind == "smooth":
        img = np.stack([128 + 100 * np.sin(xx / 17.) * np.cos(yy / 23.), 128 + 90 * np.sin(xx / 7. + yy / 11.),
                        (xx * yy / 1024.) % 256], -1) + rng.normal(0, 2, (h, w, 3))
    elif kind == "noise":
        img = rng.uniform(0, 255, (h, w, 3))
    elif kind == "flat":  # constant 4 x 4 blocks (zero-length lines), the upper half with one grey level of noise
        img = np.zeros((h, w, 3)) + rng.randint(0, 256, ((h + 3) // 4, (w + 3) // 4, 3)).repeat(4, 0).repeat(4, 1)[:h, :w]
        img[:h // 2] += rng.randint(0, 2, (h // 2, w, 3))
    elif kind == "edges":
        img = np.where(((xx // 5 + yy // 7) % 2)[..., None] > 0, rng.uniform(0, 255, 3), rng.uniform(0, 255, 3)) + \
            rng.normal(0, 3, (h, w, 3))
    elif kind == "twocolour":  # every block: two random colours split along a random line (what the partitions are for)
        by, bx = (h + 3) // 4, (w + 3) // 4
        c0 = rng.randint(0, 256, (by, bx, 3)).repeat(4, 0).repeat(4, 1)[:h, :w]
        c1 = rng.randint(0, 256, (by, bx, 3)).repeat(4, 0).repeat(4, 1)[:h, :w]
        ang = rng.uniform(0, np.pi, (by, bx)).repeat(4, 0).repeat(4, 1)[:h, :w]
        side = ((xx % 4) - 1.5) * np.cos(ang) + ((yy % 4) - 1.5) * np.sin(ang) > 0
        img = np.where(side[..., None], c0, c1).astype(np.float64)
    elif kind == "extremes":  # 0 / 255 per pixel and channel: clamps and saturated end points
        img = rng.randint(0, 2, (h, w, 3)) * 255.0
    elif kind == "nearflat":  # constant blocks with one pixel one grey level off: zero-length lines, degenerate covariances
        img = np.zeros((h, w, 3)) + rng.randint(0, 256, ((h + 3) // 4, (w + 3) // 4, 3)).repeat(4, 0).repeat(4, 1)[:h, :w]
        img[1::4, 2::4, 1] += 1
    elif kind == "channel":  # only red varies
        img = np.stack([rng.uniform(0, 255, (h, w)), np.full((h, w), 77.0), np.full((h, w), 200.0)], -1)
    elif kind == "dark":  # values 0..3 and 252..255: the parity-bit grid at both ends of the range
        img = rng.randint(0, 4, (h, w, 3)).astype(np.float64)
        img[:, w // 2:] += 252
    else:  # ramp: black to white, saturated corners
        img = np.stack([xx * 255. / max(w - 1, 1), yy * 255. / max(h - 1, 1), (xx + yy) * 255. / max(w + h - 2, 1)], -1)
    return np.concatenate([img.clip(0, 255).astype(np.uint8), np.full((h, w, 1), 255, np.uint8)], -1).copy()


def _call(lib, name, rgba):
    h, w, _ = rgba.shape
    out = np.full(w * h, 0xAB, np.uint8)
    f = getattr(lib, name)
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    assert f(rgba.ctypes.data, w, h, out.ctypes.data) == 0
    return out


def host_blocks(rgba):
    """The product's encoder source, host instantiation (IEEE arithmetic) — what the CUDA kernel computes."""
    from facebook360_dep_b200 import capi
    return _call(capi.load_cuda().lib, "derp_test_bc7_blocks_host", rgba)


_X86 = {}


def x86_blocks(rgba):
    path = os.path.join(oracle_libs.ROOT, "oracle", "libbc7_x86.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(oracle_libs.ROOT, "oracle"), "libbc7_x86.so"])
    if "lib" not in _X86:
        _X86["lib"] = C.CDLL(path)
    return _call(_X86["lib"], "derp_x86_bc7_blocks", rgba)


def ref_blocks(ref, rgba):
    h, w, _ = rgba.shape
    out = np.full(w * h, 0xAB, np.uint8)
    ref.check(ref.lib.derp_bc7_compress(0, rgba.ctypes.data, w, h, out.ctypes.data))
    return out


def psnr(decoded, rgba):
    h, w = decoded.shape[:2]
    mse = ((decoded[..., :3].astype(np.float64) - rgba[:h, :w, :3]) ** 2).mean()
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


@pytest.mark.parametrize("kind", KINDS + HARD_KINDS)
@pytest.mark.parametrize("size", [(128, 96), (70, 50)])  # the second: neither side a multiple of 4
def test_same_arithmetic_gives_the_reference_bytes(ref, kind, size):
    rgba = surface(3, size[0], size[1], kind)
    a, b = ref_blocks(ref, rgba), x86_blocks(rgba)
    assert np.array_equal(a, b), "blocks differ at bytes %s" % np.where(a != b)[0][:8]


def test_same_arithmetic_on_random_surfaces(ref):
    """150 more surfaces of random sizes (1 to 79 pixels a side, mostly not multiples of 4) and adversarial statistics —
    white noise, a constant with +-3 noise, ramps, five extreme values, 2 x 2-replicated pixels (exact ties everywhere), grey
    with +-1 colour noise — and random alpha (ignored by the opaque profile): the x86-estimate build of the encoder source is
    byte-identical to the reference's encoder on all of them."""
    rng = np.random.RandomState(123)
    for it in range(150):
        w, h = int(rng.randint(1, 20)) * 4 + int(rng.randint(0, 4)), int(rng.randint(1, 20)) * 4 + int(rng.randint(0, 4))
        mode = it % 6
        if mode == 0:
            img = rng.randint(0, 256, (h, w, 3))
        elif mode == 1:
            img = rng.randint(0, 256, (1, 1, 3)) + rng.randint(-3, 4, (h, w, 3))
        elif mode == 2:
            yy, xx = np.mgrid[0:h, 0:w]
            img = np.stack([xx * rng.uniform(0, 8), yy * rng.uniform(0, 8), (xx + yy) * rng.uniform(0, 4)], -1) + rng.randint(0, 50)
        elif mode == 3:
            img = rng.choice([0, 255, 128, 1, 254], (h, w, 3))
        elif mode == 4:
            img = np.repeat(np.repeat(rng.randint(0, 256, ((h + 1) // 2, (w + 1) // 2, 3)), 2, 0), 2, 1)[:h, :w]
        else:
            img = rng.randint(0, 256, (h, w, 1)).repeat(3, -1) + rng.randint(-1, 2, (h, w, 3))
        rgba = np.concatenate([np.clip(img, 0, 255).astype(np.uint8), rng.randint(0, 256, (h, w, 1)).astype(np.uint8)], -1).copy()
        assert np.array_equal(ref_blocks(ref, rgba), x86_blocks(rgba)), (it, w, h, mode)


@pytest.mark.parametrize("kind,min_identical", [("smooth", 0.97), ("noise", 0.99), ("flat", 0.6), ("edges", 0.8), ("ramp", 0.9)])
def test_ieee_encoder_against_reference(ref, kind, min_identical):
    w, h = 128, 96
    rgba = surface(4, w, h, kind)
    a, b = ref_blocks(ref, rgba), host_blocks(rgba)
    identical = (a.reshape(-1, 16) == b.reshape(-1, 16)).all(1).mean()
    da, db = bc7_decode.decode_surface(a, w, h), bc7_decode.decode_surface(b, w, h)
    pa, pb = psnr(da, rgba), psnr(db, rgba)
    print("%s: %.1f %% of the blocks identical, decoded PSNR reference %.3f dB / product %.3f dB" % (kind, 100 * identical, pa, pb))
    assert identical >= min_identical
    assert abs(pa - pb) <= 0.15
    modes = {bc7_decode.block_mode(x) for x in b.reshape(-1, 16)}
    assert modes <= {1, 3, 6}
    assert set(np.unique(db[..., 3])) <= {254, 255}  # the reference's opaque alpha: 127 << 1 | parity


def test_output_layout_and_partial_blocks(ref):
    """Block row r starts at byte r * width * 4; bytes that no block covers stay zero (kernel.ispc:152-159, BC7Util.h:73)."""
    w, h = 70, 50
    rgba = surface(5, w, h, "smooth")
    b = host_blocks(rgba)
    used = np.zeros(w * h, bool)
    for by in range(h // 4):
        used[by * w * 4: by * w * 4 + (w // 4) * 16] = True
    assert not b[~used].any() and b[used].reshape(-1, 16)[:, 0].all()
    full = host_blocks(np.ascontiguousarray(rgba[:48, :68]))
    assert np.array_equal(b[used].reshape(h // 4, -1), full.reshape(h // 4, -1))


def test_decoder_on_hand_made_blocks():
    """Pins tests/bc7_decode.py itself: blocks assembled bit by bit from the format's field layout."""
    def pack(fields):
        v, pos = 0, 0
        for value, bits in fields:
            v |= value << pos
            pos += bits
        assert pos == 128
        return np.frombuffer(v.to_bytes(16, "little"), np.uint8)
    # mode 6: end points (10, 20, 30, 127 | p = 0) and (100, 110, 120, 127 | p = 1), indexes k (first 3 bits)
    f = [(64, 7)] + [(v, 7) for pair in ((10, 100), (20, 110), (30, 120), (127, 127)) for v in pair] + [(0, 1), (1, 1)]
    f += [(0, 3)] + [(k, 4) for k in range(1, 16)]
    d = bc7_decode.decode_block(pack(f))
    e0, e1 = np.array([20, 40, 60, 254]), np.array([201, 221, 241, 255])
    for k in range(16):
        wk = bc7_decode.WEIGHTS[4][k]
        assert np.array_equal(d[k], (e0 * (64 - wk) + e1 * wk + 32) >> 6)
    # mode 1, partition 13 (upper half subset 0, lower half subset 1; anchors 0 and 15), shared parity bits 1 / 0
    f = [(2, 2), (13, 6)] + [(v, 6) for c in range(3) for v in (5 + c, 50 + c, 20 + c, 60 + c)] + [(1, 1), (0, 1)]
    idx = [3, 1, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 3, 2, 1, 3]
    f += [(q, 2 if k in (0, 15) else 3) for k, q in enumerate(idx)]
    d = bc7_decode.decode_block(pack(f))
    for k in range(16):
        s = k >= 8
        ends = [[(5 + c) * 2 + 1 for c in range(3)], [(50 + c) * 2 + 1 for c in range(3)]] if not s else \
            [[(20 + c) * 2 for c in range(3)], [(60 + c) * 2 for c in range(3)]]
        ends = [[(v << 1) | (v >> 6) for v in e] for e in ends]
        wk = bc7_decode.WEIGHTS[3][idx[k]]
        assert list(d[k, :3]) == [(a * (64 - wk) + b * wk + 32) >> 6 for a, b in zip(*ends)] and d[k, 3] == 255


@pytest.mark.parametrize("bits", [8, 16])
@pytest.mark.parametrize("gamma", [2.2 / 1.8, 1.0, 0.45])
def test_gamma_table_equals_reference_gamma_correct(ref, bits, gamma):
    """The lookup table the kernel reads = bc7_util::gammaCorrect (BC7Util.h:41-43) of cv_util::convertTo's float value
    (CvUtil.h:196-207), for EVERY stored channel value — compared with the reference's function itself."""
    from facebook360_dep_b200 import capi
    n = 1 << bits
    lut = np.zeros(n, np.uint8)
    f = capi.load_cuda().lib.derp_test_bc7_gamma_table
    f.restype, f.argtypes = C.c_int, [C.c_int, C.c_float, C.c_void_p]
    assert f(bits, gamma, lut.ctypes.data) == 0
    g = ref.lib.derp_ref_gamma_correct
    g.restype, g.argtypes = C.c_uint8, [C.c_float, C.c_float]
    scale = np.float32(1.0) / np.float32(n - 1)
    values = (np.arange(n, dtype=np.float32) * scale).astype(np.float32)
    step = 1 if bits == 8 else 7
    want = np.array([g(float(v), gamma) for v in values[::step]], np.uint8)
    assert np.array_equal(lut[::step], want)
    assert lut[0] == 0 and lut[-1] == 255


def _stored_image(dtype, channels, w=96, h=64, seed=8):
    rng = np.random.RandomState(seed)
    top = np.iinfo(dtype).max
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([0.5 + 0.4 * np.sin(xx / 19.), 0.5 + 0.4 * np.cos(yy / 13.), (xx + yy) / float(w + h)], -1)
    img = ((img + rng.normal(0, 0.01, img.shape)).clip(0, 1) * top).astype(dtype)
    if channels == 4:
        img = np.concatenate([img, np.full((h, w, 1), top, dtype)], -1)
    return img


def _table_mapped_surface(img, gamma):
    """What the kernel's fused load produces: stored B, G, R values through the gamma table -> R, G, B, 255."""
    from facebook360_dep_b200 import capi
    lut = np.zeros(int(np.iinfo(img.dtype).max) + 1, np.uint8)
    f = capi.load_cuda().lib.derp_test_bc7_gamma_table
    f.restype, f.argtypes = C.c_int, [C.c_int, C.c_float, C.c_void_p]
    assert f(8 * img.dtype.itemsize, gamma, lut.ctypes.data) == 0
    h, w = img.shape[:2]
    return np.concatenate([lut[img[..., 2::-1]], np.full((h, w, 1), 255, np.uint8)], -1).copy()


@pytest.mark.parametrize("dtype,channels", [(np.uint16, 3), (np.uint8, 3), (np.uint16, 4), (np.uint8, 4)])
def test_table_packing_equals_reference_compress_bc7(ref, dtype, channels):
    """bc7_util::compressBC7 itself (conversion to float, gamma, packing, encoder, file writer; BC7Util.h:45-98) on a stored
    image == the reference's encoder on the table-mapped RGBA8 surface: the fused load feeds the encoder the same bytes."""
    img = _stored_image(dtype, channels)
    gamma = 2.2 / 1.8
    assert np.array_equal(ref.bc7_compress_image(img, gamma), ref_blocks(ref, _table_mapped_surface(img, gamma)))


def test_no_cpu_path_behind_the_abi():
    """derp_bc7_compress is CUDA only: without a device it reports the CUDA error instead of encoding on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from facebook360_dep_b200 import capi
    prod = capi.load_cuda()
    with pytest.raises(capi.DerpError) as e:
        prod.bc7_compress(surface(0, 16, 16, "smooth"))
    assert e.value.code == capi.ECUDA
    with pytest.raises(capi.DerpError):
        prod.bc7_compress_image(np.zeros((8, 8, 3), np.uint16))


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_gpu_blocks_equal_host_instantiation(cuda, kind):
    for (w, h) in ((256, 192), (70, 50)):
        rgba = surface(6, w, h, kind)
        assert np.array_equal(cuda.bc7_compress(rgba), host_blocks(rgba))


@pytest.mark.gpu
def test_gpu_device_pointers_and_full_size(cuda):
    import torch
    w, h = 2048, 2048
    rgba = surface(7, w, h, "smooth")
    src = torch.from_numpy(rgba).cuda()
    out = torch.empty(w * h, dtype=torch.uint8, device="cuda")
    cuda.check(cuda.lib.derp_bc7_compress(0, src.data_ptr(), w, h, out.data_ptr()))
    got = out.cpu().numpy()
    rows = slice(0, 64 * w)  # the first 16 block rows against the host instantiation
    assert np.array_equal(got[rows], host_blocks(np.ascontiguousarray(rgba[:64]))[rows])
    assert np.array_equal(got, cuda.bc7_compress(rgba))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,channels", [(np.uint16, 3), (np.uint8, 3), (np.uint16, 4), (np.uint8, 4)])
def test_gpu_image_entry_against_reference_compress_bc7(cuda, ref, dtype, channels):
    """derp_bc7_compress_image against bc7_util::compressBC7 itself (the reference's conversion, gamma, packing, encoder and
    file writer): the packed RGBA8 surface is the same (checked through the host instantiation on the table-mapped image),
    the blocks agree as far as IEEE vs estimate arithmetic allows."""
    w, h = 160, 120
    img = _stored_image(dtype, channels, w, h)
    gamma = 2.2 / 1.8
    got = cuda.bc7_compress_image(img, gamma)
    rgba = _table_mapped_surface(img, gamma)
    assert np.array_equal(got, host_blocks(rgba))
    want = ref.bc7_compress_image(img, gamma)
    identical = (got.reshape(-1, 16) == want.reshape(-1, 16)).all(1).mean()
    pa, pb = psnr(bc7_decode.decode_surface(want, w, h), rgba), psnr(bc7_decode.decode_surface(got, w, h), rgba)
    print("compressBC7: %.1f %% blocks identical, PSNR reference %.3f / product %.3f" % (100 * identical, pa, pb))
    # the reference side depends on the host CPU's RCPPS / RSQRTPS tables (98.7-99.1 % on the Intel boxes measured): keep slack
    assert identical >= 0.85 and abs(pa - pb) <= 0.1
