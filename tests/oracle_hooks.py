"""ctypes access to the oracle-only test hooks (oracle/derp_oracle.cpp, bottom)."""
import ctypes as C

import numpy as np

from facebook360_dep_b200 import capi


def _lib(oracle):
    return oracle.lib


def camera_project(oracle, desc, pts, normalized=False):
    pts = np.ascontiguousarray(pts, np.float64)
    n = len(pts)
    pix = np.empty((n, 2), np.float64)
    sees = np.empty(n, np.uint8)
    f = _lib(oracle).oracle_camera_project
    f.restype = C.c_int
    f.argtypes = [C.POINTER(capi.CameraDesc), C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    oracle.check(f(C.byref(desc), int(normalized), pts.ctypes.data, n, pix.ctypes.data, sees.ctypes.data))
    return pix, sees.astype(bool)


def camera_unproject(oracle, desc, pix, depth):
    pix = np.ascontiguousarray(pix, np.float64)
    n = len(pix)
    out = np.empty((n, 3), np.float64)
    oc = np.empty(n, np.uint8)
    f = _lib(oracle).oracle_camera_unproject
    f.restype = C.c_int
    f.argtypes = [C.POINTER(capi.CameraDesc), C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    oracle.check(f(C.byref(desc), pix.ctypes.data, n, float(depth), out.ctypes.data, oc.ctypes.data))
    return out, oc.astype(bool)


def camera_info(oracle, desc):
    rot = np.empty(9, np.float64)
    dm, cf = C.c_double(), C.c_double()
    f = _lib(oracle).oracle_camera_info
    f.restype = C.c_int
    f.argtypes = [C.POINTER(capi.CameraDesc), C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    oracle.check(f(C.byref(desc), rot.ctypes.data, C.byref(dm), C.byref(cf)))
    return rot.reshape(3, 3), dm.value, cf.value


def camera_distort(oracle, desc, r):
    r = np.ascontiguousarray(r, np.float64)
    a = np.empty_like(r)
    b = np.empty_like(r)
    f = _lib(oracle).oracle_camera_distort
    f.restype = C.c_int
    f.argtypes = [C.POINTER(capi.CameraDesc), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    oracle.check(f(C.byref(desc), r.ctypes.data, len(r), a.ctypes.data, b.ctypes.data))
    return a, b


def remap_bicubic(oracle, src, mp):
    src = np.ascontiguousarray(src, np.uint16)
    mp = np.ascontiguousarray(mp, np.float32)
    dh, dw = mp.shape[:2]
    dst = np.empty((dh, dw, 3), np.uint16)
    f = _lib(oracle).oracle_remap_bicubic
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    f(src.ctypes.data, src.shape[1], src.shape[0], mp.ctypes.data, dw, dh, dst.ctypes.data)
    return dst


def blur3(oracle, src):
    src = np.ascontiguousarray(src, np.uint16)
    dst = np.empty_like(src)
    f = _lib(oracle).oracle_blur3
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    f(src.ctypes.data, src.shape[1], src.shape[0], dst.ctypes.data)
    return dst


def resize_area(oracle, src, dw, dh):
    src = np.ascontiguousarray(src, np.uint16)
    dst = np.empty((dh, dw, 3), np.uint16)
    f = _lib(oracle).oracle_resize_area
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    assert f(src.ctypes.data, src.shape[1], src.shape[0], dst.ctypes.data, dw, dh) == 0
    return dst


def variance(oracle, src):
    src = np.ascontiguousarray(src, np.uint16)
    dst = np.empty(src.shape[:2], np.float32)
    f = _lib(oracle).oracle_variance
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    f(src.ctypes.data, src.shape[1], src.shape[0], dst.ctypes.data)
    return dst


def _resize(oracle, name, src, dw, dh):
    src = np.ascontiguousarray(src, np.float32)
    dst = np.empty((dh, dw), np.float32)
    f = getattr(_lib(oracle), name)
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    f(src.ctypes.data, src.shape[1], src.shape[0], dst.ctypes.data, dw, dh)
    return dst


def lanczos4(oracle, src, dw, dh):
    return _resize(oracle, "oracle_lanczos4", src, dw, dh)


def nearest(oracle, src, dw, dh):
    return _resize(oracle, "oracle_nearest_f32", src, dw, dh)


def minstd_uniform(oracle, seed, n, a, b):
    out = np.empty(n, np.float32)
    f = _lib(oracle).oracle_minstd_uniform
    f.restype = None
    f.argtypes = [C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_void_p]
    f(seed, n, a, b, out.ctypes.data)
    return out


def nth_element_sum(oracle, first, second, keep):
    first = np.ascontiguousarray(first, np.float32)
    second = np.ascontiguousarray(second, np.float32)
    f = _lib(oracle).oracle_nth_element_sum
    f.restype = C.c_float
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    return f(first.ctypes.data, second.ctypes.data, len(first), keep)
