"""Shared helpers for the GPU parity tests: drive the CUDA library and the oracle with the same calls."""
import functools

import numpy as np

from facebook360_dep_b200 import capi, synth


@functools.lru_cache(maxsize=16)
def scene_inputs(num_cams, width, height, kind, hfov_deg=None, distorted=False, noise=True, seed=42):
    dist = (-0.03413328161902581, 0.0004374554953464843, -0.0018843963208481174) if distorted else None
    if kind == "SPHERE":  # the reference's test-rig layout (synth.sphere_rig)
        rig = synth.sphere_rig(num_cams, width, height)
    else:
        rig = synth.ring_rig(num_cams, width, height, kind=kind, hfov_deg=hfov_deg, distortion=dist)
    colors, true_disp = synth.render_rig(rig, width, height, scene=synth.Scene(seed=seed), noise=noise)
    return rig, colors, true_disp


def make_pair(cuda, oracle, rig, dst_to_src=None):
    descs = capi.rig_descs(rig)
    return capi.Context(cuda, descs, dst_to_src), capi.Context(oracle, descs, dst_to_src)


def both(ctxs, name, *a, **k):
    return [getattr(c, name)(*a, **k) for c in ctxs]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_float_bits(a, b):
    """Bitwise equality with all NaNs considered equal."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))


def mismatch_fraction(a, b):
    return 1.0 - same_float_bits(a, b).mean()
