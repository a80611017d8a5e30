"""Loaders of the two CHECKER libraries (test infrastructure; never imported by the product package).

  load_oracle() -> oracle/libderp_oracle.so      the CPU restatement of the depth path (oracle/derp_oracle.cpp)
  load_ref()    -> oracle/_ref/libderp_ref.so    the reference's OWN sources compiled against stand-in headers
                                                 (oracle/ref_bridge.cpp, oracle/refshim/); None when it has not
                                                 been built (it can only be built where /root/reference exists;
                                                 the prebuilt file travels to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call these.
"""
import os
import subprocess

from facebook360_dep_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "libderp_oracle.so")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libderp_ref.so")

_cache = {}


def load_oracle():
    if "oracle" not in _cache:
        if not os.path.exists(ORACLE_LIB):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libderp_oracle.so"])
        lib = capi.Library(ORACLE_LIB)
        if lib.backend != "oracle-cpu":
            raise RuntimeError("unexpected oracle backend %r" % lib.backend)
        _cache["oracle"] = lib
    return _cache["oracle"]


def load_ref():
    """The compiled reference, or None if oracle/_ref has not been built."""
    if "ref" not in _cache:
        if not os.path.exists(REF_LIB) and os.path.exists("/root/reference/source/depth_estimation/Derp.cpp"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
        if not os.path.exists(REF_LIB):
            _cache["ref"] = None
        else:
            lib = capi.Library(REF_LIB)
            if lib.backend != "reference-cpu":
                raise RuntimeError("unexpected reference backend %r" % lib.backend)
            _cache["ref"] = lib
    return _cache["ref"]
