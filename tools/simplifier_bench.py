#!/usr/bin/env python3
"""Host simplifier (derp_simplify.h through the derp_test_simplify hook, no GPU involved) against the reference's own
MeshSimplifier.cpp (oracle/_ref) on a smooth sheet that reaches its target and on a torn mesh that cannot: seconds on one
host thread and whether vertex bits and faces are identical."""
import sys, time, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import oracle_libs, test_mesh
from facebook360_dep_b200 import capi
oracle = oracle_libs.load_oracle(); ref = oracle_libs.load_ref(); prod = capi.load_cuda()
for smooth in (True, False):
    xyz, idx = test_mesh._surface_mesh(oracle, 3, 768, 768, smooth)
    t=time.time(); pv, pi = test_mesh._simplify(prod, "derp_test_simplify", xyz, idx, 150000); tp=time.time()-t
    t=time.time(); rv, ri = test_mesh._simplify(ref, "derp_ref_simplify", xyz, idx, 150000); tr=time.time()-t
    print("smooth" if smooth else "torn", len(idx), "->", len(pi), "product %.2fs reference %.2fs"%(tp,tr), np.array_equal(pi,ri) and np.array_equal(pv.view(np.uint64), rv.view(np.uint64)))
