#!/usr/bin/env python
"""Builds the full Jacobian N times in fresh contexts and compares every build with a reference build made without seed
horizons / overflow concurrency (HANK_NO_SKIP=1).  This is the check that exposed the relaxed cluster hand-shake of the
row-split kernels (one column in ~15 builds off by 3e-7..1e-4 when the overflow clusters ran next to the main wave).
usage: python tools/jacobian_repeat_check.py [N]     (HANK_RS_RELAXED=1 reproduces the failure)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "julia-newtonraphsonhank_b200"))
from bench import load_fixture
from hankb200 import HouseholdBlock
fx = load_fixture("ks_500x7_T300"); g = fx["g"]; n, P = fx["n"], fx["P"]
def mk():
    b = HouseholdBlock(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), fx["T"])
    b.set_terminal(g["ss_value"]); b.set_initial_dist(g["ss_D"]); b.ks_configure(*fx["ks"]); return b
os.environ["HANK_NO_SKIP"] = "1"
ref = mk(); ref.linearize(fx["x0"], fx["Z"]); Jref = ref.jacobian_columns(1, n + 1); ref.close()
del os.environ["HANK_NO_SKIP"]
bad = 0
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for it in range(N):
    b = mk()
    b.linearize(fx["x0"], fx["Z"])
    J = b.jacobian_columns(1, n + 1)
    err = np.abs(J - Jref) / (1e-12 * max(1.0, np.abs(Jref).max()) + 1e-10 * np.abs(Jref))
    if not np.isfinite(J).all() or err.max() > 1:
        bad += 1
        cols = np.unique(np.where((err > 1) | ~np.isfinite(J))[1])
        print("build", it, "MISMATCH: max error / tolerance", float(np.nanmax(err)), "columns (0-based)", cols[:20], flush=True)
    b.close()
print("builds", N, "mismatching", bad)
