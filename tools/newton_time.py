#!/usr/bin/env python
"""Times hank_newton_solve variants on the bench workload. usage: python tools/newton_time.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "julia-newtonraphsonhank_b200"))
from bench import load_fixture
from hankb200 import HouseholdBlock
fx = load_fixture("ks_500x7_T300"); g = fx["g"]; n, P = fx["n"], fx["P"]
blk = HouseholdBlock(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), fx["T"])
blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"]); blk.ks_configure(*fx["ks"])
blk.reserve_lanes(n // 2)
blk.linearize(fx["x0"], np.ones(P))
for _ in range(3):
    t0 = time.perf_counter(); J = blk.jacobian_columns(1, n + 1); print("jacobian e2e ms", round(1e3 * (time.perf_counter() - t0), 2))
for sv in ("lu", "lu_batched", "lu_batched", "lu"):
    blk.profile(True); blk.kernel_times(reset=True)
    t0 = time.perf_counter(); x, st = blk.newton_solve(J, fx["x0"], fx["Z"], solver=sv); dt = time.perf_counter() - t0
    kt = blk.kernel_times(reset=True); blk.profile(False)
    print(sv, "ms", round(1e3 * dt, 1), "outer", st["outer"], "jvps", st["jvps"], {k: (round(v[0], 1), v[1]) for k, v in kt.items()})
