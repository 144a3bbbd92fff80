#!/usr/bin/env python
"""Per-kernel CUDA-event times of the four sweep kernels for a workload and lane count.
usage: python tools/kernel_times.py [--workload ks_500x7_T300] [--lanes 1 4 148 592] [--reps 5]
Also the command ncu profiles (tools/ is measurement tooling, not product code)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "julia-newtonraphsonhank_b200"))
from bench import load_fixture  # noqa: E402
from hankb200 import HouseholdBlock  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="ks_500x7_T300")
ap.add_argument("--lanes", type=int, nargs="+", default=[1, 4, 148, 592])
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
fx = load_fixture(args.workload)
g = fx["g"]
blk = HouseholdBlock(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), fx["T"])
blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"]); blk.ks_configure(*fx["ks"])
G = fx["n_a"] * fx["n_e"]; P = fx["P"]
rng = np.random.default_rng(0)
for K in args.lanes:
    V = rng.standard_normal((K, fx["n"]))
    blk.linearize(fx["x0"], fx["Z"]); blk.jvp(V)   # warm-up
    blk.profile(True); blk.kernel_times(reset=True)
    for _ in range(args.reps):
        blk.linearize(fx["x0"], fx["Z"]); blk.jvp(V)
    kt = blk.kernel_times(reset=True); blk.profile(False)
    per = {k: v[0] / max(v[1], 1) for k, v in kt.items()}
    alg = 8.0 * G * P * K
    print(json.dumps({"workload": args.workload, "K": K, "ms": {k: round(v, 4) for k, v in per.items()},
                      "us_per_period": {k: round(1e3 * v / P, 3) for k, v in per.items()},
                      "GBps": {k: round(alg / (per[k] * 1e-3) / 1e9, 1) for k in ("backward_tangent", "forward_tangent")}}))
