#!/usr/bin/env python
"""Per-kernel CUDA-event times of the tangent sweeps on a synthetic grid for a list of lane counts.
usage: python tools/sweep_times.py --shape 500 7 300 --lanes 1 18 64 [--reps 3]
Measurement tooling (not product code); environment switches (HANK_NO_ROWSPLIT, HANK_RS_MAXK, ...) apply."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "julia-newtonraphsonhank_b200"))
from hankb200 import HouseholdBlock, model as M  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", type=int, nargs=3, default=[500, 7, 300], metavar=("n_a", "n_e", "T"))
ap.add_argument("--lanes", type=int, nargs="+", default=[1, 18, 64])
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--tag", default="")
args = ap.parse_args()
HBM = 6650.0
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    HBM = json.load(open(pk))["hbm_gbs"]
n_a, n_e, T = args.shape
P = T - 1
g = M.double_exponential(n_a, 0.0, 200.0)
z, Pi = M.rouwenhorst_discretization(n_e, 0.966, 0.283)
blk = HouseholdBlock(g, z, Pi, 0.98, 2.0, 0.0, T)
rbar, wbar, c0 = 0.015, 1.35, 0.1
vT = (1 + rbar) * ((rbar * g[None, :] + wbar * z[:, None]) + c0) ** (-2.0)
blk.set_terminal(vT); blk.set_initial_dist(np.full((n_e, n_a), 1.0 / (n_a * n_e)))
t = np.arange(1, P + 1)
r = rbar * (1 + 0.1 * 0.9 ** t); w = wbar * (1 + 0.05 * 0.9 ** t)
rng = np.random.default_rng(1234)
env = {k: v for k, v in os.environ.items() if k.startswith("HANK_")}
for K in args.lanes:
    dr = rng.standard_normal((K, P)); dw = rng.standard_normal((K, P))
    blk.block(r, w, dr, dw)
    blk.profile(True); blk.kernel_times(reset=True)
    for _ in range(args.reps):
        blk.block(r, w, dr, dw)
    kt = blk.kernel_times(reset=True); blk.profile(False)
    per = {k: v[0] / max(v[1], 1) for k, v in kt.items()}
    alg = 8.0 * n_a * n_e * P * K
    gb = {k: alg / (per[k] * 1e-3) / 1e9 for k in ("backward_tangent", "forward_tangent")}
    print(json.dumps({"shape": args.shape, "K": K, "env": env, "tag": args.tag,
                      "ms": {k: round(v, 4) for k, v in per.items()},
                      "us_per_period": {k: round(1e3 * v / P, 3) for k, v in per.items()},
                      "GBps": {k: round(v, 1) for k, v in gb.items()},
                      "frac_of_measured_hbm": {k: round(v / HBM, 4) for k, v in gb.items()}}), flush=True)
blk.close()
