#!/usr/bin/env python
"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel family once on small shapes.
usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "julia-newtonraphsonhank_b200"))
from hankb200 import HouseholdBlock, model as M

for (n_a, n_e, T, K) in ((60, 3, 6, 5), (500, 7, 5, 9), (1000, 7, 4, 3), (2000, 11, 3, 2)):
    g = M.double_exponential(n_a, 0.0, 200.0); z, Pi = M.rouwenhorst_discretization(n_e, 0.966, 0.283)
    blk = HouseholdBlock(g, z, Pi, 0.98, 2.0, 0.0, T)
    P = T - 1
    vT = 1.015 * ((0.015 * g[None, :] + 1.35 * z[:, None]) + 0.1) ** -2.0
    blk.set_terminal(vT); blk.set_initial_dist(np.full((n_e, n_a), 1.0 / (n_a * n_e)))
    blk.ks_configure(0.36, 0.08, 8.0)
    rng = np.random.default_rng(0)
    r = np.full(P, 0.015); w = np.full(P, 1.35)
    KD, dKD = blk.block(r, w, rng.standard_normal((K, P)), rng.standard_normal((K, P)))
    x = np.tile([2.1, 8.0, 0.015, 1.35], P)
    F = blk.linearize(x, np.ones(P)); JV = blk.jvp(rng.standard_normal((K, 4 * P)))
    J = blk.jacobian_columns(1, 4 * P + 1)
    val, pol, dv, dp, it = blk.vfi(0.02, 1.3, [1.0, 0.0], [0.0, 1.0], eps=1e-3, max_iter=50)
    print(n_a, n_e, T, K, "ok", float(KD[0]), float(np.abs(JV).max()), it)
    blk.close()
