#!/usr/bin/env python
"""Device time of the hand-written dense inverse (hank_dense_inverse, host buffers: the copies are inside) for a few n.
usage: python tools/dense_time.py [n ...]   (ncu -k regex:k_gj gives the per-kernel split)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "julia-newtonraphsonhank_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hankb200 import HouseholdBlock, model as M
g = M.double_exponential(50, 0.0, 200.0); z, Pi = M.rouwenhorst_discretization(3, 0.966, 0.283)
blk = HouseholdBlock(g, z, Pi, 0.98, 2.0, 0.0, 4)
for n in [int(a) for a in sys.argv[1:]] or [1196]:
    A = np.random.default_rng(n).standard_normal((n, n)) + 3 * np.eye(n)
    blk.dense_inverse(A)
    best = 1e9
    for _ in range(3):
        blk.sync(); blk.timer_start(); X = blk.dense_inverse(A); best = min(best, blk.timer_stop())
    print("n", n, "ms incl. copies", round(best, 3), "resid", np.abs(X @ A - np.eye(n)).max())
