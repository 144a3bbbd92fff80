#!/usr/bin/env python
"""One rank's share of an N-way sharded Jacobian build on ONE GPU (no NCCL): linearise + the rank-0 column list of the
round-robin period partition, with the per-kernel split.  usage: python tools/jacobian_shard_times.py [N ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "julia-newtonraphsonhank_b200"))
import ctypes as C
from bench import load_fixture
from hankb200 import HouseholdBlock
from hankb200.sharding import period_round_robin
fx = load_fixture("ks_500x7_T300"); g = fx["g"]; n, P = fx["n"], fx["P"]
blk = HouseholdBlock(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), fx["T"])
blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"]); blk.ks_configure(*fx["ks"])
blk.reserve_lanes(n // 2)
ones = np.ones(P)
for N in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    cols = period_round_robin(n, N, 0)
    for rep in range(3):
        blk.linearize(fx["x0"], ones); blk.jacobian_column_list(cols)
    blk.profile(True); blk.kernel_times(reset=True)
    best = 1e9
    for rep in range(5):
        blk.sync(); blk.timer_start()
        blk.linearize(fx["x0"], ones); J = blk.jacobian_column_list(cols)
        best = min(best, blk.timer_stop())
    kt = blk.kernel_times(reset=True); blk.profile(False)
    print("N", N, "cols", len(cols), "lanes", int(np.sum((cols - 1) % 4 >= 2)), "ms (host API, copies in)", round(best, 3),
          {k: round(v[0] / max(v[1], 1), 3) for k, v in kt.items()})
