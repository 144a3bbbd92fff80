#!/usr/bin/env python
"""Strong scaling of the full sequence-space Jacobian build (BASELINE config 3): the n columns are split
over the ranks, every rank recomputes the primal, builds its column block with unit-seed lanes and the
blocks are all-gathered with NCCL (hank_allgather_columns_dev).  Launch with torchrun; rank 0 prints JSON.
usage: python -m torch.distributed.run --nproc-per-node N tools/jacobian_scaling.py"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "julia-newtonraphsonhank_b200"))
from bench import load_fixture  # noqa: E402
from hankb200 import HouseholdBlock  # noqa: E402
from hankb200.sharding import column_partition  # noqa: E402

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
fx = load_fixture("ks_500x7_T300"); g = fx["g"]; n, P = fx["n"], fx["P"]
blk = HouseholdBlock(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), fx["T"], device=local)
blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"]); blk.ks_configure(*fx["ks"])
L = blk._L; h = blk.handle
if world > 1:
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt = torch.tensor(list(HouseholdBlock.comm_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(idt, 0)
    blk.comm_init(world, rank, bytes(idt.cpu().numpy().tolist()))
parts = column_partition(n, world)
kmax = max(e - b for b, e in parts)
b, e = parts[rank]
blk.reserve_lanes(kmax)
xd = torch.from_numpy(fx["x0"]).to(dev); Zd = torch.ones(P, dtype=torch.float64, device=dev)
Fd = torch.empty(n, dtype=torch.float64, device=dev)
loc = torch.zeros((kmax, n), dtype=torch.float64, device=dev)        # column-major block, padded to kmax columns
allb = torch.empty((world * kmax, n), dtype=torch.float64, device=dev)
vp = lambda t: C.c_void_p(t.data_ptr())


def build():
    blk._ck(L.hank_ks_linearize_dev(h, vp(xd), vp(Zd), vp(Fd)))
    blk._ck(L.hank_ks_jacobian_columns_dev(h, b, e, vp(loc)))
    blk._ck(L.hank_allgather_columns_dev(h, vp(loc), kmax * n, vp(allb)))


for _ in range(3):
    build()
blk.sync(); torch.cuda.synchronize()
if world > 1:
    dist.barrier()
best = 1e9
for _ in range(5):
    blk.sync()
    if world > 1:
        dist.barrier()
    blk.timer_start(); build(); ms = blk.timer_stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    best = min(best, float(t.item()))
# reassemble and check against rank 0's own single-GPU columns
J = torch.cat([allb[r * kmax: r * kmax + (parts[r][1] - parts[r][0])] for r in range(world)], 0)
chk = None
if rank == 0:
    ref = blk.jacobian_columns(1, 9)
    chk = float(np.max(np.abs(J[:8].cpu().numpy().T - ref)))
    print(json.dumps({"config": "C3 full Jacobian build, KS 500x7 T=300, 1196 columns (598 household lanes)", "n_gpus": world,
                      "ms": round(best, 3), "columns_per_rank": kmax, "allgather_bytes_per_rank": kmax * n * 8,
                      "max_abs_diff_first_columns_vs_single_gpu": chk}))
if world > 1:
    dist.destroy_process_group()
