"""GPU vs oracle at the full BASELINE.json shapes (through the C ABI), 1e-10 relative / 1e-12 absolute.

C2/C3: Krusell-Smith 500 x 7, T = 300 (model-consistent inputs: steady-state record of tests/golden/ss_500x7_T300.npz,
shock Z_t = 1 + 0.8^t, RunMain.jl:50-51) for lane counts that select every tangent-kernel shape; the oracle
evaluates a spread of eight lanes of each pass (lanes are independent of each other).
C5: 1000 x 7, T = 300, 64 lanes.  C4: 2000 x 11, T = 500 (synthetic-throughput inputs), 2 lanes,
and 64 lanes through size-independent properties (linearity across lanes, lanes equal to the 2-lane pass).
Multi-GPU: the NCCL-gathered Jacobian is identical on every rank and equals the one-GPU Jacobian (skipped with fewer
than 2 GPUs)."""
import os
import sys

import numpy as np
import pytest

from common import close, maxerr, make_block, make_oracle, synthetic
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _fixture(name):
    g = np.load(os.path.join(GOLD, name))
    T = int(g["T"]); P = T - 1
    m = dict(grid=g["grid"], z=g["z"], Pi=g["Pi"], beta=float(g["beta"]), gamma=float(g["gamma"]), borrow_cons=float(g["borrow_cons"]))
    ks = (float(g["alpha"]), float(g["delta"]), float(g["ss_vars"][1]))
    x0 = np.tile(g["ss_vars"][:4], P); Z = 1.0 + 0.8 ** np.arange(1, P + 1)
    return g, m, T, P, ks, x0, Z


def _ks_block(g, m, T, ks):
    blk = make_block(m, T)
    blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"])
    blk.ks_configure(*ks)
    return blk


# lane counts -> kernel shape at 500 rows on a 148-SM part: 1, 9: one lane per 8-CTA row-split cluster; 40: one lane per
# 2-CTA row-split cluster; 100: one lane per CTA; 200: 2 lanes per CTA; 592: 4 lanes per CTA (one wave); 700: 6 lanes
@pytest.mark.parametrize("K", [1, 9, 40, 100, 200, 592, 700])
def test_c2_500x7_T300_every_lane_shape(K):
    g, m, T, P, ks, x0, Z = _fixture("ss_500x7_T300.npz")
    n = 4 * P
    rng = np.random.default_rng(100 + K)
    V = rng.standard_normal((K, n))
    blk = _ks_block(g, m, T, ks)
    F, JV = blk.fjvp(x0, Z, V)
    blk.close()
    sel = np.unique(np.linspace(0, K - 1, min(K, 8)).astype(int))
    orc = O.Oracle(g["grid"], g["z"], g["Pi"], m["beta"], m["gamma"], m["borrow_cons"], T)
    Fo, JVo = orc.ks_fjvp(ks, g["ss_value"], g["ss_D"], Z, x0, V[sel])
    assert close(F, Fo), maxerr(F, Fo)
    assert close(JV[sel], JVo), maxerr(JV[sel], JVo)
    assert np.all(np.isfinite(JV))


def test_c2_policies_distributions_brackets_500x7_T300():
    """Policy functions, distributions and lottery brackets of the full-size primal path."""
    g, m, T, P, ks, x0, Z = _fixture("ss_500x7_T300.npz")
    xm = x0.reshape(P, 4)
    r = xm[:, 2] * (1 + 0.05 * 0.9 ** np.arange(P)); w = xm[:, 3] * (1 + 0.02 * 0.9 ** np.arange(P))
    orc = make_oracle(m, T)
    pol_o, _, v1_o, _ = orc.backward(g["ss_value"], r, w)
    KD_o, _, Dp_o, _ = orc.forward(g["ss_D"].reshape(len(m["z"]), -1), pol_o, want_path=True)
    blk = make_block(m, T)
    blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"])
    KD, _ = blk.block(r, w)
    assert close(KD, KD_o), maxerr(KD, KD_o)
    assert close(blk.value_first(), v1_o)
    flips = 0
    for t in (1, 2, 77, 150, 298, 299):
        assert close(blk.policy(t), pol_o[t - 1]), (t, maxerr(blk.policy(t), pol_o[t - 1]))
        assert close(blk.dist(t), Dp_o[t - 1]), (t, maxerr(blk.dist(t), Dp_o[t - 1]))
        # identical policy inputs: bit-exact brackets; end to end: count flips from 1-ulp pow differences
        mo, _ = orc.lottery(pol_o[t - 1])
        mg, _ = blk.lottery(pol_o[t - 1])
        assert np.array_equal(mg, mo)
        flips += int(np.sum(blk.brackets(t) != mo))
    assert flips <= 2, flips
    blk.close()


def test_c5_1000x7_T300_64_lanes():
    g, m, T, P, ks, x0, Z = _fixture("ss_1000x7_T300.npz")
    n = 4 * P
    V = np.random.default_rng(5).standard_normal((64, n))
    blk = _ks_block(g, m, T, ks)
    F, JV = blk.fjvp(x0, Z, V)
    blk.close()
    sel = np.arange(0, 64, 4)
    orc = O.Oracle(g["grid"], g["z"], g["Pi"], m["beta"], m["gamma"], m["borrow_cons"], T)
    Fo, JVo = orc.ks_fjvp(ks, g["ss_value"], g["ss_D"], Z, x0, V[sel])
    assert close(F, Fo), maxerr(F, Fo)
    assert close(JV[sel], JVo), maxerr(JV[sel], JVo)


def test_c4_2000x11_T500():
    s = synthetic(2000, 11, 500, 2)
    orc = make_oracle(s["m"], 500)
    pol_o, dpol_o, v1_o, _ = orc.backward(s["vT"], s["r"], s["w"], s["dr"], s["dw"])
    KD_o, dKD_o = orc.forward(s["D0"], pol_o, dpol_o)
    blk = make_block(s["m"], 500)
    blk.set_terminal(s["vT"]); blk.set_initial_dist(s["D0"])
    KD, dKD = blk.block(s["r"], s["w"], s["dr"], s["dw"])
    assert close(KD, KD_o), maxerr(KD, KD_o)
    assert close(dKD, dKD_o), maxerr(dKD, dKD_o)
    for t in (1, 250, 499):
        assert close(blk.policy(t), pol_o[t - 1])
        for l in (0, 1):
            assert close(blk.policy(t, l + 1), dpol_o[l][t - 1]), (t, l, maxerr(blk.policy(t, l + 1), dpol_o[l][t - 1]))
    # 64 lanes (two lanes per 4-CTA row-split cluster): lanes 0, 1 repeat the 2-lane pass, lane 63 is a combination
    rng = np.random.default_rng(9)
    dr = rng.standard_normal((64, 499)); dw = rng.standard_normal((64, 499))
    dr[:2] = s["dr"]; dw[:2] = s["dw"]
    dr[63] = 0.5 * dr[0] - 2.0 * dr[7]; dw[63] = 0.5 * dw[0] - 2.0 * dw[7]
    KD64, dKD64 = blk.block(s["r"], s["w"], dr, dw)
    assert np.array_equal(KD64, KD)
    assert close(dKD64[:2], dKD_o), maxerr(dKD64[:2], dKD_o)
    assert close(dKD64[63], 0.5 * dKD64[0] - 2.0 * dKD64[7], rtol=1e-9, atol=1e-11)
    blk.close()


# ---------------------------------------------------------------- multi-GPU
def _rank_main(rank, world, port, out):
    import ctypes as C
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "julia-newtonraphsonhank_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from hankb200 import HouseholdBlock
    from hankb200.sharding import column_partition, period_round_robin, round_robin_permutation
    g = np.load(os.path.join(GOLD, "ks_200x7_T40.npz"))
    T = int(g["T"]); P = T - 1; n = 4 * P
    blk = HouseholdBlock(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), T, device=rank)
    blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"])
    blk.ks_configure(float(g["alpha"]), float(g["delta"]), float(g["ss_vars"][1]))
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt = torch.tensor(list(HouseholdBlock.comm_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(idt, 0)
    blk.comm_init(world, rank, bytes(idt.cpu().numpy().tolist()))
    blk.linearize(g["x0"], np.ones(P))
    L = blk._L; h = blk.handle
    vp = lambda t: C.c_void_p(t.data_ptr())
    res = {}
    # contiguous column blocks
    parts = column_partition(n, world); kmax = max(e - b for b, e in parts); b, e = parts[rank]
    loc = torch.zeros((kmax, n), dtype=torch.float64, device=dev); allb = torch.empty((world * kmax, n), dtype=torch.float64, device=dev)
    blk._ck(L.hank_ks_jacobian_columns_dev(h, b, e, vp(loc)))
    blk._ck(L.hank_allgather_columns_dev(h, vp(loc), kmax * n, vp(allb)))
    blk.sync()
    res["contig"] = torch.cat([allb[r * kmax: r * kmax + (parts[r][1] - parts[r][0])] for r in range(world)], 0).cpu().numpy().T
    # round-robin periods (balanced seed horizons)
    cols = period_round_robin(n, world, rank); kmax = max(len(period_round_robin(n, world, r)) for r in range(world))
    loc = torch.zeros((kmax, n), dtype=torch.float64, device=dev); allb = torch.empty((world * kmax, n), dtype=torch.float64, device=dev)
    blk._ck(L.hank_ks_jacobian_column_list_dev(h, len(cols), cols.ctypes.data_as(C.POINTER(C.c_int)), vp(loc)))
    blk._ck(L.hank_allgather_columns_dev(h, vp(loc), kmax * n, vp(allb)))
    blk.sync()
    blocks = [allb[r * kmax: r * kmax + len(period_round_robin(n, world, r))] for r in range(world)]
    J = np.empty((n, n)); J[:, round_robin_permutation(n, world)] = torch.cat(blocks, 0).cpu().numpy().T
    res["balanced"] = J
    res["single"] = blk.jacobian_columns(1, n + 1)
    np.savez(os.path.join(out, f"r{rank}.npz"), **res)
    blk.close()
    dist.destroy_process_group()


def test_multi_gpu_jacobian_bit_identical(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world, port = 2, 29600 + os.getpid() % 2000
    mp.start_processes(_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    g = np.load(os.path.join(GOLD, "ks_200x7_T40.npz"))
    r0 = np.load(tmp_path / "r0.npz"); r1 = np.load(tmp_path / "r1.npz")
    for key in ("contig", "balanced"):
        assert np.array_equal(r0[key], r1[key]), key                       # every rank holds the same gathered matrix
        # and it is the one-GPU Jacobian: bit for bit when both builds run the same kernel shape, else to the last
        # bits (a pass with fewer lanes may run the row-split kernels, whose K̇D partial sums add in another order)
        assert close(r0[key], r0["single"], rtol=1e-12, atol=1e-14), (key, maxerr(r0[key], r0["single"]))
    assert close(r0["single"], g["Jbar"]), maxerr(r0["single"], g["Jbar"])  # which matches the oracle's
