"""world_size-2 gloo test (CPU) of the multi-GPU host logic: column partition, per-rank column blocks,
all-gather and reassembly.  The per-rank blocks come from the oracle here (no GPU in this container); on the
GPU box the same functions are driven with HouseholdBlock.jacobian_columns."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


class OracleBlock:
    """Duck-types HouseholdBlock.jacobian_columns with the CPU oracle."""

    def __init__(self, g):
        from oracle import oracle as O
        self.g = g
        self.P = int(g["T"]) - 1
        self.orc = O.Oracle(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), int(g["T"]))
        self.ks = (float(g["alpha"]), float(g["delta"]), float(g["ss_vars"][1]))

    def jacobian_columns(self, b, e):
        g = self.g
        return self.orc.jacobian(self.ks, g["ss_value"], g["ss_D"], np.ones(self.P), g["x0"], np.arange(b - 1, e - 1))

    def jacobian_column_list(self, cols):
        g = self.g
        return self.orc.jacobian(self.ks, g["ss_value"], g["ss_D"], np.ones(self.P), g["x0"], np.asarray(cols) - 1)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "julia-newtonraphsonhank_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hankb200.sharding import jacobian_distributed
    g = np.load(os.path.join(GOLD, "ks_100x3_T30.npz"))
    n = 4 * (int(g["T"]) - 1)
    J = jacobian_distributed(OracleBlock(g), n, world, rank)
    np.save(os.path.join(out, f"J{rank}.npy"), J)
    Jb = jacobian_distributed(OracleBlock(g), n, world, rank, balanced=True)
    np.save(os.path.join(out, f"Jb{rank}.npy"), Jb)
    dist.destroy_process_group()


def test_partition_rules():
    from hankb200.sharding import column_partition, lane_slice
    for n, w in ((1196, 8), (116, 2), (7, 3), (5, 8)):
        parts = column_partition(n, w)
        assert parts[0][0] == 1 and parts[-1][1] == n + 1
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [e - b for b, e in parts]
        assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
        sl = [lane_slice(n, w, r) for r in range(w)]
        assert sl[0][0] == 0 and sl[-1][1] == n and all(sl[i][1] == sl[i + 1][0] for i in range(w - 1))


def test_sharded_jacobian_gloo(tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    g = np.load(os.path.join(GOLD, "ks_100x3_T30.npz"))
    J0 = np.load(tmp_path / "J0.npy"); J1 = np.load(tmp_path / "J1.npy")
    assert np.array_equal(J0, J1)
    assert J0.shape == g["Jbar"].shape and np.allclose(J0, g["Jbar"], rtol=1e-13, atol=1e-14)
    # round-robin periods (balanced seed horizons): same matrix after the inverse permutation
    Jb0 = np.load(tmp_path / "Jb0.npy"); Jb1 = np.load(tmp_path / "Jb1.npy")
    assert np.array_equal(Jb0, Jb1) and np.array_equal(Jb0, J0)


def test_round_robin_rules():
    from hankb200.sharding import period_round_robin, round_robin_permutation
    for n, w in ((1196, 8), (116, 2), (28, 3), (8, 4)):
        blocks = [period_round_robin(n, w, r) for r in range(w)]
        allc = np.concatenate(blocks)
        assert sorted(allc.tolist()) == list(range(1, n + 1))
        assert all(np.all(np.diff(b) > 0) for b in blocks if len(b))
        sizes = [len(b) for b in blocks]
        assert max(sizes) - min(sizes) <= 4
        # every rank holds early and late periods alike: mean seed period within one stride of the others
        means = [np.mean((b - 1) // 4) for b in blocks if len(b)]
        assert max(means) - min(means) <= w
        assert np.array_equal(round_robin_permutation(n, w), allc - 1)
