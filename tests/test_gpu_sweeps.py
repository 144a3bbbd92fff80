"""GPU parity of the sweeps against the CPU oracle (through the C ABI). Tolerance: 1e-10 relative /
1e-12 absolute (north star); lottery brackets bit-exact on identical policy inputs."""
import numpy as np
import pytest

from common import synthetic, make_oracle, make_block, close, maxerr

pytestmark = pytest.mark.gpu

SHAPES = [(60, 3, 12, 2.0), (200, 7, 40, 2.0), (500, 7, 30, 2.0), (300, 7, 25, 1.5), (1000, 7, 12, 2.0),
          (700, 11, 10, 3.0), (2000, 11, 8, 2.0)]


@pytest.mark.parametrize("n_a,n_e,T,gamma", SHAPES)
@pytest.mark.parametrize("K", [0, 1, 3, 5])
def test_block_matches_oracle(n_a, n_e, T, gamma, K):
    s = synthetic(n_a, n_e, T, K, gamma)
    orc = make_oracle(s["m"], T)
    pol_o, dpol_o, v1_o, _ = orc.backward(s["vT"], s["r"], s["w"], s["dr"] if K else None, s["dw"] if K else None)
    KD_o, dKD_o, Dp_o, _ = orc.forward(s["D0"], pol_o, dpol_o if K else None, want_path=True)
    blk = make_block(s["m"], T)
    blk.set_terminal(s["vT"]); blk.set_initial_dist(s["D0"])
    KD, dKD = blk.block(s["r"], s["w"], s["dr"] if K else None, s["dw"] if K else None)
    pol = blk.policies(0)
    assert close(pol, pol_o), maxerr(pol, pol_o)
    assert close(blk.value_first(), v1_o), maxerr(blk.value_first(), v1_o)
    assert close(KD, KD_o), maxerr(KD, KD_o)
    for t in (1, T // 2, T - 1):
        assert close(blk.dist(t), Dp_o[t - 1]), (t, maxerr(blk.dist(t), Dp_o[t - 1]))
        assert abs(blk.dist(t).sum() - 1.0) < 1e-12
    if K:
        for l in range(K):
            dp = blk.policies(l + 1)
            assert close(dp, dpol_o[l]), (l, maxerr(dp, dpol_o[l]))
        assert close(dKD, dKD_o), maxerr(dKD, dKD_o)
    blk.close()


@pytest.mark.parametrize("n_a,n_e", [(60, 3), (500, 7), (2000, 11)])
def test_lottery_brackets_bit_exact(n_a, n_e):
    s = synthetic(n_a, n_e, 4)
    orc = make_oracle(s["m"], 4)
    g = s["m"]["grid"]
    rng = np.random.default_rng(7)
    pol = rng.uniform(-1.0, g[-1] * 1.05, size=(n_e, n_a))
    # exact grid hits, +-0.0, one ulp either side of nodes, below / above the grid
    pol[0, :n_a] = g
    pol[1, : n_a - 1] = np.nextafter(g[1:], -np.inf)
    pol[2, : n_a - 1] = np.nextafter(g[:-1], np.inf)
    pol[-1, 0] = -0.0; pol[-1, 1] = 0.0; pol[-1, 2] = g[-1]; pol[-1, 3] = np.nextafter(g[-1], np.inf)
    m_o, om_o = orc.lottery(pol)
    blk = make_block(s["m"], 4)
    m, om = blk.lottery(pol)
    assert np.array_equal(m, m_o)
    assert np.array_equal(om, om_o)  # same IEEE subtraction and division
    blk.close()
