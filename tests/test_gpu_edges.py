"""GPU edge cases and full-size property tests.

Edge cases against the oracle: grid sizes on the padding boundaries (n_a == LDA), the minimal horizon T = 2,
a binding non-zero borrowing constraint, lane counts that do not divide the lanes per CTA, multi-wave lane
counts, every compiled n_e.  Full BASELINE sizes (500x7 T=300, 2000x11 T=500) through size-independent
properties: linearity of the JVP, agreement with central finite differences, Jacobian column == unit-seed JVP,
mass conservation, Newton residual."""
import os

import numpy as np
import pytest

from common import synthetic, make_oracle, make_block, close, maxerr, model_inputs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compare(n_a, n_e, T, K, gamma=2.0, borrow_cons=0.0, amax=200.0):
    s = synthetic(n_a, n_e, T, K, gamma)
    if borrow_cons:
        s["m"]["borrow_cons"] = borrow_cons
    orc = make_oracle(s["m"], T)
    pol_o, dpol_o, v1_o, _ = orc.backward(s["vT"], s["r"], s["w"], s["dr"] if K else None, s["dw"] if K else None)
    KD_o, dKD_o = orc.forward(s["D0"], pol_o, dpol_o if K else None)
    blk = make_block(s["m"], T)
    blk.set_terminal(s["vT"]); blk.set_initial_dist(s["D0"])
    KD, dKD = blk.block(s["r"], s["w"], s["dr"] if K else None, s["dw"] if K else None)
    assert close(blk.policies(0), pol_o) and close(blk.value_first(), v1_o) and close(KD, KD_o), maxerr(KD, KD_o)
    if K:
        assert close(dKD, dKD_o), maxerr(dKD, dKD_o)
        for l in (0, K // 2, K - 1):
            dp = blk.policies(l + 1)
            assert close(dp, dpol_o[l]), (l, maxerr(dp, dpol_o[l]))
    blk.close()
    return pol_o


@pytest.mark.parametrize("n_a", [2, 33, 256, 257, 512, 1024, 2048])
def test_grid_sizes_on_padding_boundaries(n_a):
    _compare(n_a, 3, 5, 3)


def test_minimal_horizon():
    _compare(200, 7, 2, 5)          # P = 1: one EGM step, one lottery step


def test_binding_nonzero_borrowing_constraint():
    pol = _compare(300, 7, 10, 4, borrow_cons=0.2)
    assert np.mean(pol == 0.2) > 0.05    # the constraint binds on part of the grid


@pytest.mark.parametrize("n_a,n_e,K", [(120, 3, 301), (500, 7, 7), (500, 7, 160), (500, 7, 310), (500, 7, 893), (1000, 7, 149), (1000, 7, 445)])
def test_lane_counts_and_shapes(n_a, n_e, K):
    """1-, 2-, 4-, 6- (and 3-) lane CTAs, partially filled last CTA, more than one wave."""
    _compare(n_a, n_e, 6, K)


@pytest.mark.parametrize("switch,n_a,K", [("HANK_NO_TMA", 500, 9), ("HANK_NO_TMA", 1000, 5), ("HANK_NO_CLUSTER", 500, 3),
                                          ("HANK_NO_DSMEM", 500, 3), ("HANK_NO_DSMEM", 1000, 2), ("HANK_NO_WIDE", 500, 700),
                                          ("HANK_NO_OVERLAP", 500, 5), ("HANK_NO_ROWSPLIT", 500, 5),
                                          ("HANK_NO_ROWSPLIT", 1000, 3), ("HANK_NO_ROWSPLIT", 2000, 2),
                                          ("HANK_NO_RS_ST", 500, 3), ("HANK_NO_RS_PUSH", 500, 3), ("HANK_NO_RS_ST", 200, 2),
                                          ("HANK_NO_PIPE", 500, 700)])
def test_fallback_kernels_agree_with_oracle(switch, n_a, K, monkeypatch):
    """The A/B switches read at hank_ctx_create select the fallback kernels (register-prefetch tangents,
    single-CTA and global-exchange primal sweeps, no 6-lane shape, no side stream): same parity bar."""
    monkeypatch.setenv(switch, "1")
    _compare(n_a, 7, 6, K)


@pytest.mark.parametrize("n_a,n_e,T,K", [(200, 7, 9, 1), (200, 7, 9, 18), (500, 7, 12, 1), (500, 7, 12, 13), (500, 3, 8, 18),
                                         (500, 7, 9, 40), (500, 7, 9, 74), (500, 5, 9, 33), (500, 11, 6, 20),
                                         (1000, 7, 9, 2), (1000, 7, 9, 14), (1000, 7, 9, 19), (1000, 7, 9, 64), (1000, 11, 7, 61),
                                         (2000, 11, 7, 3), (2000, 11, 7, 21), (2000, 11, 7, 64), (2000, 7, 30, 37)])
def test_row_split_cluster_shapes(n_a, n_e, T, K):
    """Few lanes: the rows of a lane group are split over a thread-block cluster (hank_tangent_rowsplit.cuh): one lane per
    cluster of 8 CTAs with a whole period per exchange (or column by column at 2000 rows) while the clusters fit one
    wave, then one lane per cluster of 2 (500 rows) or two lanes per cluster of 4 (1000 / 2000 rows); partially filled
    last clusters, rings that wrap many times."""
    _compare(n_a, n_e, T, K)


def test_row_split_matches_one_cta_kernels(monkeypatch):
    """Same pass through the row-split clusters and through the one-CTA kernels (HANK_NO_ROWSPLIT=1)."""
    s = synthetic(500, 7, 40, 6)
    out = []
    for off in (False, True):
        if off:
            monkeypatch.setenv("HANK_NO_ROWSPLIT", "1")
        blk = make_block(s["m"], 40)
        blk.set_terminal(s["vT"]); blk.set_initial_dist(s["D0"])
        KD, dKD = blk.block(s["r"], s["w"], s["dr"], s["dw"])
        out.append((KD, dKD, np.stack([blk.policies(l + 1) for l in range(6)])))
        blk.close()
    assert np.array_equal(out[0][0], out[1][0])                       # primal: same kernels
    assert np.array_equal(out[0][2], out[1][2])                       # backward tangents: same arithmetic per point
    assert close(out[0][1], out[1][1], rtol=1e-12, atol=1e-14)        # forward: per-CTA partial sums differ in order


@pytest.mark.parametrize("n_e", [3, 5, 7, 9, 11])
def test_every_compiled_income_grid(n_e):
    _compare(200, n_e, 6, 5)


@pytest.mark.parametrize("n_e", [2, 4, 6, 8, 10])
def test_income_grids_between_the_compiled_ones(n_e):
    """Counts that have no kernel instantiation run on the next compiled count with absorbing zero-mass padding states
    (hank_ctx_create): same policies, distributions and tangents as the oracle run at the true n_e, and the lottery
    brackets of the real states bit-exact."""
    pol_o = _compare(150, n_e, 6, 5)
    s = synthetic(150, n_e, 6, 0)
    blk = make_block(s["m"], 6)
    blk.set_terminal(s["vT"]); blk.set_initial_dist(s["D0"])
    blk.block(s["r"], s["w"])
    m_dev = blk.brackets(3)
    m_api, _ = blk.lottery(blk.policy(3))
    assert m_dev.shape == (n_e, 150) and np.array_equal(m_dev, m_api)
    D = blk.dist(6 - 1)
    assert D.shape == (n_e, 150) and abs(D.sum() - 1.0) < 1e-12      # no mass leaks into the padding states
    blk.close()


def test_two_exogenous_processes_by_kronecker_product():
    """Two independent exogenous dimensions (ForwardIteration.jl:280-284: Λ_exog = kron(Π_2', kron(Π_1', I))) are one
    income dimension with the Kronecker transition: 2 x 3 = 6 states (run on 7), checked against the oracle."""
    from hankb200 import model as M
    from oracle import oracle as O
    z1, P1 = M.rouwenhorst_discretization(3, 0.9, 0.2)
    z2, P2 = M.rouwenhorst_discretization(2, 0.5, 0.1)
    z, Pi = M.kronecker_exogenous([(z1, P1), (z2, P2)])
    assert Pi.shape == (6, 6) and np.allclose(Pi.sum(axis=1), 1.0)
    T, K, n_a = 8, 3, 120
    m = dict(grid=O.double_exponential(n_a, 0.0, 200.0), z=z, Pi=Pi, beta=0.98, gamma=2.0, borrow_cons=0.0)
    P = T - 1
    vT = 1.015 * ((0.015 * m["grid"][None, :] + 1.35 * z[:, None]) + 0.1) ** (-2.0)
    D0 = np.full((6, n_a), 1.0 / (6 * n_a))
    t = np.arange(1, P + 1); r = 0.015 * (1 + 0.1 * 0.9 ** t); w = 1.35 * (1 + 0.05 * 0.9 ** t)
    rng = np.random.default_rng(5); dr = rng.standard_normal((K, P)); dw = rng.standard_normal((K, P))
    orc = make_oracle(m, T)
    pol_o, dpol_o, _, _ = orc.backward(vT, r, w, dr, dw)
    KD_o, dKD_o = orc.forward(D0, pol_o, dpol_o)
    blk = make_block(m, T)
    blk.set_terminal(vT); blk.set_initial_dist(D0)
    KD, dKD = blk.block(r, w, dr, dw)
    assert close(blk.policies(0), pol_o) and close(KD, KD_o) and close(dKD, dKD_o), maxerr(dKD, dKD_o)
    blk.close()


@pytest.mark.parametrize("gamma", [1.0, 1.5, 2.0, 3.0, 4.5])
def test_risk_aversion_values(gamma):
    _compare(150, 7, 8, 2, gamma=gamma)


# ---------------------------------------------------------------- full-size properties
def _ks_block(fixture):
    g = np.load(os.path.join(ROOT, "tests", "golden", fixture))
    T = int(g["T"]); P = T - 1
    m = dict(grid=g["grid"], z=g["z"], Pi=g["Pi"], beta=float(g["beta"]), gamma=float(g["gamma"]), borrow_cons=float(g["borrow_cons"]))
    blk = make_block(m, T)
    blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"])
    blk.ks_configure(float(g["alpha"]), float(g["delta"]), float(g["ss_vars"][1]))
    x0 = np.tile(g["ss_vars"][:4], P); Z = 1.0 + 0.8 ** np.arange(1, P + 1)
    return blk, x0, Z, P


def test_full_size_properties_500x7_T300():
    blk, x0, Z, P = _ks_block("ss_500x7_T300.npz")
    n = 4 * P
    rng = np.random.default_rng(3)
    F = blk.linearize(x0, Z)
    for t in (1, P // 2, P):
        D = blk.dist(t)
        assert abs(D.sum() - 1.0) < 1e-12 and D.min() >= 0.0          # mass conservation, positivity
    U = rng.standard_normal((2, n))
    a, b = 0.7, -1.3
    JV = blk.jvp(np.vstack([U, (a * U[0] + b * U[1])[None]]))
    assert close(JV[2], a * JV[0] + b * JV[1], rtol=1e-9, atol=1e-11)   # linearity
    v = U[0] / np.linalg.norm(U[0]); h = 1e-6
    fd = (blk.linearize(x0 + h * v, Z) - blk.linearize(x0 - h * v, Z)) / (2 * h)
    blk.linearize(x0, Z)
    jv = blk.jvp(v)[0]
    assert np.linalg.norm(fd - jv) / np.linalg.norm(jv) < 1e-6          # JVP vs central differences
    cols = [3, 4, n // 2 + 2, n - 1, n]                                 # 1-based columns
    E = np.zeros((len(cols), n)); E[np.arange(len(cols)), np.array(cols) - 1] = 1.0
    Jc = np.stack([blk.jacobian_columns(c, c + 1)[:, 0] for c in cols])
    assert close(Jc, blk.jvp(E))                                        # unit-seed columns == generic JVPs
    # Y and KS columns have no household term: only the direct residual entries
    jy = blk.jacobian_columns(1, 2)[:, 0]
    assert jy[0] == 1.0 and np.count_nonzero(jy) == 1
    # Newton path converges on the full-size problem (batched mode; the sequential mode is covered at small T)
    blk.linearize(x0, np.ones(P))
    Jbar = blk.jacobian_columns(1, n + 1)
    x, st = blk.newton_solve(Jbar, x0, Z, solver="lu_batched")
    assert st["outer"] == 5 and st["inner"] == [38, 48, 47, 38, 22]     # SURVEY Appendix C probe counts
    assert np.linalg.norm(blk.linearize(x, Z)) < 1e-8
    blk.close()


def test_side_stream_overlap_does_not_change_results(monkeypatch):
    """hank_ks_linearize runs the forward primal sweep on a side stream under the backward tangent;
    HANK_NO_OVERLAP=1 serialises it.  Same kernels, same results bit for bit."""
    blk, x0, Z, P = _ks_block("ss_500x7_T300.npz")
    V = np.random.default_rng(5).standard_normal((9, 4 * P))
    F = blk.linearize(x0, Z); JV = blk.jvp(V)
    blk.close()
    monkeypatch.setenv("HANK_NO_OVERLAP", "1")
    blk2, _, _, _ = _ks_block("ss_500x7_T300.npz")
    F2 = blk2.linearize(x0, Z); JV2 = blk2.jvp(V)
    blk2.close()
    assert np.array_equal(F, F2) and np.array_equal(JV, JV2)


@pytest.mark.parametrize("K", [700, 592])
def test_pipelined_linearisation_does_not_change_results(K, monkeypatch):
    """After a many-lane pass, hank_ks_linearize puts the backward primal sweep on the side stream and lets the next
    many-lane backward tangent sweep follow its per-period progress counters; after a multi-wave pass (K = 700) both
    primal sweeps are one launch (k_primal_ds_both), after a one-wave pass (K = 592) two.  Three linearisations at DIFFERENT points (a tangent sweep that ran ahead of the primal would
    read the previous point's tape), many-lane and few-lane passes in between; HANK_NO_PIPE=1 serialises.  Same
    kernels per point, same results bit for bit."""
    rng = np.random.default_rng(17)
    def run():
        blk, x0, Z, P = _ks_block("ss_500x7_T300.npz")
        V = rng.standard_normal((K, 4 * P)); V1 = V[:3]
        out = []
        for i in range(3):
            x = x0 * (1.0 + 1e-3 * i * np.sin(np.arange(x0.size)))
            out.append(blk.linearize(x, Z)); out.append(blk.jvp(V))
            if i == 1:
                out.append(blk.jvp(V1))                  # a few-lane (row-split) pass at the same linearisation
        out.append(blk.linearize(x0, Z)); out.append(blk.jvp(V1))   # few lanes straight after a (fused) linearisation
        out.append(blk.policies(0)); out.append(blk.dist(P))
        out.append(blk.jvp(V))                                       # many lanes again: the next linearisation is pipelined,
        out.append(blk.linearize(x0 * (1.0 + 2e-3 * np.cos(np.arange(x0.size))), Z))
        out.append(blk.jacobian_columns(1, 4 * P + 1))               # and a seed-horizon pass (+ overflow lanes) follows it
        blk.close()
        return out
    a = run()
    rng = np.random.default_rng(17)
    monkeypatch.setenv("HANK_NO_PIPE", "1")
    b = run()
    assert len(a) == len(b)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    assert not np.array_equal(a[0], a[2])                # (the points really differ)


def test_jacobian_seed_horizons_do_not_change_columns(monkeypatch):
    """hank_ks_jacobian_columns starts each group of unit-seed lanes at its seed period and stages zeros
    beyond it (policy tangents there are exactly zero); HANK_NO_SKIP=1 sweeps every period instead."""
    blk, x0, Z, P = _ks_block("ss_500x7_T300.npz")
    n = 4 * P
    blk.linearize(x0, Z)
    J = blk.jacobian_columns(1, n + 1)
    Jsub = blk.jacobian_columns(403, 431)
    blk.close()
    monkeypatch.setenv("HANK_NO_SKIP", "1")
    blk2, _, _, _ = _ks_block("ss_500x7_T300.npz")
    blk2.linearize(x0, Z)
    J2 = blk2.jacobian_columns(1, n + 1)
    cols = [3, 4, 7, n // 2 + 3, n - 1, n]
    E = np.zeros((len(cols), n)); E[np.arange(len(cols)), np.array(cols) - 1] = 1.0
    JE = blk2.jvp(E)
    blk2.close()
    assert np.all(np.isfinite(J))
    assert close(J, J2), maxerr(J, J2)
    assert close(Jsub, J2[:, 402:430]), maxerr(Jsub, J2[:, 402:430])
    assert close(J[:, np.array(cols) - 1].T, JE)


@pytest.mark.parametrize("fixture,switch", [("ss_500x7_T300.npz", "HANK_NO_TMA"), ("ss_1000x7_T300.npz", None),
                                            ("ss_1000x7_T300.npz", "HANK_NO_TMA")])
def test_jacobian_seed_horizons_other_kernels(fixture, switch, monkeypatch):
    """Seed horizons in the register-prefetch kernels and in the two-rows-per-thread shapes: a column range
    from the middle of the horizon equals generic JVPs of the same unit seeds (which sweep every period)."""
    if switch:
        monkeypatch.setenv(switch, "1")
    blk, x0, Z, P = _ks_block(fixture)
    n = 4 * P
    blk.linearize(x0, Z)
    lo, hi = 2 * P + 1, 2 * P + 41                      # 40 columns, 20 household lanes
    Jc = blk.jacobian_columns(lo, hi)
    E = np.zeros((hi - lo, n)); E[np.arange(hi - lo), np.arange(lo - 1, hi - 1)] = 1.0
    JE = blk.jvp(E)
    blk.close()
    assert np.all(np.isfinite(Jc))
    assert close(Jc.T, JE), maxerr(Jc.T, JE)


def test_fjvp_multi_wave_matches_jvp():
    """hank_ks_fjvp cuts a multi-wave pass at CTA-wave boundaries and downloads finished waves on the copy
    stream; every column must equal the one-pass hank_ks_jvp result."""
    blk, x0, Z, P = _ks_block("ss_500x7_T300.npz")
    n = 4 * P
    rng = np.random.default_rng(11)
    for K in (570, 1190):                      # one wave + a few lanes; two waves + a few lanes
        V = rng.standard_normal((K, n))
        F = blk.linearize(x0, Z)
        JV = blk.jvp(V)
        F2, JV2 = blk.fjvp(x0, Z, V)
        assert np.array_equal(F2, F)
        assert close(JV2, JV), maxerr(JV2, JV)
        assert np.all(np.isfinite(JV2))
    blk.close()


def test_full_size_properties_2000x11_T500():
    s = synthetic(2000, 11, 500, 2)
    blk = make_block(s["m"], 500)
    blk.set_terminal(s["vT"]); blk.set_initial_dist(s["D0"])
    KD, dKD = blk.block(s["r"], s["w"], s["dr"], s["dw"])
    for t in (1, 250, 499):
        assert abs(blk.dist(t).sum() - 1.0) < 1e-12
    pol = blk.policy(250)
    assert np.all(np.diff(pol, axis=1) >= 0) and pol.min() >= 0.0 and pol.max() <= s["m"]["grid"][-1]   # monotone, in range
    h = 1e-6
    KDp, _ = blk.block(s["r"] + h * s["dr"][0], s["w"] + h * s["dw"][0])
    KDm, _ = blk.block(s["r"] - h * s["dr"][0], s["w"] - h * s["dw"][0])
    fd = (KDp - KDm) / (2 * h)
    assert np.linalg.norm(fd - dKD[0]) / np.linalg.norm(fd) < 1e-6
    # linearity across lanes: lane 3 = 2*lane 1 - lane 2
    dr3 = np.vstack([s["dr"], 2 * s["dr"][0] - s["dr"][1]]); dw3 = np.vstack([s["dw"], 2 * s["dw"][0] - s["dw"][1]])
    _, d3 = blk.block(s["r"], s["w"], dr3, dw3)
    assert close(d3[2], 2 * d3[0] - d3[1], rtol=1e-9, atol=1e-11)
    blk.close()


def test_lane_counts_in_any_order_at_one_linearisation():
    """Passes of different lane counts at the SAME linearisation use different cluster shapes, hence different
    row-block layouts of the tape copy: every order must give the same columns (regression: a K = 1 pass after a
    K = 64 pass read the 256-row layout with the 64-row kernel and trapped)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "ss_500x7_T300.npz"))
    from hankb200 import HouseholdBlock
    T = int(g["T"]); P = T - 1; n = 4 * P
    blk = HouseholdBlock(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), T)
    blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"])
    blk.ks_configure(float(g["alpha"]), float(g["delta"]), float(g["ss_vars"][1]))
    x0 = np.tile(g["ss_vars"][:4], P); Z = 1.0 + 0.8 ** np.arange(1, P + 1)
    V = np.random.default_rng(11).standard_normal((300, n))
    blk.linearize(x0, Z)
    first = {}
    for K in (64, 1, 300, 4, 64, 18, 1, 150):
        JV = blk.jvp(V[:K])
        assert np.all(np.isfinite(JV))
        if K in first:
            assert np.array_equal(JV, first[K]), K
        first[K] = JV
        assert close(JV[0], first[64][0], rtol=1e-9), (K, maxerr(JV[0], first[64][0]))   # lane 0 whatever the shape
    blk.close()
