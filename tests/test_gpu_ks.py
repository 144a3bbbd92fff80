"""GPU parity of the Krusell-Smith full function, JVPs, Jacobian columns and the Newton driver against
the oracle and the committed golden files (all through the C ABI)."""
import glob
import os

import numpy as np
import pytest

from common import close, maxerr, make_block, make_oracle, synthetic
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _block_from_golden(g):
    m = dict(grid=g["grid"], z=g["z"], Pi=g["Pi"], beta=float(g["beta"]), gamma=float(g["gamma"]),
             borrow_cons=float(g["borrow_cons"]))
    T = int(g["T"])
    blk = make_block(m, T)
    return m, T, blk


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "sweep_*.npz"))))
def test_golden_sweeps_gpu(path):
    g = np.load(path)
    m, T, blk = _block_from_golden(g)
    blk.set_terminal(g["vT"]); blk.set_initial_dist(g["D0"])
    KD, dKD = blk.block(g["r"], g["w"], g["dr"], g["dw"])
    assert close(blk.policies(0), g["pol"]) and close(KD, g["KD"]) and close(dKD, g["dKD"])
    for l in range(int(g["K"])):
        assert close(blk.policies(l + 1), g["dpol"][l]), maxerr(blk.policies(l + 1), g["dpol"][l])
    assert close(blk.dist(T - 1), g["D_last"])
    # end-to-end brackets: count flips instead of asserting zero (1-ulp pow differences)
    flips = int(np.sum(blk.brackets(1) != g["m1"]))
    assert flips <= 1, flips
    # identical policy inputs: bit-exact brackets and weights
    mm, om = blk.lottery(g["pol"][0])
    assert np.array_equal(mm, g["m1"]) and np.array_equal(om, g["om1"])
    # caller-supplied policies (ForwardIteration(policy_seqs, ...) entry)
    KD2, dKD2 = blk.forward_policies(g["pol"], g["dpol"])
    assert close(KD2, g["KD"]) and close(dKD2, g["dKD"])
    blk.close()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ks_*.npz"))))
def test_golden_ks_gpu(path):
    g = np.load(path)
    m, T, blk = _block_from_golden(g)
    P = T - 1; n = 4 * P
    blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"])
    blk.ks_configure(float(g["alpha"]), float(g["delta"]), float(g["ss_vars"][1]))
    F0 = blk.linearize(g["x0"], g["Z"])
    assert close(F0, g["F0"]), maxerr(F0, g["F0"])
    JV = blk.jvp(g["V"])
    assert close(JV, g["JV"]), maxerr(JV, g["JV"])
    F2, JV2 = blk.fjvp(g["x0"], g["Z"], g["V"])           # fused host entry
    assert np.array_equal(F2, F0) and np.array_equal(JV2, JV)
    # Jacobian columns at the steady-state path (Z = 1): unit-seed JVPs
    blk.linearize(g["x0"], np.ones(P))
    J = blk.jacobian_columns(1, n + 1)
    assert close(J, g["Jbar"]), maxerr(J, g["Jbar"])
    Jc = blk.jacobian_columns(5, 12)
    assert close(Jc, g["Jbar"][:, 4:11])
    # unit-seed columns agree with generic JVPs of the same seeds (test_SteadyState.jl:205-224 pattern)
    cols = g["cols"]
    E = np.zeros((len(cols), n)); E[np.arange(len(cols)), cols] = 1.0
    assert close(blk.jvp(E).T, g["Jbar"][:, cols])
    # Newton path, LU preconditioner solve: same iteration counts, same converged path
    x, st = blk.newton_solve(g["Jbar"], g["x0"], g["Z"], solver="lu")
    assert st["inner"] == list(g["newton_inner"]), (st, list(g["newton_inner"]))
    assert close(x, g["x_newton"]), maxerr(x, g["x_newton"])
    # batched mode: J(x) assembled once per outer iteration from unit-seed lanes, inner J(x)y by GEMV
    xb, sb = blk.newton_solve(g["Jbar"], g["x0"], g["Z"], solver="lu_batched")
    assert sb["inner"] == st["inner"], (sb, st)
    assert close(xb, x), maxerr(xb, x)
    blk.close()


def test_newton_gmres_matches_oracle():
    """Reference-faithful inner solver (restarted GMRES, IterativeSolvers defaults) on a small horizon."""
    g = np.load(os.path.join(GOLD, "ks_100x3_T30.npz"))
    m, T, blk = _block_from_golden(g)
    P = T - 1
    orc = O.Oracle(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), T)
    ks = (float(g["alpha"]), float(g["delta"]), float(g["ss_vars"][1]))
    xo, so = orc.newton(ks, g["ss_value"], g["ss_D"], g["Z"], g["Jbar"], g["x0"], solver="gmres")
    blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"])
    blk.ks_configure(*ks)
    x, st = blk.newton_solve(g["Jbar"], g["x0"], g["Z"], solver="gmres")
    assert st["outer"] == so["outer"] and st["inner"] == so["inner"], (st, so)
    # both runs stop at ||y|| <= 1e-9; they agree far inside that
    assert np.max(np.abs(x - xo)) < 1e-9 and close(x, xo, rtol=1e-8, atol=1e-10)
    blk.close()


def test_error_semantics_gpu():
    from hankb200 import HankError
    s = synthetic(60, 3, 4)
    blk = make_block(s["m"], 4)
    v = s["vT"].copy(); v[1, 10] = -5.0
    with pytest.raises(HankError) as ei:
        blk.egm_step(v, 0.015, 1.35)
    assert ei.value.code == 2 and "DomainError" in ei.value.msg
    v = s["vT"].copy(); v[:, 20] *= 1e-3
    with pytest.raises(HankError) as ei:
        blk.egm_step(v, 0.015, 1.35)
    assert ei.value.code == 3
    # the context stays usable after an error
    val, pol, _, _ = blk.egm_step(s["vT"], 0.015, 1.35)
    vo, po, _, _ = make_oracle(s["m"], 4).egm_step(s["vT"], 0.015, 1.35)
    assert close(val, vo) and close(pol, po)
    with pytest.raises(HankError):
        blk.forward()          # call order: forward before backward
    # a policy that is not monotone in a (the reference accepts any policy): scatter path for the primal,
    # rejected when tangent lanes ride along
    pol3 = np.tile(pol[None], (3, 1, 1)); pol3[1, 0, 5] = 150.0; pol3[2, 1, 40] = 0.0
    blk.set_initial_dist(s["D0"])
    KD3, _ = blk.forward_policies(pol3)
    orc3 = make_oracle(s["m"], 4)
    KD3o, _, Dp3o, _ = orc3.forward(s["D0"], pol3, want_path=True)
    assert close(KD3, KD3o), maxerr(KD3, KD3o)
    assert close(blk.dist(3), Dp3o[2]) and np.array_equal(blk.brackets(2), orc3.lottery(pol3[1])[0])
    with pytest.raises(HankError) as ei:
        blk.forward_policies(pol3, np.zeros((1,) + pol3.shape))
    assert ei.value.code == 6
    blk.close()


@pytest.mark.parametrize("n_a,n_e,gamma,K", [(60, 3, 2.0, 4), (500, 7, 2.0, 4), (300, 7, 1.5, 2), (2000, 11, 2.0, 1)])
def test_egm_step_lanes(n_a, n_e, gamma, K):
    """value_fn plug-in with ForwardDiff lanes on the incoming value (find_ss's VFI pattern)."""
    s = synthetic(n_a, n_e, 4, K, gamma)
    orc = make_oracle(s["m"], 4)
    rng = np.random.default_rng(11)
    dv = rng.standard_normal((K, n_e, n_a)) * 0.01
    dr = rng.standard_normal(K); dw = rng.standard_normal(K)
    o = orc.egm_step(s["vT"], 0.02, 1.3, dv, dr, dw)
    blk = make_block(s["m"], 4)
    b = blk.egm_step(s["vT"], 0.02, 1.3, dv, dr, dw)
    for x, y in zip(b, o):
        assert close(x, y), maxerr(x, y)
    blk.close()


def test_newton_rejects_singular_jbar():
    """A singular preconditioner must be reported by the LU (getrf info), not surface as a diverged iteration."""
    from hankb200 import HankError
    g = np.load(os.path.join(GOLD, "ks_100x3_T30.npz"))
    m, T, blk = _block_from_golden(g)
    blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"])
    blk.ks_configure(float(g["alpha"]), float(g["delta"]), float(g["ss_vars"][1]))
    J = np.array(g["Jbar"]); J[:, 3] = 0.0; J[:, 7] = 0.0      # exactly singular: zero pivots in U
    with pytest.raises(HankError) as ei:
        blk.newton_solve(J, g["x0"], g["Z"], solver="lu")
    assert "singular" in ei.value.msg, ei.value.msg
    # the context (and the preconditioner cache) recover: same J̅ twice gives the same path, second call cached
    x1, s1 = blk.newton_solve(g["Jbar"], g["x0"], g["Z"], solver="lu")
    x2, s2 = blk.newton_solve(g["Jbar"], g["x0"], g["Z"], solver="lu")
    assert np.array_equal(x1, x2) and s1["inner"] == s2["inner"] == list(g["newton_inner"])
    blk.close()


def test_relinearize_back_to_back_is_race_free():
    """Two linearisations in a row with no tangent pass in between (zero-iteration Newton step, device-API users):
    the second must not overwrite x / Z while the first one's side-stream forward sweep still reads them."""
    g = np.load(os.path.join(GOLD, "ks_200x7_T40.npz"))
    m, T, blk = _block_from_golden(g)
    P = T - 1
    blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"])
    blk.ks_configure(float(g["alpha"]), float(g["delta"]), float(g["ss_vars"][1]))
    Fa = blk.linearize(g["x0"], g["Z"])
    xb = g["x0"] * (1.0 + 1e-3 * np.cos(np.arange(4 * P)))
    Fb = blk.linearize(xb, np.ones(P))
    for _ in range(5):
        assert np.array_equal(blk.linearize(g["x0"], g["Z"]), Fa)
        assert np.array_equal(blk.linearize(xb, np.ones(P)), Fb)
    blk.close()
