"""CPU tests (-m "not gpu"): pin the C++ oracle against the committed golden files, an independent
numpy restatement, finite differences and structural invariants.  The reference has no golden
vectors and Julia cannot run here (parity unpinned), so these are the strongest checks available."""
import glob
import os

import numpy as np
import pytest

import np_restatement as N
from common import synthetic, make_oracle, close, maxerr, KS
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_grid_and_rouwenhorst():
    g = O.double_exponential(200, 0.0, 200.0)
    assert g[0] == 0.0 and abs(g[-1] - 200.0) < 1e-12 and np.all(np.diff(g) > 0)
    u = np.linspace(0, np.log(1 + np.log(201.0)), 200)
    assert np.allclose(g, np.exp(np.exp(u) - 1) - 1, rtol=1e-14, atol=1e-15)
    z, Pi, D = O.rouwenhorst(7, 0.966, 0.283)
    assert np.allclose(Pi.sum(1), 1.0, atol=1e-15) and np.all(Pi > 0)
    assert np.allclose(D @ Pi, D, atol=1e-15) and abs(D.sum() - 1) < 1e-15
    assert np.allclose(D, [1 / 64, 6 / 64, 15 / 64, 20 / 64, 15 / 64, 6 / 64, 1 / 64], atol=1e-14)
    assert abs(np.dot(z, D) - 1.0) < 1e-15
    assert np.allclose(np.diff(np.log(z)), 2 * 0.283 / np.sqrt(6), atol=1e-14)


def test_julia_pow_rules():
    assert O.jl_pow(3.0, -2.0) == (1.0 / 3.0) * (1.0 / 3.0)      # inv(x)^2 path
    assert O.jl_pow(1.7, 3.0) == 1.7 * 1.7 * 1.7
    assert abs(O.jl_pow(2.5, -0.5) - 2.5 ** -0.5) <= 1e-16
    assert abs(O.jl_pow(1.3, -5.0) / (1.3 ** -5.0) - 1) < 4e-16


@pytest.mark.parametrize("n_a,n_e,T,gamma", [(60, 3, 12, 2.0), (200, 7, 20, 2.0), (150, 7, 15, 1.5), (300, 11, 8, 3.0)])
def test_oracle_matches_numpy_restatement(n_a, n_e, T, gamma):
    s = synthetic(n_a, n_e, T, 2, gamma)
    orc = make_oracle(s["m"], T)
    pol, dpol, v1, dv1 = orc.backward(s["vT"], s["r"], s["w"], s["dr"], s["dw"])
    KD, dKD = orc.forward(s["D0"], pol, dpol)
    pol2, dpol2, v2, dv2 = N.backward(s["m"], s["vT"], s["r"], s["w"], s["dr"], s["dw"])
    KD2, dKD2 = N.forward(s["m"], s["D0"], pol, dpol)
    for a, b in ((pol2, pol), (dpol2, dpol), (v2, v1), (dv2, dv1), (KD2, KD), (dKD2, dKD)):
        assert close(a, b), maxerr(a, b)


def test_lottery_rule_and_mass():
    s = synthetic(120, 7, 6)
    orc = make_oracle(s["m"], 6)
    g = s["m"]["grid"]
    rng = np.random.default_rng(3)
    pol = rng.uniform(-1.0, g[-1] * 1.05, size=(7, 120))
    pol[0] = g; pol[1, 0] = -0.0; pol[1, 1] = g[-1]
    m, om = orc.lottery(pol)
    assert np.array_equal(m, np.searchsorted(g, pol, side="left") + 1)
    assert m[0, 0] == 1 and np.array_equal(m[0, 1:], np.arange(2, 121)) and np.all(om[0] == 1.0)
    assert m[1, 0] == 1 and m[1, 1] == 120 and om[1, 1] == 1.0
    assert np.all(m[pol > g[-1]] == 121)
    D = rng.uniform(size=(7, 120)); D /= D.sum()
    Dn, _ = orc.forward_step(pol, D)
    assert abs(Dn.sum() - 1.0) < 1e-13 and np.all(Dn >= 0)
    Dn2, _ = N.forward_step(s["m"], pol, D)
    assert close(Dn, Dn2)


def test_tangents_match_finite_differences():
    s = synthetic(150, 7, 12, 1)
    orc = make_oracle(s["m"], 12)
    pol, dpol, _, _ = orc.backward(s["vT"], s["r"], s["w"], s["dr"], s["dw"])
    KD, dKD = orc.forward(s["D0"], pol, dpol)
    h = 1e-6
    out = []
    for sg in (+1, -1):
        p2, _, _, _ = orc.backward(s["vT"], s["r"] + sg * h * s["dr"][0], s["w"] + sg * h * s["dw"][0])
        out.append(orc.forward(s["D0"], p2)[0])
    fd = (out[0] - out[1]) / (2 * h)
    assert np.linalg.norm(fd - dKD[0]) / np.linalg.norm(fd) < 1e-6


def test_error_semantics():
    s = synthetic(60, 3, 4)
    orc = make_oracle(s["m"], 4)
    v = s["vT"].copy(); v[1, 10] = -5.0      # negative β·EV under ^(-1/2): Julia DomainError
    with pytest.raises(O.OracleError) as ei:
        orc.egm_step(v, 0.015, 1.35)
    assert ei.value.code == 2
    v = s["vT"].copy(); v[:, 20] *= 1e-3      # non-monotone endogenous grid: Interpolations error
    with pytest.raises(O.OracleError) as ei:
        orc.egm_step(v, 0.015, 1.35)
    assert ei.value.code == 3


def test_gmres_defaults():
    rng = np.random.default_rng(5)
    A = np.eye(60) + 0.1 * rng.standard_normal((60, 60)); b = rng.standard_normal(60)
    x, it = O.gmres(A, np.ones(60), b)
    r0 = np.linalg.norm(b - A @ np.ones(60))
    assert np.linalg.norm(b - A @ x) <= 1.0001 * np.sqrt(np.finfo(float).eps) * r0 and 20 < it < 60
    A = np.diag(np.logspace(0, 6, 40)) + 1e-3 * rng.standard_normal((40, 40))
    x, it = O.gmres(A, np.ones(40), rng.standard_normal(40))
    assert it == 40                            # maxiter = n with restart 20: stalls, like the reference would


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "sweep_*.npz"))))
def test_golden_sweeps(path):
    g = np.load(path)
    m = dict(grid=g["grid"], z=g["z"], Pi=g["Pi"], beta=float(g["beta"]), gamma=float(g["gamma"]), borrow_cons=float(g["borrow_cons"]))
    T = int(g["T"])
    orc = make_oracle(m, T)
    pol, dpol, v1, dv1 = orc.backward(g["vT"], g["r"], g["w"], g["dr"], g["dw"])
    KD, dKD, Dp, dDl = orc.forward(g["D0"], pol, dpol, want_path=True)
    mm, om = orc.lottery(pol[0])
    assert np.array_equal(mm, g["m1"])
    for a, b in ((pol, g["pol"]), (dpol, g["dpol"]), (v1, g["v1"]), (KD, g["KD"]), (dKD, g["dKD"]), (Dp[-1], g["D_last"]), (dDl, g["dD_last"])):
        assert close(a, b, rtol=1e-13, atol=1e-14), maxerr(a, b)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ks_*.npz"))))
def test_golden_ks(path):
    g = np.load(path)
    T = int(g["T"]); P = T - 1
    orc = O.Oracle(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), T)
    ks = (float(g["alpha"]), float(g["delta"]), float(g["ss_vars"][1]))
    F0, JV = orc.ks_fjvp(ks, g["ss_value"], g["ss_D"], g["Z"], g["x0"], g["V"])
    assert close(F0, g["F0"], 1e-13, 1e-14) and close(JV, g["JV"], 1e-13, 1e-14)
    Jc = orc.jacobian(ks, g["ss_value"], g["ss_D"], np.ones(P), g["x0"], g["cols"])
    assert close(Jc, g["Jcols"], 1e-13, 1e-14)
    # Jacobian column == JVP with the unit seed (test_SteadyState.jl:205-224 pattern, at 1e-10 not 1e-5)
    assert close(Jc, g["Jbar"][:, g["cols"]])
    # steady-state residual < 10 eps on the first three equations (test_SteadyState.jl:75-84); the 4th (KS-KD)
    # carries the VFI tolerance
    Fss, _ = orc.ks_fjvp(ks, g["ss_value"], g["ss_D"], np.ones(P), g["x0"])
    assert np.max(np.abs(Fss.reshape(P, 4)[:, :3])) < 1e-5
    x, st = orc.newton(ks, g["ss_value"], g["ss_D"], g["Z"], g["Jbar"], g["x0"], solver="lu")
    assert st["inner"] == list(g["newton_inner"]) and close(x, g["x_newton"], 1e-12, 1e-13)
    Fx, _ = orc.ks_fjvp(ks, g["ss_value"], g["ss_D"], g["Z"], x)
    assert np.linalg.norm(Fx) < 1e-8


def test_steady_state_matches_survey_probe():
    from oracle.steady_state import find_ss
    ss, orc, info = find_ss(200, 7, 150)
    # SURVEY.md Appendix C (restatement probe): r / w / Y / KS at 200x7
    assert abs(ss.vars["r"] - 0.0152588) < 1e-6 and abs(ss.vars["w"] - 1.351967) < 1e-5
    assert abs(ss.vars["Y"] - 2.112448) < 1e-5 and abs(ss.vars["KS"] - 7.983319) < 1e-5
    assert info["resnorm"] < 1e-6 and abs(ss.D.sum() - 1) < 1e-12
    assert int((ss.policy <= 0).sum()) == 5 and ss.policy.max() == 200.0
