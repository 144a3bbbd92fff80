"""GPU tests of the host mirror of the reference's model API: YAML -> model -> steady state (device VFI) ->
Jacobian -> Newton, compared with the oracle's restatement of the same chain."""
import os

import numpy as np
import pytest

import yaml

from common import close, maxerr, ks_yaml_dict

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def YAML(tmp_path):
    p = tmp_path / "KrusellSmith.yaml"
    p.write_text(yaml.safe_dump(ks_yaml_dict(), allow_unicode=True), encoding="utf-8")
    return str(p)


def test_grids_match_oracle():
    from hankb200 import model as M
    from oracle import oracle as O
    assert np.allclose(M.double_exponential(200, 0.0, 200.0), O.double_exponential(200, 0.0, 200.0), rtol=1e-14, atol=1e-16)
    z, Pi = M.rouwenhorst_discretization(7, 0.966, 0.283)
    zo, Pio, _ = O.rouwenhorst(7, 0.966, 0.283)
    assert np.allclose(z, zo, rtol=1e-15, atol=0) and np.allclose(Pi, Pio, rtol=1e-15, atol=1e-18)


def test_vfi_matches_oracle(YAML):
    from hankb200 import model as M
    from oracle.steady_state import get_xvals, exog_kron
    from oracle import oracle as O
    mod = M.build_model_from_yaml(YAML, {"wealth.n": 100, "productivity.n": 3, "T": 30})
    blk = mod.household_block(T=2)
    w, pr = mod.heterogeneity["wealth"], mod.heterogeneity["productivity"]
    orc = O.Oracle(w.grid, pr.grid, pr.transition, 0.98, 2.0, 0.0, 2)
    r, wg = 0.02, 1.3
    dr = np.array([0, 0, 1.0, 0]); dw = np.array([0, 0, 0, 1.0])
    val, pol, dval, dpol, steps = blk.vfi(r, wg, dr, dw, eps=1e-6)
    # oracle: the same loop (oracle/steady_state.py get_xvals), re-run here to expose the lanes
    v = np.ones((3, 100)); dv = np.zeros((4, 3, 100))
    res = orc.egm_step(v, r, wg, dv, dr, dw); n = 0
    for _ in range(10_000):
        tol = np.max(np.abs(res[0] - v)); v, dv = res[0], res[2]
        if tol < 1e-6:
            break
        res = orc.egm_step(v, r, wg, dv, dr, dw); n += 1
    assert steps == n
    for a, b in ((val, res[0]), (pol, res[1]), (dval, res[2]), (dpol, res[3])):
        assert close(a, b), maxerr(a, b)
    blk.close()


def test_yaml_to_newton_chain(YAML):
    from hankb200 import model as M
    from hankb200.steady_state import get_SteadyStates, find_ss
    from hankb200.newton import TransitionProblem, JVP, directJVPJacobian, NewtonRaphsonHANK
    g = np.load(os.path.join(ROOT, "tests", "golden", "ks_100x3_T30.npz"))
    mod = M.build_model_from_yaml(YAML, {"wealth.n": 100, "productivity.n": 3, "T": 30})
    assert mod.var_names == ("Y", "KS", "r", "w", "KD", "Z") and mod.compspec.n_v == 6 and mod.compspec.n_endog == 4
    ss, info = find_ss(mod, mod.ss_initial, "initial")
    assert info["iterations"] == int(g["ss_iterations"]) and info["resnorm"] < 1e-6
    ref = dict(zip(("Y", "KS", "r", "w", "KD", "Z"), g["ss_vars"]))
    for k in ref:
        assert abs(ss.vars[k] - ref[k]) <= 1e-9 * max(1.0, abs(ref[k])), (k, ss.vars[k], ref[k])
    assert close(ss.value, g["ss_value"], rtol=1e-8, atol=1e-10) and close(ss.D, g["ss_D"].reshape(-1), rtol=1e-7, atol=1e-10)
    # residual norm at the steady state < 10 eps on the aggregate equations (test_SteadyState.jl:75-84)
    P = mod.compspec.T - 1
    prob = TransitionProblem(mod, ss, ss, {"Z": np.ones(P)})
    Fss = prob.fullFunction(prob.x_steady())
    assert np.max(np.abs(Fss.reshape(P, 4)[:, :3])) < 10 * mod.compspec.eps
    # Jacobian column == JVP with the unit seed (test_SteadyState.jl:205-224)
    J = directJVPJacobian(prob)
    n = prob.n
    for j in (0, 1, n // 2, n - 2, n - 1):
        e = np.zeros(n); e[j] = 1.0
        assert close(JVP(prob, prob.x_steady(), e), J[:, j])
    prob2 = TransitionProblem(mod, ss, ss, {"Z": 1.0 + 0.8 ** np.arange(1, P + 1)}, blk=prob.blk)
    x, st = NewtonRaphsonHANK(prob2.x_steady(), J, prob2, solver="lu", verbose=False)
    assert st["outer"] <= 8 and np.linalg.norm(prob2.fullFunction(x)) < 1e-8
    assert np.max(np.abs(x - g["x_newton"])) < 1e-7   # fixture steady state differs at the VFI tolerance
    prob.blk.close()
