"""Host logic of the equation compiler (hankb200/equations.py) — no GPU: the bytecode of the reference's own equations
(KrusellSmith.yaml:90-94) and of a model with leads, lags, functions and unary minus, interpreted in NumPy, against
direct evaluation of the equation strings."""
import numpy as np
import pytest

import common  # noqa: F401  (puts the package on sys.path)
from hankb200.equations import EquationProgram, OP_VAR, OP_SUB
from np_equations import padded_xmat, residuals_direct, residuals_bytecode

KS_EQ = ["Y = Z * KS(-1)^α", "r + δ = α * Z * KS(-1)^(α-1)", "w = (1-α) * Z * KS(-1)^α", "KS = KD"]
KS_NAMES = ("Y", "KS", "r", "w", "KD", "Z")
PARAMS = {"α": 0.36, "δ": 0.08, "τ": 0.2}
GOV_EQ = ["Y = Z * KS(-1)^α", "r + δ = α * Z * KS(-1)^(α-1)", "w = (1-τ) * (1-α) * Z * KS(-1)^α", "KS = KD",
          "G = τ * w(+1) - -0.1 * log(Y(-2)) + sqrt(exp(-r)) / 2^-Z"]
GOV_NAMES = ("Y", "KS", "r", "w", "G", "KD", "Z")


def _case(eqs, names, n_endog, seed):
    rng = np.random.default_rng(seed)
    P = 17
    prog = EquationProgram(eqs, names, PARAMS)
    x = rng.uniform(0.5, 2.0, P * n_endog); KD = rng.uniform(1.0, 4.0, P); Z = rng.uniform(0.9, 1.1, (1, P))
    ss0 = rng.uniform(0.5, 2.0, len(names)); ss1 = rng.uniform(0.5, 2.0, len(names))
    xm = padded_xmat(x, KD, Z, n_endog, ss0, ss1, prog.max_lag, prog.max_lead)
    return prog, xm, P


def test_krusell_smith_program():
    prog, xm, P = _case(KS_EQ, KS_NAMES, 4, 0)
    assert (prog.max_lag, prog.max_lead, prog.n_eq) == (1, 0, 4)
    eq_off, code, consts = prog.arrays()
    assert code[eq_off[3]:eq_off[4]].tolist() == [OP_VAR, 1, 0, OP_VAR, 4, 0, OP_SUB]     # KS - KD
    a = residuals_bytecode(prog, xm, prog.max_lag, P); b = residuals_direct(KS_EQ, KS_NAMES, PARAMS, xm, prog.max_lag, P)
    assert np.array_equal(a, b)   # same operations in the same order


def test_leads_lags_functions_and_unary_minus():
    prog, xm, P = _case(GOV_EQ, GOV_NAMES, 5, 1)
    assert (prog.max_lag, prog.max_lead, prog.n_eq) == (2, 1, 5)
    a = residuals_bytecode(prog, xm, prog.max_lag, P); b = residuals_direct(GOV_EQ, GOV_NAMES, PARAMS, xm, prog.max_lag, P)
    assert np.allclose(a, b, rtol=1e-15, atol=0) and np.isfinite(a).all()


def test_rejects_what_the_reference_rejects():
    with pytest.raises(ValueError, match="exactly one '='"):
        EquationProgram(["Y = Z = KS"], KS_NAMES, PARAMS)
    with pytest.raises(ValueError, match="unknown symbol"):
        EquationProgram(["Y = Z * undefined_thing"], KS_NAMES, PARAMS)
    with pytest.raises(ValueError, match="not supported"):
        EquationProgram(["Y = sin(Z)"], KS_NAMES, PARAMS)
