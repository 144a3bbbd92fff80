"""Equilibrium equations as device bytecode (csrc/hank_eq.cu, hank_eq_configure) — the reference's compile_residuals +
assemble_full_xMat + the ForwardDiff pass through them (ModelParser.jl:217-259, GeneralStructures.jl:329-377).

(1) The reference's own four equations compiled to bytecode reproduce the built-in Krusell-Smith block: F to the last
    bit or two (same operations in the same order; the built-in kernel is FMA-contracted), JVPs / Jacobian columns to rounding, the Newton path with the same inner counts.
(2) A five-equation model with a tax, a lead, a two-period lag, log / exp / sqrt and a unary minus: F against direct
    NumPy evaluation of the equation strings on the padded variable matrix, JVPs against central finite differences of
    the device's own F, Jacobian columns against unit-seed JVPs, and a Newton solve that converges to F = 0."""
import os

import numpy as np
import pytest

from common import close, maxerr
from hankb200 import HouseholdBlock
from hankb200.equations import EquationProgram
from np_equations import padded_xmat, residuals_direct

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KS_EQ = ["Y = Z * KS(-1)^α", "r + δ = α * Z * KS(-1)^(α-1)", "w = (1-α) * Z * KS(-1)^α", "KS = KD"]
KS_NAMES = ("Y", "KS", "r", "w", "KD", "Z")


def _block(g):
    blk = HouseholdBlock(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), int(g["T"]))
    blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"])
    return blk


def test_krusell_smith_equations_as_bytecode_match_the_builtin_block():
    g = np.load(os.path.join(GOLD, "ks_100x3_T30.npz"))
    params = {"α": float(g["alpha"]), "δ": float(g["delta"])}
    ref = _block(g); ref.ks_configure(params["α"], params["δ"], float(g["ss_vars"][1]))
    gen = _block(g); gen.eq_configure(EquationProgram(KS_EQ, KS_NAMES, params), 4, 2, 3, g["ss_vars"], g["ss_vars"])
    n = ref.n
    F1 = ref.linearize(g["x0"], g["Z"]); F2 = gen.linearize(g["x0"], g["Z"])
    assert np.abs(F1 - F2).max() <= 4e-16 * max(1.0, np.abs(F1).max())   # same operations; the built-in kernel contracts a*b-c to FMAs
    JV1 = ref.jvp(g["V"]); JV2 = gen.jvp(g["V"])
    assert np.abs(JV1 - JV2).max() <= 1e-13 * np.abs(JV1).max()
    J1 = ref.jacobian_columns(1, n + 1); J2 = gen.jacobian_columns(1, n + 1)
    assert np.abs(J1 - J2).max() <= 1e-13 * np.abs(J1).max()
    cols = np.array([2, 3, 4, 7, 60, n - 1, n], dtype=np.int32)
    assert np.array_equal(gen.jacobian_column_list(cols), J2[:, cols - 1])
    for solver in ("lu", "lu_batched"):
        x1, s1 = ref.newton_solve(g["Jbar"], g["x0"], g["Z"], solver=solver)
        x2, s2 = gen.newton_solve(g["Jbar"], g["x0"], g["Z"], solver=solver)
        assert s1["inner"] == s2["inner"] == list(g["newton_inner"]) and close(x2, x1), (solver, maxerr(x2, x1))
    with pytest.raises(Exception, match="hank_eq_configure"):
        gen.ks_configure(0.36, 0.08, 1.0)
    ref.close(); gen.close()


def test_model_with_tax_lead_lag_and_functions():
    g = np.load(os.path.join(GOLD, "ks_100x3_T30.npz"))
    P = int(g["T"]) - 1
    params = {"α": float(g["alpha"]), "δ": float(g["delta"]), "τ": 0.1}
    names = ("Y", "KS", "r", "w", "G", "KD", "Z")
    eqs = ["Y = Z * KS(-1)^α", "r + δ = α * Z * KS(-1)^(α-1)", "w = (1-τ) * (1-α) * Z * KS(-1)^α", "KS = KD",
           "G = τ * w(+1) - -0.1 * log(Y(-2)) + sqrt(exp(-r)) / 2^Z"]
    prog = EquationProgram(eqs, names, params)
    assert (prog.max_lag, prog.max_lead) == (2, 1)
    Y, KS, r, w, KDs, Zs = g["ss_vars"]
    ss = np.array([Y, KS, r, w, 0.3, KDs, Zs])
    blk = _block(g); blk.eq_configure(prog, 5, 2, 3, ss, ss)
    n = blk.n
    assert n == 5 * P
    rng = np.random.default_rng(3)
    x = np.tile(ss[:5], P) * (1.0 + 0.01 * rng.standard_normal(n))
    Z = np.asarray(g["Z"])
    F = blk.linearize(x, Z)
    # the household aggregate of the same r, w paths from a second context (same kernels), then the equations in NumPy
    hh = _block(g)
    X = x.reshape(P, 5)
    KD, _ = hh.block(X[:, 2].copy(), X[:, 3].copy())
    xm = padded_xmat(x, KD, Z[None, :], 5, ss, ss, prog.max_lag, prog.max_lead)
    F_np = residuals_direct(eqs, names, params, xm, prog.max_lag, P)
    assert close(F, F_np), maxerr(F, F_np)
    # JVPs: linear in V, and equal to central differences of F
    V = rng.standard_normal((3, n))
    JV = blk.jvp(V)
    assert close(blk.jvp(2.0 * V[0] - V[1])[0], 2.0 * JV[0] - JV[1], rtol=1e-9)
    h = 1e-6
    v = V[2] / np.linalg.norm(V[2])
    fd = (blk.linearize(x + h * v, Z) - blk.linearize(x - h * v, Z)) / (2 * h)
    blk.linearize(x, Z)
    jv = blk.jvp(v)[0]
    assert np.linalg.norm(fd - jv) / np.linalg.norm(jv) < 1e-6
    # Jacobian columns == unit-seed JVPs (every variable kind, first / middle / last period)
    cols = np.array([1, 2, 3, 4, 5, 5 * 7 + 3, 5 * 7 + 5, n - 2, n], dtype=np.int32)
    Jc = blk.jacobian_column_list(cols)
    E = np.zeros((len(cols), n)); E[np.arange(len(cols)), cols - 1] = 1.0
    JE = blk.jvp(E)
    assert close(Jc.T, JE, rtol=1e-9), maxerr(Jc.T, JE)
    # Newton: the Jacobian at the steady-state path as preconditioner; converges to a root of the five equations
    x0 = np.tile(ss[:5], P)
    blk.linearize(x0, np.ones(P))
    Jbar = blk.jacobian_columns(1, n + 1)
    xs, st = blk.newton_solve(Jbar, x0, Z, solver="lu_batched")
    assert np.linalg.norm(blk.linearize(xs, Z)) < 1e-7 and st["outer"] < 30
    blk.close(); hh.close()


def test_bad_programs_are_rejected_on_the_host():
    from hankb200 import HankError
    g = np.load(os.path.join(GOLD, "ks_100x3_T30.npz"))
    blk = _block(g)
    prog = EquationProgram(KS_EQ, KS_NAMES, {"α": 0.36, "δ": 0.08})
    prog.code[0] = 99
    with pytest.raises(HankError, match="opcode"):
        blk.eq_configure(prog, 4, 2, 3, g["ss_vars"], g["ss_vars"])
    prog = EquationProgram(KS_EQ, KS_NAMES, {"α": 0.36, "δ": 0.08})
    prog.code.append(2)   # a dangling '+' after the last equation's program
    prog.eq_off[-1] += 1
    with pytest.raises(HankError, match="underflow|one value"):
        blk.eq_configure(prog, 4, 2, 3, g["ss_vars"], g["ss_vars"])
    blk.close()
