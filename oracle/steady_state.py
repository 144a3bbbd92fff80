"""Steady-state solve of the reference, restated for the oracle — TEST INFRASTRUCTURE.

Follows SteadyState.jl:111-233 (get_xVals / find_ss) and ForwardIteration.jl:436-442, :480-558
(invariant_dist and its ForwardDiff overload) for the Krusell-Smith model of KrusellSmith.yaml.
The EGM step is the C++ oracle's (4 tangent lanes = ForwardDiff.jacobian's chunk of 4,
SteadyState.jl:195); the sparse solves use scipy's SuperLU where Julia uses UMFPACK.
PARITY UNPINNED (see hank_oracle.cpp header).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .oracle import Oracle, OracleError, double_exponential, rouwenhorst, jl_pow

KS_YAML = dict(beta=0.98, borrow_cons=0.0, gamma=2.0, alpha=0.36, delta=0.08,
               grid_min=0.0, grid_max=200.0, rho=0.966, sigma=0.283, eps=1e-6)


def endog_transition(orc, policy, dpolicy=None):
    """make_endogenous_transition (ForwardIteration.jl:37-78) as CSC; optional lane tangents of
    the non-zeros (same sparsity)."""
    na, ne, G = orc.n_a, orc.n_e, orc.G
    m, om = orc.lottery(policy)
    m = m.reshape(-1); om = om.reshape(-1)
    cols = np.arange(G); e_of = cols // na
    edge = (m == 1) | (m > na)
    rows_lo = e_of * na + np.clip(m - 2, 0, na - 1)
    rows_hi = e_of * na + np.clip(m - 1, 0, na - 1)
    I = np.concatenate([rows_hi[edge], rows_lo[~edge], rows_hi[~edge]])
    J = np.concatenate([cols[edge], cols[~edge], cols[~edge]])
    V = np.concatenate([np.ones(edge.sum()), 1.0 - om[~edge], om[~edge]])
    L = sp.csc_matrix((V, (I, J)), shape=(G, G))
    if dpolicy is None:
        return L
    dLs = []
    g = orc.grid
    dg = np.where(edge, 1.0, g[np.clip(m - 1, 0, na - 1)] - g[np.clip(m - 2, 0, na - 1)])
    for dp in dpolicy:
        dom = np.where(edge, 0.0, dp.reshape(-1) / dg)
        dV = np.concatenate([np.zeros(edge.sum()), -dom[~edge], dom[~edge]])
        dLs.append(sp.csc_matrix((dV, (I, J)), shape=(G, G)))
    return L, dLs


def exog_kron(orc):
    """Λ_exog = kron(sparse(Π'), I_{n_a}) (SteadyState.jl:86-89)."""
    return sp.kron(sp.csc_matrix(orc.Pi.T), sp.identity(orc.n_a, format="csc"), format="csc")


def invariant_dist(Lam, dLams=None):
    """invariant_dist(Λ') — ForwardIteration.jl:436-442 and the Dual overload :480-558.
    Lam is the column-stochastic Λ (= ΠT in the reference's naming)."""
    Lam = sp.csc_matrix(Lam)
    n = Lam.shape[0]
    M = (sp.identity(n - 1, format="csc") - Lam[1:, 1:]).tocsc()
    u = np.asarray(Lam[1:, 0].todense()).reshape(-1)
    fac = spla.splu(M)
    y2 = fac.solve(u)
    if dLams is None:
        D = np.concatenate([[1.0], y2])
        return D / D.sum()
    s = 1.0 + y2.sum()
    D0 = np.concatenate([[1.0], y2]) / s
    dDs = []
    for dL in dLams:
        rhs = (dL @ D0)[1:]
        y1 = fac.solve(rhs)
        tail = y1 - y2 * (y1.sum() / s)
        dDs.append(np.concatenate([[-tail.sum()], tail]))
    return D0, dDs


class SteadyState:
    """Mirror of the reference's SteadyState record (SteadyState.jl:21-27)."""

    def __init__(self, vars_, policy, Lam, D, value):
        self.vars = vars_      # dict Y,KS,r,w,KD,Z
        self.policy = policy   # (n_e, n_a)
        self.Lam = Lam
        self.D = D             # (G,)
        self.value = value     # (n_e, n_a)


def get_xvals(orc, Lexog, p, Z, eps, lanes=False):
    """get_xVals (SteadyState.jl:111-154). p = (Y, KS, r, w). With lanes=True carries the 4
    unit tangents of ForwardDiff.jacobian through the VFI and the invariant distribution."""
    Y, KS, r, w = p
    ne, na = orc.n_e, orc.n_a
    K = 4 if lanes else 0
    dr = np.array([0.0, 0.0, 1.0, 0.0]) if lanes else None
    dw = np.array([0.0, 0.0, 0.0, 1.0]) if lanes else None
    value = np.ones((ne, na)); dvalue = np.zeros((K, ne, na)) if lanes else None
    res = orc.egm_step(value, r, w, dvalue, dr, dw)
    for _ in range(10_000):
        value_new, dvalue_new = res[0], res[2]
        tol = np.max(np.abs(value_new - value))
        value, dvalue = value_new, (dvalue_new if lanes else None)
        if tol < eps:
            break
        res = orc.egm_step(value, r, w, dvalue, dr, dw)
    policy, dpolicy = res[1], res[3]
    if lanes:
        Lend, dLend = endog_transition(orc, policy, dpolicy)
        D, dD = invariant_dist(Lexog @ Lend, [Lexog @ d for d in dLend])
        KD = float(np.dot(policy.reshape(-1), D))
        dKDv = np.array([np.dot(D, dpolicy[k].reshape(-1)) + np.dot(policy.reshape(-1), dD[k]) for k in range(4)])
    else:
        Lend = endog_transition(orc, policy)
        D = invariant_dist(Lexog @ Lend)
        KD = float(np.dot(policy.reshape(-1), D)); dKDv = None
    return dict(Y=Y, KS=KS, r=r, w=w, KD=KD, Z=Z), res[0], dKDv


def ks_ss_residual(xv, alpha, delta, dKD=None):
    """Compiled residuals on the all-columns-identical padded matrix (SteadyState.jl:163-169,
    KrusellSmith.yaml:90-94). Returns z (4,) and, if dKD given, the 4x4 Jacobian wrt (Y,KS,r,w)."""
    Y, KS, r, w, KD, Z = (xv[k] for k in ("Y", "KS", "r", "w", "KD", "Z"))
    Ka, Ka1 = jl_pow(KS, alpha), jl_pow(KS, alpha - 1.0)
    z = np.array([Y - (Z * Ka), (r + delta) - ((alpha * Z) * Ka1), w - (((1 - alpha) * Z) * Ka), KS - KD])
    if dKD is None:
        return z
    Ka2 = jl_pow(KS, alpha - 2.0)
    J = np.zeros((4, 4))  # columns: dY, dKS, dr, dw
    dKa = alpha * Ka1; dKa1 = (alpha - 1.0) * Ka2
    J[0, 0] = 1.0; J[0, 1] = -(Z * dKa)
    J[1, 2] = 1.0; J[1, 1] = -((alpha * Z) * dKa1)
    J[2, 3] = 1.0; J[2, 1] = -(((1 - alpha) * Z) * dKa)
    J[3, 1] = 1.0; J[3, :] -= dKD
    return z, J


def find_ss(n_a=200, n_e=7, T=150, Z=1.0, guesses=None, params=None, verbose=False):
    """find_ss (SteadyState.jl:184-233) for the KS model. Returns (SteadyState, Oracle, info)."""
    P_ = dict(KS_YAML); P_.update(params or {})
    guesses = guesses or dict(Y=1.5, KS=3.5, r=0.04, w=1.0)  # KrusellSmith.yaml:103-107
    grid = double_exponential(n_a, P_["grid_min"], P_["grid_max"])
    z, Pi, _ = rouwenhorst(n_e, P_["rho"], P_["sigma"])
    orc = Oracle(grid, z, Pi, P_["beta"], P_["gamma"], P_["borrow_cons"], T)
    Lexog = exog_kron(orc)
    eps, alpha, delta = P_["eps"], P_["alpha"], P_["delta"]
    p = np.array([guesses["Y"], guesses["KS"], guesses["r"], guesses["w"]], dtype=np.float64)

    def F(q):
        xv, _, _ = get_xvals(orc, Lexog, q, Z, eps)
        return ks_ss_residual(xv, alpha, delta)

    def safe_eval(q):
        try:
            return F(q)
        except (OracleError, RuntimeError, FloatingPointError):
            return np.full(4, np.inf)

    zres = F(p)
    it = 0
    while np.linalg.norm(zres) > eps and it < 100:
        if verbose:
            print(f"  [ss] Iteration {it}: residual norm = {np.linalg.norm(zres)}")
        xv, _, dKD = get_xvals(orc, Lexog, p, Z, eps, lanes=True)
        _, J = ks_ss_residual(xv, alpha, delta, dKD)
        step = np.linalg.solve(J, zres)
        eta = 1.0
        znorm = np.linalg.norm(zres)
        p_new = p - eta * step
        z_new = safe_eval(p_new)
        while (not np.isfinite(np.linalg.norm(z_new))) or np.linalg.norm(z_new) > znorm:
            eta /= 2
            if not eta > 1e-8:
                break
            p_new = p - eta * step
            z_new = safe_eval(p_new)
        p, zres = p_new, z_new
        it += 1
    xv, ss_value, _ = get_xvals(orc, Lexog, p, Z, eps)
    value, policy, _, _ = orc.egm_step(ss_value, xv["r"], xv["w"])
    Lss = Lexog @ endog_transition(orc, policy)
    D = invariant_dist(Lss)
    return SteadyState(xv, policy, Lss, D, ss_value), orc, dict(iterations=it, resnorm=float(np.linalg.norm(zres)))
