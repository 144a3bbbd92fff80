"""Steady-state solve on the host with the device value function (SteadyState.jl:111-259).

As in the reference the solver is host code: Newton over the free endogenous variables with
ForwardDiff-style tangents (4 lanes, SteadyState.jl:195), a halving line search that treats device-side
failures as infinite residuals (:199-207), and a sparse direct solve for the invariant distribution and
its tangents (ForwardIteration.jl:436-442, :480-558; SciPy SuperLU where Julia uses UMFPACK).  The inner
VFI — ~470 EGM steps per evaluation — runs on the GPU (`hank_vfi`).
"""
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from ._lib import HankError


@dataclass
class SteadyState:                   # SteadyState.jl:21-27
    vars: dict
    policies: dict
    Lam: sp.csc_matrix
    D: np.ndarray
    value: np.ndarray


def _endog_transition(blk, policy, dpolicy=None):
    """make_endogenous_transition (ForwardIteration.jl:37-78) from the device bracket kernel."""
    na, G = blk.n_a, blk.G
    m, om = blk.lottery(policy)
    m = m.reshape(-1).astype(np.int64); om = om.reshape(-1)
    cols = np.arange(G); e_of = cols // na
    edge = (m == 1) | (m > na)
    lo = e_of * na + np.clip(m - 2, 0, na - 1); hi = e_of * na + np.clip(m - 1, 0, na - 1)
    I = np.concatenate([hi[edge], lo[~edge], hi[~edge]]); J = np.concatenate([cols[edge], cols[~edge], cols[~edge]])
    L = sp.csc_matrix((np.concatenate([np.ones(edge.sum()), 1.0 - om[~edge], om[~edge]]), (I, J)), shape=(G, G))
    if dpolicy is None:
        return L
    g = blk.grid
    dg = np.where(edge, 1.0, g[np.clip(m - 1, 0, na - 1)] - g[np.clip(m - 2, 0, na - 1)])
    dLs = []
    for dp in dpolicy:
        dom = np.where(edge, 0.0, dp.reshape(-1) / dg)
        dLs.append(sp.csc_matrix((np.concatenate([np.zeros(edge.sum()), -dom[~edge], dom[~edge]]), (I, J)), shape=(G, G)))
    return L, dLs


def invariant_dist(Lam, dLams=None):
    """invariant_dist(Λ') and its Sherman-Morrison tangents."""
    Lam = sp.csc_matrix(Lam); n = Lam.shape[0]
    fac = spla.splu((sp.identity(n - 1, format="csc") - Lam[1:, 1:]).tocsc())
    y2 = fac.solve(np.asarray(Lam[1:, 0].todense()).reshape(-1))
    s = 1.0 + y2.sum()
    D0 = np.concatenate([[1.0], y2]) / s
    if dLams is None:
        D = np.concatenate([[1.0], y2])
        return D / D.sum()
    out = []
    for dL in dLams:
        y1 = fac.solve((dL @ D0)[1:])
        tail = y1 - y2 * (y1.sum() / s)
        out.append(np.concatenate([[-tail.sum()], tail]))
    return D0, out


def _residual(xv, alpha, delta, dKD=None):
    Y, KS, r, w, KD, Z = (xv[k] for k in ("Y", "KS", "r", "w", "KD", "Z"))
    Ka, Ka1 = KS ** alpha, KS ** (alpha - 1.0)
    z = np.array([Y - Z * Ka, (r + delta) - (alpha * Z) * Ka1, w - ((1 - alpha) * Z) * Ka, KS - KD])
    if dKD is None:
        return z
    J = np.zeros((4, 4))
    J[0, 0] = 1.0; J[0, 1] = -Z * alpha * Ka1
    J[1, 2] = 1.0; J[1, 1] = -(alpha * Z) * (alpha - 1.0) * KS ** (alpha - 2.0)
    J[2, 3] = 1.0; J[2, 1] = -((1 - alpha) * Z) * alpha * Ka1
    J[3, 1] = 1.0; J[3, :] -= dKD
    return z, J


def find_ss(model, spec, label="initial", blk=None, verbose=False):
    """find_ss (SteadyState.jl:184-233).  Returns (SteadyState, info)."""
    own = blk is None
    blk = blk or model.household_block(T=2)
    pr = model.heterogeneity["productivity"]
    Lexog = sp.kron(sp.csc_matrix(pr.transition.T), sp.identity(blk.n_a, format="csc"), format="csc")
    eps, alpha, delta = model.compspec.eps, model.params["α"], model.params["δ"]
    Z = float(spec["fixed"]["Z"])
    p = np.array([spec["guesses"].get(k, 1.0) for k in ("Y", "KS", "r", "w")], dtype=np.float64)

    def xvals(q, lanes):
        Y, KS, r, w = q
        if lanes:
            val, pol, dval, dpol, _ = blk.vfi(r, w, [0, 0, 1.0, 0], [0, 0, 0, 1.0], eps)
            Lend, dLend = _endog_transition(blk, pol, dpol)
            D, dD = invariant_dist(Lexog @ Lend, [Lexog @ d for d in dLend])
            dKD = np.array([D @ dpol[k].reshape(-1) + pol.reshape(-1) @ dD[k] for k in range(4)])
        else:
            val, pol, _, _, _ = blk.vfi(r, w, eps=eps)
            D = invariant_dist(Lexog @ _endog_transition(blk, pol)); dKD = None
        return dict(Y=Y, KS=KS, r=r, w=w, KD=float(pol.reshape(-1) @ D), Z=Z), val, dKD

    def safe_eval(q):
        try:
            return _residual(xvals(q, False)[0], alpha, delta)
        except HankError:
            return np.full(4, np.inf)

    z = _residual(xvals(p, False)[0], alpha, delta)
    it = 0
    while np.linalg.norm(z) > eps and it < 100:
        if verbose:
            print(f"  [{label}] Iteration {it}: residual norm = {np.linalg.norm(z)}")
        xv, _, dKD = xvals(p, True)
        _, J = _residual(xv, alpha, delta, dKD)
        step = np.linalg.solve(J, z)
        eta, znorm = 1.0, np.linalg.norm(z)
        p_new = p - eta * step; z_new = safe_eval(p_new)
        while (not np.isfinite(np.linalg.norm(z_new))) or np.linalg.norm(z_new) > znorm:
            eta /= 2
            if not eta > 1e-8:
                break
            p_new = p - eta * step; z_new = safe_eval(p_new)
        p, z = p_new, z_new
        it += 1
    xv, ss_value, _ = xvals(p, False)
    _, policy, _, _ = blk.egm_step(ss_value, xv["r"], xv["w"])
    Lss = Lexog @ _endog_transition(blk, policy)
    D = invariant_dist(Lss)
    if own:
        blk.close()
    return SteadyState(xv, {"KD": policy}, Lss, D, ss_value), dict(iterations=it, resnorm=float(np.linalg.norm(z)))


def get_SteadyStates(model, blk=None, verbose=False):
    """get_SteadyStates (SteadyState.jl:245-259)."""
    ss_i, _ = find_ss(model, model.ss_initial, "initial", blk, verbose)
    if model.ss_initial is model.ss_ending:
        return ss_i, ss_i
    ss_e, _ = find_ss(model, model.ss_ending, "ending", blk, verbose)
    return ss_i, ss_e
