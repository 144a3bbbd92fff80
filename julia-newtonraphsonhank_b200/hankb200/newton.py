"""The reference's solver-level call signatures over the device block (NewtonRaphson.jl, SteadyState.jl:296-320,
GeneralStructures.jl:542-550), so scripts written against the reference read the same."""
import numpy as np


class TransitionProblem:
    """Binds (model, ss_initial, ss_ending, exog_paths) to one device block: the closure `fullFunction`
    of y_Iteration (NewtonRaphson.jl:77-83)."""

    def __init__(self, model, ss_initial, ss_ending, exog_paths, device=0, blk=None):
        self.model, self.ss_initial, self.ss_ending = model, ss_initial, ss_ending
        self.P = model.compspec.T - 1
        self.n = model.compspec.n_endog * self.P
        self.Z = np.ascontiguousarray(exog_paths["Z"], dtype=np.float64)
        self.blk = blk or model.household_block(device)
        self.blk.set_terminal(ss_ending.value)
        self.blk.set_initial_dist(ss_initial.D)
        self.blk.ks_configure(model.params["α"], model.params["δ"], ss_initial.vars["KS"])
        self._x = None

    def fullFunction(self, x):
        self._x = np.array(x, dtype=np.float64)
        return self.blk.linearize(self._x, self.Z)

    def x_steady(self):
        return np.tile([self.ss_initial.vars[k] for k in self.model.endogenous], self.P)


def JVP(problem, primal, tangent):
    """JVP(func, primal, tangent) (GeneralStructures.jl:542-550) for func = problem.fullFunction.  A (K, n)
    array of tangents rides as K lanes in one pass.  The linearisation is reused while `primal` is unchanged."""
    primal = np.asarray(primal, dtype=np.float64)
    if problem._x is None or not np.array_equal(problem._x, primal):
        problem.fullFunction(primal)
    V = np.asarray(tangent, dtype=np.float64)
    out = problem.blk.jvp(V)
    return out[0] if V.ndim == 1 else out


def directJVPJacobian(problem, x=None, cols=None):
    """directJVPJacobian (SteadyState.jl:296-320) generalised to any 1-based half-open column range
    (default: all n columns = the full sequence-space Jacobian, BASELINE config 3)."""
    x = problem.x_steady() if x is None else x
    problem.fullFunction(x)
    b, e = (1, problem.n + 1) if cols is None else cols
    return problem.blk.jacobian_columns(b, e)


def NewtonRaphsonHANK(x_0, Jbar, problem, eps=1e-9, solver="lu", verbose=True):
    """NewtonRaphsonHANK(x_0, J̅, exog_paths, mod, ss_initial, ss_ending; ε) (NewtonRaphson.jl:27-46); the model,
    steady states and exogenous paths are bound in `problem`.  solver: "gmres" (reference), "lu", "lu_batched"."""
    x, st = problem.blk.newton_solve(Jbar, x_0, problem.Z, eps=eps, solver=solver)
    if verbose:
        print(f"Iteration: {st['outer'] + 1}, norm(y): {st['ynorm']}  (inner iterations {st['inner']})")
    return x, st
