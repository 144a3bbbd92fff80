"""Model definition: the reference's YAML contract and grid builders, host side.

Mirrors `build_model_from_yaml` (ModelParser.jl:296-379) for models whose household block is the
Krusell-Smith `ValueFunction` (KrusellSmith.jl:43-83) — the one device plug-in this library ships — and the
out-of-the-box grid functions of GeneralStructures.jl:242-261, :474-525.  Equation strings are parsed only
far enough to check that they are the KS set compiled into the device residual kernel
(KrusellSmith.yaml:90-94); arbitrary equation compilation (ModelParser.jl:54-259) is out of scope.
"""
from dataclasses import dataclass, field

import numpy as np
import yaml


def double_exponential(n, grid_min, grid_max):
    """make_DoubleExponentialGrid (GeneralStructures.jl:474-483)."""
    U = np.log(1.0 + np.log(1.0 + grid_max - grid_min))
    u = (np.arange(n, dtype=np.longdouble) * np.longdouble(U) / np.longdouble(n - 1)).astype(np.float64)
    return grid_min + np.exp(np.exp(u) - 1.0) - 1.0


def invariant_dist_dense(Pi):
    """invariant_dist (ForwardIteration.jl:436-442) for a small dense row-stochastic chain."""
    PT = np.asarray(Pi, dtype=np.float64).T
    n = PT.shape[0]
    M = np.eye(n - 1) - PT[1:, 1:]
    D = np.concatenate([[1.0], np.linalg.solve(M, PT[1:, 0])])
    return D / D.sum()


def rouwenhorst_discretization(n, rho, sigma):
    """get_RouwenhorstDiscretization (GeneralStructures.jl:500-525): returns (z, Pi) with E[z] = 1."""
    p = (1.0 + rho) / 2.0
    Pi = np.array([[p, 1 - p], [1 - p, p]])
    for i in range(3, n + 1):
        old = Pi
        Pi = np.zeros((i, i))
        Pi[: i - 1, : i - 1] += p * old
        Pi[: i - 1, 1:] += (1 - p) * old
        Pi[1:, : i - 1] += (1 - p) * old
        Pi[1:, 1:] += p * old
        Pi[1 : i - 1, :] /= 2
    D = invariant_dist_dense(Pi)
    alpha = 2.0 * (sigma / np.sqrt(n - 1))
    z = np.exp(alpha * np.arange(n))
    z = z / np.sum(z * D)
    return z, Pi


def kronecker_exogenous(processes):
    """Several independent exogenous processes [(z_1, Pi_1), (z_2, Pi_2), ...] as ONE income dimension: the state index
    runs over the first process fastest, e = e_1 + n_1 e_2 + ..., exactly the ordering of the reference's
    Λ_exog = kron(Π_K', kron(…, kron(Π_1', I_{n_a}))) (ForwardIteration.jl:280-284), so Pi = kron(Pi_K, …, Pi_1) and the
    income level of a combined state is the product of its components' levels.  The device block takes the result like
    any other (z, Pi); state counts without a kernel instantiation are padded inside hank_ctx_create."""
    z = np.ones(1); Pi = np.ones((1, 1))
    for zi, Pii in processes:
        z = np.kron(np.asarray(zi, dtype=np.float64), z)
        Pi = np.kron(np.asarray(Pii, dtype=np.float64), Pi)
    return z, Pi


GRID_FUNCTIONS = {"double_exponential": double_exponential, "rouwenhorst_discretization": rouwenhorst_discretization}
KS_EQUATIONS = ("Y=Z*KS(-1)^α", "r+δ=α*Z*KS(-1)^(α-1)", "w=(1-α)*Z*KS(-1)^α", "KS=KD")


@dataclass
class HeterogeneityDimension:        # GeneralStructures.jl:43-49
    name: str
    dim_type: str
    n: int
    grid: np.ndarray
    transition: np.ndarray = None
    policy_var: str = None


@dataclass
class ComputationalSpec:             # GeneralStructures.jl:166-174
    T: int
    eps: float
    dx: float
    n_v: int
    n_endog: int
    max_lag: int
    max_lead: int


@dataclass
class SequenceModel:                 # GeneralStructures.jl:216-226
    params: dict
    compspec: ComputationalSpec
    heterogeneity: dict
    endogenous: tuple
    heterogeneous: tuple
    exogenous: tuple
    equations: tuple
    ss_initial: dict
    ss_ending: dict
    value_fn: str = "ValueFunction"
    var_names: tuple = field(default=())

    def household_block(self, device=0, T=None):
        from .household import HouseholdBlock
        w, pr, p = self.heterogeneity["wealth"], self.heterogeneity["productivity"], self.params
        return HouseholdBlock(w.grid, pr.grid, pr.transition, p["β"], p["γ"], p["borrow_cons"],
                              self.compspec.T if T is None else T, device=device)


def build_model_from_yaml(path, overrides=None):
    """ModelParser.jl:296-379.  `overrides` may replace dimension sizes / T, e.g.
    {"T": 300, "wealth.n": 500} (the reference edits the YAML for that)."""
    y = yaml.safe_load(open(path, encoding="utf-8"))
    ov = overrides or {}
    params = {p["name"]: float(p["value"]) for p in y["parameters"]["model"]}
    cs = {p["name"]: p["value"] for p in y["parameters"].get("computational", [])}
    T = int(ov.get("T", cs.get("T", 150))); eps = float(cs.get("ε", 1e-6)); dx = float(cs.get("dx", 1e-8))
    dims = {}
    for d in y["dimensions"]:
        fn = GRID_FUNCTIONS.get(d["grid_function"])
        if fn is None:
            raise ValueError(f"grid function {d['grid_function']!r} is not one of {sorted(GRID_FUNCTIONS)}")
        kw = dict(d["params"])
        if f"{d['name']}.n" in ov:
            kw["n"] = int(ov[f"{d['name']}.n"])
        if d["type"] == "endogenous":
            grid = fn(int(kw["n"]), float(kw["grid_min"]), float(kw["grid_max"]))
            dims[d["name"]] = HeterogeneityDimension(d["name"], "endogenous", len(grid), grid, None, d.get("policy_var"))
        else:
            z, Pi = fn(int(kw["n"]), float(kw["ρ"]), float(kw["σ"]))
            dims[d["name"]] = HeterogeneityDimension(d["name"], "exogenous", len(z), z, Pi)
    v = y["variables"]
    endog = tuple(x["name"] for x in v.get("endogenous", []))
    het = tuple(x["name"] for x in v.get("heterogeneous", []) if "name" in x)
    fns = [x["function"] for x in v.get("heterogeneous", []) if "function" in x]
    if len(fns) != 1:
        raise ValueError("the 'heterogeneous' section must contain exactly one 'function' entry")
    if fns[0] != "ValueFunction":
        raise NotImplementedError(f"value function {fns[0]!r}: only the Krusell-Smith ValueFunction has a device kernel")
    exog = tuple(x["name"] for x in v.get("exogenous", []))
    eqs = tuple(str(e) for e in y["equations"])
    if tuple(e.replace(" ", "") for e in eqs) != KS_EQUATIONS:
        raise NotImplementedError("only the Krusell-Smith equations (KrusellSmith.yaml:90-94) are compiled into the device residuals")
    ss = y["steady_states"]
    parse = lambda s: dict(fixed=dict(s.get("fixed", {})), guesses=dict(s.get("guesses", {})))
    ss_i = parse(ss["initial"]); ss_e = parse(ss["ending"]) if "ending" in ss else ss_i
    names = endog + het + exog   # ModelParser.jl:357
    return SequenceModel(params, ComputationalSpec(T, eps, dx, len(names), len(endog), 1, 0), dims, endog, het, exog,
                         eqs, ss_i, ss_e, "ValueFunction", names)
