"""Shared deterministic inputs for the tests.  Test infrastructure: uses the oracle."""
import numpy as np

from oracle import oracle as O

KS = dict(beta=0.98, gamma=2.0, borrow_cons=0.0, alpha=0.36, delta=0.08, rho=0.966, sigma=0.283)

# parity tolerance of the north star: 1e-10 relative / 1e-12 absolute in FP64
RTOL, ATOL = 1e-10, 1e-12


def close(a, b, rtol=RTOL, atol=ATOL):
    """|a-b| <= atol*max(1, ||b||_inf) + rtol*|b| elementwise.

    The absolute term is scaled by the array's inf-norm because tangent arrays cross zero:
    there a 1-ulp perturbation of the inputs already moves the oracle's own output by
    ~2.5e-11 pointwise-relative (7e-12 absolute on a scale of 4e2; measured at 2000x11), so a
    pointwise relative test would compare conditioning, not implementations."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = max(1.0, float(np.max(np.abs(b)))) if b.size else 1.0
    return bool(np.all(np.abs(a - b) <= atol * scale + rtol * np.abs(b)))


def maxerr(a, b):
    """Largest |a-b| / (atol*scale + rtol*|b|): <= 1 passes."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = max(1.0, float(np.max(np.abs(b)))) if b.size else 1.0
    return float(np.max(np.abs(a - b) / (ATOL * scale + RTOL * np.abs(b)))) if b.size else 0.0


def model_inputs(n_a, n_e, gamma=2.0, borrow_cons=0.0, amax=200.0):
    grid = O.double_exponential(n_a, 0.0, amax)
    z, Pi, _ = O.rouwenhorst(n_e, KS["rho"], KS["sigma"])
    return dict(grid=grid, z=z, Pi=Pi, beta=KS["beta"], gamma=gamma, borrow_cons=borrow_cons)


def synthetic(n_a, n_e, T, K=0, gamma=2.0, seed=1234, rbar=0.015, wbar=1.35, c0=0.1):
    """Synthetic-throughput regime (ii): closed-form monotone terminal value, uniform D0,
    smooth r/w paths, N(0,1) tangent seeds."""
    m = model_inputs(n_a, n_e, gamma)
    P = T - 1
    g, z = m["grid"], m["z"]
    vT = (1 + rbar) * ((rbar * g[None, :] + wbar * z[:, None]) + c0) ** (-gamma)
    D0 = np.full((n_e, n_a), 1.0 / (n_a * n_e))
    t = np.arange(1, P + 1)
    r = rbar * (1 + 0.1 * 0.9 ** t); w = wbar * (1 + 0.05 * 0.9 ** t)
    rng = np.random.default_rng(seed)
    dr = rng.standard_normal((K, P)); dw = rng.standard_normal((K, P))
    return dict(m=m, T=T, P=P, vT=vT, D0=D0, r=r, w=w, dr=dr, dw=dw)


def make_oracle(m, T):
    return O.Oracle(m["grid"], m["z"], m["Pi"], m["beta"], m["gamma"], m["borrow_cons"], T)


def make_block(m, T, device=0):
    from hankb200.household import HouseholdBlock
    return HouseholdBlock(m["grid"], m["z"], m["Pi"], m["beta"], m["gamma"], m["borrow_cons"], T, device=device)


def ks_yaml_dict(n_a=200, n_e=7, T=150):
    """The reference's model-file contract (schema and values of KrusellSmith.yaml:12-116), built as a dict so the
    tests can write it to a temporary YAML file for build_model_from_yaml."""
    P = lambda name, value: {"name": name, "value": value}
    return {
        "file": {"name": "Krusell Smith Model", "function_file": "KrusellSmith.jl"},
        "parameters": {"model": [P("β", 0.98), P("borrow_cons", 0.0), P("γ", 2.0), P("α", 0.36), P("δ", 0.08)],
                       "computational": [P("T", T), P("ε", 1.0e-6), P("dx", 0.001)]},
        "dimensions": [
            {"name": "wealth", "type": "endogenous", "policy_var": "KD", "grid_function": "double_exponential",
             "params": {"n": n_a, "grid_min": 0.0, "grid_max": 200.0}},
            {"name": "productivity", "type": "exogenous", "grid_function": "rouwenhorst_discretization",
             "params": {"n": n_e, "ρ": 0.966, "σ": 0.283}}],
        "variables": {"endogenous": [{"name": k} for k in ("Y", "KS", "r", "w")],
                      "exogenous": [{"name": "Z", "seq_function": "exogenousZ"}],
                      "heterogeneous": [{"name": "KD"}, {"function": "ValueFunction"}]},
        "equations": ["Y = Z * KS(-1)^α", "r + δ = α * Z * KS(-1)^(α-1)", "w = (1-α) * Z * KS(-1)^α", "KS = KD"],
        "steady_states": {"initial": {"fixed": {"Z": 1.0}, "guesses": {"r": 0.04, "w": 1.0, "Y": 1.5, "KS": 3.5}},
                          "ending": {"fixed": {"Z": 2.0}, "guesses": {"r": 0.04, "w": 1.5, "Y": 2.0, "KS": 5.0}}},
    }
