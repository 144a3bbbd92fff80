#!/usr/bin/env python
"""Measures the five BASELINE.json configurations on one GPU and prints one JSON line each.
Synthetic-throughput regime for C4/C5; model-consistent regime for C1-C3.
usage: python tools/configs_report.py [--only C1 C4 ...]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "julia-newtonraphsonhank_b200"))
from hankb200 import HouseholdBlock, model as M  # noqa: E402

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def synthetic_block(n_a, n_e, T, gamma=2.0, rbar=0.015, wbar=1.35, c0=0.1):
    g = M.double_exponential(n_a, 0.0, 200.0)
    z, Pi = M.rouwenhorst_discretization(n_e, 0.966, 0.283)
    blk = HouseholdBlock(g, z, Pi, 0.98, gamma, 0.0, T)
    P = T - 1
    vT = (1 + rbar) * ((rbar * g[None, :] + wbar * z[:, None]) + c0) ** (-gamma)
    blk.set_terminal(vT); blk.set_initial_dist(np.full((n_e, n_a), 1.0 / (n_a * n_e)))
    t = np.arange(1, P + 1)
    return blk, rbar * (1 + 0.1 * 0.9 ** t), wbar * (1 + 0.05 * 0.9 ** t)


def sweep_times(blk, r, w, K, reps=3):
    P = blk.P
    rng = np.random.default_rng(1234)
    dr = rng.standard_normal((K, P)) if K else None; dw = rng.standard_normal((K, P)) if K else None
    blk.block(r, w, dr, dw)
    blk.profile(True); blk.kernel_times(reset=True)
    for _ in range(reps):
        blk.block(r, w, dr, dw)
    kt = blk.kernel_times(reset=True); blk.profile(False)
    per = {k: v[0] / max(v[1], 1) for k, v in kt.items()}
    out = {"ms": {k: round(v, 4) for k, v in per.items()}, "us_per_period": {k: round(1e3 * v / P, 3) for k, v in per.items()}}
    if K:
        alg = 8.0 * blk.G * P * K
        out["GBps"] = {k: round(alg / (per[k] * 1e-3) / 1e9, 1) for k in ("backward_tangent", "forward_tangent")}
        out["frac_of_measured_hbm"] = {k: round(v / HBM, 4) for k, v in out["GBps"].items()}
        out["sweep_pair_GBps"] = round(2 * alg / ((per["backward_tangent"] + per["forward_tangent"]) * 1e-3) / 1e9, 1)
    return out


def model_case(fixture, label):
    from bench import load_fixture, WORKLOADS
    WORKLOADS.setdefault(label, dict(fixture=fixture, desc=label))
    fx = load_fixture(label); g = fx["g"]
    blk = HouseholdBlock(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), fx["T"])
    blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"]); blk.ks_configure(*fx["ks"])
    return blk, fx


ap = argparse.ArgumentParser(); ap.add_argument("--only", nargs="*", default=None)
ap.add_argument("--gmres", action="store_true", help="also time the reference-faithful GMRES Newton at C1 (tens of seconds)")
args = ap.parse_args()
want = lambda c: args.only is None or c in args.only

if want("C1"):   # YAML default grid: 200x7, T=150 — steady state, Jacobian and Newton through the host API mirror
    import yaml, tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from hankb200.steady_state import find_ss
    from hankb200.newton import TransitionProblem, directJVPJacobian, NewtonRaphsonHANK
    d = dict(file={"name": "KS", "function_file": "KrusellSmith.jl"},
             parameters={"model": [{"name": k, "value": v} for k, v in (("β", .98), ("borrow_cons", 0.0), ("γ", 2.0), ("α", .36), ("δ", .08))],
                         "computational": [{"name": "T", "value": 150}, {"name": "ε", "value": 1e-6}]},
             dimensions=[{"name": "wealth", "type": "endogenous", "policy_var": "KD", "grid_function": "double_exponential", "params": {"n": 200, "grid_min": 0.0, "grid_max": 200.0}},
                         {"name": "productivity", "type": "exogenous", "grid_function": "rouwenhorst_discretization", "params": {"n": 7, "ρ": .966, "σ": .283}}],
             variables={"endogenous": [{"name": k} for k in ("Y", "KS", "r", "w")], "exogenous": [{"name": "Z"}], "heterogeneous": [{"name": "KD"}, {"function": "ValueFunction"}]},
             equations=["Y = Z * KS(-1)^α", "r + δ = α * Z * KS(-1)^(α-1)", "w = (1-α) * Z * KS(-1)^α", "KS = KD"],
             steady_states={"initial": {"fixed": {"Z": 1.0}, "guesses": {"r": .04, "w": 1.0, "Y": 1.5, "KS": 3.5}}})
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False, encoding="utf-8") as f:
        yaml.safe_dump(d, f, allow_unicode=True)
    mod = M.build_model_from_yaml(f.name)
    t0 = time.perf_counter(); ss, info = find_ss(mod, mod.ss_initial); t_ss = time.perf_counter() - t0
    P = mod.compspec.T - 1
    prob = TransitionProblem(mod, ss, ss, {"Z": 1.0 + 0.8 ** np.arange(1, P + 1)})
    x0 = prob.x_steady()
    probJ = TransitionProblem(mod, ss, ss, {"Z": np.ones(P)}, blk=prob.blk)
    directJVPJacobian(probJ)
    t_J = 1e9
    for _ in range(3):   # min of 3: a single 3 ms host-timed sample is noisy
        t0 = time.perf_counter(); J = directJVPJacobian(probJ); t_J = min(t_J, time.perf_counter() - t0)
    prob = TransitionProblem(mod, ss, ss, {"Z": 1.0 + 0.8 ** np.arange(1, P + 1)}, blk=prob.blk)
    res = {}
    for sv in ("lu", "lu_batched"):
        NewtonRaphsonHANK(x0, J, prob, solver=sv, verbose=False)
        t0 = time.perf_counter(); x, st = NewtonRaphsonHANK(x0, J, prob, solver=sv, verbose=False)
        res[sv] = {"ms": round(1e3 * (time.perf_counter() - t0), 2), "outer": st["outer"], "jvps": st["jvps"], "inner": st["inner"]}
    if args.gmres:   # reference-faithful inner solver: restarted GMRES(20), maxiter = n per call
        t0 = time.perf_counter(); xg, stg = NewtonRaphsonHANK(x0, J, prob, solver="gmres", verbose=False)
        res["gmres"] = {"ms": round(1e3 * (time.perf_counter() - t0), 1), "outer": stg["outer"], "jvps": stg["jvps"],
                        "gmres_iters": stg["gmres_iters"], "max_abs_diff_vs_lu": float(np.max(np.abs(xg - x)))}
    print(json.dumps({"config": "C1 KS 200x7 T=150 (YAML default) via the host API", "steady_state": {"s": round(t_ss, 2), **info, "r": ss.vars["r"], "KS": ss.vars["KS"]},
                      "jacobian_build_ms": round(1e3 * t_J, 2), "newton": res, "residual_norm": float(np.linalg.norm(prob.fullFunction(x)))}))
    prob.blk.close()

if want("C2") or want("C3"):
    blk, fx = model_case("ss_500x7_T300.npz", "ks_500x7_T300")
    n, P = fx["n"], fx["P"]
    blk.reserve_lanes(n // 2)
    blk.linearize(fx["x0"], np.ones(P)); blk.jacobian_columns(1, 9)
    t0 = time.perf_counter(); F = blk.linearize(fx["x0"], fx["Z"]); t_F = time.perf_counter() - t0
    rng = np.random.default_rng(0); v = rng.standard_normal((1, n))
    blk.jvp(v); t0 = time.perf_counter(); blk.jvp(v); t_jvp = time.perf_counter() - t0
    blk.linearize(fx["x0"], np.ones(P))
    t_J = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); J = blk.jacobian_columns(1, n + 1); t_J = min(t_J, time.perf_counter() - t0)
    res = {}
    for sv in ("lu", "lu_batched"):
        blk.newton_solve(J, fx["x0"], fx["Z"], solver=sv)
        t0 = time.perf_counter(); x, st = blk.newton_solve(J, fx["x0"], fx["Z"], solver=sv)
        res[sv] = {"ms": round(1e3 * (time.perf_counter() - t0), 2), "outer": st["outer"], "jvps": st["jvps"]}
    print(json.dumps({"config": "C2/C3 KS 500x7 T=300", "F_ms_e2e": round(1e3 * t_F, 3), "jvp_k1_ms_e2e": round(1e3 * t_jvp, 3),
                      "jacobian_1196_columns_ms_e2e": round(1e3 * t_J, 2), "newton": res, "sweeps_k1": sweep_times(blk, fx["x0"].reshape(P, 4)[:, 2], fx["x0"].reshape(P, 4)[:, 3], 1),
                      "sweeps_k592": sweep_times(blk, fx["x0"].reshape(P, 4)[:, 2], fx["x0"].reshape(P, 4)[:, 3], 592)}))
    blk.close()

if want("C4"):
    blk, r, w = synthetic_block(2000, 11, 500)
    out = {"config": "C4 large grid 2000x11 T=500 (synthetic regime ii)", "primal_K0": sweep_times(blk, r, w, 0), "K64": sweep_times(blk, r, w, 64)}
    print(json.dumps(out)); blk.close()

if want("C5"):
    blk, r, w = synthetic_block(1000, 7, 300)
    out = {"config": "C5 1000x7 T=300, 64 lanes per GPU per pass (synthetic regime ii)", "K64": sweep_times(blk, r, w, 64), "K444": sweep_times(blk, r, w, 444)}
    print(json.dumps(out)); blk.close()
