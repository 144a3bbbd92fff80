"""Generates tests/golden/*.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference ships no golden vectors and Julia cannot
run here, so these files pin the ORACLE's behaviour (regression) and give the GPU tests fixed
inputs/outputs that travel to the GPU box; they are not reference outputs (parity unpinned)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import synthetic, make_oracle  # noqa: E402
from oracle.steady_state import find_ss  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sweep_case(name, n_a, n_e, T, K, gamma):
    s = synthetic(n_a, n_e, T, K, gamma)
    orc = make_oracle(s["m"], T)
    pol, dpol, v1, dv1 = orc.backward(s["vT"], s["r"], s["w"], s["dr"], s["dw"])
    KD, dKD, Dp, dDl = orc.forward(s["D0"], pol, dpol, want_path=True)
    m, om = orc.lottery(pol[0])
    np.savez_compressed(os.path.join(HERE, name), n_a=n_a, n_e=n_e, T=T, K=K, gamma=gamma, grid=s["m"]["grid"],
                        z=s["m"]["z"], Pi=s["m"]["Pi"], beta=s["m"]["beta"], borrow_cons=s["m"]["borrow_cons"],
                        vT=s["vT"], D0=s["D0"], r=s["r"], w=s["w"], dr=s["dr"], dw=s["dw"], pol=pol, dpol=dpol,
                        v1=v1, dv1=dv1, KD=KD, dKD=dKD, D_last=Dp[-1], dD_last=dDl, m1=m, om1=om)


def ks_case(name, n_a, n_e, T, ncols=6):
    """Model-consistent regime (i): oracle steady state, RunMain.jl:50-51 shock, F, JVPs, a few
    Jacobian columns and the Newton path (LU preconditioner solve)."""
    ss, orc, info = find_ss(n_a, n_e, T)
    P = T - 1; n = 4 * P
    ks = (0.36, 0.08, ss.vars["KS"])
    x0 = np.tile([ss.vars[k] for k in ("Y", "KS", "r", "w")], P)
    Z = 1.0 + 0.8 ** np.arange(1, P + 1)
    rng = np.random.default_rng(42)
    V = rng.standard_normal((3, n))
    F0, JV = orc.ks_fjvp(ks, ss.value, ss.D, Z, x0, V)
    cols = np.array([0, 1, 2, 3, n // 2 + 2, n - 1])[:ncols]
    Jc = orc.jacobian(ks, ss.value, ss.D, np.ones(P), x0, cols)
    J = orc.jacobian(ks, ss.value, ss.D, np.ones(P), x0)
    xs, st = orc.newton(ks, ss.value, ss.D, Z, J, x0, solver="lu")
    np.savez_compressed(os.path.join(HERE, name), n_a=n_a, n_e=n_e, T=T, grid=orc.grid, z=orc.z, Pi=orc.Pi,
                        beta=orc.beta, gamma=orc.gamma, borrow_cons=orc.borrow_cons, alpha=0.36, delta=0.08,
                        ss_vars=np.array([ss.vars[k] for k in ("Y", "KS", "r", "w", "KD", "Z")]),
                        ss_value=ss.value, ss_D=ss.D.reshape(n_e, n_a), ss_policy=ss.policy, x0=x0, Z=Z, V=V, F0=F0,
                        JV=JV, cols=cols, Jcols=Jc, Jbar=J, x_newton=xs, newton_inner=np.array(st["inner"]),
                        newton_jvps=st["jvps"], ss_iterations=info["iterations"])


def ss_case(name, n_a, n_e, T):
    """Steady-state record only (the input of bench.py's model-consistent workload)."""
    ss, orc, info = find_ss(n_a, n_e, T)
    np.savez_compressed(os.path.join(HERE, name), n_a=n_a, n_e=n_e, T=T, grid=orc.grid, z=orc.z, Pi=orc.Pi,
                        beta=orc.beta, gamma=orc.gamma, borrow_cons=orc.borrow_cons, alpha=0.36, delta=0.08,
                        ss_vars=np.array([ss.vars[k] for k in ("Y", "KS", "r", "w", "KD", "Z")]),
                        ss_value=ss.value, ss_D=ss.D.reshape(n_e, n_a), ss_policy=ss.policy,
                        ss_iterations=info["iterations"])


if __name__ == "__main__":
    ss_case("ss_500x7_T300.npz", 500, 7, 300)
    ss_case("ss_1000x7_T300.npz", 1000, 7, 300)
    sweep_case("sweep_60x3_T12_K2.npz", 60, 3, 12, 2, 2.0)
    sweep_case("sweep_200x7_T20_K3_g15.npz", 200, 7, 20, 3, 1.5)
    ks_case("ks_100x3_T30.npz", 100, 3, 30)
    ks_case("ks_200x7_T40.npz", 200, 7, 40)
    print("golden files written to", HERE)
