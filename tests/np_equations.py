"""Test infrastructure: two independent evaluations of a model's equilibrium equations on the padded variable matrix of
assemble_full_xMat (GeneralStructures.jl:329-377) — direct NumPy evaluation of the equation strings (what the
reference's compiled Julia function computes, ModelParser.jl:217-259), and a NumPy interpreter of the device bytecode."""
import re

import numpy as np


def padded_xmat(x, KD, Z, n_endog, ss_start, ss_end, max_lag, max_lead):
    """x: (P*n_endog,) variable-fastest; KD: (P,); Z: (n_exog, P) -> xMat (nv, max_lag + P + max_lead)."""
    P = len(KD)
    mid = np.vstack([np.asarray(x).reshape(P, n_endog).T, np.asarray(KD)[None, :], np.atleast_2d(Z)])
    return np.hstack([np.repeat(np.asarray(ss_start)[:, None], max_lag, 1), mid, np.repeat(np.asarray(ss_end)[:, None], max_lead, 1)])


def residuals_direct(equations, names, params, xmat, max_lag, P):
    """LHS - RHS of every equation over the P transition columns, equation-fastest like the reference's output."""
    out = np.empty((P, len(equations)))
    pat = "|".join(sorted((re.escape(n) for n in names), key=len, reverse=True))
    for i, eq in enumerate(equations):
        lhs, rhs = eq.split("=")
        vals = []
        for side in (lhs, rhs):
            s = side.strip().replace("^", "**")
            s = re.sub(rf"\b({pat})\s*\(\s*([+-]?\d+)\s*\)", lambda m: f"__S('{m.group(1)}', {int(m.group(2))})", s)
            env = {n: xmat[j, max_lag:max_lag + P] for j, n in enumerate(names)}
            env.update(params)
            env.update(exp=np.exp, log=np.log, sqrt=np.sqrt,
                       __S=lambda n, k: xmat[names.index(n), max_lag + k:max_lag + k + P])
            vals.append(eval(s, {"__builtins__": {}}, env) + np.zeros(P))
        out[:, i] = vals[0] - vals[1]
    return out.reshape(-1)


def residuals_bytecode(prog, xmat, max_lag, P):
    """Interprets hankb200.equations bytecode with NumPy vectors over the periods as stack entries."""
    from hankb200 import equations as E
    eq_off, code, consts = prog.arrays()
    out = np.empty((P, prog.n_eq))
    for i in range(prog.n_eq):
        st = []
        pc = eq_off[i]
        while pc < eq_off[i + 1]:
            op = code[pc]; pc += 1
            if op == E.OP_CONST:
                st.append(np.full(P, consts[code[pc]])); pc += 1
            elif op == E.OP_VAR:
                v, k = code[pc], code[pc + 1]; pc += 2
                st.append(xmat[v, max_lag + k:max_lag + k + P].copy())
            elif op in (E.OP_NEG, E.OP_EXP, E.OP_LOG, E.OP_SQRT):
                a = st.pop()
                st.append({E.OP_NEG: np.negative, E.OP_EXP: np.exp, E.OP_LOG: np.log, E.OP_SQRT: np.sqrt}[op](a))
            else:
                b = st.pop(); a = st.pop()
                st.append({E.OP_ADD: np.add, E.OP_SUB: np.subtract, E.OP_MUL: np.multiply, E.OP_DIV: np.divide,
                           E.OP_POW: np.power}[op](a, b))
        assert len(st) == 1
        out[:, i] = st[0]
    return out.reshape(-1)
