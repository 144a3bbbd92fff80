"""Independent vectorised numpy restatement of the household block, used only to pin the C++
oracle (tests/test_oracle.py).  It is written differently on purpose: np.interp for the EGM
interpolation values, np.searchsorted for brackets, a dense scatter for the lottery, and closed-form
"coefficient" tangents (q̇ = −s·[(1−δ)k̇_i + δ k̇_{i+1}], s = Δg/Δk) instead of dual-number rules.
Reference: KrusellSmith.jl:43-83, ForwardIteration.jl:37-99, :253-311."""
import numpy as np


def egm_step(m, vnext, r, w, dvnext=None, dr=None, dw=None):
    g, z, Pi = m["grid"], m["z"], m["Pi"]
    beta, gamma, bc = m["beta"], m["gamma"], m["borrow_cons"]
    ne, na = vnext.shape
    K = 0 if dr is None else len(dr)
    EV = Pi @ vnext                                   # EV[e,a] = Σ_e2 Π[e,e2] V[e2,a]
    B = beta * EV
    c = B ** (-1.0 / gamma)
    rho = 1.0 / (1.0 + r)
    S = (c - w * z[:, None]) + g[None, :]
    k = rho * S
    pol = np.empty_like(k); val = np.empty_like(k)
    idx = np.empty(k.shape, dtype=int); delta = np.empty_like(k); live = np.empty(k.shape, dtype=bool)
    for e in range(ne):
        q = np.interp(g, k[e], g)                     # Flat extrapolation is np.interp's default
        i = np.clip(np.searchsorted(k[e], np.clip(g, k[e, 0], k[e, -1]), side="left"), 1, na - 1) - 1
        idx[e] = i
        xh = np.clip(g, k[e, 0], k[e, -1])
        delta[e] = (xh - k[e, i]) / (k[e, i + 1] - k[e, i])
        live[e] = (g >= k[e, 0]) & (g <= k[e, -1]) & (q >= bc)
        pol[e] = np.maximum(q, bc)
    cg = (1 + r) * g[None, :] + w * z[:, None] - pol
    val = (1 + r) * cg ** (-gamma)
    if K == 0:
        return val, pol, None, None
    dval = np.empty((K, ne, na)); dpol = np.empty((K, ne, na))
    for l in range(K):
        dEV = Pi @ dvnext[l] if dvnext is not None else np.zeros_like(EV)
        dc = (-1.0 / gamma) * c / B * (beta * dEV)
        dk = rho * (dc - dw[l] * z[:, None]) - rho * rho * dr[l] * S
        for e in range(ne):
            i = idx[e]
            s = (g[i + 1] - g[i]) / (k[e, i + 1] - k[e, i])
            dq = -s * ((1 - delta[e]) * dk[e, i] + delta[e] * dk[e, i + 1])
            dpol[l, e] = np.where(live[e], dq, 0.0)
        dcg = dr[l] * g[None, :] + dw[l] * z[:, None] - dpol[l]
        dval[l] = dr[l] * cg ** (-gamma) + (1 + r) * (-gamma) * cg ** (-gamma - 1) * dcg
    return val, pol, dval, dpol


def backward(m, vT, r, w, dr=None, dw=None):
    P = len(r)
    K = 0 if dr is None else len(dr)
    pol = np.empty((P,) + vT.shape); dpol = np.empty((K, P) + vT.shape)
    v = vT; dv = None
    for t in range(P - 1, -1, -1):
        v, pol[t], dv, dp = egm_step(m, v, r[t], w[t], dv, None if K == 0 else dr[:, t], None if K == 0 else dw[:, t])
        if K:
            dpol[:, t] = dp
    return pol, dpol, v, dv


def lottery(g, pol):
    na = len(g)
    m = np.searchsorted(g, pol, side="left") + 1     # Julia searchsortedfirst, 1-based
    inner = (m > 1) & (m <= na)
    lo = np.clip(m - 2, 0, na - 1); hi = np.clip(m - 1, 0, na - 1)
    dg = np.where(inner, g[hi] - g[lo], 1.0)
    om = np.where(inner, (pol - g[lo]) / dg, 1.0)
    return m, om, inner, lo, hi, dg


def forward_step(m_, pol, D, dpol=None, dD=None):
    g, Pi = m_["grid"], m_["Pi"]
    ne, na = pol.shape
    m, om, inner, lo, hi, dg = lottery(g, pol)
    row_hi = np.where(m > na, na - 1, hi)            # m == 1 -> node 0, m > n_a -> node n_a-1
    tmp = np.zeros((ne, na))
    for e in range(ne):
        np.add.at(tmp[e], row_hi[e], np.where(inner[e], om[e], 1.0) * D[e])
        np.add.at(tmp[e], lo[e], np.where(inner[e], (1 - om[e]) * D[e], 0.0))
    Dn = Pi.T @ tmp                                   # D+[e2,a] = Σ_e Π[e,e2] tmp[e,a]
    if dpol is None:
        return Dn, None
    K = len(dpol)
    dDn = np.empty((K, ne, na))
    for l in range(K):
        dom = np.where(inner, dpol[l] / dg, 0.0)
        t2 = np.zeros((ne, na))
        for e in range(ne):
            np.add.at(t2[e], row_hi[e], np.where(inner[e], om[e], 1.0) * dD[l, e] + dom[e] * D[e])
            np.add.at(t2[e], lo[e], np.where(inner[e], (1 - om[e]) * dD[l, e], 0.0) - dom[e] * D[e])
        dDn[l] = Pi.T @ t2
    return Dn, dDn


def forward(m_, D0, pol, dpol=None):
    P = len(pol)
    K = 0 if dpol is None else len(dpol)
    D = D0; dD = np.zeros((K,) + D0.shape) if K else None
    KD = np.empty(P); dKD = np.empty((K, P))
    for t in range(P):
        D, dD = forward_step(m_, pol[t], D, None if K == 0 else dpol[:, t], dD)
        KD[t] = np.sum(pol[t] * D)
        for l in range(K):
            dKD[l, t] = np.sum(dpol[l, t] * D) + np.sum(pol[t] * dD[l])
    return KD, dKD
