"""ctypes binding of oracle/libhank_oracle.so — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see oracle/hank_oracle.cpp header).  PARITY UNPINNED: no golden vectors
exist in the reference and Julia is not installed; see DESIGN.md "Oracle".

Array conventions follow the reference (Julia column-major): an (n_a, n_e) Julia matrix is
passed as a numpy array of shape (n_e, n_a) C-order (a fastest); lane tangents are lane-major
with the lane as the leading axis.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_i32p = C.POINTER(C.c_int32)


def build(force=False):
    so = os.path.join(_HERE, "libhank_oracle.so")
    src = os.path.join(_HERE, "hank_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libhank_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.hanko_create.restype = C.c_void_p
        L.hanko_create.argtypes = [C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp, C.c_double, C.c_double, C.c_double]
        L.hanko_destroy.argtypes = [C.c_void_p]
        L.hanko_double_exponential_grid.argtypes = [C.c_double, C.c_double, C.c_int, c_dp]
        L.hanko_rouwenhorst.argtypes = [C.c_int, C.c_double, C.c_double, c_dp, c_dp, c_dp]
        L.hanko_egm_step.argtypes = [C.c_void_p, c_dp, c_dp, C.c_double, C.c_double, C.c_int, c_dp, c_dp,
                                     c_dp, c_dp, c_dp, c_dp, c_ip, c_i32p]
        L.hanko_lottery.argtypes = [C.c_void_p, c_dp, c_i32p, c_dp]
        L.hanko_forward_step.argtypes = [C.c_void_p, c_dp, c_dp, C.c_int, c_dp, c_dp, c_dp, c_dp]
        L.hanko_backward.argtypes = [C.c_void_p, c_dp, c_dp, c_dp, C.c_int, c_dp, c_dp, c_dp, c_dp, c_ip, c_dp, c_dp]
        L.hanko_forward.argtypes = [C.c_void_p, c_dp, c_dp, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp]
        L.hanko_ks_residuals.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, c_dp, c_dp, c_dp,
                                         C.c_int, c_dp, c_dp, c_dp, c_dp]
        L.hanko_ks_fjvp.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, c_dp, c_dp, c_dp, c_dp,
                                    C.c_int, c_dp, c_dp, c_dp, c_ip]
        L.hanko_gmres.argtypes = [c_dp, C.c_int, c_dp, c_dp, c_ip]
        L.hanko_newton.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, c_dp, c_dp, c_dp, c_dp, c_dp,
                                   C.c_double, C.c_double, C.c_int, C.c_int, c_dp, c_dp, c_ip, c_ip]
        L.hanko_lu_solve_dense.argtypes = [c_dp, C.c_int, C.c_int, c_dp]
        L.hanko_jl_pow.restype = C.c_double
        L.hanko_jl_pow.argtypes = [C.c_double, C.c_double]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _f(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


class OracleError(RuntimeError):
    def __init__(self, code, info):
        kinds = {1: "bad argument", 2: "DomainError: negative base in ^", 3: "knot-vectors must be unique and sorted"}
        super().__init__(f"oracle status {code} ({kinds.get(code, '?')}), info(kind,a,e,t)={list(info)}")
        self.code = code
        self.info = list(info)


def double_exponential(n, grid_min=0.0, grid_max=200.0):
    out = np.empty(n)
    lib().hanko_double_exponential_grid(float(grid_min), float(grid_max), int(n), _p(out))
    return out


def rouwenhorst(n, rho, sigma):
    """Returns (z, Pi, D): Pi[e, e2] row-stochastic (numpy indexing = Julia Π[e,e2])."""
    Pi = np.empty((n, n)); D = np.empty(n); z = np.empty(n)
    rc = lib().hanko_rouwenhorst(int(n), float(rho), float(sigma), _p(Pi), _p(D), _p(z))
    assert rc == 0
    return z, Pi.T.copy(), D  # column-major buffer -> Pi[e,e2]


class Oracle:
    """CPU oracle of the household block for one (grid, z, Π, β, γ, borrow_cons, T)."""

    def __init__(self, grid, z, Pi, beta, gamma, borrow_cons, T):
        self.grid = _f(grid); self.z = _f(z); self.Pi = _f(Pi)
        self.n_a = len(self.grid); self.n_e = len(self.z); self.T = int(T); self.P = self.T - 1
        self.G = self.n_a * self.n_e
        self.beta, self.gamma, self.borrow_cons = float(beta), float(gamma), float(borrow_cons)
        pi_cm = np.ascontiguousarray(self.Pi.T)  # column-major buffer of Π
        self.h = lib().hanko_create(self.n_a, self.n_e, self.T, _p(self.grid), _p(self.z), _p(pi_cm),
                                    self.beta, self.gamma, self.borrow_cons)
        assert self.h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().hanko_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # value_fn plug-in (KrusellSmith.jl:43-83) with K tangent lanes
    def egm_step(self, value_next, r, w, dvalue_next=None, dr=None, dw=None, want_interval=False):
        ne, na, G = self.n_e, self.n_a, self.G
        vn = _f(value_next, (ne, na))
        K = 0 if dr is None else len(dr)
        dvn = None if dvalue_next is None else _f(dvalue_next, (K, ne, na))
        drr = None if dr is None else _f(dr); dww = None if dw is None else _f(dw)
        value = np.empty((ne, na)); policy = np.empty((ne, na))
        dvalue = np.empty((K, ne, na)); dpolicy = np.empty((K, ne, na))
        info = np.zeros(4, dtype=np.int32)
        interval = np.zeros((ne, na), dtype=np.int32)
        rc = lib().hanko_egm_step(self.h, _p(vn), _p(dvn), float(r), float(w), K, _p(drr), _p(dww), _p(value),
                                  _p(policy), _p(dvalue), _p(dpolicy), info.ctypes.data_as(c_ip),
                                  interval.ctypes.data_as(c_i32p))
        if rc:
            raise OracleError(rc, info)
        out = (value, policy, dvalue, dpolicy)
        return out + (interval,) if want_interval else out

    def lottery(self, policy):
        pol = _f(policy, (self.n_e, self.n_a))
        m = np.empty((self.n_e, self.n_a), dtype=np.int32); om = np.empty((self.n_e, self.n_a))
        lib().hanko_lottery(self.h, _p(pol), m.ctypes.data_as(c_i32p), _p(om))
        return m, om

    def forward_step(self, policy, D, dpolicy=None, dD=None):
        pol = _f(policy, (self.n_e, self.n_a)); D = _f(D).reshape(self.n_e, self.n_a)
        K = 0 if dpolicy is None else len(dpolicy)
        dp = None if K == 0 else _f(dpolicy, (K, self.n_e, self.n_a))
        dd = None if K == 0 else _f(dD, (K, self.n_e, self.n_a))
        Dn = np.empty((self.n_e, self.n_a)); dDn = np.empty((K, self.n_e, self.n_a))
        lib().hanko_forward_step(self.h, _p(pol), _p(D), K, _p(dp), _p(dd), _p(Dn), _p(dDn))
        return Dn, dDn

    def backward(self, value_T, r, w, dr=None, dw=None):
        """BackwardIteration.jl:46-116. r,w: (P,); dr,dw: (K,P). Returns policy (P,n_e,n_a),
        dpolicy (K,P,n_e,n_a), value at t=1 and its tangents."""
        P, ne, na = self.P, self.n_e, self.n_a
        vT = _f(value_T, (ne, na)); r = _f(r, (P,)); w = _f(w, (P,))
        K = 0 if dr is None else len(dr)
        drr = np.zeros((max(K, 1), P)) if K == 0 else _f(dr, (K, P))
        dww = np.zeros((max(K, 1), P)) if K == 0 else _f(dw, (K, P))
        policy = np.empty((P, ne, na)); dpolicy = np.empty((K, P, ne, na))
        v1 = np.empty((ne, na)); dv1 = np.empty((K, ne, na))
        info = np.zeros(4, dtype=np.int32)
        rc = lib().hanko_backward(self.h, _p(vT), _p(r), _p(w), K, _p(drr), _p(dww), _p(policy), _p(dpolicy),
                                  info.ctypes.data_as(c_ip), _p(v1), _p(dv1))
        if rc:
            raise OracleError(rc, info)
        return policy, dpolicy, v1, dv1

    def forward(self, D0, policy, dpolicy=None, want_path=False):
        """ForwardIteration.jl:253-311. Returns KD (P,), dKD (K,P) [, D_path (P,n_e,n_a), dD_last]."""
        P, ne, na = self.P, self.n_e, self.n_a
        D0 = _f(D0).reshape(ne, na); policy = _f(policy, (P, ne, na))
        K = 0 if dpolicy is None else len(dpolicy)
        dp = None if K == 0 else _f(dpolicy, (K, P, ne, na))
        KD = np.empty(P); dKD = np.empty((K, P))
        Dp = np.empty((P, ne, na)) if want_path else None
        dDl = np.empty((K, ne, na))
        lib().hanko_forward(self.h, _p(D0), _p(policy), K, _p(dp), _p(KD), _p(dKD), _p(Dp), _p(dDl))
        return (KD, dKD, Dp, dDl) if want_path else (KD, dKD)

    def ks_fjvp(self, ks, value_T, D0, Z, x, V=None):
        """fullFunction(x) and J(x)·V (NewtonRaphson.jl:77-83, GeneralStructures.jl:542-550).
        ks = (alpha, delta, ss_start_KS); V: (K, n) lane-major."""
        n = 4 * self.P
        x = _f(x, (n,)); Z = _f(Z, (self.P,))
        K = 0 if V is None else len(V)
        Vv = None if K == 0 else _f(V, (K, n))
        F = np.empty(n); JV = np.empty((K, n))
        info = np.zeros(4, dtype=np.int32)
        vT = _f(value_T, (self.n_e, self.n_a)); D0 = _f(D0).reshape(self.n_e, self.n_a)
        rc = lib().hanko_ks_fjvp(self.h, float(ks[0]), float(ks[1]), float(ks[2]), _p(vT), _p(D0), _p(Z), _p(x),
                                 K, _p(Vv), _p(F), _p(JV), info.ctypes.data_as(c_ip))
        if rc:
            raise OracleError(rc, info)
        return F, JV

    def jacobian(self, ks, value_T, D0, Z, x, cols=None, chunk=32):
        """Brute-force JVP Jacobian columns (directJVPJacobian pattern, SteadyState.jl:296-320,
        generalised to any column set). Returns (n, len(cols)) column-major-like array J[:, j]."""
        n = 4 * self.P
        cols = np.arange(n) if cols is None else np.asarray(cols)
        J = np.empty((n, len(cols)))
        for s in range(0, len(cols), chunk):
            cc = cols[s:s + chunk]
            V = np.zeros((len(cc), n)); V[np.arange(len(cc)), cc] = 1.0
            _, JV = self.ks_fjvp(ks, value_T, D0, Z, x, V)
            J[:, s:s + len(cc)] = JV.T
        return J

    def newton(self, ks, value_T, D0, Z, Jbar, x0, eps=1e-9, eps_inner=1e-9, solver="gmres", max_inner=0):
        """NewtonRaphsonHANK (NewtonRaphson.jl:27-114). Jbar: (n,n) numpy with Jbar[i,j]."""
        n = 4 * self.P
        Jcm = np.ascontiguousarray(np.asarray(Jbar, dtype=np.float64).T)  # column-major buffer
        x0 = _f(x0, (n,)); Z = _f(Z, (self.P,))
        xo = np.empty(n); stats = np.zeros(8); inner = np.zeros(100, dtype=np.int32)
        info = np.zeros(4, dtype=np.int32)
        vT = _f(value_T, (self.n_e, self.n_a)); D0 = _f(D0).reshape(self.n_e, self.n_a)
        rc = lib().hanko_newton(self.h, float(ks[0]), float(ks[1]), float(ks[2]), _p(vT), _p(D0), _p(Z), _p(Jcm),
                                _p(x0), float(eps), float(eps_inner), {"gmres": 0, "lu": 1}[solver], int(max_inner),
                                _p(xo), _p(stats), inner.ctypes.data_as(c_ip), info.ctypes.data_as(c_ip))
        if rc:
            raise OracleError(rc, info)
        outer = int(stats[0])
        return xo, dict(outer=outer, jvps=int(stats[1]), fevals=int(stats[2]), ynorm=float(stats[3]),
                        gmres_iters=int(stats[4]), inner=[int(v) for v in inner[:outer]])


def ks_residuals(P, alpha, delta, ss_start_KS, x, KD, Z, dx=None, dKD=None):
    x = _f(x, (4 * P,)); KD = _f(KD, (P,)); Z = _f(Z, (P,))
    K = 0 if dx is None else len(dx)
    dxx = None if K == 0 else _f(dx, (K, 4 * P)); dkd = None if K == 0 else _f(dKD, (K, P))
    F = np.empty(4 * P); dF = np.empty((K, 4 * P))
    lib().hanko_ks_residuals(P, float(alpha), float(delta), float(ss_start_KS), _p(x), _p(KD), _p(Z), K,
                             _p(dxx), _p(dkd), _p(F), _p(dF))
    return F, dF


def gmres(A, x0, b):
    A = np.asarray(A, dtype=np.float64); n = len(b)
    Acm = np.ascontiguousarray(A.T); x = _f(x0).copy(); b = _f(b)
    it = C.c_int(0)
    lib().hanko_gmres(_p(Acm), n, _p(x), _p(b), C.byref(it))
    return x, it.value


def jl_pow(x, y):
    return lib().hanko_jl_pow(float(x), float(y))
