"""Host-side handle of the device household block (one hank_ctx).

Array conventions are the reference's (Julia column-major): an (n_a, n_e) Julia matrix is a numpy
array of shape (n_e, n_a) in C order (a fastest); tangent lanes are the leading axis; the Newton
vector x is reshape(x, n_endog, P) column-major, i.e. numpy x.reshape(P, 4) with columns
(Y, KS, r, w).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import HankError, c_dp, c_i32p, c_ip


def _f(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(c_dp)


class HouseholdBlock:
    """The heterogeneous-household block of a SequenceModel on one GPU.

    Replaces, for Float64 and ForwardDiff.Dual inputs alike, the reference's
    `model.value_fn` (KrusellSmith.jl:43-83), `BackwardIteration` (BackwardIteration.jl:46-116),
    `ForwardIteration` (ForwardIteration.jl:253-311) and — with `ks_configure` — the composite
    `fullFunction` and `JVP` (NewtonRaphson.jl:77-83, GeneralStructures.jl:542-550).
    """

    def __init__(self, grid, z, Pi, beta, gamma, borrow_cons, T, device=0):
        self._L = _lib.load()
        self.grid = _f(grid); self.z = _f(z); self.Pi = _f(Pi)
        self.n_a, self.n_e, self.T = len(self.grid), len(self.z), int(T)
        self.P = self.T - 1
        self.G = self.n_a * self.n_e
        self.n_endog, self.n_exog = 4, 1   # the built-in Krusell-Smith block (Y, KS, r, w; Z) until eq_configure
        self.n = 4 * self.P
        self.beta, self.gamma, self.borrow_cons = float(beta), float(gamma), float(borrow_cons)
        self._h = C.c_void_p()
        pi_cm = np.ascontiguousarray(self.Pi.T)  # column-major buffer of Π[e,e2]
        rc = self._L.hank_ctx_create(C.byref(self._h), int(device), self.n_a, self.n_e, self.T, _p(self.grid),
                                     _p(self.z), _p(pi_cm), self.beta, self.gamma, self.borrow_cons)
        if rc:
            msg = self._L.hank_last_error(self._h).decode() if self._h else "context allocation failed"
            if self._h:
                self._L.hank_ctx_destroy(self._h)
                self._h = None
            raise HankError(rc, msg)

    # -- plumbing ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.hank_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise HankError(rc, self._L.hank_last_error(self._h).decode())

    @property
    def handle(self):
        return self._h

    def sync(self):
        self._ck(self._L.hank_sync(self._h))

    def timer_start(self):
        self._ck(self._L.hank_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float(0)
        self._ck(self._L.hank_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self._L.hank_launch_count(self._h))

    def profile(self, enable=True):
        self._ck(self._L.hank_profile(self._h, int(bool(enable))))

    def kernel_times(self, reset=True):
        """{kernel: (total ms, launches)} of the four sweep kernels since the last reset."""
        ms = np.zeros(4); cnt = np.zeros(4, dtype=np.int64)
        self._ck(self._L.hank_kernel_times(self._h, _p(ms), cnt.ctypes.data_as(C.POINTER(C.c_int64)), int(reset)))
        names = ("backward_primal", "backward_tangent", "forward_primal", "forward_tangent")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(names)}

    def reserve_lanes(self, K):
        self._ck(self._L.hank_reserve_lanes(self._h, int(K)))

    # -- steady-state record ----------------------------------------------------------------
    def set_terminal(self, value_T):
        """ss_ending.value, the terminal ∂V/∂a (BackwardIteration.jl:85)."""
        self._ck(self._L.hank_set_terminal(self._h, _p(_f(value_T, (self.n_e, self.n_a)))))

    def set_initial_dist(self, D0):
        """ss_initial.D (ForwardIteration.jl:293)."""
        self._ck(self._L.hank_set_initial_dist(self._h, _p(_f(np.asarray(D0).reshape(self.n_e, self.n_a)))))

    # -- value_fn plug-in -------------------------------------------------------------------
    def egm_step(self, value_next, r, w, dvalue_next=None, dr=None, dw=None):
        """model.value_fn(value_next, xVals, model) -> (Value, KD[, dValue, dKD]) with K lanes."""
        ne, na = self.n_e, self.n_a
        vn = _f(value_next, (ne, na))
        K = 0 if dr is None else len(dr)
        dvn = None if dvalue_next is None else _f(dvalue_next, (K, ne, na))
        drr = None if K == 0 else _f(dr, (K,)); dww = None if K == 0 else _f(dw, (K,))
        value = np.empty((ne, na)); policy = np.empty((ne, na))
        dvalue = np.empty((K, ne, na)); dpolicy = np.empty((K, ne, na))
        self._ck(self._L.hank_egm_step(self._h, _p(vn), _p(dvn), float(r), float(w), K, _p(drr), _p(dww),
                                       _p(value), _p(policy), _p(dvalue), _p(dpolicy)))
        return value, policy, dvalue, dpolicy

    def vfi(self, r, w, dr=None, dw=None, eps=1e-6, max_iter=10_000):
        """Inner VFI of get_xVals (SteadyState.jl:132-141) on the device: returns
        (Value, KD, dValue, dKD, steps)."""
        ne, na = self.n_e, self.n_a
        K = 0 if dr is None else len(dr)
        drr = None if K == 0 else _f(dr, (K,)); dww = None if K == 0 else _f(dw, (K,))
        value = np.empty((ne, na)); policy = np.empty((ne, na))
        dvalue = np.empty((K, ne, na)); dpolicy = np.empty((K, ne, na))
        it = C.c_int(0)
        self._ck(self._L.hank_vfi(self._h, float(r), float(w), K, _p(drr), _p(dww), float(eps), int(max_iter),
                                  _p(value), _p(policy), _p(dvalue), _p(dpolicy), C.byref(it)))
        return value, policy, dvalue, dpolicy, it.value

    # -- sweeps -----------------------------------------------------------------------------
    def backward(self, r, w, dr=None, dw=None):
        """BackwardIteration: policies (and K tangent lanes) stay on the device."""
        P = self.P
        r = _f(r, (P,)); w = _f(w, (P,))
        K = 0 if dr is None else len(dr)
        drr = None if K == 0 else _f(dr, (K, P)); dww = None if K == 0 else _f(dw, (K, P))
        self._ck(self._L.hank_backward(self._h, _p(r), _p(w), K, _p(drr), _p(dww)))
        self._K = K

    def forward(self):
        """ForwardIteration on the device-resident policies: returns KD (P,), dKD (K, P)."""
        K = getattr(self, "_K", 0)
        KD = np.empty(self.P); dKD = np.empty((K, self.P))
        self._ck(self._L.hank_forward(self._h, _p(KD), _p(dKD)))
        return KD, dKD

    def forward_policies(self, policy, dpolicy=None):
        P, ne, na = self.P, self.n_e, self.n_a
        pol = _f(policy, (P, ne, na))
        K = 0 if dpolicy is None else len(dpolicy)
        dp = None if K == 0 else _f(dpolicy, (K, P, ne, na))
        KD = np.empty(P); dKD = np.empty((K, P))
        self._ck(self._L.hank_forward_policies(self._h, _p(pol), K, _p(dp), _p(KD), _p(dKD)))
        self._K = K
        return KD, dKD

    def block(self, r, w, dr=None, dw=None):
        P = self.P
        r = _f(r, (P,)); w = _f(w, (P,))
        K = 0 if dr is None else len(dr)
        drr = None if K == 0 else _f(dr, (K, P)); dww = None if K == 0 else _f(dw, (K, P))
        KD = np.empty(P); dKD = np.empty((K, P))
        self._ck(self._L.hank_block(self._h, _p(r), _p(w), K, _p(drr), _p(dww), _p(KD), _p(dKD)))
        self._K = K
        return KD, dKD

    def policy(self, t, lane=0):
        out = np.empty((self.n_e, self.n_a))
        self._ck(self._L.hank_get_policy(self._h, int(t), int(lane), _p(out)))
        return out

    def policies(self, lane=0):
        """Array(...) of the device policy handle: (P, n_e, n_a)."""
        return np.stack([self.policy(t, lane) for t in range(1, self.P + 1)])

    def dist(self, t):
        out = np.empty((self.n_e, self.n_a))
        self._ck(self._L.hank_get_dist(self._h, int(t), _p(out)))
        return out

    def value_first(self):
        out = np.empty((self.n_e, self.n_a))
        self._ck(self._L.hank_get_value_first(self._h, 0, _p(out)))
        return out

    def brackets(self, t):
        out = np.empty((self.n_e, self.n_a), dtype=np.int32)
        self._ck(self._L.hank_get_brackets(self._h, int(t), out.ctypes.data_as(c_i32p)))
        return out

    def lottery(self, policy):
        pol = _f(policy, (self.n_e, self.n_a))
        m = np.empty((self.n_e, self.n_a), dtype=np.int32); om = np.empty((self.n_e, self.n_a))
        self._ck(self._L.hank_lottery(self._h, _p(pol), m.ctypes.data_as(c_i32p), _p(om)))
        return m, om

    # -- Krusell-Smith full function ----------------------------------------------------------
    def ks_configure(self, alpha, delta, ss_start_KS):
        self._ck(self._L.hank_ks_configure(self._h, float(alpha), float(delta), float(ss_start_KS)))

    def eq_configure(self, program, n_endog, ir, iw, ss_start, ss_end):
        """The model's equilibrium equations as device bytecode (hankb200.equations.EquationProgram) instead of the
        built-in Krusell-Smith block: x becomes (P, n_endog), Z (n_exog, P); `ir`, `iw` are the rows of the household
        block's inputs r and w; ss_start / ss_end the boundary values of all variables (assemble_full_xMat)."""
        nv = len(program.names)
        n_exog = nv - n_endog - 1
        if program.n_eq != n_endog:
            raise ValueError("one equation per endogenous variable")
        eq_off, code, consts = program.arrays()
        ss0 = _f(ss_start, (nv,)); ss1 = _f(ss_end, (nv,))
        self._ck(self._L.hank_eq_configure(self._h, int(n_endog), int(n_exog), int(ir), int(iw), eq_off.ctypes.data_as(c_ip),
                                           code.ctypes.data_as(c_ip), len(program.consts), _p(consts), _p(ss0), _p(ss1)))
        self.n_endog, self.n_exog = int(n_endog), int(n_exog)
        self.n = self.n_endog * self.P

    def linearize(self, x, Z):
        """fullFunction(x): returns F(x) and keeps the linearisation for later JVPs at this x."""
        x = _f(x, (self.n,)); Z = _f(Z).reshape(-1)
        if Z.size != self.n_exog * self.P:
            raise ValueError("Z must hold n_exog paths of P periods")
        F = np.empty(self.n)
        self._ck(self._L.hank_ks_linearize(self._h, _p(x), _p(Z), _p(F)))
        return F

    def jvp(self, V):
        """J(x)·V for V of shape (K, n) (lane-major = Julia n x K), at the linearisation point."""
        V = _f(V)
        if V.ndim == 1:
            V = V[None]
        if V.shape[1] != self.n:
            raise ValueError("V must have shape (K, n)")
        JV = np.empty_like(V)
        self._ck(self._L.hank_ks_jvp(self._h, V.shape[0], _p(V), _p(JV)))
        return JV

    def fjvp(self, x, Z, V):
        """fullFunction(x) and J(x)·V in one call: returns (F, JV)."""
        x = _f(x, (self.n,)); Z = _f(Z).reshape(-1); V = _f(V)
        if V.ndim == 1:
            V = V[None]
        F = np.empty(self.n); JV = np.empty_like(V)
        self._ck(self._L.hank_ks_fjvp(self._h, _p(x), _p(Z), V.shape[0], _p(V), _p(F), _p(JV)))
        return F, JV

    def jacobian_columns(self, col_begin, col_end):
        """Columns col_begin..col_end-1 (1-based) of the Jacobian at the linearisation point,
        returned as (n, ncols) with J[:, j] the column (NumPy layout)."""
        ncols = col_end - col_begin
        out = np.empty((ncols, self.n))
        self._ck(self._L.hank_ks_jacobian_columns(self._h, int(col_begin), int(col_end), _p(out)))
        return out.T

    def jacobian_column_list(self, cols):
        """Jacobian columns `cols` (1-based, ascending) at the linearisation point: (n, len(cols))."""
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        out = np.empty((len(cols), self.n))
        self._ck(self._L.hank_ks_jacobian_column_list(self._h, len(cols), cols.ctypes.data_as(c_ip), _p(out)))
        return out.T

    def newton_solve(self, Jbar, x0, Z, eps=1e-9, eps_inner=1e-9, solver="lu"):
        """NewtonRaphsonHANK on the device. Jbar[i, j] NumPy (n, n)."""
        n = self.n
        Jcm = np.ascontiguousarray(np.asarray(Jbar, dtype=np.float64).T)
        if Jcm.shape != (n, n):
            raise ValueError("Jbar must be (n, n)")
        x0 = _f(x0, (n,)); Z = _f(Z, (self.P,))
        xo = np.empty(n); stats = np.zeros(8); inner = np.zeros(100, dtype=np.int32)
        rc = self._L.hank_newton_solve(self._h, _p(Jcm), _p(x0), _p(Z), float(eps), float(eps_inner),
                                       {"gmres": 0, "lu": 1, "lu_batched": 2}[solver], _p(xo), _p(stats), inner.ctypes.data_as(c_ip))
        self._ck(rc)
        outer = int(stats[0])
        return xo, dict(outer=outer, jvps=int(stats[1]), fevals=int(stats[2]), ynorm=float(stats[3]),
                        gmres_iters=int(stats[4]), inner=[int(v) for v in inner[:outer]])

    def dense_inverse(self, A):
        """A^-1 by the device's blocked Gauss-Jordan inverse (the preconditioner solve of the Newton driver)."""
        A = np.asarray(A, dtype=np.float64)
        n = A.shape[0]
        if A.shape != (n, n):
            raise ValueError("A must be square")
        Acm = np.ascontiguousarray(A.T)
        out = np.empty((n, n))
        self._ck(self._L.hank_dense_inverse(self._h, n, _p(Acm), _p(out)))
        return out.T

    # -- multi-GPU ------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        rc = _lib.load().hank_comm_unique_id(buf)
        if rc:
            raise HankError(rc, "ncclGetUniqueId failed (libnccl.so.2 not loadable?)")
        return bytes(buf)

    def comm_init(self, nranks, rank, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        self._ck(self._L.hank_comm_init(self._h, int(nranks), int(rank), buf))
