"""ctypes binding of libhankb200.so (include/hankb200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present the
import / context creation raises.  Build with `python julia-newtonraphsonhank_b200/build.py`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libhankb200.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_i32p = C.POINTER(C.c_int32)
c_fp = C.POINTER(C.c_float)
ctx_p = C.c_void_p

# name -> (restype, argtypes); every symbol include/hankb200.h declares
SIGNATURES = {
    "hank_ctx_create": (C.c_int, [C.POINTER(ctx_p), C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp,
                                  C.c_double, C.c_double, C.c_double]),
    "hank_ctx_destroy": (None, [ctx_p]),
    "hank_last_error": (C.c_char_p, [ctx_p]),
    "hank_version": (C.c_char_p, []),
    "hank_sync": (C.c_int, [ctx_p]),
    "hank_timer_start": (C.c_int, [ctx_p]),
    "hank_timer_stop": (C.c_int, [ctx_p, c_fp]),
    "hank_launch_count": (C.c_int64, [ctx_p]),
    "hank_reserve_lanes": (C.c_int, [ctx_p, C.c_int]),
    "hank_profile": (C.c_int, [ctx_p, C.c_int]),
    "hank_kernel_times": (C.c_int, [ctx_p, c_dp, C.POINTER(C.c_int64), C.c_int]),
    "hank_set_terminal": (C.c_int, [ctx_p, c_dp]),
    "hank_set_initial_dist": (C.c_int, [ctx_p, c_dp]),
    "hank_egm_step": (C.c_int, [ctx_p, c_dp, c_dp, C.c_double, C.c_double, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "hank_vfi": (C.c_int, [ctx_p, C.c_double, C.c_double, C.c_int, c_dp, c_dp, C.c_double, C.c_int, c_dp, c_dp, c_dp, c_dp, c_ip]),
    "hank_backward": (C.c_int, [ctx_p, c_dp, c_dp, C.c_int, c_dp, c_dp]),
    "hank_forward": (C.c_int, [ctx_p, c_dp, c_dp]),
    "hank_forward_policies": (C.c_int, [ctx_p, c_dp, C.c_int, c_dp, c_dp, c_dp]),
    "hank_block": (C.c_int, [ctx_p, c_dp, c_dp, C.c_int, c_dp, c_dp, c_dp, c_dp]),
    "hank_block_dev": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hank_get_policy": (C.c_int, [ctx_p, C.c_int, C.c_int, c_dp]),
    "hank_get_dist": (C.c_int, [ctx_p, C.c_int, c_dp]),
    "hank_get_value_first": (C.c_int, [ctx_p, C.c_int, c_dp]),
    "hank_get_brackets": (C.c_int, [ctx_p, C.c_int, c_i32p]),
    "hank_lottery": (C.c_int, [ctx_p, c_dp, c_i32p, c_dp]),
    "hank_ks_configure": (C.c_int, [ctx_p, C.c_double, C.c_double, C.c_double]),
    "hank_eq_configure": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip, C.c_int, c_dp, c_dp, c_dp]),
    "hank_ks_linearize": (C.c_int, [ctx_p, c_dp, c_dp, c_dp]),
    "hank_ks_jvp": (C.c_int, [ctx_p, C.c_int, c_dp, c_dp]),
    "hank_ks_fjvp": (C.c_int, [ctx_p, c_dp, c_dp, C.c_int, c_dp, c_dp, c_dp]),
    "hank_ks_linearize_dev": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hank_ks_jvp_dev": (C.c_int, [ctx_p, C.c_int, C.c_void_p, C.c_void_p]),
    "hank_ks_jacobian_columns": (C.c_int, [ctx_p, C.c_int, C.c_int, c_dp]),
    "hank_ks_jacobian_columns_dev": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_void_p]),
    "hank_ks_jacobian_column_list_dev": (C.c_int, [ctx_p, C.c_int, c_ip, C.c_void_p]),
    "hank_ks_jacobian_column_list": (C.c_int, [ctx_p, C.c_int, c_ip, c_dp]),
    "hank_newton_solve": (C.c_int, [ctx_p, c_dp, c_dp, c_dp, C.c_double, C.c_double, C.c_int, c_dp, c_dp, c_ip]),
    "hank_dense_inverse": (C.c_int, [ctx_p, C.c_int, c_dp, c_dp]),
    "hank_comm_unique_id": (C.c_int, [C.c_void_p]),
    "hank_comm_init": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_void_p]),
    "hank_allgather_columns_dev": (C.c_int, [ctx_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "hank_allgather_columns": (C.c_int, [ctx_p, c_dp, C.c_size_t, c_dp]),
    "hank_comm_destroy": (C.c_int, [ctx_p]),
}

_lib = None


def load():
    """Loads libhankb200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `python julia-newtonraphsonhank_b200/build.py` "
                              "(hankb200 has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class HankError(RuntimeError):
    """Raised for any non-zero status, like the Julia wrapper's error(hank_last_error())."""

    def __init__(self, code, msg):
        super().__init__(f"hankb200 status {code}: {msg}")
        self.code = code
        self.msg = msg
