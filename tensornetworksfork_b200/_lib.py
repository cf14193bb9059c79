"""ctypes binding of the C ABI declared in include/tn_b200.h.

The shared library is built in-tree by ``make`` (or ``__graft_entry__.build()``).  There is no
CPU fallback: if the library is missing, or a CUDA tensor is required and absent, the caller gets
an exception.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtn_b200.so")

c_double_p = ctypes.c_void_p
c_int_p = ctypes.c_void_p
i64 = ctypes.c_int64
i32 = ctypes.c_int
f64 = ctypes.c_double
vp = ctypes.c_void_p


class tn_factor(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("ld", ctypes.c_int64), ("m", ctypes.c_int32), ("div", ctypes.c_int32),
                ("map_kind", ctypes.c_int32), ("_pad", ctypes.c_int32)]


class tn_operator(ctypes.Structure):
    """Mirror of tn_operator (include/tn_b200.h); the two callbacks are stored as raw addresses (0 = NULL)."""
    _fields_ = [("fa", ctypes.POINTER(tn_factor)), ("fb", ctypes.POINTER(tn_factor)), ("fc", ctypes.POINTER(tn_factor)),
                ("w", ctypes.c_void_p), ("rows", ctypes.c_int64), ("apply", ctypes.c_void_p), ("apply_ctx", ctypes.c_void_p),
                ("allreduce", ctypes.c_void_p), ("allreduce_ctx", ctypes.c_void_p), ("sigma", ctypes.c_void_p),
                ("ridge", ctypes.c_double), ("P", ctypes.c_int64), ("apply_is_global", ctypes.c_int32), ("_pad", ctypes.c_int32)]


APPLY_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p)
ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p)
OP = ctypes.POINTER(tn_operator)

FP = ctypes.POINTER(tn_factor)
IP = ctypes.POINTER(ctypes.c_int)

# name -> (restype, argtypes); mirrors include/tn_b200.h one to one
PROTOTYPES = {
    "tn_version": (i32, []),
    "tn_last_error": (ctypes.c_char_p, []),
    "tn_sm_count": (i32, []),
    "tn_launch_count": (i64, []),
    "tn_env_update": (i32, [vp, i64, i32, vp, i64, i32, i32, i32, vp, vp, i64, vp, i64, i32, vp, i64, i32, i32, vp]),
    "tn_class_rows": (i32, [vp, vp, vp, vp, vp, i64, i32, i32, i32, vp]),
    "tn_gram_ksplit": (i32, [i64, i32, i32, i32, i32]),
    "tn_gram_kr3": (i32, [i32, FP, FP, FP, vp, i64, vp, vp, i32, i32, vp]),
    "tn_gram_tc_flush_rows": (i32, [i32]),
    "tn_rhs_ksplit": (i32, [i64, i32, i32, i32]),
    "tn_rhs_kr3": (i32, [FP, FP, FP, vp, i64, vp, vp, i32, i32, vp]),
    "tn_generic_ksplit": (i32, [i64, i32, i32]),
    "tn_gram_generic": (i32, [FP, FP, FP, vp, vp, vp, i32, vp, i64, vp, i32, vp, i32, i32, vp]),
    "tn_gram_sigma": (i32, [vp, IP, IP, vp, vp]),
    "tn_gram_expand": (i32, [vp, IP, IP, vp, f64, vp, i64, vp]),
    "tn_rhs_prepare": (i32, [vp, vp, vp, f64, vp, i64, vp]),
    "tn_cholesky_work_elems": (i64, [i64]),
    "tn_cholesky_solve": (i32, [vp, i64, i64, vp, vp, vp, vp]),
    "tn_cholesky_mixed_work_elems": (i64, [i64]),
    "tn_cholesky_solve_mixed": (i32, [vp, i64, i64, vp, vp, vp, f64, i32, vp, vp]),
    "tn_update_node": (i32, [vp, vp, i64, f64, i32, f64, vp, vp]),
    "tn_qr": (i32, [vp, i32, i32, vp, vp]),
    "tn_matvec_work_elems": (i64, [i64, i32, i32, i32]),
    "tn_bmm": (i32, [vp, i64, i64, i64, vp, i64, i64, i64, vp, i64, i32, i32, i32, i32, vp]),
    "tn_outer_rows": (i32, [vp, i64, i32, i32, vp, i64, i32, vp, i64, vp, i32, vp]),
    "tn_rows_dot": (i32, [vp, i64, i32, vp, i64, i32, i64, vp, i64, vp]),
    "tn_matvec_kr3": (i32, [FP, FP, FP, vp, i64, vp, vp, vp, vp]),
    "tn_cg_work_elems": (i64, [OP]),
    "tn_cg": (i32, [OP, vp, i64, vp, vp, vp, vp, i32, i32, f64, i32, vp, vp, vp]),
    "tn_minres_work_elems": (i64, [OP]),
    "tn_minres": (i32, [OP, vp, vp, i32, i32, f64, i32, vp, vp, vp]),
    "tn_lanczos_work_elems": (i64, [OP, i32]),
    "tn_lanczos": (i32, [OP, vp, vp, vp, i32, f64, i32, vp, vp, vp]),
    "tn_cholesky_factor": (i32, [vp, i64, i64, i32, vp, vp, vp]),
    "tn_cholesky_apply": (i32, [vp, i64, i64, vp, vp, vp, vp]),
    "tn_gram_trace": (i32, [FP, FP, FP, vp, i64, vp, i32, vp]),
}

_lib = None


class TnError(RuntimeError):
    pass


def load():
    """Load libtn_b200.so; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TnError(f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()) first; there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().tn_last_error().decode("utf-8", "replace")
        raise TnError(f"{what} failed (rc={rc}): {msg}")
