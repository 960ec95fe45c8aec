"""ctypes binding of libatacom_b200.so (C ABI declared in include/atacom_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at
import time, and every compute entry point raises when the library reports an
error (e.g. no CUDA device).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ATACOM_B200_LIB") or os.path.join(_HERE, "libatacom_b200.so")   # override: kernel experiments

MAX_Q, MAX_F, MAX_G, MAX_C, ENV_PARAMS = 8, 4, 16, 20, 24

OK = 0
ST_RANK_DEFICIENT, ST_COLUMN_DROPPED, ST_SLACK_PIVOT, ST_NONFINITE, ST_DENSE_PATH = 1, 2, 4, 8, 16
VARIANT_ATACOM, VARIANT_ERROR_CORRECTION = 0, 1
BIAS_JDOT_QDOT, BIAS_OMEGA_X_V = 0, 1
BASIS_LAPACK, BASIS_CANONICAL = 0, 1
ST_LAPACK_PATH = 32
HOST_AUTO, HOST_STAGED, HOST_ZERO_COPY, HOST_HYBRID = 0, 1, 2, 3


class AtacomParams(ctypes.Structure):
    """Mirror of `struct AtacomParams` (include/atacom_b200.h)."""
    _fields_ = [
        ("K_f", ctypes.c_float * MAX_F),
        ("K_g", ctypes.c_float * MAX_G),
        ("K_c", ctypes.c_float * MAX_C),
        ("K_q", ctypes.c_float * MAX_Q),
        ("vel_max", ctypes.c_float * MAX_Q),
        ("acc_max", ctypes.c_float * MAX_Q),
        ("dt", ctypes.c_float),
        ("rref_tol", ctypes.c_float),
        ("variant", ctypes.c_int32),
        ("bias_mode", ctypes.c_int32),
        ("clip_acc", ctypes.c_int32),
        ("basis_mode", ctypes.c_int32),
        ("env", ctypes.c_double * ENV_PARAMS),
    ]

    def copy(self):
        other = AtacomParams()
        ctypes.memmove(ctypes.byref(other), ctypes.byref(self), ctypes.sizeof(self))
        return other

    def flat(self):
        """All fields flattened in declaration order (ints as floats) — the layout the host
        test harness takes."""
        out = []
        for name, ctype in self._fields_:
            v = getattr(self, name)
            out.extend(list(v) if hasattr(v, "__len__") else [v])
        return [float(x) for x in out]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "rl_on_manifold_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C rl_on_manifold_b200/csrc`. There is no CPU fallback." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

_f = ctypes.c_void_p       # device / host float*
_u8 = ctypes.c_void_p
_i64 = ctypes.c_int64
_P = ctypes.POINTER(AtacomParams)
_stream = ctypes.c_void_p

# name -> argtypes; every symbol include/atacom_b200.h declares
SIGNATURES = {
    "atacom_version": ([], ctypes.c_char_p),
    "atacom_error_string": ([ctypes.c_int], ctypes.c_char_p),
    "atacom_launch_count": ([], ctypes.c_int64),
    "atacom_circle_default_params": ([_P], ctypes.c_int),
    "atacom_planar_default_params": ([_P], ctypes.c_int),
    "atacom_iiwa_default_params": ([_P, ctypes.c_int], ctypes.c_int),
    "atacom_point_reach_default_params": ([_P], ctypes.c_int),
    "atacom_circle_step": ([_f, _f, _f, _f, _f, _f, _u8, _f, _i64, _P, _stream], ctypes.c_int),
    "atacom_planar_step": ([_f, _f, _f, _f, _f, _f, _u8, _f, _i64, _P, _stream], ctypes.c_int),
    "atacom_iiwa_step": ([ctypes.c_int, _f, _f, _f, _f, _f, _f, _u8, _f, _i64, _P, _stream], ctypes.c_int),
    "atacom_iiwa_step_gather": ([ctypes.c_int, _f, _f, _f, _f, _f, _f, _u8, _i64, _P, _stream,
                                 ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _i64], ctypes.c_int),
    "atacom_iiwa_step_gather_sync": ([ctypes.c_int, _f, _f, _f, _f, _f, _f, _u8, _i64, _P, _stream,
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _i64,
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_int], ctypes.c_int),
    "atacom_iiwa_substeps_workspace_doubles": ([ctypes.c_int], ctypes.c_int),
    "atacom_iiwa_step_substeps": ([ctypes.c_int, ctypes.c_int, _f, _f, _f, _f, _f, _f, _u8, ctypes.c_void_p, _i64, _P,
                                   _stream], ctypes.c_int),
    "atacom_circle_slack_init": ([_f, _f, _f, _u8, _i64, _P, _stream], ctypes.c_int),
    "atacom_planar_slack_init": ([_f, _f, _f, _u8, _i64, _P, _stream], ctypes.c_int),
    "atacom_iiwa_slack_init": ([ctypes.c_int, _f, _f, _f, _u8, _i64, _P, _stream], ctypes.c_int),
    "atacom_point_reach_step": ([ctypes.c_int, _f, _f, _f, _f, _f, _f, _f, _f, _u8, _f, _i64, _P, _stream],
                                ctypes.c_int),
    "atacom_point_reach_slack_init": ([ctypes.c_int, _f, _f, _f, _u8, _i64, _P, _stream], ctypes.c_int),
    "atacom_circle_constraint_stats": ([_f, _f, _f, _f, _i64, _P, _stream], ctypes.c_int),
    "atacom_planar_constraint_stats": ([_f, _f, _f, _f, _i64, _P, _stream], ctypes.c_int),
    "atacom_iiwa_constraint_stats": ([ctypes.c_int, _f, _f, _f, _f, _i64, _P, _stream], ctypes.c_int),
    "atacom_circle_rollout": ([_f, _f, _f, _f, _f, _u8, _i64, ctypes.c_int, _P, _stream], ctypes.c_int),
    "atacom_point_reach_rollout": ([ctypes.c_int, _f, _f, _f, _f, _f, ctypes.c_double, _f, _f, _u8, _i64, ctypes.c_int,
                                    _P, _stream], ctypes.c_int),
    "atacom_generic_supported": ([ctypes.c_int, ctypes.c_int, ctypes.c_int], ctypes.c_int),
    "atacom_generic_step": ([ctypes.c_int, ctypes.c_int, ctypes.c_int, _f, _f, _f, _f, _f, _f, _f, _f, _u8, _f,
                             _i64, _P, _stream], ctypes.c_int),
    "atacom_host_ctx_create": ([ctypes.POINTER(ctypes.c_void_p), _i64, ctypes.c_int], ctypes.c_int),
    "atacom_host_ctx_set_mode": ([ctypes.c_void_p, ctypes.c_int], ctypes.c_int),
    "atacom_host_ctx_destroy": ([ctypes.c_void_p], ctypes.c_int),
    "atacom_iiwa_step_host": ([ctypes.c_void_p, ctypes.c_int, _f, _f, _f, _f, _f, _f, _u8, _i64, _P],
                              ctypes.c_int),
    "atacom_circle_step_host": ([ctypes.c_void_p, _f, _f, _f, _f, _f, _f, _u8, _i64, _P], ctypes.c_int),
    "atacom_planar_step_host": ([ctypes.c_void_p, _f, _f, _f, _f, _f, _f, _u8, _i64, _P], ctypes.c_int),
    "atacom_point_reach_step_host": ([ctypes.c_void_p, ctypes.c_int, _f, _f, _f, _f, _f, _f, _f, _f, _u8, _i64, _P],
                                     ctypes.c_int),
    "atacom_generic_step_host": ([ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f, _f, _f, _f, _f, _f,
                                  _f, _f, _u8, _i64, _P], ctypes.c_int),
    "atacom_spin_timeouts": ([ctypes.POINTER(ctypes.c_uint)], ctypes.c_int),
}

for _name, (_args, _res) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header/library mismatch
    _fn.argtypes = _args
    _fn.restype = _res


class AtacomError(RuntimeError):
    pass


def check(code):
    if code != OK:
        raise AtacomError("libatacom_b200: %s (code %d)" % (lib.atacom_error_string(code).decode(), code))


def version():
    return lib.atacom_version().decode()


def launch_count():
    return int(lib.atacom_launch_count())


def spin_timeouts():
    """Device-side waits that ran out of their clock budget on the current device (0 in a healthy run)."""
    n = ctypes.c_uint(0)
    check(lib.atacom_spin_timeouts(ctypes.byref(n)))
    return int(n.value)


def default_params(family, n_ctrl_joints=6):
    p = AtacomParams()
    if family == "circle":
        check(lib.atacom_circle_default_params(ctypes.byref(p)))
    elif family == "planar":
        check(lib.atacom_planar_default_params(ctypes.byref(p)))
    elif family == "iiwa":
        check(lib.atacom_iiwa_default_params(ctypes.byref(p), n_ctrl_joints))
    elif family == "point_reach":
        check(lib.atacom_point_reach_default_params(ctypes.byref(p)))
    else:
        raise ValueError("unknown env family %r" % (family,))
    return p
