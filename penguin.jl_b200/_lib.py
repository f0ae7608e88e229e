"""ctypes binding of libpenguin_b200.so (include/penguin_b200.h).  There is no CPU fallback: a missing library or a
missing CUDA device raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("PB200_LIB") or os.path.join(HERE, "libpenguin_b200.so")   # (PB200_LIB: kernel-variant experiments)

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class LevelSetC(C.Structure):
    _fields_ = [("kind", C.c_int), ("nballs", C.c_int), ("centers", dp), ("radii", dp), ("fluid_inside", C.c_int),
                ("hs_dim", C.c_int), ("hs_c", C.c_double)]


class SolverDesc(C.Structure):
    _fields_ = [("phase_type", C.c_int), ("time_type", C.c_int), ("ops1", C.c_void_p), ("ops2", C.c_void_p),
                ("D1", C.c_double), ("D2", C.c_double), ("D1_arr", dp), ("D2_arr", dp),
                ("ifc_kind", C.c_int), ("alpha", C.c_double), ("beta", C.c_double),
                ("alpha1", C.c_double), ("alpha2", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double)]


class StepIn(C.Structure):
    _fields_ = [("scheme", C.c_int), ("dt", C.c_double), ("f_const", (C.c_double * 2) * 2), ("f_arr", (dp * 2) * 2),
                ("g_const", C.c_double * 2), ("g_arr", dp * 2)]


class KrylovOpts(C.Structure):
    _fields_ = [("method", C.c_int), ("rtol", C.c_double), ("atol", C.c_double), ("maxit", C.c_int),
                ("warm_start", C.c_int), ("check_every", C.c_int), ("path", C.c_int), ("precond", C.c_int)]


class StepStats(C.Structure):
    _fields_ = [("iters", C.c_int), ("converged", C.c_int), ("rnorm", C.c_double), ("bnorm", C.c_double),
                ("solve_ms", C.c_double), ("setup_ms", C.c_double), ("dof_bulk", C.c_int64), ("dof_ifc", C.c_int64),
                ("launches", C.c_int64), ("apply_ms", C.c_double), ("apply_launches", C.c_int64),
                ("apply_cells_uniform", C.c_int64), ("apply_cells_general", C.c_int64),
                ("kernel_ms", C.c_double * 8), ("kernel_launches", C.c_int64 * 8), ("apply_cells_fast", C.c_int64), ("band_cells", C.c_int64), ("band_rows", C.c_int64)]


# every symbol include/penguin_b200.h declares (tests/test_abi.py checks the .so exports them all)
SYMBOLS = {
    "pb200_init": ([C.POINTER(C.c_void_p), C.c_int], C.c_int),
    "pb200_nccl_unique_id": ([C.c_char_p], C.c_int),
    "pb200_init_dist": ([C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_char_p], C.c_int),
    "pb200_init_multi": ([C.POINTER(C.c_void_p), ip, C.c_int], C.c_int),
    "pb200_finalize": ([C.c_void_p], C.c_int),
    "pb200_last_error": ([C.c_void_p], C.c_char_p),
    "pb200_sync": ([C.c_void_p], C.c_int),
    "pb200_launch_count": ([C.c_void_p], C.c_int64),
    "pb200_stream": ([C.c_void_p], C.c_uint64),
    "pb200_set_profiling": ([C.c_void_p, C.c_int], C.c_int),
    "pb200_capacity_create": ([C.c_void_p, C.c_int, ip, dp, dp, C.POINTER(LevelSetC), C.c_int, C.POINTER(C.c_void_p)], C.c_int),
    "pb200_capacity_import": ([C.c_void_p, C.c_int, ip, dp, dp] + [dp] * 8 + [C.POINTER(C.c_void_p)], C.c_int),
    "pb200_capacity_export": ([C.c_void_p] + [dp] * 8, C.c_int),
    "pb200_capacity_local": ([C.c_void_p, ip, ip, C.POINTER(C.c_int64)], C.c_int),
    "pb200_capacity_destroy": ([C.c_void_p], C.c_int),
    "pb200_ops_create": ([C.c_void_p, C.POINTER(C.c_void_p)], C.c_int),
    "pb200_ops_grad": ([C.c_void_p, dp, dp], C.c_int),
    "pb200_ops_set_convection": ([C.c_void_p, dp, dp], C.c_int),
    "pb200_ops_export_convection": ([C.c_void_p, dp, dp], C.c_int),
    "pb200_ops_div": ([C.c_void_p, dp, dp, dp], C.c_int),
    "pb200_ops_export_wdag": ([C.c_void_p, dp], C.c_int),
    "pb200_ops_destroy": ([C.c_void_p], C.c_int),
    "pb200_solver_create": ([C.c_void_p, C.POINTER(SolverDesc), C.POINTER(C.c_void_p)], C.c_int),
    "pb200_solver_set_border": ([C.c_void_p, C.c_int, C.c_int, C.c_double, dp], C.c_int),
    "pb200_solver_set_state": ([C.c_void_p, dp], C.c_int),
    "pb200_solver_get_state": ([C.c_void_p, dp], C.c_int),
    "pb200_solver_get_state_async": ([C.c_void_p, dp], C.c_int),
    "pb200_solver_error_norms": ([C.c_void_p, C.c_int, dp, C.c_double, C.c_int, dp], C.c_int),
    "pb200_poly_coefs": ([C.c_double, C.c_double, dp], C.c_int),
    "pb200_solver_wait_state": ([C.c_void_p], C.c_int),
    "pb200_solver_step": ([C.c_void_p, C.POINTER(StepIn), C.POINTER(KrylovOpts), C.POINTER(StepStats)], C.c_int),
    "pb200_solver_destroy": ([C.c_void_p], C.c_int),
}

_lib = None


class PenguinB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpenguin_b200 error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise ImportError(f"{SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(libpenguin_b200 is the only compute path; there is no CPU fallback)")
        l = C.CDLL(SO)
        for name, (args, res) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.argtypes, fn.restype = args, res
        _lib = l
    return _lib


def check(rc, ctx=None, allow=()):
    if rc != 0 and rc not in allow:
        msg = lib().pb200_last_error(ctx)
        raise PenguinB200Error(rc, msg.decode() if msg else "?")
    return rc
