"""ctypes loader of libpiqp_b200.so (the C-ABI in include/piqp_b200.h).

The library is the product; there is NO CPU fallback: if it is missing or fails to load, importing the
solver classes raises.  `build()` compiles it in-tree with nvcc for sm_100a.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpiqp_b200.so")
_lib = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class Settings(C.Structure):
    """b200qp_settings == piqp_settings (interfaces/c/include/piqp_typedef.h:75-104)."""
    _fields_ = [
        ("rho_init", C.c_double), ("delta_init", C.c_double), ("eps_abs", C.c_double), ("eps_rel", C.c_double),
        ("check_duality_gap", C.c_int), ("eps_duality_gap_abs", C.c_double), ("eps_duality_gap_rel", C.c_double),
        ("infeasibility_threshold", C.c_double), ("reg_lower_limit", C.c_double), ("reg_finetune_lower_limit", C.c_double),
        ("reg_finetune_primal_update_threshold", C.c_int), ("reg_finetune_dual_update_threshold", C.c_int),
        ("max_iter", C.c_int), ("max_factor_retires", C.c_int), ("preconditioner_scale_cost", C.c_int),
        ("preconditioner_reuse_on_update", C.c_int), ("preconditioner_iter", C.c_int), ("tau", C.c_double),
        ("kkt_solver", C.c_int), ("iterative_refinement_always_enabled", C.c_int),
        ("iterative_refinement_eps_abs", C.c_double), ("iterative_refinement_eps_rel", C.c_double),
        ("iterative_refinement_max_iter", C.c_int), ("iterative_refinement_min_improvement_rate", C.c_double),
        ("iterative_refinement_static_regularization_eps", C.c_double),
        ("iterative_refinement_static_regularization_rel", C.c_double),
        ("verbose", C.c_int), ("compute_timings", C.c_int),
    ]


class Info(C.Structure):
    """b200qp_info == piqp_info (interfaces/c/include/piqp_typedef.h:116-159)."""
    _fields_ = [
        ("status", C.c_int), ("iter", C.c_int),
        ("rho", C.c_double), ("delta", C.c_double), ("mu", C.c_double), ("sigma", C.c_double),
        ("primal_step", C.c_double), ("dual_step", C.c_double),
        ("primal_res", C.c_double), ("primal_res_rel", C.c_double), ("dual_res", C.c_double), ("dual_res_rel", C.c_double),
        ("primal_res_reg", C.c_double), ("primal_res_reg_rel", C.c_double), ("dual_res_reg", C.c_double), ("dual_res_reg_rel", C.c_double),
        ("primal_prox_inf", C.c_double), ("dual_prox_inf", C.c_double), ("prev_primal_res", C.c_double), ("prev_dual_res", C.c_double),
        ("primal_obj", C.c_double), ("dual_obj", C.c_double), ("duality_gap", C.c_double), ("duality_gap_rel", C.c_double),
        ("factor_retires", C.c_int), ("reg_limit", C.c_double), ("no_primal_update", C.c_int), ("no_dual_update", C.c_int),
        ("setup_time", C.c_double), ("update_time", C.c_double), ("solve_time", C.c_double),
        ("kkt_factor_time", C.c_double), ("kkt_solve_time", C.c_double), ("run_time", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("factor_calls", C.c_longlong), ("kkt_solve_calls", C.c_longlong), ("backend_solves", C.c_longlong),
        ("ip_iterations", C.c_longlong), ("lockstep_iterations", C.c_int),
        ("factor_ms", C.c_double), ("solve_ms", C.c_double), ("total_ms", C.c_double),
        ("kernel_launches", C.c_ulonglong),
        ("assemble_ms", C.c_double), ("cholesky_ms", C.c_double), ("backend_solve_ms", C.c_double),
        ("assemble_launches", C.c_longlong), ("cholesky_calls", C.c_longlong), ("backend_solve_launch_groups", C.c_longlong),
    ]


# every symbol include/piqp_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "b200_last_error", "b200_kernel_launch_count", "b200_timeline_dump", "b200_device_count",
    "b200kkt_dense_create", "b200kkt_sparse_create", "b200kkt_sparse_info", "b200_sparse_ldlt_symbolic", "b200_sparse_ldlt_symbolic_mode", "b200kkt_multistage_create", "b200kkt_update_data",
    "b200kkt_factor", "b200kkt_solve", "b200kkt_eval_P_x", "b200kkt_eval_A_xn_and_AT_xt", "b200kkt_eval_G_xn_and_GT_xt",
    "b200kkt_clone", "b200kkt_print_info", "b200kkt_destroy", "b200kkt_dense_get_kkt",
    "b200qp_set_default_settings_dense", "b200qp_set_default_settings_sparse", "b200qp_setup_dense",
    "b200qp_update_dense", "b200qp_setup_sparse", "b200qp_setup_sparse_ex", "b200qp_get_sparse_perm", "b200qp_update_sparse", "b200qp_multistage_blocks", "b200kkt_multistage_blocks", "b200qp_update_settings", "b200qp_solve", "b200qp_get_result", "b200qp_get_info",
    "b200qp_get_stats", "b200qp_get_trace", "b200qp_set_profiling", "b200qp_get_work", "b200qp_cleanup", "b200qp_bench_factor_solve",
]


def build(verbose=False):
    """Compile libpiqp_b200.so in-tree (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libpiqp_b200.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libpiqp_b200.so is missing (%s): run piqp_b200.build() / __graft_entry__.build(). "
                           "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.b200_last_error.restype = C.c_char_p
    L.b200_kernel_launch_count.restype = C.c_ulonglong
    L.b200kkt_clone.restype = C.c_void_p
    L.b200kkt_clone.argtypes = [C.c_void_p]
    for f in ("b200kkt_destroy", "b200kkt_print_info", "b200qp_cleanup"):
        getattr(L, f).restype = None
        getattr(L, f).argtypes = [C.c_void_p]
    _lib = L
    return L


def check(code, what=""):
    if code < 0:
        raise RuntimeError("%s failed (%d): %s" % (what, code, lib().b200_last_error().decode()))
    return code
