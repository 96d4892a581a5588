"""ctypes loader for libtinyopt_b200.so (the C-ABI of include/tinyopt_b200.h).

The shared library is built in-tree by `make -C tinyopt_b200/csrc` (see __graft_entry__.build()).
There is deliberately no fallback: if the library is missing, or no CUDA device is present when a
context is created, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TOB200_LIB_OVERRIDE: development A/B switch (tools/build_variant.sh builds libtinyopt_b200_<name>.so with
# experiment flags); unset in every test / bench run that is reported
LIB_PATH = os.environ.get("TOB200_LIB_OVERRIDE") or os.path.join(_HERE, "libtinyopt_b200.so")


class Options(C.Structure):
    """tob200_options == numeric subset of tinyopt::Options (optimizers/options.h:18-156)."""
    _fields_ = [
        ("solver_type", C.c_int32), ("check_final_cost", C.c_int32),
        ("use_step_quality_approx", C.c_int32), ("grad_clipping", C.c_float),
        ("use_ldlt", C.c_int32), ("H_is_full", C.c_int32), ("check_min_H_diag", C.c_float),
        ("save_last", C.c_int32), ("use_squared_norm", C.c_int32), ("downscale_by_2", C.c_int32),
        ("normalize", C.c_int32), ("max_iters", C.c_int32), ("min_error", C.c_float),
        ("min_rerr_dec", C.c_float), ("min_step_norm2", C.c_float), ("min_grad_norm2", C.c_float),
        ("max_total_failures", C.c_int32), ("max_consec_failures", C.c_int32),
        ("damping_init", C.c_float), ("damping_min", C.c_float), ("damping_max", C.c_float),
        ("good_factor", C.c_float), ("bad_factor", C.c_float),
    ]


class Result(C.Structure):
    """tob200_result == POD subset of tinyopt::Output (output.h:122-142)."""
    _fields_ = [
        ("final_cost", C.c_double), ("final_rerr_dec", C.c_double), ("last_lambda", C.c_double),
        ("last_prev_lambda", C.c_double), ("final_num_residuals", C.c_int32),
        ("stop_reason", C.c_int32), ("num_iters", C.c_int32), ("num_failures", C.c_int32),
        ("num_consec_failures", C.c_int32), ("num_builds", C.c_int32),
    ]


RESULT_FIELDS = [(n, "f8" if t is C.c_double else "i4") for n, t in Result._fields_]

# every symbol include/tinyopt_b200.h declares: (name, restype, argtypes)
_vp, _i, _i64, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
_f, _d = C.c_float, C.c_double
_PO = C.POINTER(Options)
SYMBOLS = {
    "tob200_version": (_i, []),
    "tob200_create": (_i, [C.POINTER(_vp), _i, _vp]),
    "tob200_destroy": (_i, [_vp]),
    "tob200_sync": (_i, [_vp]),
    "tob200_last_error": (C.c_char_p, [_vp]),
    "tob200_launch_count": (_i64, [_vp]),
    "tob200_last_elapsed_ms": (_i, [_vp, C.POINTER(_f)]),
    "tob200_set_exact": (_i, [_vp, _i]),
    "tob200_options_default": (None, [_PO]),
    "tob200_device_alloc": (_i, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "tob200_device_free": (_i, [_vp, _vp]),
    "tob200_copy_to_device": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "tob200_copy_to_host": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "tob200_tiled_elems": (_i64, [_i64, _i, _i]),
    "tob200_kernel_family": (_i, [_i, _i]),
    "tob200_retile_f32": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "tob200_retile_f64": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "tob200_build_solve_f32": (_i, [_vp, _vp, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tob200_build_solve_f64": (_i, [_vp, _vp, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tob200_jtj_f32": (_i, [_vp, _vp, _vp, _i64, _i, _i, _vp]),
    "tob200_solve_ldlt_f32": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp]),
    "tob200_inv_cov_f32": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _vp]),
    "tob200_inv_cov_f64": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _vp]),
    "tob200_last_phase_ms": (_i, [_vp, _i, C.POINTER(_f), C.POINTER(_i)]),
    "tob200_lm_run_f32": (_i, [_vp, _PO, _vp, _vp, _f, _i, _i64, _i, _i, _vp, _vp]),
    "tob200_lm_run_f64": (_i, [_vp, _PO, _vp, _vp, _d, _i, _i64, _i, _i, _vp, _vp]),
    "tob200_lm_run_ex_f32": (_i, [_vp, _PO, _vp, _vp, _f, _i, _i64, _i, _i, _vp, _vp, _vp]),
    "tob200_lm_run_ex_f64": (_i, [_vp, _PO, _vp, _vp, _d, _i, _i64, _i, _i, _vp, _vp, _vp]),
    "tob200_lm_run_host_f32": (_i, [_vp, _PO, _vp, _vp, _f, _i, _i64, _i, _i, _vp, _vp]),
    "tob200_lm_run_host_f64": (_i, [_vp, _PO, _vp, _vp, _d, _i, _i64, _i, _i, _vp, _vp]),
    "tob200_solver_create": (_i, [_vp, _i, _i64, _i, _PO, C.POINTER(_vp)]),
    "tob200_solver_destroy": (_i, [_vp]),
    "tob200_solver_reset": (_i, [_vp, _vp]),
    "tob200_solver_x": (_vp, [_vp]),
    "tob200_solver_needs": (_vp, [_vp]),
    "tob200_solver_step_f32": (_i, [_vp, _vp, _vp, _i, _i]),
    "tob200_solver_step_f64": (_i, [_vp, _vp, _vp, _i, _i]),
    "tob200_solver_create_ex": (_i, [_vp, _i, _i64, _i, _PO, _i, C.POINTER(_vp)]),
    "tob200_solver_step_cost_f32": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "tob200_solver_step_cost_f64": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "tob200_solver_step_hg_sparse_f32": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "tob200_solver_step_hg_sparse_f64": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "tob200_solver_step_hg_f32": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "tob200_solver_step_hg_f64": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "tob200_solver_num_active": (_i, [_vp, C.POINTER(_i64)]),
    "tob200_solver_results": (_i, [_vp, _vp]),
    "tob200_solver_final_hessian": (_i, [_vp, _vp]),
    "tob200_solver_covariance": (_i, [_vp, _vp, _vp, _vp]),
    "tob200_synth_generate_f32": (_i, [_vp, _u64, _i64, _i64, _i, _i, _f, _f, _i, _vp, _vp, _vp, _vp]),
    "tob200_synth_generate_f64": (_i, [_vp, _u64, _i64, _i64, _i, _i, _d, _d, _i, _vp, _vp, _vp, _vp]),
    "tob200_synth_eval_f32": (_i, [_vp, _vp, _vp, _f, _i, _i64, _i, _i, _vp, _vp, _vp]),
    "tob200_synth_eval_f64": (_i, [_vp, _vp, _vp, _d, _i, _i64, _i, _i, _vp, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA extension; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `make -C tinyopt_b200/csrc` "
                "(or __graft_entry__.build()); tinyopt_b200 has no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
