"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/tinyopt_b200.h declares (and nothing the header lacks), mirrors tinyopt's defaults, and
fails loudly — never falls back — when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tinyopt_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(tob200_\w+)\s*\(", src))


def test_header_and_binding_agree():
    from tinyopt_b200 import _lib
    assert header_functions() == set(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from tinyopt_b200 import _lib
    lib = _lib.load()  # resolves every entry of SYMBOLS or raises
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (tob200_\w+)", out))
    assert header_functions() <= exported
    assert exported <= header_functions(), "exported C symbols missing from the header"
    assert lib.tob200_version() == 100


def test_options_default_matches_reference_defaults():
    """optimizers/options.h:18-156 defaults; also identical to the oracle's."""
    from oracle import oracle as O
    import tinyopt_b200 as tb
    o, ref = tb.options(), O.default_options()
    for name, _ in o._fields_:
        assert getattr(o, name) == getattr(ref, name), name
    assert (o.max_iters, o.max_consec_failures, o.max_total_failures) == (50, 5, 0)
    assert o.min_error == C.c_float(1e-12).value and o.min_rerr_dec == C.c_float(1e-10).value
    assert o.min_step_norm2 == C.c_float(1e-14).value and o.min_grad_norm2 == C.c_float(1e-18).value
    assert o.damping_init == C.c_float(1e-4).value and o.good_factor == C.c_float(1.0 / 3.0).value
    assert (o.use_ldlt, o.H_is_full, o.save_last, o.use_squared_norm) == (1, 1, 1, 1)
    with pytest.raises(AttributeError):
        tb.options(no_such_field=1)


def test_struct_layouts():
    from tinyopt_b200 import _lib, api
    assert C.sizeof(_lib.Options) == 23 * 4
    assert C.sizeof(_lib.Result) == 56 == api.RESULT_DTYPE.itemsize
    assert api.RESULT_DTYPE.fields["stop_reason"][1] == _lib.Result.stop_reason.offset
    assert api.RESULT_DTYPE.fields["num_builds"][1] == _lib.Result.num_builds.offset


def test_helpers_without_gpu():
    from tinyopt_b200 import _lib
    lib = _lib.load()
    assert lib.tob200_tiled_elems(100, 30, 6) == 4 * 32 * 30 * 6
    assert lib.tob200_tiled_elems(0, 30, 6) == 0
    assert lib.tob200_kernel_family(1, 6) == 1 and lib.tob200_kernel_family(0, 12) == 1
    assert lib.tob200_kernel_family(1, 0) == 0 and lib.tob200_kernel_family(7, 3) == 0


def test_no_cpu_fallback():
    """Without a CUDA device context creation must fail with an error, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import tinyopt_b200 as tb
    from tinyopt_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.tob200_create(C.byref(h), 0, None)
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in lib.tob200_last_error(None)
    with pytest.raises(tb.TinyoptB200Error):
        tb.Context()
    # every compute entry point rejects a NULL context instead of doing anything
    assert lib.tob200_lm_run_f64(None, None, None, None, 0.1, 0, 1, 1, 1, None, None) == -1
    assert lib.tob200_build_solve_f32(None, None, None, 0, 1, 1, 1, None, None, None, None, None, None) == -1


def test_product_does_not_touch_the_oracle():
    """The product path may not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "tinyopt_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("CPU oracle's", "").replace("the oracle", "").replace("CPU oracle", "").lower() \
                    or f in (), f"{f} references the oracle"
    from tinyopt_b200 import _lib
    needed = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in needed
