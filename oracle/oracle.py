"""ctypes binding of the CPU oracle (oracle/libtinyopt_oracle.so).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` leg.  The product package (tinyopt_b200/) never imports this.
See oracle/tinyopt_oracle.h for what the oracle restates and its parity status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtinyopt_oracle.so")

# stop_reasons.h:14-43
STOP = {
    -4: "kOutOfMemory", -3: "kSolverFailed", -2: "kSystemHasNaNOrInf", -1: "kSkipped",
    0: "kNone", 1: "kMinError", 2: "kMinRelError", 3: "kMinDeltaNorm", 4: "kMinGradNorm",
    5: "kMaxIters", 6: "kMaxNoDecr", 7: "kMaxConsecNoDecr", 8: "kTimedOut", 9: "kUserStopped",
}
K = {v: k for k, v in STOP.items()}


class Options(C.Structure):
    """too_options == numeric subset of tinyopt::Options (optimizers/options.h:18-156)."""
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
    _fields_ = [
        ("final_cost", C.c_double), ("final_rerr_dec", C.c_double),
        ("final_num_residuals", C.c_int32), ("stop_reason", C.c_int32), ("num_iters", C.c_int32),
        ("num_failures", C.c_int32), ("num_consec_failures", C.c_int32),
        ("history_len", C.c_int32), ("last_lambda", C.c_double), ("last_prev_lambda", C.c_double),
        ("min_margin", C.c_double), ("sign_margin", C.c_double), ("thr_margin", C.c_double),
    ]


RESULT_DTYPE = np.dtype([
    ("final_cost", "f8"), ("final_rerr_dec", "f8"), ("final_num_residuals", "i4"),
    ("stop_reason", "i4"), ("num_iters", "i4"), ("num_failures", "i4"),
    ("num_consec_failures", "i4"), ("history_len", "i4"), ("last_lambda", "f8"),
    ("last_prev_lambda", "f8"), ("min_margin", "f8"), ("sign_margin", "f8"),
    ("thr_margin", "f8")], align=True)
assert RESULT_DTYPE.itemsize == C.sizeof(Result)


class Trace(C.Structure):
    _fields_ = [("cap", C.c_int32), ("errs", C.POINTER(C.c_double)),
                ("deltas2", C.POINTER(C.c_double)), ("successes", C.POINTER(C.c_int32)),
                ("lambdas", C.POINTER(C.c_double)), ("rebuilt", C.POINTER(C.c_int32)),
                ("xs", C.POINTER(C.c_double))]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc)."""
    srcs = [os.path.join(_HERE, f) for f in ("tinyopt_oracle.c", "oracle_impl.inc", "tinyopt_oracle.h", "Makefile")]
    sos = [_SO] + [os.path.join(_HERE, f"libtinyopt_oracle_fast_{v}.so") for v in ("v3", "v4")]
    stale = any((not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs) for so in sos)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None
_lib_fast = None


def _declare(l: C.CDLL) -> C.CDLL:
    l.too_max_threads.restype = C.c_int
    l.too_options_default.argtypes = [C.POINTER(Options)]
    for suf in ("f32", "f64"):
        for name in ("too_solve_ldlt", "too_inv_cov", "too_build_solve", "too_optimize",
                     "too_synth_lm_run"):
            getattr(l, f"{name}_{suf}").restype = C.c_int
        for name in ("too_synth_generate", "too_synth_eval"):
            getattr(l, f"{name}_{suf}").restype = None
    return l


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:  # no compiler on the box: use the prebuilt .so that travelled
            if not os.path.exists(_SO):
                raise
        _lib = _declare(C.CDLL(_SO))
    return _lib


def fast_variant() -> str:
    """'v4' (AVX-512) if this host's cores have it, else 'v3' (AVX2 + FMA)."""
    try:
        with open("/proc/cpuinfo") as f:
            flags = f.read()
        return "v4" if (" avx512f" in flags and " avx512vl" in flags and " avx512bw" in flags and " avx512dq" in flags) else "v3"
    except OSError:
        return "v3"


def lib_fast() -> C.CDLL:
    """The -DTOO_FAST build: the same arithmetic scheduled for a wide core (oracle/Makefile); its
    results are bit-identical to lib()'s, it only serves as the faster CPU baseline."""
    global _lib_fast
    if _lib_fast is None:
        lib()  # builds everything
        _lib_fast = _declare(C.CDLL(os.path.join(_HERE, f"libtinyopt_oracle_fast_{fast_variant()}.so")))
    return _lib_fast


def default_options(**kw) -> Options:
    o = Options()
    lib().too_options_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def _suf(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(dtype)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ct(dtype):
    return C.c_float if np.dtype(dtype) == np.float32 else C.c_double


def max_threads() -> int:
    return lib().too_max_threads()


def solve_ldlt(A: np.ndarray, b: np.ndarray):
    """math.h:232-240.  Returns x or None (not positive / numerical issue)."""
    A = np.ascontiguousarray(A)
    b = np.ascontiguousarray(b, dtype=A.dtype)
    n = b.shape[0]
    x = np.empty_like(b)
    ok = getattr(lib(), f"too_solve_ldlt_{_suf(A.dtype)}")(C.c_int(n), _ptr(A), _ptr(b), _ptr(x))
    return x if ok else None


def inv_cov(A: np.ndarray):
    """math.h:44-57."""
    A = np.ascontiguousarray(A)
    n = A.shape[0]
    out = np.empty_like(A)
    ok = getattr(lib(), f"too_inv_cov_{_suf(A.dtype)}")(C.c_int(n), _ptr(A), _ptr(out))
    return out if ok else None


def build_solve(J: np.ndarray, r: np.ndarray, lam: float = 0.0):
    """One Build+Solve for one problem.  Returns dict(dx, cost, H, g, status)."""
    J = np.ascontiguousarray(J)
    r = np.ascontiguousarray(r, dtype=J.dtype)
    m, n = J.shape
    dx = np.zeros(n, J.dtype)
    H = np.zeros((n, n), J.dtype)
    g = np.zeros(n, J.dtype)
    cost = C.c_double(0)
    st = getattr(lib(), f"too_build_solve_{_suf(J.dtype)}")(
        C.c_int(m), C.c_int(n), _ptr(J), _ptr(r), _ct(J.dtype)(lam), _ptr(dx), C.byref(cost),
        _ptr(H), _ptr(g))
    return dict(dx=dx, cost=cost.value, H=H, g=g, status=st)


@dataclass
class Output:
    """POD mirror of tinyopt::Output (output.h:26-145)."""
    x: np.ndarray
    final_cost: float
    final_rerr_dec: float
    final_num_residuals: int
    stop_reason: int
    num_iters: int
    num_failures: int
    num_consec_failures: int
    last_lambda: float
    last_prev_lambda: float
    min_margin: float
    errs: list = field(default_factory=list)
    deltas2: list = field(default_factory=list)
    successes: list = field(default_factory=list)
    lambdas: list = field(default_factory=list)
    rebuilt: list = field(default_factory=list)
    xs: list = field(default_factory=list)
    final_hessian: np.ndarray | None = None

    @property
    def stop_name(self) -> str:
        return STOP[self.stop_reason]

    def Succeeded(self) -> bool:  # output.h:30
        return self.stop_reason >= 0

    def Converged(self) -> bool:  # output.h:33-35
        return 1 <= self.stop_reason < 5

    def Covariance(self):  # output.h:81-96 (rescaled = false)
        if self.final_hessian is None:
            return None
        return inv_cov(self.final_hessian)


def optimize(x0, acc, options: Options | None = None, dtype=np.float64) -> Output:
    """tinyopt::Optimize(x, acc, options) through the oracle (optimizer.h:243-327).

    `acc(x, grad, H)` follows docs/API.md:37-57: grad/H are numpy views to fill in place (H is n*n,
    zeroed) or both None for the cost-only call; it returns cost, or (cost, num_residuals), or a
    residual vector (-> squared norm, size; cost.h:27-30).
    """
    dtype = np.dtype(dtype)
    ct = _ct(dtype)
    x = np.array(np.atleast_1d(x0), dtype=dtype).copy()
    n = x.shape[0]
    opt = options if options is not None else default_options()

    CB = C.CFUNCTYPE(None, C.POINTER(ct), C.c_int, C.POINTER(ct), C.POINTER(ct),
                     C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p)

    def _cb(xp, n_, gp, hp, costp, nresp, _user):
        xv = np.ctypeslib.as_array(xp, shape=(n_,)) if n_ > 0 else np.zeros(0, dtype)
        if gp:
            g = np.ctypeslib.as_array(gp, shape=(n_,))
            H = np.ctypeslib.as_array(hp, shape=(n_, n_))
        else:
            g = H = None
        ret = acc(xv, g, H)
        if isinstance(ret, tuple):
            costp[0], nresp[0] = float(ret[0]), int(ret[1])
        elif isinstance(ret, np.ndarray):
            rv = ret.astype(dtype).ravel()
            c = dtype.type(0)
            for v in rv:  # Cost(residuals): squaredNorm in Scalar (cost.h:27-30)
                c = dtype.type(c + v * v)
            costp[0], nresp[0] = float(c), int(rv.size)
        else:
            costp[0], nresp[0] = float(ret), 1  # Cost(Scalar) (cost.h:22)

    cb = CB(_cb)
    cap = int(opt.max_iters) + 3
    errs = np.zeros(cap); d2 = np.zeros(cap); succ = np.zeros(cap, np.int32)
    lams = np.zeros(cap); reb = np.zeros(cap, np.int32); xs = np.zeros((cap, max(n, 1)))
    tr = Trace(cap, errs.ctypes.data_as(C.POINTER(C.c_double)), d2.ctypes.data_as(C.POINTER(C.c_double)),
               succ.ctypes.data_as(C.POINTER(C.c_int32)), lams.ctypes.data_as(C.POINTER(C.c_double)),
               reb.ctypes.data_as(C.POINTER(C.c_int32)), xs.ctypes.data_as(C.POINTER(C.c_double)))
    res = Result()
    fh = np.zeros((n, n)) if (n > 0 and n <= 4096) else None
    getattr(lib(), f"too_optimize_{_suf(dtype)}")(
        _ptr(x) if n > 0 else None, C.c_int(n), cb, None, C.byref(opt), C.byref(res), C.byref(tr), _ptr(fh))
    h, it = res.history_len, res.num_iters
    return Output(
        x=x, final_cost=res.final_cost, final_rerr_dec=res.final_rerr_dec,
        final_num_residuals=res.final_num_residuals, stop_reason=res.stop_reason,
        num_iters=res.num_iters, num_failures=res.num_failures,
        num_consec_failures=res.num_consec_failures, last_lambda=res.last_lambda,
        last_prev_lambda=res.last_prev_lambda, min_margin=res.min_margin,
        errs=list(errs[:h]), deltas2=list(d2[:h]), successes=[bool(v) for v in succ[:h]],
        lambdas=list(lams[:it]), rebuilt=[bool(v) for v in reb[:it]], xs=[xs[i, :n].copy() for i in range(min(it, cap))],
        final_hessian=fh if (fh is not None and opt.save_last and res.stop_reason not in (-1, -4)) else None)


# ---- synthetic family -----------------------------------------------------------
SEED = 20261017
ALPHA = 0.1
SIGMA = 1e-2


def synth_generate(B: int, m: int, n: int, dtype, p0: int = 0, seed: int = SEED,
                   alpha: float = ALPHA, sigma: float = SIGMA):
    """Returns A[B,m,n], y[B,m], xstar[B,n], x0[B,n] (problem-major)."""
    dtype = np.dtype(dtype)
    A = np.empty((B, m, n), dtype); y = np.empty((B, m), dtype)
    xs = np.empty((B, n), dtype); x0 = np.empty((B, n), dtype)
    ct = _ct(dtype)
    getattr(lib(), f"too_synth_generate_{_suf(dtype)}")(
        C.c_uint64(seed), C.c_int64(p0), C.c_int64(B), C.c_int(m), C.c_int(n), ct(alpha), ct(sigma),
        _ptr(A), _ptr(y), _ptr(xs), _ptr(x0))
    return A, y, xs, x0


def synth_eval(A: np.ndarray, y: np.ndarray, x: np.ndarray, alpha: float = ALPHA):
    """r[B,m], J[B,m,n] of the family at x[B,n]."""
    B, m, n = A.shape
    r = np.empty((B, m), A.dtype); J = np.empty((B, m, n), A.dtype)
    fn = getattr(lib(), f"too_synth_eval_{_suf(A.dtype)}")
    ct = _ct(A.dtype)
    x = np.ascontiguousarray(x, dtype=A.dtype)
    for b in range(B):
        fn(C.c_int(m), C.c_int(n), _ptr(A[b]), _ptr(y[b]), ct(alpha), _ptr(x[b]), _ptr(r[b]), _ptr(J[b]))
    return r, J


def synth_lm_run(A: np.ndarray, y: np.ndarray, x0: np.ndarray, options: Options | None = None,
                 alpha: float = ALPHA, nthreads: int = 0, fast: bool = False, reverse_rows: bool = False,
                 want_hessian: bool = False, robust: tuple | None = None, numdiff: tuple | None = None):
    """Batched LM over the family.  Returns (x[B,n], results structured array, threads used).
    fast=True runs the -DTOO_FAST build (same results, scheduled for throughput).  reverse_rows=True
    (census only) sums the residual rows m-1 .. 0: another valid order, NOT the canonical one."""
    A = np.ascontiguousarray(A); y = np.ascontiguousarray(y, dtype=A.dtype)
    B, m, n = A.shape
    x = np.array(x0, dtype=A.dtype, order="C", copy=True)
    res = np.zeros(B, RESULT_DTYPE)
    opt = options if options is not None else default_options()
    l = lib_fast() if fast else lib()
    l.too_census_set_row_order(C.c_int(1 if reverse_rows else 0))
    # robust = (kind 1..7, th2): every residual goes through that M-estimator inside the accumulation
    l.too_synth_set_robust(C.c_int(robust[0] if robust else 0), C.c_double(robust[1] if robust else 1.0))
    # numdiff = (method 1 kForward / 2 kCentral / 3 kFastCentral, h or 0 for FloatEpsilon): diff/num_diff.h instead of the analytic J
    l.too_synth_set_numdiff(C.c_int(numdiff[0] if numdiff else 0), C.c_double(numdiff[1] if numdiff else 0.0))
    fh = np.zeros((B, n, n)) if want_hessian else None
    try:
        fn = getattr(l, f"too_synth_lm_run_fh_{_suf(A.dtype)}")
        fn.restype = C.c_int
        used = fn(C.c_int64(B), C.c_int(m), C.c_int(n), _ptr(A), _ptr(y), _ct(A.dtype)(alpha), _ptr(x),
                  C.byref(opt), _ptr(res), C.c_int(nthreads), _ptr(fh))
    finally:
        l.too_census_set_row_order(C.c_int(0))
        l.too_synth_set_robust(C.c_int(0), C.c_double(1.0))
        l.too_synth_set_numdiff(C.c_int(0), C.c_double(0.0))
    if want_hessian:
        return x, res, used, fh   # Output::final_hessian per problem (optimizer.h:313-316), un-damped
    return x, res, used
