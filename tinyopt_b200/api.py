"""Host-side interface of tinyopt_b200 over the C-ABI (ctypes + torch tensors for device memory).

Names follow the reference: `Options` (optimizers/options.h), `Output` (output.h), `StopReason`
(stop_reasons.h), `Context.optimize_batch` == one `tinyopt::Optimize()` per problem
(optimize.h:17-77), `BatchSolver` == a batch of `Optimizer_<SolverLM>` driven from the host
(optimizers/optimizer.h:332-539 Step).
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import Options  # noqa: F401

TILE32 = 0
PROBLEM_MAJOR = 1


class TinyoptB200Error(RuntimeError):
    pass


class StopReason(enum.IntEnum):
    """stop_reasons.h:14-43."""
    kOutOfMemory = -4
    kSolverFailed = -3
    kSystemHasNaNOrInf = -2
    kSkipped = -1
    kNone = 0
    kMinError = 1
    kMinRelError = 2
    kMinDeltaNorm = 3
    kMinGradNorm = 4
    kMaxIters = 5
    kMaxNoDecr = 6
    kMaxConsecNoDecr = 7
    kTimedOut = 8
    kUserStopped = 9


def options(**kw) -> Options:
    """tinyopt::Options{} with overrides (flattened names: damping_init == lm.damping_init ...)."""
    o = Options()
    _lib.load().tob200_options_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(f"tinyopt::Options has no numeric field '{k}'")
        setattr(o, k, v)
    return o


RESULT_DTYPE = np.dtype(_lib.RESULT_FIELDS, align=True)
assert RESULT_DTYPE.itemsize == C.sizeof(_lib.Result)


@dataclass
class Output:
    """Batched tinyopt::Output (output.h:26-145): one entry per problem."""
    x: torch.Tensor            # [B, n] solutions (device)
    results: np.ndarray        # structured array, RESULT_DTYPE, host
    final_hessian: torch.Tensor | None = None   # [B, n, n] float64 (device): Output::final_hessian, un-damped

    @property
    def num_iters(self):
        return self.results["num_iters"]

    @property
    def stop_reason(self):
        return self.results["stop_reason"]

    @property
    def final_cost(self):
        return self.results["final_cost"]

    def Succeeded(self):  # output.h:30
        return self.results["stop_reason"] >= 0

    def Converged(self):  # output.h:33-35
        sr = self.results["stop_reason"]
        return (sr >= 1) & (sr < 5)


def _suf(dtype: torch.dtype) -> str:
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise TypeError(f"tinyopt_b200 computes in float32 or float64, got {dtype}")


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _layout_of(J: torch.Tensor, layout: int | None) -> int:
    """TILE32 tensors are 4-D [tiles, m, n, 32], problem-major ones 3-D [B, m, n]."""
    if layout is not None:
        return layout
    if J.dim() == 4 and J.shape[-1] == 32:
        return TILE32
    if J.dim() == 3:
        return PROBLEM_MAJOR
    raise ValueError(f"cannot infer the layout of a tensor of shape {tuple(J.shape)}")


def to_tile32(a: torch.Tensor) -> torch.Tensor:
    """[B, m, n] or [B, m] problem-major -> TILE32 [ceil(B/32), m, n, 32] (pure torch; host helper
    for tests — the library re-tiles PROBLEM_MAJOR inputs itself on the device)."""
    squeeze = a.dim() == 2
    if squeeze:
        a = a.unsqueeze(-1)
    B, m, n = a.shape
    nt = (B + 31) // 32
    pad = torch.zeros((nt * 32, m, n), dtype=a.dtype, device=a.device)
    pad[:B] = a
    out = pad.view(nt, 32, m, n).permute(0, 2, 3, 1).contiguous()
    return out.view(nt, m, 32) if squeeze else out


def from_tile32(a: torch.Tensor, B: int) -> torch.Tensor:
    """Inverse of to_tile32."""
    if a.dim() == 3:
        nt, m, _ = a.shape
        return a.permute(0, 2, 1).reshape(nt * 32, m)[:B].contiguous()
    nt, m, n, _ = a.shape
    return a.permute(0, 3, 1, 2).reshape(nt * 32, m, n)[:B].contiguous()


class Context:
    """One tob200_ctx: one GPU, one stream.  Not thread-safe."""

    def __init__(self, device: int | torch.device | None = None, stream: torch.cuda.Stream | None = None):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise TinyoptB200Error("no CUDA device: tinyopt_b200 has no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.stream = stream if stream is not None else torch.cuda.current_stream(self.device)
        h = C.c_void_p()
        # torch's default stream has handle 0, which the C-ABI reads as "create your own stream":
        # name the legacy default stream explicitly (cudaStreamLegacy == 0x1) so that the kernels are
        # ordered with the torch ops that produce / consume their buffers
        handle = self.stream.cuda_stream or 1
        rc = self._lib.tob200_create(C.byref(h), self.device.index or 0, C.c_void_p(handle))
        if rc != 0:
            raise TinyoptB200Error(f"tob200_create failed ({rc}): {self._lib.tob200_last_error(None).decode()}")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tob200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int, what: str):
        if rc != 0:
            raise TinyoptB200Error(f"{what} failed ({rc}): {self._lib.tob200_last_error(self._h).decode()}")

    def sync(self):
        self._ck(self._lib.tob200_sync(self._h), "tob200_sync")

    @property
    def launch_count(self) -> int:
        return int(self._lib.tob200_launch_count(self._h))

    def set_exact(self, exact: bool | int = True):
        """Mid-n float runs (13 <= n <= 55): bit-exact warp-per-problem kernel instead of the tensor-core one.
        exact=2: float 56 <= n <= 512 on the general (bit-exact) family as well, instead of the tcgen05 one."""
        self._ck(self._lib.tob200_set_exact(self._h, int(exact)), "set_exact")

    def last_elapsed_ms(self) -> float:
        ms = C.c_float(0)
        self._ck(self._lib.tob200_last_elapsed_ms(self._h, C.byref(ms)), "tob200_last_elapsed_ms")
        return float(ms.value)

    def kernel_family(self, dtype: torch.dtype, n: int) -> int:
        return int(self._lib.tob200_kernel_family(0 if dtype == torch.float32 else 1, n))

    # ---- layout -------------------------------------------------------------------------------
    def retile(self, a: torch.Tensor) -> torch.Tensor:
        """PROBLEM_MAJOR [B,m,n] / [B,m] -> TILE32 on the device (tob200_retile_*)."""
        squeeze = a.dim() == 2
        B, m = a.shape[0], a.shape[1]
        n = 1 if squeeze else a.shape[2]
        a = a.contiguous()
        nt = (B + 31) // 32
        out = torch.empty((nt, m, 32) if squeeze else (nt, m, n, 32), dtype=a.dtype, device=a.device)
        fn = getattr(self._lib, f"tob200_retile_{_suf(a.dtype)}")
        self._ck(fn(self._h, _p(a), B, m, n, _p(out)), "tob200_retile")
        return out

    # ---- a1+a3+a5+a6 ----------------------------------------------------------------------------
    def build_solve(self, J: torch.Tensor, r: torch.Tensor, lam: torch.Tensor | None = None, *,
                    B: int | None = None, layout: int | None = None, want_H: bool = False, want_g: bool = False):
        """One Build + Solve per problem from materialised residual blocks.

        PROBLEM_MAJOR: J [B,m,n], r [B,m].  TILE32: J [nt,m,n,32], r [nt,m,32] and B given.
        Returns dict(dx [B,n], cost [B] f64, status [B] i32, H [B,n,n]?, g [B,n]?).
        """
        layout = _layout_of(J, layout)
        if layout == PROBLEM_MAJOR:
            B, m, n = J.shape
        else:
            _, m, n, _ = J.shape
            if B is None:
                raise ValueError("TILE32 input needs B (the number of problems)")
        dt, dev = J.dtype, J.device
        J = J.contiguous(); r = r.contiguous()
        dx = torch.zeros((B, n), dtype=dt, device=dev)
        cost = torch.zeros((B,), dtype=torch.float64, device=dev)
        status = torch.zeros((B,), dtype=torch.int32, device=dev)
        H = torch.zeros((B, n, n), dtype=dt, device=dev) if want_H else None
        g = torch.zeros((B, n), dtype=dt, device=dev) if want_g else None
        if lam is not None:
            lam = lam.to(dtype=dt, device=dev).contiguous()
        fn = getattr(self._lib, f"tob200_build_solve_{_suf(dt)}")
        self._ck(fn(self._h, _p(J), _p(r), layout, B, m, n, _p(lam), _p(dx), _p(cost), _p(H), _p(g), _p(status)),
                 "tob200_build_solve")
        out = dict(dx=dx, cost=cost, status=status)
        if want_H:
            out["H"] = H
        if want_g:
            out["g"] = g
        return out

    # ---- a1 / a6 alone (the reference's secondary seams) ------------------------------------------
    def jtj(self, J: torch.Tensor, row_scale: torch.Tensor | None = None) -> torch.Tensor:
        """H = Jᵀ diag(s²) J on the tensor cores (tob200_jtj_f32).  J [B,m,n] float32."""
        B, m, n = J.shape
        H = torch.empty((B, n, n), dtype=torch.float32, device=J.device)
        rs = None if row_scale is None else row_scale.to(dtype=torch.float32, device=J.device).contiguous()
        self._ck(self._lib.tob200_jtj_f32(self._h, _p(J.contiguous()), _p(rs), B, m, n, _p(H)), "tob200_jtj_f32")
        return H

    def solve_ldlt(self, A: torch.Tensor, b: torch.Tensor):
        """tinyopt::SolveLDLT(A, b) per problem (tob200_solve_ldlt_f32).  Returns (x [B,n], status [B])."""
        B, n, _ = A.shape
        x = torch.zeros((B, n), dtype=torch.float32, device=A.device)
        status = torch.zeros((B,), dtype=torch.int32, device=A.device)
        self._ck(self._lib.tob200_solve_ldlt_f32(self._h, _p(A.contiguous()), _p(b.contiguous()), B, n, _p(x), _p(status)),
                 "tob200_solve_ldlt_f32")
        return x, status

    def inv_cov(self, H: torch.Tensor, want_cov: bool = True, want_max_std: bool = True):
        """tinyopt::InvCov(H) and MaxStdDev per problem (tob200_inv_cov_*; math.h:44-57,
        solvers/lm.h:176-187).  H [B,n,n] float32 / float64, only the upper triangle is read.
        Returns (cov [B,n,n] or None, max_std [B] or None, status [B]: 1 where the reference gives nullopt)."""
        B, n, _ = H.shape
        suf = "f64" if H.dtype == torch.float64 else "f32"
        cov = torch.zeros_like(H) if want_cov else None
        ms = torch.zeros((B,), dtype=H.dtype, device=H.device) if want_max_std else None
        status = torch.zeros((B,), dtype=torch.int32, device=H.device)
        fn = getattr(self._lib, f"tob200_inv_cov_{suf}")
        self._ck(fn(self._h, _p(H.contiguous()), B, n, _p(cov) if cov is not None else None,
                    _p(ms) if ms is not None else None, _p(status)), f"tob200_inv_cov_{suf}")
        return cov, ms, status

    def last_phase_ms(self, phase: int):
        """(ms, launches) the last large-n call spent in phase 0 eval / 1 JᵀJ / 2 solve."""
        ms, cnt = C.c_float(0), C.c_int(0)
        self._ck(self._lib.tob200_last_phase_ms(self._h, phase, C.byref(ms), C.byref(cnt)), "tob200_last_phase_ms")
        return float(ms.value), int(cnt.value)

    # ---- a7-a10 ---------------------------------------------------------------------------------
    def optimize_batch(self, A: torch.Tensor, y: torch.Tensor, x0: torch.Tensor, opt: Options | None = None, *,
                       alpha: float = 0.1, layout: int | None = None, results: torch.Tensor | None = None,
                       sync: bool = True, want_hessian: bool = False) -> Output:
        """One tinyopt::Optimize() per problem of the polynomial family, device resident
        (tob200_lm_run_*).  x0 [B,n] is copied; the returned Output holds the solutions.
        want_hessian: also Output.final_hessian (tob200_lm_run_ex_*; needs options.save_last, the default)."""
        opt = opt if opt is not None else options()
        layout = _layout_of(A, layout)
        B, n = x0.shape
        m = A.shape[1]
        dt, dev = A.dtype, A.device
        x = x0.to(dtype=dt, device=dev).clone().contiguous()
        if results is None:
            results = torch.empty((B, C.sizeof(_lib.Result)), dtype=torch.uint8, device=dev)
        ct = C.c_float if dt == torch.float32 else C.c_double
        fh = None
        if want_hessian:
            fh = torch.zeros((B, n, n), dtype=torch.float64, device=dev)
            fn = getattr(self._lib, f"tob200_lm_run_ex_{_suf(dt)}")
            self._ck(fn(self._h, C.byref(opt), _p(A.contiguous()), _p(y.contiguous()), ct(alpha), layout, B, m, n,
                        _p(x), _p(results), _p(fh)), "tob200_lm_run_ex")
        else:
            fn = getattr(self._lib, f"tob200_lm_run_{_suf(dt)}")
            self._ck(fn(self._h, C.byref(opt), _p(A.contiguous()), _p(y.contiguous()), ct(alpha), layout, B, m, n,
                        _p(x), _p(results)), "tob200_lm_run")
        if not sync:
            return Output(x=x, results=results, final_hessian=fh)  # raw device buffer; caller decodes after sync
        self.sync()
        return Output(x=x, results=decode_results(results), final_hessian=fh)

    def optimize_batch_host(self, A: np.ndarray, y: np.ndarray, x: np.ndarray, opt: Options | None = None, *,
                            alpha: float = 0.1, layout: int | None = None, B: int | None = None,
                            results: np.ndarray | None = None):
        """Same through HOST buffers (numpy, ideally pinned): H2D + run + D2H inside
        (tob200_lm_run_host_*).  x is updated in place; returns the results array."""
        opt = opt if opt is not None else options()
        if layout is None:
            layout = TILE32 if A.ndim == 4 else PROBLEM_MAJOR
        if B is None:
            B = x.shape[0]
        n = x.shape[1]
        m = A.shape[1]
        if results is None:
            results = np.zeros(B, RESULT_DTYPE)
        suf = "f32" if A.dtype == np.float32 else "f64"
        fn = getattr(self._lib, f"tob200_lm_run_host_{suf}")
        ct = C.c_float if A.dtype == np.float32 else C.c_double
        self._ck(fn(self._h, C.byref(opt), A.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), ct(alpha),
                    layout, B, m, n, x.ctypes.data_as(C.c_void_p), results.ctypes.data_as(C.c_void_p)),
                 "tob200_lm_run_host")
        return results

    # ---- synthetic family -----------------------------------------------------------------------
    def synth_generate(self, B: int, m: int, n: int, dtype: torch.dtype, *, p0: int = 0, seed: int = 20261017,
                       alpha: float = 0.1, sigma: float = 1e-2, layout: int = TILE32):
        """A, y (in `layout`), xstar, x0 ([B,n]) generated on the device (tob200_synth_generate_*)."""
        nt = (B + 31) // 32
        dev = self.device
        if layout == TILE32:
            A = torch.empty((nt, m, n, 32), dtype=dtype, device=dev)
            y = torch.empty((nt, m, 32), dtype=dtype, device=dev)
        else:
            A = torch.empty((B, m, n), dtype=dtype, device=dev)
            y = torch.empty((B, m), dtype=dtype, device=dev)
        xs = torch.empty((B, n), dtype=dtype, device=dev)
        x0 = torch.empty((B, n), dtype=dtype, device=dev)
        fn = getattr(self._lib, f"tob200_synth_generate_{_suf(dtype)}")
        ct = C.c_float if dtype == torch.float32 else C.c_double
        self._ck(fn(self._h, seed, p0, B, m, n, ct(alpha), ct(sigma), layout, _p(A), _p(y), _p(xs), _p(x0)),
                 "tob200_synth_generate")
        return A, y, xs, x0

    def synth_eval(self, A: torch.Tensor, y: torch.Tensor, x: torch.Tensor, *, alpha: float = 0.1,
                   layout: int | None = None):
        """Residual blocks r, J of the family at x, in `layout` (tob200_synth_eval_*)."""
        layout = _layout_of(A, layout)
        B, n = x.shape
        m = A.shape[1]
        r = torch.empty_like(y)
        J = torch.empty_like(A)
        fn = getattr(self._lib, f"tob200_synth_eval_{_suf(A.dtype)}")
        ct = C.c_float if A.dtype == torch.float32 else C.c_double
        self._ck(fn(self._h, _p(A), _p(y), ct(alpha), layout, B, m, n, _p(x.contiguous()), _p(r), _p(J)),
                 "tob200_synth_eval")
        return r, J


def decode_results(buf: torch.Tensor) -> np.ndarray:
    """Device byte buffer of tob200_result[B] -> host structured array."""
    return buf.cpu().numpy().view(RESULT_DTYPE).reshape(-1).copy()


class BatchSolver:
    """A batch of `Optimizer_<SolverLM>` whose state lives on the device; the caller evaluates the
    residual blocks (its own lambda, AD, ...) at `x` and feeds them to `step`."""

    def __init__(self, ctx: Context, B: int, n: int, dtype: torch.dtype, opt: Options | None = None, *,
                 general: bool = False):
        """Any n <= 2048 (n > 55 runs on the general kernel family); `general=True` forces that family for a small n
        too (TOB200_SOLVER_GENERAL: needed by `step(..., cost=...)`)."""
        self.ctx, self.B, self.n, self.dtype = ctx, B, n, dtype
        self.opt = opt if opt is not None else options()
        h = C.c_void_p()
        ctx._ck(ctx._lib.tob200_solver_create_ex(ctx._h, 0 if dtype == torch.float32 else 1, B, n,
                                                 C.byref(self.opt), 1 if general else 0, C.byref(h)),
                "tob200_solver_create_ex")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._lib.tob200_solver_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, x0: torch.Tensor):
        x0 = x0.to(dtype=self.dtype, device=self.ctx.device).contiguous()
        self.ctx._ck(self.ctx._lib.tob200_solver_reset(self._h, _p(x0)), "tob200_solver_reset")

    def _wrap(self, ptr: int, shape, dtype):
        """Zero-copy torch view of solver-owned device memory."""
        typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4"}[dtype]

        class _Holder:
            __cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Holder(), device=self.ctx.device)

    @property
    def x(self) -> torch.Tensor:
        """View of the solver's current x [B, n] (device memory owned by the solver)."""
        return self._wrap(self.ctx._lib.tob200_solver_x(self._h), (self.B, self.n), self.dtype)

    @property
    def needs(self) -> torch.Tensor:
        """[B] int32: 1 rebuild (J and r), 0 cost only (r), -1 finished."""
        return self._wrap(self.ctx._lib.tob200_solver_needs(self._h), (self.B,), torch.int32)

    def step(self, J: torch.Tensor, r: torch.Tensor, layout: int | None = None, cost: torch.Tensor | None = None):
        """One Step from the residual blocks at the current x.  `cost` ([B] float64): the pass's Cost::cost as the
        caller's accumulation functor returns it (tob200_solver_step_cost_*; e.g. the residual NORM of
        diff/num_diff.h:300-305) instead of r^T r - general family only."""
        layout = _layout_of(J, layout)
        m = r.shape[1]
        if cost is not None:
            cost = cost.to(dtype=torch.float64, device=self.ctx.device).contiguous()
            fn = getattr(self.ctx._lib, f"tob200_solver_step_cost_{_suf(self.dtype)}")
            self.ctx._ck(fn(self._h, _p(J.contiguous()), _p(r.contiguous()), layout, m, _p(cost)), "tob200_solver_step_cost")
            self._keep = (cost,)
            return
        fn = getattr(self.ctx._lib, f"tob200_solver_step_{_suf(self.dtype)}")
        self.ctx._ck(fn(self._h, _p(J.contiguous()), _p(r.contiguous()), layout, m), "tob200_solver_step")

    def step_hg(self, grad: torch.Tensor, H: torch.Tensor, cost: torch.Tensor, num_residuals: torch.Tensor | None = None):
        """One Step from USER-FILLED accumulators (tob200_solver_step_hg_*): what the reference's
        `acc(x, grad, H) -> Cost` lambda leaves in grad_ / H_ and returns (docs/API.md:37-57,137-170).
        grad [B,n], H [B,n,n] (only the upper triangle is read), cost [B] (float64), num_residuals [B] int32
        (default 1: a scalar cost, cost.h:22)."""
        dev = self.ctx.device
        grad = grad.to(dtype=self.dtype, device=dev).contiguous()
        H = H.to(dtype=self.dtype, device=dev).contiguous()
        cost = cost.to(dtype=torch.float64, device=dev).contiguous()
        if num_residuals is None:
            num_residuals = torch.ones((self.B,), dtype=torch.int32, device=dev)
        nres = num_residuals.to(dtype=torch.int32, device=dev).contiguous()
        fn = getattr(self.ctx._lib, f"tob200_solver_step_hg_{_suf(self.dtype)}")
        self.ctx._ck(fn(self._h, _p(grad), _p(H), _p(cost), _p(nres)), "tob200_solver_step_hg")
        self._keep = (grad, H, cost, nres)  # the launch is asynchronous: keep the converted copies alive

    def step_hg_sparse(self, grad: torch.Tensor, rows, cols, values: torch.Tensor, cost: torch.Tensor,
                       num_residuals: torch.Tensor | None = None):
        """One Step from a SPARSE user-filled H (tob200_solver_step_hg_sparse_*; the reference's
        `acc(x, grad, SparseMatrix &H)` signature, tests/sparse.cpp): rows / cols [nnz] (host, one pattern for the batch),
        values [B, nnz]; duplicates are summed in triplet order, entries below the diagonal ignored."""
        dev = self.ctx.device
        rows = np.ascontiguousarray(np.asarray(rows, dtype=np.int32))
        cols = np.ascontiguousarray(np.asarray(cols, dtype=np.int32))
        nnz = int(rows.shape[0])
        grad = grad.to(dtype=self.dtype, device=dev).contiguous()
        values = values.to(dtype=self.dtype, device=dev).contiguous()
        cost = cost.to(dtype=torch.float64, device=dev).contiguous()
        if num_residuals is None:
            num_residuals = torch.ones((self.B,), dtype=torch.int32, device=dev)
        nres = num_residuals.to(dtype=torch.int32, device=dev).contiguous()
        fn = getattr(self.ctx._lib, f"tob200_solver_step_hg_sparse_{_suf(self.dtype)}")
        self.ctx._ck(fn(self._h, _p(grad), rows.ctypes.data_as(C.c_void_p), cols.ctypes.data_as(C.c_void_p), nnz, _p(values),
                        _p(cost), _p(nres)), "tob200_solver_step_hg_sparse")
        self._keep = (grad, values, cost, nres)

    def num_active(self) -> int:
        v = C.c_int64(0)
        self.ctx._ck(self.ctx._lib.tob200_solver_num_active(self._h, C.byref(v)), "tob200_solver_num_active")
        return int(v.value)

    def results(self) -> np.ndarray:
        buf = torch.empty((self.B, C.sizeof(_lib.Result)), dtype=torch.uint8, device=self.ctx.device)
        self.ctx._ck(self.ctx._lib.tob200_solver_results(self._h, _p(buf)), "tob200_solver_results")
        self.ctx.sync()
        return decode_results(buf)

    def final_hessian(self) -> torch.Tensor:
        H = torch.empty((self.B, self.n, self.n), dtype=torch.float64, device=self.ctx.device)
        self.ctx._ck(self.ctx._lib.tob200_solver_final_hessian(self._h, _p(H)), "tob200_solver_final_hessian")
        return H

    def covariance(self, rescaled: bool = False):
        """Output::Covariance(rescaled) (output.h:81-103): InvCov of the un-damped final Hessian in
        double; `rescaled` multiplies by final_cost^2 / (num_residuals - n) when num_residuals > n.
        Returns (cov [B,n,n] float64, status [B]: 1 where the reference returns nullopt)."""
        cov = torch.zeros((self.B, self.n, self.n), dtype=torch.float64, device=self.ctx.device)
        status = torch.zeros((self.B,), dtype=torch.int32, device=self.ctx.device)
        self.ctx._ck(self.ctx._lib.tob200_solver_covariance(self._h, _p(cov), None, _p(status)),
                     "tob200_solver_covariance")
        if rescaled:
            res = self.results()
            fc = torch.from_numpy(res["final_cost"].astype("float64")).to(cov.device)
            nr = torch.from_numpy(res["final_num_residuals"].astype("float64")).to(cov.device)
            f = torch.where(nr > self.n, fc * fc / (nr - self.n).clamp(min=1.0), torch.ones_like(fc))
            cov = cov * f[:, None, None]
        return cov, status

    def max_std_dev(self) -> torch.Tensor:
        """SolverLM::MaxStdDev(use_damped=false) (solvers/lm.h:176-187) per problem; 0 where InvCov fails."""
        ms = torch.zeros((self.B,), dtype=torch.float64, device=self.ctx.device)
        status = torch.zeros((self.B,), dtype=torch.int32, device=self.ctx.device)
        self.ctx._ck(self.ctx._lib.tob200_solver_covariance(self._h, None, _p(ms), _p(status)),
                     "tob200_solver_covariance")
        return ms
