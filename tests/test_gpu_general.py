"""GPU parity tests of the GENERAL kernel family (gn.cuh, family 4): double precision above n = 55 (tinyopt's
default scalar), any precision above n = 512 (the reference's dynamic-size solver has no size cap, math.h:232-240)
and `hessian.use_ldlt = false` above n = 55 (solvers/gn.h:157-163).  The family keeps the oracle's canonical
operation sequence, so everything is compared BIT FOR BIT — solutions, costs, lambda, iteration counts, stop
reasons — which is stronger than the north star's 1e-10 (double) / 1e-4 (float)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

FLOAT_OPTS = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)
TDT = {np.float32: torch.float32, np.float64: torch.float64}


@pytest.fixture(scope="module")
def ctx():
    import tinyopt_b200 as tb
    c = tb.Context(0)
    yield c
    c.close()


def run_both(ctx, dtype, B, m, n, p0=0, want_hessian=False, **optkw):
    import tinyopt_b200 as tb
    kw = {**(FLOAT_OPTS if dtype == np.float32 else {}), **optkw}
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=p0)
    ref = O.synth_lm_run(A, y, x0, O.default_options(**kw), fast=True, want_hessian=want_hessian)
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, TDT[dtype], p0=p0, layout=tb.PROBLEM_MAJOR)
    out = ctx.optimize_batch(dA, dy, dx0, tb.options(**kw), layout=tb.PROBLEM_MAJOR, want_hessian=want_hessian)
    return ref, out


def assert_exact(ref, out):
    xo, ro = ref[0], ref[1]
    rg = out.results
    assert np.array_equal(rg["num_iters"], ro["num_iters"])
    assert np.array_equal(rg["stop_reason"], ro["stop_reason"])
    assert np.array_equal(rg["num_failures"], ro["num_failures"])
    assert np.array_equal(out.x.cpu().numpy(), xo)
    assert np.array_equal(rg["final_cost"], ro["final_cost"])
    assert np.array_equal(rg["last_lambda"], ro["last_lambda"])
    assert np.array_equal(rg["final_rerr_dec"], ro["final_rerr_dec"])


@pytest.mark.parametrize("n,B", [(56, 6), (57, 4), (64, 4), (100, 3), (200, 3), (333, 2), (512, 2)])
def test_double_above_55(ctx, n, B):
    """tinyopt deduces the scalar from x and defaults to double (optimize.h:20-33): 56 <= n <= 512 in double,
    default options, Output::final_hessian included."""
    assert ctx.kernel_family(torch.float64, n) == 4
    ref, out = run_both(ctx, np.float64, B, 2 * n + 7, n, p0=n, want_hessian=True)
    assert_exact(ref, out)
    assert np.array_equal(out.final_hessian.cpu().numpy(), ref[3])
    assert (ref[1]["stop_reason"] > 0).all()


@pytest.mark.parametrize("dtype,n,m,B", [(np.float32, 513, 600, 3), (np.float32, 777, 900, 2), (np.float64, 520, 640, 2),
                                         (np.float32, 1024, 1100, 2), (np.float64, 1030, 1100, 1), (np.float32, 2048, 2100, 1)])
def test_above_512(ctx, dtype, n, m, B):
    """SolveLDLT on dynamic matrices has no size cap in the reference (math.h:232-240; the only limit is bad_alloc,
    optimizers/optimizer.h:75-86): n up to 2048 in both precisions."""
    assert ctx.kernel_family(TDT[dtype], n) == 4
    ref, out = run_both(ctx, dtype, B, m, n, p0=7)
    assert_exact(ref, out)


def test_size_cap_is_reported(ctx):
    import tinyopt_b200 as tb
    assert ctx.kernel_family(torch.float32, 2049) == 0 and ctx.kernel_family(torch.float64, 2049) == 0
    dA, dy, _, dx0 = ctx.synth_generate(1, 4, 2049, torch.float32, layout=tb.PROBLEM_MAJOR)
    with pytest.raises(tb.api.TinyoptB200Error, match="largest supported size"):
        ctx.optimize_batch(dA, dy, dx0, tb.options(), layout=tb.PROBLEM_MAJOR)


@pytest.mark.parametrize("dtype,n,B", [(np.float32, 56, 5), (np.float32, 100, 3), (np.float32, 200, 2), (np.float64, 60, 4),
                                       (np.float64, 130, 2), (np.float32, 600, 1)])
def test_inverse_path_above_55(ctx, dtype, n, B):
    """hessian.use_ldlt = false: dx = -H.inverse() * grad (solvers/gn.h:157-163; Eigen's PartialPivLU has no size
    cap) — the oracle's partial-pivot LU in its canonical order, bit for bit."""
    ref, out = run_both(ctx, dtype, B, 2 * n + 3, n, p0=3, use_ldlt=0)
    assert_exact(ref, out)


@pytest.mark.parametrize("optkw", [dict(solver_type=1), dict(max_iters=2), dict(damping_init=10.0, max_consec_failures=2),
                                   dict(check_min_H_diag=1e4), dict(grad_clipping=0.05, max_iters=6),
                                   dict(normalize=1, downscale_by_2=1), dict(use_squared_norm=0),
                                   dict(min_rerr_dec=0.0, min_step_norm2=0.0, min_error=0.0, min_grad_norm2=0.0, max_iters=25)])
def test_option_variants_double_n60(ctx, optkw):
    """Gauss-Newton, the iteration cap, heavy damping with early give-up, the diagonal check (kSolverFailed),
    gradient clipping, the cost conventions, and a run into the noise floor (rejected steps, roll-backs, cost-only
    passes with the stale re-damped H_)."""
    ref, out = run_both(ctx, np.float64, 8, 150, 60, p0=1, **optkw)
    assert_exact(ref, out)
    if "min_step_norm2" in optkw:   # every stop test off: the run ends in kMaxConsecNoDecr at the noise floor
        assert (out.results["num_builds"] < out.results["num_iters"]).all()   # cost-only passes happened
        assert (out.results["stop_reason"] == 7).all() and (out.results["num_failures"] >= 5).all()


@pytest.mark.parametrize("dtype,B,m,n", [(np.float64, 4, 170, 80), (np.float32, 2, 650, 600), (np.float64, 3, 20, 60)])
def test_build_solve_general(ctx, dtype, B, m, n):
    """One Build + Solve from materialised J, r (a1 + a5 + a6); the last shape has m < n: J^T J is rank deficient,
    zero / negative pivots and Eigen's accept / reject rules decide the status — identically on both sides."""
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=5)
    r, J = O.synth_eval(A, y, x0)
    lam = np.full(B, 1e-4, dtype)
    lam[0] = 0
    out = ctx.build_solve(torch.from_numpy(J).cuda(), torch.from_numpy(r).cuda(), torch.from_numpy(lam).cuda(), want_H=True, want_g=True)
    ctx.sync()
    for p in range(B):
        o = O.build_solve(J[p], r[p], float(lam[p]))
        assert out["status"][p].item() == o["status"], (p, out["status"][p].item(), o["status"])
        assert out["cost"][p].item() == o["cost"]
        assert np.array_equal(out["g"][p].cpu().numpy(), o["g"])
        assert np.array_equal(out["H"][p].cpu().numpy(), o["H"])
        if o["status"] == 0:
            assert np.array_equal(out["dx"][p].cpu().numpy(), o["dx"], equal_nan=True)
