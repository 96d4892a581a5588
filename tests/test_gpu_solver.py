"""GPU tests of the host-driven seam (tob200_solver_*): the caller evaluates its own residual
lambda at x, the device runs Build / Solve / Step / the x update.  Restates the reference tests
that use arbitrary user lambdas (sqrt2.cpp, basic.cpp) and compares with the CPU oracle.
"""
import math

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import tinyopt_b200 as tb
    c = tb.Context(0)
    yield c
    c.close()


def drive(ctx, x0, residual_fn, opt, dtype=torch.float64, max_steps=200):
    """OptimizeBatch with a host/torch residual lambda: r, J = residual_fn(x) -> [B,m], [B,m,n]."""
    import tinyopt_b200 as tb
    B, n = x0.shape
    s = tb.BatchSolver(ctx, B, n, dtype, opt)
    s.reset(x0)
    steps = 0
    while s.num_active() > 0 and steps < max_steps:
        r, J = residual_fn(s.x.clone())
        s.step(J, r)
        steps += 1
    res = s.results()
    x = s.x.clone()
    H = s.final_hessian().cpu().numpy()
    s.close()
    return x.cpu().numpy(), res, H


def test_c1_sqrt2_double(ctx):
    """config C1: sqrt2 scalar LM (x*x - 2), double, default options (README.md:77-95).
    Golden: 5 Steps, kMinError, x -> 1.4142135623730951."""
    import tinyopt_b200 as tb
    x0 = torch.tensor([[1.0], [float(np.float32(-0.3))], [float(np.float32(3.2))]], dtype=torch.float64, device="cuda")

    def f(x):
        return (x * x - 2.0), (2.0 * x).unsqueeze(-1)

    x, res, H = drive(ctx, x0[:1], f, tb.options())
    assert res["stop_reason"][0] == tb.StopReason.kMinError and res["num_iters"][0] == 5
    assert x[0, 0] == 1.4142135623730951
    # tests/sqrt2.cpp:22-28 options, its three starting points
    kw = dict(max_iters=20, max_consec_failures=0)
    x, res, H = drive(ctx, x0, f, tb.options(**kw))
    for b in range(3):
        o = O.optimize(float(x0[b, 0]), lambda xv, g, Hm: _sqrt2_acc(xv, g, Hm), O.default_options(**kw))
        assert res["num_iters"][b] == o.num_iters and res["stop_reason"][b] == o.stop_reason
        assert x[b, 0] == o.x[0]
        assert res["final_cost"][b] == o.final_cost
        assert abs(abs(x[b, 0]) - math.sqrt(2)) < 1e-5  # tests/sqrt2.cpp
        assert H[b, 0, 0] == o.final_hessian[0, 0]


def _sqrt2_acc(x, g, H):
    r = x[0] * x[0] - 2.0
    J = 2 * x[0]
    if g is not None:
        g[0] = J * r
        H[0, 0] = J * J
    return r * r


def test_bad_step_path(ctx):
    """r = atan x from x0 = 3: 6 Steps, kMaxConsecNoDecr, x rolled back to 3.0,
    3 rebuilds then 3 cost-only passes (tests/solvers.cpp:74-109 skip-rebuild contract)."""
    import tinyopt_b200 as tb
    x0 = torch.tensor([[3.0]], dtype=torch.float64, device="cuda")

    def f(x):
        return torch.atan(x), (1.0 / (1.0 + x * x)).unsqueeze(-1)

    x, res, _ = drive(ctx, x0, f, tb.options())
    assert res["stop_reason"][0] == tb.StopReason.kMaxConsecNoDecr
    assert res["num_iters"][0] == 6 and res["num_builds"][0] == 3
    assert x[0, 0] == 3.0
    assert res["last_lambda"][0] == pytest.approx(float(np.float32(1e-4)) * 32768, rel=1e-12)


def test_nan_and_failures(ctx):
    """tests/basic.cpp:158-232 through the residual-block interface: NaN in J, Inf residual, and a
    healthy problem side by side; every outcome must equal the oracle's (for n == 1 Eigen's
    size<=1 special case accepts a NaN pivot and the pseudo-inverse returns dx = 0)."""
    import tinyopt_b200 as tb
    x0 = torch.ones((3, 1), dtype=torch.float64, device="cuda")

    def f(x):
        r = x - 2.0
        J = torch.ones((3, 1, 1), dtype=torch.float64, device="cuda")
        J[0] = float("nan")
        r = r.clone()
        r[1] = float("inf")
        return r, J

    x, res, _ = drive(ctx, x0, f, tb.options())

    def make_acc(kind):
        def acc(xv, g, H):
            r = xv[0] - 2.0
            J = float("nan") if kind == 0 else 1.0
            if kind == 1:
                r = float("inf")
            if g is not None:
                g[0] = J * r
                H[0, 0] = J * J
            return r * r
        return acc

    for b in range(3):
        o = O.optimize(1.0, make_acc(b))
        assert res["stop_reason"][b] == o.stop_reason, b
        assert res["num_iters"][b] == o.num_iters, b
        assert x[b, 0] == o.x[0], b
    assert res["stop_reason"][1] == tb.StopReason.kSystemHasNaNOrInf and x[1, 0] == 1.0
    assert res["stop_reason"][2] > 0


def test_solver_failure_retries(ctx):
    """Rank-1 Jacobians under Gauss-Newton (no damping): rounding makes the second pivot land on
    either side of zero, so some problems are rejected by `isPositive()`, retried
    (optimizer.h:356-393) and end in kSolverFailed.  Same op sequence -> same outcome per problem."""
    import tinyopt_b200 as tb
    rng = np.random.default_rng(5)
    B = 256
    u = rng.uniform(-1, 1, (B, 3)); v = rng.uniform(-1, 1, (B, 2))
    Jn = u[:, :, None] * v[:, None, :]
    Jd = torch.from_numpy(Jn).cuda()
    x0 = torch.ones((B, 2), dtype=torch.float64, device="cuda")

    def f(x):
        return torch.einsum("bij,bj->bi", Jd, x) - 1.0, Jd

    opt_kw = dict(solver_type=1)
    x, res, _ = drive(ctx, x0, f, tb.options(**opt_kw))
    n_failed = 0
    for b in range(B):
        # status depends on H = J^T J only; the oracle's accumulate + LDLT decide it
        r0 = (Jd[b] @ x0[b] - 1.0).cpu().numpy()
        ob = O.build_solve(Jn[b], r0, 0.0)
        if ob["status"] == 1:
            n_failed += 1
            assert res["stop_reason"][b] == tb.StopReason.kSolverFailed, b
            assert res["num_iters"][b] == 1 and res["num_failures"][b] == 5
            assert (x[b] == 1.0).all()
        else:
            assert res["stop_reason"][b] != tb.StopReason.kSolverFailed or res["num_iters"][b] > 1, b
    assert n_failed > 0


def _prior_acc(x, g, H):
    r = x[0] - 2.0
    if g is not None:
        g[0] = r
        H[0, 0] = 1.0
    return r * r


@pytest.mark.parametrize("dtype,B,m,n", [(np.float64, 300, 30, 6), (np.float32, 300, 30, 6),
                                         (np.float32, 64, 90, 20), (np.float64, 64, 90, 20),
                                         (np.float32, 40, 120, 50), (np.float64, 24, 100, 40),
                                         (np.float64, 33, 40, 9), (np.float32, 17, 131, 55)])
def test_solver_matches_fused_run(ctx, dtype, B, m, n):
    """The host-driven loop (the SolverType seam, thread- and warp-per-problem families) fed with the
    family's residual blocks must reproduce tob200_lm_run and the oracle exactly (same op sequence once J
    is materialised: sc * a_j is rounded before use in both)."""
    import tinyopt_b200 as tb
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    layout = tb.TILE32 if ctx.kernel_family(tdt, n) == 1 else tb.PROBLEM_MAJOR
    kw = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9) if dtype == np.float32 else {}
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=3)
    xo, ro, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw))
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, tdt, p0=3, layout=layout)
    s = tb.BatchSolver(ctx, B, n, tdt, tb.options(**kw))
    s.reset(dx0)
    steps = 0
    while s.num_active() > 0 and steps < 200:
        r, J = ctx.synth_eval(dA, dy, s.x, layout=layout)
        s.step(J, r, layout=layout)
        steps += 1
    res = s.results()
    assert np.array_equal(res["num_iters"], ro["num_iters"])
    assert np.array_equal(res["stop_reason"], ro["stop_reason"])
    assert np.array_equal(s.x.cpu().numpy(), xo)
    assert np.array_equal(res["final_cost"], ro["final_cost"])
    # Output::final_hessian (un-damped J^T J of the last rebuild) and Output::Covariance() = its inverse
    H = s.final_hessian().cpu().numpy()
    cov, st = s.covariance()
    ctx.sync()
    assert (st.cpu().numpy() == 0).all()
    assert np.array_equal(H, np.swapaxes(H, 1, 2))
    eye_err = np.abs(np.einsum("bij,bjk->bik", cov.cpu().numpy(), H) - np.eye(n)).max()
    assert eye_err < (1e-9 if dtype == np.float64 else 1e-4), eye_err
    for b in range(min(B, 3)):
        assert np.array_equal(cov[b].cpu().numpy(), O.inv_cov(H[b]))
    s.close()


def test_final_hessian_covariance(ctx):
    """tests/cov.cpp:66-91: whitened prior L^T (x - y); the un-damped final Hessian's inverse is the
    prior covariance (+-1e-5) -> pins Hessian() un-damping (solvers/lm.h:157-171)."""
    import tinyopt_b200 as tb
    Cy = np.array([[10.0, 2.0], [2.0, 4.0]])
    Lt = np.linalg.cholesky(np.linalg.inv(Cy)).T
    y = 2 * np.array([0.25, -0.6])
    Ltd = torch.from_numpy(Lt).cuda(); yd = torch.from_numpy(y).cuda()

    def f(x):
        r = (x - yd) @ Ltd.T
        return r, Ltd.unsqueeze(0).expand(x.shape[0], 2, 2).contiguous()

    x, res, H = drive(ctx, torch.zeros((2, 2), dtype=torch.float64, device="cuda"), f, tb.options())
    assert (res["stop_reason"] >= 1).all() and (res["stop_reason"] < 5).all()
    assert np.abs(np.linalg.inv(H[0]) - Cy).max() < 1e-5
    assert np.abs(x[0] - y).max() < 1e-6


def _random_solver_cases(seed, count):
    rng = np.random.default_rng(seed)
    return [(np.float64 if rng.random() < 0.5 else np.float32, int(rng.integers(1, 50)), 0, int(rng.integers(1, 56)))
            for _ in range(count)]


@pytest.mark.parametrize("dtype,B,_m,n", _random_solver_cases(4242, 16))
def test_random_solver_matches_oracle(ctx, dtype, B, _m, n):
    """The SolverType seam over ragged shapes (both families, both precisions, m around n): host-driven
    loop == oracle bit for bit."""
    import tinyopt_b200 as tb
    rng = np.random.default_rng(n * 131 + B)
    m = int(rng.integers(max(1, n // 2), 3 * n + 8))
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    layout = tb.TILE32 if ctx.kernel_family(tdt, n) == 1 else tb.PROBLEM_MAJOR
    kw = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9) if dtype == np.float32 else {}
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=17)
    xo, ro, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw))
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, tdt, p0=17, layout=layout)
    s = tb.BatchSolver(ctx, B, n, tdt, tb.options(**kw))
    s.reset(dx0)
    steps = 0
    while s.num_active() > 0 and steps < 200:
        r, J = ctx.synth_eval(dA, dy, s.x, layout=layout)
        s.step(J, r, layout=layout)
        steps += 1
    res = s.results()
    assert np.array_equal(res["num_iters"], ro["num_iters"])
    assert np.array_equal(res["stop_reason"], ro["stop_reason"])
    assert np.array_equal(s.x.cpu().numpy(), xo)
    assert np.array_equal(res["final_cost"], ro["final_cost"])
    s.close()
