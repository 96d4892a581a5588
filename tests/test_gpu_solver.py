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


# ---- the manual accumulation contract `acc(x, grad, H) -> Cost` with USER-FILLED accumulators ----
def drive_hg(ctx, x0, acc_fn, opt, dtype, max_steps=400, poison_lower=True):
    """Host-driven loop over tob200_solver_step_hg: `acc_fn(x_p, want_grad) -> (grad, H, cost, nres)` is the
    caller's lambda (numpy, one problem at a time: it is the SAME function the oracle calls back, so both see
    identical accumulators).  The strict lower triangle handed to the device is poisoned with NaN: only the
    upper one may be read (docs/API.md:170)."""
    import tinyopt_b200 as tb
    B, n = x0.shape
    npdt = np.float64 if dtype == torch.float64 else np.float32
    s = tb.BatchSolver(ctx, B, n, dtype, opt)
    s.reset(torch.from_numpy(x0).cuda())
    steps = 0
    while s.num_active() > 0 and steps < max_steps:
        x = s.x.cpu().numpy()
        needs = s.needs.cpu().numpy()
        g = np.zeros((B, n), npdt); H = np.zeros((B, n, n), npdt); c = np.zeros(B); nr = np.ones(B, np.int32)
        for p in range(B):
            if needs[p] < 0:
                continue
            gp, Hp, cp, nrp = acc_fn(p, x[p], needs[p] == 1)
            c[p], nr[p] = cp, nrp
            if needs[p] == 1:
                g[p], H[p] = gp, Hp
                if poison_lower and opt.use_ldlt:
                    H[p][np.tril_indices(n, -1)] = np.nan
        s.step_hg(torch.from_numpy(g), torch.from_numpy(H), torch.from_numpy(c), torch.from_numpy(nr))
        steps += 1
    res = s.results()
    x = s.x.cpu().numpy()
    Hf = s.final_hessian().cpu().numpy()
    s.close()
    return x, res, Hf


def oracle_hg(x0, acc_fn, kw, npdt):
    outs = []
    for p in range(x0.shape[0]):
        def acc(xv, g, H, p=p):
            gp, Hp, cp, nrp = acc_fn(p, xv.astype(npdt), g is not None)
            if g is not None:
                g[:] = gp
                H[:, :] = Hp
            return cp, nrp
        outs.append(O.optimize(x0[p], acc, O.default_options(**kw), dtype=npdt))
    return outs


def assert_hg_parity(x, res, outs):
    for p, o in enumerate(outs):
        assert res["num_iters"][p] == o.num_iters and res["stop_reason"][p] == o.stop_reason, (p, res["num_iters"][p], o.num_iters)
        assert res["num_failures"][p] == o.num_failures
        assert np.array_equal(x[p], o.x), (p, x[p], o.x)
        assert res["final_cost"][p] == o.final_cost and res["last_lambda"][p] == o.last_lambda


def test_hg_rosenbrock_true_hessian(ctx):
    """tests/optimize_easy.cpp:35-80: Rosenbrock through a lambda that fills the TRUE Hessian (not J^T J) and
    the gradient, options max_iters = 200, min_rerr_dec = 0, max_consec_failures = 20; the reference asserts
    Succeeded, Converged and x = (1, 1) +- 1e-5.  A batch of perturbed starting points; device == oracle bit for
    bit (same accumulators in, same canonical Build / Solve / Step sequence)."""
    import tinyopt_b200 as tb

    def acc(p, v, want):
        xv, yv = float(v[0]), float(v[1])
        t1 = 1.0 - xv
        t2 = yv - xv * xv
        g = H = None
        if want:
            g = np.array([-2.0 * t1 - 400.0 * xv * t2, 200.0 * t2])
            H = np.array([[2.0 - 400.0 * yv + 1200.0 * xv * xv, -400.0 * xv], [-400.0 * xv, 200.0]])
        return g, H, t1 * t1 + 100.0 * t2 * t2, 1

    rng = np.random.default_rng(3)
    x0 = np.array([-1.2, 1.0]) + 0.05 * rng.standard_normal((40, 2))
    x0[0] = [-1.2, 1.0]
    kw = dict(max_iters=200, min_rerr_dec=0.0, max_consec_failures=20)
    x, res, _ = drive_hg(ctx, x0, acc, tb.options(**kw), torch.float64)
    outs = oracle_hg(x0, acc, kw, np.float64)
    assert_hg_parity(x, res, outs)
    assert outs[0].Succeeded() and outs[0].Converged()
    assert abs(x[0, 0] - 1.0) < 1e-5 and abs(x[0, 1] - 1.0) < 1e-5           # the reference's own assertion
    ok = res["stop_reason"] > 0     # (a few perturbed starts end in kSolverFailed - in the oracle and on the device alike)
    assert ok.mean() > 0.8 and np.abs(x[ok] - 1.0).max() < 1e-4
    assert (res["num_failures"] > 0).any()    # the indefinite true Hessian makes LM reject / re-damp on the way


@pytest.mark.parametrize("dtype,n", [(torch.float64, 3), (torch.float64, 6), (torch.float64, 12), (torch.float32, 12),
                                     (torch.float32, 30), (torch.float64, 40)])
def test_hg_prior_diagonal_hessian(ctx, dtype, n):
    """benchmarks/dense.cpp:57-66 "Prior n" (the reference's published benchmark rows): whitened prior
    res = (x - y) / stdevs, grad = J res with J = diag(1 / stdevs), ONLY H.diagonal() is filled (the rest of
    the pre-zeroed H_ stays zero), cost = res.squaredNorm() as a scalar Cost (one "residual")."""
    import tinyopt_b200 as tb
    npdt = np.float64 if dtype == torch.float64 else np.float32
    rng = np.random.default_rng(n)
    B = 37
    y = rng.uniform(-1, 1, (B, n)).astype(npdt)
    sd = rng.uniform(-1, 1, (B, n)).astype(npdt)
    sd[np.abs(sd) < 0.05] = npdt(0.3)

    def acc(p, v, want):
        res = ((v - y[p]) / sd[p]).astype(npdt)
        cost = npdt(0)
        for r in res:
            cost = npdt(cost + npdt(r * r))
        g = H = None
        if want:
            g = (res / sd[p]).astype(npdt)
            H = np.diag((npdt(1) / sd[p]) ** 2).astype(npdt)
        return g, H, float(cost), 1

    x0 = rng.uniform(-1, 1, (B, n)).astype(npdt)
    kw = {} if npdt == np.float64 else dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)
    x, res, Hf = drive_hg(ctx, x0, acc, tb.options(**kw), dtype)
    outs = oracle_hg(x0, acc, kw, npdt)
    assert_hg_parity(x, res, outs)
    assert np.abs(x - y).max() < (1e-6 if npdt == np.float64 else 1e-3) and (res["stop_reason"] > 0).all()
    for p in (0, B - 1):   # Output::final_hessian (lm.h:157-171): the un-damped user H
        assert np.allclose(np.diag(Hf[p]), (1.0 / sd[p].astype(np.float64)) ** 2, rtol=1e-6 if npdt == np.float32 else 1e-14)


def test_hg_singular_normal_matrix_float(ctx):
    """tests/types.cpp:97-108: three Vec2f stacked into 6 parameters, res = x0 + x1 + x2 - 10: H = J^T J is 6 x 6
    of rank 2, solved only thanks to the LM damping; the lambda returns the residual VECTOR (Cost = its squared
    norm in float, 2 residuals).  Reference assertion: |x0 + x1 + x2 - 10| < 1e-5."""
    import tinyopt_b200 as tb
    J = np.array([[1, 0, 1, 0, 1, 0], [0, 1, 0, 1, 0, 1]], np.float32)
    JtJ = (J.T @ J).astype(np.float32)

    def acc4(p, v, want):
        r = (v[0:2] + v[2:4] + v[4:6] - np.float32(10)).astype(np.float32)
        cost = np.float32(np.float32(r[0] * r[0]) + np.float32(r[1] * r[1]))
        if want:
            return (J.T @ r).astype(np.float32), JtJ, float(cost), 2
        return None, None, float(cost), 2

    rng = np.random.default_rng(0)
    x0 = rng.uniform(-1, 1, (16, 6)).astype(np.float32)
    x, res, _ = drive_hg(ctx, x0, acc4, tb.options(), torch.float32)
    outs = oracle_hg(x0, acc4, {}, np.float32)
    assert_hg_parity(x, res, outs)
    assert np.abs(x[:, 0:2] + x[:, 2:4] + x[:, 4:6] - 10).max() < 1e-5


# ---- the seam above n = 55: the general kernel family behind tob200_solver_* (gn.cuh) --------------------------------
@pytest.mark.parametrize("dtype,B,m,n", [(np.float64, 9, 150, 56), (np.float32, 7, 200, 72), (np.float64, 5, 260, 130),
                                         (np.float32, 3, 640, 300), (np.float64, 2, 300, 257), (np.float32, 2, 700, 600)])
def test_solver_seam_above_55_matches_oracle(ctx, dtype, B, m, n):
    """`Optimizer_<SolverLM>` with a user lambda has no size cap (math.h:232-240): the host-driven loop on the general
    family, fed with the family's residual blocks, == the oracle bit for bit in both precisions; Output::final_hessian and
    Output::Covariance() (InvCov of the un-damped H_, in double) follow."""
    import tinyopt_b200 as tb
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    kw = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9) if dtype == np.float32 else {}
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=5)
    xo, ro, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw))
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, tdt, p0=5, layout=tb.PROBLEM_MAJOR)
    s = tb.BatchSolver(ctx, B, n, tdt, tb.options(**kw))
    s.reset(dx0)
    assert s.num_active() == B and (s.needs.cpu().numpy() == 1).all()
    steps = 0
    while s.num_active() > 0 and steps < 200:
        r, J = ctx.synth_eval(dA, dy, s.x, layout=tb.PROBLEM_MAJOR)
        s.step(J, r, layout=tb.PROBLEM_MAJOR)
        steps += 1
    res = s.results()
    assert (s.needs.cpu().numpy() == -1).all()
    assert np.array_equal(res["num_iters"], ro["num_iters"])
    assert np.array_equal(res["stop_reason"], ro["stop_reason"])
    assert np.array_equal(s.x.cpu().numpy(), xo)
    assert np.array_equal(res["final_cost"], ro["final_cost"])
    assert np.array_equal(res["last_lambda"], ro["last_lambda"])
    H = s.final_hessian().cpu().numpy()
    assert np.array_equal(H, np.swapaxes(H, 1, 2))
    cov, st = s.covariance()
    ctx.sync()
    assert (st.cpu().numpy() == 0).all()
    assert np.array_equal(cov[0].cpu().numpy(), O.inv_cov(H[0]))
    eye_err = np.abs(np.einsum("bij,bjk->bik", cov.cpu().numpy(), H) - np.eye(n)).max()
    assert eye_err < 1e-8, eye_err
    ms = s.max_std_dev().cpu().numpy()
    assert np.array_equal(ms, np.sqrt(cov.cpu().numpy().reshape(B, -1).max(axis=1)))
    s.close()


@pytest.mark.parametrize("dtype,n", [(torch.float64, 64), (torch.float32, 100), (torch.float64, 300)])
def test_hg_prior_above_55(ctx, dtype, n):
    """"Prior n" (benchmarks/dense.cpp:57-66) with user-filled accumulators above n = 55: only H.diagonal() is filled, the
    lower triangle is poisoned, cost is a scalar Cost; bit for bit against the oracle run with the same numpy lambda."""
    import tinyopt_b200 as tb
    npdt = np.float64 if dtype == torch.float64 else np.float32
    rng = np.random.default_rng(n)
    B = 5
    y = rng.uniform(-1, 1, (B, n)).astype(npdt)
    sd = rng.uniform(-1, 1, (B, n)).astype(npdt)
    sd[np.abs(sd) < 0.05] = npdt(0.3)

    def acc(p, v, want):
        res = ((v - y[p]) / sd[p]).astype(npdt)
        cost = npdt(0)
        for r in res:
            cost = npdt(cost + npdt(r * r))
        g = H = None
        if want:
            g = (res / sd[p]).astype(npdt)
            H = np.diag((npdt(1) / sd[p]) ** 2).astype(npdt)
        return g, H, float(cost), 1

    x0 = rng.uniform(-1, 1, (B, n)).astype(npdt)
    kw = {} if npdt == np.float64 else dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)
    x, res, Hf = drive_hg(ctx, x0, acc, tb.options(**kw), dtype)
    outs = oracle_hg(x0, acc, kw, npdt)
    assert_hg_parity(x, res, outs)
    assert np.abs(x - y).max() < (1e-6 if npdt == np.float64 else 1e-3) and (res["stop_reason"] > 0).all()


def test_solver_seam_above_55_option_variants(ctx):
    """Gauss-Newton, use_ldlt = false, cost normalisation and gradient clipping through the seam at n = 64 (double)."""
    import tinyopt_b200 as tb
    B, m, n = 4, 160, 64
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float64, p0=11)
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, torch.float64, p0=11, layout=tb.PROBLEM_MAJOR)
    for kw in (dict(solver_type=1), dict(use_ldlt=0), dict(normalize=1, downscale_by_2=1), dict(grad_clipping=0.05, max_iters=8),
               dict(damping_init=10.0)):
        xo, ro, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw))
        s = tb.BatchSolver(ctx, B, n, torch.float64, tb.options(**kw))
        s.reset(dx0)
        steps = 0
        while s.num_active() > 0 and steps < 200:
            r, J = ctx.synth_eval(dA, dy, s.x, layout=tb.PROBLEM_MAJOR)
            s.step(J, r, layout=tb.PROBLEM_MAJOR)
            steps += 1
        res = s.results()
        assert np.array_equal(res["num_iters"], ro["num_iters"]), kw
        assert np.array_equal(res["stop_reason"], ro["stop_reason"]), kw
        assert np.array_equal(s.x.cpu().numpy(), xo), kw
        assert np.array_equal(res["final_cost"], ro["final_cost"]), kw
        s.close()


# ---- sparse H in the accumulation signature (tests/sparse.cpp:19-57) -------------------------------------------------
@pytest.mark.parametrize("dtype,n", [(torch.float64, 100), (torch.float32, 10), (torch.float64, 12)])
def test_hg_sparse_reference_case(ctx, dtype, n):
    """tests/sparse.cpp "tinyopt_sparse" (n = 100, double) and "tinyopt_sparse_ad" (n = 10, float): res = 10 x - 2,
    grad = J^T res with J = 10 I, H = J^T J handed over as TRIPLETS (here with a duplicated diagonal entry, 60 + 40, and
    junk below the diagonal that a `SimplicialLDLT<_, Upper>` never reads), Cost(res.norm(), res.size()).  Reference
    assertions: Succeeded, Converged, min / max of x == 0.2 +- 1e-5; and bit for bit against the oracle's dense run of the
    same lambda."""
    import tinyopt_b200 as tb
    npdt = np.float64 if dtype == torch.float64 else np.float32
    rng = np.random.default_rng(7 + n)
    B = 5
    x0 = rng.uniform(-1, 1, (B, n)).astype(npdt)
    rows = np.concatenate([np.arange(n), np.arange(n), np.arange(1, n)]).astype(np.int32)   # diag twice + sub-diagonal junk
    cols = np.concatenate([np.arange(n), np.arange(n), np.arange(0, n - 1)]).astype(np.int32)

    def acc(p, v, want):
        res = (npdt(10) * v - npdt(2)).astype(npdt)
        c = npdt(0)
        for r in res:
            c = npdt(c + npdt(r * r))
        cost = float(np.sqrt(c))
        if want:
            return (npdt(10) * res).astype(npdt), np.diag(np.full(n, 100, npdt)), cost, n
        return None, None, cost, n

    kw = dict(check_final_cost=0)
    if npdt == np.float32:
        kw.update(min_rerr_dec=1e-5, min_step_norm2=1e-9)
    s = tb.BatchSolver(ctx, B, n, dtype, tb.options(**kw))
    s.reset(torch.from_numpy(x0).cuda())
    steps = 0
    while s.num_active() > 0 and steps < 100:
        x = s.x.cpu().numpy()
        needs = s.needs.cpu().numpy()
        g = np.zeros((B, n), npdt); c = np.zeros(B); nr = np.full(B, n, np.int32)
        vals = np.zeros((B, 3 * n - 1), npdt)
        for p in range(B):
            if needs[p] < 0:
                continue
            gp, Hp, c[p], nr[p] = acc(p, x[p], needs[p] == 1)
            if needs[p] == 1:
                g[p] = gp
                vals[p, :n] = 60; vals[p, n:2 * n] = 40; vals[p, 2 * n:] = np.nan
        s.step_hg_sparse(torch.from_numpy(g), rows, cols, torch.from_numpy(vals), torch.from_numpy(c), torch.from_numpy(nr))
        steps += 1
    res = s.results()
    x = s.x.cpu().numpy()
    s.close()
    assert (res["stop_reason"] > 0).all()
    assert abs(x.min() - 0.2) < 1e-5 and abs(x.max() - 0.2) < 1e-5
    assert_hg_parity(x, res, oracle_hg(x0, acc, kw, npdt))


# ---- tob200_solver_step_cost: the Cost is what the caller's accumulation functor returns (diff/num_diff.h:284-309) ----
def test_step_cost_numdiff_host_lambda(ctx):
    """`CreateNumDiffFunc2(x, residuals)` handed to the optimizer, with the residuals as a HOST lambda (numpy, float): an
    exponential fit r_i = a exp(b t_i) + c - y_i, J from central differences with h = FloatEpsilon<float> = 1e-4f
    (diff/num_diff.h:57-126), grad = J^T r, H = J^T J, Cost(res.norm(), res.size()) (:305).  The device gets J, r and that
    cost (tob200_solver_step_cost_f32, general family); the oracle runs the same lambda, its g / H formed with the
    canonical fma chains (float fma emulated exactly in double).  Bit for bit."""
    import tinyopt_b200 as tb
    f32 = np.float32
    B, n, m = 6, 3, 24
    rng = np.random.default_rng(5)
    t = np.linspace(0, 1, m).astype(f32)
    truth = np.stack([rng.uniform(0.5, 2.0, B), rng.uniform(-1.5, 1.0, B), rng.uniform(-0.5, 0.5, B)], 1).astype(f32)
    yv = (truth[:, 0:1] * np.exp(truth[:, 1:2] * t) + truth[:, 2:3] + 0.01 * rng.standard_normal((B, m))).astype(f32)
    x0 = (truth + 0.2 * rng.uniform(-1, 1, (B, n))).astype(f32)
    h = f32(1e-4)

    def res(p, x):
        return (f32(x[0]) * np.exp(f32(x[1]) * t).astype(f32) + f32(x[2]) - yv[p]).astype(f32)

    def numeval(p, x):   # NumEval, Method::kCentral
        x = x.astype(f32)
        r = res(p, x)
        J = np.zeros((m, n), f32)
        for k in range(n):
            yp = x.copy(); yp[k] = f32(x[k] + h)
            ym = x.copy(); ym[k] = f32(x[k] + f32(-h))
            J[:, k] = ((res(p, yp) - res(p, ym)).astype(f32) / f32(f32(2) * h)).astype(f32)
        return r, J

    def fma32(a, b, c):
        return f32(np.float64(a) * np.float64(b) + np.float64(c))

    def norm32(r):
        c = f32(0)
        for v in r:
            c = fma32(v, v, c)
        return float(np.sqrt(c, dtype=f32))

    kw = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)
    s = tb.BatchSolver(ctx, B, n, torch.float32, tb.options(**kw), general=True)
    s.reset(torch.from_numpy(x0).cuda())
    steps = 0
    while s.num_active() > 0 and steps < 100:
        x = s.x.cpu().numpy()
        r = np.zeros((B, m), f32); J = np.zeros((B, m, n), f32); c = np.zeros(B)
        for p in range(B):
            r[p], J[p] = numeval(p, x[p])
            c[p] = norm32(r[p])
        s.step(torch.from_numpy(J).cuda(), torch.from_numpy(r).cuda(), layout=tb.PROBLEM_MAJOR, cost=torch.from_numpy(c))
        steps += 1
    resu = s.results()
    xg = s.x.cpu().numpy()
    s.close()
    # a small-n solver without the flag refuses the cost override (the fused step kernels form r^T r themselves)
    s2 = tb.BatchSolver(ctx, B, n, torch.float32, tb.options(**kw))
    s2.reset(torch.from_numpy(x0).cuda())
    with pytest.raises(RuntimeError):
        s2.step(torch.from_numpy(J).cuda(), torch.from_numpy(r).cuda(), layout=tb.PROBLEM_MAJOR, cost=torch.from_numpy(c))
    s2.close()
    for p in range(B):
        def acc(xv, g, H, p=p):
            r_, J_ = numeval(p, xv)
            if g is not None:
                for j in range(n):
                    gj = f32(0)
                    for i in range(m):
                        gj = fma32(J_[i, j], r_[i], gj)
                    g[j] = gj
                    for k in range(j, n):
                        hjk = f32(0)
                        for i in range(m):
                            hjk = fma32(J_[i, j], J_[i, k], hjk)
                        H[j, k] = hjk
                        H[k, j] = hjk
            return norm32(r_), m
        o = O.optimize(x0[p], acc, O.default_options(**kw), dtype=f32)
        assert resu["num_iters"][p] == o.num_iters and resu["stop_reason"][p] == o.stop_reason, (p, resu["num_iters"][p], o.num_iters)
        assert np.array_equal(xg[p], o.x), (p, xg[p], o.x)
        assert resu["final_cost"][p] == o.final_cost and resu["final_num_residuals"][p] == m
        assert o.stop_reason > 0 and np.abs(o.x - truth[p]).max() < 0.3
