"""Pin the CPU oracle against every known answer the reference's own tests hold for the hot path
(SURVEY.md §8c).  Each case restates one reference test (file:line in the docstring) through the
oracle's callback API and asserts the reference's own tolerance.  AD-path tests are restated with
analytic Jacobians (forward-mode AD is exact up to rounding), i.e. the `acc` that
diff/optimize_autodiff.h:91-166 builds: grad = J^T r, H = J^T J, cost = |r|^2.
"""
import math

import numpy as np
import pytest

from oracle import oracle as O

K = O.K


def nlls_acc(residuals_and_jac):
    """AD adaptor restated: r, J -> grad, H, Cost(|r|^2, m)."""
    def acc(x, g, H):
        r, J = residuals_and_jac(x, g is not None)
        r = np.atleast_1d(np.asarray(r, dtype=x.dtype))
        if g is not None:
            J = np.asarray(J, dtype=x.dtype).reshape(r.size, x.size)
            g[:] = J.T @ r
            H[:, :] = J.T @ J
        if r.size == 1:  # scalar residual: `return res.a * res.a` -> Cost(scalar), n = 1
            return float(r[0] * r[0])
        return r
    return acc


# ---- tests/sqrt2.cpp:22-113 ----------------------------------------------------------------------
def sqrt2_options(**kw):
    return O.default_options(max_iters=20, max_consec_failures=0, **kw)  # sqrt2.cpp:22-28


@pytest.mark.parametrize("x0", [1.0, -0.3, 3.2])
def test_sqrt2_manual_float(x0):
    """tests/sqrt2.cpp:30-57 (float, manual accumulation)."""
    f = np.float32

    def loss(x, g, H):
        res = f(x[0] * x[0] - f(2))
        J = f(2) * x[0]
        if g is not None:
            g[0] = J * res
            H[0, 0] = J * J
        return float(f(res * res))

    out = O.optimize(x0, loss, sqrt2_options(), dtype=np.float32)
    assert out.Succeeded() and out.Converged()
    assert abs(out.x[0]) == pytest.approx(math.sqrt(2.0), abs=1e-5)


@pytest.mark.parametrize("x0", [1.0, -0.3, 3.2])
@pytest.mark.parametrize("variant", ["jet_half", "jet2", "jet"])
def test_sqrt2_jet(x0, variant):
    """tests/sqrt2.cpp:59-104: AD double; downscale_by_2 variant; two-residual variant."""
    x0 = float(np.float32(x0))
    if variant == "jet" and x0 <= 0:
        pytest.skip("sqrt2.cpp:112 runs this variant only for x0 > 0")
    if variant == "jet2":
        acc = nlls_acc(lambda x, _: ([x[0] * x[0] - 2.0, 0.1 * (x[0] * x[0] - 2.0)],
                                     [[2 * x[0]], [0.2 * x[0]]]))
        opt = sqrt2_options()
    else:
        acc = nlls_acc(lambda x, _: (x[0] * x[0] - 2.0, [[2 * x[0]]]))
        opt = sqrt2_options(downscale_by_2=1) if variant == "jet_half" else sqrt2_options()
    out = O.optimize(x0, acc, opt)
    assert out.Succeeded() and out.Converged()
    assert abs(out.x[0]) == pytest.approx(math.sqrt(2.0), abs=1e-5)


def test_sqrt2_readme_trajectory():
    """README.md:77-95 / SURVEY.md §8c golden: x0 = 1, defaults, 5 Steps, kMinError."""
    out = O.optimize(1.0, nlls_acc(lambda x, _: (x[0] * x[0] - 2.0, [[2 * x[0]]])))
    assert out.stop_reason == K["kMinError"] and out.num_iters == 5
    xs = [float(v[0]) for v in out.xs]
    gold = [1.499950005001, 1.416672217686, 1.414215777771, 1.414213562399, 1.4142135623730951]
    assert xs == pytest.approx(gold, rel=1e-12)
    assert out.errs == pytest.approx([1.0, 6.242503e-2, 4.844400e-5, 3.926395e-11, 5.554745e-21], rel=1e-6)
    lam0 = float(np.float32(1e-4))
    assert out.lambdas == pytest.approx([lam0 * float(np.float32(1.0) / np.float32(3.0)) ** k for k in range(5)], rel=1e-12)
    # README.md:91-95 prints |dx| of the first three steps
    steps = np.abs(np.diff([1.0] + xs))
    assert steps[:3] == pytest.approx([5.00e-1, 8.33e-2, 2.45e-3], rel=5e-3)


# ---- tests/basic.cpp -----------------------------------------------------------------------------
def loss_xm2(x, g, H):
    res = x[0] - 2
    if g is not None:
        H[0, 0] = 1
        g[0] = res
    return abs(res)


def success_checks(out, expected, lo=2, hi=5):
    """tests/basic.cpp:22-37."""
    assert out.Succeeded()
    assert lo <= out.num_iters <= hi
    if lo > 0:
        assert out.final_cost < 1e-5 and out.Converged()
        assert len(out.errs) == out.num_iters == len(out.successes) == len(out.deltas2)
    assert out.final_hessian is not None and out.final_hessian[0, 0] > 0
    assert out.stop_reason == expected


def test_basic_lm():
    """tests/basic.cpp:41-54; SURVEY §8c: 3 Steps, kMinDeltaNorm, cost 9.998e-9."""
    out = O.optimize(1.0, loss_xm2)
    success_checks(out, K["kMinDeltaNorm"])
    assert out.num_iters == 3
    assert out.x[0] == pytest.approx(1.9999999999996667, rel=1e-13)
    assert out.final_cost == pytest.approx(9.998e-9, rel=1e-3)


def test_basic_gn():
    """tests/basic.cpp:73-88 (GaussNewton) and :106-122 (min_error = 1e-2)."""
    success_checks(O.optimize(1.0, loss_xm2, O.default_options(solver_type=1)), K["kMinError"])
    success_checks(O.optimize(1.0, loss_xm2, O.default_options(solver_type=1, min_error=1e-2)), K["kMinError"])


def failure_checks(out, expected, max_iters=1):
    """tests/basic.cpp:147-156."""
    assert not out.Succeeded() and not out.Converged()
    assert out.num_iters <= max_iters
    assert out.errs == [] and out.successes == [] and out.deltas2 == []
    assert out.stop_reason == expected


def test_basic_failures():
    """tests/basic.cpp:158-258."""
    def nan_grad(x, g, H):
        if g is not None:
            H[0, 0] = 1; g[0] = float("nan")
        return abs(x[0] - 2)

    def inf_grad(x, g, H):
        if g is not None:
            H[0, 0] = 1; g[0] = float("inf")
        return abs(x[0] - 2)

    def inf_res(x, g, H):
        if g is not None:
            H[0, 0] = 1; g[0] = float("inf")
        return float("inf")

    def inf_cost(x, g, H):
        if g is not None:
            H[0, 0] = 1; g[0] = x[0] + 1
        return float("inf")

    for f in (nan_grad, inf_grad, inf_res, inf_cost):
        failure_checks(O.optimize(1.0, f), K["kSystemHasNaNOrInf"])

    def forgot(x, g, H):  # basic.cpp:219-232
        return abs(x[0] - 2)
    failure_checks(O.optimize(1.0, forgot, O.default_options(solver_type=1, check_min_H_diag=1e-7)),
                   K["kSolverFailed"], 3)

    def no_res(x, g, H):  # basic.cpp:234-242
        return np.zeros(0)
    failure_checks(O.optimize(1.0, no_res), K["kSkipped"])

    failure_checks(O.optimize(np.zeros(0, np.float32), loss_xm2, dtype=np.float32), K["kSkipped"])  # :244-258
    out = O.optimize(np.zeros(100000), loss_xm2)  # :260-281
    failure_checks(out, K["kOutOfMemory"])


# ---- tests/solvers.cpp ---------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_solver_single_build_solve(dtype):
    """tests/solvers.cpp:20-45: r = x - y from x = 0, one Build+Solve: dx ~ y (+-1e-2)."""
    y = np.array([4, 5], dtype)
    J = np.eye(2, dtype=dtype)
    r = -y
    lm = O.build_solve(J, r, lam=float(np.float32(1e-4)))
    gn = O.build_solve(J, r, lam=0.0)
    for res in (lm, gn):
        assert res["status"] == 0
        assert res["dx"] == pytest.approx(y, abs=1e-2)
    assert gn["dx"] == pytest.approx(y, abs=1e-12)


def test_solver_skip_rebuild():
    """tests/solvers.cpp:74-109: after Rebuild(false) Build must not re-run the gradient branch.
    Through the optimizer: the bad-step trace of SURVEY §8c (atan) has rebuild flags T,T,T,F,F,F and
    the gradient branch runs exactly 3 times."""
    calls = {"grad": 0, "cost_only": 0}

    def acc(x, g, H):
        r = math.atan(x[0]); J = 1 / (1 + x[0] ** 2)
        if g is not None:
            calls["grad"] += 1
            H[0, 0] = J * J; g[0] = J * r
        else:
            calls["cost_only"] += 1
        return r * r

    out = O.optimize(3.0, acc)
    assert out.stop_reason == K["kMaxConsecNoDecr"] and out.num_iters == 6
    assert out.x[0] == 3.0
    assert out.rebuilt == [True, True, True, False, False, False]
    assert calls == {"grad": 3, "cost_only": 3}
    lam0 = float(np.float32(1e-4))
    assert out.lambdas == pytest.approx([lam0, 2 * lam0, 8 * lam0, 64 * lam0, 1024 * lam0, 32768 * lam0], rel=1e-12)


# ---- tests/optimizers.cpp:24-112 (float sqrt2 through Optimizer_<SolverLM<Mat1f>>) ---------------
def test_optimizer_float_sqrt2():
    f = np.float32
    acc = nlls_acc(lambda x, _: (f(x[0] * x[0] - f(2)), [[f(2) * x[0]]]))
    out = O.optimize(1.0, acc, dtype=np.float32)
    assert out.Succeeded()
    assert out.x[0] == pytest.approx(math.sqrt(2.0), abs=1e-5)


# ---- tests/types.cpp -----------------------------------------------------------------------------
def test_types_vectors():
    """tests/types.cpp:17-65: priors on Vec2 / VecXf(3), sum constraints on 3 params."""
    out = O.optimize(np.ones(2), nlls_acc(lambda x, _: (x - 2.0, np.eye(2))))
    assert np.abs(out.x - 2).sum() == pytest.approx(0, abs=1e-5)
    out = O.optimize(np.ones(2), nlls_acc(lambda x, _: (x[0] + x[1] - 10.0, [[1, 1]])))
    assert out.x.sum() == pytest.approx(10, abs=1e-5)
    out = O.optimize(np.ones(3), nlls_acc(lambda x, _: (x - np.float32(2), np.eye(3))), dtype=np.float32)
    assert np.abs(out.x - 2).sum() == pytest.approx(0, abs=1e-5)
    out = O.optimize([1, 2, 3], nlls_acc(lambda x, _: (x.sum() - 10.0, [[1, 1, 1]])))
    assert out.x.sum() == pytest.approx(10, abs=1e-5)
    out = O.optimize([1, 2, 3], nlls_acc(lambda x, _: (x.sum() - np.float32(10), [[1, 1, 1]])), dtype=np.float32)
    assert out.x.sum() == pytest.approx(10, abs=1e-5)


def test_types_singular_6x6():
    """tests/types.cpp:94-108: 6x6 rank-2 J^T J, float, solved thanks to the LM damping; pins the
    pivoted LDLT on a PSD-singular matrix."""
    rng = np.random.default_rng(0)
    x0 = rng.uniform(-1, 1, 6).astype(np.float32)
    J = np.array([[1, 0, 1, 0, 1, 0], [0, 1, 0, 1, 0, 1]], np.float32)

    def acc(x, g, H):
        res = (x[0:2] + x[2:4] + x[4:6] - np.float32(10)).astype(np.float32)
        if g is not None:
            H[:, :] = J.T @ J
            g[:] = J.T @ res
        return res

    out = O.optimize(x0, acc, dtype=np.float32)
    x = out.x
    assert np.linalg.norm(x[0:2] + x[2:4] + x[4:6] - 10) == pytest.approx(0, abs=1e-5)
    # undamped, the same matrix is only PSD: Eigen's LDLT still reports Success/isPositive and the
    # solve uses the pseudo-inverse of D
    H = (J.T @ J).astype(np.float64)
    sol = O.solve_ldlt(H, H @ np.arange(6.0))
    assert sol is not None and H @ sol == pytest.approx(H @ np.arange(6.0), abs=1e-12)


# ---- tests/cov.cpp -------------------------------------------------------------------------------
def test_cov_iso_manual():
    """tests/cov.cpp:22-47: H.diagonal = 1/sigma^2; Covariance() recovers sigma to 1e-7."""
    y = 10 * np.array([0.3, -0.7]); sd = np.array([4.2, 4.2])

    def loss(x, g, H):
        res = (x - y) / sd
        if g is not None:
            g[:] = res / sd
            H[0, 0], H[1, 1] = 1 / sd[0] ** 2, 1 / sd[1] ** 2
        return float(np.linalg.norm(res))

    out = O.optimize(np.zeros(2), loss)
    assert out.Succeeded() and out.Converged()
    Cm = out.Covariance()
    assert np.abs(np.sqrt(np.diag(Cm)) - sd).max() == pytest.approx(0, abs=1e-7)


def test_cov_full_matrix():
    """tests/cov.cpp:66-91,118-141: whitened residual L^T (x-y), Cy = [[10,2],[2,4]] recovered 1e-5."""
    y = 2 * np.array([0.25, -0.6]); Cy = np.array([[10.0, 2.0], [2.0, 4.0]])
    Lt = np.linalg.cholesky(np.linalg.inv(Cy)).T
    out = O.optimize(np.zeros(2), nlls_acc(lambda x, _: (Lt @ (x - y), Lt)))
    assert out.Succeeded() and out.Converged()
    assert np.abs(out.Covariance() - Cy).max() == pytest.approx(0, abs=1e-5)
    assert out.x == pytest.approx(y, abs=1e-6)


# ---- tests/optimize_easy.cpp ---------------------------------------------------------------------
def test_rosenbrock():
    """tests/optimize_easy.cpp:35-80: true (indefinite-capable) Hessian, <= 200 iters, (1,1) +-1e-5."""
    def loss(v, g, H):
        x, y = v
        t1, t2 = 1 - x, y - x * x
        if g is not None:
            g[0] = -2 * t1 - 400 * x * t2
            g[1] = 200 * t2
            H[0, 0] = 2 - 400 * y + 1200 * x * x
            H[0, 1] = H[1, 0] = -400 * x
            H[1, 1] = 200
        return t1 * t1 + 100 * t2 * t2

    out = O.optimize([-1.2, 1.0], loss, O.default_options(max_iters=200, min_rerr_dec=0, max_consec_failures=20))
    assert out.Succeeded() and out.Converged()
    assert out.x == pytest.approx([1, 1], abs=1e-5)


def test_easom_plateau():
    """tests/optimize_easy.cpp:89-146: (pi, pi) +-1e-4 from (3,3), damping_init 1e-6."""
    PI = math.pi

    def loss(v, g, H):
        dx, dy = v[0] - PI, v[1] - PI
        ex = math.exp(-(dx * dx + dy * dy))
        cx, cy, sx, sy = math.cos(v[0]), math.cos(v[1]), math.sin(v[0]), math.sin(v[1])
        if g is not None:
            g[0] = cy * ex * (sx + 2 * dx * cx)
            g[1] = cx * ex * (sy + 2 * dy * cy)
            H[0, 0] = cy * ex * (cx - 4 * dx * sx + (2 - 4 * dx * dx) * cx)
            H[1, 1] = cx * ex * (cy - 4 * dy * sy + (2 - 4 * dy * dy) * cy)
            H[0, 1] = H[1, 0] = ex * (sx + 2 * dx * cx) * (sy + 2 * dy * cy)
        return 1 - cx * cy * ex

    out = O.optimize([3.0, 3.0], loss, O.default_options(damping_init=1e-6))
    assert out.Succeeded()
    assert out.x == pytest.approx([PI, PI], abs=1e-4)


def test_powell_singular():
    """tests/optimize_easy.cpp:155-221: singular Hessian at the solution; |x_i| < 1e-3."""
    def loss(v, g, H):
        x1, x2, x3, x4 = v
        t1, t2, t3, t4 = x1 + 10 * x2, x3 - x4, x2 - 2 * x3, x1 - x4
        if g is not None:
            g[0] = 2 * t1 + 40 * t4 ** 3
            g[1] = 20 * t1 + 4 * t3 ** 3
            g[2] = 10 * t2 - 8 * t3 ** 3
            g[3] = -10 * t2 - 40 * t4 ** 3
            H[:, :] = 0
            H[0, 0], H[0, 1], H[1, 0], H[1, 1] = 2, 20, 20, 200
            H[2, 2] += 10; H[2, 3] += -10; H[3, 2] += -10; H[3, 3] += 10
            d3 = 12 * t3 * t3
            H[1, 1] += d3; H[1, 2] += -2 * d3; H[2, 1] += -2 * d3; H[2, 2] += 4 * d3
            d4 = 120 * t4 * t4
            H[0, 0] += d4; H[0, 3] += -d4; H[3, 0] += -d4; H[3, 3] += d4
        return t1 * t1 + 5 * t2 * t2 + t3 ** 4 + 10 * t4 ** 4

    out = O.optimize([3.0, -1.0, 0.0, 1.0], loss,
                     O.default_options(max_iters=200, max_consec_failures=0, min_error=1e-30,
                                       min_rerr_dec=1e-30, damping_init=1e-1))
    assert out.Succeeded()
    assert np.abs(out.x).max() < 1e-3


# ---- tests/optimize_hard.cpp ---------------------------------------------------------------------
def test_beale():
    """tests/optimize_hard.cpp:34-62: AD, 3 residuals, (3, 0.5) +-1e-4."""
    def rj(v, _):
        x, y = v
        r = [1.5 - x + x * y, 2.25 - x + x * y * y, 2.625 - x + x * y ** 3]
        J = [[-1 + y, x], [-1 + y * y, 2 * x * y], [-1 + y ** 3, 3 * x * y * y]]
        return r, J
    out = O.optimize([1.0, 1.0], nlls_acc(rj),
                     O.default_options(max_iters=200, max_consec_failures=0, min_error=1e-30, damping_init=1e-3))
    assert out.Succeeded()
    assert out.x == pytest.approx([3.0, 0.5], abs=1e-4)


def test_himmelblau():
    """tests/optimize_hard.cpp:71-99: (3, 2) +-1e-4 from (3.5, 2.5)."""
    def rj(v, _):
        x, y = v
        return [x * x + y - 11, x + y * y - 7], [[2 * x, 1], [1, 2 * y]]
    out = O.optimize([3.5, 2.5], nlls_acc(rj),
                     O.default_options(max_iters=200, max_consec_failures=0, min_error=1e-30, damping_init=1e-4))
    assert out.x == pytest.approx([3.0, 2.0], abs=1e-4)


def test_jennrich_sampson():
    """tests/optimize_hard.cpp:222-283: 10 residuals, x0 ~ x1 +-1e-5 (~0.2578)."""
    def rj(v, _):
        i = np.arange(1, 11, dtype=float)
        e0, e1 = np.exp(i * v[0]), np.exp(i * v[1])
        return 2 + 2 * i - (e0 + e1), np.stack([-i * e0, -i * e1], axis=1)
    out = O.optimize([0.3, 0.4], nlls_acc(rj),
                     O.default_options(max_iters=500, max_consec_failures=0, min_error=1e-30,
                                       min_rerr_dec=0, damping_init=1e-6))
    assert out.Succeeded()
    assert out.x[0] == pytest.approx(out.x[1], abs=1e-5)
    assert out.x[0] == pytest.approx(0.2578, abs=1e-3)


# ---- tests/circle.cpp:32-66 ----------------------------------------------------------------------
def test_fit_circle():
    """10 points on a circle (r = 2, centre (2,7), noise 1e-5), damping_init = 1e1, +-1e-5."""
    rng = np.random.default_rng(1)
    ang = np.arange(10) * 2 * math.pi / 9
    obs = np.stack([2 + 2 * np.cos(ang), 7 + 2 * np.sin(ang)]) + 1e-5 * rng.uniform(-1, 1, (2, 10))
    obs = obs.astype(np.float32).astype(np.float64)

    def rj(x, _):
        d = obs - x[:2, None]
        r = (d * d).sum(0) - x[2] * x[2]
        J = np.stack([-2 * d[0], -2 * d[1], np.full(10, -2 * x[2])], axis=1)
        return r, J

    out = O.optimize([0.0, 0.0, 1.0], nlls_acc(rj), O.default_options(damping_init=1e1))
    assert out.Succeeded()
    assert out.x == pytest.approx([2, 7, 2], abs=1e-5)


# ---- hessian.use_ldlt = false: dx = -H.inverse() * grad (solvers/gn.h:157-163) ---------------------
def test_rectangle_inverse_path():
    """tests/userdef_params_jet.cpp:85-125 (float, `use_ldlt = false`, damping_init = 0.1).  The
    manifold of params_trait<Rectangle>::PlusEq (:68-76) is additive in u = (p1.x, p1.y, width,
    height), so the test is a plain 4-parameter problem in u; `area()` there is the diagonal's norm."""
    f = np.float32

    def rj(u, want_j):
        w, h = u[2], u[3]
        d = np.sqrt(w * w + h * h)
        hm = max(h, f(1e-8))
        r = np.array([d - f(200), f(100) * (w / hm - f(2)), u[0] + f(0.5) * w - f(1), u[1] + f(0.5) * h - f(2)], f)
        J = None
        if want_j:
            J = np.array([[0, 0, w / d, h / d],
                          [0, 0, f(100) / hm, -f(100) * w / (hm * hm)],
                          [1, 0, 0.5, 0],
                          [0, 1, 0, 0.5]], f)
        return r, J

    out = O.optimize(np.array([0, 0, 1, 1], f), nlls_acc(rj), O.default_options(use_ldlt=0, damping_init=1e-1),
                     dtype=np.float32)
    assert out.Succeeded()
    u = out.x.astype(np.float64)
    assert math.hypot(u[2], u[3]) == pytest.approx(200, abs=1e-3)      # the reference asserts 1e-5 on its
    assert u[0] + 0.5 * u[2] == pytest.approx(1, abs=1e-4)              # own float trajectory; the restated
    assert u[1] + 0.5 * u[3] == pytest.approx(2, abs=1e-4)              # Jacobian is analytic, not Jets
    assert u[2] == pytest.approx(2 * u[3], abs=1e-3)
    # and the inverse path lands where the LDLT path does
    ref = O.optimize(np.array([0, 0, 1, 1], f), nlls_acc(rj), O.default_options(damping_init=1e-1), dtype=np.float32)
    assert ref.Succeeded() and out.x == pytest.approx(ref.x, rel=1e-4)


@pytest.mark.parametrize("dtype,x0", [(np.float32, 1.0), (np.float64, 0.7), (np.float64, -0.9)])
def test_sqrt2_inverse_path(dtype, x0):
    """benchmarks/dense.cpp:27-52: the sqrt(2) benchmarks run with `use_ldlt = false`, i.e. the
    Dims == 1 branch `H(0,0) > FloatEpsilon ? -H.inverse() * grad : 0` (gn.h:158-160)."""
    f = dtype

    def loss(x, want_j):
        return f(x[0] * x[0] - f(2)), (f(2) * x[0] if want_j else None)

    out = O.optimize(x0, nlls_acc(loss), O.default_options(use_ldlt=0), dtype=dtype)
    assert out.Succeeded() and out.Converged()
    assert abs(out.x[0]) == pytest.approx(math.sqrt(2.0), abs=1e-5)
    ref = O.optimize(x0, nlls_acc(loss), O.default_options(), dtype=dtype)
    assert out.num_iters == ref.num_iters and out.x[0] == pytest.approx(ref.x[0], rel=1e-6)


def test_inverse_path_guard_and_singular_system():
    """gn.h:158-160: a scalar H at or below FloatEpsilon gives dx = 0 instead of a division;
    gn.h:162: above one dimension there is no check at all, so a singular H ends the run through
    Step's NaN/Inf guard (optimizer.h:405-425) where the LDLT path solves the damped system."""
    def flat(x, g, H):  # J = 0: H = 0 <= FloatEpsilon
        if g is not None:
            g[0] = 0.0
            H[0, 0] = 0.0
        return 1.0
    out = O.optimize(1.0, flat, O.default_options(use_ldlt=0))
    assert out.x[0] == 1.0 and out.deltas2[0] == 0.0

    def rank1(x, g, H):  # J = [1 0]: second row and column of H are zero
        r = x[0] - 2.0
        if g is not None:
            g[:] = (r, 0.0)
            H[:, :] = ((1.0, 0.0), (0.0, 0.0))
        return r * r
    bad = O.optimize(np.array([1.0, 1.0]), rank1, O.default_options(use_ldlt=0))
    assert bad.stop_reason == K["kSystemHasNaNOrInf"]
    good = O.optimize(np.array([1.0, 1.0]), rank1, O.default_options())
    assert good.Converged() and good.x[0] == pytest.approx(2.0, abs=1e-6) and good.x[1] == 1.0


# ---- diff/num_diff.h: numeric differentiation handed to the optimizer ---------------------------------------------------
def _create_numdiff_func2(residuals, method="central", h=None, dtype=np.float64):
    """`CreateNumDiffFunc2(x, residuals, method, h)` (diff/num_diff.h:284-309) over `NumEval` (:57-126) as an accumulation
    lambda for O.optimize: J column by column from f(x + h e_r) and f(x - h e_r) (kCentral), (f(x + h e_r) - f(x)) / h
    (kForward) or the second point at (x_r + h) - 2h (kFastCentral); grad = J^T res, H = J^T J;
    Cost(res.norm(), res.size()) - the NORM.  h defaults to FloatEpsilon (math.h:297-301)."""
    T = np.dtype(dtype).type
    h = T(np.float32(1e-7 if dtype == np.float64 else 1e-4)) if h is None else T(h)

    def acc(x, g, H):
        x = np.asarray(x, dtype)
        res = np.atleast_1d(residuals(x)).astype(dtype)
        if g is not None:
            J = np.zeros((res.size, x.size), dtype)
            for r in range(x.size):
                y = x.copy(); y[r] = T(x[r] + h)
                rp = np.atleast_1d(residuals(y)).astype(dtype)
                if method == "forward":
                    J[:, r] = (rp - res) / h
                else:
                    if method == "central":
                        y = x.copy(); y[r] = T(x[r] + T(-h))
                    else:
                        y[r] = T(y[r] + T(-2) * h)
                    rm = np.atleast_1d(residuals(y)).astype(dtype)
                    J[:, r] = (rp - rm) / (T(2) * h)
            g[:] = J.T @ res
            H[:, :] = J.T @ J
        return float(np.sqrt((res * res).sum(dtype=dtype))), int(res.size)
    return acc


def test_numdiff_gradient_matches_reference_assertions():
    """tests/diff.cpp:59-86 `TestCreateNumDiffFunc2`: loss = 2 (x - y_prior) at x = 0: g == 2 res (+-1e-5), for a
    random prior; the scalar float case r = x - 2: g == res (+-1e-3)."""
    rng = np.random.default_rng(3)
    yp = rng.uniform(-1, 1, 3)
    acc = _create_numdiff_func2(lambda x: 2 * (x - yp))
    g = np.zeros(3); H = np.zeros((3, 3))
    cost, nres = acc(np.zeros(3), g, H)
    res = 2 * (np.zeros(3) - yp)
    assert np.abs(g - 2 * res).max() < 1e-5 and nres == 3
    assert abs(cost - np.linalg.norm(res)) < 1e-12          # the NORM, not its square (num_diff.h:305)
    assert np.abs(H - 4 * np.eye(3)).max() < 1e-5
    accf = _create_numdiff_func2(lambda x: x - np.float32(2), dtype=np.float32)
    g1 = np.zeros(1, np.float32); H1 = np.zeros((1, 1), np.float32)
    accf(np.zeros(1, np.float32), g1, H1)
    assert abs(g1[0] - (-2.0)) < 1e-3


@pytest.mark.parametrize("method", ["central", "forward", "fast"])
def test_numdiff_optimizer_converges(method):
    """tests/optimizers.cpp:100-120: loss = x - y_prior, y_prior = (3, 2, 1), x0 = 0, `CreateNumDiffFunc2` handed to
    `Optimizer_<SolverLM<Mat3>>`: Succeeded and Converged (and it lands on the prior)."""
    yp = np.array([3.0, 2.0, 1.0])
    o = O.optimize(np.zeros(3), _create_numdiff_func2(lambda x: x - yp, method), O.default_options())
    assert o.stop_reason > 0                                  # Succeeded() && Converged()
    assert np.abs(o.x - yp).max() < 1e-5
    assert o.final_cost < 1e-5 and o.num_iters <= 10


def test_sparse_reference_case_through_the_dense_restatement():
    """tests/sparse.cpp:19-57 "tinyopt_sparse": res = 10 x - 2 over 100 parameters, H = J^T J (diagonal, handed over as a
    sparse matrix in the reference), Cost(res.norm(), res.size()), check_final_cost = false.  The product solves the
    scattered triplets with the dense pivoted LDLT (tob200_solver_step_hg_sparse_*); for this positive definite H that is
    the sparse factorisation's answer: the reference's assertions hold for the dense restatement (Succeeded, Converged,
    min / max of x == 0.2 +- 1e-5)."""
    rng = np.random.default_rng(11)
    n = 100
    x0 = rng.uniform(-1, 1, n)

    def acc(x, g, H):
        res = 10 * x - 2.0
        if g is not None:
            g[:] = 10 * res
            H[:, :] = 0
            H[np.arange(n), np.arange(n)] = 100.0
        return float(np.linalg.norm(res)), n

    o = O.optimize(x0, acc, O.default_options(check_final_cost=0))
    assert o.stop_reason > 0
    assert abs(o.x.min() - 0.2) < 1e-5 and abs(o.x.max() - 0.2) < 1e-5
