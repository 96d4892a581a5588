"""CPU checks of the oracle's numerics that do not come from a reference test: the restated Eigen
LDLT against numpy on random matrices, its accept / reject semantics, the
synthetic family's golden fixture and its statistical properties."""
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synth_family.npz")


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-11), (np.float32, 2e-3)])
@pytest.mark.parametrize("n", [1, 2, 3, 6, 12, 50, 129])
def test_ldlt_solves_spd(dtype, tol, n):
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n + 5, n))
    A = (M.T @ M + 0.1 * np.eye(n)).astype(dtype)
    b = rng.standard_normal(n).astype(dtype)
    x = O.solve_ldlt(np.triu(A), b)  # only the upper triangle may be read
    assert x is not None
    ref = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
    assert np.abs(x - ref).max() <= tol * max(1.0, np.abs(ref).max()) * np.linalg.cond(A.astype(np.float64))
    inv = O.inv_cov(A)
    assert np.abs(inv.astype(np.float64) @ A.astype(np.float64) - np.eye(n)).max() < tol * 1e3 * n


def test_ldlt_accept_reject_semantics():
    """isPositive() && info()==Success (math.h:236): indefinite and negative-definite are rejected,
    PSD-singular and all-zero are accepted (pseudo-inverse of D), size 1 by the sign of A00."""
    assert O.solve_ldlt(np.array([[2.0, 0], [0, -1.0]]), np.ones(2)) is None
    assert O.solve_ldlt(-np.eye(3), np.ones(3)) is None
    assert O.solve_ldlt(np.array([[1.0, 2.0], [2.0, 1.0]]), np.ones(2)) is None  # indefinite
    x = O.solve_ldlt(np.zeros((3, 3)), np.ones(3))
    assert x is not None and np.all(x == 0)
    x = O.solve_ldlt(np.diag([4.0, 0.0, 1.0]), np.array([4.0, 7.0, 2.0]))
    assert list(x) == [1.0, 0.0, 2.0]
    assert O.solve_ldlt(np.array([[-1.0]]), np.ones(1)) is None
    assert O.solve_ldlt(np.array([[0.0]]), np.ones(1))[0] == 0.0
    # pivoting: the largest diagonal goes first, result independent of the ordering
    A = np.array([[1e-8, 1e-4], [1e-4, 4.0]])
    assert O.solve_ldlt(A, A @ np.array([1.0, 2.0])) == pytest.approx([1.0, 2.0], rel=1e-6)
    # zero diagonal with non-zero off-diagonal: NumericalIssue
    assert O.solve_ldlt(np.array([[0.0, 1.0], [1.0, 0.0]]), np.ones(2)) is None


def test_synth_family_golden_fixture():
    """tests/golden/synth_family.npz (made by tests/golden/make_golden.py from the oracle at the
    commit that pinned it): generator and LM results must not drift."""
    g = np.load(GOLDEN)
    for tag, dtype, (B, m, n), kw in (("c2", np.float64, (8, 30, 6), {}),
                                      ("c3", np.float32, (8, 200, 12), dict(min_rerr_dec=1e-5, min_step_norm2=1e-9))):
        A, y, xs, x0 = O.synth_generate(B, m, n, dtype)
        x, res, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw), nthreads=1)
        assert np.array_equal(A, g[f"{tag}_A"]) and np.array_equal(y, g[f"{tag}_y"])
        assert np.array_equal(x0, g[f"{tag}_x0"]) and np.array_equal(xs, g[f"{tag}_xstar"])
        assert np.array_equal(x, g[f"{tag}_x"])
        assert np.array_equal(res["num_iters"], g[f"{tag}_num_iters"])
        assert np.array_equal(res["stop_reason"], g[f"{tag}_stop_reason"])
        assert np.array_equal(res["final_cost"], g[f"{tag}_final_cost"])


def test_synth_family_statistics():
    """A ~ U(-1,1)/sqrt(n), x* ~ U(-1,1), x0 = x* + 0.3 U, noise sigma = 1e-2,
    cond(J^T J) small, seeded and stateless (any slice regenerates identically)."""
    B, m, n = 4000, 30, 6
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float64)
    assert abs(A.mean()) < 2e-3 and A.std() == pytest.approx(1 / np.sqrt(3 * n), rel=2e-2)
    assert np.abs(A).max() <= 1 / np.sqrt(n)
    assert np.abs(x0 - xs).max() <= 0.3 and (x0 - xs).std() == pytest.approx(0.3 / np.sqrt(3), rel=3e-2)
    t = np.einsum("bij,bj->bi", A, xs)
    noise = y - (t + 0.1 * t ** 3)
    assert noise.std() == pytest.approx(1e-2, rel=3e-2) and np.abs(noise).max() <= 1e-2 * np.sqrt(3) * 1.0001
    A2, y2, xs2, x02 = O.synth_generate(100, m, n, np.float64, p0=1234)
    assert np.array_equal(A2, A[1234:1334]) and np.array_equal(y2, y[1234:1334]) and np.array_equal(x02, x0[1234:1334])
    r, J = O.synth_eval(A[:200], y[:200], x0[:200])
    conds = [np.linalg.cond(J[b].T @ J[b]) for b in range(200)]
    assert max(conds) < 60
    # float and double families share the construction (not the bits)
    Af, yf, xsf, x0f = O.synth_generate(50, m, n, np.float32)
    assert Af.dtype == np.float32 and np.abs(Af).max() <= np.float32(1 / np.sqrt(n)) * 1.000001


def test_batch_runner_equals_single_calls():
    """too_synth_lm_run (OpenMP over problems) == one too_optimize per problem, any thread count."""
    B, m, n = 64, 30, 6
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float64, p0=7)
    x1, r1, _ = O.synth_lm_run(A, y, x0, nthreads=1)
    x2, r2, used = O.synth_lm_run(A, y, x0, nthreads=0)
    assert np.array_equal(x1, x2) and np.array_equal(r1, r2) and used >= 1
    assert (r1["stop_reason"] > 0).all() and (r1["num_iters"] >= 2).all()
    # and equals the callback path with the same accumulation written in numpy-free python
    for b in (0, 17):
        def acc(x, g, H, b=b):
            r, J = O.synth_eval(A[b:b + 1], y[b:b + 1], x[None, :])
            r, J = r[0], J[0]
            c = np.float64(0)
            for i in range(m):
                c = np.float64(np.fma(r[i], r[i], c)) if hasattr(np, "fma") else c + r[i] * r[i]
            if g is not None:
                g[:] = 0
                H[:, :] = 0
                H[:, :] = J.T @ J
                g[:] = J.T @ r
            return float(c), m
        o = O.optimize(x0[b], acc)
        assert o.num_iters == r1["num_iters"][b] and o.stop_reason == r1["stop_reason"][b]
        assert o.x == pytest.approx(x1[b], rel=1e-9)


@pytest.mark.parametrize("B,m,n,dtype,kw", [
    (300, 30, 6, np.float64, {}), (200, 200, 12, np.float32, dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)),
    (40, 500, 50, np.float32, dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)), (40, 61, 33, np.float32, {}),
    (12, 300, 57, np.float64, {}), (3, 517, 130, np.float32, dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)),
    (50, 45, 9, np.float64, dict(solver_type=1)), (30, 20, 17, np.float32, dict(damping_init=10.0))])
def test_fast_build_is_bit_identical(B, m, n, dtype, kw):
    """The -DTOO_FAST build (rows in blocks of 8, interleaved chains, vectorised k loops: the CPU arm
    bench.py reports beside the canonical one) gives the SAME bits as the canonical restatement —
    every accumulator still receives its terms in row order."""
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=5)
    opt = O.default_options(**kw)
    x1, r1, _ = O.synth_lm_run(A, y, x0, opt)
    for nt in (1, 3):
        x2, r2, _ = O.synth_lm_run(A, y, x0, opt, nthreads=nt, fast=True)
        assert np.array_equal(x1, x2)
        for k in r1.dtype.names:
            assert np.array_equal(r1[k], r2[k]), k


@pytest.mark.parametrize("dtype,method", [(np.float64, 2), (np.float64, 1), (np.float64, 3), (np.float32, 2)])
def test_numdiff_variant_of_the_family(dtype, method):
    """diff/num_diff.h restated (NumEval :57-126, CreateNumDiffFunc2 :284-309): the Jacobian from central / forward /
    fast-central differences with h = FloatEpsilon, the Cost the residual NORM.  Pinned the way tests/num_diff.cpp pins
    the reference: the estimate agrees with the analytic Jacobian to the method's order (same minimiser as the analytic
    run), and here additionally: final_cost is the norm (sqrt of the analytic run's r^T r at the same point)."""
    B, m, n = 6, 40, 7
    kw = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9) if dtype == np.float32 else {}
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype)
    xa, ra, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw))
    xn, rn, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw), numdiff=(method, 0.0))
    assert (rn["stop_reason"] > 0).all() and (ra["stop_reason"] > 0).all()
    tol = {(np.float64, 2): 1e-7, (np.float64, 1): 1e-5, (np.float64, 3): 1e-7, (np.float32, 2): 5e-3}[(dtype, method)]
    assert np.abs(xn - xa).max() / np.abs(xa).max() < tol
    # the cost of the numdiff run is a norm: compare with the analytic run's squared norm at (nearly) the same point
    assert np.allclose(rn["final_cost"] ** 2, ra["final_cost"], rtol=1e-3)
    assert (rn["final_num_residuals"] == m).all()
    # the switch is off again: the analytic family is what runs by default
    xb, rb, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw))
    assert np.array_equal(xa, xb)
