"""GPU parity of `hessian.use_ldlt = false` (include/tinyopt/solvers/gn.h:157-163: dx = -H.inverse() *
grad, no invertibility check; options.h:59) against the oracle's partial-pivot LU, bit for bit.
The reference's own sqrt(2) benchmarks (benchmarks/dense.cpp:27-52) and tests/userdef_params_jet.cpp
run with this option.  Thread-per-problem (n <= 12 / 8) and warp-per-problem (n <= 55) families.
"""
import math

import numpy as np
import pytest
import torch

from oracle import oracle as O
from test_gpu_parity import assert_lm_parity, run_both
from test_gpu_solver import drive

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import tinyopt_b200 as tb
    c = tb.Context(0)
    yield c
    c.close()


F64 = [(500, 30, 6), (33, 4, 1), (64, 12, 3), (50, 9, 7), (100, 20, 8),       # thread per problem
       (40, 40, 10), (25, 60, 21), (21, 77, 30), (12, 120, 50), (9, 131, 55)]  # warp per problem
F32 = [(500, 200, 12), (45, 30, 6), (10, 5, 1), (77, 17, 11),
       (40, 64, 13), (33, 90, 27), (35, 100, 28), (20, 77, 40), (9, 131, 55)]


@pytest.mark.parametrize("B,m,n", F64)
def test_inverse_path_parity_f64(ctx, B, m, n):
    xo, ro, out = run_both(ctx, np.float64, B, m, n, use_ldlt=0)
    assert_lm_parity(np.float64, xo, ro, out)
    assert (ro["stop_reason"] > 0).all()


@pytest.mark.parametrize("B,m,n", F32)
def test_inverse_path_parity_f32(ctx, B, m, n):
    xo, ro, out = run_both(ctx, np.float32, B, m, n, use_ldlt=0)
    assert_lm_parity(np.float32, xo, ro, out)


@pytest.mark.parametrize("dtype,B,m,n", [(np.float64, 200, 30, 6), (np.float32, 30, 90, 27)])
@pytest.mark.parametrize("optkw", [
    dict(solver_type=1),                              # Gauss-Newton through H.inverse()
    dict(H_is_full=0),                                # upper -> lower mirror before the inverse (lm.h:89-94)
    dict(damping_init=10.0, max_consec_failures=2),
    dict(min_rerr_dec=1e-10, min_step_norm2=1e-14),   # float: rejects, roll-backs, stale re-damped H_
    dict(grad_clipping=0.05),
])
def test_inverse_path_option_variants(ctx, dtype, B, m, n, optkw):
    xo, ro, out = run_both(ctx, dtype, B, m, n, use_ldlt=0, **optkw)
    assert_lm_parity(dtype, xo, ro, out)


def test_inverse_path_layouts_agree(ctx):
    import tinyopt_b200 as tb
    for dtype, shape in ((np.float64, (100, 30, 6)), (np.float32, (40, 64, 13))):
        _, _, out_t = run_both(ctx, dtype, *shape, layout=tb.TILE32, use_ldlt=0)
        _, _, out_p = run_both(ctx, dtype, *shape, layout=tb.PROBLEM_MAJOR, use_ldlt=0)
        assert torch.equal(out_t.x, out_p.x)
        assert np.array_equal(out_t.results, out_p.results)


def test_inverse_path_large_n_runs_in_the_general_family(ctx):
    """n > 55: H.inverse() is served by the general kernel family's partial-pivot LU (gn.cuh; bit-exact parity in
    tests/test_gpu_general.py::test_inverse_path_above_55) - never a silent LDLT: the two solves give different
    bits."""
    import tinyopt_b200 as tb
    B, m, n = 4, 128, 64
    fl = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, torch.float32, layout=tb.PROBLEM_MAJOR)
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32)
    xo, ro, _ = O.synth_lm_run(A, y, x0, O.default_options(use_ldlt=0, **fl))
    out = ctx.optimize_batch(dA, dy, dx0, tb.options(use_ldlt=0, **fl), layout=tb.PROBLEM_MAJOR)
    assert np.array_equal(out.x.cpu().numpy(), xo) and np.array_equal(out.results["num_iters"], ro["num_iters"])
    out2 = ctx.optimize_batch(dA, dy, dx0, tb.options(**fl), layout=tb.PROBLEM_MAJOR)
    assert (out2.results["stop_reason"] > 0).all() and not torch.equal(out.x, out2.x)


def test_sqrt2_benchmark_options(ctx):
    """benchmarks/dense.cpp:27-52 through the solver seam: scalar LM with `use_ldlt = false`, i.e. the
    guarded Dims == 1 branch (gn.h:158-160); equals the oracle step for step."""
    import tinyopt_b200 as tb
    starts = [1.0, 0.7, -0.9, 3.2]
    x0 = torch.tensor([[s] for s in starts], dtype=torch.float64, device="cuda")

    def f(x):
        return (x * x - 2.0), (2.0 * x).unsqueeze(-1)

    x, res, H = drive(ctx, x0, f, tb.options(use_ldlt=0))

    def acc(xv, g, Hm):
        r = xv[0] * xv[0] - 2.0
        J = 2 * xv[0]
        if g is not None:
            g[0] = J * r
            Hm[0, 0] = J * J
        return r * r

    for b, s in enumerate(starts):
        o = O.optimize(s, acc, O.default_options(use_ldlt=0))
        assert res["num_iters"][b] == o.num_iters and res["stop_reason"][b] == o.stop_reason
        assert x[b, 0] == o.x[0] and res["final_cost"][b] == o.final_cost
        assert abs(abs(x[b, 0]) - math.sqrt(2)) < 1e-5


def test_inverse_path_guard_and_singular_system(ctx):
    """gn.h:158-160: scalar H <= FloatEpsilon -> dx = 0; gn.h:162: a singular H above one dimension
    is not caught by the solver and ends in kSystemHasNaNOrInf (optimizer.h:405-425), where the LDLT
    path converges.  Both families (n = 2: registers, n = 14: warp), against the oracle."""
    import tinyopt_b200 as tb
    K = O.K
    # flat scalar problem
    x0 = torch.ones((2, 1), dtype=torch.float64, device="cuda")

    def flat(x):
        return torch.ones_like(x), torch.zeros((x.shape[0], 1, 1), dtype=x.dtype, device="cuda")

    x, res, _ = drive(ctx, x0, flat, tb.options(use_ldlt=0))

    def flat_acc(xv, g, H):
        if g is not None:
            g[0] = 0.0
            H[0, 0] = 0.0
        return 1.0
    o = O.optimize(1.0, flat_acc, O.default_options(use_ldlt=0))
    assert (res["num_iters"] == o.num_iters).all() and (res["stop_reason"] == o.stop_reason).all()
    assert (x == 1.0).all()

    for n in (2, 14):
        x0 = torch.ones((3, n), dtype=torch.float64, device="cuda")

        def rank1(x, n=n):  # J = e_0^T: every other row and column of H is zero
            J = torch.zeros((x.shape[0], 1, n), dtype=x.dtype, device="cuda")
            J[:, 0, 0] = 1.0
            return x[:, :1] - 2.0, J

        def rank1_acc(xv, g, H, n=n):
            r = xv[0] - 2.0
            if g is not None:
                g[:] = 0.0
                g[0] = r
                H[:, :] = 0.0
                H[0, 0] = 1.0
            return r * r

        x, res, _ = drive(ctx, x0, rank1, tb.options(use_ldlt=0))
        o = O.optimize(np.ones(n), rank1_acc, O.default_options(use_ldlt=0))
        assert o.stop_reason == K["kSystemHasNaNOrInf"]
        assert (res["stop_reason"] == o.stop_reason).all() and (res["num_iters"] == o.num_iters).all()
        assert np.array_equal(x, np.tile(o.x, (3, 1)))
        x, res, _ = drive(ctx, x0, rank1, tb.options())
        o = O.optimize(np.ones(n), rank1_acc, O.default_options())
        assert o.Converged() and (res["stop_reason"] == o.stop_reason).all()
        assert np.array_equal(x, np.tile(o.x, (3, 1)))


@pytest.mark.parametrize("dtype,B,m,n", [(np.float64, 200, 30, 6), (np.float32, 100, 50, 10),
                                         (np.float32, 40, 90, 20), (np.float64, 24, 100, 40),
                                         (np.float32, 17, 131, 55)])
def test_inverse_path_solver_seam(ctx, dtype, B, m, n):
    """The host-driven loop (tob200_solver_*: tpp_step_kernel / the kWppStepInv kernels) with
    `use_ldlt = false`, fed with the synthetic family's residual blocks, equals the oracle exactly;
    the final Hessian it exports is still the un-damped J^T J (solvers/lm.h:157-171)."""
    import tinyopt_b200 as tb
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    layout = tb.TILE32 if ctx.kernel_family(tdt, n) == 1 else tb.PROBLEM_MAJOR
    kw = dict(use_ldlt=0, **(dict(min_rerr_dec=1e-5, min_step_norm2=1e-9) if dtype == np.float32 else {}))
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=5)
    xo, ro, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw))
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, tdt, p0=5, layout=layout)
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
    H = s.final_hessian().cpu().numpy()
    assert np.array_equal(H, np.swapaxes(H, 1, 2)) and (np.einsum("bii->bi", H) > 0).all()
    s.close()
