"""GPU parity tests of the mid-n tensor-core family (wtc.cuh: 13 <= n <= 55, float; config C4): JᵀJ off-diagonal on
tcgen05 (FP16 hi / lo split with per-column power-of-two scales), g / diag / cost / t as FP32 sums, a
latency-optimised LDLᵀ with the exact pivoted routine behind it — against the CPU oracle and against the bit-exact
warp-per-problem kernel (`Context.set_exact`).

Bars (BASELINE.json north_star, float): x and the final cost within 1e-4 relative PER PROBLEM; iteration counts and
stop reasons identical wherever every accept / stop decision of the run clears FP32 noise (the oracle re-run in double
reports its decision margins), at most one Step apart elsewhere.  The rest of the suite runs with TOB200_WPP_TC=0
(tests/conftest.py), i.e. on the bit-exact kernels; this module is the one that pins the default path.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

FLOAT_OPTS = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)


@pytest.fixture(scope="module")
def ctx():
    import tinyopt_b200 as tb
    c = tb.Context(0)
    c.set_exact(False)  # the tensor-core kernel (the library default; conftest.py selects the exact ones for the other modules)
    yield c
    c.close()


def rel_err_rows(a, b):
    a = np.asarray(a, np.float64).reshape(len(a), -1); b = np.asarray(b, np.float64).reshape(len(b), -1)
    return np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-300)


def oracle_run(A, y, x0, kw):
    xo, ro, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw), fast=True)
    _, r64, _ = O.synth_lm_run(A.astype(np.float64), y.astype(np.float64), x0.astype(np.float64), O.default_options(**kw), fast=True)
    robust = (r64["sign_margin"] > 2e-5) & (r64["thr_margin"] > 0.1)
    return xo, ro, robust


def gpu_run(ctx, A, y, x0, kw, exact=False):
    import tinyopt_b200 as tb
    ctx.set_exact(exact)
    try:
        out = ctx.optimize_batch(torch.from_numpy(A).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(x0).cuda(), tb.options(**kw),
                                 layout=tb.PROBLEM_MAJOR)
    finally:
        ctx.set_exact(False)
    return out.x.cpu().numpy(), out.results


def check(xg, rg, xo, ro, robust, min_robust=0.0, tol=1e-4):
    assert robust.mean() >= min_robust, robust.mean()
    assert np.array_equal(rg["num_iters"][robust], ro["num_iters"][robust])
    assert np.array_equal(rg["stop_reason"][robust], ro["stop_reason"][robust])
    assert np.abs(rg["num_iters"].astype(int) - ro["num_iters"]).max() <= 1
    same = rg["num_iters"] == ro["num_iters"]
    assert rel_err_rows(xg[same], xo[same]).max() <= tol
    assert (np.abs(rg["final_cost"][same] - ro["final_cost"][same]) / np.maximum(np.abs(ro["final_cost"][same]), 1e-300)).max() <= tol
    # a problem that is one noise-floor Step apart has still converged to the same point
    assert rel_err_rows(xg, xo).max() <= 10 * tol


@pytest.mark.parametrize("B,m,n", [(1, 500, 50), (7, 500, 50), (9, 500, 50), (300, 500, 50),   # partial CTAs, idle slots, odd pair counts
                                   (64, 192, 28), (40, 196, 32), (33, 200, 32), (33, 204, 33),   # n <= 32: one drain warp per side
                                   (50, 231, 52), (40, 260, 55), (25, 501, 44),                  # rows not a multiple of the 32-row stage
                                   (30, 300, 51), (30, 268, 35)])                                # odd n: scalar t-chain, single last LDLT column
def test_wtc_lm_run_parity(ctx, B, m, n):
    assert (m * n) % 4 == 0  # the shapes the tensor-core path takes (others fall back to the exact kernel: test below)
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32, p0=1000 * n + m)
    xo, ro, robust = oracle_run(A, y, x0, FLOAT_OPTS)
    xg, rg = gpu_run(ctx, A, y, x0, FLOAT_OPTS)
    check(xg, rg, xo, ro, robust)
    assert (rg["stop_reason"] > 0).all()


def test_wtc_c4_sample_against_oracle_and_exact_kernel(ctx):
    """4096 problems of the C4 shape: the oracle, the exact kernel (== oracle bit for bit) and the tensor-core kernel."""
    B, m, n = 4096, 500, 50
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32)
    xo, ro, robust = oracle_run(A, y, x0, FLOAT_OPTS)
    xe, re_ = gpu_run(ctx, A, y, x0, FLOAT_OPTS, exact=True)
    assert np.array_equal(xe, xo) and np.array_equal(re_["num_iters"], ro["num_iters"])
    xg, rg = gpu_run(ctx, A, y, x0, FLOAT_OPTS)
    check(xg, rg, xo, ro, robust, min_robust=0.98)
    assert (rg["num_iters"] == ro["num_iters"]).mean() >= 0.999
    assert np.median(rel_err_rows(xg, xo)) <= 1e-6
    assert abs(int(rg["num_builds"].sum()) - int(re_["num_builds"].sum())) <= 4


def test_wtc_is_deterministic(ctx):
    B, m, n = 777, 300, 48
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32, p0=5)
    x1, r1 = gpu_run(ctx, A, y, x0, FLOAT_OPTS)
    x2, r2 = gpu_run(ctx, A, y, x0, FLOAT_OPTS)
    assert np.array_equal(x1, x2) and np.array_equal(r1, r2)  # which slot / pair a problem lands in does not matter


@pytest.mark.parametrize("kw", [dict(), dict(solver_type=1), dict(damping_init=1.0), dict(use_squared_norm=0),
                                dict(downscale_by_2=1, normalize=1), dict(grad_clipping=0.5), dict(max_iters=2),
                                dict(check_min_H_diag=1e-6), dict(min_error=1e-2), dict(max_consec_failures=1)],
                         ids=lambda k: ",".join(f"{a}={b}" for a, b in k.items()) or "float-defaults")
def test_wtc_option_variants(ctx, kw):
    B, m, n = 96, 256, 40
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32, p0=77)
    kw = {**FLOAT_OPTS, **kw}
    xo, ro, robust = oracle_run(A, y, x0, kw)
    xg, rg = gpu_run(ctx, A, y, x0, kw)
    check(xg, rg, xo, ro, robust)


def test_wtc_noise_floor_runs(ctx):
    """tinyopt's default thresholds drive float problems into the FP32 noise floor: rejected steps, roll-backs, cost-only
    passes on the stale re-damped H_ (the exact routine on the slot's persistent copy).  Iteration counts are noise there
    in any float implementation; what must hold: every problem converges to the oracle's point and stops for a legal reason."""
    B, m, n = 256, 500, 50
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32, p0=123)
    xo, ro, robust = oracle_run(A, y, x0, {})
    xg, rg = gpu_run(ctx, A, y, x0, {})
    assert rel_err_rows(xg, xo).max() <= 1e-4
    assert (np.abs(rg["final_cost"] - ro["final_cost"]) / ro["final_cost"]).max() <= 1e-4
    assert (rg["stop_reason"] > 0).all()
    assert (rg["num_builds"] < rg["num_iters"]).any()          # cost-only passes did happen
    assert np.array_equal(rg["num_iters"][robust], ro["num_iters"][robust])


def test_wtc_badly_scaled_columns_and_overflow_redo(ctx):
    """Columns of A spread over 2^±8 (per-column power-of-two scales keep every column at full FP16-split precision: the error
    is relative to each column's own magnitude) and rows that grow by 2^20 after the first 32 (the first-pass scale estimate
    overflows FP16: the pass is detected and repeated with smaller scales)."""
    B, m, n = 64, 320, 36
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32, p0=9)
    scale = np.exp2(np.linspace(-8, 8, n)).astype(np.float32)
    A1 = A * scale                      # x_j scales by 1 / scale_j: the fit itself is unchanged
    x01 = x0 / scale
    xo, ro, robust = oracle_run(A1, y, x01, FLOAT_OPTS)
    xg, rg = gpu_run(ctx, A1, y, x01, FLOAT_OPTS)
    same = rg["num_iters"] == ro["num_iters"]
    assert np.array_equal(rg["num_iters"][robust], ro["num_iters"][robust]), (robust.mean(), same.mean())
    assert same.mean() >= 0.5 and (rg["stop_reason"] > 0).all()
    colrel = np.abs((xg - xo) * scale).max(axis=1) / np.abs(xo * scale).max(axis=1)  # in the units of the unscaled fit
    assert colrel[same].max() <= 1e-4, colrel[same].max()
    assert (np.abs(rg["final_cost"] - ro["final_cost"]) / ro["final_cost"]).max() <= 1e-4
    A2 = A.copy(); A2[:, 32:, :] *= np.float32(2.0 ** 20); y2 = y.copy(); y2[:, 32:] *= np.float32(2.0 ** 20)
    xo, ro, robust = oracle_run(A2, y2, x0, FLOAT_OPTS)
    xg, rg = gpu_run(ctx, A2, y2, x0, FLOAT_OPTS)
    assert np.isfinite(xg).all() and (rg["stop_reason"] > 0).all()
    same = rg["num_iters"] == ro["num_iters"]
    assert np.array_equal(rg["num_iters"][robust], ro["num_iters"][robust]), (robust.mean(), same.mean())
    assert rel_err_rows(xg[same], xo[same]).max() <= 1e-4
    assert (np.abs(rg["final_cost"] - ro["final_cost"]) / ro["final_cost"]).max() <= 1e-3


def test_wtc_degenerate_systems_take_the_exact_route(ctx):
    """A zero column (zero pivot: Eigen's D⁺), two equal columns (rank deficient JᵀJ, solved through the damping), a NaN and
    an Inf in A: the outcomes of the exact kernel, problem by problem."""
    B, m, n = 48, 256, 30
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32, p0=31)
    A[0, :, 7] = 0.0
    A[1, :, 9] = A[1, :, 3]
    A[2, 17, 5] = np.nan
    A[3, 164, 2] = np.inf
    A[4] = 0.0
    xe, re_ = gpu_run(ctx, A, y, x0, FLOAT_OPTS, exact=True)
    xg, rg = gpu_run(ctx, A, y, x0, FLOAT_OPTS)
    # NaN, Inf, the all-zero problem: the same verdict, the same Step
    assert np.array_equal(rg["stop_reason"][2:5], re_["stop_reason"][2:5]), (rg["stop_reason"][:5], re_["stop_reason"][:5])
    assert np.array_equal(rg["num_iters"][2:5], re_["num_iters"][2:5])
    # the zero column / the duplicated column: both kernels converge (which of the two last-Step stop tests fires is FP32
    # noise, and x is not unique along the null direction) to the same cost; the null coordinate of problem 0 never moves
    assert (rg["stop_reason"][:2] > 0).all() and (re_["stop_reason"][:2] > 0).all()
    assert np.abs(rg["num_iters"][:2].astype(int) - re_["num_iters"][:2]).max() <= 1
    assert (np.abs(rg["final_cost"][:2] - re_["final_cost"][:2]) / re_["final_cost"][:2]).max() <= 1e-4
    assert xg[0, 7] == x0[0, 7] == xe[0, 7]
    fin = np.isfinite(xe).all(axis=1)
    assert np.array_equal(fin, np.isfinite(xg).all(axis=1))
    ok = fin.copy(); ok[:2] = False
    assert rel_err_rows(xg[ok], xe[ok]).max() <= 2e-4
    assert np.array_equal(rg["num_iters"][5:], re_["num_iters"][5:])


def test_shapes_outside_the_tensor_core_path_run_the_exact_kernel(ctx):
    """(m n) % 4 != 0, short problems (m < 192), n < 28, double precision and `use_ldlt = 0` never reach wtc.cuh: bit-identical
    to the oracle."""
    import tinyopt_b200 as tb
    for (B, m, n, kw) in [(20, 225, 31, {}), (20, 100, 50, {}), (20, 300, 20, {}), (20, 256, 40, dict(use_ldlt=0))]:
        A, y, xs, x0 = O.synth_generate(B, m, n, np.float32, p0=3)
        kw = {**FLOAT_OPTS, **kw}
        xo, ro, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw))
        xg, rg = gpu_run(ctx, A, y, x0, kw)
        assert np.array_equal(xg, xo) and np.array_equal(rg["num_iters"], ro["num_iters"]), (B, m, n, kw)
    A, y, xs, x0 = O.synth_generate(16, 128, 20, np.float64, p0=3)
    xo, ro, _ = O.synth_lm_run(A, y, x0, O.default_options())
    out = ctx.optimize_batch(torch.from_numpy(A).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(x0).cuda(), tb.options(), layout=tb.PROBLEM_MAJOR)
    assert np.array_equal(out.x.cpu().numpy(), xo)


def test_wtc_full_c4_shard_properties(ctx):
    """A 65 536-problem shard of C4 on the device (6.5 GB): whole-batch properties that need no oracle — every problem
    converges for a legal reason in 3..5 Steps, the final cost sits at the noise level of the generator (sigma = 1e-2:
    cost ≈ m sigma² (1 - n / m)), the solution is within the fit's own uncertainty of x*, and the tensor-core kernel agrees
    with the exact kernel on the iteration count of ≥ 99.9 % of the problems and on x to 1e-4 for all of them."""
    import tinyopt_b200 as tb
    B, m, n = 65536, 500, 50
    A, y, xs, x0 = ctx.synth_generate(B, m, n, torch.float32, layout=tb.PROBLEM_MAJOR)
    opt = tb.options(**FLOAT_OPTS)
    out = ctx.optimize_batch(A, y, x0, opt, layout=tb.PROBLEM_MAJOR)
    r = out.results
    assert (r["stop_reason"] > 0).all() and r["num_iters"].min() >= 3 and r["num_iters"].max() <= 6
    expect = m * 1e-4 * (1 - n / m)
    assert abs(np.median(r["final_cost"]) / expect - 1) < 0.05
    err = (out.x - xs).abs().max(dim=1).values.cpu().numpy()
    assert np.quantile(err, 0.999) < 0.04 and np.median(err) < 0.02
    ctx.set_exact(True)
    ref = ctx.optimize_batch(A, y, x0, opt, layout=tb.PROBLEM_MAJOR)
    ctx.set_exact(False)
    assert (ref.results["num_iters"] == r["num_iters"]).mean() >= 0.999
    assert rel_err_rows(out.x.cpu().numpy(), ref.x.cpu().numpy()).max() <= 1e-4


# ---- tob200_build_solve_f32 on the tensor-core kernel (materialised J, r: the SolverType seam's Build + Solve) ----------
def oracle_build_solve_batch(J, r, lam):
    B, m, n = J.shape
    dx = np.zeros((B, n), J.dtype); cost = np.zeros(B); st = np.zeros(B, np.int32)
    H = np.zeros((B, n, n), J.dtype); g = np.zeros((B, n), J.dtype)
    for p in range(B):
        o = O.build_solve(J[p], r[p], float(lam[p]))
        st[p] = o["status"]; cost[p] = o["cost"]; H[p] = o["H"]; g[p] = o["g"]
        if o["status"] == 0:
            dx[p] = o["dx"]
    return dx, cost, st, H, g


@pytest.mark.parametrize("B,m,n", [(64, 500, 50), (9, 231, 52), (40, 260, 55), (33, 204, 33), (50, 192, 28), (30, 300, 51)])
def test_wtc_build_solve_parity(ctx, B, m, n):
    import tinyopt_b200 as tb
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32, p0=11)
    r, J = O.synth_eval(A, y, x0)
    lam = np.full(B, np.float32(1e-4), np.float32)
    lam[::3] = 0
    if B > 8:
        J[5, :, 3] = 0.0      # a zero column: zero pivot, Eigen's D+ (the exact route, after one repeated pass)
        J[7] = 0.0            # the all-zero system
    dx, cost, st, H, g = oracle_build_solve_batch(J, r, lam)
    for want in (False, True):
        out = ctx.build_solve(torch.from_numpy(J).cuda(), torch.from_numpy(r).cuda(), torch.from_numpy(lam).cuda(), layout=tb.PROBLEM_MAJOR,
                              want_H=want, want_g=want)
        ctx.sync()
        assert np.array_equal(out["status"].cpu().numpy(), st), (B, m, n, out["status"].cpu().numpy(), st)
        ok = st == 0
        assert rel_err_rows(out["dx"].cpu().numpy()[ok], dx[ok]).max() <= 1e-4, (B, m, n)
        assert (np.abs(out["cost"].cpu().numpy() - cost) / np.maximum(cost, 1e-30)).max() <= 1e-5
        if want:
            assert rel_err_rows(out["g"].cpu().numpy(), g).max() <= 1e-5 or np.abs(g).max() == 0
            assert rel_err_rows(out["H"].cpu().numpy(), H).max() <= 1e-5


@pytest.mark.parametrize("n", list(range(28, 56)))
def test_wtc_every_size(ctx, n):
    """Every n of the family (pitch of the LDLᵀ matrix, last step of 1..4 columns, one or two drain warps per side, odd n)."""
    B, m = 24, 256 if (256 * n) % 4 == 0 else 260
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32, p0=n)
    xo, ro, robust = oracle_run(A, y, x0, FLOAT_OPTS)
    xg, rg = gpu_run(ctx, A, y, x0, FLOAT_OPTS)
    check(xg, rg, xo, ro, robust)
    r, J = O.synth_eval(A, y, x0)
    lam = np.full(B, np.float32(1e-4), np.float32)
    dx, cost, st, H, g = oracle_build_solve_batch(J, r, lam)
    out = ctx.build_solve(torch.from_numpy(J).cuda(), torch.from_numpy(r).cuda(), torch.from_numpy(lam).cuda(), want_H=True, want_g=True)
    ctx.sync()
    assert np.array_equal(out["status"].cpu().numpy(), st)
    assert rel_err_rows(out["dx"].cpu().numpy(), dx).max() <= 1e-4
    assert rel_err_rows(out["H"].cpu().numpy(), H).max() <= 1e-5 and rel_err_rows(out["g"].cpu().numpy(), g).max() <= 1e-5
