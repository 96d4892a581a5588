"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on the same
seeded inputs.  Run on the B200 box with `pytest -m gpu`.

Bars (BASELINE.json north_star): dx within 1e-10 relative (double) / 1e-4 (float), iteration
counts identical.  The thread-per-problem kernels implement the oracle's canonical op sequence
(DESIGN.md §4), so these tests additionally assert BIT equality, which is stronger.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-10, np.float32: 1e-4}
TDT = {np.float64: torch.float64, np.float32: torch.float32}

# float configs run with float-appropriate thresholds (SURVEY.md §7 hard part 2b / §8d)
FLOAT_OPTS = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)


@pytest.fixture(scope="module")
def ctx():
    import tinyopt_b200 as tb
    c = tb.Context(0)
    yield c
    c.close()


def both_options(dtype, **kw):
    import tinyopt_b200 as tb
    if dtype == np.float32:
        kw = {**FLOAT_OPTS, **kw}
    return O.default_options(**kw), tb.options(**kw)


def rel_err(a, b):
    """Worst PER-PROBLEM relative error: leading axis = problems, each problem's max |a - b| over its own
    entries divided by its own max |b| (a batch-global ratio would let a problem with a small solution be
    off by far more than the bar)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    if a.ndim == 0:
        return float(np.abs(a - b) / max(np.abs(b), 1e-300))
    a = a.reshape(a.shape[0], -1); b = b.reshape(b.shape[0], -1)
    if a.shape[0] == 0:
        return 0.0
    return float((np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-300)).max())


# ---- synthetic generator: device == oracle, bit for bit ------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("B,m,n", [(1, 1, 1), (33, 30, 6), (64, 7, 12), (100, 13, 5)])
def test_synth_generate_bitexact(ctx, dtype, B, m, n):
    import tinyopt_b200 as tb
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=17)
    for layout in (tb.TILE32, tb.PROBLEM_MAJOR):
        dA, dy, dxs, dx0 = ctx.synth_generate(B, m, n, TDT[dtype], p0=17, layout=layout)
        if layout == tb.TILE32:
            dA, dy = tb.from_tile32(dA, B), tb.from_tile32(dy, B)
        assert np.array_equal(dA.cpu().numpy(), A)
        assert np.array_equal(dy.cpu().numpy(), y)
        assert np.array_equal(dxs.cpu().numpy(), xs)
        assert np.array_equal(dx0.cpu().numpy(), x0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_retile_roundtrip(ctx, dtype):
    import tinyopt_b200 as tb
    g = torch.Generator().manual_seed(0)
    a = torch.rand((70, 9, 5), generator=g, dtype=TDT[dtype]).cuda()
    t = ctx.retile(a)
    assert torch.equal(t, tb.to_tile32(a))
    assert torch.equal(tb.from_tile32(t, 70), a)
    r = torch.rand((70, 9), generator=g, dtype=TDT[dtype]).cuda()
    assert torch.equal(ctx.retile(r), tb.to_tile32(r))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_synth_eval_bitexact(ctx, dtype):
    import tinyopt_b200 as tb
    B, m, n = 70, 11, 4
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype)
    r, J = O.synth_eval(A, y, x0)
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, TDT[dtype], layout=tb.TILE32)
    dr, dJ = ctx.synth_eval(dA, dy, dx0, layout=tb.TILE32)
    assert np.array_equal(tb.from_tile32(dr, B).cpu().numpy(), r)
    assert np.array_equal(tb.from_tile32(dJ, B).cpu().numpy(), J)


# ---- a1+a3+a5+a6: one Build+Solve -----------------------------------------------------------------
SHAPES64 = [(1, 1, 1), (37, 5, 2), (96, 30, 6), (1000, 30, 6), (65, 40, 8), (50, 9, 3), (40, 3, 7)]
SHAPES32 = SHAPES64 + [(500, 200, 12), (70, 50, 9), (33, 25, 11), (64, 64, 10)]


def oracle_build_solve_batch(J, r, lam):
    B, m, n = J.shape
    dx = np.zeros((B, n), J.dtype); cost = np.zeros(B); st = np.zeros(B, np.int32)
    H = np.zeros((B, n, n), J.dtype); g = np.zeros((B, n), J.dtype)
    for b in range(B):
        o = O.build_solve(J[b], r[b], float(lam[b]))
        dx[b], cost[b], st[b], H[b], g[b] = o["dx"], o["cost"], o["status"], o["H"], o["g"]
    return dx, cost, st, H, g


@pytest.mark.parametrize("dtype,shapes", [(np.float64, SHAPES64), (np.float32, SHAPES32)])
def test_build_solve_parity(ctx, dtype, shapes):
    import tinyopt_b200 as tb
    for (B, m, n) in shapes:
        A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=5)
        r, J = O.synth_eval(A, y, x0)
        lam = np.full(B, np.float32(1e-4), dtype)
        lam[::3] = 0  # Gauss-Newton rows
        lam[1::7] = dtype(0.5)
        dx, cost, st, H, g = oracle_build_solve_batch(J, r, lam)
        for layout in (tb.PROBLEM_MAJOR, tb.TILE32):
            Jd, rd = torch.from_numpy(J).cuda(), torch.from_numpy(r).cuda()
            if layout == tb.TILE32:
                Jd, rd = tb.to_tile32(Jd), tb.to_tile32(rd)
            out = ctx.build_solve(Jd, rd, torch.from_numpy(lam).cuda(), B=B, layout=layout, want_H=True, want_g=True)
            ctx.sync()
            assert np.array_equal(out["status"].cpu().numpy(), st), (B, m, n)
            # the stated bar
            assert rel_err(out["dx"].cpu().numpy(), dx) <= TOL[dtype], (B, m, n)
            assert rel_err(out["cost"].cpu().numpy(), cost) <= TOL[dtype]
            # the stronger property of this kernel family: same op sequence -> same bits
            assert np.array_equal(out["dx"].cpu().numpy(), dx), (B, m, n, layout)
            assert np.array_equal(out["cost"].cpu().numpy(), cost)
            assert np.array_equal(out["g"].cpu().numpy(), g)
            assert np.array_equal(out["H"].cpu().numpy(), H)


def test_build_solve_edge_cases(ctx):
    """zero J (H == 0: Eigen's all-zero-diagonal branch -> success, dx = 0), NaN in J (-> status 1),
    rank-deficient J with and without damping (tests/types.cpp:94-108 situation)."""
    import tinyopt_b200 as tb
    B, m, n = 4, 6, 6
    J = np.zeros((B, m, n)); r = np.ones((B, m))
    Jt = np.array([[1, 0, 1, 0, 1, 0], [0, 1, 0, 1, 0, 1]], float)
    J[1, :2] = Jt                      # rank 2, undamped: PSD-singular
    J[2, :2] = Jt                      # rank 2, damped
    J[3, 0, 0] = np.nan
    lam = np.array([0, 0, 1e-4, 0.0])
    dx, cost, st, H, g = oracle_build_solve_batch(J, r, lam)
    out = ctx.build_solve(torch.from_numpy(J).cuda(), torch.from_numpy(r).cuda(), torch.from_numpy(lam).cuda())
    ctx.sync()
    assert np.array_equal(out["status"].cpu().numpy(), st)
    assert list(st) == [0, 0, 0, 1]
    ok = st == 0
    assert np.array_equal(out["dx"].cpu().numpy()[ok], dx[ok])
    assert np.all(dx[0] == 0)


# ---- a7-a10: the whole LM loop ---------------------------------------------------------------------
def run_both(ctx, dtype, B, m, n, layout=None, p0=0, **optkw):
    import tinyopt_b200 as tb
    layout = tb.TILE32 if layout is None else layout
    oo, go = both_options(dtype, **optkw)
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=p0)
    xo, ro, _ = O.synth_lm_run(A, y, x0, oo)
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, TDT[dtype], p0=p0, layout=layout)
    out = ctx.optimize_batch(dA, dy, dx0, go, layout=layout)
    return xo, ro, out


def assert_lm_parity(dtype, xo, ro, out, exact=True):
    rg = out.results
    assert np.array_equal(rg["num_iters"], ro["num_iters"])          # iteration counts identical
    assert np.array_equal(rg["stop_reason"], ro["stop_reason"])
    assert np.array_equal(rg["num_failures"], ro["num_failures"])
    xg = out.x.cpu().numpy()
    assert rel_err(xg, xo) <= TOL[dtype]
    assert rel_err(rg["final_cost"], ro["final_cost"]) <= TOL[dtype]
    if exact:
        assert np.array_equal(xg, xo)
        assert np.array_equal(rg["final_cost"], ro["final_cost"])
        assert np.array_equal(rg["last_lambda"], ro["last_lambda"])
        assert np.array_equal(rg["final_rerr_dec"], ro["final_rerr_dec"])


@pytest.mark.parametrize("B,m,n", [(2000, 30, 6), (33, 30, 6), (1, 4, 1), (257, 12, 3), (100, 20, 8), (64, 9, 7)])
def test_lm_run_parity_f64(ctx, B, m, n):
    """config C2 shape (n=6, m=30, double, tinyopt default options) and neighbours."""
    xo, ro, out = run_both(ctx, np.float64, B, m, n)
    assert_lm_parity(np.float64, xo, ro, out)
    assert (ro["stop_reason"] > 0).all()


@pytest.mark.parametrize("B,m,n", [(1000, 200, 12), (45, 200, 12), (300, 50, 10), (128, 30, 6), (77, 17, 11), (10, 5, 1)])
def test_lm_run_parity_f32(ctx, B, m, n):
    """config C3 shape (n=12, m=200, float, float-tuned thresholds) and neighbours."""
    xo, ro, out = run_both(ctx, np.float32, B, m, n)
    assert_lm_parity(np.float32, xo, ro, out)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_lm_run_default_options_stress(ctx, dtype):
    """float with tinyopt's *default* (double-scaled) thresholds runs into the fp32 noise floor:
    rejects, roll-backs, eval-only passes with a stale re-damped H.  Because the op
    sequence is identical this must still agree exactly."""
    import tinyopt_b200 as tb
    B, m, n = 512, 40, 6
    oo, go = O.default_options(), tb.options()
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=99)
    xo, ro, _ = O.synth_lm_run(A, y, x0, oo)
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, TDT[dtype], p0=99)
    out = ctx.optimize_batch(dA, dy, dx0, go)
    assert_lm_parity(dtype, xo, ro, out)
    if dtype == np.float32:
        assert (out.results["num_builds"] < out.results["num_iters"]).any()  # eval-only passes happened


@pytest.mark.parametrize("optkw", [
    dict(solver_type=1),                                   # Gauss-Newton
    dict(max_iters=2),                                     # kMaxIters
    dict(downscale_by_2=1, normalize=1),                   # cost scaling (base.h:41-45)
    dict(use_squared_norm=0),
    dict(damping_init=10.0, max_consec_failures=2),
    dict(grad_clipping=0.05),
    dict(check_min_H_diag=1e3),                            # Build fails -> kSolverFailed
    dict(min_error=1.0),                                   # kMinError at once
    dict(max_total_failures=1, min_step_norm2=0, min_rerr_dec=0, min_error=0, min_grad_norm2=0),
    dict(check_final_cost=1, max_iters=3),
    dict(use_step_quality_approx=1),
])
def test_lm_run_option_variants(ctx, optkw):
    xo, ro, out = run_both(ctx, np.float64, 200, 30, 6, **optkw)
    # use_step_quality_approx goes through pow(): tolerance-level only (documented)
    assert_lm_parity(np.float64, xo, ro, out, exact="use_step_quality_approx" not in optkw)


def test_lm_run_layouts_agree(ctx):
    import tinyopt_b200 as tb
    xo, ro, out_t = run_both(ctx, np.float64, 100, 30, 6, layout=tb.TILE32)
    _, _, out_p = run_both(ctx, np.float64, 100, 30, 6, layout=tb.PROBLEM_MAJOR)
    assert torch.equal(out_t.x, out_p.x)
    assert np.array_equal(out_t.results, out_p.results)


def test_lm_run_no_residuals_and_errors(ctx):
    """m == 0 -> kSkipped (tests/basic.cpp:234-242); unsupported n -> error code, never a fallback."""
    import tinyopt_b200 as tb
    B, n = 40, 3
    x0 = torch.ones((B, n), dtype=torch.float64, device="cuda")
    res = torch.empty((B, 56), dtype=torch.uint8, device="cuda")
    # a valid non-null pointer is required even for m == 0
    dummy = torch.zeros(64, dtype=torch.float64, device="cuda")
    import ctypes as C
    rc = ctx._lib.tob200_lm_run_f64(ctx._h, C.byref(tb.options()), C.c_void_p(dummy.data_ptr()),
                                    C.c_void_p(dummy.data_ptr()), C.c_double(0.1), tb.TILE32, B, 0, n,
                                    C.c_void_p(x0.data_ptr()), C.c_void_p(res.data_ptr()))
    assert rc == 0
    ctx.sync()
    from tinyopt_b200.api import decode_results
    r = decode_results(res)
    assert (r["stop_reason"] == tb.StopReason.kSkipped).all() and (r["num_iters"] == 1).all()
    assert torch.equal(x0, torch.ones_like(x0))
    with pytest.raises(tb.TinyoptB200Error):
        ctx.optimize_batch(torch.zeros((1, 4, 2100, 32), device="cuda"), torch.zeros((1, 4, 32), device="cuda"),
                           torch.zeros((3, 2100), device="cuda"))   # n > 2048: no kernel family


# ---- BASELINE.json config sizes: exact on a slice, properties on the whole batch -------------------
@pytest.mark.parametrize("dtype,B,m,n", [(np.float64, 100_000, 30, 6), (np.float32, 100_000, 200, 12)])
def test_full_size_configs(ctx, dtype, B, m, n):
    """configs[1] (C2) and configs[2] (C3) at full size."""
    import tinyopt_b200 as tb
    oo, go = both_options(dtype)
    dA, dy, dxs, dx0 = ctx.synth_generate(B, m, n, TDT[dtype])
    out = ctx.optimize_batch(dA, dy, dx0, go)
    r = out.results
    assert out.Succeeded().all() and out.Converged().all()
    # (1) the first and last 2048 problems against the oracle, exactly
    for lo in (0, B - 2048):
        A, y, xs, x0 = O.synth_generate(2048, m, n, dtype, p0=lo)
        xo, ro, _ = O.synth_lm_run(A, y, x0, oo)
        sl = slice(lo, lo + 2048)
        assert np.array_equal(r["num_iters"][sl], ro["num_iters"])
        assert np.array_equal(r["stop_reason"][sl], ro["stop_reason"])
        assert np.array_equal(out.x[sl].cpu().numpy(), xo)
    # (2) size-independent properties on all B problems
    x = out.x
    err0 = (dx0 - dxs).norm(dim=1)
    err1 = (x - dxs).norm(dim=1)
    assert (err1 < err0).float().mean() > 0.999            # moved towards the truth
    assert err1.max().item() < 0.1                          # noise-limited accuracy (sigma = 1e-2)
    # first-order optimality: g = J^T r ~ 0 at the solution
    rr, JJ = ctx.synth_eval(dA, dy, x)
    bs = ctx.build_solve(JJ, rr, None, B=B, layout=tb.TILE32, want_g=True)
    ctx.sync()
    gn = bs["g"].norm(dim=1)
    assert gn.max().item() < (1e-6 if dtype == np.float64 else 2e-3)
    # cost reported == cost recomputed from the returned x (the lag of optimizer.h:428: final_cost is
    # the cost at the last accepted x *before* the final step; the recomputed one can only be lower)
    assert (bs["cost"].cpu().numpy() <= r["final_cost"] * (1 + (1e-9 if dtype == np.float64 else 1e-4))).all()
    # idempotence: restarting from the solution stops at once with a tiny step
    out2 = ctx.optimize_batch(dA, dy, x, go)
    assert (out2.results["num_iters"] <= 3).all()
    assert (out2.x - x).abs().max().item() < (1e-6 if dtype == np.float64 else 1e-3)
    # total work
    assert r["num_iters"].sum() == r["num_iters"].astype(np.int64).sum()
    assert 2 <= r["num_iters"].min() and r["num_iters"].max() <= 12


# ---- warp-per-problem family (13 <= n <= 55, float; config C4 shape n = 50, m = 500) ---------------
WPP_SHAPES = [(40, 500, 50), (33, 37, 13), (17, 64, 20), (9, 100, 27), (21, 96, 28), (12, 45, 32),
              (10, 130, 40), (7, 33, 47), (6, 70, 55), (5, 7, 13), (3, 1, 29)]


def test_wpp_kernel_family(ctx):
    assert ctx.kernel_family(torch.float32, 12) == 1 and ctx.kernel_family(torch.float32, 13) == 2
    assert ctx.kernel_family(torch.float32, 55) == 2 and ctx.kernel_family(torch.float32, 56) == 3
    assert ctx.kernel_family(torch.float32, 57) == 3 and ctx.kernel_family(torch.float32, 512) == 3
    assert ctx.kernel_family(torch.float32, 513) == 4 and ctx.kernel_family(torch.float32, 2048) == 4
    assert ctx.kernel_family(torch.float32, 2049) == 0
    assert ctx.kernel_family(torch.float64, 8) == 1 and ctx.kernel_family(torch.float64, 9) == 2
    assert ctx.kernel_family(torch.float64, 55) == 2 and ctx.kernel_family(torch.float64, 56) == 4


WPP_SHAPES_F64 = [(64, 60, 9), (50, 37, 12), (40, 64, 13), (33, 90, 27), (35, 100, 28), (24, 200, 50), (9, 131, 55)]


@pytest.mark.parametrize("B,m,n", WPP_SHAPES_F64)
def test_wpp_double_lm_run_parity(ctx, B, m, n):
    """double for 9 <= n <= 55 (the warp-per-problem family's scalar paths): the same canonical op
    sequence as the oracle -> bit-exact x, costs, lambda, iteration counts under tinyopt's defaults."""
    import tinyopt_b200 as tb
    for layout in (tb.PROBLEM_MAJOR, tb.TILE32):
        xo, ro, out = run_both(ctx, np.float64, B, m, n, layout=layout)
        assert_lm_parity(np.float64, xo, ro, out)
    assert (ro["stop_reason"] > 0).all()


@pytest.mark.parametrize("B,m,n", [(30, 40, 10), (21, 77, 30), (12, 120, 50)])
def test_wpp_double_build_solve_parity(ctx, B, m, n):
    import tinyopt_b200 as tb
    dtype = np.float64
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=3)
    r, J = O.synth_eval(A, y, x0)
    lam = np.full(B, 1e-4, dtype)
    lam[::3] = 0
    dx, cost, st, H, g = oracle_build_solve_batch(J, r, lam)
    out = ctx.build_solve(torch.from_numpy(J).cuda(), torch.from_numpy(r).cuda(), torch.from_numpy(lam).cuda(), B=B,
                          layout=tb.PROBLEM_MAJOR, want_H=True, want_g=True)
    ctx.sync()
    assert np.array_equal(out["status"].cpu().numpy(), st)
    assert np.array_equal(out["cost"].cpu().numpy(), cost)
    assert np.array_equal(out["g"].cpu().numpy(), g)
    assert np.array_equal(out["H"].cpu().numpy(), H)
    assert np.array_equal(out["dx"].cpu().numpy(), dx)
    assert rel_err(out["dx"].cpu().numpy(), dx) <= 1e-10


def test_wpp_double_degenerate(ctx):
    """ties on the diagonal (Eigen's first-maximum order, the serial replay), a zero matrix and an indefinite
    one through the double warp kernels."""
    import tinyopt_b200 as tb
    n, m, B = 20, 25, 4
    rng = np.random.default_rng(3)
    J = rng.standard_normal((B, m, n))
    J[0] = 0; J[0, :n, :] = np.eye(n) * 2.0          # H = 4 I: every diagonal entry ties
    J[1] = 0                                         # zero matrix: ZeroSign, accepted, dx = 0
    J[2, :, 5] = J[2, :, 7]                          # two identical columns: singular, tie at (5, 7)
    r = rng.standard_normal((B, m))
    lam = np.zeros(B)
    dx, cost, st, H, g = oracle_build_solve_batch(J, r, lam)
    out = ctx.build_solve(torch.from_numpy(J).cuda(), torch.from_numpy(r).cuda(), torch.from_numpy(lam).cuda(), B=B,
                          layout=tb.PROBLEM_MAJOR, want_H=True, want_g=True)
    ctx.sync()
    assert np.array_equal(out["status"].cpu().numpy(), st)
    ok = st == 0
    assert np.array_equal(out["dx"].cpu().numpy()[ok], dx[ok], equal_nan=True)
    assert np.array_equal(out["H"].cpu().numpy(), H)


@pytest.mark.parametrize("B,m,n", WPP_SHAPES)
def test_wpp_build_solve_parity(ctx, B, m, n):
    import tinyopt_b200 as tb
    dtype = np.float32
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=11)
    r, J = O.synth_eval(A, y, x0)
    lam = np.full(B, np.float32(1e-4), dtype)
    lam[::3] = 0
    dx, cost, st, H, g = oracle_build_solve_batch(J, r, lam)
    for layout in (tb.PROBLEM_MAJOR, tb.TILE32):
        Jd, rd = torch.from_numpy(J).cuda(), torch.from_numpy(r).cuda()
        if layout == tb.TILE32:
            Jd, rd = tb.to_tile32(Jd), tb.to_tile32(rd)
        out = ctx.build_solve(Jd, rd, torch.from_numpy(lam).cuda(), B=B, layout=layout, want_H=True, want_g=True)
        ctx.sync()
        assert np.array_equal(out["status"].cpu().numpy(), st), (B, m, n)
        ok = st == 0
        assert rel_err(out["dx"].cpu().numpy()[ok], dx[ok]) <= TOL[dtype], (B, m, n)
        assert rel_err(out["cost"].cpu().numpy(), cost) <= TOL[dtype]
        # same canonical op sequence as the oracle -> same bits
        assert np.array_equal(out["cost"].cpu().numpy(), cost)
        assert np.array_equal(out["g"].cpu().numpy(), g)
        assert np.array_equal(out["H"].cpu().numpy(), H)
        assert np.array_equal(out["dx"].cpu().numpy()[ok], dx[ok]), (B, m, n, layout)


@pytest.mark.parametrize("B,m,n", [(300, 500, 50), (40, 64, 13), (33, 90, 27), (35, 100, 28), (20, 77, 40), (9, 131, 55)])
def test_wpp_lm_run_parity(ctx, B, m, n):
    """config C4 shape (n=50, m=500, float, float-tuned thresholds) and the family's corners;
    m*n % 4 != 0 cases take the non-TMA loader."""
    import tinyopt_b200 as tb
    xo, ro, out = run_both(ctx, np.float32, B, m, n, layout=tb.PROBLEM_MAJOR)
    assert_lm_parity(np.float32, xo, ro, out)
    assert (ro["stop_reason"] > 0).all()


@pytest.mark.parametrize("optkw", [dict(min_rerr_dec=1e-10, min_step_norm2=1e-14),  # tinyopt defaults: fp32 noise floor
                                   dict(solver_type=1), dict(max_iters=2), dict(damping_init=10.0, max_consec_failures=2),
                                   dict(check_min_H_diag=1e3), dict(grad_clipping=0.05)])
def test_wpp_lm_run_option_variants(ctx, optkw):
    import tinyopt_b200 as tb
    xo, ro, out = run_both(ctx, np.float32, 96, 120, 30, layout=tb.PROBLEM_MAJOR, **optkw)
    assert_lm_parity(np.float32, xo, ro, out)
    if "min_rerr_dec" in optkw:
        assert (out.results["num_builds"] < out.results["num_iters"]).any()  # cost-only passes happened


def test_wpp_layouts_agree(ctx):
    import tinyopt_b200 as tb
    _, _, out_p = run_both(ctx, np.float32, 70, 60, 33, layout=tb.PROBLEM_MAJOR)
    _, _, out_t = run_both(ctx, np.float32, 70, 60, 33, layout=tb.TILE32)
    assert torch.equal(out_t.x, out_p.x) and np.array_equal(out_t.results, out_p.results)


# ---- seeded sweep over ragged shapes: every family below n = 56, both precisions ---------------------
def _random_cases(seed, count):
    rng = np.random.default_rng(seed)
    cases = []
    for _ in range(count):
        dtype = np.float64 if rng.random() < 0.5 else np.float32
        n = int(rng.integers(1, 56))
        m = int(rng.integers(max(1, n // 2), 3 * n + 8))   # under- and over-determined, odd sizes, m % 4 != 0
        B = int(rng.integers(1, 70))
        layout = int(rng.integers(0, 2))
        cases.append((dtype, B, m, n, layout))
    return cases


@pytest.mark.parametrize("dtype,B,m,n,layout", _random_cases(20261017, 48))
def test_random_shapes_bitexact(ctx, dtype, B, m, n, layout):
    """Ragged shapes through tob200_lm_run_* (thread- and warp-per-problem families, TMA and non-TMA
    loaders, both layouts, rank-deficient m < n systems that only the LM damping makes solvable): bit-exact
    x, costs, lambdas and iteration counts against the oracle."""
    import tinyopt_b200 as tb
    xo, ro, out = run_both(ctx, dtype, B, m, n, layout=tb.TILE32 if layout == 0 else tb.PROBLEM_MAJOR, p0=n * 1000 + m)
    assert_lm_parity(dtype, xo, ro, out)


def test_shape_sequence_keeps_shared_memory_limit(ctx):
    """Regression: one kernel instantiation serves shapes with different shared-memory footprints; a small
    shape must not lower the limit under a larger one whose launch geometry is cached."""
    import tinyopt_b200 as tb
    for dtype, seq in ((np.float32, [(20, 82, 39), (20, 64, 36), (20, 82, 39), (9, 300, 12), (9, 5, 12), (9, 300, 12)]),
                       (np.float64, [(12, 66, 26), (12, 30, 25), (12, 66, 26)])):
        for (B, m, n) in seq:
            xo, ro, out = run_both(ctx, dtype, B, m, n, layout=tb.PROBLEM_MAJOR)
            assert_lm_parity(dtype, xo, ro, out)


@pytest.mark.parametrize("dtype,B,m,n,layout", _random_cases(777, 24))
def test_random_build_solve_bitexact(ctx, dtype, B, m, n, layout):
    """Ragged shapes through tob200_build_solve_* (Build + damping + LDLT alone): status, cost, g, damped H
    and dx bit-exact against the oracle, LM and Gauss-Newton rows mixed."""
    import tinyopt_b200 as tb
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=7 * n + m)
    r, J = O.synth_eval(A, y, x0)
    lam = np.full(B, 1e-4, dtype)
    lam[::3] = 0
    lam[1::5] = dtype(0.25)
    dx, cost, st, H, g = oracle_build_solve_batch(J, r, lam)
    Jd, rd = torch.from_numpy(J).cuda(), torch.from_numpy(r).cuda()
    lay = tb.PROBLEM_MAJOR if layout else tb.TILE32
    if lay == tb.TILE32:
        Jd, rd = tb.to_tile32(Jd), tb.to_tile32(rd)
    out = ctx.build_solve(Jd, rd, torch.from_numpy(lam).cuda(), B=B, layout=lay, want_H=True, want_g=True)
    ctx.sync()
    assert np.array_equal(out["status"].cpu().numpy(), st)
    ok = st == 0
    assert np.array_equal(out["cost"].cpu().numpy(), cost)
    assert np.array_equal(out["g"].cpu().numpy(), g)
    assert np.array_equal(out["H"].cpu().numpy(), H)
    assert np.array_equal(out["dx"].cpu().numpy()[ok], dx[ok])


@pytest.mark.parametrize("dtype,B,m,n,tiled", [(np.float64, 30007, 30, 6, True), (np.float64, 30007, 30, 6, False),
                                               (np.float32, 9001, 120, 20, False), (np.float32, 70, 40, 12, True)])
def test_host_entry_chunked_pipeline(ctx, dtype, B, m, n, tiled):
    """tob200_lm_run_host_* (host buffers; inputs above 32 MB go through the 4-chunk upload / solve / download
    pipeline, B is not a multiple of the chunk or tile size): identical to the device-resident run, and to
    the oracle on a slice."""
    import tinyopt_b200 as tb
    oo, go = both_options(dtype)
    layout = tb.TILE32 if tiled else tb.PROBLEM_MAJOR
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, TDT[dtype], layout=layout)
    out = ctx.optimize_batch(dA, dy, dx0, go, layout=layout)
    Ah, yh, xh = dA.cpu().numpy(), dy.cpu().numpy(), dx0.cpu().numpy().copy()
    res = ctx.optimize_batch_host(Ah, yh, xh, go, layout=layout, B=B)
    assert np.array_equal(xh, out.x.cpu().numpy())
    for k in ("num_iters", "stop_reason", "final_cost", "last_lambda", "num_builds"):
        assert np.array_equal(res[k], out.results[k]), k
    sl = slice(B - 40, B)   # the tail: the last, partial chunk / tile
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype)
    xo, ro, _ = O.synth_lm_run(A[sl], y[sl], x0[sl], oo)
    assert np.array_equal(xh[sl], xo) and np.array_equal(res["num_iters"][sl], ro["num_iters"])


# ---- Output::final_hessian from the device-resident loop (tob200_lm_run_ex_*) -------------------
@pytest.mark.parametrize("dtype,B,m,n,exact", [(np.float64, 70, 30, 6, True), (np.float32, 70, 60, 12, True),
                                               (np.float32, 40, 200, 50, True), (np.float64, 20, 64, 20, True),
                                               (np.float32, 40, 64, 20, True), (np.float32, 5, 256, 64, False),
                                               (np.float32, 4, 300, 57, False)])
def test_final_hessian_of_lm_run(ctx, dtype, B, m, n, exact):
    """optimizer.h:313-316 / solvers/lm.h:157-171: the last H_ with its damping removed (diagonal / (1 +
    prev_lambda_) in Scalar), as doubles.  Families 1 and 2 equal the oracle bit for bit; the tensor-core family
    (n >= 56, incl. the zero-padded n % 4 != 0 path) to FP32 accuracy.  hessian.save_last = false (the option the
    reference's benchmarks run with) leaves the buffer untouched."""
    import tinyopt_b200 as tb
    oo, go = both_options(dtype)
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype, p0=11)
    xo, ro, _, fho = O.synth_lm_run(A, y, x0, oo, want_hessian=True)
    layout = tb.TILE32 if ctx.kernel_family(TDT[dtype], n) == 1 else tb.PROBLEM_MAJOR
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, TDT[dtype], p0=11, layout=layout)
    out = ctx.optimize_batch(dA, dy, dx0, go, layout=layout, want_hessian=True)
    fh = out.final_hessian.cpu().numpy()
    assert np.array_equal(fh, np.swapaxes(fh, 1, 2))
    if exact:
        assert np.array_equal(out.x.cpu().numpy(), xo)
        assert np.array_equal(fh, fho)
    else:
        same = out.results["num_iters"] == ro["num_iters"]     # same number of Builds -> the same last H_
        assert same.any()
        scale = np.abs(fho).max(axis=(1, 2))
        assert (np.abs(fh - fho).max(axis=(1, 2)) / scale)[same].max() < 2e-5
    oo2, go2 = both_options(dtype, save_last=0)
    out2 = ctx.optimize_batch(dA, dy, dx0, go2, layout=layout, want_hessian=True)
    assert not out2.final_hessian.any().item()
    assert np.array_equal(out2.x.cpu().numpy(), out.x.cpu().numpy())
