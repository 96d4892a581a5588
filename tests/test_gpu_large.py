"""GPU parity tests of the large-n family (56 <= n <= 512, float; config C5): tensor-core JᵀJ
(tcgen05, 3xTF32), blocked LDLT, host-orchestrated LM loop — against the CPU oracle.

Bars: the LDLT solver alone is BIT exact (same operation order as the oracle); JᵀJ from the tensor
cores is FP32-level accurate (checked against float64); δx and x within 1e-4 relative for the full
path and iteration counts identical wherever the decisions clear FP32 noise (robust_decisions below;
BASELINE.json north_star, float).
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
    yield c
    c.close()


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


def spd(rng, n, cond=50.0):
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    d = np.geomspace(1.0, cond, n)
    return ((q * d) @ q.T).astype(np.float32)


# ---- a6 alone: SolveLDLT, bit exact against the oracle ---------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 5, 31, 32, 33, 64, 100, 257, 512])
def test_solve_ldlt_bitexact(ctx, n):
    rng = np.random.default_rng(n)
    B = 5
    A = np.stack([spd(rng, n) for _ in range(B)])
    A[1] *= 1e-3                                   # a different scale
    A[2][np.tril_indices(n, -1)] = 7.0             # only the upper triangle may be read
    b = rng.standard_normal((B, n)).astype(np.float32)
    x, st = ctx.solve_ldlt(torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda())
    ctx.sync()
    x, st = x.cpu().numpy(), st.cpu().numpy()
    for p in range(B):
        xo = O.solve_ldlt(A[p], b[p])
        assert xo is not None and st[p] == 0
        assert np.array_equal(x[p], xo), (n, p, np.abs(x[p] - xo).max())


def test_solve_ldlt_ties_and_rejections(ctx):
    """equal diagonal entries (Eigen's first-maximum-wins order), indefinite / negative matrices
    (isPositive() false -> status 1), the all-zero matrix (ZeroSign: success, x = 0), a PSD-singular
    matrix (zero pivot followed by zero column: success through the pseudo-inverse of D)."""
    rng = np.random.default_rng(7)
    n = 96
    mats = []
    a = spd(rng, n); np.fill_diagonal(a, 3.0); a += 10 * np.eye(n, dtype=np.float32); mats.append(a)   # all ties
    a = spd(rng, n); a[5, 5] = a[40, 40] = a[41, 41] = 77.0; mats.append(a)                            # some ties
    mats.append(-spd(rng, n))                                                                          # negative
    a = spd(rng, n); a[3, 3] = -5.0; mats.append(a)                                                    # indefinite
    mats.append(np.zeros((n, n), np.float32))                                                          # zero
    v = rng.standard_normal((n, 3)).astype(np.float32); mats.append((v @ v.T).astype(np.float32))      # rank 3
    a = np.zeros((n, n), np.float32); a[0, 1] = 1.0; mats.append(a)                                    # zero diag, nonzero off
    A = np.stack(mats)
    b = rng.standard_normal((len(mats), n)).astype(np.float32)
    x, st = ctx.solve_ldlt(torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda())
    ctx.sync()
    x, st = x.cpu().numpy(), st.cpu().numpy()
    for p in range(len(mats)):
        xo = O.solve_ldlt(A[p], b[p])
        assert (xo is None) == (st[p] == 1), p
        if xo is not None:
            assert np.array_equal(x[p], xo, equal_nan=True), p
    assert st[2] == 1 and st[3] == 1 and st[4] == 0 and st[6] == 1


# ---- a1 alone: JᵀJ on the tensor cores ---------------------------------------------------------------
@pytest.mark.parametrize("B,m,n", [(3, 8, 64), (2, 100, 128), (3, 37, 200), (2, 1000, 256), (2, 333, 320), (2, 2048, 512), (5, 64, 4), (2, 50, 60),
                                   (3, 40, 7), (2, 90, 61), (2, 700, 257), (4, 33, 1)])  # n % 4 != 0: padded copy
def test_jtj_matches_float64(ctx, B, m, n):
    rng = np.random.default_rng(m * n)
    J = rng.uniform(-1, 1, (B, m, n)).astype(np.float32)
    s = rng.uniform(0.5, 1.5, (B, m)).astype(np.float32)
    for scale in (None, s):
        H = ctx.jtj(torch.from_numpy(J).cuda(), None if scale is None else torch.from_numpy(scale).cuda())
        ctx.sync()
        H = H.cpu().numpy()
        Js = J.astype(np.float64) if scale is None else (J * scale[:, :, None]).astype(np.float32).astype(np.float64)
        ref = np.einsum("bmi,bmj->bij", Js, Js)
        assert np.array_equal(H, np.swapaxes(H, 1, 2))          # exactly symmetric (one triangle computed)
        scale_h = np.abs(ref).max()
        di = np.arange(n)
        derr = np.abs(H[:, di, di] - ref[:, di, di]).max() / scale_h
        off = np.abs(H - ref); off[:, di, di] = 0
        oerr = off.max() / scale_h
        # FP32-level: diagonal accumulated in FP32 (rows in order), off-diagonals 3xTF32 on the tensor
        # cores (their accumulator truncates, which only matters for long same-sign sums)
        assert derr < 5e-6 and oerr < 4e-6, (B, m, n, derr, oerr)


@pytest.mark.parametrize("B,m,n", [(2, 2048, 512), (7, 333, 400), (3, 100, 509), (149, 40, 512)])
def test_jtj_cluster_multicast_is_identical(ctx, B, m, n, monkeypatch):
    """TOB200_LG_MC=1: the J^T J kernel launched as clusters of two CTAs, the raw stages both units of a problem read
    multicast by the TMA unit (A leaves HBM / L2 once for the two wide strips).  Same operands, same MMAs, same order:
    the result must be BIT-identical to the default launch (ragged rows, padded n, more problems than clusters)."""
    import tinyopt_b200 as tb
    rng = np.random.default_rng(m + n)
    J = torch.from_numpy(rng.uniform(-1, 1, (B, m, n)).astype(np.float32)).cuda()
    s = torch.from_numpy(rng.uniform(0.5, 1.5, (B, m)).astype(np.float32)).cuda()
    H0 = ctx.jtj(J, s)
    ctx.sync()
    monkeypatch.setenv("TOB200_LG_MC", "1")
    c2 = tb.Context(0)
    try:
        H1 = c2.jtj(J, s)
        c2.sync()
        assert torch.equal(H0, H1)
    finally:
        c2.close()


@pytest.mark.parametrize("B,m,n", [(5, 200, 64), (3, 600, 257), (2, 1100, 512)])
def test_set_exact_2_runs_the_bit_exact_family(ctx, B, m, n):
    """tob200_set_exact(ctx, 2): float 56 <= n <= 512 through the general family - x, iteration counts, stop reasons,
    final costs, lambda and Build + Solve bit for bit against the oracle (the default tensor-core family is tolerance-held)."""
    import tinyopt_b200 as tb
    kw = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32, p0=9)
    xo, ro, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw))
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, torch.float32, p0=9, layout=tb.PROBLEM_MAJOR)
    ctx.set_exact(2)
    try:
        out = ctx.optimize_batch(dA, dy, dx0, tb.options(**kw))
        ctx.sync()
    finally:
        ctx.set_exact(True)   # what the suite runs with (conftest: TOB200_WPP_TC=0)
    assert np.array_equal(out.results["num_iters"], ro["num_iters"])
    assert np.array_equal(out.results["stop_reason"], ro["stop_reason"])
    assert np.array_equal(out.x.cpu().numpy(), xo)
    assert np.array_equal(out.results["final_cost"], ro["final_cost"])
    assert np.array_equal(out.results["last_lambda"], ro["last_lambda"])


# ---- a1 + a5 + a6: one Build + Solve -----------------------------------------------------------------
@pytest.mark.parametrize("B,m,n", [(4, 256, 64), (3, 300, 100), (3, 1024, 256), (2, 2048, 512),
                                   (4, 256, 57), (3, 300, 101), (2, 900, 258), (2, 1100, 511)])  # n % 4 != 0: padded copy
def test_lg_build_solve_parity(ctx, B, m, n):
    assert ctx.kernel_family(torch.float32, n) == 3
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32, p0=3)
    r, J = O.synth_eval(A, y, x0)
    lam = np.full(B, np.float32(1e-4), np.float32)
    lam[0] = 0
    out = ctx.build_solve(torch.from_numpy(J).cuda(), torch.from_numpy(r).cuda(), torch.from_numpy(lam).cuda(),
                          want_H=True, want_g=True)
    ctx.sync()
    for p in range(B):
        o = O.build_solve(J[p], r[p], float(lam[p]))
        assert out["status"][p].item() == o["status"] == 0
        assert rel_err(out["dx"][p].cpu().numpy()[None], o["dx"][None]) <= 1e-4, (p, rel_err(out["dx"][p].cpu().numpy()[None], o["dx"][None]))
        assert rel_err(out["cost"][p].item(), o["cost"]) <= 1e-5
        assert rel_err(out["g"][p].cpu().numpy()[None], o["g"][None]) <= 1e-5
        assert rel_err(out["H"][p].cpu().numpy()[None], o["H"][None]) <= 1e-5


# ---- a7-a10: the whole LM loop -----------------------------------------------------------------------
def run_both(ctx, B, m, n, p0=0, **optkw):
    import tinyopt_b200 as tb
    kw = {**FLOAT_OPTS, **optkw}
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32, p0=p0)
    xo, ro, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw), fast=m * n >= 1 << 18)  # same bits, sooner
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, torch.float32, p0=p0, layout=tb.PROBLEM_MAJOR)
    out = ctx.optimize_batch(dA, dy, dx0, tb.options(**kw), layout=tb.PROBLEM_MAJOR)
    return xo, ro, out


def robust_decisions(B, m, n, p0=0, **optkw):
    """Problems whose accept / stop decisions do not ride on FP32 rounding.  The large-n family is not
    bit-identical to the oracle (tensor-core JᵀJ, tree-summed dot products), so "iteration counts
    identical" is only well defined where every branch decision of the run clears FP32 noise.  The
    oracle is re-run in DOUBLE on the same float inputs and reports its decision margins: the
    accept/reject test `derr < 0` (optimizer.h:429) must clear the relative FP32 noise of a cost
    (~1e-6 for a sum of ~1e3 squares, x20), every stop threshold (optimizer.h:518-528) must be 10 %
    away.  For the rest, a decision at the noise floor may fall either way in ANY float
    implementation, the reference's included (the float oracle itself disagrees with the double one
    on many of them)."""
    kw = {**FLOAT_OPTS, **optkw}
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32, p0=p0)
    _, r64, _ = O.synth_lm_run(A.astype(np.float64), y.astype(np.float64), x0.astype(np.float64), O.default_options(**kw),
                               fast=m * n >= 1 << 18)
    return (r64["sign_margin"] > 2e-5) & (r64["thr_margin"] > 0.1)


@pytest.mark.parametrize("B,m,n,min_robust", [(6, 256, 64, 0), (150, 300, 60, 0.5), (5, 700, 128, 0), (3, 1500, 320, 0), (3, 2048, 512, 0),
                                              (150, 300, 57, 0.5), (6, 500, 99, 0), (3, 1200, 258, 0)])  # n % 4 != 0: padded copy
def test_lg_lm_run_parity(ctx, B, m, n, min_robust):
    xo, ro, out = run_both(ctx, B, m, n)
    rg = out.results
    robust = robust_decisions(B, m, n)
    assert robust.mean() >= min_robust, robust.mean()
    assert np.array_equal(rg["num_iters"][robust], ro["num_iters"][robust])      # iteration counts identical
    assert np.array_equal(rg["stop_reason"][robust], ro["stop_reason"][robust])
    assert np.abs(rg["num_iters"].astype(int) - ro["num_iters"]).max() <= 1      # noise-floor decisions: one Step at most
    assert rel_err(out.x.cpu().numpy(), xo) <= 1e-4
    assert rel_err(rg["final_cost"], ro["final_cost"]) <= 1e-4
    assert (ro["stop_reason"] > 0).all() and (rg["stop_reason"] > 0).all()


def test_lg_lm_run_parity_real_c5_shape(ctx):
    """The REAL C5 shape (n = 512, m = 4096).  The decision-margin census of the whole C5 batch
    (profiles/r2_margin_census.json, tools/margin_census.py) shows what this workload is: every problem
    converges in 3 Steps and takes a 4th whose cost change is ~1e-8 relative — below FP32 resolution of a
    sum of 4096 squares — so the SIGN of that last derr (optimizer.h:429) is rounding noise for 99.8 % of
    the problems, in ANY float implementation (the float oracle agrees with its own double re-run on the
    stop reason of only 50 %).  The margin-robust set is therefore empty by construction and cannot be the
    filter here.  What IS well defined, and asserted unconditionally on every problem: the iteration count
    (both outcomes of the noise decision stop at Step 4: accepted -> kMinRelError, rejected -> roll-back +
    kMinDeltaNorm), the solution to 1e-4 per problem (the two outcomes differ by the last, < 3e-5, step),
    the final cost, and that every stop is one of those two."""
    B, m, n = 16, 4096, 512
    xo, ro, out = run_both(ctx, B, m, n)
    rg = out.results
    assert np.array_equal(rg["num_iters"], ro["num_iters"])                     # ALL problems, no filter
    assert set(np.unique(rg["stop_reason"])) <= {2, 3} and set(np.unique(ro["stop_reason"])) <= {2, 3}
    assert rel_err(out.x.cpu().numpy(), xo) <= 1e-4
    assert rel_err(rg["final_cost"], ro["final_cost"]) <= 1e-4
    assert np.array_equal(rg["num_builds"], rg["num_iters"])                     # every pass rebuilt H (no eval-only pass)
    # where the decisions ARE robust (if any problem is), the stop reasons agree too
    robust = robust_decisions(B, m, n)
    assert np.array_equal(rg["stop_reason"][robust], ro["stop_reason"][robust])


@pytest.mark.parametrize("optkw", [dict(solver_type=1), dict(max_iters=2), dict(damping_init=10.0, max_consec_failures=2),
                                   dict(check_min_H_diag=1e4), dict(min_rerr_dec=1e-10, min_step_norm2=1e-14, max_iters=8)])
def test_lg_lm_run_option_variants(ctx, optkw):
    """Gauss-Newton, the iteration cap, heavy damping with early give-up, the diagonal check
    (kSolverFailed), and tinyopt's default thresholds (fp32 noise floor: rejected steps, roll-backs
    and cost-only passes with the stale re-damped H_).  Off the noise floor the counts must be
    identical; on it (last variant) decisions ride on rounding, so only the solution is compared."""
    xo, ro, out = run_both(ctx, 8, 300, 64, **optkw)
    rg = out.results
    if "min_rerr_dec" not in optkw:
        robust = robust_decisions(8, 300, 64, **optkw)
        assert np.array_equal(rg["num_iters"][robust], ro["num_iters"][robust])
        assert np.array_equal(rg["stop_reason"][robust], ro["stop_reason"][robust])
        assert np.abs(rg["num_iters"].astype(int) - ro["num_iters"]).max() <= 1
    else:
        assert (rg["num_builds"] < rg["num_iters"]).any()            # cost-only passes happened
    ok = ro["stop_reason"] > 0
    if ok.any():
        assert rel_err(out.x.cpu().numpy()[ok], xo[ok]) <= 1e-4


def test_lg_phase_timers_and_launch_count(ctx):
    l0 = ctx.launch_count
    run_both(ctx, 4, 256, 64)
    assert ctx.launch_count - l0 >= 4
    ms = [ctx.last_phase_ms(k) for k in range(3)]
    assert all(t > 0 and c >= 1 for t, c in ms)


@pytest.mark.parametrize("n", [56, 59, 61, 66, 127, 129, 190, 255, 333, 385, 450, 509])
def test_lg_every_n_lm_run(ctx, n):
    """Odd sizes of the large-n family (strip boundaries +-1, n % 4 != 0 through the padded copy, partial
    TMA boxes): x within 1e-4 of the oracle, iteration counts identical where the decisions are robust."""
    B, m = 5, 2 * n + 37
    xo, ro, out = run_both(ctx, B, m, n, p0=n)
    robust = robust_decisions(B, m, n, p0=n)
    rg = out.results
    assert np.array_equal(rg["num_iters"][robust], ro["num_iters"][robust])
    assert (np.abs(rg["num_iters"].astype(int) - ro["num_iters"].astype(int)) <= 1).all()
    same = rg["num_iters"] == ro["num_iters"]
    if same.any():
        assert rel_err(out.x.cpu().numpy()[same], xo[same]) <= 1e-4
