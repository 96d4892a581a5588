"""GPU parity tests of SURVEY.md §8(f) rank 2 — the step after the path: batched InvCov / MaxStdDev /
Output::Covariance (math.h:44-57, solvers/lm.h:157-187, output.h:81-103; reference tests:
tests/cov.cpp) against the CPU oracle's too_inv_cov.  The factorisation and the substitutions run in
the oracle's operation order, so the bar is BIT equality."""
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


def spd(rng, n, dtype, cond=50.0):
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    d = np.geomspace(1.0, cond, n)
    return ((q * d) @ q.T).astype(dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 3, 6, 12, 31, 32, 33, 50, 64])
def test_inv_cov_small_bitexact(ctx, n, dtype):
    rng = np.random.default_rng(100 + n)
    B = 37
    A = np.stack([spd(rng, n, dtype) for _ in range(B)])
    A[1] *= 1e-3
    if n > 1:
        A[2][np.tril_indices(n, -1)] = 7.0          # only the upper triangle may be read
        A[3] = -A[3]                                # not positive -> nullopt
        A[4][:] = 0                                 # ZeroSign: accepted, inverse 0 (pseudo-inverse of D)
        a = A[5]; np.fill_diagonal(a, a.diagonal().max())   # ties on the diagonal: Eigen's first-max order
    cov, ms, st = ctx.inv_cov(torch.from_numpy(A).cuda())
    ctx.sync()
    cov, ms, st = cov.cpu().numpy(), ms.cpu().numpy(), st.cpu().numpy()
    for p in range(B):
        ref = O.inv_cov(A[p])
        assert (ref is None) == (st[p] == 1), (n, p)
        if ref is not None:
            assert np.array_equal(cov[p], ref, equal_nan=True), (n, p, np.abs(cov[p] - ref).max())
            assert ms[p] == np.sqrt(ref.max()).astype(dtype) or np.isclose(ms[p], np.sqrt(ref.max()), rtol=1e-6)
        else:
            assert ms[p] == 0
    if n > 1:
        assert st[3] == 1 and st[4] == 0


@pytest.mark.parametrize("n", [65, 96, 200, 512])
def test_inv_cov_large_bitexact(ctx, n):
    rng = np.random.default_rng(n)
    B = 3
    A = np.stack([spd(rng, n, np.float32) for _ in range(B)])
    A[1][np.tril_indices(n, -1)] = -3.0
    A[2] = -A[2]
    cov, ms, st = ctx.inv_cov(torch.from_numpy(A).cuda())
    ctx.sync()
    cov, ms, st = cov.cpu().numpy(), ms.cpu().numpy(), st.cpu().numpy()
    for p in range(B):
        ref = O.inv_cov(A[p])
        assert (ref is None) == (st[p] == 1), (n, p)
        if ref is not None:
            assert np.array_equal(cov[p], ref), (n, p, np.abs(cov[p] - ref).max())
            assert np.isclose(ms[p], np.sqrt(ref.max()), rtol=1e-6)
    assert st[2] == 1 and ms[2] == 0


@pytest.mark.parametrize("dtype,n", [(np.float64, 65), (np.float64, 96), (np.float64, 200), (np.float64, 513),
                                     (np.float32, 513), (np.float32, 700)])
def test_inv_cov_general_family_bitexact(ctx, dtype, n):
    """InvCov above the specialised kernels' sizes (double n > 64, float n > 512): the general family's LDLT + every
    column of the identity at once, each entry in the oracle's update order."""
    rng = np.random.default_rng(n)
    B = 3
    A = np.stack([spd(rng, n, dtype) for _ in range(B)])
    A[1][np.tril_indices(n, -1)] = -3.0
    A[2] = -A[2]
    cov, ms, st = ctx.inv_cov(torch.from_numpy(A).cuda())
    ctx.sync()
    cov, ms, st = cov.cpu().numpy(), ms.cpu().numpy(), st.cpu().numpy()
    for p in range(B):
        ref = O.inv_cov(A[p])
        assert (ref is None) == (st[p] == 1), (n, p)
        if ref is not None:
            assert np.array_equal(cov[p], ref), (n, p, np.abs(cov[p] - ref).max())
            assert ms[p] == np.sqrt(ref.max())
    assert st[2] == 1 and ms[2] == 0


def test_inv_cov_is_the_inverse(ctx):
    """tests/cov.cpp:20-42 style: InvCov(H) * H = I to 1e-5 (double)."""
    rng = np.random.default_rng(5)
    A = np.stack([spd(rng, 6, np.float64, cond=20.0) for _ in range(64)])
    cov, _, st = ctx.inv_cov(torch.from_numpy(A).cuda(), want_max_std=False)
    ctx.sync()
    assert (st.cpu().numpy() == 0).all()
    err = np.abs(np.einsum("bij,bjk->bik", cov.cpu().numpy(), A) - np.eye(6)).max()
    assert err < 1e-10, err


def test_solver_covariance_matches_prior(ctx):
    """tests/cov.cpp:66-91: whitened prior L^T (x - y): Output::Covariance() is the prior covariance
    (+-1e-5); rescaled multiplies by final_cost^2 / (m - n) only when m > n (output.h:94-97)."""
    import tinyopt_b200 as tb
    Cy = np.array([[10.0, 2.0], [2.0, 4.0]])
    Lt = np.linalg.cholesky(np.linalg.inv(Cy)).T
    y = 2 * np.array([0.25, -0.6])
    Ltd = torch.from_numpy(Lt).cuda(); yd = torch.from_numpy(y).cuda()
    B = 5
    s = tb.BatchSolver(ctx, B, 2, torch.float64, tb.options())
    s.reset(torch.zeros((B, 2), dtype=torch.float64, device="cuda"))
    steps = 0
    while s.num_active() > 0 and steps < 100:
        x = s.x.clone()
        s.step(Ltd.unsqueeze(0).expand(B, 2, 2).contiguous(), (x - yd) @ Ltd.T)
        steps += 1
    H = s.final_hessian().cpu().numpy()
    cov, st = s.covariance()
    cov_r, _ = s.covariance(rescaled=True)
    ms = s.max_std_dev().cpu().numpy()
    ctx.sync()
    cov, st = cov.cpu().numpy(), st.cpu().numpy()
    assert (st == 0).all()
    assert np.abs(cov[0] - Cy).max() < 1e-5
    for p in range(B):
        ref = O.inv_cov(H[p])
        assert np.array_equal(cov[p], ref)
        assert np.isclose(ms[p], np.sqrt(ref.max()), rtol=1e-12)
    assert np.array_equal(cov_r.cpu().numpy(), cov)   # m == n: no rescaling
    s.close()
