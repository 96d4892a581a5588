// C++ adaptor test (tinyopt_b200.hpp over the C-ABI).  Reads like the reference's own tests:
// tests/sqrt2.cpp, tests/solvers.cpp:25-45, tests/circle.cpp, tests/basic.cpp:41-54.
// Exit code 0 = all checks passed, 3 = no CUDA device (the library has no CPU fallback), 1 = failure.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tinyopt_b200.hpp"

using namespace tinyopt::b200;

static int g_failures = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);  \
      ++g_failures;                                                        \
    }                                                                      \
  } while (0)

// tests/sqrt2.cpp: r = x*x - 2, J = 2x, from several x0
static void test_sqrt2(const Context &ctx) {
  const int64_t B = 96;
  std::vector<double> xs(B);
  for (int64_t p = 0; p < B; ++p) xs[p] = (p % 3 == 0) ? 1.0 : (p % 3 == 1 ? -0.3 : 3.2);
  auto residuals = [](size_t, const double *x, double *r, double *J) {
    r[0] = x[0] * x[0] - 2.0;
    if (J) J[0] = 2.0 * x[0];
  };
  Options options;  // tests/sqrt2.cpp:22-28
  options.max_iters = 20;
  options.max_consec_failures = 0;
  auto outs = OptimizeBatch<double>(ctx, xs.data(), B, 1, 1, residuals, options);
  for (int64_t p = 0; p < B; ++p) {
    CHECK(outs[p].Succeeded());
    CHECK(outs[p].Converged());
    CHECK(std::fabs(std::fabs(xs[p]) - std::sqrt(2.0)) < 1e-5);
    CHECK(outs[p].has_final_hessian() && outs[p].final_hessian[0] > 0);
  }
  // SURVEY.md §8(c) golden vector: x0 = 1 -> 5 Steps, kMinError (the same under default Options)
  CHECK(outs[0].num_iters == 5);
  CHECK(outs[0].stop_reason == StopReason::kMinError);
  {  // default Options: x0 = -0.3 overshoots and gives up after 5 consecutive failures, as in the reference
    std::vector<double> x1 = {1.0, -0.3};
    auto o2 = OptimizeBatch<double>(ctx, x1.data(), 2, 1, 1, residuals, Options());
    CHECK(o2[0].num_iters == 5 && o2[0].stop_reason == StopReason::kMinError);
    CHECK(o2[1].stop_reason == StopReason::kMaxConsecNoDecr);
  }
  {  // benchmarks/dense.cpp:39-52: the reference's own sqrt(2) benchmark runs with hessian.use_ldlt = false
    Options inv = options;
    inv.hessian.use_ldlt = false;
    std::vector<double> x2(B);
    for (int64_t p = 0; p < B; ++p) x2[p] = (p % 3 == 0) ? 1.0 : (p % 3 == 1 ? -0.3 : 3.2);
    auto o3 = OptimizeBatch<double>(ctx, x2.data(), B, 1, 1, residuals, inv);
    for (int64_t p = 0; p < B; ++p) {
      CHECK(o3[p].Converged());
      CHECK(o3[p].num_iters == outs[p].num_iters);
      CHECK(std::fabs(x2[p] - xs[p]) < 1e-12);
    }
  }
  std::printf("sqrt2: x[0]=%.16g iters=%d stop=%d\n", xs[0], (int)outs[0].num_iters, (int)outs[0].stop_reason);
}

// tests/basic.cpp:41-54: r = x - 2, H = 1 -> kMinDeltaNorm within 2..5 iterations (float and double)
template <typename T>
static void test_basic() {
  Context ctx(0);
  const int64_t B = 40;
  std::vector<T> xs(B, (T)1);
  auto residuals = [](size_t, const T *x, T *r, T *J) {
    r[0] = x[0] - (T)2;
    if (J) J[0] = (T)1;
  };
  Options options;
  options.cost.use_squared_norm = false;  // Cost(|r|): the reference's lambda returns the norm
  auto outs = OptimizeBatch<T>(ctx, xs.data(), B, 1, 1, residuals, options);
  for (int64_t p = 0; p < B; ++p) {
    CHECK(outs[p].Succeeded());
    CHECK(outs[p].num_iters >= 2 && outs[p].num_iters <= 5);
    CHECK(outs[p].final_cost.cost < 1e-5);
    CHECK(std::fabs((double)xs[p] - 2.0) < 1e-5);
  }
  std::printf("basic<%s>: x=%.9g iters=%d stop=%d\n", sizeof(T) == 4 ? "float" : "double", (double)xs[0],
              (int)outs[0].num_iters, (int)outs[0].stop_reason);
}

// tests/solvers.cpp:25-45: one Build + Solve on r = x - y from x = 0 gives dx ~= y
static void test_build_solve(const Context &ctx) {
  const int64_t B = 50;
  const int n = 2, m = 2;
  std::vector<double> J(B * m * n, 0.0), r(B * m), lam(B, 1e-4), dx(B * n), cost(B);
  for (int64_t p = 0; p < B; ++p) {
    J[p * 4 + 0] = 1.0;
    J[p * 4 + 3] = 1.0;
    r[p * 2 + 0] = -(4.0 + p);  // r = x - y at x = 0
    r[p * 2 + 1] = -(5.0 + p);
  }
  auto st = BuildSolve<double>(ctx, J.data(), r.data(), B, m, n, lam.data(), dx.data(), cost.data());
  for (int64_t p = 0; p < B; ++p) {
    CHECK(st[p] == 0);
    CHECK(std::fabs(dx[p * 2] - (4.0 + p)) < 1e-2 && std::fabs(dx[p * 2 + 1] - (5.0 + p)) < 1e-2);
    CHECK(std::fabs(cost[p] - ((4.0 + p) * (4.0 + p) + (5.0 + p) * (5.0 + p))) < 1e-9);
  }
  std::printf("build_solve: dx[0]=(%.6f, %.6f)\n", dx[0], dx[1]);
}

// tests/circle.cpp: fit centre + radius to 10 noisy points per problem
static void test_circle(const Context &ctx) {
  const int64_t B = 200;
  const int n = 3, m = 10;
  std::vector<double> pts(B * m * 2), xs(B * n), truth(B * n);
  uint64_t s = 12345;
  auto rnd = [&]() {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(s >> 11) / 9007199254740992.0;
  };
  for (int64_t p = 0; p < B; ++p) {
    const double cx = 2 * rnd() - 1, cy = 2 * rnd() - 1, rad = 1 + rnd();
    truth[p * 3] = cx; truth[p * 3 + 1] = cy; truth[p * 3 + 2] = rad;
    for (int i = 0; i < m; ++i) {
      const double a = 2 * M_PI * (i + rnd() * 0.5) / m;
      pts[(p * m + i) * 2] = cx + rad * std::cos(a);
      pts[(p * m + i) * 2 + 1] = cy + rad * std::sin(a);
    }
    xs[p * 3] = cx + 0.2; xs[p * 3 + 1] = cy - 0.1; xs[p * 3 + 2] = rad * 1.2;
  }
  auto residuals = [&](size_t p, const double *x, double *r, double *J) {
    for (int i = 0; i < m; ++i) {
      const double dx = pts[(p * m + i) * 2] - x[0], dy = pts[(p * m + i) * 2 + 1] - x[1];
      const double d = std::sqrt(dx * dx + dy * dy);
      r[i] = d - x[2];
      if (J) {
        J[i * 3] = -dx / d;
        J[i * 3 + 1] = -dy / d;
        J[i * 3 + 2] = -1.0;
      }
    }
  };
  auto outs = OptimizeBatch<double>(ctx, xs.data(), B, n, m, residuals);
  for (int64_t p = 0; p < B; ++p) {
    CHECK(outs[p].Succeeded());
    for (int j = 0; j < 3; ++j) CHECK(std::fabs(xs[p * 3 + j] - truth[p * 3 + j]) < 1e-5);
  }
  std::printf("circle: iters[0]=%d stop[0]=%d\n", (int)outs[0].num_iters, (int)outs[0].stop_reason);
}

// The polynomial family through the host-driven solver (user lambda on the host, canonical fma
// order) must equal the fused device-resident loop bit for bit, iteration counts included.
template <typename T>
static void test_family_equivalence(const Context &ctx, int n, int m) {
  const int64_t B = 70;
  const T alpha = (T)0.1;
  std::vector<T> A(B * m * n), y(B * m), x1(B * n), x2;
  uint64_t s = 99;
  auto rnd = [&]() {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (T)((double)(s >> 11) / 9007199254740992.0 * 2 - 1);
  };
  std::vector<T> xt(n);
  for (int64_t p = 0; p < B; ++p) {
    for (int j = 0; j < n; ++j) xt[j] = rnd();
    for (int i = 0; i < m; ++i) {
      T t = 0;
      for (int j = 0; j < n; ++j) {
        A[(p * m + i) * n + j] = rnd() / std::sqrt((T)n);
        t = std::fma(A[(p * m + i) * n + j], xt[j], t);
      }
      y[p * m + i] = t + alpha * t * t * t + (T)0.01 * rnd();
    }
    for (int j = 0; j < n; ++j) x1[p * n + j] = xt[j] + (T)0.3 * rnd();
  }
  x2 = x1;
  Options options;
  if (sizeof(T) == 4) {
    options.min_rerr_dec = 1e-5f;
    options.min_step_norm2 = 1e-9f;
  }
  options.hessian.save_last = false;
  auto residuals = [&](size_t p, const T *x, T *r, T *J) {
    for (int i = 0; i < m; ++i) {
      const T *a = &A[(p * m + i) * n];
      T t = 0;
      for (int j = 0; j < n; ++j) t = std::fma(a[j], x[j], t);
      const T t2 = t * t;
      r[i] = std::fma(t, std::fma(alpha, t2, (T)1), -y[p * m + i]);
      if (J) {
        const T sc = std::fma((T)3 * alpha, t2, (T)1);
        for (int j = 0; j < n; ++j) J[i * n + j] = sc * a[j];
      }
    }
  };
  auto o1 = OptimizeBatch<T>(ctx, x1.data(), B, n, m, residuals, options);
  auto o2 = OptimizePolynomialBatch<T>(ctx, A.data(), y.data(), alpha, x2.data(), B, m, n, options);
  int total = 0;
  for (int64_t p = 0; p < B; ++p) {
    CHECK(o1[p].num_iters == o2[p].num_iters);
    CHECK(o1[p].stop_reason == o2[p].stop_reason);
    CHECK(o1[p].final_cost.cost == o2[p].final_cost.cost);
    total += o1[p].num_iters;
  }
  CHECK(std::memcmp(x1.data(), x2.data(), x1.size() * sizeof(T)) == 0);
  std::printf("family<%s> n=%d m=%d: total iters=%d, host-driven == device-resident\n",
              sizeof(T) == 4 ? "float" : "double", n, m, total);
}

// tests/cov.cpp:66-91: whitened prior L^T (x - y); Output::Covariance() == the prior covariance
static void test_covariance(const Context &ctx) {
  const double Cy[4] = {10.0, 2.0, 2.0, 4.0};
  // L^T of inv(Cy) = [[a, b], [0, c]] with inv(Cy) = 1/36 * [[4, -2], [-2, 10]]
  const double i00 = 4.0 / 36.0, i01 = -2.0 / 36.0, i11 = 10.0 / 36.0;
  const double l00 = std::sqrt(i00), l10 = i01 / l00, l11 = std::sqrt(i11 - l10 * l10);
  const double Lt[4] = {l00, l10, 0.0, l11};  // L^T, row-major
  const double y[2] = {0.5, -1.2};
  const int64_t B = 4;
  std::vector<double> xs((size_t)B * 2, 0.0);
  auto residuals = [&](size_t, const double *x, double *r, double *J) {
    const double d0 = x[0] - y[0], d1 = x[1] - y[1];
    r[0] = Lt[0] * d0 + Lt[1] * d1;
    r[1] = Lt[2] * d0 + Lt[3] * d1;
    if (J) { J[0] = Lt[0]; J[1] = Lt[1]; J[2] = Lt[2]; J[3] = Lt[3]; }
  };
  auto outs = OptimizeBatch<double>(ctx, xs.data(), B, 2, 2, residuals, Options());
  auto covs = CovarianceBatch(ctx, outs, 2);
  for (int64_t p = 0; p < B; ++p) {
    CHECK(outs[p].Converged());
    CHECK(covs[p].size() == 4);
    if (covs[p].size() == 4)
      for (int e = 0; e < 4; ++e) CHECK(std::fabs(covs[p][e] - Cy[e]) < 1e-5);
    CHECK(std::fabs(xs[2 * p] - y[0]) < 1e-6 && std::fabs(xs[2 * p + 1] - y[1]) < 1e-6);
  }
  std::printf("covariance: cov[0]=(%.6f, %.6f; %.6f, %.6f)\n", covs[0][0], covs[0][1], covs[0][2], covs[0][3]);
}

static void test_misuse(const Context &ctx) {
  bool threw = false;
  try {
    std::vector<double> x(4);
    OptimizeBatch<double>(ctx, x.data(), 4, 0, 1, [](size_t, const double *, double *, double *) {});
  } catch (const std::invalid_argument &) {
    threw = true;
  }
  CHECK(threw);
}

// tests/optimize_easy.cpp:35-80: Rosenbrock through the manual accumulation contract with the TRUE Hessian
// (docs/API.md:37-57); the reference asserts Succeeded, Converged, x = (1, 1) +- 1e-5.  And benchmarks/dense.cpp:57-66
// "Prior n": only H.diagonal() is filled.
static void test_accumulation_contract(const Context &ctx) {
  const int64_t B = 33;
  std::vector<double> xs(B * 2);
  for (int64_t p = 0; p < B; ++p) { xs[2 * p] = -1.2 + 0.01 * (double)(p % 7); xs[2 * p + 1] = 1.0 - 0.01 * (double)(p % 5); }
  Options options;
  options.max_iters = 200;
  options.min_rerr_dec = 0;
  options.max_consec_failures = 20;
  auto loss = [](size_t, const double *v, double *grad, double *H) {
    const double x = v[0], y = v[1], t1 = 1.0 - x, t2 = y - x * x;
    if (grad) {
      grad[0] = -2.0 * t1 - 400.0 * x * t2;
      grad[1] = 200.0 * t2;
      H[0] = 2.0 - 400.0 * y + 1200.0 * x * x;
      H[1] = -400.0 * x;  // (the lower triangle is left at zero: only the upper one is read, docs/API.md:170)
      H[3] = 200.0;
    }
    Cost c;
    c.cost = t1 * t1 + 100.0 * t2 * t2;
    c.num_resisuals = 1;
    return c;
  };
  auto outs = OptimizeBatchAcc<double>(ctx, xs.data(), B, 2, loss, options);
  CHECK(outs[0].Succeeded() && outs[0].Converged());
  CHECK(std::fabs(xs[0] - 1.0) < 1e-5 && std::fabs(xs[1] - 1.0) < 1e-5);
  CHECK(outs[0].num_iters == 59 && outs[0].num_failures == 24);   // the oracle's trajectory for x0 = (-1.2, 1)
  CHECK(outs[0].has_final_hessian() && std::fabs(outs[0].final_hessian[3] - 200.0) < 1e-9);
  int conv = 0;
  for (int64_t p = 0; p < B; ++p) conv += outs[p].Converged();
  CHECK(conv >= B * 3 / 4);
  // "Prior n"
  const int n = 12;
  std::vector<float> y(B * n), sd(B * n), x(B * n);
  uint64_t s = 5;
  auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (float)((double)(s >> 11) / 9007199254740992.0 * 2 - 1); };
  for (auto &v : y) v = rnd();
  for (auto &v : sd) { v = rnd(); if (std::fabs(v) < 0.05f) v = 0.3f; }
  for (auto &v : x) v = rnd();
  Options fo;
  fo.min_rerr_dec = 1e-5f;
  fo.min_step_norm2 = 1e-9f;
  auto prior = [&](size_t p, const float *xv, float *grad, float *H) {
    Cost c;
    float acc = 0.f;
    for (int j = 0; j < n; ++j) {
      const float res = (xv[j] - y[p * n + j]) / sd[p * n + j];
      acc += res * res;
      if (grad) {
        grad[j] = res / sd[p * n + j];
        H[j * n + j] = (1.f / sd[p * n + j]) * (1.f / sd[p * n + j]);
      }
    }
    c.cost = acc;
    c.num_resisuals = 1;
    return c;
  };
  auto po = OptimizeBatchAcc<float>(ctx, x.data(), B, n, prior, fo);
  float worst = 0.f;
  for (size_t i = 0; i < x.size(); ++i) worst = std::fmax(worst, std::fabs(x[i] - y[i]));
  CHECK(worst < 1e-3f);
  for (int64_t p = 0; p < B; ++p) CHECK(po[p].Converged());
  // the same contract above n = 55 (the reference's dynamic-size solver has no cap, math.h:232-240): "Prior 80" in double
  {
    const int nl = 80;
    const int64_t Bl = 6;
    std::vector<double> yl(Bl * nl), sl(Bl * nl), xl(Bl * nl);
    for (auto &v : yl) v = rnd();
    for (auto &v : sl) { v = rnd(); if (std::fabs(v) < 0.05) v = 0.3; }
    for (auto &v : xl) v = rnd();
    auto prior_l = [&](size_t p, const double *xv, double *grad, double *H) {
      Cost c;
      double acc = 0.0;
      for (int j = 0; j < nl; ++j) {
        const double res = (xv[j] - yl[p * nl + j]) / sl[p * nl + j];
        acc += res * res;
        if (grad) {
          grad[j] = res / sl[p * nl + j];
          H[j * nl + j] = (1.0 / sl[p * nl + j]) * (1.0 / sl[p * nl + j]);
        }
      }
      c.cost = acc;
      c.num_resisuals = 1;
      return c;
    };
    auto pl = OptimizeBatchAcc<double>(ctx, xl.data(), Bl, nl, prior_l);
    double wl = 0.0;
    for (size_t i = 0; i < xl.size(); ++i) wl = std::fmax(wl, std::fabs(xl[i] - yl[i]));
    CHECK(wl < 1e-6);
    for (int64_t p = 0; p < Bl; ++p) CHECK(pl[p].Converged());
    std::printf("accumulation contract above n = 55: Prior 80 (double): max |x - y| = %.2e\n", wl);
  }
  // tests/sparse.cpp:19-57 "tinyopt_sparse": res = 10 x - 2 over 100 parameters, H = J^T J handed over as triplets (diagonal),
  // Cost(res.norm(), res.size()), check_final_cost = false; reference assertions: converged, min / max of x == 0.2 +- 1e-5
  {
    const int ns = 100;
    const int64_t Bs = 3;
    std::vector<double> xsp(Bs * ns);
    for (auto &v : xsp) v = rnd();
    std::vector<int32_t> rws(ns), cls(ns);
    for (int i = 0; i < ns; ++i) rws[i] = cls[i] = i;
    Options so;
    so.check_final_cost = false;
    auto sparse_acc = [&](size_t, const double *xv, double *grad, double *vals) {
      Cost c;
      double acc = 0.0;
      for (int j = 0; j < ns; ++j) {
        const double res = 10.0 * xv[j] - 2.0;
        acc += res * res;
        if (grad) {
          grad[j] = 10.0 * res;
          vals[j] = 100.0;
        }
      }
      c.cost = std::sqrt(acc);
      c.num_resisuals = ns;
      return c;
    };
    auto sp = OptimizeBatchAccSparse<double>(ctx, xsp.data(), Bs, ns, rws, cls, sparse_acc, so);
    double lo = 1e300, hi = -1e300;
    for (double v : xsp) { lo = std::fmin(lo, v); hi = std::fmax(hi, v); }
    for (int64_t p = 0; p < Bs; ++p) CHECK(sp[p].Succeeded() && sp[p].Converged());
    CHECK(std::fabs(lo - 0.2) < 1e-5 && std::fabs(hi - 0.2) < 1e-5);
    std::printf("sparse accumulation signature (tests/sparse.cpp): min x = %.9f, max x = %.9f\n", lo, hi);
  }
  std::printf("accumulation contract: Rosenbrock (true Hessian) iters[0]=%d failures[0]=%d, converged %d/%d; Prior 12: max |x - y| = %.2e\n",
              (int)outs[0].num_iters, (int)outs[0].num_failures, conv, (int)B, (double)worst);
}

// every visible GPU behind one call: contiguous shards, one context and one host thread per device
static void test_multi_device() {
  const std::vector<int> devs = MultiContext::all_devices();
  MultiContext mc(devs);
  const int64_t B = 1000;
  const int n = 6, m = 30;
  std::vector<double> A(B * m * n), y(B * m), x1(B * n), x2;
  uint64_t s = 17;
  auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (double)(s >> 11) / 9007199254740992.0 * 2 - 1; };
  for (auto &v : A) v = rnd() / std::sqrt((double)n);
  for (auto &v : y) v = 0.3 * rnd();
  for (auto &v : x1) v = 0.5 * rnd();
  x2 = x1;
  Context c0(0);
  auto o1 = OptimizePolynomialBatch<double>(c0, A.data(), y.data(), 0.1, x1.data(), B, m, n);
  auto o2 = OptimizePolynomialBatch<double>(mc, A.data(), y.data(), 0.1, x2.data(), B, m, n);
  CHECK(std::memcmp(x1.data(), x2.data(), x1.size() * sizeof(double)) == 0);
  for (int64_t p = 0; p < B; ++p) CHECK(o1[p].num_iters == o2[p].num_iters && o1[p].stop_reason == o2[p].stop_reason);
  std::printf("multi-device: %d device(s), %lld problems in contiguous shards == single context\n", mc.size(), (long long)B);
}

int main() {
  try {
    Context ctx(0);
    test_sqrt2(ctx);
    test_basic<double>();
    test_basic<float>();
    test_build_solve(ctx);
    test_circle(ctx);
    test_covariance(ctx);
    test_family_equivalence<double>(ctx, 6, 30);
    test_family_equivalence<float>(ctx, 12, 40);
    test_family_equivalence<float>(ctx, 20, 64);    // warp-per-problem family behind the same SolverType seam
    test_family_equivalence<double>(ctx, 30, 90);
    test_family_equivalence<double>(ctx, 64, 150);  // above n = 55: the general family behind the seam and behind lm_run
    test_accumulation_contract(ctx);
    test_multi_device();
    test_misuse(ctx);
  } catch (const Error &e) {
    std::printf("tinyopt::b200::Error (%d): %s\n", e.code, e.what());
    return e.code == TOB200_ERR_CUDA ? 3 : 1;
  }
  if (g_failures) {
    std::printf("%d check(s) failed\n", g_failures);
    return 1;
  }
  std::printf("all C++ adaptor checks passed\n");
  return 0;
}
