// tinyopt_b200.hpp — header-only C++17 adaptor over the C-ABI (include/tinyopt_b200.h).
//
// This is the host side a tinyopt user / maintainer sees.  It mirrors the reference's interface
// for the batched dense-NLLS path: same `Options` (include/tinyopt/optimizers/options.h:18-156,
// numeric subset, same nesting and defaults), same `StopReason` (stop_reasons.h:14-43), same
// `Output` fields (output.h:26-145), same accumulation contract — the user's residual callable is
// unchanged, it is simply invoked once per problem.  No Eigen, no CUDA headers: device memory goes
// through tob200_device_alloc / tob200_copy_to_*.  Link with -ltinyopt_b200.
//
//   tinyopt::b200::Context ctx(0);
//   auto outs = tinyopt::b200::OptimizeBatch<double>(ctx, xs, B, n, m, residuals, options);
//
// `residuals(p, x, r, J)` fills r[m] and, when J != nullptr, the row-major m x n Jacobian
// (J == nullptr is the reference's cost-only `acc(x, nullptr, H)` call, solvers/gn.h:98-105).
// Errors: misuse throws std::invalid_argument (as tinyopt does, solvers/gn.h:51,66), CUDA / device
// failures throw tinyopt::b200::Error; numerical failures are reported through StopReason.
#pragma once

#include <array>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <memory>
#include <thread>
#include <vector>

#include "tinyopt_b200.h"

namespace tinyopt::b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

/// stop_reasons.h:14-43
enum class StopReason : int {
  kOutOfMemory = -4,
  kSolverFailed = -3,
  kSystemHasNaNOrInf = -2,
  kSkipped = -1,
  kNone = 0,
  kMinError,
  kMinRelError,
  kMinDeltaNorm,
  kMinGradNorm,
  kMaxIters,
  kMaxNoDecr,
  kMaxConsecNoDecr,
  kTimedOut,
  kUserStopped
};

/// optimizers/options.h:18-156 (numeric subset; logging, callbacks and GD are host-only concerns)
struct Options {
  enum Solver { LevenbergMarquardt = 0, GaussNewton };
  Solver solver_type = LevenbergMarquardt;
  Options(Solver type = LevenbergMarquardt) : solver_type(type) {}
  bool check_final_cost = false;
  bool use_step_quality_approx = false;
  float grad_clipping = 0;
  struct Hessian {
    bool use_ldlt = true;
    bool H_is_full = true;
    float check_min_H_diag = 0;
    bool save_last = true;
  } hessian;
  struct CostScaling {
    bool use_squared_norm = true;
    bool downscale_by_2 = false;
    bool normalize = false;
  } cost;
  uint16_t max_iters = 50;
  float min_error = 1e-12f;
  float min_rerr_dec = 1e-10f;
  float min_step_norm2 = 1e-14f;
  float min_grad_norm2 = 1e-18f;
  uint8_t max_total_failures = 0;
  uint8_t max_consec_failures = 5;
  struct LM {
    float damping_init = 1e-4f;
    std::array<float, 2> damping_range{{1e-9f, 1e9f}};
    float good_factor = 1.0f / 3.0f;
    float bad_factor = 2.0f;
  } lm;

  tob200_options pod() const {
    tob200_options o;
    o.solver_type = (int)solver_type;
    o.check_final_cost = check_final_cost;
    o.use_step_quality_approx = use_step_quality_approx;
    o.grad_clipping = grad_clipping;
    o.use_ldlt = hessian.use_ldlt;
    o.H_is_full = hessian.H_is_full;
    o.check_min_H_diag = hessian.check_min_H_diag;
    o.save_last = hessian.save_last;
    o.use_squared_norm = cost.use_squared_norm;
    o.downscale_by_2 = cost.downscale_by_2;
    o.normalize = cost.normalize;
    o.max_iters = max_iters;
    o.min_error = min_error;
    o.min_rerr_dec = min_rerr_dec;
    o.min_step_norm2 = min_step_norm2;
    o.min_grad_norm2 = min_grad_norm2;
    o.max_total_failures = max_total_failures;
    o.max_consec_failures = max_consec_failures;
    o.damping_init = lm.damping_init;
    o.damping_min = lm.damping_range[0];
    o.damping_max = lm.damping_range[1];
    o.good_factor = lm.good_factor;
    o.bad_factor = lm.bad_factor;
    return o;
  }
};

/// cost.h:18-25
struct Cost {
  double cost = std::numeric_limits<double>::max();
  int num_resisuals = 0;  // (sic) the reference's spelling
};

/// output.h:26-145, one per problem
struct Output {
  bool Succeeded() const { return stop_reason >= StopReason::kNone; }
  bool Converged() const { return stop_reason >= StopReason::kMinError && stop_reason < StopReason::kMaxIters; }
  Cost final_cost;
  double final_rerr_dec = std::numeric_limits<double>::max();
  StopReason stop_reason = StopReason::kNone;
  uint16_t num_residuals = 0;
  uint16_t num_iters = 0;
  uint8_t num_failures = 0;
  uint8_t num_consec_failures = 0;
  std::vector<double> final_hessian;  ///< n*n row-major, un-damped (options.hessian.save_last), else empty
  bool has_final_hessian() const { return !final_hessian.empty(); }
};

inline Output to_output(const tob200_result &r) {
  Output o;
  o.final_cost.cost = r.final_cost;
  o.final_cost.num_resisuals = r.final_num_residuals;
  o.final_rerr_dec = r.final_rerr_dec;
  o.stop_reason = (StopReason)r.stop_reason;
  o.num_residuals = (uint16_t)r.final_num_residuals;
  o.num_iters = (uint16_t)r.num_iters;
  o.num_failures = (uint8_t)r.num_failures;
  o.num_consec_failures = (uint8_t)r.num_consec_failures;
  return o;
}

template <typename Scalar>
struct scalar_traits;
template <>
struct scalar_traits<float> {
  static constexpr int dtype = TOB200_F32;
};
template <>
struct scalar_traits<double> {
  static constexpr int dtype = TOB200_F64;
};

/// One GPU, one stream.  Not thread-safe; distinct contexts are independent (like distinct
/// `Optimizer_` instances in the reference).
class Context {
 public:
  explicit Context(int device = 0) {
    const int rc = tob200_create(&ctx_, device, nullptr);
    if (rc != TOB200_OK) throw Error(rc, std::string("tob200_create: ") + tob200_last_error(nullptr));
  }
  ~Context() { tob200_destroy(ctx_); }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  tob200_ctx *get() const { return ctx_; }
  void check(int rc, const char *what) const {
    if (rc == TOB200_OK) return;
    const std::string msg = std::string(what) + ": " + tob200_last_error(ctx_);
    if (rc == TOB200_ERR_INVALID) throw std::invalid_argument(msg);
    throw Error(rc, msg);
  }
  void sync() const { check(tob200_sync(ctx_), "tob200_sync"); }
  /// Mid-n float runs (13 <= n <= 55): the bit-exact warp-per-problem kernel instead of the tensor-core one (tob200_set_exact)
  void set_exact(bool exact = true) const { check(tob200_set_exact(ctx_, exact ? 1 : 0), "tob200_set_exact"); }

 private:
  tob200_ctx *ctx_ = nullptr;
};

/// RAII device buffer of `count` elements
template <typename T>
class DeviceBuffer {
 public:
  DeviceBuffer(const Context &ctx, size_t count) : ctx_(ctx), count_(count) {
    ctx_.check(tob200_device_alloc(ctx_.get(), count * sizeof(T), &ptr_), "tob200_device_alloc");
  }
  ~DeviceBuffer() { tob200_device_free(ctx_.get(), ptr_); }
  DeviceBuffer(const DeviceBuffer &) = delete;
  DeviceBuffer &operator=(const DeviceBuffer &) = delete;
  T *data() const { return static_cast<T *>(ptr_); }
  size_t size() const { return count_; }
  void upload(const T *src, size_t count) const {
    ctx_.check(tob200_copy_to_device(ctx_.get(), ptr_, src, count * sizeof(T)), "tob200_copy_to_device");
  }
  void download(T *dst, size_t count) const {
    ctx_.check(tob200_copy_to_host(ctx_.get(), dst, ptr_, count * sizeof(T)), "tob200_copy_to_host");
  }

 private:
  const Context &ctx_;
  void *ptr_ = nullptr;
  size_t count_;
};

namespace detail {
inline int build_solve(tob200_ctx *c, const float *J, const float *r, int64_t B, int m, int n, const float *lam, float *dx,
                       double *cost, float *H, float *g, int32_t *st) {
  return tob200_build_solve_f32(c, J, r, TOB200_LAYOUT_PROBLEM_MAJOR, B, m, n, lam, dx, cost, H, g, st);
}
inline int build_solve(tob200_ctx *c, const double *J, const double *r, int64_t B, int m, int n, const double *lam,
                       double *dx, double *cost, double *H, double *g, int32_t *st) {
  return tob200_build_solve_f64(c, J, r, TOB200_LAYOUT_PROBLEM_MAJOR, B, m, n, lam, dx, cost, H, g, st);
}
inline int solver_step(tob200_solver *s, const float *J, const float *r, int m) {
  return tob200_solver_step_f32(s, J, r, TOB200_LAYOUT_PROBLEM_MAJOR, m);
}
inline int solver_step(tob200_solver *s, const double *J, const double *r, int m) {
  return tob200_solver_step_f64(s, J, r, TOB200_LAYOUT_PROBLEM_MAJOR, m);
}
inline int solver_step_hg(tob200_solver *s, const float *g, const float *H, const double *c, const int32_t *nr) {
  return tob200_solver_step_hg_f32(s, g, H, c, nr);
}
inline int solver_step_hg(tob200_solver *s, const double *g, const double *H, const double *c, const int32_t *nr) {
  return tob200_solver_step_hg_f64(s, g, H, c, nr);
}
inline int solver_step_hg_sparse(tob200_solver *s, const float *g, const int32_t *rows, const int32_t *cols, int nnz,
                                 const float *v, const double *c, const int32_t *nr) {
  return tob200_solver_step_hg_sparse_f32(s, g, rows, cols, nnz, v, c, nr);
}
inline int solver_step_hg_sparse(tob200_solver *s, const double *g, const int32_t *rows, const int32_t *cols, int nnz,
                                 const double *v, const double *c, const int32_t *nr) {
  return tob200_solver_step_hg_sparse_f64(s, g, rows, cols, nnz, v, c, nr);
}
inline int lm_run_host(tob200_ctx *c, const tob200_options *o, const float *A, const float *y, float alpha, int64_t B,
                       int m, int n, float *x, tob200_result *res) {
  return tob200_lm_run_host_f32(c, o, A, y, alpha, TOB200_LAYOUT_PROBLEM_MAJOR, B, m, n, x, res);
}
inline int lm_run_host(tob200_ctx *c, const tob200_options *o, const double *A, const double *y, double alpha,
                       int64_t B, int m, int n, double *x, tob200_result *res) {
  return tob200_lm_run_host_f64(c, o, A, y, alpha, TOB200_LAYOUT_PROBLEM_MAJOR, B, m, n, x, res);
}
}  // namespace detail

/// A batch of `SolverLM` (solvers/lm.h:22) behind the `SolverType` seam of `Optimizer_`
/// (optimizers/optimizer.h:34-43): one Build (JᵀJ, Jᵀr, cost, damping) + Solve (LDLT) per problem
/// from host residual blocks.  J: [B][m][n] row-major, r: [B][m], lambda: [B] or nullptr
/// (Gauss-Newton).  Returns per-problem status (0 ok, 1 = SolveLDLT rejected: std::nullopt in the
/// reference).
template <typename Scalar>
std::vector<int32_t> BuildSolve(const Context &ctx, const Scalar *J, const Scalar *r, int64_t B, int m, int n,
                                const Scalar *lambda, Scalar *dx, double *cost) {
  if (B < 0 || m < 0 || n < 1) throw std::invalid_argument("BuildSolve: need B >= 0, m >= 0, n >= 1");
  std::vector<int32_t> status((size_t)B, 0);
  if (B == 0) return status;
  DeviceBuffer<Scalar> dJ(ctx, (size_t)B * m * n), dr(ctx, (size_t)B * m), ddx(ctx, (size_t)B * n);
  DeviceBuffer<Scalar> dl(ctx, lambda ? (size_t)B : 0);
  DeviceBuffer<double> dc(ctx, (size_t)B);
  DeviceBuffer<int32_t> ds(ctx, (size_t)B);
  dJ.upload(J, dJ.size());
  dr.upload(r, dr.size());
  if (lambda) dl.upload(lambda, (size_t)B);
  std::vector<Scalar> zero((size_t)B * n, (Scalar)0);
  ddx.upload(zero.data(), zero.size());
  ctx.check(detail::build_solve(ctx.get(), dJ.data(), dr.data(), B, m, n, lambda ? dl.data() : nullptr, ddx.data(),
                                dc.data(), nullptr, nullptr, ds.data()),
            "tob200_build_solve");
  ddx.download(dx, (size_t)B * n);
  dc.download(cost, (size_t)B);
  ds.download(status.data(), (size_t)B);
  return status;
}

/// `tinyopt::Optimize(x, residuals, options)` (optimize.h:17-77) for a batch of B independent
/// problems with n parameters and m residuals each.  xs: [B][n], updated in place.
/// `residuals(p, x, r, J)`: evaluate problem p at x; fill r[m]; fill J (m x n row-major) unless
/// J == nullptr.  The LM state (x, lambda, H_, grad_, counters) lives on the device; each round
/// trip evaluates the still-running problems on the host and runs one Optimizer_::Step +
/// OptimizeAcc update for all of them (optimizers/optimizer.h:266-309).
template <typename Scalar, typename Residuals>
std::vector<Output> OptimizeBatch(const Context &ctx, Scalar *xs, int64_t B, int n, int m, Residuals &&residuals,
                                  const Options &options = Options()) {
  if (B < 0 || n < 1 || m < 0) throw std::invalid_argument("OptimizeBatch: need B >= 0, n >= 1, m >= 0");
  std::vector<Output> outs((size_t)B);
  if (B == 0) return outs;
  const tob200_options pod = options.pod();
  tob200_solver *solver = nullptr;
  ctx.check(tob200_solver_create(ctx.get(), scalar_traits<Scalar>::dtype, B, n, &pod, &solver), "tob200_solver_create");
  struct Guard {
    tob200_solver *s;
    ~Guard() { tob200_solver_destroy(s); }
  } guard{solver};
  {
    DeviceBuffer<Scalar> x0(ctx, (size_t)B * n);
    x0.upload(xs, (size_t)B * n);
    ctx.check(tob200_solver_reset(solver, x0.data()), "tob200_solver_reset");
    ctx.sync();
  }
  const size_t mm = (size_t)(m > 0 ? m : 1);
  std::vector<Scalar> J((size_t)B * mm * n, (Scalar)0), r((size_t)B * mm, (Scalar)0), x((size_t)B * n);
  std::vector<int32_t> needs((size_t)B);
  DeviceBuffer<Scalar> dJ(ctx, J.size()), dr(ctx, r.size());
  int64_t active = B;
  while (active > 0) {
    ctx.check(tob200_copy_to_host(ctx.get(), x.data(), tob200_solver_x(solver), x.size() * sizeof(Scalar)), "copy x");
    ctx.check(tob200_copy_to_host(ctx.get(), needs.data(), tob200_solver_needs(solver), needs.size() * sizeof(int32_t)),
              "copy needs");
    for (int64_t p = 0; p < B; ++p) {
      if (needs[(size_t)p] < 0) continue;  // finished
      residuals((size_t)p, (const Scalar *)&x[(size_t)p * n], &r[(size_t)p * mm],
                needs[(size_t)p] ? &J[(size_t)p * mm * n] : (Scalar *)nullptr);
    }
    dJ.upload(J.data(), J.size());
    dr.upload(r.data(), r.size());
    ctx.check(detail::solver_step(solver, dJ.data(), dr.data(), m), "tob200_solver_step");
    ctx.check(tob200_solver_num_active(solver, &active), "tob200_solver_num_active");
  }
  DeviceBuffer<tob200_result> dres(ctx, (size_t)B);
  ctx.check(tob200_solver_results(solver, dres.data()), "tob200_solver_results");
  std::vector<tob200_result> res((size_t)B);
  dres.download(res.data(), (size_t)B);
  ctx.check(tob200_copy_to_host(ctx.get(), xs, tob200_solver_x(solver), (size_t)B * n * sizeof(Scalar)), "copy x");
  std::vector<double> H;
  if (options.hessian.save_last) {  // optimizer.h:313-316
    DeviceBuffer<double> dH(ctx, (size_t)B * n * n);
    ctx.check(tob200_solver_final_hessian(solver, dH.data()), "tob200_solver_final_hessian");
    H.resize((size_t)B * n * n);
    dH.download(H.data(), H.size());
  }
  for (int64_t p = 0; p < B; ++p) {
    outs[(size_t)p] = to_output(res[(size_t)p]);
    if (!H.empty()) outs[(size_t)p].final_hessian.assign(H.begin() + (size_t)p * n * n, H.begin() + (size_t)(p + 1) * n * n);
  }
  return outs;
}

/// The MANUAL ACCUMULATION contract of the reference (docs/API.md:37-57,137-170) for a batch:
/// `Cost acc(p, x, grad, H)` is the user's lambda, unchanged in meaning — it fills the solver-owned, pre-zeroed
/// `grad` (n) and `H` (n x n row-major; filling only the upper triangle is enough, docs/API.md:170) and returns
/// the cost (`Cost{cost, num_resisuals}`; a plain scalar cost counts as one residual, cost.h:22).  grad == H ==
/// nullptr is the reference's `nullptr_t` cost-only call (solvers/gn.h:98-105).  The user may fill a true
/// Hessian (tests/optimize_easy.cpp:35-80), a diagonal prior (benchmarks/dense.cpp:57-66), per-block
/// `J^T J` sums (tests/types.cpp:97-108) ... — the device runs Build's tail, damping, Solve, Step and the
/// OptimizeAcc update for every still-running problem (tob200_solver_step_hg_*).  Every n <= 2048 (above n = 55 on the
/// general kernel family).
template <typename Scalar, typename Acc>
std::vector<Output> OptimizeBatchAcc(const Context &ctx, Scalar *xs, int64_t B, int n, Acc &&acc,
                                     const Options &options = Options()) {
  if (B < 0 || n < 1) throw std::invalid_argument("OptimizeBatchAcc: need B >= 0, n >= 1");
  std::vector<Output> outs((size_t)B);
  if (B == 0) return outs;
  const tob200_options pod = options.pod();
  tob200_solver *solver = nullptr;
  ctx.check(tob200_solver_create(ctx.get(), scalar_traits<Scalar>::dtype, B, n, &pod, &solver), "tob200_solver_create");
  struct Guard {
    tob200_solver *s;
    ~Guard() { tob200_solver_destroy(s); }
  } guard{solver};
  {
    DeviceBuffer<Scalar> x0(ctx, (size_t)B * n);
    x0.upload(xs, (size_t)B * n);
    ctx.check(tob200_solver_reset(solver, x0.data()), "tob200_solver_reset");
    ctx.sync();
  }
  const size_t nn = (size_t)n * n;
  std::vector<Scalar> g((size_t)B * n), H((size_t)B * nn), x((size_t)B * n);
  std::vector<double> cost((size_t)B, 0.0);
  std::vector<int32_t> nres((size_t)B, 1), needs((size_t)B);
  DeviceBuffer<Scalar> dg(ctx, g.size()), dH(ctx, H.size());
  DeviceBuffer<double> dc(ctx, cost.size());
  DeviceBuffer<int32_t> dn(ctx, nres.size());
  int64_t active = B;
  while (active > 0) {
    ctx.check(tob200_copy_to_host(ctx.get(), x.data(), tob200_solver_x(solver), x.size() * sizeof(Scalar)), "copy x");
    ctx.check(tob200_copy_to_host(ctx.get(), needs.data(), tob200_solver_needs(solver), needs.size() * sizeof(int32_t)),
              "copy needs");
    for (int64_t p = 0; p < B; ++p) {
      if (needs[(size_t)p] < 0) continue;
      Scalar *gp = nullptr, *Hp = nullptr;
      if (needs[(size_t)p] == 1) {  // Build re-accumulates: the solver's clear() (solvers/gn.h:77-81)
        gp = &g[(size_t)p * n];
        Hp = &H[(size_t)p * nn];
        std::fill(gp, gp + n, (Scalar)0);
        std::fill(Hp, Hp + nn, (Scalar)0);
      }
      const Cost c = acc((size_t)p, (const Scalar *)&x[(size_t)p * n], gp, Hp);
      cost[(size_t)p] = c.cost;
      nres[(size_t)p] = c.num_resisuals;
    }
    dg.upload(g.data(), g.size());
    dH.upload(H.data(), H.size());
    dc.upload(cost.data(), cost.size());
    dn.upload(nres.data(), nres.size());
    ctx.check(detail::solver_step_hg(solver, dg.data(), dH.data(), dc.data(), dn.data()), "tob200_solver_step_hg");
    ctx.check(tob200_solver_num_active(solver, &active), "tob200_solver_num_active");
  }
  DeviceBuffer<tob200_result> dres(ctx, (size_t)B);
  ctx.check(tob200_solver_results(solver, dres.data()), "tob200_solver_results");
  std::vector<tob200_result> res((size_t)B);
  dres.download(res.data(), (size_t)B);
  ctx.check(tob200_copy_to_host(ctx.get(), xs, tob200_solver_x(solver), (size_t)B * n * sizeof(Scalar)), "copy x");
  std::vector<double> Hf;
  if (options.hessian.save_last) {
    DeviceBuffer<double> dHf(ctx, (size_t)B * nn);
    ctx.check(tob200_solver_final_hessian(solver, dHf.data()), "tob200_solver_final_hessian");
    Hf.resize((size_t)B * nn);
    dHf.download(Hf.data(), Hf.size());
  }
  for (int64_t p = 0; p < B; ++p) {
    outs[(size_t)p] = to_output(res[(size_t)p]);
    if (!Hf.empty()) outs[(size_t)p].final_hessian.assign(Hf.begin() + (size_t)p * nn, Hf.begin() + (size_t)(p + 1) * nn);
  }
  return outs;
}

/// The accumulation contract with a SPARSE H (the reference selects its sparse solver from the lambda's H type,
/// optimize.h:27-33; tests/sparse.cpp:19-57): the batch shares one triplet pattern (rows, cols: nnz entries; duplicates
/// are summed in triplet order as Eigen's setFromTriplets does, entries below the diagonal are ignored as
/// `SimplicialLDLT<_, Upper>` ignores them), `Cost acc(p, x, grad, values)` fills grad (n) and the nnz triplet VALUES of
/// its problem (both pre-zeroed; nullptr on cost-only calls).  Solved by the dense pivoted LDL^T of the path
/// (tob200_solver_step_hg_sparse_*).
template <typename Scalar, typename Acc>
std::vector<Output> OptimizeBatchAccSparse(const Context &ctx, Scalar *xs, int64_t B, int n, const std::vector<int32_t> &rows,
                                           const std::vector<int32_t> &cols, Acc &&acc, const Options &options = Options()) {
  if (B < 0 || n < 1 || rows.size() != cols.size()) throw std::invalid_argument("OptimizeBatchAccSparse: bad sizes");
  std::vector<Output> outs((size_t)B);
  if (B == 0) return outs;
  const int nnz = (int)rows.size();
  const tob200_options pod = options.pod();
  tob200_solver *solver = nullptr;
  ctx.check(tob200_solver_create(ctx.get(), scalar_traits<Scalar>::dtype, B, n, &pod, &solver), "tob200_solver_create");
  struct Guard {
    tob200_solver *s;
    ~Guard() { tob200_solver_destroy(s); }
  } guard{solver};
  {
    DeviceBuffer<Scalar> x0(ctx, (size_t)B * n);
    x0.upload(xs, (size_t)B * n);
    ctx.check(tob200_solver_reset(solver, x0.data()), "tob200_solver_reset");
    ctx.sync();
  }
  std::vector<Scalar> g((size_t)B * n), v((size_t)B * (nnz > 0 ? nnz : 1)), x((size_t)B * n);
  std::vector<double> cost((size_t)B, 0.0);
  std::vector<int32_t> nres((size_t)B, 1), needs((size_t)B);
  DeviceBuffer<Scalar> dg(ctx, g.size()), dv(ctx, v.size());
  DeviceBuffer<double> dc(ctx, cost.size());
  DeviceBuffer<int32_t> dn(ctx, nres.size());
  int64_t active = B;
  while (active > 0) {
    ctx.check(tob200_copy_to_host(ctx.get(), x.data(), tob200_solver_x(solver), x.size() * sizeof(Scalar)), "copy x");
    ctx.check(tob200_copy_to_host(ctx.get(), needs.data(), tob200_solver_needs(solver), needs.size() * sizeof(int32_t)),
              "copy needs");
    for (int64_t p = 0; p < B; ++p) {
      if (needs[(size_t)p] < 0) continue;
      Scalar *gp = nullptr, *vp = nullptr;
      if (needs[(size_t)p] == 1) {
        gp = &g[(size_t)p * n];
        vp = &v[(size_t)p * nnz];
        std::fill(gp, gp + n, (Scalar)0);
        std::fill(vp, vp + nnz, (Scalar)0);
      }
      const Cost c = acc((size_t)p, (const Scalar *)&x[(size_t)p * n], gp, vp);
      cost[(size_t)p] = c.cost;
      nres[(size_t)p] = c.num_resisuals;
    }
    dg.upload(g.data(), g.size());
    dv.upload(v.data(), v.size());
    dc.upload(cost.data(), cost.size());
    dn.upload(nres.data(), nres.size());
    ctx.check(detail::solver_step_hg_sparse(solver, dg.data(), rows.data(), cols.data(), nnz, dv.data(), dc.data(), dn.data()),
              "tob200_solver_step_hg_sparse");
    ctx.check(tob200_solver_num_active(solver, &active), "tob200_solver_num_active");
  }
  DeviceBuffer<tob200_result> dres(ctx, (size_t)B);
  ctx.check(tob200_solver_results(solver, dres.data()), "tob200_solver_results");
  std::vector<tob200_result> res((size_t)B);
  dres.download(res.data(), (size_t)B);
  ctx.check(tob200_copy_to_host(ctx.get(), xs, tob200_solver_x(solver), (size_t)B * n * sizeof(Scalar)), "copy x");
  for (int64_t p = 0; p < B; ++p) outs[(size_t)p] = to_output(res[(size_t)p]);
  return outs;
}

/// Output::Covariance(rescaled) (output.h:81-103) for a batch: InvCov (math.h:44-57) of every
/// final_hessian through tob200_inv_cov_f64.  Entry p is the n*n row-major covariance, or empty where
/// the reference returns nullopt (no final Hessian, or the LDLT rejects it).  `rescaled` multiplies by
/// final_cost^2 / (num_residuals - n) when num_residuals > n, as the reference does.
inline std::vector<std::vector<double>> CovarianceBatch(const Context &ctx, const std::vector<Output> &outs, int n,
                                                        bool rescaled = false) {
  const int64_t B = (int64_t)outs.size();
  std::vector<std::vector<double>> covs((size_t)B);
  if (B == 0 || n < 1) return covs;
  const size_t nn = (size_t)n * n;
  std::vector<double> H((size_t)B * nn, 0.0);
  for (int64_t p = 0; p < B; ++p) {
    if (outs[(size_t)p].final_hessian.size() == nn)
      std::copy(outs[(size_t)p].final_hessian.begin(), outs[(size_t)p].final_hessian.end(), H.begin() + (size_t)p * nn);
    else
      for (int i = 0; i < n; ++i) H[(size_t)p * nn + (size_t)i * n + i] = -1.0;  // rejected: not positive
  }
  DeviceBuffer<double> dH(ctx, H.size()), dC(ctx, H.size());
  DeviceBuffer<int32_t> dS(ctx, (size_t)B);
  dH.upload(H.data(), H.size());
  ctx.check(tob200_inv_cov_f64(ctx.get(), dH.data(), B, n, dC.data(), nullptr, dS.data()), "tob200_inv_cov_f64");
  std::vector<int32_t> st((size_t)B);
  dC.download(H.data(), H.size());
  dS.download(st.data(), st.size());
  for (int64_t p = 0; p < B; ++p) {
    const Output &o = outs[(size_t)p];
    if (st[(size_t)p] != 0 || o.final_hessian.size() != nn) continue;
    covs[(size_t)p].assign(H.begin() + (size_t)p * nn, H.begin() + (size_t)(p + 1) * nn);
    if (rescaled && (int)o.num_residuals > n) {
      const double f = o.final_cost.cost * o.final_cost.cost / (double)((int)o.num_residuals - n);
      for (double &v : covs[(size_t)p]) v *= f;
    }
  }
  return covs;
}

/// The device-resident loop for the polynomial residual family r_i = t_i + alpha t_i^3 - y_i,
/// t = A x (tob200_lm_run_host_*): A [B][m][n], y [B][m], xs [B][n] in/out, all host memory.
template <typename Scalar>
std::vector<Output> OptimizePolynomialBatch(const Context &ctx, const Scalar *A, const Scalar *y, Scalar alpha,
                                            Scalar *xs, int64_t B, int m, int n, const Options &options = Options()) {
  std::vector<tob200_result> res((size_t)(B > 0 ? B : 0));
  const tob200_options pod = options.pod();
  ctx.check(detail::lm_run_host(ctx.get(), &pod, A, y, alpha, B, m, n, xs, res.data()), "tob200_lm_run_host");
  std::vector<Output> outs;
  outs.reserve(res.size());
  for (const auto &r : res) outs.push_back(to_output(r));
  return outs;
}

/// One context per visible GPU (or per listed device): the batch of a call is split into CONTIGUOUS shards of
/// independent problems, one per device, solved concurrently (one host thread per device; no data-path
/// communication: the reference has no cross-problem term anywhere).
class MultiContext {
 public:
  explicit MultiContext(const std::vector<int> &devices) {
    for (int d : devices) ctxs_.emplace_back(new Context(d));
    if (ctxs_.empty()) throw std::invalid_argument("MultiContext: no device");
  }
  /// every CUDA device the process can see
  static std::vector<int> all_devices() {
    std::vector<int> d;
    for (int i = 0; i < 64; ++i) {
      tob200_ctx *c = nullptr;
      if (tob200_create(&c, i, nullptr) != TOB200_OK) break;
      tob200_destroy(c);
      d.push_back(i);
    }
    return d;
  }
  int size() const { return (int)ctxs_.size(); }
  const Context &operator[](int i) const { return *ctxs_[(size_t)i]; }
  /// contiguous block [lo, hi) of ceil(B / size()) problems of shard g (last shards may be short or empty)
  void shard(int64_t B, int g, int64_t *lo, int64_t *hi) const {
    const int64_t per = (B + size() - 1) / size();
    *lo = std::min<int64_t>(B, (int64_t)g * per);
    *hi = std::min<int64_t>(B, *lo + per);
  }
  /// run fn(g, ctx, lo, hi) for every shard concurrently; the first exception is rethrown
  template <typename Fn>
  void for_each_shard(int64_t B, Fn &&fn) const {
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> err((size_t)size());
    for (int g = 0; g < size(); ++g) {
      int64_t lo, hi;
      shard(B, g, &lo, &hi);
      if (hi <= lo) continue;
      th.emplace_back([&, g, lo, hi] {
        try {
          fn(g, *ctxs_[(size_t)g], lo, hi);
        } catch (...) {
          err[(size_t)g] = std::current_exception();
        }
      });
    }
    for (auto &t : th) t.join();
    for (auto &e : err)
      if (e) std::rethrow_exception(e);
  }

 private:
  std::vector<std::unique_ptr<Context>> ctxs_;
};

/// OptimizePolynomialBatch over several GPUs: shard g solves problems [lo_g, hi_g) on its own device.
template <typename Scalar>
std::vector<Output> OptimizePolynomialBatch(const MultiContext &mc, const Scalar *A, const Scalar *y, Scalar alpha,
                                            Scalar *xs, int64_t B, int m, int n, const Options &options = Options()) {
  std::vector<Output> outs((size_t)(B > 0 ? B : 0));
  mc.for_each_shard(B, [&](int, const Context &ctx, int64_t lo, int64_t hi) {
    auto part = OptimizePolynomialBatch<Scalar>(ctx, A + (size_t)lo * m * n, y + (size_t)lo * m, alpha, xs + (size_t)lo * n,
                                                hi - lo, m, n, options);
    std::move(part.begin(), part.end(), outs.begin() + lo);
  });
  return outs;
}

/// OptimizeBatch (host residual lambda) over several GPUs; `residuals` is called with GLOBAL problem indices and must
/// be safe to call concurrently for different problems.
template <typename Scalar, typename Residuals>
std::vector<Output> OptimizeBatch(const MultiContext &mc, Scalar *xs, int64_t B, int n, int m, Residuals &&residuals,
                                  const Options &options = Options()) {
  std::vector<Output> outs((size_t)(B > 0 ? B : 0));
  mc.for_each_shard(B, [&](int, const Context &ctx, int64_t lo, int64_t hi) {
    auto shifted = [&, lo](size_t p, const Scalar *x, Scalar *r, Scalar *J) { residuals(p + (size_t)lo, x, r, J); };
    auto part = OptimizeBatch<Scalar>(ctx, xs + (size_t)lo * n, hi - lo, n, m, shifted, options);
    std::move(part.begin(), part.end(), outs.begin() + lo);
  });
  return outs;
}

}  // namespace tinyopt::b200
