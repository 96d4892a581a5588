// lm_state.cuh — per-problem Levenberg-Marquardt state machine, device side.
//
// One `LmState` is one `Optimizer_<SolverLM<...>>` + its `Output` for one problem
// (include/tinyopt/optimizers/optimizer.h:243-539, solvers/lm.h:37-171, solvers/gn.h:77-171,
// solvers/base.h:30-45).  The data pass (accumulating H, g, cost over the residual rows) is done
// by the caller kernel; `lm_after_pass` is everything the reference does with the accumulated
// system: Build's tail (normalise, validity, clamp, diagonal check, damping), the solve-retry loop,
// Step's accept / reject bookkeeping and stop tests, and OptimizeAcc's x update / roll-back.
//
// Storage-agnostic: the solver's persistent H_ (damped, packed upper triangle) and grad_ are
// reached through an accessor object (`HG`) with ld_h(i)/st_h(i,v)/ld_g(j)/st_g(j,v), so the
// thread-per-problem kernels keep them in shared memory and the step kernels in global memory.
#pragma once

#include <math_constants.h>

#include "../../include/tinyopt_b200.h"
#include "common.cuh"
#include "ldlt_reg.cuh"

namespace tob200 {

// device copy of tob200_options, thresholds pre-widened exactly as the reference widens them
template <typename T>
struct DevOptions {
  int solver_type, check_final_cost, use_step_quality_approx, use_ldlt;
  int use_squared_norm, downscale_by_2, normalize;
  int max_iters, max_total_failures, max_consec_failures;
  T grad_clipping, check_min_H_diag;
  T damping_init, damping_min, damping_max, good_factor, bad_factor;
  double min_error, min_rerr_dec, min_step_norm2, min_grad_norm2;  // (double)float
  float min_error_f, min_rerr_dec_f, min_step_norm2_f, min_grad_norm2_f;
};

template <typename T>
inline DevOptions<T> make_dev_options(const tob200_options &o) {
  DevOptions<T> d;
  d.solver_type = o.solver_type;
  d.check_final_cost = o.check_final_cost;
  d.use_step_quality_approx = o.use_step_quality_approx;
  d.use_ldlt = o.use_ldlt;
  d.use_squared_norm = o.use_squared_norm;
  d.downscale_by_2 = o.downscale_by_2;
  d.normalize = o.normalize;
  d.max_iters = o.max_iters;
  d.max_total_failures = o.max_total_failures;
  d.max_consec_failures = o.max_consec_failures;
  d.grad_clipping = (T)o.grad_clipping;
  d.check_min_H_diag = (T)o.check_min_H_diag;
  d.damping_init = (T)o.damping_init;
  d.damping_min = (T)o.damping_min;
  d.damping_max = (T)o.damping_max;
  d.good_factor = (T)o.good_factor;
  d.bad_factor = (T)o.bad_factor;
  d.min_error = (double)o.min_error;
  d.min_rerr_dec = (double)o.min_rerr_dec;
  d.min_step_norm2 = (double)o.min_step_norm2;
  d.min_grad_norm2 = (double)o.min_grad_norm2;
  d.min_error_f = o.min_error;
  d.min_rerr_dec_f = o.min_rerr_dec;
  d.min_step_norm2_f = o.min_step_norm2;
  d.min_grad_norm2_f = o.min_grad_norm2;
  return d;
}

// flag bits of LmState::flags
enum : uint32_t {
  kFlagRebuild = 1u,         // SolverLM::rebuild_linear_system_ (solvers/lm.h:194)
  kFlagHasLastDx = 2u,       // `last_dx` engaged (optimizer.h:262)
  kFlagLastWasSuccess = 4u,  // optimizer.h:263
  kFlagDone = 8u,            // the OptimizeAcc loop has exited for this problem
};

// Scalar part of one `Optimizer_<SolverLM<...>>` + its `Output` (what does not depend on n).
template <typename T>
struct LmScalars {
  T lambda, prev_lambda, bad_factor;  // solvers/lm.h:191-193
  double final_cost;                  // Output::final_cost.cost (output.h:122)
  double final_rerr_dec;              // output.h:123
  int final_nres;
  int stop_reason;
  uint32_t flags;
  uint16_t num_iters;                         // output.h:133
  uint8_t num_failures, num_consec_failures;  // output.h:134-136 (uint8 wrap-around kept)
  int num_builds;
  int iter;  // OptimizeAcc's loop counter (== num_iters while running)

  __device__ __forceinline__ void reset_scalars(const DevOptions<T> &o) {
    lambda = o.damping_init;  // lm.h:46-52
    prev_lambda = (T)0;
    bad_factor = o.bad_factor;
    final_cost = 1.7976931348623157e+308;  // output.h:122
    final_rerr_dec = 1.7976931348623157e+308;
    final_nres = 0;
    stop_reason = TOB200_STOP_NONE;
    flags = kFlagRebuild | kFlagLastWasSuccess;
    num_iters = 0;
    num_failures = 0;
    num_consec_failures = 0;
    num_builds = 0;
    iter = 0;
  }
  __device__ __forceinline__ bool done() const { return flags & kFlagDone; }
  __device__ __forceinline__ bool rebuild() const { return flags & kFlagRebuild; }
};

// thread-per-problem state: scalars + x and last_dx in registers
template <typename T, int N>
struct LmState : LmScalars<T> {
  T x[N];
  T last_dx[N];
  __device__ __forceinline__ void reset(const DevOptions<T> &o) {
    this->reset_scalars(o);
#pragma unroll
    for (int j = 0; j < N; ++j) last_dx[j] = (T)0;
  }
};

template <typename T>
__device__ __forceinline__ T clamp_t(T v, T lo, T hi) {  // std::clamp
  return v < lo ? lo : (hi < v ? hi : v);
}

// solvers/lm.h:123-137
template <typename T>
__device__ __forceinline__ void lm_good_step(LmScalars<T> &s, const DevOptions<T> &o, T quality) {
  if (o.solver_type != 0) return;  // base.h:52: no-op for Gauss-Newton
  T sc = o.good_factor;
  if (quality != (T)0) {
    // `1.0f - std::pow(2.0f * quality - 1.0f, 3.0f)` in Scalar; cube written out (off by default;
    // tolerance-level, not bit-level, agreement with libm's pow)
    const T b = Ops<T>::sub(Ops<T>::mul((T)2, quality), (T)1);
    const T q = Ops<T>::sub((T)1, Ops<T>::mul(Ops<T>::mul(b, b), b));
    sc = sc > q ? sc : q;
  }
  if (s.bad_factor != o.bad_factor) sc = Ops<T>::div(sc, s.bad_factor);
  s.prev_lambda = s.lambda;
  s.lambda = clamp_t(Ops<T>::mul(s.lambda, sc), o.damping_min, o.damping_max);
  s.bad_factor = o.bad_factor;
}

// solvers/lm.h:140-148 (FailedStep == BadStep)
template <typename T>
__device__ __forceinline__ void lm_bad_step(LmScalars<T> &s, const DevOptions<T> &o) {
  if (o.solver_type != 0) return;
  const T sc = s.bad_factor;
  s.prev_lambda = s.lambda;
  s.lambda = clamp_t(Ops<T>::mul(s.lambda, sc), o.damping_min, o.damping_max);
  s.bad_factor = Ops<T>::mul(s.bad_factor, o.bad_factor);
}

__device__ __forceinline__ bool is_nan_or_inf(double v) { return isnan(v) || isinf(v); }

// ---- the storage-independent pieces of one OptimizeAcc iteration -------------------------------
// They are shared by the thread-per-problem kernels (vectors in registers) and the warp-per-problem
// kernels (vectors in shared memory), so both run the very same decision sequence.

// NormalizeCost (base.h:41-45) + Cost::isValid (cost.h:83)
template <typename T>
__device__ __forceinline__ bool lm_normalize_cost(const DevOptions<T> &o, T cost_t, int nres, double &cost) {
  cost = (double)cost_t;
  if (!o.use_squared_norm) cost = sqrt(cost);
  if (o.downscale_by_2) cost *= 0.5f;
  if (o.normalize && nres > 0) cost /= nres;
  return nres > 0 && cost != 1.7976931348623157e+308;
}

// the same for a cost that arrives as the reference's `Cost::cost` itself (a double: user-filled accumulators,
// tob200_solver_step_hg_*)
template <typename T>
__device__ __forceinline__ bool lm_normalize_cost_d(const DevOptions<T> &o, double cost_in, int nres, double &cost) {
  cost = cost_in;
  if (!o.use_squared_norm) cost = sqrt(cost);
  if (o.downscale_by_2) cost *= 0.5f;
  if (o.normalize && nres > 0) cost /= nres;
  return nres > 0 && cost != 1.7976931348623157e+308;
}

// optimizer.h:356-357
template <typename T>
__device__ __forceinline__ uint8_t lm_max_tries(const DevOptions<T> &o) {
  return o.max_consec_failures > 0 ? (uint8_t)(o.max_consec_failures > 1 ? o.max_consec_failures : 1) : 255;
}

// Build's damping factor (lm.h:108-117): returns false when no damping applies
template <typename T>
__device__ __forceinline__ bool lm_damping_scale(const LmScalars<T> &s, const DevOptions<T> &o, bool pass_rebuilt,
                                                 double &sc) {
  if (o.solver_type != 0 || !(s.lambda > (T)0)) return false;
  sc = pass_rebuilt ? 1.0 + (double)s.lambda : (1.0 + (double)s.lambda) / (1.0 + (double)s.prev_lambda);
  return true;
}

enum LmFailureAction { kLmRetry = 0, kLmBreak = 1, kLmEarlyReturn = 2 };

// One failed Build-or-Solve inside Step's retry loop (optimizer.h:370-392)
template <typename T>
__device__ __forceinline__ int lm_on_solver_failure(LmScalars<T> &s, const DevOptions<T> &o, double cost, int nres) {
  s.num_consec_failures++;
  s.num_failures++;
  if (nres == 0) {  // :374-377
    s.stop_reason = TOB200_STOP_SKIPPED;
    return kLmEarlyReturn;
  }
  if (is_nan_or_inf(cost)) {  // :378-381
    s.stop_reason = TOB200_STOP_SYSTEM_HAS_NAN_OR_INF;
    return kLmEarlyReturn;
  }
  if (o.max_consec_failures > 0 && s.num_consec_failures >= o.max_consec_failures) {  // :382-386
    if (s.final_cost < (double)Ops<T>::max_value()) s.stop_reason = TOB200_STOP_MAX_CONSEC_NO_DECR;
    return kLmBreak;
  }
  lm_bad_step(s, o);  // FailedStep (:389)
  return kLmRetry;
}

// The rest of Step once the retry loop is over (optimizer.h:395-538): NaN guards, accept / reject,
// lambda schedule, failure counters, stop tests.  Returns through success / has_dx the pair Step
// hands back to OptimizeAcc.
template <typename T>
__device__ __forceinline__ void lm_finish_step(LmScalars<T> &s, const DevOptions<T> &o, bool early_return,
                                               bool solver_failed, double cost, int nres, double dx_norm2,
                                               double grad_norm2, bool &success, bool &has_dx) {
  using O = Ops<T>;
  success = false;
  has_dx = false;
  const int iter = s.num_iters;
  if (early_return) return;  // stop_reason already set; status = {false, nullopt}
  if (solver_failed) {       // :396-399
    s.stop_reason = TOB200_STOP_SOLVER_FAILED;
    return;
  }
  const double err = cost;
  if (is_nan_or_inf(err) || is_nan_or_inf(dx_norm2)) {  // :405-409, :416-425
    s.stop_reason = TOB200_STOP_SYSTEM_HAS_NAN_OR_INF;
    return;
  }
  const double derr = err - s.final_cost;  // :428
  const bool is_good_step = derr < 0.0;    // :429
  const double rel_derr = (s.final_cost > (double)O::float_eps() && s.final_cost < (double)O::max_value())
                              ? (s.final_cost - err) / s.final_cost
                              : 0.0;  // :431-434
  if (is_good_step || iter == 0) {  // :441-446
    if (iter > 0) lm_good_step(s, o, o.use_step_quality_approx ? (T)rel_derr : (T)0);
    s.num_consec_failures = 0;
    s.final_cost = cost;
    s.final_nres = nres;
    s.final_rerr_dec = rel_derr;
  } else {  // :447-460
    lm_bad_step(s, o);
    s.num_failures++;
    s.num_consec_failures++;
    if (o.max_consec_failures > 0 && s.num_consec_failures >= o.max_consec_failures) {
      s.stop_reason = TOB200_STOP_MAX_CONSEC_NO_DECR;
      return;
    }
    if (o.max_total_failures > 0 && s.num_failures >= o.max_total_failures) {
      s.stop_reason = TOB200_STOP_MAX_NO_DECR;
      return;
    }
  }
  // :518-528
  if (o.min_error_f > 0 && err < o.min_error) s.stop_reason = TOB200_STOP_MIN_ERROR;
  else if (o.min_rerr_dec_f > 0 && rel_derr > 0.0 && rel_derr < o.min_rerr_dec) s.stop_reason = TOB200_STOP_MIN_REL_ERROR;
  else if (o.min_step_norm2_f > 0 && dx_norm2 < o.min_step_norm2) s.stop_reason = TOB200_STOP_MIN_DELTA_NORM;
  else if (o.min_grad_norm2_f > 0 && grad_norm2 < o.min_grad_norm2) s.stop_reason = TOB200_STOP_MIN_GRAD_NORM;
  success = is_good_step;  // :536-538
  has_dx = true;
}

enum LmUpdateAction { kLmNoMove = 0, kLmApplyDx = 1, kLmRollBack = 2, kLmProbeDx = 3 };

// OptimizeAcc's reaction to Step's result (optimizer.h:269-309, 320-321): which move to apply to x
// (kLmApplyDx / kLmProbeDx: x += dx and last_dx = dx; kLmRollBack: x -= last_dx), the rebuild flag
// of the next Build, the iteration counter and the loop exit.
template <typename T>
__device__ __forceinline__ int lm_update_action(LmScalars<T> &s, const DevOptions<T> &o, bool success, bool has_dx) {
  const int max_iters = o.max_iters + 1 + (o.check_final_cost ? 1 : 0);  // :248-250
  bool eval_only = false;
  int action = kLmNoMove;
  if (success) {  // :271-279
    action = kLmApplyDx;
    s.flags |= kFlagHasLastDx | kFlagLastWasSuccess;
    if (o.check_final_cost && s.iter + 1 == max_iters) eval_only = true;
  } else {  // :281-297
    if (s.flags & kFlagHasLastDx) {
      action = kLmRollBack;
      s.flags &= ~kFlagHasLastDx;
    } else if (has_dx) {
      action = kLmProbeDx;
      s.flags |= kFlagHasLastDx;
    }
    eval_only = !(s.flags & kFlagLastWasSuccess);
    s.flags &= ~kFlagLastWasSuccess;
  }
  if (eval_only) s.flags &= ~kFlagRebuild;  // :299 solver_.Rebuild(!eval_only)
  else s.flags |= kFlagRebuild;
  s.num_iters++;  // :307
  s.iter++;
  if (s.stop_reason != TOB200_STOP_NONE) {
    s.flags |= kFlagDone;  // :309
  } else if (s.iter >= max_iters) {
    s.stop_reason = TOB200_STOP_MAX_ITERS;  // :320-321
    s.flags |= kFlagDone;
  }
  return action;
}

// ---- thread-per-problem: everything after the data pass of one OptimizeAcc iteration -----------
//   pass_rebuilt : the pass accumulated H and g (true) or only the cost (false)
//   hu, g        : undamped accumulated upper triangle / gradient (valid iff pass_rebuilt)
//   cost_t       : sum r^2 accumulated in T, nres : number of residuals the pass saw
//   hg           : persistent H_ / grad_ storage of the solver
template <typename T, int N, class HG>
__device__ __forceinline__ void lm_after_pass(LmState<T, N> &s, const DevOptions<T> &o,
                                              bool pass_rebuilt, T (&hu)[tri_count(N)], T (&g)[N],
                                              T cost_t, int nres, HG &hg, const double *cost_d = nullptr) {
  using O = Ops<T>;
  using L = LdltReg<T, N>;
  constexpr int NT = tri_count(N);

  // ---- Build, first attempt: cost_ = acc(...); NormalizeCost ----
  double cost;
  // (cost_d: the caller's accumulation lambda returned the cost itself, as a double: docs/API.md:37-57)
  bool built_ok = cost_d ? lm_normalize_cost_d(o, *cost_d, nres, cost) : lm_normalize_cost(o, cost_t, nres, cost);
  T diag0[N];  // undamped diagonal, for the re-damping of a retry after a rebuild
  if (pass_rebuilt) {
    s.num_builds++;
    if (built_ok) {
      if (o.grad_clipping != (T)0) {  // base.h:30-38
#pragma unroll
        for (int j = 0; j < N; ++j) {
          T v = g[j];
          v = v < -o.grad_clipping ? -o.grad_clipping : v;
          v = v > o.grad_clipping ? o.grad_clipping : v;
          g[j] = v;
        }
      }
      if (o.check_min_H_diag > (T)0) {  // lm.h:82-86
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (O::abs(hu[tri_index(N, j, j)]) < o.check_min_H_diag) built_ok = false;
      }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      diag0[j] = hu[tri_index(N, j, j)];
      hg.st_g(j, g[j]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      g[j] = hg.ld_g(j);
      diag0[j] = (T)0;
    }
  }

  // ---- solve-retry loop (optimizer.h:354-399) ----
  bool solver_failed = true, early_return = false;
  T dx[N];
  int tr[N];
  const uint8_t max_tries = lm_max_tries(o);
  for (int attempt = 0; s.num_consec_failures <= max_tries; ++attempt) {
    // A failed Build (invalid cost / diagonal check) returns before the damping, exactly like the
    // early `return false`s of lm.h:72-76,85,102.
    if (built_ok) {
      // `hu` is factorised in place below, so a retry (and every cost-only pass) starts from the
      // persistent copy of H_ instead of keeping a second triangle alive in registers
      if (!pass_rebuilt || attempt > 0) {
#pragma unroll
        for (int i = 0; i < NT; ++i) hu[i] = hg.ld_h(i);
      }
      double sc;
      if (lm_damping_scale(s, o, pass_rebuilt, sc)) {
        // re-accumulating at the same x gives the same H, so a retry re-damps the saved diagonal
#pragma unroll
        for (int j = 0; j < N; ++j)
          hu[tri_index(N, j, j)] = (T)((double)(pass_rebuilt ? diag0[j] : hu[tri_index(N, j, j)]) * sc);
      } else if (pass_rebuilt) {
#pragma unroll
        for (int j = 0; j < N; ++j) hu[tri_index(N, j, j)] = diag0[j];
      }
      // H_ now holds the damped matrix: publish it to the persistent storage
      if (pass_rebuilt && attempt == 0) {
#pragma unroll
        for (int i = 0; i < NT; ++i) hg.st_h(i, hu[i]);
      } else {
#pragma unroll
        for (int j = 0; j < N; ++j) hg.st_h(tri_index(N, j, j), hu[tri_index(N, j, j)]);
      }
      // Solve (gn.h:150-156)
      if (o.use_ldlt) {
        if (L::factor(hu, tr)) {
#pragma unroll
          for (int j = 0; j < N; ++j) dx[j] = -g[j];
          L::solve(hu, tr, dx);
          solver_failed = false;
        }
      } else {  // gn.h:157-163: -H^-1 g, never fails
        solve_inverse_reg<T, N>(hu, g, dx);
        solver_failed = false;
      }
    }
    if (!solver_failed) break;
    const int act = lm_on_solver_failure(s, o, cost, nres);
    if (act == kLmEarlyReturn) early_return = true;
    if (act != kLmRetry) break;
    // the reference would retry forever when max_consec_failures == 0 and H never becomes
    // positive; give up after 100000 retries like the oracle does
    if (attempt >= 100000) break;
  }

  double dx_norm2 = 0.0, grad_norm2 = 0.0;
  if (!solver_failed) {
    T dn = (T)0;
#pragma unroll
    for (int j = 0; j < N; ++j) dn = O::fma(dx[j], dx[j], dn);
    dx_norm2 = (double)dn;  // optimizer.h:412
    if (o.min_grad_norm2_f > 0.0f) {  // :413-415
      T gn = (T)0;
#pragma unroll
      for (int j = 0; j < N; ++j) gn = O::fma(g[j], g[j], gn);
      grad_norm2 = (double)gn;
    }
  }
  bool success, has_dx;
  lm_finish_step(s, o, early_return, solver_failed, cost, nres, dx_norm2, grad_norm2, success, has_dx);

  // ---- OptimizeAcc's update (traits.h:162,184-190 PlusEq) ----
  const int action = lm_update_action(s, o, success, has_dx);
  if (action == kLmApplyDx || action == kLmProbeDx) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      s.x[j] = O::add(s.x[j], dx[j]);
      s.last_dx[j] = dx[j];
    }
  } else if (action == kLmRollBack) {
#pragma unroll
    for (int j = 0; j < N; ++j) s.x[j] = O::add(s.x[j], -s.last_dx[j]);
  }
}

template <typename T>
__device__ __forceinline__ void lm_write_result(const LmScalars<T> &s, tob200_result *r) {
  r->final_cost = s.final_cost;
  r->final_rerr_dec = s.final_rerr_dec;
  r->last_lambda = (double)s.lambda;
  r->last_prev_lambda = (double)s.prev_lambda;
  r->final_num_residuals = s.final_nres;
  r->stop_reason = s.stop_reason;
  r->num_iters = s.num_iters;
  r->num_failures = s.num_failures;
  r->num_consec_failures = s.num_consec_failures;
  r->num_builds = s.num_builds;
}

}  // namespace tob200
