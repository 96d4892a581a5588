/*
 * tinyopt_oracle.h — CPU ORACLE for the batched dense-NLLS Levenberg-Marquardt path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under tinyopt_b200/ (the product) may include, link or
 * call this.  Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
 * `--impl reference` leg.
 *
 * What it is: a plain-C restatement of the reference's algorithm for the hot path
 * (SURVEY.md §8a rows a1-a12).  Every function cites the reference file:line it follows
 * (paths relative to /root/reference/include/tinyopt/).
 *
 * PARITY STATUS: the reference itself cannot be compiled in this image (its arithmetic lives in
 * Eigen >= 3.4, un-vendored, absent, no network - cmake/ThirdParties.cmake:16-22).  The oracle is
 * pinned against every known answer the reference's own tests hold for this path
 * (tests/basic.cpp, sqrt2.cpp, solvers.cpp, cov.cpp, types.cpp:94-108, optimize_easy.cpp,
 * optimize_hard.cpp, README.md:91-95 - see tests/test_oracle_reference_cases.py), whose tolerances
 * are 1e-2..1e-7.  At the 1e-10 level against real Eigen output it is "PARITY UNPINNED": Eigen's
 * GEMM / reduction order is implementation-defined and no Eigen build exists here to generate
 * fixtures.  The dense pivoted LDLT follows Eigen 3.4 `LDLT<Matrix,Upper>` semantics as published
 * (SURVEY.md Appendix A).
 *
 * Arithmetic contract ("canonical op sequence"): every sum is a left-to-right
 * chain of IEEE fused multiply-adds in the problem's scalar type, starting from +0; the file is
 * compiled with -ffp-contract=off so nothing else is fused.  The CUDA thread-per-problem kernels
 * implement the same sequence independently, which is why they can be compared bit-for-bit.
 */
#ifndef TINYOPT_ORACLE_H
#define TINYOPT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* stop_reasons.h:14-43 */
enum too_stop_reason {
  TOO_OUT_OF_MEMORY = -4,
  TOO_SOLVER_FAILED = -3,
  TOO_SYSTEM_HAS_NAN_OR_INF = -2,
  TOO_SKIPPED = -1,
  TOO_NONE = 0,
  TOO_MIN_ERROR = 1,
  TOO_MIN_REL_ERROR = 2,
  TOO_MIN_DELTA_NORM = 3,
  TOO_MIN_GRAD_NORM = 4,
  TOO_MAX_ITERS = 5,
  TOO_MAX_NO_DECR = 6,
  TOO_MAX_CONSEC_NO_DECR = 7,
  TOO_TIMED_OUT = 8,
  TOO_USER_STOPPED = 9
};

/* Numeric subset of tinyopt::Options (optimizers/options.h:18-156), same defaults.
 * All thresholds are `float` in the reference and are widened at the point of use. */
typedef struct too_options {
  int32_t solver_type;            /* 0 = LevenbergMarquardt, 1 = GaussNewton (options.h:24-30) */
  int32_t check_final_cost;       /* options.h:43 */
  int32_t use_step_quality_approx;/* options.h:46 */
  float grad_clipping;            /* options.h:49 */
  int32_t use_ldlt;               /* options.h:59 */
  int32_t H_is_full;              /* options.h:61 */
  float check_min_H_diag;         /* options.h:63 */
  int32_t save_last;              /* options.h:66 */
  int32_t use_squared_norm;       /* options.h:76 */
  int32_t downscale_by_2;         /* options.h:77 */
  int32_t normalize;              /* options.h:79 */
  int32_t max_iters;              /* uint16 in the reference, options.h:89 */
  float min_error;                /* options.h:90 */
  float min_rerr_dec;             /* options.h:91 */
  float min_step_norm2;           /* options.h:92 */
  float min_grad_norm2;           /* options.h:93 */
  int32_t max_total_failures;     /* uint8, options.h:94 */
  int32_t max_consec_failures;    /* uint8, options.h:95 */
  float damping_init;             /* options.h:133 */
  float damping_min;              /* options.h:136 [0] */
  float damping_max;              /* options.h:136 [1] */
  float good_factor;              /* options.h:138 */
  float bad_factor;               /* options.h:139 */
} too_options;

/* POD subset of tinyopt::Output (output.h:122-142) */
typedef struct too_result {
  double final_cost;          /* Output::final_cost.cost */
  double final_rerr_dec;      /* Output::final_rerr_dec */
  int32_t final_num_residuals;/* Output::final_cost.num_resisuals */
  int32_t stop_reason;        /* Output::stop_reason */
  int32_t num_iters;          /* uint16 in the reference */
  int32_t num_failures;       /* uint8 in the reference (wraps the same way here) */
  int32_t num_consec_failures;/* uint8 */
  int32_t history_len;        /* errs.size() == deltas2.size() == successes.size() */
  double last_lambda;         /* SolverLM::lambda_ at exit */
  double last_prev_lambda;    /* SolverLM::prev_lambda_ at exit */
  double min_margin;          /* smallest relative distance of any branch decision from its
                                 threshold (diagnostic: how fragile the iteration count is) */
  double sign_margin;         /* ... of the accept/reject decisions only: min |derr| / |err| */
  double thr_margin;          /* ... of the stop-threshold decisions only: min |v - thr| / thr */
} too_result;

/* Optional per-iteration trace (caller allocates `cap` entries of each non-NULL array). */
typedef struct too_trace {
  int32_t cap;
  double *errs;        /* Output::errs */
  double *deltas2;     /* Output::deltas2 */
  int32_t *successes;  /* Output::successes */
  double *lambdas;     /* lambda_ after the Step */
  int32_t *rebuilt;    /* rebuild_linear_system_ used by the Step's Build */
  double *xs;          /* x after the OptimizeAcc update, n per iteration (or NULL) */
} too_trace;

void too_options_default(too_options *o);

/* Accumulation callback == the reference's `acc(x, grad, H)` contract (docs/API.md:37-57).
 * grad == NULL && H == NULL  <=> the nullptr_t cost-only call.  H is n*n row-major and arrives
 * zeroed (gn.h:77-81); only the upper triangle is read back when use_ldlt.  Return through
 * cost / num_res (Cost, cost.h:18-25). */
typedef void (*too_acc_f64)(const double *x, int n, double *grad, double *H, double *cost,
                            int *num_res, void *user);
typedef void (*too_acc_f32)(const float *x, int n, float *grad, float *H, double *cost,
                            int *num_res, void *user);

/* math.h:232-240 SolveLDLT == Eigen LDLT<_,Upper>: returns 1 iff info()==Success && isPositive(),
 * in which case x = A^-1 b.  A: n*n row-major, only i<=j read. */
int too_solve_ldlt_f64(int n, const double *A, const double *b, double *x);
int too_solve_ldlt_f32(int n, const float *A, const float *b, float *x);

/* math.h:44-57 DenseInvCov: inverse through the same factorisation. returns 1 on success. */
int too_inv_cov_f64(int n, const double *A, double *Ainv);
int too_inv_cov_f32(int n, const float *A, float *Ainv);

/* One Build + Solve from a materialised residual block (a1 + a3 + a5 + a6):
 *   g = J^T r, H = J^T J (diff/optimize_autodiff.h:151-157), cost = |r|^2 (:164),
 *   H_ii *= (1+lambda) (solvers/lm.h:108-117), dx = SolveLDLT(H, -g) (solvers/gn.h:150-156).
 * J is m*n row-major (row i = d r_i / d x).  H_out (n*n row-major, damped, full symmetric) and
 * g_out may be NULL.  Returns status: 0 ok, 1 solve failed (not positive / numerical issue). */
int too_build_solve_f64(int m, int n, const double *J, const double *r, double lambda, double *dx,
                        double *cost, double *H_out, double *g_out);
int too_build_solve_f32(int m, int n, const float *J, const float *r, float lambda, float *dx,
                        double *cost, float *H_out, float *g_out);

/* optimizers/optimizer.h:243-327 OptimizeAcc with SolverLM / SolverGN.  x is updated in place.
 * final_hessian: n*n row-major doubles or NULL (optimizer.h:313-316, lm.h:157-171). */
int too_optimize_f64(double *x, int n, too_acc_f64 acc, void *user, const too_options *opt,
                     too_result *out, too_trace *trace, double *final_hessian);
int too_optimize_f32(float *x, int n, too_acc_f32 acc, void *user, const too_options *opt,
                     too_result *out, too_trace *trace, double *final_hessian);

/* ---- synthetic problem family ------------------------------------------------
 * problem p: A (m*n), truth x*, y = t* + alpha t*^3 + sigma z, x0 = x* + 0.3 u;
 * r_i(x) = t_i + alpha t_i^3 - y_i, t = A x.  All arrays problem-major: A[p][i][j], y[p][i],
 * x[p][j].  Any output pointer may be NULL. */
void too_synth_generate_f64(uint64_t seed, int64_t p0, int64_t B, int m, int n, double alpha,
                            double sigma, double *A, double *y, double *xstar, double *x0);
void too_synth_generate_f32(uint64_t seed, int64_t p0, int64_t B, int m, int n, float alpha,
                            float sigma, float *A, float *y, float *xstar, float *x0);

/* residual + Jacobian of the family at x for one problem (J m*n row-major, may be NULL) */
void too_synth_eval_f64(int m, int n, const double *A, const double *y, double alpha,
                        const double *x, double *r, double *J);
void too_synth_eval_f32(int m, int n, const float *A, const float *y, float alpha, const float *x,
                        float *r, float *J);

/* Full LM run over a batch of the family; `nthreads` OpenMP threads over problems
 * (<=0: all available).  x: [B][n] in/out, results: [B].  Returns the threads used. */
int too_synth_lm_run_f64(int64_t B, int m, int n, const double *A, const double *y, double alpha,
                         double *x, const too_options *opt, too_result *results, int nthreads);
int too_synth_lm_run_f32(int64_t B, int m, int n, const float *A, const float *y, float alpha,
                         float *x, const too_options *opt, too_result *results, int nthreads);

/* the same, also returning Output::final_hessian per problem ([B][n][n] doubles, optimizer.h:313-316) */
int too_synth_lm_run_fh_f64(int64_t B, int m, int n, const double *A, const double *y, double alpha,
                            double *x, const too_options *opt, too_result *results, int nthreads,
                            double *final_hessian);
int too_synth_lm_run_fh_f32(int64_t B, int m, int n, const float *A, const float *y, float alpha,
                            float *x, const too_options *opt, too_result *results, int nthreads,
                            double *final_hessian);

/* Robust variant of the family (losses/robust_norms.h:16-29): while kind != 0 every residual of
 * too_synth_lm_run_* goes through M-estimator `kind` (1 Truncated, 2 Huber, 3 Tukey, 4 Arctan, 5 Cauchy,
 * 6 GemanMcClure, 7 BlakeZisserman) with squared threshold th2:  cost += loss, grad += J^T r * scale, H += J^T J. */
void too_synth_set_robust(int kind, double th2);
/* Numeric-differentiation variant (diff/num_diff.h:57-126 NumEval, :284-309 CreateNumDiffFunc2): while method != 0 the
 * family's Jacobian is estimated column by column from the residual function (1 kForward, 2 kCentral - the
 * reference's default -, 3 kFastCentral; h <= 0: FloatEpsilon<Scalar>()), grad = J^T res, H = J^T J, and the Cost is
 * the residual NORM with m residuals, as the reference's lambda returns it.  Global state like set_robust. */
void too_synth_set_numdiff(int method, double h);

int too_max_threads(void);

/* Decision-margin census only: rev != 0 makes too_synth_lm_run_* accumulate the residual rows in
 * reverse order (a different, equally valid summation order - what Eigen's GEMM is to this
 * restatement).  Never set by tests of the canonical sequence. */
void too_census_set_row_order(int rev);

#ifdef __cplusplus
}
#endif
#endif
