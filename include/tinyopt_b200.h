/*
 * tinyopt_b200.h — C-ABI of the B200-native batched dense-NLLS Levenberg-Marquardt hot path.
 *
 * The reference (julien-michot/tinyopt, header-only C++ on Eigen) has no FFI of its own; its
 * extension seam is the `SolverType` concept consumed by `Optimizer_<SolverType>`
 * (include/tinyopt/optimizers/optimizer.h:34-43).  Every entry point below names the reference
 * interface it replaces (paths relative to /root/reference/include/tinyopt/).  INTEGRATION.md shows
 * the binding a tinyopt maintainer would add on their side.
 *
 * Conventions: plain pointers and sizes only; return 0 (TOB200_OK) or a negative tob200_status;
 * nothing throws across the ABI; all data pointers are DEVICE pointers owned by the caller unless
 * the name says `_host`; calls are asynchronous on the context's stream (tob200_sync to wait);
 * one context per GPU, a context is not thread-safe, different contexts are independent.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * TOB200_ERR_CUDA.
 */
#ifndef TINYOPT_B200_H
#define TINYOPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOB200_VERSION 100 /* 0.1.0 */

typedef enum tob200_status {
  TOB200_OK = 0,
  TOB200_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, misaligned buffer) */
  TOB200_ERR_CUDA = -2,        /* CUDA runtime error (no device, launch failure, ...) */
  TOB200_ERR_UNSUPPORTED = -3, /* shape / dtype has no kernel (see tob200_kernel_family) */
  TOB200_ERR_NOMEM = -4        /* device allocation failed (reference: kOutOfMemory) */
} tob200_status;

typedef enum tob200_dtype { TOB200_F32 = 0, TOB200_F64 = 1 } tob200_dtype;

/* Memory layout of a batch of residual blocks J (m x n per problem) and r (m per problem).
 *   TILE32        : problems interleaved in tiles of 32 — J[tile][i][j][lane], r[tile][i][lane],
 *                   problem p = 32*tile + lane.  Buffers hold ceil(B/32) full tiles
 *                   (tob200_tiled_elems); pad lanes are never read back.  Native layout: one
 *                   contiguous TMA bulk copy per row chunk, conflict-free shared-memory reads.
 *   PROBLEM_MAJOR : J[p][i][j] (row i = d r_i / d x, the rows `J.row(i) = res[i].v` gathers in
 *                   diff/optimize_autodiff.h:127-148), r[p][i].  Re-tiled on the device first. */
typedef enum tob200_layout { TOB200_LAYOUT_TILE32 = 0, TOB200_LAYOUT_PROBLEM_MAJOR = 1 } tob200_layout;

/* stop_reasons.h:14-43 (same values) */
typedef enum tob200_stop_reason {
  TOB200_STOP_OUT_OF_MEMORY = -4,
  TOB200_STOP_SOLVER_FAILED = -3,
  TOB200_STOP_SYSTEM_HAS_NAN_OR_INF = -2,
  TOB200_STOP_SKIPPED = -1,
  TOB200_STOP_NONE = 0,
  TOB200_STOP_MIN_ERROR = 1,
  TOB200_STOP_MIN_REL_ERROR = 2,
  TOB200_STOP_MIN_DELTA_NORM = 3,
  TOB200_STOP_MIN_GRAD_NORM = 4,
  TOB200_STOP_MAX_ITERS = 5,
  TOB200_STOP_MAX_NO_DECR = 6,
  TOB200_STOP_MAX_CONSEC_NO_DECR = 7,
  TOB200_STOP_TIMED_OUT = 8,
  TOB200_STOP_USER_STOPPED = 9
} tob200_stop_reason;

/* POD mirror of the numeric subset of tinyopt::Options (optimizers/options.h:18-156), same
 * defaults (tob200_options_default).  Thresholds are `float` as in the reference and are widened
 * at the point of use exactly where the reference widens them. */
typedef struct tob200_options {
  int32_t solver_type;             /* options.h:24-30: 0 LevenbergMarquardt, 1 GaussNewton */
  int32_t check_final_cost;        /* options.h:43 */
  int32_t use_step_quality_approx; /* options.h:46 */
  float grad_clipping;             /* options.h:49 */
  int32_t use_ldlt;                /* options.h:59; 0 = H.inverse() path (gn.h:157-163): every n <= 2048 */
  int32_t H_is_full;               /* options.h:61 */
  float check_min_H_diag;          /* options.h:63 */
  int32_t save_last;               /* options.h:66 */
  int32_t use_squared_norm;        /* options.h:76 */
  int32_t downscale_by_2;          /* options.h:77 */
  int32_t normalize;               /* options.h:79 */
  int32_t max_iters;               /* options.h:89 */
  float min_error;                 /* options.h:90 */
  float min_rerr_dec;              /* options.h:91 */
  float min_step_norm2;            /* options.h:92 */
  float min_grad_norm2;            /* options.h:93 */
  int32_t max_total_failures;      /* options.h:94 */
  int32_t max_consec_failures;     /* options.h:95 */
  float damping_init;              /* options.h:133 */
  float damping_min;               /* options.h:136 */
  float damping_max;               /* options.h:136 */
  float good_factor;               /* options.h:138 */
  float bad_factor;                /* options.h:139 */
} tob200_options;

/* POD subset of tinyopt::Output (output.h:122-142), one per problem. */
typedef struct tob200_result {
  double final_cost;           /* Output::final_cost.cost */
  double final_rerr_dec;       /* Output::final_rerr_dec */
  double last_lambda;          /* SolverLM::lambda_ at exit (solvers/lm.h:191) */
  double last_prev_lambda;     /* SolverLM::prev_lambda_ at exit (solvers/lm.h:192) */
  int32_t final_num_residuals; /* Output::final_cost.num_resisuals */
  int32_t stop_reason;         /* Output::stop_reason, tob200_stop_reason */
  int32_t num_iters;           /* Output::num_iters (Step calls) */
  int32_t num_failures;        /* Output::num_failures */
  int32_t num_consec_failures; /* Output::num_consec_failures */
  int32_t num_builds;          /* passes that rebuilt H and g (the rest were cost-only) */
} tob200_result;

typedef struct tob200_ctx tob200_ctx;
typedef struct tob200_solver tob200_solver;

/* ---- context ---------------------------------------------------------------------------------- */
int tob200_version(void);
/* device: CUDA ordinal.  stream: a cudaStream_t to run on (use cudaStreamLegacy / cudaStreamPerThread
 * to name a default stream), or NULL to let the context create its own non-blocking stream. */
int tob200_create(tob200_ctx **out, int device, void *stream);
int tob200_destroy(tob200_ctx *ctx);
int tob200_sync(tob200_ctx *ctx);
/* Message of the last failure on this context (ctx == NULL: last failure of tob200_create). */
const char *tob200_last_error(const tob200_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t tob200_launch_count(const tob200_ctx *ctx);
/* Device time in ms of the last compute entry point (CUDA events on the context's stream;
 * blocks until that work has finished). */
int tob200_last_elapsed_ms(tob200_ctx *ctx, float *ms);
/* exact != 0: tob200_lm_run_f32 and tob200_build_solve_f32 for 28 <= n <= 55 (and m >= 192, m n % 4 == 0, use_ldlt) run the warp-per-problem FFMA kernel, whose every sum is the CPU
 * restatement's canonical fma chain (bit-identical results), instead of the default tensor-core kernel (wtc.cuh:
 * H = J^T J off-diagonal from tcgen05.mma with an FP16 hi / lo split, held to the float tolerance of 1e-4).  The
 * environment variable TOB200_WPP_TC=0 selects the same at tob200_create time.
 * exact >= 2 additionally runs float 56 <= n <= 512 (tob200_lm_run_f32 / tob200_build_solve_f32: by default the
 * tcgen05 J^T J family, tolerance-held) on the general kernel family, whose results are bit-identical to the CPU
 * restatement - several times slower, for callers who want reproducibility down to the last bit (env TOB200_LG_EXACT=1). */
int tob200_set_exact(tob200_ctx *ctx, int exact);

void tob200_options_default(tob200_options *opt); /* == tinyopt::Options{} */

/* ---- device memory for callers without the CUDA toolkit (the C++ adaptor include/tinyopt_b200.hpp,
 * a cgo / JNI / ctypes binding): plain cudaMalloc / cudaFree / stream-ordered cudaMemcpyAsync on the
 * context's device and stream.  tob200_copy_to_host returns once the bytes are in host memory;
 * tob200_copy_to_device is asynchronous when `src` is pinned (keep `src` alive until tob200_sync). */
int tob200_device_alloc(tob200_ctx *ctx, size_t bytes, void **out);
int tob200_device_free(tob200_ctx *ctx, void *ptr);
int tob200_copy_to_device(tob200_ctx *ctx, void *dst_device, const void *src_host, size_t bytes);
int tob200_copy_to_host(tob200_ctx *ctx, void *dst_host, const void *src_device, size_t bytes);

/* Elements (not bytes) of a TILE32 buffer: ceil(B/32)*32*m*n.  r / y buffers: n = 1. */
int64_t tob200_tiled_elems(int64_t B, int m, int n);
/* Which kernel family serves (dtype, n): 1 thread-per-problem registers (n <= 12 float, n <= 8 double),
 * 2 warp-per-problem shared-memory tiles (n <= 55, float and double), 3 tensor-core J^T J + blocked
 * LDLT (56 <= n <= 512, float; n % 4 != 0 runs on a zero-padded copy), 4 the general family (double 56 <= n <= 2048,
 * float 513 <= n <= 2048, and `use_ldlt = 0` above n = 55: canonical operation order, bit-identical to the CPU
 * restatement, no tensor cores), 0 none (n > 2048: TOB200_ERR_UNSUPPORTED, never a fallback). */
int tob200_kernel_family(int dtype, int n);

/* ---- layout ----------------------------------------------------------------------------------- */
/* src[B][m][n] (problem-major) -> dst TILE32.  Use n = 1 for r / y. */
int tob200_retile_f32(tob200_ctx *ctx, const float *src, int64_t B, int m, int n, float *dst);
int tob200_retile_f64(tob200_ctx *ctx, const double *src, int64_t B, int m, int n, double *dst);

/* ---- a1+a3+a5+a6: one Build + Solve for a batch of materialised residual blocks ---------------
 * Replaces, per problem: `grad = J^T r; H = J^T J` (diff/optimize_autodiff.h:151-157, cost :164),
 * SolverLM::Build's damping `H_ii *= 1 + lambda` (solvers/lm.h:108-117) and SolverGN::Solve ->
 * SolveLDLT(H, -grad) (solvers/gn.h:150-156, math.h:232-240).
 *   lambda : [B] per-problem damping, or NULL for Gauss-Newton (no damping)
 *   dx     : [B][n]     cost : [B] (= |r|^2 accumulated in T, widened)     status : [B]
 *            status 0 ok, 1 LDLT rejected (info != Success or not positive; dx untouched)
 *   H_out  : optional [B][n][n] damped H_, full symmetric, row-major   g_out : optional [B][n] */
int tob200_build_solve_f32(tob200_ctx *ctx, const float *J, const float *r, int layout, int64_t B,
                           int m, int n, const float *lambda, float *dx, double *cost,
                           float *H_out, float *g_out, int32_t *status);
int tob200_build_solve_f64(tob200_ctx *ctx, const double *J, const double *r, int layout,
                           int64_t B, int m, int n, const double *lambda, double *dx, double *cost,
                           double *H_out, double *g_out, int32_t *status);

/* ---- a1 alone: the product site H = J^T diag(s^2) J on the tensor cores (tcgen05, 3xTF32) --------
 * Replaces `H = J.transpose() * J` (diff/optimize_autodiff.h:156, diff/num_diff.h:297) for
 * 1 <= n <= 512 (float; n % 4 != 0 through a zero-padded copy).  J: [B][m][n] problem-major; row_scale: [B][m] or NULL
 * (J_i = s_i a_i: the polynomial family's Jacobian from A);  H: [B][n][n] full symmetric. */
int tob200_jtj_f32(tob200_ctx *ctx, const float *J, const float *row_scale, int64_t B, int m, int n,
                   float *H);

/* ---- a6 alone: tinyopt::SolveLDLT(A, b) (math.h:232-240) == Eigen LDLT<_, Upper> + solve ----------
 * A: [B][n][n] row-major, only the upper triangle is read; b, x: [B][n]; status: [B], 0 ok,
 * 1 rejected (info() != Success or not positive; x untouched).  1 <= n <= 512 (float).  Blocked
 * left-looking factorisation with the oracle's operation order: bit-exact for the same A. */
int tob200_solve_ldlt_f32(tob200_ctx *ctx, const float *A, const float *b, int64_t B, int n, float *x,
                          int32_t *status);

/* ---- SURVEY.md §8(f) rank 2, the step after the path: tinyopt::InvCov(H) (math.h:44-57 DenseInvCov:
 * `H.selfadjointView<Upper>().ldlt()` then `chol.solve(Identity)`) and MaxStdDev (solvers/lm.h:176-187,
 * solvers/gn.h:177: sqrt of the largest coefficient of InvCov(H)) for a batch, through the same
 * diagonal-pivoted LDL^T as the solve of the path.
 *   H       : [B][n][n] row-major, only the upper triangle is read
 *   cov     : [B][n][n] out (untouched where status != 0), or NULL
 *   max_std : [B] out (0 where status != 0, as the reference returns), or NULL
 *   status  : [B]: 0 ok, 1 rejected (info() != Success or not positive: the reference's nullopt)
 * n <= 64 (float, double; bit-exact with the oracle) or 64 < n <= 512 (float). */
int tob200_inv_cov_f32(tob200_ctx *ctx, const float *H, int64_t B, int n, float *cov, float *max_std,
                       int32_t *status);
int tob200_inv_cov_f64(tob200_ctx *ctx, const double *H, int64_t B, int n, double *cov, double *max_std,
                       int32_t *status);

/* Device time (ms, CUDA events) the last large-n call (n > 55) spent in one kernel class:
 * phase 0 residual/gradient pass, 1 tensor-core J^T J, 2 factor + solve + LM step. */
int tob200_last_phase_ms(tob200_ctx *ctx, int phase, float *ms, int *launches);

/* ---- a7-a10: the whole LM loop, device resident, for the polynomial residual family ------------
 * r_i(x) = t_i + alpha t_i^3 - y_i, t = A x.  One call == one tinyopt::Optimize()
 * per problem (optimizers/optimizer.h:243-327 OptimizeAcc + :332-539 Step + solvers/lm.h), the
 * Jacobian (1 + 3 alpha t_i^2) a_i^T is formed on the fly so J never exists in HBM.
 *   A : TILE32 or PROBLEM_MAJOR [B][m][n]    y : same layout, [B][m]
 *   x : [B][n] in (x0) / out (solution)      results : [B] */
int tob200_lm_run_f32(tob200_ctx *ctx, const tob200_options *opt, const float *A, const float *y,
                      float alpha, int layout, int64_t B, int m, int n, float *x,
                      tob200_result *results);
int tob200_lm_run_f64(tob200_ctx *ctx, const tob200_options *opt, const double *A, const double *y,
                      double alpha, int layout, int64_t B, int m, int n, double *x,
                      tob200_result *results);
/* The same run that also returns Output::final_hessian (optimizer.h:313-316): the last H_ of every problem,
 * un-damped as SolverLM::Hessian() does (solvers/lm.h:157-171: diagonal / (1 + prev_lambda_)), widened to
 * double, [B][n][n] row-major, full symmetric.  Honoured only when options.save_last != 0 (options.h:66, the
 * reference's default); final_hessian may be NULL.  Every kernel family (n <= 2048, float and double). */
int tob200_lm_run_ex_f32(tob200_ctx *ctx, const tob200_options *opt, const float *A, const float *y,
                         float alpha, int layout, int64_t B, int m, int n, float *x,
                         tob200_result *results, double *final_hessian);
int tob200_lm_run_ex_f64(tob200_ctx *ctx, const tob200_options *opt, const double *A, const double *y,
                         double alpha, int layout, int64_t B, int m, int n, double *x,
                         tob200_result *results, double *final_hessian);
/* Same call with HOST buffers (pageable or pinned): H2D of A, y, x0, the run, D2H of x and
 * results, all inside; returns when the results are in host memory. */
int tob200_lm_run_host_f32(tob200_ctx *ctx, const tob200_options *opt, const float *A_host,
                           const float *y_host, float alpha, int layout, int64_t B, int m, int n,
                           float *x_host, tob200_result *results_host);
int tob200_lm_run_host_f64(tob200_ctx *ctx, const tob200_options *opt, const double *A_host,
                           const double *y_host, double alpha, int layout, int64_t B, int m, int n,
                           double *x_host, tob200_result *results_host);

/* ---- the SolverType seam for user-evaluated residuals (host-driven loop) -----------------------
 * A batched `Optimizer_<SolverLM>` whose per-problem state (x, lambda_, prev_lambda_, bad_factor_,
 * rebuild flag, H_, grad_, Output counters) lives on the device.  Each tob200_solver_step is one
 * Optimizer_::Step + the OptimizeAcc update (optimizer.h:266-309) for every still-running problem,
 * fed with the residual blocks the caller's lambda produced at the current x.
 * Every n <= 2048, float and double: n <= 55 on kernel families 1 and 2 (bit-identical to tob200_lm_run_* for
 * the same residual blocks), above on the general family 4 in both precisions (the reference's dynamic-size
 * solver has no size cap, math.h:232-240; bit-identical to the CPU oracle fed with the same J, r). */
int tob200_solver_create(tob200_ctx *ctx, int dtype, int64_t B, int n, const tob200_options *opt,
                         tob200_solver **out);
/* flags: TOB200_SOLVER_GENERAL runs the solver on the general family whatever n is (needed by
 * tob200_solver_step_cost_*). */
#define TOB200_SOLVER_GENERAL 1
int tob200_solver_create_ex(tob200_ctx *ctx, int dtype, int64_t B, int n, const tob200_options *opt, int flags,
                            tob200_solver **out);
int tob200_solver_destroy(tob200_solver *s);
/* Reset all problems (solvers/lm.h:46-52 reset()) and set x <- x0 ([B][n], device). */
int tob200_solver_reset(tob200_solver *s, const void *x0);
/* Device pointer to the current x [B][n] (dtype of the solver). */
void *tob200_solver_x(tob200_solver *s);
/* Device pointer to [B] int32: 1 = the next step must carry J for this problem (rebuild),
 * 0 = cost-only (r suffices), -1 = finished. */
const int32_t *tob200_solver_needs(tob200_solver *s);
int tob200_solver_step_f32(tob200_solver *s, const float *J, const float *r, int layout, int m);
int tob200_solver_step_f64(tob200_solver *s, const double *J, const double *r, int layout, int m);
/* tob200_solver_step with the pass's `Cost::cost` supplied by the caller ([B] doubles, device) instead of
 * being formed as r^T r: accumulation functors return what they like - diff/num_diff.h:300-305 returns the
 * NORM of the residuals, not its square.  grad_ = J^T r and H_ = J^T J as in tob200_solver_step; the residual
 * count is m.  General family only (tob200_solver_create_ex with TOB200_SOLVER_GENERAL), else
 * TOB200_ERR_UNSUPPORTED. */
int tob200_solver_step_cost_f32(tob200_solver *s, const float *J, const float *r, int layout, int m,
                                const double *cost);
int tob200_solver_step_cost_f64(tob200_solver *s, const double *J, const double *r, int layout, int m,
                                const double *cost);
/* The manual accumulation contract `acc(x, grad, H) -> Cost` (docs/API.md:37-57,137-170; examples
 * tests/optimize_easy.cpp:35-80 (Rosenbrock with its true Hessian), tests/types.cpp:97-108,
 * benchmarks/dense.cpp:57-66 ("Prior n": a diagonal H)): the caller's lambda has filled, for every problem,
 * what the reference hands it as the solver-owned grad_ and H_ (solvers/gn.h:109-113) and returned the cost.
 *   grad [B][n], H [B][n][n] row-major - only the UPPER triangle (row <= col) is read, as
 *   `H.selfadjointView<Upper>()` does (docs/API.md:170); cost [B] = Cost::cost (a double, cost.h:93),
 *   num_residuals [B] = Cost::num_resisuals (1 for a scalar cost, cost.h:22).
 * Problems whose `needs` entry is 0 (cost-only Step: the nullptr_t call of the reference) read only cost and
 * num_residuals; finished ones (-1) read nothing.  Same Step / OptimizeAcc bookkeeping as tob200_solver_step. */
int tob200_solver_step_hg_f32(tob200_solver *s, const float *grad, const float *H, const double *cost,
                              const int32_t *num_residuals);
int tob200_solver_step_hg_f64(tob200_solver *s, const double *grad, const double *H, const double *cost,
                              const int32_t *num_residuals);
/* The same contract for accumulation functors whose H is SPARSE (the reference selects its sparse solver from the
 * lambda's H type, optimize.h:27-33; tests/sparse.cpp:19-85): H arrives as triplets with ONE pattern for the whole batch -
 * rows / cols: [nnz] int32 HOST arrays; values: [B][nnz] device.  Duplicates are summed in triplet order (Eigen's
 * setFromTriplets), entries below the diagonal are ignored (`SimplicialLDLT<_, Upper>`, math.h:267-277).  The values are
 * scattered into a dense upper triangle and solved by the dense pivoted LDL^T of the path: the same solution as the sparse
 * factorisation's up to rounding for the positive definite systems LM produces; an indefinite H is rejected (the dense
 * reference semantics) where SimplicialLDLT would go on.  n <= 2048.  Synchronises the stream once (pattern upload). */
int tob200_solver_step_hg_sparse_f32(tob200_solver *s, const float *grad, const int32_t *rows, const int32_t *cols, int nnz,
                                     const float *values, const double *cost, const int32_t *num_residuals);
int tob200_solver_step_hg_sparse_f64(tob200_solver *s, const double *grad, const int32_t *rows, const int32_t *cols, int nnz,
                                     const double *values, const double *cost, const int32_t *num_residuals);
/* Number of problems still running (synchronises the stream). */
int tob200_solver_num_active(tob200_solver *s, int64_t *n_active);
/* Copy the per-problem results to `results` ([B], device). */
int tob200_solver_results(tob200_solver *s, tob200_result *results);
/* Final un-damped Hessian (solvers/lm.h:157-171 Hessian()) as doubles, [B][n][n] device. */
int tob200_solver_final_hessian(tob200_solver *s, double *H);
/* Output::Covariance() (output.h:81-103) / SolverLM::Covariance() (solvers/lm.h:173): InvCov of the
 * un-damped final Hessian, in double as the reference computes it.  cov: [B][n][n] device or NULL;
 * max_std: [B] or NULL; status: [B] (1 where the reference returns nullopt).  `rescaled` is a host-side
 * scalar (final_cost^2 / (num_residuals - n)): see the C++ / Python adaptors. */
int tob200_solver_covariance(tob200_solver *s, double *cov, double *max_std, int32_t *status);

/* ---- synthetic problem family, generated on the device (bit-identical to the CPU oracle's) ----
 * Any output may be NULL.  A, y in `layout`; xstar, x0: [B][n]. */
int tob200_synth_generate_f32(tob200_ctx *ctx, uint64_t seed, int64_t p0, int64_t B, int m, int n,
                              float alpha, float sigma, int layout, float *A, float *y,
                              float *xstar, float *x0);
int tob200_synth_generate_f64(tob200_ctx *ctx, uint64_t seed, int64_t p0, int64_t B, int m, int n,
                              double alpha, double sigma, int layout, double *A, double *y,
                              double *xstar, double *x0);
/* Residual blocks of the family at x: r and J in `layout` (what a user lambda + AD would hand to
 * tob200_build_solve / tob200_solver_step). */
int tob200_synth_eval_f32(tob200_ctx *ctx, const float *A, const float *y, float alpha, int layout,
                          int64_t B, int m, int n, const float *x, float *r, float *J);
int tob200_synth_eval_f64(tob200_ctx *ctx, const double *A, const double *y, double alpha,
                          int layout, int64_t B, int m, int n, const double *x, double *r,
                          double *J);

#ifdef __cplusplus
}
#endif
#endif /* TINYOPT_B200_H */
