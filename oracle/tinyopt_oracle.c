/*
 * tinyopt_oracle.c — CPU ORACLE (test infrastructure only; see tinyopt_oracle.h for the scope,
 * the reference citations and the parity status).  Build: `make -C oracle` (gcc, -ffp-contract=off).
 */
#include "tinyopt_oracle.h"

#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* optimizers/options.h defaults */
void too_options_default(too_options *o) {
  o->solver_type = 0;
  o->check_final_cost = 0;
  o->use_step_quality_approx = 0;
  o->grad_clipping = 0;
  o->use_ldlt = 1;
  o->H_is_full = 1;
  o->check_min_H_diag = 0;
  o->save_last = 1;
  o->use_squared_norm = 1;
  o->downscale_by_2 = 0;
  o->normalize = 0;
  o->max_iters = 50;
  o->min_error = 1e-12f;
  o->min_rerr_dec = 1e-10f;
  o->min_step_norm2 = 1e-14f;
  o->min_grad_norm2 = 1e-18f;
  o->max_total_failures = 0;
  o->max_consec_failures = 5;
  o->damping_init = 1e-4f;
  o->damping_min = 1e-9f;
  o->damping_max = 1e9f;
  o->good_factor = 1.0f / 3.0f;
  o->bad_factor = 2.0f;
}

int too_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* census only (tools/margin_census.py): 1 = accumulate the rows in reverse order */
static int too_census_row_order_ = 0;
void too_census_set_row_order(int rev) { too_census_row_order_ = rev; }

/* robust variant of the synthetic family: kind 0 = none, 1..7 = the M-estimators of
 * losses/robust_norms.h applied per residual inside the accumulation; th2 = squared threshold */
static int too_synth_robust_kind_ = 0;
static double too_synth_robust_th2_ = 1.0;
void too_synth_set_robust(int kind, double th2) { too_synth_robust_kind_ = kind; too_synth_robust_th2_ = th2; }
/* numeric-differentiation variant of the family (diff/num_diff.h:57-126, 284-309): method 0 = off (analytic J),
 * 1 = kForward, 2 = kCentral, 3 = kFastCentral; h <= 0 = FloatEpsilon<Scalar>() (math.h:297-301) */
static int too_synth_numdiff_method_ = 0;
static double too_synth_numdiff_h_ = 0.0;
void too_synth_set_numdiff(int method, double h) { too_synth_numdiff_method_ = method; too_synth_numdiff_h_ = h; }

/* stateless counter RNG of the synthetic family:
 * u(seed,p,k) = splitmix64-finaliser(seed ^ (p * golden + k)) */
static inline uint64_t too_hash64_(uint64_t seed, uint64_t p, uint64_t k) {
  uint64_t z = seed ^ (p * 0x9E3779B97F4A7C15ull + k);
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* math.h:298-301: FloatEpsilon<float>() = 1e-4f, FloatEpsilon<double>() = (double)1e-7f */
#define T double
#define SUF _f64
#define FMA fma
#define SQRT sqrt
#define FABS fabs
#define POW pow
#define T_MAX DBL_MAX
#define T_MIN DBL_MIN
#define FLOAT_EPS ((double)1e-7f)
#include "oracle_impl.inc"
#undef T
#undef SUF
#undef FMA
#undef SQRT
#undef FABS
#undef POW
#undef T_MAX
#undef T_MIN
#undef FLOAT_EPS

#define T float
#define T_IS_FLOAT 1
#define SUF _f32
#define FMA fmaf
#define SQRT sqrtf
#define FABS fabsf
#define POW powf
#define T_MAX FLT_MAX
#define T_MIN FLT_MIN
#define FLOAT_EPS (1e-4f)
#include "oracle_impl.inc"
