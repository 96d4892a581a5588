// tinyopt_b200_device.cuh — SURVEY.md §8(f) rank 1, the step BEFORE the path: the user's residual
// functor evaluated on the device and fused into the normal-equations accumulation, so that the
// Jacobian never exists in HBM.
//
// Header-only CUDA C++ (compile the translation unit that includes it with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --extended-lambda -fmad=false
//        -I<repo>/include -I<repo>/tinyopt_b200/csrc
// and link nothing else: the kernels below are instantiated in YOUR translation unit, like tinyopt's own
// templates).  What it mirrors in the reference:
//   * tinyopt::Optimize(x, residuals, options) with automatic differentiation
//     (optimize.h:17-77 -> diff/optimize_autodiff.h:33-166): the residual functor is templated on
//     its scalar and is called with Jet<T, N> (3rdparty/ceres/jet.h:217) when the solver rebuilds the
//     system and with plain T for cost-only iterations (solvers/gn.h:98-105);
//   * the accumulation contract `acc(x, grad, H)` (docs/API.md:37-57) for functors that provide their
//     own Jacobian rows.
// Differences forced by the device: residuals are EMITTED one at a time (`emit(r)`) instead of being
// returned as a vector — only one Jet is live per thread — and the functor receives the problem index.
// Everything after the accumulation is the library's own device code (lm_state.cuh, ldlt_reg.cuh): same
// LM semantics, same stop tests, same tob200_result as tob200_lm_run_*.
//
// Kernel family: thread per problem (n <= 12 float, n <= 8 double — the register-resident LDLT of
// ldlt_reg.cuh).  Each accumulator receives its terms in emission order as a chain of IEEE fmas
// (DESIGN.md §4), so a functor that emits the canonical op sequence is bit-identical to the CPU oracle.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "tinyopt_b200.h"
#include "lm_state.cuh"  // -I<repo>/tinyopt_b200/csrc
#include "wpp.cuh"       // the warp-per-problem building blocks (n <= 55)

namespace tinyopt {
namespace b200 {
namespace device {

using tob200::Ops;
using tob200::tri_count;
using tob200::tri_index;

// ---- forward-mode dual number: value + N partials (3rdparty/ceres/jet.h:217) ---------------------
template <typename T, int N>
struct Jet {
  T a;
  T v[N];
  __host__ __device__ Jet() : a((T)0) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = (T)0;
  }
  __host__ __device__ Jet(T value) : a(value) {  // NOLINT: implicit, as ceres::Jet
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = (T)0;
  }
  __host__ __device__ Jet(T value, int k) : a(value) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = (j == k) ? (T)1 : (T)0;
  }
};

#define TOB200_JET_FN template <typename T, int N> __host__ __device__ inline
TOB200_JET_FN Jet<T, N> operator+(const Jet<T, N> &f, const Jet<T, N> &g) {
  Jet<T, N> h;
  h.a = f.a + g.a;
#pragma unroll
  for (int j = 0; j < N; ++j) h.v[j] = f.v[j] + g.v[j];
  return h;
}
TOB200_JET_FN Jet<T, N> operator-(const Jet<T, N> &f, const Jet<T, N> &g) {
  Jet<T, N> h;
  h.a = f.a - g.a;
#pragma unroll
  for (int j = 0; j < N; ++j) h.v[j] = f.v[j] - g.v[j];
  return h;
}
TOB200_JET_FN Jet<T, N> operator-(const Jet<T, N> &f) {
  Jet<T, N> h;
  h.a = -f.a;
#pragma unroll
  for (int j = 0; j < N; ++j) h.v[j] = -f.v[j];
  return h;
}
TOB200_JET_FN Jet<T, N> operator*(const Jet<T, N> &f, const Jet<T, N> &g) {
  Jet<T, N> h;
  h.a = f.a * g.a;
#pragma unroll
  for (int j = 0; j < N; ++j) h.v[j] = f.a * g.v[j] + f.v[j] * g.a;
  return h;
}
TOB200_JET_FN Jet<T, N> operator/(const Jet<T, N> &f, const Jet<T, N> &g) {
  Jet<T, N> h;
  const T gi = (T)1 / g.a, q = f.a * gi;
  h.a = q;
#pragma unroll
  for (int j = 0; j < N; ++j) h.v[j] = (f.v[j] - q * g.v[j]) * gi;
  return h;
}
TOB200_JET_FN Jet<T, N> operator+(const Jet<T, N> &f, T s) { Jet<T, N> h = f; h.a = f.a + s; return h; }
TOB200_JET_FN Jet<T, N> operator+(T s, const Jet<T, N> &f) { return f + s; }
TOB200_JET_FN Jet<T, N> operator-(const Jet<T, N> &f, T s) { Jet<T, N> h = f; h.a = f.a - s; return h; }
TOB200_JET_FN Jet<T, N> operator-(T s, const Jet<T, N> &f) { return (-f) + s; }
TOB200_JET_FN Jet<T, N> operator*(const Jet<T, N> &f, T s) {
  Jet<T, N> h;
  h.a = f.a * s;
#pragma unroll
  for (int j = 0; j < N; ++j) h.v[j] = f.v[j] * s;
  return h;
}
TOB200_JET_FN Jet<T, N> operator*(T s, const Jet<T, N> &f) { return f * s; }
TOB200_JET_FN Jet<T, N> operator/(const Jet<T, N> &f, T s) { return f * ((T)1 / s); }
TOB200_JET_FN Jet<T, N> operator/(T s, const Jet<T, N> &g) { return Jet<T, N>(s) / g; }
TOB200_JET_FN Jet<T, N> &operator+=(Jet<T, N> &f, const Jet<T, N> &g) { f = f + g; return f; }
TOB200_JET_FN Jet<T, N> &operator-=(Jet<T, N> &f, const Jet<T, N> &g) { f = f - g; return f; }
TOB200_JET_FN Jet<T, N> &operator*=(Jet<T, N> &f, const Jet<T, N> &g) { f = f * g; return f; }
TOB200_JET_FN Jet<T, N> &operator+=(Jet<T, N> &f, T s) { f.a += s; return f; }
TOB200_JET_FN Jet<T, N> &operator*=(Jet<T, N> &f, T s) { f = f * s; return f; }
TOB200_JET_FN bool operator<(const Jet<T, N> &f, const Jet<T, N> &g) { return f.a < g.a; }
TOB200_JET_FN bool operator>(const Jet<T, N> &f, const Jet<T, N> &g) { return f.a > g.a; }
TOB200_JET_FN bool operator<(const Jet<T, N> &f, T s) { return f.a < s; }
TOB200_JET_FN bool operator>(const Jet<T, N> &f, T s) { return f.a > s; }

// chain rule helper: h = (fa, dfa * f.v)
TOB200_JET_FN Jet<T, N> jet_chain(const Jet<T, N> &f, T fa, T dfa) {
  Jet<T, N> h;
  h.a = fa;
#pragma unroll
  for (int j = 0; j < N; ++j) h.v[j] = dfa * f.v[j];
  return h;
}
TOB200_JET_FN Jet<T, N> sqrt(const Jet<T, N> &f) { const T s = ::sqrt(f.a); return jet_chain(f, s, (T)1 / ((T)2 * s)); }
TOB200_JET_FN Jet<T, N> exp(const Jet<T, N> &f) { const T e = ::exp(f.a); return jet_chain(f, e, e); }
TOB200_JET_FN Jet<T, N> log(const Jet<T, N> &f) { return jet_chain(f, ::log(f.a), (T)1 / f.a); }
TOB200_JET_FN Jet<T, N> sin(const Jet<T, N> &f) { return jet_chain(f, ::sin(f.a), ::cos(f.a)); }
TOB200_JET_FN Jet<T, N> cos(const Jet<T, N> &f) { return jet_chain(f, ::cos(f.a), -::sin(f.a)); }
TOB200_JET_FN Jet<T, N> atan(const Jet<T, N> &f) { return jet_chain(f, ::atan(f.a), (T)1 / ((T)1 + f.a * f.a)); }
TOB200_JET_FN Jet<T, N> tanh(const Jet<T, N> &f) { const T t = ::tanh(f.a); return jet_chain(f, t, (T)1 - t * t); }
TOB200_JET_FN Jet<T, N> abs(const Jet<T, N> &f) { return f.a < (T)0 ? -f : f; }
#undef TOB200_JET_FN
// the same spellings for plain scalars, so a functor templated on its scalar type compiles for both
__host__ __device__ inline float sqrt(float x) { return ::sqrtf(x); }
__host__ __device__ inline double sqrt(double x) { return ::sqrt(x); }
__host__ __device__ inline float exp(float x) { return ::expf(x); }
__host__ __device__ inline double exp(double x) { return ::exp(x); }
__host__ __device__ inline float log(float x) { return ::logf(x); }
__host__ __device__ inline double log(double x) { return ::log(x); }
__host__ __device__ inline float sin(float x) { return ::sinf(x); }
__host__ __device__ inline double sin(double x) { return ::sin(x); }
__host__ __device__ inline float cos(float x) { return ::cosf(x); }
__host__ __device__ inline double cos(double x) { return ::cos(x); }
__host__ __device__ inline float atan(float x) { return ::atanf(x); }
__host__ __device__ inline double atan(double x) { return ::atan(x); }

// ---- M-estimators: losses/robust_norms.h restated for device functors ----
// Each takes the squared norm n2 of a residual block and the squared threshold th2 and returns the
// robust loss together with the scale s = d loss / d n2 the reference returns as `J_scale`
// (robust_norms.h: "the scale can then be used to solve JtJ * dx = Jt * res * s").  Templated on the
// scalar, so they also differentiate through Jets.  Pinned by tests/robust_norms.cpp's closed forms and
// its scale == autodiff-derivative check (tests/cuda/test_device_functor.cu, host side).
namespace losses {
template <typename S>
struct Robust {
  S loss, scale;
};
template <typename T, int N> __host__ __device__ inline T value_of(const Jet<T, N> &s) { return s.a; }
__host__ __device__ inline float value_of(float s) { return s; }
__host__ __device__ inline double value_of(double s) { return s; }

/// std::numeric_limits<T>::min(): the reference clamps the outlier scales of Huber / Arctan / Cauchy with
/// max<T>(numeric_limits<T>::min(), scale) so that an extreme outlier never zeroes a Jacobian row
/// (robust_norms.h:95,184,221; for Jets the comparison is on the value part, as ceres::Jet's operator<)
__host__ __device__ inline float min_normal_of(float) { return 1.175494351e-38f; }
__host__ __device__ inline double min_normal_of(double) { return 2.2250738585072014e-308; }
template <typename S>
__host__ __device__ inline S clamp_scale(const S &scale) {
  const auto lo = min_normal_of(value_of(scale));
  return value_of(scale) < lo ? S(lo) : scale;  // max(lo, scale): lo wins only if strictly larger
}

/// robust_norms.h:32-55: loss = min(n2, th2), scale in {1, 0}
template <typename S, typename T>
__host__ __device__ inline Robust<S> Truncated(const S &n2, T th2) {
  if (value_of(n2) <= th2) return {n2, S((T)1)};
  return {S(th2), S((T)0)};
}
/// robust_norms.h:67-103: loss = n2 (inlier) or 2 th n - th2, scale = th / n
template <typename S, typename T>
__host__ __device__ inline Robust<S> Huber(const S &n2, T th2) {
  if (value_of(n2) <= th2) return {n2, S((T)1)};
  const T th = sqrt(th2);
  const S n = sqrt(n2);
  return {(T)2 * th * n - th2, clamp_scale(S(th / n))};
}
/// robust_norms.h:118-152: loss = th2 (1 - (1 - n2 / th2)^3) (inlier) or th2, scale = 3 (th2 - n2)^2 / th2^2 or 0
template <typename S, typename T>
__host__ __device__ inline Robust<S> Tukey(const S &n2, T th2) {
  if (value_of(n2) <= th2) {
    const S s = (T)1 - n2 / th2;
    const S d = th2 - n2;
    return {th2 * ((T)1 - s * s * s), (T)3 * d * d / (th2 * th2)};
  }
  return {S(th2), S((T)0)};
}
/// robust_norms.h:165-191: loss = th atan2(n2, th), scale = 1 / (n2^2 / th2 + 1)
template <typename S, typename T>
__host__ __device__ inline Robust<S> Arctan(const S &n2, T th2) {
  const T th = sqrt(th2);
  return {th * atan(n2 / th), clamp_scale(S((T)1 / (n2 * n2 / th2 + (T)1)))};  // th > 0: atan2(n2, th) == atan(n2 / th)
}
/// robust_norms.h:204-228: loss = th2 log(1 + n2 / th2), scale = 1 / (1 + n2 / th2)
template <typename S, typename T>
__host__ __device__ inline Robust<S> Cauchy(const S &n2, T th2) {
  const S s = (T)1 + n2 / th2;
  return {th2 * log(s), clamp_scale(S((T)1 / s))};
}
/// robust_norms.h:241-265: loss = n2 / (n2 + th2), scale = th2 / (n2 + th2)^2
template <typename S, typename T>
__host__ __device__ inline Robust<S> GemanMcClure(const S &n2, T th2) {
  const S e = n2 + th2;
  return {n2 / e, th2 / (e * e)};
}
/// robust_norms.h:278-303: loss = -log(exp(-n2) + exp(-th2)), scale = 1 / (exp(-th2) exp(n2) + 1)
template <typename S, typename T>
__host__ __device__ inline Robust<S> BlakeZisserman(const S &n2, T th2) {
  const T eps = exp(-th2);
  return {-log(exp(-n2) + eps), (T)1 / (eps * exp(n2) + (T)1)};
}
}  // namespace losses

// ---- the accumulation site (a1: grad = J^T r, H = J^T J, cost = |r|^2; diff/optimize_autodiff.h:151-164)
// One Emit lives in the registers of one thread for one pass.  Every accumulator receives its terms
// in emission order: cost = fma(r, r, cost); g_j = fma(J_j, r, g_j); H_jk = fma(J_j, J_k, H_jk), j <= k.
template <typename T, int N>
struct Emit {
  static constexpr int NT = tri_count(N);
  T hu[NT], g[N], cost;
  int nres;
  bool want_j;
  __device__ explicit Emit(bool rebuild) : cost((T)0), nres(0), want_j(rebuild) {
#pragma unroll
    for (int i = 0; i < NT; ++i) hu[i] = (T)0;
#pragma unroll
    for (int j = 0; j < N; ++j) g[j] = (T)0;
  }
  /// residual with its Jacobian row d r / d x (manual derivatives, docs/API.md:37-57)
  __device__ __forceinline__ void operator()(T r, const T (&J)[N]) {
    using O = Ops<T>;
    cost = O::fma(r, r, cost);
    ++nres;
    if (!want_j) return;
#pragma unroll
    for (int j = 0; j < N; ++j) g[j] = O::fma(J[j], r, g[j]);
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
      for (int k = j; k < N; ++k) hu[tri_index(N, j, k)] = O::fma(J[j], J[k], hu[tri_index(N, j, k)]);
    }
  }
  /// residual as a Jet (automatic differentiation): J.row(i) = res[i].v (diff/optimize_autodiff.h:138-141)
  __device__ __forceinline__ void operator()(const Jet<T, N> &r) { (*this)(r.a, r.v); }
  /// cost-only residual (passes that do not rebuild: solvers/gn.h:98-105)
  __device__ __forceinline__ void operator()(T r) {
    cost = Ops<T>::fma(r, r, cost);
    ++nres;
  }
  /// ROBUST residual: the M-estimator re-weights the accumulation inside the same pass —
  /// `rb = losses::Huber(r * r, th2)` etc. gives the loss and the scale of losses/robust_norms.h, and, as its
  /// header says (robust_norms.h:16-29: "the scale can then be used to solve JtJ * dx = Jt * res * s"):
  ///   cost += loss ;  g_j = fma(J_j, r * scale, g_j) ;  H_jk = fma(J_j, J_k, H_jk)
  __device__ __forceinline__ void robust(T r, const T (&J)[N], T loss, T scale) {
    using O = Ops<T>;
    cost = O::add(cost, loss);
    ++nres;
    if (!want_j) return;
    const T rs = O::mul(r, scale);
#pragma unroll
    for (int j = 0; j < N; ++j) g[j] = O::fma(J[j], rs, g[j]);
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
      for (int k = j; k < N; ++k) hu[tri_index(N, j, k)] = O::fma(J[j], J[k], hu[tri_index(N, j, k)]);
    }
  }
  /// robust residual of a cost-only pass
  __device__ __forceinline__ void robust(T loss) {
    cost = Ops<T>::add(cost, loss);
    ++nres;
  }
};

namespace detail {

template <typename T, int N>
struct ThreadHG {  // persistent damped H_ / grad_ of one problem (solvers/gn.h:200-201), tile-interleaved in HBM
  T *h, *g;
  __device__ __forceinline__ T ld_h(int i) const { return h[i * 32]; }
  __device__ __forceinline__ void st_h(int i, T v) { h[i * 32] = v; }
  __device__ __forceinline__ T ld_g(int j) const { return g[j * 32]; }
  __device__ __forceinline__ void st_g(int j, T v) { g[j * 32] = v; }
};

// kAutoDiff: f(p, x, emit) is templated on the scalar of x (Jet<T, N> on rebuild passes, T otherwise);
// else:      f(p, x, emit, want_jacobian) with x plain T and emit(r, Jrow) / emit(r).
template <typename T, int N, bool kAutoDiff, typename F>
__global__ void __launch_bounds__(128) functor_lm_run_kernel(F f, tob200::DevOptions<T> opt, T *x,
                                                             tob200_result *results, int64_t B, T *hg_store) {
  constexpr int NT = tri_count(N);
  const int lane = threadIdx.x & 31;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t Bpad = (B + 31) / 32 * 32;
  const bool is_lm = opt.solver_type == 0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < Bpad; p += nthreads) {
    if (p >= B) continue;
    tob200::LmState<T, N> s;
    s.reset(opt);
#pragma unroll
    for (int j = 0; j < N; ++j) s.x[j] = x[p * N + j];
    T *base = hg_store + (size_t)(p / 32) * (NT + N) * 32 + lane;
    ThreadHG<T, N> hg{base, base + (size_t)NT * 32};
    while (!s.done()) {
      const bool do_rebuild = !is_lm || s.rebuild();  // GN's Build always re-accumulates (gn.h:118-131)
      Emit<T, N> emit(do_rebuild);
      if constexpr (kAutoDiff) {
        if (do_rebuild) {
          Jet<T, N> xj[N];
#pragma unroll
          for (int j = 0; j < N; ++j) xj[j] = Jet<T, N>(s.x[j], j);
          f(p, xj, emit);
        } else {
          f(p, s.x, emit);
        }
      } else {
        f(p, s.x, emit, do_rebuild);
      }
      tob200::lm_after_pass<T, N>(s, opt, do_rebuild, emit.hu, emit.g, emit.cost, emit.nres, hg);
    }
#pragma unroll
    for (int j = 0; j < N; ++j) x[p * N + j] = s.x[j];
    tob200::lm_write_result(s, &results[p]);
  }
}

template <typename T, int N, bool kAutoDiff, typename F>
cudaError_t launch(const F &f, T *x, int64_t B, const tob200_options &options, tob200_result *results,
                   cudaStream_t stream) {
  static_assert(N >= 1 && N <= (sizeof(T) == 8 ? 8 : 12),
                "thread-per-problem family: n <= 12 (float) / n <= 8 (double)");
  if (B <= 0) return cudaSuccess;
  constexpr int NT = tri_count(N);
  const int64_t Bpad = (B + 31) / 32 * 32;
  T *hg = nullptr;
  cudaError_t e = cudaMallocAsync(&hg, (size_t)Bpad * (NT + N) * sizeof(T), stream);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((Bpad + 127) / 128);
  functor_lm_run_kernel<T, N, kAutoDiff, F>
      <<<grid, 128, 0, stream>>>(f, tob200::make_dev_options<T>(options), x, results, B, hg);
  e = cudaGetLastError();
  cudaFreeAsync(hg, stream);
  return e;
}

}  // namespace detail

/// tinyopt::Optimize(x, residuals, options) with automatic differentiation, one problem per thread.
///   f : `template <typename S> __device__ void operator()(int64_t p, const S (&x)[N], Emit<T, N> &emit)`
///       — evaluate the residuals of problem p at x and call emit(r) for each (S is Jet<T, N> or T)
///   x : [B][N] device, in/out          results : [B] device
/// Asynchronous on `stream`; returns the launch status.
template <int N, typename T, typename F>
cudaError_t OptimizeBatchAutoDiff(const F &f, T *x, int64_t B, const tob200_options &options,
                                  tob200_result *results, cudaStream_t stream = nullptr) {
  return detail::launch<T, N, true, F>(f, x, B, options, results, stream);
}

/// The accumulation contract with user-provided derivatives:
///   f : `__device__ void operator()(int64_t p, const T (&x)[N], Emit<T, N> &emit, bool want_jacobian)`
///       — call emit(r, Jrow) per residual (or emit(r) when !want_jacobian)
template <int N, typename T, typename F>
cudaError_t OptimizeBatchManual(const F &f, T *x, int64_t B, const tob200_options &options,
                                tob200_result *results, cudaStream_t stream = nullptr) {
  return detail::launch<T, N, false, F>(f, x, B, options, results, stream);
}

// ================================================================================================
// Warp per problem (n <= 55, float and double): the same functor source, run in LOCKSTEP by the 32
// lanes of a warp for one problem.  Derivatives are warp-distributed: every lane carries the value and
// the two partials d/dx_lane, d/dx_{lane+32} (a Jet<T, 2> seeded per lane), so a Jet operation costs three
// scalar operations per lane whatever n is.  emit(r) writes the augmented row [J_i | r_i] into the packed
// row buffer of wpp.cuh (lane l stores its two entries); every 32 rows the warp folds the buffer into
// its register blocks of [J|r]^T [J|r] exactly as the library's own warp-per-problem kernel does, and the
// pass ends in the same wpp_after_pass (pivoted LDL^T, LM state machine).  Accumulation order = emission
// order, so the result is bit-identical to tob200_lm_run_* for a functor that emits the same values.
// The functor must be templated on the type of x and of emit:
//   template <typename X, typename E> __device__ void operator()(int64_t p, const X &x, E &emit) const
// (x[j] yields Jet<T, 2> on rebuild passes and T on cost-only passes) and must be free of lane-dependent
// control flow.
// ================================================================================================
template <typename T>
struct WarpXJet {
  const T *xs;
  int lane;
  __device__ __forceinline__ Jet<T, 2> operator[](int j) const {
    Jet<T, 2> r;
    r.a = xs[j];
    r.v[0] = (lane == j) ? (T)1 : (T)0;
    r.v[1] = (lane + 32 == j) ? (T)1 : (T)0;
    return r;
  }
};
template <typename T>
struct WarpXScalar {
  const T *xs;
  __device__ __forceinline__ T operator[](int j) const { return xs[j]; }
};

template <typename T, int N>
struct WarpEmit {
  static constexpr int BLK = tob200::wpp_blk_for(N), NB = tob200::wpp_nb_for(N), NP = NB * BLK,
                       NPS = tob200::wpp_nps(NP);
  static constexpr bool kF32 = sizeof(T) == 4;
  T *jbuf;
  int lane, bi, bj;
  bool has_block, want_j;
  unsigned long long acc2[BLK][BLK / 2];  // float: packed pairs for FFMA2
  T acc[BLK][BLK];
  T cost;
  int nres, fill;
  bool robust_cost = false;  // the pass emitted robust residuals: `cost` (sum of losses) is the pass's cost

  __device__ WarpEmit(T *jbuf_, int lane_, int bi_, int bj_, bool has_block_, bool rebuild)
      : jbuf(jbuf_), lane(lane_), bi(bi_), bj(bj_), has_block(has_block_), want_j(rebuild), cost((T)0), nres(0), fill(0) {
#pragma unroll
    for (int u = 0; u < BLK; ++u) {
#pragma unroll
      for (int h = 0; h < BLK / 2; ++h) acc2[u][h] = 0ull;
#pragma unroll
      for (int v = 0; v < BLK; ++v) acc[u][v] = (T)0;
    }
    // the factor matrix aliases the row buffer: clear the pad columns of all 32 rows before the pass
    constexpr int w = NP - (N + 1);
    if (w > 0)
      for (int e = lane; e < 32 * w; e += 32) jbuf[(e / w) * NPS + tob200::wpp_col<BLK>(N + 1 + (e % w))] = (T)0;
    __syncwarp();
  }
  __device__ __forceinline__ void row_done() {
    ++nres;
    if (++fill == 32) flush();
  }
  /// residual as a warp-distributed Jet (rebuild passes of OptimizeBatchAutoDiffWarp)
  __device__ __forceinline__ void operator()(const Jet<T, 2> &r) {
    T *row = jbuf + fill * NPS;
    if (lane < N) row[tob200::wpp_col<BLK>(lane)] = r.v[0];
    if (lane + 32 < N) row[tob200::wpp_col<BLK>(lane + 32)] = r.v[1];
    if (lane == 0) row[tob200::wpp_col<BLK>(N)] = r.a;
    row_done();
  }
  /// residual with its full Jacobian row (every lane holds the same row; rebuild passes of ...ManualWarp)
  __device__ __forceinline__ void operator()(T r, const T (&J)[N]) {
    T *row = jbuf + fill * NPS;
#pragma unroll
    for (int j = 0; j < N; ++j)
      if ((j & 31) == lane) row[tob200::wpp_col<BLK>(j)] = J[j];
    if (lane == 0) row[tob200::wpp_col<BLK>(N)] = r;
    row_done();
  }
  /// cost-only residual
  __device__ __forceinline__ void operator()(T r) {
    cost = Ops<T>::fma(r, r, cost);
    ++nres;
  }
  /// ROBUST residual (see Emit::robust): the row goes into the block accumulation as [J | r * scale], so that
  /// its augmented column gives g = sum J^T (r * scale); the cost is the sum of the losses, kept by every lane
  /// (the corner of the augmented matrix would hold sum (r * scale)^2 instead: `robust_cost` tells the kernel)
  __device__ __forceinline__ void robust(T r, const T (&J)[N], T loss, T scale) {
    cost = Ops<T>::add(cost, loss);
    robust_cost = true;
    if (!want_j) { ++nres; return; }
    T *row = jbuf + fill * NPS;
    const T rs = Ops<T>::mul(r, scale);
#pragma unroll
    for (int j = 0; j < N; ++j)
      if ((j & 31) == lane) row[tob200::wpp_col<BLK>(j)] = J[j];
    if (lane == 0) row[tob200::wpp_col<BLK>(N)] = rs;
    row_done();
  }
  __device__ __forceinline__ void robust(T loss) {
    cost = Ops<T>::add(cost, loss);
    robust_cost = true;
    ++nres;
  }
  // fold the buffered rows into the register blocks (wpp.cuh phase 2, same operation order)
  __device__ __forceinline__ void flush() {
    __syncwarp();
    if (has_block) {
      const T *pa = jbuf + tob200::wpp_col<BLK>(bi * BLK), *pb = jbuf + tob200::wpp_col<BLK>(bj * BLK);
      if constexpr (kF32) {
#pragma unroll 2
        for (int i = 0; i < fill; ++i) {
          T a[BLK], b[BLK];
#pragma unroll
          for (int q = 0; q < BLK / 4; ++q) {
            const float4 av = *reinterpret_cast<const float4 *>(pa + i * NPS + 4 * q);
            const float4 bv = *reinterpret_cast<const float4 *>(pb + i * NPS + 4 * q);
            a[4 * q] = av.x; a[4 * q + 1] = av.y; a[4 * q + 2] = av.z; a[4 * q + 3] = av.w;
            b[4 * q] = bv.x; b[4 * q + 1] = bv.y; b[4 * q + 2] = bv.z; b[4 * q + 3] = bv.w;
          }
#pragma unroll
          for (int u = 0; u < BLK; ++u)
#pragma unroll
            for (int h = 0; h < BLK / 2; ++h) tob200::ffma2_bcast(acc2[u][h], a[u], b[2 * h], b[2 * h + 1]);
        }
      } else {
        for (int i = 0; i < fill; ++i) {
          T a[BLK], b[BLK];
#pragma unroll
          for (int u = 0; u < BLK; ++u) { a[u] = pa[i * NPS + u]; b[u] = pb[i * NPS + u]; }
#pragma unroll
          for (int u = 0; u < BLK; ++u)
#pragma unroll
            for (int v = 0; v < BLK; ++v) acc[u][v] = Ops<T>::fma(a[u], b[v], acc[u][v]);
        }
      }
    }
    __syncwarp();
    fill = 0;
  }
  __device__ __forceinline__ void finish() {
    if (fill) flush();
    if constexpr (kF32) {
#pragma unroll
      for (int u = 0; u < BLK; ++u)
#pragma unroll
        for (int h = 0; h < BLK / 2; ++h) {
          acc[u][2 * h] = __uint_as_float((uint32_t)acc2[u][h]);
          acc[u][2 * h + 1] = __uint_as_float((uint32_t)(acc2[u][h] >> 32));
        }
    }
  }
};

namespace detail {

template <typename T, int N, bool kAutoDiff, typename F>
__global__ void __launch_bounds__(tob200::kWppThreads, sizeof(T) == 4 ? 2 : 1)
    functor_warp_lm_run_kernel(F f, tob200::DevOptions<T> opt, tob200::WppSmem L, T *x, tob200_result *results, int64_t B,
                               unsigned long long *counter, T *hpersist) {
  using namespace tob200;
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int BLK = wpp_blk_for(N), NB = wpp_nb_for(N), NP = NB * BLK, LDW = wpp_ldw(NP);
  const int lane = threadIdx.x & 31, wid = threadIdx.x / 32;
  unsigned char *ws = smem + (size_t)wid * L.total;
  T *xs = reinterpret_cast<T *>(ws + L.xs);
  T *last_dx = reinterpret_cast<T *>(ws + L.last_dx);
  T *jbuf = reinterpret_cast<T *>(ws + L.jbuf);
  T *hp = hpersist + ((size_t)blockIdx.x * (kWppThreads / 32) + wid) * (NP * LDW);
  int bi, bj;
  bool has_block;
  wpp_block_of_lane<NB>(lane, bi, bj, has_block);
  const bool is_lm = opt.solver_type == 0;
  WppData<T> d;  // what wpp_after_pass reads: n, the residual count of the pass, the layout
  d.A = nullptr; d.y = nullptr; d.B = B; d.m = 0; d.n = N; d.stages = 0; d.use_tma = 0; d.L = L; d.counter = counter;
  d.hpersist = hpersist;

  for (int64_t pr = wpp_next(counter, lane); pr < B; pr = wpp_next(counter, lane)) {
    for (int j = lane; j < NP; j += 32) {
      xs[j] = j < N ? x[pr * N + j] : (T)0;
      last_dx[j] = (T)0;
    }
    LmScalars<T> s;
    s.reset_scalars(opt);
    __syncwarp();
    while (!s.done()) {
      const bool do_rebuild = !is_lm || s.rebuild();
      WarpEmit<T, N> emit(jbuf, lane, bi, bj, has_block, do_rebuild);
      if constexpr (kAutoDiff) {
        if (do_rebuild) f(pr, WarpXJet<T>{xs, lane}, emit);
        else f(pr, WarpXScalar<T>{xs}, emit);
      } else {
        f(pr, WarpXScalar<T>{xs}, emit, do_rebuild);
      }
      emit.finish();
      d.m = emit.nres;
      const double cost_d = (double)emit.cost;  // robust passes: the sum of the losses replaces the corner r^T r
      const double *cd = emit.robust_cost ? &cost_d : nullptr;
      if (opt.use_ldlt) wpp_after_pass<T, NB, BLK, false>(s, opt, d, ws, hp, do_rebuild, bi, bj, has_block, emit.acc, emit.cost, lane, false, cd, emit.nres);
      else wpp_after_pass<T, NB, BLK, true>(s, opt, d, ws, hp, do_rebuild, bi, bj, has_block, emit.acc, emit.cost, lane, false, cd, emit.nres);
    }
    for (int j = lane; j < N; j += 32) x[pr * N + j] = xs[j];
    if (lane == 0) lm_write_result(s, &results[pr]);
    __syncwarp();
  }
}

template <typename T, int N, bool kAutoDiff, typename F>
cudaError_t launch_warp(const F &f, T *x, int64_t B, const tob200_options &options, tob200_result *results,
                        cudaStream_t stream) {
  using namespace tob200;
  static_assert(N >= 1 && N <= 55, "warp-per-problem family: n <= 55");
  if (B <= 0) return cudaSuccess;
  constexpr int BLK = wpp_blk_for(N), NB = wpp_nb_for(N), NP = NB * BLK, LDW = wpp_ldw(NP);
  static_assert(NB <= 7, "n + 1 columns must fit 28 register blocks");
  const WppSmem L = wpp_smem_layout(N, NP, 0, (uint32_t)sizeof(T));
  const int warps = kWppThreads / 32;
  const size_t smem = (size_t)L.total * warps;
  auto kern = functor_warp_lm_run_kernel<T, N, kAutoDiff, F>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0, dev = 0, sms = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWppThreads, smem)) != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t grid = (int64_t)per_sm * sms;
  const int64_t need = (B + warps - 1) / warps;
  if (grid > need) grid = need;
  unsigned long long *counter = nullptr;
  T *hp = nullptr;
  if ((e = cudaMallocAsync(&counter, sizeof(unsigned long long), stream)) != cudaSuccess) return e;
  if ((e = cudaMallocAsync(&hp, (size_t)grid * warps * NP * LDW * sizeof(T), stream)) != cudaSuccess) return e;
  cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
  kern<<<(unsigned)grid, kWppThreads, smem, stream>>>(f, make_dev_options<T>(options), L, x, results, B, counter, hp);
  e = cudaGetLastError();
  cudaFreeAsync(hp, stream);
  cudaFreeAsync(counter, stream);
  return e;
}

}  // namespace detail

/// tinyopt::Optimize(x, residuals, options) with automatic differentiation for n <= 55, one problem per warp.
/// f : `template <typename X, typename E> __device__ void operator()(int64_t p, const X &x, E &emit) const`
template <int N, typename T, typename F>
cudaError_t OptimizeBatchAutoDiffWarp(const F &f, T *x, int64_t B, const tob200_options &options,
                                      tob200_result *results, cudaStream_t stream = nullptr) {
  return detail::launch_warp<T, N, true, F>(f, x, B, options, results, stream);
}
/// The accumulation contract for n <= 55, one problem per warp.
/// f : `template <typename X, typename E> __device__ void operator()(int64_t p, const X &x, E &emit, bool want_jacobian) const`
///     — emit(r, Jrow) per residual (every lane passes the same full row), or emit(r) when !want_jacobian
template <int N, typename T, typename F>
cudaError_t OptimizeBatchManualWarp(const F &f, T *x, int64_t B, const tob200_options &options,
                                    tob200_result *results, cudaStream_t stream = nullptr) {
  return detail::launch_warp<T, N, false, F>(f, x, B, options, results, stream);
}


// ================================================================================================
// Any n up to 2048, run-time n (the reference's dynamic-size path: `Optimize(x, residuals)` with a VecX, no size
// cap, math.h:232-240): the same functor source as the warp family above, evaluated by one warp per problem into
// MATERIALISED residual blocks J [B][m][n], r [B][m] in HBM, which then go through the SolverType seam of the
// library (tob200_solver_*: n <= 55 on the fused step kernels, above on the general family) in a host-driven
// Step loop — the structure of `Optimizer_::OptimizeAcc` itself (optimizer.h:243-327).  Unlike the kernels above
// this one links the library (-ltinyopt_b200).  Three ways to the Jacobian:
//   * OptimizeBatchAutoDiffLarge   — Jets (diff/optimize_autodiff.h:33-166): 64 parameters per sweep (every lane
//     carries d/dx_{base+lane}, d/dx_{base+lane+32}), ceil(n / 64) sweeps of the functor per rebuild pass;
//   * OptimizeBatchManualLarge     — the functor supplies its own rows: emit(r, [&](int j) { return J_ij; });
//   * OptimizeBatchNumDiffLarge    — numeric differentiation (diff/num_diff.h:57-126, 284-309: kForward, kCentral
//     (default), kFastCentral, h = FloatEpsilon = 1e-7f / 1e-4f): every lane perturbs ITS OWN parameter, so 32
//     columns of J come out of one lockstep sweep; the Cost is the residual NORM with m residuals, as the
//     reference's `CreateNumDiffFunc2` returns it (num_diff.h:300-305), handed over by tob200_solver_step_cost_*.
// Cost-only Steps (solvers/gn.h:98-105) sweep the functor once with plain T.  The functor must emit exactly m
// residuals per sweep and be free of lane-dependent control flow.
// ================================================================================================
enum NumDiffMethod { kForward = 0, kCentral = 1, kFastCentral = 2 };  // diff/num_diff.h:20-52

template <typename T>
struct LargeXJet {
  const T *xg;
  int lane, base;
  __device__ __forceinline__ Jet<T, 2> operator[](int j) const {
    Jet<T, 2> r;
    r.a = xg[j];
    r.v[0] = (base + lane == j) ? (T)1 : (T)0;
    r.v[1] = (base + lane + 32 == j) ? (T)1 : (T)0;
    return r;
  }
};
template <typename T>
struct LargeXScalar {
  const T *xg;
  __device__ __forceinline__ T operator[](int j) const { return xg[j]; }
};
template <typename T>
struct LargeXNum {  // PlusEq(y, dx) with dx = +-h e_mine (num_diff.h:96-113)
  const T *xg;
  int mine;
  T step;
  __device__ __forceinline__ T operator[](int j) const {
    const T v = xg[j];
    return j == mine ? Ops<T>::add(v, step) : v;
  }
};

template <typename T>
struct LargeXNumFast {  // kFastCentral's second point: y = (x + h e_mine) + (-2h e_mine)
  const T *xg;
  int mine;
  T h, m2h;
  __device__ __forceinline__ T operator[](int j) const {
    const T v = xg[j];
    return j == mine ? Ops<T>::add(Ops<T>::add(v, h), m2h) : v;
  }
};

template <typename T>
struct LargeEmit {
  enum Mode { kScalar = 0, kJet = 1, kManual = 2, kNumPlus = 3, kNumMinus = 4, kNumForward = 5 };
  T *Jp, *rp;  // the problem's J [m][n] and r [m]
  int n, lane, base, mode, row;
  T inv_den;   // numeric differentiation: the divisor 2h or h
  __device__ __forceinline__ void operator()(const Jet<T, 2> &r) {
    T *Jr = Jp + (size_t)row * n;
    if (base + lane < n) Jr[base + lane] = r.v[0];
    if (base + lane + 32 < n) Jr[base + lane + 32] = r.v[1];
    if (base == 0 && lane == 0) rp[row] = r.a;
    ++row;
  }
  __device__ __forceinline__ void operator()(T r) {
    const int c = base + lane;
    T *Je = Jp + (size_t)row * n + c;
    if (mode == kScalar) {
      if (lane == 0) rp[row] = r;
    } else if (c < n) {
      if (mode == kNumPlus) *Je = r;                                               // res_plus
      else if (mode == kNumMinus) *Je = Ops<T>::div(Ops<T>::sub(*Je, r), inv_den);  // (res_plus - res_minus) / (2 h)
      else *Je = Ops<T>::div(Ops<T>::sub(r, rp[row]), inv_den);                     // (res_plus - res) / h
    }
    ++row;
  }
  /// residual with its own Jacobian row: jf(j) = d r / d x_j (every lane passes the same r and the same jf)
  template <typename JF>
  __device__ __forceinline__ void operator()(T r, JF jf) {
    T *Jr = Jp + (size_t)row * n;
    for (int j = lane; j < n; j += 32) Jr[j] = jf(j);
    if (lane == 0) rp[row] = r;
    ++row;
  }
};

namespace detail {

// kind 0: Jets, 1: manual rows, 2: numeric differentiation
template <typename T, int kKind, typename F>
__global__ void __launch_bounds__(128) functor_eval_kernel(F f, const T *x, const int32_t *needs, T *J, T *r, int64_t B, int n,
                                                          int m, int method, T h) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t pr = warp0; pr < B; pr += nwarps) {
    const int need = needs[pr];
    if (need < 0) continue;  // finished
    const T *xg = x + (size_t)pr * n;
    LargeEmit<T> emit;
    emit.Jp = J + (size_t)pr * m * n;
    emit.rp = r + (size_t)pr * m;
    emit.n = n; emit.lane = lane; emit.base = 0; emit.row = 0; emit.inv_den = (T)1;
    if constexpr (kKind == 1) {
      emit.mode = need ? LargeEmit<T>::kManual : LargeEmit<T>::kScalar;
      f(pr, LargeXScalar<T>{xg}, emit, need != 0);
    } else if constexpr (kKind == 0) {
      if (!need) {
        emit.mode = LargeEmit<T>::kScalar;
        f(pr, LargeXScalar<T>{xg}, emit);
      } else {
        emit.mode = LargeEmit<T>::kJet;
        for (int base = 0; base < n; base += 64) {
          emit.base = base;
          emit.row = 0;
          f(pr, LargeXJet<T>{xg, lane, base}, emit);
        }
      }
    } else {
      emit.mode = LargeEmit<T>::kScalar;
      f(pr, LargeXScalar<T>{xg}, emit);  // `const auto res = residuals(x)` (num_diff.h:293)
      if (need) {
        __syncwarp();  // kNumForward reads lane 0's r
        const T two_h = Ops<T>::mul((T)2, h);
        for (int base = 0; base < n; base += 32) {
          const int mine = base + lane < n ? base + lane : -1;
          emit.base = base;
          emit.row = 0;
          if (method == kForward) {
            emit.mode = LargeEmit<T>::kNumForward;
            emit.inv_den = h;
            f(pr, LargeXNum<T>{xg, mine, h}, emit);
          } else {
            emit.mode = LargeEmit<T>::kNumPlus;
            f(pr, LargeXNum<T>{xg, mine, h}, emit);
            emit.row = 0;
            emit.mode = LargeEmit<T>::kNumMinus;
            emit.inv_den = two_h;
            if (method == kCentral) {
              f(pr, LargeXNum<T>{xg, mine, -h}, emit);  // y = x; dx[r] = -h
            } else {  // kFastCentral: y = (x + h) + (-2h) (num_diff.h:104-106)
              f(pr, LargeXNumFast<T>{xg, mine, h, Ops<T>::mul((T)-2, h)}, emit);
            }
          }
        }
      }
    }
    __syncwarp();
  }
}

// Cost(res.norm(), res.size()) of num_diff.h:305: sqrt of the canonical chain over the rows, one thread per problem
template <typename T>
__global__ void norm_cost_kernel(const T *r, const int32_t *needs, int64_t B, int m, double *cost) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B || needs[p] < 0) return;
  const T *rp = r + (size_t)p * m;
  T c = (T)0;
  for (int i = 0; i < m; ++i) c = Ops<T>::fma(rp[i], rp[i], c);
  cost[p] = (double)sqrt(c);
}

inline int large_step(tob200_solver *s, const float *J, const float *r, int m, const double *cost) {
  return cost ? tob200_solver_step_cost_f32(s, J, r, TOB200_LAYOUT_PROBLEM_MAJOR, m, cost)
              : tob200_solver_step_f32(s, J, r, TOB200_LAYOUT_PROBLEM_MAJOR, m);
}
inline int large_step(tob200_solver *s, const double *J, const double *r, int m, const double *cost) {
  return cost ? tob200_solver_step_cost_f64(s, J, r, TOB200_LAYOUT_PROBLEM_MAJOR, m, cost)
              : tob200_solver_step_f64(s, J, r, TOB200_LAYOUT_PROBLEM_MAJOR, m);
}

template <typename T, int kKind, typename F>
int optimize_large(tob200_ctx *ctx, const F &f, T *x, int64_t B, int n, int m, const tob200_options &options,
                   tob200_result *results, int method, T h) {
  if (!ctx || !x || !results || B < 0 || n < 1 || m < 1) return TOB200_ERR_INVALID;
  if (B == 0) return TOB200_OK;
  tob200_solver *s = nullptr;
  int rc = tob200_solver_create_ex(ctx, sizeof(T) == 4 ? TOB200_F32 : TOB200_F64, B, n, &options,
                                   kKind == 2 ? TOB200_SOLVER_GENERAL : 0, &s);
  if (rc != TOB200_OK) return rc;
  T *J = nullptr, *r = nullptr;
  double *cost = nullptr;
  auto done = [&](int code) {
    cudaFree(J); cudaFree(r); cudaFree(cost);
    tob200_solver_destroy(s);
    return code;
  };
  if (cudaMalloc(&J, (size_t)B * m * n * sizeof(T)) != cudaSuccess) return done(TOB200_ERR_NOMEM);
  if (cudaMalloc(&r, (size_t)B * m * sizeof(T)) != cudaSuccess) return done(TOB200_ERR_NOMEM);
  if (kKind == 2 && cudaMalloc(&cost, (size_t)B * sizeof(double)) != cudaSuccess) return done(TOB200_ERR_NOMEM);
  if ((rc = tob200_solver_reset(s, x)) != TOB200_OK) return done(rc);
  const T *xs = static_cast<const T *>(tob200_solver_x(s));
  const int32_t *needs = tob200_solver_needs(s);
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t grid = (B + 3) / 4;
  if (grid > 8 * (int64_t)sms) grid = 8 * (int64_t)sms;
  const int max_steps = options.max_iters + 2;  // optimizer.h:248-250
  for (int step = 0; step < max_steps; ++step) {
    int64_t active = 0;
    if ((rc = tob200_solver_num_active(s, &active)) != TOB200_OK) return done(rc);  // synchronises the library's stream
    if (active == 0) break;
    functor_eval_kernel<T, kKind, F><<<(unsigned)grid, 128>>>(f, xs, needs, J, r, B, n, m, method, h);
    if (kKind == 2) norm_cost_kernel<T><<<(unsigned)((B + 127) / 128), 128>>>(r, needs, B, m, cost);
    if (cudaDeviceSynchronize() != cudaSuccess) return done(TOB200_ERR_CUDA);
    if ((rc = large_step(s, J, r, m, cost)) != TOB200_OK) return done(rc);
  }
  if ((rc = tob200_solver_results(s, results)) != TOB200_OK) return done(rc);
  if ((rc = tob200_sync(ctx)) != TOB200_OK) return done(rc);
  if (cudaMemcpy(x, xs, (size_t)B * n * sizeof(T), cudaMemcpyDeviceToDevice) != cudaSuccess) return done(TOB200_ERR_CUDA);
  return done(TOB200_OK);
}

}  // namespace detail

/// tinyopt::Optimize(x, residuals, options) with automatic differentiation, any n <= 2048 (run-time n, m residuals).
/// f : `template <typename X, typename E> __device__ void operator()(int64_t p, const X &x, E &emit) const`
/// Returns a tob200_status; x [B][n] and results [B] are device arrays.
template <typename T, typename F>
int OptimizeBatchAutoDiffLarge(tob200_ctx *ctx, const F &f, T *x, int64_t B, int n, int m, const tob200_options &options,
                               tob200_result *results) {
  return detail::optimize_large<T, 0, F>(ctx, f, x, B, n, m, options, results, 0, (T)0);
}
/// The accumulation contract with the functor's own Jacobian rows, any n <= 2048.
/// f : `template <typename X, typename E> __device__ void operator()(int64_t p, const X &x, E &emit, bool want_jacobian) const`
///     — emit(r, [&](int j) { return J_ij; }) per residual, or emit(r) when !want_jacobian
template <typename T, typename F>
int OptimizeBatchManualLarge(tob200_ctx *ctx, const F &f, T *x, int64_t B, int n, int m, const tob200_options &options,
                             tob200_result *results) {
  return detail::optimize_large<T, 1, F>(ctx, f, x, B, n, m, options, results, 0, (T)0);
}
/// tinyopt::Optimize with numeric differentiation (diff/num_diff.h: `CreateNumDiffFunc2(x, residuals, method, h)` handed
/// to the optimizer), any n <= 2048.  f as for OptimizeBatchAutoDiffLarge (it is only ever called with plain T).
/// h <= 0 selects the reference's default FloatEpsilon<T>() (math.h:297-301: 1e-7f in double, 1e-4f in float).
template <typename T, typename F>
int OptimizeBatchNumDiffLarge(tob200_ctx *ctx, const F &f, T *x, int64_t B, int n, int m, const tob200_options &options,
                              tob200_result *results, NumDiffMethod method = kCentral, T h = (T)0) {
  if (!(h > (T)0)) h = sizeof(T) == 4 ? (T)1e-4f : (T)1e-7f;
  return detail::optimize_large<T, 2, F>(ctx, f, x, B, n, m, options, results, (int)method, h);
}

}  // namespace device
}  // namespace b200
}  // namespace tinyopt
