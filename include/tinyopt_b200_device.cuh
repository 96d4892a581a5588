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
  return {(T)2 * th * n - th2, th / n};
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
  return {th * atan(n2 / th), (T)1 / (n2 * n2 / th2 + (T)1)};  // th > 0: atan2(n2, th) == atan(n2 / th)
}
/// robust_norms.h:204-228: loss = th2 log(1 + n2 / th2), scale = 1 / (1 + n2 / th2)
template <typename S, typename T>
__host__ __device__ inline Robust<S> Cauchy(const S &n2, T th2) {
  const S s = (T)1 + n2 / th2;
  return {th2 * log(s), (T)1 / s};
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
  if (options.use_ldlt != 1) return cudaErrorInvalidValue;  // only the LDLT path (options.h:59) exists
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

}  // namespace device
}  // namespace b200
}  // namespace tinyopt
