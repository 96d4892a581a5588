// ldlt_reg.cuh — per-thread dense diagonal-pivoted LDL^T on a register-resident packed triangle.
//
// Replaces `tinyopt::SolveLDLT` (include/tinyopt/math.h:232-240), i.e. Eigen 3.4
// `A.selfadjointView<Upper>().ldlt()` + info()/isPositive() + solve(b), for compile-time n.
// Semantics restated from Eigen's published algorithm: unblocked, in place,
// largest-|diagonal| pivot (first maximum wins), sign tracking, failure only when a non-zero pivot
// follows a zero pivot, pseudo-inverse of D in the solve.  Every loop is fully unrolled so that all
// indices into the triangle are static (registers); the run-time pivot row is handled by
// conditional selects over the unrolled swap pattern (`cswp`), never by a branch.
#pragma once

#include "common.cuh"

namespace tob200 {

template <typename T, int N>
struct LdltReg {
  static constexpr int NT = tri_count(N);
  using O = Ops<T>;

  // element (i, j), j <= i, of the lower-triangular working matrix == upper (j, i) of H
  static __device__ __forceinline__ constexpr int idx(int i, int j) { return tri_index(N, j, i); }

  // conditional swap as two selects: the run-time pivot row never becomes a branch, so the lanes of
  // a warp (32 different problems, 32 different pivots) do not diverge
  static __device__ __forceinline__ void cswp(bool c, T &a, T &b) {
    const T ta = c ? b : a;
    const T tb = c ? a : b;
    a = ta;
    b = tb;
  }

  // In-place factorisation.  Returns true iff info()==Success && isPositive().
  static __device__ __forceinline__ bool factor(T (&w)[NT], int (&tr)[N]) {
    if (N == 1) {
      tr[0] = 0;
      return !(w[0] < (T)0);  // PositiveSemiDef or ZeroSign (NaN compares false both ways -> Zero)
    }
    int sign = 0;  // 0 Zero, 1 PositiveSemiDef, -1 NegativeSemiDef, 2 Indefinite
    bool found_zero_pivot = false, ret = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      int p = k;
      T best = O::abs(w[idx(k, k)]);
#pragma unroll
      for (int i = k + 1; i < N; ++i) {
        const T v = O::abs(w[idx(i, i)]);
        if (v > best) {
          best = v;
          p = i;
        }
      }
      tr[k] = p;
#pragma unroll
      for (int pp = k + 1; pp < N; ++pp) {
        const bool c = (p == pp);
#pragma unroll
        for (int j = 0; j < k; ++j) cswp(c, w[idx(k, j)], w[idx(pp, j)]);
#pragma unroll
        for (int i = pp + 1; i < N; ++i) cswp(c, w[idx(i, k)], w[idx(i, pp)]);
        cswp(c, w[idx(k, k)], w[idx(pp, pp)]);
#pragma unroll
        for (int i = k + 1; i < pp; ++i) cswp(c, w[idx(i, k)], w[idx(pp, i)]);
      }
      if (k > 0) {
        T temp[N];
#pragma unroll
        for (int j = 0; j < k; ++j) temp[j] = O::mul(w[idx(j, j)], w[idx(k, j)]);
        {
          T s = (T)0;
#pragma unroll
          for (int j = 0; j < k; ++j) s = O::fma(w[idx(k, j)], temp[j], s);
          w[idx(k, k)] = O::sub(w[idx(k, k)], s);
        }
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
          T s = (T)0;
#pragma unroll
          for (int j = 0; j < k; ++j) s = O::fma(w[idx(i, j)], temp[j], s);
          w[idx(i, k)] = O::sub(w[idx(i, k)], s);
        }
      }
      const T akk = w[idx(k, k)];
      const bool pivot_is_valid = O::abs(akk) > (T)0;
      if (k == 0 && !pivot_is_valid) {
        // the whole diagonal is zero: success iff the strict triangle is zero as well
#pragma unroll
        for (int j = 0; j < N; ++j) {
          tr[j] = j;
#pragma unroll
          for (int i = j + 1; i < N; ++i) ret = ret && (w[idx(i, j)] == (T)0);
        }
        return ret;  // sign == ZeroSign -> isPositive()
      }
      if (k < N - 1) {
        if (pivot_is_valid) {
#pragma unroll
          for (int i = k + 1; i < N; ++i) w[idx(i, k)] = O::div(w[idx(i, k)], akk);
        } else {
#pragma unroll
          for (int i = k + 1; i < N; ++i) ret = ret && (w[idx(i, k)] == (T)0);
        }
      }
      if (found_zero_pivot && pivot_is_valid) ret = false;
      else if (!pivot_is_valid) found_zero_pivot = true;

      if (sign == 1) { if (akk < (T)0) sign = 2; }
      else if (sign == -1) { if (akk > (T)0) sign = 2; }
      else if (sign == 0) { if (akk > (T)0) sign = 1; else if (akk < (T)0) sign = -1; }
    }
    return ret && (sign == 1 || sign == 0);
  }

  // y <- P^T L^-T D^+ L^-1 P y
  static __device__ __forceinline__ void solve(const T (&w)[NT], const int (&tr)[N], T (&y)[N]) {
    if (N > 1) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int pp = k + 1; pp < N; ++pp) cswp(tr[k] == pp, y[k], y[pp]);
      }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      T s = y[i];
#pragma unroll
      for (int j = 0; j < i; ++j) s = O::fma(-w[idx(i, j)], y[j], s);
      y[i] = s;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const T d = w[idx(i, i)];
      y[i] = (O::abs(d) > O::min_normal()) ? O::div(y[i], d) : (T)0;
    }
    // L^T y = y, updates applied to y_i in the order j = N-1 .. i+1 (canonical order)
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
      T s = y[i];
#pragma unroll
      for (int j = N - 1; j > i; --j) s = O::fma(-w[idx(j, i)], y[j], s);
      y[i] = s;
    }
    if (N > 1) {
#pragma unroll
      for (int k = N - 1; k >= 0; --k) {
#pragma unroll
        for (int pp = k + 1; pp < N; ++pp) cswp(tr[k] == pp, y[k], y[pp]);
      }
    }
  }
};

// `hessian.use_ldlt = false` (include/tinyopt/solvers/gn.h:157-163): dx = -H^-1 g with no check on
// invertibility, restated as a partial-pivot LU solve in the canonical order of the oracle
// (right-looking elimination, first largest |entry| of the column is the pivot, upper solve with the
// updates of x_i applied for j = n-1 .. i+1).  The option is off the hot path, so the full matrix
// lives in thread-local memory behind a real call and the default LDLT path keeps its registers.
template <typename T, int N>
__device__ __noinline__ void lu_solve_local(T *M, T *b) {
  using O = Ops<T>;
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    int p = k;
    T best = O::abs(M[k * N + k]);
#pragma unroll 1
    for (int i = k + 1; i < N; ++i) {
      const T v = O::abs(M[i * N + k]);
      if (v > best) {
        best = v;
        p = i;
      }
    }
    if (p != k) {
#pragma unroll 1
      for (int j = 0; j < N; ++j) {
        const T t = M[k * N + j];
        M[k * N + j] = M[p * N + j];
        M[p * N + j] = t;
      }
      const T t = b[k];
      b[k] = b[p];
      b[p] = t;
    }
    const T piv = M[k * N + k];
#pragma unroll 1
    for (int i = k + 1; i < N; ++i) {
      const T f = O::div(M[i * N + k], piv);
#pragma unroll 1
      for (int j = k + 1; j < N; ++j) M[i * N + j] = O::fma(-f, M[k * N + j], M[i * N + j]);
      b[i] = O::fma(-f, b[k], b[i]);
    }
  }
#pragma unroll 1
  for (int i = N - 1; i >= 0; --i) {
    T s = b[i];
#pragma unroll 1
    for (int j = N - 1; j > i; --j) s = O::fma(-M[i * N + j], b[j], s);
    b[i] = O::div(s, M[i * N + i]);
  }
}

// hu: packed upper triangle of the (damped) symmetric H_, g: gradient -> dx
template <typename T, int N>
__device__ __forceinline__ void solve_inverse_reg(const T (&hu)[tri_count(N)], const T (&g)[N], T (&dx)[N]) {
  using O = Ops<T>;
  if (N == 1) {  // gn.h:158-160: guarded scalar inverse, zero step otherwise
    dx[0] = hu[0] > O::float_eps() ? O::mul(-O::div((T)1, hu[0]), g[0]) : (T)0;
    return;
  }
  T M[N * N], b[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = 0; j < N; ++j) M[i * N + j] = hu[tri_index(N, i < j ? i : j, i < j ? j : i)];
    b[i] = -g[i];
  }
  lu_solve_local<T, N>(M, b);
#pragma unroll
  for (int i = 0; i < N; ++i) dx[i] = b[i];
}

}  // namespace tob200
