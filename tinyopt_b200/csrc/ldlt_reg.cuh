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

}  // namespace tob200
