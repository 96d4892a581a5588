// gn.cuh — the GENERAL kernel family (family 4): any scalar type, any n up to kGnMaxN.
//
// It covers what the specialised families do not: double precision above n = 55 (tinyopt's default
// scalar is double, optimize.h:20-33) and n above 512 in either precision (the reference's dynamic-size
// solver has no size cap, math.h:232-240).  One CTA per problem, matrices in HBM / L2, and the CANONICAL
// operation sequence of the CPU oracle kept exactly — every sum a left-to-right chain of
// IEEE fma in the problem's scalar type, rows in order — so the results are bit-identical to the oracle in
// float and in double: the 1e-10 bar of the north star holds with margin.  It is a coverage path, not a
// roofline path: no tensor cores (they would break the double / bit-exact contract), plain FMA tiles.
//
//   gn_accum_kernel  a1 / a3: t = A x (chain over the columns), r, the Jacobian row scale, cost (chain over
//                    the rows), g = J^T r and the upper triangle of H = J^T J as 64 x 64 tiles of 4 x 4
//                    register blocks whose accumulators take their rows in order i = 0 .. m-1
//                    (diff/optimize_autodiff.h:151-164, solvers/gn.h:77-113).
//   gn_solve_kernel  a4 .. a10: Build's tail and damping (solvers/lm.h:60-120), Eigen's pivot order, the
//                    unblocked left-looking LDL^T of the oracle with thread = row and a column-major factor
//                    (every dot product the oracle's chain j = 0 .. k-1), LDLT::solve, or the partial-pivot LU
//                    of `use_ldlt = false` (solvers/gn.h:157-163), then Step / OptimizeAcc (lm_state.cuh).
#pragma once

#include "common.cuh"
#include "lm_state.cuh"

namespace tob200 {

constexpr int kGnMaxN = 2048;
constexpr int kGnThreads = 256;
constexpr int kGnTile = 64;  // H tile edge: 256 threads x (4 x 4) accumulators
constexpr int kGnRows = 16;  // rows of [J] staged per step of the tile loop

template <typename T>
struct GnAccumParams {
  const T *A;   // [B][m][n] (J when synth == 0)
  const T *y;   // [B][m]    (r when synth == 0)
  const T *x;   // [B][n]
  const LmScalars<T> *rec;  // per-problem state or nullptr (every problem, rebuild pass)
  T *rs;        // [B][2][m] scratch: r_i, then the Jacobian row scale s_i
  T *g;         // [B][n]
  T *H;         // [B][n][n]: upper triangle (row <= col) written
  T *cost;      // [B]
  int64_t B;
  int m, n;
  int synth, is_lm;
  T alpha, alpha3;
};

// one CTA per problem (grid-stride)
template <typename T>
__global__ void __launch_bounds__(kGnThreads) gn_accum_kernel(const __grid_constant__ GnAccumParams<T> p) {
  using O = Ops<T>;
  extern __shared__ __align__(16) unsigned char gn_smem[];
  T *xs = reinterpret_cast<T *>(gn_smem);                    // [n]
  T *ta = xs + p.n;                                          // [kGnRows][kGnTile + 1] tile of block column bi
  T *tb = ta + kGnRows * (kGnTile + 1);                      // ... of block column bj
  T *tile = tb + kGnRows * (kGnTile + 1);                    // [kGnThreads][33]: A tile of the t-chain pass
  const int tid = threadIdx.x, m = p.m, n = p.n;
  for (int64_t pr = blockIdx.x; pr < p.B; pr += gridDim.x) {
    bool do_rebuild = true;
    if (p.rec) {
      const LmScalars<T> s = p.rec[pr];
      if (s.done()) continue;
      do_rebuild = !p.is_lm || s.rebuild();
    }
    const T *Ap = p.A + (size_t)pr * m * n, *yp = p.y + (size_t)pr * m;
    T *rp = p.rs + (size_t)pr * 2 * m, *sp = rp + m;
    __syncthreads();
    for (int j = tid; j < n; j += kGnThreads) xs[j] = p.synth ? p.x[(size_t)pr * n + j] : (T)0;
    __syncthreads();
    // ---- rows: t_i = a_i . x (canonical chain over j), r_i, s_i.  thread = row of a block of 256 rows; the
    // ---- columns stream through a [256][32] shared tile so that the global loads are coalesced ----
    for (int i0 = 0; i0 < m; i0 += kGnThreads) {
      const int i = i0 + tid;
      T t = (T)0;
      if (p.synth) {
        for (int j0 = 0; j0 < n; j0 += 32) {
          __syncthreads();
          for (int e = tid; e < kGnThreads * 32; e += kGnThreads) {
            const int rr = e >> 5, c = e & 31;
            tile[rr * 33 + c] = (i0 + rr < m && j0 + c < n) ? Ap[(size_t)(i0 + rr) * n + j0 + c] : (T)0;
          }
          __syncthreads();
          const int jn = n - j0 < 32 ? n - j0 : 32;
          for (int c = 0; c < jn; ++c) t = O::fma(tile[tid * 33 + c], xs[j0 + c], t);
        }
      }
      if (i < m) {
        T ri, sc = (T)1;
        if (p.synth) {
          const T t2 = O::mul(t, t);
          ri = O::fma(t, O::fma(p.alpha, t2, (T)1), -yp[i]);
          sc = O::fma(p.alpha3, t2, (T)1);
        } else {
          ri = yp[i];
        }
        rp[i] = ri;
        sp[i] = sc;
      }
    }
    __syncthreads();
    // ---- cost = sum r_i^2, rows in order: one chain ----
    if (tid == 0) {
      T c = (T)0;
      for (int i = 0; i < m; ++i) c = O::fma(rp[i], rp[i], c);
      p.cost[pr] = c;
    }
    if (!do_rebuild) continue;
    // ---- g_j = sum_i J_ij r_i, rows in order; thread = column (coalesced) ----
    for (int j = tid; j < n; j += kGnThreads) {
      T gj = (T)0;
      for (int i = 0; i < m; ++i) {
        const T a = Ap[(size_t)i * n + j];
        gj = O::fma(p.synth ? O::mul(sp[i], a) : a, rp[i], gj);
      }
      p.g[(size_t)pr * n + j] = gj;
    }
    // ---- H = J^T J, upper tiles; 256 threads x 4 x 4 accumulators, rows in order ----
    const int nb = (n + kGnTile - 1) / kGnTile;
    const int tr = tid >> 4, tc = tid & 15;  // my 4 x 4 block: rows 4 tr .., columns 4 tc ..
    T *Hp = p.H + (size_t)pr * n * n;
    for (int bi = 0; bi < nb; ++bi) {
      for (int bj = bi; bj < nb; ++bj) {
        T acc[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[u][v] = (T)0;
        for (int i0 = 0; i0 < m; i0 += kGnRows) {
          __syncthreads();
          for (int e = tid; e < kGnRows * kGnTile; e += kGnThreads) {
            const int rr = e / kGnTile, c = e % kGnTile, i = i0 + rr;
            const int ja = bi * kGnTile + c, jb = bj * kGnTile + c;
            T va = (T)0, vb = (T)0;
            if (i < m) {
              const T sc = sp[i];
              if (ja < n) { va = Ap[(size_t)i * n + ja]; if (p.synth) va = O::mul(sc, va); }
              if (jb < n) { vb = Ap[(size_t)i * n + jb]; if (p.synth) vb = O::mul(sc, vb); }
            }
            ta[rr * (kGnTile + 1) + c] = va;
            tb[rr * (kGnTile + 1) + c] = vb;
          }
          __syncthreads();
          const int rn = m - i0 < kGnRows ? m - i0 : kGnRows;
          for (int rr = 0; rr < rn; ++rr) {
            T a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              a[u] = ta[rr * (kGnTile + 1) + 4 * tr + u];
              b[u] = tb[rr * (kGnTile + 1) + 4 * tc + u];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int v = 0; v < 4; ++v) acc[u][v] = O::fma(a[u], b[v], acc[u][v]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int row = bi * kGnTile + 4 * tr + u, col = bj * kGnTile + 4 * tc + v;
            if (row <= col && col < n) Hp[(size_t)row * n + col] = acc[u][v];
          }
      }
    }
  }
}

__host__ __device__ inline size_t gn_accum_smem_bytes(int n, size_t es) {
  return ((size_t)n + 2 * kGnRows * (kGnTile + 1) + (size_t)kGnThreads * 33) * es + 16;
}

template <typename T>
struct GnSolveParams {
  T *H;         // [B][n][n] upper triangle, undamped diagonal after a rebuild
  T *hd;        // [B][n] persistent damped diagonal of H_ (solvers/lm.h keeps H_ damped), or nullptr
  T *g;         // [B][n] grad_
  const T *cost;  // [B]
  T *W;         // [grid][n][n] factor workspace
  int64_t B;
  int n, nres;
  int mode;     // 0: LM loop step, 1: one Build + Solve (lambda / dx / status)
  DevOptions<T> opt;
  LmScalars<T> *rec;
  T *x, *last_dx;
  tob200_result *results;
  unsigned long long *n_active;
  const T *lambda;
  T *dx;
  double *cost_out;
  int32_t *status;
  int use_ldlt;
  // the SolverType seam above n = 55 (tob200_solver_*, family 4): what the next Step needs from the caller, and the
  // cost / residual count of user-filled accumulators (tob200_solver_step_hg_*: Cost::cost is a double, cost.h:93)
  int32_t *needs = nullptr;           // [B] or nullptr
  const double *cost_d = nullptr;     // [B] or nullptr: replaces cost[]
  const int32_t *nres_arr = nullptr;  // [B] or nullptr: replaces nres
  // mode 3 (InvCov, math.h:44-57): H -> cov = LDLT(H).solve(Identity), MaxStdDev (solvers/lm.h:176-187)
  T *Y = nullptr;        // [grid][n][n] second workspace
  T *cov = nullptr;      // [B][n][n] or nullptr
  T *max_std = nullptr;  // [B] or nullptr
};

// shared memory of gn_solve_kernel, in scalars: dd | temp | rhs / y | keys(perm, inv as ints share the tail)
__host__ __device__ inline size_t gn_solve_smem_bytes(int n, size_t es) { return (size_t)n * (3 * es + 8) + 64; }

// Eigen's pivot order on the diagonal dd (see wpp_pivot_order): position = number of larger |d| when all keys
// are distinct and none is NaN, else the literal replay by one thread.
template <typename T>
__device__ void gn_pivot_order(const T *dd, int n, int *perm, int *inv, int *flag) {
  using O = Ops<T>;
  const int tid = threadIdx.x;
  if (tid == 0) *flag = 0;
  __syncthreads();
  for (int i = tid; i < n; i += kGnThreads) {
    const T a = O::abs(dd[i]);
    int gt = 0, eq = 0;
    for (int j = 0; j < n; ++j) {
      const T v = O::abs(dd[j]);
      gt += v > a;
      eq += v == a;
    }
    if (eq != 1) *flag = 1;  // a tie, or a NaN (which equals nothing, not even itself)
    else { perm[gt] = i; inv[i] = gt; }
  }
  __syncthreads();
  if (*flag) {
    if (tid == 0) {
      for (int i = 0; i < n; ++i) perm[i] = i;
      for (int k = 0; k < n; ++k) {
        int pb = k;
        T best = O::abs(dd[perm[k]]);
        for (int i = k + 1; i < n; ++i) {
          const T v = O::abs(dd[perm[i]]);
          if (v > best) { best = v; pb = i; }
        }
        const int t = perm[k]; perm[k] = perm[pb]; perm[pb] = t;
      }
      for (int k = 0; k < n; ++k) inv[perm[k]] = k;
    }
    __syncthreads();
  }
}

// unpivoted left-looking LDL^T of the permuted matrix, W column-major lower (W(i, j) at W[j * n + i], i >= j):
// the oracle's ldlt_factor_ with the rows spread over the threads.  Returns info()==Success && isPositive().
template <typename T>
__device__ bool gn_ldlt_factor(T *W, int n, T *temp, int *misc) {
  using O = Ops<T>;
  const int tid = threadIdx.x;
#define GW(i, j) W[(size_t)(j) * n + (i)]
  if (n == 1) return !(GW(0, 0) < (T)0);
  if (tid == 0) { misc[0] = 0; misc[1] = 0; misc[2] = 1; misc[3] = 0; }  // sign, found_zero_pivot, ret, early
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    if (k > 0) {
      for (int j = tid; j < k; j += kGnThreads) temp[j] = O::mul(GW(j, j), GW(k, j));
      __syncthreads();
      for (int i = k + tid; i < n; i += kGnThreads) {
        T s = (T)0;
        for (int j = 0; j < k; ++j) s = O::fma(GW(i, j), temp[j], s);
        GW(i, k) = O::sub(GW(i, k), s);
      }
      __syncthreads();
    }
    const T akk = GW(k, k);
    const bool valid = O::abs(akk) > (T)0;
    if (k == 0 && !valid) {  // the whole diagonal is zero: success iff the strict triangle is zero too
      int nz = 0;
      for (int j = 0; j < n; ++j)
        for (int i = j + 1 + tid; i < n; i += kGnThreads) nz |= (GW(i, j) != (T)0);
      return !__syncthreads_or(nz);
    }
    if (k < n - 1) {
      int nz = 0;
      for (int i = k + 1 + tid; i < n; i += kGnThreads) {
        if (valid) GW(i, k) = O::div(GW(i, k), akk);
        else nz |= (GW(i, k) != (T)0);
      }
      if (!valid) { if (__syncthreads_or(nz) && tid == 0) misc[2] = 0; }
    }
    if (tid == 0) {
      if (misc[1] && valid) misc[2] = 0;
      else if (!valid) misc[1] = 1;
      int sign = misc[0];
      if (sign == 1) { if (akk < (T)0) sign = 2; }
      else if (sign == -1) { if (akk > (T)0) sign = 2; }
      else if (sign == 0) { if (akk > (T)0) sign = 1; else if (akk < (T)0) sign = -1; }
      misc[0] = sign;
    }
    __syncthreads();
  }
  const bool ok = misc[2] && (misc[0] == 1 || misc[0] == 0);
  __syncthreads();
  return ok;
#undef GW
}

// ys (shared, position order) holds P b on entry; on exit x (original order) = P^T L^-T D^+ L^-1 P b
template <typename T>
__device__ void gn_ldlt_solve(const T *W, int n, const int *perm, T *ys, T *xout) {
  using O = Ops<T>;
  const int tid = threadIdx.x;
#define GW(i, j) W[(size_t)(j) * n + (i)]
  for (int j = 0; j < n; ++j) {  // L y = y: y_i takes its updates in the order j = 0 .. i-1
    const T yj = ys[j];
    __syncthreads();
    for (int i = j + 1 + tid; i < n; i += kGnThreads) ys[i] = O::fma(-GW(i, j), yj, ys[i]);
    __syncthreads();
  }
  for (int i = tid; i < n; i += kGnThreads) {
    const T d = GW(i, i);
    ys[i] = (O::abs(d) > O::min_normal()) ? O::div(ys[i], d) : (T)0;
  }
  __syncthreads();
  for (int j = n - 1; j >= 0; --j) {  // L^T y = y: y_i takes its updates in the order j = n-1 .. i+1
    const T yj = ys[j];
    __syncthreads();
    for (int i = tid; i < j; i += kGnThreads) ys[i] = O::fma(-GW(j, i), yj, ys[i]);
    __syncthreads();
  }
  for (int i = tid; i < n; i += kGnThreads) xout[perm[i]] = ys[i];
  __syncthreads();
#undef GW
}

// `hessian.use_ldlt = false` (solvers/gn.h:157-163): x = -H^-1 g by the oracle's partial-pivot LU (solve_inverse_).
// M row-major full n x n in W (destroyed); b (shared) holds -g on entry and x on exit.
template <typename T>
__device__ void gn_lu_solve(T *M, int n, T *b, T *fcol, int *misc) {
  using O = Ops<T>;
  const int tid = threadIdx.x;
  __shared__ int s_p;
  for (int k = 0; k < n; ++k) {
    if (tid == 0) {  // first largest |entry| of column k at or below the diagonal (a NaN diagonal keeps p = k)
      int pb = k;
      T best = O::abs(M[(size_t)k * n + k]);
      for (int i = k + 1; i < n; ++i) {
        const T v = O::abs(M[(size_t)i * n + k]);
        if (v > best) { best = v; pb = i; }
      }
      s_p = pb;
    }
    __syncthreads();
    const int pv = s_p;
    if (pv != k) {
      for (int j = tid; j < n; j += kGnThreads) {
        const T t = M[(size_t)k * n + j];
        M[(size_t)k * n + j] = M[(size_t)pv * n + j];
        M[(size_t)pv * n + j] = t;
      }
      if (tid == 0) { const T t = b[k]; b[k] = b[pv]; b[pv] = t; }
      __syncthreads();
    }
    const T piv = M[(size_t)k * n + k];
    for (int i = k + 1 + tid; i < n; i += kGnThreads) fcol[i] = O::div(M[(size_t)i * n + k], piv);
    __syncthreads();
    const int w = n - k - 1;  // trailing block (k+1.., k+1..) plus the rhs as column n
    for (int e = tid; e < w * (w + 1); e += kGnThreads) {
      const int i = k + 1 + e / (w + 1), c = e % (w + 1);
      const T f = fcol[i];
      if (c == w) b[i] = O::fma(-f, b[k], b[i]);
      else M[(size_t)i * n + k + 1 + c] = O::fma(-f, M[(size_t)k * n + k + 1 + c], M[(size_t)i * n + k + 1 + c]);
    }
    __syncthreads();
  }
  (void)misc;
  // U x = b from the last row, x_i taking its updates in the order j = n-1 .. i+1, then the division
  for (int j = n - 1; j >= 0; --j) {
    if (tid == 0) b[j] = O::div(b[j], M[(size_t)j * n + j]);
    __syncthreads();
    const T xj = b[j];
    for (int i = tid; i < j; i += kGnThreads) b[i] = O::fma(-M[(size_t)i * n + j], xj, b[i]);
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(kGnThreads) gn_solve_kernel(const __grid_constant__ GnSolveParams<T> p) {
  using O = Ops<T>;
  extern __shared__ __align__(16) unsigned char gn_smem[];
  const int tid = threadIdx.x, n = p.n;
  T *dd = reinterpret_cast<T *>(gn_smem);
  T *temp = dd + n;
  T *ys = temp + n;
  int *perm = reinterpret_cast<int *>(ys + n);
  int *inv = perm + n;
  T *W = p.W + (size_t)blockIdx.x * n * n;
  __shared__ LmScalars<T> s;
  __shared__ int sh_act, sh_flag, sh_built_ok, misc[4];
  __shared__ double sh_cost, sh_norm[2];
  unsigned long long local_active = 0;
  for (int64_t pr = blockIdx.x; pr < p.B; pr += gridDim.x) {
    T *Hp = p.H + (size_t)pr * n * n;
    T *hd = p.hd ? p.hd + (size_t)pr * n : nullptr;
    T *gp = p.g + (size_t)pr * n;
    bool pass_rebuilt = true;
    __syncthreads();
    if (p.mode == 3) {  // InvCov(H) = H.ldlt().solve(Identity) (math.h:44-57), MaxStdDev (solvers/lm.h:176-187)
      T *Cp = p.cov ? p.cov + (size_t)pr * n * n : nullptr;
      if (n == 1) {  // m.inverse() of a 1 x 1 (math.h:49-50), unprotected as in the reference
        if (tid == 0) {
          const T v = O::div((T)1, Hp[0]);
          if (Cp) Cp[0] = v;
          if (p.max_std) p.max_std[pr] = sqrt(v);
          p.status[pr] = 0;
        }
        continue;
      }
      for (int j = tid; j < n; j += kGnThreads) dd[j] = Hp[(size_t)j * n + j];
      __syncthreads();
      gn_pivot_order<T>(dd, n, perm, inv, &sh_flag);
      for (int64_t e = tid; e < (int64_t)n * n; e += kGnThreads) {
        const int b = (int)(e / n), a = (int)(e % n);
        if (a < b) continue;
        const int ia = perm[a], ib = perm[b];
        W[e] = ia < ib ? Hp[(size_t)ia * n + ib] : Hp[(size_t)ib * n + ia];
      }
      __syncthreads();
      const bool ok = gn_ldlt_factor<T>(W, n, temp, misc);
      T best = (T)0;
      bool have = false;
      if (ok) {
        // every column of the identity at once: thread = right-hand side c, Y(i, c) at Y[i * n + c] (coalesced over
        // c), each entry taking its updates in the oracle's order (ldlt_solve_: j = 0 .. i-1, then j = n-1 .. i+1)
        T *Y = p.Y + (size_t)blockIdx.x * n * n;
#define GW(i, j) W[(size_t)(j) * n + (i)]
        for (int c = tid; c < n; c += kGnThreads) {
          for (int i = 0; i < n; ++i) {
            T sacc = perm[i] == c ? (T)1 : (T)0;  // P e_c
            for (int j = 0; j < i; ++j) sacc = O::fma(-GW(i, j), Y[(size_t)j * n + c], sacc);
            Y[(size_t)i * n + c] = sacc;
          }
          for (int i = 0; i < n; ++i) {
            const T d = GW(i, i);
            const T v = Y[(size_t)i * n + c];
            Y[(size_t)i * n + c] = (O::abs(d) > O::min_normal()) ? O::div(v, d) : (T)0;
          }
          for (int i = n - 1; i >= 0; --i) {
            T sacc = Y[(size_t)i * n + c];
            for (int j = n - 1; j > i; --j) sacc = O::fma(-GW(j, i), Y[(size_t)j * n + c], sacc);
            Y[(size_t)i * n + c] = sacc;
          }
          for (int i = 0; i < n; ++i) {  // x = P^T y
            const T v = Y[(size_t)i * n + c];
            if (Cp) Cp[(size_t)perm[i] * n + c] = v;
            if (!have || v > best) { best = v; have = true; }  // maxCoeff
          }
        }
#undef GW
      }
      if (p.max_std) {
        __shared__ T red_best[kGnThreads];
        __shared__ int red_have[kGnThreads];
        red_best[tid] = best;
        red_have[tid] = have;
        __syncthreads();
        if (tid == 0) {
          for (int t = 1; t < kGnThreads; ++t)
            if (red_have[t] && (!have || red_best[t] > best)) { best = red_best[t]; have = true; }
          p.max_std[pr] = ok ? sqrt(best) : (T)0;
        }
      }
      if (tid == 0) p.status[pr] = ok ? 0 : 1;
      continue;
    }
    if (p.mode == 0) {
      if (tid == 0) s = p.rec[pr];
      __syncthreads();
      if (s.done()) continue;
      pass_rebuilt = p.opt.solver_type != 0 || s.rebuild();
    }
    double cost = 0.0;
    bool built_ok = true;
    const int nres = p.nres_arr ? p.nres_arr[pr] : p.nres;
    if (p.mode == 0) {
      if (tid == 0) {
        double c;
        const bool ok = p.cost_d ? lm_normalize_cost_d(p.opt, p.cost_d[pr], nres, c)
                                 : lm_normalize_cost(p.opt, p.cost[pr], nres, c);
        if (pass_rebuilt) s.num_builds++;
        sh_cost = c;
        sh_built_ok = ok;
      }
      __syncthreads();
      cost = sh_cost;
      built_ok = sh_built_ok;
      if (pass_rebuilt && built_ok) {
        if (p.opt.grad_clipping != (T)0)
          for (int j = tid; j < n; j += kGnThreads) {
            T v = gp[j];
            v = v < -p.opt.grad_clipping ? -p.opt.grad_clipping : v;
            v = v > p.opt.grad_clipping ? p.opt.grad_clipping : v;
            gp[j] = v;
          }
        if (p.opt.check_min_H_diag > (T)0) {
          int low = 0;
          for (int j = tid; j < n; j += kGnThreads) low |= O::abs(Hp[(size_t)j * n + j]) < p.opt.check_min_H_diag;
          if (__syncthreads_or(low)) built_ok = false;
        }
      }
    }
    bool solver_failed = true, early_return = false;
    const uint8_t max_tries = p.mode == 0 ? lm_max_tries(p.opt) : 0;
    for (int attempt = 0;; ++attempt) {
      if (p.mode == 0 && !(s.num_consec_failures <= max_tries)) break;
      if (built_ok) {
        double sc = 1.0;
        bool damp = false;
        if (p.mode == 0) damp = lm_damping_scale(s, p.opt, pass_rebuilt, sc);
        else if (p.lambda) { const T lam = p.lambda[pr]; damp = lam > (T)0; sc = 1.0 + (double)lam; }
        for (int j = tid; j < n; j += kGnThreads) {
          const T base = (pass_rebuilt || !hd) ? Hp[(size_t)j * n + j] : hd[j];
          dd[j] = damp ? (T)((double)base * sc) : base;
        }
        __syncthreads();
        if (hd) for (int j = tid; j < n; j += kGnThreads) hd[j] = dd[j];  // H_ keeps the damped diagonal
        if (p.use_ldlt) {
          gn_pivot_order<T>(dd, n, perm, inv, &sh_flag);
          // W = P H P^T, lower, column major: W(a, b) = H(perm[a], perm[b]) from the upper triangle
          for (int64_t e = tid; e < (int64_t)n * n; e += kGnThreads) {
            const int b = (int)(e / n), a = (int)(e % n);
            if (a < b) continue;
            const int ia = perm[a], ib = perm[b];
            W[e] = (a == b) ? dd[ia] : (ia < ib ? Hp[(size_t)ia * n + ib] : Hp[(size_t)ib * n + ia]);
          }
          __syncthreads();
          if (gn_ldlt_factor<T>(W, n, temp, misc)) {
            for (int i = tid; i < n; i += kGnThreads) ys[i] = -gp[perm[i]];  // gn.h:155, P b
            __syncthreads();
            gn_ldlt_solve<T>(W, n, perm, ys, dd);  // dd is dead once W is laid out: dx goes there
            solver_failed = false;
          }
        } else {  // gn.h:157-163: -H^-1 g, never fails (n == 1 keeps the reference's guard)
          if (n == 1) {
            if (tid == 0) dd[0] = dd[0] > O::float_eps() ? O::mul(-O::div((T)1, dd[0]), gp[0]) : (T)0;
            __syncthreads();
          } else {
            for (int64_t e = tid; e < (int64_t)n * n; e += kGnThreads) {
              const int i = (int)(e / n), j = (int)(e % n);
              W[e] = (i == j) ? dd[i] : (i < j ? Hp[e] : Hp[(size_t)j * n + i]);
            }
            for (int i = tid; i < n; i += kGnThreads) ys[i] = -gp[i];
            __syncthreads();
            gn_lu_solve<T>(W, n, ys, temp, misc);
            for (int i = tid; i < n; i += kGnThreads) dd[i] = ys[i];
            __syncthreads();
          }
          solver_failed = false;
        }
      }
      if (!solver_failed) break;
      if (p.mode != 0) break;
      if (tid == 0) sh_act = lm_on_solver_failure(s, p.opt, cost, nres);
      __syncthreads();
      const int act = sh_act;
      if (act == kLmEarlyReturn) early_return = true;
      if (act != kLmRetry) break;
      if (attempt >= 100000) break;
    }
    const T *dxs = dd;
    if (p.mode != 0) {
      if (!solver_failed) for (int j = tid; j < n; j += kGnThreads) p.dx[(size_t)pr * n + j] = dxs[j];
      if (tid == 0) {
        p.status[pr] = solver_failed ? 1 : 0;
        if (p.cost_out) p.cost_out[pr] = (double)p.cost[pr];
      }
      continue;
    }
    if (tid == 0) {  // optimizer.h:412-415: squaredNorm in Scalar, canonical chains
      T dn = (T)0, gn = (T)0;
      if (!solver_failed) {
        for (int j = 0; j < n; ++j) dn = O::fma(dxs[j], dxs[j], dn);
        if (p.opt.min_grad_norm2_f > 0.0f)
          for (int j = 0; j < n; ++j) gn = O::fma(gp[j], gp[j], gn);
      }
      bool success, has_dx;
      lm_finish_step(s, p.opt, early_return, solver_failed, cost, nres, (double)dn, (double)gn, success, has_dx);
      sh_act = lm_update_action(s, p.opt, success, has_dx);
    }
    __syncthreads();
    const int action = sh_act;
    T *xp = p.x + (size_t)pr * n, *lp = p.last_dx + (size_t)pr * n;
    for (int j = tid; j < n; j += kGnThreads) {
      if (action == kLmApplyDx || action == kLmProbeDx) {
        xp[j] = O::add(xp[j], dxs[j]);
        lp[j] = dxs[j];
      } else if (action == kLmRollBack) {
        xp[j] = O::add(xp[j], -lp[j]);
      }
    }
    if (tid == 0) {
      p.rec[pr] = s;
      if (s.done()) { if (p.results) lm_write_result(s, &p.results[pr]); }
      else local_active++;
      if (p.needs) p.needs[pr] = s.done() ? -1 : ((s.rebuild() || p.opt.solver_type != 0) ? 1 : 0);
    }
  }
  (void)sh_norm;
  if (tid == 0 && local_active && p.n_active) atomicAdd(p.n_active, local_active);
}

template <typename T>
__global__ void gn_init_kernel(LmScalars<T> *rec, DevOptions<T> opt, T *last_dx, int64_t B, int n, int32_t *needs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    LmScalars<T> s;
    s.reset_scalars(opt);
    rec[i] = s;
    if (needs) needs[i] = 1;
  }
  if (i < B * n) last_dx[i] = (T)0;
}

// tob200_solver_step_hg_* above n = 55: the caller's grad / H (upper triangle read, docs/API.md:170) become the
// solver-owned grad_ / H_ of the problems whose Step rebuilds; the others keep theirs (stale H_, re-damped)
template <typename T>
__global__ void gn_import_hg_kernel(const T *grad, const T *Hin, const LmScalars<T> *rec, int is_lm, int64_t B, int n, T *g,
                                    T *H) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * n * n) return;
  const int64_t pr = e / ((int64_t)n * n);
  const LmScalars<T> s = rec[pr];
  if (s.done() || (is_lm && !s.rebuild())) return;
  const int ij = (int)(e % ((int64_t)n * n));
  const int i = ij / n, j = ij % n;
  if (i <= j) H[e] = Hin[e];
  if (i == 0) g[(size_t)pr * n + j] = grad[(size_t)pr * n + j];
}

// the Output scalars of every problem, finished or not (tob200_solver_results)
template <typename T>
__global__ void gn_results_kernel(const LmScalars<T> *rec, int64_t B, tob200_result *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) lm_write_result(rec[i], &out[i]);
}

// Output::final_hessian / H_out: upper triangle + persistent (damped) diagonal -> full symmetric
template <typename T, typename OutT>
__global__ void gn_export_h_kernel(const T *H, const T *hd, const LmScalars<T> *rec, const T *lambda, int solver_type, int64_t B,
                                   int n, OutT *out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * n * n) return;
  const int64_t pr = e / ((int64_t)n * n);
  const int ij = (int)(e % ((int64_t)n * n));
  const int i = ij / n, j = ij % n;
  const T *Hp = H + (size_t)pr * n * n;
  T v;
  if (i == j) {
    v = hd ? hd[(size_t)pr * n + i] : Hp[(size_t)i * n + i];
    if (rec) {  // SolverLM::Hessian() (lm.h:157-171)
      const T pl = rec[pr].prev_lambda;
      if (solver_type == 0 && pl > (T)0) v = Ops<T>::div(v, Ops<T>::add((T)1, pl));
    } else if (lambda) {  // build_solve's H_out: damped (lm.h:108-117)
      const T lam = lambda[pr];
      if (lam > (T)0) v = (T)((double)v * (1.0 + (double)lam));
    }
  } else {
    v = i < j ? Hp[(size_t)i * n + j] : Hp[(size_t)j * n + i];
  }
  out[e] = (OutT)v;
}

}  // namespace tob200
