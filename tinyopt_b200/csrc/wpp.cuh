// wpp.cuh — "warp per problem" kernels for mid n (13 <= n <= 55, float): config C4 (n = 50).
//
// One warp owns one problem.  Rows of the residual block (PROBLEM_MAJOR: A[p][i][j], dense) are
// streamed HBM -> shared memory in chunks of 32 rows by TMA bulk copies (one contiguous copy per
// chunk) through a warp-private mbarrier ring.  Per chunk:
//   phase 1 (lane = row): t_i = a_i . x as the canonical left-to-right fma chain, r_i, the scale
//            (1 + 3 alpha t_i^2), and the AUGMENTED row [J_i | r_i] is written to a packed buffer
//            whose pitch keeps 16-byte vector loads aligned and conflict-free;
//   phase 2 (lane = 8x8 register block of the upper triangle of [J|r]^T [J|r]): per row two
//            broadcast vector loads per operand and 64 fmas.  The augmented column gives
//            g = J^T r and the corner gives cost = r^T r for free, and every accumulator still
//            receives its terms in row order i = 0..m-1 — the same sequence as the CPU oracle.
// Then the damped matrix is laid out in shared memory and factorised by the warp with the same
// diagonal-pivoted LDL^T as ldlt_reg.cuh (left-looking, lane = row, sequential dots; column-
// oriented substitutions whose per-element update order equals the oracle's), and the LM state
// machine (lm_state.cuh) runs warp-uniformly.  The persistent copy of H_ that tinyopt keeps for
// cost-only iterations (solvers/lm.h:96-117) is written to a per-warp global scratch only when the
// coming step can actually be followed by one.
//
// Reference path replaced: diff/optimize_autodiff.h:151-157, solvers/lm.h:60-120,
// solvers/gn.h:150-171, math.h:232-240, optimizers/optimizer.h:243-539.
#pragma once

#include <cstdio>

#include "common.cuh"
#include "lm_state.cuh"

namespace tob200 {

// optional per-phase cycle counters of warp 0 of block 0 (build with -DTOB200_WPP_TIMING; printed at kernel end)
#ifdef TOB200_WPP_TIMING
__device__ long long g_wpp_tm[16];
#define WPP_T0() long long wpp_t0__ = clock64()
#define WPP_T(k) do { const long long t__ = clock64(); if (blockIdx.x == 0 && threadIdx.x == 0) g_wpp_tm[k] += t__ - wpp_t0__; wpp_t0__ = t__; } while (0)
#else
#define WPP_T0()
#define WPP_T(k)
#endif

constexpr int kWppRows = 32;      // rows per chunk == lanes
constexpr int kWppThreads = 128;  // 4 independent warps per CTA
constexpr int kWppMaxStages = 4;
#ifndef TOB200_WPP_AREG
#define TOB200_WPP_AREG 0  // measured: C4 13.07 M it/s with the row of A held in registers vs 13.58 M without
#endif
#ifndef TOB200_WPP_FOLD_PIPE
#define TOB200_WPP_FOLD_PIPE 0
#endif
#ifndef TOB200_WPP_LDLT_COLS
#define TOB200_WPP_LDLT_COLS 2
#endif
constexpr int kWppLdltCols = TOB200_WPP_LDLT_COLS;  // columns per step of the warp LDL^T fast path (<= 4; 2 measured best)

// geometry: the (n+1) columns of [J|r] are cut into NB blocks of BLK columns (BLK = 4 for n <= 27,
// 8 above), NP = NB*BLK padded columns, NB*(NB+1)/2 <= 28 upper-triangular blocks = busy lanes
__host__ __device__ constexpr int wpp_blk_for(int n) { return n + 1 <= 28 ? 4 : 8; }
__host__ __device__ constexpr int wpp_nb_for(int n) { return (n + 1 + wpp_blk_for(n) - 1) / wpp_blk_for(n); }
// pitch of the LDLT matrix: multiple of 4 floats with pitch/4 odd, so that a lane can read its own row
// with aligned, conflict-free 16-byte loads in the dot products
__host__ __device__ constexpr int wpp_ldw(int np) { return ((np / 4) & 1) ? np : np + 4; }
// pitch of the packed [J|r] rows: multiple of 4 floats (16-byte loads) with pitch/4 odd
// (conflict-free lane-per-row stores)
__host__ __device__ constexpr int wpp_nps(int np) { return ((np / 4) & 1) ? np : np + 4; }

// position of column c of [J|r] inside a packed row.  With 8-column blocks a lane reads its block row and its
// block column with 16-byte loads at float offsets 8 bi / 8 bj: blocks b and b + 4 would sit 32 banks apart
// (a 2-way conflict on every operand load of the quarter-warps that hold both: 16.7 % of all shared-memory
// wavefronts of the C4 kernel, which is bound by exactly those wavefronts).  A 4-float gap after column 31
// moves blocks 4..6 by 4 banks: every operand load is conflict-free.  (It is the pad that used to sit at the
// end of the row, so the pitch does not change.)
template <int BLK>
__host__ __device__ constexpr int wpp_col(int c) { return BLK == 8 ? c + 4 * (c >> 5) : c; }

struct WppSmem {  // byte offsets inside one warp's shared memory
  uint32_t bars, xs, last_dx, g, dxs, temp, dg, dd, perm, inv, tx, stages, stage_bytes, jbuf, total;
};
// es = sizeof(scalar): 4 (float) or 8 (double)
__host__ __device__ inline WppSmem wpp_smem_layout(int n, int np_, int stages, uint32_t es = 4) {
  WppSmem L;
  const uint32_t np = (uint32_t)np_;
  uint32_t o = 0;
  L.bars = o; o += 64;
  L.xs = o; o += np * es;
  L.last_dx = o; o += np * es;
  L.g = o; o += np * es;
  L.temp = o; o += np * es;
  L.dg = o; o += np * es;
  L.perm = o; o += np * es;
  // dxs, dd, inv and tx are dead while the factorisation runs: together they are its kWppLdltCols x np
  // buffer of D_j L_{k+c,j} rows
  L.dxs = o; o += np * es;
  L.dd = o; o += np * es;
  L.inv = o; o += np * es;
  L.tx = o; o += np * es;
  o = (o + 127u) & ~127u;
  L.stages = o;
  L.stage_bytes = ((uint32_t)kWppRows * (uint32_t)(n + 1) * es + 127u) & ~127u;  // A rows, then y
  o += (uint32_t)stages * L.stage_bytes;
  L.jbuf = o; o += (uint32_t)kWppRows * (uint32_t)wpp_nps(np_) * es;
  // the LDLT working matrix aliases [stages .. jbuf end): make sure it fits
  const uint32_t wbytes = np * (uint32_t)wpp_ldw(np_) * es;
  if (o - L.stages < wbytes) o = L.stages + wbytes;
  L.total = (o + 127u) & ~127u;
  return L;
}

template <typename T>
struct WppData {
  const T *A;  // [B][m][n] dense (J for build_solve)
  const T *y;  // [B][m]    (r for build_solve)
  int64_t B;
  int m, n;
  int stages;
  int use_tma;  // 1: every chunk is 16-byte aligned; 0: cooperative loads
  WppSmem L;
  unsigned long long *counter;  // dynamic problem queue, zeroed before the launch
  T *hpersist;                  // [warps in grid][NP * LDW] persistent damped H_
};

// ---- warp-private chunk loader ---------------------------------------------------------------------
template <typename T>
struct WppPipe {
  uint64_t *bars;
  unsigned char *stages;
  uint32_t stage_bytes, nstages, stage, phase;
  int use_tma;

  __device__ __forceinline__ void init(unsigned char *ws, const WppData<T> &d, int lane) {
    bars = reinterpret_cast<uint64_t *>(ws + d.L.bars);
    stages = ws + d.L.stages;
    stage_bytes = d.L.stage_bytes;
    nstages = (uint32_t)d.stages;
    stage = 0;
    phase = 0;
    use_tma = d.use_tma;
    if (lane == 0) {
      for (uint32_t s = 0; s < nstages; ++s) mbar_init(&bars[s], 1);
      mbar_fence_init();
    }
    __syncwarp();
  }
  __device__ __forceinline__ T *stage_ptr(uint32_t st) const { return reinterpret_cast<T *>(stages + (size_t)st * stage_bytes); }
  __device__ __forceinline__ void advance() {
    if (++stage == nstages) {
      stage = 0;
      phase ^= 1u;
    }
  }
  // all lanes call; rows [row0, row0 + nrows) of the problem whose rows start at ap / yp
  __device__ __forceinline__ void issue(const T *ap, const T *yp, int n, int row0, int nrows, uint32_t st, int lane) {
    T *sa = stage_ptr(st);
    T *sy = sa + (size_t)kWppRows * n;
    // (y is NOT staged: the 32 values of a chunk are one coalesced load straight into the lanes' registers,
    //  a chunk ahead of their use — one bulk operation per chunk instead of two, and no alignment
    //  requirement on m; measured performance-neutral on C4)
    (void)sy; (void)yp;
    if (use_tma) {
      if (lane == 0) {
        const uint32_t ba = (uint32_t)nrows * (uint32_t)n * (uint32_t)sizeof(T);
        mbar_expect_tx(&bars[st], ba);
        tma_bulk_g2s(sa, ap + (size_t)row0 * n, ba, &bars[st]);
      }
    } else {  // unaligned shapes: plain coalesced loads, completed before anyone reads
      const T *ga = ap + (size_t)row0 * n;
      for (int e = lane; e < nrows * n; e += 32) sa[e] = ga[e];
    }
  }
  __device__ __forceinline__ void wait() {
    if (use_tma) mbar_wait(&bars[stage], phase);
    else __syncwarp();
  }
};

// ---- warp-cooperative pivoted LDL^T, same semantics and the same per-element operation order as
// ---- LdltReg / the CPU oracle -----------------------------------------------------------------------
// Eigen's unblocked LDLT is left-looking: step k only finishes column k, so the trailing diagonal it
// searches for the next pivot is still the ORIGINAL (permuted) diagonal.  The whole pivot sequence is
// therefore a function of the diagonal alone: it is computed first (in registers, one REDUX + two
// ballots per step), the matrix is laid out already permuted, and the factorisation itself runs
// without any search or swap.  P A P^T = L D L^T with the same P, the same L, the same bits.

// pos2orig of the pivot order for the diagonal dd[0..n) (n <= 64); writes perm[pos] = orig and
// inv[orig] = pos.  "First maximum wins" and NaN handling as in Eigen's maxCoeff visitor.
// Eigen's pivot search replayed literally on the diagonal by lane 0 (any scalar type): at step k the FIRST
// largest |d| among positions k..n-1 is swapped to k (maxCoeff visitor: strict >, so a NaN never wins
// and a NaN sitting at k stays).  The slow path of the double instantiation and of cov_kernels.cu.
template <typename T>
__device__ void wpp_pivot_order_serial(const T *dd, int n, int *perm, int *inv, int lane) {
  using O = Ops<T>;
  if (lane == 0) {
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
      int p = k;
      T best = O::abs(dd[perm[k]]);
      for (int i = k + 1; i < n; ++i) {
        const T v = O::abs(dd[perm[i]]);
        if (v > best) { best = v; p = i; }
      }
      const int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
    }
    for (int k = 0; k < n; ++k) inv[perm[k]] = k;
  }
  __syncwarp();
}

// double: the same fast path (position = number of larger |d|, valid when all |d| are distinct and
// none is NaN — a NaN equals nothing, not even itself, so its own count gives it away), else the replay
__device__ __forceinline__ void wpp_pivot_order(const double *dd, int n, int *perm, int *inv, int lane) {
  const int i0 = lane, i1 = lane + 32;
  const double a0 = i0 < n ? fabs(dd[i0]) : 0.0, a1 = i1 < n ? fabs(dd[i1]) : 0.0;
  int gt0 = 0, gt1 = 0, eq0 = 0, eq1 = 0;
#pragma unroll 4
  for (int j = 0; j < n; ++j) {
    const double v = fabs(dd[j]);
    gt0 += v > a0; eq0 += v == a0;
    gt1 += v > a1; eq1 += v == a1;
  }
  const bool amb = (i0 < n && eq0 != 1) || (i1 < n && eq1 != 1);
  if (!__any_sync(0xffffffffu, amb)) {
    if (i0 < n) { perm[gt0] = i0; inv[i0] = gt0; }
    if (i1 < n) { perm[gt1] = i1; inv[i1] = gt1; }
    __syncwarp();
    return;
  }
  wpp_pivot_order_serial<double>(dd, n, perm, inv, lane);
}

__device__ __forceinline__ void wpp_pivot_order(const float *dd, int n, int *perm, int *inv, int lane) {
  const int i0 = lane, i1 = lane + 32;
  uint32_t k0 = 0, k1 = 0;  // 0: NaN (never wins), otherwise 1 + bits(|d|) (monotone in |d|)
  if (i0 < n) { const float v = fabsf(dd[i0]); k0 = (v != v) ? 0u : __float_as_uint(v) + 1u; }
  if (i1 < n) { const float v = fabsf(dd[i1]); k1 = (v != v) ? 0u : __float_as_uint(v) + 1u; }
  // Fast path: with distinct keys (no ties, no NaN) "the first largest of the rest goes to k" is the
  // descending sort, so position = rank = number of larger keys (dd is in shared memory: broadcast
  // loads).  Ties make the winner depend on where earlier swaps moved the candidates: replay below.
  {
    int gt0 = 0, gt1 = 0, eq0 = 0, eq1 = 0;
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      const float v = fabsf(dd[j]);
      const uint32_t kj = (v != v) ? 0u : __float_as_uint(v) + 1u;
      gt0 += kj > k0; eq0 += kj == k0;
      gt1 += kj > k1; eq1 += kj == k1;
    }
    const bool amb = (i0 < n && (eq0 != 1 || k0 == 0u)) || (i1 < n && (eq1 != 1 || k1 == 0u));
    if (!__any_sync(0xffffffffu, amb)) {
      if (i0 < n) { perm[gt0] = i0; inv[i0] = gt0; }
      if (i1 < n) { perm[gt1] = i1; inv[i1] = gt1; }
      __syncwarp();
      return;
    }
  }
  int o0 = i0, o1 = i1;
  for (int k = 0; k < n; ++k) {
    const uint32_t kk = __shfl_sync(0xffffffffu, k < 32 ? k0 : k1, k & 31);
    int p = k;
    if (kk != 0u) {  // a NaN sitting at k stays (nothing compares greater than it)
      const bool e0 = i0 >= k && i0 < n, e1 = i1 >= k && i1 < n;
      const uint32_t c0 = e0 ? k0 : 0u, c1 = e1 ? k1 : 0u;
      const uint32_t m = __reduce_max_sync(0xffffffffu, c0 > c1 ? c0 : c1);
      const uint32_t b0 = __ballot_sync(0xffffffffu, e0 && k0 == m);
      const uint32_t b1 = __ballot_sync(0xffffffffu, e1 && k1 == m);
      p = b0 ? (__ffs(b0) - 1) : (32 + __ffs(b1) - 1);
    }
    if (p != k) {  // swap the contents of positions k and p (warp-uniform branch)
      const int ok = __shfl_sync(0xffffffffu, k < 32 ? o0 : o1, k & 31);
      const int op = __shfl_sync(0xffffffffu, p < 32 ? o0 : o1, p & 31);
      const uint32_t kp = __shfl_sync(0xffffffffu, p < 32 ? k0 : k1, p & 31);
      if (i0 == k) { k0 = kp; o0 = op; }
      if (i1 == k) { k1 = kp; o1 = op; }
      if (i0 == p) { k0 = kk; o0 = ok; }
      if (i1 == p) { k1 = kk; o1 = ok; }
    }
  }
  if (i0 < n) { perm[i0] = o0; inv[o0] = i0; }
  if (i1 < n) { perm[i1] = o1; inv[o1] = i1; }
  __syncwarp();
}

// unpivoted left-looking LDL^T of the already permuted W (lower triangle, pitch ldw), lane = row, one
// column per step, scalar loads: any scalar type.  Same bookkeeping (sign, zero pivots) and the same
// per-element operation order as the float fast path below / the oracle's ldlt_factor_.
template <typename T>
__device__ bool wpp_ldlt_factor_generic(T *W, int ldw, int n, T *temp, int lane) {
  using O = Ops<T>;
#define WW(i, j) W[(i) * ldw + (j)]
  if (n == 1) return !(WW(0, 0) < (T)0);
  int sign = 0;
  bool found_zero_pivot = false, ret = true;
  for (int k = 0; k < n; ++k) {
    if (k > 0) {
      for (int j = lane; j < k; j += 32) temp[j] = O::mul(WW(j, j), WW(k, j));
      __syncwarp();
      for (int i = k + lane; i < n; i += 32) {  // row k itself: A_kk -= A10 . temp
        T s = (T)0;
        for (int j = 0; j < k; ++j) s = O::fma(WW(i, j), temp[j], s);
        WW(i, k) = O::sub(WW(i, k), s);
      }
      __syncwarp();
    }
    const T akk = WW(k, k);
    const bool pivot_is_valid = O::abs(akk) > (T)0;
    if (k == 0 && !pivot_is_valid) {  // the whole diagonal is zero
      bool z = true;
      for (int j = 0; j < n; ++j)
        for (int i = j + 1 + lane; i < n; i += 32) z = z && (WW(i, j) == (T)0);
      return __all_sync(0xffffffffu, z);
    }
    if (k < n - 1) {
      if (pivot_is_valid) {
        for (int i = k + 1 + lane; i < n; i += 32) WW(i, k) = O::div(WW(i, k), akk);
      } else {
        bool z = true;
        for (int i = k + 1 + lane; i < n; i += 32) z = z && (WW(i, k) == (T)0);
        ret = ret && __all_sync(0xffffffffu, z);
      }
      __syncwarp();
    }
    if (found_zero_pivot && pivot_is_valid) ret = false;
    else if (!pivot_is_valid) found_zero_pivot = true;
    if (sign == 1) { if (akk < (T)0) sign = 2; }
    else if (sign == -1) { if (akk > (T)0) sign = 2; }
    else if (sign == 0) { if (akk > (T)0) sign = 1; else if (akk < (T)0) sign = -1; }
  }
  return ret && (sign == 1 || sign == 0);
#undef WW
}

// unpivoted left-looking LDL^T of the already permuted W (lower triangle, pitch ldw).
//
// Fast path: kWppLdltCols columns per step while the pivots are valid.  The dot product of column
// k + c is sum_{j<k} L_ij (D_j L_{k+c,j}) followed by the terms j = k .. k + c - 1, and its first k
// terms do not depend on the columns k .. k + c - 1 — so the columns of a step share one sweep over the
// rows of L (one load of L_ij feeds every column's fma chain: with two columns 4 loads per 16 fmas instead
// of 6 and four independent chains per lane instead of two; 3 and 4 columns measured slower), the pivots travel by shuffle instead of through shared
// memory, and the remaining terms are appended as the columns before become final.  Every chain
// still receives its terms in the order j = 0, 1, .., so the factor is bit-identical to the one-column
// loop below, which remains as the path for zero / NaN pivots and for the last n % kWppLdltCols columns.
template <typename T>
__device__ __forceinline__ bool wpp_ldlt_factor(T *W, int ldw, int n, T *temp, T *tbuf, int tp, int lane) {
  using O = Ops<T>;
  if constexpr (sizeof(T) != 4) {  // the vector loads below are float specific
    return wpp_ldlt_factor_generic<T>(W, ldw, n, temp, lane);
  } else {
#define WW(i, j) W[(i) * ldw + (j)]
  if (n == 1) return !(WW(0, 0) < (T)0);
  int sign = 0;
  bool found_zero_pivot = false, ret = true;
  int k = 0;
  constexpr int C = kWppLdltCols;
  bool bail = false;
  for (; k + C <= n && !bail; ) {
    // T_c[j] = D_j L_{k+c,j}, j < k, c < C (tbuf rows of pitch tp)
    for (int j = lane; j < k; j += 32) {
      const T d = WW(j, j);
#pragma unroll
      for (int c = 0; c < C; ++c) tbuf[c * tp + j] = O::mul(d, WW(k + c, j));
    }
    __syncwarp();
    const int r0 = k + lane, r1 = k + lane + 32;  // lane c owns row k + c
    const bool h0 = r0 < n, h1 = r1 < n;
    const T *w0 = &WW(h0 ? r0 : k, 0);
    const T *w1 = &WW(h1 ? r1 : k, 0);
    T s0[C], s1[C];  // chains of columns k .. k + C - 1 for rows r0, r1
#pragma unroll
    for (int c = 0; c < C; ++c) { s0[c] = (T)0; s1[c] = (T)0; }
    int j = 0;
#pragma unroll 2
    for (; j + 4 <= k; j += 4) {
      const float4 a4 = *reinterpret_cast<const float4 *>(w0 + j);
      const float4 b4 = *reinterpret_cast<const float4 *>(w1 + j);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float4 t4 = *reinterpret_cast<const float4 *>(tbuf + c * tp + j);
        s0[c] = O::fma(a4.x, t4.x, s0[c]); s1[c] = O::fma(b4.x, t4.x, s1[c]);
        s0[c] = O::fma(a4.y, t4.y, s0[c]); s1[c] = O::fma(b4.y, t4.y, s1[c]);
        s0[c] = O::fma(a4.z, t4.z, s0[c]); s1[c] = O::fma(b4.z, t4.z, s1[c]);
        s0[c] = O::fma(a4.w, t4.w, s0[c]); s1[c] = O::fma(b4.w, t4.w, s1[c]);
      }
    }
    for (; j < k; ++j) {
      const T a = w0[j], b = w1[j];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const T t = tbuf[c * tp + j];
        s0[c] = O::fma(a, t, s0[c]);
        s1[c] = O::fma(b, t, s1[c]);
      }
    }
    // finish the C columns one after the other; column c first takes the terms j = k .. k + c - 1
    T l0[C], l1[C];  // the final L values of my rows in columns k .. k + C - 1 (lane c: D_{k+c} in l0[c])
    T dk[C];
    int done = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (!bail) {
#pragma unroll
        for (int cc = 0; cc < c; ++cc) {
          const T tk = O::mul(dk[cc], __shfl_sync(0xffffffffu, l0[cc], c));  // D_{k+cc} L_{k+c,k+cc}
          if (lane > cc) s0[c] = O::fma(l0[cc], tk, s0[c]);
          s1[c] = O::fma(l1[cc], tk, s1[c]);
        }
        const T e0 = O::sub(h0 ? w0[k + c] : (T)0, s0[c]), e1 = O::sub(h1 ? w1[k + c] : (T)0, s1[c]);
        const T piv = __shfl_sync(0xffffffffu, e0, c);
        if (!(O::abs(piv) > (T)0)) {
          bail = true;  // columns k .. k + c - 1 are final and stored; the rest goes through the loop below
        } else {
          if (sign == 1) { if (piv < (T)0) sign = 2; }
          else if (sign == -1) { if (piv > (T)0) sign = 2; }
          else if (sign == 0) { if (piv > (T)0) sign = 1; else sign = -1; }
          dk[c] = piv;
          l0[c] = lane == c ? piv : O::div(e0, piv);
          l1[c] = O::div(e1, piv);
          if (lane >= c && h0) WW(r0, k + c) = l0[c];
          if (h1) WW(r1, k + c) = l1[c];
          done = c + 1;
        }
      }
    }
    __syncwarp();
    k += done;
  }
  for (; k < n; ++k) {
    if (k > 0) {
      for (int j = lane; j < k; j += 32) temp[j] = O::mul(WW(j, j), WW(k, j));
      __syncwarp();
      // rows k+lane and k+lane+32 together (row k itself gives A_kk -= A10 . temp); rows and temp
      // are 16-byte aligned, so four terms per load, consumed in order
      const int r0 = k + lane, r1 = k + lane + 32;
      if (r0 < n) {
        const T *w0 = &WW(r0, 0);
        const bool two = r1 < n;
        const T *w1 = two ? &WW(r1, 0) : w0;
        T s0 = (T)0, s1 = (T)0;
        int j = 0;
#pragma unroll 2
        for (; j + 4 <= k; j += 4) {
          const float4 t4 = *reinterpret_cast<const float4 *>(temp + j);
          const float4 a4 = *reinterpret_cast<const float4 *>(w0 + j);
          const float4 b4 = *reinterpret_cast<const float4 *>(w1 + j);
          s0 = O::fma(a4.x, t4.x, s0); s1 = O::fma(b4.x, t4.x, s1);
          s0 = O::fma(a4.y, t4.y, s0); s1 = O::fma(b4.y, t4.y, s1);
          s0 = O::fma(a4.z, t4.z, s0); s1 = O::fma(b4.z, t4.z, s1);
          s0 = O::fma(a4.w, t4.w, s0); s1 = O::fma(b4.w, t4.w, s1);
        }
        for (; j < k; ++j) {
          const T t = temp[j];
          s0 = O::fma(w0[j], t, s0);
          s1 = O::fma(w1[j], t, s1);
        }
        WW(r0, k) = O::sub(WW(r0, k), s0);
        if (two) WW(r1, k) = O::sub(WW(r1, k), s1);
      }
      __syncwarp();
    }
    const T akk = WW(k, k);
    const bool pivot_is_valid = O::abs(akk) > (T)0;
    if (k == 0 && !pivot_is_valid) {  // the whole diagonal is zero (the pivot order is the identity)
      bool z = true;
      for (int j = 0; j < n; ++j)
        for (int i = j + 1 + lane; i < n; i += 32) z = z && (WW(i, j) == (T)0);
      return __all_sync(0xffffffffu, z);
    }
    if (k < n - 1) {
      if (pivot_is_valid) {
        for (int i = k + 1 + lane; i < n; i += 32) WW(i, k) = O::div(WW(i, k), akk);
      } else {
        bool z = true;
        for (int i = k + 1 + lane; i < n; i += 32) z = z && (WW(i, k) == (T)0);
        ret = ret && __all_sync(0xffffffffu, z);
      }
      __syncwarp();
    }
    if (found_zero_pivot && pivot_is_valid) ret = false;
    else if (!pivot_is_valid) found_zero_pivot = true;
    if (sign == 1) { if (akk < (T)0) sign = 2; }
    else if (sign == -1) { if (akk > (T)0) sign = 2; }
    else if (sign == 0) { if (akk > (T)0) sign = 1; else if (akk < (T)0) sign = -1; }
  }
  return ret && (sign == 1 || sign == 0);
#undef WW
  }
}

// x (shared, n values, original order) <- P^T L^-T D^+ L^-1 P b;  n <= 64
template <typename T>
__device__ __forceinline__ void wpp_ldlt_solve(const T *W, int ldw, int n, const int *perm, const T *b, T *x, int lane) {
  using O = Ops<T>;
#define WW(i, j) W[(i) * ldw + (j)]
  T y0 = lane < n ? b[perm[lane]] : (T)0, y1 = lane + 32 < n ? b[perm[lane + 32]] : (T)0;
  // L y = y, column oriented: y_i takes its updates in the order j = 0 .. i-1
  for (int j = 0; j < n; ++j) {
    const T yj = __shfl_sync(0xffffffffu, j < 32 ? y0 : y1, j & 31);
    if (lane > j && lane < n) y0 = O::fma(-WW(lane, j), yj, y0);
    if (lane + 32 > j && lane + 32 < n) y1 = O::fma(-WW(lane + 32, j), yj, y1);
  }
  if (lane < n) { const T d = WW(lane, lane); y0 = (O::abs(d) > O::min_normal()) ? O::div(y0, d) : (T)0; }
  if (lane + 32 < n) { const T d = WW(lane + 32, lane + 32); y1 = (O::abs(d) > O::min_normal()) ? O::div(y1, d) : (T)0; }
  // L^T y = y, column oriented from the last column: y_i takes its updates in the order j = n-1 .. i+1
  for (int j = n - 1; j >= 0; --j) {
    const T yj = __shfl_sync(0xffffffffu, j < 32 ? y0 : y1, j & 31);
    if (lane < j) y0 = O::fma(-WW(j, lane), yj, y0);
    if (lane + 32 < j) y1 = O::fma(-WW(j, lane + 32), yj, y1);
  }
  __syncwarp();
  if (lane < n) x[perm[lane]] = y0;
  if (lane + 32 < n) x[perm[lane + 32]] = y1;
  __syncwarp();
#undef WW
}

// `hessian.use_ldlt = false` (include/tinyopt/solvers/gn.h:157-163): x <- -H^-1 g by partial-pivot LU
// in the oracle's canonical order (see lu_solve_local in ldlt_reg.cuh), warp-cooperative.  W holds
// the lower triangle of the unpermuted damped H_ on entry and is destroyed; b is n scratch values.
// Lanes own columns in the elimination (the rhs is column n) and rows in the upper solve; n <= 64.
// Off the hot path, so a real call: the LDLT path keeps its registers.
template <typename T>
__device__ __noinline__ void wpp_lu_solve(T *W, int ldw, int n, const T *g, T *b, T *x, int lane) {
  using O = Ops<T>;
  constexpr unsigned kFull = 0xffffffffu;
#define WW(i, j) W[(i) * ldw + (j)]
  if (n == 1) {  // gn.h:158-160
    if (lane == 0) x[0] = WW(0, 0) > O::float_eps() ? O::mul(-O::div((T)1, WW(0, 0)), g[0]) : (T)0;
    __syncwarp();
    return;
  }
  for (int e = lane; e < n * n; e += 32) {
    const int i = e / n, j = e - i * n;
    if (j > i) WW(i, j) = WW(j, i);
  }
  for (int j = lane; j < n; j += 32) b[j] = -g[j];
  __syncwarp();
  for (int k = 0; k < n; ++k) {
    // first largest |entry| of column k at or below the diagonal (a NaN diagonal keeps p = k)
    T best = O::abs(WW(k, k));
    int p = k;
    for (int i = k + 1 + lane; i < n; i += 32) {
      const T v = O::abs(WW(i, k));
      if (v > best) {
        best = v;
        p = i;
      }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      const T ob = __shfl_xor_sync(kFull, best, off);
      const int op = __shfl_xor_sync(kFull, p, off);
      if (ob > best || (ob == best && op < p)) {
        best = ob;
        p = op;
      }
    }
    if (p != k) {
      for (int j = lane; j < n; j += 32) {
        const T t = WW(k, j);
        WW(k, j) = WW(p, j);
        WW(p, j) = t;
      }
      if (lane == 0) {
        const T t = b[k];
        b[k] = b[p];
        b[p] = t;
      }
      __syncwarp();
    }
    const T piv = WW(k, k);
    for (int j = k + 1 + lane; j <= n; j += 32) {
      const bool rhs = j == n;
      const T akj = rhs ? b[k] : WW(k, j);
      for (int i = k + 1; i < n; ++i) {
        const T f = O::div(WW(i, k), piv);
        if (rhs) b[i] = O::fma(-f, akj, b[i]);
        else WW(i, j) = O::fma(-f, akj, WW(i, j));
      }
    }
    __syncwarp();
  }
  // U x = b from the last row, x_i taking its updates in the order j = n-1 .. i+1
  T s0 = lane < n ? b[lane] : (T)0, s1 = lane + 32 < n ? b[lane + 32] : (T)0;
  for (int j = n - 1; j >= 0; --j) {
    T xj = j < 32 ? s0 : s1;
    if (lane == (j & 31)) xj = O::div(xj, WW(j, j));
    xj = __shfl_sync(kFull, xj, j & 31);
    if (lane == (j & 31)) {
      if (j < 32) s0 = xj;
      else s1 = xj;
    }
    if (lane < j) s0 = O::fma(-WW(lane, j), xj, s0);
    if (lane + 32 < j) s1 = O::fma(-WW(lane + 32, j), xj, s1);
  }
  if (lane < n) x[lane] = s0;
  if (lane + 32 < n) x[lane + 32] = s1;
  __syncwarp();
#undef WW
}

// sequential (canonical) sum of squares of a shared vector, computed by lane 0, broadcast
template <typename T>
__device__ __forceinline__ T wpp_sqnorm(const T *v, int n, int lane) {
  T s = (T)0;
  if (lane == 0) {
#pragma unroll 8
    for (int j = 0; j < n; ++j) s = Ops<T>::fma(v[j], v[j], s);
  }
  return __shfl_sync(0xffffffffu, s, 0);
}

// ---- per-warp register block bookkeeping ---------------------------------------------------------------
template <int NB>
__device__ __forceinline__ void wpp_block_of_lane(int lane, int &bi, int &bj, bool &has_block) {
  // upper-triangular blocks enumerated row by row: (0,0) (0,1) .. (0,NB-1) (1,1) ..
  bi = 0;
  bj = 0;
  has_block = false;
  int rem = lane;
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    const int len = NB - r;
    if (!has_block) {
      if (rem < len) {
        bi = r;
        bj = r + rem;
        has_block = true;
      } else {
        rem -= len;
      }
    }
  }
}

// One streaming pass over the rows of a problem: accumulates the register block of
// [J|r]^T [J|r] (do_rebuild) or only the cost (cost-only pass: returns it in cost_only).
template <typename T, int NB, int BLK, bool kSynth>
__device__ __forceinline__ void wpp_pass(WppPipe<T> &pipe, const WppData<T> &d, unsigned char *ws, int64_t p, int lane,
                                         bool do_rebuild, T alpha, T alpha3, int bi, int bj, bool has_block,
                                         T (&acc)[BLK][BLK], T &cost_only) {
  using O = Ops<T>;
  constexpr int NP = NB * BLK, NPS = wpp_nps(NP);
  const int m = d.m, n = d.n;
  const T *xs = reinterpret_cast<const T *>(ws + d.L.xs);
  T *jbuf = reinterpret_cast<T *>(ws + d.L.jbuf);
  // accumulators as packed FP32 pairs (acc[u][2 h], acc[u][2 h + 1]): FFMA2 (fma.rn.f32x2) does two
  // IEEE fused multiply-adds per issue slot, same roundings as two scalar fmas
  constexpr bool kF32 = sizeof(T) == 4;  // the vector / FFMA2 paths are float specific; double takes scalar ones
  unsigned long long acc2[BLK][BLK / 2];
#pragma unroll
  for (int u = 0; u < BLK; ++u)
#pragma unroll
    for (int h = 0; h < BLK / 2; ++h) acc2[u][h] = 0ull;
  if constexpr (!kF32) {
#pragma unroll
    for (int u = 0; u < BLK; ++u)
#pragma unroll
      for (int v = 0; v < BLK; ++v) acc[u][v] = (T)0;
  }
  cost_only = (T)0;

  const int nchunks = (m + kWppRows - 1) / kWppRows;
  const T *ap = d.A + (size_t)p * m * n;
  const T *yp = d.y + (size_t)p * m;
  {
    if (lane == 0) fence_proxy_async();
    const int pre = nchunks < (int)pipe.nstages ? nchunks : (int)pipe.nstages;
    uint32_t st = pipe.stage;
    for (int c = 0; c < pre; ++c) {
      const int row0 = c * kWppRows;
      pipe.issue(ap, yp, n, row0, (m - row0 < kWppRows) ? (m - row0) : kWppRows, st, lane);
      if (++st == pipe.nstages) st = 0;
    }
  }
  WPP_T0();
  T ycur = lane < m ? yp[lane] : (T)0;  // y (or r) of my row of chunk 0
  for (int c = 0; c < nchunks; ++c) {
    const int row0 = c * kWppRows;
    const int nrows = (m - row0 < kWppRows) ? (m - row0) : kWppRows;
    const T ynext = (row0 + kWppRows + lane < m) ? yp[row0 + kWppRows + lane] : (T)0;  // in flight during this chunk
    pipe.wait();
    WPP_T(0);
    const T *sa = pipe.stage_ptr(pipe.stage);
    // ---- phase 1 (lane = row): canonical t-chain, residual r_i, Jacobian scale sc_i, then the
    // ---- augmented packed row [sc_i a_i | r_i | 0 ..] written with 16-byte stores (pitch / 4 odd:
    // ---- conflict free); the pad columns are rewritten for every row, so they never go stale ----
#if TOB200_WPP_AREG
    // float, even n: the row of A is read from shared memory ONCE, into registers, and serves both the
    // t-chain and the scaled copy (the kernel is bound by shared-memory wavefronts: this saves n / 2 8-byte
    // loads per row).  Columns below NP - BLK are always Jacobian columns (n + 1 > NP - BLK), so only the last
    // block needs run-time masks.
    bool row_done = false;
    if constexpr (kF32) {
      if ((n & 1) == 0) {
        row_done = true;
        if (lane < nrows) {
          constexpr int HQ = NP / 2, HS = (NP - BLK) / 2;
          const int h = n / 2;
          const float2 *a2 = reinterpret_cast<const float2 *>(sa + lane * n);
          const float2 *x2 = reinterpret_cast<const float2 *>(xs);
          float2 ar[HQ];
#pragma unroll
          for (int q = 0; q < HQ; ++q) ar[q] = (q < HS || q < h) ? a2[q] : make_float2(0.f, 0.f);
          T ri, sc = (T)1;
          if (kSynth) {
            T t = (T)0;
#pragma unroll
            for (int q = 0; q < HQ; ++q) {
              if (q < HS || q < h) {
                const float2 xv = x2[q];
                t = O::fma(ar[q].x, xv.x, t);
                t = O::fma(ar[q].y, xv.y, t);
              }
            }
            const T t2 = O::mul(t, t);
            ri = O::fma(t, O::fma(alpha, t2, (T)1), -ycur);
            sc = O::fma(alpha3, t2, (T)1);
          } else {
            ri = ycur;
          }
          if (do_rebuild) {
            float4 *jrow = reinterpret_cast<float4 *>(jbuf + lane * NPS);
#pragma unroll
            for (int q4 = 0; q4 < NP / 4; ++q4) {
              float2 lo = ar[2 * q4], hi = ar[2 * q4 + 1];
              const int q0 = 2 * q4, q1 = 2 * q4 + 1;
              if (kSynth) {  // (pairs at or beyond h hold zeros: scaling them is harmless and branch free)
                lo.x = O::mul(sc, lo.x); lo.y = O::mul(sc, lo.y);
                hi.x = O::mul(sc, hi.x); hi.y = O::mul(sc, hi.y);
              }
              if (q0 >= HS && q0 == h) lo = make_float2(ri, 0.f);
              if (q1 >= HS && q1 == h) hi = make_float2(ri, 0.f);
              jrow[wpp_col<BLK>(4 * q4) / 4] = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
          } else {
            jbuf[lane * NPS + wpp_col<BLK>(n)] = ri;
          }
        }
      }
    }
    if (!row_done)
#endif
    if (lane < nrows) {
      const T *arow = sa + lane * n;
      T ri, sc = (T)1;
      if (kSynth) {
        T t = (T)0;
        bool done = false;
        if constexpr (kF32) {
          if ((n & 1) == 0) {  // even n: rows are 8-byte aligned, two columns per load
            const float2 *a2 = reinterpret_cast<const float2 *>(arow);
            const float2 *x2 = reinterpret_cast<const float2 *>(xs);
#pragma unroll 8
            for (int j = 0; j < n / 2; ++j) {
              const float2 av = a2[j], xv = x2[j];
              t = O::fma(av.x, xv.x, t);
              t = O::fma(av.y, xv.y, t);
            }
            done = true;
          }
        }
        if (!done) {
#pragma unroll 8
          for (int j = 0; j < n; ++j) t = O::fma(arow[j], xs[j], t);
        }
        const T t2 = O::mul(t, t);
        ri = O::fma(t, O::fma(alpha, t2, (T)1), -ycur);
        sc = O::fma(alpha3, t2, (T)1);
      } else {
        ri = ycur;
      }
      if (do_rebuild) {
       if constexpr (!kF32) {
        T *jrow = jbuf + lane * NPS;
#pragma unroll 4
        for (int j = 0; j < NP; ++j) {
          T v = (T)0;
          if (j < n) v = kSynth ? O::mul(sc, arow[j]) : arow[j];
          else if (j == n) v = ri;
          jrow[wpp_col<BLK>(j)] = v;
        }
       } else {
        float4 *jrow = reinterpret_cast<float4 *>(jbuf + lane * NPS);
        if ((n & 1) == 0) {
          const float2 *a2 = reinterpret_cast<const float2 *>(arow);
          const int h = n / 2;  // pairs (2 q, 2 q + 1) hold J for q < h, (r_i, 0) for q == h, zeros after
#pragma unroll 7
          for (int q4 = 0; q4 < NP / 4; ++q4) {
            float2 lo = make_float2((T)0, (T)0), hi = lo;
            const int q0 = 2 * q4, q1 = 2 * q4 + 1;
            if (q0 < h) { lo = a2[q0]; if (kSynth) { lo.x = O::mul(sc, lo.x); lo.y = O::mul(sc, lo.y); } }
            else if (q0 == h) lo.x = ri;
            if (q1 < h) { hi = a2[q1]; if (kSynth) { hi.x = O::mul(sc, hi.x); hi.y = O::mul(sc, hi.y); } }
            else if (q1 == h) hi.x = ri;
            jrow[wpp_col<BLK>(4 * q4) / 4] = make_float4(lo.x, lo.y, hi.x, hi.y);
          }
        } else {
#pragma unroll 7
          for (int q4 = 0; q4 < NP / 4; ++q4) {
            T v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 4 * q4 + e;
              v[e] = (T)0;
              if (j < n) v[e] = kSynth ? O::mul(sc, arow[j]) : arow[j];
              else if (j == n) v[e] = ri;
            }
            jrow[wpp_col<BLK>(4 * q4) / 4] = make_float4(v[0], v[1], v[2], v[3]);
          }
        }
       }
      } else {
        jbuf[lane * NPS + wpp_col<BLK>(n)] = ri;
      }
    }
    __syncwarp();
    WPP_T(1);
    // the stage is free again: refill it with chunk c + nstages
    if (c + (int)pipe.nstages < nchunks) {
      if (lane == 0) fence_proxy_async();
      const int nrow0 = (c + (int)pipe.nstages) * kWppRows;
      pipe.issue(ap, yp, n, nrow0, (m - nrow0 < kWppRows) ? (m - nrow0) : kWppRows, pipe.stage, lane);
    }
    pipe.advance();
    // ---- phase 2: lane = 8x8 block of the upper triangle ----
    if (do_rebuild) {
      if (has_block) {
        const T *pa = jbuf + wpp_col<BLK>(bi * BLK);
        const T *pb = jbuf + wpp_col<BLK>(bj * BLK);
       if constexpr (!kF32) {
        for (int i = 0; i < nrows; ++i) {
          T a[BLK], b[BLK];
#pragma unroll
          for (int u = 0; u < BLK; ++u) { a[u] = pa[i * NPS + u]; b[u] = pb[i * NPS + u]; }
#pragma unroll
          for (int u = 0; u < BLK; ++u)
#pragma unroll
            for (int v = 0; v < BLK; ++v) acc[u][v] = O::fma(a[u], b[v], acc[u][v]);
        }
       } else if (TOB200_WPP_FOLD_PIPE) {
        // software pipeline: the operands of row i + 1 are requested before the 32 FFMA2 of row i issue
        float4 av[BLK / 4], bv[BLK / 4];
#pragma unroll
        for (int q = 0; q < BLK / 4; ++q) {
          av[q] = *reinterpret_cast<const float4 *>(pa + 4 * q);
          bv[q] = *reinterpret_cast<const float4 *>(pb + 4 * q);
        }
        for (int i = 0; i < nrows; ++i) {
          T a[BLK], b[BLK];
#pragma unroll
          for (int q = 0; q < BLK / 4; ++q) {
            a[4 * q] = av[q].x; a[4 * q + 1] = av[q].y; a[4 * q + 2] = av[q].z; a[4 * q + 3] = av[q].w;
            b[4 * q] = bv[q].x; b[4 * q + 1] = bv[q].y; b[4 * q + 2] = bv[q].z; b[4 * q + 3] = bv[q].w;
          }
          const int in = i + 1 < nrows ? i + 1 : i;  // (the last row re-reads itself: unused)
#pragma unroll
          for (int q = 0; q < BLK / 4; ++q) {
            av[q] = *reinterpret_cast<const float4 *>(pa + in * NPS + 4 * q);
            bv[q] = *reinterpret_cast<const float4 *>(pb + in * NPS + 4 * q);
          }
#pragma unroll
          for (int u = 0; u < BLK; ++u)
#pragma unroll
            for (int h = 0; h < BLK / 2; ++h) ffma2_bcast(acc2[u][h], a[u], b[2 * h], b[2 * h + 1]);
        }
       } else {
#pragma unroll 2
        for (int i = 0; i < nrows; ++i) {
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
            for (int h = 0; h < BLK / 2; ++h) ffma2_bcast(acc2[u][h], a[u], b[2 * h], b[2 * h + 1]);
        }
       }
      }
    } else if (lane == 0) {  // cost-only pass (solvers/gn.h:98-105): sum r_i^2 in row order
      for (int i = 0; i < nrows; ++i) {
        const T ri = jbuf[i * NPS + wpp_col<BLK>(n)];
        cost_only = O::fma(ri, ri, cost_only);
      }
    }
    __syncwarp();
    WPP_T(2);
    ycur = ynext;
  }
  cost_only = __shfl_sync(0xffffffffu, cost_only, 0);
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

// ---- moving the register blocks of [J|r]^T [J|r] out ---------------------------------------------------
// g = J^T r (column n of the augmented matrix), cost = r^T r (its corner), undamped diagonal of H
template <typename T, int NB, int BLK>
__device__ __forceinline__ void wpp_extract(const T (&acc)[BLK][BLK], int bi, int bj, bool has_block, int n, T *g,
                                            T *dg, T *cost_slot) {
  if (!has_block) return;
  const int vn = n - (NB - 1) * BLK;  // column n sits in the last block column
#pragma unroll
  for (int u = 0; u < BLK; ++u) {
    const int row = bi * BLK + u;
    if (bi == bj && row < n) dg[row] = acc[u][u];
    if (bj == NB - 1) {
#pragma unroll
      for (int v = 0; v < BLK; ++v) {
        if (v == vn) {
          if (row < n) g[row] = acc[u][v];
          else if (row == n) *cost_slot = acc[u][v];
        }
      }
    }
  }
}

// W <- P H P^T (lower triangle) with the damped diagonal dd; optionally the unpermuted damped H_ to
// the persistent global copy hp (lower-triangular layout W(i,j) = H(j,i))
template <typename T, int NB, int BLK>
__device__ __forceinline__ void wpp_store_permuted(T *W, const T (&acc)[BLK][BLK], int bi, int bj, bool has_block,
                                                   int n, const T *dd, const int *inv, T *hp) {
  constexpr int LDW = wpp_ldw(NB * BLK);
  if (!has_block) return;
  int pr[BLK], pc[BLK];
#pragma unroll
  for (int u = 0; u < BLK; ++u) {
    const int row = bi * BLK + u, col = bj * BLK + u;
    pr[u] = row < n ? inv[row] : 0;
    pc[u] = col < n ? inv[col] : 0;
  }
#pragma unroll
  for (int u = 0; u < BLK; ++u)
#pragma unroll
    for (int v = 0; v < BLK; ++v) {
      const int row = bi * BLK + u, col = bj * BLK + v;  // upper element (row, col)
      if (row <= col && col < n) {
        const T val = row == col ? dd[row] : acc[u][v];
        const int a = pr[u], b = pc[v];
        W[(a > b ? a : b) * LDW + (a > b ? b : a)] = val;
        if (hp) hp[col * LDW + row] = val;
      }
    }
}

// Everything after the data pass for one problem, warp-cooperative; mirrors lm_after_pass.
// INV selects the `hessian.use_ldlt = false` solve at compile time: the real call to wpp_lu_solve
// inside the retry loop would otherwise cost the LDLT kernels registers (measured -3% on C4).
template <typename T, int NB, int BLK, bool INV = false>
__device__ __forceinline__ void wpp_after_pass(LmScalars<T> &s, const DevOptions<T> &o, const WppData<T> &d,
                                               unsigned char *ws, T *hp, bool pass_rebuilt, int bi, int bj,
                                               bool has_block, const T (&acc)[BLK][BLK], T cost_only, int lane,
                                               bool always_persist = false, const double *cost_d = nullptr,
                                               int nres_d = 0) {
  using O = Ops<T>;
  constexpr int LDW = wpp_ldw(NB * BLK);
  const int n = d.n, nres = cost_d ? nres_d : d.m;
  T *W = reinterpret_cast<T *>(ws + d.L.stages);
  T *xs = reinterpret_cast<T *>(ws + d.L.xs);
  T *last_dx = reinterpret_cast<T *>(ws + d.L.last_dx);
  T *g = reinterpret_cast<T *>(ws + d.L.g);
  T *dxs = reinterpret_cast<T *>(ws + d.L.dxs);
  T *temp = reinterpret_cast<T *>(ws + d.L.temp);
  T *dg = reinterpret_cast<T *>(ws + d.L.dg);
  T *dd = reinterpret_cast<T *>(ws + d.L.dd);
  int *perm = reinterpret_cast<int *>(ws + d.L.perm);
  int *inv = reinterpret_cast<int *>(ws + d.L.inv);

  WPP_T0();
  T cost_t = cost_only;
  if (pass_rebuilt) {
    wpp_extract<T, NB, BLK>(acc, bi, bj, has_block, n, g, dg, temp);  // temp[0] <- r^T r
    __syncwarp();
    cost_t = temp[0];
    __syncwarp();
  }
  double cost;
  bool built_ok = cost_d ? lm_normalize_cost_d(o, *cost_d, nres, cost) : lm_normalize_cost(o, cost_t, nres, cost);
  if (pass_rebuilt) {
    s.num_builds++;
    if (built_ok) {
      if (o.grad_clipping != (T)0) {  // base.h:30-38
        for (int j = lane; j < n; j += 32) {
          T v = g[j];
          v = v < -o.grad_clipping ? -o.grad_clipping : v;
          v = v > o.grad_clipping ? o.grad_clipping : v;
          g[j] = v;
        }
        __syncwarp();
      }
      if (o.check_min_H_diag > (T)0) {  // lm.h:82-86
        bool low = false;
        for (int j = lane; j < n; j += 32) low = low || (O::abs(dg[j]) < o.check_min_H_diag);
        if (__any_sync(0xffffffffu, low)) built_ok = false;
      }
    }
  }
  // Will a cost-only iteration possibly follow this one?  Only then does H_ have to outlive the
  // next pass (optimizer.h:295: eval_only = !last_was_success, after a step that did not succeed).
  // (the step solver keeps H_ in HBM between calls and exports it as Output::final_hessian: always)
  const bool may_need_stale_h =
      always_persist || !pass_rebuilt || (!(cost - s.final_cost < 0.0) && !(s.flags & kFlagLastWasSuccess));

  bool solver_failed = true, early_return = false;
  const uint8_t max_tries = lm_max_tries(o);
  for (int attempt = 0; s.num_consec_failures <= max_tries; ++attempt) {
    if (built_ok) {
      // damped diagonal of H_ (lm.h:108-117): from the undamped one after a rebuild (a retry
      // re-accumulates the same H), cumulative on the stale H_ otherwise
      double sc;
      const bool damp = lm_damping_scale(s, o, pass_rebuilt, sc);
      for (int j = lane; j < n; j += 32) {
        const T base = pass_rebuilt ? dg[j] : hp[j * LDW + j];
        dd[j] = damp ? (T)((double)base * sc) : base;
      }
      __syncwarp();
      if constexpr (!INV) {
        wpp_pivot_order(dd, n, perm, inv, lane);
      } else {  // the LU of the inverse path pivots by itself: lay H_ out unpermuted
        for (int j = lane; j < n; j += 32) {
          perm[j] = j;
          inv[j] = j;
        }
        __syncwarp();
      }
      if (pass_rebuilt) {
        wpp_store_permuted<T, NB, BLK>(W, acc, bi, bj, has_block, n, dd, inv, may_need_stale_h ? hp : nullptr);
      } else {  // cost-only pass, or a retry of one: lay the persistent H_ out, publish its new diagonal
        for (int e = lane; e < n * LDW; e += 32) {
          const int i = e / LDW, j = e - i * LDW;
          if (j <= i) {
            const T val = i == j ? dd[i] : hp[e];
            const int a = inv[i], b = inv[j];
            W[(a > b ? a : b) * LDW + (a > b ? b : a)] = val;
          }
        }
        __syncwarp();
        for (int j = lane; j < n; j += 32) hp[j * LDW + j] = dd[j];
      }
      __syncwarp();
      WPP_T(3);
      if constexpr (!INV) {
        const bool fact_ok = wpp_ldlt_factor<T>(W, LDW, n, temp, dxs, NB * BLK, lane);  // gn.h:150-156
        WPP_T(4);
        if (fact_ok) {
          for (int j = lane; j < n; j += 32) temp[j] = -g[j];
          __syncwarp();
          wpp_ldlt_solve<T>(W, LDW, n, perm, temp, dxs, lane);
          solver_failed = false;
        }
      } else {  // gn.h:157-163: -H^-1 g, never fails
        wpp_lu_solve<T>(W, LDW, n, g, temp, dxs, lane);
        solver_failed = false;
      }
      WPP_T(5);
    }
    if (!solver_failed) break;
    const int act = lm_on_solver_failure(s, o, cost, nres);
    if (act == kLmEarlyReturn) early_return = true;
    if (act != kLmRetry) break;
    if (attempt >= 100000) break;
  }

  double dx_norm2 = 0.0, grad_norm2 = 0.0;
  if (!solver_failed) {
    dx_norm2 = (double)wpp_sqnorm<T>(dxs, n, lane);
    if (o.min_grad_norm2_f > 0.0f) grad_norm2 = (double)wpp_sqnorm<T>(g, n, lane);
  }
  bool success, has_dx;
  lm_finish_step(s, o, early_return, solver_failed, cost, nres, dx_norm2, grad_norm2, success, has_dx);
  const int action = lm_update_action(s, o, success, has_dx);
  if (action == kLmApplyDx || action == kLmProbeDx) {
    for (int j = lane; j < n; j += 32) {
      xs[j] = O::add(xs[j], dxs[j]);
      last_dx[j] = dxs[j];
    }
  } else if (action == kLmRollBack) {
    for (int j = lane; j < n; j += 32) xs[j] = O::add(xs[j], -last_dx[j]);
  }
  __syncwarp();
  WPP_T(6);
}

__device__ __forceinline__ int64_t wpp_next(unsigned long long *counter, int lane) {
  unsigned long long t = 0;
  if (lane == 0) t = atomicAdd(counter, 1ull);
  return (int64_t)__shfl_sync(0xffffffffu, t, 0);
}

// ---------------------------------------------------------------------------------------------
// tob200_lm_run_* for mid n
// ---------------------------------------------------------------------------------------------
template <typename T>
struct WppRunParams {
  WppData<T> d;
  DevOptions<T> opt;
  T alpha, alpha3;
  T *x;                    // [B][n] in/out
  tob200_result *results;  // [B]
  double *final_hessian;   // [B][n][n] or nullptr: Output::final_hessian (optimizer.h:313-316, lm.h:157-171)
};

template <typename T, int NB, int BLK, bool INV = false>
__global__ void __launch_bounds__(kWppThreads, sizeof(T) == 4 ? 3 : 1) wpp_lm_run_kernel(const __grid_constant__ WppRunParams<T> p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x / 32;
  unsigned char *ws = smem + (size_t)wid * p.d.L.total;
  const int n = p.d.n;
  constexpr int NP = NB * BLK, NPS = wpp_nps(NP), LDW = wpp_ldw(NP);
  WppPipe<T> pipe;
  pipe.init(ws, p.d, lane);
  T *xs = reinterpret_cast<T *>(ws + p.d.L.xs);
  T *last_dx = reinterpret_cast<T *>(ws + p.d.L.last_dx);
  T *jbuf = reinterpret_cast<T *>(ws + p.d.L.jbuf);
  T *hp = p.d.hpersist + ((size_t)blockIdx.x * (kWppThreads / 32) + wid) * (NP * LDW);
  int bi, bj;
  bool has_block;
  wpp_block_of_lane<NB>(lane, bi, bj, has_block);
  const bool is_lm = p.opt.solver_type == 0;

  for (int64_t pr = wpp_next(p.d.counter, lane); pr < p.d.B; pr = wpp_next(p.d.counter, lane)) {
    for (int j = lane; j < NP; j += 32) {
      xs[j] = j < n ? p.x[pr * n + j] : (T)0;
      last_dx[j] = (T)0;
    }
    LmScalars<T> s;
    s.reset_scalars(p.opt);
    __syncwarp();
    while (!s.done()) {
      const bool do_rebuild = !is_lm || s.rebuild();
      T acc[BLK][BLK], cost_only;
      wpp_pass<T, NB, BLK, true>(pipe, p.d, ws, pr, lane, do_rebuild, p.alpha, p.alpha3, bi, bj, has_block, acc, cost_only);
      wpp_after_pass<T, NB, BLK, INV>(s, p.opt, p.d, ws, hp, do_rebuild, bi, bj, has_block, acc, cost_only, lane,
                                      p.final_hessian != nullptr);  // the final Hessian needs H_ of every Build kept
    }
    for (int j = lane; j < n; j += 32) p.x[pr * n + j] = xs[j];
    if (lane == 0) lm_write_result(s, &p.results[pr]);
    if (p.final_hessian) {  // SolverLM::Hessian() (lm.h:157-171) from the persistent damped H_, widened to double
      __syncwarp();
      const bool undamp = is_lm && s.prev_lambda > (T)0;
      const T sc = Ops<T>::add((T)1, s.prev_lambda);
      double *o = p.final_hessian + (size_t)pr * n * n;
      for (int e = lane; e < n * n; e += 32) {
        const int j = e / n, k = e - j * n;
        if (j > k) continue;
        T v = hp[k * LDW + j];
        if (j == k && undamp) v = Ops<T>::div(v, sc);
        o[j * n + k] = (double)v;
        o[k * n + j] = (double)v;
      }
    }
    __syncwarp();
  }
#ifdef TOB200_WPP_TIMING
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    long long tot = 0;
    for (int k = 0; k < 7; ++k) tot += g_wpp_tm[k];
    printf("wpp warp 0: kcycles: tma wait %lld phase1 %lld phase2 %lld extract+pivot+layout %lld factor %lld solve %lld "
           "state+update %lld total %lld\n", g_wpp_tm[0] / 1000, g_wpp_tm[1] / 1000, g_wpp_tm[2] / 1000, g_wpp_tm[3] / 1000,
           g_wpp_tm[4] / 1000, g_wpp_tm[5] / 1000, g_wpp_tm[6] / 1000, tot / 1000);
    for (int k = 0; k < 16; ++k) g_wpp_tm[k] = 0;
  }
#endif
}

// ---------------------------------------------------------------------------------------------
// tob200_build_solve_* for mid n
// ---------------------------------------------------------------------------------------------
template <typename T>
struct WppBuildSolveParams {
  WppData<T> d;
  const T *lambda;
  T *dx;
  double *cost;
  T *H_out;
  T *g_out;
  int32_t *status;
};

template <typename T, int NB, int BLK>
__global__ void __launch_bounds__(kWppThreads, sizeof(T) == 4 ? 2 : 1) wpp_build_solve_kernel(const __grid_constant__ WppBuildSolveParams<T> p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x / 32;
  unsigned char *ws = smem + (size_t)wid * p.d.L.total;
  const int n = p.d.n;
  constexpr int NP = NB * BLK, NPS = wpp_nps(NP), LDW = wpp_ldw(NP);
  WppPipe<T> pipe;
  pipe.init(ws, p.d, lane);
  T *W = reinterpret_cast<T *>(ws + p.d.L.stages);
  T *g = reinterpret_cast<T *>(ws + p.d.L.g);
  T *dxs = reinterpret_cast<T *>(ws + p.d.L.dxs);
  T *temp = reinterpret_cast<T *>(ws + p.d.L.temp);
  T *dg = reinterpret_cast<T *>(ws + p.d.L.dg);
  T *dd = reinterpret_cast<T *>(ws + p.d.L.dd);
  int *perm = reinterpret_cast<int *>(ws + p.d.L.perm);
  int *inv = reinterpret_cast<int *>(ws + p.d.L.inv);
  T *jbuf = reinterpret_cast<T *>(ws + p.d.L.jbuf);
  int bi, bj;
  bool has_block;
  wpp_block_of_lane<NB>(lane, bi, bj, has_block);

  for (int64_t pr = wpp_next(p.d.counter, lane); pr < p.d.B; pr = wpp_next(p.d.counter, lane)) {
    T acc[BLK][BLK], cost_only;
    wpp_pass<T, NB, BLK, false>(pipe, p.d, ws, pr, lane, true, (T)0, (T)0, bi, bj, has_block, acc, cost_only);
    wpp_extract<T, NB, BLK>(acc, bi, bj, has_block, n, g, dg, temp);
    __syncwarp();
    const T cost_t = temp[0];
    const T lam = p.lambda ? p.lambda[pr] : (T)0;
    const double sc = 1.0 + (double)lam;  // solvers/lm.h:108-117
    for (int j = lane; j < n; j += 32) dd[j] = lam > (T)0 ? (T)((double)dg[j] * sc) : dg[j];
    __syncwarp();
    if (lane == 0) p.cost[pr] = (double)cost_t;
    if (p.g_out)
      for (int j = lane; j < n; j += 32) p.g_out[pr * n + j] = g[j];
    if (p.H_out && has_block) {  // damped H_, full symmetric
      T *Ho = p.H_out + (size_t)pr * n * n;
#pragma unroll
      for (int u = 0; u < BLK; ++u)
#pragma unroll
        for (int v = 0; v < BLK; ++v) {
          const int row = bi * BLK + u, col = bj * BLK + v;
          if (row <= col && col < n) {
            const T val = row == col ? dd[row] : acc[u][v];
            Ho[row * n + col] = val;
            Ho[col * n + row] = val;
          }
        }
    }
    wpp_pivot_order(dd, n, perm, inv, lane);
    wpp_store_permuted<T, NB, BLK>(W, acc, bi, bj, has_block, n, dd, inv, nullptr);
    __syncwarp();
    const bool ok = wpp_ldlt_factor<T>(W, LDW, n, temp, dxs, NP, lane);  // math.h:232-240
    if (ok) {
      for (int j = lane; j < n; j += 32) temp[j] = -g[j];  // solvers/gn.h:155
      __syncwarp();
      wpp_ldlt_solve<T>(W, LDW, n, perm, temp, dxs, lane);
      for (int j = lane; j < n; j += 32) p.dx[pr * n + j] = dxs[j];
    }
    if (lane == 0) p.status[pr] = ok ? 0 : 1;
    __syncwarp();
  }
}

}  // namespace tob200
