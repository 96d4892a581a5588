// cov_kernels.cu — SURVEY.md §8(f) rank 2: the step AFTER the path.  Batched tinyopt::InvCov(H)
// (math.h:44-57 DenseInvCov: `H.selfadjointView<Upper>().ldlt()` then `chol.solve(Identity)`) and
// MaxStdDev (solvers/lm.h:176-187, solvers/gn.h:177: sqrt of the largest coefficient of InvCov(H)),
// i.e. what Output::Covariance() (output.h:81-103) and SolverLM::Covariance() (lm.h:173) evaluate.
//
// n <= 64, float or double: one warp per problem, the matrix lives in shared memory.  The
// factorisation is the same diagonal-pivoted LDL^T as the solve of the hot path (wpp.cuh: pivot order
// from the diagonal, W = P H P^T laid out permuted, left-looking, every dot product a left-to-right
// fma chain) and every column of the inverse is one LDLT::solve of a unit vector with the
// substitutions in the oracle's update order, so the result is bit-identical to the CPU oracle's
// too_inv_cov.  Larger n (float) goes through lg_solve_kernel mode 3 (lg_solve.cuh).
#include "internal.h"
#include "wpp.cuh"

namespace tob200 {

constexpr int kCovMaxN = 64;

// Eigen's pivot search replayed literally on the diagonal (lane 0; n <= 64): at step k the FIRST
// largest |d| among positions k..n-1 is swapped to k (maxCoeff visitor: strict >, so a NaN never
// wins and a NaN sitting at k stays).  perm[pos] = original index, inv[orig] = pos.
template <typename T>
__device__ void cov_pivot_order(const T *dd, int n, int *perm, int *inv, int lane) {
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

// unpivoted left-looking LDL^T of the permuted lower matrix W (pitch ldw), lane = row; same
// bookkeeping (sign, zero pivots) as wpp_ldlt_factor / the oracle's ldlt_factor_
template <typename T>
__device__ bool cov_ldlt_factor(T *W, int ldw, int n, T *temp, int lane) {
  using O = Ops<T>;
#define WW(i, j) W[(i) * ldw + (j)]
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

template <typename T>
struct CovParams {
  const T *H;       // [B][n][n] row-major, only the upper triangle is read
  T *cov;           // [B][n][n] out (untouched where status != 0), or nullptr
  T *max_std;       // [B] out: sqrt(max coefficient of the inverse), 0 where status != 0; or nullptr
  int32_t *status;  // [B]: 0 ok, 1 rejected (info() != Success or not positive)
  int64_t B;
  int n, ldw;
  uint32_t warp_bytes;
};

template <typename T>
__global__ void __launch_bounds__(128) cov_warp_kernel(const __grid_constant__ CovParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int n = p.n, ldw = p.ldw;
  unsigned char *ws = smem + (size_t)wid * p.warp_bytes;
  T *W = reinterpret_cast<T *>(ws);
  T *dd = W + (size_t)n * ldw;
  T *temp = dd + n;
  T *rhs = temp + n;
  T *xs = rhs + n;
  int *perm = reinterpret_cast<int *>(xs + n);
  int *inv = perm + n;
  using O = Ops<T>;
  for (int64_t pr = (int64_t)blockIdx.x * wpc + wid; pr < p.B; pr += (int64_t)gridDim.x * wpc) {
    const T *Hp = p.H + (size_t)pr * n * n;
    T *Cp = p.cov ? p.cov + (size_t)pr * n * n : nullptr;
    if (n == 1) {  // DenseInvCov: m.inverse() for a 1 x 1 (math.h:49-50), unprotected as in the reference
      if (lane == 0) {
        const T v = O::div((T)1, Hp[0]);
        if (Cp) Cp[0] = v;
        if (p.max_std) p.max_std[pr] = sqrt(v);
        p.status[pr] = 0;
      }
      continue;
    }
    for (int j = lane; j < n; j += 32) dd[j] = Hp[(size_t)j * n + j];
    __syncwarp();
    cov_pivot_order<T>(dd, n, perm, inv, lane);
    for (int e = lane; e < n * n; e += 32) {  // W <- P H P^T, lower triangle, from the upper one of H
      const int i = e / n, j = e - i * n;
      if (j >= i) {
        const int a = inv[i], b = inv[j];
        W[(a > b ? a : b) * ldw + (a > b ? b : a)] = Hp[e];
      }
    }
    __syncwarp();
    const bool ok = cov_ldlt_factor<T>(W, ldw, n, temp, lane);
    T best = (T)0;
    bool have = false;
    if (ok) {
      for (int c = 0; c < n; ++c) {
        for (int j = lane; j < n; j += 32) rhs[j] = (j == c) ? (T)1 : (T)0;
        __syncwarp();
        wpp_ldlt_solve<T>(W, ldw, n, perm, rhs, xs, lane);  // column c of the inverse
        for (int i = lane; i < n; i += 32) {
          const T v = xs[i];
          if (Cp) Cp[(size_t)i * n + c] = v;
          if (!have || v > best) { best = v; have = true; }  // maxCoeff
        }
        __syncwarp();
      }
    }
    if (p.max_std) {
      // warp maximum of the lanes that saw a value
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const T ob = __shfl_xor_sync(0xffffffffu, best, off);
        const bool oh = __shfl_xor_sync(0xffffffffu, (int)have, off) != 0;
        if (oh && (!have || ob > best)) { best = ob; have = true; }
      }
      if (lane == 0) p.max_std[pr] = ok ? sqrt(best) : (T)0;
    }
    if (lane == 0) p.status[pr] = ok ? 0 : 1;
    __syncwarp();
  }
}

template <typename T>
cudaError_t launch_cov_warp(const T *H, int64_t B, int n, T *cov, T *max_std, int32_t *status, int num_sms,
                            cudaStream_t st) {
  CovParams<T> p;
  p.H = H; p.cov = cov; p.max_std = max_std; p.status = status; p.B = B; p.n = n;
  p.ldw = n | 1;  // odd pitch: a lane-per-row walk along a column is conflict free
  size_t wb = ((size_t)n * p.ldw + 4 * (size_t)n) * sizeof(T) + 2 * (size_t)n * sizeof(int);
  wb = (wb + 15) & ~(size_t)15;
  p.warp_bytes = (uint32_t)wb;
  int wpc = 4;
  while (wpc > 1 && wpc * wb > 200 * 1024) wpc >>= 1;
  const size_t smem = wpc * wb;
  cudaError_t e = cudaFuncSetAttribute(cov_warp_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int64_t grid = (B + wpc - 1) / wpc;
  const int64_t cap = (int64_t)num_sms * 8;
  if (grid > cap) grid = cap;
  cov_warp_kernel<T><<<(unsigned)grid, wpc * 32, smem, st>>>(p);
  return cudaGetLastError();
}

template cudaError_t launch_cov_warp<float>(const float *, int64_t, int, float *, float *, int32_t *, int, cudaStream_t);
template cudaError_t launch_cov_warp<double>(const double *, int64_t, int, double *, double *, int32_t *, int,
                                             cudaStream_t);

}  // namespace tob200
