// misc_kernels.cu — layout conversion and the synthetic problem family on device.
// The generator is bit-identical to the CPU oracle's: same counter hash, same op sequence.
#include "../../include/tinyopt_b200.h"
#include "common.cuh"
#include "internal.h"

#include <cmath>

namespace tob200 {

cudaError_t raise_smem_limit(const void *fn, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<int, const void *>, size_t> limit;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  size_t &cur = limit[std::make_pair(dev, fn)];
  if (smem > cur) {
    if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    cur = smem;
  }
  return cudaSuccess;
}


// ---- PROBLEM_MAJOR [B][mn] -> TILE32 [ntiles][mn][32]: 32x32 transposes through shared memory ----
template <typename T>
__global__ void retile_kernel(const T *__restrict__ src, int64_t B, int64_t mn, T *__restrict__ dst) {
  __shared__ T tile[32][33];
  // element chunks on grid.x (2^31 - 1 blocks), tiles on a grid-stride loop over grid.y (<= 65535)
  const int64_t e0 = (int64_t)blockIdx.x * 32;
  const int64_t ntiles = (B + 31) / 32;
  for (int64_t t = blockIdx.y; t < ntiles; t += gridDim.y) {
    for (int pl = threadIdx.y; pl < 32; pl += blockDim.y) {
      const int64_t p = t * 32 + pl, e = e0 + threadIdx.x;
      tile[pl][threadIdx.x] = (p < B && e < mn) ? src[p * mn + e] : (T)0;
    }
    __syncthreads();
    for (int el = threadIdx.y; el < 32; el += blockDim.y) {
      const int64_t e = e0 + el;
      if (e < mn) dst[(t * mn + e) * 32 + threadIdx.x] = tile[threadIdx.x][el];
    }
    __syncthreads();
  }
}

template <typename T>
cudaError_t launch_retile(const T *src, int64_t B, int m, int n, T *dst, cudaStream_t st) {
  const int64_t mn = (int64_t)m * n;
  if (B <= 0 || mn <= 0) return cudaSuccess;
  const int64_t ntiles = (B + 31) / 32;
  dim3 grid((unsigned)((mn + 31) / 32), (unsigned)(ntiles < 65535 ? ntiles : 65535)), block(32, 8);
  retile_kernel<T><<<grid, block, 0, st>>>(src, B, mn, dst);
  return cudaGetLastError();
}

// ---- TILE32 -> PROBLEM_MAJOR (inverse) ----
template <typename T>
__global__ void untile_kernel(const T *__restrict__ src, int64_t B, int64_t mn, T *__restrict__ dst) {
  __shared__ T tile[32][33];
  const int64_t e0 = (int64_t)blockIdx.x * 32;
  const int64_t ntiles = (B + 31) / 32;
  for (int64_t t = blockIdx.y; t < ntiles; t += gridDim.y) {
    for (int el = threadIdx.y; el < 32; el += blockDim.y) {
      const int64_t e = e0 + el;
      tile[el][threadIdx.x] = e < mn ? src[(t * mn + e) * 32 + threadIdx.x] : (T)0;
    }
    __syncthreads();
    for (int pl = threadIdx.y; pl < 32; pl += blockDim.y) {
      const int64_t p = t * 32 + pl, e = e0 + threadIdx.x;
      if (p < B && e < mn) dst[p * mn + e] = tile[threadIdx.x][pl];
    }
    __syncthreads();
  }
}

template <typename T>
cudaError_t launch_untile(const T *src, int64_t B, int m, int n, T *dst, cudaStream_t st) {
  const int64_t mn = (int64_t)m * n;
  if (B <= 0 || mn <= 0) return cudaSuccess;
  const int64_t ntiles = (B + 31) / 32;
  dim3 grid((unsigned)((mn + 31) / 32), (unsigned)(ntiles < 65535 ? ntiles : 65535)), block(32, 8);
  untile_kernel<T><<<grid, block, 0, st>>>(src, B, mn, dst);
  return cudaGetLastError();
}

// ---- counter RNG: splitmix64 finaliser of seed ^ (p * golden + k) ----
__device__ __forceinline__ uint64_t hash64(uint64_t seed, uint64_t p, uint64_t k) {
  uint64_t z = seed ^ (p * 0x9E3779B97F4A7C15ull + k);
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
template <typename T> __device__ __forceinline__ T unif01(uint64_t h);
template <> __device__ __forceinline__ float unif01<float>(uint64_t h) {
  return __fmul_rn((float)(uint32_t)(h >> 40), 1.0f / 16777216.0f);
}
template <> __device__ __forceinline__ double unif01<double>(uint64_t h) {
  return __dmul_rn((double)(h >> 11), 1.0 / 9007199254740992.0);
}
template <typename T>
__device__ __forceinline__ T unif_pm1(uint64_t seed, uint64_t p, uint64_t k) {
  return Ops<T>::sub(Ops<T>::mul((T)2, unif01<T>(hash64(seed, p, k))), (T)1);
}

__device__ __forceinline__ size_t at_J(int layout, int64_t b, int i, int j, int m, int n) {
  return layout == TOB200_LAYOUT_TILE32 ? ((((size_t)(b / 32) * m + i) * n + j) * 32 + (size_t)(b % 32))
                                        : (((size_t)b * m + i) * n + j);
}
__device__ __forceinline__ size_t at_r(int layout, int64_t b, int i, int m) {
  return layout == TOB200_LAYOUT_TILE32 ? (((size_t)(b / 32) * m + i) * 32 + (size_t)(b % 32)) : ((size_t)b * m + i);
}

// thread (b, i): row i of problem p0 + b.  Threads are ordered lane-fastest inside a tile so the
// TILE32 stores coalesce.
template <typename T>
__global__ void synth_rows_kernel(uint64_t seed, int64_t p0, int64_t B, int m, int n, T alpha, T sigma, T rsn, T sqrt3,
                                  int layout, T *A, T *y) {
  using O = Ops<T>;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t ntiles = (B + 31) / 32;
  if (gid >= ntiles * m * 32) return;
  const int lane = (int)(gid % 32);
  const int i = (int)((gid / 32) % m);
  const int64_t b = (gid / 32 / m) * 32 + lane;
  if (b >= B) {
    if (layout == TOB200_LAYOUT_TILE32) {  // keep the pad lanes of the last tile finite
      if (A) for (int j = 0; j < n; ++j) A[at_J(layout, b, i, j, m, n)] = (T)0;
      if (y) y[at_r(layout, b, i, m)] = (T)0;
    }
    return;
  }
  const uint64_t p = (uint64_t)(p0 + b);
  const uint64_t kx = (uint64_t)m * n, kz = kx + n;
  T t = (T)0;
  for (int j = 0; j < n; ++j) {
    const T a = O::mul(unif_pm1<T>(seed, p, (uint64_t)i * n + j), rsn);
    if (A) A[at_J(layout, b, i, j, m, n)] = a;
    t = O::fma(a, unif_pm1<T>(seed, p, kx + j), t);
  }
  if (y) {
    const T w = O::fma(alpha, O::mul(t, t), (T)1);
    const T z = O::mul(unif_pm1<T>(seed, p, kz + i), sqrt3);
    y[at_r(layout, b, i, m)] = O::fma(sigma, z, O::mul(t, w));
  }
}

template <typename T>
__global__ void synth_x_kernel(uint64_t seed, int64_t p0, int64_t B, int m, int n, T *xstar, T *x0) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * n) return;
  const int64_t b = gid / n;
  const int j = (int)(gid % n);
  const uint64_t p = (uint64_t)(p0 + b);
  const uint64_t kx = (uint64_t)m * n, k0 = kx + n + m;
  const T xs = unif_pm1<T>(seed, p, kx + j);
  if (xstar) xstar[gid] = xs;
  if (x0) x0[gid] = Ops<T>::fma((T)0.3, unif_pm1<T>(seed, p, k0 + j), xs);
}

template <typename T>
cudaError_t launch_synth_generate(uint64_t seed, int64_t p0, int64_t B, int m, int n, T alpha, T sigma, int layout,
                                  T *A, T *y, T *xstar, T *x0, cudaStream_t st, int *launches) {
  *launches = 0;
  if (B <= 0) return cudaSuccess;
  const T rsn = (T)1 / std::sqrt((T)n);
  const T sqrt3 = std::sqrt((T)3);
  if ((A || y) && m > 0) {
    const int64_t total = ((B + 31) / 32) * m * 32;
    synth_rows_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(seed, p0, B, m, n, alpha, sigma, rsn, sqrt3,
                                                                        layout, A, y);
    ++*launches;
  }
  if (xstar || x0) {
    const int64_t total = B * n;
    synth_x_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(seed, p0, B, m, n, xstar, x0);
    ++*launches;
  }
  return cudaGetLastError();
}

// r_i = t (1 + alpha t^2) - y_i, J_ij = (1 + 3 alpha t^2) A_ij, t = A_i . x   (thread per row)
template <typename T>
__global__ void synth_eval_kernel(const T *A, const T *y, T alpha, T alpha3, int layout, int64_t B, int m, int n,
                                  const T *x, T *r, T *J) {
  using O = Ops<T>;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t ntiles = (B + 31) / 32;
  if (gid >= ntiles * m * 32) return;
  const int lane = (int)(gid % 32);
  const int i = (int)((gid / 32) % m);
  const int64_t b = (gid / 32 / m) * 32 + lane;
  if (b >= B) {
    if (layout == TOB200_LAYOUT_TILE32) {
      if (J) for (int j = 0; j < n; ++j) J[at_J(layout, b, i, j, m, n)] = (T)0;
      if (r) r[at_r(layout, b, i, m)] = (T)0;
    }
    return;
  }
  T t = (T)0;
  for (int j = 0; j < n; ++j) t = O::fma(A[at_J(layout, b, i, j, m, n)], x[b * n + j], t);
  const T t2 = O::mul(t, t);
  if (r) r[at_r(layout, b, i, m)] = O::fma(t, O::fma(alpha, t2, (T)1), -y[at_r(layout, b, i, m)]);
  if (J) {
    const T sc = O::fma(alpha3, t2, (T)1);
    for (int j = 0; j < n; ++j) J[at_J(layout, b, i, j, m, n)] = O::mul(sc, A[at_J(layout, b, i, j, m, n)]);
  }
}

template <typename T>
cudaError_t launch_synth_eval(const T *A, const T *y, T alpha, int layout, int64_t B, int m, int n, const T *x, T *r,
                              T *J, cudaStream_t st) {
  if (B <= 0 || m <= 0) return cudaSuccess;
  const int64_t total = ((B + 31) / 32) * m * 32;
  synth_eval_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(A, y, alpha, (T)3 * alpha, layout, B, m, n, x,
                                                                      r, J);
  return cudaGetLastError();
}

// dst[b][i][c] (i < rd, c < pd) <- src[b][i][c] where that exists (i < rs, c < ps), else 0: pads / strips
// rows and columns of a batch of row-major matrices (the large-n family wants n % 4 == 0)
__global__ void repitch_kernel(const float *src, int64_t B, int rs, int ps, float *dst, int rd, int pd) {
  const int64_t total = B * (int64_t)rd * pd;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % pd);
    const int64_t q = e / pd;
    const int i = (int)(q % rd);
    const int64_t b = q / rd;
    dst[e] = (i < rs && c < ps) ? src[(b * rs + i) * (int64_t)ps + c] : 0.f;
  }
}
cudaError_t launch_repitch(const float *src, int64_t B, int rs, int ps, float *dst, int rd, int pd, cudaStream_t st) {
  const int64_t total = B * (int64_t)rd * pd;
  if (total <= 0) return cudaSuccess;
  int64_t grid = (total + 255) / 256;
  if (grid > 148 * 32) grid = 148 * 32;
  repitch_kernel<<<(unsigned)grid, 256, 0, st>>>(src, B, rs, ps, dst, rd, pd);
  return cudaGetLastError();
}

// dst[b][0..nd) <- src[b][0..nd) for the problems with status[b] == 0 (dx of a rejected solve stays untouched)
__global__ void repitch_masked_kernel(const float *src, const int32_t *status, int64_t B, int ps, float *dst, int nd) {
  const int64_t total = B * (int64_t)nd;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / nd;
    const int c = (int)(e % nd);
    if (status[b] == 0) dst[e] = src[b * ps + c];
  }
}
cudaError_t launch_repitch_masked(const float *src, const int32_t *status, int64_t B, int ps, float *dst, int nd,
                                  cudaStream_t st) {
  const int64_t total = B * (int64_t)nd;
  if (total <= 0) return cudaSuccess;
  int64_t grid = (total + 255) / 256;
  if (grid > 148 * 32) grid = 148 * 32;
  repitch_masked_kernel<<<(unsigned)grid, 256, 0, st>>>(src, status, B, ps, dst, nd);
  return cudaGetLastError();
}

#define INST(T)                                                                                                   \
  template cudaError_t launch_retile<T>(const T *, int64_t, int, int, T *, cudaStream_t);                          \
  template cudaError_t launch_untile<T>(const T *, int64_t, int, int, T *, cudaStream_t);                          \
  template cudaError_t launch_synth_generate<T>(uint64_t, int64_t, int64_t, int, int, T, T, int, T *, T *, T *, T *, \
                                                cudaStream_t, int *);                                              \
  template cudaError_t launch_synth_eval<T>(const T *, const T *, T, int, int64_t, int, int, const T *, T *, T *,  \
                                            cudaStream_t);
INST(float)
INST(double)

}  // namespace tob200
