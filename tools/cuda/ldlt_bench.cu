// Latency of the warp LDL^T (wpp_ldlt_factor / wpp_ldlt_solve / wpp_pivot_order) in isolation:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I tinyopt_b200/csrc tools/cuda/ldlt_bench.cu -o gpurun_out/ldlt_bench
// usage: ldlt_bench [n] [warps per block] [blocks]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "wpp.cuh"
// #define WTC_LDLT_PROF 1
#include "wtc.cuh"
using namespace tob200;

__global__ void bench(const float *H, int n, int ldw, int reps, long long *out) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float *W = sm + (size_t)w * ((n + 1) * ldw + 16 * 64);
  float *vec = W + (n + 1) * ldw;
  float *temp = vec, *tb = vec + 64, *dd = vec + 3 * 64, *g = vec + 4 * 64, *x = vec + 5 * 64;
  int *perm = (int *)(vec + 6 * 64), *inv = (int *)(vec + 7 * 64);
  long long tf = 0, ts = 0, tp = 0;
  for (int j = lane; j < n; j += 32) { dd[j] = H[j * n + j]; g[j] = 1.f + j; }
  __syncwarp();
  bool ok = true;
  for (int r = 0; r < reps; ++r) {
    long long t0 = clock64();
    wpp_pivot_order(dd, n, perm, inv, lane);
    long long t1 = clock64();
    for (int e = lane; e < n * n; e += 32) {
      const int i = e / n, j = e % n;
      if (j <= i) { const int a = inv[i], b = inv[j]; W[(a > b ? a : b) * ldw + (a > b ? b : a)] = H[i * n + j]; }
    }
    __syncwarp();
    long long t2 = clock64();
    ok = wpp_ldlt_factor<float>(W, ldw, n, temp, tb, 64, lane) && ok;
    long long t3 = clock64();
    wpp_ldlt_solve<float>(W, ldw, n, perm, g, x, lane);
    long long t4 = clock64();
    tp += t1 - t0; tf += t3 - t2; ts += t4 - t3;
  }
  // the latency-optimised factorisation of wtc.cuh (right-hand side as row n) + back substitution
  long long tf2 = 0, ts2 = 0;
  float *cs1 = vec + 2 * 64;
  for (int j = lane; j < 64; j += 32) cs1[j] = 1.f;
  for (int e = lane; e < (n + 1) * ldw; e += 32) W[e] = 0.f;
  __syncwarp();
  for (int r = 0; r < reps; ++r) {
    for (int e = lane; e < n * n; e += 32) {
      const int i = e / n, j = e % n;
      if (j <= i) { const int a = inv[i], b = inv[j]; W[(a > b ? a : b) * ldw + (a > b ? b : a)] = H[i * n + j]; }
    }
    for (int j = lane; j < n; j += 32) W[n * ldw + inv[j]] = g[j];
    __syncwarp();
    long long t2 = clock64();
    ok = wtc_ldlt_fast(W, ldw, n, temp, tb, lane) && ok;
    long long t3 = clock64();
    wtc_back_subst(W, ldw, n, perm, cs1, x, lane);
    long long t4 = clock64();
    tf2 += t3 - t2; ts2 += t4 - t3;
  }
  long long tf4 = 0;
  for (int r = 0; r < reps; ++r) {
    for (int e = lane; e < (n + 1) * ldw; e += 32) W[e] = 0.f;
    __syncwarp();
    for (int e = lane; e < n * n; e += 32) {
      const int i = e / n, j = e % n;
      if (j <= i) { const int a = inv[i], b = inv[j]; W[(a > b ? a : b) * ldw + (a > b ? b : a)] = H[i * n + j]; }
    }
    for (int j = lane; j < n; j += 32) W[n * ldw + inv[j]] = g[j];
    __syncwarp();
    long long t2 = clock64();
    ok = wtc_ldlt_fast4(W, ldw, n, temp, vec + 8 * 64, lane) && ok;
    long long t3 = clock64();
    wtc_back_subst(W, ldw, n, perm, cs1, x, lane);
    tf4 += t3 - t2;
  }
  if (lane == 0 && blockIdx.x == 0 && w == 0) { out[113] = tf4 / reps; out[114] = (long long)(x[0] * 1e6f); }
  if (lane == 0 && blockIdx.x == 0 && w == 0) { out[110] = tf2 / reps; out[111] = ts2 / reps; out[112] = (long long)(x[0] * 1e6f); }
  if (lane == 0 && blockIdx.x == 0) { out[3 * w] = tp / reps; out[3 * w + 1] = tf / reps; out[3 * w + 2] = ts / reps; }
  if (!ok && lane == 0) out[100] = 1;
  if (lane == 0 && w == 0 && blockIdx.x == 0) out[101] = (long long)(x[0] * 1e6f);
}

int main(int argc, char **argv) {
  int n = argc > 1 ? atoi(argv[1]) : 50, warps = argc > 2 ? atoi(argv[2]) : 1, blocks = argc > 3 ? atoi(argv[3]) : 1;
  const int ldw = wtc_ldw(n);
  std::vector<float> H(n * n);
  // SPD: A^T A + I with a pseudo-random A
  std::vector<float> A(200 * n);
  unsigned s = 12345;
  for (auto &v : A) { s = s * 1664525u + 1013904223u; v = ((s >> 8) & 0xffff) / 65536.f - 0.5f; }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double acc = i == j ? 1.0 : 0.0; for (int k = 0; k < 200; ++k) acc += (double)A[k * n + i] * A[k * n + j]; H[i * n + j] = (float)acc; }
  float *dH; long long *dout;
  cudaMalloc(&dH, H.size() * 4); cudaMemcpy(dH, H.data(), H.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&dout, 128 * 8); cudaMemset(dout, 0, 128 * 8);
  size_t smem = (size_t)warps * ((n + 1) * ldw + 16 * 64) * 4;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bench<<<blocks, warps * 32, smem>>>(dH, n, ldw, 50, dout);
  cudaError_t e = cudaDeviceSynchronize();
  long long out[128]; cudaMemcpy(out, dout, sizeof(out), cudaMemcpyDeviceToHost);
  printf("n=%d warps=%d blocks=%d (%s): warp 0 cycles: pivot order %lld factor %lld solve %lld  fail=%lld x0=%lld\n", n, warps, blocks,
         cudaGetErrorString(e), out[0], out[1], out[2], out[100], out[101]);
#ifdef WTC_LDLT_PROF
  { long long pr[8]; cudaMemcpyFromSymbol(pr, g_ldlt_prof, sizeof(pr)); int steps = 50 * warps * blocks * ((n + 1) / 2);
    printf("   fast factor per step (cycles, lane 0): sweep + next T %lld pivot block %lld finish + store + sync %lld\n", pr[1] / steps, pr[2] / steps, pr[3] / steps); }
#endif
  printf("   fast4: factor %lld x0=%lld\n", out[113], out[114]);
  printf("   fast: factor %lld back-substitution %lld x0=%lld\n", out[110], out[111], out[112]);
  if (warps > 1) printf("   last warp: pivot %lld factor %lld solve %lld\n", out[3 * (warps - 1)], out[3 * (warps - 1) + 1], out[3 * (warps - 1) + 2]);
  return 0;
}
