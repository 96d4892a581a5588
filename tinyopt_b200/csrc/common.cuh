// common.cuh — device helpers shared by every kernel: pinned-rounding arithmetic, mbarrier + TMA
// bulk-copy PTX wrappers, packed-triangle indexing.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tob200 {

constexpr int kTile = 32;  // problems per TILE32 tile == lanes per warp

// ---- arithmetic with the rounding and (non-)fusion written out --------------------------------
// The canonical op sequence is a chain of IEEE round-to-nearest fma / mul / add /
// div in the problem's scalar type.  These intrinsics are never re-associated or contracted by
// nvcc, so the sequence in the source is the sequence in SASS (the TU is also built -fmad=false).
template <typename T> struct Ops;
template <> struct Ops<float> {
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
  static __device__ __forceinline__ float max_value() { return 3.402823466e+38f; }
  static __device__ __forceinline__ float min_normal() { return 1.175494351e-38f; }
  static __device__ __forceinline__ float float_eps() { return 1e-4f; }  // math.h:298-301
};
template <> struct Ops<double> {
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double abs(double a) { return fabs(a); }
  static __device__ __forceinline__ double max_value() { return 1.7976931348623157e+308; }
  static __device__ __forceinline__ double min_normal() { return 2.2250738585072014e-308; }
  static __device__ __forceinline__ double float_eps() { return (double)1e-7f; }  // math.h:298-301
};

// acc (two packed floats) <- fma(a, (b0, b1), acc): Blackwell FFMA2, one issue slot for two IEEE fmas
__device__ __forceinline__ void ffma2_bcast(unsigned long long &acc, float a, float b0, float b1) {
  unsigned long long av, bv;
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bv) : "f"(b0), "f"(b1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(av), "l"(bv));
}

// packed upper triangle of an N x N symmetric matrix, row-major: (j,k), j <= k
__host__ __device__ constexpr int tri_count(int n) { return n * (n + 1) / 2; }
__host__ __device__ constexpr int tri_index(int n, int j, int k) { return j * n - j * (j - 1) / 2 + (k - j); }

// ---- mbarrier / TMA bulk copy (cp.async.bulk, SASS: UBLKCP) ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// generic-proxy accesses to shared memory -> later async-proxy (TMA) accesses to the same bytes
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// global -> shared bulk copy, completion counted in bytes on `bar`.  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                             uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace tob200
