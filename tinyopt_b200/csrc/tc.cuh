// tc.cuh — tcgen05 / TMEM / TMA PTX wrappers shared by the tensor-core kernels (lg.cuh: large n,
// wtc.cuh: mid n).  sm_100a only.
#pragma once

#include <cuda.h>  // CUtensorMap (types only)
#include <cuda_fp16.h>

#include "common.cuh"

namespace tob200 {

// ---- tcgen05 wrappers ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by one thread
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same for kind::f16 (A, B = FP16, D = FP32): K = 16 per instruction, the same 32 bytes of K per operand row
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle.  Canonical layout (16-byte units):
// ((8, n), 2) : ((1, SBO), LBO) — a core matrix is 8 MN rows x 16 bytes (4 tf32 of K), rows 16 bytes
// apart; SBO = byte distance between 8-row groups along MN, LBO = between the two 4-element K chunks
// of one K = 8 instruction.  (Probed on the B200 with tools/tc_probe.cu: MN-major tf32 operands
// produce no output, K-major ones are exact.)
__device__ __forceinline__ uint64_t tc_desc_k_major(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  return d;         // layout type 0: no swizzle
}
// round-to-nearest TF32 (low 13 mantissa bits cleared): hi part of the 3xTF32 split; lo = v - hi is
// then exact and of either sign, so the hardware's truncation of lo does not bias the products
__device__ __forceinline__ float tc_round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128
__device__ __forceinline__ uint32_t tc_idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
// instruction descriptor: D = F32, A = B = FP16 (format 0), both K-major, M = 128
__device__ __forceinline__ uint32_t tc_idesc_f16(int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
// power-of-two scale of the FP16 split: |J_ij| * 2^e < 2^12 for every entry of the problem, so hi = fp16(v 2^e) never
// overflows (fp16 max 65504), lo = fp16(v 2^e - hi) is a normal fp16 for every entry that matters, the products
// hi hi' + hi lo' + lo hi' carry 22 significant bits - the accuracy class of the 3xTF32 split - and the FP32
// accumulator stays far from its range (2^24 per product, m of them).  H = acc * 2^(-2e).
__device__ __forceinline__ int lg_fp16_exp(float amax) {
  if (!(amax > 0.f) || !(amax < 3.0e38f)) return 0;  // zero, NaN or Inf: nothing sensible to scale
  int e;
  frexpf(amax, &e);  // amax = f * 2^e, f in [0.5, 1)
  e = 12 - e;
  return e < -60 ? -60 : (e > 60 ? 60 : e);
}
__device__ __forceinline__ uint32_t lg_pack_h2(float a, float b) {  // two FP16 (round to nearest) in one register
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float lg_h_lo(uint32_t h2) { return __half2float(__ushort_as_half((unsigned short)(h2 & 0xffffu))); }
__device__ __forceinline__ float lg_h_hi(uint32_t h2) { return __half2float(__ushort_as_half((unsigned short)(h2 >> 16))); }
__device__ __forceinline__ void sts_v4u(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// explicit shared-space accesses (the stage pointers go through an integer round-trip for alignment, so
// the compiler would otherwise emit generic LD / ST, which are tracked like global accesses)
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t saddr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 2-D tiled TMA: one box of the tensor map (column c, row r = innermost-first coordinates) -> shared memory
__device__ __forceinline__ void tma_load_box(void *smem_dst, const CUtensorMap *tmap, int c, int r, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c), "r"(r), "r"(smem_u32(bar))
      : "memory");
}
// the same box delivered to every CTA of the cluster named in `cta_mask` (same CTA-relative destination and mbarrier
// offsets in each of them; each destination's mbarrier receives the box's bytes)
__device__ __forceinline__ void tma_load_box_multicast(void *smem_dst, const CUtensorMap *tmap, int c, int r, uint64_t *bar,
                                                       uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], "
      "[%4], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c), "r"(r), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// ---- thread-block cluster helpers ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same CTA-relative address in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// wait with cluster-scope acquire (the arrival comes from another CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_prefetch_box(const CUtensorMap *tmap, int c, int r) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(c), "r"(r)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2(const void *gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace tob200
