// wtc.cuh — mid-n tensor-core family (13 <= n <= 55, float; config C4: n = 50, m = 500).
//
// tob200_lm_run_f32 for the sizes where H = J^T J is too big for FFMA register blocks to reach the HBM roofline
// (the warp-per-problem kernel of wpp.cuh sits at the FP32 ridge and is bound by the shared-memory crossbar: DESIGN.md
// §5.2) but too small to fill a tensor-core tile on its own.  One persistent CTA per SM keeps EIGHT problems ("slots")
// in flight; two slots share one 128 x 128 FP32 accumulator in TMEM, whose two 64 x 64 diagonal blocks are their
// Hessians (the operand rows 0..63 are the columns of slot A's Jacobian, rows 64..127 those of slot B's; the
// off-diagonal blocks are computed and ignored).  Warp-specialised:
//
//   loader   (1 warp)   32-row chunks of both problems' A with cp.async.bulk through an mbarrier ring, L2 prefetch ahead
//   t-warps  (2 warps)  lane = row: the canonical t = a_i . x chain, r_i, the row scale s_i, cost = sum r_i^2 in row order
//   columns  (8 warps)  thread = column j: J_ij = s_i a_ij, g_j += J_ij r_i and H_jj += J_ij^2 in FP32, then
//                       v = J_ij 2^e_j (a per-column power of two), hi = fp16(v), lo = fp16(v - hi), both stored
//                       K-major into the UMMA operand ring
//   MMA      (1 thread) tcgen05.mma.cta_group::1.kind::f16, M = N = 128, K = 16: hi hi' + hi lo' + lo hi' (22 significant
//                       bits, the accuracy class of FP32 sums) into the pair's TMEM accumulator
//   drains   (4 warps)  a warp can only read its own quarter of the TMEM lanes: warp k moves rows [32 k, 32 k + 32) of
//                       every accumulator (tcgen05.ld) straight into the slot's pivoted LDL^T layout in shared memory
//                       and into the slot's persistent copy of H_ (global, L2 resident)
//   solvers  (8 warps)  one per slot, two per scheduler: Build's tail, damping, Eigen's pivot order (from the FP32
//                       diagonal, before the accumulator is complete), a latency-optimised two-column LDL^T that carries
//                       the right-hand side along as an extra row, the back substitution, and the shared LM state
//                       machine (lm_state.cuh).  A pivot that is not positive sends the problem through the exact
//                       pivoted routine of the warp-per-problem family (wpp.cuh) on the persistent copy.
//
// The data pass of one pair overlaps the solves of the other three.  Parity: g, diag(H), cost and t are FP32 sums
// (cost and t in the oracle's canonical order), the off-diagonal of H comes from the tensor core, so the family is
// tolerance-held (1e-4 on x and the costs, iteration counts where the decisions clear FP32 noise), like the large-n
// family; TOB200_WPP_TC=0 / tob200_set_exact() keeps the bit-exact warp-per-problem kernel.
//
// Reference path replaced: diff/optimize_autodiff.h:151-157, solvers/lm.h:60-120, solvers/gn.h:150-171,
// math.h:232-240, optimizers/optimizer.h:243-539.
#pragma once

#include <cstdio>

#include "common.cuh"
#include "lm_state.cuh"
#include "tc.cuh"
#include "wpp.cuh"
#include "wtc_params.h"

namespace tob200 {

#ifndef TOB200_WTC_SWEEP_UNROLL
#define TOB200_WTC_SWEEP_UNROLL 1  // the LDL^T sweep's j loop (2 measured: see DESIGN.md)
#endif
constexpr int kWtcSweepUnroll = TOB200_WTC_SWEEP_UNROLL;

enum WtcVec {
  kVx = 0, kVlastdx, kVg, kVdg, kVdd, kVtemp, kVdxs, kVtb1, kVtb2, kVtb3, kVperm, kVinv, kVcs, kVci,
  kVgp0, kVgp1, kVdp0, kVdp1, kVmp0, kVmp1, kVmisc
};
static_assert(kVmisc + 1 == kWtcVecs, "vector count");
static_assert(kVtb1 == kVdxs + 1 && kVtb3 == kVdxs + 3, "the LDLT's T rows are dxs .. tb3");

enum WtcBar {
  kBRawFull = 0, kBRawEmpty = 5, kBRsFull = 10, kBOpFull = 15, kBOpEmpty = 18,
  kBAccFull = 21, kBFrontDone = 25, kBPairReady = 29, kBPermReady = 33, kBWReady = 41, kBCount = 49
};

// warp roles (24 warps, six per scheduler).  A warp reads the TMEM lanes [32 (warp % 4), +32): the drains are warps
// 0..3; the solvers are warps 4..11 (two per scheduler: their code is one long dependency chain).
#ifdef TOB200_WTC_TIMING
__device__ long long g_wtc_tm[32];
#define WTC_T0() long long wtc_t0__ = clock64()
#define WTC_T(k) do { const long long t__ = clock64(); if (blockIdx.x == 0 && lane == 0) g_wtc_tm[k] += t__ - wtc_t0__; wtc_t0__ = t__; } while (0)
#else
#define WTC_T0()
#define WTC_T(k)
#endif

constexpr int kWtcDrainWarps = 4, kWtcSolveWarp0 = 4, kWtcLoadWarp = 12, kWtcMmaWarp = 13, kWtcTWarp0 = 14, kWtcTWarp1 = 15,
              kWtcColWarp0 = 16, kWtcTWarp2 = 24, kWtcTWarp3 = 25;
__device__ __forceinline__ int wtc_col_warp(int warp) { return (warp >= kWtcColWarp0 && warp < kWtcColWarp0 + kWtcColWarps) ? warp - kWtcColWarp0 : -1; }
// t-warps: 14, 15 take the even chunks of side 0 / 1, 24, 25 the odd ones (-> side + 2 * parity, or -1)
__device__ __forceinline__ int wtc_t_warp(int warp) {
  return warp == kWtcTWarp0 ? 0 : (warp == kWtcTWarp1 ? 1 : ((warp >= kWtcTWarp2 && warp < kWtcTWarp2 + kWtcTWarps - 2) ? 2 + warp - kWtcTWarp2 : -1));
}

// mbarrier wait with an exponentially growing sleep between polls.  The waits of this kernel are long (a data pass, a
// solve) and there are twenty waiting warps: polled without a pause they took 28 % of all issued instructions, and with
// the suspend-time hint of try_wait (NANOSLEEP.SYNCS wakes on every barrier event of the CTA) still 40 % of them plus a
// third of the shared-memory pipe, which is the resource this kernel runs out of.  ns0: first sleep, doubled up to nsmax.
constexpr int kWtcColArrive = kWtcColWarps;  // column warps that work on one chunk (all of them: alternating sets measured slower)
#ifndef TOB200_WTC_SLEEP_DIV
#define TOB200_WTC_SLEEP_DIV 1
#endif
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity, unsigned ns0, unsigned nsmax = 0) {
  if (mbar_try_wait(bar, parity)) return;
  if (nsmax == 0) nsmax = 8 * ns0;
  ns0 /= TOB200_WTC_SLEEP_DIV;
  nsmax /= TOB200_WTC_SLEEP_DIV;
  unsigned ns = ns0;
  do {
    __nanosleep(ns);
    ns = ns * 2 > nsmax ? nsmax : ns * 2;
  } while (!mbar_try_wait(bar, parity));
}
__device__ __forceinline__ float wtc_rcp(float x) {  // MUFU.RCP: 1 ulp, the LDL^T below is tolerance-held anyway
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// packed FP32 pairs (Blackwell FMUL2 / FFMA2 / FADD2: two IEEE operations per issue slot)
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// four consecutive floats of shared memory as two packed pairs
__device__ __forceinline__ void lds_f2x2(uint32_t saddr, unsigned long long &a, unsigned long long &b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(saddr));
}
__device__ __forceinline__ uint32_t h2_absmax(uint32_t acc, uint32_t v) {  // per half: max(acc, |v|)
  const __half2 r = __hmax2(*reinterpret_cast<const __half2 *>(&acc), __habs2(*reinterpret_cast<const __half2 *>(&v)));
  return *reinterpret_cast<const uint32_t *>(&r);
}

// 2^e as a float, |e| <= 100
__device__ __forceinline__ float wtc_pow2(int e) { return __int_as_float((127 + e) << 23); }

// per-column power of two: base 2^e in [2^(target-1), 2^target)
__device__ __forceinline__ int wtc_scale_exp(float base, int target) {
  if (!(base > 0.f) || !(base < 3.0e38f)) return 0;
  int ex;
  frexpf(base, &ex);
  const int e = target - ex;
  return e < -100 ? -100 : (e > 100 ? 100 : e);
}

// ---- latency-optimised LDL^T of the already permuted, column-scaled system ------------------------------------------
// W: rows 0..n of pitch ldw (row n = right-hand side), lower triangle; on return the strict lower triangle holds L, row n
// holds z = D^-1 L^-1 b (the forward substitution rides along as one more row of the sweep), dvec holds D.  Left-looking,
// two columns per step; T (2 x 64 floats) takes the rows D_j L_{k+c,j} of the step, zero padded to a multiple of four so
// that the sweep has no tail loop (the never-written upper triangle of W is zeroed once per kernel).  The 2 x 2 pivot
// block is broadcast by three shuffles and factorised redundantly in every lane with MUFU.RCP instead of two divisions
// behind two dependent shuffles: the serial chain of a step is shuffle - rcp - fma - rcp - multiply.
// Measured in isolation (tools/cuda/ldlt_bench.cu, one warp, n = 50): 31 k cycles against 46 k for the bit-exact routine
// of wpp.cuh; a single warp retires one instruction per ~5 cycles here, so the instruction count of a step is what
// matters (a variant with fixed row ownership, T rows produced a step ahead and eight fma chains was SLOWER: 36 k).
// Returns false when a pivot is not positive (NaN included): the caller then runs the exact routine, which decides
// what Eigen would have decided (zero pivots, sign of D).
__device__ __forceinline__ bool wtc_ldlt_fast(float *W, int ldw, int n, float *dvec, float *T, int lane) {
  constexpr unsigned kFull = 0xffffffffu;
  bool bad = false;
  const int nr = n + 1;
  for (int k = 0; k < n; k += 2) {
    const bool two = k + 1 < n;
    const int k4 = (k + 3) & ~3;
    for (int j = lane; j < k4; j += 32) {
      const float d = j < k ? dvec[j] : 0.f;
      T[j] = j < k ? __fmul_rn(d, W[k * ldw + j]) : 0.f;
      T[kWtcNP + j] = (two && j < k) ? __fmul_rn(d, W[(k + 1) * ldw + j]) : 0.f;
    }
    __syncwarp();
    const int r0 = k + lane, r1 = r0 + 32;
    const bool h0 = r0 < nr, h1 = r1 < nr;
    float *w0 = W + (h0 ? r0 : k) * ldw, *w1 = W + (h1 ? r1 : k) * ldw;
    float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
#pragma unroll 2
    for (int j = 0; j < k4; j += 4) {
      const float4 a4 = *reinterpret_cast<const float4 *>(w0 + j);
      const float4 b4 = *reinterpret_cast<const float4 *>(w1 + j);
      const float4 t4 = *reinterpret_cast<const float4 *>(T + j);
      const float4 u4 = *reinterpret_cast<const float4 *>(T + kWtcNP + j);
      s00 = __fmaf_rn(a4.x, t4.x, s00); s01 = __fmaf_rn(a4.x, u4.x, s01); s10 = __fmaf_rn(b4.x, t4.x, s10); s11 = __fmaf_rn(b4.x, u4.x, s11);
      s00 = __fmaf_rn(a4.y, t4.y, s00); s01 = __fmaf_rn(a4.y, u4.y, s01); s10 = __fmaf_rn(b4.y, t4.y, s10); s11 = __fmaf_rn(b4.y, u4.y, s11);
      s00 = __fmaf_rn(a4.z, t4.z, s00); s01 = __fmaf_rn(a4.z, u4.z, s01); s10 = __fmaf_rn(b4.z, t4.z, s10); s11 = __fmaf_rn(b4.z, u4.z, s11);
      s00 = __fmaf_rn(a4.w, t4.w, s00); s01 = __fmaf_rn(a4.w, u4.w, s01); s10 = __fmaf_rn(b4.w, t4.w, s10); s11 = __fmaf_rn(b4.w, u4.w, s11);
    }
    const float e00 = __fsub_rn(w0[k], s00), e10 = __fsub_rn(w1[k], s10);
    const float e01 = __fsub_rn(two ? w0[k + 1] : 0.f, s01), e11 = __fsub_rn(two ? w1[k + 1] : 0.f, s11);
    // the pivot block [a b; b c] (rows k, k + 1 are lanes 0, 1)
    const float a = __shfl_sync(kFull, e00, 0), b = __shfl_sync(kFull, e00, 1), c = __shfl_sync(kFull, e01, 1);
    const float ra = wtc_rcp(a);
    const float lb = __fmul_rn(b, ra);
    const float c2 = __fmaf_rn(-lb, b, c);
    const float rc = wtc_rcp(c2);
    bad = bad || !(a > 0.f) || (two && !(c2 > 0.f));
    const float l00 = __fmul_rn(e00, ra), l10 = __fmul_rn(e10, ra);
    const float l01 = __fmul_rn(__fmaf_rn(-l00, b, e01), rc), l11 = __fmul_rn(__fmaf_rn(-l10, b, e11), rc);
    if (h0 && lane > 0) w0[k] = l00;
    if (h1) w1[k] = l10;
    if (two) {
      if (h0 && lane > 1) w0[k + 1] = l01;
      if (h1) w1[k + 1] = l11;
    }
    if (lane == 0) {
      dvec[k] = a;
      if (two) dvec[k + 1] = c2;
    }
    __syncwarp();
  }
  return !bad;
}

// The same factorisation, FOUR columns per step (13 steps at n = 50 instead of 25: the per-step overhead — T rows, two warp
// barriers, loop control — is what a lone warp pays for, tools/cuda/ldlt_bench.cu).  The 4 x 4 pivot block travels by ten
// shuffles and is factorised redundantly in every lane (four MUFU.RCP in the chain); k is a multiple of four, so a row's
// four new entries are one 16-byte load and one 16-byte store.  T: 4 x 64 floats.  Same result contract as wtc_ldlt_fast.
__device__ __forceinline__ bool wtc_ldlt_fast4(float *W, int ldw, int n, float *dvec, float *T, int lane) {
  constexpr unsigned kFull = 0xffffffffu;
  bool bad = false;
  const int nr = n + 1;
  for (int k = 0; k < n; k += 4) {
    const int ncol = n - k < 4 ? n - k : 4;  // columns of this step (the last step may have fewer)
    for (int j = lane; j < k; j += 32) {
      const float d = dvec[j];
#pragma unroll
      for (int c = 0; c < 4; ++c) T[c * kWtcNP + j] = c < ncol ? __fmul_rn(d, W[(k + c) * ldw + j]) : 0.f;
    }
    __syncwarp();
    const int r0 = k + lane, r1 = r0 + 32;
    const bool h0 = r0 < nr, h1 = r1 < nr;
    float *w0 = W + (h0 ? r0 : k) * ldw, *w1 = W + (h1 ? r1 : k) * ldw;
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll kWtcSweepUnroll
    for (int j = 0; j < k; j += 4) {
      const float4 a4 = *reinterpret_cast<const float4 *>(w0 + j), b4 = *reinterpret_cast<const float4 *>(w1 + j);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 t4 = *reinterpret_cast<const float4 *>(T + c * kWtcNP + j);
        s0[c] = __fmaf_rn(a4.x, t4.x, s0[c]); s1[c] = __fmaf_rn(b4.x, t4.x, s1[c]);
        s0[c] = __fmaf_rn(a4.y, t4.y, s0[c]); s1[c] = __fmaf_rn(b4.y, t4.y, s1[c]);
        s0[c] = __fmaf_rn(a4.z, t4.z, s0[c]); s1[c] = __fmaf_rn(b4.z, t4.z, s1[c]);
        s0[c] = __fmaf_rn(a4.w, t4.w, s0[c]); s1[c] = __fmaf_rn(b4.w, t4.w, s1[c]);
      }
    }
    const float4 a4 = *reinterpret_cast<const float4 *>(w0 + k), b4 = *reinterpret_cast<const float4 *>(w1 + k);
    float e0[4] = {a4.x - s0[0], a4.y - s0[1], a4.z - s0[2], a4.w - s0[3]};
    float e1[4] = {b4.x - s1[0], b4.y - s1[1], b4.z - s1[2], b4.w - s1[3]};
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c >= ncol) { e0[c] = 0.f; e1[c] = 0.f; }
    // pivot block P (lower triangle; rows k .. k + 3 are lanes 0 .. 3); missing columns become identity columns
    const float p00 = __shfl_sync(kFull, e0[0], 0);
    float p10 = __shfl_sync(kFull, e0[0], 1), p11 = __shfl_sync(kFull, e0[1], 1);
    float p20 = __shfl_sync(kFull, e0[0], 2), p21 = __shfl_sync(kFull, e0[1], 2), p22 = __shfl_sync(kFull, e0[2], 2);
    float p30 = __shfl_sync(kFull, e0[0], 3), p31 = __shfl_sync(kFull, e0[1], 3), p32 = __shfl_sync(kFull, e0[2], 3),
          p33 = __shfl_sync(kFull, e0[3], 3);
    if (ncol < 2) { p10 = 0.f; p11 = 1.f; }   // rows k + ncol .. are not pivot rows (row n is the right-hand side)
    if (ncol < 3) { p20 = 0.f; p21 = 0.f; p22 = 1.f; }
    if (ncol < 4) { p30 = 0.f; p31 = 0.f; p32 = 0.f; p33 = 1.f; }
    const float q0 = wtc_rcp(p00);
    const float l10 = __fmul_rn(p10, q0), l20 = __fmul_rn(p20, q0), l30 = __fmul_rn(p30, q0);
    const float d1 = __fmaf_rn(-l10, p10, p11);
    const float q1 = wtc_rcp(d1);
    const float m21 = __fmaf_rn(-l20, p10, p21), m31 = __fmaf_rn(-l30, p10, p31);
    const float l21 = __fmul_rn(m21, q1), l31 = __fmul_rn(m31, q1);
    const float d2 = __fmaf_rn(-l21, m21, __fmaf_rn(-l20, p20, p22));
    const float q2 = wtc_rcp(d2);
    const float m32 = __fmaf_rn(-l31, m21, __fmaf_rn(-l30, p20, p32));
    const float l32 = __fmul_rn(m32, q2);
    const float d3 = __fmaf_rn(-l32, m32, __fmaf_rn(-l31, m31, __fmaf_rn(-l30, p30, p33)));
    const float q3 = wtc_rcp(d3);
    bad = bad || !(p00 > 0.f) || !(d1 > 0.f) || !(d2 > 0.f) || !(d3 > 0.f);
    // my rows: L_rc = M_rc / D_c, M_rc = e_rc - sum_{c' < c} L_rc' M_{k+c, c'}
    float x0[4], x1[4];
    x0[0] = __fmul_rn(e0[0], q0);
    x1[0] = __fmul_rn(e1[0], q0);
    x0[1] = __fmul_rn(__fmaf_rn(-x0[0], p10, e0[1]), q1);
    x1[1] = __fmul_rn(__fmaf_rn(-x1[0], p10, e1[1]), q1);
    x0[2] = __fmul_rn(__fmaf_rn(-x0[1], m21, __fmaf_rn(-x0[0], p20, e0[2])), q2);
    x1[2] = __fmul_rn(__fmaf_rn(-x1[1], m21, __fmaf_rn(-x1[0], p20, e1[2])), q2);
    x0[3] = __fmul_rn(__fmaf_rn(-x0[2], m32, __fmaf_rn(-x0[1], m31, __fmaf_rn(-x0[0], p30, e0[3]))), q3);
    x1[3] = __fmul_rn(__fmaf_rn(-x1[2], m32, __fmaf_rn(-x1[1], m31, __fmaf_rn(-x1[0], p30, e1[3]))), q3);
    if (ncol == 4 && lane >= 4) {
      if (h0) *reinterpret_cast<float4 *>(w0 + k) = make_float4(x0[0], x0[1], x0[2], x0[3]);
    } else if (h0) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < ncol && lane > c) w0[k + c] = x0[c];
    }
    if (h1) {
      if (ncol == 4) {
        *reinterpret_cast<float4 *>(w1 + k) = make_float4(x1[0], x1[1], x1[2], x1[3]);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < ncol) w1[k + c] = x1[c];
      }
    }
    if (lane == 0) {
      dvec[k] = p00;
      if (ncol > 1) dvec[k + 1] = d1;
      if (ncol > 2) dvec[k + 2] = d2;
      if (ncol > 3) dvec[k + 3] = d3;
    }
    __syncwarp();
  }
  return !bad;
}

// x <- L^-T z (z = row n of W), two columns per step, then dx[perm[i]] = x_i * cs[perm[i]] (undoing the column scaling)
__device__ __forceinline__ void wtc_back_subst(const float *W, int ldw, int n, const int *perm, const float *cs, float *dx,
                                               int lane) {
  constexpr unsigned kFull = 0xffffffffu;
  float x0 = lane < n ? W[n * ldw + lane] : 0.f, x1 = lane + 32 < n ? W[n * ldw + lane + 32] : 0.f;
  for (int j = n - 1; j >= 1; j -= 2) {
    const float u = __shfl_sync(kFull, j < 32 ? x0 : x1, j & 31);               // x_j, final
    float v = __shfl_sync(kFull, j - 1 < 32 ? x0 : x1, (j - 1) & 31);           // x_{j-1} before its last update
    v = __fmaf_rn(-W[j * ldw + j - 1], u, v);
    if (lane == ((j - 1) & 31)) {
      if (j - 1 < 32) x0 = v;
      else x1 = v;
    }
    if (lane < j - 1) x0 = __fmaf_rn(-W[(j - 1) * ldw + lane], v, __fmaf_rn(-W[j * ldw + lane], u, x0));
    if (lane + 32 < j - 1) x1 = __fmaf_rn(-W[(j - 1) * ldw + lane + 32], v, __fmaf_rn(-W[j * ldw + lane + 32], u, x1));
  }
  if (lane < n) { const int o = perm[lane]; dx[o] = __fmul_rn(x0, cs[o]); }
  if (lane + 32 < n) { const int o = perm[lane + 32]; dx[o] = __fmul_rn(x1, cs[o]); }
  __syncwarp();
}

// The exact route: Eigen's pivoted LDL^T semantics (zero pivots, sign of D, D^+) on the UNSCALED damped H_ rebuilt from the
// slot's persistent copy.  Cold: cost-only passes, retries after a solver failure, systems that are not positive
// definite.  Returns false when the factorisation reports failure.
__device__ __noinline__ bool wtc_solve_exact(float *W, int ldw, int n, float *V, float *hp, int lane) {
  float *g = V + kVg * kWtcNP, *dd = V + kVdd * kWtcNP, *temp = V + kVtemp * kWtcNP, *dxs = V + kVdxs * kWtcNP;
  int *perm = reinterpret_cast<int *>(V + kVperm * kWtcNP), *inv = reinterpret_cast<int *>(V + kVinv * kWtcNP);
  const float *hps = hp + n * ldw;
  wpp_pivot_order(dd, n, perm, inv, lane);
  for (int e = lane; e < n * ldw; e += 32) {
    const int i = e / ldw, j = e - i * ldw;
    if (j <= i) {
      const float val = i == j ? dd[i] : __fmul_rn(__fmul_rn(__ldcg(hp + e), __ldcg(hps + i)), __ldcg(hps + j));
      const int a = inv[i], b = inv[j];
      W[(a > b ? a : b) * ldw + (a > b ? b : a)] = val;
    }
  }
  __syncwarp();
  for (int j = lane; j < n; j += 32) hp[j * ldw + j] = dd[j];
  __syncwarp();
  const bool ok = wpp_ldlt_factor<float>(W, ldw, n, temp, dxs, kWtcNP, lane);  // gn.h:150-156
  if (ok) {
    for (int j = lane; j < n; j += 32) temp[j] = -g[j];
    __syncwarp();
    wpp_ldlt_solve<float>(W, ldw, n, perm, temp, dxs, lane);
  }
  // the fast path relies on a zero upper triangle: the routine above never writes it
  return ok;
}

// damped diagonal of H_ (lm.h:108-117): from the undamped FP32 one after a rebuild, cumulative on the stale H_ otherwise
__device__ __forceinline__ void wtc_damp(const LmScalars<float> &s, const DevOptions<float> &o, bool pass_rebuilt, int n, int ldw,
                                         const float *dg, const float *hp, float *dd, int lane) {
  double sc;
  const bool damp = lm_damping_scale(s, o, pass_rebuilt, sc);
  for (int j = lane; j < n; j += 32) {
    const float base = pass_rebuilt ? dg[j] : __ldcg(hp + j * ldw + j);
    dd[j] = damp ? (float)((double)base * sc) : base;
  }
  __syncwarp();
}

// sum of squares of a shared vector: lane shares + a fixed butterfly (the bit-exact kernels sum sequentially: 50 dependent fmas
// of one lane; here the order only has to be deterministic)
__device__ __forceinline__ float wtc_sqnorm(const float *v, int n, int lane) {
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s = __fmaf_rn(v[j], v[j], s);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, off));
  return s;
}

// Eigen's pivot order of the damped diagonal (pos2orig in perm, orig2pos in inv), rank-count fast path of
// wpp_pivot_order with half the work per key: position = number of larger keys; ties and NaNs show up as two keys landing
// on one position (checked after the fact: perm[inv[i]] == i) or as a zero key, and go through the exact replay of wpp.cuh.
// keys: 64 words of scratch.
__device__ __forceinline__ void wtc_pivot_order(const float *dd, int n, int *perm, int *inv, uint32_t *keys, int lane) {
  const int i0 = lane, i1 = lane + 32;
  uint32_t k0 = 0, k1 = 0;  // 0: NaN or beyond n, otherwise 1 + bits(|d|) (monotone in |d|)
  if (i0 < n) { const float v = fabsf(dd[i0]); k0 = (v != v) ? 0u : __float_as_uint(v) + 1u; }
  if (i1 < n) { const float v = fabsf(dd[i1]); k1 = (v != v) ? 0u : __float_as_uint(v) + 1u; }
  keys[i0] = k0;
  keys[i1] = k1;
  __syncwarp();
  int gt0 = 0, gt1 = 0;
  const int n4 = (n + 3) & ~3;  // keys beyond n are 0: never greater
#pragma unroll 2
  for (int j = 0; j < n4; j += 4) {
    const uint4 kj = *reinterpret_cast<const uint4 *>(keys + j);
    gt0 += (kj.x > k0) + (kj.y > k0) + (kj.z > k0) + (kj.w > k0);
    gt1 += (kj.x > k1) + (kj.y > k1) + (kj.z > k1) + (kj.w > k1);
  }
  if (i0 < n) { perm[gt0] = i0; inv[i0] = gt0; }
  if (i1 < n) { perm[gt1] = i1; inv[i1] = gt1; }
  __syncwarp();
  const bool amb = (i0 < n && (k0 == 0u || perm[gt0] != i0)) || (i1 < n && (k1 == 0u || perm[gt1] != i1));
  if (__any_sync(0xffffffffu, amb)) wpp_pivot_order(dd, n, perm, inv, lane);
}

struct WtcSolverCtx {
  uint64_t *perm_ready, *w_ready;
  volatile int *cmd;  // this visit: bit 0 = the drains lay the accumulator out (else they only pass), bit 1 = they also
                      // write the slot's persistent copy of H_
  uint32_t sv;        // visits of this slot so far (parity of the two barriers above)
  bool force_hp;      // the next pass must keep a persistent copy (the last one turned out to need it and had none)
};

// Everything after the data pass of one problem (mirrors wpp_after_pass / lm_after_pass).  Called when the FP32 sums of
// the pass (g, diag, cost) are complete; the accumulator itself is waited for as late as possible.  Returns true when
// the pass has to be REPEATED with new column scales (an FP16 operand overflowed: nothing of the state was touched).
__device__ __forceinline__ bool wtc_after_pass(LmScalars<float> &s, const DevOptions<float> &o, int n, int nres, int ldw,
                                               float *V, float *W, float *hp, WtcSolverCtx &sx, bool pass_rebuilt, int lane,
                                               int debug) {
  using O = Ops<float>;
  float *xs = V + kVx * kWtcNP, *last_dx = V + kVlastdx * kWtcNP, *g = V + kVg * kWtcNP, *dg = V + kVdg * kWtcNP;
  float *dd = V + kVdd * kWtcNP, *dvec = V + kVtemp * kWtcNP, *dxs = V + kVdxs * kWtcNP;
  int *perm = reinterpret_cast<int *>(V + kVperm * kWtcNP), *inv = reinterpret_cast<int *>(V + kVinv * kWtcNP);
  float *cs = V + kVcs * kWtcNP, *ci = V + kVci * kWtcNP;
  const float cost_t = __fadd_rn(__fadd_rn(V[kVmisc * kWtcNP], V[kVmisc * kWtcNP + 1]), V[kVmisc * kWtcNP + 2]);  // the t-warps' shares
#ifdef TOB200_WTC_TIMING
  const bool tm_on = (threadIdx.x >> 5) == kWtcSolveWarp0;
#define WTC_TA(k) do { if (tm_on) WTC_T(k); } while (0)
  WTC_T0();
#else
#define WTC_TA(k)
#endif
  // hands the visit to the drain warps (lay out / pass) and waits until both have passed
  auto release_drains = [&](int lay_out) {
    __syncwarp();  // every lane's writes of W, perm, hp precede the release
    if (lane == 0) {
      *sx.cmd = lay_out;
      mbar_arrive(sx.perm_ready);
    }
    __syncwarp();
  };
  auto wait_drains = [&]() {
    mbar_wait_sleep(sx.w_ready, sx.sv & 1u, 200, 800);
    ++sx.sv;
  };

  int new_e[2] = {0, 0};
  bool new_ok[2] = {false, false};
  if (pass_rebuilt) {
    bool ovf = false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h;
      if (j < n) {
        // the two row halves of every stage were summed by different warps: fixed combine order
        g[j] = __fadd_rn(V[kVgp0 * kWtcNP + j], V[kVgp1 * kWtcNP + j]);
        dg[j] = __fadd_rn(V[kVdp0 * kWtcNP + j], V[kVdp1 * kWtcNP + j]);
        const float cms = fmaxf(V[kVmp0 * kWtcNP + j], V[kVmp1 * kWtcNP + j]);  // max_i |J_ij| 2^e_j (Inf: overflow)
        if (!(cms < 60000.f)) {
          // an operand of this column overflowed FP16: the true maximum is unknown, step the exponent down and repeat
          // the pass — unless the exponent is at its floor already (the data itself is Inf: let the solver see it)
          int ex;
          frexpf(cs[j], &ex);  // cs = 2^(ex - 1)
          if (ex - 1 > -100) {
            ovf = true;
            new_e[h] = ex - 1 - 8 < -100 ? -100 : ex - 1 - 8;
            new_ok[h] = true;
          }
        } else if (cms > 0.f) {
          new_e[h] = wtc_scale_exp(__fmul_rn(cms, ci[j]), 11);
          new_ok[h] = true;
        }
      }
    }
    if (__any_sync(0xffffffffu, ovf)) {
      release_drains(0);
      wait_drains();
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (new_ok[h]) {
          cs[lane + 32 * h] = wtc_pow2(new_e[h]);
          ci[lane + 32 * h] = wtc_pow2(-new_e[h]);
        }
      __syncwarp();
      return true;
    }
    __syncwarp();
  }

  double cost;
  bool built_ok = lm_normalize_cost(o, cost_t, nres, cost);
  if (pass_rebuilt) {
    s.num_builds++;
    if (built_ok) {
      if (o.grad_clipping != 0.f) {  // base.h:30-38
        for (int j = lane; j < n; j += 32) {
          float v = g[j];
          v = v < -o.grad_clipping ? -o.grad_clipping : v;
          v = v > o.grad_clipping ? o.grad_clipping : v;
          g[j] = v;
        }
        __syncwarp();
      }
      if (o.check_min_H_diag > 0.f) {  // lm.h:82-86
        bool low = false;
        for (int j = lane; j < n; j += 32) low = low || (O::abs(dg[j]) < o.check_min_H_diag);
        if (__any_sync(0xffffffffu, low)) built_ok = false;
      }
    }
  }
  WTC_TA(21);

  // The pivot order depends on the damped diagonal only: it is ready before the accumulator is.  The solver writes the
  // diagonal and the right-hand side row (column-scaled like the accumulator: H_s = S H S, b_s = -S g), the drains add
  // the off-diagonal entries at their pivoted positions.
  const bool fast = pass_rebuilt && built_ok;
  // Will a cost-only iteration possibly follow this one (optimizer.h:295)?  Only then — or when the last attempt found the
  // system not positive definite and had no copy to hand to the exact routine — does H_ have to outlive the accumulator:
  // the 10 KB per problem the drains would otherwise write every pass are 7 % of the kernel's DRAM traffic and 8 % of its
  // shared-memory-pipe wavefronts.
#ifdef TOB200_WTC_ALWAYS_PERSIST
  const bool persist = fast;
#else
  const bool persist = fast && (sx.force_hp || (!(cost - s.final_cost < 0.0) && !(s.flags & kFlagLastWasSuccess)));
#endif
  sx.force_hp = false;
  if (fast) {
    wtc_damp(s, o, true, n, ldw, dg, hp, dd, lane);
    // (re-using the previous pass's order when it still sorts the new diagonal was measured slower: the diagonal entries of a
    //  well-scaled problem are nearly equal, so their order changes with every step and the check is pure overhead)
    wtc_pivot_order(dd, n, perm, inv, reinterpret_cast<uint32_t *>(dxs), lane);  // (dxs is free until the factorisation)
    for (int j = lane; j < n; j += 32) {
      const int a = inv[j];
      const float c = cs[j];
      W[a * ldw + a] = __fmul_rn(__fmul_rn(dd[j], c), c);
      W[n * ldw + a] = __fmul_rn(-g[j], c);
      if (persist) {
        hp[j * ldw + j] = dd[j];
        hp[n * ldw + j] = ci[j];
      }
    }
  }
  release_drains(fast ? (persist ? 3 : 1) : 0);
  WTC_TA(22);
  wait_drains();
  WTC_TA(23);

  bool solver_failed = true, early_return = false;
  const uint8_t max_tries = lm_max_tries(o);
  for (int attempt = 0; s.num_consec_failures <= max_tries; ++attempt) {
    if (built_ok) {
      bool ok = false;
      if (fast && attempt == 0) {
        if (debug & 2) {  // timing experiment: no factorisation (results invalid)
          for (int j = lane; j < n; j += 32) dxs[j] = 0.f;
          __syncwarp();
          ok = true;
        } else {
          ok = wtc_ldlt_fast4(W, ldw, n, dvec, dxs, lane);
          WTC_TA(24);
          if (ok) wtc_back_subst(W, ldw, n, perm, cs, dxs, lane);
          WTC_TA(25);
        }
        if (!ok) {  // not positive definite (or NaN): let the exact routine decide; W's upper triangle must stay zero
          if (!persist) {
            // no copy of H_ was kept: repeat the data pass once with the copy requested (rare: singular / indefinite systems)
            for (int e = lane; e < (n + 1) * ldw; e += 32) W[e] = 0.f;
            __syncwarp();
            s.num_builds--;
            sx.force_hp = true;
            return true;
          }
          ok = wtc_solve_exact(W, ldw, n, V, hp, lane);
          for (int e = lane; e < (n + 1) * ldw; e += 32) W[e] = 0.f;
          __syncwarp();
        }
      } else {
        wtc_damp(s, o, pass_rebuilt, n, ldw, dg, hp, dd, lane);
        ok = wtc_solve_exact(W, ldw, n, V, hp, lane);
        for (int e = lane; e < (n + 1) * ldw; e += 32) W[e] = 0.f;
        __syncwarp();
      }
      if (ok) solver_failed = false;
    }
    if (!solver_failed) break;
    const int act = lm_on_solver_failure(s, o, cost, nres);
    if (act == kLmEarlyReturn) early_return = true;
    if (act != kLmRetry) break;
    if (attempt >= 100000) break;
  }

  double dx_norm2 = 0.0, grad_norm2 = 0.0;
  if (!solver_failed) {
    dx_norm2 = (double)wtc_sqnorm(dxs, n, lane);
    if (o.min_grad_norm2_f > 0.0f) grad_norm2 = (double)wtc_sqnorm(g, n, lane);
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
  // column scales of the next pass from this pass's column maxima of J
  if (pass_rebuilt) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (new_ok[h]) {
        cs[lane + 32 * h] = wtc_pow2(new_e[h]);
        ci[lane + 32 * h] = wtc_pow2(-new_e[h]);
      }
  }
  __syncwarp();
  WTC_TA(26);
  return false;
}

// tob200_build_solve_f32 on this family: one Build + Solve of a problem whose J and r were handed in (SolverLM::Build,
// solvers/lm.h:60-120, + SolverGN::Solve, gn.h:150-171): cost, g, the damped H_, dx.  Same machinery as the LM pass, no state.
// Returns true when the pass has to be repeated (FP16 overflow of a column, or a copy of H_ needed after all).
__device__ __forceinline__ bool wtc_build_solve_pass(const WtcParams &p, long long prob, int n, int ldw, float *V, float *W,
                                                     float *hp, WtcSolverCtx &sx, int lane) {
  float *g = V + kVg * kWtcNP, *dg = V + kVdg * kWtcNP, *dd = V + kVdd * kWtcNP, *dvec = V + kVtemp * kWtcNP;
  float *dxs = V + kVdxs * kWtcNP, *cs = V + kVcs * kWtcNP, *ci = V + kVci * kWtcNP;
  int *perm = reinterpret_cast<int *>(V + kVperm * kWtcNP), *inv = reinterpret_cast<int *>(V + kVinv * kWtcNP);
  const float cost_t = __fadd_rn(__fadd_rn(V[kVmisc * kWtcNP], V[kVmisc * kWtcNP + 1]), V[kVmisc * kWtcNP + 2]);
  auto release_drains = [&](int cmd) {
    __syncwarp();
    if (lane == 0) {
      *sx.cmd = cmd;
      mbar_arrive(sx.perm_ready);
    }
    __syncwarp();
  };
  auto wait_drains = [&]() {
    mbar_wait_sleep(sx.w_ready, sx.sv & 1u, 200, 800);
    ++sx.sv;
  };
  bool ovf = false;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
    if (j < n) {
      g[j] = __fadd_rn(V[kVgp0 * kWtcNP + j], V[kVgp1 * kWtcNP + j]);
      dg[j] = __fadd_rn(V[kVdp0 * kWtcNP + j], V[kVdp1 * kWtcNP + j]);
      const float cms = fmaxf(V[kVmp0 * kWtcNP + j], V[kVmp1 * kWtcNP + j]);
      if (!(cms < 60000.f)) {
        int ex;
        frexpf(cs[j], &ex);
        if (ex - 1 > -100) {
          ovf = true;
          const int e = ex - 1 - 8 < -100 ? -100 : ex - 1 - 8;
          cs[j] = wtc_pow2(e);
          ci[j] = wtc_pow2(-e);
        }
      }
    }
  }
  if (__any_sync(0xffffffffu, ovf)) {
    release_drains(0);
    wait_drains();
    return true;
  }
  __syncwarp();
  const float lam = p.lambda ? p.lambda[prob] : 0.f;
  const double sc = 1.0 + (double)lam;  // solvers/lm.h:108-117
  for (int j = lane; j < n; j += 32) dd[j] = lam > 0.f ? (float)((double)dg[j] * sc) : dg[j];
  __syncwarp();
  if (lane == 0) p.cost_out[prob] = (double)cost_t;
  if (p.g_out)
    for (int j = lane; j < n; j += 32) p.g_out[(size_t)prob * n + j] = g[j];
  const bool persist = sx.force_hp || p.H_out != nullptr;
  sx.force_hp = false;
  wtc_pivot_order(dd, n, perm, inv, reinterpret_cast<uint32_t *>(dxs), lane);
  for (int j = lane; j < n; j += 32) {
    const int a = inv[j];
    const float c = cs[j];
    W[a * ldw + a] = __fmul_rn(__fmul_rn(dd[j], c), c);
    W[n * ldw + a] = __fmul_rn(-g[j], c);
    if (persist) {
      hp[j * ldw + j] = dd[j];
      hp[n * ldw + j] = ci[j];
    }
  }
  release_drains(persist ? 3 : 1);
  wait_drains();
  if (p.H_out) {  // damped H_, full symmetric, from the persistent copy (column-scaled: undone here)
    float *Ho = p.H_out + (size_t)prob * n * n;
    for (int e = lane; e < n * n; e += 32) {
      const int i = e / n, j = e - i * n;
      const int a = i > j ? i : j, b = i > j ? j : i;
      Ho[e] = i == j ? dd[i] : __fmul_rn(__fmul_rn(__ldcg(hp + a * ldw + b), ci[a]), ci[b]);
    }
  }
  bool ok = wtc_ldlt_fast4(W, ldw, n, dvec, dxs, lane);
  if (ok) {
    wtc_back_subst(W, ldw, n, perm, cs, dxs, lane);
  } else {
    if (!persist) {  // no copy of H_ for the exact routine: repeat the pass with the copy requested
      for (int e = lane; e < (n + 1) * ldw; e += 32) W[e] = 0.f;
      __syncwarp();
      sx.force_hp = true;
      return true;
    }
    ok = wtc_solve_exact(W, ldw, n, V, hp, lane);
    for (int e = lane; e < (n + 1) * ldw; e += 32) W[e] = 0.f;
    __syncwarp();
  }
  if (ok)
    for (int j = lane; j < n; j += 32) p.dx[(size_t)prob * n + j] = dxs[j];
  if (lane == 0) p.status[prob] = ok ? 0 : 1;
  __syncwarp();
  return false;
}

// kMode 0: tob200_lm_run_f32 (the whole LM loop of the polynomial family), 1: tob200_build_solve_f32 (materialised J, r)
template <int kMode>
__global__ void __launch_bounds__(kWtcThreads, 1) wtc_lm_run_kernel(const __grid_constant__ WtcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const WtcSmem &L = p.L;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.bars + kBCount * 8);
  volatile long long *slot_prob = reinterpret_cast<volatile long long *>(smem + L.desc);  // [2][kWtcSlots]
  volatile int *slot_cmd = reinterpret_cast<volatile int *>(smem + L.desc + 128);         // [kWtcSlots]
  const int n = p.n, m = p.m;
  const int nchunks = (m + kWtcRows - 1) / kWtcRows;
  const uint32_t R = (uint32_t)L.raw_stages, S = (uint32_t)L.op_stages;
  float *rsr = reinterpret_cast<float *>(smem + L.rsr);  // [2: s, r][raw stage][side][32 rows]
  constexpr int kRsrHalf = kWtcMaxRawStages * 2 * kWtcRows;
  const int ldw = wtc_ldw(n);

  if (tid == 0) {
    for (int s = 0; s < kWtcMaxRawStages; ++s) {
      mbar_init(&bars[kBRawFull + s], 1);
      mbar_init(&bars[kBRawEmpty + s], kWtcColArrive);
      mbar_init(&bars[kBRsFull + s], 2);
    }
    for (int s = 0; s < kWtcMaxOpStages; ++s) {
      mbar_init(&bars[kBOpFull + s], kWtcColArrive);
      mbar_init(&bars[kBOpEmpty + s], 1);
    }
    for (int q = 0; q < kWtcPairs; ++q) {
      mbar_init(&bars[kBAccFull + q], 1);
      mbar_init(&bars[kBFrontDone + q], kWtcColWarps + kWtcTWarps);
      mbar_init(&bars[kBPairReady + q], 2);
    }
    for (int s = 0; s < kWtcSlots; ++s) {
      mbar_init(&bars[kBPermReady + s], 1);
      mbar_init(&bars[kBWReady + s], 2);
    }
    mbar_fence_init();
  }
  if (warp == 0) {  // the whole TMEM of this SM: four 128 x 128 FP32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool is_drain = warp < kWtcDrainWarps;
  const bool is_solver = warp >= kWtcSolveWarp0 && warp < kWtcSolveWarp0 + kWtcSlots;
  const int cw = wtc_col_warp(warp);
  const int tw = wtc_t_warp(warp);
  const bool is_front = warp == kWtcLoadWarp || warp == kWtcMmaWarp || tw >= 0 || cw >= 0;

  if (is_solver) {
    // ===================== solver: one warp per slot =====================
    const int slot = warp - kWtcSolveWarp0, q = slot >> 1;
    float *V = reinterpret_cast<float *>(smem + L.vec + (size_t)slot * L.vec_stride);
    float *W = reinterpret_cast<float *>(smem + L.w + (size_t)slot * L.w_stride);
    float *hp = p.hpersist + ((size_t)blockIdx.x * kWtcSlots + slot) * (size_t)wtc_hp_floats(n);
    const bool is_lm = p.opt.solver_type == 0;
    LmScalars<float> s;
    s.reset_scalars(p.opt);
    long long prob = -1;
    WtcSolverCtx sx;
    sx.perm_ready = &bars[kBPermReady + slot];
    sx.w_ready = &bars[kBWReady + slot];
    sx.cmd = &slot_cmd[slot];
    sx.sv = 0;
    sx.force_hp = false;
    for (int e = lane; e < (n + 1) * ldw; e += 32) W[e] = 0.f;  // the fast LDL^T reads (and multiplies by zero) above the diagonal

    // the first chunks of the problem's next pass on their way into L2 while the other pairs stream
    auto prefetch_head = [&]() {
      if (lane < 4 && lane < nchunks) {
        const int row0 = lane * kWtcRows;
        const int rows = (m - row0 < kWtcRows) ? (m - row0) : kWtcRows;
        tma_prefetch_l2(p.A + ((size_t)prob * m + row0) * n, (uint32_t)rows * (uint32_t)n * 4u);
      }
    };
    auto fetch = [&]() {
      unsigned long long t = 0;
      if (lane == 0) t = atomicAdd(p.counter, 1ull);
      t = __shfl_sync(0xffffffffu, t, 0);
      if ((int64_t)t >= p.B) {
        prob = -1;
        return;
      }
      prob = (long long)t;
        prefetch_head();
      for (int j = lane; j < kWtcNP; j += 32) {
        V[kVx * kWtcNP + j] = (j < n && p.x) ? p.x[(size_t)prob * n + j] : 0.f;
        V[kVlastdx * kWtcNP + j] = 0.f;
      }
      s.reset_scalars(p.opt);
      // first estimate of the column scales: max |a_ij| over the first rows (s_i is unknown yet: 2^7 of headroom;
      // an overflow is detected after the pass and the pass repeated with smaller scales)
      const float *Ap = p.A + (size_t)prob * m * n;
      const int re = m < 32 ? m : 32;
      float m0 = 0.f, m1 = 0.f;
      for (int i = 0; i < re; ++i) {
        if (lane < n) m0 = fmaxf(m0, fabsf(Ap[(size_t)i * n + lane]));
        if (lane + 32 < n) m1 = fmaxf(m1, fabsf(Ap[(size_t)i * n + lane + 32]));
      }
      float gm = fmaxf(m0, m1);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, off));
      const int e0 = lane < n ? wtc_scale_exp(m0 > 0.f ? m0 : gm, 9) : 0;
      const int e1 = lane + 32 < n ? wtc_scale_exp(m1 > 0.f ? m1 : gm, 9) : 0;
      V[kVcs * kWtcNP + lane] = wtc_pow2(e0);
      V[kVci * kWtcNP + lane] = wtc_pow2(-e0);
      V[kVcs * kWtcNP + lane + 32] = wtc_pow2(e1);
      V[kVci * kWtcNP + lane + 32] = wtc_pow2(-e1);
    };
    fetch();
    uint32_t v = 0;
    WTC_T0();
    for (;;) {
      if (lane == 0) slot_prob[(v & 1u) * kWtcSlots + slot] = prob;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[kBPairReady + q]);
      mbar_wait_sleep(&bars[kBPairReady + q], v & 1u, 200, 1600);
      const long long pa = slot_prob[(v & 1u) * kWtcSlots + 2 * q], pb = slot_prob[(v & 1u) * kWtcSlots + 2 * q + 1];
      if (pa < 0 && pb < 0) break;
      if (slot == 0) WTC_T(0);  // waiting for the pair to be ready (the partner's solve)
      if (prob >= 0) {
        mbar_wait_sleep(&bars[kBFrontDone + q], v & 1u, 400, 3200);
        if (slot == 0) WTC_T(1);  // waiting for the data pass
        if constexpr (kMode == 1) {  // tob200_build_solve: one pass per problem
          const bool again = wtc_build_solve_pass(p, prob, n, ldw, V, W, hp, sx, lane);
          if (!again) fetch();
          ++v;
          continue;
        }
        const bool do_rebuild = !is_lm || s.rebuild();
        const bool redo = wtc_after_pass(s, p.opt, n, m, ldw, V, W, hp, sx, do_rebuild, lane, p.debug);
        if (slot == 0) WTC_T(2);  // after-pass
        if (!redo && s.done()) {
          for (int j = lane; j < n; j += 32) p.x[(size_t)prob * n + j] = V[kVx * kWtcNP + j];
          if (lane == 0) lm_write_result(s, &p.results[prob]);
          __syncwarp();
          fetch();
        } else {
          prefetch_head();
        }
        if (slot == 0) WTC_T(3);  // write-back + fetch
      }
      ++v;
    }
  } else if (is_drain) {
    // ===================== drain: TMEM lanes [32 warp, +32) of every accumulator =====================
    // warp k: side b = k / 2 (accumulator rows 64 b ..), rows [32 hh, 32 hh + 32) of the slot's H, hh = k % 2.  Only the
    // upper triangle (col > row) is moved: to W(max, min) of the pivoted positions and to the persistent copy
    // hp(col, row) (coalesced: lane = row).  The values stay column-scaled; the solver keeps the scale vector beside them.
    const int b = warp >> 1, hh = warp & 1;
    uint32_t vpar = 0, fin = 0, svm = 0;  // svm: parity of each slot's visit count (bit q)
    for (int q = 0; fin != 0xFu; q = (q + 1) & 3) {
      if ((fin >> q) & 1u) continue;
      const uint32_t par = (vpar >> q) & 1u;
      mbar_wait_sleep(&bars[kBPairReady + q], par, 200, 1600);
      const long long pa = slot_prob[par * kWtcSlots + 2 * q], pb = slot_prob[par * kWtcSlots + 2 * q + 1];
      if (pa < 0 && pb < 0) {
        fin |= 1u << q;
        continue;
      }
      vpar ^= 1u << q;
      const int slot = 2 * q + b;
      if ((b ? pb : pa) < 0) continue;
      mbar_wait_sleep(&bars[kBPermReady + slot], (svm >> q) & 1u, 400, 1600);
      svm ^= 1u << q;
      const int lay_out = slot_cmd[slot];
      mbar_wait_sleep(&bars[kBAccFull + q], par, 100, 800);
      tc_fence_after();
      if ((lay_out & 1) && 32 * hh < n) {
        const bool keep = (lay_out & 2) != 0;
        const float *V = reinterpret_cast<const float *>(smem + L.vec + (size_t)slot * L.vec_stride);
        const int *inv = reinterpret_cast<const int *>(V + kVinv * kWtcNP);
        float *W = reinterpret_cast<float *>(smem + L.w + (size_t)slot * L.w_stride);
        float *hp = p.hpersist + ((size_t)blockIdx.x * kWtcSlots + slot) * (size_t)wtc_hp_floats(n);
        const int j = 32 * hh + lane;  // my row of H
        const int aj = j < n ? inv[j] : 0;
        const uint32_t taddr = tmem_base + ((uint32_t)(64 * b + 32 * hh) << 16) + (uint32_t)(128 * q + 64 * b);
#pragma unroll 1
        for (int cb = 32 * hh; cb < n; cb += 32) {
          uint32_t v[32];
          tc_ld32(taddr + (uint32_t)cb, v);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int col = cb + c;
            if (col > j && col < n) {
              const int bb = inv[col];
              W[(aj > bb ? aj : bb) * ldw + (aj > bb ? bb : aj)] = __uint_as_float(v[c]);
            }
          }
          if (keep) {  // (its own loop: a predicate on every store of the loop above cost 2 % of the kernel)
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int col = cb + c;
              if (col > j && col < n) hp[col * ldw + j] = __uint_as_float(v[c]);
            }
          }
        }
      }
      tc_fence_before();  // my tcgen05.ld of the accumulator precede the pair's next MMAs
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[kBWReady + slot]);
    }
  } else if (is_front) {
    // ===================== the data pass: loader, t-warps, column warps, MMA issuer =====================
    const uint32_t ops_u32 = smem_u32(smem + L.ops);
    uint32_t vpar = 0, fin = 0;
    uint32_t st = 0, ph = 0, os = 0, oph = 0;  // raw ring stage / phase, operand ring stage / phase
    WTC_T0();
    const int tmk = warp == kWtcLoadWarp ? 4 : (warp == kWtcMmaWarp ? 8 : (tw == 0 ? 12 : (cw == 0 ? 16 : 28)));
    (void)tmk;
    for (int q = 0; fin != 0xFu; q = (q + 1) & 3) {
      if ((fin >> q) & 1u) continue;
      const uint32_t par = (vpar >> q) & 1u;
      mbar_wait_sleep(&bars[kBPairReady + q], par, 200, 1600);
      WTC_T(tmk);  // waiting for a ready pair
      const long long pa = slot_prob[par * kWtcSlots + 2 * q], pb = slot_prob[par * kWtcSlots + 2 * q + 1];
      if (pa < 0 && pb < 0) {
        fin |= 1u << q;
        continue;
      }
      vpar ^= 1u << q;

      if (warp == kWtcLoadWarp) {
        for (int c = 0; c < nchunks; ++c) {
          const int row0 = c * kWtcRows;
          const int rows = (m - row0 < kWtcRows) ? (m - row0) : kWtcRows;
          const uint32_t bytes = (uint32_t)rows * (uint32_t)n * 4u;
          if (lane == 0) {
            mbar_wait_sleep(&bars[kBRawEmpty + st], ph ^ 1u, 100, 400);
            WTC_T(5);
            fence_proxy_async();  // the column warps' generic reads of this stage precede the async writes
            // (a run without these copies - stale stages, results invalid - is no faster: 17.0 against 17.2 M it/s)
            mbar_expect_tx(&bars[kBRawFull + st], bytes * (uint32_t)((pa >= 0) + (pb >= 0)));
            unsigned char *dst = smem + L.raw + (size_t)st * L.raw_stage;
            if (pa >= 0) tma_bulk_g2s(dst, p.A + ((size_t)pa * m + row0) * n, bytes, &bars[kBRawFull + st]);
            if (pb >= 0) tma_bulk_g2s(dst + L.raw_side, p.A + ((size_t)pb * m + row0) * n, bytes, &bars[kBRawFull + st]);
          } else if (lane <= 2) {  // L2 prefetch of the same rows p.prefetch stages ahead
            const long long pp = lane == 1 ? pa : pb;
            const int cc = c + p.prefetch;
            if (pp >= 0 && cc < nchunks) {
              const int r0 = cc * kWtcRows;
              const int rr = (m - r0 < kWtcRows) ? (m - r0) : kWtcRows;
              tma_prefetch_l2(p.A + ((size_t)pp * m + r0) * n, (uint32_t)rr * (uint32_t)n * 4u);
            }
          }
          __syncwarp();
          WTC_T(6);
          if (++st == R) { st = 0; ph ^= 1u; }
        }
      } else if (tw >= 0) {
        // ---- lane = row: t = a_i . x, r_i, s_i, my share of the cost; this warp takes every other chunk of its side ----
        const int b = tw & 1, tpar = tw >> 1;
        const long long prob = b ? pb : pa;
        const bool active = prob >= 0;
        const int slot = 2 * q + b;
        float *V = reinterpret_cast<float *>(smem + L.vec + (size_t)slot * L.vec_stride);
        const float *xs = V + kVx * kWtcNP;
        const float *yp = p.y + (size_t)(active ? prob : 0) * m;
        float cost = 0.f;  // my rows' share of sum r_i^2
        constexpr int kTStride = kWtcTWarps / 2;  // this warp takes every kTStride-th chunk of its side
        float ynext = (active && tpar * kWtcRows + lane < m) ? yp[tpar * kWtcRows + lane] : 0.f;
        for (int c = 0; c < nchunks; ++c) {
          if (kTStride > 1 && (c % kTStride) != tpar) {  // another warp's chunk
            if (++st == R) { st = 0; ph ^= 1u; }
            continue;
          }
          const int row0 = c * kWtcRows;
          const int rows = (m - row0 < kWtcRows) ? (m - row0) : kWtcRows;
          const float ycur = ynext;
          ynext = (active && row0 + kTStride * kWtcRows + lane < m) ? yp[row0 + kTStride * kWtcRows + lane] : 0.f;
          mbar_wait_sleep(&bars[kBRawFull + st], ph, 100, 400);
          if (tw == 0) WTC_T(13);
          float *arow = reinterpret_cast<float *>(smem + L.raw + (size_t)st * L.raw_stage + (size_t)b * L.raw_side) + lane * n;
          float ri = 0.f, sc = 0.f;
          if (kMode == 1 && active && lane < rows) {  // materialised blocks: the row IS the Jacobian row, y the residual: the row IS the Jacobian row, y the residual
            ri = ycur;
            sc = 1.f;
            cost = __fmaf_rn(ri, ri, cost);
          } else if (active && lane < rows) {
            // t = a_i . x as four interleaved partial sums (fixed combine order): the chain is the latency of this warp.
            // (x held in registers - all of it, or the first 32 entries - was measured SLOWER: 2.0 k cycles per chunk
            // against 1.4 k; the unrolled predicated code and its spills cost more than the broadcast loads)
            float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
            if ((n & 1) == 0) {
              const float2 *a2 = reinterpret_cast<const float2 *>(arow);
              const float2 *x2 = reinterpret_cast<const float2 *>(xs);
              const int h = n / 2;
              int j = 0;
#pragma unroll 4
              for (; j + 2 <= h; j += 2) {
                const float2 av0 = a2[j], av1 = a2[j + 1], xv0 = x2[j], xv1 = x2[j + 1];
                t0 = __fmaf_rn(av0.x, xv0.x, t0);
                t1 = __fmaf_rn(av0.y, xv0.y, t1);
                t2 = __fmaf_rn(av1.x, xv1.x, t2);
                t3 = __fmaf_rn(av1.y, xv1.y, t3);
              }
              if (j < h) {
                const float2 av0 = a2[j], xv0 = x2[j];
                t0 = __fmaf_rn(av0.x, xv0.x, t0);
                t1 = __fmaf_rn(av0.y, xv0.y, t1);
              }
            } else {
              int j = 0;
#pragma unroll 2
              for (; j + 4 <= n; j += 4) {
                t0 = __fmaf_rn(arow[j], xs[j], t0);
                t1 = __fmaf_rn(arow[j + 1], xs[j + 1], t1);
                t2 = __fmaf_rn(arow[j + 2], xs[j + 2], t2);
                t3 = __fmaf_rn(arow[j + 3], xs[j + 3], t3);
              }
              for (; j < n; ++j) t0 = __fmaf_rn(arow[j], xs[j], t0);
            }
            const float t = __fadd_rn(__fadd_rn(t0, t1), __fadd_rn(t2, t3));
            const float tt = __fmul_rn(t, t);
            ri = __fmaf_rn(t, __fmaf_rn(p.alpha, tt, 1.f), -ycur);
            sc = __fmaf_rn(p.alpha3, tt, 1.f);
            cost = __fmaf_rn(ri, ri, cost);
          } else if (lane >= rows) {
            // a partial last chunk: the rows the bulk copy did not write hold another chunk's data: clear them, so that
            // the column warps need no row guard (s_i = 0 alone would let a NaN through)
            for (int j = 0; j < n; ++j) arow[j] = 0.f;
          }
          float *ssp = rsr + ((int)st * 2 + b) * kWtcRows, *srp = ssp + kRsrHalf;
          ssp[lane] = sc;
          srp[lane] = ri;
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[kBRsFull + st]);
          if (tw == 0) WTC_T(14);
          if (++st == R) { st = 0; ph ^= 1u; }
        }
        // cost: the lanes' shares combined by a fixed butterfly
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) cost = __fadd_rn(cost, __shfl_xor_sync(0xffffffffu, cost, off));
        if (lane == 0) {
          V[kVmisc * kWtcNP + tpar] = cost;  // the solver adds the two warps' shares
          if (kTStride == 1) V[kVmisc * kWtcNP + 1] = 0.f;
          if (kTStride < 3) V[kVmisc * kWtcNP + 2 + tpar] = 0.f;
          mbar_arrive(&bars[kBFrontDone + q]);
        }
        __syncwarp();
      } else if (cw >= 0) {
        // ---- thread = operand row (slot side, column j): FP32 sums of g_j and H_jj, FP16 hi / lo operands ----
        // Warp (grp, quarter): rows [16 grp, 16 grp + 16) of every stage, operand rows [32 quarter, +32).  Two rows at
        // a time as packed FP32 pairs.  Pad columns (j >= n) and idle slots read column n - 1 / stale bytes and are
        // multiplied by zero: whatever they produce stays inside accumulator rows / columns nobody reads.
        const int grp = cw >> 2;
        const int rr = 32 * (cw & 3) + lane, b = rr >> 6, j = rr & 63;
        const long long prob = b ? pb : pa;
        const int slot = 2 * q + b;
        float *V = reinterpret_cast<float *>(smem + L.vec + (size_t)slot * L.vec_stride);
        const bool colok = prob >= 0 && j < n;
        const float fs = colok ? V[kVcs * kWtcNP + j] : 0.f;
        const unsigned long long fs2 = f2_pack(fs, fs);
        unsigned long long g2 = 0ull, d2 = 0ull;  // (even rows, odd rows) partial sums
        uint32_t mxh = 0u;                        // max |hi| of my column, per half
        const uint32_t n4 = (uint32_t)n * 4u;
        const uint32_t off0 = (uint32_t)(2 * grp) * 2048u + (uint32_t)(rr >> 3) * 128u + (uint32_t)(rr & 7) * 16u;
        const uint32_t raw0 = smem_u32(smem + L.raw) + (uint32_t)b * L.raw_side + (uint32_t)((j < n ? j : n - 1) + 16 * grp * n) * 4u;
        const uint32_t rs0 = smem_u32(rsr) + (uint32_t)(b * kWtcRows + 16 * grp) * 4u;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait_sleep(&bars[kBRsFull + st], ph, 100, 400);
          mbar_wait(&bars[kBRawFull + st], ph);
          const uint32_t sa = raw0 + st * L.raw_stage;
          if (cw == 0) WTC_T(17);
          const uint32_t ss = rs0 + st * (uint32_t)(2 * kWtcRows * 4), sr = ss + (uint32_t)kRsrHalf * 4u;
          uint32_t hi[2][4], lo[2][4];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            unsigned long long sc2[4], r2[4];
            lds_f2x2(ss + (uint32_t)h * 32u, sc2[0], sc2[1]);
            lds_f2x2(ss + (uint32_t)h * 32u + 16u, sc2[2], sc2[3]);
            lds_f2x2(sr + (uint32_t)h * 32u, r2[0], r2[1]);
            lds_f2x2(sr + (uint32_t)h * 32u + 16u, r2[2], r2[3]);
            float a[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) a[t] = lds_f32(sa + (uint32_t)(8 * h + t) * n4);
#pragma unroll
            for (int t2 = 0; t2 < 4; ++t2) {
              const unsigned long long jv = f2_mul(f2_pack(a[2 * t2], a[2 * t2 + 1]), sc2[t2]);  // J_ij = s_i a_ij
              g2 = f2_fma(jv, r2[t2], g2);
              d2 = f2_fma(jv, jv, d2);
              const unsigned long long vv = f2_mul(jv, fs2);  // * 2^e_j (exact)
              float v0, v1;
              f2_unpack(vv, v0, v1);
              const uint32_t hh = lg_pack_h2(v0, v1);
              const unsigned long long ll = f2_sub(vv, f2_pack(lg_h_lo(hh), lg_h_hi(hh)));  // exact
              float l0, l1;
              f2_unpack(ll, l0, l1);
              hi[h][t2] = hh;
              lo[h][t2] = lg_pack_h2(l0, l1);
              mxh = h2_absmax(mxh, hh);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[kBRawEmpty + st]);
          if (cw == 0) WTC_T(18);
          mbar_wait_sleep(&bars[kBOpEmpty + os], oph ^ 1u, 100, 400);
          if (cw == 0) WTC_T(19);
          const uint32_t sb = ops_u32 + os * (uint32_t)kWtcOpStageBytes + off0;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            sts_v4u(sb + (uint32_t)h * 2048u, hi[h][0], hi[h][1], hi[h][2], hi[h][3]);
            sts_v4u(sb + (uint32_t)(kWtcOpStageBytes / 2) + (uint32_t)h * 2048u, lo[h][0], lo[h][1], lo[h][2], lo[h][3]);
          }
          fence_proxy_async();  // generic-proxy stores -> the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[kBOpFull + os]);
          if (cw == 0) WTC_T(20);
          if (++st == R) { st = 0; ph ^= 1u; }
          if (++os == S) { os = 0; oph ^= 1u; }
        }
        {
          float ge, go, de, dO;
          f2_unpack(g2, ge, go);
          f2_unpack(d2, de, dO);
          V[(kVgp0 + grp) * kWtcNP + j] = __fadd_rn(ge, go);
          V[(kVdp0 + grp) * kWtcNP + j] = __fadd_rn(de, dO);
          V[(kVmp0 + grp) * kWtcNP + j] = fmaxf(lg_h_lo(mxh), lg_h_hi(mxh));  // max_i |J_ij| 2^e_j as the FP16 operand saw it
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[kBFrontDone + q]);
      } else {
        // ---- MMA issuer: one thread ----
        if (lane == 0) {
          tc_fence_after();
          const uint32_t idesc = tc_idesc_f16(128);
          const uint32_t d = tmem_base + (uint32_t)(128 * q);
          for (int c = 0; c < nchunks; ++c) {
            mbar_wait_sleep(&bars[kBOpFull + os], oph, 100, 400);
            WTC_T(9);
            tc_fence_after();
            const uint32_t sb0 = ops_u32 + os * (uint32_t)kWtcOpStageBytes;
            if (!(p.debug & 1)) {
#pragma unroll
              for (int kk = 0; kk < kWtcRows / 16; ++kk) {
                const uint32_t sb = sb0 + (uint32_t)(2 * kk) * 2048u;  // this K step: 8-row chunks 2 kk, 2 kk + 1
                const uint64_t d_hi = tc_desc_k_major(sb, 2048u, 128u);
                const uint64_t d_lo = tc_desc_k_major(sb + (uint32_t)(kWtcOpStageBytes / 2), 2048u, 128u);
                tc_mma_f16(d, d_hi, d_hi, idesc, (c > 0 || kk > 0) ? 1u : 0u);
                tc_mma_f16(d, d_hi, d_lo, idesc, 1u);
                tc_mma_f16(d, d_lo, d_hi, idesc, 1u);
              }
            }
            tc_commit(&bars[kBOpEmpty + os]);  // arrives when the MMAs above have read the stage
            if (c == nchunks - 1) tc_commit(&bars[kBAccFull + q]);
            WTC_T(10);
            if (++os == S) { os = 0; oph ^= 1u; }
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef TOB200_WTC_TIMING
  if (blockIdx.x == 0 && tid == 0) {
    printf("wtc block 0 kcycles: solver0 wait-pair %lld wait-data %lld after-pass %lld fetch %lld | loader wait-pair %lld wait-empty %lld "
           "issue %lld | mma wait-pair %lld wait-full %lld issue %lld | twarp wait-pair %lld wait-raw %lld work %lld | col wait-pair %lld "
           "wait-rs %lld compute %lld wait-op %lld store %lld\n",
           g_wtc_tm[0] / 1000, g_wtc_tm[1] / 1000, g_wtc_tm[2] / 1000, g_wtc_tm[3] / 1000, g_wtc_tm[4] / 1000, g_wtc_tm[5] / 1000,
           g_wtc_tm[6] / 1000, g_wtc_tm[8] / 1000, g_wtc_tm[9] / 1000, g_wtc_tm[10] / 1000, g_wtc_tm[12] / 1000, g_wtc_tm[13] / 1000,
           g_wtc_tm[14] / 1000, g_wtc_tm[16] / 1000, g_wtc_tm[17] / 1000, g_wtc_tm[18] / 1000, g_wtc_tm[19] / 1000, g_wtc_tm[20] / 1000);
    printf("   after-pass of solver 0: gather+damp %lld pivot order %lld layout %lld factor %lld solve %lld state %lld\n", g_wtc_tm[21] / 1000,
           g_wtc_tm[22] / 1000, g_wtc_tm[23] / 1000, g_wtc_tm[24] / 1000, g_wtc_tm[25] / 1000, g_wtc_tm[26] / 1000);
    for (int k = 0; k < 32; ++k) g_wtc_tm[k] = 0;
  }
#endif
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

}  // namespace tob200
