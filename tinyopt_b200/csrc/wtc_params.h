// wtc_params.h — parameter block, launch geometry and shared-memory plan of the mid-n tensor-core
// family (kernel: wtc.cuh; launcher: wtc_kernels.cu; dispatch: api.cu).
#pragma once

#include <stdint.h>

#include "lm_state.cuh"

namespace tob200 {

constexpr int kWtcMinN = 28;       // below, the 4 x 4-block warp-per-problem kernels are faster (measured: DESIGN.md §5.5); the kernel itself runs n >= 13
constexpr int kWtcMaxN = 55;       // above: large-n family
constexpr int kWtcMinM = 192;      // shorter problems are dominated by the solve: the exact kernel is as fast (measured)
constexpr int kWtcSlots = 8;       // problems in flight per CTA
constexpr int kWtcPairs = 4;       // two slots share one 128 x 128 accumulator (its two 64 x 64 diagonal blocks)
constexpr int kWtcNP = 64;         // operand rows per slot
constexpr int kWtcRows = 32;       // rows of A per stage == two K = 16 steps of tcgen05.mma.kind::f16
constexpr int kWtcMaxRawStages = 5;
constexpr int kWtcMaxOpStages = 3;
constexpr int kWtcColWarps = 8;    // (row half of the stage) x (32 operand rows)
#ifndef TOB200_WTC_TWARPS
#define TOB200_WTC_TWARPS 4
#endif
constexpr int kWtcTWarps = TOB200_WTC_TWARPS;             // 4: two per side, alternating chunks; 2: one per side
constexpr int kWtcThreads = (24 + kWtcTWarps - 2) * 32;   // 24 / 26 / 28 warps (80 / 72 / 72 registers), roles in wtc.cuh
constexpr int kWtcOpStageBytes = 2 * 128 * kWtcRows * 2;  // hi + lo, 128 operand rows, FP16
constexpr int kWtcVecs = 21;       // 64-float vectors per slot (enum WtcVec in wtc.cuh)

// pitch of the LDLT working matrix (floats): multiple of 4 with pitch / 4 odd (wpp.cuh: wpp_ldw)
__host__ __device__ constexpr int wtc_ldw(int n) {
  const int n4 = (n + 3) & ~3;
  return ((n4 / 4) & 1) ? n4 : n4 + 4;
}

// per-slot persistent copy of the last accumulated H: n x ldw matrix (as accumulated: column-scaled), then the 64 scale
// factors that undo the scaling
__host__ __device__ constexpr int wtc_hp_floats(int n) { return n * wtc_ldw(n) + kWtcNP; }

struct WtcSmem {  // byte offsets
  uint32_t bars, desc, vec, vec_stride, rsr, w, w_stride, raw, raw_side, raw_stage, ops, total;
  int raw_stages, op_stages;
};

__host__ __device__ inline WtcSmem wtc_smem_plan(int n, int raw_stages, int op_stages) {
  WtcSmem L;
  uint32_t o = 0;
  L.bars = o; o += 448;              // 49 mbarriers + the TMEM base address
  L.desc = o; o += 192;              // per-visit slot descriptors (double buffered)
  o = (o + 127u) & ~127u;
  L.vec = o; L.vec_stride = kWtcVecs * kWtcNP * 4u; o += kWtcSlots * L.vec_stride;
  L.rsr = o; o += 2u * kWtcMaxRawStages * 2u * kWtcRows * 4u;  // row scale and residual per raw stage and side
  o = (o + 127u) & ~127u;
  // LDLT working matrix of a slot: rows 0 .. n (row n carries the right-hand side through the factorisation)
  L.w = o; L.w_stride = ((uint32_t)(n + 1) * (uint32_t)wtc_ldw(n) * 4u + 15u) & ~15u; o += kWtcSlots * L.w_stride;
  o = (o + 127u) & ~127u;
  L.raw = o; L.raw_side = (uint32_t)kWtcRows * (uint32_t)n * 4u; L.raw_stage = 2u * L.raw_side;
  o += (uint32_t)raw_stages * L.raw_stage;
  o = (o + 127u) & ~127u;
  L.ops = o; o += (uint32_t)op_stages * (uint32_t)kWtcOpStageBytes;
  L.total = o;
  L.raw_stages = raw_stages;
  L.op_stages = op_stages;
  return L;
}

// deepest rings that fit the 227 KB of an SM
__host__ inline WtcSmem wtc_smem_best(int n) {
  const int tries[7][2] = {{5, 3}, {4, 3}, {4, 2}, {3, 3}, {3, 2}, {2, 2}, {2, 1}};
  WtcSmem L{};
  for (int t = 0; t < 7; ++t) {
    L = wtc_smem_plan(n, tries[t][0], tries[t][1]);
    if (L.total <= 232448u) break;
  }
  return L;
}

struct WtcParams {
  const float *A;          // [B][m][n]
  const float *y;          // [B][m]
  float *x;                // [B][n] in / out
  tob200_result *results;  // [B]
  int64_t B;
  int m, n;
  DevOptions<float> opt;
  float alpha, alpha3;
  unsigned long long *counter;  // problem queue, zeroed before the launch
  float *hpersist;              // [grid][kWtcSlots][wtc_hp_floats(n)]: H_ of the slot's last Build (L2 resident)
  WtcSmem L;
  int prefetch;                 // L2 prefetch distance of the loader, in stages
  int debug;                    // timing experiments (env TOB200_WTC_DEBUG): 1 no MMAs, 2 no LDLT (results invalid)
  // mode 1: one Build + Solve per problem from materialised blocks (tob200_build_solve_f32): A = J, y = r, no LM state
  int mode;
  const float *lambda;          // [B] or nullptr
  float *dx;                    // [B][n]   (written for accepted problems only)
  double *cost_out;             // [B] sum r^2
  float *H_out;                 // [B][n][n] damped H_, full symmetric, or nullptr
  float *g_out;                 // [B][n] or nullptr
  int32_t *status;              // [B] 0 solved, 1 the factorisation rejected the system
};

}  // namespace tob200
