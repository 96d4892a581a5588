// wtc_params.h — parameter block, launch geometry and shared-memory plan of the mid-n tensor-core
// family (kernel: wtc.cuh; launcher: wtc_kernels.cu; dispatch: api.cu).
#pragma once

#include <stdint.h>

#include "lm_state.cuh"

namespace tob200 {

constexpr int kWtcMinN = 13;       // below: thread-per-problem family
constexpr int kWtcMaxN = 55;       // above: large-n family
constexpr int kWtcMinM = 64;       // shorter problems are not worth a tensor-core pass
constexpr int kWtcSlots = 8;       // problems in flight per CTA
constexpr int kWtcPairs = 4;       // two slots share one 128 x 128 accumulator (its two 64 x 64 diagonal blocks)
constexpr int kWtcNP = 64;         // operand rows per slot
constexpr int kWtcRows = 32;       // rows of A per stage == two K = 16 steps of tcgen05.mma.kind::f16
constexpr int kWtcMaxRawStages = 3;
constexpr int kWtcMaxOpStages = 3;
constexpr int kWtcColWarps = 8;    // (row half of the stage) x (32 operand rows)
constexpr int kWtcThreads = 704;   // 22 warps, roles in wtc.cuh
constexpr int kWtcOpStageBytes = 2 * 128 * kWtcRows * 2;  // hi + lo, 128 operand rows, FP16
constexpr int kWtcStgPitch = 24;   // rows / columns 32 .. 55 of a slot's H travel through a 24 x 24 staging tile
constexpr int kWtcVecs = 19;       // 64-float vectors per slot (enum WtcVec in wtc.cuh)

// pitch of the LDLT working matrix (floats): multiple of 4 with pitch / 4 odd (wpp.cuh: wpp_ldw)
__host__ __device__ constexpr int wtc_ldw(int n) {
  const int n4 = (n + 3) & ~3;
  return ((n4 / 4) & 1) ? n4 : n4 + 4;
}

struct WtcSmem {  // byte offsets
  uint32_t bars, desc, vec, vec_stride, stg, stg_stride, rsr, w, w_stride, raw, raw_side, raw_stage, ops, total;
  int raw_stages, op_stages;
};

__host__ __device__ inline WtcSmem wtc_smem_plan(int n, int raw_stages, int op_stages) {
  WtcSmem L;
  uint32_t o = 0;
  L.bars = o; o += 320;              // 35 mbarriers + the TMEM base address
  L.desc = o; o += 192;              // per-visit slot descriptors (double buffered)
  o = (o + 127u) & ~127u;
  L.vec = o; L.vec_stride = kWtcVecs * kWtcNP * 4u; o += kWtcSlots * L.vec_stride;
  L.stg = o; L.stg_stride = kWtcStgPitch * kWtcStgPitch * 4u; o += kWtcSlots * L.stg_stride;
  L.rsr = o; o += 2u * kWtcMaxRawStages * 2u * kWtcRows * 4u;  // row scale and residual per raw stage and side
  o = (o + 127u) & ~127u;
  L.w = o; L.w_stride = ((uint32_t)n * (uint32_t)wtc_ldw(n) * 4u + 15u) & ~15u; o += kWtcSlots * L.w_stride;
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
  const int tries[4][2] = {{3, 3}, {3, 2}, {2, 2}, {2, 1}};
  WtcSmem L{};
  for (int t = 0; t < 4; ++t) {
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
  float *hpersist;              // [grid][kWtcSlots][n * ldw]: persistent damped H_ of the problems that need it
  WtcSmem L;
  int prefetch;                 // L2 prefetch distance of the loader, in stages
  int debug;                    // timing experiments (env TOB200_WTC_DEBUG): 1 no MMAs, 2 no LDLT (results invalid)
};

}  // namespace tob200
