// lg_params.h — parameter blocks and launch geometry of the large-n family (kernels: lg.cuh,
// lg_solve.cuh; launchers: lg_kernels.cu; orchestration: api.cu).
#pragma once

#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "lm_state.cuh"

namespace tob200 {

constexpr int kLgMinN = 56;
constexpr int kLgMaxN = 512;
__host__ __device__ constexpr int lg_np(int n) { return (n + 31) / 32 * 32; }  // padded order (pitch of H, W)

constexpr int kLgEvalThreads = 512;
constexpr int kLgEvalRows = 16;  // rows per TMA chunk == warps per CTA
constexpr int kLgEvalStages = 3;

struct LgEvalParams {
  const float *A;   // [B][m][n] (J when synth == 0)
  const float *y;   // [B][m]    (r when synth == 0)
  const float *x;   // [B][n]    (unused when synth == 0)
  const LmScalars<float> *rec;  // per-problem state or nullptr (nullptr: every problem, full pass)
  const float *scale_in;  // [B][m] row scale of a materialised J (synth == 0), or nullptr
  float *scale;     // [B][m] out: s_i (only written when synth && the pass rebuilds)
  float *g;         // [B][n] out (rebuild passes)
  float *dg;        // [B][n] out (rebuild passes): diag(J^T J) in FP32, rows in order (the tensor core
                    // accumulates with truncation, which biases the long same-sign sums of the diagonal)
  float *cost;      // [B] out: sum r_i^2
  float *amax;      // [B] out (rebuild passes) or nullptr: max |J_ij| of the problem — the FP16-split J^T J kernel
                    // scales by a power of two derived from it
  int64_t B;
  int m, n;
  int synth;        // 1: polynomial family evaluated from A, y, x; 0: materialised J, r
  int is_lm;
  float alpha, alpha3;
};

__host__ __device__ inline size_t lg_eval_smem_bytes(int n) {
  // [bars 64 | x n | w 2 x 16 | s 2 x 16 | cost 16 | stages * 16 rows * n]
  return 64 + (size_t)lg_np(n) * 4 + 5 * 64 + (size_t)kLgEvalStages * kLgEvalRows * n * 4 + 128;
}

constexpr int kLgEpiWarps = 4;    // warps 0..3: TMEM lane quadrant == warp id
constexpr int kLgMmaWarp = 4;     // warp 4: lane 0 issues every tcgen05.mma
constexpr int kLgLoadWarp = 5;    // warp 5: lane 0 streams raw rows of A into the raw ring (TMA bulk copies)
constexpr int kLgProdWarps = 16;  // warps 6..21: 4 K rows x 64 columns of a stage each
constexpr int kLgPrefetchStages = 6;  // L2 prefetch distance beyond the raw ring, in stages
constexpr int kLgRawStages = 2;   // default raw ring depth of the widest strip (LgSyrkParams::raw_stages; env TOB200_LG_RAW_STAGES)
constexpr int kLgSyrkThreads = (kLgEpiWarps + 2 + kLgProdWarps) * 32;
constexpr int kLgMmaK = 8;        // K extent of one tf32 tcgen05.mma
constexpr int kLgStageK = 16;     // rows per stage == two MMA K steps (halves the barrier hand-offs per row)
constexpr int kLgMaxStages = 8;

constexpr int kLgBoxCols = 128;   // raw stages are built from TMA boxes of kLgStageK rows x 128 columns (8 KB)

struct LgSyrkParams {
  alignas(64) CUtensorMap tmap;  // A as a 2-D tensor {n columns, B * m rows}, box {128, kLgStageK}; valid iff use_tmap
  int use_tmap;        // 0: per-row bulk copies into the same box layout (tensor map could not be encoded)
  const float *A;      // [B][m][n]
  const float *scale;  // [B][m] row scale, or nullptr (materialised J)
  const LmScalars<float> *rec;  // nullptr: every problem
  float *H;            // [B][np][np]: strip r writes rows [128 r, 128 r + 128), columns >= 128 r
  int64_t B;
  int m, n, np, nstrips;
  int stages;          // operand ring depth (stages of the widest strip)
  int raw_stages;      // raw ring depth (stages of the widest strip)
  const float *amax;   // [B] max |J_ij| per problem (fp16 != 0), from lg_eval_kernel
  int fp16;            // 1: FP16 hi / lo split, tcgen05.mma.kind::f16 (K = 16 per instruction: twice the TF32 rate,
                       // half the operand bytes); 0: TF32 split, kind::tf32
  int terms;           // 3: hi*hi + hi*lo + lo*hi; 1: plain TF32
  int is_lm;
  int debug;           // timing experiments only (env TOB200_LG_DEBUG): 1 no MMAs, 2 no transform, 4 no copies
  uint32_t half_bytes; // bytes of the hi (== lo) part of a stage == of a raw stage: kLgStageK x max(128, np) floats
  int mc;              // 1: launched as clusters of two CTAs (the two units of a problem); while both stream their first,
                       // wide strips (0 and 1 of four), CTA 0 loads every raw stage ONCE and the TMA unit multicasts the
                       // three column boxes both need to CTA 1 (nstrips == 4, fp16, tensor map only)
};

__host__ __device__ inline uint32_t lg_syrk_half_bytes(int np, int fp16 = 0) {
  return (uint32_t)(np < 128 ? 128 : np) * (uint32_t)kLgStageK * (fp16 ? 2u : 4u);
}
// bytes of one raw stage of the widest strip: whole boxes
__host__ __device__ inline uint32_t lg_syrk_raw_bytes(int np) {
  const int w = np < 128 ? 128 : np;
  return (uint32_t)((w + kLgBoxCols - 1) / kLgBoxCols) * (uint32_t)(kLgBoxCols * kLgStageK * 4);
}
__host__ __device__ inline size_t lg_syrk_smem_bytes(int np, int stages, int raw_stages, int fp16) {
  // [1 KB alignment slack | operand stages (hi + lo) | raw stages | barriers]
  return 1024 + (size_t)stages * 2 * lg_syrk_half_bytes(np, fp16) + (size_t)raw_stages * lg_syrk_raw_bytes(np) + 512;
}

constexpr int kLgSolveThreads = 512;
constexpr int kLgPanel = 32;  // factorisation panel width (16 / 2 CTAs per SM was measured slower: more tiles, more barriers)
constexpr int kLgBlk = 32;    // substitution block == warp
constexpr int kLgTilePitch = 36;  // pitch of the phase-1 W tiles: 16-byte aligned rows (cp.async 16, LDS.128 of L)

struct LgSolveParams {
  // ---- per problem inputs ----
  float *H;            // [B][np][np] upper triangle (row <= col), undamped diagonal after a rebuild
  const float *dg;     // [B][n] FP32 diagonal of the last rebuild (lg_eval), or nullptr: use H's own
  float *hd;           // [B][np] persistent damped diagonal of H_ (solvers/lm.h keeps H_ damped)
  float *g;            // [B][n] grad_
  const float *cost;   // [B] sum r^2 of the pass
  float *W;            // [grid][np][np] factor workspace (lower, permuted)
  int64_t B;
  int n, np, nres;
  int mode;            // 0: LM loop step (rec/x/last_dx/results), 1: one Build+Solve (lambda/dx/status), 2: plain SolveLDLT(H, b),
                       // 3: InvCov(H) (dx = [B][n][n] inverse or nullptr, max_std, status)
  // ---- mode 0 ----
  DevOptions<float> opt;
  LmScalars<float> *rec;
  float *x, *last_dx;
  tob200_result *results;
  unsigned long long *n_active;
  // ---- mode 1 / 2 ----
  const float *lambda;  // [B] or nullptr
  const float *b;       // mode 2: right-hand sides [B][n]
  float *dx;            // [B][n]  (mode 3: [B][n][n])
  float *max_std;       // mode 3: [B] sqrt(max coefficient of the inverse), or nullptr
  double *cost_out;     // [B] (mode 1)
  int32_t *status;      // [B]
};

struct LgSolveSmem {
  // offsets in floats
  int dd, dsm, ysm, rhs, perm, inv, tt, tile, misc, total;
};
__host__ __device__ inline LgSolveSmem lg_solve_smem(int np) {
  LgSolveSmem L;
  const int vl = kLgMaxN;  // vectors are sized for the largest n: the pivot sort pads to a power of two
  int o = 0;
  L.dd = o; o += vl;
  L.dsm = o; o += vl;
  L.ysm = o; o += vl;
  L.rhs = o; o += vl;
  L.perm = o; o += vl;
  L.inv = o; o += vl;
  L.tt = o; o += kLgPanel * kLgPanel;
  {  // double-buffered W tile; also the 16 per-warp 32 x 33 transpose tiles of lg_mirror_upper
    const int a = 2 * np * kLgTilePitch, b = (kLgSolveThreads / 32) * 32 * 33;
    L.tile = o; o += a > b ? a : b;
  }
  L.misc = o; o += 32;
  L.total = o;
  return L;
}

}  // namespace tob200
