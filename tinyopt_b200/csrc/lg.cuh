// lg.cuh — "large n" kernels (56 <= n <= 512, float): config C5 (n = 512, m = 4096).
//
// At this size H = J^T J is a true dense contraction (2 m n^2 flop per problem, 128-256 flop/B), so
// it runs on the 5th-generation tensor cores; everything else stays FP32 CUDA-core work.  One LM
// iteration of the batch is three kernels over the still-running problems (host-orchestrated, the
// per-problem state lives in HBM; a whole-batch iteration is tens of milliseconds, so launch latency
// is irrelevant here):
//
//   lg_eval_kernel   t = A x, r, cost, the Jacobian row scale s_i = 1 + 3 alpha t_i^2 and
//                    g = J^T r.  One CTA per problem, rows streamed by TMA bulk copies
//                    (cp.async.bulk, 16-row chunks of 32 KB through a 3-stage mbarrier ring).  HBM bound.
//   lg_syrk_kernel   H = A^T diag(s^2) A on tcgen05: warp-specialised, one CTA per (problem, 128-row
//                    strip of H); producer warps load rows of A, scale them by s_i, split every value
//                    into TF32 hi + lo parts and store both, transposed, into the K-major UMMA
//                    core-matrix layout; one thread issues tcgen05.mma.kind::tf32 (hi*hi + hi*lo + lo*hi: "3xTF32",
//                    FP32-level accuracy) into a 128 x (n - 128 r) FP32 accumulator in TMEM — only the
//                    blocks on or above the diagonal are computed; four epilogue warps drain TMEM with
//                    tcgen05.ld and write the strip of H.  Tensor-pipe bound.
//   lg_solve_kernel  damping, Eigen's pivot order, P H P^T, blocked left-looking LDL^T (32-column
//                    panels, thread per row, every dot product a left-to-right fma chain: the same
//                    operation sequence as the CPU oracle's unblocked factorisation, so the solver is
//                    bit-exact for a given H), blocked substitutions, then the LM state machine of
//                    lm_state.cuh and the x update.  One CTA per problem.
//
// Reference path replaced: diff/optimize_autodiff.h:151-157, solvers/lm.h:60-120,
// solvers/gn.h:150-171, math.h:232-240, optimizers/optimizer.h:243-539.
#pragma once

#include <cstdio>

#include <cuda_fp16.h>

#include "common.cuh"
#include "lg_params.h"
#include "tc.cuh"
#include "lm_state.cuh"

namespace tob200 {


// ================================================================================================
// lg_eval_kernel
// ================================================================================================

__global__ void __launch_bounds__(kLgEvalThreads, 2) lg_eval_kernel(const __grid_constant__ LgEvalParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.n, m = p.m;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
  float *xs = reinterpret_cast<float *>(smem + 64);
  float *wbuf2 = xs + lg_np(n);  // [2][16] row weights s_i r_i, double buffered: ONE barrier per chunk (below)
  float *sbuf2 = wbuf2 + 32;     // [2][16] row scales s_i
  float *cbuf = sbuf2 + 32;
  // (an OFFSET rounded up, not the address: a pointer that went through an integer keeps no address space and every read of
  //  a stage became a generic LD, scoreboarded like a global access, instead of an LDS)
  const size_t stage_off = ((size_t)(reinterpret_cast<unsigned char *>(cbuf + 16) - smem) + 127) & ~(size_t)127;
  float *stages = reinterpret_cast<float *>(smem + stage_off);
  const uint32_t stage_elems = (uint32_t)kLgEvalRows * n;
  if (tid == 0) {
    for (int s = 0; s < kLgEvalStages; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  uint32_t it = 0;  // chunks consumed so far (ring position / parity)
  const int nchunks = (m + kLgEvalRows - 1) / kLgEvalRows;

  for (int64_t pr = blockIdx.x; pr < p.B; pr += gridDim.x) {
    bool rebuild = true;
    if (p.rec) {
      const uint32_t fl = p.rec[pr].flags;
      if (fl & kFlagDone) continue;
      rebuild = !p.is_lm || (fl & kFlagRebuild);
    }
    const float *Ap = p.A + (size_t)pr * m * n;
    const float *yp = p.y + (size_t)pr * m;
    if (p.synth)
      for (int j = tid; j < n; j += kLgEvalThreads) xs[j] = p.x[(size_t)pr * n + j];
    __syncthreads();  // xs ready; every thread is past the previous problem's stages
    auto issue = [&](int c, uint32_t st) {
      const int row0 = c * kLgEvalRows;
      const int nrows = (m - row0 < kLgEvalRows) ? (m - row0) : kLgEvalRows;
      const uint32_t bytes = (uint32_t)nrows * (uint32_t)n * 4u;
      mbar_expect_tx(&bars[st], bytes);
      tma_bulk_g2s(stages + (size_t)st * stage_elems, Ap + (size_t)row0 * n, bytes, &bars[st]);
    };
    if (tid == 0) {
      fence_proxy_async();
      for (int c = 0; c < kLgEvalStages && c < nchunks; ++c) issue(c, (it + c) % kLgEvalStages);
    }
    float gacc = 0.f;   // thread j < n: g_j, rows in order
    float dacc = 0.f;   // thread j < n: (J^T J)_jj, rows in order
    float amx = 0.f;    // thread j < n: max_i |J_ij|
    float cacc = 0.f;   // lane 0 of warp w: sum of r_i^2 over its rows
    for (int c = 0; c < nchunks; ++c, ++it) {
      const uint32_t st = it % kLgEvalStages, ph = (it / kLgEvalStages) & 1u;
      const int row = c * kLgEvalRows + warp;
      const float yrow = (row < m && p.y) ? yp[row] : 0.f;  // in flight while the chunk lands
      mbar_wait(&bars[st], ph);
      const float *sa = stages + (size_t)st * stage_elems;
      float *wbuf = wbuf2 + 16 * (c & 1), *sbuf = sbuf2 + 16 * (c & 1);
      // ---- warp = row: t_i, r_i, s_i ----
      if (row < m) {
        float ri, sc = 1.f;
        if (p.synth) {
          const float *arow = sa + warp * n;
          float t = 0.f;
          for (int j = lane * 4; j < n; j += 128) {
            const float4 a4 = *reinterpret_cast<const float4 *>(arow + j);
            const float4 x4 = *reinterpret_cast<const float4 *>(xs + j);
            t = __fmaf_rn(a4.x, x4.x, t);
            t = __fmaf_rn(a4.y, x4.y, t);
            t = __fmaf_rn(a4.z, x4.z, t);
            t = __fmaf_rn(a4.w, x4.w, t);
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, off));
          const float t2 = __fmul_rn(t, t);
          ri = __fmaf_rn(t, __fmaf_rn(p.alpha, t2, 1.f), -yrow);
          sc = __fmaf_rn(p.alpha3, t2, 1.f);
        } else {
          ri = yrow;
          if (p.scale_in) sc = p.scale_in[(size_t)pr * m + row];
        }
        if (lane == 0) {
          cacc = __fmaf_rn(ri, ri, cacc);
          wbuf[warp] = __fmul_rn(sc, ri);
          sbuf[warp] = sc;
          if (p.synth && rebuild) p.scale[(size_t)pr * m + row] = sc;
        }
      } else if (lane == 0) {
        wbuf[warp] = 0.f;
        sbuf[warp] = 0.f;
      }
      // The only barrier of the chunk: this chunk's row weights are visible, and every thread is through the column
      // pass of the chunk before - whose stage can be refilled now, and whose weight buffer the NEXT row pass may
      // overwrite (the weights alternate between two buffers, so the row pass of chunk c + 1 runs beside the column
      // pass of chunk c).
      __syncthreads();
      if (tid == 0 && c >= 1 && c - 1 + kLgEvalStages < nchunks) {
        fence_proxy_async();
        issue(c - 1 + kLgEvalStages, (it - 1) % kLgEvalStages);
      }
      // ---- thread = column: g_j += sum_i (s_i r_i) a_ij, rows in order ----
      if (rebuild && tid < n) {
        // (the kernel issues ~60 % of its cycles: the per-row weights come in as four 16-byte broadcast loads each,
        //  the column walks the stage with a running pointer)
        const int nrows = (m - c * kLgEvalRows < kLgEvalRows) ? (m - c * kLgEvalRows) : kLgEvalRows;
        float wv[kLgEvalRows], sv[kLgEvalRows];
#pragma unroll
        for (int q = 0; q < kLgEvalRows / 4; ++q) {
          const float4 w4 = *reinterpret_cast<const float4 *>(wbuf + 4 * q);
          const float4 s4 = *reinterpret_cast<const float4 *>(sbuf + 4 * q);
          wv[4 * q] = w4.x; wv[4 * q + 1] = w4.y; wv[4 * q + 2] = w4.z; wv[4 * q + 3] = w4.w;
          sv[4 * q] = s4.x; sv[4 * q + 1] = s4.y; sv[4 * q + 2] = s4.z; sv[4 * q + 3] = s4.w;
        }
        const float *ap = sa + tid;
        if (nrows == kLgEvalRows) {  // every chunk but the last: no per-row guard, the sixteen loads issue back to back
          float av[kLgEvalRows];
#pragma unroll
          for (int i = 0; i < kLgEvalRows; ++i) av[i] = ap[(size_t)i * n];
#pragma unroll
          for (int i = 0; i < kLgEvalRows; ++i) {
            gacc = __fmaf_rn(av[i], wv[i], gacc);
            const float jv = __fmul_rn(sv[i], av[i]);
            dacc = __fmaf_rn(jv, jv, dacc);
            amx = fmaxf(amx, fabsf(jv));
          }
        } else {
#pragma unroll
          for (int i = 0; i < kLgEvalRows; ++i) {
            if (i >= nrows) break;  // (rows the TMA did not write may hold anything)
            const float a = *ap;
            ap += n;
            gacc = __fmaf_rn(a, wv[i], gacc);
            const float jv = __fmul_rn(sv[i], a);
            dacc = __fmaf_rn(jv, jv, dacc);
            amx = fmaxf(amx, fabsf(jv));
          }
        }
      }
    }
    __syncthreads();  // the last column pass is through: the stages and both weight buffers are free
    float *wbuf = wbuf2;
    if (rebuild && tid < n) {
      p.g[(size_t)pr * n + tid] = gacc;
      if (p.dg) p.dg[(size_t)pr * n + tid] = dacc;
    }
    if (lane == 0) cbuf[warp] = cacc;
    if (p.amax && rebuild) {  // block maximum of |J_ij| (wbuf is free after the last chunk's barrier)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) amx = fmaxf(amx, __shfl_xor_sync(0xffffffffu, amx, off));
      if (lane == 0) wbuf[warp] = amx;
    }
    __syncthreads();
    if (tid == 0) {
      float cs = 0.f;
#pragma unroll
      for (int w = 0; w < kLgEvalRows; ++w) cs = __fadd_rn(cs, cbuf[w]);
      p.cost[pr] = cs;
      if (p.amax && rebuild) {
        float mx = 0.f;
#pragma unroll
        for (int w = 0; w < kLgEvalRows; ++w) mx = fmaxf(mx, wbuf[w]);
        p.amax[pr] = mx;
      }
    }
  }
}

// ================================================================================================
// lg_syrk_kernel (tcgen05 / TMEM)
// ================================================================================================

// (tcgen05 / TMEM / TMA wrappers: tc.cuh)

__device__ __forceinline__ bool lg_skip(const LgSyrkParams &p, int64_t pr) {
  if (!p.rec) return false;
  const uint32_t fl = p.rec[pr].flags;
  if (fl & kFlagDone) return true;
  return p.is_lm && !(fl & kFlagRebuild);
}

// timing experiment (TOB200_LG_DEBUG & 16): cycles the MMA thread of block 0 spends per strip index
__device__ long long g_syrk_strip_cycles[8];
__device__ int g_syrk_strip_count[8];

__global__ void __launch_bounds__(kLgSyrkThreads, 1) lg_syrk_kernel(const __grid_constant__ LgSyrkParams p) {
  extern __shared__ unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // stages 1024-byte aligned (swizzle atoms), barriers behind them
  unsigned char *stages = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = 2u * p.half_bytes;
  float *raw = reinterpret_cast<float *>(stages + (size_t)p.stages * stage_bytes);  // kLgRawStages x half_bytes
  uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(raw) + (size_t)p.raw_stages * lg_syrk_raw_bytes(p.np));
  uint64_t *full = bars, *empty = bars + kLgMaxStages, *tmem_full = bars + 2 * kLgMaxStages, *tmem_empty = tmem_full + 1;
  uint64_t *raw_full = tmem_empty + 1, *raw_empty = raw_full + kLgMaxStages;
  uint64_t *peer_empty = raw_empty + kLgMaxStages;  // multicast mode, CTA 0: CTA 1 has released raw stage s
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(peer_empty + kLgMaxStages);
  const bool mc = p.mc != 0;
  const uint32_t crank = mc ? cluster_ctarank() : 0u;
  // Ring geometry per strip.  The operand region (p.stages stages of the widest strip) and the raw region
  // (kLgRawStages of them) are re-cut for every strip into stages of ITS width ncs, so a strip of 128 / 256
  // columns runs 8 / 4 stages deep in the bytes that give the 512-column strip two: the narrow strips
  // need few tensor cycles per stage and are otherwise bound by the hand-off and TMA latency.  Stage s
  // of a strip owns barrier s; every role walks the same (strip, stage) sequence, so the use count of
  // each barrier — whose parity is what a wait needs — is tracked per barrier in a bit mask.  Stages
  // of different strips overlap at different offsets: a new strip starts only after the MMAs of the one
  // before have completed (tmem_full), which drains both rings.
  const int np_ = p.np;
  const uint32_t op_region = (uint32_t)p.stages * stage_bytes, raw_region = (uint32_t)p.raw_stages * lg_syrk_raw_bytes(p.np);
  constexpr uint32_t kBoxBytes = kLgBoxCols * kLgStageK * 4;  // one TMA box: 16 rows x 128 columns
  // A stage holds RS rows: 16 for the wide strips, 32 / 64 for strips of <= 256 / 128 columns when the widest
  // strip has 512, so that every stage carries about the same bytes and the per-stage hand-off cost
  // (~1400 cycles whatever the width) is paid per byte, not per 16 rows.
  const int np_boxes = ((np_ < 128 ? 128 : np_) + kLgBoxCols - 1) / kLgBoxCols;
  auto ring_geom = [&](int ncs, int &RS, uint32_t &hb, uint32_t &S, uint32_t &R) {
    const int ncg = (ncs + kLgBoxCols - 1) / kLgBoxCols;  // column groups == boxes per 16 rows
    int rmul = np_boxes / ncg;                             // whole 16-row boxes that fit the widest raw stage
    if (rmul > 4) rmul = 4;
    for (;; rmul >>= 1) {
      if (rmul < 1) rmul = 1;
      RS = kLgStageK * rmul;
      hb = (uint32_t)ncs * (uint32_t)RS * (p.fp16 ? 2u : 4u);  // bytes of the hi (== lo) part of an operand stage
      S = op_region / (2u * hb);
      R = raw_region / ((uint32_t)(rmul * ncg) * kBoxBytes);  // raw stage = whole boxes
      if ((S >= 2 && R >= 2) || rmul == 1) break;
    }
    if (S > (uint32_t)kLgMaxStages) S = kLgMaxStages;
    if (R > (uint32_t)kLgMaxStages) R = kLgMaxStages;
  };

  if (tid == 0) {
    for (int s = 0; s < kLgMaxStages; ++s) {
      mbar_init(&full[s], kLgProdWarps);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, kLgEpiWarps);
    for (int s = 0; s < kLgMaxStages; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], kLgProdWarps);
      mbar_init(&peer_empty[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == 0) {  // the whole TMEM of this SM: 128 lanes x 512 FP32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (mc) cluster_sync_all();  // the peer's barriers are initialised before anybody arrives on them remotely
  const uint32_t tmem_base = *tmem_slot;

  // Work unit = (problem, pair of strips {h, nstrips - 1 - h}): both units of a problem cost the same
  // (strip r has nstrips - r blocks) and are taken by neighbouring CTAs at the same time, so the
  // second reader of a row of A hits L2 (strip-major order re-read A from HBM 2.5 times at n = 512).
  const int upp = (p.nstrips + 1) / 2;  // units per problem
  const int64_t total = (int64_t)upp * p.B;
  // odd strip counts leave one lighter unit: rotate which CTA gets it from round to round
  const bool rotate = (gridDim.x % upp) == 0;
  auto unit_of = [&](int64_t idx, int64_t &pr, int &h) {
    pr = idx / upp;
    h = (int)(idx % upp);
    if (rotate && !mc) h = (int)((h + (idx - blockIdx.x) / gridDim.x) % upp);  // (multicast: h == rank in the cluster)
  };
  const int m = p.m, n = p.n, np = p.np;

  if (warp == kLgLoadWarp) {
    // ===================== loader: streams raw rows of A with 2-D tensor-map TMA =====================
    // A raw stage holds the kLgStageK rows of one stage, columns [c0, c0 + ncs) of the strip, as boxes of
    // 128 columns (box b = [16 rows][128 floats], 8 KB).  Columns >= n are zero-filled by the TMA unit,
    // rows >= m belong to the next problem: the producers select 0 for both.
    // lane 0 owns the barrier; lane b < nbox issues the copy of box b, lane 16 + b its L2 prefetch
    uint32_t rawe_bits = 0, item = 0;  // bit s: parity of the use count of raw stage s
    uint32_t peer_bits = 0;            // multicast mode, CTA 0: the same for peer_empty
    for (int64_t idx = blockIdx.x; idx < total; idx += gridDim.x) {
      int64_t pr; int h;
      unit_of(idx, pr, h);
      if (lg_skip(p, pr)) continue;
      for (int half = 0; half < 2; ++half, ++item) {
        const int r = half == 0 ? h : p.nstrips - 1 - h;
        if (half == 1 && r == h) break;
        const int c0 = 128 * r;
        const int ncs = (np - c0 < 128) ? 128 : (np - c0);
        uint32_t hb, S, R, rs = 0;
        int RS;
        ring_geom(ncs, RS, hb, S, R);
        const bool mc1 = mc && half == 0;  // strips 0 (CTA 0) and 1 (CTA 1) in lockstep: CTA 0 loads for both
        if (mc1 && crank == 1) {
          // CTA 1 only relays: when its producers have released raw stage rs it posts the bytes it expects (its three
          // boxes) on its own raw_full and tells CTA 0, which issues the multicast copies.  Same ring depth as CTA 0.
          uint32_t hb0, S0, R0;
          int RS0;
          ring_geom(np, RS0, hb0, S0, R0);
          const int ksteps1 = (m + RS - 1) / RS;
          if (item > 0 && lane == 0) mbar_wait(tmem_full, (item - 1u) & 1u);
          __syncwarp();
          for (int ks = 0; ks < ksteps1; ++ks) {
            if (lane == 0) {
              mbar_wait(&raw_empty[rs], ((rawe_bits >> rs) & 1u) ^ 1u);
              fence_proxy_async();
              mbar_expect_tx(&raw_full[rs], (uint32_t)((ncs + kLgBoxCols - 1) / kLgBoxCols) * kBoxBytes);
              mbar_arrive_remote(&peer_empty[rs], 0u);
            }
            __syncwarp();
            rawe_bits ^= 1u << rs;
            if (++rs == R0) rs = 0;
          }
          continue;
        }
        const int ksteps = (m + RS - 1) / RS;
        const int ncg = (ncs + kLgBoxCols - 1) / kLgBoxCols;
        const int nbox = ncg * (RS / kLgStageK);  // boxes per stage (<= 4): box b = (row box b / ncg, column box b % ncg)
        const uint32_t raw_stage = (uint32_t)nbox * kBoxBytes;
        const int ccnt = (n - c0 < ncs) ? (n - c0) : ncs;  // real columns (> 0: c0 <= np - 32 < n)
        const float *Ap = p.A + (size_t)pr * m * n + c0;
        const int grow0 = (int)(pr * m);  // first row of the problem in the {n, B * m} tensor
        const int pf = (int)R + kLgPrefetchStages;  // L2 prefetch distance in stages
        // The TMA unit retires roughly one bulk operation per 70 cycles per SM whatever its size, so a stage
        // is ONE tensor copy per 128-column box (16 rows x 512 B, <= 4 per stage; 16 per-row copies + 16
        // per-row prefetches made the loader the bottleneck at ~2300 cycles per stage for every strip width).
        // L2 prefetch (same boxes) runs `pf` stages ahead: no shared memory needed.
        if (p.use_tmap) {
          for (int d = 0; d < pf; ++d)
            if (lane < nbox && d * RS + kLgStageK * (lane / ncg) < m)
              tma_prefetch_box(&p.tmap, c0 + kLgBoxCols * (lane % ncg), grow0 + d * RS + kLgStageK * (lane / ncg));
        }
        // drain (the rings are re-cut for this strip) — after the prefetches above, so that the first
        // stages of the new strip are on their way into L2 while the old strip finishes
        if (item > 0 && lane == 0) mbar_wait(tmem_full, (item - 1u) & 1u);
        __syncwarp();
        for (int ks = 0; ks < ksteps; ++ks) {
          const int row0 = ks * RS;
          const int rows = (m - row0 < RS) ? (m - row0) : RS;
          if (lane == 0) {
            mbar_wait(&raw_empty[rs], ((rawe_bits >> rs) & 1u) ^ 1u);
            if (mc1) {  // ... and CTA 1 has released the same stage of ITS ring
              mbar_wait_cluster(&peer_empty[rs], (peer_bits >> rs) & 1u);
              peer_bits ^= 1u << rs;
            }
            fence_proxy_async();  // the producers' generic reads of this stage precede the async writes
            if (p.debug & 4) mbar_arrive(&raw_full[rs]);  // timing experiment: no copies
            else if (p.use_tmap) mbar_expect_tx(&raw_full[rs], raw_stage);  // OOB parts of a box count too
            else mbar_expect_tx(&raw_full[rs], (uint32_t)rows * (uint32_t)ccnt * 4u);
          }
          __syncwarp();
          if (!(p.debug & 4)) {
            unsigned char *dst = reinterpret_cast<unsigned char *>(raw) + (size_t)rs * raw_stage;
            if (p.use_tmap) {
              if (lane < nbox) {
                if (mc1 && (lane % ncg) > 0)  // columns >= 128: strip 1 needs them too
                  tma_load_box_multicast(dst + (size_t)lane * kBoxBytes, &p.tmap, c0 + kLgBoxCols * (lane % ncg),
                                         grow0 + row0 + kLgStageK * (lane / ncg), &raw_full[rs], (uint16_t)3);
                else
                  tma_load_box(dst + (size_t)lane * kBoxBytes, &p.tmap, c0 + kLgBoxCols * (lane % ncg),
                               grow0 + row0 + kLgStageK * (lane / ncg), &raw_full[rs]);
              } else if (lane >= 16 && lane - 16 < nbox) {
                const int prow = row0 + pf * RS + kLgStageK * ((lane - 16) / ncg);
                if (prow < m) tma_prefetch_box(&p.tmap, c0 + kLgBoxCols * ((lane - 16) % ncg), grow0 + prow);
              }
            } else {  // fallback: one bulk copy per row and column box, same layout
              for (int rw = lane; rw < rows; rw += 32) {
                for (int bx = 0; bx < ncg; ++bx) {
                  const int cc = ccnt - kLgBoxCols * bx;
                  if (cc > 0)
                    tma_bulk_g2s(dst + (size_t)((rw / kLgStageK) * ncg + bx) * kBoxBytes + (size_t)(rw % kLgStageK) * (kLgBoxCols * 4),
                                 Ap + (size_t)(row0 + rw) * n + kLgBoxCols * bx,
                                 (uint32_t)(cc < kLgBoxCols ? cc : kLgBoxCols) * 4u, &raw_full[rs]);
                }
              }
            }
          }
          rawe_bits ^= 1u << rs;
          if (++rs == R) rs = 0;
        }
      }
    }
  } else if (warp > kLgLoadWarp && p.fp16) {
    // ===================== producers, FP16 split =====================
    // Same protocol as the TF32 producers below; a 16-byte K chunk of the operand now holds EIGHT rows, so the
    // stage is cut into chunks of 8 rows and, per 128-column group, two halves of 64 columns: warp w owns (chunk kc,
    // group cg, half hq), lane l the columns 64 hq + l + 32 q, q = 0, 1 - sixteen elements per lane and stage as
    // before, one 16-byte store per column and part.  Every value is scaled by the problem's power of two (exact),
    // hi = fp16(v), lo = fp16(v - hi).
    const int w = warp - (kLgLoadWarp + 1);
    uint32_t empty_bits = 0, rawf_bits = 0, item = 0;
    for (int64_t idx = blockIdx.x; idx < total; idx += gridDim.x) {
      int64_t pr; int h;
      unit_of(idx, pr, h);
      if (lg_skip(p, pr)) continue;
      const float fs = ldexpf(1.f, lg_fp16_exp(p.amax ? p.amax[pr] : 1.f));
     for (int half = 0; half < 2; ++half, ++item) {
      const int r = half == 0 ? h : p.nstrips - 1 - h;
      if (half == 1 && r == h) break;
      const int c0 = 128 * r;
      const int ncs = (np - c0 < 128) ? 128 : (np - c0);
      const uint32_t lbo = (uint32_t)ncs * 16u;
      uint32_t hb, S, R, st = 0, rs = 0;
      int RS;
      ring_geom(ncs, RS, hb, S, R);
      const int ksteps = (m + RS - 1) / RS;
      const int ncg = (ncs + kLgBoxCols - 1) / kLgBoxCols;
      uint32_t raw_stage = (uint32_t)(ncg * (RS / kLgStageK)) * kBoxBytes;
      uint32_t box_shift = 0;
      if (mc && half == 0 && crank == 1) {  // multicast phase: the raw stages have CTA 0's geometry (strip 0: one box more,
        uint32_t hb0, S0;                   // in front) and ring depth
        int RS0;
        ring_geom(np, RS0, hb0, S0, R);
        raw_stage = (uint32_t)((ncg + 1) * (RS / kLgStageK)) * kBoxBytes;
        box_shift = 1;
      }
      const int per = 2 * ncg;
      const int kc = w / per, cg = (w % per) >> 1, hq = w & 1;
      const bool mine = kc < RS / 8;
      if (item > 0) mbar_wait(tmem_full, (item - 1u) & 1u);  // drain: the rings are re-cut for this strip
      const int rr0 = kLgBoxCols * cg + 64 * hq + lane;  // my operand rows: rr0 + 32 q, q = 0, 1
      const float *sp = p.scale ? p.scale + (size_t)pr * m : nullptr;
      auto load_scale = [&](int ks) {
        const int row = ks * RS + 8 * kc + (lane & 7);
        return (mine && ks < ksteps && row < m) ? (sp ? sp[row] : 1.f) : 0.f;
      };
      float sc_a = load_scale(0), sc_b = load_scale(1);
      const uint32_t raw_u32 = smem_u32(raw), stages_u32 = smem_u32(stages);
      for (int ks = 0; ks < ksteps; ++ks) {
        const float sc_cur = sc_a;
        sc_a = sc_b;
        sc_b = load_scale(ks + 2);
        float sc[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) sc[t] = __shfl_sync(0xffffffffu, sc_cur, t);
        mbar_wait(&raw_full[rs], (rawf_bits >> rs) & 1u);
        // row 8 kc + t of the stage sits in row box kc / 2, row 8 (kc % 2) + t of the box
        const uint32_t rsrc = raw_u32 + rs * raw_stage + (uint32_t)((kc >> 1) * (ncg + (int)box_shift)) * kBoxBytes +
                              box_shift * kBoxBytes + (uint32_t)(8 * (kc & 1)) * (kLgBoxCols * 4u);
        const int row0 = ks * RS + 8 * kc;
        if ((p.debug & 2) || !mine) {  // idle warp (or timing experiment): barrier protocol only
          mbar_wait(&empty[st], ((empty_bits >> st) & 1u) ^ 1u);
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&full[st]);
            mbar_arrive(&raw_empty[rs]);
          }
          empty_bits ^= 1u << st;
          rawf_bits ^= 1u << rs;
          if (++st == S) st = 0;
          if (++rs == R) rs = 0;
          continue;
        }
        float bv[2][8];  // [q][t]: column rr0 + 32 q, row 8 kc + t
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int rr = rr0 + 32 * q;
          const bool cok = rr < ncs && c0 + rr < n;
#pragma unroll
          for (int t = 0; t < 8; ++t)
            bv[q][t] = (cok && row0 + t < m)
                           ? lds_f32(rsrc + (uint32_t)cg * kBoxBytes + (uint32_t)(t * kLgBoxCols + 64 * hq + 32 * q + lane) * 4u)
                           : 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[rs]);
        mbar_wait(&empty[st], ((empty_bits >> st) & 1u) ^ 1u);
        const uint32_t sb = stages_u32 + st * 2u * hb;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int rr = rr0 + 32 * q;  // operand row (column of A relative to c0)
          if (rr < ncs) {
            float v[8];
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = __fmul_rn(__fmul_rn(bv[q][t], sc[t]), fs);  // J_ij, then * 2^e (exact)
#pragma unroll
            for (int t2 = 0; t2 < 4; ++t2) {
              hi[t2] = lg_pack_h2(v[2 * t2], v[2 * t2 + 1]);
              lo[t2] = lg_pack_h2(__fsub_rn(v[2 * t2], lg_h_lo(hi[t2])), __fsub_rn(v[2 * t2 + 1], lg_h_hi(hi[t2])));
            }
            const uint32_t off = (uint32_t)kc * lbo + (uint32_t)(rr >> 3) * 128u + (uint32_t)(rr & 7) * 16u;
            sts_v4u(sb + off, hi[0], hi[1], hi[2], hi[3]);
            sts_v4u(sb + hb + off, lo[0], lo[1], lo[2], lo[3]);
          }
        }
        fence_proxy_async();  // generic-proxy stores -> the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[st]);
        empty_bits ^= 1u << st;
        rawf_bits ^= 1u << rs;
        if (++st == S) st = 0;
        if (++rs == R) rs = 0;
      }
     }
    }
  } else if (warp > kLgLoadWarp) {
    // ===================== producers =====================
    // An operand stage is the K-major image of RS rows x ncs columns of diag(s) A (RS / 8 MMA K steps):
    // operand row = column j of A, K = row i.  The stage is cut into K chunks of 4 rows and column groups
    // of 128; warp w owns chunk kc = w / ncg of group cg = w % ncg (RS / 4 * ncg <= 16 warps), lane l the
    // columns l + 32 q of the group — conflict-free 4-byte reads of the raw boxes (lane = consecutive
    // column), scaled by s_i, split into TF32 hi + lo, conflict-free 16-byte stores (one per column and
    // part).  Sixteen elements per lane and stage whatever the strip width.
    const int w = warp - (kLgLoadWarp + 1);
    uint32_t empty_bits = 0, rawf_bits = 0, item = 0;  // per-barrier use-count parities (see ring_geom)
    for (int64_t idx = blockIdx.x; idx < total; idx += gridDim.x) {
      int64_t pr; int h;
      unit_of(idx, pr, h);
      if (lg_skip(p, pr)) continue;
     for (int half = 0; half < 2; ++half, ++item) {
      const int r = half == 0 ? h : p.nstrips - 1 - h;
      if (half == 1 && r == h) break;
      const int c0 = 128 * r;
      const int ncs = (np - c0 < 128) ? 128 : (np - c0);  // columns staged (the A operand needs 128)
      const uint32_t lbo = (uint32_t)ncs * 16u;
      uint32_t hb, S, R, st = 0, rs = 0;
      int RS;
      ring_geom(ncs, RS, hb, S, R);
      const int ksteps = (m + RS - 1) / RS;
      const int ncg = (ncs + kLgBoxCols - 1) / kLgBoxCols;
      const uint32_t raw_stage = (uint32_t)(ncg * (RS / kLgStageK)) * kBoxBytes;
      const int kc = w / ncg, cg = w % ncg;
      const bool mine = kc < RS / 4;  // warp uniform; idle warps only keep the barrier protocol
      if (item > 0) mbar_wait(tmem_full, (item - 1u) & 1u);  // drain: the rings are re-cut for this strip
      const int rr0 = kLgBoxCols * cg + lane;  // my operand rows: rr0 + 32 q, q = 0..3
      const float *sp = p.scale ? p.scale + (size_t)pr * m : nullptr;
      // row scales of my four rows: lanes 0..3 load them two stages ahead of their use, shuffled out below
      auto load_scale = [&](int ks) {
        const int row = ks * RS + 4 * kc + (lane & 3);
        return (mine && ks < ksteps && row < m) ? (sp ? sp[row] : 1.f) : 0.f;
      };
      float sc_a = load_scale(0), sc_b = load_scale(1);
      const uint32_t raw_u32 = smem_u32(raw), stages_u32 = smem_u32(stages);
      for (int ks = 0; ks < ksteps; ++ks) {
        const float sc_cur = sc_a;
        sc_a = sc_b;
        sc_b = load_scale(ks + 2);
        float sc[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) sc[t] = __shfl_sync(0xffffffffu, sc_cur, t);
        mbar_wait(&raw_full[rs], (rawf_bits >> rs) & 1u);
        // row 4 kc + t of the stage sits in row box kc / 4, row 4 (kc % 4) + t of the box
        const uint32_t rsrc = raw_u32 + rs * raw_stage + (uint32_t)((kc >> 2) * ncg) * kBoxBytes +
                              (uint32_t)(4 * (kc & 3)) * (kLgBoxCols * 4u);
        const int row0 = ks * RS + 4 * kc;
        if ((p.debug & 2) || !mine) {  // idle warp (or timing experiment): barrier protocol only
          mbar_wait(&empty[st], ((empty_bits >> st) & 1u) ^ 1u);
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&full[st]);
            mbar_arrive(&raw_empty[rs]);
          }
          empty_bits ^= 1u << st;
          rawf_bits ^= 1u << rs;
          if (++st == S) st = 0;
          if (++rs == R) rs = 0;
          continue;
        }
        float bv[4][4];  // [q][t]: column rr0 + 32 q, row 4 kc + t
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int rr = rr0 + 32 * q;
          const bool cok = rr < ncs && c0 + rr < n;  // rr < ncs is warp uniform (ncs is a multiple of 32)
#pragma unroll
          for (int t = 0; t < 4; ++t)
            bv[q][t] = (cok && row0 + t < m)
                           ? lds_f32(rsrc + (uint32_t)cg * kBoxBytes + (uint32_t)(t * kLgBoxCols + 32 * q + lane) * 4u)
                           : 0.f;
        }
        // the raw stage is consumed once the values are in registers: release it before the transform
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[rs]);
        mbar_wait(&empty[st], ((empty_bits >> st) & 1u) ^ 1u);
        const uint32_t sb = stages_u32 + st * 2u * hb;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int rr = rr0 + 32 * q;  // operand row (column of A relative to c0)
          if (rr < ncs) {
            float v[4], hi[4], lo[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              v[t] = __fmul_rn(bv[q][t], sc[t]);
              hi[t] = (p.terms == 3) ? tc_round_tf32(v[t]) : v[t];
              lo[t] = __fsub_rn(v[t], hi[t]);
            }
            const uint32_t off = (uint32_t)kc * lbo + (uint32_t)(rr >> 3) * 128u + (uint32_t)(rr & 7) * 16u;
            sts_v4(sb + off, hi[0], hi[1], hi[2], hi[3]);
            if (p.terms == 3) sts_v4(sb + hb + off, lo[0], lo[1], lo[2], lo[3]);
          }
        }
        fence_proxy_async();  // generic-proxy stores -> the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[st]);
        empty_bits ^= 1u << st;
        rawf_bits ^= 1u << rs;
        if (++st == S) st = 0;
        if (++rs == R) rs = 0;
      }
     }
    }
  } else if (warp == kLgMmaWarp) {
    // ===================== MMA issuer: one thread =====================
    uint32_t full_bits = 0, item = 0;
    for (int64_t idx = blockIdx.x; idx < total; idx += gridDim.x) {
      int64_t pr; int h;
      unit_of(idx, pr, h);
      if (lg_skip(p, pr)) continue;
     for (int half = 0; half < 2; ++half) {
      const int r = half == 0 ? h : p.nstrips - 1 - h;
      if (half == 1 && r == h) break;
      const int nb = np - 128 * r;  // accumulator columns of this strip
      const int ncs = nb < 128 ? 128 : nb;
      uint32_t hb, S, R, st = 0;
      int RS;
      ring_geom(ncs, RS, hb, S, R);
      const int ksteps = (m + RS - 1) / RS;
      const long long t_strip = (p.debug & 16) ? clock64() : 0;
      if (lane == 0) {
        mbar_wait(tmem_empty, (item & 1u) ^ 1u);  // the epilogue has drained the previous strip
        tc_fence_after();
      }
      __syncwarp();
      for (int ks = 0; ks < ksteps; ++ks) {
        if (lane == 0) {
          mbar_wait(&full[st], (full_bits >> st) & 1u);
          tc_fence_after();
          const uint32_t sb0 = smem_u32(stages) + st * 2u * hb;
          const uint32_t lbo = (uint32_t)ncs * 16u;  // K chunk c + 1 follows all the core matrices of chunk c
          const int mma_k = p.fp16 ? 2 * kLgMmaK : kLgMmaK;  // rows per instruction: 32 bytes of K in either format
          for (int kk = 0; kk < RS / mma_k && !(p.debug & 1); ++kk) {  // debug & 1: timing experiment, no MMAs
            const uint32_t sb = sb0 + (uint32_t)(2 * kk) * lbo;  // this K step: chunks 2 kk, 2 kk + 1
            const uint64_t a_hi = tc_desc_k_major(sb, lbo, 128u);
            const uint64_t a_lo = tc_desc_k_major(sb + hb, lbo, 128u);
            for (int n0 = 0; n0 < nb; n0 += 256) {
              const int N = (nb - n0 < 256) ? (nb - n0) : 256;
              const uint64_t b_hi = tc_desc_k_major(sb + (uint32_t)(n0 / 8) * 128u, lbo, 128u);
              const uint64_t b_lo = tc_desc_k_major(sb + hb + (uint32_t)(n0 / 8) * 128u, lbo, 128u);
              const uint32_t d = tmem_base + (uint32_t)n0;
              if (p.fp16) {
                const uint32_t idesc = tc_idesc_f16(N);
                tc_mma_f16(d, a_hi, b_hi, idesc, (ks > 0 || kk > 0) ? 1u : 0u);
                tc_mma_f16(d, a_hi, b_lo, idesc, 1u);
                tc_mma_f16(d, a_lo, b_hi, idesc, 1u);
              } else {
                const uint32_t idesc = tc_idesc_tf32(N);
                tc_mma_tf32(d, a_hi, b_hi, idesc, (ks > 0 || kk > 0) ? 1u : 0u);
                if (p.terms == 3) {
                  tc_mma_tf32(d, a_hi, b_lo, idesc, 1u);
                  tc_mma_tf32(d, a_lo, b_hi, idesc, 1u);
                }
              }
            }
          }
          tc_commit(&empty[st]);  // arrives when the MMAs above have read the stage
          if (ks == ksteps - 1) tc_commit(tmem_full);
        }
        __syncwarp();
        full_bits ^= 1u << st;
        if (++st == S) st = 0;
      }
      if ((p.debug & 16) && lane == 0 && blockIdx.x == 0) {
        g_syrk_strip_cycles[r & 7] += (clock64() - t_strip) * kLgStageK / RS;  // per 16 rows
        g_syrk_strip_count[r & 7] += 1;
      }
      ++item;
     }
    }
  } else {
    // ===================== epilogue: warp w drains TMEM lanes [32 w, 32 w + 32) =====================
    uint32_t item = 0;
    for (int64_t idx = blockIdx.x; idx < total; idx += gridDim.x) {
      int64_t pr; int h;
      unit_of(idx, pr, h);
      if (lg_skip(p, pr)) continue;
     for (int half = 0; half < 2; ++half) {
      const int r = half == 0 ? h : p.nstrips - 1 - h;
      if (half == 1 && r == h) break;
      const int c0 = 128 * r, nb = np - c0;
      // FP16 split: the accumulator holds H * 2^(2e): two exact multiplications by 2^-e bring it back
      const float inv = p.fp16 ? ldexpf(1.f, -lg_fp16_exp(p.amax ? p.amax[pr] : 1.f)) : 1.f;
      mbar_wait(tmem_full, item & 1u);
      tc_fence_after();
      const int row = c0 + 32 * warp + lane;
      float *hrow = p.H + ((size_t)pr * np + row) * np + c0;
      for (int cb = 0; cb < nb; cb += 32) {
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)cb, v);
        tc_wait_ld();
        if (row < np) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4 *>(hrow + cb + 4 * q) =
                make_float4(__uint_as_float(v[4 * q]) * inv * inv, __uint_as_float(v[4 * q + 1]) * inv * inv,
                            __uint_as_float(v[4 * q + 2]) * inv * inv, __uint_as_float(v[4 * q + 3]) * inv * inv);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
      ++item;
     }
    }
  }
  tc_fence_before();
  __syncthreads();
  if ((p.debug & 16) && blockIdx.x == 0 && tid == 0) {
    for (int r = 0; r < p.nstrips && r < 8; ++r) {
      if (g_syrk_strip_count[r])
        printf("syrk block 0: strip %d: %d x %lld kcycles (%lld cycles per 16-row stage)\n", r, g_syrk_strip_count[r],
               g_syrk_strip_cycles[r] / g_syrk_strip_count[r] / 1000,
               g_syrk_strip_cycles[r] / g_syrk_strip_count[r] / ((m + kLgStageK - 1) / kLgStageK));
      g_syrk_strip_cycles[r] = 0;
      g_syrk_strip_count[r] = 0;
    }
  }
  if (mc) cluster_sync_all();  // nobody leaves while its peer may still signal it
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

}  // namespace tob200
