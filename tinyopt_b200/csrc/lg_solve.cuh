// lg_solve.cuh — damping + dense pivoted LDL^T solve + LM state machine for large n (one CTA per
// problem, 512 threads, thread = row).  See lg.cuh for the pipeline this kernel closes.
//
// The factorisation is Eigen 3.4's `LDLT<_, Upper>` (math.h:232-240 SolveLDLT) restated as a BLOCKED
// left-looking algorithm with the same arithmetic as the CPU oracle's unblocked one:
//  * Eigen's unblocked LDLT only finishes column k at step k, so the diagonal it searches for the next
//    pivot is the ORIGINAL diagonal: the whole pivot order is a function of diag(H) alone.  It is
//    computed first, W = P H P^T is laid out already permuted, and no search or swap remains.
//  * every element (i, k) of L is (A_ik - sum_{j<k} L_ij (D_j L_kj)) / D_k with the sum a left-to-right
//    fma chain from +0.  Panels of 32 columns: phase 1 accumulates the chain over the columns left of
//    the panel for the whole panel at once (W rows staged through shared memory, 32 accumulators per
//    thread), phase 2 continues it column by column inside the panel (values in registers).
//    Same operands, same order, same roundings => bit-identical to the oracle for the same H.
//  * substitutions in blocks of 32 with each y_i updated in the oracle's order (ascending j forward,
//    descending j backward).
#pragma once

#include <cstdio>

#include "lg.cuh"

namespace tob200 {

// optional per-phase cycle counters of block 0 (build with -DTOB200_LG_TIMING; printed at kernel end)
#ifdef TOB200_LG_TIMING
__device__ long long g_lg_tm[32];
__device__ long long g_lg_t0;
#define LG_T0() do { if (threadIdx.x == 0 && blockIdx.x == 0) g_lg_t0 = clock64(); } while (0)
#define LG_T(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long t = clock64(); g_lg_tm[k] += t - g_lg_t0; g_lg_t0 = t; } } while (0)
#else
#define LG_T0()
#define LG_T(k)
#endif

// monotone key of |d| for the pivot search: 0 for NaN (never greater than anything)
__device__ __forceinline__ uint32_t lg_key(float d) {
  const float v = fabsf(d);
  return (v != v) ? 0u : __float_as_uint(v) + 1u;
}

// Pivot order of Eigen's LDLT for the diagonal dd[0..n): perm[pos] = original index, inv[orig] = pos.
// Eigen picks, at step k, the FIRST largest |d| among positions k..n-1 and swaps it to position k.
// With distinct keys that is a descending sort (bitonic, CTA wide).  With ties (or NaNs) the winner
// among equal keys depends on where earlier swaps have moved them, so the swaps are replayed exactly
// by one thread — in O(n + sum of squared tie-group sizes): the sorted order says which key value is
// due at step k, only the members of that tie group are compared by current position.
__device__ void lg_pivot_order(const float *dd, int n, int np, int *perm, int *inv, float *scratch_keys, int *spare, int *flag) {
  const int tid = threadIdx.x;
  uint32_t *keys = reinterpret_cast<uint32_t *>(scratch_keys);  // kLgMaxN entries
  int sz = 1;  // sort size: next power of two >= n (<= 512 == blockDim)
  while (sz < n) sz <<= 1;
  if (tid < sz) {
    keys[tid] = tid < n ? lg_key(dd[tid]) : 0u;
    perm[tid] = tid;
  }
  if (tid == 0) *flag = 0;
  __syncthreads();
  // bitonic sort, descending by key (pad keys 0 sink to the end; real keys are >= 1 unless NaN)
  for (int k = 2; k <= sz; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (tid < sz) {
        const int ixj = tid ^ j;
        if (ixj > tid) {
          const uint32_t a = keys[tid], b = keys[ixj];
          const bool desc = (tid & k) == 0;
          if (desc ? (a < b) : (a > b)) {
            keys[tid] = b; keys[ixj] = a;
            const int t = perm[tid]; perm[tid] = perm[ixj]; perm[ixj] = t;
          }
        }
      }
      __syncthreads();
    }
  }
  LG_T(10);
  // ties among real entries, or a NaN (key 0 inside the first n) -> exact replay
  if (tid < n) {
    const bool tie = (tid + 1 < n && keys[tid] == keys[tid + 1]) || keys[tid] == 0u;
    if (tie) *flag = 1;
  }
  __syncthreads();
  LG_T(11);
#ifdef TOB200_LG_TIMING
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    g_lg_tm[20] += *flag;
    if (*flag && g_lg_tm[21] < 3) {
      g_lg_tm[21]++;
      int shown = 0;
      for (int i = 0; i + 1 < n && shown < 4; ++i)
        if (keys[i] == keys[i + 1] || keys[i] == 0u) { printf("tie at sorted %d: keys %u %u orig %d %d dd %g %g\n", i, keys[i], keys[i + 1], perm[i], perm[i + 1], dd[perm[i]], dd[perm[i + 1]]); ++shown; }
    }
  }
#endif
  if (*flag) {
    // srt = the sorted originals (moved to spare scratch), perm is rebuilt as at[pos] (element now at
    // that position), inv as pos[orig]; keys stay sorted and delimit the tie groups
    int *srt = spare;
    if (tid < n) srt[tid] = perm[tid];
    __syncthreads();
    if (tid < n) { perm[tid] = tid; inv[tid] = tid; }
    __syncthreads();
    if (tid == 0) {
      if (keys[n - 1] == 0u) {
        // a NaN on the diagonal: replay Eigen's search literally (a NaN sitting at position k stays,
        // elsewhere it never wins); the factorisation fails anyway, speed is irrelevant
        for (int k = 0; k < n; ++k) {
          int pbest = k;
          uint32_t best = lg_key(dd[perm[k]]);
          if (best != 0u)
            for (int i = k + 1; i < n; ++i) {
              const uint32_t kv = lg_key(dd[perm[i]]);
              if (kv > best) { best = kv; pbest = i; }
            }
          const int ek = perm[k]; perm[k] = perm[pbest]; perm[pbest] = ek;
        }
        for (int k = 0; k < n; ++k) inv[perm[k]] = k;
      } else {
        uint32_t prev = 0u, cur = keys[0];
        for (int k = 0; k < n; ++k) {
          const uint32_t next = (k + 1 < n) ? keys[k + 1] : 0u;
          int e;
          if (cur != prev && cur != next) {
            e = srt[k];  // singleton: the k-th largest key
          } else {       // tie group: the untaken member that currently sits first
            int gs = k;
            while (gs > 0 && keys[gs - 1] == cur) --gs;
            e = -1;
            int best_pos = n, slot = gs;
            for (int i = gs; i < n && keys[i] == cur; ++i) {
              const int o = srt[i];
              if (o >= 0 && inv[o] < best_pos) { best_pos = inv[o]; e = o; slot = i; }
            }
            srt[slot] = -1;
          }
          const int pp = inv[e], f = perm[k];  // swap the contents of positions k and pp
          perm[k] = e; perm[pp] = f;
          inv[e] = k; inv[f] = pp;
          prev = cur; cur = next;
        }
      }
    }
    __syncthreads();
  } else {
    if (tid < n) inv[perm[tid]] = tid;
    __syncthreads();
  }
  LG_T(12);
}

// 16-byte asynchronous global -> shared copy (LDGSTS): the W tiles of phase 1 are double buffered
__device__ __forceinline__ void lg_cp_async16(float *dst, const float *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void lg_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void lg_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Packed FP32 pairs: Blackwell's FFMA2 (fma.rn.f32x2) does two IEEE fused multiply-adds per issue slot
// — same roundings as two scalar fmas, half the instructions in the FMA-issue-bound loops below.
__device__ __forceinline__ void lg_ffma2(unsigned long long &acc, float w, float b0, float b1) {
  unsigned long long a, b;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ float lg_lo(unsigned long long v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float lg_hi(unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ unsigned long long lg_pack(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}

// Blocked left-looking LDL^T of the permuted lower matrix W (pitch np) in global memory; D is left on
// the diagonal of W and in dsm.  Returns info()==Success && isPositive() (uniform across the CTA).
//
// Panel kb (32 columns), thread = row kb + tid:
//   phase 1   S_ic = sum_{j < kb} L_ij (D_j L_{kb+c,j}), j ascending: W tiles [rows x 32 j] stream through
//             a double-buffered shared tile (cp.async), T[j][c] = D_j L_{kb+c,j} is formed once per tile,
//             32 accumulators per thread.
//   phase 2a  the 32 x 32 diagonal block, by warp 0 alone, right-looking in registers: after column jj
//             is final every lane folds it into its pending sums S_r[c] (c > jj) — each S_r[c] still
//             receives its terms in ascending jj, i.e. the oracle's order — with T2[c][jj] = D_jj L_{c,jj}
//             exchanged by shuffles and left in shared memory for
//   phase 2b  the rows below the block: same recurrence, no further synchronisation.
__device__ bool lg_ldlt_factor(float *W, int n, int np, float *sm, const LgSolveSmem &L) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float *dsm = sm + L.dsm;
  float *tt = sm + L.tt;      // [32 j][32 c]: T of phase 1, then T2 of phase 2
  float *tile0 = sm + L.tile;  // 2 x [rows][kLgTilePitch]
  const int tile_elems = np * kLgTilePitch;
  int *misc = reinterpret_cast<int *>(sm + L.misc);  // 0: sign, 1: found_zero_pivot, 2: ret, 3: nonzero-below flag
  if (n == 1) {
    const float a = W[0];
    if (tid == 0) dsm[0] = a;
    __syncthreads();
    return !(a < 0.f);
  }
  if (tid == 0) { misc[0] = 0; misc[1] = 0; misc[2] = 1; misc[3] = 0; }
  __syncthreads();
  for (int kb = 0; kb < n; kb += kLgPanel) {
    const int pw = (n - kb < kLgPanel) ? (n - kb) : kLgPanel;  // panel width
    const int nrows = n - kb;
    const int i = kb + tid;  // my row
    const bool active = tid < nrows;
    float areg[kLgPanel], S[kLgPanel];
    // phase-1 accumulators: a 4 x 8 register tile (rows 4 rg .. 4 rg + 3 below kb, panel columns 8 cg .. 8 cg + 7)
    // as packed pairs, so that every L value feeds 8 and every T value 4 fmas (a 1 x 32 tile loaded 9 words
    // per 32 fmas and was LSU bound); the sums change hands to thread = row through shared memory afterwards
    const int rg = tid >> 2, cg = tid & 3;
    unsigned long long S2[4][4];
#pragma unroll
    for (int c = 0; c < kLgPanel; ++c) { areg[c] = 0.f; S[c] = 0.f; }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int h = 0; h < 4; ++h) S2[a][h] = 0ull;
    auto issue_tile = [&](int jc, float *tile) {
      for (int idx = tid; idx < nrows * (kLgPanel / 4); idx += kLgSolveThreads) {
        const int rr = idx / (kLgPanel / 4), q = idx % (kLgPanel / 4);
        lg_cp_async16(tile + rr * kLgTilePitch + 4 * q, W + (size_t)(kb + rr) * np + jc + 4 * q);
      }
      lg_cp_async_commit();
    };
    if (kb > 0) issue_tile(0, tile0);
    if (active) {
      const float4 *wr = reinterpret_cast<const float4 *>(W + (size_t)i * np + kb);
#pragma unroll
      for (int q = 0; q < kLgPanel / 4; ++q) {
        const float4 v = wr[q];
        areg[4 * q] = v.x; areg[4 * q + 1] = v.y; areg[4 * q + 2] = v.z; areg[4 * q + 3] = v.w;
      }
    }
    LG_T(13);
    // ---- phase 1 ----
    for (int jc = 0, tb = 0; jc < kb; jc += kLgPanel, tb ^= 1) {
      float *tile = tile0 + tb * tile_elems;
      lg_cp_async_wait<0>();
      __syncthreads();  // the tile has landed for everybody; the other buffer and tt are fully consumed
      if (jc + kLgPanel < kb) issue_tile(jc + kLgPanel, tile0 + (tb ^ 1) * tile_elems);  // overlaps the FMAs below
      for (int e = tid; e < kLgPanel * kLgPanel; e += kLgSolveThreads) {
        const int j = e / kLgPanel, c = e % kLgPanel;
        tt[j * kLgPanel + c] = (c < pw) ? __fmul_rn(dsm[jc + j], tile[c * kLgTilePitch + j]) : 0.f;
      }
      __syncthreads();
      if (4 * rg < nrows) {
        const float *lrow = tile + (4 * rg) * kLgTilePitch;
#pragma unroll 2
        for (int j4 = 0; j4 < kLgPanel; j4 += 4) {
          float4 w[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) w[a] = *reinterpret_cast<const float4 *>(lrow + a * kLgTilePitch + j4);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float4 *t4 = reinterpret_cast<const float4 *>(tt + (j4 + jj) * kLgPanel + 8 * cg);
            const float4 t0 = t4[0], t1 = t4[1];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const float wv = jj == 0 ? w[a].x : (jj == 1 ? w[a].y : (jj == 2 ? w[a].z : w[a].w));
              lg_ffma2(S2[a][0], wv, t0.x, t0.y);
              lg_ffma2(S2[a][1], wv, t0.z, t0.w);
              lg_ffma2(S2[a][2], wv, t1.x, t1.y);
              lg_ffma2(S2[a][3], wv, t1.z, t1.w);
            }
          }
        }
      }
    }
    __syncthreads();  // tt and both tile buffers are free
    if (kb > 0) {
      // hand the sums over: [row][33] (thread = row reads its 32 values conflict free)
      float *sbuf = tile0;
      if (4 * rg < nrows) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float *d = sbuf + (4 * rg + a) * (kLgPanel + 1) + 8 * cg + 2 * h;
            d[0] = lg_lo(S2[a][h]);
            d[1] = lg_hi(S2[a][h]);
          }
      }
      __syncthreads();
      if (active) {
#pragma unroll
        for (int c = 0; c < kLgPanel; ++c) S[c] = sbuf[tid * (kLgPanel + 1) + c];
      }
      __syncthreads();  // sbuf (== tile buffer 0) is reused by the next panel's first tile
    }
    LG_T(14);
    // ---- phase 2a: diagonal block (lanes < kLgPanel) and the rows of warp 0 below it, warp 0 ----
    if (warp == 0) {
      int sign = misc[0], fzp = misc[1], ret = misc[2], bad = 0;
#pragma unroll
      for (int jj = 0; jj < kLgPanel; ++jj) {
        if (jj < pw) {  // uniform
          const float d = __fsub_rn(areg[jj], S[jj]);
          const float akk = __shfl_sync(0xffffffffu, d, jj);
          const bool valid = fabsf(akk) > 0.f;
          float tm = 0.f;
          if (lane == jj) {
            areg[jj] = akk;
            dsm[kb + jj] = akk;
          } else if (lane > jj) {
            float v = d;
            if (valid) v = __fdiv_rn(d, akk);
            else if (v != 0.f && active) bad = 1;  // a zero pivot with a non-zero column below it
            areg[jj] = v;
            tm = __fmul_rn(akk, v);  // T2[lane][jj] = D_jj L_{lane,jj}
          }
          if (lane < kLgPanel) tt[jj * kLgPanel + lane] = tm;
          // Eigen LDLT bookkeeping (identical in every lane: akk is uniform)
          if (fzp && valid) ret = 0;
          else if (!valid) fzp = 1;
          if (sign == 1) { if (akk < 0.f) sign = 2; }
          else if (sign == -1) { if (akk > 0.f) sign = 2; }
          else if (sign == 0) { if (akk > 0.f) sign = 1; else if (akk < 0.f) sign = -1; }
#pragma unroll
          for (int c = jj + 1; c < kLgPanel; ++c) {
            // no "lane >= c" guard: S[c] of a row above column c is never read
            S[c] = __fmaf_rn(areg[jj], __shfl_sync(0xffffffffu, tm, c), S[c]);
          }
        }
      }
      bad = __any_sync(0xffffffffu, bad);
      if (lane == 0) {
        misc[0] = sign; misc[1] = fzp; misc[2] = ret;
        if (bad) misc[3] = 1;
      }
    }
    __syncthreads();
    LG_T(15);
    // ---- phase 2b: the rows below the block ----
    if (active && tid >= 32) {
#pragma unroll
      for (int jj = 0; jj < kLgPanel; ++jj) {
        if (jj < pw) {
          const float akk = dsm[kb + jj];
          float v = __fsub_rn(areg[jj], S[jj]);
          if (fabsf(akk) > 0.f) v = __fdiv_rn(v, akk);
          else if (v != 0.f) misc[3] = 1;
          areg[jj] = v;
#pragma unroll
          for (int c = jj + 1; c < kLgPanel; ++c) S[c] = __fmaf_rn(v, tt[jj * kLgPanel + c], S[c]);
        }
      }
    }
    LG_T(16);
    // ---- write the panel back (L below the diagonal, D on it) ----
    if (active) {
      float4 *wr = reinterpret_cast<float4 *>(W + (size_t)i * np + kb);
#pragma unroll
      for (int q = 0; q < kLgPanel / 4; ++q) wr[q] = make_float4(areg[4 * q], areg[4 * q + 1], areg[4 * q + 2], areg[4 * q + 3]);
    }
    __syncthreads();
    LG_T(17);
  }
  const bool ok = misc[2] && !misc[3] && (misc[0] == 1 || misc[0] == 0);
  __syncthreads();
  return ok;
}

// x (shared, original order) <- P^T L^-T D^+ L^-1 P b.  b in shared (original order).
__device__ void lg_ldlt_solve(const float *W, int n, int np, const int *perm, const float *b, float *xout, float *sm,
                              const LgSolveSmem &L) {
  const int tid = threadIdx.x, lane = tid & 31;
  float *ysm = sm + L.ysm;
  const float *dsm = sm + L.dsm;
  const int i = tid;
  const bool active = i < n;
  float yv = active ? b[perm[i]] : 0.f;
  // forward: y_i takes its updates in the order j = 0 .. i-1
  for (int jb = 0; jb < n; jb += kLgBlk) {
    // diagonal block rows are exactly the lanes of warp jb / 32
    if ((tid >> 5) == (jb >> 5)) {
      float lrow[kLgBlk];
      if (active) {
        const float4 *wr = reinterpret_cast<const float4 *>(W + (size_t)i * np + jb);
#pragma unroll
        for (int q = 0; q < kLgBlk / 4; ++q) {
          const float4 v = wr[q];
          lrow[4 * q] = v.x; lrow[4 * q + 1] = v.y; lrow[4 * q + 2] = v.z; lrow[4 * q + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < kLgBlk; ++q) lrow[q] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < kLgBlk; ++j) {
        const float yj = __shfl_sync(0xffffffffu, yv, j);
        if (lane > j && active) yv = __fmaf_rn(-lrow[j], yj, yv);
      }
      ysm[i] = yv;  // i < jb + 32 <= np
    }
    __syncthreads();
    if (active && i >= jb + kLgBlk) {
      const float4 *wr = reinterpret_cast<const float4 *>(W + (size_t)i * np + jb);
#pragma unroll
      for (int q = 0; q < kLgBlk / 4; ++q) {
        const float4 v = wr[q];
        const int j0 = jb + 4 * q;
        if (j0 < n) yv = __fmaf_rn(-v.x, ysm[j0], yv);
        if (j0 + 1 < n) yv = __fmaf_rn(-v.y, ysm[j0 + 1], yv);
        if (j0 + 2 < n) yv = __fmaf_rn(-v.z, ysm[j0 + 2], yv);
        if (j0 + 3 < n) yv = __fmaf_rn(-v.w, ysm[j0 + 3], yv);
      }
    }
  }
  LG_T(18);
  // D^+ (pseudo-inverse with tolerance min())
  if (active) {
    const float d = dsm[i];
    yv = (fabsf(d) > 1.175494351e-38f) ? __fdiv_rn(yv, d) : 0.f;
  }
  __syncthreads();
  // backward: y_i takes its updates in the order j = n-1 .. i+1
  const int last = ((n - 1) / kLgBlk) * kLgBlk;
  for (int jb = last; jb >= 0; jb -= kLgBlk) {
    if ((tid >> 5) == (jb >> 5)) {
#pragma unroll
      for (int j = kLgBlk - 1; j >= 0; --j) {
        const float yj = __shfl_sync(0xffffffffu, yv, j);
        if (lane < j && jb + j < n) yv = __fmaf_rn(-W[(size_t)(jb + j) * np + i], yj, yv);
      }
      ysm[i] = yv;
    }
    __syncthreads();
    if (i < jb) {
      const int jend = (n - jb < kLgBlk) ? (n - jb) : kLgBlk;
      for (int j = jend - 1; j >= 0; --j) yv = __fmaf_rn(-W[(size_t)(jb + j) * np + i], ysm[jb + j], yv);
    }
  }
  if (active) xout[perm[i]] = yv;
  __syncthreads();
  LG_T(19);
}

// deterministic CTA-wide sum of v[0..n) squared (fixed tree)
__device__ float lg_sqnorm(const float *v, int n, float *red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float s = (tid < n) ? __fmul_rn(v[tid], v[tid]) : 0.f;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, off));
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < kLgSolveThreads / 32; ++w) t = __fadd_rn(t, red[w]);
  __syncthreads();
  return t;
}

// lower triangle of H <- transpose of the upper one (32 x 32 tiles through per-warp shared tiles)
__device__ void lg_mirror_upper(float *H, int n, int np, float *scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *t = scratch + warp * (32 * 33);
  const int nb = (n + 31) / 32;
  const int ntiles = nb * (nb + 1) / 2;
  for (int e = warp; e < ntiles; e += kLgSolveThreads / 32) {
    int bi = 0, rem = e;  // tile (bi, bj), bi <= bj, enumerated row by row
    while (rem >= nb - bi) { rem -= nb - bi; ++bi; }
    const int bj = bi + rem;
    for (int r = 0; r < 32; ++r) t[r * 33 + lane] = H[(size_t)(32 * bi + r) * np + 32 * bj + lane];
    __syncwarp();
    for (int r = 0; r < 32; ++r) {
      // element (32 bj + r, 32 bi + lane) <- (32 bi + lane, 32 bj + r); diagonal tiles: strictly lower part only
      if (bi != bj || lane < r) H[(size_t)(32 * bj + r) * np + 32 * bi + lane] = t[lane * 33 + r];
    }
    __syncwarp();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kLgSolveThreads, 1) lg_solve_kernel(const __grid_constant__ LgSolveParams p) {
  extern __shared__ __align__(16) float sm[];
  const LgSolveSmem L = lg_solve_smem(p.np);
  const int tid = threadIdx.x;
  const int n = p.n, np = p.np;
  float *dd = sm + L.dd, *rhs = sm + L.rhs, *ysm = sm + L.ysm;
  int *perm = reinterpret_cast<int *>(sm + L.perm), *inv = reinterpret_cast<int *>(sm + L.inv);
  float *W = p.W + (size_t)blockIdx.x * np * np;
  __shared__ LmScalars<float> s;
  __shared__ int sh_act, sh_flag, sh_built_ok;
  __shared__ double sh_cost;
  unsigned long long local_active = 0;

#ifdef TOB200_LG_TIMING
  int lg_cnt = 0;
  if (tid == 0 && blockIdx.x == 0) for (int i = 0; i < 32; ++i) g_lg_tm[i] = 0;
#endif
  for (int64_t pr = blockIdx.x; pr < p.B; pr += gridDim.x) {
    LG_T0();
#ifdef TOB200_LG_TIMING
    ++lg_cnt;
#endif
    float *Hp = p.H + (size_t)pr * np * np;
    float *hd = p.hd ? p.hd + (size_t)pr * np : nullptr;
    float *gp = p.g ? p.g + (size_t)pr * n : nullptr;
    bool pass_rebuilt = true;
    __syncthreads();
    if (p.mode == 0) {
      if (tid == 0) s = p.rec[pr];
      __syncthreads();
      if (s.done()) continue;  // uniform
      pass_rebuilt = p.opt.solver_type != 0 || s.rebuild();
    }
    // ---- Build's tail: cost, validity, clipping, diagonal check (lm.h:69-86) ----
    double cost = 0.0;
    bool built_ok = true;
    if (p.mode == 0) {
      if (tid == 0) {
        double c;
        bool ok = lm_normalize_cost(p.opt, p.cost[pr], p.nres, c);
        if (pass_rebuilt) s.num_builds++;
        sh_cost = c;
        sh_built_ok = ok;
      }
      __syncthreads();
      cost = sh_cost;
      built_ok = sh_built_ok;
      if (pass_rebuilt && built_ok) {
        if (p.opt.grad_clipping != 0.f && tid < n) {
          float v = gp[tid];
          v = v < -p.opt.grad_clipping ? -p.opt.grad_clipping : v;
          v = v > p.opt.grad_clipping ? p.opt.grad_clipping : v;
          gp[tid] = v;
        }
        if (p.opt.check_min_H_diag > 0.f) {
          const int low = (tid < n) && (fabsf(p.dg ? p.dg[(size_t)pr * n + tid] : Hp[(size_t)tid * np + tid]) < p.opt.check_min_H_diag);
          if (__syncthreads_or(low)) built_ok = false;
        }
      }
    }
    // right-hand side: -grad (gn.h:155), or b
    if (tid < n) rhs[tid] = (p.mode == 2) ? p.b[(size_t)pr * n + tid] : (p.mode == 3 ? 0.f : -gp[tid]);
    __syncthreads();

    bool solver_failed = true, early_return = false;
    bool mirrored = (p.mode == 0) && !pass_rebuilt;  // a stale H_ was mirrored when it was built
    const uint8_t max_tries = p.mode == 0 ? lm_max_tries(p.opt) : 0;
    for (int attempt = 0;; ++attempt) {
      if (p.mode == 0 && !(s.num_consec_failures <= max_tries)) break;
      if (built_ok) {
        // damped diagonal (lm.h:108-117)
        double sc = 1.0;
        bool damp = false;
        if (p.mode == 0) damp = lm_damping_scale(s, p.opt, pass_rebuilt, sc);
        else if (p.mode == 1 && p.lambda) { const float lam = p.lambda[pr]; damp = lam > 0.f; sc = 1.0 + (double)lam; }
        if (tid < n) {
          const float base = (pass_rebuilt || !hd) ? (p.dg ? p.dg[(size_t)pr * n + tid] : Hp[(size_t)tid * np + tid]) : hd[tid];
          const float v = damp ? (float)((double)base * sc) : base;
          dd[tid] = v;
        }
        __syncthreads();
        if (hd && tid < n) hd[tid] = dd[tid];  // H_ keeps the damped diagonal
        LG_T(0);
        lg_pivot_order(dd, n, np, perm, inv, ysm, reinterpret_cast<int *>(sm + L.dsm), &sh_flag);
        LG_T(1);
        // H is made fully symmetric in place once per rebuild (the upper triangle is the canonical
        // one), so that a row of P H P^T is a gather from ONE row of H: W(a, b) = H(perm[a], perm[b])
        if (!mirrored) {
          lg_mirror_upper(Hp, n, np, sm + L.tile);
          mirrored = true;
        }
        LG_T(2);
        // row gather with eight independent loads in flight per lane (two rows x four 32-column chunks):
        // the gathers hit L2 at random, so memory-level parallelism is what this loop runs on
        for (int a0 = tid >> 5; a0 < n; a0 += 2 * (kLgSolveThreads / 32)) {
          const int a1 = a0 + kLgSolveThreads / 32;
          const bool two = a1 < n;
          const int ia0 = perm[a0], ia1 = two ? perm[a1] : ia0;
          const float *h0 = Hp + (size_t)ia0 * np, *h1 = Hp + (size_t)ia1 * np;
          const int last = two ? a1 : a0;  // a1 > a0: the longer row bounds the chunk loop
          for (int b0 = 0; b0 <= last; b0 += 128) {
            float v0[4], v1[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int b = b0 + 32 * q + (tid & 31);
              const int jb = (b <= last) ? perm[b] : 0;
              v0[q] = (b <= a0) ? ((ia0 == jb) ? dd[ia0] : h0[jb]) : 0.f;
              v1[q] = (two && b <= a1) ? ((ia1 == jb) ? dd[ia1] : h1[jb]) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int b = b0 + 32 * q + (tid & 31);
              if (b <= a0) W[(size_t)a0 * np + b] = v0[q];
              if (two && b <= a1) W[(size_t)a1 * np + b] = v1[q];
            }
          }
        }
        __syncthreads();
        // all-zero diagonal: success iff the strict triangle is zero as well (ZeroSign)
        bool ok;
        if (n > 1 && !(fabsf(dd[perm[0]]) > 0.f)) {
          int nz = 0;
          for (int a = tid >> 5; a < n; a += kLgSolveThreads / 32)
            for (int b = tid & 31; b < a; b += 32) nz |= (W[(size_t)a * np + b] != 0.f);
          ok = !__syncthreads_or(nz);
          if (tid < n) (sm + L.dsm)[tid] = 0.f;
          __syncthreads();
        } else {
          LG_T(3);
          ok = lg_ldlt_factor(W, n, np, sm, L);
          LG_T(4);
        }
        if (ok && p.mode == 3) {
          // InvCov (math.h:44-57): chol.solve(Identity), one unit vector at a time through the factor
          float best = -3.402823466e+38f;
          for (int c = 0; c < n; ++c) {
            if (tid < n) rhs[tid] = (tid == c) ? 1.f : 0.f;
            __syncthreads();
            lg_ldlt_solve(W, n, np, perm, rhs, dd, sm, L);
            if (tid < n) {
              const float v = dd[tid];
              if (p.dx) p.dx[((size_t)pr * n + tid) * n + c] = v;
              best = v > best ? v : best;
            }
            __syncthreads();
          }
          if (p.max_std) {  // MaxStdDev (solvers/lm.h:176-187): sqrt of the largest coefficient
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
              const float o = __shfl_xor_sync(0xffffffffu, best, off);
              best = o > best ? o : best;
            }
            if ((tid & 31) == 0) ysm[tid >> 5] = best;
            __syncthreads();
            if (tid == 0) {
              for (int w = 1; w < kLgSolveThreads / 32; ++w) best = ysm[w] > best ? ysm[w] : best;
              p.max_std[pr] = sqrtf(best);
            }
            __syncthreads();
          }
          solver_failed = false;
        } else if (ok) {
          lg_ldlt_solve(W, n, np, perm, rhs, dd, sm, L);  // dd is dead once W is laid out: dx goes there
          LG_T(5);
          solver_failed = false;
        }
      }
      if (!solver_failed) break;
      if (p.mode != 0) break;
      if (tid == 0) sh_act = lm_on_solver_failure(s, p.opt, cost, p.nres);
      __syncthreads();
      const int act = sh_act;
      if (act == kLmEarlyReturn) early_return = true;
      if (act != kLmRetry) break;
      if (attempt >= 100000) break;
    }
    const float *dxs = dd;  // the solution, original order
    if (p.mode != 0) {
      if (!solver_failed && tid < n && p.mode != 3) p.dx[(size_t)pr * n + tid] = dxs[tid];
      if (tid == 0) {
        p.status[pr] = solver_failed ? 1 : 0;
        if (p.mode == 3 && solver_failed && p.max_std) p.max_std[pr] = 0.f;
        if (p.mode == 1 && p.cost_out) p.cost_out[pr] = (double)p.cost[pr];
      }
      continue;
    }
    double dx_norm2 = 0.0, grad_norm2 = 0.0;
    if (!solver_failed) {
      dx_norm2 = (double)lg_sqnorm(dxs, n, ysm);
      if (p.opt.min_grad_norm2_f > 0.0f) {
        if (tid < n) rhs[tid] = gp[tid];
        __syncthreads();
        grad_norm2 = (double)lg_sqnorm(rhs, n, ysm);
      }
    }
    if (tid == 0) {
      bool success, has_dx;
      lm_finish_step(s, p.opt, early_return, solver_failed, cost, p.nres, dx_norm2, grad_norm2, success, has_dx);
      sh_act = lm_update_action(s, p.opt, success, has_dx);
    }
    __syncthreads();
    const int action = sh_act;
    if (tid < n) {
      float *xp = p.x + (size_t)pr * n, *lp = p.last_dx + (size_t)pr * n;
      if (action == kLmApplyDx || action == kLmProbeDx) {
        xp[tid] = __fadd_rn(xp[tid], dxs[tid]);
        lp[tid] = dxs[tid];
      } else if (action == kLmRollBack) {
        xp[tid] = __fadd_rn(xp[tid], -lp[tid]);
      }
    }
    if (tid == 0) {
      p.rec[pr] = s;
      if (s.done()) lm_write_result(s, &p.results[pr]);
      else local_active++;
    }
  }
#ifdef TOB200_LG_TIMING
  if (tid == 0 && blockIdx.x == 0)
    printf("lg_solve block 0: %d problems; kcycles/problem: pre %lld pivot %lld (sort %lld tiechk %lld walk %lld) mirror %lld gatherW %lld factor %lld (load %lld ph1 %lld ph2a %lld ph2b %lld wb %lld) solve %lld (fwd %lld bwd %lld) tie-path taken %lld\n", lg_cnt,
           g_lg_tm[0] / lg_cnt / 1000, g_lg_tm[1] / lg_cnt / 1000, g_lg_tm[10] / lg_cnt / 1000, g_lg_tm[11] / lg_cnt / 1000, g_lg_tm[12] / lg_cnt / 1000,
           g_lg_tm[2] / lg_cnt / 1000, g_lg_tm[3] / lg_cnt / 1000, g_lg_tm[4] / lg_cnt / 1000, g_lg_tm[13] / lg_cnt / 1000, g_lg_tm[14] / lg_cnt / 1000,
           g_lg_tm[15] / lg_cnt / 1000, g_lg_tm[16] / lg_cnt / 1000, g_lg_tm[17] / lg_cnt / 1000, g_lg_tm[5] / lg_cnt / 1000, g_lg_tm[18] / lg_cnt / 1000, g_lg_tm[19] / lg_cnt / 1000, g_lg_tm[20]);
#endif
  if (tid == 0 && local_active && p.n_active) atomicAdd(p.n_active, local_active);
}

}  // namespace tob200
