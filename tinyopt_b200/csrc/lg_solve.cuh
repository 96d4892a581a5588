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

#include "lg.cuh"

namespace tob200 {


// monotone key of |d| for the pivot search: 0 for NaN (never greater than anything)
__device__ __forceinline__ uint32_t lg_key(float d) {
  const float v = fabsf(d);
  return (v != v) ? 0u : __float_as_uint(v) + 1u;
}

// Pivot order of Eigen's LDLT for the diagonal dd[0..n): perm[pos] = original index, inv[orig] = pos.
// Distinct keys: descending sort (bitonic, CTA wide).  Ties or NaNs: exact sequential simulation of
// "first maximum wins + swap" by one thread (rare).
__device__ void lg_pivot_order(const float *dd, int n, int np, int *perm, int *inv, float *scratch_keys, int *flag) {
  const int tid = threadIdx.x;
  uint32_t *keys = reinterpret_cast<uint32_t *>(scratch_keys);  // np entries
  // sort size: next power of two >= n (<= 512 == blockDim)
  int sz = 1;
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
  // ties among real entries, or a NaN (key 0 inside the first n) -> exact path
  if (tid < n) {
    const bool tie = (tid + 1 < n && keys[tid] == keys[tid + 1]) || keys[tid] == 0u;
    if (tie) *flag = 1;
  }
  __syncthreads();
  if (*flag) {
    if (tid == 0) {
      for (int i = 0; i < n; ++i) { keys[i] = lg_key(dd[i]); perm[i] = i; }
      for (int k = 0; k < n; ++k) {
        int pbest = k;
        uint32_t best = keys[k];
        if (best != 0u) {  // a NaN at k stays: nothing compares greater
          for (int i = k + 1; i < n; ++i)
            if (keys[i] > best) { best = keys[i]; pbest = i; }
        }
        if (pbest != k) {
          const uint32_t tk = keys[k]; keys[k] = keys[pbest]; keys[pbest] = tk;
          const int tp = perm[k]; perm[k] = perm[pbest]; perm[pbest] = tp;
        }
      }
    }
    __syncthreads();
  }
  if (tid < n) inv[perm[tid]] = tid;
  __syncthreads();
}

// Blocked left-looking LDL^T of the permuted lower matrix W (pitch np) in global memory; D is left on
// the diagonal of W and in dsm.  Returns info()==Success && isPositive() (uniform across the CTA).
__device__ bool lg_ldlt_factor(float *W, int n, int np, float *sm, const LgSolveSmem &L) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float *dsm = sm + L.dsm;
  float *temp = sm + L.temp;  // [2][32]
  float *tt = sm + L.tt;      // [32 j][32 c]
  float *tile = sm + L.tile;  // [rows][33]
  int *misc = reinterpret_cast<int *>(sm + L.misc);  // 0: sign, 1: found_zero_pivot, 2: ret, 3: nonzero-below flag
  float *piv = sm + L.misc + 8;                       // pivot broadcast [2]
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
#pragma unroll
    for (int c = 0; c < kLgPanel; ++c) { areg[c] = 0.f; S[c] = 0.f; }
    if (active) {
      const float4 *wr = reinterpret_cast<const float4 *>(W + (size_t)i * np + kb);
#pragma unroll
      for (int q = 0; q < kLgPanel / 4; ++q) {
        const float4 v = wr[q];
        areg[4 * q] = v.x; areg[4 * q + 1] = v.y; areg[4 * q + 2] = v.z; areg[4 * q + 3] = v.w;
      }
    }
    // ---- phase 1: S_ic = sum_{j < kb} W_ij (D_j W_{kb+c, j}), j ascending ----
    for (int jc = 0; jc < kb; jc += kLgPanel) {
      __syncthreads();  // previous tile / tt fully consumed
      for (int rr = warp; rr < nrows; rr += kLgSolveThreads / 32)
        tile[rr * (kLgPanel + 1) + lane] = W[(size_t)(kb + rr) * np + jc + lane];
      __syncthreads();
      for (int e = tid; e < kLgPanel * kLgPanel; e += kLgSolveThreads) {
        const int j = e >> 5, c = e & 31;
        tt[j * kLgPanel + c] = (c < pw) ? __fmul_rn(dsm[jc + j], tile[c * (kLgPanel + 1) + j]) : 0.f;
      }
      __syncthreads();
      if (active) {
        const float *trow = tile + tid * (kLgPanel + 1);
#pragma unroll 4
        for (int j = 0; j < kLgPanel; ++j) {
          const float w = trow[j];
          const float4 *t4 = reinterpret_cast<const float4 *>(tt + j * kLgPanel);
#pragma unroll
          for (int q = 0; q < kLgPanel / 4; ++q) {
            const float4 tv = t4[q];
            S[4 * q] = __fmaf_rn(w, tv.x, S[4 * q]);
            S[4 * q + 1] = __fmaf_rn(w, tv.y, S[4 * q + 1]);
            S[4 * q + 2] = __fmaf_rn(w, tv.z, S[4 * q + 2]);
            S[4 * q + 3] = __fmaf_rn(w, tv.w, S[4 * q + 3]);
          }
        }
      }
    }
    __syncthreads();
    // ---- phase 2: the panel, column by column (k = kb + c) ----
#pragma unroll
    for (int c = 0; c < kLgPanel; ++c) {
      if (c < pw) {  // uniform
        float *tb = temp + (c & 1) * kLgPanel;
        if (tid == c) {  // the diagonal row finishes its own chain and publishes temp, pivot
          float s = S[c];
#pragma unroll
          for (int jj = 0; jj < c; ++jj) {
            const float tv = __fmul_rn(dsm[kb + jj], areg[jj]);
            tb[jj] = tv;
            s = __fmaf_rn(areg[jj], tv, s);
          }
          const float akk = __fsub_rn(areg[c], s);
          areg[c] = akk;
          dsm[kb + c] = akk;
          piv[c & 1] = akk;
          // sign / zero-pivot bookkeeping (Eigen LDLT: m_sign, found_zero_pivot, ret)
          const bool valid = fabsf(akk) > 0.f;
          if (misc[1] && valid) misc[2] = 0;
          else if (!valid) misc[1] = 1;
          int sign = misc[0];
          if (sign == 1) { if (akk < 0.f) sign = 2; }
          else if (sign == -1) { if (akk > 0.f) sign = 2; }
          else if (sign == 0) { if (akk > 0.f) sign = 1; else if (akk < 0.f) sign = -1; }
          misc[0] = sign;
        }
        __syncthreads();
        if (active && tid > c) {
          const float akk = piv[c & 1];
          float s = S[c];
#pragma unroll
          for (int jj = 0; jj < c; ++jj) s = __fmaf_rn(areg[jj], tb[jj], s);
          float v = __fsub_rn(areg[c], s);
          if (fabsf(akk) > 0.f) v = __fdiv_rn(v, akk);
          else if (v != 0.f) misc[3] = 1;  // a zero pivot with a non-zero column below it
          areg[c] = v;
        }
      }
    }
    // ---- write the panel back (L below the diagonal, D on it) ----
    if (active) {
      float4 *wr = reinterpret_cast<float4 *>(W + (size_t)i * np + kb);
#pragma unroll
      for (int q = 0; q < kLgPanel / 4; ++q) wr[q] = make_float4(areg[4 * q], areg[4 * q + 1], areg[4 * q + 2], areg[4 * q + 3]);
    }
    __syncthreads();
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
  for (int jb = 0; jb < n; jb += kLgPanel) {
    // diagonal block rows are exactly the lanes of warp jb / 32
    if ((tid >> 5) == (jb >> 5)) {
      float lrow[kLgPanel];
      if (active) {
        const float4 *wr = reinterpret_cast<const float4 *>(W + (size_t)i * np + jb);
#pragma unroll
        for (int q = 0; q < kLgPanel / 4; ++q) {
          const float4 v = wr[q];
          lrow[4 * q] = v.x; lrow[4 * q + 1] = v.y; lrow[4 * q + 2] = v.z; lrow[4 * q + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < kLgPanel; ++q) lrow[q] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < kLgPanel; ++j) {
        const float yj = __shfl_sync(0xffffffffu, yv, j);
        if (lane > j && active) yv = __fmaf_rn(-lrow[j], yj, yv);
      }
      ysm[i] = yv;  // i < jb + 32 <= np
    }
    __syncthreads();
    if (active && i >= jb + kLgPanel) {
      const float4 *wr = reinterpret_cast<const float4 *>(W + (size_t)i * np + jb);
#pragma unroll
      for (int q = 0; q < kLgPanel / 4; ++q) {
        const float4 v = wr[q];
        const int j0 = jb + 4 * q;
        if (j0 < n) yv = __fmaf_rn(-v.x, ysm[j0], yv);
        if (j0 + 1 < n) yv = __fmaf_rn(-v.y, ysm[j0 + 1], yv);
        if (j0 + 2 < n) yv = __fmaf_rn(-v.z, ysm[j0 + 2], yv);
        if (j0 + 3 < n) yv = __fmaf_rn(-v.w, ysm[j0 + 3], yv);
      }
    }
  }
  // D^+ (pseudo-inverse with tolerance min())
  if (active) {
    const float d = dsm[i];
    yv = (fabsf(d) > 1.175494351e-38f) ? __fdiv_rn(yv, d) : 0.f;
  }
  __syncthreads();
  // backward: y_i takes its updates in the order j = n-1 .. i+1
  const int last = ((n - 1) / kLgPanel) * kLgPanel;
  for (int jb = last; jb >= 0; jb -= kLgPanel) {
    if ((tid >> 5) == (jb >> 5)) {
#pragma unroll
      for (int j = kLgPanel - 1; j >= 0; --j) {
        const float yj = __shfl_sync(0xffffffffu, yv, j);
        if (lane < j && jb + j < n) yv = __fmaf_rn(-W[(size_t)(jb + j) * np + i], yj, yv);
      }
      ysm[i] = yv;
    }
    __syncthreads();
    if (i < jb) {
      const int jend = (n - jb < kLgPanel) ? (n - jb) : kLgPanel;
      for (int j = jend - 1; j >= 0; --j) yv = __fmaf_rn(-W[(size_t)(jb + j) * np + i], ysm[jb + j], yv);
    }
  }
  if (active) xout[perm[i]] = yv;
  __syncthreads();
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

  for (int64_t pr = blockIdx.x; pr < p.B; pr += gridDim.x) {
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
    if (tid < n) rhs[tid] = (p.mode == 2) ? p.b[(size_t)pr * n + tid] : -gp[tid];
    __syncthreads();

    bool solver_failed = true, early_return = false;
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
        lg_pivot_order(dd, n, np, perm, inv, ysm, &sh_flag);
        // W <- P H P^T, lower triangle: W(a, b) = H(min(i,j), max(i,j)), i = perm[a], j = perm[b]
        for (int a = tid >> 5; a < n; a += kLgSolveThreads / 32) {
          const int ia = perm[a];
          for (int b = tid & 31; b <= a; b += 32) {
            const int jb = perm[b];
            const float v = (ia == jb) ? dd[ia] : (ia < jb ? Hp[(size_t)ia * np + jb] : Hp[(size_t)jb * np + ia]);
            W[(size_t)a * np + b] = v;
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
          ok = lg_ldlt_factor(W, n, np, sm, L);
        }
        if (ok) {
          lg_ldlt_solve(W, n, np, perm, rhs, dd, sm, L);  // dd is dead once W is laid out: dx goes there
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
      if (!solver_failed && tid < n) p.dx[(size_t)pr * n + tid] = dxs[tid];
      if (tid == 0) {
        p.status[pr] = solver_failed ? 1 : 0;
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
  if (tid == 0 && local_active && p.n_active) atomicAdd(p.n_active, local_active);
}

}  // namespace tob200
