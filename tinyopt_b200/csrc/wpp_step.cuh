// wpp_step.cuh — the SolverType seam (tob200_solver_*) for the warp-per-problem family: ONE Step of
// Optimizer_::OptimizeAcc per call from materialised residual blocks J, r, with the per-problem state
// of the optimizer (scalars, x, last_dx, the persistent damped H_ and grad_ of solvers/gn.h:200-201) in
// HBM between calls.  Same pass and same wpp_after_pass as the device-resident loop, so a host-driven
// run is bit-identical to tob200_lm_run_* for the same residual blocks.
// Reference: optimizers/optimizer.h:266-309 (the loop body), :332-539 (Step).
#pragma once

#include "tpp.cuh"  // StateRec
#include "wpp.cuh"

namespace tob200 {

template <typename T>
struct WppStepParams {
  WppData<T> d;
  DevOptions<T> opt;
  StateRec<T> *rec;  // [B]
  T *x;              // [B][n]
  T *last_dx;        // [B][n]
  T *H;              // [B][NP * LDW]: H_ as wpp_store_permuted leaves it (element (row, col), row <= col, at col * LDW + row)
  T *g;              // [B][NP]
  int32_t *needs;    // [B]
  unsigned long long *n_active;  // device counter, zeroed by the host before the launch
  int reset;         // 1: initialise the state instead of stepping (x already holds x0)
  // tob200_solver_step_hg_*: caller-filled accumulators (see TppStepParams)
  const T *hg_grad, *hg_H;
  const double *hg_cost;
  const int32_t *hg_nres;
};

template <typename T>
__device__ __forceinline__ void wpp_state_load(LmScalars<T> &s, const StateRec<T> &r) {
  s.final_cost = r.final_cost; s.final_rerr_dec = r.final_rerr_dec;
  s.lambda = r.lambda; s.prev_lambda = r.prev_lambda; s.bad_factor = r.bad_factor;
  s.final_nres = r.final_nres; s.stop_reason = r.stop_reason; s.flags = r.flags;
  s.num_builds = r.num_builds; s.iter = r.iter; s.num_iters = r.num_iters;
  s.num_failures = r.num_failures; s.num_consec_failures = r.num_consec_failures;
}
template <typename T>
__device__ __forceinline__ void wpp_state_store(const LmScalars<T> &s, StateRec<T> &r) {
  r.final_cost = s.final_cost; r.final_rerr_dec = s.final_rerr_dec;
  r.lambda = s.lambda; r.prev_lambda = s.prev_lambda; r.bad_factor = s.bad_factor;
  r.final_nres = s.final_nres; r.stop_reason = s.stop_reason; r.flags = s.flags;
  r.num_builds = s.num_builds; r.iter = s.iter; r.num_iters = s.num_iters;
  r.num_failures = s.num_failures; r.num_consec_failures = s.num_consec_failures;
}

template <typename T, int NB, int BLK, bool INV = false>
__global__ void __launch_bounds__(kWppThreads, sizeof(T) == 4 ? 2 : 1) wpp_step_kernel(const __grid_constant__ WppStepParams<T> p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x / 32;
  unsigned char *ws = smem + (size_t)wid * p.d.L.total;
  const int n = p.d.n;
  constexpr int NP = NB * BLK, LDW = wpp_ldw(NP);
  WppPipe<T> pipe;
  pipe.init(ws, p.d, lane);
  T *xs = reinterpret_cast<T *>(ws + p.d.L.xs);
  T *last_dx = reinterpret_cast<T *>(ws + p.d.L.last_dx);
  T *g = reinterpret_cast<T *>(ws + p.d.L.g);
  int bi, bj;
  bool has_block;
  wpp_block_of_lane<NB>(lane, bi, bj, has_block);
  const bool is_lm = p.opt.solver_type == 0;
  unsigned long long local_active = 0;  // counted by lane 0

  for (int64_t pr = wpp_next(p.d.counter, lane); pr < p.d.B; pr = wpp_next(p.d.counter, lane)) {
    LmScalars<T> s;
    if (p.reset) {
      s.reset_scalars(p.opt);
      if (lane == 0) {
        wpp_state_store(s, p.rec[pr]);
        p.needs[pr] = 1;
        local_active++;
      }
      for (int j = lane; j < n; j += 32) p.last_dx[pr * n + j] = (T)0;
      continue;
    }
    wpp_state_load(s, p.rec[pr]);
    if (s.done()) continue;  // warp uniform: a finished problem costs no traffic
    for (int j = lane; j < NP; j += 32) {
      xs[j] = j < n ? p.x[pr * n + j] : (T)0;
      last_dx[j] = j < n ? p.last_dx[pr * n + j] : (T)0;
      g[j] = j < n ? p.g[pr * NP + j] : (T)0;  // grad_ of the last rebuild: a cost-only Step re-solves with it
    }
    __syncwarp();
    const bool do_rebuild = !is_lm || s.rebuild();  // GN's Build always re-accumulates (gn.h:118-131)
    T acc[BLK][BLK], cost_only;
    T *hp = p.H + (size_t)pr * (NP * LDW);
    if (p.hg_cost) {
      // user-filled accumulators (docs/API.md:37-57,137-170): the lane's block of the augmented matrix
      // [H g; . cost] straight from the caller's H (upper triangle only is read) and grad
      cost_only = (T)0;
      const double cost_d = p.hg_cost[pr];
      const T *Hu = p.hg_H + (size_t)pr * n * n, *gu = p.hg_grad + (size_t)pr * n;
#pragma unroll
      for (int u = 0; u < BLK; ++u)
#pragma unroll
        for (int v = 0; v < BLK; ++v) {
          const int row = bi * BLK + u, col = bj * BLK + v;
          T val = (T)0;
          if (do_rebuild && has_block) {
            if (row <= col && col < n) val = Hu[row * n + col];
            else if (row < n && col == n) val = gu[row];
          }
          acc[u][v] = val;
        }
      wpp_after_pass<T, NB, BLK, INV>(s, p.opt, p.d, ws, hp, do_rebuild, bi, bj, has_block, acc, cost_only, lane, true, &cost_d,
                                      p.hg_nres[pr]);
    } else {
      wpp_pass<T, NB, BLK, false>(pipe, p.d, ws, pr, lane, do_rebuild, (T)0, (T)0, bi, bj, has_block, acc, cost_only);
      wpp_after_pass<T, NB, BLK, INV>(s, p.opt, p.d, ws, hp, do_rebuild, bi, bj, has_block, acc, cost_only, lane, true);
    }
    for (int j = lane; j < n; j += 32) {
      p.x[pr * n + j] = xs[j];
      p.last_dx[pr * n + j] = last_dx[j];
      p.g[pr * NP + j] = g[j];
    }
    if (lane == 0) {
      wpp_state_store(s, p.rec[pr]);
      p.needs[pr] = s.done() ? -1 : ((s.rebuild() || !is_lm) ? 1 : 0);
      if (!s.done()) local_active++;
    }
    __syncwarp();
  }
  if (lane == 0 && local_active) atomicAdd(p.n_active, local_active);
}

// un-damped final Hessian (solvers/lm.h:157-171) from the persistent H_ of the step solver
template <typename T>
__global__ void wpp_final_hessian_kernel(const T *H, const StateRec<T> *rec, int solver_type, int64_t B, int n, int np,
                                         int ldw, double *out) {
  const int64_t pr = blockIdx.x;
  if (pr >= B) return;
  const T *h = H + (size_t)pr * np * ldw;
  const T pl = rec[pr].prev_lambda;
  const bool undamp = solver_type == 0 && pl > (T)0;
  const T sc = Ops<T>::add((T)1, pl);
  double *o = out + (size_t)pr * n * n;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int j = e / n, k = e - j * n;
    if (j > k) continue;
    T v = h[k * ldw + j];
    if (j == k && undamp) v = Ops<T>::div(v, sc);
    o[j * n + k] = (double)v;
    o[k * n + j] = (double)v;
  }
}

}  // namespace tob200
