// tpp.cuh — "thread per problem" kernels for small n (n <= 12 float, n <= 8 double).
//
// Mapping: one warp owns one TILE32 tile (32 problems, lane l = problem 32*tile + l).  The rows of
// the residual blocks are streamed HBM -> shared memory by TMA bulk copies (cp.async.bulk, one
// contiguous copy per row chunk because TILE32 interleaves the 32 problems) through a warp-private
// multi-stage mbarrier ring; every lane then reads its own column of the stage (stride-1 across lanes:
// conflict-free) and accumulates its problem's cost, g = J^T r and the upper triangle of
// H = J^T J in registers, rows in order i = 0..m-1, one fma per term (the canonical op sequence,
// identical to the CPU oracle's).  Damping, the pivoted LDL^T, the solve and the whole LM state
// machine then run per lane in registers (lm_state.cuh, ldlt_reg.cuh) — H never touches HBM.
//
// Reference path replaced: the product site (diff/optimize_autodiff.h:151-157), SolverLM::Build /
// SolverGN::Solve (solvers/lm.h:60-120, solvers/gn.h:150-171), SolveLDLT (math.h:232-240),
// Optimizer_::Step / OptimizeAcc (optimizers/optimizer.h:243-539).
#pragma once

#include "common.cuh"
#include "lm_state.cuh"

namespace tob200 {

constexpr int kTppMaxStages = 8;
constexpr int kTppBarBytes = 128;  // up to 8 mbarriers, padded so the stages stay 128-B aligned
constexpr int kTppThreads = 128;   // 4 warps per CTA, each warp independent

// shared-memory layout of one warp: [mbarriers | stages | persistent H_ + grad_ (lane-interleaved)]
__host__ __device__ inline size_t tpp_stage_bytes(int n, int rows, size_t elt) { return (size_t)rows * (n + 1) * kTile * elt; }
__host__ __device__ inline size_t tpp_warp_smem_bytes(int n, int rows, int stages, size_t elt) {
  size_t b = kTppBarBytes + (size_t)stages * tpp_stage_bytes(n, rows, elt) + (size_t)(tri_count(n) + n) * kTile * elt;
  return (b + 127) & ~(size_t)127;
}

// Resident CTAs per SM each kernel is compiled for (__launch_bounds__): fixes the register budget
// (65536 / (128 * k)).  Chosen from ptxas -v so that nothing spills.
#ifndef TOB200_TPP_MINB_F64_MID
#define TOB200_TPP_MINB_F64_MID 3  // resident CTAs per SM for double, 3 <= n <= 6 (C2)
#endif
template <typename T, int N>
__host__ __device__ constexpr int tpp_min_blocks() {
  if (sizeof(T) == 8) return N <= 2 ? 4 : (N <= 6 ? TOB200_TPP_MINB_F64_MID : 2);
  return N <= 6 ? 4 : (N <= 8 ? 3 : 2);
}

template <typename T>
struct TppData {      // one batch of residual blocks in TILE32 layout
  const T *J;         // [ntiles][m][n][32]  (A for the polynomial family)
  const T *r;         // [ntiles][m][32]     (y for the polynomial family)
  int64_t B;          // problems
  int64_t ntiles;     // ceil(B / 32)
  int m, rows;        // residuals per problem, rows per pipeline stage
  int stages;         // pipeline depth (2..kTppMaxStages)
  uint32_t warp_smem; // bytes of shared memory per warp
  unsigned long long *tile_counter;  // dynamic tile scheduler, zeroed before the launch
};

template <typename T, int N>
struct SmemHG {  // persistent H_ / grad_ of the 32 problems of a warp, lane-interleaved
  T *base;
  int lane;
  __device__ __forceinline__ T ld_h(int i) const { return base[i * kTile + lane]; }
  __device__ __forceinline__ void st_h(int i, T v) { base[i * kTile + lane] = v; }
  __device__ __forceinline__ T ld_g(int j) const { return base[(tri_count(N) + j) * kTile + lane]; }
  __device__ __forceinline__ void st_g(int j, T v) { base[(tri_count(N) + j) * kTile + lane] = v; }
};

template <typename T, int N>
struct GlobalHG {  // same, in global memory (tile-interleaved) for the host-driven step kernel
  T *h;            // [ntiles][NT][32] + lane offset applied
  T *g;            // [ntiles][N][32]
  __device__ __forceinline__ T ld_h(int i) const { return h[i * kTile]; }
  __device__ __forceinline__ void st_h(int i, T v) { h[i * kTile] = v; }
  __device__ __forceinline__ T ld_g(int j) const { return g[j * kTile]; }
  __device__ __forceinline__ void st_g(int j, T v) { g[j * kTile] = v; }
};

// warp-private TMA ring
template <typename T, int N>
struct TppPipe {
  uint64_t *bars;
  unsigned char *stages;
  uint32_t stage_bytes;
  uint32_t r_off;    // element offset of the r / y rows inside a stage
  uint32_t nstages;
  uint32_t stage;    // next stage to consume
  uint32_t phase;    // its mbarrier phase parity

  __device__ __forceinline__ void init(unsigned char *warp_smem, int rows, int nst, int lane) {
    bars = reinterpret_cast<uint64_t *>(warp_smem);
    stages = warp_smem + kTppBarBytes;
    stage_bytes = (uint32_t)tpp_stage_bytes(N, rows, sizeof(T));
    r_off = (uint32_t)rows * N * kTile;
    nstages = (uint32_t)nst;
    stage = 0;
    phase = 0;
    if (lane == 0) {
      for (uint32_t s = 0; s < nstages; ++s) mbar_init(&bars[s], 1);
      mbar_fence_init();
    }
    __syncwarp();
  }
  __device__ __forceinline__ T *stage_ptr(uint32_t st) const {
    return reinterpret_cast<T *>(stages + (size_t)st * stage_bytes);
  }
  __device__ __forceinline__ void advance() {
    if (++stage == nstages) {
      stage = 0;
      phase ^= 1u;
    }
  }
  // lane 0 only: fill stage `st` with `nrows` rows starting at row `row0` of the tile whose J / r
  // rows start at jt / rt
  __device__ __forceinline__ void issue(const T *jt, const T *rt, int row0, int nrows, uint32_t st) {
    const uint32_t bj = (uint32_t)nrows * (N * kTile * (uint32_t)sizeof(T));
    const uint32_t br = (uint32_t)nrows * (kTile * (uint32_t)sizeof(T));
    T *sj = stage_ptr(st);
    mbar_expect_tx(&bars[st], bj + br);
    tma_bulk_g2s(sj, jt + (uint32_t)row0 * (N * kTile), bj, &bars[st]);
    tma_bulk_g2s(sj + r_off, rt + (uint32_t)row0 * kTile, br, &bars[st]);
  }
};

// ---- the per-row arithmetic (canonical op sequence) -------------------------------
// kSynth: (a, yv) are a row of A and y of the polynomial family; the lane evaluates
// r_i = t (1 + alpha t^2) - y_i and J_i = (1 + 3 alpha t^2) a_i at its x.
// Otherwise (a, yv) are the row of J and r themselves.
template <typename T, int N, bool kSynth>
__device__ __forceinline__ void tpp_row_residual(T (&a)[N], T yv, const T (&x)[N], T alpha, T alpha3, bool want_j,
                                                 T &ri) {
  using O = Ops<T>;
  if (kSynth) {
    T t = (T)0;
#pragma unroll
    for (int j = 0; j < N; ++j) t = O::fma(a[j], x[j], t);
    const T t2 = O::mul(t, t);
    ri = O::fma(t, O::fma(alpha, t2, (T)1), -yv);
    if (want_j) {
      const T sc = O::fma(alpha3, t2, (T)1);
#pragma unroll
      for (int j = 0; j < N; ++j) a[j] = O::mul(sc, a[j]);
    }
  } else {
    ri = yv;
  }
}

template <typename T, int N>
__device__ __forceinline__ void tpp_row_accumulate(const T (&a)[N], T ri, T (&hu)[tri_count(N)], T (&g)[N]) {
  using O = Ops<T>;
#pragma unroll
  for (int j = 0; j < N; ++j) g[j] = O::fma(a[j], ri, g[j]);
#pragma unroll
  for (int j = 0; j < N; ++j) {
#pragma unroll
    for (int k = j; k < N; ++k) hu[tri_index(N, j, k)] = O::fma(a[j], a[k], hu[tri_index(N, j, k)]);
  }
}

// One streaming pass over the m rows of a tile: cost, and (do_rebuild) g and the upper triangle of H.
// Rows are consumed two at a time so that the two dependent t-chains interleave (ILP); the
// accumulation order into every sum is still row 0, 1, 2, ...
template <typename T, int N, bool kSynth>
__device__ __forceinline__ void tpp_pass(TppPipe<T, N> &pipe, const TppData<T> &d, int64_t tile, int lane,
                                         bool active, bool do_rebuild, const T (&x)[N], T alpha, T alpha3,
                                         T (&hu)[tri_count(N)], T (&g)[N], T &cost) {
  using O = Ops<T>;
  constexpr int NT = tri_count(N);
#pragma unroll
  for (int i = 0; i < NT; ++i) hu[i] = (T)0;  // solvers/gn.h:77-81 clear()
#pragma unroll
  for (int j = 0; j < N; ++j) g[j] = (T)0;
  cost = (T)0;

  const int m = d.m, R = d.rows;
  const int nchunks = (m + R - 1) / R;
  const T *jt = d.J + (size_t)tile * m * (N * kTile);
  const T *rt = d.r + (size_t)tile * m * kTile;
  if (lane == 0) {
    fence_proxy_async();
    const int pre = nchunks < (int)pipe.nstages ? nchunks : (int)pipe.nstages;
    uint32_t st = pipe.stage;
    for (int c = 0; c < pre; ++c) {
      const int row0 = c * R;
      pipe.issue(jt, rt, row0, (m - row0 < R) ? (m - row0) : R, st);
      if (++st == pipe.nstages) st = 0;
    }
  }
  for (int c = 0; c < nchunks; ++c) {
    mbar_wait(&pipe.bars[pipe.stage], pipe.phase);
    const int row0 = c * R;
    const int nrows = (m - row0 < R) ? (m - row0) : R;
    if (active) {
      const T *sj = pipe.stage_ptr(pipe.stage) + lane;
      const T *sr = sj + pipe.r_off;
      if (do_rebuild) {
        int rr = 0;
#ifdef TOB200_TPP_ROWS4
        for (; rr + 4 <= nrows; rr += 4) {  // four rows in flight: four independent t-chains
          T a0[N], a1[N], a2[N], a3[N];
#pragma unroll
          for (int j = 0; j < N; ++j) {
            a0[j] = sj[(rr * N + j) * kTile];
            a1[j] = sj[((rr + 1) * N + j) * kTile];
            a2[j] = sj[((rr + 2) * N + j) * kTile];
            a3[j] = sj[((rr + 3) * N + j) * kTile];
          }
          T r0, r1, r2, r3;
          tpp_row_residual<T, N, kSynth>(a0, sr[rr * kTile], x, alpha, alpha3, true, r0);
          tpp_row_residual<T, N, kSynth>(a1, sr[(rr + 1) * kTile], x, alpha, alpha3, true, r1);
          tpp_row_residual<T, N, kSynth>(a2, sr[(rr + 2) * kTile], x, alpha, alpha3, true, r2);
          tpp_row_residual<T, N, kSynth>(a3, sr[(rr + 3) * kTile], x, alpha, alpha3, true, r3);
          cost = O::fma(r0, r0, cost);
          cost = O::fma(r1, r1, cost);
          cost = O::fma(r2, r2, cost);
          cost = O::fma(r3, r3, cost);
          tpp_row_accumulate<T, N>(a0, r0, hu, g);
          tpp_row_accumulate<T, N>(a1, r1, hu, g);
          tpp_row_accumulate<T, N>(a2, r2, hu, g);
          tpp_row_accumulate<T, N>(a3, r3, hu, g);
        }
#endif
        for (; rr + 2 <= nrows; rr += 2) {
          T a0[N], a1[N];
#pragma unroll
          for (int j = 0; j < N; ++j) {
            a0[j] = sj[(rr * N + j) * kTile];
            a1[j] = sj[((rr + 1) * N + j) * kTile];
          }
          const T y0 = sr[rr * kTile], y1 = sr[(rr + 1) * kTile];
          T r0, r1;
          tpp_row_residual<T, N, kSynth>(a0, y0, x, alpha, alpha3, true, r0);
          tpp_row_residual<T, N, kSynth>(a1, y1, x, alpha, alpha3, true, r1);
          cost = O::fma(r0, r0, cost);
          cost = O::fma(r1, r1, cost);
          tpp_row_accumulate<T, N>(a0, r0, hu, g);
          tpp_row_accumulate<T, N>(a1, r1, hu, g);
        }
        if (rr < nrows) {
          T a0[N];
#pragma unroll
          for (int j = 0; j < N; ++j) a0[j] = sj[(rr * N + j) * kTile];
          T r0;
          tpp_row_residual<T, N, kSynth>(a0, sr[rr * kTile], x, alpha, alpha3, true, r0);
          cost = O::fma(r0, r0, cost);
          tpp_row_accumulate<T, N>(a0, r0, hu, g);
        }
      } else {  // cost-only pass (solvers/gn.h:98-105 Evaluate)
        for (int rr = 0; rr < nrows; ++rr) {
          T a0[N];
#pragma unroll
          for (int j = 0; j < N; ++j) a0[j] = sj[(rr * N + j) * kTile];
          T r0;
          tpp_row_residual<T, N, kSynth>(a0, sr[rr * kTile], x, alpha, alpha3, false, r0);
          cost = O::fma(r0, r0, cost);
        }
      }
    }
    __syncwarp();
    if (lane == 0 && c + (int)pipe.nstages < nchunks) {
      fence_proxy_async();
      const int nrow0 = (c + (int)pipe.nstages) * R;
      pipe.issue(jt, rt, nrow0, (m - nrow0 < R) ? (m - nrow0) : R, pipe.stage);
    }
    pipe.advance();
  }
}

// dynamic tile scheduler: one atomic per warp per tile, broadcast by shuffle
__device__ __forceinline__ int64_t tpp_next_tile(unsigned long long *counter, int lane) {
  unsigned long long t = 0;
  if (lane == 0) t = atomicAdd(counter, 1ull);
  return (int64_t)__shfl_sync(0xffffffffu, t, 0);
}

__device__ __forceinline__ unsigned char *tpp_warp_smem(unsigned char *smem, uint32_t warp_smem) {
  return smem + (size_t)(threadIdx.x / 32) * warp_smem;
}

// ---------------------------------------------------------------------------------------------
// tob200_lm_run_*: the whole LM loop per problem, polynomial family evaluated on the fly
// ---------------------------------------------------------------------------------------------
template <typename T>
struct TppRunParams {
  TppData<T> d;
  DevOptions<T> opt;
  T alpha, alpha3;
  T *x;                    // [B][n] in/out
  tob200_result *results;  // [B]
  double *final_hessian;   // [B][n][n] or nullptr: Output::final_hessian (optimizer.h:313-316, lm.h:157-171)
};

template <typename T, int N>
__global__ void __launch_bounds__(kTppThreads, tpp_min_blocks<T, N>())
    tpp_lm_run_kernel(const __grid_constant__ TppRunParams<T> p) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NT = tri_count(N);
  const int lane = threadIdx.x & 31;
  unsigned char *ws = tpp_warp_smem(smem, p.d.warp_smem);
  TppPipe<T, N> pipe;
  pipe.init(ws, p.d.rows, p.d.stages, lane);
  SmemHG<T, N> hg{reinterpret_cast<T *>(ws + kTppBarBytes + (size_t)pipe.nstages * pipe.stage_bytes), lane};

  const bool is_lm = p.opt.solver_type == 0;

  for (int64_t tile = tpp_next_tile(p.d.tile_counter, lane); tile < p.d.ntiles;
       tile = tpp_next_tile(p.d.tile_counter, lane)) {
    const int64_t pidx = tile * kTile + lane;
    const bool valid = pidx < p.d.B;
    LmState<T, N> s;
    s.reset(p.opt);
    if (valid) {
#pragma unroll
      for (int j = 0; j < N; ++j) s.x[j] = p.x[pidx * N + j];
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) s.x[j] = (T)0;
      s.flags |= kFlagDone;
    }
    while (__any_sync(0xffffffffu, !s.done())) {
      const bool active = !s.done();
      const bool do_rebuild = !is_lm || s.rebuild();  // GN's Build always re-accumulates (gn.h:118-131)
      T hu[NT], g[N], cost;
      tpp_pass<T, N, true>(pipe, p.d, tile, lane, active, do_rebuild, s.x, p.alpha, p.alpha3, hu, g, cost);
      if (active) lm_after_pass<T, N>(s, p.opt, do_rebuild, hu, g, cost, p.d.m, hg);
    }
    if (valid) {
#pragma unroll
      for (int j = 0; j < N; ++j) p.x[pidx * N + j] = s.x[j];
      lm_write_result(s, &p.results[pidx]);
      if (p.final_hessian) {
        // SolverLM::Hessian() (lm.h:157-171): the persistent damped H_ with its diagonal divided by
        // 1 + prev_lambda_ (in Scalar), widened to double (optimizer.h:315)
        const bool undamp = is_lm && s.prev_lambda > (T)0;
        const T sc = Ops<T>::add((T)1, s.prev_lambda);
        double *o = p.final_hessian + (size_t)pidx * N * N;
#pragma unroll
        for (int j = 0; j < N; ++j) {
#pragma unroll
          for (int k = j; k < N; ++k) {
            T v = hg.ld_h(tri_index(N, j, k));
            if (j == k && undamp) v = Ops<T>::div(v, sc);
            o[j * N + k] = (double)v;
            o[k * N + j] = (double)v;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// tob200_build_solve_*: one Build + Solve from materialised J, r
// ---------------------------------------------------------------------------------------------
template <typename T>
struct TppBuildSolveParams {
  TppData<T> d;
  const T *lambda;  // [B] or nullptr
  T *dx;            // [B][n]
  double *cost;     // [B]
  T *H_out;         // [B][n][n] or nullptr
  T *g_out;         // [B][n] or nullptr
  int32_t *status;  // [B]
};

template <typename T, int N>
__global__ void __launch_bounds__(kTppThreads, tpp_min_blocks<T, N>())
    tpp_build_solve_kernel(const __grid_constant__ TppBuildSolveParams<T> p) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NT = tri_count(N);
  using L = LdltReg<T, N>;
  const int lane = threadIdx.x & 31;
  unsigned char *ws = tpp_warp_smem(smem, p.d.warp_smem);
  TppPipe<T, N> pipe;
  pipe.init(ws, p.d.rows, p.d.stages, lane);


  for (int64_t tile = tpp_next_tile(p.d.tile_counter, lane); tile < p.d.ntiles;
       tile = tpp_next_tile(p.d.tile_counter, lane)) {
    const int64_t pidx = tile * kTile + lane;
    const bool valid = pidx < p.d.B;
    T hu[NT], g[N], cost, x[N];
#pragma unroll
    for (int j = 0; j < N; ++j) x[j] = (T)0;
    tpp_pass<T, N, false>(pipe, p.d, tile, lane, valid, true, x, (T)0, (T)0, hu, g, cost);
    if (valid) {
      const T lam = p.lambda ? p.lambda[pidx] : (T)0;
      if (lam > (T)0) {  // solvers/lm.h:108-117
        const double sc = 1.0 + (double)lam;
#pragma unroll
        for (int j = 0; j < N; ++j) hu[tri_index(N, j, j)] = (T)((double)hu[tri_index(N, j, j)] * sc);
      }
      p.cost[pidx] = (double)cost;
      if (p.g_out) {
#pragma unroll
        for (int j = 0; j < N; ++j) p.g_out[pidx * N + j] = g[j];
      }
      if (p.H_out) {
        T *Ho = p.H_out + pidx * N * N;
#pragma unroll
        for (int j = 0; j < N; ++j) {
#pragma unroll
          for (int k = j; k < N; ++k) {
            Ho[j * N + k] = hu[tri_index(N, j, k)];
            Ho[k * N + j] = hu[tri_index(N, j, k)];
          }
        }
      }
      int tr[N];
      T dx[N];
      const bool ok = L::factor(hu, tr);  // math.h:232-240
      if (ok) {
#pragma unroll
        for (int j = 0; j < N; ++j) dx[j] = -g[j];  // solvers/gn.h:155
        L::solve(hu, tr, dx);
#pragma unroll
        for (int j = 0; j < N; ++j) p.dx[pidx * N + j] = dx[j];
      }
      p.status[pidx] = ok ? 0 : 1;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// tob200_solver_step_*: one Optimizer_::Step + OptimizeAcc update, state in global memory
// ---------------------------------------------------------------------------------------------
template <typename T>
struct StateRec {  // per-problem scalars of LmState (x, last_dx, H_, grad_ live in their own arrays)
  double final_cost, final_rerr_dec;
  T lambda, prev_lambda, bad_factor;
  int final_nres, stop_reason;
  uint32_t flags;
  int num_builds, iter;
  uint16_t num_iters;
  uint8_t num_failures, num_consec_failures;
};

template <typename T>
struct TppStepParams {
  TppData<T> d;
  DevOptions<T> opt;
  StateRec<T> *rec;  // [B]
  T *x;              // [B][n]
  T *last_dx;        // [B][n]
  T *H;              // [ntiles][NT][32]
  T *g;              // [ntiles][n][32]
  int32_t *needs;    // [B]
  unsigned long long *n_active;  // device counter, zeroed by the host before the launch
  int reset;         // 1: initialise the state instead of stepping (x already holds x0)
  // tob200_solver_step_hg_*: the accumulators arrive filled by the caller (docs/API.md:37-57,137-170) instead of
  // being formed from J, r: grad [B][n], H [B][n][n] (upper triangle read), cost [B], num_residuals [B]
  const T *hg_grad, *hg_H;
  const double *hg_cost;
  const int32_t *hg_nres;
};

template <typename T, int N>
__device__ __forceinline__ void state_load(LmState<T, N> &s, const StateRec<T> &r) {
  s.final_cost = r.final_cost; s.final_rerr_dec = r.final_rerr_dec;
  s.lambda = r.lambda; s.prev_lambda = r.prev_lambda; s.bad_factor = r.bad_factor;
  s.final_nres = r.final_nres; s.stop_reason = r.stop_reason; s.flags = r.flags;
  s.num_builds = r.num_builds; s.iter = r.iter; s.num_iters = r.num_iters;
  s.num_failures = r.num_failures; s.num_consec_failures = r.num_consec_failures;
}
template <typename T, int N>
__device__ __forceinline__ void state_store(const LmState<T, N> &s, StateRec<T> &r) {
  r.final_cost = s.final_cost; r.final_rerr_dec = s.final_rerr_dec;
  r.lambda = s.lambda; r.prev_lambda = s.prev_lambda; r.bad_factor = s.bad_factor;
  r.final_nres = s.final_nres; r.stop_reason = s.stop_reason; r.flags = s.flags;
  r.num_builds = s.num_builds; r.iter = s.iter; r.num_iters = s.num_iters;
  r.num_failures = s.num_failures; r.num_consec_failures = s.num_consec_failures;
}

template <typename T, int N>
__global__ void __launch_bounds__(kTppThreads, tpp_min_blocks<T, N>())
    tpp_step_kernel(const __grid_constant__ TppStepParams<T> p) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NT = tri_count(N);
  const int lane = threadIdx.x & 31;
  unsigned char *ws = tpp_warp_smem(smem, p.d.warp_smem);
  TppPipe<T, N> pipe;
  pipe.init(ws, p.d.rows, p.d.stages, lane);

  const bool is_lm = p.opt.solver_type == 0;
  unsigned long long local_active = 0;

  for (int64_t tile = tpp_next_tile(p.d.tile_counter, lane); tile < p.d.ntiles;
       tile = tpp_next_tile(p.d.tile_counter, lane)) {
    const int64_t pidx = tile * kTile + lane;
    const bool valid = pidx < p.d.B;
    LmState<T, N> s;
    if (p.reset) {
      if (valid) {
        s.reset(p.opt);
        state_store(s, p.rec[pidx]);
#pragma unroll
        for (int j = 0; j < N; ++j) p.last_dx[pidx * N + j] = (T)0;
        p.needs[pidx] = 1;
        local_active++;
      }
      continue;
    }
    if (valid) {
      state_load(s, p.rec[pidx]);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        s.x[j] = p.x[pidx * N + j];
        s.last_dx[j] = p.last_dx[pidx * N + j];
      }
    } else {
      s.reset(p.opt);
      s.flags |= kFlagDone;
#pragma unroll
      for (int j = 0; j < N; ++j) s.x[j] = (T)0;
    }
    if (!__any_sync(0xffffffffu, !s.done())) continue;  // whole tile finished: no traffic at all
    const bool active = !s.done();
    const bool do_rebuild = !is_lm || s.rebuild();
    T hu[NT], g[N], cost;
    double cost_d = 0.0;
    int nres = p.d.m;
    if (p.hg_cost) {  // user-filled accumulators: what `acc(x, grad, H)` left in grad_, H_ and returned
      cost = (T)0;
      if (active) {
        cost_d = p.hg_cost[pidx];
        nres = p.hg_nres[pidx];
        if (do_rebuild) {
          const T *Hu = p.hg_H + (size_t)pidx * N * N;
#pragma unroll
          for (int j = 0; j < N; ++j) {
            g[j] = p.hg_grad[pidx * N + j];
#pragma unroll
            for (int k = j; k < N; ++k) hu[tri_index(N, j, k)] = Hu[j * N + k];
          }
        }
      }
    } else {
      tpp_pass<T, N, false>(pipe, p.d, tile, lane, active, do_rebuild, s.x, (T)0, (T)0, hu, g, cost);
    }
    if (active) {
      GlobalHG<T, N> hg{p.H + (size_t)tile * NT * kTile + lane, p.g + (size_t)tile * N * kTile + lane};
      lm_after_pass<T, N>(s, p.opt, do_rebuild, hu, g, cost, nres, hg, p.hg_cost ? &cost_d : nullptr);
      state_store(s, p.rec[pidx]);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        p.x[pidx * N + j] = s.x[j];
        p.last_dx[pidx * N + j] = s.last_dx[j];
      }
      p.needs[pidx] = s.done() ? -1 : (s.rebuild() || !is_lm ? 1 : 0);
      if (!s.done()) local_active++;
    }
  }
  // one atomic per warp
  for (int off = 16; off > 0; off >>= 1) local_active += __shfl_xor_sync(0xffffffffu, local_active, off);
  if (lane == 0 && local_active) atomicAdd(p.n_active, local_active);
}

// un-damped final Hessian (solvers/lm.h:157-171) from the step solver's persistent H_
template <typename T>
__global__ void final_hessian_kernel(const T *H, const StateRec<T> *rec, int solver_type, int64_t B, int n,
                                     double *out) {
  const int64_t pidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pidx >= B) return;
  const int64_t tile = pidx / kTile;
  const int lane = (int)(pidx % kTile);
  const int nt = tri_count(n);
  const T *h = H + (size_t)tile * nt * kTile + lane;
  const T pl = rec[pidx].prev_lambda;
  const bool undamp = solver_type == 0 && pl > (T)0;
  const T sc = Ops<T>::add((T)1, pl);
  double *o = out + (size_t)pidx * n * n;
  for (int j = 0; j < n; ++j)
    for (int k = j; k < n; ++k) {
      T v = h[tri_index(n, j, k) * kTile];
      if (j == k && undamp) v = Ops<T>::div(v, sc);
      o[j * n + k] = (double)v;
      o[k * n + j] = (double)v;
    }
}

}  // namespace tob200
