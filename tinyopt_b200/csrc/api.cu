// api.cu — the C-ABI of include/tinyopt_b200.h: context, argument checking, kernel selection and
// launch.  No CPU fallback anywhere: every compute entry point either launches a CUDA kernel or
// returns an error.
#include "../../include/tinyopt_b200.h"

#include <cuda_runtime.h>

#include <algorithm>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <tuple>
#include <vector>

#include "internal.h"
#include "gn.cuh"
#include "lg_params.h"
#include "tpp_inst.cuh"
#include "wpp_inst.cuh"
#include "wtc_params.h"

using namespace tob200;

static thread_local std::string g_create_error;

struct tob200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 0;
  std::string err;
  int64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // scratch for PROBLEM_MAJOR -> TILE32 conversion and for the *_host entry points
  // (slots 0-1 layout conversion, 2-5 host entry points, 7 warp-per-problem H_, 8-19 large-n family)
  static constexpr int kScratchSlots = 28;
  void *scratch[kScratchSlots] = {};
  size_t scratch_bytes[kScratchSlots] = {};
  std::map<std::tuple<int, int, int, int, size_t>, int> occupancy;  // (dtype, n, kind, block, smem) -> CTAs/SM
  // tuning knobs (env: TOB200_TPP_STAGE_BYTES, TOB200_TPP_STAGES, TOB200_TPP_CTAS_PER_SM)
  int tpp_stage_bytes = 16384;  // upper bound of one pipeline stage
  int tpp_stages = 2;  // measured best on B200 (tools/tune_tpp.py): few, large stages
  int tpp_ctas_per_sm = 0;  // 0: what the kernel was compiled for (__launch_bounds__)
  // work-queue counters: a zeroed ring, one fresh counter per launch (re-zeroed when exhausted), so
  // that a launch costs no extra memset
  static constexpr int kNumCounters = 4096;
  unsigned long long *counters = nullptr;
  int next_counter = 0;
  unsigned long long *tile_counter = nullptr;  // the counter of the launch being configured
  int wpp_stages = 1;  // env TOB200_WPP_STAGES (1: three CTAs per SM fit, measured best)
  int wpp_tc = 1;      // env TOB200_WPP_TC / tob200_set_exact: 1 = mid-n float lm_run on the tensor-core kernel (wtc.cuh)
  int wtc_prefetch = 6;  // env TOB200_WTC_PREFETCH: L2 prefetch distance of its loader, in 32-row stages
  int wtc_debug = 0;     // env TOB200_WTC_DEBUG (timing experiments)
  int wtc_raw_stages = 0, wtc_op_stages = 0;  // env TOB200_WTC_RAW / TOB200_WTC_OPS: ring depths (0: the deepest that fit)
  // *_host entry points: upload / solve / download pipeline over chunks of the batch
  static constexpr int kMaxChunks = 16;
  int host_chunks = 4;  // env TOB200_HOST_CHUNKS
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t ev_h2d[kMaxChunks] = {}, ev_run[kMaxChunks] = {}, ev_side = nullptr;
  // large-n family: 3 = 3xTF32 (hi*hi + hi*lo + lo*hi, FP32-level accuracy), 1 = plain TF32
  int lg_tf32_terms = 3;  // env TOB200_LG_TF32_TERMS
  int lg_raw_stages = kLgRawStages;  // env TOB200_LG_RAW_STAGES (2..6)
  int lg_exact = 0;       // tob200_set_exact(ctx, 2) / env TOB200_LG_EXACT: float 56 <= n <= 512 on the general (bit-exact) family
  int lg_mc = 0;          // env TOB200_LG_MC: 1 = JtJ kernel as clusters of two CTAs with TMA multicast of the shared raw stages
  int lg_fp16 = 1;        // env TOB200_LG_FP16: 1 = FP16 hi / lo split (kind::f16), 0 = TF32 split (kind::tf32)
  // device time of the last large-n call by phase (0 eval, 1 syrk, 2 solve): CUDA event pairs
  static constexpr int kMaxPhaseEvents = 3 * 80;
  std::vector<cudaEvent_t> phase_ev;  // 2 events per (iteration, phase)
  int phase_used = 0;
  int phase_kind[kMaxPhaseEvents] = {};
  // host-orchestrated LM loops (large-n and general families): the number of still-running problems of pass k is
  // read back asynchronously into pinned memory and only looked at after pass k + 1 has been queued, so the device
  // never idles on a host round trip between passes (a surplus pass finds every problem done and exits at once)
  unsigned long long *active_host = nullptr;  // [2] pinned
  cudaEvent_t ev_active[2] = {};
};

namespace {

int fail(tob200_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->err = msg;
  else g_create_error = msg;
  return code;
}
int fail_cuda(tob200_ctx *ctx, cudaError_t e, const char *what) {
  cudaGetLastError();  // a failed launch leaves its (non-sticky) error behind: do not let the next call trip on it
  return fail(ctx, e == cudaErrorMemoryAllocation ? TOB200_ERR_NOMEM : TOB200_ERR_CUDA,
              std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(expr)                                                 \
  do {                                                           \
    cudaError_t e__ = (expr);                                    \
    if (e__ != cudaSuccess) return fail_cuda(ctx, e__, #expr);   \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int env_int(const char *name, int dflt) {
  const char *v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

int ensure_scratch(tob200_ctx *ctx, int slot, size_t bytes) {
  if (ctx->scratch_bytes[slot] >= bytes) return TOB200_OK;
  if (ctx->scratch[slot]) {
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(ctx->scratch[slot]));
    ctx->scratch[slot] = nullptr;
    ctx->scratch_bytes[slot] = 0;
  }
  CK(cudaMalloc(&ctx->scratch[slot], bytes));
  ctx->scratch_bytes[slot] = bytes;
  return TOB200_OK;
}

// hands out a zeroed work-queue counter for the next launch
int next_counter(tob200_ctx *ctx) {
  if (ctx->next_counter >= tob200_ctx::kNumCounters) {
    CK(cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long) * tob200_ctx::kNumCounters, ctx->stream));
    ctx->next_counter = 0;
  }
  ctx->tile_counter = ctx->counters + ctx->next_counter++;
  return TOB200_OK;
}

template <typename T> constexpr int dtype_of();
template <> constexpr int dtype_of<float>() { return TOB200_F32; }
template <> constexpr int dtype_of<double>() { return TOB200_F64; }

template <typename T>
TppEntry tpp_entry_for(int n) {
  if (sizeof(T) == 4) {
    if (n >= 1 && n <= 4) return tpp_entry_f32_a;
    if (n <= 8) return tpp_entry_f32_b;
    if (n <= 10) return tpp_entry_f32_c;
    if (n <= kTppMaxN_f32) return tpp_entry_f32_d;
  } else {
    if (n >= 1 && n <= 4) return tpp_entry_f64_a;
    if (n <= 6) return tpp_entry_f64_b;
    if (n <= kTppMaxN_f64) return tpp_entry_f64_c;
  }
  return nullptr;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
int tpp_min_blocks_rt(int n) {  // tpp_min_blocks<T, N>() for a run-time n
  if (sizeof(T) == 8) return n <= 2 ? 4 : (n <= 6 ? TOB200_TPP_MINB_F64_MID : 2);
  return n <= 6 ? 4 : (n <= 8 ? 3 : 2);
}

// launch geometry of a thread-per-problem kernel: pipeline shape from the shared-memory budget of
// the occupancy the kernel was compiled for, persistent grid = resident CTAs, dynamic tile queue
template <typename T>
int tpp_configure(tob200_ctx *ctx, int n, int m, int64_t B, int kind, TppData<T> *d, TppLaunch *cfg) {
  TppEntry entry = tpp_entry_for<T>(n);
  if (!entry) return fail(ctx, TOB200_ERR_UNSUPPORTED, "no thread-per-problem kernel for this n");
  const int warps = kTppThreads / 32;
  int minb = tpp_min_blocks_rt<T>(n);
  if (ctx->tpp_ctas_per_sm > 0 && ctx->tpp_ctas_per_sm < minb) minb = ctx->tpp_ctas_per_sm;
  const size_t row_bytes = (size_t)(n + 1) * kTile * sizeof(T);
  const size_t fixed = kTppBarBytes + (size_t)(tri_count(n) + n) * kTile * sizeof(T) + 128;
  const size_t budget = (232448 - (size_t)minb * 1024) / ((size_t)minb * warps);
  int stages = ctx->tpp_stages;
  if (stages < 2) stages = 2;
  if (stages > kTppMaxStages) stages = kTppMaxStages;
  size_t avail = budget > fixed ? budget - fixed : 0;
  int rows = (int)(avail / ((size_t)stages * row_bytes));
  while (rows < 1 && stages > 2) rows = (int)(avail / ((size_t)(--stages) * row_bytes));
  if (rows < 1) rows = 1;  // occupancy will drop below minb; the query below reports what fits
  if ((size_t)rows * row_bytes > (size_t)ctx->tpp_stage_bytes) rows = (int)(ctx->tpp_stage_bytes / row_bytes);
  if (rows < 1) rows = 1;
  if (rows > m) rows = m > 0 ? m : 1;
  d->B = B;
  d->ntiles = (B + kTile - 1) / kTile;
  d->m = m;
  d->rows = rows;
  d->stages = stages;
  d->warp_smem = (uint32_t)tpp_warp_smem_bytes(n, rows, stages, sizeof(T));
  d->tile_counter = ctx->tile_counter;
  cfg->block = kTppThreads;
  cfg->smem = (size_t)d->warp_smem * warps;
  cfg->stream = ctx->stream;
  const auto key = std::make_tuple(dtype_of<T>(), n, kind, cfg->block, cfg->smem);
  auto it = ctx->occupancy.find(key);
  int per_sm = 0;
  if (it == ctx->occupancy.end()) {
    cudaError_t e = entry(kTppQuery, n, kind, nullptr, *cfg, &per_sm);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "occupancy query");
    if (per_sm < 1) return fail(ctx, TOB200_ERR_CUDA, "kernel does not fit on an SM (shared memory / registers)");
    ctx->occupancy[key] = per_sm;
  } else {
    per_sm = it->second;
  }
  if (ctx->tpp_ctas_per_sm > 0 && ctx->tpp_ctas_per_sm < per_sm) per_sm = ctx->tpp_ctas_per_sm;
  int64_t grid = (int64_t)per_sm * ctx->num_sms;
  const int64_t need = (d->ntiles + warps - 1) / warps;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  cfg->grid = (int)grid;
  // the tile queue of this launch: a fresh zeroed counter
  int rcq = next_counter(ctx);
  if (rcq != TOB200_OK) return rcq;
  d->tile_counter = ctx->tile_counter;
  return TOB200_OK;
}

// Each kernel family has a native layout (1: TILE32, 2: PROBLEM_MAJOR); inputs in the other layout
// are converted on the device into context scratch (slots 0: J/A, 1: r/y)
template <typename T>
int to_native_layout(tob200_ctx *ctx, int family, int layout, int64_t B, int m, int n, const T **J, const T **r) {
  if (layout != TOB200_LAYOUT_TILE32 && layout != TOB200_LAYOUT_PROBLEM_MAJOR)
    return fail(ctx, TOB200_ERR_INVALID, "unknown layout");
  const int native = family == 1 ? TOB200_LAYOUT_TILE32 : TOB200_LAYOUT_PROBLEM_MAJOR;
  if (layout == native || B == 0 || m == 0) return TOB200_OK;
  const size_t ej = (size_t)tob200_tiled_elems(B, m, n), er = (size_t)tob200_tiled_elems(B, m, 1);
  int rc;
  if ((rc = ensure_scratch(ctx, 0, ej * sizeof(T))) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, 1, er * sizeof(T))) != TOB200_OK) return rc;
  if (native == TOB200_LAYOUT_TILE32) {
    CK(launch_retile<T>(*J, B, m, n, (T *)ctx->scratch[0], ctx->stream));
    CK(launch_retile<T>(*r, B, m, 1, (T *)ctx->scratch[1], ctx->stream));
  } else {
    CK(launch_untile<T>(*J, B, m, n, (T *)ctx->scratch[0], ctx->stream));
    CK(launch_untile<T>(*r, B, m, 1, (T *)ctx->scratch[1], ctx->stream));
  }
  ctx->launches += 2;
  *J = (const T *)ctx->scratch[0];
  *r = (const T *)ctx->scratch[1];
  return TOB200_OK;
}

// launch geometry of a warp-per-problem kernel (family 2)
typedef cudaError_t (*WppEntry)(int, int, int, const void *, const TppLaunch &, int *);
template <typename T>
inline WppEntry wpp_entry_for(int n, int kind) {
  if (kind == kWppRunInv || kind == kWppStepInv) {
    if (sizeof(T) == 8) return wpp_blk_for(n) == 4 ? wpp_entry_f64_blk4_inv : wpp_entry_f64_blk8_inv;
    return wpp_blk_for(n) == 4 ? wpp_entry_f32_blk4_inv : wpp_entry_f32_blk8_inv;
  }
  if (sizeof(T) == 8) return wpp_blk_for(n) == 4 ? wpp_entry_f64_blk4 : wpp_entry_f64_blk8;
  return wpp_blk_for(n) == 4 ? wpp_entry_f32_blk4 : wpp_entry_f32_blk8;
}

template <typename T>
int wpp_configure(tob200_ctx *ctx, int n, int m, int64_t B, int kind, const T *A, const T *y, WppData<T> *d,
                  TppLaunch *cfg) {
  const int blk = wpp_blk_for(n), nb = wpp_nb_for(n), np = nb * blk;
  const int warps = kWppThreads / 32;
  int stages = ctx->wpp_stages;
  if (stages < 1) stages = 1;
  if (stages > kWppMaxStages) stages = kWppMaxStages;
  d->A = A;
  d->y = y;
  d->B = B;
  d->m = m;
  d->n = n;
  d->stages = stages;
  d->L = wpp_smem_layout(n, np, stages, (uint32_t)sizeof(T));
  // TMA bulk copies need 16-byte aligned sources and sizes for every chunk of every problem
  const int64_t per16 = 16 / (int64_t)sizeof(T);
  d->use_tma = (aligned16(A) && ((int64_t)m * n) % per16 == 0) ? 1 : 0;  // y is read straight from global
  d->counter = ctx->tile_counter;
  cfg->block = kWppThreads;
  cfg->smem = (size_t)d->L.total * warps;
  cfg->stream = ctx->stream;
  const auto key = std::make_tuple(100 * (int)sizeof(T) / 4 + blk, nb, kind, cfg->block, cfg->smem);
  auto it = ctx->occupancy.find(key);
  int per_sm = 0;
  if (it == ctx->occupancy.end()) {
    cudaError_t e = wpp_entry_for<T>(n, kind)(kTppQuery, nb, kind, nullptr, *cfg, &per_sm);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "occupancy query");
    if (per_sm < 1) return fail(ctx, TOB200_ERR_CUDA, "kernel does not fit on an SM (shared memory / registers)");
    ctx->occupancy[key] = per_sm;
  } else {
    per_sm = it->second;
  }
  int64_t grid = (int64_t)per_sm * ctx->num_sms;
  const int64_t need = (B + warps - 1) / warps;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  cfg->grid = (int)grid;
  int rc = ensure_scratch(ctx, 7, (size_t)grid * warps * np * wpp_ldw(np) * sizeof(T));
  if (rc != TOB200_OK) return rc;
  d->hpersist = (T *)ctx->scratch[7];
  if ((rc = next_counter(ctx)) != TOB200_OK) return rc;
  d->counter = ctx->tile_counter;
  return TOB200_OK;
}

template <typename T>
int wpp_launch(tob200_ctx *ctx, int n, int kind, const void *params, const TppLaunch &cfg) {
  CK(wpp_entry_for<T>(n, kind)(kTppLaunch, wpp_nb_for(n), kind, params, cfg, nullptr));
  ctx->launches++;
  return TOB200_OK;
}


// launch geometry / scratch of the mid-n tensor-core kernel (wtc.cuh)
inline bool wtc_takes(const tob200_ctx *ctx, int elt, int64_t m, int n, const void *A) {
  return elt == 4 && ctx->wpp_tc && n >= kWtcMinN && n <= kWtcMaxN && m >= kWtcMinM && (m * n) % 4 == 0 && aligned16(A);
}
int wtc_configure(tob200_ctx *ctx, const float *A, const float *y, int64_t B, int m, int n, WtcParams *p, int64_t *grid_out) {
  p->A = A;
  p->y = y;
  p->B = B;
  p->m = m;
  p->n = n;
  p->L = wtc_smem_best(n);
  if (ctx->wtc_raw_stages > 0 && ctx->wtc_op_stages > 0) {  // tuning override (env TOB200_WTC_RAW / TOB200_WTC_OPS)
    const WtcSmem L2 = wtc_smem_plan(n, ctx->wtc_raw_stages, ctx->wtc_op_stages);
    if (ctx->wtc_raw_stages <= kWtcMaxRawStages && ctx->wtc_op_stages <= kWtcMaxOpStages && L2.total <= 232448u) p->L = L2;
  }
  p->prefetch = ctx->wtc_prefetch;
  p->debug = ctx->wtc_debug;
  int64_t grid = ctx->num_sms;
  const int64_t need = (B + kWtcSlots - 1) / kWtcSlots;
  if (grid > need) grid = need;
  int rc = ensure_scratch(ctx, 7, (size_t)grid * kWtcSlots * wtc_hp_floats(n) * sizeof(float));
  if (rc != TOB200_OK) return rc;
  p->hpersist = (float *)ctx->scratch[7];
  if ((rc = next_counter(ctx)) != TOB200_OK) return rc;
  p->counter = ctx->tile_counter;
  *grid_out = grid;
  return TOB200_OK;
}

// ---- large-n family (lg.cuh): host-orchestrated eval -> syrk -> solve per LM iteration ------------
enum LgSlot { kLgH = 8, kLgHd, kLgG, kLgCost, kLgScale, kLgRec, kLgLastDx, kLgW, kLgActive, kLgDg };
constexpr int kLgAmaxSlot = 6;  // [B] max |J_ij| per problem for the FP16-split J^T J (a free scratch slot)

int lg_phase_begin(tob200_ctx *ctx, int kind) {
  if (ctx->phase_used + 2 > tob200_ctx::kMaxPhaseEvents * 2) return -1;
  while ((int)ctx->phase_ev.size() < ctx->phase_used + 2) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return -1;
    ctx->phase_ev.push_back(e);
  }
  ctx->phase_kind[ctx->phase_used / 2] = kind;
  cudaEventRecord(ctx->phase_ev[ctx->phase_used], ctx->stream);
  return ctx->phase_used;
}
void lg_phase_end(tob200_ctx *ctx, int slot) {
  if (slot < 0) return;
  cudaEventRecord(ctx->phase_ev[slot + 1], ctx->stream);
  ctx->phase_used = slot + 2;
}

// pass-loop helper: queue the read-back of the active counter of pass `pass`; return the count of pass `pass - 1`
// (already complete or nearly so: it was queued one whole pass earlier), or -1 for the first pass
int lm_loop_poll(tob200_ctx *ctx, const unsigned long long *n_active_dev, int pass, long long *prev_active) {
  if (!ctx->active_host) {
    CK(cudaMallocHost((void **)&ctx->active_host, 2 * sizeof(unsigned long long)));
    for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&ctx->ev_active[i], cudaEventDisableTiming));
  }
  const int slot = pass & 1;
  CK(cudaMemcpyAsync(&ctx->active_host[slot], n_active_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaEventRecord(ctx->ev_active[slot], ctx->stream));
  *prev_active = -1;
  if (pass > 0) {
    CK(cudaEventSynchronize(ctx->ev_active[slot ^ 1]));
    *prev_active = (long long)ctx->active_host[slot ^ 1];
  }
  return TOB200_OK;
}

struct LgBuffers {
  float *H, *hd, *g, *dg, *cost, *scale, *W, *last_dx, *amax;
  LmScalars<float> *rec;
  unsigned long long *n_active;
  int np, solve_grid;
};

int lg_prepare(tob200_ctx *ctx, int64_t B, int m, int n, bool need_state, bool need_scale, LgBuffers *b) {
  const int np = lg_np(n);
  b->np = np;
  b->solve_grid = (int)(B < ctx->num_sms ? B : ctx->num_sms);  // lg_solve_kernel: 1 CTA per SM (128 registers)
  int rc;
  if ((rc = ensure_scratch(ctx, kLgH, (size_t)B * np * np * 4)) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kLgHd, (size_t)B * np * 4)) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kLgG, (size_t)B * n * 4)) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kLgDg, (size_t)B * n * 4)) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kLgCost, (size_t)B * 4)) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kLgW, (size_t)b->solve_grid * np * np * 4)) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kLgActive, 64)) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kLgAmaxSlot, (size_t)B * 4)) != TOB200_OK) return rc;
  if (need_scale && (rc = ensure_scratch(ctx, kLgScale, (size_t)B * (m > 0 ? m : 1) * 4)) != TOB200_OK) return rc;
  if (need_state) {
    if ((rc = ensure_scratch(ctx, kLgRec, (size_t)B * sizeof(LmScalars<float>))) != TOB200_OK) return rc;
    if ((rc = ensure_scratch(ctx, kLgLastDx, (size_t)B * n * 4)) != TOB200_OK) return rc;
  }
  b->H = (float *)ctx->scratch[kLgH];
  b->hd = (float *)ctx->scratch[kLgHd];
  b->g = (float *)ctx->scratch[kLgG];
  b->dg = (float *)ctx->scratch[kLgDg];
  b->cost = (float *)ctx->scratch[kLgCost];
  b->scale = (float *)ctx->scratch[kLgScale];
  b->W = (float *)ctx->scratch[kLgW];
  b->rec = (LmScalars<float> *)ctx->scratch[kLgRec];
  b->last_dx = (float *)ctx->scratch[kLgLastDx];
  b->n_active = (unsigned long long *)ctx->scratch[kLgActive];
  b->amax = (float *)ctx->scratch[kLgAmaxSlot];
  return TOB200_OK;
}

// A [B][m][n] as the 2-D tensor {n columns, B * m rows} with boxes of {128 columns, kLgStageK rows} for the
// TMA loader of lg_syrk_kernel.  cuTensorMapEncodeTiled is a driver entry point: fetched through the
// runtime so that the library does not link libcuda directly.
bool lg_encode_tmap(CUtensorMap *tm, const float *A, int64_t B, int m, int n) {
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                               const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  static std::once_flag looked_up;
  std::call_once(looked_up, [] {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      encode = (EncodeFn)fn;
  });
  const int64_t rows = B * (int64_t)m;
  if (!encode || m <= 0 || rows >= (int64_t)1 << 31 || (n % 4) != 0 || ((uintptr_t)A & 15) != 0) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)n * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kLgBoxCols, (cuuint32_t)kLgStageK};
  const cuuint32_t estride[2] = {1, 1};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(A), gdim, gstride, box, estride,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

LgSyrkParams lg_syrk_params(tob200_ctx *ctx, const LgBuffers &b, const float *A, const float *scale,
                            const LmScalars<float> *rec, int64_t B, int m, int n, int is_lm) {
  LgSyrkParams sp;
  std::memset(&sp.tmap, 0, sizeof(sp.tmap));
  sp.use_tmap = (env_int("TOB200_LG_NO_TMAP", 0) == 0 && lg_encode_tmap(&sp.tmap, A, B, m, n)) ? 1 : 0;
  sp.A = A;
  sp.scale = scale;
  sp.rec = rec;
  sp.H = b.H;
  sp.B = B;
  sp.m = m;
  sp.n = n;
  sp.np = b.np;
  sp.nstrips = (b.np + 127) / 128;
  sp.fp16 = ctx->lg_fp16;
  sp.raw_stages = ctx->lg_raw_stages;
  // the raw ring may not crowd out the operand ring: at least two operand stages of the widest strip must remain
  while (sp.raw_stages > 2 && (size_t)sp.raw_stages * lg_syrk_raw_bytes(b.np) + 4 * (size_t)lg_syrk_half_bytes(b.np, sp.fp16) > 200 * 1024)
    --sp.raw_stages;
  sp.stages = lg_syrk_stages(b.np, sp.raw_stages, sp.fp16);
  sp.terms = ctx->lg_tf32_terms;
  sp.amax = b.amax;
  sp.is_lm = is_lm;
  sp.debug = env_int("TOB200_LG_DEBUG", 0);
  sp.half_bytes = lg_syrk_half_bytes(b.np, sp.fp16);
  // cluster multicast of the raw stages both units of a problem read (lg.cuh): four strips, FP16 split, tensor map, m rows
  // in whole 16-row stages for both wide strips (they are: same RS), an even grid
  sp.mc = (ctx->lg_mc && sp.nstrips == 4 && sp.fp16 && sp.use_tmap && !(sp.debug & 4) && B >= 1 && (ctx->num_sms % 2) == 0) ? 1 : 0;
  return sp;
}

// the whole LM loop for 56 <= n <= 512 (float): three kernels per iteration over the active problems
int lg_lm_run(tob200_ctx *ctx, const tob200_options *opt, const float *A, const float *y, float alpha, int64_t B, int m,
              int n, float *x, tob200_result *results, double *final_hessian = nullptr, int n_out = 0) {
  LgBuffers b;
  int rc = lg_prepare(ctx, B, m, n, true, true, &b);
  if (rc != TOB200_OK) return rc;
  const DevOptions<float> dopt = make_dev_options<float>(*opt);
  const int is_lm = opt->solver_type == 0;
  ctx->phase_used = 0;
  CK(launch_lg_init(b.rec, dopt, b.last_dx, B, n, ctx->stream));
  ctx->launches++;
  LgEvalParams ep;
  ep.A = A; ep.y = y; ep.x = x; ep.rec = b.rec; ep.scale_in = nullptr; ep.scale = b.scale; ep.g = b.g; ep.dg = b.dg; ep.cost = b.cost; ep.amax = b.amax;
  ep.B = B; ep.m = m; ep.n = n; ep.synth = 1; ep.is_lm = is_lm; ep.alpha = alpha; ep.alpha3 = 3.f * alpha;
  const LgSyrkParams sp = lg_syrk_params(ctx, b, A, b.scale, b.rec, B, m, n, is_lm);
  LgSolveParams vp;
  vp.H = b.H; vp.dg = b.dg; vp.hd = b.hd; vp.g = b.g; vp.cost = b.cost; vp.W = b.W; vp.B = B; vp.n = n; vp.np = b.np; vp.nres = m;
  vp.mode = 0; vp.opt = dopt; vp.rec = b.rec; vp.x = x; vp.last_dx = b.last_dx; vp.results = results;
  vp.n_active = b.n_active; vp.lambda = nullptr; vp.b = nullptr; vp.dx = nullptr; vp.cost_out = nullptr; vp.status = nullptr; vp.max_std = nullptr;
  const int max_passes = opt->max_iters + 1 + (opt->check_final_cost ? 1 : 0);  // optimizer.h:248-250
  for (int pass = 0; pass < max_passes; ++pass) {
    int ev = lg_phase_begin(ctx, 0);
    if (m > 0) {
      CK(launch_lg_eval(ep, ctx->num_sms, ctx->stream));
      ctx->launches++;
    } else {
      CK(cudaMemsetAsync(b.cost, 0, (size_t)B * 4, ctx->stream));
      CK(cudaMemsetAsync(b.g, 0, (size_t)B * n * 4, ctx->stream));
      CK(cudaMemsetAsync(b.dg, 0, (size_t)B * n * 4, ctx->stream));
    }
    lg_phase_end(ctx, ev);
    ev = lg_phase_begin(ctx, 1);
    if (m > 0) {
      CK(launch_lg_syrk(sp, ctx->num_sms, ctx->stream));
      ctx->launches++;
    } else {
      CK(cudaMemsetAsync(b.H, 0, (size_t)B * b.np * b.np * 4, ctx->stream));
    }
    lg_phase_end(ctx, ev);
    ev = lg_phase_begin(ctx, 2);
    CK(cudaMemsetAsync(b.n_active, 0, sizeof(unsigned long long), ctx->stream));
    CK(launch_lg_solve(vp, b.solve_grid, ctx->stream));
    ctx->launches++;
    lg_phase_end(ctx, ev);
    long long prev_active;
    if ((rc = lm_loop_poll(ctx, b.n_active, pass, &prev_active)) != TOB200_OK) return rc;
    if (prev_active == 0) break;  // pass - 1 already finished every problem: this pass was a no-op
  }
  if (final_hessian) {  // Output::final_hessian (optimizer.h:313-316)
    CK(launch_lg_final_hessian(b.H, b.hd, b.rec, opt->solver_type, B, n_out > 0 ? n_out : n, b.np, final_hessian, ctx->stream));
    ctx->launches++;
  }
  return TOB200_OK;
}

// one Build + Solve from materialised J, r for 56 <= n <= 512 (float)
int lg_build_solve(tob200_ctx *ctx, const float *J, const float *r, int64_t B, int m, int n, const float *lambda,
                   float *dx, double *cost, float *H_out, float *g_out, int32_t *status) {
  LgBuffers b;
  int rc = lg_prepare(ctx, B, m, n, false, false, &b);
  if (rc != TOB200_OK) return rc;
  ctx->phase_used = 0;
  LgEvalParams ep;
  ep.A = J; ep.y = r; ep.x = nullptr; ep.rec = nullptr; ep.scale_in = nullptr; ep.scale = nullptr; ep.g = b.g; ep.dg = b.dg; ep.cost = b.cost; ep.amax = b.amax;
  ep.B = B; ep.m = m; ep.n = n; ep.synth = 0; ep.is_lm = 1; ep.alpha = 0.f; ep.alpha3 = 0.f;
  int ev = lg_phase_begin(ctx, 0);
  if (m > 0) {
    CK(launch_lg_eval(ep, ctx->num_sms, ctx->stream));
    ctx->launches++;
  } else {
    CK(cudaMemsetAsync(b.cost, 0, (size_t)B * 4, ctx->stream));
    CK(cudaMemsetAsync(b.g, 0, (size_t)B * n * 4, ctx->stream));
    CK(cudaMemsetAsync(b.dg, 0, (size_t)B * n * 4, ctx->stream));
  }
  lg_phase_end(ctx, ev);
  ev = lg_phase_begin(ctx, 1);
  if (m > 0) {
    const LgSyrkParams sp = lg_syrk_params(ctx, b, J, nullptr, nullptr, B, m, n, 1);
    CK(launch_lg_syrk(sp, ctx->num_sms, ctx->stream));
    ctx->launches++;
  } else {
    CK(cudaMemsetAsync(b.H, 0, (size_t)B * b.np * b.np * 4, ctx->stream));
  }
  lg_phase_end(ctx, ev);
  if (H_out) {
    CK(launch_lg_export_h(b.H, b.dg, lambda, B, n, b.np, H_out, ctx->stream));
    ctx->launches++;
  }
  if (g_out) CK(cudaMemcpyAsync(g_out, b.g, (size_t)B * n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  LgSolveParams vp;
  vp.H = b.H; vp.dg = b.dg; vp.hd = nullptr; vp.g = b.g; vp.cost = b.cost; vp.W = b.W; vp.B = B; vp.n = n; vp.np = b.np; vp.nres = m;
  vp.mode = 1; vp.opt = DevOptions<float>(); vp.rec = nullptr; vp.x = nullptr; vp.last_dx = nullptr; vp.results = nullptr;
  vp.n_active = nullptr; vp.lambda = lambda; vp.b = nullptr; vp.dx = dx; vp.cost_out = cost; vp.status = status; vp.max_std = nullptr;
  ev = lg_phase_begin(ctx, 2);
  CK(launch_lg_solve(vp, b.solve_grid, ctx->stream));
  ctx->launches++;
  lg_phase_end(ctx, ev);
  return TOB200_OK;
}

// ---- general family (gn.cuh): double above n = 55, any precision above n = 512, use_ldlt = false above 55 ----
enum GnSlot { kGnH = 8, kGnHd, kGnG, kGnCost, kGnRs, kGnRec, kGnLastDx, kGnW, kGnActive };  // shares the large-n slots

template <typename T>
struct GnBuffers {
  T *H, *hd, *g, *cost, *rs, *W, *last_dx;
  LmScalars<T> *rec;
  unsigned long long *n_active;
  int solve_grid;
};

template <typename T>
int gn_prepare(tob200_ctx *ctx, int64_t B, int m, int n, GnBuffers<T> *b) {
  b->solve_grid = (int)(B < 2 * ctx->num_sms ? B : 2 * ctx->num_sms);
  int rc;
  if ((rc = ensure_scratch(ctx, kGnH, (size_t)B * n * n * sizeof(T))) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kGnHd, (size_t)B * n * sizeof(T))) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kGnG, (size_t)B * n * sizeof(T))) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kGnCost, (size_t)B * sizeof(T))) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kGnRs, (size_t)B * 2 * (m > 0 ? m : 1) * sizeof(T))) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kGnRec, (size_t)B * sizeof(LmScalars<T>))) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kGnLastDx, (size_t)B * n * sizeof(T))) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kGnW, (size_t)b->solve_grid * n * n * sizeof(T))) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, kGnActive, 64)) != TOB200_OK) return rc;
  b->H = (T *)ctx->scratch[kGnH];
  b->hd = (T *)ctx->scratch[kGnHd];
  b->g = (T *)ctx->scratch[kGnG];
  b->cost = (T *)ctx->scratch[kGnCost];
  b->rs = (T *)ctx->scratch[kGnRs];
  b->rec = (LmScalars<T> *)ctx->scratch[kGnRec];
  b->last_dx = (T *)ctx->scratch[kGnLastDx];
  b->W = (T *)ctx->scratch[kGnW];
  b->n_active = (unsigned long long *)ctx->scratch[kGnActive];
  return TOB200_OK;
}

// the whole LM loop: two kernels per iteration over the still-running problems (host orchestrated, like lg_lm_run)
template <typename T>
int gn_lm_run(tob200_ctx *ctx, const tob200_options *opt, const T *A, const T *y, T alpha, int64_t B, int m, int n, T *x,
              tob200_result *results, double *final_hessian) {
  GnBuffers<T> b;
  int rc = gn_prepare<T>(ctx, B, m, n, &b);
  if (rc != TOB200_OK) return rc;
  const DevOptions<T> dopt = make_dev_options<T>(*opt);
  const int is_lm = opt->solver_type == 0;
  ctx->phase_used = 0;
  CK(launch_gn_init<T>(b.rec, dopt, b.last_dx, B, n, ctx->stream));
  ctx->launches++;
  GnAccumParams<T> ap;
  ap.A = A; ap.y = y; ap.x = x; ap.rec = b.rec; ap.rs = b.rs; ap.g = b.g; ap.H = b.H; ap.cost = b.cost;
  ap.B = B; ap.m = m; ap.n = n; ap.synth = 1; ap.is_lm = is_lm; ap.alpha = alpha; ap.alpha3 = (T)3 * alpha;
  GnSolveParams<T> vp;
  vp.H = b.H; vp.hd = b.hd; vp.g = b.g; vp.cost = b.cost; vp.W = b.W; vp.B = B; vp.n = n; vp.nres = m; vp.mode = 0; vp.opt = dopt;
  vp.rec = b.rec; vp.x = x; vp.last_dx = b.last_dx; vp.results = results; vp.n_active = b.n_active; vp.lambda = nullptr;
  vp.dx = nullptr; vp.cost_out = nullptr; vp.status = nullptr; vp.use_ldlt = opt->use_ldlt;
  const int max_passes = opt->max_iters + 1 + (opt->check_final_cost ? 1 : 0);  // optimizer.h:248-250
  for (int pass = 0; pass < max_passes; ++pass) {
    int ev = lg_phase_begin(ctx, 0);
    CK(launch_gn_accum<T>(ap, ctx->num_sms, ctx->stream));
    ctx->launches++;
    lg_phase_end(ctx, ev);
    ev = lg_phase_begin(ctx, 2);
    CK(cudaMemsetAsync(b.n_active, 0, sizeof(unsigned long long), ctx->stream));
    CK(launch_gn_solve<T>(vp, b.solve_grid, ctx->stream));
    ctx->launches++;
    lg_phase_end(ctx, ev);
    long long prev_active;
    if ((rc = lm_loop_poll(ctx, b.n_active, pass, &prev_active)) != TOB200_OK) return rc;
    if (prev_active == 0) break;  // pass - 1 already finished every problem: this pass was a no-op
  }
  if (final_hessian) {
    CK((launch_gn_export_h<T, double>(b.H, b.hd, b.rec, nullptr, opt->solver_type, B, n, final_hessian, ctx->stream)));
    ctx->launches++;
  }
  return TOB200_OK;
}

// one Build + Solve from materialised J, r
template <typename T>
int gn_build_solve(tob200_ctx *ctx, const T *J, const T *r, int64_t B, int m, int n, const T *lambda, T *dx, double *cost,
                   T *H_out, T *g_out, int32_t *status) {
  GnBuffers<T> b;
  int rc = gn_prepare<T>(ctx, B, m, n, &b);
  if (rc != TOB200_OK) return rc;
  ctx->phase_used = 0;
  GnAccumParams<T> ap;
  ap.A = J; ap.y = r; ap.x = nullptr; ap.rec = nullptr; ap.rs = b.rs; ap.g = b.g; ap.H = b.H; ap.cost = b.cost;
  ap.B = B; ap.m = m; ap.n = n; ap.synth = 0; ap.is_lm = 1; ap.alpha = (T)0; ap.alpha3 = (T)0;
  CK(launch_gn_accum<T>(ap, ctx->num_sms, ctx->stream));
  ctx->launches++;
  if (g_out) CK(cudaMemcpyAsync(g_out, b.g, (size_t)B * n * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
  if (H_out) {
    CK((launch_gn_export_h<T, T>(b.H, nullptr, nullptr, lambda, 0, B, n, H_out, ctx->stream)));
    ctx->launches++;
  }
  GnSolveParams<T> vp;
  vp.H = b.H; vp.hd = nullptr; vp.g = b.g; vp.cost = b.cost; vp.W = b.W; vp.B = B; vp.n = n; vp.nres = m; vp.mode = 1;
  vp.opt = DevOptions<T>(); vp.rec = nullptr; vp.x = nullptr; vp.last_dx = nullptr; vp.results = nullptr; vp.n_active = nullptr;
  vp.lambda = lambda; vp.dx = dx; vp.cost_out = cost; vp.status = status; vp.use_ldlt = 1;
  CK(launch_gn_solve<T>(vp, b.solve_grid, ctx->stream));
  ctx->launches++;
  return TOB200_OK;
}

int check_options(tob200_ctx *ctx, const tob200_options *o) {
  if (!o) return fail(ctx, TOB200_ERR_INVALID, "options is NULL");
  if (o->solver_type != 0 && o->solver_type != 1)
    return fail(ctx, TOB200_ERR_UNSUPPORTED, "solver_type must be LevenbergMarquardt (0) or GaussNewton (1)");
  if (o->max_iters < 0 || o->max_iters > 65535) return fail(ctx, TOB200_ERR_INVALID, "max_iters out of range");
  if (o->max_consec_failures < 0 || o->max_consec_failures > 255 || o->max_total_failures < 0 ||
      o->max_total_failures > 255)
    return fail(ctx, TOB200_ERR_INVALID, "failure limits are uint8 in tinyopt::Options");
  return TOB200_OK;
}

template <typename T>
int build_solve_impl(tob200_ctx *ctx, const T *J, const T *r, int layout, int64_t B, int m, int n, const T *lambda,
                     T *dx, double *cost, T *H_out, T *g_out, int32_t *status) {
  if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");
  if (B < 0 || m < 0 || n < 1) return fail(ctx, TOB200_ERR_INVALID, "need B >= 0, m >= 0, n >= 1");
  if (B == 0) return TOB200_OK;
  if (!J || !r || !dx || !cost || !status) return fail(ctx, TOB200_ERR_INVALID, "NULL buffer");
  if (!aligned16(J) || !aligned16(r)) return fail(ctx, TOB200_ERR_INVALID, "J and r must be 16-byte aligned");
  DeviceGuard guard(ctx->device);
  int family = tob200_kernel_family(dtype_of<T>(), n);
  if (family == 0) return fail(ctx, TOB200_ERR_UNSUPPORTED, "build_solve: n is above the largest supported size (2048)");
  if (family == 3 && ctx->lg_exact) family = 4;  // tob200_set_exact(ctx, 2): the bit-exact general family
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = to_native_layout<T>(ctx, family, layout, B, m, n, &J, &r);
  if (rc != TOB200_OK) return rc;
  TppLaunch cfg;
  if (family == 4) {
    if ((rc = gn_build_solve<T>(ctx, J, r, B, m, n, lambda, dx, cost, H_out, g_out, status)) != TOB200_OK) return rc;
  } else if (family == 1) {
    TppBuildSolveParams<T> p;
    if ((rc = tpp_configure<T>(ctx, n, m, B, kTppBuildSolve, &p.d, &cfg)) != TOB200_OK) return rc;
    p.d.J = J;
    p.d.r = r;
    p.lambda = lambda;
    p.dx = dx;
    p.cost = cost;
    p.H_out = H_out;
    p.g_out = g_out;
    p.status = status;
    CK(tpp_entry_for<T>(n)(kTppLaunch, n, kTppBuildSolve, &p, cfg, nullptr));
    ctx->launches++;
  } else if (family == 3) {
    if (sizeof(T) != 4) return fail(ctx, TOB200_ERR_UNSUPPORTED, "large-n kernels are float only");
    if (n % 4 != 0) {  // zero-padded copy, see lm_run_impl
      const int n4 = (n + 3) & ~3;
      if ((rc = ensure_scratch(ctx, 21, (size_t)B * m * n4 * 4)) != TOB200_OK) return rc;
      if ((rc = ensure_scratch(ctx, 22, (size_t)B * n4 * 4)) != TOB200_OK) return rc;
      if (H_out && (rc = ensure_scratch(ctx, 23, (size_t)B * n4 * n4 * 4)) != TOB200_OK) return rc;
      if (g_out && (rc = ensure_scratch(ctx, 20, (size_t)B * n4 * 4)) != TOB200_OK) return rc;
      float *J4 = (float *)ctx->scratch[21], *dx4 = (float *)ctx->scratch[22];
      float *H4 = H_out ? (float *)ctx->scratch[23] : nullptr, *g4 = g_out ? (float *)ctx->scratch[20] : nullptr;
      CK(launch_repitch((const float *)J, B, m, n, J4, m, n4, ctx->stream));
      CK(cudaMemsetAsync(dx4, 0, (size_t)B * n4 * 4, ctx->stream));
      if ((rc = lg_build_solve(ctx, J4, (const float *)r, B, m, n4, (const float *)lambda, dx4, cost, H4, g4, status)) !=
          TOB200_OK)
        return rc;
      // rejected problems leave dx untouched (the caller's contract): copy only into the accepted ones
      CK(launch_repitch_masked(dx4, status, B, n4, (float *)dx, n, ctx->stream));
      if (H_out) CK(launch_repitch(H4, B, n4, n4, (float *)H_out, n, n, ctx->stream));
      if (g_out) CK(launch_repitch(g4, 1, (int)B, n4, (float *)g_out, (int)B, n, ctx->stream));
      ctx->launches += 3;
    } else if ((rc = lg_build_solve(ctx, (const float *)J, (const float *)r, B, m, n, (const float *)lambda, (float *)dx,
                                    cost, (float *)H_out, (float *)g_out, status)) != TOB200_OK) {
      return rc;
    }
  } else if (wtc_takes(ctx, (int)sizeof(T), m, n, J)) {
    // the tensor-core kernel in its Build + Solve mode (materialised J, r)
    WtcParams p{};
    int64_t grid = 0;
    if ((rc = wtc_configure(ctx, (const float *)J, (const float *)r, B, m, n, &p, &grid)) != TOB200_OK) return rc;
    tob200_options dflt;
    tob200_options_default(&dflt);
    p.opt = make_dev_options<float>(dflt);
    p.mode = 1;
    p.lambda = (const float *)lambda;
    p.dx = (float *)dx;
    p.cost_out = cost;
    p.H_out = (float *)H_out;
    p.g_out = (float *)g_out;
    p.status = status;
    CK(launch_wtc_lm_run(p, (int)grid, ctx->stream));
    ctx->launches++;
  } else {
    WppBuildSolveParams<T> p;
    if ((rc = wpp_configure<T>(ctx, n, m, B, kWppBuildSolve, J, r, &p.d, &cfg)) != TOB200_OK) return rc;
    p.lambda = lambda;
    p.dx = dx;
    p.cost = cost;
    p.H_out = H_out;
    p.g_out = g_out;
    p.status = status;
    if ((rc = wpp_launch<T>(ctx, n, kWppBuildSolve, &p, cfg)) != TOB200_OK) return rc;
  }
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  return TOB200_OK;
}

template <typename T>
int lm_run_impl(tob200_ctx *ctx, const tob200_options *opt, const T *A, const T *y, T alpha, int layout, int64_t B,
                int m, int n, T *x, tob200_result *results, bool record_events = true, double *final_hessian = nullptr) {
  if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");
  int rc = check_options(ctx, opt);
  if (rc != TOB200_OK) return rc;
  if (B < 0 || m < 0 || n < 1) return fail(ctx, TOB200_ERR_INVALID, "need B >= 0, m >= 0, n >= 1");
  if (B == 0) return TOB200_OK;
  if (!A || !y || !x || !results) return fail(ctx, TOB200_ERR_INVALID, "NULL buffer");
  if (!aligned16(A) || !aligned16(y)) return fail(ctx, TOB200_ERR_INVALID, "A and y must be 16-byte aligned");
  DeviceGuard guard(ctx->device);
  int family = tob200_kernel_family(dtype_of<T>(), n);
  if (family == 0) return fail(ctx, TOB200_ERR_UNSUPPORTED, "lm_run: n is above the largest supported size (2048)");
  if (!opt->use_ldlt && family == 3) family = 4;  // hessian.use_ldlt = false (H.inverse()): the general family's LU
  if (family == 3 && ctx->lg_exact) family = 4;   // tob200_set_exact(ctx, 2): the bit-exact general family
  if (record_events) CK(cudaEventRecord(ctx->ev0, ctx->stream));
  if ((rc = to_native_layout<T>(ctx, family, layout, B, m, n, &A, &y)) != TOB200_OK) return rc;
  if (!opt->save_last) final_hessian = nullptr;  // options.h:66 (hessian.save_last)
  TppLaunch cfg;
  if (family == 4) {
    if ((rc = gn_lm_run<T>(ctx, opt, A, y, alpha, B, m, n, x, results, final_hessian)) != TOB200_OK) return rc;
  } else if (family == 1) {
    TppRunParams<T> p;
    if ((rc = tpp_configure<T>(ctx, n, m, B, kTppRun, &p.d, &cfg)) != TOB200_OK) return rc;
    p.d.J = A;
    p.d.r = y;
    p.opt = make_dev_options<T>(*opt);
    p.alpha = alpha;
    p.alpha3 = (T)3 * alpha;
    p.x = x;
    p.results = results;
    p.final_hessian = final_hessian;
    CK(tpp_entry_for<T>(n)(kTppLaunch, n, kTppRun, &p, cfg, nullptr));
    ctx->launches++;
  } else if (family == 3) {
    if (sizeof(T) != 4) return fail(ctx, TOB200_ERR_UNSUPPORTED, "large-n kernels are float only");
    if (n % 4 != 0 && opt->check_min_H_diag > 0.f)
      return fail(ctx, TOB200_ERR_UNSUPPORTED, "check_min_H_diag with n % 4 != 0 (n > 55): the zero pad columns would trip it");
    if (n % 4 != 0) {
      // The TMA / vector paths want 16-byte aligned rows: run the problem with n4 = n rounded up to 4 on a
      // zero-padded copy of A and x.  The pad columns of J are zero, so H gets zero rows / columns there:
      // they sort last in Eigen's pivot order, the leading n x n factorisation is untouched, the zero
      // pivots are legal (zero column below them) and D^+ gives dx = 0 for the pads.
      const int n4 = (n + 3) & ~3;
      if ((rc = ensure_scratch(ctx, 21, (size_t)B * m * n4 * 4)) != TOB200_OK) return rc;
      if ((rc = ensure_scratch(ctx, 22, (size_t)B * n4 * 4)) != TOB200_OK) return rc;
      float *A4 = (float *)ctx->scratch[21], *x4 = (float *)ctx->scratch[22];
      CK(launch_repitch((const float *)A, B, m, n, A4, m, n4, ctx->stream));
      CK(launch_repitch((const float *)x, 1, (int)B, n, x4, (int)B, n4, ctx->stream));
      ctx->launches += 2;
      if ((rc = lg_lm_run(ctx, opt, A4, (const float *)y, (float)alpha, B, m, n4, x4, results, final_hessian, n)) != TOB200_OK) return rc;
      CK(launch_repitch(x4, 1, (int)B, n4, (float *)x, (int)B, n, ctx->stream));
      ctx->launches++;
    } else if ((rc = lg_lm_run(ctx, opt, (const float *)A, (const float *)y, (float)alpha, B, m, n, (float *)x,
                               results, final_hessian, n)) != TOB200_OK) {
      return rc;
    }
  } else if (opt->use_ldlt && !final_hessian && wtc_takes(ctx, (int)sizeof(T), m, n, A)) {
    // mid-n tensor-core family (wtc.cuh): one persistent CTA per SM, eight problems in flight each
    WtcParams p{};
    int64_t grid = 0;
    if ((rc = wtc_configure(ctx, (const float *)A, (const float *)y, B, m, n, &p, &grid)) != TOB200_OK) return rc;
    p.x = (float *)x;
    p.results = results;
    p.opt = make_dev_options<float>(*opt);
    p.alpha = (float)alpha;
    p.alpha3 = 3.f * (float)alpha;
    p.mode = 0;
    CK(launch_wtc_lm_run(p, (int)grid, ctx->stream));
    ctx->launches++;
  } else {
    WppRunParams<T> p;
    const int kind = opt->use_ldlt ? kWppRun : kWppRunInv;  // options.h:59
    if ((rc = wpp_configure<T>(ctx, n, m, B, kind, A, y, &p.d, &cfg)) != TOB200_OK) return rc;
    p.opt = make_dev_options<T>(*opt);
    p.alpha = alpha;
    p.alpha3 = (T)3 * alpha;
    p.x = x;
    p.results = results;
    p.final_hessian = final_hessian;
    if ((rc = wpp_launch<T>(ctx, n, kind, &p, cfg)) != TOB200_OK) return rc;
  }
  if (record_events) CK(cudaEventRecord(ctx->ev1, ctx->stream));
  return TOB200_OK;
}

template <typename T>
int lm_run_host_impl(tob200_ctx *ctx, const tob200_options *opt, const T *A, const T *y, T alpha, int layout,
                     int64_t B, int m, int n, T *x, tob200_result *results) {
  if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");
  {  // before any allocation or copy is queued
    const int rc0 = check_options(ctx, opt);
    if (rc0 != TOB200_OK) return rc0;
  }
  if (B < 0 || m < 0 || n < 1) return fail(ctx, TOB200_ERR_INVALID, "need B >= 0, m >= 0, n >= 1");
  if (B == 0) return TOB200_OK;
  if (!A || !y || !x || !results) return fail(ctx, TOB200_ERR_INVALID, "NULL buffer");
  if (layout != TOB200_LAYOUT_TILE32 && layout != TOB200_LAYOUT_PROBLEM_MAJOR)
    return fail(ctx, TOB200_ERR_INVALID, "unknown layout");
  if (tob200_kernel_family(dtype_of<T>(), n) == 0)
    return fail(ctx, TOB200_ERR_UNSUPPORTED, "lm_run: n has no kernel yet for this dtype");
  DeviceGuard guard(ctx->device);
  const bool tiled = layout == TOB200_LAYOUT_TILE32;
  const size_t ea = tiled ? (size_t)tob200_tiled_elems(B, m, n) : (size_t)B * m * n;
  const size_t ey = tiled ? (size_t)tob200_tiled_elems(B, m, 1) : (size_t)B * m;
  int rc;
  if ((rc = ensure_scratch(ctx, 2, ea * sizeof(T))) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, 3, ey * sizeof(T))) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, 4, (size_t)B * n * sizeof(T))) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, 5, (size_t)B * sizeof(tob200_result))) != TOB200_OK) return rc;
  // Chunked three-stage pipeline (problems are independent): every host->device copy is queued up
  // front on its own stream, the solve of chunk c starts when its inputs have landed and overlaps the
  // upload of chunk c + 1, and its results go back on a third stream while the next chunk runs (PCIe is
  // full duplex).  Chunks are tile aligned; small inputs go in one piece.
  if (!ctx->h2d_stream) {
    CK(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    for (int c = 0; c < tob200_ctx::kMaxChunks; ++c) {
      CK(cudaEventCreateWithFlags(&ctx->ev_h2d[c], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&ctx->ev_run[c], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&ctx->ev_side, cudaEventDisableTiming));
  }
  int nch = (ea * sizeof(T) >= ((size_t)32 << 20)) ? ctx->host_chunks : 1;
  if (nch > tob200_ctx::kMaxChunks) nch = tob200_ctx::kMaxChunks;
  if (nch < 1) nch = 1;
  const int64_t per = ((B + nch - 1) / nch + 31) / 32 * 32;
  T *dA = (T *)ctx->scratch[2], *dy = (T *)ctx->scratch[3], *dx = (T *)ctx->scratch[4];
  tob200_result *dres = (tob200_result *)ctx->scratch[5];
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  // the scratch buffers may still be read by earlier work on the compute stream
  CK(cudaEventRecord(ctx->ev_side, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->h2d_stream, ctx->ev_side, 0));
  CK(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_side, 0));
  int used = 0;
  for (int c = 0; c < nch; ++c, ++used) {
    const int64_t p0 = (int64_t)c * per;
    const int64_t pb = (B - p0 < per) ? (B - p0) : per;
    if (pb <= 0) break;
    const size_t oa = tiled ? (size_t)(p0 / 32) * m * n * 32 : (size_t)p0 * m * n;
    const size_t oy = tiled ? (size_t)(p0 / 32) * m * 32 : (size_t)p0 * m;
    const size_t ca = tiled ? (size_t)tob200_tiled_elems(pb, m, n) : (size_t)pb * m * n;
    const size_t cy = tiled ? (size_t)tob200_tiled_elems(pb, m, 1) : (size_t)pb * m;
    CK(cudaMemcpyAsync(dA + oa, A + oa, ca * sizeof(T), cudaMemcpyHostToDevice, ctx->h2d_stream));
    CK(cudaMemcpyAsync(dy + oy, y + oy, cy * sizeof(T), cudaMemcpyHostToDevice, ctx->h2d_stream));
    CK(cudaMemcpyAsync(dx + (size_t)p0 * n, x + (size_t)p0 * n, (size_t)pb * n * sizeof(T), cudaMemcpyHostToDevice,
                       ctx->h2d_stream));
    CK(cudaEventRecord(ctx->ev_h2d[c], ctx->h2d_stream));
  }
  for (int c = 0; c < used; ++c) {
    const int64_t p0 = (int64_t)c * per;
    const int64_t pb = (B - p0 < per) ? (B - p0) : per;
    const size_t oa = tiled ? (size_t)(p0 / 32) * m * n * 32 : (size_t)p0 * m * n;
    const size_t oy = tiled ? (size_t)(p0 / 32) * m * 32 : (size_t)p0 * m;
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[c], 0));
    rc = lm_run_impl<T>(ctx, opt, dA + oa, dy + oy, alpha, layout, pb, m, n, dx + (size_t)p0 * n, dres + p0, false);
    if (rc != TOB200_OK) {
      cudaStreamSynchronize(ctx->h2d_stream);
      cudaStreamSynchronize(ctx->d2h_stream);
      return rc;
    }
    CK(cudaEventRecord(ctx->ev_run[c], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_run[c], 0));
    CK(cudaMemcpyAsync(x + (size_t)p0 * n, dx + (size_t)p0 * n, (size_t)pb * n * sizeof(T), cudaMemcpyDeviceToHost,
                       ctx->d2h_stream));
    CK(cudaMemcpyAsync(results + p0, dres + p0, (size_t)pb * sizeof(tob200_result), cudaMemcpyDeviceToHost,
                       ctx->d2h_stream));
  }
  CK(cudaEventRecord(ctx->ev_side, ctx->d2h_stream));
  CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0));
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return TOB200_OK;
}

}  // namespace

// ================================================================================================
struct tob200_solver {
  tob200_ctx *ctx = nullptr;
  // copies of what destroy needs: the solver may outlive its context (a garbage-collected Python
  // BatchSolver after Context.close()); destroy must not dereference ctx
  int device = 0;
  int dtype = 0, n = 0;
  int64_t B = 0;
  tob200_options opt;
  void *rec = nullptr, *x = nullptr, *last_dx = nullptr, *H = nullptr, *g = nullptr;
  int32_t *needs = nullptr;
  unsigned long long *n_active = nullptr;  // device
  bool is_reset = false;
  int family = 1;          // 1: thread per problem (H_, grad_ tile-interleaved), 2: warp per problem ([B][NP * LDW], [B][NP]),
                           // 4: general family, n > 55 (H_ [B][n][n] upper triangle, grad_ [B][n]; rec = LmScalars<T>)
  size_t h_bytes = 0, g_bytes = 0;
  void *hd = nullptr, *cost = nullptr;  // family 4: the damped diagonal of H_ [B][n], the pass's cost [B]
};

namespace {

template <typename T>
struct StepHG {  // tob200_solver_step_hg_*: caller-filled accumulators instead of J, r
  const T *grad = nullptr, *H = nullptr;
  const double *cost = nullptr;
  const int32_t *nres = nullptr;
};

// The seam above n = 55 on the general kernel family (gn.cuh): gn_accum_kernel forms cost, grad_, H_ from the
// caller's residual blocks (canonical chains, rows in order: bit-identical to the oracle fed with the same J, r),
// gn_solve_kernel runs Build's tail, Solve, Step and the OptimizeAcc update.
template <typename T>
int gn_solver_step(tob200_solver *s, const T *J, const T *r, int m, int reset, const StepHG<T> *hg, const double *cost_in) {
  tob200_ctx *ctx = s->ctx;
  const int n = s->n;
  const int64_t B = s->B;
  const DevOptions<T> dopt = make_dev_options<T>(s->opt);
  LmScalars<T> *rec = (LmScalars<T> *)s->rec;
  if (reset) {
    CK(launch_gn_init<T>(rec, dopt, (T *)s->last_dx, B, n, ctx->stream, s->needs));
    ctx->launches++;
    const unsigned long long all = (unsigned long long)B;
    CK(cudaMemcpyAsync(s->n_active, &all, sizeof(all), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // `all` is a stack variable
    return TOB200_OK;
  }
  const int grid = (int)(B < 2 * ctx->num_sms ? B : 2 * ctx->num_sms);
  int rc;
  if ((rc = ensure_scratch(ctx, kGnW, (size_t)grid * n * n * sizeof(T))) != TOB200_OK) return rc;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  const int is_lm = s->opt.solver_type == 0;
  if (hg) {
    CK(launch_gn_import_hg<T>(hg->grad, hg->H, rec, is_lm, B, n, (T *)s->g, (T *)s->H, ctx->stream));
  } else {
    if ((rc = ensure_scratch(ctx, kGnRs, (size_t)B * 2 * (m > 0 ? m : 1) * sizeof(T))) != TOB200_OK) return rc;
    GnAccumParams<T> ap;
    ap.A = J; ap.y = r; ap.x = nullptr; ap.rec = rec; ap.rs = (T *)ctx->scratch[kGnRs]; ap.g = (T *)s->g; ap.H = (T *)s->H;
    ap.cost = (T *)s->cost; ap.B = B; ap.m = m; ap.n = n; ap.synth = 0; ap.is_lm = is_lm; ap.alpha = (T)0; ap.alpha3 = (T)0;
    CK(launch_gn_accum<T>(ap, ctx->num_sms, ctx->stream));
  }
  ctx->launches++;
  GnSolveParams<T> vp;
  vp.H = (T *)s->H; vp.hd = (T *)s->hd; vp.g = (T *)s->g; vp.cost = (const T *)s->cost; vp.W = (T *)ctx->scratch[kGnW];
  vp.B = B; vp.n = n; vp.nres = m; vp.mode = 0; vp.opt = dopt; vp.rec = rec; vp.x = (T *)s->x; vp.last_dx = (T *)s->last_dx;
  vp.results = nullptr; vp.n_active = s->n_active; vp.lambda = nullptr; vp.dx = nullptr; vp.cost_out = nullptr;
  vp.status = nullptr; vp.use_ldlt = s->opt.use_ldlt; vp.needs = s->needs;
  if (hg) { vp.cost_d = hg->cost; vp.nres_arr = hg->nres; }
  else if (cost_in) vp.cost_d = cost_in;
  CK(cudaMemsetAsync(s->n_active, 0, sizeof(unsigned long long), ctx->stream));
  CK(launch_gn_solve<T>(vp, grid, ctx->stream));
  ctx->launches++;
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  return TOB200_OK;
}

template <typename T>
int solver_step_impl(tob200_solver *s, const T *J, const T *r, int layout, int m, int reset, const StepHG<T> *hg = nullptr,
                     const double *cost_in = nullptr) {
  if (!s) return fail(nullptr, TOB200_ERR_INVALID, "solver is NULL");
  tob200_ctx *ctx = s->ctx;
  if (s->dtype != dtype_of<T>()) return fail(ctx, TOB200_ERR_INVALID, "solver dtype mismatch");
  DeviceGuard guard(ctx->device);
  const int n = s->n;
  const int64_t B = s->B;
  if (hg) {
    if (!s->is_reset) return fail(ctx, TOB200_ERR_INVALID, "tob200_solver_reset must be called first");
    if (!hg->grad || !hg->H || !hg->cost || !hg->nres) return fail(ctx, TOB200_ERR_INVALID, "NULL buffer");
    J = nullptr;
    r = nullptr;
    m = 1;
  } else if (!reset) {
    if (!s->is_reset) return fail(ctx, TOB200_ERR_INVALID, "tob200_solver_reset must be called first");
    if (m < 0) return fail(ctx, TOB200_ERR_INVALID, "m < 0");
    if (!J || !r) return fail(ctx, TOB200_ERR_INVALID, "NULL buffer");
    if (!aligned16(J) || !aligned16(r)) return fail(ctx, TOB200_ERR_INVALID, "J and r must be 16-byte aligned");
    int rc = to_native_layout<T>(ctx, s->family, layout, B, m, n, &J, &r);
    if (rc != TOB200_OK) return rc;
  }
  if (s->family == 4) return gn_solver_step<T>(s, J, r, m, reset, hg, cost_in);
  if (cost_in) return fail(ctx, TOB200_ERR_UNSUPPORTED, "tob200_solver_step_cost: general family only (TOB200_SOLVER_GENERAL)");
  if (s->family == 2) {  // warp per problem (wpp_step.cuh)
    WppStepParams<T> p;
    TppLaunch cfg;
    const int kind = s->opt.use_ldlt ? kWppStep : kWppStepInv;  // options.h:59
    int rc = wpp_configure<T>(ctx, n, reset ? 1 : m, B, kind, J, r, &p.d, &cfg);
    if (rc != TOB200_OK) return rc;
    p.opt = make_dev_options<T>(s->opt);
    p.rec = (StateRec<T> *)s->rec;
    p.x = (T *)s->x;
    p.last_dx = (T *)s->last_dx;
    p.H = (T *)s->H;
    p.g = (T *)s->g;
    p.needs = s->needs;
    p.n_active = s->n_active;
    p.reset = reset;
    p.hg_grad = hg ? hg->grad : nullptr;
    p.hg_H = hg ? hg->H : nullptr;
    p.hg_cost = hg ? hg->cost : nullptr;
    p.hg_nres = hg ? hg->nres : nullptr;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    CK(cudaMemsetAsync(s->n_active, 0, sizeof(unsigned long long), ctx->stream));
    if ((rc = wpp_launch<T>(ctx, n, kind, &p, cfg)) != TOB200_OK) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    return TOB200_OK;
  }
  TppStepParams<T> p;
  TppLaunch cfg;
  int rc = tpp_configure<T>(ctx, n, reset ? 1 : m, B, kTppStep, &p.d, &cfg);
  if (rc != TOB200_OK) return rc;
  p.d.J = J;
  p.d.r = r;
  p.opt = make_dev_options<T>(s->opt);
  p.rec = (StateRec<T> *)s->rec;
  p.x = (T *)s->x;
  p.last_dx = (T *)s->last_dx;
  p.H = (T *)s->H;
  p.g = (T *)s->g;
  p.needs = s->needs;
  p.n_active = s->n_active;
  p.reset = reset;
  p.hg_grad = hg ? hg->grad : nullptr;
  p.hg_H = hg ? hg->H : nullptr;
  p.hg_cost = hg ? hg->cost : nullptr;
  p.hg_nres = hg ? hg->nres : nullptr;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  CK(cudaMemsetAsync(s->n_active, 0, sizeof(unsigned long long), ctx->stream));
  CK(tpp_entry_for<T>(n)(kTppLaunch, n, kTppStep, &p, cfg, nullptr));
  ctx->launches++;
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  return TOB200_OK;
}

}  // namespace

template <typename T>
__global__ void results_kernel(const StateRec<T> *rec, int64_t B, tob200_result *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const StateRec<T> r = rec[i];
  tob200_result o;
  o.final_cost = r.final_cost;
  o.final_rerr_dec = r.final_rerr_dec;
  o.last_lambda = (double)r.lambda;
  o.last_prev_lambda = (double)r.prev_lambda;
  o.final_num_residuals = r.final_nres;
  // a problem that is still running reports kNone, or kMaxIters semantics are applied by the loop
  o.stop_reason = r.stop_reason;
  o.num_iters = r.num_iters;
  o.num_failures = r.num_failures;
  o.num_consec_failures = r.num_consec_failures;
  o.num_builds = r.num_builds;
  out[i] = o;
}

// ---- InvCov / MaxStdDev from the same factorisation -------------------------
namespace {
template <typename T>
int inv_cov_impl(tob200_ctx *ctx, const T *H, int64_t B, int n, T *cov, T *max_std, int32_t *status) {
  if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");
  if (B < 0 || n < 1) return fail(ctx, TOB200_ERR_INVALID, "need B >= 0, n >= 1");
  if (B == 0) return TOB200_OK;
  if (!H || !status || (!cov && !max_std)) return fail(ctx, TOB200_ERR_INVALID, "NULL buffer");
  DeviceGuard guard(ctx->device);
  if (n <= 64) {
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    CK(launch_cov_warp<T>(H, B, n, cov, max_std, status, ctx->num_sms, ctx->stream));
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    return TOB200_OK;
  }
  if (sizeof(T) != 4 || n > kLgMaxN) {  // general family: double above n = 64, float above n = 512 (gn.cuh, mode 3)
    if (n > kGnMaxN) return fail(ctx, TOB200_ERR_UNSUPPORTED, "InvCov: n above 2048 has no kernel");
    const int grid = (int)(B < 2 * ctx->num_sms ? B : 2 * ctx->num_sms);
    int rc;
    if ((rc = ensure_scratch(ctx, kGnW, (size_t)grid * n * n * sizeof(T))) != TOB200_OK) return rc;
    if ((rc = ensure_scratch(ctx, 24, (size_t)grid * n * n * sizeof(T))) != TOB200_OK) return rc;
    GnSolveParams<T> vp;
    vp.H = const_cast<T *>(H); vp.hd = nullptr; vp.g = nullptr; vp.cost = nullptr; vp.W = (T *)ctx->scratch[kGnW]; vp.B = B;
    vp.n = n; vp.nres = 0; vp.mode = 3; vp.opt = DevOptions<T>(); vp.rec = nullptr; vp.x = nullptr; vp.last_dx = nullptr;
    vp.results = nullptr; vp.n_active = nullptr; vp.lambda = nullptr; vp.dx = nullptr; vp.cost_out = nullptr; vp.status = status;
    vp.use_ldlt = 1; vp.Y = (T *)ctx->scratch[24]; vp.cov = cov; vp.max_std = max_std;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    CK(launch_gn_solve<T>(vp, grid, ctx->stream));
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    return TOB200_OK;
  }
  LgBuffers b;
  int rc = lg_prepare(ctx, B, 0, n, false, false, &b);
  if (rc != TOB200_OK) return rc;
  ctx->phase_used = 0;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  CK(launch_lg_import_h((const float *)H, B, n, b.np, b.H, ctx->stream));
  LgSolveParams vp;
  vp.H = b.H; vp.dg = nullptr; vp.hd = nullptr; vp.g = nullptr; vp.cost = nullptr; vp.W = b.W; vp.B = B; vp.n = n; vp.np = b.np; vp.nres = 0;
  vp.mode = 3; vp.opt = DevOptions<float>(); vp.rec = nullptr; vp.x = nullptr; vp.last_dx = nullptr; vp.results = nullptr;
  vp.n_active = nullptr; vp.lambda = nullptr; vp.b = nullptr; vp.dx = (float *)cov; vp.cost_out = nullptr; vp.status = status;
  vp.max_std = (float *)max_std;
  int ev = lg_phase_begin(ctx, 2);
  CK(launch_lg_solve(vp, b.solve_grid, ctx->stream));
  lg_phase_end(ctx, ev);
  ctx->launches += 2;
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  return TOB200_OK;
}
}  // namespace

// ================================================================================================
extern "C" {

int tob200_version(void) { return TOB200_VERSION; }

void tob200_options_default(tob200_options *o) {  // optimizers/options.h defaults
  if (!o) return;
  o->solver_type = 0;
  o->check_final_cost = 0;
  o->use_step_quality_approx = 0;
  o->grad_clipping = 0;
  o->use_ldlt = 1;
  o->H_is_full = 1;
  o->check_min_H_diag = 0;
  o->save_last = 1;
  o->use_squared_norm = 1;
  o->downscale_by_2 = 0;
  o->normalize = 0;
  o->max_iters = 50;
  o->min_error = 1e-12f;
  o->min_rerr_dec = 1e-10f;
  o->min_step_norm2 = 1e-14f;
  o->min_grad_norm2 = 1e-18f;
  o->max_total_failures = 0;
  o->max_consec_failures = 5;
  o->damping_init = 1e-4f;
  o->damping_min = 1e-9f;
  o->damping_max = 1e9f;
  o->good_factor = 1.0f / 3.0f;
  o->bad_factor = 2.0f;
}

int64_t tob200_tiled_elems(int64_t B, int m, int n) {
  if (B <= 0 || m <= 0 || n <= 0) return 0;
  return ((B + kTile - 1) / kTile) * kTile * (int64_t)m * n;
}

int tob200_kernel_family(int dtype, int n) {
  if (n < 1) return 0;
  if (dtype == TOB200_F32) {
    if (n <= kTppMaxN_f32) return 1;
    if (n <= kWppMaxN_f32) return 2;
    if (n <= kLgMaxN) return 3;  // n % 4 != 0 goes through a zero-padded copy (rows 16-byte aligned for TMA)
    return n <= kGnMaxN ? 4 : 0;
  }
  if (dtype == TOB200_F64) return n <= kTppMaxN_f64 ? 1 : (n <= kWppMaxN_f64 ? 2 : (n <= kGnMaxN ? 4 : 0));
  return 0;
}

int tob200_create(tob200_ctx **out, int device, void *stream) {
  if (!out) return fail(nullptr, TOB200_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, TOB200_ERR_CUDA,
                std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                    " (tinyopt_b200 has no CPU fallback)");
  if (device < 0 || device >= count) return fail(nullptr, TOB200_ERR_INVALID, "device ordinal out of range");
  tob200_ctx *ctx = new (std::nothrow) tob200_ctx();
  if (!ctx) return fail(nullptr, TOB200_ERR_NOMEM, "host allocation failed");
  ctx->device = device;
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    delete ctx;
    return fail_cuda(nullptr, e, "cudaGetDeviceProperties");
  }
  if (prop.major < 10) {
    delete ctx;
    return fail(nullptr, TOB200_ERR_UNSUPPORTED, "tinyopt_b200 is built for sm_100a (Blackwell B200) only");
  }
  ctx->num_sms = prop.multiProcessorCount;
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
      delete ctx;
      return fail_cuda(nullptr, e, "cudaStreamCreate");
    }
    ctx->own_stream = true;
  }
  cudaEventCreate(&ctx->ev0);
  cudaEventCreate(&ctx->ev1);
  ctx->tpp_stage_bytes = env_int("TOB200_TPP_STAGE_BYTES", ctx->tpp_stage_bytes);
  ctx->tpp_stages = env_int("TOB200_TPP_STAGES", ctx->tpp_stages);
  ctx->tpp_ctas_per_sm = env_int("TOB200_TPP_CTAS_PER_SM", 0);
  ctx->wpp_stages = env_int("TOB200_WPP_STAGES", ctx->wpp_stages);
  ctx->wpp_tc = env_int("TOB200_WPP_TC", ctx->wpp_tc);
  ctx->wtc_prefetch = env_int("TOB200_WTC_PREFETCH", ctx->wtc_prefetch);
  ctx->wtc_debug = env_int("TOB200_WTC_DEBUG", ctx->wtc_debug);
  ctx->wtc_raw_stages = env_int("TOB200_WTC_RAW", 0);
  ctx->wtc_op_stages = env_int("TOB200_WTC_OPS", 0);
  ctx->lg_tf32_terms = env_int("TOB200_LG_TF32_TERMS", ctx->lg_tf32_terms) == 1 ? 1 : 3;
  ctx->lg_fp16 = (env_int("TOB200_LG_FP16", 1) != 0 && ctx->lg_tf32_terms == 3) ? 1 : 0;
  ctx->lg_mc = env_int("TOB200_LG_MC", 0);
  ctx->lg_exact = env_int("TOB200_LG_EXACT", 0);
  ctx->lg_raw_stages = env_int("TOB200_LG_RAW_STAGES", kLgRawStages);  // 2..5 measured equal on C5 (12.43 .. 12.57 ms): not the limiter
  if (ctx->lg_raw_stages < 2) ctx->lg_raw_stages = 2;
  if (ctx->lg_raw_stages > 6) ctx->lg_raw_stages = 6;
  ctx->host_chunks = env_int("TOB200_HOST_CHUNKS", ctx->host_chunks);
  if ((e = cudaMalloc((void **)&ctx->counters, sizeof(unsigned long long) * tob200_ctx::kNumCounters)) != cudaSuccess) {
    tob200_destroy(ctx);
    return fail_cuda(nullptr, e, "cudaMalloc(counters)");
  }
  ctx->next_counter = tob200_ctx::kNumCounters;  // forces the first memset
  *out = ctx;
  return TOB200_OK;
}

int tob200_destroy(tob200_ctx *ctx) {
  if (!ctx) return TOB200_OK;
  DeviceGuard guard(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < tob200_ctx::kScratchSlots; ++i)
    if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  for (cudaEvent_t e : ctx->phase_ev) cudaEventDestroy(e);
  if (ctx->counters) cudaFree(ctx->counters);
  if (ctx->active_host) cudaFreeHost(ctx->active_host);
  for (int i = 0; i < 2; ++i)
    if (ctx->ev_active[i]) cudaEventDestroy(ctx->ev_active[i]);
  if (ctx->h2d_stream) {
    cudaStreamDestroy(ctx->h2d_stream);
    cudaStreamDestroy(ctx->d2h_stream);
    for (int c = 0; c < tob200_ctx::kMaxChunks; ++c) {
      if (ctx->ev_h2d[c]) cudaEventDestroy(ctx->ev_h2d[c]);
      if (ctx->ev_run[c]) cudaEventDestroy(ctx->ev_run[c]);
    }
    if (ctx->ev_side) cudaEventDestroy(ctx->ev_side);
  }
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return TOB200_OK;
}

int tob200_sync(tob200_ctx *ctx) {
  if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");
  DeviceGuard guard(ctx->device);
  CK(cudaStreamSynchronize(ctx->stream));
  return TOB200_OK;
}

const char *tob200_last_error(const tob200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int64_t tob200_launch_count(const tob200_ctx *ctx) { return ctx ? ctx->launches : 0; }

int tob200_set_exact(tob200_ctx *ctx, int exact) {
  if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");
  ctx->wpp_tc = exact ? 0 : 1;
  ctx->lg_exact = exact >= 2 ? 1 : 0;
  return TOB200_OK;
}

int tob200_last_elapsed_ms(tob200_ctx *ctx, float *ms) {
  if (!ctx || !ms) return fail(ctx, TOB200_ERR_INVALID, "NULL argument");
  DeviceGuard guard(ctx->device);
  CK(cudaEventSynchronize(ctx->ev1));
  CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return TOB200_OK;
}

int tob200_device_alloc(tob200_ctx *ctx, size_t bytes, void **out) {
  if (!ctx || !out) return fail(ctx, TOB200_ERR_INVALID, "NULL argument");
  *out = nullptr;
  if (bytes == 0) return TOB200_OK;
  DeviceGuard guard(ctx->device);
  CK(cudaMalloc(out, bytes));
  return TOB200_OK;
}

int tob200_device_free(tob200_ctx *ctx, void *ptr) {
  if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");
  if (!ptr) return TOB200_OK;
  DeviceGuard guard(ctx->device);
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaFree(ptr));
  return TOB200_OK;
}

int tob200_copy_to_device(tob200_ctx *ctx, void *dst, const void *src, size_t bytes) {
  if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");
  if (bytes == 0) return TOB200_OK;
  if (!dst || !src) return fail(ctx, TOB200_ERR_INVALID, "NULL buffer");
  DeviceGuard guard(ctx->device);
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return TOB200_OK;
}

int tob200_copy_to_host(tob200_ctx *ctx, void *dst, const void *src, size_t bytes) {
  if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");
  if (bytes == 0) return TOB200_OK;
  if (!dst || !src) return fail(ctx, TOB200_ERR_INVALID, "NULL buffer");
  DeviceGuard guard(ctx->device);
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return TOB200_OK;
}

#define RETILE(SUF, T)                                                                                  \
  int tob200_retile_##SUF(tob200_ctx *ctx, const T *src, int64_t B, int m, int n, T *dst) {              \
    if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");                                   \
    if (B < 0 || m < 0 || n < 0) return fail(ctx, TOB200_ERR_INVALID, "negative size");                  \
    if (B == 0 || m == 0 || n == 0) return TOB200_OK;                                                    \
    if (!src || !dst) return fail(ctx, TOB200_ERR_INVALID, "NULL buffer");                               \
    DeviceGuard guard(ctx->device);                                                                      \
    CK(launch_retile<T>(src, B, m, n, dst, ctx->stream));                                                \
    ctx->launches++;                                                                                     \
    return TOB200_OK;                                                                                    \
  }
RETILE(f32, float)
RETILE(f64, double)

int tob200_build_solve_f32(tob200_ctx *ctx, const float *J, const float *r, int layout, int64_t B, int m, int n,
                           const float *lambda, float *dx, double *cost, float *H_out, float *g_out,
                           int32_t *status) {
  return build_solve_impl<float>(ctx, J, r, layout, B, m, n, lambda, dx, cost, H_out, g_out, status);
}
int tob200_build_solve_f64(tob200_ctx *ctx, const double *J, const double *r, int layout, int64_t B, int m, int n,
                           const double *lambda, double *dx, double *cost, double *H_out, double *g_out,
                           int32_t *status) {
  return build_solve_impl<double>(ctx, J, r, layout, B, m, n, lambda, dx, cost, H_out, g_out, status);
}

int tob200_jtj_f32(tob200_ctx *ctx, const float *J, const float *row_scale, int64_t B, int m, int n, float *H) {
  if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");
  if (B < 0 || m < 0 || n < 1 || n > kLgMaxN) return fail(ctx, TOB200_ERR_INVALID, "need 1 <= n <= 512");
  if (B == 0) return TOB200_OK;
  if (!J || !H) return fail(ctx, TOB200_ERR_INVALID, "NULL buffer");
  if (!aligned16(J)) return fail(ctx, TOB200_ERR_INVALID, "J must be 16-byte aligned");
  DeviceGuard guard(ctx->device);
  if (n % 4 != 0) {  // zero-padded copy (16-byte aligned rows for the TMA path), result stripped again
    const int n4 = (n + 3) & ~3;
    int rc;
    if ((rc = ensure_scratch(ctx, 21, (size_t)B * m * n4 * 4)) != TOB200_OK) return rc;
    if ((rc = ensure_scratch(ctx, 23, (size_t)B * n4 * n4 * 4)) != TOB200_OK) return rc;
    float *J4 = (float *)ctx->scratch[21], *H4 = (float *)ctx->scratch[23];
    CK(launch_repitch(J, B, m, n, J4, m, n4, ctx->stream));
    if ((rc = tob200_jtj_f32(ctx, J4, row_scale, B, m, n4, H4)) != TOB200_OK) return rc;
    CK(launch_repitch(H4, B, n4, n4, H, n, n, ctx->stream));
    ctx->launches += 2;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    return TOB200_OK;
  }
  LgBuffers b;
  int rc = lg_prepare(ctx, B, m, n, false, false, &b);
  if (rc != TOB200_OK) return rc;
  ctx->phase_used = 0;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  // the diagonal in FP32 (lg.cuh: the tensor core truncates its long same-sign sums) and max |J_ij| per problem
  // (the power-of-two scale of the FP16 split) first, then the tensor-core product
  const float *dg = nullptr;
  int ev;
  if (m > 0) {
    LgEvalParams ep;
    ep.A = J; ep.y = nullptr; ep.x = nullptr; ep.rec = nullptr; ep.scale_in = row_scale; ep.scale = nullptr; ep.g = b.g; ep.dg = b.dg; ep.cost = b.cost; ep.amax = b.amax;
    ep.B = B; ep.m = m; ep.n = n; ep.synth = 0; ep.is_lm = 1; ep.alpha = 0.f; ep.alpha3 = 0.f;
    ev = lg_phase_begin(ctx, 0);
    CK(launch_lg_eval(ep, ctx->num_sms, ctx->stream));
    lg_phase_end(ctx, ev);
    ctx->launches++;
    dg = b.dg;
  }
  ev = lg_phase_begin(ctx, 1);
  if (m > 0) {
    const LgSyrkParams sp = lg_syrk_params(ctx, b, J, row_scale, nullptr, B, m, n, 1);
    CK(launch_lg_syrk(sp, ctx->num_sms, ctx->stream));
  } else {
    CK(cudaMemsetAsync(b.H, 0, (size_t)B * b.np * b.np * 4, ctx->stream));
  }
  lg_phase_end(ctx, ev);
  CK(launch_lg_export_h(b.H, dg, nullptr, B, n, b.np, H, ctx->stream));
  ctx->launches += 2;
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  return TOB200_OK;
}

int tob200_solve_ldlt_f32(tob200_ctx *ctx, const float *A, const float *bvec, int64_t B, int n, float *x,
                          int32_t *status) {
  if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");
  if (B < 0 || n < 1 || n > kLgMaxN) return fail(ctx, TOB200_ERR_INVALID, "need 1 <= n <= 512");
  if (B == 0) return TOB200_OK;
  if (!A || !bvec || !x || !status) return fail(ctx, TOB200_ERR_INVALID, "NULL buffer");
  DeviceGuard guard(ctx->device);
  LgBuffers b;
  int rc = lg_prepare(ctx, B, 0, n, false, false, &b);
  if (rc != TOB200_OK) return rc;
  ctx->phase_used = 0;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  CK(launch_lg_import_h(A, B, n, b.np, b.H, ctx->stream));
  LgSolveParams vp;
  vp.H = b.H; vp.dg = nullptr; vp.hd = nullptr; vp.g = nullptr; vp.cost = nullptr; vp.W = b.W; vp.B = B; vp.n = n; vp.np = b.np; vp.nres = 0;
  vp.mode = 2; vp.opt = DevOptions<float>(); vp.rec = nullptr; vp.x = nullptr; vp.last_dx = nullptr; vp.results = nullptr;
  vp.n_active = nullptr; vp.lambda = nullptr; vp.b = bvec; vp.dx = x; vp.cost_out = nullptr; vp.status = status; vp.max_std = nullptr;
  int ev = lg_phase_begin(ctx, 2);
  CK(launch_lg_solve(vp, b.solve_grid, ctx->stream));
  lg_phase_end(ctx, ev);
  ctx->launches += 2;
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  return TOB200_OK;
}

int tob200_inv_cov_f32(tob200_ctx *ctx, const float *H, int64_t B, int n, float *cov, float *max_std, int32_t *status) {
  return inv_cov_impl<float>(ctx, H, B, n, cov, max_std, status);
}
int tob200_inv_cov_f64(tob200_ctx *ctx, const double *H, int64_t B, int n, double *cov, double *max_std,
                       int32_t *status) {
  return inv_cov_impl<double>(ctx, H, B, n, cov, max_std, status);
}

int tob200_last_phase_ms(tob200_ctx *ctx, int phase, float *ms, int *launches) {
  if (!ctx || !ms) return fail(ctx, TOB200_ERR_INVALID, "NULL argument");
  DeviceGuard guard(ctx->device);
  CK(cudaStreamSynchronize(ctx->stream));
  float total = 0.f;
  int count = 0;
  for (int i = 0; i + 1 < ctx->phase_used; i += 2) {
    if (ctx->phase_kind[i / 2] != phase) continue;
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, ctx->phase_ev[i], ctx->phase_ev[i + 1]));
    total += t;
    ++count;
  }
  *ms = total;
  if (launches) *launches = count;
  return TOB200_OK;
}

int tob200_lm_run_f32(tob200_ctx *ctx, const tob200_options *opt, const float *A, const float *y, float alpha,
                      int layout, int64_t B, int m, int n, float *x, tob200_result *results) {
  return lm_run_impl<float>(ctx, opt, A, y, alpha, layout, B, m, n, x, results);
}
int tob200_lm_run_f64(tob200_ctx *ctx, const tob200_options *opt, const double *A, const double *y, double alpha,
                      int layout, int64_t B, int m, int n, double *x, tob200_result *results) {
  return lm_run_impl<double>(ctx, opt, A, y, alpha, layout, B, m, n, x, results);
}
int tob200_lm_run_ex_f32(tob200_ctx *ctx, const tob200_options *opt, const float *A, const float *y, float alpha,
                         int layout, int64_t B, int m, int n, float *x, tob200_result *results, double *final_hessian) {
  return lm_run_impl<float>(ctx, opt, A, y, alpha, layout, B, m, n, x, results, true, final_hessian);
}
int tob200_lm_run_ex_f64(tob200_ctx *ctx, const tob200_options *opt, const double *A, const double *y, double alpha,
                         int layout, int64_t B, int m, int n, double *x, tob200_result *results, double *final_hessian) {
  return lm_run_impl<double>(ctx, opt, A, y, alpha, layout, B, m, n, x, results, true, final_hessian);
}
int tob200_lm_run_host_f32(tob200_ctx *ctx, const tob200_options *opt, const float *A, const float *y, float alpha,
                           int layout, int64_t B, int m, int n, float *x, tob200_result *results) {
  return lm_run_host_impl<float>(ctx, opt, A, y, alpha, layout, B, m, n, x, results);
}
int tob200_lm_run_host_f64(tob200_ctx *ctx, const tob200_options *opt, const double *A, const double *y,
                           double alpha, int layout, int64_t B, int m, int n, double *x, tob200_result *results) {
  return lm_run_host_impl<double>(ctx, opt, A, y, alpha, layout, B, m, n, x, results);
}

// ---- solver object -------------------------------------------------------------------------------
int tob200_solver_create(tob200_ctx *ctx, int dtype, int64_t B, int n, const tob200_options *opt,
                         tob200_solver **out) {
  return tob200_solver_create_ex(ctx, dtype, B, n, opt, 0, out);
}

int tob200_solver_create_ex(tob200_ctx *ctx, int dtype, int64_t B, int n, const tob200_options *opt, int flags,
                            tob200_solver **out) {
  if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");
  if (!out) return fail(ctx, TOB200_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int rc = check_options(ctx, opt);
  if (rc != TOB200_OK) return rc;
  if (B < 1 || n < 1) return fail(ctx, TOB200_ERR_INVALID, "need B >= 1 and n >= 1");
  int family = tob200_kernel_family(dtype, n);
  if (family == 0) return fail(ctx, TOB200_ERR_UNSUPPORTED, "solver: n above 2048 has no kernel");
  if (family == 3) family = 4;  // the seam above n = 55 runs on the general (bit-exact) family in both precisions
  if (flags & TOB200_SOLVER_GENERAL) family = 4;
  DeviceGuard guard(ctx->device);
  tob200_solver *s = new (std::nothrow) tob200_solver();
  if (!s) return fail(ctx, TOB200_ERR_NOMEM, "host allocation failed");
  s->ctx = ctx;
  s->device = ctx->device;
  s->dtype = dtype;
  s->n = n;
  s->B = B;
  s->opt = *opt;
  const size_t elt = dtype == TOB200_F32 ? 4 : 8;
  const size_t ntiles = (size_t)((B + kTile - 1) / kTile);
  size_t rec_bytes = dtype == TOB200_F32 ? sizeof(StateRec<float>) : sizeof(StateRec<double>);
  if (family == 4) rec_bytes = dtype == TOB200_F32 ? sizeof(LmScalars<float>) : sizeof(LmScalars<double>);
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void **p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
  };
  alloc(&s->rec, rec_bytes * B);
  alloc(&s->x, elt * B * n);
  alloc(&s->last_dx, elt * B * n);
  s->family = family;
  if (family == 1) {
    s->h_bytes = elt * ntiles * tri_count(n) * kTile;
    s->g_bytes = elt * ntiles * n * kTile;
  } else if (family == 2) {
    const int np = wpp_nb_for(n) * wpp_blk_for(n);
    s->h_bytes = elt * (size_t)B * np * wpp_ldw(np);
    s->g_bytes = elt * (size_t)B * np;
  } else {
    s->h_bytes = elt * (size_t)B * n * n;
    s->g_bytes = elt * (size_t)B * n;
    alloc(&s->hd, elt * B * n);
    alloc(&s->cost, elt * B);
  }
  alloc(&s->H, s->h_bytes);
  alloc(&s->g, s->g_bytes);
  alloc((void **)&s->needs, sizeof(int32_t) * B);
  alloc((void **)&s->n_active, sizeof(unsigned long long));
  if (e != cudaSuccess) {
    tob200_solver_destroy(s);
    return fail_cuda(ctx, e, "solver allocation");
  }
  *out = s;
  return TOB200_OK;
}

int tob200_solver_destroy(tob200_solver *s) {
  if (!s) return TOB200_OK;
  DeviceGuard guard(s->device);
  cudaDeviceSynchronize();  // not ctx->stream: the context may already be gone (cudaFree synchronises anyway)
  cudaFree(s->rec);
  cudaFree(s->x);
  cudaFree(s->last_dx);
  cudaFree(s->H);
  cudaFree(s->g);
  cudaFree(s->needs);
  cudaFree(s->n_active);
  cudaFree(s->hd);
  cudaFree(s->cost);
  delete s;
  return TOB200_OK;
}

int tob200_solver_reset(tob200_solver *s, const void *x0) {
  if (!s) return fail(nullptr, TOB200_ERR_INVALID, "solver is NULL");
  tob200_ctx *ctx = s->ctx;
  if (!x0) return fail(ctx, TOB200_ERR_INVALID, "x0 is NULL");
  DeviceGuard guard(ctx->device);
  const size_t elt = s->dtype == TOB200_F32 ? 4 : 8;
  CK(cudaMemcpyAsync(s->x, x0, elt * s->B * s->n, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemsetAsync(s->H, 0, s->h_bytes, ctx->stream));
  CK(cudaMemsetAsync(s->g, 0, s->g_bytes, ctx->stream));
  if (s->hd) CK(cudaMemsetAsync(s->hd, 0, elt * s->B * s->n, ctx->stream));
  int rc = s->dtype == TOB200_F32 ? solver_step_impl<float>(s, nullptr, nullptr, 0, 0, 1)
                                  : solver_step_impl<double>(s, nullptr, nullptr, 0, 0, 1);
  if (rc == TOB200_OK) s->is_reset = true;
  return rc;
}

void *tob200_solver_x(tob200_solver *s) { return s ? s->x : nullptr; }
const int32_t *tob200_solver_needs(tob200_solver *s) { return s ? s->needs : nullptr; }

int tob200_solver_step_f32(tob200_solver *s, const float *J, const float *r, int layout, int m) {
  return solver_step_impl<float>(s, J, r, layout, m, 0);
}
int tob200_solver_step_f64(tob200_solver *s, const double *J, const double *r, int layout, int m) {
  return solver_step_impl<double>(s, J, r, layout, m, 0);
}

int tob200_solver_step_cost_f32(tob200_solver *s, const float *J, const float *r, int layout, int m, const double *cost) {
  if (!cost) return fail(s ? s->ctx : nullptr, TOB200_ERR_INVALID, "cost is NULL");
  return solver_step_impl<float>(s, J, r, layout, m, 0, nullptr, cost);
}
int tob200_solver_step_cost_f64(tob200_solver *s, const double *J, const double *r, int layout, int m, const double *cost) {
  if (!cost) return fail(s ? s->ctx : nullptr, TOB200_ERR_INVALID, "cost is NULL");
  return solver_step_impl<double>(s, J, r, layout, m, 0, nullptr, cost);
}

int tob200_solver_step_hg_f32(tob200_solver *s, const float *grad, const float *H, const double *cost,
                              const int32_t *num_residuals) {
  StepHG<float> hg;
  hg.grad = grad; hg.H = H; hg.cost = cost; hg.nres = num_residuals;
  return solver_step_impl<float>(s, nullptr, nullptr, 0, 0, 0, &hg);
}
int tob200_solver_step_hg_f64(tob200_solver *s, const double *grad, const double *H, const double *cost,
                              const int32_t *num_residuals) {
  StepHG<double> hg;
  hg.grad = grad; hg.H = H; hg.cost = cost; hg.nres = num_residuals;
  return solver_step_impl<double>(s, nullptr, nullptr, 0, 0, 0, &hg);
}

}  // extern "C"

// ---- sparse H in the accumulation signature (optimize.h:27-33 selects the solver from the lambda's H type; tests/sparse.cpp)
// One triplet pattern for the whole batch, values per problem.  Duplicated (row, col) entries are summed in triplet order
// (what Eigen's setFromTriplets does), entries below the diagonal are ignored (`SimplicialLDLT<_, Upper>` reads the upper
// triangle, math.h:270).  The values are scattered into a dense upper triangle and go through the dense pivoted LDL^T of the
// path (tob200_solver_step_hg_*): for the positive definite systems LM produces (J^T J, damped) that is the same solution as
// the sparse factorisation's up to rounding; unlike SimplicialLDLT it REJECTS an indefinite H (the dense reference semantics).
namespace {
template <typename T>
__global__ void sparse_to_dense_kernel(const T *vals, const int32_t *dest, const int32_t *start, const int32_t *order, int64_t B,
                                       int nnz, int nd, int n, T *H) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * nd) return;
  const int64_t pr = e / nd;
  const int d = (int)(e % nd);
  const T *v = vals + (size_t)pr * nnz;
  T sum = v[order[start[d]]];
  for (int k = start[d] + 1; k < start[d + 1]; ++k) sum = Ops<T>::add(sum, v[order[k]]);
  H[(size_t)pr * n * n + dest[d]] = sum;
}

template <typename T>
int step_hg_sparse_impl(tob200_solver *s, const T *grad, const int32_t *rows, const int32_t *cols, int nnz, const T *values,
                        const double *cost, const int32_t *nres) {
  if (!s) return fail(nullptr, TOB200_ERR_INVALID, "solver is NULL");
  tob200_ctx *ctx = s->ctx;
  if (nnz < 0 || (nnz > 0 && (!rows || !cols || !values))) return fail(ctx, TOB200_ERR_INVALID, "bad triplet arrays");
  const int n = s->n;
  const int64_t B = s->B;
  // group the kept triplets by destination, original order inside a group
  std::vector<std::pair<int32_t, int32_t>> key;  // (dest, triplet)
  key.reserve((size_t)nnz);
  for (int k = 0; k < nnz; ++k) {
    if (rows[k] < 0 || cols[k] < 0 || rows[k] >= n || cols[k] >= n) return fail(ctx, TOB200_ERR_INVALID, "triplet index out of range");
    if (rows[k] <= cols[k]) key.emplace_back(rows[k] * n + cols[k], k);
  }
  std::stable_sort(key.begin(), key.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
  std::vector<int32_t> pack;  // dest[nd] | start[nd + 1] | order[kept]
  std::vector<int32_t> dest, start, order;
  for (size_t k = 0; k < key.size(); ++k) {
    if (k == 0 || key[k].first != key[k - 1].first) { dest.push_back(key[k].first); start.push_back((int32_t)k); }
    order.push_back(key[k].second);
  }
  start.push_back((int32_t)key.size());
  const int nd = (int)dest.size();
  pack.insert(pack.end(), dest.begin(), dest.end());
  pack.insert(pack.end(), start.begin(), start.end());
  pack.insert(pack.end(), order.begin(), order.end());
  DeviceGuard guard(ctx->device);
  int rc;
  if ((rc = ensure_scratch(ctx, 25, (size_t)B * n * n * sizeof(T))) != TOB200_OK) return rc;
  if ((rc = ensure_scratch(ctx, 26, pack.size() * sizeof(int32_t) + 16)) != TOB200_OK) return rc;
  T *Hd = (T *)ctx->scratch[25];
  int32_t *dp = (int32_t *)ctx->scratch[26];
  CK(cudaMemsetAsync(Hd, 0, (size_t)B * n * n * sizeof(T), ctx->stream));
  if (nd > 0) {
    CK(cudaMemcpyAsync(dp, pack.data(), pack.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // `pack` is a local
    const int64_t total = B * nd;
    sparse_to_dense_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(values, dp, dp + nd, dp + 2 * nd + 1, B, nnz,
                                                                                       nd, n, Hd);
    CK(cudaGetLastError());
    ctx->launches++;
  }
  StepHG<T> hg;
  hg.grad = grad; hg.H = Hd; hg.cost = cost; hg.nres = nres;
  return solver_step_impl<T>(s, nullptr, nullptr, 0, 0, 0, &hg);
}
}  // namespace

extern "C" {

int tob200_solver_step_hg_sparse_f32(tob200_solver *s, const float *grad, const int32_t *rows, const int32_t *cols, int nnz,
                                     const float *values, const double *cost, const int32_t *num_residuals) {
  return step_hg_sparse_impl<float>(s, grad, rows, cols, nnz, values, cost, num_residuals);
}
int tob200_solver_step_hg_sparse_f64(tob200_solver *s, const double *grad, const int32_t *rows, const int32_t *cols, int nnz,
                                     const double *values, const double *cost, const int32_t *num_residuals) {
  return step_hg_sparse_impl<double>(s, grad, rows, cols, nnz, values, cost, num_residuals);
}

int tob200_solver_num_active(tob200_solver *s, int64_t *n_active) {
  if (!s || !n_active) return fail(s ? s->ctx : nullptr, TOB200_ERR_INVALID, "NULL argument");
  tob200_ctx *ctx = s->ctx;
  DeviceGuard guard(ctx->device);
  unsigned long long v = 0;
  CK(cudaMemcpyAsync(&v, s->n_active, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *n_active = (int64_t)v;
  return TOB200_OK;
}

int tob200_solver_results(tob200_solver *s, tob200_result *results) {
  if (!s || !results) return fail(s ? s->ctx : nullptr, TOB200_ERR_INVALID, "NULL argument");
  tob200_ctx *ctx = s->ctx;
  DeviceGuard guard(ctx->device);
  if (s->family == 4) {
    if (s->dtype == TOB200_F32) CK(launch_gn_results<float>((const LmScalars<float> *)s->rec, s->B, results, ctx->stream));
    else CK(launch_gn_results<double>((const LmScalars<double> *)s->rec, s->B, results, ctx->stream));
    ctx->launches++;
    return TOB200_OK;
  }
  const unsigned grid = (unsigned)((s->B + 255) / 256);
  if (s->dtype == TOB200_F32)
    results_kernel<float><<<grid, 256, 0, ctx->stream>>>((const StateRec<float> *)s->rec, s->B, results);
  else
    results_kernel<double><<<grid, 256, 0, ctx->stream>>>((const StateRec<double> *)s->rec, s->B, results);
  CK(cudaGetLastError());
  ctx->launches++;
  return TOB200_OK;
}

int tob200_solver_final_hessian(tob200_solver *s, double *H) {
  if (!s || !H) return fail(s ? s->ctx : nullptr, TOB200_ERR_INVALID, "NULL argument");
  tob200_ctx *ctx = s->ctx;
  DeviceGuard guard(ctx->device);
  if (s->family == 4) {
    if (s->dtype == TOB200_F32)
      CK((launch_gn_export_h<float, double>((const float *)s->H, (const float *)s->hd, (const LmScalars<float> *)s->rec, nullptr,
                                            s->opt.solver_type, s->B, s->n, H, ctx->stream)));
    else
      CK((launch_gn_export_h<double, double>((const double *)s->H, (const double *)s->hd, (const LmScalars<double> *)s->rec,
                                             nullptr, s->opt.solver_type, s->B, s->n, H, ctx->stream)));
    ctx->launches++;
    return TOB200_OK;
  }
  if (s->family == 2) {
    const int np = wpp_nb_for(s->n) * wpp_blk_for(s->n), ldw = wpp_ldw(np);
    if (s->dtype == TOB200_F32)
      wpp_final_hessian_kernel<float><<<(unsigned)s->B, 128, 0, ctx->stream>>>(
          (const float *)s->H, (const StateRec<float> *)s->rec, s->opt.solver_type, s->B, s->n, np, ldw, H);
    else
      wpp_final_hessian_kernel<double><<<(unsigned)s->B, 128, 0, ctx->stream>>>(
          (const double *)s->H, (const StateRec<double> *)s->rec, s->opt.solver_type, s->B, s->n, np, ldw, H);
    CK(cudaGetLastError());
    ctx->launches++;
    return TOB200_OK;
  }
  const unsigned grid = (unsigned)((s->B + 127) / 128);
  if (s->dtype == TOB200_F32)
    final_hessian_kernel<float><<<grid, 128, 0, ctx->stream>>>((const float *)s->H, (const StateRec<float> *)s->rec,
                                                             s->opt.solver_type, s->B, s->n, H);
  else
    final_hessian_kernel<double><<<grid, 128, 0, ctx->stream>>>((const double *)s->H,
                                                              (const StateRec<double> *)s->rec, s->opt.solver_type,
                                                              s->B, s->n, H);
  CK(cudaGetLastError());
  ctx->launches++;
  return TOB200_OK;
}

int tob200_solver_covariance(tob200_solver *s, double *cov, double *max_std, int32_t *status) {
  if (!s || !status || (!cov && !max_std)) return fail(s ? s->ctx : nullptr, TOB200_ERR_INVALID, "NULL argument");
  tob200_ctx *ctx = s->ctx;
  DeviceGuard guard(ctx->device);
  int rc = ensure_scratch(ctx, 20, (size_t)s->B * s->n * s->n * sizeof(double));
  if (rc != TOB200_OK) return rc;
  if ((rc = tob200_solver_final_hessian(s, (double *)ctx->scratch[20])) != TOB200_OK) return rc;
  return inv_cov_impl<double>(ctx, (const double *)ctx->scratch[20], s->B, s->n, cov, max_std, status);
}

// ---- synthetic family ----------------------------------------------------------------------------
#define SYNTH(SUF, T)                                                                                              \
  int tob200_synth_generate_##SUF(tob200_ctx *ctx, uint64_t seed, int64_t p0, int64_t B, int m, int n, T alpha,     \
                                  T sigma, int layout, T *A, T *y, T *xstar, T *x0) {                               \
    if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");                                              \
    if (B < 0 || m < 0 || n < 1) return fail(ctx, TOB200_ERR_INVALID, "need B >= 0, m >= 0, n >= 1");               \
    if (layout != TOB200_LAYOUT_TILE32 && layout != TOB200_LAYOUT_PROBLEM_MAJOR)                                    \
      return fail(ctx, TOB200_ERR_INVALID, "unknown layout");                                                       \
    DeviceGuard guard(ctx->device);                                                                                 \
    int launches = 0;                                                                                               \
    CK(launch_synth_generate<T>(seed, p0, B, m, n, alpha, sigma, layout, A, y, xstar, x0, ctx->stream, &launches)); \
    ctx->launches += launches;                                                                                      \
    return TOB200_OK;                                                                                               \
  }                                                                                                                 \
  int tob200_synth_eval_##SUF(tob200_ctx *ctx, const T *A, const T *y, T alpha, int layout, int64_t B, int m,       \
                              int n, const T *x, T *r, T *J) {                                                      \
    if (!ctx) return fail(nullptr, TOB200_ERR_INVALID, "ctx is NULL");                                              \
    if (B < 0 || m < 0 || n < 1) return fail(ctx, TOB200_ERR_INVALID, "need B >= 0, m >= 0, n >= 1");               \
    if (!A || !y || !x) return fail(ctx, TOB200_ERR_INVALID, "NULL buffer");                                        \
    if (layout != TOB200_LAYOUT_TILE32 && layout != TOB200_LAYOUT_PROBLEM_MAJOR)                                    \
      return fail(ctx, TOB200_ERR_INVALID, "unknown layout");                                                       \
    DeviceGuard guard(ctx->device);                                                                                 \
    CK(launch_synth_eval<T>(A, y, alpha, layout, B, m, n, x, r, J, ctx->stream));                                   \
    ctx->launches++;                                                                                                \
    return TOB200_OK;                                                                                               \
  }
SYNTH(f32, float)
SYNTH(f64, double)

}  // extern "C"
