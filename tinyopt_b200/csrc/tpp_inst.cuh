// tpp_inst.cuh — instantiation + launch plumbing for the thread-per-problem kernels.  Each
// tpp_inst_*.cu defines one `tpp_entry_*` function covering a range of n, so the heavy fully
// unrolled kernels compile in parallel.
#pragma once

#include "internal.h"
#include "tpp.cuh"

namespace tob200 {

enum TppKind { kTppRun = 0, kTppBuildSolve = 1, kTppStep = 2 };
enum TppOp { kTppLaunch = 0, kTppQuery = 1 };

struct TppLaunch {
  int grid, block;
  size_t smem;
  cudaStream_t stream;
};

// op == kTppLaunch: launch `kind` for this n with `params` (a Tpp*Params<T>*).
// op == kTppQuery : set the dynamic-smem attribute and return resident CTAs per SM in *out.
using TppEntry = cudaError_t (*)(int op, int n, int kind, const void *params, const TppLaunch &cfg, int *out);

template <typename T, int N>
cudaError_t tpp_entry_one(int op, int kind, const void *params, const TppLaunch &cfg, int *out) {
  const void *fn = nullptr;
  switch (kind) {
    case kTppRun: fn = (const void *)tpp_lm_run_kernel<T, N>; break;
    case kTppBuildSolve: fn = (const void *)tpp_build_solve_kernel<T, N>; break;
    case kTppStep: fn = (const void *)tpp_step_kernel<T, N>; break;
    default: return cudaErrorInvalidValue;
  }
  {  // per (device, kernel), monotonic, thread safe: internal.h
    cudaError_t e = raise_smem_limit(fn, cfg.smem);
    if (e != cudaSuccess) return e;
  }
  if (op == kTppQuery) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, fn, cfg.block, cfg.smem);
  void *args[] = {const_cast<void *>(params)};
  return cudaLaunchKernel(fn, dim3(cfg.grid), dim3(cfg.block), args, cfg.smem, cfg.stream);
}

template <typename T, int LO, int HI>
cudaError_t tpp_entry_range(int op, int n, int kind, const void *params, const TppLaunch &cfg, int *out) {
  if (n == LO) return tpp_entry_one<T, LO>(op, kind, params, cfg, out);
  if constexpr (LO < HI) return tpp_entry_range<T, LO + 1, HI>(op, n, kind, params, cfg, out);
  return cudaErrorInvalidValue;
}

#define TOB200_TPP_ENTRY_DECL(name) \
  cudaError_t name(int op, int n, int kind, const void *params, const TppLaunch &cfg, int *out)
#define TOB200_TPP_ENTRY_DEFINE(name, T, LO, HI)                 \
  TOB200_TPP_ENTRY_DECL(name) { return tpp_entry_range<T, LO, HI>(op, n, kind, params, cfg, out); }

// limits of the register-resident family
constexpr int kTppMaxN_f32 = 12;
constexpr int kTppMaxN_f64 = 8;

TOB200_TPP_ENTRY_DECL(tpp_entry_f32_a);  // n = 1..4
TOB200_TPP_ENTRY_DECL(tpp_entry_f32_b);  // n = 5..8
TOB200_TPP_ENTRY_DECL(tpp_entry_f32_c);  // n = 9..10
TOB200_TPP_ENTRY_DECL(tpp_entry_f32_d);  // n = 11..12
TOB200_TPP_ENTRY_DECL(tpp_entry_f64_a);  // n = 1..4
TOB200_TPP_ENTRY_DECL(tpp_entry_f64_b);  // n = 5..6
TOB200_TPP_ENTRY_DECL(tpp_entry_f64_c);  // n = 7..8

}  // namespace tob200
