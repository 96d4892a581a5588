// wpp_inst.cuh — instantiation + launch plumbing for the warp-per-problem kernels (wpp.cuh).
#pragma once

#include "tpp_inst.cuh"  // TppLaunch, TppOp
#include "wpp.cuh"
#include "wpp_step.cuh"

namespace tob200 {

// the *Inv kinds are the `hessian.use_ldlt = false` variants (wpp_lu_solve); they are instantiated in
// their own translation units (wpp_inst_*_inv.cu) behind the *_inv entries
enum WppKind { kWppRun = 0, kWppBuildSolve = 1, kWppStep = 2, kWppRunInv = 3, kWppStepInv = 4 };

template <typename T, int NB, int BLK>
cudaError_t wpp_entry_one(int op, int kind, const void *params, const TppLaunch &cfg, int *out) {
  const void *fn = nullptr;
  switch (kind) {
#ifdef TOB200_WPP_INV_TU
    case kWppRunInv: fn = (const void *)wpp_lm_run_kernel<T, NB, BLK, true>; break;
    case kWppStepInv: fn = (const void *)wpp_step_kernel<T, NB, BLK, true>; break;
#else
    case kWppRun: fn = (const void *)wpp_lm_run_kernel<T, NB, BLK>; break;
    case kWppBuildSolve: fn = (const void *)wpp_build_solve_kernel<T, NB, BLK>; break;
    case kWppStep: fn = (const void *)wpp_step_kernel<T, NB, BLK>; break;
#endif
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

// one entry per block edge; nb = 4..7
#define TOB200_WPP_ENTRY_DECL(name) \
  cudaError_t name(int op, int nb, int kind, const void *params, const TppLaunch &cfg, int *out)
#define TOB200_WPP_ENTRY_DEFINE(name, T, BLK)                                              \
  TOB200_WPP_ENTRY_DECL(name) {                                                            \
    switch (nb) {                                                                          \
      case 4: return wpp_entry_one<T, 4, BLK>(op, kind, params, cfg, out);                 \
      case 5: return wpp_entry_one<T, 5, BLK>(op, kind, params, cfg, out);                 \
      case 6: return wpp_entry_one<T, 6, BLK>(op, kind, params, cfg, out);                 \
      case 7: return wpp_entry_one<T, 7, BLK>(op, kind, params, cfg, out);                 \
      default: return cudaErrorInvalidValue;                                               \
    }                                                                                      \
  }

// the double instantiation also serves n = 9..12 (n + 1 <= 12: three blocks of 4)
#define TOB200_WPP_ENTRY_DEFINE3(name, T, BLK)                                             \
  TOB200_WPP_ENTRY_DECL(name) {                                                            \
    switch (nb) {                                                                          \
      case 3: return wpp_entry_one<T, 3, BLK>(op, kind, params, cfg, out);                 \
      case 4: return wpp_entry_one<T, 4, BLK>(op, kind, params, cfg, out);                 \
      case 5: return wpp_entry_one<T, 5, BLK>(op, kind, params, cfg, out);                 \
      case 6: return wpp_entry_one<T, 6, BLK>(op, kind, params, cfg, out);                 \
      case 7: return wpp_entry_one<T, 7, BLK>(op, kind, params, cfg, out);                 \
      default: return cudaErrorInvalidValue;                                               \
    }                                                                                      \
  }

constexpr int kWppMinN_f32 = kTppMaxN_f32 + 1;  // 13
constexpr int kWppMaxN_f32 = 55;                // n + 1 <= 7 * 8
constexpr int kWppMinN_f64 = kTppMaxN_f64 + 1;  // 9
constexpr int kWppMaxN_f64 = 55;

TOB200_WPP_ENTRY_DECL(wpp_entry_f32_blk4);  // n = 13..27
TOB200_WPP_ENTRY_DECL(wpp_entry_f32_blk8);  // n = 28..55
TOB200_WPP_ENTRY_DECL(wpp_entry_f64_blk4);  // n = 9..27, scalar paths (functional coverage of double)
TOB200_WPP_ENTRY_DECL(wpp_entry_f64_blk8);  // n = 28..55
TOB200_WPP_ENTRY_DECL(wpp_entry_f32_blk4_inv);  // kWppRunInv / kWppStepInv of the same ranges
TOB200_WPP_ENTRY_DECL(wpp_entry_f32_blk8_inv);
TOB200_WPP_ENTRY_DECL(wpp_entry_f64_blk4_inv);
TOB200_WPP_ENTRY_DECL(wpp_entry_f64_blk8_inv);

}  // namespace tob200
