// wtc_kernels.cu — the mid-n tensor-core family (wtc.cuh): instantiation + launch.
#include "internal.h"
#include "wtc.cuh"

namespace tob200 {

cudaError_t launch_wtc_lm_run(const WtcParams &p, int grid, cudaStream_t st) {
  const void *fn = p.mode == 1 ? (const void *)wtc_lm_run_kernel<1> : (const void *)wtc_lm_run_kernel<0>;
  {
    cudaError_t e = raise_smem_limit(fn, p.L.total);
    if (e != cudaSuccess) return e;
  }
  if (p.mode == 1) wtc_lm_run_kernel<1><<<(unsigned)grid, kWtcThreads, p.L.total, st>>>(p);
  else wtc_lm_run_kernel<0><<<(unsigned)grid, kWtcThreads, p.L.total, st>>>(p);
  return cudaGetLastError();
}

}  // namespace tob200
