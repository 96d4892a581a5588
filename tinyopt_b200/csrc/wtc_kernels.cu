// wtc_kernels.cu — the mid-n tensor-core family (wtc.cuh): instantiation + launch.
#include "internal.h"
#include "wtc.cuh"

namespace tob200 {

cudaError_t launch_wtc_lm_run(const WtcParams &p, int grid, cudaStream_t st) {
  {
    cudaError_t e = raise_smem_limit((const void *)wtc_lm_run_kernel, p.L.total);
    if (e != cudaSuccess) return e;
  }
  wtc_lm_run_kernel<<<(unsigned)grid, kWtcThreads, p.L.total, st>>>(p);
  return cudaGetLastError();
}

}  // namespace tob200
