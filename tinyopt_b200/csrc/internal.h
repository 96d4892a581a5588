// internal.h — host-side declarations shared by api.cu and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tinyopt_b200.h"

#include <map>
#include <mutex>
#include <utility>

namespace tob200 {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: the raised limit is
// remembered per (device, kernel) — a second context on another GPU of the same process must set it again
// — and raised monotonically (the same instantiation serves shapes with different footprints; lowering it
// for a small one would make a later launch of a larger, occupancy-cached one fail with "invalid
// argument").  Thread safe: contexts on different devices may be driven by different threads.
// Defined once in misc_kernels.cu.
cudaError_t raise_smem_limit(const void *fn, size_t smem);

// misc_kernels.cu
template <typename T>
cudaError_t launch_retile(const T *src, int64_t B, int m, int n, T *dst, cudaStream_t st);
template <typename T>
cudaError_t launch_untile(const T *src, int64_t B, int m, int n, T *dst, cudaStream_t st);
template <typename T>
cudaError_t launch_synth_generate(uint64_t seed, int64_t p0, int64_t B, int m, int n, T alpha, T sigma, int layout,
                                  T *A, T *y, T *xstar, T *x0, cudaStream_t st, int *launches);
template <typename T>
cudaError_t launch_synth_eval(const T *A, const T *y, T alpha, int layout, int64_t B, int m, int n, const T *x, T *r,
                              T *J, cudaStream_t st);

cudaError_t launch_repitch(const float *src, int64_t B, int rs, int ps, float *dst, int rd, int pd, cudaStream_t st);

cudaError_t launch_repitch_masked(const float *src, const int32_t *status, int64_t B, int ps, float *dst, int nd,
                                  cudaStream_t st);

// cov_kernels.cu: InvCov / MaxStdDev, one warp per problem (n <= 64)
template <typename T>
cudaError_t launch_cov_warp(const T *H, int64_t B, int n, T *cov, T *max_std, int32_t *status, int num_sms,
                            cudaStream_t st);

}  // namespace tob200

// lg_kernels.cu (large-n family; parameter structs in lg.cuh / lg_solve.cuh)
namespace tob200 {
template <typename T> struct LmScalars;
template <typename T> struct DevOptions;
struct LgEvalParams;
struct LgSyrkParams;
struct LgSolveParams;
cudaError_t launch_lg_init(LmScalars<float> *rec, const DevOptions<float> &opt, float *last_dx, int64_t B, int n,
                           cudaStream_t st);
cudaError_t launch_lg_export_h(const float *H, const float *dg, const float *lambda, int64_t B, int n, int np, float *out,
                               cudaStream_t st);
cudaError_t launch_lg_final_hessian(const float *H, const float *hd, const LmScalars<float> *rec, int solver_type, int64_t B,
                                    int n_out, int np, double *out, cudaStream_t st);
cudaError_t launch_lg_import_h(const float *in, int64_t B, int n, int np, float *H, cudaStream_t st);
cudaError_t launch_lg_eval(const LgEvalParams &p, int num_sms, cudaStream_t st);
cudaError_t launch_lg_syrk(const LgSyrkParams &p, int num_sms, cudaStream_t st);
cudaError_t launch_lg_solve(const LgSolveParams &p, int grid, cudaStream_t st);
int lg_syrk_stages(int np, int raw_stages, int fp16);
}  // namespace tob200

// wtc_kernels.cu (mid-n tensor-core family; parameter struct in wtc_params.h)
namespace tob200 {
struct WtcParams;
cudaError_t launch_wtc_lm_run(const WtcParams &p, int grid, cudaStream_t st);
}  // namespace tob200

// gn_kernels.cu (general family; parameter structs in gn.cuh)
namespace tob200 {
template <typename T> struct GnAccumParams;
template <typename T> struct GnSolveParams;
template <typename T>
cudaError_t launch_gn_init(LmScalars<T> *rec, const DevOptions<T> &opt, T *last_dx, int64_t B, int n, cudaStream_t st,
                           int32_t *needs = nullptr);
template <typename T>
cudaError_t launch_gn_import_hg(const T *grad, const T *Hin, const LmScalars<T> *rec, int is_lm, int64_t B, int n, T *g, T *H,
                                cudaStream_t st);
template <typename T>
cudaError_t launch_gn_results(const LmScalars<T> *rec, int64_t B, tob200_result *out, cudaStream_t st);
template <typename T>
cudaError_t launch_gn_accum(const GnAccumParams<T> &p, int num_sms, cudaStream_t st);
template <typename T>
cudaError_t launch_gn_solve(const GnSolveParams<T> &p, int grid, cudaStream_t st);
template <typename T, typename OutT>
cudaError_t launch_gn_export_h(const T *H, const T *hd, const LmScalars<T> *rec, const T *lambda, int solver_type, int64_t B,
                               int n, OutT *out, cudaStream_t st);
}  // namespace tob200
