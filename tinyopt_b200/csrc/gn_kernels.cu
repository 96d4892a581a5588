// gn_kernels.cu — launchers of the general kernel family (gn.cuh): double precision above n = 55, any
// precision above n = 512, and `hessian.use_ldlt = false` above n = 55.
#include "../../include/tinyopt_b200.h"
#include "gn.cuh"
#include "internal.h"

namespace tob200 {

template <typename T>
cudaError_t launch_gn_init(LmScalars<T> *rec, const DevOptions<T> &opt, T *last_dx, int64_t B, int n, cudaStream_t st,
                           int32_t *needs) {
  const int64_t total = B * n;
  gn_init_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(rec, opt, last_dx, B, n, needs);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_gn_import_hg(const T *grad, const T *Hin, const LmScalars<T> *rec, int is_lm, int64_t B, int n, T *g, T *H,
                                cudaStream_t st) {
  const int64_t total = B * n * n;
  if (total <= 0) return cudaSuccess;
  gn_import_hg_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(grad, Hin, rec, is_lm, B, n, g, H);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_gn_results(const LmScalars<T> *rec, int64_t B, tob200_result *out, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  gn_results_kernel<T><<<(unsigned)((B + 255) / 256), 256, 0, st>>>(rec, B, out);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_gn_accum(const GnAccumParams<T> &p, int num_sms, cudaStream_t st) {
  const size_t smem = gn_accum_smem_bytes(p.n, sizeof(T));
  cudaError_t e = raise_smem_limit((const void *)gn_accum_kernel<T>, smem);
  if (e != cudaSuccess) return e;
  int64_t grid = 2 * (int64_t)num_sms;
  if (grid > p.B) grid = p.B;
  gn_accum_kernel<T><<<(unsigned)grid, kGnThreads, smem, st>>>(p);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_gn_solve(const GnSolveParams<T> &p, int grid, cudaStream_t st) {
  const size_t smem = gn_solve_smem_bytes(p.n, sizeof(T));
  cudaError_t e = raise_smem_limit((const void *)gn_solve_kernel<T>, smem);
  if (e != cudaSuccess) return e;
  gn_solve_kernel<T><<<(unsigned)grid, kGnThreads, smem, st>>>(p);
  return cudaGetLastError();
}

template <typename T, typename OutT>
cudaError_t launch_gn_export_h(const T *H, const T *hd, const LmScalars<T> *rec, const T *lambda, int solver_type, int64_t B,
                               int n, OutT *out, cudaStream_t st) {
  const int64_t total = B * n * n;
  if (total <= 0) return cudaSuccess;
  gn_export_h_kernel<T, OutT><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(H, hd, rec, lambda, solver_type, B, n, out);
  return cudaGetLastError();
}

#define TOB200_GN_INST(T)                                                                                                   \
  template cudaError_t launch_gn_init<T>(LmScalars<T> *, const DevOptions<T> &, T *, int64_t, int, cudaStream_t, int32_t *); \
  template cudaError_t launch_gn_import_hg<T>(const T *, const T *, const LmScalars<T> *, int, int64_t, int, T *, T *,       \
                                              cudaStream_t);                                                                \
  template cudaError_t launch_gn_results<T>(const LmScalars<T> *, int64_t, tob200_result *, cudaStream_t);                   \
  template cudaError_t launch_gn_accum<T>(const GnAccumParams<T> &, int, cudaStream_t);                                      \
  template cudaError_t launch_gn_solve<T>(const GnSolveParams<T> &, int, cudaStream_t);                                      \
  template cudaError_t launch_gn_export_h<T, double>(const T *, const T *, const LmScalars<T> *, const T *, int, int64_t,    \
                                                     int, double *, cudaStream_t);
TOB200_GN_INST(float)
TOB200_GN_INST(double)
template cudaError_t launch_gn_export_h<float, float>(const float *, const float *, const LmScalars<float> *, const float *, int,
                                                      int64_t, int, float *, cudaStream_t);

}  // namespace tob200
