// lg_kernels.cu — launchers of the large-n family (lg.cuh, lg_solve.cuh) and its small helpers.
#include "../../include/tinyopt_b200.h"
#include "internal.h"
#include "lg_solve.cuh"

namespace tob200 {

// reset of the per-problem LM state (solvers/lm.h:46-52 reset(), output.h defaults)
__global__ void lg_init_kernel(LmScalars<float> *rec, DevOptions<float> opt, float *last_dx, int64_t B, int n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    LmScalars<float> s;
    s.reset_scalars(opt);
    rec[i] = s;
  }
  if (i < B * n) last_dx[i] = 0.f;
}

// H scratch [np][np] (upper triangle valid) -> dense n x n symmetric, diagonal damped by (1 + lambda)
__global__ void lg_export_h_kernel(const float *H, const float *dg, const float *lambda, int64_t B, int n, int np,
                                   float *out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * n * n) return;
  const int64_t pr = e / ((int64_t)n * n);
  const int ij = (int)(e % ((int64_t)n * n));
  const int i = ij / n, j = ij % n;
  const float *Hp = H + (size_t)pr * np * np;
  float v = i <= j ? Hp[(size_t)i * np + j] : Hp[(size_t)j * np + i];
  if (i == j && dg) v = dg[(size_t)pr * n + i];
  if (i == j && lambda) {
    const float lam = lambda[pr];
    if (lam > 0.f) v = (float)((double)v * (1.0 + (double)lam));  // solvers/lm.h:108-117
  }
  out[e] = v;
}

// dense n x n (upper triangle read) -> H scratch [np][np]
__global__ void lg_import_h_kernel(const float *in, int64_t B, int n, int np, float *H) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * n * n) return;
  const int64_t pr = e / ((int64_t)n * n);
  const int ij = (int)(e % ((int64_t)n * n));
  const int i = ij / n, j = ij % n;
  if (i <= j) H[(size_t)pr * np * np + (size_t)i * np + j] = in[e];
}

// Output::final_hessian for the large-n family (optimizer.h:313-316, lm.h:157-171): off-diagonals from the upper
// triangle of H_, the persistent damped diagonal hd divided by 1 + prev_lambda_ (in float), widened to double;
// n_out <= n leading rows / columns (n % 4 != 0 runs on a zero-padded copy)
__global__ void lg_final_hessian_kernel(const float *H, const float *hd, const LmScalars<float> *rec, int solver_type,
                                        int64_t B, int n_out, int np, double *out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * n_out * n_out) return;
  const int64_t pr = e / ((int64_t)n_out * n_out);
  const int ij = (int)(e % ((int64_t)n_out * n_out));
  const int i = ij / n_out, j = ij % n_out;
  const float *Hp = H + (size_t)pr * np * np;
  float v;
  if (i == j) {
    v = hd[(size_t)pr * np + i];
    const float pl = rec[pr].prev_lambda;
    if (solver_type == 0 && pl > 0.f) v = __fdiv_rn(v, __fadd_rn(1.f, pl));
  } else {
    v = i < j ? Hp[(size_t)i * np + j] : Hp[(size_t)j * np + i];
  }
  out[e] = (double)v;
}

cudaError_t launch_lg_final_hessian(const float *H, const float *hd, const LmScalars<float> *rec, int solver_type, int64_t B,
                                    int n_out, int np, double *out, cudaStream_t st) {
  const int64_t total = B * n_out * n_out;
  if (total <= 0) return cudaSuccess;
  lg_final_hessian_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(H, hd, rec, solver_type, B, n_out, np, out);
  return cudaGetLastError();
}

cudaError_t launch_lg_init(LmScalars<float> *rec, const DevOptions<float> &opt, float *last_dx, int64_t B, int n,
                           cudaStream_t st) {
  const int64_t total = B * n;
  lg_init_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(rec, opt, last_dx, B, n);
  return cudaGetLastError();
}

cudaError_t launch_lg_export_h(const float *H, const float *dg, const float *lambda, int64_t B, int n, int np, float *out,
                               cudaStream_t st) {
  const int64_t total = B * n * n;
  lg_export_h_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(H, dg, lambda, B, n, np, out);
  return cudaGetLastError();
}

cudaError_t launch_lg_import_h(const float *in, int64_t B, int n, int np, float *H, cudaStream_t st) {
  const int64_t total = B * n * n;
  lg_import_h_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, B, n, np, H);
  return cudaGetLastError();
}

cudaError_t launch_lg_eval(const LgEvalParams &p, int num_sms, cudaStream_t st) {
  const size_t smem = lg_eval_smem_bytes(p.n);
  {
    cudaError_t e = raise_smem_limit((const void *)lg_eval_kernel, smem);
    if (e != cudaSuccess) return e;
  }
  int64_t grid = 2 * (int64_t)num_sms;
  if (grid > p.B) grid = p.B;
  lg_eval_kernel<<<(unsigned)grid, kLgEvalThreads, smem, st>>>(p);
  return cudaGetLastError();
}

int lg_syrk_stages(int np, int raw_stages, int fp16) {
  const size_t stage = 2 * (size_t)lg_syrk_half_bytes(np, fp16);
  const size_t budget = 200 * 1024, raw = (size_t)raw_stages * lg_syrk_raw_bytes(np);
  int s = raw + 2 * stage <= budget ? (int)((budget - raw) / stage) : 2;
  if (s > kLgMaxStages) s = kLgMaxStages;
  if (s < 2) s = 2;
  return s;
}

cudaError_t launch_lg_syrk(const LgSyrkParams &p, int num_sms, cudaStream_t st) {
  const size_t smem = lg_syrk_smem_bytes(p.np, p.stages, p.raw_stages, p.fp16);
  {
    cudaError_t e = raise_smem_limit((const void *)lg_syrk_kernel, smem);
    if (e != cudaSuccess) return e;
  }
  int64_t grid = num_sms;  // one CTA per SM: each owns the SM's whole TMEM
  const int64_t total = (int64_t)((p.nstrips + 1) / 2) * p.B;  // work units: strip pairs
  if (grid > total) grid = total;
  if (p.mc) {  // clusters of two CTAs = the two units of a problem (total is even here)
    grid &= ~(int64_t)1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kLgSyrkThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, lg_syrk_kernel, p);
  }
  lg_syrk_kernel<<<(unsigned)grid, kLgSyrkThreads, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_lg_solve(const LgSolveParams &p, int grid, cudaStream_t st) {
  const size_t smem = (size_t)lg_solve_smem(p.np).total * 4;
  {
    cudaError_t e = raise_smem_limit((const void *)lg_solve_kernel, smem);
    if (e != cudaSuccess) return e;
  }
  lg_solve_kernel<<<(unsigned)grid, kLgSolveThreads, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace tob200
