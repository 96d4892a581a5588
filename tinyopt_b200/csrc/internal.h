// internal.h — host-side declarations shared by api.cu and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tob200 {

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

}  // namespace tob200
