// tc_probe.cu — stand-alone probe of tcgen05.mma.kind::tf32 shared-memory descriptor semantics.
// The host builds the shared-memory image of A (128 x 8) and B (N x 8) for a candidate layout, the
// kernel issues ONE MMA and dumps the 128 x N accumulator; the host reports which candidates match.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu && ./tc_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Probe {
  uint64_t a_desc_rel, b_desc_rel;  // descriptors with start address RELATIVE to the image base (added on device)
  uint32_t idesc;
  int N;
  int image_bytes;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const unsigned char *image, Probe pr, float *D) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *img = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < pr.image_bytes / 4; i += 128) ((uint32_t *)img)[i] = ((const uint32_t *)image)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t base = smem_u32(img);
    const uint64_t a = pr.a_desc_rel + (uint64_t)((base >> 4) & 0x3FFF);
    const uint64_t b = pr.b_desc_rel + (uint64_t)((base >> 4) & 0x3FFF);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
        "l"(a), "l"(b), "r"(pr.idesc), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait for the MMA
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(&bar)), "r"(0u)
          : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int cb = 0; cb < pr.N; cb += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)cb)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int q = 0; q < 8; ++q) D[(32 * warp + lane) * pr.N + cb + q] = __uint_as_float(v[q]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

static uint64_t make_desc(uint32_t rel_addr, uint32_t lbo, uint32_t sbo, int layout_type, int version) {
  uint64_t d = (uint64_t)((rel_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)version << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
static uint32_t make_idesc(int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
}

int main() {
  const int M = 128, N = 128, K = 8;
  std::vector<float> A(M * K), B(N * K), ref(M * N);
  for (int i = 0; i < M; ++i) for (int k = 0; k < K; ++k) A[i * K + k] = (float)((i * 7 + k * 3) % 11 - 5);
  for (int j = 0; j < N; ++j) for (int k = 0; k < K; ++k) B[j * K + k] = (float)((j * 5 + k * 2) % 13 - 6);
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
    float s = 0; for (int k = 0; k < K; ++k) s += A[i * K + k] * B[j * K + k]; ref[i * N + j] = s;
  }
  unsigned char *d_img; float *d_D;
  cudaMalloc(&d_img, 65536); cudaMalloc(&d_D, M * N * 4);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  struct Cand { const char *name; int mn_major; int swz; uint32_t lbo, sbo; int version; };
  // image layouts: element (r, k) of an operand with R rows (MN) -> byte offset
  auto off_mn_sw128 = [](int r, int k) { return (r / 32) * 1024 + k * 128 + ((((r % 32) / 4) ^ k) * 16) + (r % 4) * 4; };
  auto off_mn_none = [](int r, int k, int R) { (void)R; return (r / 4) * 128 + k * 16 + (r % 4) * 4; };          // core = 4 MN x 8 K, cores along MN
  auto off_k_none = [](int r, int k, int R) { return (k / 4) * (R / 8) * 128 + (r / 8) * 128 + (r % 8) * 16 + (k % 4) * 4; };  // K chunks outer
  auto off_k_sw32 = [](int r, int k) { return (r / 8) * 256 + (r % 8) * 32 + ((((k / 4)) ^ ((r % 8) >> 2 & 1)) * 16) + (k % 4) * 4; };
  std::vector<Cand> cands = {
      {"MN sw128 lbo=1024 sbo=4096 v1", 1, 2, 1024, 4096, 1}, {"MN sw128 lbo=4096 sbo=1024 v1", 1, 2, 4096, 1024, 1},
      {"MN sw128 lbo=1024 sbo=4096 v0", 1, 2, 1024, 4096, 0}, {"MN none  lbo=128 sbo=128 v1", 1, 0, 128, 128, 1},
      {"MN none  lbo=4096 sbo=128 v1", 1, 0, 4096, 128, 1},  {"MN none  lbo=128 sbo=4096 v1", 1, 0, 128, 4096, 1},
      {"K  none  lbo=2048 sbo=128 v1", 0, 0, 2048, 128, 1},  {"K  none  lbo=128 sbo=2048 v1", 0, 0, 128, 2048, 1},
      {"K  sw32  lbo=16 sbo=256 v1", 0, 6, 16, 256, 1},      {"K  none  lbo=2048 sbo=128 v0", 0, 0, 2048, 128, 0},
  };
  for (const Cand &c : cands) {
    std::vector<unsigned char> img(65536, 0);
    const int a_base = 0, b_base = 16384;
    for (int op = 0; op < 2; ++op) {
      const std::vector<float> &X = op ? B : A;
      const int R = op ? N : M, base = op ? b_base : a_base;
      for (int r = 0; r < R; ++r) for (int k = 0; k < K; ++k) {
        int o;
        if (c.mn_major) o = c.swz == 2 ? off_mn_sw128(r, k) : off_mn_none(r, k, R);
        else o = c.swz == 6 ? off_k_sw32(r, k) : off_k_none(r, k, R);
        memcpy(&img[base + o], &X[r * K + k], 4);
      }
    }
    cudaMemcpy(d_img, img.data(), 65536, cudaMemcpyHostToDevice);
    cudaMemset(d_D, 0xFF, M * N * 4);
    Probe pr;
    pr.a_desc_rel = make_desc(a_base, c.lbo, c.sbo, c.swz, c.version);
    pr.b_desc_rel = make_desc(b_base, c.lbo, c.sbo, c.swz, c.version);
    pr.idesc = make_idesc(N, c.mn_major, c.mn_major);
    pr.N = N; pr.image_bytes = 65536;
    probe_kernel<<<1, 128, 70000>>>(d_img, pr, d_D);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D(M * N);
    cudaMemcpy(D.data(), d_D, M * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0, nz = 0; for (int i = 0; i < M * N; ++i) { bad += D[i] != ref[i]; nz += D[i] != 0.f; }
    printf("%-34s err=%s mismatches=%d/%d nonzero=%d  D[0][0..3]=%g %g %g %g (ref %g %g %g %g) D[1][0]=%g (ref %g) D[40][70]=%g (ref %g)\n", c.name,
           cudaGetErrorString(e), bad, M * N, nz, D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3], D[N], ref[N], D[40 * N + 70], ref[40 * N + 70]);
  }
  return 0;
}
