// Stand-alone check of tcgen05.mma kind::tf32 with NO-swizzle K-major shared-memory
// operands (dev microbenchmark): D[128 x N] = A[128 x K] * B[N x K]^T, fp32 in TMEM.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 32;  // K floats = 4 MMA k-steps of 8

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  // layout_type [61,64) = 0: no swizzle (interleave)
  return d;
}

__global__ void __launch_bounds__(128) umma_test(const float* A, const float* B, float* D) {
  // layout [chunk = k/4][row][k%4] floats: 8 rows x 16 B core matrices contiguous
  __shared__ __align__(128) float sA[(K / 4) * M * 4];
  __shared__ __align__(128) float sB[(K / 4) * N * 4];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M * K; i += 128) {
    int r = i / K, k = i % K;
    sA[((k >> 2) * M + r) * 4 + (k & 3)] = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    int r = i / K, k = i % K;
    sB[((k >> 2) * N + r) * 4 + (k & 3)] = B[i];
  }
  if (tid == 0) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(&mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(&tmem_base);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // generic-proxy smem writes -> visible to the async proxy (tensor core)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sA);
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sB);
    for (int s = 0; s < K / 8; ++s) {
      // one MMA consumes K=8 floats = 2 chunks; chunk stride (LBO) = rows*16 B, 8-row group stride (SBO) = 128 B
      uint64_t da = make_desc(sa + s * 2 * M * 16, M * 16, 128);
      uint64_t db = make_desc(sb + s * 2 * N * 16, N * 16, 128);
      uint32_t acc = s > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(taddr),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb) : "memory");
  }
  {
    uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar);
    asm volatile(
        "{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra Dn;\n\tbra W;\n\tDn:\n\t}\n" ::"r"(mb)
        : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // each warp reads its 32 lanes (rows), 8 columns at a time
  const int row = warp * 32 + (tid & 31);
  for (int c = 0; c < N; c += 8) {
    uint32_t v[8];
    const uint32_t ta = taddr + ((uint32_t)(warp * 32) << 16) + c;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(ta));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[row * N + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(64));
}

int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N), R(M * N);
  srand(1);
  for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.f;
  for (auto& x : B) x = (rand() % 2001 - 1000) / 1000.f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
      R[m * N + n] = (float)s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, D.size() * 4);
  umma_test<<<1, 128>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs(D[i] - R[i])); maxref = fmax(maxref, fabs(R[i])); }
  printf("max abs err %.3e  (max ref %.3f)  rel %.3e  D[0..3]= %f %f %f %f  R= %f %f %f %f\n", maxerr, maxref, maxerr / maxref,
         D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
  printf("%s\n", maxerr / maxref < 5e-3 ? "UMMA_TF32 PASS" : "UMMA_TF32 FAIL");
  return 0;
}
