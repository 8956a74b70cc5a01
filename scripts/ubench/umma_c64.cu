// complex64 tile product on tcgen05 (dev check): C[128 x N] = A[128 x K] * B[K x N], complex,
// via the real embedding  C'[128 x 2N] = A'[128 x 2K] * B'[2N x 2K]^T  and the 3xTF32 split
//   D += A'lo*B'hi + A'hi*B'lo + A'hi*B'hi     (A'hi = raw fp32, hardware truncates to tf32)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <complex>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 64;        // complex sizes
constexpr int KT = 16;                        // complex k per stage -> 32 floats = 4 MMA k-steps
constexpr int NP = 2 * N;                     // B' rows / TMEM columns

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Ap: [stage][chunk 0..7][row 0..127][4 floats]; Bp: [stage][hi|lo][chunk 0..7][row 0..NP-1][4 floats]
__global__ void __launch_bounds__(128) k(const float* Ap, const float* Bp, float* D) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* sAh = (float*)smem;              // 8*128*4 floats = 16 KB
  float* sAl = sAh + 8 * M * 4;           // 16 KB
  float* sBh = sAl + 8 * M * 4;           // 8*NP*4 floats
  float* sBl = sBh + 8 * NP * 4;
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(&mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(&tmem_base);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(NP));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem_base;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  uint32_t phase = 0;
  for (int s = 0; s < K / KT; ++s) {
    for (int i = tid; i < 8 * M * 4; i += 128) {
      float v = Ap[(size_t)s * 8 * M * 4 + i];
      sAh[i] = v;
      sAl[i] = v - trunc_tf32(v);
    }
    for (int i = tid; i < 8 * NP * 4; i += 128) {
      sBh[i] = Bp[((size_t)s * 2 + 0) * 8 * NP * 4 + i];
      sBl[i] = Bp[((size_t)s * 2 + 1) * 8 * NP * 4 + i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      const uint32_t ah = (uint32_t)__cvta_generic_to_shared(sAh), al = (uint32_t)__cvta_generic_to_shared(sAl);
      const uint32_t bh = (uint32_t)__cvta_generic_to_shared(sBh), bl = (uint32_t)__cvta_generic_to_shared(sBl);
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a0 = pass == 0 ? al : ah, b0 = pass == 1 ? bl : bh;  // lo*hi, hi*lo, hi*hi
        for (int q = 0; q < 4; ++q) {
          uint64_t da = make_desc(a0 + q * 2 * M * 16, M * 16, 128);
          uint64_t db = make_desc(b0 + q * 2 * NP * 16, NP * 16, 128);
          uint32_t acc = (s | pass | q) != 0;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(taddr),
              "l"(da), "l"(db), "r"(idesc), "r"(acc)
              : "memory");
        }
      }
      uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb) : "memory");
    }
    {  // everyone waits for this stage's MMAs before overwriting the operand tiles
      uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar);
      asm volatile(
          "{\n\t.reg .pred p;\n\tW%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D%=;\n\tbra W%=;\n\tD%=:\n\t}\n" ::"r"(mb), "r"(phase)
          : "memory");
      phase ^= 1;
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = warp * 32 + (tid & 31);
  for (int c = 0; c < NP; c += 8) {
    uint32_t v[8];
    const uint32_t ta = taddr + ((uint32_t)(warp * 32) << 16) + c;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(ta));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[row * NP + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(NP));
}

static float trunc_tf32_h(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  typedef std::complex<double> cd;
  std::vector<std::complex<float>> A(M * K), B(K * N);
  srand(3);
  auto rnd = [] { return (rand() % 20001 - 10000) / 10000.f; };
  for (auto& x : A) x = {rnd(), rnd()};
  for (auto& x : B) x = {rnd(), rnd()};
  const int S = K / KT;
  std::vector<float> Ap((size_t)S * 8 * M * 4), Bp((size_t)S * 2 * 8 * NP * 4);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      int s = k / KT, kk = k % KT;
      for (int p = 0; p < 2; ++p) {
        int kp = 2 * kk + p;  // float index within the stage
        Ap[(size_t)s * 8 * M * 4 + ((kp >> 2) * M + m) * 4 + (kp & 3)] = p ? A[m * K + k].imag() : A[m * K + k].real();
      }
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      int s = k / KT, kk = k % KT;
      float br = B[k * N + n].real(), bi = B[k * N + n].imag();
      // B'[2n][2kk] = Br, B'[2n][2kk+1] = -Bi, B'[2n+1][2kk] = Bi, B'[2n+1][2kk+1] = Br
      float vals[2][2] = {{br, -bi}, {bi, br}};
      for (int q = 0; q < 2; ++q)
        for (int p = 0; p < 2; ++p) {
          int row = 2 * n + q, kp = 2 * kk + p;
          float v = vals[q][p], hi = trunc_tf32_h(v), lo = v - hi;
          size_t off = ((size_t)(kp >> 2) * NP + row) * 4 + (kp & 3);
          Bp[((size_t)s * 2 + 0) * 8 * NP * 4 + off] = v;   // hi: raw (hardware truncates)
          Bp[((size_t)s * 2 + 1) * 8 * NP * 4 + off] = lo;
        }
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, Ap.size() * 4); cudaMalloc(&dB, Bp.size() * 4); cudaMalloc(&dD, (size_t)M * NP * 4);
  cudaMemcpy(dA, Ap.data(), Ap.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bp.data(), Bp.size() * 4, cudaMemcpyHostToDevice);
  size_t smem = (size_t)(2 * 8 * M * 4 + 2 * 8 * NP * 4) * 4;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<1, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  std::vector<float> D((size_t)M * NP);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      cd s = 0;
      for (int k2 = 0; k2 < K; ++k2) s += cd(A[m * K + k2]) * cd(B[k2 * N + n]);
      cd g(D[(size_t)m * NP + 2 * n], D[(size_t)m * NP + 2 * n + 1]);
      maxerr = fmax(maxerr, std::abs(g - s));
      maxref = fmax(maxref, std::abs(s));
    }
  printf("complex 3xTF32 on tcgen05: max abs err %.3e, max ref %.3f, rel %.3e\n", maxerr, maxref, maxerr / maxref);
  printf("%s\n", maxerr / maxref < 1e-5 ? "UMMA_C64 PASS" : "UMMA_C64 FAIL");
  return 0;
}
