// gett_kernels.cuh -- permutation-fused pairwise tensor contraction for sm_100a.
//
// Replaces cotengra/contract.py:364-411 (`_do_contraction_via_bmm`: transpose ->
// reshape(copy) -> matmul -> reshape/transpose) by ONE kernel:
//   * operands are read where they lie: every tile element is fetched with a
//     cp.async (LDGSTS) from  base(tile) + kbase(k-step) + delta(element),
//     the three terms being sums of digit*stride over the GRID, K-GRID and TILE
//     dims of the descriptor (gett_desc.h);  no permuted copy of A, B or C is
//     ever materialised in HBM;
//   * a 3-4 stage shared-memory ring hides HBM/L2 latency;
//   * the compute policy is pluggable: FMA register tiles (any dtype), per-thread
//     k partial sums (dot-product-like nodes), or fp64 tensor-core mma.sync
//     (DMMA m8n8k4) for float64 / complex128;
//   * results are stored straight into the parent's index order (strided C),
//     optionally accumulated (slice sums, core.py:3842-3844) or atomically
//     added (split-K).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gett_desc.h"

namespace ctgb {

// ------------------------------------------------------------------ elements
__device__ __forceinline__ void mac(float& c, float a, float b) { c = fmaf(a, b, c); }
__device__ __forceinline__ void mac(double& c, double a, double b) { c = fma(a, b, c); }
__device__ __forceinline__ void mac(float2& c, float2 a, float2 b) {
  c.x = fmaf(a.x, b.x, c.x);
  c.x = fmaf(-a.y, b.y, c.x);
  c.y = fmaf(a.x, b.y, c.y);
  c.y = fmaf(a.y, b.x, c.y);
}
__device__ __forceinline__ void mac(double2& c, double2 a, double2 b) {
  c.x = fma(a.x, b.x, c.x);
  c.x = fma(-a.y, b.y, c.x);
  c.y = fma(a.x, b.y, c.y);
  c.y = fma(a.y, b.x, c.y);
}
template <typename T> __device__ __forceinline__ T zero_of();
template <> __device__ __forceinline__ float zero_of<float>() { return 0.f; }
template <> __device__ __forceinline__ double zero_of<double>() { return 0.0; }
template <> __device__ __forceinline__ float2 zero_of<float2>() { return make_float2(0.f, 0.f); }
template <> __device__ __forceinline__ double2 zero_of<double2>() { return make_double2(0.0, 0.0); }

__device__ __forceinline__ float add_of(float a, float b) { return a + b; }
__device__ __forceinline__ double add_of(double a, double b) { return a + b; }
__device__ __forceinline__ float2 add_of(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 add_of(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

__device__ __forceinline__ void atomic_add_of(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_of(double* p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_of(float2* p, float2 v) {
  atomicAdd(&p->x, v.x);
  atomicAdd(&p->y, v.y);
}
__device__ __forceinline__ void atomic_add_of(double2* p, double2 v) {
  atomicAdd(&p->x, v.x);
  atomicAdd(&p->y, v.y);
}

// two adjacent elements with one store; 256-bit (STG.E.ENL2.256) for complex128
template <typename T>
__device__ __forceinline__ void store_pair_of(T* p, T v0, T v1) {
  p[0] = v0;
  p[1] = v1;
}
template <>
__device__ __forceinline__ void store_pair_of<double2>(double2* p, double2 v0, double2 v1) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v0.x), "d"(v0.y), "d"(v1.x), "d"(v1.y)
               : "memory");
}

__device__ __forceinline__ float shfl_down_of(float v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_down_of(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ float2 shfl_down_of(float2 v, int d) {
  return make_float2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}
__device__ __forceinline__ double2 shfl_down_of(double2 v, int d) {
  return make_double2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}

// ------------------------------------------------------------------ cp.async
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(void* smem_dst, const void* gsrc, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  int src_bytes = valid ? BYTES : 0;
  if constexpr (BYTES == 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes) : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(s), "l"(gsrc), "n"(BYTES), "r"(src_bytes)
                 : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// ------------------------------------------------------------------ policies
// A policy fixes the CTA tile (MT x NT x KT), the pipeline depth, the shared
// memory layout of the operand tiles and how the tile product is computed.

// Generic FMA register-tile policy: works for every dtype and every extent.
template <typename T, int MT_, int NT_, int KT_, int STAGES_>
struct SimtPolicy {
  static constexpr int MT = MT_, NT = NT_, KT = KT_, STAGES = STAGES_;
  static constexpr int THREADS = 256;
  static constexpr int TM = MT / 16, TN = NT / 16;
  static constexpr int A_ELEMS = MT * KT, B_ELEMS = NT * KT;
  static constexpr int SCRATCH_ELEMS = 0;
  static constexpr int MIN_BLOCKS = 2;
  struct Acc {
    T v[TM][TN];
  };
  __device__ static __forceinline__ int idxA(int r, int kk) { return kk * MT + r; }
  __device__ static __forceinline__ int idxB(int c, int kk) { return kk * NT + c; }
  __device__ static __forceinline__ void clear(Acc& acc) {
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc.v[i][j] = zero_of<T>();
  }
  __device__ static __forceinline__ void compute(const T* __restrict__ sA, const T* __restrict__ sB, Acc& acc,
                                                 int kvalid) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      T a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = sA[kk * MT + ty + 16 * i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = sB[kk * NT + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) mac(acc.v[i][j], a[i], b[j]);
    }
  }
  template <typename F, typename F2>
  __device__ static __forceinline__ void epilogue(Acc& acc, T* scratch, F&& store, F2&& store_pair, bool pair_ok) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) store(ty + 16 * i, tx + 16 * j, acc.v[i][j]);
  }
};

// Tiny M x N with a huge contracted space (the final inner product of an
// amplitude network is M = N = 1, K = 2^30): every thread owns a strided set
// of k and all (r, c) pairs; partial sums are reduced across the block once.
template <typename T, int MT_, int NT_, int KT_, int STAGES_>
struct KredPolicy {
  static constexpr int MT = MT_, NT = NT_, KT = KT_, STAGES = STAGES_;
  static constexpr int THREADS = 256;
  static constexpr int A_ELEMS = MT * KT, B_ELEMS = NT * KT;
  static constexpr int SCRATCH_ELEMS = MT * NT * (THREADS / 32);
  static constexpr int MIN_BLOCKS = 1;
  struct Acc {
    T v[MT][NT];
  };
  __device__ static __forceinline__ int idxA(int r, int kk) { return r * KT + kk; }
  __device__ static __forceinline__ int idxB(int c, int kk) { return c * KT + kk; }
  __device__ static __forceinline__ void clear(Acc& acc) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) acc.v[i][j] = zero_of<T>();
  }
  __device__ static __forceinline__ void compute(const T* __restrict__ sA, const T* __restrict__ sB, Acc& acc,
                                                 int kvalid) {
#pragma unroll
    for (int kk = threadIdx.x; kk < KT; kk += THREADS) {
      T a[MT], b[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) a[i] = sA[i * KT + kk];
#pragma unroll
      for (int j = 0; j < NT; ++j) b[j] = sB[j * KT + kk];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) mac(acc.v[i][j], a[i], b[j]);
    }
  }
  template <typename F, typename F2>
  __device__ static __forceinline__ void epilogue(Acc& acc, T* scratch, F&& store, F2&& store_pair, bool pair_ok) {
    // block reduction of every (r, c) partial sum: shuffles, then 8 warps via smem
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        T v = acc.v[i][j];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = add_of(v, shfl_down_of(v, d));
        if (lane == 0) scratch[(i * NT + j) * (THREADS / 32) + warp] = v;
      }
    __syncthreads();
    if (threadIdx.x < MT * NT) {
      T v = zero_of<T>();
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) v = add_of(v, scratch[threadIdx.x * (THREADS / 32) + w]);
      store(threadIdx.x / NT, threadIdx.x % NT, v);
    }
    __syncthreads();
  }
};

// Skinny nodes (few kept indices on the small operand: N <= 8) are HBM-bound:
// one output row per thread, the small operand broadcast from shared memory,
// small footprint so that many CTAs per SM keep enough loads in flight.
template <typename T, int MT_, int NT_, int KT_, int STAGES_>
struct RowPolicy {
  static constexpr int MT = MT_, NT = NT_, KT = KT_, STAGES = STAGES_;
  static constexpr int THREADS = MT;
  static constexpr int A_ELEMS = MT * KT, B_ELEMS = NT * KT;
  static constexpr int SCRATCH_ELEMS = 0;
  static constexpr int MIN_BLOCKS = 4;
  struct Acc {
    T v[NT];
  };
  __device__ static __forceinline__ int idxA(int r, int kk) { return kk * MT + r; }
  __device__ static __forceinline__ int idxB(int c, int kk) { return kk * NT + c; }
  __device__ static __forceinline__ void clear(Acc& acc) {
#pragma unroll
    for (int j = 0; j < NT; ++j) acc.v[j] = zero_of<T>();
  }
  __device__ static __forceinline__ void compute(const T* __restrict__ sA, const T* __restrict__ sB, Acc& acc,
                                                 int kvalid) {
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      if (kk >= kvalid) break;
      const T a = sA[kk * MT + threadIdx.x];
#pragma unroll
      for (int j = 0; j < NT; ++j) mac(acc.v[j], a, sB[kk * NT + j]);
    }
  }
  template <typename F, typename F2>
  __device__ static __forceinline__ void epilogue(Acc& acc, T* scratch, F&& store, F2&& store_pair, bool pair_ok) {
    if (pair_ok) {
      // the row's columns are adjacent in C: 32-byte (256-bit) stores, full sectors
#pragma unroll
      for (int j = 0; j < NT; j += 2) store_pair((int)threadIdx.x, j, acc.v[j], acc.v[j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < NT; ++j) store((int)threadIdx.x, j, acc.v[j]);
    }
  }
};

// fp64 tensor-core policy: mma.sync.aligned.m8n8k4 (DMMA).  tcgen05 has no f64
// kind (cute/arch/mma_sm100_umma.hpp exposes f16/tf32/f8f6f4/i8/mx* only), so
// the double-precision tensor path on sm_100a is the warp-level DMMA.
// Complex products are four real DMMAs per (A-frag, B-frag) pair:
//   Cr += Ar*Br;  Cr += (-Ai)*Bi;  Ci += Ar*Bi;  Ci += Ai*Br.
__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <typename T, int WARPS_M, int WARPS_N, int FM, int FN, int KT_, int STAGES_>
struct DmmaPolicy {
  // T is double (real) or double2 (complex)
  static constexpr bool CPLX = sizeof(T) == 16;
  static constexpr int MT = WARPS_M * FM * 8, NT = WARPS_N * FN * 8, KT = KT_, STAGES = STAGES_;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int A_ELEMS = MT * KT, B_ELEMS = NT * KT;
  static constexpr int SCRATCH_ELEMS = 0;
  static constexpr int MIN_BLOCKS = 1;
  static_assert(KT % 4 == 0, "KT must be a multiple of the DMMA k");
  struct Acc {
    double re[FM][FN][2];
    double im[CPLX ? FM : 1][CPLX ? FN : 1][2];
  };
  // [k/4][row][k%4]: the 4 k of one fragment row are contiguous (64 B complex),
  // fragment rows contiguous -> conflict-free LDS.128 / LDS.64 fragment loads.
  __device__ static __forceinline__ int idxA(int r, int kk) { return ((kk >> 2) * MT + r) * 4 + (kk & 3); }
  __device__ static __forceinline__ int idxB(int c, int kk) { return ((kk >> 2) * NT + c) * 4 + (kk & 3); }
  __device__ static __forceinline__ void clear(Acc& acc) {
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
      for (int j = 0; j < FN; ++j) {
        acc.re[i][j][0] = acc.re[i][j][1] = 0.0;
        if constexpr (CPLX) acc.im[i][j][0] = acc.im[i][j][1] = 0.0;
      }
  }
  __device__ static __forceinline__ void compute(const T* __restrict__ sA, const T* __restrict__ sB, Acc& acc,
                                                 int kvalid) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int frow = lane >> 2, fk = lane & 3;
    const T* pa = sA + ((wm * FM * 8 + frow) * 4 + fk);
    const T* pb = sB + ((wn * FN * 8 + frow) * 4 + fk);
#pragma unroll
    for (int k4 = 0; k4 < KT / 4; ++k4) {
      if (k4 * 4 >= kvalid) break;  // uniform: trailing k of a ragged step are zero
      T a[FM], b[FN];
#pragma unroll
      for (int i = 0; i < FM; ++i) a[i] = pa[(k4 * MT + i * 8) * 4];
#pragma unroll
      for (int j = 0; j < FN; ++j) b[j] = pb[(k4 * NT + j * 8) * 4];
      if constexpr (CPLX) {
        // four passes of FM*FN independent DMMAs: the two updates of one
        // accumulator are FM*FN*2 instructions apart, so the tensor pipe never
        // waits on its own result
        double nai[FM];
#pragma unroll
        for (int i = 0; i < FM; ++i) nai[i] = -a[i].y;
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma8x8x4(acc.re[i][j][0], acc.re[i][j][1], a[i].x, b[j].x);
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma8x8x4(acc.im[i][j][0], acc.im[i][j][1], a[i].x, b[j].y);
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma8x8x4(acc.re[i][j][0], acc.re[i][j][1], nai[i], b[j].y);
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma8x8x4(acc.im[i][j][0], acc.im[i][j][1], a[i].y, b[j].x);
      } else {
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma8x8x4(acc.re[i][j][0], acc.re[i][j][1], a[i], b[j]);
      }
    }
  }
  template <typename F, typename F2>
  __device__ static __forceinline__ void epilogue(Acc& acc, T* scratch, F&& store, F2&& store_pair, bool pair_ok) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int frow = lane >> 2, fc = (lane & 3) * 2;
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
      for (int j = 0; j < FN; ++j) {
        const int r = (wm * FM + i) * 8 + frow;
        const int c = (wn * FN + j) * 8 + fc;
        if constexpr (CPLX) {
          // a lane owns two adjacent columns of the fragment: one 256-bit store
          if (pair_ok) {
            store_pair(r, c, make_double2(acc.re[i][j][0], acc.im[i][j][0]),
                       make_double2(acc.re[i][j][1], acc.im[i][j][1]));
          } else {
            store(r, c, make_double2(acc.re[i][j][0], acc.im[i][j][0]));
            store(r, c + 1, make_double2(acc.re[i][j][1], acc.im[i][j][1]));
          }
        } else {
          store(r, c, acc.re[i][j][0]);
          store(r, c + 1, acc.re[i][j][1]);
        }
      }
  }
};

// ------------------------------------------------------------------ skeleton
constexpr int KCHUNK = 128;  // k-steps whose base offsets are tabulated once per kernel

template <class P>
struct GettSmem {
  static constexpr int NA = (P::A_ELEMS + P::THREADS - 1) / P::THREADS;
  static constexpr int NB = (P::B_ELEMS + P::THREADS - 1) / P::THREADS;
  static constexpr int TI = P::STAGES + 1;  // tile-info ring
  template <typename T>
  static constexpr size_t bytes() {
    return sizeof(T) * ((size_t)P::STAGES * (P::A_ELEMS + P::B_ELEMS) + P::SCRATCH_ELEMS)  // ring + scratch
           + 8 * (size_t)(NA + NB) * P::THREADS                                            // element deltas
           + 8 * (size_t)(P::MT + P::NT)                                                   // C offsets
           + 8 * (size_t)2 * KCHUNK                                                        // k-step bases
           + 8 * (size_t)3 * TI                                                            // tile bases
           + 4 * (size_t)(NA + NB) * P::THREADS                                            // element (r, kk)
           + 4 * (size_t)KCHUNK + 4 * (size_t)2 * TI + 4 * (size_t)P::STAGES               // valid counts
           + 64;
  }
};

// One CTA walks its work items (tile x k-split) as ONE stream of k-steps: the
// cp.async ring keeps prefetching across tile boundaries, so tiles with few
// k-steps (K <= 64 on Sycamore trees) and single-step HBM-bound tiles never
// drain the pipeline.
template <typename T, class P>
__global__ void __launch_bounds__(P::THREADS, P::MIN_BLOCKS)
gett_kernel(const int64_t* __restrict__ D, const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C) {
  constexpr int MT = P::MT, NT = P::NT, STAGES = P::STAGES, THREADS = P::THREADS;
  constexpr int NA = GettSmem<P>::NA, NB = GettSmem<P>::NB, TI = GettSmem<P>::TI;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sA = reinterpret_cast<T*>(smem_raw);
  T* sB = sA + STAGES * P::A_ELEMS;
  T* scratch = sB + STAGES * P::B_ELEMS;
  long long* gA = reinterpret_cast<long long*>(scratch + P::SCRATCH_ELEMS);
  long long* gB = gA + NA * THREADS;
  long long* offMC = gB + NB * THREADS;
  long long* offNC = offMC + MT;
  long long* kbA = offNC + NT;
  long long* kbB = kbA + KCHUNK;
  long long* ti_base = kbB + KCHUNK;  // [TI][3]: A, B, C
  unsigned* metaA = reinterpret_cast<unsigned*>(ti_base + 3 * TI);
  unsigned* metaB = metaA + NA * THREADS;
  int* kval = reinterpret_cast<int*>(metaB + NB * THREADS);
  int* ti_valid = kval + KCHUNK;       // [TI][2]: m_valid, n_valid
  int* stage_kv = ti_valid + 2 * TI;   // [STAGES]

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  // ---- header (uniform loads through the read-only path) ----
  const int n_tm = (int)D[W_NTM], n_tn = (int)D[W_NTN];
  const int n_gm = (int)D[W_NGM], n_gn = (int)D[W_NGN], n_gk = (int)D[W_NGK], n_gb = (int)D[W_NGB];
  const int MTa = (int)D[W_MTA], NTa = (int)D[W_NTA], KTa = (int)D[W_KTA];
  const unsigned tiles_m = (unsigned)D[W_TILES_M], tiles_n = (unsigned)D[W_TILES_N], tiles_b = (unsigned)D[W_TILES_B];
  const unsigned steps_k = (unsigned)D[W_STEPS_K], splitk = (unsigned)D[W_SPLITK];
  const int pgm = (int)D[W_PGM], pgn = (int)D[W_PGN], pgk = (int)D[W_PGK];
  const bool accumulate = (D[W_FLAGS] & 1) != 0;
  const bool atomic = splitk > 1;
  // bit1: every pair of columns (2q, 2q+1) is adjacent in C and 32-byte aligned
  const bool pair_ok = (D[W_FLAGS] & 2) != 0 && !atomic && !accumulate && sizeof(T) == 16;
  const bool ktab = steps_k <= (unsigned)KCHUNK;

  // ---- one-time tables ----
  // zero the operand ring: rows/cols/k beyond the actual tile are never loaded
  for (int i = tid; i < STAGES * (P::A_ELEMS + P::B_ELEMS); i += THREADS) sA[i] = zero_of<T>();
  // per-slot element tables, enumerated in operand-memory order for coalescing
  {
    const int n_lda = (int)D[W_NLDA], n_ldb = (int)D[W_NLDB];
    for (int i = 0; i < NA; ++i) {
      unsigned e = tid + i * THREADS;
      long long g = 0;
      unsigned r = 0, kk = 0xFFFFu;
      if (e < (unsigned)(MTa * KTa)) {
        kk = 0;
        for (int d = 0; d < n_lda; ++d) {
          const int64_t* L = D + OFF_LDA + d * 4;
          unsigned ext = (unsigned)L[0];
          unsigned dig = e % ext;
          e /= ext;
          g += (long long)dig * L[1];
          r += dig * (unsigned)L[2];
          kk += dig * (unsigned)L[3];
        }
      }
      gA[i * THREADS + tid] = g;
      metaA[i * THREADS + tid] = r | (kk << 16);
    }
    for (int i = 0; i < NB; ++i) {
      unsigned e = tid + i * THREADS;
      long long g = 0;
      unsigned c = 0, kk = 0xFFFFu;
      if (e < (unsigned)(NTa * KTa)) {
        kk = 0;
        for (int d = 0; d < n_ldb; ++d) {
          const int64_t* L = D + OFF_LDB + d * 4;
          unsigned ext = (unsigned)L[0];
          unsigned dig = e % ext;
          e /= ext;
          g += (long long)dig * L[1];
          kk += dig * (unsigned)L[2];
          c += dig * (unsigned)L[3];
        }
      }
      gB[i * THREADS + tid] = g;
      metaB[i * THREADS + tid] = c | (kk << 16);
    }
  }
  // local C offsets of every tile row / column
  for (int r = tid; r < MT; r += THREADS) {
    long long o = 0;
    if (r < MTa) {
      unsigned e = r;
      for (int d = 0; d < n_tm; ++d) {
        const int64_t* L = D + OFF_TM + d * 3;
        unsigned ext = (unsigned)L[0];
        o += (long long)(e % ext) * L[2];
        e /= ext;
      }
    }
    offMC[r] = o;
  }
  for (int c = tid; c < NT; c += THREADS) {
    long long o = 0;
    if (c < NTa) {
      unsigned e = c;
      for (int d = 0; d < n_tn; ++d) {
        const int64_t* L = D + OFF_TN + d * 3;
        unsigned ext = (unsigned)L[0];
        o += (long long)(e % ext) * L[2];
        e /= ext;
      }
    }
    offNC[c] = o;
  }
  // k-step bases: a function of the absolute step index only
  auto kstep_bases = [&](unsigned step, long long& a, long long& b, int& kv) {
    a = 0;
    b = 0;
    kv = KTa;
    for (int j = 0; j < n_gk; ++j) {
      const int64_t* G = D + OFF_GK + j * 4;
      unsigned dig = (step / (unsigned)G[1]) % (unsigned)G[0];
      a += (long long)dig * G[2];
      b += (long long)dig * G[3];
      if (j == pgk)
        kv = (int)min((long long)D[W_KTEXT], (long long)D[W_KFULL] - (long long)dig * (long long)D[W_KTEXT]) *
             (int)D[W_KW];
    }
  };
  if (ktab) {
    for (unsigned s = tid; s < steps_k; s += THREADS) {
      long long a, b;
      int kv;
      kstep_bases(s, a, b, kv);
      kbA[s] = a;
      kbB[s] = b;
      kval[s] = kv;
    }
  }
  __syncthreads();

  // the host guarantees total_work < 2^31 (lowering.py)
  const unsigned tiles_all = tiles_m * tiles_n * tiles_b;
  const unsigned total_work = tiles_all * splitk;
  const unsigned steps_per_split = (steps_k + splitk - 1) / splitk;
  const unsigned nw = blockIdx.x < total_work ? (total_work - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
  if (nw == 0) return;

  auto work_krange = [&](unsigned j, unsigned& k0, unsigned& k1) {
    const unsigned w = blockIdx.x + j * gridDim.x;
    const unsigned ks = w / tiles_all;
    k0 = ks * steps_per_split;
    k1 = min(steps_k, k0 + steps_per_split);
  };

  // grid-base offsets of work item j -> tile-info slot j % TI.  One warp, one
  // lane per grid dim; n fastest so neighbouring CTAs share A tiles in L2.
  auto decode_tile = [&](unsigned j) {
    if (warp == 0) {
      unsigned t = (blockIdx.x + j * gridDim.x) % tiles_all;
      const unsigned in_ = t % tiles_n;
      t /= tiles_n;
      const unsigned im_ = t % tiles_m;
      const unsigned ib_ = t / tiles_m;
      long long a = 0, b = 0, c = 0;
      int vm = 0, vn = 0;
      for (int q = lane; q < n_gm; q += 32) {
        const int64_t* G = D + OFF_GM + q * 4;
        unsigned dig = (im_ / (unsigned)G[1]) % (unsigned)G[0];
        a += (long long)dig * G[2];
        c += (long long)dig * G[3];
        if (q == pgm)
          vm = (int)min((long long)D[W_MTEXT], (long long)D[W_MFULL] - (long long)dig * (long long)D[W_MTEXT]) *
               (int)D[W_MW];
      }
      for (int q = lane; q < n_gn; q += 32) {
        const int64_t* G = D + OFF_GN + q * 4;
        unsigned dig = (in_ / (unsigned)G[1]) % (unsigned)G[0];
        b += (long long)dig * G[2];
        c += (long long)dig * G[3];
        if (q == pgn)
          vn = (int)min((long long)D[W_NTEXT], (long long)D[W_NFULL] - (long long)dig * (long long)D[W_NTEXT]) *
               (int)D[W_NW];
      }
      for (int q = lane; q < n_gb; q += 32) {
        const int64_t* G = D + OFF_GB + q * 5;
        unsigned dig = (ib_ / (unsigned)G[1]) % (unsigned)G[0];
        a += (long long)dig * G[2];
        b += (long long)dig * G[3];
        c += (long long)dig * G[4];
      }
      a = warp_sum_ll(a);
      b = warp_sum_ll(b);
      c = warp_sum_ll(c);
      vm = warp_sum_i(vm);
      vn = warp_sum_i(vn);
      if (lane == 0) {
        const int slot = (int)(j % TI);
        ti_base[slot * 3 + 0] = a;
        ti_base[slot * 3 + 1] = b;
        ti_base[slot * 3 + 2] = c;
        ti_valid[slot * 2 + 0] = pgm < 0 ? MTa : vm;
        ti_valid[slot * 2 + 1] = pgn < 0 ? NTa : vn;
      }
    }
    __syncthreads();
  };

  // long contracted ranges: the k table is a window of KCHUNK steps, refilled
  // cooperatively (one step per thread) whenever the loader leaves it
  unsigned ktab_base = 0;
  auto refill_ktab = [&](unsigned first, unsigned last) {
    __syncthreads();  // every warp is done reading the previous window
    ktab_base = first;
    for (unsigned s = tid; s < (unsigned)KCHUNK && first + s < last; s += THREADS) {
      long long a, b;
      int kv;
      kstep_bases(first + s, a, b, kv);
      kbA[s] = a;
      kbB[s] = b;
      kval[s] = kv;
    }
    __syncthreads();
  };

  auto issue = [&](int st, int slot, unsigned step) {
    T* dA = sA + st * P::A_ELEMS;
    T* dB = sB + st * P::B_ELEMS;
    long long ka, kb;
    int kvi;
    ka = kbA[step - ktab_base];
    kb = kbB[step - ktab_base];
    kvi = kval[step - ktab_base];
    stage_kv[st] = kvi;  // every thread writes the same value
    const unsigned kv = (unsigned)kvi;
    const T* srcA = A + ti_base[slot * 3 + 0] + ka;
    const T* srcB = B + ti_base[slot * 3 + 1] + kb;
    const unsigned m_valid = (unsigned)ti_valid[slot * 2 + 0], n_valid = (unsigned)ti_valid[slot * 2 + 1];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const unsigned meta = metaA[i * THREADS + tid];
      const unsigned r = meta & 0xFFFFu, kk = meta >> 16;
      if (kk != 0xFFFFu) {
        const bool ok = (r < m_valid) && (kk < kv);
        cp_async_zfill<sizeof(T)>(dA + P::idxA(r, kk), ok ? (srcA + gA[i * THREADS + tid]) : A, ok);
      }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const unsigned meta = metaB[i * THREADS + tid];
      const unsigned c = meta & 0xFFFFu, kk = meta >> 16;
      if (kk != 0xFFFFu) {
        const bool ok = (c < n_valid) && (kk < kv);
        cp_async_zfill<sizeof(T)>(dB + P::idxB(c, kk), ok ? (srcB + gB[i * THREADS + tid]) : B, ok);
      }
    }
  };

  // loader cursor (runs ahead) and compute cursor over the same stream
  unsigned l_j = 0, l_step, l_k0, l_k1, g_issue = 0;
  unsigned c_j = 0, c_step, c_k1, g_comp = 0;
  work_krange(0, l_k0, l_k1);
  l_step = l_k0;
  c_step = l_k0;
  c_k1 = l_k1;

  auto loader_advance = [&]() {
    if (l_j < nw) {
      if (l_step == l_k0) decode_tile(l_j);  // entering a new tile (uniform branch)
      if (!ktab && (l_step == l_k0 || l_step >= ktab_base + KCHUNK)) refill_ktab(l_step, l_k1);
      issue((int)(g_issue % STAGES), (int)(l_j % TI), l_step);
      ++g_issue;
      ++l_step;
      if (l_step >= l_k1) {
        ++l_j;
        if (l_j < nw) {
          work_krange(l_j, l_k0, l_k1);
          l_step = l_k0;
        }
      }
    }
    cp_async_commit();
  };

  typename P::Acc acc;
  P::clear(acc);
#pragma unroll 1
  for (int s = 0; s < STAGES - 1; ++s) loader_advance();

  while (c_j < nw) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    loader_advance();
    const int st = (int)(g_comp % STAGES);
    P::compute(sA + st * P::A_ELEMS, sB + st * P::B_ELEMS, acc, stage_kv[st]);
    ++g_comp;
    ++c_step;
    if (c_step >= c_k1) {
      // ---- epilogue of tile c_j: store in the parent's index order (strided C)
      const int slot = (int)(c_j % TI);
      const long long baseC = ti_base[slot * 3 + 2];
      const int m_valid = ti_valid[slot * 2 + 0], n_valid = ti_valid[slot * 2 + 1];
      P::epilogue(
          acc, scratch,
          [&](int r, int c, T v) {
            if (r < m_valid && c < n_valid) {
              T* p = C + baseC + offMC[r] + offNC[c];
              if (atomic) {
                atomic_add_of(p, v);
              } else if (accumulate) {
                *p = add_of(*p, v);
              } else {
                *p = v;
              }
            }
          },
          [&](int r, int c, T v0, T v1) {
            // only called when pair_ok: columns c, c+1 are adjacent and 32B aligned
            if (r < m_valid && c < n_valid) store_pair_of(C + baseC + offMC[r] + offNC[c], v0, v1);
          },
          pair_ok);
      P::clear(acc);
      ++c_j;
      if (c_j < nw) {
        unsigned k0;
        work_krange(c_j, k0, c_k1);
        c_step = k0;
      }
    }
  }
  cp_async_wait<0>();
}


// ------------------------------------------------------------------ single operand
// out[o] = sum_s X[off_o(o) + off_s(s)]  (diag via summed strides; contract.py:332-361)
template <typename T>
__global__ void single_kernel(const int64_t* __restrict__ D, const T* __restrict__ X, T* __restrict__ out) {
  const int n_o = (int)D[S_NO], n_s = (int)D[S_NS];
  const long long out_elems = D[S_OUT_ELEMS], sum_elems = D[S_SUM_ELEMS];
  const bool accumulate = (D[S_FLAGS] & 1) != 0;
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < out_elems;
       o += (long long)gridDim.x * blockDim.x) {
    long long e = o, xo = 0, oo = 0;
    for (int d = 0; d < n_o; ++d) {
      const int64_t* L = D + OFF_SO + d * 3;
      long long dig = e % L[0];
      e /= L[0];
      xo += dig * L[1];
      oo += dig * L[2];
    }
    T acc = zero_of<T>();
    for (long long s = 0; s < sum_elems; ++s) {
      long long e2 = s, xs = 0;
      for (int d = 0; d < n_s; ++d) {
        const int64_t* L = D + OFF_SS + d * 2;
        xs += (e2 % L[0]) * L[1];
        e2 /= L[0];
      }
      acc = add_of(acc, X[xo + xs]);
    }
    out[oo] = accumulate ? add_of(out[oo], acc) : acc;
  }
}

// ------------------------------------------------------------------ strip_exponent helpers
// contract.py:816-829: factor = max|p|; exponent += log10(factor); p /= factor.
__device__ __forceinline__ double abs_of(float v) { return fabs((double)v); }
__device__ __forceinline__ double abs_of(double v) { return fabs(v); }
__device__ __forceinline__ double abs_of(float2 v) { return hypot((double)v.x, (double)v.y); }
__device__ __forceinline__ double abs_of(double2 v) { return hypot(v.x, v.y); }

template <typename T>
__global__ void absmax_kernel(const T* __restrict__ p, long long n, unsigned long long* __restrict__ slot) {
  double m = 0.0;
  bool nan = false;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double a = abs_of(p[i]);
    if (a != a) nan = true;
    m = fmax(m, a);
  }
  if (nan) m = __longlong_as_double(0x7ff8000000000000LL);
  // non-negative doubles (and +NaN) order like their bit patterns
  unsigned long long bits = (unsigned long long)__double_as_longlong(m);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    unsigned long long o = __shfl_xor_sync(0xffffffffu, bits, d);
    bits = o > bits ? o : bits;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(slot, bits);
}

__device__ __forceinline__ float scale_of(float v, double s) { return (float)(v / (float)s); }
__device__ __forceinline__ double scale_of(double v, double s) { return v / s; }
__device__ __forceinline__ float2 scale_of(float2 v, double s) {
  float f = (float)s;
  return make_float2(v.x / f, v.y / f);
}
__device__ __forceinline__ double2 scale_of(double2 v, double s) { return make_double2(v.x / s, v.y / s); }

// p /= factor ; exponent += log10(factor)   (block 0 / thread 0 updates the exponent)
template <typename T>
__global__ void strip_kernel(T* __restrict__ p, long long n, const unsigned long long* __restrict__ slot,
                             double* __restrict__ exponent) {
  const double f = __longlong_as_double((long long)*slot);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = scale_of(p[i], f);
  if (blockIdx.x == 0 && threadIdx.x == 0) *exponent += log10(f);
}

__device__ __forceinline__ float mulr_of(float v, double s) { return (float)(v * s); }
__device__ __forceinline__ double mulr_of(double v, double s) { return v * s; }
__device__ __forceinline__ float2 mulr_of(float2 v, double s) { return make_float2((float)(v.x * s), (float)(v.y * s)); }
__device__ __forceinline__ double2 mulr_of(double2 v, double s) { return make_double2(v.x * s, v.y * s); }

// Exponent-aware slice accumulation (core.py:163-170):
//   e = max(E, es);  out = out * 10^(E - e) (+ chunk: m * 10^(es - e));  E = e
// Phase 0 rescales the whole output (early-out when the scale is exactly 1),
// phase 1 adds the slice mantissa into its chunk, phase 2 commits E.
template <typename T>
__global__ void rescale_out_kernel(T* __restrict__ out, long long n, const double* __restrict__ E,
                                   const double* __restrict__ es) {
  const double e = fmax(*E, *es);
  const double so = (*E == e) ? 1.0 : pow(10.0, *E - e);
  if (so == 1.0) return;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = mulr_of(out[i], so);
}
template <typename T>
__global__ void add_chunk_kernel(const int64_t* __restrict__ D, T* __restrict__ out, const T* __restrict__ m,
                                 const double* __restrict__ E, const double* __restrict__ es) {
  // D: single-operand descriptor mapping the dense slice result onto the chunk
  const double e = fmax(*E, *es);
  const double sn = (*es == e) ? 1.0 : pow(10.0, *es - e);
  const int n_o = (int)D[S_NO];
  const long long n = D[S_OUT_ELEMS];
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < n; o += (long long)gridDim.x * blockDim.x) {
    long long t = o, xo = 0, oo = 0;
    for (int d = 0; d < n_o; ++d) {
      const int64_t* L = D + OFF_SO + d * 3;
      long long dig = t % L[0];
      t /= L[0];
      xo += dig * L[1];
      oo += dig * L[2];
    }
    out[oo] = add_of(out[oo], mulr_of(m[xo], sn));
  }
}
__global__ void commit_exponent_kernel(double* __restrict__ E, const double* __restrict__ es) {
  *E = fmax(*E, *es);
}
__global__ void set_double_kernel(double* p, double v) { *p = v; }
__global__ void copy_double_kernel(double* dst, const double* src) { *dst = *src; }
__global__ void zero_slot_kernel(unsigned long long* p) { *p = 0ull; }

}  // namespace ctgb
