// dotstream.cuh -- the final inner product of a contraction tree (M = N = 1, one batch):
// out = sum_k A[k] * B[k] over two operands of up to 2^30 elements whose index orders differ.
//
// The staged KRED policy moves every element with its own cp.async; for 8-byte types that
// is one LSU wavefront per lane (3.5 TB/s for complex64, ncu: LSU bound).  Here a thread
// owns the same DS_U tile-local elements of every 2048-element tile: their offsets in A and
// in B (the tile dims' digits times the strides) are launch-invariant and live in registers,
// the tile base (grid dims) is computed once per tile by each warp (one lane per grid dim +
// warp reduction), and the operands come straight from global memory with 2 * DS_U loads in
// flight per thread -- neighbouring lanes read neighbouring elements of A, and a permutation
// of them in B that stays inside the same few sectors (L1 serves the rest).
// Block partial sums are added atomically into the (zeroed) output.
// (included inside namespace ctgb)
#pragma once

constexpr int DOT_KT = 2048, DOT_THREADS = 256, DOT_U = DOT_KT / DOT_THREADS;

template <typename T>
__global__ void __launch_bounds__(DOT_THREADS, 2)
dotstream_kernel(const int64_t* __restrict__ D, const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C) {
  __shared__ T s_part[DOT_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31;
  const int n_tk = (int)D[W_NTK], n_gk = (int)D[W_NGK];
  const int KTa = (int)D[W_KTA];
  const unsigned steps = (unsigned)D[W_STEPS_K];  // the host guarantees < 2^31
  // tile-local offsets of this thread's elements (tile dims: dim 0 fastest)
  long long la[DOT_U], lb[DOT_U];
  bool in_tile[DOT_U];
#pragma unroll
  for (int j = 0; j < DOT_U; ++j) {
    unsigned e = (unsigned)(tid + j * DOT_THREADS);
    in_tile[j] = e < (unsigned)KTa;
    if (!in_tile[j]) e = 0;
    long long a = 0, b = 0;
    for (int d = 0; d < n_tk; ++d) {
      const int64_t* L = D + OFF_TK + d * 3;
      const unsigned ext = (unsigned)L[0];
      a += (long long)(e % ext) * L[1];
      b += (long long)(e % ext) * L[2];
      e /= ext;
    }
    la[j] = a;
    lb[j] = b;
  }
  // this lane's grid dims (at most 2 per lane: MAX_G = 40 <= 64)
  unsigned g_ext[2] = {1u, 1u};
  unsigned g_div[2] = {1u, 1u};
  long long g_sa[2] = {0, 0}, g_sb[2] = {0, 0};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int q = lane + 32 * h;
    if (q < n_gk) {
      const int64_t* G = D + OFF_GK + q * 4;
      g_ext[h] = (unsigned)G[0];
      g_div[h] = (unsigned)G[1];
      g_sa[h] = G[2];
      g_sb[h] = G[3];
    }
  }
  T acc = zero_of<T>();
  for (unsigned t = blockIdx.x; t < steps; t += gridDim.x) {
    long long ta = 0, tb = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const unsigned dig = (t / g_div[h]) % g_ext[h];
      ta += (long long)dig * g_sa[h];
      tb += (long long)dig * g_sb[h];
    }
    ta = warp_sum_ll(ta);
    tb = warp_sum_ll(tb);
    T a[DOT_U], b[DOT_U];
#pragma unroll
    for (int j = 0; j < DOT_U; ++j) {
      a[j] = in_tile[j] ? A[ta + la[j]] : zero_of<T>();
      b[j] = in_tile[j] ? B[tb + lb[j]] : zero_of<T>();
    }
#pragma unroll
    for (int j = 0; j < DOT_U; ++j) mac(acc, a[j], b[j]);
  }
  // block reduction, one atomic per block
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc = add_of(acc, shfl_down_of(acc, d));
  if (lane == 0) s_part[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    T v = s_part[0];
    for (int w = 1; w < DOT_THREADS / 32; ++w) v = add_of(v, s_part[w]);
    atomic_add_of(C, v);
  }
}
