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
// the same stream with a small kept space on both operands (M, N <= 4): the last small tensor
// of a stem peeled over the final inner product (cotengra_b200/fusion.py),
//   R[m, n] = sum_k A[k, m] * B[k, n],
// 16 accumulators per thread, 4 k per thread and tile (16 + 16 loads in flight: 128 KB per SM,
// what the 1x1 kernel needs for 6.9 TB/s; with 2 k it stopped at 5.6 TB/s)
constexpr int DOT4_MN = 4;
// k per thread and tile: 4 for 16-byte elements, 8 for narrower ones (the same 128 KB in flight)
template <typename T> constexpr int dot4_u() { return sizeof(T) >= 16 ? 4 : 8; }
template <typename T> constexpr int dot4_kt() { return dot4_u<T>() * DOT_THREADS; }

template <typename T, int MT, int NT, int U>
__global__ void __launch_bounds__(DOT_THREADS, (MT * NT > 1) ? 1 : 2)
dotstream_kernel(const int64_t* __restrict__ D, const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C) {
  __shared__ T s_part[DOT_THREADS / 32][MT * NT];
  const int tid = threadIdx.x, lane = tid & 31;
  const int n_tk = (int)D[W_NTK], n_gk = (int)D[W_NGK];
  const int KTa = (int)D[W_KTA], MTa = (int)D[W_MTA], NTa = (int)D[W_NTA];
  const unsigned steps = (unsigned)D[W_STEPS_K];  // the host guarantees < 2^31
  // offsets of the kept indices (all of M and N sit inside the tile)
  long long am[MT], bn[NT];
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    long long a = 0;
    unsigned e = (unsigned)i;
    if (MT > 1 && i < MTa)
      for (int d = 0; d < (int)D[W_NTM]; ++d) {
        const int64_t* L = D + OFF_TM + d * 3;
        a += (long long)(e % (unsigned)L[0]) * L[1];
        e /= (unsigned)L[0];
      }
    am[i] = a;
  }
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    long long b = 0;
    unsigned e = (unsigned)i;
    if (NT > 1 && i < NTa)
      for (int d = 0; d < (int)D[W_NTN]; ++d) {
        const int64_t* L = D + OFF_TN + d * 3;
        b += (long long)(e % (unsigned)L[0]) * L[1];
        e /= (unsigned)L[0];
      }
    bn[i] = b;
  }
  // tile-local offsets of this thread's elements (tile dims: dim 0 fastest)
  long long la[U], lb[U];
  bool in_tile[U];
#pragma unroll
  for (int j = 0; j < U; ++j) {
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
  T acc[MT][NT];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int c = 0; c < NT; ++c) acc[i][c] = zero_of<T>();
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
    T a[U][MT], b[U][NT];
#pragma unroll
    for (int j = 0; j < U; ++j) {
#pragma unroll
      for (int i = 0; i < MT; ++i) a[j][i] = (in_tile[j] && i < MTa) ? A[ta + la[j] + am[i]] : zero_of<T>();
#pragma unroll
      for (int c = 0; c < NT; ++c) b[j][c] = (in_tile[j] && c < NTa) ? B[tb + lb[j] + bn[c]] : zero_of<T>();
    }
#pragma unroll
    for (int j = 0; j < U; ++j)
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int c = 0; c < NT; ++c) mac(acc[i][c], a[j][i], b[j][c]);
  }
  // block reduction, one atomic per block and output element
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int c = 0; c < NT; ++c) {
      T v = acc[i][c];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v = add_of(v, shfl_down_of(v, d));
      if (lane == 0) s_part[tid >> 5][i * NT + c] = v;
    }
  __syncthreads();
  if (tid < MT * NT) {
    const int i = tid / NT, c = tid % NT;
    if (i < MTa && c < NTa) {
      T v = s_part[0][tid];
      for (int w = 1; w < DOT_THREADS / 32; ++w) v = add_of(v, s_part[w][tid]);
      StripCtx sctx = strip_begin(D);  // strip_exponent: partial sums are only scaled here
      if (sctx.scale) v = strip_apply(sctx, v);
      long long oc = 0;
      unsigned e = (unsigned)i;
      for (int d = 0; d < (int)D[W_NTM]; ++d) {
        const int64_t* L = D + OFF_TM + d * 3;
        oc += (long long)(e % (unsigned)L[0]) * L[2];
        e /= (unsigned)L[0];
      }
      e = (unsigned)c;
      for (int d = 0; d < (int)D[W_NTN]; ++d) {
        const int64_t* L = D + OFF_TN + d * 3;
        oc += (long long)(e % (unsigned)L[0]) * L[2];
        e /= (unsigned)L[0];
      }
      atomic_add_of(C + oc, v);
    }
  }
}
