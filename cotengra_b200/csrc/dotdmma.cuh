// dotdmma.cuh -- a small result over a huge contracted range on the fp64 tensor cores
// (complex128, M, N <= 32, one batch):
//     R[m, n] = sum_k A[k, m] * B[k, n],      K ~ 2^25
// the node stem fusion produces when it peels the last few small tensors of a stem over the
// final inner product (cotengra_b200/fusion.py): two 16 GiB operands are read once, the 32 x 32
// result stays in DMMA accumulator fragments.  Every warp streams its own k: a lane of the m8k4
// A fragment is one complex element of A (k = its fragment column, m = its fragment row), a lane
// of the k4n8 B fragment one element of B, both straight from global memory; 4 real DMMAs per
// fragment pair as in the staged policies.  Intensity at M = N = 32 is 8 flop/B, just above the
// fp64 ridge, so the kernel wants both pipes busy: 16 loads in flight per lane ahead of 128 DMMAs.
// Warp partial sums are combined through shared memory, one atomic per element and block.
// (included inside namespace ctgb)
#pragma once

constexpr int DD_WARPS = 8, DD_S = 4, DD_KT = DD_WARPS * DD_S * 4;  // 128 k per tile and block

template <int FM, int FN>
__global__ void __launch_bounds__(DD_WARPS * 32, 1)
dotdmma_kernel(const int64_t* __restrict__ D, const double2* __restrict__ A, const double2* __restrict__ B,
               double2* __restrict__ C) {
  __shared__ double2 s_red[FM * 8][FN * 8 + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frow = lane >> 2, fk = lane & 3;
  const int n_tk = (int)D[W_NTK], n_gk = (int)D[W_NGK];
  const int MTa = (int)D[W_MTA], NTa = (int)D[W_NTA];
  const unsigned steps = (unsigned)D[W_STEPS_K];  // the host guarantees < 2^31 and exact tiles
  auto decode = [&](unsigned e, int off, int n, int col) -> long long {
    long long o = 0;
    for (int d = 0; d < n; ++d) {
      const int64_t* L = D + off + d * 3;
      o += (long long)(e % (unsigned)L[0]) * L[col];
      e /= (unsigned)L[0];
    }
    return o;
  };
  // this lane's rows of A (m = 8 i + frow) and columns of B (n = 8 j + frow); tile-local offsets
  // are 32-bit (the host only picks this kernel for operands below 2^32 elements): 255 registers
  // with 64-bit ones, and spills inside the loop
  unsigned am[FM], bn[FN];
  bool mok[FM], nok[FN];
#pragma unroll
  for (int i = 0; i < FM; ++i) {
    mok[i] = 8 * i + frow < MTa;
    am[i] = mok[i] ? (unsigned)decode((unsigned)(8 * i + frow), OFF_TM, (int)D[W_NTM], 1) : 0u;
  }
#pragma unroll
  for (int j = 0; j < FN; ++j) {
    nok[j] = 8 * j + frow < NTa;
    bn[j] = nok[j] ? (unsigned)decode((unsigned)(8 * j + frow), OFF_TN, (int)D[W_NTN], 1) : 0u;
  }
  // tile-local k of this lane: e = warp * 16 + 4 s + fk
  unsigned la[DD_S], lb[DD_S];
#pragma unroll
  for (int s = 0; s < DD_S; ++s) {
    const unsigned e = (unsigned)(warp * (DD_S * 4) + 4 * s + fk);
    la[s] = (unsigned)decode(e, OFF_TK, n_tk, 1);
    lb[s] = (unsigned)decode(e, OFF_TK, n_tk, 2);
  }
  // grid dims of the contracted space, at most two per lane (MAX_G = 40 <= 64)
  unsigned g_ext[2] = {1u, 1u}, g_div[2] = {1u, 1u};
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
  double re[FM][FN][2], im[FM][FN][2];
#pragma unroll
  for (int i = 0; i < FM; ++i)
#pragma unroll
    for (int j = 0; j < FN; ++j) re[i][j][0] = re[i][j][1] = im[i][j][0] = im[i][j][1] = 0.0;
  const double2 zero = make_double2(0.0, 0.0);
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
#pragma unroll
    for (int s0 = 0; s0 < DD_S; s0 += 2) {
      // two k4 steps of loads in flight (2 x (FM + FN) x 16 B per lane), then their DMMAs
      double2 a[2][FM], b[2][FN];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
#pragma unroll
        for (int i = 0; i < FM; ++i) a[u][i] = mok[i] ? A[ta + (long long)(la[s0 + u] + am[i])] : zero;
#pragma unroll
        for (int j = 0; j < FN; ++j) b[u][j] = nok[j] ? B[tb + (long long)(lb[s0 + u] + bn[j])] : zero;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma8x8x4(re[i][j][0], re[i][j][1], a[u][i].x, b[u][j].x);
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma8x8x4(im[i][j][0], im[i][j][1], a[u][i].x, b[u][j].y);
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma8x8x4(re[i][j][0], re[i][j][1], -a[u][i].y, b[u][j].y);
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma8x8x4(im[i][j][0], im[i][j][1], a[u][i].y, b[u][j].x);
      }
    }
  }
  // ---- combine the warps' partial sums: a lane owns (m = 8 i + frow, n = 8 j + 2 fk + {0, 1})
  for (int w = 0; w < DD_WARPS; ++w) {
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            double2& dst = s_red[8 * i + frow][8 * j + 2 * fk + e];
            const double2 v = make_double2(re[i][j][e], im[i][j][e]);
            dst = (w == 0) ? v : make_double2(dst.x + v.x, dst.y + v.y);
          }
    }
    __syncthreads();
  }
  StripCtx sctx = strip_begin(D);  // strip_exponent: block partial sums are only scaled
  for (int idx = tid; idx < FM * 8 * FN * 8; idx += DD_WARPS * 32) {
    const int m = idx / (FN * 8), n = idx % (FN * 8);
    if (m < MTa && n < NTa) {
      double2 v = s_red[m][n];
      if (sctx.scale) v = strip_apply(sctx, v);
      atomic_add_of(C + decode((unsigned)m, OFF_TM, (int)D[W_NTM], 2) + decode((unsigned)n, OFF_TN, (int)D[W_NTN], 2), v);
    }
  }
}
