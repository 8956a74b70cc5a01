// dmmastream.cuh -- fp64 tensor-core streaming kernel for narrow complex128 nodes
// (N <= 8*NJ, K <= 64, no batch): the rowstream idea with DMMA fragments.
//
// Every warp owns blocks of 32 output rows and runs them start to finish on its own:
// A fragments straight from global memory into registers (a lane of the m8k4 A fragment
// holds one complex element: one LDG.128, 16 of them in flight per lane), B fragments from
// a zero-padded shared-memory copy made once per CTA, 4 real DMMAs per fragment pair, 256-bit
// stores.  No operand staging, no producer warps, no CTA barriers in the loop: the warps of
// an SM drift apart, so loads, DMMAs and stores of different row blocks overlap by themselves
// -- which the staged 256x16 policy could not do (all consumer warps share one phase: ncu
// showed its N=16 K=16 node at 3.3 TB/s, compute and HBM time adding up instead of overlapping).
// (included inside namespace ctgb)
#pragma once

constexpr int DS_KMAX = 64;

// NJ = column fragments (N <= 8*NJ).  128 threads x 3 blocks (NJ <= 2: <= 170 registers) or x 2 blocks
// (NJ = 4: 128 accumulator registers), 16 loads in flight per lane.  NJ = 1 serves skinny nodes
// whose contracted space is too long for the row-stream kernel (N <= 8, 8 < K <= 64: the staged
// row policy ran the M = 2^22, N = 8, K = 64 node of the Sycamore slice at 0.57 of its roofline)
template <int NJ, bool STRIP = false>
__global__ void __launch_bounds__(128, NJ <= 2 ? 3 : 2)
dmmastream_kernel(const int64_t* __restrict__ D, const double2* __restrict__ A, const double2* __restrict__ B,
                  double2* __restrict__ C) {
  constexpr int DS_NMAX = NJ * 8;
  __shared__ long long s_akoff[DS_KMAX], s_bkoff[DS_KMAX], s_bnoff[DS_NMAX], s_cnoff[DS_NMAX];
  __shared__ long long s_msA[RS_MAXDIMS], s_msC[RS_MAXDIMS];
  __shared__ unsigned s_mext[RS_MAXDIMS];
  __shared__ double2 s_B[DS_KMAX * DS_NMAX];  // [k][n], zero beyond (K, N)
  const int tid = threadIdx.x, lane = tid & 31;
  const int n_tm = (int)D[W_NTM], n_gm = (int)D[W_NGM], n_tk = (int)D[W_NTK], n_tn = (int)D[W_NTN];
  const int K = (int)D[W_KTA], N = (int)D[W_NTA];
  const int n_m = n_tm + n_gm;
  const bool accumulate = (D[W_FLAGS] & 1) != 0;
  const bool pair_ok = (D[W_FLAGS] & 2) != 0 && !accumulate;
  const bool pow2 = (D[W_FLAGS] & 8) != 0;  // every m dim (tile and grid) is a power of two
  // m dims in enumeration order: tile dims (dim 0 fastest) then grid dims
  for (int d = tid; d < n_m; d += blockDim.x) {
    if (d < n_tm) {
      const int64_t* L = D + OFF_TM + d * 3;
      s_mext[d] = (unsigned)L[0];
      s_msA[d] = L[1];
      s_msC[d] = L[2];
    } else {
      const int64_t* G = D + OFF_GM + (d - n_tm) * 4;
      s_mext[d] = (unsigned)G[0];
      s_msA[d] = G[2];
      s_msC[d] = G[3];
    }
  }
  if (tid < DS_KMAX) {
    long long a = 0, b = 0;
    if (tid < K) {
      unsigned e = tid;
      for (int d = 0; d < n_tk; ++d) {
        const int64_t* L = D + OFF_TK + d * 3;
        const unsigned ext = (unsigned)L[0];
        a += (long long)(e % ext) * L[1];
        b += (long long)(e % ext) * L[2];
        e /= ext;
      }
    }
    s_akoff[tid] = a;
    s_bkoff[tid] = b;
  }
  if (tid >= 32 && tid < 32 + DS_NMAX) {
    const int c = tid - 32;
    long long b = 0, o = 0;
    if (c < N) {
      unsigned e = c;
      for (int d = 0; d < n_tn; ++d) {
        const int64_t* L = D + OFF_TN + d * 3;
        const unsigned ext = (unsigned)L[0];
        b += (long long)(e % ext) * L[1];
        o += (long long)(e % ext) * L[2];
        e /= ext;
      }
    }
    s_bnoff[c] = b;
    s_cnoff[c] = o;
  }
  __syncthreads();
  for (int i = tid; i < DS_KMAX * DS_NMAX; i += blockDim.x) {
    const int kk = i / DS_NMAX, c = i % DS_NMAX;
    s_B[i] = (kk < K && c < N) ? B[s_bkoff[kk] + s_bnoff[c]] : make_double2(0.0, 0.0);
  }
  __syncthreads();

  [[maybe_unused]] StripCtx sctx;  // fused strip_exponent: its own instantiation (register-bound loop)
  if constexpr (STRIP) sctx = strip_begin(D);
  const int frow = lane >> 2, fk = lane & 3, fc = (lane & 3) * 2;
  const int n8s = (N + 7) >> 3;       // column fragments in use
  const int kchunks = (K + 15) >> 4;  // chunks of 16 k (4 k4-steps each)
  const unsigned long long M = (unsigned long long)D[W_MTA] * (unsigned long long)D[W_TILES_M];
  const unsigned long long nblk = (M + 31) >> 5;
  const unsigned long long wstride = (unsigned long long)gridDim.x * (blockDim.x >> 5);
  for (unsigned long long blk = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (tid >> 5); blk < nblk;
       blk += wstride) {
    // the lane's four rows (one per m8 fragment)
    long long oa[4], oc[4];
    bool live[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned long long m = blk * 32 + (unsigned)(i * 8 + frow);
      live[i] = m < M;
      unsigned e = live[i] ? (unsigned)m : 0u;
      long long xa = 0, xc = 0;
      if (pow2) {
        for (int d = 0; d < n_m; ++d) {
          const unsigned ext = s_mext[d];
          const unsigned dig = e & (ext - 1);
          e >>= 31 - __clz(ext);
          xa += (long long)dig * s_msA[d];
          xc += (long long)dig * s_msC[d];
        }
      } else {
        for (int d = 0; d < n_m; ++d) {
          const unsigned ext = s_mext[d];
          const unsigned dig = e % ext;
          e /= ext;
          xa += (long long)dig * s_msA[d];
          xc += (long long)dig * s_msC[d];
        }
      }
      oa[i] = xa;
      oc[i] = xc;
    }
    double re[4][NJ][2], im[4][NJ][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) re[i][j][0] = re[i][j][1] = im[i][j][0] = im[i][j][1] = 0.0;
    for (int kc = 0; kc < kchunks; ++kc) {
      // 16 independent 128-bit loads per lane: element (row i*8 + frow, k = kc*16 + k4*4 + fk)
      double2 a[4][4];
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const int kk = kc * 16 + k4 * 4 + fk;
        const long long ko = s_akoff[kk & (DS_KMAX - 1)];
        const bool kin = kk < K;
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i][k4] = kin ? A[oa[i] + ko] : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        if (kc * 16 + k4 * 4 >= K) break;  // uniform
        // B fragment: lane holds B[k = .. + fk][n = j*8 + frow]
        double2 b[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) b[j] = s_B[(kc * 16 + k4 * 4 + fk) * DS_NMAX + j * 8 + frow];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j)
            if (j < n8s) dmma8x8x4(re[i][j][0], re[i][j][1], a[i][k4].x, b[j].x);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j)
            if (j < n8s) dmma8x8x4(im[i][j][0], im[i][j][1], a[i][k4].x, b[j].y);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j)
            if (j < n8s) dmma8x8x4(re[i][j][0], re[i][j][1], -a[i][k4].y, b[j].y);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j)
            if (j < n8s) dmma8x8x4(im[i][j][0], im[i][j][1], a[i][k4].y, b[j].x);
      }
    }
    if constexpr (STRIP) {
      if (!sctx.scale) {
        // max|C|: integer scan over the accumulators, then (rarely) the values (see gett_ws.cuh)
        int hmax = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e)
              hmax = max(hmax, max(strip_hi(re[i][j][e]), strip_hi(im[i][j][e])));
        if (strip_hot<double2>(sctx, hmax)) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
              for (int e = 0; e < 2; ++e)
                if (live[i] && j * 8 + fc + e < N) strip_track(sctx, re[i][j][e], im[i][j][e]);
        }
      }
    }
    // a lane owns columns (fc, fc+1) of fragment j in row i*8 + frow
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!live[i]) continue;
      double2* crow = C + oc[i];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c = j * 8 + fc;
        if (c >= N) continue;
        double2 v0 = make_double2(re[i][j][0], im[i][j][0]), v1 = make_double2(re[i][j][1], im[i][j][1]);
        if constexpr (STRIP) {
          if (sctx.scale) {
            v0 = strip_apply(sctx, v0);
            if (c + 1 < N) v1 = strip_apply(sctx, v1);
          }
        }
        if (pair_ok) {
          store_pair_of(crow + s_cnoff[c], v0, v1);
        } else {
          double2* p0 = crow + s_cnoff[c];
          *p0 = accumulate ? add_of(*p0, v0) : v0;
          if (c + 1 < N) {
            double2* p1 = crow + s_cnoff[c + 1];
            *p1 = accumulate ? add_of(*p1, v1) : v1;
          }
        }
      }
    }
  }
  if constexpr (STRIP) strip_end(sctx);
}
