// tc05_policy.cuh -- complex64 dense nodes on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM).  Included inside namespace ctgb.
//
// A complex tile product C[128 x NT] += A[128 x 16] * B[16 x NT] is run as the real
// product  C'[128 x 2NT] += A'[128 x 32] * B'[2NT x 32]^T  where A' is A's own
// (re, im)-interleaved memory image (so the strided gather of A lands directly in
// UMMA's K-major no-swizzle canonical layout) and B' is the 2x2-block embedding
//   B'[2n][2k] = Br   B'[2n][2k+1] = -Bi   B'[2n+1][2k] = Bi   B'[2n+1][2k+1] = Br
// so that C' is C's own interleaved image.  fp32 accuracy comes from the 3xTF32
// split  D += A'lo*B'hi + A'hi*B'lo + A'hi*B'hi  (the tensor core truncates its
// operands to tf32, so "hi" is the raw fp32 word and lo = x - trunc_tf32(x)):
// 1.3e-6 relative on a K = 64 tile (scripts/ubench/umma_c64.cu).
//
//   * B' (hi and lo, already in shared-memory tile order) is prepared once per
//     launch by bprime_kernel -- B is the small operand, <= a few MB -- and each
//     stage's pair of tiles arrives with ONE TMA bulk copy (cp.async.bulk,
//     mbarrier complete_tx);
//   * A'lo is produced from the gathered A' stage by the consumer warps; when the A tile
//     is made of long contiguous runs the producers fetch the runs with TMA bulk copies
//     into a staging area and the same pass scatters them into the UMMA layout (the
//     8-byte LDGSTS gather costs one LSU wavefront per lane: 66 % LSU pipe under ncu);
//   * one elected thread issues the 12 UMMAs of a stage (3 passes x 4 k-steps of
//     M128 x N(2NT) x K8) and commits them to the stage's "empty" mbarrier;
//   * the epilogue reads TMEM with tcgen05.ld (32 lanes x 8 columns = 4 complex
//     per thread and instruction) and writes 32-byte sectors.
#pragma once

template <int NT_, int STAGES_>
struct Tc05Policy {
  static constexpr bool IS_TC05 = true;
  static constexpr int MT = 128, NT = NT_, KT = 16, STAGES = STAGES_;
  static constexpr int THREADS = 256;
  static constexpr int TILE_FLOATS = 8 * (2 * NT) * 4;       // one B' tile: [8 chunks][2NT rows][4 floats]
  static constexpr int PAIR_BYTES = 2 * TILE_FLOATS * 4;     // hi + lo
  static constexpr int A_GATHER = MT * KT, B_GATHER = 0;      // float2 elements fetched by the producers
  static constexpr int A_ELEMS = 3 * MT * KT;                 // A'hi | A'lo | bulk-copy staging (float2 units)
  static constexpr int B_ELEMS = TILE_FLOATS;                 // hi + lo tiles (2*TILE_FLOATS floats)
  static constexpr int SCRATCH_ELEMS = 0;
  static constexpr int CONSUMER_REGS = 0, PRODUCER_REGS = 0;
  static constexpr bool HAS_BCACHE = false;
  static constexpr int MIN_BLOCKS = 1;
  static constexpr int TMEM_COLS = 2 * NT;                    // fp32 columns of the accumulator
  static_assert(TMEM_COLS == 64 || TMEM_COLS == 128, "two accumulators of TMEM_COLS columns must fit 512 and be a power of two");
  struct Acc {};
  // [chunk = k'/4][row][k'%4] floats == [kk/2][row][kk%2] complex elements
  __device__ static __forceinline__ int idxA(int r, int kk) { return ((kk >> 1) * MT + r) * 2 + (kk & 1); }
  __device__ static __forceinline__ int idxB(int c, int kk) { return 0; }
  __device__ static __forceinline__ void clear(Acc&) {}
  __device__ static __forceinline__ void compute(const float2*, const float2*, Acc&, int, int) {}
  template <typename F, typename F2>
  __device__ static __forceinline__ void epilogue(Acc&, float2*, F&&, F2&&, bool, int) {}
};

__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
  // version=1 [46,48), layout_type [61,64) = 0 (no swizzle / interleave)
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// B -> B'hi / B'lo in shared-memory tile order:
//   Bp[((ib*tiles_n + in)*steps_k + step)][hi|lo][chunk 0..7][row 0..2NT-1][4 floats]
template <int NT>
__global__ void __launch_bounds__(256) bprime_kernel(const int64_t* __restrict__ D, const float2* __restrict__ B,
                                                     float* __restrict__ Bp) {
  constexpr int TILE = 8 * (2 * NT) * 4;
  const int n_tn = (int)D[W_NTN], n_tk = (int)D[W_NTK], n_gn = (int)D[W_NGN], n_gk = (int)D[W_NGK], n_gb = (int)D[W_NGB];
  const unsigned tiles_n = (unsigned)D[W_TILES_N], tiles_b = (unsigned)D[W_TILES_B], steps_k = (unsigned)D[W_STEPS_K];
  const unsigned long long total = (unsigned long long)tiles_b * tiles_n * steps_k * TILE;
  for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned e = (unsigned)(idx % TILE);
    unsigned t = (unsigned)(idx / TILE);
    const unsigned step = t % steps_k;
    t /= steps_k;
    const unsigned in_ = t % tiles_n, ib_ = t / tiles_n;
    const unsigned j = e & 3, row = (e >> 2) % (2 * NT), chunk = (e >> 2) / (2 * NT);
    const unsigned kp = chunk * 4 + j, kk = kp >> 1, p = kp & 1, c = row >> 1, q = row & 1;
    long long off = 0;
    {
      unsigned x = c;
      for (int d = 0; d < n_tn; ++d) {
        const int64_t* L = D + OFF_TN + d * 3;
        off += (long long)(x % (unsigned)L[0]) * L[1];
        x /= (unsigned)L[0];
      }
      x = kk;
      for (int d = 0; d < n_tk; ++d) {
        const int64_t* L = D + OFF_TK + d * 3;
        off += (long long)(x % (unsigned)L[0]) * L[2];
        x /= (unsigned)L[0];
      }
      for (int d = 0; d < n_gn; ++d) {
        const int64_t* G = D + OFF_GN + d * 4;
        off += (long long)((in_ / (unsigned)G[1]) % (unsigned)G[0]) * G[2];
      }
      for (int d = 0; d < n_gk; ++d) {
        const int64_t* G = D + OFF_GK + d * 4;
        off += (long long)((step / (unsigned)G[1]) % (unsigned)G[0]) * G[3];
      }
      for (int d = 0; d < n_gb; ++d) {
        const int64_t* G = D + OFF_GB + d * 5;
        off += (long long)((ib_ / (unsigned)G[1]) % (unsigned)G[0]) * G[3];
      }
    }
    const float2 b = B[off];
    const float v = (q == p) ? b.x : (q == 0 ? -b.y : b.y);
    const unsigned long long base = (idx / TILE) * (2ull * TILE);
    Bp[base + e] = v;                        // hi: raw fp32 (the tensor core truncates)
    Bp[base + TILE + e] = v - trunc_tf32(v); // lo
  }
}
