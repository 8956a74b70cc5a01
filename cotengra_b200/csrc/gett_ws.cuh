// gett_ws.cuh -- warp-specialised skeleton of the permutation-fused contraction
// kernel (included from gett_kernels.cuh, after the compute policies).
//
//   producer warps (PRODUCER_THREADS, the last warps of the CTA)
//       decode tile bases, look up per-element gather offsets, issue cp.async
//       into the shared-memory ring and signal "full" mbarriers through
//       cp.async.mbarrier.arrive -- all address generation lives here;
//   consumer warps (P::THREADS, the first warps)
//       wait "full", run the policy's tile product (DMMA / FMA), release the
//       stage through the "empty" mbarrier, and store finished tiles straight
//       into the parent's index order.
//
// One CTA walks its work items (tile x k-split) as ONE stream of k-steps, so
// the ring keeps prefetching across tile boundaries and no warp ever waits at a
// CTA-wide barrier in steady state.
#pragma once
// (included inside namespace ctgb)

constexpr int KCHUNK = 128;          // k-steps whose base offsets are tabulated at once
constexpr int PRODUCER_THREADS = 128;  // one warpgroup (setmaxnreg granularity)

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(a) : "memory");
}
// arrive once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void mbar_arrive_cp_async(unsigned long long* bar) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
template <int ID, int COUNT>
__device__ __forceinline__ void named_sync() {
  asm volatile("bar.sync %0, %1;\n" ::"n"(ID), "n"(COUNT) : "memory");
}
template <int REGS>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS));
}
template <int REGS>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS));
}

template <class P, bool HAS>
struct BCacheOf {
  struct type {};
};
template <class P>
struct BCacheOf<P, true> {
  using type = typename P::BCache;
};

template <class P>
struct GettSmem {
  // slots of the per-element gather tables (elements the producers fetch one by one)
  static constexpr int NA = (P::A_GATHER + PRODUCER_THREADS - 1) / PRODUCER_THREADS;
  static constexpr int NB = (P::B_GATHER + PRODUCER_THREADS - 1) / PRODUCER_THREADS;
  // tile-info ring.  STAGES + 2 slots: the producer decodes tile j+TI right after
  // issuing tile j+TI-1, whose first stage needed the "empty" arrival of every
  // consumer warp for tile j+1 -- i.e. all of them are past the epilogue of tile
  // j, the last reader of slot j % TI.  (STAGES + 1 is one too few: the decode
  // runs BEFORE the producer waits on the stage it will fill.)
  // (the tcgen05 policy's epilogue group may lag one more tile behind: two TMEM accumulators)
  static constexpr int TI = P::STAGES + (P::IS_TC05 ? 4 : 2);
  template <typename T>
  static constexpr size_t bytes() {
    return sizeof(T) * ((size_t)P::STAGES * (P::A_ELEMS + P::B_ELEMS) + P::SCRATCH_ELEMS)  // ring + scratch
           + 8 * (size_t)(NA + NB) * PRODUCER_THREADS                                      // element deltas
           + 8 * (size_t)(P::MT + P::NT)                                                   // C offsets
           + 8 * (size_t)2 * KCHUNK                                                        // k-step bases
           + 8 * (size_t)4 * TI                                                            // tile bases
           + 8 * (size_t)2 * P::STAGES                                                     // mbarriers
           + 4 * (size_t)(NA + NB) * PRODUCER_THREADS                                      // element (r, kk)
           + 4 * (size_t)KCHUNK + 4 * (size_t)2 * TI                                       // valid counts
           + 64;
  }
};

// Consumer side of the tcgen05 policy, itself specialised:
//   warps 0-3  "MMA group": build A'lo for the stage, issue the 12 UMMAs (one elected
//              thread), commit them to the stage's "empty" barrier;
//   warps 4-7  "epilogue group": TMEM -> registers -> C for the previous tile while the
//              MMA group already works on the next one (two TMEM accumulators).
template <typename T, class P>
__device__ __forceinline__ void tc05_consumer(const int64_t* __restrict__ D, T* __restrict__ C, T* sA, T* sB,
                                              unsigned long long* bar_full, unsigned long long* bar_empty,
                                              unsigned long long* bar_tmem, unsigned* tmem_slot, const long long* ti_base,
                                              const int* ti_valid, const long long* offMC, const long long* offNC,
                                              unsigned nw, unsigned tiles_all, unsigned steps_k, unsigned steps_per_split,
                                              bool accumulate, bool atomic, const unsigned* metaA, bool bulk_a,
                                              bool exactA) {
  constexpr int MT = P::MT, NT = P::NT, STAGES = P::STAGES, NCONS = P::THREADS;
  constexpr int TI = GettSmem<P>::TI;
  constexpr int GROUP = 128;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long* tmem_full = bar_tmem;       // [2] count 1 (tcgen05.commit)
  unsigned long long* tmem_empty = bar_tmem + 2;  // [2] count 4 (epilogue warps)
  if (warp == 0) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(tmem_slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(a), "r"(2 * P::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  named_sync<2, NCONS>();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned taddr = *tmem_slot;
  // InstrDescriptor: D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2, K-major A/B, N>>3 [17,23), M>>4 [24,29)
  constexpr unsigned idesc =
      (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(P::TMEM_COLS >> 3) << 17) | ((unsigned)(MT >> 4) << 24);

  if (warp < 4) {
    // ------------------------------------------------------------ MMA group
    // scatter map of the bulk-copy mode: the same for every stage, kept in registers
    // (tables of launches with a blocked dim hold (r, kk) pairs instead of the index)
    constexpr int NSCAT = (MT * P::KT) / GROUP;
    unsigned ureg[NSCAT];
    if (bulk_a) {
#pragma unroll
      for (int i = 0; i < NSCAT; ++i) {
        const unsigned meta = metaA[tid + i * GROUP];
        ureg[i] = exactA ? meta : (unsigned)P::idxA((int)(meta & 0xFFFFu), (int)(meta >> 16));
      }
    }
    unsigned g = 0;
    for (unsigned j = 0; j < nw; ++j) {
      const unsigned w = blockIdx.x + j * gridDim.x;
      const unsigned ks = w / tiles_all;
      const unsigned k0 = ks * steps_per_split, k1 = min(steps_k, k0 + steps_per_split);
      const unsigned buf = j & 1;
      mbar_wait(&tmem_empty[buf], ((j >> 1) & 1) ^ 1);  // epilogue of tile j-2 has drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      for (unsigned step = k0; step < k1; ++step, ++g) {
        const int st = (int)(g % STAGES);
        mbar_wait(&bar_full[st], (g / STAGES) & 1);
        // A'lo = A' - trunc_tf32(A') for the whole stage (same tile order)
        float4* ah = reinterpret_cast<float4*>(sA + st * P::A_ELEMS);
        float4* al = ah + (MT * P::KT) / 2;
        if (bulk_a) {
          // staging holds the tile in A-memory order; ureg[i] is the UMMA index of element tid + i*GROUP
          const float2* stg = reinterpret_cast<const float2*>(al + (MT * P::KT) / 2);
          float2* hi2 = reinterpret_cast<float2*>(ah);
          float2* lo2 = reinterpret_cast<float2*>(al);
#pragma unroll
          for (int i = 0; i < NSCAT; ++i) {
            const float2 v = stg[tid + i * GROUP];
            hi2[ureg[i]] = v;
            lo2[ureg[i]] = make_float2(v.x - trunc_tf32(v.x), v.y - trunc_tf32(v.y));
          }
        } else {
#pragma unroll
          for (int i = tid; i < (MT * P::KT) / 2; i += GROUP) {
            const float4 v = ah[i];
            al[i] = make_float4(v.x - trunc_tf32(v.x), v.y - trunc_tf32(v.y), v.z - trunc_tf32(v.z), v.w - trunc_tf32(v.w));
          }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> tensor core
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        named_sync<3, GROUP>();
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        if (tid == 0) {
          const unsigned a_hi = (unsigned)__cvta_generic_to_shared(ah);
          const unsigned a_lo = (unsigned)__cvta_generic_to_shared(al);
          const unsigned b_hi = (unsigned)__cvta_generic_to_shared(sB + st * P::B_ELEMS);
          const unsigned b_lo = b_hi + P::TILE_FLOATS * 4;
          const unsigned dcol = taddr + buf * P::TMEM_COLS;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const unsigned a0 = pass == 0 ? a_lo : a_hi, b0 = pass == 1 ? b_lo : b_hi;  // lo*hi, hi*lo, hi*hi
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              // one UMMA eats K = 8 floats = 2 chunks; chunk stride (LBO) = rows*16 B, 8-row group stride (SBO) = 128 B
              const uint64_t da = umma_desc_kmajor(a0 + q * 2 * MT * 16, MT * 16, 128);
              const uint64_t db = umma_desc_kmajor(b0 + q * 2 * (2 * NT) * 16, (2 * NT) * 16, 128);
              const unsigned acc = (step != k0 || pass != 0 || q != 0) ? 1u : 0u;
              asm volatile(
                  "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dcol),
                  "l"(da), "l"(db), "r"(idesc), "r"(acc)
                  : "memory");
            }
          }
          // the stage may be refilled once these UMMAs have read it
          const unsigned mb = (unsigned)__cvta_generic_to_shared(&bar_empty[st]);
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mb)
                       : "memory");
          if (step + 1 == k1) {
            const unsigned mf = (unsigned)__cvta_generic_to_shared(&tmem_full[buf]);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mf)
                         : "memory");
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue group
    const bool quad_ok = (D[W_FLAGS] & 16) != 0 && !accumulate && !atomic;
    const int quad = warp & 3;  // TMEM lane quadrant of this warp
    const int r = quad * 32 + lane;
    for (unsigned j = 0; j < nw; ++j) {
      const unsigned buf = j & 1;
      mbar_wait(&tmem_full[buf], (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const int slot = (int)(j % TI);
      const long long baseC = ti_base[slot * 4 + 2];
      const int m_valid = ti_valid[slot * 2 + 0], n_valid = ti_valid[slot * 2 + 1];
      T* crow = C + baseC + offMC[r];
      // 32 fp32 columns (16 complex) per tcgen05.ld: one TMEM round trip per 128 bytes of a row
#pragma unroll 1
      for (int col = 0; col < 2 * NT; col += 32) {  // fp32 column; complex column = col / 2
        unsigned v[32];
        const unsigned ta = taddr + buf * P::TMEM_COLS + ((unsigned)(quad * 32) << 16) + (unsigned)col;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (r < m_valid) {
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {  // groups of 4 complex columns
            const int c0 = (col >> 1) + s4 * 4;
            const unsigned* w = v + s4 * 8;
            if (c0 >= n_valid) continue;
            if (quad_ok) {
              const unsigned long long q0 = ((unsigned long long)w[1] << 32) | w[0], q1 = ((unsigned long long)w[3] << 32) | w[2];
              const unsigned long long q2 = ((unsigned long long)w[5] << 32) | w[4], q3 = ((unsigned long long)w[7] << 32) | w[6];
              asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};\n" ::"l"(crow + offNC[c0]), "l"(q0), "l"(q1), "l"(q2),
                           "l"(q3)
                           : "memory");
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (c0 + e < n_valid) {
                  T* p = crow + offNC[c0 + e];
                  const T val = make_float2(__uint_as_float(w[2 * e]), __uint_as_float(w[2 * e + 1]));
                  if (atomic) {
                    atomic_add_of(p, val);
                  } else if (accumulate) {
                    *p = add_of(*p, val);
                  } else {
                    *p = val;
                  }
                }
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
  }
  named_sync<2, NCONS>();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(2 * P::TMEM_COLS)
                 : "memory");
}

template <typename T, class P>
__global__ void __launch_bounds__(P::THREADS + PRODUCER_THREADS, P::MIN_BLOCKS)
gett_kernel(const int64_t* __restrict__ D, const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C) {
  constexpr int MT = P::MT, NT = P::NT, STAGES = P::STAGES;
  constexpr int NCONS = P::THREADS, NPROD = PRODUCER_THREADS, NTHR = NCONS + NPROD;
  constexpr int NA = GettSmem<P>::NA, NB = GettSmem<P>::NB, TI = GettSmem<P>::TI;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sA = reinterpret_cast<T*>(smem_raw);
  T* sB = sA + STAGES * P::A_ELEMS;
  T* scratch = sB + STAGES * P::B_ELEMS;
  long long* gA = reinterpret_cast<long long*>(scratch + P::SCRATCH_ELEMS);
  long long* gB = gA + NA * NPROD;
  long long* offMC = gB + NB * NPROD;
  long long* offNC = offMC + MT;
  long long* kbA = offNC + NT;
  long long* kbB = kbA + KCHUNK;
  long long* ti_base = kbB + KCHUNK;  // [TI][4]: A, B, C, B'-tile index
  unsigned long long* bar_full = reinterpret_cast<unsigned long long*>(ti_base + 4 * TI);
  unsigned long long* bar_empty = bar_full + STAGES;
  unsigned* metaA = reinterpret_cast<unsigned*>(bar_empty + STAGES);
  unsigned* metaB = metaA + NA * NPROD;
  int* kval = reinterpret_cast<int*>(metaB + NB * NPROD);
  int* ti_valid = kval + KCHUNK;  // [TI][2]: m_valid, n_valid

  __shared__ __align__(8) unsigned long long bar_tile_storage[4];  // tc05: tmem_full[2], tmem_empty[2]
  __shared__ unsigned tmem_slot_storage;
  [[maybe_unused]] unsigned long long* bar_tile = bar_tile_storage;
  [[maybe_unused]] unsigned* tmem_slot = &tmem_slot_storage;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const bool is_producer = tid >= NCONS;
  const int ptid = tid - NCONS;  // producer-local thread id

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
  // bit2: every tile-grid extent is a power of two -> digits by shift/mask, no idiv
  const bool g_pow2 = (D[W_FLAGS] & 4) != 0;
  auto digit_of = [&](unsigned idx, unsigned div, unsigned ext) -> unsigned {
    return g_pow2 ? ((idx >> (31 - __clz(div))) & (ext - 1)) : ((idx / div) % ext);
  };
  // tcgen05 policy: A tile = contiguous runs of run_a elements fetched by TMA bulk copies (flags bit6)
  [[maybe_unused]] const bool bulk_a =
      P::IS_TC05 && (D[W_FLAGS] & 64) != 0 && (reinterpret_cast<unsigned long long>(A) & 15ull) == 0;
  [[maybe_unused]] const unsigned run_a = (unsigned)D[W_RUNA];
  // no blocked (partial) dim touches the operand: every tabulated element is always valid
  const bool exactA = pgm < 0 && pgk < 0, exactB = pgn < 0 && pgk < 0;

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

  // ---- one-time tables (all threads) ----
  // zero the operand ring: rows/cols/k beyond the actual tile are never loaded
  for (int i = tid; i < STAGES * (P::A_ELEMS + P::B_ELEMS); i += NTHR) sA[i] = zero_of<T>();
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bar_full[s], NPROD);
      mbar_init(&bar_empty[s], P::IS_TC05 ? 1 : NCONS / 32);  // tc05: one tcgen05.commit per stage
    }
    if constexpr (P::IS_TC05) {
      mbar_init(&bar_tile[0], 1);
      mbar_init(&bar_tile[1], 1);
      mbar_init(&bar_tile[2], 4);
      mbar_init(&bar_tile[3], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (is_producer) {
    // per-slot element tables, enumerated in operand-memory order for coalescing
    const int n_lda = (int)D[W_NLDA], n_ldb = (int)D[W_NLDB];
    for (int i = 0; i < NA; ++i) {
      unsigned e = ptid + i * NPROD;
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
      gA[i * NPROD + ptid] = g;
      // exact tiles (no blocked dim): store the shared-memory index directly
      metaA[i * NPROD + ptid] = exactA ? (kk == 0xFFFFu ? 0xFFFFFFFFu : (unsigned)P::idxA((int)r, (int)kk))
                                       : (r | (kk << 16));
    }
    for (int i = 0; i < NB; ++i) {
      unsigned e = ptid + i * NPROD;
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
      gB[i * NPROD + ptid] = g;
      metaB[i * NPROD + ptid] = exactB ? (kk == 0xFFFFu ? 0xFFFFFFFFu : (unsigned)P::idxB((int)c, (int)kk))
                                       : (c | (kk << 16));
    }
    if (ktab) {
      for (unsigned s = ptid; s < steps_k; s += NPROD) {
        long long a, b;
        int kv;
        kstep_bases(s, a, b, kv);
        kbA[s] = a;
        kbB[s] = b;
        kval[s] = kv;
      }
    }
  } else {
    // local C offsets of every tile row / column (used by the consumers' epilogue)
    for (int r = tid; r < MT; r += NCONS) {
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
    for (int c = tid; c < NT; c += NCONS) {
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
  }
  __syncthreads();

  // the host guarantees total_work < 2^31 (lowering.py)
  const unsigned tiles_all = tiles_m * tiles_n * tiles_b;
  const unsigned total_work = tiles_all * splitk;
  const unsigned steps_per_split = (steps_k + splitk - 1) / splitk;
  const unsigned nw = blockIdx.x < total_work ? (total_work - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;

  auto work_krange = [&](unsigned j, unsigned& k0, unsigned& k1) {
    const unsigned w = blockIdx.x + j * gridDim.x;
    const unsigned ks = w / tiles_all;
    k0 = ks * steps_per_split;
    k1 = min(steps_k, k0 + steps_per_split);
  };

  if (is_producer) {
    // ===================================================== PRODUCER WARPS
    if constexpr (P::CONSUMER_REGS > 0) reg_dealloc<P::PRODUCER_REGS>();
    unsigned ktab_base = 0;
    unsigned g = 0;  // global stage counter
    for (unsigned j = 0; j < nw; ++j) {
      unsigned k0, k1;
      work_krange(j, k0, k1);
      // ---- grid-base offsets of work item j -> tile-info slot j % TI (one lane
      // per grid dim; n fastest so neighbouring CTAs share A tiles in L2)
      const int slot = (int)(j % TI);
      if (ptid < 32) {
        unsigned t = blockIdx.x + j * gridDim.x;
        if (splitk > 1) t %= tiles_all;
        unsigned in_, im_, ib_;
        if (g_pow2) {
          in_ = t & (tiles_n - 1);
          t >>= 31 - __clz(tiles_n);
          im_ = t & (tiles_m - 1);
          ib_ = t >> (31 - __clz(tiles_m));
        } else {
          in_ = t % tiles_n;
          t /= tiles_n;
          im_ = t % tiles_m;
          ib_ = t / tiles_m;
        }
        long long a = 0, b = 0, c = 0;
        int vm = 0, vn = 0;
        for (int q = lane; q < n_gm; q += 32) {
          const int64_t* G = D + OFF_GM + q * 4;
          unsigned dig = digit_of(im_, (unsigned)G[1], (unsigned)G[0]);
          a += (long long)dig * G[2];
          c += (long long)dig * G[3];
          if (q == pgm)
            vm = (int)min((long long)D[W_MTEXT], (long long)D[W_MFULL] - (long long)dig * (long long)D[W_MTEXT]) *
                 (int)D[W_MW];
        }
        for (int q = lane; q < n_gn; q += 32) {
          const int64_t* G = D + OFF_GN + q * 4;
          unsigned dig = digit_of(in_, (unsigned)G[1], (unsigned)G[0]);
          b += (long long)dig * G[2];
          c += (long long)dig * G[3];
          if (q == pgn)
            vn = (int)min((long long)D[W_NTEXT], (long long)D[W_NFULL] - (long long)dig * (long long)D[W_NTEXT]) *
                 (int)D[W_NW];
        }
        for (int q = lane; q < n_gb; q += 32) {
          const int64_t* G = D + OFF_GB + q * 5;
          unsigned dig = digit_of(ib_, (unsigned)G[1], (unsigned)G[0]);
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
          ti_base[slot * 4 + 0] = a;
          ti_base[slot * 4 + 1] = b;
          ti_base[slot * 4 + 2] = c;
          ti_base[slot * 4 + 3] = (long long)ib_ * tiles_n + in_;
          ti_valid[slot * 2 + 0] = pgm < 0 ? MTa : vm;
          ti_valid[slot * 2 + 1] = pgn < 0 ? NTa : vn;
          __threadfence_block();
        }
      }
      named_sync<1, NPROD>();
      const long long tA = ti_base[slot * 4 + 0], tB = ti_base[slot * 4 + 1];
      [[maybe_unused]] const long long tBp = ti_base[slot * 4 + 3];
      const unsigned m_valid = (unsigned)ti_valid[slot * 2 + 0], n_valid = (unsigned)ti_valid[slot * 2 + 1];

      for (unsigned step = k0; step < k1; ++step, ++g) {
        if (!ktab && (step == k0 || step >= ktab_base + KCHUNK)) {
          // long contracted ranges: the k table is a window of KCHUNK steps
          named_sync<1, NPROD>();  // every producer is done reading the old window
          ktab_base = step;
          for (unsigned s = ptid; s < (unsigned)KCHUNK && step + s < k1; s += NPROD) {
            long long a, b;
            int kv;
            kstep_bases(step + s, a, b, kv);
            kbA[s] = a;
            kbB[s] = b;
            kval[s] = kv;
          }
          named_sync<1, NPROD>();
        }
        const int st = (int)(g % STAGES);
        mbar_wait(&bar_empty[st], ((g / STAGES) & 1) ^ 1);
        T* dA = sA + st * P::A_ELEMS;
        T* dB = sB + st * P::B_ELEMS;
        const unsigned ti = step - ktab_base;
        const T* srcA = A + tA + kbA[ti];
        const T* srcB = B + tB + kbB[ti];
        const unsigned kv = (unsigned)kval[ti];
        bool arrived = false;
        if constexpr (P::IS_TC05) {
          if (bulk_a) {
            // the A tile is a set of long contiguous runs: TMA bulk copies into the staging
            // area (memory order); the MMA group scatters them into the UMMA layout
            const unsigned bar = (unsigned)__cvta_generic_to_shared(&bar_full[st]);
            const unsigned nruns = (unsigned)(MTa * KTa) / run_a;
            const unsigned mine = ptid < nruns ? (nruns - ptid + NPROD - 1) / NPROD : 0u;
            const unsigned bytes = mine * run_a * (unsigned)sizeof(T) + (ptid == 0 ? (unsigned)P::PAIR_BYTES : 0u);
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(bar),
                         "r"(bytes)
                         : "memory");
            T* stg = dA + 2 * P::MT * P::KT;
            for (unsigned q = ptid; q < nruns; q += NPROD) {
              const unsigned dst = (unsigned)__cvta_generic_to_shared(stg + q * run_a);
              asm volatile(
                  "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                  "l"(srcA + gA[q * run_a]), "r"(run_a * (unsigned)sizeof(T)), "r"(bar)
                  : "memory");
            }
            arrived = true;
          }
        }
        if (arrived) {
        } else if (exactA) {
#pragma unroll
          for (int i = 0; i < NA; ++i) {
            const unsigned meta = metaA[i * NPROD + ptid];
            if (meta != 0xFFFFFFFFu) cp_async_zfill<sizeof(T)>(dA + meta, srcA + gA[i * NPROD + ptid], true);
          }
        } else {
#pragma unroll
          for (int i = 0; i < NA; ++i) {
            const unsigned meta = metaA[i * NPROD + ptid];
            const unsigned r = meta & 0xFFFFu, kk = meta >> 16;
            if (kk != 0xFFFFu) {
              const bool ok = (r < m_valid) && (kk < kv);
              cp_async_zfill<sizeof(T)>(dA + P::idxA(r, kk), ok ? (srcA + gA[i * NPROD + ptid]) : A, ok);
            }
          }
        }
        if constexpr (P::IS_TC05) {
          // the stage's B'hi + B'lo tiles are contiguous in the prepared buffer: one TMA bulk copy
          if (ptid == 0) {
            const unsigned bar = (unsigned)__cvta_generic_to_shared(&bar_full[st]);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(dB);
            const char* src = reinterpret_cast<const char*>(B) +
                              ((unsigned long long)tBp * steps_k + step) * (unsigned long long)P::PAIR_BYTES;
            if (!arrived)  // (bulk mode already announced these bytes with its arrival)
              asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(P::PAIR_BYTES)
                           : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                "l"(src), "r"(P::PAIR_BYTES), "r"(bar)
                : "memory");
          }
        } else if (exactB) {
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            const unsigned meta = metaB[i * NPROD + ptid];
            if (meta != 0xFFFFFFFFu) cp_async_zfill<sizeof(T)>(dB + meta, srcB + gB[i * NPROD + ptid], true);
          }
        } else {
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            const unsigned meta = metaB[i * NPROD + ptid];
            const unsigned c = meta & 0xFFFFu, kk = meta >> 16;
            if (kk != 0xFFFFu) {
              const bool ok = (c < n_valid) && (kk < kv);
              cp_async_zfill<sizeof(T)>(dB + P::idxB(c, kk), ok ? (srcB + gB[i * NPROD + ptid]) : B, ok);
            }
          }
        }
        if (!arrived) mbar_arrive_cp_async(&bar_full[st]);
      }
    }
    cp_async_commit();
    cp_async_wait<0>();  // do not exit with copies in flight
  } else {
    // ===================================================== CONSUMER WARPS
    if constexpr (P::IS_TC05) {
      tc05_consumer<T, P>(D, C, sA, sB, bar_full, bar_empty, bar_tile, tmem_slot, ti_base, ti_valid, offMC, offNC, nw,
                          tiles_all, steps_k, steps_per_split, accumulate, atomic, metaA, bulk_a, exactA);
      return;
    }
    if constexpr (P::CONSUMER_REGS > 0) reg_alloc<P::CONSUMER_REGS>();
    // valid k of a step: only a blocked (partial) k dim can shorten it
    const unsigned pk_div = pgk >= 0 ? (unsigned)D[OFF_GK + pgk * 4 + 1] : 1u;
    const unsigned pk_ext = pgk >= 0 ? (unsigned)D[OFF_GK + pgk * 4 + 0] : 1u;
    const int ktext = (int)D[W_KTEXT], kfull = (int)D[W_KFULL], kw = (int)D[W_KW];
    typename P::Acc acc;
    P::clear(acc);
    unsigned g = 0;
    // the small operand's tile is identical for every work item of this launch
    [[maybe_unused]] const bool b_invariant = steps_k == 1 && n_gn == 0 && n_gb == 0;
    [[maybe_unused]] bool b_loaded = false;
    [[maybe_unused]] typename BCacheOf<P, P::HAS_BCACHE>::type bcache;
    for (unsigned j = 0; j < nw; ++j) {
      unsigned k0, k1;
      work_krange(j, k0, k1);
      for (unsigned step = k0; step < k1; ++step, ++g) {
        const int st = (int)(g % STAGES);
        int kv = KTa;
        if (pgk >= 0) {
          const int dig = (int)((step / pk_div) % pk_ext);
          kv = min(ktext, kfull - dig * ktext) * kw;
        }
        mbar_wait(&bar_full[st], (g / STAGES) & 1);
        if constexpr (P::HAS_BCACHE) {
          if (b_invariant) {
            if (!b_loaded) {
              P::load_b(sB + st * P::B_ELEMS, bcache);
              b_loaded = true;
            }
            P::compute_cached(sA + st * P::A_ELEMS, bcache, acc, kv, NTa);
          } else {
            P::compute(sA + st * P::A_ELEMS, sB + st * P::B_ELEMS, acc, kv, NTa);
          }
        } else
          P::compute(sA + st * P::A_ELEMS, sB + st * P::B_ELEMS, acc, kv, NTa);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[st]);
      }
      // ---- epilogue of tile j: store in the parent's index order (strided C)
      const int slot = (int)(j % TI);
      const long long baseC = ti_base[slot * 4 + 2];
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
          pair_ok, n_valid);
      P::clear(acc);
    }
  }
}

