// bf16 GEMM on the 5th-generation tensor cores (sm_100a): tcgen05.mma with fp32 accumulators in TMEM,
// operands staged in shared memory by TMA (128-byte swizzle), warp-specialised:
//   warp 0   : TMA producer (one elected lane)
//   warp 1   : TMEM allocator + MMA issuer (one elected lane)
//   warps 2-5: epilogue (tcgen05.ld -> bias/ReLU/padding columns -> shared-memory transpose -> coalesced global
//              stores with the dgrad ReLU mask / split-K reduction applied on the way out)
//
// Two kernels share the descriptors and the epilogue:
//   gemm_tc2_kernel : the workhorse.  A CTA PAIR (cluster of 2, tcgen05 cta_group::2) owns a 256 x BN output
//                     tile: each CTA stages its own 128 rows of A and HALF of the B tile, so every operand byte
//                     is fetched from L2 once per pair (these GEMMs are L2->SM bandwidth bound, not FLOP bound:
//                     M = batch = 4096 with N, K of 512..4096).  Persistent over (tile, k-split) work units with
//                     a double-buffered TMEM accumulator: the epilogue of unit i overlaps the main loop of i+1.
//   gemm_tc_kernel  : one CTA, one 128 x BN tile; used for narrow outputs (N < 128: the latent heads) and
//                     matrices with fewer than 256 rows.
//
// C[M,N] = epilogue(op(A) . op(B)) with both operand majors supported through the UMMA descriptors, so
// the forward (A K-major, W MN-major), dgrad (both K-major) and wgrad (both MN-major) GEMMs of a dense
// layer all read the SAME row-major activation / weight buffers - no transposed copies in HBM.
//
// Descriptor encodings follow the PTX ISA "tcgen05 shared memory descriptor" / "instruction descriptor"
// tables (bit layout as in cute/arch/mma_sm100_desc.hpp of the vendored CUTLASS headers).
#include "tc_device.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// single-CTA kernel: one 128 x BN tile per CTA
// ------------------------------------------------------------------------------------------------
template <int A_MN, int B_MN, int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, int M, int N, int K, int kb_per_split, int stages,
               const EpiParams ep) {
  constexpr int B_TILE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 1];
  __shared__ uint32_t tmem_slot;

  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B needs 1024-byte alignment
  const uint32_t epi_stage = tiles + stages * STAGE_BYTES;          // 1024-byte aligned (STAGE_BYTES % 1024 == 0)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkb_total = (K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(nkb_total, kb_begin + kb_per_split);
  const int nkb = kb_end - kb_begin;

  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[kMaxStages + s]); };
  const uint32_t tmem_full_bar = smem_u32(&bars[2 * kMaxStages]);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();                  // the prologue above overlapped the previous kernel; operands / outputs are touched below
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % stages;
        const uint32_t ph = (uint32_t)(i / stages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), STAGE_BYTES);
        const uint32_t sa = tiles + s * STAGE_BYTES, sb = sa + A_TILE_BYTES;
        const int k0 = (kb_begin + i) * BK;
        if (A_MN == 0) {
          tma_load_2d(sa, &tmA, full_bar(s), k0, m0);                       // box {64 k, 128 m}
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, &tmA, full_bar(s), m0 + 64 * j, k0);   // box {64 m, 64 k}
        }
        if (B_MN == 0) {
          tma_load_2d(sb, &tmB, full_bar(s), k0, n0);                       // box {64 k, BN n}
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, full_bar(s), n0 + 64 * j, k0);   // box {64 n, 64 k}
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<A_MN, B_MN>(BM, BN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % stages;
        const uint32_t ph = (uint32_t)(i / stages) & 1u;
        mbar_wait(full_bar(s), ph);
        tcgen05_fence_after();
        const uint32_t sa = tiles + s * STAGE_BYTES, sb = sa + A_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major: 8-row groups are 1024 B apart (SBO); advance 32 B per UMMA_K inside the swizzle row.
          // MN-major: 64-element MN groups are 8192 B apart (LBO), 8-k groups 1024 B apart (SBO); advance 16 k-rows.
          const uint64_t ad = A_MN ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
          const uint64_t bd = B_MN ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
          tcgen05_mma_f16<1>(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        tcgen05_commit(empty_bar(s));        // smem slot reusable once these MMAs retire
      }
      tcgen05_commit(tmem_full_bar);         // accumulator complete
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    if (nkb > 0) {
      epilogue_warp<false>(tmem_base + ((uint32_t)(q * 32) << 16), 0, BN, m0 + q * 32, n0, M, N, &tmC, ep, blockIdx.z == 0,
                    epi_stage + q * kEpiStageBytes, lane, 0u, nullptr, tmem_full_bar, 0u);
    } else {
      mbar_wait(tmem_full_bar, 0);
      tcgen05_fence_after();
    }
  }
  // ===================== teardown =====================
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair kernel: persistent, 256 x BN tile per pair (cta_group::2), double-buffered TMEM accumulator
// ------------------------------------------------------------------------------------------------
template <int BN>
struct Pair {
  static constexpr int BNH = BN / 2;                          // B columns staged by each CTA
  static constexpr int B_TILE_BYTES = BNH * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES = BN == 256 ? 6 : 8;            // 192 KiB of operand stages per CTA
  static constexpr int TMEM_COLS = 2 * BN;                    // two accumulator buffers
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + kEpiBytes2 + 1024;
};

template <int A_MN, int B_MN, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, int M, int N, int K, int kb_per_split, int nsplit,
                const EpiParams ep) {
  using P = Pair<BN>;
  constexpr int STAGES = P::STAGES;
  constexpr int STAGE_BYTES = P::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t tmem_slot;

  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_stage = tiles + STAGES * STAGE_BYTES;          // 1024-byte aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                    // 0 = leader (issues the MMAs)
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int tiles_m = (M + 2 * BM - 1) / (2 * BM), tiles_n = (N + BN - 1) / BN;
  const int n_tiles = tiles_m * tiles_n;
  const int n_units = n_tiles * nsplit;
  const int nkb_total = (K + BK - 1) / BK;

  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[STAGES + s]); };
  auto tfull_bar = [&](int a) { return smem_u32(&bars[2 * STAGES + a]); };
  auto tempty_bar = [&](int a) { return smem_u32(&bars[2 * STAGES + 2 + a]); };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);          // leader's producer: arrive.expect_tx for both CTAs' bytes
      mbar_init(empty_bar(s), 1);         // multicast tcgen05.commit of the leader
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);         // multicast tcgen05.commit of the leader
      mbar_init(tempty_bar(a), 16);       // 8 epilogue warps of each CTA (used in the leader only)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(P::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();                     // barriers of both CTAs initialised, TMEM allocated
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();                             // the prologue above overlapped the previous kernel's tail
  pdl_launch_dependents();

  // work unit u -> (split, tile); tiles walk M fastest so that concurrently running pairs share B columns
  auto decode = [&](int u, int& m_blk, int& n_blk, int& kb0, int& kb1) {
    const int split = u / n_tiles, t = u - split * n_tiles;
    n_blk = t / tiles_m;
    m_blk = t - n_blk * tiles_m;
    kb0 = split * kb_per_split;
    kb1 = min(nkb_total, kb0 + kb_per_split);
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      const uint32_t lead_full0 = mapa_u32(full_bar(0), 0);
      uint32_t it = 0;
      for (int u = pair_id; u < n_units; u += n_pairs) {
        int m_blk, n_blk, kb0, kb1;
        decode(u, m_blk, n_blk, kb0, kb1);
        const int m0 = m_blk * 2 * BM + (int)rank * BM;       // this CTA's 128 rows of A
        const int nb0 = n_blk * BN + (int)rank * P::BNH;      // this CTA's half of the B tile
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (rank == 0) mbar_expect_tx(full_bar(s), 2 * STAGE_BYTES);
          const uint32_t fb = lead_full0 + (uint32_t)(s * 8);
          const uint32_t sa = tiles + s * STAGE_BYTES, sb = sa + A_TILE_BYTES;
          const int k0 = kb * BK;
          if (A_MN == 0) {
            tma_load_2d_pair(sa, &tmA, fb, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d_pair(sa + j * 8192, &tmA, fb, m0 + 64 * j, k0);
          }
          if (B_MN == 0) {
            tma_load_2d_pair(sb, &tmB, fb, k0, nb0);
          } else {
#pragma unroll
            for (int j = 0; j < P::BNH / 64; ++j) tma_load_2d_pair(sb + j * 8192, &tmB, fb, nb0 + 64 * j, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc<A_MN, B_MN>(2 * BM, BN);
      uint32_t it = 0, ui = 0;
      for (int u = pair_id; u < n_units; u += n_pairs, ++ui) {
        int m_blk, n_blk, kb0, kb1;
        decode(u, m_blk, n_blk, kb0, kb1);
        const uint32_t as = ui & 1u, aph = (ui >> 1) & 1u;
        mbar_wait(tempty_bar(as), aph ^ 1u);                  // both CTAs' epilogues drained this buffer
        tcgen05_fence_after();
        const uint32_t acc = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(full_bar(s), ph);
          tcgen05_fence_after();
          const uint32_t sa = tiles + s * STAGE_BYTES, sb = sa + A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = A_MN ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
            tcgen05_mma_f16<2>(acc, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tcgen05_commit_pair(empty_bar(s), 3);               // frees this stage in both CTAs
        }
        tcgen05_commit_pair(tfull_bar(as), 3);                // accumulator complete, both CTAs' epilogues
      }
    }
  } else {
    // ===================== epilogue (both CTAs; each drains its own 128 accumulator rows) =====================
    // warps 2..9: lane quarter = warp % 4 (the TMEM access rule), column half = (warp - 2) / 4
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t lead_tempty0 = mapa_u32(tempty_bar(0), 0);
    const uint32_t my_stage = epi_stage + (uint32_t)(warp - 2) * kEpiStageBytes;
    uint32_t ui = 0;
    for (int u = pair_id; u < n_units; u += n_pairs, ++ui) {
      int m_blk, n_blk, kb0, kb1;
      decode(u, m_blk, n_blk, kb0, kb1);
      const uint32_t as = ui & 1u, aph = (ui >> 1) & 1u;
      epilogue_warp<(A_MN == 0 && B_MN == 1)>(tmem_base + as * BN + ((uint32_t)(q * 32) << 16), half * (BN / 2), (half + 1) * (BN / 2),
                    m_blk * 2 * BM + (int)rank * BM + q * 32, n_blk * BN, M, N, &tmC, ep, kb0 == 0, my_stage, lane,
                    lead_tempty0 + as * 8, nullptr, tfull_bar(as), aph);
    }
  }
  // ===================== teardown =====================
  tcgen05_fence_before();
  cluster_sync_all();                     // no CTA leaves while its peer may still read its smem / signal its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps and launch
// ------------------------------------------------------------------------------------------------
// operand tensor maps; bn_box = number of B columns one CTA stages per k-block
template <int A_MN, int B_MN>
int make_tmaps(dmvae_ctx* ctx, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int out_dtype,
               int M, int N, int K, int bn_box, CUtensorMap* ta, CUtensorMap* tb, CUtensorMap* tc) {
  int rc;
  // C: [M, N] row-major; the epilogue stores 32-row slabs of 128 bytes (64 bf16 / 32 fp32 columns)
  if (out_dtype == DMVAE_BF16) rc = get_tmap(ctx, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, 64, 32, tc, 2);
  else rc = get_tmap(ctx, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, 32, 32, tc, 4);
  if (rc) return rc;
  // A: K-major -> stored [M, K] (inner K); MN-major -> stored [K, M] (inner M)
  if (A_MN == 0) rc = get_tmap(ctx, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM, ta);
  else rc = get_tmap(ctx, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK, ta);
  if (rc) return rc;
  // B: K-major -> stored [N, K] (inner K); MN-major -> stored [K, N] (inner N)
  if (B_MN == 0) rc = get_tmap(ctx, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, (uint32_t)bn_box, tb);
  else rc = get_tmap(ctx, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK, tb);
  return rc;
}

template <int A_MN, int B_MN, int BN>
int launch_tc(dmvae_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int M, int N, int K,
              int split, const EpiParams& ep, cudaStream_t st) {
  constexpr int STAGE_BYTES = A_TILE_BYTES + BN * BK * 2;
  const int nkb = (K + BK - 1) / BK;
  int kps = (nkb + split - 1) / split;
  if (kps < 1) kps = 1;
  split = (nkb + kps - 1) / kps;      // every split owns at least one k-block
  if (split < 1) split = 1;
  // short K: two CTAs per SM (one tile's epilogue overlaps the other's main loop).  Long K with few tiles (the latent
  // heads: 32 CTAs streaming 2048-deep operands): the CTA is latency bound on its TMA loads, so take the whole SM's
  // shared memory for stages - bytes in flight / latency is its bandwidth
  int stages = BN == 64 ? 4 : 3;
  if (kps >= 12) stages = BN == 64 ? 8 : 6;
  const size_t smem = (size_t)stages * STAGE_BYTES + kEpiBytes + 1024;
  auto kern = gemm_tc_kernel<A_MN, B_MN, BN>;
  static size_t smem_opted = 0;       // per instantiation
  if (smem > smem_opted) {
    DMVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_opted = smem;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, split);
  dmvae_launch(kern, grid, dim3(kThreads), smem, st, true, ta, tb, tc, M, N, K, kps, stages, ep);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

template <int A_MN, int B_MN, int BN>
int launch_tc2(dmvae_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int M, int N, int K,
               int split, const EpiParams& ep, cudaStream_t st) {
  using P = Pair<BN>;
  const int nkb = (K + BK - 1) / BK;
  int kps = (nkb + split - 1) / split;
  if (kps < 1) kps = 1;
  split = (nkb + kps - 1) / kps;
  if (split < 1) split = 1;
  auto kern = gemm_tc2_kernel<A_MN, B_MN, BN>;
  static bool opted = false;          // per instantiation
  if (!opted) {
    DMVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM));
    opted = true;
  }
  const int units = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN) * split;
  const int pairs = std::max(1, std::min(units, ctx->sm_count / 2));
  dmvae_launch(kern, dim3(2 * pairs), dim3(kThreads2), P::SMEM, st, true, ta, tb, tc, M, N, K, kps, split, ep);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

// DMVAE_GEMM_PAIR=0 forces the single-CTA kernel everywhere (A/B measurements)
bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DMVAE_GEMM_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int A_MN, int B_MN>
int dispatch_bn(dmvae_ctx* ctx, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int M, int N, int K,
                int split, const EpiParams& ep, cudaStream_t st) {
  CUtensorMap ta, tb, tc;
  int rc;
  const int pairs_avail = ctx->sm_count / 2;
  if (pair_enabled() && M >= 2 * BM && N >= 128) {
    // 256-wide tiles halve the L2 traffic per FLOP; take them when they still give most pairs a work unit
    const int tm = (M + 2 * BM - 1) / (2 * BM);
    const long long units256 = (long long)tm * ((N + 255) / 256) * split;
    const int bn = (N >= 256 && units256 * 4 >= (long long)pairs_avail * 3) ? 256 : 128;
    rc = make_tmaps<A_MN, B_MN>(ctx, A, lda, B, ldb, C, ldc, ep.out_dtype, M, N, K, bn / 2, &ta, &tb, &tc);
    if (rc) return rc;
    if (bn == 256) return launch_tc2<A_MN, B_MN, 256>(ctx, ta, tb, tc, M, N, K, split, ep, st);
    return launch_tc2<A_MN, B_MN, 128>(ctx, ta, tb, tc, M, N, K, split, ep, st);
  }
  // single-CTA tiles: 128-wide unless that would leave SMs idle
  const int mt = (M + BM - 1) / BM;
  int bn = 64;
  if (N > 64 && (long long)mt * ((N + 127) / 128) * split >= (long long)ctx->sm_count) bn = 128;
  rc = make_tmaps<A_MN, B_MN>(ctx, A, lda, B, ldb, C, ldc, ep.out_dtype, M, N, K, bn, &ta, &tb, &tc);
  if (rc) return rc;
  if (bn == 64) return launch_tc<A_MN, B_MN, 64>(ctx, ta, tb, tc, M, N, K, split, ep, st);
  return launch_tc<A_MN, B_MN, 128>(ctx, ta, tb, tc, M, N, K, split, ep, st);
}

}  // namespace

int dmvae_gemm_bf16_tc(dmvae_ctx* ctx, int trans_a, int trans_b, const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
                       int64_t ldc, int M, int N, int K, const dmvae_gemm_epilogue* epi, cudaStream_t st) {
  if (!dmvae_ctx_has_tcgen05(ctx)) {
    dmvae_set_error("gemm(bf16): the tcgen05 path needs an sm_100 device (found sm_%d%d) - there is no fallback", ctx->cc_major,
                    ctx->cc_minor);
    return DMVAE_ERR_UNSUPPORTED;
  }
  DMVAE_CHECK_ARG(K > 0, "gemm(bf16): K must be positive");
  DMVAE_CHECK_ARG(N % 8 == 0 && ldc % 8 == 0, "gemm(bf16): N (%d) and ldc (%lld) must be multiples of 8", N, (long long)ldc);
  DMVAE_CHECK_ARG(epi->out_dtype == DMVAE_BF16 || epi->out_dtype == DMVAE_F32, "gemm(bf16): output must be bf16 or fp32");
  DMVAE_CHECK_ARG(((uintptr_t)C & 15) == 0, "gemm(bf16): C must be 16-byte aligned");
  if (epi->relu_mask) DMVAE_CHECK_ARG(epi->ld_mask % 8 == 0 && ((uintptr_t)epi->relu_mask & 15) == 0, "gemm(bf16): mask must be 16-byte aligned with ld % 8 == 0");
  EpiParams ep = make_epi_params(*epi, DMVAE_BF16);
  DMVAE_CHECK_ARG(epi->n_valid >= epi->n_block || epi->n_block % 32 == 0, "gemm(bf16): n_block (%d) must be a multiple of 32", epi->n_block);
  if (epi->recon) {
    const dmvae_recon_fuse* r = epi->recon;
    DMVAE_CHECK_ARG(epi->out_dtype == DMVAE_BF16 && epi->act == DMVAE_ACT_NONE && epi->split_k <= 1 && !epi->accumulate && !epi->relu_mask &&
                        epi->n_valid >= epi->n_block,
                    "gemm(bf16): the fused reconstruction term needs a plain bf16 output layer (no activation, mask, split-K, padding columns)");
    DMVAE_CHECK_ARG(!trans_a && !trans_b && M >= 2 * BM && N >= 128 && pair_enabled(),
                    "gemm(bf16): the fused reconstruction term exists in the CTA-pair forward kernel only (M >= 256, N >= 128, no transposes)");
    DMVAE_CHECK_ARG(r->X && r->r_part && r->D > 0 && r->D <= N && r->r_parts >= (N + 31) / 32 &&
                        (r->input_type == DMVAE_INPUT_BINARY || r->input_type == DMVAE_INPUT_REAL),
                    "gemm(bf16): recon fuse: X, r_part, 0 < D <= N, r_parts >= ceil(N / 32), binary | real");
    DMVAE_CHECK_ARG((r->x_dtype == DMVAE_U8 && r->D % 16 == 0 && r->ldx % 16 == 0 && ((uintptr_t)r->X & 15) == 0) ||
                        (r->x_dtype == DMVAE_F32 && r->D % 4 == 0 && r->ldx % 4 == 0 && ((uintptr_t)r->X & 15) == 0),
                    "gemm(bf16): recon fuse: targets must be uint8 (D, ldx multiples of 16) or fp32 (multiples of 4), 16-byte aligned");
  }
  const int split = epi->split_k;
  // op(A) [M,K]: trans_a=0 -> stored [M,K] = K-major; trans_a=1 -> stored [K,M] = MN-major
  // op(B) [K,N]: trans_b=0 -> stored [K,N] = MN-major; trans_b=1 -> stored [N,K] = K-major
  const int a_mn = trans_a ? 1 : 0, b_mn = trans_b ? 0 : 1;
  if (!a_mn && !b_mn) return dispatch_bn<0, 0>(ctx, A, lda, B, ldb, C, ldc, M, N, K, split, ep, st);
  if (!a_mn && b_mn) return dispatch_bn<0, 1>(ctx, A, lda, B, ldb, C, ldc, M, N, K, split, ep, st);
  if (a_mn && !b_mn) return dispatch_bn<1, 0>(ctx, A, lda, B, ldb, C, ldc, M, N, K, split, ep, st);
  return dispatch_bn<1, 1>(ctx, A, lda, B, ldb, C, ldc, M, N, K, split, ep, st);
}
