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
#include "common.cuh"
#include "epilogue.cuh"

#include <stdlib.h>

#include <algorithm>

namespace {

constexpr int BM = 128;          // UMMA M per CTA
constexpr int BK = 64;           // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;          // single-CTA kernel: producer, MMA, 4 epilogue warps
constexpr int kThreads2 = 320;         // CTA-pair kernel : producer, MMA, 8 epilogue warps (two per TMEM lane quarter)
constexpr int kMaxStages = 8;
constexpr int A_TILE_BYTES = BM * BK * 2;   // 16 KiB
constexpr int kEpiStageBytes = 4096;        // per epilogue warp: 32 rows x 128 B
constexpr int kEpiBytes = 4 * kEpiStageBytes;
constexpr int kEpiBytes2 = 8 * kEpiStageBytes;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// address of the same shared-memory object in CTA `rank` of this cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// pair variant: data lands in this CTA's shared memory, the transaction bytes are counted on the mbarrier at
// cluster address `bar` (the leader CTA's "full" barrier)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrives on the barrier at the same shared-memory offset in every CTA of `mask` once all prior MMAs of the pair retire
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
template <int CG>
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  if (CG == 1) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, SWIZZLE_128B, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);           // [0,14)  start address >> 4
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;      // [16,30) leading-dimension byte offset >> 4
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;      // [32,46) stride-dimension byte offset >> 4
  d |= (uint64_t)1 << 46;                                 // [46,48) descriptor version = 1
  d |= (uint64_t)2 << 61;                                 // [61,64) layout type = SWIZZLE_128B
  return d;
}

template <int A_MN, int B_MN>
__device__ __forceinline__ constexpr uint32_t make_idesc(int umma_m, int umma_n) {
  return (1u << 4)                  // D format  = F32
         | (1u << 7)                // A format  = BF16
         | (1u << 10)               // B format  = BF16
         | ((uint32_t)A_MN << 15)   // A major   (0 = K, 1 = MN)
         | ((uint32_t)B_MN << 16)   // B major
         | ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)(umma_m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// epilogue: one warp drains 32 accumulator rows (its TMEM lane quarter) x [col_begin, col_end) of a tile
// ------------------------------------------------------------------------------------------------
// TMEM -> registers (thread = row) -> bias / ReLU / dgrad ReLU-mask / padding columns -> bf16 pack -> a 4 KiB
// shared-memory slab laid out exactly as a SWIZZLE_128B TMA box {128 bytes, 32 rows} -> one TMA store (or TMA
// reduce-add for split-K / accumulating outputs) per slab.  The TMA unit does the coalescing and clips rows >= M
// and columns >= N, so the warp spends its issue slots on the value transform only.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}

// taddr      : TMEM address of (lane quarter base, column 0 of the tile's accumulator)
// m_base     : global row of lane 0;  n0: global column of accumulator column 0
// stage      : shared-memory address (1024-byte aligned) of this warp's private 4 KiB slab
// arrive_bar : cluster address of the "accumulator drained" barrier (0: none); signalled right after the last
//              TMEM read so that the MMA issuer can reuse the buffer while this warp still converts and stores
template <bool OUT_BF16, bool RELU, bool MASK>
__device__ __forceinline__ void epilogue_warp_cols(uint32_t taddr, int col_begin, int col_end, int m_base, int n0, int M, int N,
                                                   const CUtensorMap* tmC, const EpiParams& ep, bool first_split,
                                                   uint32_t stage, int lane, uint32_t arrive_bar) {
  constexpr int CP = OUT_BF16 ? 64 : 32;                  // columns per 128-byte slab row
  const bool padded = ep.n_valid < ep.n_block;
  const uint32_t st_row = stage + lane * 128;
  const int m = m_base + lane;
  const __nv_bfloat16* mrow = MASK ? reinterpret_cast<const __nv_bfloat16*>(ep.mask) + (int64_t)m * ep.ld_mask : nullptr;
  const bool m_ok = m < M;
  if (col_end > N - n0) col_end = N - n0;                 // N % 8 == 0
  bool arrived = false;
#pragma unroll 1
  for (int c0 = col_begin; c0 < col_end; c0 += CP) {
    // the TMA store of the previous slab must have finished READING shared memory before it is overwritten
    if (c0 != col_begin) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    }
#pragma unroll
    for (int sub = 0; sub < CP; sub += 32) {
      const int n = n0 + c0 + sub;
      uint4 mk[4];
      if (MASK) {
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          mk[g4] = make_uint4(0u, 0u, 0u, 0u);
          if (m_ok && n + g4 * 8 < N) mk[g4] = __ldg(reinterpret_cast<const uint4*>(mrow + n) + g4);
        }
      }
      uint32_t raw[32];
      tmem_ld32(taddr + (uint32_t)(c0 + sub), raw);
      if (arrive_bar != 0 && c0 + CP >= col_end && sub + 32 >= CP) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(arrive_bar) : "memory");
        arrived = true;
      }
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
      if (ep.bias && first_split) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n + j < N) v[j] += __ldg(ep.bias + n + j);
      }
      if (RELU && !OUT_BF16) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (padded) {
        const int jb = n % ep.n_block;                    // n_block % 32 == 0: the chunk stays inside one block
        if (jb + 32 > ep.n_valid) {                       // the chunk touches the ones / zero padding columns
          const float one = RELU ? fmaxf(ep.pad_one, 0.f) : ep.pad_one;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int t = jb + j;
            v[j] = t < ep.n_valid ? v[j] : (t == ep.n_valid ? one : 0.f);
          }
        }
      }
      if (OUT_BF16) {
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          uint32_t pk[4];
          const uint32_t mw[4] = {mk[g4].x, mk[g4].y, mk[g4].z, mk[g4].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[g4 * 8 + 2 * i], v[g4 * 8 + 2 * i + 1]);
            if (RELU) h = __hmax2(h, __float2bfloat162_rn(0.f));
            pk[i] = *reinterpret_cast<uint32_t*>(&h);
            if (MASK) {
              const __nv_bfloat162 mv = *reinterpret_cast<const __nv_bfloat162*>(&mw[i]);
              pk[i] &= __hgt2_mask(mv, __float2bfloat162_rn(0.f));     // 0xffff per half where mask > 0
            }
          }
          const int ch = (sub >> 3) + g4;                 // 16-byte chunk index inside the 128-byte row
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + (uint32_t)((ch ^ (lane & 7)) << 4)),
                       "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                       : "memory");
        }
      } else {
        if (MASK) {
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const uint32_t mw[4] = {mk[g4].x, mk[g4].y, mk[g4].z, mk[g4].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              v[g4 * 8 + 2 * i] = ((mw[i] & 0xffffu) - 1u) < 0x7fffu ? v[g4 * 8 + 2 * i] : 0.f;
              v[g4 * 8 + 2 * i + 1] = ((mw[i] >> 16) - 1u) < 0x7fffu ? v[g4 * 8 + 2 * i + 1] : 0.f;
            }
          }
        }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + (uint32_t)((ch ^ (lane & 7)) << 4)),
                       "r"(__float_as_uint(v[ch * 4])), "r"(__float_as_uint(v[ch * 4 + 1])),
                       "r"(__float_as_uint(v[ch * 4 + 2])), "r"(__float_as_uint(v[ch * 4 + 3]))
                       : "memory");
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA unit
    __syncwarp();
    if (lane == 0) {
      if (ep.accumulate) tma_reduce_add_2d(tmC, stage, n0 + c0, m_base);
      else tma_store_2d(tmC, stage, n0 + c0, m_base);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (arrive_bar != 0 && !arrived) {                      // nothing to drain (tile column range beyond N)
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(arrive_bar) : "memory");
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}

// warp-uniform dispatch to the specialised epilogues
__device__ __forceinline__ void epilogue_warp(uint32_t taddr, int col_begin, int col_end, int m_base, int n0, int M, int N,
                                              const CUtensorMap* tmC, const EpiParams& ep, bool first_split, uint32_t stage,
                                              int lane, uint32_t arrive_bar) {
  const bool relu = ep.act == DMVAE_ACT_RELU, mask = ep.mask != nullptr;
#define EPI_GO(B, R, K) epilogue_warp_cols<B, R, K>(taddr, col_begin, col_end, m_base, n0, M, N, tmC, ep, first_split, stage, lane, arrive_bar)
  if (ep.out_dtype == DMVAE_BF16) {
    if (mask) EPI_GO(true, false, true);                  // dgrad (activation already applied upstream)
    else if (relu) EPI_GO(true, true, false);             // forward hidden layer
    else EPI_GO(true, false, false);
  } else {
    if (mask) EPI_GO(false, false, true);
    else if (relu) EPI_GO(false, true, false);
    else EPI_GO(false, false, false);                     // heads, dZ, weight gradients
  }
#undef EPI_GO
}

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
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    if (nkb > 0)
      epilogue_warp(tmem_base + ((uint32_t)(q * 32) << 16), 0, BN, m0 + q * 32, n0, M, N, &tmC, ep, blockIdx.z == 0,
                    epi_stage + q * kEpiStageBytes, lane, 0u);
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
      mbar_wait(tfull_bar(as), aph);
      tcgen05_fence_after();
      epilogue_warp(tmem_base + as * BN + ((uint32_t)(q * 32) << 16), half * (BN / 2), (half + 1) * (BN / 2),
                    m_blk * 2 * BM + (int)rank * BM + q * 32, n_blk * BN, M, N, &tmC, ep, kb0 == 0, my_stage, lane,
                    lead_tempty0 + as * 8);
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D tensor [outer, inner] of esz-byte elements (2: bf16, 4: fp32) with row pitch ld elements; box {b_inner, b_outer};
// 128-byte swizzle; OOB reads give 0, OOB writes are dropped.
int get_tmap(dmvae_ctx* ctx, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t b_inner, uint32_t b_outer,
             CUtensorMap* out, uint32_t esz = 2) {
  TmapKey key{(uint64_t)(uintptr_t)ptr, inner, outer, ld, b_inner, b_outer, esz, 128u};
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    auto it = ctx->tmaps.find(key);
    if (it != ctx->tmaps.end()) {
      *out = it->second;
      return DMVAE_OK;
    }
  }
  DMVAE_CHECK_ARG(((uintptr_t)ptr & 15) == 0, "gemm(bf16): operand pointer must be 16-byte aligned");
  DMVAE_CHECK_ARG((ld * esz) % 16 == 0, "gemm(bf16): leading dimension (%llu) must be a multiple of 16 bytes", (unsigned long long)ld);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * esz};
  cuuint32_t box[2] = {b_inner, b_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap tm;
  CUresult r = ((EncodeTiledFn)ctx->encode_tiled)(&tm, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                                                  const_cast<void*>(ptr), dims, strides,
                                                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    dmvae_set_error("cuTensorMapEncodeTiled failed with CUresult %d (ptr=%p inner=%llu outer=%llu ld=%llu box=%ux%u)", (int)r, ptr,
                    (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, b_inner, b_outer);
    return DMVAE_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    if (ctx->tmaps.size() > 4096) ctx->tmaps.clear();
    ctx->tmaps[key] = tm;
  }
  *out = tm;
  return DMVAE_OK;
}

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
  int stages = BN == 64 ? 4 : 3;      // two CTAs per SM: one tile's epilogue overlaps the other's main loop
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
  const int split = epi->split_k;
  // op(A) [M,K]: trans_a=0 -> stored [M,K] = K-major; trans_a=1 -> stored [K,M] = MN-major
  // op(B) [K,N]: trans_b=0 -> stored [K,N] = MN-major; trans_b=1 -> stored [N,K] = K-major
  const int a_mn = trans_a ? 1 : 0, b_mn = trans_b ? 0 : 1;
  if (!a_mn && !b_mn) return dispatch_bn<0, 0>(ctx, A, lda, B, ldb, C, ldc, M, N, K, split, ep, st);
  if (!a_mn && b_mn) return dispatch_bn<0, 1>(ctx, A, lda, B, ldb, C, ldc, M, N, K, split, ep, st);
  if (a_mn && !b_mn) return dispatch_bn<1, 0>(ctx, A, lda, B, ldb, C, ldc, M, N, K, split, ep, st);
  return dispatch_bn<1, 1>(ctx, A, lda, B, ldb, C, ldc, M, N, K, split, ep, st);
}
