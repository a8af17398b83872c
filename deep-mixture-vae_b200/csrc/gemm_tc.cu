// bf16 GEMM on the 5th-generation tensor cores (sm_100a): tcgen05.mma with fp32 accumulators in TMEM,
// operands staged in shared memory by TMA (128-byte swizzle), warp-specialised:
//   warp 0   : TMA producer (one elected lane)
//   warp 1   : TMEM allocator + MMA issuer (one elected lane)
//   warps 2-5: epilogue (tcgen05.ld -> bias/ReLU/mask/padding columns -> global)
//
// C[M,N] = epilogue(op(A) . op(B)) with both operand majors supported through the UMMA descriptors, so
// the forward (A K-major, W MN-major), dgrad (both K-major) and wgrad (both MN-major) GEMMs of a dense
// layer all read the SAME row-major activation / weight buffers - no transposed copies in HBM.
//
// Descriptor encodings follow the PTX ISA "tcgen05 shared memory descriptor" / "instruction descriptor"
// tables (bit layout as in cute/arch/mma_sm100_desc.hpp of the vendored CUTLASS headers).
#include "common.cuh"
#include "epilogue.cuh"

namespace {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 64;           // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;
constexpr int A_TILE_BYTES = BM * BK * 2;   // 16 KiB

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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
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

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
template <int A_MN, int B_MN, int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, void* __restrict__ C,
               int64_t ldc, int M, int N, int K, int kb_per_split, int stages, const EpiParams ep) {
  constexpr int B_TILE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 1];
  __shared__ uint32_t tmem_slot;

  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B needs 1024-byte alignment
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
      constexpr uint32_t idesc = (1u << 4)                  // D format  = F32
                                 | (1u << 7)                // A format  = BF16
                                 | (1u << 10)               // B format  = BF16
                                 | ((uint32_t)A_MN << 15)   // A major   (0 = K, 1 = MN)
                                 | ((uint32_t)B_MN << 16)   // B major
                                 | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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
          tcgen05_mma_f16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        tcgen05_commit(empty_bar(s));        // smem slot reusable once these MMAs retire
      }
      tcgen05_commit(tmem_full_bar);         // accumulator complete
    }
  } else {
    // ===================== epilogue =====================
    // Each thread owns one output row and walks the tile in 32-column chunks straight out of TMEM.  All
    // decisions (activation, mask, padding columns, store flavour) are warp-uniform and taken once per chunk;
    // the per-element work of a pure data chunk is a max / select, a bf16 pack and 16-byte stores.
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < M && nkb > 0;
    const bool padded = ep.n_valid < ep.n_block;
    int jb = padded ? (n0 % ep.n_block) : 0;  // column index inside the padding block (n_block % 32 == 0)
    const __nv_bfloat16* mrow = ep.mask ? reinterpret_cast<const __nv_bfloat16*>(ep.mask) + (int64_t)m * ep.ld_mask : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      const int n = n0 + c0;
      if (n >= N) break;
      uint32_t raw[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, raw);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
      const int ncols = min(32, N - n);
      if (row_ok) {
        if (ep.bias && blockIdx.z == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncols) v[j] += __ldg(ep.bias + n + j);
        }
        if (ep.act == DMVAE_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (mrow) {
          if (ncols == 32) {
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const uint4 w = __ldg(reinterpret_cast<const uint4*>(mrow + n) + g4);
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                // bf16 > 0  <=>  sign bit clear and magnitude non-zero
                const uint32_t lo = ww[i] & 0xffffu, hi = ww[i] >> 16;
                v[g4 * 8 + 2 * i] = (lo - 1u < 0x7fffu) ? v[g4 * 8 + 2 * i] : 0.f;
                v[g4 * 8 + 2 * i + 1] = (hi - 1u < 0x7fffu) ? v[g4 * 8 + 2 * i + 1] : 0.f;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) v[j] = __bfloat162float(mrow[n + j]) > 0.f ? v[j] : 0.f;
          }
        }
        if (padded && jb + 32 > ep.n_valid) {          // the chunk touches the ones / zero padding columns
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int t = jb + j;
            v[j] = t < ep.n_valid ? v[j] : (t == ep.n_valid ? ep.pad_one : 0.f);
          }
        }
        if (ep.out_dtype == DMVAE_BF16) {
          __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(C) + (int64_t)m * ldc + n;
          if (ncols == 32) {
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              uint32_t pk[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[g4 * 8 + 2 * i], v[g4 * 8 + 2 * i + 1]);
                pk[i] = *reinterpret_cast<uint32_t*>(&h);
              }
              reinterpret_cast<uint4*>(c)[g4] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) c[j] = __float2bfloat16_rn(v[j]);
          }
        } else {
          float* c = reinterpret_cast<float*>(C) + (int64_t)m * ldc + n;
          if (ncols == 32) {
#pragma unroll
            for (int g8 = 0; g8 < 8; ++g8) {
              float4 o = make_float4(v[g8 * 4], v[g8 * 4 + 1], v[g8 * 4 + 2], v[g8 * 4 + 3]);
              float4* cp = reinterpret_cast<float4*>(c) + g8;
              if (ep.accumulate == 2) {
                atomicAdd(cp, o);                          // split-K: vector fp32 reduction in L2
              } else if (ep.accumulate == 1) {
                float4 old = *cp;
                *cp = make_float4(old.x + o.x, old.y + o.y, old.z + o.z, old.w + o.w);
              } else {
                *cp = o;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) {
                if (ep.accumulate == 2) atomicAdd(c + j, v[j]);
                else if (ep.accumulate == 1) c[j] += v[j];
                else c[j] = v[j];
              }
          }
        }
      }
      if (padded) {
        jb += 32;
        if (jb >= ep.n_block) jb -= ep.n_block;
      }
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
// host side: tensor maps and launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D bf16 tensor [outer, inner] with row pitch ld elements; box {b_inner, b_outer}; 128-byte swizzle; OOB reads give 0.
int get_tmap(dmvae_ctx* ctx, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t b_inner, uint32_t b_outer,
             CUtensorMap* out) {
  TmapKey key{(uint64_t)(uintptr_t)ptr, inner, outer, ld, b_inner, b_outer, 2u, 128u};
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    auto it = ctx->tmaps.find(key);
    if (it != ctx->tmaps.end()) {
      *out = it->second;
      return DMVAE_OK;
    }
  }
  DMVAE_CHECK_ARG(((uintptr_t)ptr & 15) == 0, "gemm(bf16): operand pointer must be 16-byte aligned");
  DMVAE_CHECK_ARG((ld * 2) % 16 == 0, "gemm(bf16): leading dimension (%llu) must be a multiple of 8 elements", (unsigned long long)ld);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {b_inner, b_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap tm;
  CUresult r = ((EncodeTiledFn)ctx->encode_tiled)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
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

template <int A_MN, int B_MN, int BN>
int launch_tc(dmvae_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, void* C, int64_t ldc, int M, int N, int K,
              int split, const EpiParams& ep, cudaStream_t st) {
  constexpr int STAGE_BYTES = A_TILE_BYTES + BN * BK * 2;
  const int nkb = (K + BK - 1) / BK;
  int kps = (nkb + split - 1) / split;
  if (kps < 1) kps = 1;
  split = (nkb + kps - 1) / kps;      // every split owns at least one k-block
  if (split < 1) split = 1;
  // two CTAs per SM (one tile's epilogue overlaps the other's main loop) unless the tile is 128x256
  int stages = BN == 256 ? 4 : (BN == 128 ? 3 : 4);
  if (stages > kMaxStages) stages = kMaxStages;
  const size_t smem = (size_t)stages * STAGE_BYTES + 1024;
  auto kern = gemm_tc_kernel<A_MN, B_MN, BN>;
  static size_t smem_opted = 0;       // per instantiation
  if (smem > smem_opted) {
    DMVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_opted = smem;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, split);
  kern<<<grid, kThreads, smem, st>>>(ta, tb, C, ldc, M, N, K, kps, stages, ep);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

template <int A_MN, int B_MN>
int dispatch_bn(dmvae_ctx* ctx, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int M, int N, int K,
                int split, const EpiParams& ep, cudaStream_t st) {
  // tile width: the widest that still gives every SM work
  const int mt = (M + BM - 1) / BM;
  int bn = 64;
  // 128-wide tiles unless they would leave SMs idle (two CTAs fit per SM): then 64-wide tiles double the CTA count
  if (N > 64 && (long long)mt * ((N + 127) / 128) * split >= (long long)ctx->sm_count) bn = 128;
  (void)mt;   // 128x256 tiles (one CTA per SM) need the persistent / double-buffered-TMEM variant to pay off
  CUtensorMap ta, tb;
  int rc;
  // A: K-major -> stored [M, K] (inner K); MN-major -> stored [K, M] (inner M)
  if (A_MN == 0) rc = get_tmap(ctx, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM, &ta);
  else rc = get_tmap(ctx, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK, &ta);
  if (rc) return rc;
  // B: K-major -> stored [N, K] (inner K); MN-major -> stored [K, N] (inner N)
  if (B_MN == 0) rc = get_tmap(ctx, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, (uint32_t)bn, &tb);
  else rc = get_tmap(ctx, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK, &tb);
  if (rc) return rc;
  if (bn == 64) return launch_tc<A_MN, B_MN, 64>(ctx, ta, tb, C, ldc, M, N, K, split, ep, st);
  if (bn == 128) return launch_tc<A_MN, B_MN, 128>(ctx, ta, tb, C, ldc, M, N, K, split, ep, st);
  return launch_tc<A_MN, B_MN, 256>(ctx, ta, tb, C, ldc, M, N, K, split, ep, st);
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
