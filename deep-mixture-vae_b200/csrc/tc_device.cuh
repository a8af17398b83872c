// Device-side building blocks shared by the tcgen05 GEMM kernels (gemm_tc.cu, gemm_chain.cu): PTX wrappers for
// mbarrier / TMA / tcgen05, UMMA descriptors, the TMA-store epilogue, and the cached tensor-map encoder.
#pragma once
#include "common.cuh"
#include "epilogue.cuh"
#include "philox.cuh"

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

// Fused reparameterisation (priors.py:86-89) for the latent head: the slab row of this lane holds [mean | log_var]
// (fp32, columns [0, 2L) with 2L <= 32, SWIZZLE_128B layout).  Draws eps (Philox counter = (global row, column / 4,
// step, 0), identical to reparam_fwd_kernel) or takes the injected one, writes eps and the decoder operand row
// Z = [mean + exp(lv/2) eps | 1 | 0 ...] in operand dtype bf16.
__device__ __forceinline__ void fused_reparam_row(uint32_t stage, int lane, int m, int M, const dmvae_reparam_args& a) {
  if (m >= M) return;
  const int L = a.L;
  auto slab = [&](int j) -> float {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(stage + (uint32_t)(lane * 128 + (((j >> 2) ^ (lane & 7)) << 4) + ((j & 3) << 2))));
    return v;
  };
  const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  const uint32_t step = (uint32_t)(a.step_dev ? *a.step_dev : a.step);
  const uint32_t grow = (uint32_t)(a.row_offset + (uint64_t)m);
  __nv_bfloat16* z = reinterpret_cast<__nv_bfloat16*>(a.Z_out) + (int64_t)m * a.ld_z;
  for (int c8 = 0; c8 < a.z_cols; c8 += 8) {           // z_cols % 8 == 0, row pitch 16-byte aligned
    float zz[8];
#pragma unroll
    for (int h4 = 0; h4 < 2; ++h4) {
      const int l0 = c8 + 4 * h4;
      float e[4] = {0.f, 0.f, 0.f, 0.f};
      if (l0 < L) {
        if (a.eps_in) {
#pragma unroll
          for (int i = 0; i < 4; ++i) e[i] = (l0 + i < L) ? a.eps_in[(int64_t)m * L + l0 + i] : 0.f;
        } else {
          const uint4 x = philox4x32_10(make_uint4(grow, (uint32_t)(l0 >> 2), step, 0u), key);
          box_muller(x.x, x.y, e[0], e[1]);
          box_muller(x.z, x.w, e[2], e[3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int l = l0 + i;
        float v = l == L ? 1.f : 0.f;
        if (l < L) {
          a.eps_out[(int64_t)m * L + l] = e[i];
          v = slab(l) + expf(0.5f * slab(L + l)) * e[i];
        }
        zz[4 * h4 + i] = v;
      }
    }
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 hh = __floats2bfloat162_rn(zz[2 * i], zz[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&hh);
    }
    *reinterpret_cast<uint4*>(z + c8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// taddr      : TMEM address of (lane quarter base, column 0 of the tile's accumulator)
// m_base     : global row of lane 0;  n0: global column of accumulator column 0
// stage      : shared-memory address (1024-byte aligned) of this warp's private 4 KiB slab
// arrive_bar : cluster address of the "accumulator drained" barrier (0: none); signalled right after the last
//              TMEM read so that the MMA issuer can reuse the buffer while this warp still converts and stores
template <bool OUT_BF16, bool RELU, bool MASK, bool FUSE = false, bool RECON = false>
__device__ __forceinline__ void epilogue_warp_cols(uint32_t taddr, int col_begin, int col_end, int m_base, int n0, int M, int N,
                                                   const CUtensorMap* tmC, const EpiParams& ep, bool first_split,
                                                   uint32_t stage, int lane, uint32_t arrive_bar,
                                                   const dmvae_reparam_args* fuse = nullptr, uint32_t wait_bar = 0,
                                                   uint32_t wait_parity = 0) {
  constexpr int CP = OUT_BF16 ? 64 : 32;                  // columns per 128-byte slab row
  const bool padded = ep.n_valid < ep.n_block;
  const uint32_t st_row = stage + lane * 128;
  const int m = m_base + lane;
  const __nv_bfloat16* mrow = MASK ? reinterpret_cast<const __nv_bfloat16*>(ep.mask) + (int64_t)m * ep.ld_mask : nullptr;
  const bool m_ok = m < M;
  const int part_w = col_end - col_begin;                 // RECON: width of this warp's column range = one r_part slot
  if (col_end > N - n0) col_end = N - n0;                 // N % 8 == 0
  bool arrived = false;
  float r_row = 0.f;                                      // RECON: reconstruction term of (row m, this column range)
  // ReLU-mask rows (dgrad): 64 bytes per thread and 32 columns, software-pipelined one chunk ahead; the first chunk is
  // requested BEFORE waiting for the accumulator, so its DRAM / L2 latency hides behind the main loop
  auto load_mask = [&](int n, uint4 (&mk)[4]) {
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
      mk[g4] = make_uint4(0u, 0u, 0u, 0u);
      if (m_ok && n + g4 * 8 < N) mk[g4] = __ldg(reinterpret_cast<const uint4*>(mrow + n) + g4);
    }
  };
  uint4 mk_next[4];
  if (MASK && col_begin < col_end) load_mask(n0 + col_begin, mk_next);
  // RECON: the targets of the same 32 columns.  uint8 targets (32 bytes per thread and chunk) are pipelined like the mask;
  // fp32 targets (128 bytes) are loaded where they are used - holding two chunks of them would spill
  const bool x_u8 = RECON && ep.rx_dtype == DMVAE_U8;
  const char* xrow = RECON ? reinterpret_cast<const char*>(ep.rx) + (int64_t)m * ep.rx_ld * (x_u8 ? 1 : 4) : nullptr;
  auto load_x8 = [&](int n, uint4 (&xr)[2]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      xr[q] = make_uint4(0u, 0u, 0u, 0u);
      if (m_ok && n + 16 * q < ep.rx_D) xr[q] = __ldg(reinterpret_cast<const uint4*>(xrow + n) + q);
    }
  };
  uint4 xr_next[2];
  if (RECON && x_u8 && col_begin < col_end) load_x8(n0 + col_begin, xr_next);
  if (wait_bar != 0) {
    mbar_wait(wait_bar, wait_parity);
    tcgen05_fence_after();
  }
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
        for (int g4 = 0; g4 < 4; ++g4) mk[g4] = mk_next[g4];
        if (c0 + sub + 32 < col_end) load_mask(n + 32, mk_next);
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
      if (RECON) {
        // v holds the fp32 decoder logits of (row m, columns n .. n+31): replace them by the gradient of the reconstruction
        // term and add the term itself to r_row (define_recon_loss, base_models.py:72-85; arithmetic of recon8_t in
        // elbo.cu: one tanh per element, the log term as a running product folded once per 32 columns)
        const int D = ep.rx_D;
        uint4 xr[2];
        if (x_u8) {
          xr[0] = xr_next[0]; xr[1] = xr_next[1];
          if (c0 + sub + 32 < col_end) load_x8(n + 32, xr_next);
        }
        const float sc = ep.rx_s;
        float acc = 0.f;
        const bool binary = ep.rx_input == DMVAE_INPUT_BINARY;
        if (x_u8 && binary && n + 32 <= D) {
          // Interior chunk of the common case (uint8 targets, Bernoulli likelihood), ~10 issue slots per element.
          // byte -> float without I2F: 0x4700bb00 is the float 32768 + bb (one PRMT), so with xc = x - 1/2:
          //   xc = fma(f, xs, -(32768 xs + 1/2)),   -s xc = fma(f, -s xs, s (32768 xs + 1/2))
          //   gradient s (sigmoid(d) - x) = fma(s / 2, u, -s xc),  u = sign(d) tanh(|d| / 2)
          //   term  max(d,0) - d x + log1p(e^-|d|) = |d| / 2 - d xc + ln 2 - ln(1 + |u|)
          // Rows >= M compute on zeros and are never stored.
          const float xs = ep.rx_scale, xo = -(32768.f * xs + 0.5f);
          const float nsx = -sc * xs, nso = -sc * xo, hs = 0.5f * sc;
          float accd = 0.f, prod = 1.f;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const uint32_t w = jj < 16 ? (&xr[0].x)[jj >> 2] : (&xr[1].x)[(jj - 16) >> 2];
            const float f = __uint_as_float(__byte_perm(w, 0x47000000u, 0x7404u | ((uint32_t)(jj & 3) << 4)));
            const float d = v[jj];
            const float xc = fmaf(f, xs, xo), nsxc = fmaf(f, nsx, nso), h = 0.5f * fabsf(d);
            float ua;
            asm("tanh.approx.f32 %0, %1;" : "=f"(ua) : "f"(h));
            const float u = __uint_as_float((__float_as_uint(d) & 0x80000000u) | __float_as_uint(ua));
            v[jj] = fmaf(hs, u, nsxc);
            acc += h;
            accd = fmaf(d, xc, accd);
            prod = fmaf(prod, ua, prod);
          }
          float lp;
          asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lp) : "f"(prod));
          acc = (acc - accd) + 0.6931471805599453f * (32.f - lp);
        } else {
          float x[32];
          if (x_u8) {
            const float xs = ep.rx_scale, xo = -32768.f * xs;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const uint32_t w = jj < 16 ? (&xr[0].x)[jj >> 2] : (&xr[1].x)[(jj - 16) >> 2];
              x[jj] = fmaf(__uint_as_float(__byte_perm(w, 0x47000000u, 0x7404u | ((uint32_t)(jj & 3) << 4))), xs, xo);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
              if (m_ok && n + 4 * q < D) w = __ldg(reinterpret_cast<const float4*>(xrow + (int64_t)n * 4) + q);
              x[4 * q] = w.x; x[4 * q + 1] = w.y; x[4 * q + 2] = w.z; x[4 * q + 3] = w.w;
            }
          }
          float prod = 1.f;
          int nvalid = 0;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const bool ok = m_ok && n + jj < D;
            const float d = v[jj];
            float g = 0.f;
            if (binary) {
              const float xc = x[jj] - 0.5f, h = 0.5f * fabsf(d);
              float ua;
              asm("tanh.approx.f32 %0, %1;" : "=f"(ua) : "f"(h));
              const float u = __uint_as_float((__float_as_uint(d) & 0x80000000u) | __float_as_uint(ua));
              g = sc * fmaf(0.5f, u, -xc);                // sigmoid(d) - x = u / 2 - (x - 1/2)
              if (ok) {
                acc += fmaf(-d, xc, h);                   // max(d,0) - d x = |d|/2 - d (x - 1/2)
                prod = fmaf(prod, ua, prod);              // log1p(e^{-|d|}) = ln 2 - ln(1 + u)
                ++nvalid;
              }
            } else {
              const float df = d - x[jj];                 // base_models.py:80-83
              g = sc * df;
              if (ok) acc = fmaf(0.5f * df, df, acc);
            }
            v[jj] = ok ? g : 0.f;                         // padding columns of d_decoded are zero
          }
          if (binary) {
            float lp;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lp) : "f"(prod));
            acc += 0.6931471805599453f * ((float)nvalid - lp);
          }
        }
        r_row += acc;
      }
      if (!RELU && !MASK && !RECON && ep.act == DMVAE_ACT_SIGMOID) {          // reconstructed_X = sigmoid(decoded_X), base_models.py:295-296
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __fdividef(1.f, 1.f + __expf(-v[j]));
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
    if (FUSE && !OUT_BF16 && n0 + c0 == 0) fused_reparam_row(stage, lane, m_base + lane, M, *fuse);
  }
  if (RECON && m_ok && col_begin < col_end) ep.r_part[(int64_t)m * ep.r_parts + (n0 + col_begin) / part_w] = r_row;
  if (arrive_bar != 0 && !arrived) {                      // nothing to drain (tile column range beyond N)
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(arrive_bar) : "memory");
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}

// warp-uniform dispatch to the specialised epilogues (WITH_RECON: compile the fused-reconstruction variant in)
template <bool WITH_RECON = true>
__device__ __forceinline__ void epilogue_warp(uint32_t taddr, int col_begin, int col_end, int m_base, int n0, int M, int N,
                                              const CUtensorMap* tmC, const EpiParams& ep, bool first_split, uint32_t stage,
                                              int lane, uint32_t arrive_bar, const dmvae_reparam_args* fuse = nullptr,
                                              uint32_t wait_bar = 0, uint32_t wait_parity = 0) {
  const bool relu = ep.act == DMVAE_ACT_RELU, mask = ep.mask != nullptr;
#define EPI_GO(B, R, K) epilogue_warp_cols<B, R, K>(taddr, col_begin, col_end, m_base, n0, M, N, tmC, ep, first_split, stage, lane, arrive_bar, nullptr, wait_bar, wait_parity)
  if (ep.out_dtype == DMVAE_BF16) {
    if (WITH_RECON && ep.rx != nullptr)                   // output layer with the reconstruction term fused (dmvae_recon_fuse)
      epilogue_warp_cols<true, false, false, false, true>(taddr, col_begin, col_end, m_base, n0, M, N, tmC, ep, first_split, stage,
                                                          lane, arrive_bar, nullptr, wait_bar, wait_parity);
    else if (mask) EPI_GO(true, false, true);             // dgrad (activation already applied upstream)
    else if (relu) EPI_GO(true, true, false);             // forward hidden layer
    else EPI_GO(true, false, false);
  } else {
    if (mask) EPI_GO(false, false, true);
    else if (relu) EPI_GO(false, true, false);
    else if (fuse)
      epilogue_warp_cols<false, false, false, true>(taddr, col_begin, col_end, m_base, n0, M, N, tmC, ep, first_split, stage, lane,
                                                    arrive_bar, fuse, wait_bar, wait_parity);   // latent head + fused reparameterisation
    else EPI_GO(false, false, false);                     // heads, dZ, weight gradients
  }
#undef EPI_GO
}

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

}  // namespace
