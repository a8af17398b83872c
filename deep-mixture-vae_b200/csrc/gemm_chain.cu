// A CHAIN of dependent dense-layer GEMMs in ONE persistent launch (bf16 tcgen05, CTA pairs).
//
// The step's GEMMs are tiny by Blackwell standards (M = batch = 4096, N / K = 64..4096: 0.5-10 us of tensor work each),
// so a launch per layer spends most of its time in launch latency, prologue, pipeline fill and the tail wave.  Here the
// layers of one pass (encoder -> heads (+ reparameterisation) -> decoder, or the data-gradient chain) are ONE work list
// of (layer, 256-row block, column tile, k-split) units walked by 74 persistent CTA pairs.  A unit of layer e reads
// rows of layer e-1's output, so it only has to wait for the column tiles of THAT row block: every epilogue warp
// bumps a per-(layer, row block) counter in global memory once its TMA stores have completed, and the TMA producer of a
// dependent unit spins on the counter (acquire) before issuing its loads.  Units are ordered layer-major, row-block
// major, so each pair only ever waits for units that come earlier in every pair's list: with all pairs co-resident
// (grid <= SM count, one CTA per SM) the schedule cannot deadlock, and the row blocks of consecutive layers overlap.
// The mechanics per unit (TMA -> smem stages -> tcgen05.mma cta_group::2 -> double-buffered TMEM -> TMA-store
// epilogue) are those of gemm_tc2_kernel; operand majors, tile width and epilogue are per-layer run-time values.
#include <algorithm>
#include <vector>

#include "tc_device.cuh"

namespace {

constexpr int kMaxChain = 16;
constexpr int CH_BN = 256;                                   // widest tile (TMEM: 2 x 256 columns)
constexpr int CH_STAGES = 6;
constexpr int kMaxSchedUnits = 64;           // longest per-pair unit list a scheduled launch may have (else: round robin)
constexpr int CH_STAGE_BYTES = A_TILE_BYTES + (CH_BN / 2) * BK * 2;   // 32 KiB
constexpr size_t CH_SMEM = (size_t)CH_STAGES * CH_STAGE_BYTES + kEpiBytes2 + 1024;

struct ChainLayer {
  CUtensorMap tmA, tmB, tmC;
  EpiParams ep;
  int M, N, K;
  int a_mn, b_mn, bn;
  int tiles_m, tiles_n, nsplit, kb_per_split, nkb;
  int unit_begin, unit_end;
  int dep[2];          // producer layer of an operand this layer reads (-1: none)
  int dep_all[2];      // 0: the unit needs the producer's row block m_blk only; 1: every row block (reduction over rows)
  int dep_need[2];     // counter value of one complete row block of the producer
  int dep_tiles_m[2];
  int fuse;            // 1: fused reparameterisation in the epilogue (latent head)
  int publish;         // 1: a later layer waits for this one (bump the row-block counters once the stores have landed)
};

struct ChainParams {
  int n_layers, n_units, cstride;
  const int* sched;    // unit schedule [pairs + 1 offsets][unit ids] (build_schedule), or NULL: unit u runs on pair u % pairs
  int* counters;       // [n_layers][cstride], zero at launch
  dmvae_reparam_args ra;
  ChainLayer L[kMaxChain];
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// spin until the counter reaches `need`; traps instead of hanging forever if the schedule is ever broken
__device__ __forceinline__ void wait_counter(const int* p, int need) {
  uint32_t spins = 0;
  while (ld_acquire(p) < need) {
    __nanosleep(64);
    if (++spins > (1u << 24)) __trap();
  }
}

struct Unit {
  int e, m_blk, n_blk, kb0, kb1;
};

__device__ __forceinline__ void decode_unit(const ChainParams& P, int u, int& e, Unit& out) {
  if (P.sched) e = 0;                           // scheduled units come in any order
  while (u >= P.L[e].unit_end) ++e;
  const ChainLayer& Ly = P.L[e];
  const int t = u - Ly.unit_begin;
  const int per_split = Ly.tiles_m * Ly.tiles_n;
  const int split = t / per_split, t2 = t - split * per_split;
  out.e = e;
  out.m_blk = t2 / Ly.tiles_n;                 // row-block major, column tile fastest
  out.n_blk = t2 - out.m_blk * Ly.tiles_n;
  out.kb0 = split * Ly.kb_per_split;
  out.kb1 = min(Ly.nkb, out.kb0 + Ly.kb_per_split);
}

// 160 registers (not the 168 ptxas would take): 10 warps x 160 leave room on the SM for one 256-thread block of the
// streamed Adam update, which then runs BESIDE the grouped gradient GEMMs instead of waiting for their CTAs to exit
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(152)
gemm_chain_kernel(const __grid_constant__ ChainParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * CH_STAGES + 4];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_units[kMaxSchedUnits];      // this pair's units, in execution order (scheduled launches)
  __shared__ int s_cnt;

  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_stage = tiles + CH_STAGES * CH_STAGE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[CH_STAGES + s]); };
  auto tfull_bar = [&](int a) { return smem_u32(&bars[2 * CH_STAGES + a]); };
  auto tempty_bar = [&](int a) { return smem_u32(&bars[2 * CH_STAGES + 2 + a]); };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < CH_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 16);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2 && P.sched) {                  // the table is written once, at the first use of the shape: no need to wait for the predecessor
    const int b = __ldg(P.sched + pair_id), n = __ldg(P.sched + pair_id + 1) - b;
    for (int i = lane; i < n; i += 32) s_units[i] = __ldg(P.sched + n_pairs + 1 + b + i);
    if (lane == 0) s_cnt = n;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(2 * CH_BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();
  pdl_launch_dependents();
  // this pair's units: its slice of the schedule (longest-processing-time-first, built on the host; copied into shared
  // memory by warp 2 before the cluster barrier above, so no role waits on a global load per unit), or every n_pairs-th unit
  int u_begin = pair_id, u_end = P.n_units, u_step = n_pairs;
  if (P.sched) {
    u_begin = 0;
    u_end = s_cnt;
    u_step = 1;
  }

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      const uint32_t lead_full0 = mapa_u32(full_bar(0), 0);
      uint32_t it = 0;
      int e = 0;
      for (int i = u_begin; i < u_end; i += u_step) {
        const int u = P.sched ? s_units[i] : i;
        Unit un;
        decode_unit(P, u, e, un);
        const ChainLayer& Ly = P.L[e];
        // ---- wait for the producers of this unit's operands ----
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const int pe = Ly.dep[d];
          if (pe < 0) continue;
          const int* c = P.counters + (size_t)pe * P.cstride;
          if (Ly.dep_all[d]) {
            for (int mb = 0; mb < Ly.dep_tiles_m[d]; ++mb) wait_counter(c + mb, Ly.dep_need[d]);
          } else {
            wait_counter(c + un.m_blk, Ly.dep_need[d]);
          }
        }
        asm volatile("fence.proxy.async;" ::: "memory");      // acquired data -> visible to the TMA (async proxy) loads
        const int bnh = Ly.bn >> 1;
        const uint32_t stage_tx = (uint32_t)(A_TILE_BYTES + bnh * BK * 2);
        const int m0 = un.m_blk * 2 * BM + (int)rank * BM;
        const int nb0 = un.n_blk * Ly.bn + (int)rank * bnh;
        for (int kb = un.kb0; kb < un.kb1; ++kb, ++it) {
          const int s = it % CH_STAGES;
          const uint32_t ph = (it / CH_STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (rank == 0) mbar_expect_tx(full_bar(s), 2 * stage_tx);
          const uint32_t fb = lead_full0 + (uint32_t)(s * 8);
          const uint32_t sa = tiles + s * CH_STAGE_BYTES, sb = sa + A_TILE_BYTES;
          const int k0 = kb * BK;
          if (!Ly.a_mn) {
            tma_load_2d_pair(sa, &Ly.tmA, fb, k0, m0);
          } else {
            tma_load_2d_pair(sa, &Ly.tmA, fb, m0, k0);
            tma_load_2d_pair(sa + 8192, &Ly.tmA, fb, m0 + 64, k0);
          }
          if (!Ly.b_mn) {
            tma_load_2d_pair(sb, &Ly.tmB, fb, k0, nb0);
          } else {
            for (int j = 0; j < bnh / 64; ++j) tma_load_2d_pair(sb + j * 8192, &Ly.tmB, fb, nb0 + 64 * j, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      uint32_t it = 0, ui = 0;
      int e = 0;
      for (int i = u_begin; i < u_end; i += u_step, ++ui) {
        const int u = P.sched ? s_units[i] : i;
        Unit un;
        decode_unit(P, u, e, un);
        const ChainLayer& Ly = P.L[e];
        const int a_mn = Ly.a_mn, b_mn = Ly.b_mn;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
                               ((uint32_t)(Ly.bn >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
        const uint32_t as = ui & 1u, aph = (ui >> 1) & 1u;
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tcgen05_fence_after();
        const uint32_t acc = tmem_base + as * CH_BN;
        for (int kb = un.kb0; kb < un.kb1; ++kb, ++it) {
          const int s = it % CH_STAGES;
          const uint32_t ph = (it / CH_STAGES) & 1u;
          mbar_wait(full_bar(s), ph);
          tcgen05_fence_after();
          const uint32_t sa = tiles + s * CH_STAGE_BYTES, sb = sa + A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = a_mn ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t bd = b_mn ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
            tcgen05_mma_f16<2>(acc, ad, bd, idesc, (kb > un.kb0 || k > 0) ? 1u : 0u);
          }
          tcgen05_commit_pair(empty_bar(s), 3);
        }
        tcgen05_commit_pair(tfull_bar(as), 3);
      }
    }
  } else {
    // ===================== epilogue (both CTAs) =====================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t lead_tempty0 = mapa_u32(tempty_bar(0), 0);
    const uint32_t my_stage = epi_stage + (uint32_t)(warp - 2) * kEpiStageBytes;
    uint32_t ui = 0;
    int e = 0;
    for (int i = u_begin; i < u_end; i += u_step, ++ui) {
      const int u = P.sched ? s_units[i] : i;
      Unit un;
      decode_unit(P, u, e, un);
      const ChainLayer& Ly = P.L[e];
      const uint32_t as = ui & 1u, aph = (ui >> 1) & 1u;
      const int bnh = Ly.bn >> 1;
      epilogue_warp<false>(tmem_base + as * CH_BN + ((uint32_t)(q * 32) << 16), half * bnh, (half + 1) * bnh,
                    un.m_blk * 2 * BM + (int)rank * BM + q * 32, un.n_blk * Ly.bn, Ly.M, Ly.N, &Ly.tmC, Ly.ep, un.kb0 == 0,
                    my_stage, lane, lead_tempty0 + as * 8, (Ly.fuse && half == 0) ? &P.ra : nullptr, tfull_bar(as), aph);
      if (Ly.publish) {
        // publish: this warp's part of the tile is in global memory
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
        if (Ly.fuse) __threadfence();                           // the fused rows were written with ordinary stores by every lane
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(P.counters + (size_t)e * P.cstride + un.m_blk) : "memory");
        }
      }
    }
  }
  // ===================== teardown =====================
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * CH_BN) : "memory");
  }
}

}  // namespace

// Units of a grouped launch differ a lot in length (a weight-gradient unit contracts over the batch: 64-128 k-blocks; a
// data-gradient unit over a layer width: 8-64), and "unit u on pair u % pairs" leaves some pairs with two long units while
// others hold one short one (modelled makespan up to 1.4x the mean at the 4096-row shapes).  Independent launches therefore
// get a longest-processing-time-first assignment: units sorted by cost (k-blocks x bytes per k-block + a fixed part for
// prologue / epilogue), each given to the least-loaded pair.  The table lives in device memory, keyed by the tiling.
int* build_schedule(dmvae_ctx* ctx, const ChainParams& P, int pairs, bool capturing) {
  std::vector<int> key;
  key.push_back(pairs);
  for (int i = 0; i < P.n_layers; ++i) {
    const ChainLayer& L = P.L[i];
    key.insert(key.end(), {L.tiles_m, L.tiles_n, L.nsplit, L.kb_per_split, L.nkb, L.bn});
  }
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    auto it = ctx->scheds.find(key);
    if (it != ctx->scheds.end()) return it->second;
  }
  if (capturing) return nullptr;              // no allocation inside a capture: round robin (the eager step normally came first)
  struct U { int cost, id; };
  std::vector<U> us;
  us.reserve(P.n_units);
  for (int i = 0; i < P.n_layers; ++i) {
    const ChainLayer& L = P.L[i];
    const int per_split = L.tiles_m * L.tiles_n;
    for (int t = 0; t < L.unit_end - L.unit_begin; ++t) {
      const int split = t / per_split;
      const int kb = std::min(L.nkb, (split + 1) * L.kb_per_split) - split * L.kb_per_split;
      us.push_back({16 + kb * (L.bn == 256 ? 4 : 3), L.unit_begin + t});     // 8 KB of operands per cost unit; 16 ~ four wide k-blocks
    }
  }
  // what "unit u on pair u % pairs" would give under the same cost model
  std::vector<long long> rr(pairs, 0);
  for (const U& u : us) rr[u.id % pairs] += u.cost;
  const long long rr_makespan = *std::max_element(rr.begin(), rr.end());
  std::stable_sort(us.begin(), us.end(), [](const U& a, const U& b) { return a.cost > b.cost; });
  std::vector<std::vector<int>> lists(pairs);
  std::vector<long long> load(pairs, 0);
  for (const U& u : us) {
    int best = 0;
    for (int p = 1; p < pairs; ++p)
      if (load[p] < load[best]) best = p;
    lists[best].push_back(u.id);
    load[best] += u.cost;
  }
  // Keep the plain order unless the schedule shortens the modelled makespan by more than 3 %: where both are equal (most
  // 4096-row shapes: two units per pair either way) the plain order measured 2 % faster - it runs neighbouring tiles of one
  // GEMM at the same time on neighbouring pairs.
  if (*std::max_element(load.begin(), load.end()) * 103 >= rr_makespan * 100) {
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->scheds[key] = nullptr;
    return nullptr;
  }
  // Order inside a pair: the epilogue of a unit overlaps the main loop of the NEXT one (double-buffered accumulators), the
  // last unit's epilogue is exposed.  So the units with the heavy epilogue (fp32 tiles reduced into the gradient buffer:
  // 4 bytes x 256 x bn of red.global.add) go first and the bf16 data-gradient tiles last (measured at cfg2: the other
  // order costs 3 % of the step).
  auto epi_bytes = [&](int id) {
    int e = 0;
    while (id >= P.L[e].unit_end) ++e;
    return P.L[e].bn * (P.L[e].ep.out_dtype == DMVAE_BF16 ? 2 : 4);
  };
  for (int p = 0; p < pairs; ++p)
    std::stable_sort(lists[p].begin(), lists[p].end(), [&](int a, int b) { return epi_bytes(a) > epi_bytes(b); });
  for (int p = 0; p < pairs; ++p)
    if ((int)lists[p].size() > kMaxSchedUnits) {          // very large launches: many units per pair, round robin is balanced enough
      std::lock_guard<std::mutex> g(ctx->mu);
      ctx->scheds[key] = nullptr;
      return nullptr;
    }
  std::vector<int> host(pairs + 1 + P.n_units);
  int off = 0;
  for (int p = 0; p < pairs; ++p) {
    host[p] = off;
    for (int id : lists[p]) host[pairs + 1 + off++] = id;
  }
  host[pairs] = off;
  int* dev = nullptr;
  if (cudaMalloc(&dev, host.size() * sizeof(int)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(dev, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(dev);
    return nullptr;
  }
  std::lock_guard<std::mutex> g(ctx->mu);
  ctx->scheds[key] = dev;
  return dev;
}

extern "C" int64_t dmvae_gemm_chain_counters(int n, int max_rows) {
  if (n <= 0 || max_rows <= 0) return 0;
  const int64_t tm = (max_rows + 2 * BM - 1) / (2 * BM);
  return (int64_t)n * std::max<int64_t>(tm, 64);
}

extern "C" int dmvae_gemm_chain(dmvae_ctx* ctx, const dmvae_chain_gemm* g, int n, int32_t* counters, int64_t counters_len,
                                int zero_counters, const dmvae_reparam_args* reparam, void* stream) {
  DMVAE_CHECK_ARG(ctx && g, "gemm_chain: NULL argument");
  DMVAE_CHECK_ARG(n >= 1 && n <= kMaxChain, "gemm_chain: 1..%d layers supported (got %d)", kMaxChain, n);
  if (!dmvae_ctx_has_tcgen05(ctx)) {
    dmvae_set_error("gemm_chain: needs an sm_100 device (found sm_%d%d) - there is no fallback", ctx->cc_major, ctx->cc_minor);
    return DMVAE_ERR_UNSUPPORTED;
  }
  static ChainParams P;                    // large (6 KiB); the call is serialised per process by the GIL-holding caller
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  memset(&P, 0, sizeof(P));
  int cstride = 1;
  for (int i = 0; i < n; ++i) cstride = std::max(cstride, (g[i].M + 2 * BM - 1) / (2 * BM));
  bool wants_dep = false;
  for (int i = 0; i < n; ++i) wants_dep = wants_dep || g[i].dep[0] >= 0 || g[i].dep[1] >= 0;
  DMVAE_CHECK_ARG(!wants_dep || (counters && (int64_t)n * cstride <= counters_len),
                  "gemm_chain: counters buffer missing or too small (%lld < %lld)", (long long)counters_len, (long long)n * cstride);
  P.n_layers = n;
  P.cstride = cstride;
  P.counters = counters;
  int units = 0, n_fuse = 0;
  bool any_dep = false;
  const int pairs_avail = ctx->sm_count / 2;
  for (int i = 0; i < n; ++i) {
    const dmvae_chain_gemm& e = g[i];
    ChainLayer& Ly = P.L[i];
    DMVAE_CHECK_ARG(e.A && e.B && e.C && e.M > 0 && e.N > 0 && e.K > 0, "gemm_chain[%d]: bad operands / sizes", i);
    DMVAE_CHECK_ARG(e.N % 8 == 0 && e.ldc % 8 == 0 && ((uintptr_t)e.C & 15) == 0, "gemm_chain[%d]: N, ldc multiples of 8 and C 16-byte aligned", i);
    DMVAE_CHECK_ARG(e.epi.n_valid >= e.epi.n_block || e.epi.n_block % 32 == 0, "gemm_chain[%d]: n_block must be a multiple of 32", i);
    DMVAE_CHECK_ARG(e.epi.split_k >= 1, "gemm_chain[%d]: split_k must be >= 1", i);
    DMVAE_CHECK_ARG(e.epi.split_k == 1 || (e.epi.accumulate && e.epi.out_dtype == DMVAE_F32 && !e.epi.relu_mask && e.epi.act == DMVAE_ACT_NONE),
                    "gemm_chain[%d]: split_k > 1 needs an accumulating fp32 linear epilogue", i);
    if (e.epi.relu_mask) DMVAE_CHECK_ARG(e.epi.ld_mask % 8 == 0 && ((uintptr_t)e.epi.relu_mask & 15) == 0, "gemm_chain[%d]: misaligned mask", i);
    Ly.M = e.M; Ly.N = e.N; Ly.K = e.K;
    Ly.a_mn = e.trans_a ? 1 : 0;
    Ly.b_mn = e.trans_b ? 0 : 1;
    DMVAE_CHECK_ARG(e.epi.recon == nullptr, "gemm_chain: entry %d: the fused reconstruction epilogue is dmvae_gemm's", i);
    Ly.ep = make_epi_params(e.epi, DMVAE_BF16);
    Ly.tiles_m = (e.M + 2 * BM - 1) / (2 * BM);
    Ly.nkb = (e.K + BK - 1) / BK;
    int split = e.epi.split_k;
    int kps = std::max(1, (Ly.nkb + split - 1) / split);
    split = (Ly.nkb + kps - 1) / kps;
    Ly.nsplit = split;
    Ly.kb_per_split = kps;
    // tile width: 256 when that still gives (most of) the pairs a unit of this layer, else 128
    const long long units256 = (long long)Ly.tiles_m * ((e.N + 255) / 256) * split;
    Ly.bn = (e.N >= 256 && units256 * 4 >= (long long)pairs_avail * 3) ? 256 : 128;
    Ly.tiles_n = (e.N + Ly.bn - 1) / Ly.bn;
    Ly.unit_begin = units;
    units += Ly.tiles_m * Ly.tiles_n * split;
    Ly.unit_end = units;
    Ly.fuse = e.fuse ? 1 : 0;
    if (Ly.fuse) {
      ++n_fuse;
      DMVAE_CHECK_ARG(reparam != nullptr && n_fuse == 1, "gemm_chain[%d]: one fused reparameterisation per chain, with its arguments", i);
      DMVAE_CHECK_ARG(e.epi.out_dtype == DMVAE_F32 && 2 * reparam->L <= 32 && reparam->z_dtype == DMVAE_BF16 && reparam->z_cols % 8 == 0 &&
                          reparam->ld_z % 8 == 0 && reparam->zeta_out == nullptr && reparam->rows == e.M && reparam->Z_out && reparam->eps_out,
                      "gemm_chain[%d]: fused reparameterisation needs fp32 head output, 2L <= 32, bf16 Z with 8-column granularity, no concrete sample", i);
      P.ra = *reparam;
    }
    for (int d = 0; d < 2; ++d) {
      const int pe = e.dep[d];
      Ly.dep[d] = -1;
      if (pe < 0) continue;
      DMVAE_CHECK_ARG(pe < i, "gemm_chain[%d]: dependency %d must refer to an earlier layer", i, pe);
      const ChainLayer& Pr = P.L[pe];
      Ly.dep[d] = pe;
      P.L[pe].publish = 1;
      any_dep = true;
      Ly.dep_all[d] = (e.dep_all[d] || g[pe].M != e.M) ? 1 : 0;
      Ly.dep_need[d] = Pr.tiles_n * Pr.nsplit * 16;            // 8 epilogue warps x 2 CTAs per unit
      Ly.dep_tiles_m[d] = Pr.tiles_m;
    }
    int rc = get_tmap(ctx, e.C, (uint64_t)e.N, (uint64_t)e.M, (uint64_t)e.ldc, e.epi.out_dtype == DMVAE_BF16 ? 64 : 32, 32, &Ly.tmC,
                      e.epi.out_dtype == DMVAE_BF16 ? 2 : 4);
    if (rc) return rc;
    if (!Ly.a_mn) rc = get_tmap(ctx, e.A, (uint64_t)e.K, (uint64_t)e.M, (uint64_t)e.lda, BK, BM, &Ly.tmA);
    else rc = get_tmap(ctx, e.A, (uint64_t)e.M, (uint64_t)e.K, (uint64_t)e.lda, 64, BK, &Ly.tmA);
    if (rc) return rc;
    if (!Ly.b_mn) rc = get_tmap(ctx, e.B, (uint64_t)e.K, (uint64_t)e.N, (uint64_t)e.ldb, BK, (uint32_t)(Ly.bn / 2), &Ly.tmB);
    else rc = get_tmap(ctx, e.B, (uint64_t)e.N, (uint64_t)e.K, (uint64_t)e.ldb, 64, BK, &Ly.tmB);
    if (rc) return rc;
  }
  P.n_units = units;
  cudaStream_t st = (cudaStream_t)stream;
  // a list without dependencies is a plain grouped launch: independent GEMMs share one persistent grid
  const bool memset_first = zero_counters && any_dep;
  if (memset_first) DMVAE_CUDA(cudaMemsetAsync(counters, 0, sizeof(int32_t) * (size_t)n * cstride, st));
  static bool opted = false;
  if (!opted) {
    DMVAE_CUDA(cudaFuncSetAttribute(gemm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM));
    opted = true;
  }
  // every pair must be resident at once (units wait for one another): never more pairs than the device holds
  const int pairs = std::max(1, std::min(units, pairs_avail));
  static int lpt = -1;                       // DMVAE_CHAIN_LPT=0: unit u on pair u % pairs (A/B measurements)
  if (lpt < 0) {
    const char* ev = getenv("DMVAE_CHAIN_LPT");
    lpt = (ev && ev[0] == '0') ? 0 : 1;
  }
  if (lpt && !any_dep && units > pairs) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    DMVAE_CUDA(cudaStreamIsCapturing(st, &cs));
    P.sched = build_schedule(ctx, P, pairs, cs != cudaStreamCaptureStatusNone);
  }
  DMVAE_CUDA(dmvae_launch(gemm_chain_kernel, dim3(2 * pairs), dim3(kThreads2), CH_SMEM, st, !memset_first, P));
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}
