// Fused ELBO for LARGE mixtures (K*L beyond the row-tile kernel: VaDE K=50 L=64, DMVAE K=100 L=128) - included by elbo.cu.
//
// The reference materialises [B,K,L] broadcasts for the mixture terms (priors.py:91-102, :130-145).  Here they are
// three small contractions ("[B,K,L] term at large K*L"):
//     A_bk   = c1_k - sum_l lv_bl - L + sum_l (e^{lv_bl} + mu_bl^2) iv_kl + sum_l mu_bl (-2 m_kl iv_kl)
//     s_bk   = -1/2 [ c1_k + sum_l z_bl^2 iv_kl + sum_l z_bl (-2 m_kl iv_kl) ]                      (VaDE score)
//     d mu   = mu (w . IV) + 1/2 (w . MIV2),   d lv = 1/2 (e^{lv} (w . IV) - sum_k w),   dZ_gamma = -z (ds . IV) - 1/2 (ds . MIV2)
// with iv = e^{-plv}, MIV2 = -2 m iv, c1_k = sum_l (m^2 iv + plv), evaluated with warp-level tf32 MMAs (m16n8k8) in the
// 3-term split form a_hi b_hi + a_lo b_hi + a_hi b_lo (fp32-class accuracy: the 1e-4 tier holds), one warp per 16 rows,
// the [K,L] tables resident in shared memory.  This part is a few GFLOP at most (HBM-bound kernel, not a GEMM: the
// warp-level MMA keeps it off the issue slots); the D-wide reconstruction part is a separate streaming kernel
// (elbo_recon_kernel) launched right behind it with a programmatic edge, so the two overlap.
//
// Fragment layouts (PTX ISA, mma.m16n8k8 tf32), g = lane / 4, t = lane % 4:
//   A (16x8): a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)      B (8x8): b0 (k=t, n=g)  b1 (k=t+4, n=g)
//   C (16x8): c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
// The second pass contracts over clusters with the permutation "A column t <-> cluster 2t, column t+4 <-> cluster 2t+1",
// which makes the first pass's C fragments directly usable as A fragments (no shuffles) and keeps the table reads
// bank-conflict free in both passes (table row pitch = 4 mod 32 floats).
#pragma once

// x = hi + lo (+ <= 2^-21 |x|): hi = the top 19 bits of x (truncation, the bits a tf32 operand keeps), lo = the top 19 bits
// of the exact remainder.  Three instructions (LOP3, FADD, LOP3); cvt.rna.tf32 costs ~5 SASS instructions on sm_100a and
// round-to-nearest buys nothing once the remainder is carried as a second operand.
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += a . b with both operands split: three MMAs
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1,
                                     uint32_t bl0, uint32_t bl1) {
  mma_tf32(c, al, bh0, bh1);
  mma_tf32(c, ah, bl0, bl1);
  mma_tf32(c, ah, bh0, bh1);
}
__device__ __forceinline__ float quad_sum_f(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

constexpr int kLatWarpsPerTile = 4;           // warps that share one 16-row tile
constexpr int kLatTilesPerCta = 4;
constexpr int kLatThreadsM = 32 * kLatWarpsPerTile * kLatTilesPerCta;     // 512
__host__ __device__ inline int elbo_mma_ls(int L) { return ((((L + 7) & ~7) + 27) / 32) * 32 + 4; }      // >= round8(L), == 4 (mod 32)
__host__ __device__ inline int elbo_mma_wp(int KT) { return ((KT * 8 + 23) / 32) * 32 + 8; }             // >= 8 KT, == 8 (mod 32)
// floats: tables | per tile group: q hi / lo (and d score hi / lo) [16][Wp] + exchange [4 warps][16 rows][4]
__host__ __device__ inline size_t elbo_mma_smem(int L, int KT, bool vade) {
  const size_t tab = 2 * (size_t)KT * 8 * elbo_mma_ls(L) + (size_t)KT * 8;
  const size_t grp = (size_t)(vade ? 4 : 2) * 16 * elbo_mma_wp(KT) + kLatWarpsPerTile * 16 * 4;
  return sizeof(float) * (tab + kLatTilesPerCta * grp);
}

// KT = number of 8-cluster groups (K <= 8 KT); VADE: scores from Z (priors.py:91-102) instead of the logits.
// Four warps share a 16-row tile: pass 1 is split over the cluster blocks (warp wq owns nt = wq, wq+4, ...), the row
// statistics of the softmax meet in shared memory, q (and the VaDE d score) go to shared memory already split, and
// pass 2 is split over the l-chunks (warp wq owns lc = wq, wq+4, ...).  A tile's latency - which is the kernel's
// duration, there being about as many tiles as warp schedulers - drops ~3.5x against one warp per tile.
template <int KT, bool VADE>
__global__ void __launch_bounds__(kLatThreadsM, 1) elbo_latent_mma_kernel(const ElboParams p) {
  extern __shared__ float smem[];
  constexpr int NTJ = (KT + kLatWarpsPerTile - 1) / kLatWarpsPerTile;
  const dmvae_elbo_args& a = p.a;
  const int L = a.L, K = a.K, Ls = elbo_mma_ls(L), Lc = (L + 7) >> 3;
  constexpr int Wp = ((KT * 8 + 23) / 32) * 32 + 8;
  float* IV = smem;                           // [8 KT][Ls]  e^{-plv}      (0 beyond K / L)
  float* MIV2 = IV + KT * 8 * Ls;             // [8 KT][Ls]  -2 m e^{-plv}
  float* c1 = MIV2 + KT * 8 * Ls;             // [8 KT]      sum_l (m^2 iv + plv)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int grp = warp / kLatWarpsPerTile, wq = warp % kLatWarpsPerTile;
  constexpr int kGrpFloats = (VADE ? 4 : 2) * 16 * Wp + kLatWarpsPerTile * 16 * 4;
  float* gbase = c1 + KT * 8 + grp * kGrpFloats;
  uint32_t* Wh = reinterpret_cast<uint32_t*>(gbase);          // [16][Wp] q, high part
  uint32_t* Wl = Wh + 16 * Wp;
  uint32_t* Dh = Wl + 16 * Wp;                                // VaDE: d score
  uint32_t* Dl = Dh + 16 * Wp;
  float* xch = gbase + (VADE ? 4 : 2) * 16 * Wp;              // [4 warps][16 rows][4]
  const int bar_id = 1 + grp;
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * kLatWarpsPerTile) : "memory"); };
  // The prior tables are parameters: their last writer (the previous step's update) finished long before this kernel's
  // stream predecessor started, so the table build runs BEFORE griddepcontrol.wait and overlaps that predecessor's tail.
  // Phase 1: raw tables -> shared memory (independent coalesced loads); phase 2: transform in place.
  for (int k = warp; k < KT * 8; k += kLatThreadsM / 32) {
    float* ivr = IV + k * Ls;
    float* mvr = MIV2 + k * Ls;
#pragma unroll 4
    for (int l = lane; l < Ls; l += 32) {
      const bool v = k < K && l < L;
      ivr[l] = v ? __ldg(a.prior_log_vars + k * L + l) : 0.f;
      mvr[l] = v ? __ldg(a.prior_means + k * L + l) : 0.f;
    }
    __syncwarp();
    float acc = 0.f;                            // c1_k = sum_l (m^2 e^{-plv} + plv): one warp per cluster, fixed order
    for (int l = lane; l < Ls; l += 32) {
      const bool v = k < K && l < L;
      const float plv = ivr[l], m = mvr[l];
      const float iv = v ? expf(-plv) : 0.f;
      acc += v ? m * m * iv + plv : 0.f;
      ivr[l] = iv;
      mvr[l] = -2.f * m * iv;
    }
    acc = warp_sum(acc);
    if (lane == 0) c1[k] = acc;
  }
  pdl_wait();
  pdl_launch_dependents();                    // the reconstruction kernel behind this one may start streaming right away
  __syncthreads();

  const float r = a.kl_ratio_dev ? __ldg(a.kl_ratio_dev) : a.kl_ratio;
  const float s = a.inv_global_batch, sr = s * r;
  const float logK = logf((float)K);
  const int n_tiles = (a.rows + 15) >> 4;
  // every warp of a group runs the same number of iterations (the group barriers are unconditional)
  for (int tile = grp * gridDim.x + blockIdx.x; tile < n_tiles; tile += gridDim.x * kLatTilesPerCta) {
    const int row_a = tile * 16 + g, row_b = row_a + 8;
    const bool ok_a = row_a < a.rows, ok_b = row_b < a.rows;
    const int64_t ra = ok_a ? row_a : a.rows - 1, rb = ok_b ? row_b : a.rows - 1;     // clamped for loads
    const float* mu_a = a.mean + ra * a.ld_zh;
    const float* mu_b = a.mean + rb * a.ld_zh;
    const float* lv_a = a.log_var + ra * a.ld_zh;
    const float* lv_b = a.log_var + rb * a.ld_zh;
    const float* ep_a = VADE ? a.eps + ra * a.ld_eps : nullptr;
    const float* ep_b = VADE ? a.eps + rb * a.ld_eps : nullptr;

    // ================= pass 1: A_bk (and the VaDE scores) of this warp's cluster blocks, contraction over l =================
    float accA[NTJ][4], accS[NTJ][4], accM[NTJ][4], accT[VADE ? NTJ : 1][4];
#pragma unroll
    for (int j = 0; j < NTJ; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        accA[j][c] = accS[j][c] = accM[j][c] = 0.f;
        if (VADE) accT[j][c] = 0.f;
      }
    float slv_a = 0.f, slv_b = 0.f;
    // row inputs of l-chunk lc in A-fragment order: (row a, l0) (row b, l0) (row a, l1) (row b, l1); loaded one chunk ahead
    auto load_chunk = [&](int lc, float (&mu)[4], float (&lv)[4], float (&ep)[4]) {
      const int l0 = lc * 8 + t, l1 = l0 + 4;
      const bool v0 = l0 < L, v1 = l1 < L;
      mu[0] = v0 ? __ldg(mu_a + l0) : 0.f; mu[1] = v0 ? __ldg(mu_b + l0) : 0.f;
      mu[2] = v1 ? __ldg(mu_a + l1) : 0.f; mu[3] = v1 ? __ldg(mu_b + l1) : 0.f;
      lv[0] = v0 ? __ldg(lv_a + l0) : 0.f; lv[1] = v0 ? __ldg(lv_b + l0) : 0.f;
      lv[2] = v1 ? __ldg(lv_a + l1) : 0.f; lv[3] = v1 ? __ldg(lv_b + l1) : 0.f;
      if (VADE) {
        ep[0] = v0 ? __ldg(ep_a + l0) : 0.f; ep[1] = v0 ? __ldg(ep_b + l0) : 0.f;
        ep[2] = v1 ? __ldg(ep_a + l1) : 0.f; ep[3] = v1 ? __ldg(ep_b + l1) : 0.f;
      }
    };
    float mu_n[4], lv_n[4], ep_n[4] = {0.f, 0.f, 0.f, 0.f};
    load_chunk(0, mu_n, lv_n, ep_n);
#pragma unroll 1
    for (int lc = 0; lc < Lc; ++lc) {
      const int l0 = lc * 8 + t, l1 = l0 + 4;
      const bool v0 = l0 < L, v1 = l1 < L;
      float mu[4], lv[4], ep[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { mu[i] = mu_n[i]; lv[i] = lv_n[i]; ep[i] = ep_n[i]; }
      if (lc + 1 < Lc) load_chunk(lc + 1, mu_n, lv_n, ep_n);
      slv_a += lv[0] + lv[2];
      slv_b += lv[1] + lv[3];
      uint32_t e_h[4], e_l[4], m_h[4], m_l[4], z2_h[4], z2_l[4], z_h[4], z_l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool v = (i < 2) ? v0 : v1;
        const float e2 = v ? fast_exp(lv[i]) + mu[i] * mu[i] : 0.f;
        tf32_split(e2, e_h[i], e_l[i]);
        tf32_split(mu[i], m_h[i], m_l[i]);
        if (VADE) {
          const float z = v ? mu[i] + fast_exp(0.5f * lv[i]) * ep[i] : 0.f;
          tf32_split(z * z, z2_h[i], z2_l[i]);
          tf32_split(z, z_h[i], z_l[i]);
        }
      }
      const float* ivp = IV + (wq * 8 + g) * Ls + lc * 8 + t;
      const float* mvp = MIV2 + (wq * 8 + g) * Ls + lc * 8 + t;
      // the table fragments of the warp's cluster blocks are split first, then the MMAs go out term by term so that
      // consecutive MMAs hit different accumulators (mma.sync has a long dependent-issue latency)
      uint32_t ih0[NTJ], il0[NTJ], ih1[NTJ], il1[NTJ], vh0[NTJ], vl0[NTJ], vh1[NTJ], vl1[NTJ];
#pragma unroll
      for (int j = 0; j < NTJ; ++j) {
        if (wq + 4 * j < KT) {
          tf32_split(ivp[j * 32 * Ls], ih0[j], il0[j]);
          tf32_split(ivp[j * 32 * Ls + 4], ih1[j], il1[j]);
          tf32_split(mvp[j * 32 * Ls], vh0[j], vl0[j]);
          tf32_split(mvp[j * 32 * Ls + 4], vh1[j], vl1[j]);
        }
      }
#pragma unroll
      for (int term = 0; term < 3; ++term) {
#pragma unroll
        for (int j = 0; j < NTJ; ++j) {
          if (wq + 4 * j < KT) {
            if (term == 0) {
              mma_tf32(accA[j], e_l, ih0[j], ih1[j]);
              mma_tf32(accM[j], m_l, vh0[j], vh1[j]);
              if (VADE) { mma_tf32(accS[j], z2_l, ih0[j], ih1[j]); mma_tf32(accT[j], z_l, vh0[j], vh1[j]); }
            } else if (term == 1) {
              mma_tf32(accA[j], e_h, il0[j], il1[j]);
              mma_tf32(accM[j], m_h, vl0[j], vl1[j]);
              if (VADE) { mma_tf32(accS[j], z2_h, il0[j], il1[j]); mma_tf32(accT[j], z_h, vl0[j], vl1[j]); }
            } else {
              mma_tf32(accA[j], e_h, ih0[j], ih1[j]);
              mma_tf32(accM[j], m_h, vh0[j], vh1[j]);
              if (VADE) { mma_tf32(accS[j], z2_h, ih0[j], ih1[j]); mma_tf32(accT[j], z_h, vh0[j], vh1[j]); }
            }
          }
        }
      }
    }
    slv_a = quad_sum_f(slv_a);
    slv_b = quad_sum_f(slv_b);

    // ================= softmax over K, KL terms, d score =================
    // C-fragment element c of block j: row (c < 2 ? a : b), cluster (wq + 4 j) * 8 + 2t + (c & 1)
    float mx_a = -INFINITY, mx_b = -INFINITY;
    int am_a = 1 << 30, am_b = 1 << 30;
#pragma unroll
    for (int j = 0; j < NTJ; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = (wq + 4 * j) * 8 + 2 * t + (c & 1);
        float sc = -INFINITY;
        if (k < K) {
          const float ck = c1[k];
          accA[j][c] = ck - ((c < 2) ? slv_a : slv_b) - (float)L + (accA[j][c] + accM[j][c]);       // A_bk
          if (VADE) sc = -0.5f * ((accS[j][c] + accT[j][c]) + ck);
          else sc = __ldg(a.logits + ((c < 2) ? ra : rb) * a.ld_logits + k);
        }
        accS[j][c] = sc;
        if (c < 2) { if (sc > mx_a) { mx_a = sc; am_a = k; } }           // ascending k: first maximum wins
        else { if (sc > mx_b) { mx_b = sc; am_b = k; } }
      }
    auto argmax_merge = [](float& m, int& i, float om, int oi) {
      if (om > m || (om == m && oi < i)) { m = om; i = oi; }
    };
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      argmax_merge(mx_a, am_a, __shfl_xor_sync(0xffffffffu, mx_a, o), __shfl_xor_sync(0xffffffffu, am_a, o));
      argmax_merge(mx_b, am_b, __shfl_xor_sync(0xffffffffu, mx_b, o), __shfl_xor_sync(0xffffffffu, am_b, o));
    }
    // ---- exchange 1: row maximum / argmax over the four warps ----
    if (t == 0) {
      xch[(wq * 16 + g) * 4 + 0] = mx_a; xch[(wq * 16 + g) * 4 + 1] = __int_as_float(am_a);
      xch[(wq * 16 + g + 8) * 4 + 0] = mx_b; xch[(wq * 16 + g + 8) * 4 + 1] = __int_as_float(am_b);
    }
    group_sync();
    mx_a = -INFINITY; mx_b = -INFINITY; am_a = am_b = 1 << 30;
#pragma unroll
    for (int w = 0; w < kLatWarpsPerTile; ++w) {
      argmax_merge(mx_a, am_a, xch[(w * 16 + g) * 4], __float_as_int(xch[(w * 16 + g) * 4 + 1]));
      argmax_merge(mx_b, am_b, xch[(w * 16 + g + 8) * 4], __float_as_int(xch[(w * 16 + g + 8) * 4 + 1]));
    }
    float den_a = 0.f, den_b = 0.f;
#pragma unroll
    for (int j = 0; j < NTJ; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = (wq + 4 * j) * 8 + 2 * t + (c & 1);
        const float e = (k < K) ? fast_exp(accS[j][c] - ((c < 2) ? mx_a : mx_b)) : 0.f;
        accS[j][c] = e;
        if (c < 2) den_a += e; else den_b += e;
      }
    den_a = quad_sum_f(den_a);
    den_b = quad_sum_f(den_b);
    // ---- exchange 2: softmax denominators (slot 2; slots 0/1 are still being read by slower warps) ----
    if (t == 0) { xch[(wq * 16 + g) * 4 + 2] = den_a; xch[(wq * 16 + g + 8) * 4 + 2] = den_b; }
    group_sync();
    den_a = den_b = 0.f;
#pragma unroll
    for (int w = 0; w < kLatWarpsPerTile; ++w) { den_a += xch[(w * 16 + g) * 4 + 2]; den_b += xch[(w * 16 + g + 8) * 4 + 2]; }
    const float inv_a = __fdividef(1.f, den_a), inv_b = __fdividef(1.f, den_b);
    float C_a = 0.f, C_b = 0.f, Z_a = 0.f, Z_b = 0.f, qG_a = 0.f, qG_b = 0.f, ws_a = 0.f, ws_b = 0.f;
    // accS <- q, accA <- G
#pragma unroll
    for (int j = 0; j < NTJ; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = (wq + 4 * j) * 8 + 2 * t + (c & 1);
        float q = 0.f, G = 0.f;
        if (k < K) {
          q = accS[j][c] * ((c < 2) ? inv_a : inv_b);
          const float A = accA[j][c];
          const float lq = __logf(q + kEps0);
          const float gC = lq + __fdividef(q, q + kEps0) + logK;
          G = r * (gC + 0.5f * A);
          if (VADE && a.d_gate_extra) G += __ldg(a.d_gate_extra + ((c < 2) ? ra : rb) * a.ld_dge + k) / s;
          const float cc = q * (lq + logK), zz = 0.5f * q * A, qg = q * G;     // priors.py:195-199, :137-145
          if (c < 2) { C_a += cc; Z_a += zz; qG_a += qg; ws_a += q; }
          else { C_b += cc; Z_b += zz; qG_b += qg; ws_b += q; }
        }
        accS[j][c] = q;
        accA[j][c] = G;
      }
    C_a = quad_sum_f(C_a); C_b = quad_sum_f(C_b);
    Z_a = quad_sum_f(Z_a); Z_b = quad_sum_f(Z_b);
    qG_a = quad_sum_f(qG_a); qG_b = quad_sum_f(qG_b);
    ws_a = quad_sum_f(ws_a); ws_b = quad_sum_f(ws_b);
    // ---- exchange 3: KL_c, KL_z, sum q G, sum q over the four warps (all readers of exchange 1 passed barrier 2) ----
    group_sync();                               // every warp has read the denominators
    if (t == 0) {
      float* xa = xch + (wq * 16 + g) * 4;
      float* xb = xch + (wq * 16 + g + 8) * 4;
      xa[0] = C_a; xa[1] = Z_a; xa[2] = qG_a; xa[3] = ws_a;
      xb[0] = C_b; xb[1] = Z_b; xb[2] = qG_b; xb[3] = ws_b;
    }
    group_sync();
    C_a = C_b = Z_a = Z_b = qG_a = qG_b = ws_a = ws_b = 0.f;
#pragma unroll
    for (int w = 0; w < kLatWarpsPerTile; ++w) {
      const float* xa = xch + (w * 16 + g) * 4;
      const float* xb = xch + (w * 16 + g + 8) * 4;
      C_a += xa[0]; Z_a += xa[1]; qG_a += xa[2]; ws_a += xa[3];
      C_b += xb[0]; Z_b += xb[1]; qG_b += xb[2]; ws_b += xb[3];
    }
    // outputs q, d score; shared copies (already split) for pass 2
#pragma unroll
    for (int j = 0; j < NTJ; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int nt = wq + 4 * j;
        const int k = nt * 8 + 2 * t + (c & 1);
        if (nt >= KT) continue;
        const bool okr = (c < 2) ? ok_a : ok_b;
        const int64_t row = (c < 2) ? row_a : row_b;
        const float q = accS[j][c];
        const float ds = s * q * (accA[j][c] - ((c < 2) ? qG_a : qG_b));
        const int si = (g + ((c >> 1) << 3)) * Wp + k;
        uint32_t hi, lo;
        tf32_split(q, hi, lo);
        Wh[si] = hi; Wl[si] = lo;
        if (VADE) {
          tf32_split(ds, hi, lo);
          Dh[si] = hi; Dl[si] = lo;
        }
        if (okr && k < K) {
          a.qc[row * K + k] = q;
          if (VADE) a.w_scratch[row * K + k] = ds;
          else if (a.dlogits_dtype == DMVAE_BF16)
            reinterpret_cast<__nv_bfloat16*>(a.d_logits)[row * a.ld_dlogits + k] = __float2bfloat16_rn(ds);
          else reinterpret_cast<float*>(a.d_logits)[row * a.ld_dlogits + k] = ds;
        }
      }
    if (!VADE) {                              // zero the padding columns [K, dlogits_cols) of this tile's rows
      const int npad = a.dlogits_cols - K;
      for (int i = wq * 32 + lane; i < 16 * npad; i += 32 * kLatWarpsPerTile) {
        const int rr = i / npad, cc = K + i - rr * npad;
        const int64_t row = (int64_t)tile * 16 + rr;
        if (row < a.rows) {
          if (a.dlogits_dtype == DMVAE_BF16)
            reinterpret_cast<__nv_bfloat16*>(a.d_logits)[row * a.ld_dlogits + cc] = __float2bfloat16_rn(0.f);
          else reinterpret_cast<float*>(a.d_logits)[row * a.ld_dlogits + cc] = 0.f;
        }
      }
    }
    if (t == 0 && wq == 0) {
      // the reconstruction kernel fills .x (R) and rewrites .w = recon_scale R + r (C + Zk)
      if (ok_a) {
        a.per_sample[4 * (int64_t)row_a + 1] = C_a;
        a.per_sample[4 * (int64_t)row_a + 2] = Z_a;
        a.argmax[row_a] = am_a;
      }
      if (ok_b) {
        a.per_sample[4 * (int64_t)row_b + 1] = C_b;
        a.per_sample[4 * (int64_t)row_b + 2] = Z_b;
        a.argmax[row_b] = am_b;
      }
    }
    group_sync();                               // q / d score of every cluster block are in shared memory

    // ================= pass 2: gradients wrt mean / log_var (/ Z through gamma) for this warp's l-chunks =================
    // contraction over clusters with "A column t <-> cluster 2t, column t+4 <-> cluster 2t+1" (file header)
#pragma unroll 1
    for (int lc = wq; lc < Lc; lc += kLatWarpsPerTile) {
      // cluster blocks nt = j (mod 4) accumulate into partial sums j: four independent MMA chains per output
      float wivp[4][4], wmvp[4][4], divp[VADE ? 4 : 1][4], dmvp[VADE ? 4 : 1][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          wivp[j][c] = wmvp[j][c] = 0.f;
          if (VADE) divp[j][c] = dmvp[j][c] = 0.f;
        }
      // B fragment: b0 = T[cluster 2t of the block][l = lc*8 + g], b1 = T[cluster 2t + 1][same l]
      const float* ivp = IV + (2 * t) * Ls + lc * 8 + g;
      const float* mvp = MIV2 + (2 * t) * Ls + lc * 8 + g;
      // A fragment of block nt: (row g, cluster 8nt+2t) (row g+8, 8nt+2t) (row g, 8nt+2t+1) (row g+8, 8nt+2t+1)
      const int wa = g * Wp + 2 * t, wb = (g + 8) * Wp + 2 * t;
#pragma unroll
      for (int n0 = 0; n0 < KT; n0 += 4) {
        uint32_t ih0[4], il0[4], ih1[4], il1[4], vh0[4], vl0[4], vh1[4], vl1[4];
        uint32_t qh[4][4], ql[4][4], dh[VADE ? 4 : 1][4], dl[VADE ? 4 : 1][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n0 + j < KT) {
            const int nt = n0 + j;
            tf32_split(ivp[nt * 8 * Ls], ih0[j], il0[j]);
            tf32_split(ivp[nt * 8 * Ls + Ls], ih1[j], il1[j]);
            tf32_split(mvp[nt * 8 * Ls], vh0[j], vl0[j]);
            tf32_split(mvp[nt * 8 * Ls + Ls], vh1[j], vl1[j]);
            const uint2 ha = *reinterpret_cast<const uint2*>(Wh + wa + nt * 8), hb = *reinterpret_cast<const uint2*>(Wh + wb + nt * 8);
            const uint2 la = *reinterpret_cast<const uint2*>(Wl + wa + nt * 8), lb = *reinterpret_cast<const uint2*>(Wl + wb + nt * 8);
            qh[j][0] = ha.x; qh[j][1] = hb.x; qh[j][2] = ha.y; qh[j][3] = hb.y;
            ql[j][0] = la.x; ql[j][1] = lb.x; ql[j][2] = la.y; ql[j][3] = lb.y;
            if (VADE) {
              const uint2 ea = *reinterpret_cast<const uint2*>(Dh + wa + nt * 8), eb = *reinterpret_cast<const uint2*>(Dh + wb + nt * 8);
              const uint2 fa = *reinterpret_cast<const uint2*>(Dl + wa + nt * 8), fb = *reinterpret_cast<const uint2*>(Dl + wb + nt * 8);
              dh[j][0] = ea.x; dh[j][1] = eb.x; dh[j][2] = ea.y; dh[j][3] = eb.y;
              dl[j][0] = fa.x; dl[j][1] = fb.x; dl[j][2] = fa.y; dl[j][3] = fb.y;
            }
          }
        }
#pragma unroll
        for (int term = 0; term < 3; ++term) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n0 + j < KT) {
              if (term == 0) {
                mma_tf32(wivp[j], ql[j], ih0[j], ih1[j]);
                mma_tf32(wmvp[j], ql[j], vh0[j], vh1[j]);
                if (VADE) { mma_tf32(divp[j], dl[j], ih0[j], ih1[j]); mma_tf32(dmvp[j], dl[j], vh0[j], vh1[j]); }
              } else if (term == 1) {
                mma_tf32(wivp[j], qh[j], il0[j], il1[j]);
                mma_tf32(wmvp[j], qh[j], vl0[j], vl1[j]);
                if (VADE) { mma_tf32(divp[j], dh[j], il0[j], il1[j]); mma_tf32(dmvp[j], dh[j], vl0[j], vl1[j]); }
              } else {
                mma_tf32(wivp[j], qh[j], ih0[j], ih1[j]);
                mma_tf32(wmvp[j], qh[j], vh0[j], vh1[j]);
                if (VADE) { mma_tf32(divp[j], dh[j], ih0[j], ih1[j]); mma_tf32(dmvp[j], dh[j], vh0[j], vh1[j]); }
              }
            }
          }
        }
      }
      float wiv[4], wmv[4], div[4], dmv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        wiv[c] = (wivp[0][c] + wivp[1][c]) + (wivp[2][c] + wivp[3][c]);
        wmv[c] = (wmvp[0][c] + wmvp[1][c]) + (wmvp[2][c] + wmvp[3][c]);
        div[c] = dmv[c] = 0.f;
        if (VADE) {
          div[c] = (divp[0][c] + divp[1][c]) + (divp[2][c] + divp[3][c]);
          dmv[c] = (dmvp[0][c] + dmvp[1][c]) + (dmvp[2][c] + dmvp[3][c]);
        }
      }
      // C fragment: (row a, l) (row a, l+1) (row b, l) (row b, l+1), l = lc*8 + 2t
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int l = lc * 8 + 2 * t + (c & 1);
        const bool okr = (c < 2) ? ok_a : ok_b;
        if (!okr || l >= L) continue;
        const int64_t row = (c < 2) ? row_a : row_b;
        const float mu = __ldg(((c < 2) ? mu_a : mu_b) + l), lv = __ldg(((c < 2) ? lv_a : lv_b) + l);
        a.d_mean_kl[row * a.ld_dkl + l] = sr * (mu * wiv[c] + 0.5f * wmv[c]);
        a.d_log_var_kl[row * a.ld_dkl + l] = sr * 0.5f * (fast_exp(lv) * wiv[c] - ((c < 2) ? ws_a : ws_b));
        if (VADE) {
          const float z = mu + fast_exp(0.5f * lv) * __ldg(((c < 2) ? ep_a : ep_b) + l);
          a.d_Z_gamma[row * a.ld_dzg + l] = -z * div[c] - 0.5f * dmv[c];
        }
      }
    }
    group_sync();                               // shared q / d score are free for the group's next tile
  }
}

// ---------------------------------------------------------------------------------------------------------------
// D-wide reconstruction part as a streaming kernel: one warp per row, 8 elements per lane and iteration, two
// iterations in flight.  Launched right behind the latent kernel with a programmatic edge and touches nothing that
// kernel writes until the final combine, so it only waits (griddepcontrol.wait) before reading KL_c / KL_z.
// Writes d_decoded (padding columns zeroed), per_sample.x = R and per_sample.w = recon_scale R + r (C + Zk).
// ---------------------------------------------------------------------------------------------------------------
template <typename TX, bool CENTRED>
__device__ __forceinline__ void gload8_x(const TX* p, float (&x)[8], float xs) {
  Vec8<TX>::load(p, x);
  if (sizeof(TX) == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = CENTRED ? fmaf(x[i], xs, -0.5f) : x[i] * xs;
  } else if (CENTRED) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] -= 0.5f;
  }
}

template <typename TX, typename TD, int INPUT>
__global__ void __launch_bounds__(256) elbo_recon_kernel(const ElboParams p) {
  constexpr bool FAST = sizeof(TD) == 2;
  constexpr bool CENTRED = FAST && INPUT == DMVAE_INPUT_BINARY;
  const dmvae_elbo_args& a = p.a;
  const int D = a.D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * 8;
  const float s_rec = a.inv_global_batch * a.recon_scale;
  pdl_launch_dependents();
  for (int row = blockIdx.x * 8 + warp; row < a.rows; row += warps_total) {
    const TX* xr = reinterpret_cast<const TX*>(a.X) + (int64_t)row * a.ldx;
    const TD* dr = reinterpret_cast<const TD*>(a.decoded) + (int64_t)row * a.ld_dec;
    TD* gr = reinterpret_cast<TD*>(a.d_decoded) + (int64_t)row * a.ld_ddec;
    float acc = 0.f, prod = 1.f, lacc = 0.f;
    if (p.vec_ok) {
      int j = lane * 8;
#pragma unroll 1
      for (int it = 0; j < D; ++it) {
        if (FAST) {
          float x0[8], x1[8];
          uint4 d0, d1 = make_uint4(0u, 0u, 0u, 0u);
          const bool two = j + 256 < D;
          gload8_x<TX, CENTRED>(xr + j, x0, p.xs);
          d0 = __ldg(reinterpret_cast<const uint4*>(dr + j));
          if (two) {
            gload8_x<TX, CENTRED>(xr + j + 256, x1, p.xs);
            d1 = __ldg(reinterpret_cast<const uint4*>(dr + j + 256));
          }
          uint32_t dw[4] = {d0.x, d0.y, d0.z, d0.w}, gw[4];
          recon8_t<INPUT, false>(x0, dw, gw, s_rec, acc, prod);
          *reinterpret_cast<uint4*>(gr + j) = make_uint4(gw[0], gw[1], gw[2], gw[3]);
          if (two) {
            uint32_t dw1[4] = {d1.x, d1.y, d1.z, d1.w};
            recon8_t<INPUT, false>(x1, dw1, gw, s_rec, acc, prod);
            *reinterpret_cast<uint4*>(gr + j + 256) = make_uint4(gw[0], gw[1], gw[2], gw[3]);
          }
          if (CENTRED && (it & 3) == 3) {       // fold the running product before it can overflow (<= 2^64 per fold)
            float lp;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lp) : "f"(prod));
            lacc += lp;
            prod = 1.f;
          }
          j += 512;
        } else {
          float x0[8], d0[8], g0[8];
          Vec8<TX>::load(xr + j, x0);
          Vec8<TD>::load(dr + j, d0);
          scale8<TX>(x0, p.xs);
          acc += recon8<INPUT>(x0, d0, g0, s_rec);
          Vec8<TD>::store(gr + j, g0);
          j += 256;
        }
      }
    } else {
      for (int j = lane; j < D; j += 32) {
        float g0;
        acc += recon1<INPUT>(scale1<TX>(to_f32<TX>(xr[j]), p.xs), to_f32<TD>(dr[j]), s_rec, g0);
        gr[j] = from_f32<TD>(g0);
      }
    }
    for (int j = D + lane; j < a.ddec_cols; j += 32) gr[j] = from_f32<TD>(0.f);
    if (CENTRED && p.vec_ok) {                  // + n ln 2 - ln prod (1 + u)     (recon8_t)
      float lp;
      asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lp) : "f"(prod));
      acc = fmaf(lacc + lp, -0.6931471805599453f, acc);
    }
    float R = warp_sum(acc);
    if (CENTRED && p.vec_ok) R = fmaf((float)D, 0.6931471805599453f, R);
    if (lane == 0) a.per_sample[4 * (int64_t)row] = R;
  }
  // ---- combine with the latent kernel's terms ----
  pdl_wait();
  const float r = a.kl_ratio_dev ? __ldg(a.kl_ratio_dev) : a.kl_ratio;
  __syncwarp();
  for (int row = blockIdx.x * 8 + warp; row < a.rows; row += warps_total) {
    if (lane == 0) {
      float4* ps = reinterpret_cast<float4*>(a.per_sample) + row;
      float4 v = *ps;
      v.w = a.recon_scale * v.x + r * (v.y + v.z);
      *ps = v;
    }
  }
}

inline bool elbo_mma_ok(const ElboParams& p) {
  static int enabled = -1;                    // DMVAE_ELBO_MMA=0 forces the warp-per-row kernel (A/B measurements)
  if (enabled < 0) {
    const char* e = getenv("DMVAE_ELBO_MMA");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  const dmvae_elbo_args& a = p.a;
  if (!enabled || a.mode == DMVAE_MODE_DMVAE_SAMPLED || a.K > 128) return false;
  return elbo_mma_smem(a.L, (a.K + 7) / 8 <= 2 ? 2 : (a.K + 7) / 8, a.mode == DMVAE_MODE_VADE) <= 220 * 1024;
}

template <int KT>
int launch_elbo_latent_kt(dmvae_ctx* ctx, const ElboParams& p, cudaStream_t st) {
  const bool vade = p.a.mode == DMVAE_MODE_VADE;
  const size_t smem = elbo_mma_smem(p.a.L, KT, vade);
  const int n_tiles = (p.a.rows + 15) / 16;
  // four 16-row tiles per CTA (four warps each); the tiles are spread over all SMs first
  const int blocks = std::max(1, std::min(ctx->sm_count, (n_tiles + kLatTilesPerCta - 1) / kLatTilesPerCta));
  if (vade) {
    auto kern = elbo_latent_mma_kernel<KT, true>;
    if (smem > 48 * 1024) DMVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dmvae_launch(kern, dim3(blocks), dim3(kLatThreadsM), smem, st, true, p);
  } else {
    auto kern = elbo_latent_mma_kernel<KT, false>;
    if (smem > 48 * 1024) DMVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dmvae_launch(kern, dim3(blocks), dim3(kLatThreadsM), smem, st, true, p);
  }
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

inline int launch_elbo_mma_latent(dmvae_ctx* ctx, const ElboParams& p, cudaStream_t st) {
  const int kt = (p.a.K + 7) / 8;
  if (kt <= 2) return launch_elbo_latent_kt<2>(ctx, p, st);
  if (kt <= 4) return launch_elbo_latent_kt<4>(ctx, p, st);
  if (kt <= 7) return launch_elbo_latent_kt<7>(ctx, p, st);
  if (kt <= 10) return launch_elbo_latent_kt<10>(ctx, p, st);
  if (kt <= 13) return launch_elbo_latent_kt<13>(ctx, p, st);
  return launch_elbo_latent_kt<16>(ctx, p, st);
}

template <typename TX, typename TD, int INPUT>
int launch_elbo_mma(dmvae_ctx* ctx, const ElboParams& p, cudaStream_t st) {
  int rc = launch_elbo_mma_latent(ctx, p, st);
  if (rc) return rc;
  const int blocks = std::max(1, std::min(ctx->sm_count * 8, (p.a.rows + 7) / 8));
  dmvae_launch(elbo_recon_kernel<TX, TD, INPUT>, dim3(blocks), dim3(256), 0, st, true, p);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Cross-sample reduction of the prior-table gradients as a contraction over the batch,
//     U[set][k][f] = sum_b W[b,k] F[b,f],   F = [mu | e^{lv} + mu^2 | 1]  (set 1, VaDE: W = d score, F = [z | z^2 | 1]),
// with the same split-tf32 MMAs: a CTA owns a chunk of 64 rows (operands staged in shared memory), warp w owns the
// 16-cluster block w % MT and every (8 / MT)-th 8-feature block; the chunk partials are summed in a fixed order by
// elbo_reduce_final_kernel (deterministic).  Same workspace layout as the scalar kernel it replaces.
// ---------------------------------------------------------------------------------------------------------------
// Fused-reconstruction mode (dmvae_elbo_args.r_part): the output-layer GEMM left per-column-range sums of the reconstruction
// term; the first reduce stage completes per_sample = (R, C, Zk, recon_scale R + r (C + Zk)) - one thread per row, slots
// added in index order (deterministic).
__device__ __forceinline__ float4 finish_per_sample(const dmvae_elbo_args& a, int64_t row) {
  float4 v = reinterpret_cast<const float4*>(a.per_sample)[row];
  if (a.r_part) {
    const float r = a.kl_ratio_dev ? __ldg(a.kl_ratio_dev) : a.kl_ratio;
    const float* rp = a.r_part + row * a.r_parts;
    float R = 0.f;
    for (int j = 0; j < a.r_parts; ++j) R += rp[j];
    v.x = R;
    v.w = a.recon_scale * R + r * (v.y + v.z);
    reinterpret_cast<float4*>(a.per_sample)[row] = v;
  }
  return v;
}

constexpr int kRedChunkM = 64;
// row pitch of the staged operands: >= n (n % 8 == 0) and == 8 (mod 16), i.e. == +-8 (mod 32): the MMA fragment loads
// (address t * pitch + g, t < 4, g < 8) touch 32 distinct banks either way.  The smaller pitch keeps the small-mixture
// launch under 16 KB of shared memory, which is what fits beside a resident GEMM CTA (this kernel runs on the side stream).
__host__ __device__ inline int red_pad8(int n) { return ((n + 7) / 16) * 16 + 8; }

// grid (row chunks, sets, column groups): a CTA computes U[set][:, its 8-feature blocks] of its 64 rows
template <int NTW>
__global__ void __launch_bounds__(256) elbo_reduce_mma_partial_kernel(const dmvae_elbo_args a, int G, int NTC, float* __restrict__ ws,
                                                                      int do_loss) {
  extern __shared__ float sm[];
  const int L = a.L, K = a.K, nF = 2 * L + 1;
  const int MT = (K + 15) >> 4, NT8 = (nF + 7) >> 3;
  const int Kw = red_pad8(MT * 16), Fw = red_pad8(NTC * 8);
  float* w_sm = sm;                           // [64][Kw]
  float* f_sm = sm + kRedChunkM * Kw;         // [64][Fw]: features [f_begin, f_begin + 8 NTC)
  const int gch = blockIdx.x, set = blockIdx.y, cg = blockIdx.z;
  const int nt_begin = cg * NTC, nt_end = min(NT8, nt_begin + NTC), f_begin = nt_begin * 8;
  const int b0 = gch * kRedChunkM;
  const int nb = min(kRedChunkM, a.rows - b0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  pdl_wait();
  pdl_launch_dependents();
  // flat, coalesced fills with four independent elements in flight per thread (the loads are L2 / HBM latency bound)
#pragma unroll 4
  for (int i = threadIdx.x; i < kRedChunkM * Kw; i += 256) {
    const int b = i / Kw, k = i - b * Kw;
    const int64_t row = b0 + b;
    float w = 0.f;
    if (b < nb && k < K) {
      if (a.mode == DMVAE_MODE_DMVAE_SAMPLED) w = __ldg(a.zeta + row * a.ld_zeta + k);
      else if (set == 1) w = __ldg(a.w_scratch + row * K + k);
      else w = __ldg(a.qc + row * K + k);
    }
    w_sm[i] = w;
  }
#pragma unroll 4
  for (int i = threadIdx.x; i < kRedChunkM * Fw; i += 256) {
    const int b = i / Fw, j = i - b * Fw;
    const int64_t row = b0 + b;
    const int f = f_begin + j;
    float v = 0.f;
    if (b < nb && f < nF && j < NTC * 8) {
      if (f == 2 * L) v = 1.f;
      else if (a.mode == DMVAE_MODE_DMVAE_SAMPLED) v = __ldg(a.f_scratch + row * 2 * L + f);
      else {
        const int l = f < L ? f : f - L;
        const float mu = __ldg(a.mean + row * a.ld_zh + l), lv = __ldg(a.log_var + row * a.ld_zh + l);
        if (set == 1) {
          const float z = mu + fast_exp(0.5f * lv) * __ldg(a.eps + row * a.ld_eps + l);
          v = f < L ? z : z * z;
        } else {
          v = f < L ? mu : fast_exp(lv) + mu * mu;
        }
      }
    }
    f_sm[i] = v;
  }
  __syncthreads();
  const int nsplit = max(1, 8 / MT);
  const int mt = warp % MT, slice = warp / MT;
  if (slice < nsplit) {
    float acc[NTW][4];
#pragma unroll
    for (int j = 0; j < NTW; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
#pragma unroll 1
    for (int ks = 0; ks < kRedChunkM / 8; ++ks) {
      const float* wp = w_sm + (ks * 8 + t) * Kw + mt * 16 + g;
      uint32_t ah[4], al[4];
      tf32_split(wp[0], ah[0], al[0]);
      tf32_split(wp[8], ah[1], al[1]);
      tf32_split(wp[4 * Kw], ah[2], al[2]);
      tf32_split(wp[4 * Kw + 8], ah[3], al[3]);
      const float* fp = f_sm + (ks * 8 + t) * Fw + g;
#pragma unroll
      for (int j0 = 0; j0 < NTW; j0 += 4) {
        uint32_t bh0[4], bl0[4], bh1[4], bl1[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int ntl = slice + (j0 + jj) * nsplit;               // local 8-feature block of this CTA
          if (j0 + jj < NTW && nt_begin + ntl < nt_end) {
            tf32_split(fp[ntl * 8], bh0[jj], bl0[jj]);
            tf32_split(fp[ntl * 8 + 4 * Fw], bh1[jj], bl1[jj]);
          }
        }
#pragma unroll
        for (int term = 0; term < 3; ++term)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int ntl = slice + (j0 + jj) * nsplit;
            if (j0 + jj < NTW && nt_begin + ntl < nt_end) {
              if (term == 0) mma_tf32(acc[j0 + jj], al, bh0[jj], bh1[jj]);
              else if (term == 1) mma_tf32(acc[j0 + jj], ah, bl0[jj], bl1[jj]);
              else mma_tf32(acc[j0 + jj], ah, bh0[jj], bh1[jj]);
            }
          }
      }
    }
    float* out = ws + ((size_t)set * G + gch) * (size_t)(K * nF);
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      const int nt = nt_begin + slice + j * nsplit;
      if (nt >= nt_end) continue;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = mt * 16 + g + ((c >> 1) << 3), f = nt * 8 + 2 * t + (c & 1);
        if (k < K && f < nF) out[(size_t)k * nF + f] = acc[j][c];
      }
    }
  }
  if (do_loss && set == 0 && cg == 0 && warp == 7) {
    // loss partials: per_sample[b] = (R, C, Zk, total)
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = lane; b < nb; b += 32) {
      const float4 v = finish_per_sample(a, b0 + b);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
    if (lane == 0) {
      const int nsets = (a.mode == DMVAE_MODE_VADE) ? 2 : 1;
      float* lp = ws + (size_t)nsets * G * (size_t)(K * nF) + (size_t)gch * 4;
      lp[0] = acc.x; lp[1] = acc.y; lp[2] = acc.z; lp[3] = acc.w;
    }
  }
}

inline bool elbo_reduce_mma_ok(const dmvae_elbo_args& a) {
  static int enabled = -1;                    // DMVAE_ELBO_MMA=0: scalar reduction kernel (A/B measurements)
  if (enabled < 0) {
    const char* e = getenv("DMVAE_ELBO_MMA");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  return enabled && a.K <= 128 && a.L <= 128;
}

inline int launch_elbo_reduce_mma(dmvae_ctx* ctx, const dmvae_elbo_args& a, int G, int nsets, float* ws, int do_loss, cudaStream_t st) {
  const int nF = 2 * a.L + 1, MT = (a.K + 15) / 16, NT8 = (nF + 7) / 8;
  const int nsplit = std::max(1, 8 / MT);
  // column groups: at most 9 feature blocks per warp (small code, ~100 registers) and enough CTAs to fill the chip
  int CG = 1;
  while ((NT8 + CG - 1) / CG > 9 * nsplit) ++CG;
  const int NTC = (NT8 + CG - 1) / CG;
  const int ntw = (NTC + nsplit - 1) / nsplit;
  const size_t smem = sizeof(float) * (size_t)kRedChunkM * (size_t)(red_pad8(MT * 16) + red_pad8(NTC * 8));
#define RED_GO(N)                                                                                              \
  do {                                                                                                         \
    auto kern = elbo_reduce_mma_partial_kernel<N>;                                                             \
    if (smem > 48 * 1024) DMVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    dmvae_launch(kern, dim3(G, nsets, CG), dim3(256), smem, st, true, a, G, NTC, ws, do_loss);                       \
  } while (0)
  if (ntw <= 3) RED_GO(3);
  else if (ntw <= 5) RED_GO(5);
  else RED_GO(9);
#undef RED_GO
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}
