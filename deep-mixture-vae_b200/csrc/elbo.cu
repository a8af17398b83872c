// Fused mixture-prior ELBO forward + backward (HBM-bound).
//
// One warp owns one sample row.  The D-wide reconstruction part streams X and the decoder logits
// once with 8-element vector loads and writes d_decoded; the latent part (KL_z under q(c|x), KL_c,
// softmax / logsumexp over K, and the gradients wrt mean, log_var, logits or Z) is done by the same
// warp with the [K,L] prior tables held in shared memory - the reference's [B,K,L] broadcasts
// (priors.py:131-145) are never materialised.  Cross-sample sums (prior-table gradients, the three
// loss terms) are produced by dmvae_elbo_reduce from the per-row outputs, deterministically.
//
// Formulas: SURVEY.md section 8(a'), restated and checked in oracle/closed_form.py.
#include "common.cuh"

#include <cuda_fp16.h>

#include <stdlib.h>

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxK = 128;                 // K values per lane <= 4
constexpr float kEps0 = 1e-20f;

struct ElboParams {
  dmvae_elbo_args a;
  float xs, xo;    // uint8 targets: x = byte * xs; centred form x - 1/2 = fma(2^15 + byte, xs, xo), xo = -(2^15 xs + 1/2)
  int Ls;          // padded (odd) row stride of the shared prior tables
  int vec_ok;      // 8-wide vector path usable for the streaming part
  int bulk_ok;     // rows can be moved with 16-byte-granular bulk async copies (row-tile kernel)
  int xrow, drow;  // bytes per row of the shared-memory target / logits tiles
  int contig;      // global row pitches equal xrow / drow: whole-tile bulk copies
};

template <int INPUT>
__device__ __forceinline__ float recon1(float x, float d, float s, float& g) {
  if (INPUT == DMVAE_INPUT_BINARY) {
    // max(x,0) - x z + log1p(exp(-|x|)); grad = sigmoid(x) - z          (base_models.py:74-79)
    float t = __expf(-fabsf(d));
    float inv = __fdividef(1.f, 1.f + t);
    float sig = d >= 0.f ? inv : t * inv;
    g = s * (sig - x);
    return fmaxf(d, 0.f) - d * x + __logf(1.f + t);
  } else {
    float df = d - x;                                                    // base_models.py:80-83
    g = s * df;
    return 0.5f * df * df;
  }
}

// uint8 targets carry `xs` per unit (1 for binarised data, 1/255 for 8-bit intensities); other dtypes are values already
template <typename TX>
__device__ __forceinline__ void scale8(float (&x)[8], float xs) {
  if (sizeof(TX) == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] *= xs;
  }
}
template <typename TX>
__device__ __forceinline__ float scale1(float x, float xs) { return sizeof(TX) == 1 ? x * xs : x; }

template <int INPUT>
__device__ __forceinline__ float recon8(const float (&x)[8], const float (&d)[8], float (&g)[8], float s) {
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += recon1<INPUT>(x[i], d[i], s, g[i]);
  return acc;
}

template <typename TX, typename TD, int INPUT, int kKPL>
__global__ void __launch_bounds__(kThreads) elbo_kernel(const ElboParams p) {
  extern __shared__ float smem[];
  pdl_wait();
  pdl_launch_dependents();
  const dmvae_elbo_args& a = p.a;
  const int L = a.L, K = a.K, Ls = p.Ls, D = a.D;
  const int mode = a.mode;
  float* tab_m = smem;                        // [K][Ls]
  float* tab_b = tab_m + K * Ls;              // [K][Ls]: exp(-plv) (analytic / VaDE) or plv (sampled)
  float* sum_plv = tab_b + K * Ls;            // [K]
  float* warp_base = sum_plv + ((K + 3) & ~3);
  const int Lp = (L + 3) & ~3, Kp = (K + 3) & ~3;
  const int per_warp = 4 * Lp + 2 * Kp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* mu_s = warp_base + warp * per_warp;  // [L]
  float* elv_s = mu_s + Lp;                   // [L] exp(lv)
  float* x1_s = elv_s + Lp;                   // [L] z (VaDE) or a (sampled)
  float* x2_s = x1_s + Lp;                    // [L] c (sampled)
  float* w_s = x2_s + Lp;                     // [K] weights (q, gamma or zeta)
  float* aux_s = w_s + Kp;                    // [K] d_s (VaDE)

  // ---- prior tables -> shared (once per CTA; the grid is persistent) ----
  for (int i = threadIdx.x; i < K * L; i += kThreads) {
    int k = i / L, l = i - k * L;
    float plv = a.prior_log_vars[i];
    tab_m[k * Ls + l] = a.prior_means[i];
    tab_b[k * Ls + l] = (mode == DMVAE_MODE_DMVAE_SAMPLED) ? plv : expf(-plv);
  }
  for (int k = threadIdx.x; k < K; k += kThreads) {
    float sacc = 0.f;
    for (int l = 0; l < L; ++l) sacc += a.prior_log_vars[k * L + l];
    sum_plv[k] = sacc;
  }
  __syncthreads();

  const float r = a.kl_ratio_dev ? __ldg(a.kl_ratio_dev) : a.kl_ratio;
  const float s = a.inv_global_batch, s_rec = a.inv_global_batch * a.recon_scale;
  const float logK = logf((float)K);
  const int warps_total = gridDim.x * kWarps;

  for (int row = blockIdx.x * kWarps + warp; row < a.rows; row += warps_total) {
    // latent inputs first: their latency overlaps the streaming part
    const float* mean = a.mean + (int64_t)row * a.ld_zh;
    const float* lvp = a.log_var + (int64_t)row * a.ld_zh;
    const float mu0 = lane < L ? __ldg(mean + lane) : 0.f, lv0 = lane < L ? __ldg(lvp + lane) : 0.f;
    float logit0 = -INFINITY;
    if (mode != DMVAE_MODE_VADE && lane < K) logit0 = __ldg(a.logits + (int64_t)row * a.ld_logits + lane);
    // =========================== reconstruction part (streams D) ===========================
    const TX* xr = reinterpret_cast<const TX*>(a.X) + (int64_t)row * a.ldx;
    const TD* dr = reinterpret_cast<const TD*>(a.decoded) + (int64_t)row * a.ld_dec;
    TD* gr = reinterpret_cast<TD*>(a.d_decoded) + (int64_t)row * a.ld_ddec;
    float racc = 0.f;
    if (p.vec_ok) {
      int j = lane * 8;
      // two 256-element slabs in flight per warp
      for (; j + 256 < D; j += 512) {
        float x0[8], d0[8], x1[8], d1[8], g0[8], g1[8];
        Vec8<TX>::load(xr + j, x0);
        Vec8<TD>::load(dr + j, d0);
        Vec8<TX>::load(xr + j + 256, x1);
        Vec8<TD>::load(dr + j + 256, d1);
        scale8<TX>(x0, p.xs);
        scale8<TX>(x1, p.xs);
        racc += recon8<INPUT>(x0, d0, g0, s_rec);
        racc += recon8<INPUT>(x1, d1, g1, s_rec);
        Vec8<TD>::store(gr + j, g0);
        Vec8<TD>::store(gr + j + 256, g1);
      }
      if (j < D) {
        float x0[8], d0[8], g0[8];
        Vec8<TX>::load(xr + j, x0);
        Vec8<TD>::load(dr + j, d0);
        scale8<TX>(x0, p.xs);
        racc += recon8<INPUT>(x0, d0, g0, s_rec);
        Vec8<TD>::store(gr + j, g0);
      }
    } else {
      for (int j = lane; j < D; j += 32) {
        float g0;
        racc += recon1<INPUT>(scale1<TX>(to_f32<TX>(xr[j]), p.xs), to_f32<TD>(dr[j]), s_rec, g0);
        gr[j] = from_f32<TD>(g0);
      }
    }
    for (int j = D + lane; j < a.ddec_cols; j += 32) gr[j] = from_f32<TD>(0.f);
    const float R = warp_sum(racc);

    // =========================== latent part ===========================
    float sum_lv = 0.f;
    for (int l = lane; l < L; l += 32) {
      float mu = (l == lane) ? mu0 : mean[l], lv = (l == lane) ? lv0 : lvp[l];
      mu_s[l] = mu;
      elv_s[l] = expf(lv);
      sum_lv += lv;
      if (mode == DMVAE_MODE_VADE) x1_s[l] = mu + expf(0.5f * lv) * a.eps[(int64_t)row * a.ld_eps + l];
    }
    sum_lv = warp_sum(sum_lv);
    if (mode == DMVAE_MODE_DMVAE_SAMPLED)
      for (int k = lane; k < K; k += 32) w_s[k] = a.zeta[(int64_t)row * a.ld_zeta + k];
    __syncwarp();

    float C = 0.f, Zk = 0.f;
    int amax = 0;
    if (mode != DMVAE_MODE_DMVAE_SAMPLED) {
      // ---- A_k (and the VaDE score s_k), lane <-> k ----
      float A[kKPL], sc[kKPL];
#pragma unroll
      for (int jk = 0; jk < kKPL; ++jk) {
        int k = lane + 32 * jk;
        A[jk] = 0.f;
        sc[jk] = -INFINITY;
        if (k < K) {
          const float* mk = tab_m + k * Ls;
          const float* ik = tab_b + k * Ls;
          float acc = 0.f, zacc = 0.f;
          for (int l = 0; l < L; ++l) {
            float dm = mu_s[l] - mk[l];
            acc += (elv_s[l] + dm * dm) * ik[l];
            if (mode == DMVAE_MODE_VADE) {
              float dz = x1_s[l] - mk[l];
              zacc += dz * dz * ik[l];
            }
          }
          A[jk] = sum_plv[k] - sum_lv - (float)L + acc;
          sc[jk] = (mode == DMVAE_MODE_VADE) ? -0.5f * (zacc + sum_plv[k]) : (jk == 0 ? logit0 : a.logits[(int64_t)row * a.ld_logits + k]);
        }
      }
      // ---- softmax over K (warp shuffles), argmax (first maximum wins) ----
      float mx = sc[0];
      int mi = lane;
#pragma unroll
      for (int jk = 1; jk < kKPL; ++jk)
        if (sc[jk] > mx) { mx = sc[jk]; mi = lane + 32 * jk; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float om = __shfl_xor_sync(0xffffffffu, mx, o);
        int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (om > mx || (om == mx && oi < mi)) { mx = om; mi = oi; }
      }
      amax = mi;
      float q[kKPL], den = 0.f;
#pragma unroll
      for (int jk = 0; jk < kKPL; ++jk) {
        q[jk] = (lane + 32 * jk < K) ? expf(sc[jk] - mx) : 0.f;
        den += q[jk];
      }
      den = warp_sum(den);
      const float inv_den = 1.f / den;
      float G[kKPL], qG = 0.f;
#pragma unroll
      for (int jk = 0; jk < kKPL; ++jk) {
        q[jk] *= inv_den;
        G[jk] = 0.f;
        if (lane + 32 * jk < K) {
          float lq = logf(q[jk] + kEps0);
          C += q[jk] * (lq + logK);                                        // priors.py:195-199
          float gC = lq + q[jk] / (q[jk] + kEps0) + logK;
          G[jk] = r * (gC + 0.5f * A[jk]);
          if (mode == DMVAE_MODE_VADE && a.d_gate_extra)       // another loss gated by gamma: same softmax Jacobian
            G[jk] += __ldg(a.d_gate_extra + (int64_t)row * a.ld_dge + lane + 32 * jk) / s;
          Zk += 0.5f * q[jk] * A[jk];
          qG += q[jk] * G[jk];
        }
      }
      C = warp_sum(C);
      Zk = warp_sum(Zk);
      qG = warp_sum(qG);
#pragma unroll
      for (int jk = 0; jk < kKPL; ++jk) {
        int k = lane + 32 * jk;
        if (k < K) {
          float dsc = s * q[jk] * (G[jk] - qG);          // d loss / d logits_k  (or d s_k for VaDE)
          w_s[k] = q[jk];
          a.qc[(int64_t)row * K + k] = q[jk];
          if (mode == DMVAE_MODE_VADE) {
            aux_s[k] = dsc;
            a.w_scratch[(int64_t)row * K + k] = dsc;
          } else if (a.dlogits_dtype == DMVAE_BF16) {
            reinterpret_cast<__nv_bfloat16*>(a.d_logits)[(int64_t)row * a.ld_dlogits + k] = __float2bfloat16_rn(dsc);
          } else {
            reinterpret_cast<float*>(a.d_logits)[(int64_t)row * a.ld_dlogits + k] = dsc;
          }
        }
      }
      __syncwarp();
      // ---- gradients wrt mean / log_var (and Z through gamma), lane <-> l ----
      for (int l = lane; l < L; l += 32) {
        float mu = mu_s[l], dmu = 0.f, wiv = 0.f, wsum = 0.f, dzg = 0.f;
        for (int k = 0; k < K; ++k) {
          float w = w_s[k], iv = tab_b[k * Ls + l], mk = tab_m[k * Ls + l];
          dmu += w * (mu - mk) * iv;
          wiv += w * iv;
          wsum += w;
          if (mode == DMVAE_MODE_VADE) dzg -= aux_s[k] * (x1_s[l] - mk) * iv;
        }
        a.d_mean_kl[(int64_t)row * a.ld_dkl + l] = s * r * dmu;
        a.d_log_var_kl[(int64_t)row * a.ld_dkl + l] = s * r * 0.5f * (elv_s[l] * wiv - wsum);
        if (mode == DMVAE_MODE_VADE) a.d_Z_gamma[(int64_t)row * a.ld_dzg + l] = dzg;
      }
    } else {
      // ---- cluster_sample=True: prior parameters mixed by zeta (priors.py:118-128) ----
      float zk = 0.f;
      for (int l = lane; l < L; l += 32) {
        float mbar = 0.f, pbar = 0.f;
        for (int k = 0; k < K; ++k) {
          float w = w_s[k];
          mbar += w * tab_m[k * Ls + l];
          pbar += w * tab_b[k * Ls + l];
        }
        float e = expf(-pbar), dm = mu_s[l] - mbar, lv = lvp[l];
        float t = (elv_s[l] + dm * dm) * e;
        zk += pbar - lv - 1.f + t;
        float ga = -dm * e, gc = 0.5f * (1.f - t);
        x1_s[l] = ga;
        x2_s[l] = gc;
        a.f_scratch[(int64_t)row * 2 * L + l] = ga;
        a.f_scratch[(int64_t)row * 2 * L + L + l] = gc;
        a.d_mean_kl[(int64_t)row * a.ld_dkl + l] = s * r * dm * e;
        a.d_log_var_kl[(int64_t)row * a.ld_dkl + l] = s * r * 0.5f * (elv_s[l] * e - 1.f);
      }
      Zk = 0.5f * warp_sum(zk);
      __syncwarp();
      float sc[kKPL], Gw[kKPL], zt[kKPL];
#pragma unroll
      for (int jk = 0; jk < kKPL; ++jk) {
        int k = lane + 32 * jk;
        sc[jk] = -INFINITY;
        Gw[jk] = 0.f;
        zt[jk] = 0.f;
        if (k < K) {
          sc[jk] = jk == 0 ? logit0 : a.logits[(int64_t)row * a.ld_logits + k];
          zt[jk] = w_s[k];
          float acc = 0.f;
          for (int l = 0; l < L; ++l) acc += x1_s[l] * tab_m[k * Ls + l] + x2_s[l] * tab_b[k * Ls + l];
          Gw[jk] = acc;
        }
      }
      float mx = sc[0];
      int mi = lane;
#pragma unroll
      for (int jk = 1; jk < kKPL; ++jk)
        if (sc[jk] > mx) { mx = sc[jk]; mi = lane + 32 * jk; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float om = __shfl_xor_sync(0xffffffffu, mx, o);
        int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (om > mx || (om == mx && oi < mi)) { mx = om; mi = oi; }
      }
      amax = mi;
      float q[kKPL], den = 0.f;
#pragma unroll
      for (int jk = 0; jk < kKPL; ++jk) {
        q[jk] = (lane + 32 * jk < K) ? expf(sc[jk] - mx) : 0.f;
        den += q[jk];
      }
      den = warp_sum(den);
      const float inv_den = 1.f / den;
      float gC[kKPL], qg = 0.f, zg = 0.f;
#pragma unroll
      for (int jk = 0; jk < kKPL; ++jk) {
        q[jk] *= inv_den;
        gC[jk] = 0.f;
        if (lane + 32 * jk < K) {
          float lq = logf(q[jk] + kEps0);
          C += q[jk] * (lq + logK);
          gC[jk] = lq + q[jk] / (q[jk] + kEps0) + logK;
          qg += q[jk] * gC[jk];
          zg += zt[jk] * Gw[jk];
        }
      }
      C = warp_sum(C);
      qg = warp_sum(qg);
      zg = warp_sum(zg);
      const float inv_tau = 1.f / a.tau;
#pragma unroll
      for (int jk = 0; jk < kKPL; ++jk) {
        int k = lane + 32 * jk;
        if (k < K) {
          float dsc = s * r * q[jk] * (gC[jk] - qg) + s * r * inv_tau * zt[jk] * (Gw[jk] - zg);
          a.qc[(int64_t)row * K + k] = q[jk];
          if (a.dlogits_dtype == DMVAE_BF16)
            reinterpret_cast<__nv_bfloat16*>(a.d_logits)[(int64_t)row * a.ld_dlogits + k] = __float2bfloat16_rn(dsc);
          else
            reinterpret_cast<float*>(a.d_logits)[(int64_t)row * a.ld_dlogits + k] = dsc;
        }
      }
    }
    if (mode != DMVAE_MODE_VADE && a.d_logits) {
      for (int k = K + lane; k < a.dlogits_cols; k += 32) {
        if (a.dlogits_dtype == DMVAE_BF16)
          reinterpret_cast<__nv_bfloat16*>(a.d_logits)[(int64_t)row * a.ld_dlogits + k] = __float2bfloat16_rn(0.f);
        else
          reinterpret_cast<float*>(a.d_logits)[(int64_t)row * a.ld_dlogits + k] = 0.f;
      }
    }
    if (lane == 0) {
      reinterpret_cast<float4*>(a.per_sample)[row] = make_float4(R, C, Zk, a.recon_scale * R + r * (C + Zk));
      a.argmax[row] = amax;
    }
    __syncwarp();
  }
}

// ---- shared-memory tile access and bulk async copies (row-tile kernel) ----
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "EW_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra.uni EW_DONE;\n"
      "bra.uni EW_LOOP;\n"
      "EW_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// 8 consecutive elements of a shared-memory tile <-> 8 floats
template <typename T>
struct Tile8;
template <>
struct Tile8<float> {
  static __device__ __forceinline__ void load(uint32_t a, float (&v)[8]) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a + 16));
  }
  static __device__ __forceinline__ void store(uint32_t a, const float (&v)[8]) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a + 16), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
  }
};
template <>
struct Tile8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(uint32_t a, float (&v)[8]) {
    uint32_t w[4];
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(a));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void store(uint32_t a, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
  }
};
template <>
struct Tile8<uint8_t> {
  static __device__ __forceinline__ void load(uint32_t a, float (&v)[8]) {
    uint32_t lo, hi;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a));
    // byte b -> float bits 0x4B0000bb = 2^23 + b, minus 2^23: exact, full-rate (no I2F on the XU pipe)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7540 + i)) - 8388608.f;
      v[4 + i] = __uint_as_float(__byte_perm(hi, 0x4B000000u, 0x7540 + i)) - 8388608.f;
    }
  }
};

// =================================================================================================
// Row-tile kernel for the common small-mixture case (DMVAE analytic KL, K*L small): a CTA owns 16 rows, 4 CTAs per SM.
//   warps 0-7 : stream the D-wide reconstruction part from a shared-memory tile (one bulk async copy per tensor in,
//               one out), 2 rows per warp treated as ONE list of 8-element chunks (no per-row tail iteration);
//               fp32, one MUFU (tanh) per element in the bf16 tier (recon8_t).
//   warps 8-9 : the latent part, 8 rows per warp and four lanes per row; for L <= 16 a register-resident single sweep
//               over the K components (rowtile_latent), else the generic two-pass loops.  They run beside the
//               streaming warps and only meet them at one named barrier (the row's reconstruction sum).
// Same formulas and outputs as elbo_kernel.
// =================================================================================================
constexpr int kFastRW = 2;                  // rows per reconstruction warp
constexpr int kFastRows = 8 * kFastRW;      // rows per CTA
constexpr int kLatWarps = 2;                // latent warps: 8 rows each, 4 lanes per row
constexpr int kFastThreads = 32 * (8 + kLatWarps);

// shared-memory plan of the row-tile kernel (floats)
struct RowTileSmem {
  int tab, sum_plv, ml, q, g, R, total;
  __host__ __device__ RowTileSmem(int L, int K) {
    int o = 0;
    tab = o; o += 2 * K * ((L + 3) & ~3);   // float2 (prior mean, exp(-prior log-variance)) [K][L | padded to 4 LQ]
    sum_plv = o; o += (K + 3) & ~3;
    ml = o; o += 2 * kFastRows * L;         // float2 (mean, exp(log_var)) [rows][L]
    q = o; o += (kFastRows * K + 3) & ~3;   // logits -> q(c|x), [rows][K]
    g = o; o += (kFastRows * K + 3) & ~3;   // G_k -> d_logits
    R = o; o += kFastRows;
    total = o;
  }
};

// fp32 reconstruction term of one element (fp32 decoder logits: the 1e-4 tier).  log1p(t), t = e^{-|d|} in (0,1],
// is a degree-7 polynomial (max abs error 2.4e-7) on the FMA pipe instead of lg2 on the XU pipe.
template <int INPUT>
__device__ __forceinline__ float recon1_rt(float x, float d, float s, float& g) {
  if (INPUT == DMVAE_INPUT_BINARY) {
    float t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-1.4426950408889634f * fabsf(d)));
    float pl = -0.008574675768613815f;
    pl = fmaf(pl, t, 0.044214192777872086f);
    pl = fmaf(pl, t, -0.10785368084907532f);
    pl = fmaf(pl, t, 0.17757023870944977f);
    pl = fmaf(pl, t, -0.2449961155653f);
    pl = fmaf(pl, t, 0.3327617645263672f);
    pl = fmaf(pl, t, -0.49997448921203613f);
    pl = fmaf(pl, t, 0.9999998211860657f);
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.f + t));
    const float sig = d >= 0.f ? inv : t * inv;
    g = s * (sig - x);
    return fmaf(pl, t, fmaf(-d, x, fmaxf(d, 0.f)));             // max(d,0) - d x + log1p(e^{-|d|})   (base_models.py:74-79)
  } else {
    const float df = d - x;                                     // base_models.py:80-83
    g = s * df;
    return 0.5f * df * df;
  }
}

// 8 targets of a shared-memory tile as floats, optionally centred (x - 0.5)
template <typename T, bool CENTRED>
__device__ __forceinline__ void tile8_x(uint32_t a, float (&x)[8], float xs, float xo) {
  Tile8<T>::load(a, x);
  scale8<T>(x, xs);
  if (CENTRED) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] -= 0.5f;
  }
}
template <>
__device__ __forceinline__ void tile8_x<uint8_t, true>(uint32_t a, float (&x)[8], float xs, float xo) {
  uint32_t lo, hi;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a));
  // byte b -> float bits 0x4700bb00 = 2^15 + b (ulp 2^-8); fma(2^15 + b, xs, -(2^15 xs + 0.5)) = b xs - 0.5 with ONE rounding
  // (exact for xs = 1): one PRMT + one FFMA per element whatever the scale
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[i] = fmaf(__uint_as_float(__byte_perm(lo, 0x47000000u, 0x7504 | (i << 4))), xs, xo);
    x[4 + i] = fmaf(__uint_as_float(__byte_perm(hi, 0x47000000u, 0x7504 | (i << 4))), xs, xo);
  }
}

// bf16 tier (bf16 decoder logits in, bf16 gradient out; tolerance 2e-2), 8 elements of one row.  Binary likelihood,
// with h = |d|/2, u = tanh(h) (ONE MUFU per element), xc = x - 1/2:
//   sigmoid(d) - x            = sign(d) u / 2 - xc
//   max(d,0) - d x            = h - d xc
//   log1p(e^{-|d|})           = ln 2 - ln(1 + u)          ->  sum_d = n ln 2 - ln prod_d (1 + u_d)
// so the log term costs one FFMA per element (a running product, <= 2^n) and one lg2 per lane and row.
// PRECISE keeps e^{-|d|} from a second MUFU (ex2, 2^-22) for the product instead of tanh's 2^-11.
// Everything is fp32: ~11 issue slots per element against ~19 for the earlier packed-half2 formulation (whose f16x2 MUFUs
// split into two per-half MUFUs + a PRMT anyway).
template <int INPUT, bool PRECISE>
__device__ __forceinline__ void recon8_t(const float (&xc)[8], const uint32_t (&dw)[4], uint32_t (&gw)[4], float s, float& acc,
                                         float& prod) {
  const float c = 0.5f * s, ms = -s;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float d2[2] = {__uint_as_float(dw[i] << 16), __uint_as_float(dw[i] & 0xffff0000u)};
    float g2[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float d = d2[j], x = xc[2 * i + j];
      if (INPUT == DMVAE_INPUT_BINARY) {
        const float h = 0.5f * fabsf(d);
        float ua;
        asm("tanh.approx.f32 %0, %1;" : "=f"(ua) : "f"(h));
        const float u = __uint_as_float((__float_as_uint(d) & 0x80000000u) | __float_as_uint(ua));
        g2[j] = fmaf(u, c, x * ms);                             // s (sigmoid(d) - x)
        acc += h;
        acc = fmaf(-d, x, acc);
        if (PRECISE) {
          float t;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(h * -2.8853900817779268f));
          prod = fmaf(prod, t, prod);
        } else {
          prod = fmaf(prod, ua, prod);
        }
      } else {
        const float df = d - x;                                 // base_models.py:80-83 (x is not centred here)
        g2[j] = s * df;
        acc = fmaf(0.5f * df, df, acc);
      }
    }
    __nv_bfloat162 gb = __floats2bfloat162_rn(g2[0], g2[1]);
    gw[i] = *reinterpret_cast<uint32_t*>(&gb);
  }
}

__device__ __forceinline__ float fast_exp(float x) {            // e^x, relative error ~2^-21 (ex2.approx on the XU pipe)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}

// d_logits row [K data | zeros to dlogits_cols], written by the four lanes of a row's quad
__device__ __forceinline__ void rowtile_store_dlogits(const dmvae_elbo_args& a, const float* g_r, int64_t grow, int h, int K) {
  if (a.dlogits_dtype == DMVAE_BF16 && (a.dlogits_cols & 7) == 0 && (a.ld_dlogits & 7) == 0) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.d_logits) + grow * a.ld_dlogits;
    for (int c8 = h << 3; c8 < a.dlogits_cols; c8 += 32) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c8 + j < K) ? g_r[c8 + j] : 0.f;
      Vec8<__nv_bfloat16>::store(out + c8, v);
    }
  } else {
    for (int c = h; c < a.dlogits_cols; c += 4) {
      const float v = c < K ? g_r[c] : 0.f;
      if (a.dlogits_dtype == DMVAE_BF16)
        reinterpret_cast<__nv_bfloat16*>(a.d_logits)[grow * a.ld_dlogits + c] = __float2bfloat16_rn(v);
      else
        reinterpret_cast<float*>(a.d_logits)[grow * a.ld_dlogits + c] = v;
    }
  }
}

// q(c|x) of a latent warp's rows: one flat coalesced copy
__device__ __forceinline__ void rowtile_store_q(const dmvae_elbo_args& a, const float* qs, int row_first, int n, int lane) {
  float* qo = a.qc + (int64_t)row_first * a.K;
  for (int i = lane; i < n; i += 32) qo[i] = qs[i];
}

// Latent part of the row-tile kernel for L <= 4 LQ (register-resident): lane (row, h) of a quad owns the latent
// dimensions l = h + 4 j, j < LQ (mean, exp(log_var), d_mean, sum_k w iv in registers) and ONE sweep over the K
// components produces both A_k (quad-reduced) and the KL-side gradients - q = softmax(logits) does not depend on A, so
// the two passes of the formulas (SURVEY 8a') fuse.  The prior table is zero-padded to 4 LQ columns: padded (l >= L)
// entries contribute exact zeros.  ~7 issue slots per (k, l) pair instead of ~20 for the generic two-pass loops.
template <int LQ, bool LO = false>
__device__ __forceinline__ void rowtile_latent(const dmvae_elbo_args& a, float* smem, const RowTileSmem& sm, int row0,
                                               int nrows_cta, int lw, int lane, float r, float s) {
  constexpr int Ls = 4 * LQ;
  constexpr int kLatThreads = 32 * kLatWarps;
  constexpr int kRowsPerLat = kFastRows / kLatWarps;
  const int L = a.L, K = a.K;
  float2* tab = reinterpret_cast<float2*>(smem + sm.tab);       // [K][Ls] (m, exp(-plv)), zero beyond L
  float* sum_plv = smem + sm.sum_plv;
  const float* R_s = smem + sm.R;
  const int lt = lw * 32 + lane;
  const int lr0 = lw * kRowsPerLat;
  const int nrows_l = max(0, min(kRowsPerLat, nrows_cta - lr0));
  const int rl = lr0 + (lane >> 2), h = lane & 3;
  const bool valid = (lane >> 2) < nrows_l;
  const int rsel = valid ? rl : 0;                              // idle lanes shadow the tile's first row (always present)
  const int64_t grow = row0 + rsel;
  float* q_r = smem + sm.q + rsel * K;
  float* g_r = smem + sm.g + rsel * K;
  for (int i = lt; i < K * Ls; i += kLatThreads) {
    const int k = i / Ls, l = i - k * Ls;
    float2 t = make_float2(0.f, 0.f);
    if (l < L) t = make_float2(__ldg(a.prior_means + k * L + l), fast_exp(-__ldg(a.prior_log_vars + k * L + l)));
    tab[i] = t;
  }
  for (int k = lt; k < K; k += kLatThreads) {
    float sacc = 0.f;
    for (int l = 0; l < L; ++l) sacc += __ldg(a.prior_log_vars + k * L + l);
    sum_plv[k] = sacc;
  }
  float mu[LQ], elv[LQ], sum_lv = 0.f;
  {
    const float* mp = a.mean + grow * a.ld_zh;
    const float* vp = a.log_var + grow * a.ld_zh;
    float lv[LQ];
#pragma unroll
    for (int j = 0; j < LQ; ++j) {
      const int l = h + 4 * j;
      mu[j] = l < L ? __ldg(mp + l) : 0.f;
      lv[j] = l < L ? __ldg(vp + l) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < LQ; ++j) {
      elv[j] = (h + 4 * j < L) ? fast_exp(lv[j]) : 0.f;
      sum_lv += lv[j];
    }
  }
  float mx = -INFINITY;
  int amax = K;
  {
    const float* lp = a.logits + grow * a.ld_logits;
#pragma unroll 4
    for (int k = h; k < K; k += 4) {
      const float sc = __ldg(lp + k);
      if (valid) q_r[k] = sc;
      if (sc > mx) { mx = sc; amax = k; }                       // first maximum wins
    }
  }
  auto quad_sum = [](float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
  };
  sum_lv = quad_sum(sum_lv);
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, amax, o);
    if (om > mx || (om == mx && oi < amax)) { mx = om; amax = oi; }
  }
  float den = 0.f;
  if (valid)
    for (int k = h; k < K; k += 4) {
      const float e = fast_exp(q_r[k] - mx);                    // own values: written by this thread above
      q_r[k] = e;
      den += e;
    }
  den = quad_sum(den);
  const float inv_den = __fdividef(1.f, den);
  asm volatile("bar.sync 2, %0;" ::"n"(kLatThreads) : "memory");              // prior tables and exp(logit - max) complete
  // ---- one sweep over the components ----
  float dmu[LQ], wiv[LQ];
#pragma unroll
  for (int j = 0; j < LQ; ++j) dmu[j] = wiv[j] = 0.f;
  {
    const float2* tk = tab + h;
#pragma unroll 2
    for (int k = 0; k < K; ++k, tk += Ls) {
      const float qk = q_r[k] * inv_den;
      float accA = 0.f;
#pragma unroll
      for (int j = 0; j < LQ; ++j) {
        const float2 t = tk[4 * j];
        const float dm = mu[j] - t.x;
        accA = fmaf(fmaf(dm, dm, elv[j]), t.y, accA);
        const float wi = qk * t.y;
        wiv[j] += wi;
        dmu[j] = fmaf(wi, dm, dmu[j]);
      }
      accA = quad_sum(accA);
      if (valid && h == (k & 3)) g_r[k] = accA;                 // read back below by the same lane
    }
  }
  __syncwarp();                                                 // the sweep's reads of exp(logit - max) are done
  const float logK = __logf((float)K);
  float C = 0.f, Zk = 0.f, qG = 0.f, wsum = 0.f;
  if (valid)
    for (int k = h; k < K; k += 4) {
      const float q = q_r[k] * inv_den;
      const float A = sum_plv[k] - sum_lv - (float)L + g_r[k];
      const float lq = __logf(q + kEps0);
      C = fmaf(q, lq + logK, C);                                // priors.py:195-199
      const float gC = lq + __fdividef(q, q + kEps0) + logK;
      const float G = r * fmaf(0.5f, A, gC);
      Zk = fmaf(0.5f * q, A, Zk);
      qG = fmaf(q, G, qG);
      wsum += q;
      q_r[k] = q;
      g_r[k] = G;
    }
  C = quad_sum(C);
  Zk = quad_sum(Zk);
  qG = quad_sum(qG);
  wsum = quad_sum(wsum);
  if (valid)
    for (int k = h; k < K; k += 4) g_r[k] = s * q_r[k] * (g_r[k] - qG);       // d loss / d logits_k
  __syncwarp();                                                 // q, d_logits of all four lanes visible
  if (valid) {
    float* dmp = a.d_mean_kl + grow * a.ld_dkl;
    float* dvp = a.d_log_var_kl + grow * a.ld_dkl;
    const float sr = s * r;
#pragma unroll
    for (int j = 0; j < LQ; ++j) {
      const int l = h + 4 * j;
      if (l < L) {
        dmp[l] = sr * dmu[j];
        dvp[l] = sr * 0.5f * fmaf(elv[j], wiv[j], -wsum);
      }
    }
    rowtile_store_dlogits(a, g_r, grow, h, K);
  }
  rowtile_store_q(a, smem + sm.q + lr0 * K, row0 + lr0, nrows_l * K, lane);
  if (!LO) asm volatile("bar.sync 1, %0;" ::"n"(kFastThreads) : "memory");    // every slab's R_s is written
  if (valid && h == 0) {
    const float R = LO ? 0.f : R_s[rl];       // latent-only launch: dmvae_elbo_reduce completes .x and .w from r_part
    reinterpret_cast<float4*>(a.per_sample)[row0 + rl] = make_float4(R, C, Zk, a.recon_scale * R + r * (C + Zk));
    a.argmax[row0 + rl] = amax;
  }
}

// The latent warps of the row-tile kernel alone (dmvae_elbo_args.r_part: the reconstruction term lives in the output
// layer's GEMM epilogue).  64 threads and a few KB of shared memory per 16-row tile: these CTAs fit beside the resident
// GEMM CTAs of the decoder's forward pass, which this launch overlaps.
__global__ void __launch_bounds__(32 * kLatWarps) elbo_rowtile_latent_kernel(const ElboParams p) {
  extern __shared__ __align__(128) float smem[];
  const dmvae_elbo_args& a = p.a;
  const RowTileSmem sm(a.L, a.K);
  const int lw = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kFastRows;
  const int nrows_cta = min(kFastRows, a.rows - row0);
  pdl_wait();
  pdl_launch_dependents();
  const float r = a.kl_ratio_dev ? __ldg(a.kl_ratio_dev) : a.kl_ratio;
  const float s = a.inv_global_batch;
  const int lq = (a.L + 3) >> 2;
  if (lq == 1) rowtile_latent<1, true>(a, smem, sm, row0, nrows_cta, lw, lane, r, s);
  else if (lq == 2) rowtile_latent<2, true>(a, smem, sm, row0, nrows_cta, lw, lane, r, s);
  else if (lq == 3) rowtile_latent<3, true>(a, smem, sm, row0, nrows_cta, lw, lane, r, s);
  else rowtile_latent<4, true>(a, smem, sm, row0, nrows_cta, lw, lane, r, s);
}

template <typename TX, typename TD, int INPUT, bool PRECISE>
__global__ void __launch_bounds__(kFastThreads, 4) elbo_rowtile_kernel(const ElboParams p) {
  extern __shared__ __align__(128) float smem[];
  constexpr bool FAST = sizeof(TD) == 2;
  constexpr bool CENTRED = FAST && INPUT == DMVAE_INPUT_BINARY;   // the bf16-tier binary path works on x - 1/2
  const dmvae_elbo_args& a = p.a;
  const int L = a.L, K = a.K, D = a.D;
  const RowTileSmem sm(L, K);
  float* R_s = smem + sm.R;                   // [rows]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kFastRows;
  const int nrows_cta = min(kFastRows, a.rows - row0);
  pdl_wait();                                   // everything below reads the previous kernels' outputs
  pdl_launch_dependents();
  const float r = a.kl_ratio_dev ? __ldg(a.kl_ratio_dev) : a.kl_ratio;
  const float s = a.inv_global_batch, s_rec = a.inv_global_batch * a.recon_scale;

  if (warp < 8) {
    // =========================== reconstruction warps ===========================
    // bulk-load this warp's slab (targets, decoder logits) into shared memory, transform the logits into the
    // gradient in place, bulk-store the gradient rows.  These warps never wait for the latent warps.
    uint8_t* tiles = reinterpret_cast<uint8_t*>(smem) + (((size_t)sm.total * 4 + 127) & ~(size_t)127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles);                      // 8 mbarriers (one per slab)
    const uint32_t xt = smem_addr_u32(tiles + 128);
    const uint32_t dt = xt + (uint32_t)(kFastRows * p.xrow);
    const int rw0 = row0 + warp * kFastRW;
    const int nrows_w = max(0, min(kFastRW, a.rows - rw0));
    const uint32_t xt_w = xt + (uint32_t)(warp * kFastRW * p.xrow), dt_w = dt + (uint32_t)(warp * kFastRW * p.drow);
    // contiguous tiles (row pitch == row bytes) move as ONE bulk copy per tensor and CTA: the TMA unit's cost is per
    // request (~150 cycles measured), so 48 row-sized copies per CTA had made it the bottleneck
    // ... and as TWO halves (rows 0-7 for warps 0-3, rows 8-15 for warps 4-7), each with its own mbarrier and its own
    // store: the first half's warps start computing while the second half is still in flight, and its gradient rows
    // leave while the second half is still being computed.
    const bool whole = p.contig != 0;
    constexpr int kHalfRows = kFastRows / 2;
    const int half = warp >> 2;
    const int nrows_h0 = min(nrows_cta, kHalfRows), nrows_h1 = nrows_cta - nrows_h0;
    const uint32_t bar_w = smem_addr_u32(&bars[whole ? half : warp]);
    if (lane == 0 && (!whole || warp == 0)) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_w), "r"(1));
      if (whole) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(&bars[1])), "r"(1));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const uint32_t dbytes = (uint32_t)(D * (int)sizeof(TD));
      if (whole) {
        const uint32_t xb = (uint32_t)(nrows_h0 * p.xrow), db = (uint32_t)(nrows_h0 * p.drow);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_w), "r"(xb + db) : "memory");
        bulk_g2s(xt, reinterpret_cast<const uint8_t*>(a.X) + (int64_t)row0 * p.xrow, xb, bar_w);
        bulk_g2s(dt, reinterpret_cast<const uint8_t*>(a.decoded) + (int64_t)row0 * p.drow, db, bar_w);
        if (nrows_h1 > 0) {
          const uint32_t bar1 = smem_addr_u32(&bars[1]);
          const uint32_t xb1 = (uint32_t)(nrows_h1 * p.xrow), db1 = (uint32_t)(nrows_h1 * p.drow);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar1), "r"(xb1 + db1) : "memory");
          bulk_g2s(xt + xb, reinterpret_cast<const uint8_t*>(a.X) + (int64_t)(row0 + kHalfRows) * p.xrow, xb1, bar1);
          bulk_g2s(dt + db, reinterpret_cast<const uint8_t*>(a.decoded) + (int64_t)(row0 + kHalfRows) * p.drow, db1, bar1);
        }
      } else {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_w),
                     "r"((uint32_t)nrows_w * ((uint32_t)p.xrow + dbytes))
                     : "memory");
        for (int q = 0; q < nrows_w; ++q) {
          bulk_g2s(xt_w + (uint32_t)(q * p.xrow),
                   reinterpret_cast<const uint8_t*>(a.X) + (int64_t)(rw0 + q) * a.ldx * (int64_t)sizeof(TX), (uint32_t)p.xrow, bar_w);
          bulk_g2s(dt_w + (uint32_t)(q * p.drow),
                   reinterpret_cast<const uint8_t*>(a.decoded) + (int64_t)(rw0 + q) * a.ld_dec * (int64_t)sizeof(TD), dbytes, bar_w);
        }
      }
    }
    if (whole) asm volatile("bar.sync 3, 256;" ::: "memory");   // the barrier words are initialised before anyone polls them
    else __syncwarp();
    if (nrows_w > 0) mbar_wait_parity(bar_w, 0);
    // The warp's two rows are ONE list of 8-element chunks, chunk c = lane + 32 i (no per-row tail iteration).  A lane's
    // chunks are ordered by row, so the row sums are plain per-lane accumulators that are parked once, at the
    // iteration where the lane crosses into the second row.
    static_assert(kFastRW == 2, "the reconstruction loop parks exactly one row");
    const int cpr = D >> 3;                                     // 8-element chunks per row (D % 8 == 0)
    const int total = nrows_w * cpr;
    const uint32_t dgap = (uint32_t)(p.drow - D * (int)sizeof(TD));           // the targets tile has no row gap
    float acc = 0.f, prod = 1.f, acc0 = 0.f, prod0 = 1.f;
    bool second = false;
    uint32_t xa = xt_w + (uint32_t)(lane * 8 * (int)sizeof(TX));
    uint32_t da = dt_w + (uint32_t)(lane * 8 * (int)sizeof(TD));
    // three phases with ONE copy of the loop body: iterations where every lane is in the first row, the mixed
    // iteration, the rest (second row); lanes park between phases, so the body carries no row logic
    const int nA32 = (cpr >> 5) << 5;
    int c = lane;
#pragma unroll 1
    for (int ph = 0; ph < 3; ++ph) {
      const int cstop = min(total, ph == 0 ? nA32 : (ph == 1 ? nA32 + 32 : total));
      if (c >= cpr && !second) {
        second = true;
        acc0 = acc; prod0 = prod;
        acc = 0.f; prod = 1.f;
        da += dgap;
      }
#pragma unroll 1
    for (; c < cstop; c += 32, xa += 256u * (uint32_t)sizeof(TX), da += 256u * (uint32_t)sizeof(TD)) {
      if (FAST) {
        float x[8];
        uint32_t dw[4], gw[4];
        tile8_x<TX, CENTRED>(xa, x, p.xs, p.xo);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(dw[0]), "=r"(dw[1]), "=r"(dw[2]), "=r"(dw[3]) : "r"(da));
        recon8_t<INPUT, PRECISE>(x, dw, gw, s_rec, acc, prod);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(da), "r"(gw[0]), "r"(gw[1]), "r"(gw[2]), "r"(gw[3]) : "memory");
      } else {
        float x[8], d[8], g[8];
        Tile8<TX>::load(xa, x);
        scale8<TX>(x, p.xs);
        Tile8<TD>::load(da, d);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += recon1_rt<INPUT>(x[i], d[i], s_rec, g[i]);
        Tile8<TD>::store(da, g);
      }
    }
    }
    if (!second) { acc0 = acc; prod0 = prod; acc = 0.f; prod = 1.f; }         // the lane never reached the second row
#pragma unroll
    for (int q = 0; q < kFastRW; ++q) {
      float av = q == 0 ? acc0 : acc;
      if (CENTRED) {                                            // + n ln 2 - ln prod(1 + u)   |   + ln prod(1 + t)
        float lp;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lp) : "f"(q == 0 ? prod0 : prod));
        av = fmaf(lp, PRECISE ? 0.6931471805599453f : -0.6931471805599453f, av);
      }
      float R = warp_sum(av);
      if (CENTRED && !PRECISE) R = fmaf((float)D, 0.6931471805599453f, R);
      if (lane == 0) R_s[warp * kFastRW + q] = R;
    }
    // zero the padding columns [D, ddec_cols) of the gradient rows (operand of the decoder's dgrad / wgrad GEMMs)
    const int padw = (p.drow - D * (int)sizeof(TD)) >> 2;       // 32-bit words of padding per row
    for (int q = 0; q < nrows_w; ++q)
      for (int j = lane; j < padw; j += 32)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dt_w + (uint32_t)(q * p.drow + D * (int)sizeof(TD) + 4 * j)), "r"(0u) : "memory");
    asm volatile("bar.arrive 1, %0;" ::"n"(kFastThreads) : "memory");         // R_s written (latent warps wait on it)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (whole) {
      // every slab of this half is final (named barrier 4 / 5, the half's four warps)
      if (half == 0) asm volatile("bar.sync 4, 128;" ::: "memory");
      else asm volatile("bar.sync 5, 128;" ::: "memory");
      const int nrows_h = half == 0 ? nrows_h0 : nrows_h1;
      if ((warp & 3) == 0 && lane == 0 && nrows_h > 0) {
        const int64_t r0h = row0 + half * kHalfRows;
        bulk_s2g(reinterpret_cast<uint8_t*>(a.d_decoded) + r0h * p.drow, dt + (uint32_t)(half * kHalfRows * p.drow),
                 (uint32_t)(nrows_h * p.drow));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    } else {
      __syncwarp();
      if (lane == 0) {
        for (int q = 0; q < nrows_w; ++q)
          bulk_s2g(reinterpret_cast<uint8_t*>(a.d_decoded) + (int64_t)(rw0 + q) * a.ld_ddec * (int64_t)sizeof(TD),
                   dt_w + (uint32_t)(q * p.drow), (uint32_t)p.drow);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // the slab is read by the TMA unit until then
      }
    }
    return;
  }

  // =========================== latent warps: 8 rows each, four lanes per row ===========================
  // Lane (row, h) owns the components k = h, h+4, ... in the first pass and the latent dimensions l = h, h+4, ... in
  // the second; the row's inputs go global -> registers -> shared without index arithmetic, the gradients go from
  // registers straight to global.  exp / log are single MUFU ops (ex2.approx / lg2.approx, ~2^-21 relative).
  const int lw = warp - 8;
  if (L <= 16) {                                                // register-resident fused sweep
    const int lq = (L + 3) >> 2;
    if (lq == 1) rowtile_latent<1>(a, smem, sm, row0, nrows_cta, lw, lane, r, s);
    else if (lq == 2) rowtile_latent<2>(a, smem, sm, row0, nrows_cta, lw, lane, r, s);
    else if (lq == 3) rowtile_latent<3>(a, smem, sm, row0, nrows_cta, lw, lane, r, s);
    else rowtile_latent<4>(a, smem, sm, row0, nrows_cta, lw, lane, r, s);
    return;
  }
  // generic two-pass form (any L <= 64)
  float2* tab = reinterpret_cast<float2*>(smem + sm.tab);       // [K][L] (m, exp(-plv))
  float* sum_plv = smem + sm.sum_plv;                           // [K]
  const int lt = lw * 32 + lane;                                // thread index among the latent warps
  constexpr int kLatThreads = 32 * kLatWarps;
  constexpr int kRowsPerLat = kFastRows / kLatWarps;            // 8
  const int lr0 = lw * kRowsPerLat;                             // first tile row of this warp
  const int nrows_l = max(0, min(kRowsPerLat, nrows_cta - lr0));
  const int rl = lr0 + (lane >> 2), h = lane & 3;
  const bool valid = (lane >> 2) < nrows_l;
  const int rsel = valid ? rl : 0;                              // idle lanes shadow the tile's first row (always present)
  const int64_t grow = row0 + rsel;
  float2* ml_r = reinterpret_cast<float2*>(smem + sm.ml) + rsel * L;
  float* q_r = smem + sm.q + rsel * K;
  float* g_r = smem + sm.g + rsel * K;
  for (int i = lt; i < K * L; i += kLatThreads)
    tab[i] = make_float2(__ldg(a.prior_means + i), fast_exp(-__ldg(a.prior_log_vars + i)));
  for (int k = lt; k < K; k += kLatThreads) {
    float sacc = 0.f;
    for (int l = 0; l < L; ++l) sacc += __ldg(a.prior_log_vars + k * L + l);
    sum_plv[k] = sacc;
  }
  float sum_lv = 0.f;
  {
    const float* mp = a.mean + grow * a.ld_zh;
    const float* vp = a.log_var + grow * a.ld_zh;
#pragma unroll 4
    for (int l = h; l < L; l += 4) {
      const float mu = __ldg(mp + l), lv = __ldg(vp + l);
      if (valid) ml_r[l] = make_float2(mu, fast_exp(lv));
      sum_lv += lv;
    }
  }
  float mx = -INFINITY;
  int amax = K;
  {
    const float* lp = a.logits + grow * a.ld_logits;
#pragma unroll 4
    for (int k = h; k < K; k += 4) {
      const float sc = __ldg(lp + k);
      if (valid) q_r[k] = sc;
      if (sc > mx) { mx = sc; amax = k; }                       // first maximum wins
    }
  }
  auto quad_sum = [](float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
  };
  sum_lv = quad_sum(sum_lv);
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, amax, o);
    if (om > mx || (om == mx && oi < amax)) { mx = om; amax = oi; }
  }
  float den = 0.f;
  if (valid)
    for (int k = h; k < K; k += 4) {
      const float e = fast_exp(q_r[k] - mx);                    // own values: written by this thread above
      q_r[k] = e;
      den += e;
    }
  den = quad_sum(den);
  const float inv_den = __fdividef(1.f, den);
  asm volatile("bar.sync 2, %0;" ::"n"(kLatThreads) : "memory");              // prior tables and row inputs complete
  const float logK = __logf((float)K);
  float C = 0.f, Zk = 0.f, qG = 0.f, wsum = 0.f;
  if (valid)
    for (int k = h; k < K; k += 4) {
      const float q = q_r[k] * inv_den;
      const float2* tk = tab + k * L;
      float acc = 0.f;
#pragma unroll 2
      for (int l = 0; l < L; ++l) {
        const float2 t = tk[l], m = ml_r[l];
        const float dm = m.x - t.x;
        acc = fmaf(fmaf(dm, dm, m.y), t.y, acc);
      }
      const float A = sum_plv[k] - sum_lv - (float)L + acc;
      const float lq = __logf(q + kEps0);
      C = fmaf(q, lq + logK, C);                                // priors.py:195-199
      const float gC = lq + __fdividef(q, q + kEps0) + logK;
      const float G = r * fmaf(0.5f, A, gC);
      Zk = fmaf(0.5f * q, A, Zk);
      qG = fmaf(q, G, qG);
      wsum += q;
      q_r[k] = q;
      g_r[k] = G;
    }
  C = quad_sum(C);
  Zk = quad_sum(Zk);
  qG = quad_sum(qG);
  wsum = quad_sum(wsum);
  if (valid)
    for (int k = h; k < K; k += 4) g_r[k] = s * q_r[k] * (g_r[k] - qG);       // d loss / d logits_k
  __syncwarp();                                                 // q, d_logits of all four lanes visible
  if (valid) {
    float* dmp = a.d_mean_kl + grow * a.ld_dkl;
    float* dvp = a.d_log_var_kl + grow * a.ld_dkl;
    const float sr = s * r;
    for (int l = h; l < L; l += 4) {
      const float2 m = ml_r[l];
      float dmu = 0.f, wiv = 0.f;
#pragma unroll 2
      for (int k = 0; k < K; ++k) {
        const float2 t = tab[k * L + l];
        const float wi = q_r[k] * t.y;
        wiv += wi;
        dmu = fmaf(wi, m.x - t.x, dmu);
      }
      dmp[l] = sr * dmu;
      dvp[l] = sr * 0.5f * fmaf(m.y, wiv, -wsum);
    }
    rowtile_store_dlogits(a, g_r, grow, h, K);
  }
  rowtile_store_q(a, smem + sm.q + lr0 * K, row0 + lr0, nrows_l * K, lane);
  asm volatile("bar.sync 1, %0;" ::"n"(kFastThreads) : "memory");             // every slab's R_s is written
  if (valid && h == 0) {
    const float R = R_s[rl];
    reinterpret_cast<float4*>(a.per_sample)[row0 + rl] = make_float4(R, C, Zk, a.recon_scale * R + r * (C + Zk));
    a.argmax[row0 + rl] = amax;
  }
}

size_t elbo_rowtile_smem_bytes(const ElboParams& p) {
  const size_t fl = (sizeof(float) * (size_t)RowTileSmem(p.a.L, p.a.K).total + 127) & ~(size_t)127;
  return fl + 128 + (size_t)kFastRows * (size_t)(p.xrow + p.drow);
}

bool elbo_rowtile_ok(const ElboParams& p) {
  static int enabled = -1;                    // DMVAE_ELBO_ROWTILE=0 forces the warp-per-row kernel (A/B measurements)
  if (enabled < 0) {
    const char* e = getenv("DMVAE_ELBO_ROWTILE");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled) return false;
  const dmvae_elbo_args& a = p.a;
  return a.mode == DMVAE_MODE_DMVAE && p.vec_ok && p.bulk_ok && a.K * a.L <= 512 && a.K <= 64 && a.L <= 64 &&
         ((uintptr_t)a.d_logits & 15) == 0 && elbo_rowtile_smem_bytes(p) <= 100 * 1024;
}

template <typename TX, typename TD, int INPUT>
int launch_elbo_rowtile(dmvae_ctx* ctx, const ElboParams& p, cudaStream_t st) {
  const size_t smem = elbo_rowtile_smem_bytes(p);
  static int precise = -1;                    // DMVAE_ELBO_PRECISE=1: second MUFU (ex2) for the log term of the bf16 tier
  if (precise < 0) {
    const char* e = getenv("DMVAE_ELBO_PRECISE");
    precise = (e && e[0] == '1') ? 1 : 0;
  }
  constexpr bool kHasPrecise = sizeof(TD) == 2 && INPUT == DMVAE_INPUT_BINARY;
  auto kern = elbo_rowtile_kernel<TX, TD, INPUT, false>;
  if (kHasPrecise && precise) kern = elbo_rowtile_kernel<TX, TD, INPUT, kHasPrecise>;
  const int blocks = (p.a.rows + kFastRows - 1) / kFastRows;
  if (smem > 48 * 1024) DMVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dmvae_launch(kern, dim3(blocks), dim3(kFastThreads), smem, st, true, p);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

size_t elbo_smem_bytes(int L, int K, int Ls) {
  const int Lp = (L + 3) & ~3, Kp = (K + 3) & ~3;
  return sizeof(float) * (size_t)(2 * K * Ls + Kp + kWarps * (4 * Lp + 2 * Kp));
}

template <typename TX, typename TD, int INPUT, int KPL>
int launch_elbo_k(dmvae_ctx* ctx, const ElboParams& p, cudaStream_t st) {
  const size_t smem = elbo_smem_bytes(p.a.L, p.a.K, p.Ls);
  auto kern = elbo_kernel<TX, TD, INPUT, KPL>;
  if (smem > 48 * 1024) DMVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 8;
  if (smem > 0) per_sm = (int)min((size_t)8, (size_t)(220 * 1024) / max(smem, (size_t)1));
  if (per_sm < 1) per_sm = 1;
  int blocks = min(ctx->sm_count * per_sm, (p.a.rows + kWarps - 1) / kWarps);
  dmvae_launch(kern, dim3(blocks), dim3(kThreads), smem, st, true, p);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

#include "elbo_mma.cuh"

template <typename TX, typename TD, int INPUT>
int launch_elbo(dmvae_ctx* ctx, const ElboParams& p, cudaStream_t st) {
  if (elbo_rowtile_ok(p)) return launch_elbo_rowtile<TX, TD, INPUT>(ctx, p, st);
  if (elbo_mma_ok(p)) return launch_elbo_mma<TX, TD, INPUT>(ctx, p, st);
  if (p.a.K <= 32) return launch_elbo_k<TX, TD, INPUT, 1>(ctx, p, st);
  if (p.a.K <= 64) return launch_elbo_k<TX, TD, INPUT, 2>(ctx, p, st);
  return launch_elbo_k<TX, TD, INPUT, 4>(ctx, p, st);
}

// dmvae_elbo_args.r_part: latent part only (the reconstruction term is the output-layer GEMM's, dmvae_recon_fuse)
int launch_elbo_latent_only(dmvae_ctx* ctx, const ElboParams& p, cudaStream_t st) {
  const dmvae_elbo_args& a = p.a;
  static int rt_enabled = -1;
  if (rt_enabled < 0) {
    const char* e = getenv("DMVAE_ELBO_ROWTILE");
    rt_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (rt_enabled && a.mode == DMVAE_MODE_DMVAE && a.L <= 16 && a.K <= 64 && a.K * a.L <= 512 && ((uintptr_t)a.d_logits & 15) == 0) {
    const size_t smem = sizeof(float) * (size_t)RowTileSmem(a.L, a.K).total;
    const int blocks = (a.rows + kFastRows - 1) / kFastRows;
    dmvae_launch(elbo_rowtile_latent_kernel, dim3(blocks), dim3(32 * kLatWarps), smem, st, true, p);
    DMVAE_LAUNCH_CHECK(ctx);
    return DMVAE_OK;
  }
  DMVAE_CHECK_ARG(elbo_mma_ok(p), "elbo: latent-only launch (r_part) is not available for mode %d, L=%d, K=%d", a.mode, a.L, a.K);
  return launch_elbo_mma_latent(ctx, p, st);
}

int check_elbo_args(const dmvae_elbo_args* a) {
  DMVAE_CHECK_ARG(a != nullptr, "elbo: args is NULL");
  DMVAE_CHECK_ARG(a->mode >= 0 && a->mode <= 2, "elbo: unknown mode %d", a->mode);
  DMVAE_CHECK_ARG(a->input_type == DMVAE_INPUT_BINARY || a->input_type == DMVAE_INPUT_REAL,
                  "elbo: input_type %d not implemented (binary | real)", a->input_type);   // base_models.py:84-85
  DMVAE_CHECK_ARG(a->rows >= 0 && a->D > 0 && a->L > 0 && a->K > 0, "elbo: bad sizes rows=%d D=%d L=%d K=%d", a->rows, a->D, a->L, a->K);
  DMVAE_CHECK_ARG(a->K <= kMaxK, "elbo: K=%d exceeds the supported maximum %d", a->K, kMaxK);
  const bool latent_only = a->r_part != nullptr;
  DMVAE_CHECK_ARG(!latent_only || a->r_parts > 0, "elbo: r_part needs r_parts > 0");
  DMVAE_CHECK_ARG(a->mean && a->log_var && a->prior_means && a->prior_log_vars, "elbo: NULL input");
  DMVAE_CHECK_ARG(latent_only || (a->X && a->decoded), "elbo: NULL input");
  DMVAE_CHECK_ARG(a->per_sample && a->qc && a->argmax && a->d_mean_kl && a->d_log_var_kl && (latent_only || a->d_decoded), "elbo: NULL output");
  DMVAE_CHECK_ARG(((uintptr_t)a->per_sample & 15) == 0, "elbo: per_sample must be 16-byte aligned");
  DMVAE_CHECK_ARG(latent_only || (a->ldx >= a->D && a->ld_dec >= a->D && a->ld_ddec >= a->D && a->ddec_cols <= a->ld_ddec),
                  "elbo: leading dimensions too small");
  if (a->mode == DMVAE_MODE_VADE)
    DMVAE_CHECK_ARG(a->eps && a->d_Z_gamma && a->w_scratch, "elbo(VADE): eps, d_Z_gamma and w_scratch are required");
  else
    DMVAE_CHECK_ARG(a->logits && a->d_logits && a->dlogits_cols <= a->ld_dlogits && a->ld_dlogits >= a->K, "elbo: logits / d_logits required");
  if (a->mode == DMVAE_MODE_DMVAE_SAMPLED)
    DMVAE_CHECK_ARG(a->zeta && a->f_scratch && a->tau > 0.f, "elbo(SAMPLED): zeta, f_scratch and tau > 0 are required");
  DMVAE_CHECK_ARG(a->dec_dtype == DMVAE_F32 || a->dec_dtype == DMVAE_BF16, "elbo: dec_dtype must be f32 or bf16");
  return DMVAE_OK;
}

}  // namespace

extern "C" int dmvae_elbo_fwd_bwd(dmvae_ctx* ctx, const dmvae_elbo_args* a, void* stream) {
  DMVAE_CHECK_ARG(ctx != nullptr, "elbo: ctx is NULL");
  int rc = check_elbo_args(a);
  if (rc) return rc;
  if (a->rows == 0) return DMVAE_OK;
  ElboParams p;
  p.a = *a;
  p.xs = a->x_scale == 0.f ? 1.f : a->x_scale;
  p.xo = -(32768.f * p.xs + 0.5f);
  p.Ls = a->L | 1;
  const size_t xs = dmvae_dtype_size(a->x_dtype), ds = dmvae_dtype_size(a->dec_dtype);
  p.vec_ok = (a->D % 8 == 0) && (a->ldx % 8 == 0) && (a->ld_dec % 8 == 0) && (a->ld_ddec % 8 == 0) &&
             (((uintptr_t)a->X) % (8 * xs) == 0) && (((uintptr_t)a->decoded) % (8 * ds) == 0) &&
             (((uintptr_t)a->d_decoded) % (8 * ds) == 0);
  p.xrow = (int)(a->D * xs);
  p.drow = (int)(a->ddec_cols * ds);
  p.bulk_ok = (a->D * xs) % 16 == 0 && (a->ldx * xs) % 16 == 0 && (a->D * ds) % 16 == 0 && (a->ld_dec * ds) % 16 == 0 &&
              (a->ddec_cols * ds) % 16 == 0 && (a->ld_ddec * ds) % 16 == 0 && ((uintptr_t)a->X & 15) == 0 &&
              ((uintptr_t)a->decoded & 15) == 0 && ((uintptr_t)a->d_decoded & 15) == 0 && a->ddec_cols >= a->D;
  p.contig = (int64_t)(a->ldx * xs) == p.xrow && (int64_t)(a->ld_dec * ds) == p.drow && (int64_t)(a->ld_ddec * ds) == p.drow;
  DMVAE_CHECK_ARG(elbo_mma_ok(p) || elbo_smem_bytes(a->L, a->K, p.Ls) <= 220 * 1024,
                  "elbo: K*L = %d too large for the shared prior tables", a->K * a->L);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->r_part) return launch_elbo_latent_only(ctx, p, st);
#define GO(TX, TD)                                                                            \
  return a->input_type == DMVAE_INPUT_BINARY ? launch_elbo<TX, TD, DMVAE_INPUT_BINARY>(ctx, p, st) \
                                             : launch_elbo<TX, TD, DMVAE_INPUT_REAL>(ctx, p, st)
  if (a->dec_dtype == DMVAE_F32) {
    if (a->x_dtype == DMVAE_F32) GO(float, float);
    if (a->x_dtype == DMVAE_U8) GO(uint8_t, float);
    if (a->x_dtype == DMVAE_BF16) GO(__nv_bfloat16, float);
  } else {
    if (a->x_dtype == DMVAE_F32) GO(float, __nv_bfloat16);
    if (a->x_dtype == DMVAE_U8) GO(uint8_t, __nv_bfloat16);
    if (a->x_dtype == DMVAE_BF16) GO(__nv_bfloat16, __nv_bfloat16);
  }
#undef GO
  dmvae_set_error("elbo: unsupported x_dtype %d", a->x_dtype);
  return DMVAE_ERR_INVALID;
}

// =================================================================================================
// cross-sample reductions: U[set][k][f] = sum_b W[b,k] * F[b,f], f in [0, 2L] (f = 2L is the constant 1),
// plus the loss terms; two deterministic stages (block partials, then a fixed-order final sum).
// =================================================================================================
namespace {

constexpr int kRedThreads = 256;
constexpr int kRedMaxChunk = 64;
constexpr size_t kRedSmemBudget = 96 * 1024;

inline int reduce_chunk(int L, int K) {
  size_t row_bytes = sizeof(float) * (size_t)(K + 2 * L + 1);
  int c = (int)(kRedSmemBudget / row_bytes);
  return max(1, min(kRedMaxChunk, c));
}

__global__ void __launch_bounds__(kRedThreads) elbo_reduce_partial_kernel(const dmvae_elbo_args a, int chunk, int G,
                                                                          float* __restrict__ ws, int do_loss) {
  extern __shared__ float sm[];
  pdl_wait();
  pdl_launch_dependents();
  const int L = a.L, K = a.K, nF = 2 * L + 1;
  const int g = blockIdx.x, set = blockIdx.y;
  float* w_sm = sm;                 // [chunk][K]
  float* f_sm = sm + chunk * K;     // [chunk][nF]
  const int b0 = g * chunk;
  const int nb = min(chunk, a.rows - b0);
  for (int i = threadIdx.x; i < nb * K; i += kRedThreads) {
    int b = i / K, k = i - b * K;
    const int64_t row = b0 + b;
    float w;
    if (a.mode == DMVAE_MODE_DMVAE_SAMPLED) w = a.zeta[row * a.ld_zeta + k];
    else if (set == 1) w = a.w_scratch[row * K + k];
    else w = a.qc[row * K + k];
    w_sm[b * K + k] = w;
  }
  for (int i = threadIdx.x; i < nb * nF; i += kRedThreads) {
    int b = i / nF, f = i - b * nF;
    const int64_t row = b0 + b;
    float v;
    if (f == 2 * L) v = 1.f;
    else if (a.mode == DMVAE_MODE_DMVAE_SAMPLED) v = a.f_scratch[row * 2 * L + f];
    else {
      int l = f < L ? f : f - L;
      float mu = a.mean[row * a.ld_zh + l], lv = a.log_var[row * a.ld_zh + l];
      if (set == 1) {
        float z = mu + expf(0.5f * lv) * a.eps[row * a.ld_eps + l];
        v = f < L ? z : z * z;
      } else {
        v = f < L ? mu : expf(lv) + mu * mu;
      }
    }
    f_sm[b * nF + f] = v;
  }
  __syncthreads();
  float* out = ws + ((size_t)set * G + g) * (size_t)(K * nF);
  for (int o = threadIdx.x; o < K * nF; o += kRedThreads) {
    int k = o / nF, f = o - k * nF;
    float acc = 0.f;
    for (int b = 0; b < nb; ++b) acc += w_sm[b * K + k] * f_sm[b * nF + f];
    out[o] = acc;
  }
  if (do_loss && set == 0 && threadIdx.x < 32) {
    // loss partials: per_sample[b] = (R, C, Zk, total)
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = threadIdx.x; b < nb; b += 32) {
      float4 v = finish_per_sample(a, b0 + b);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
    if (threadIdx.x == 0) {
      const int nsets = (a.mode == DMVAE_MODE_VADE) ? 2 : 1;
      float* lp = ws + (size_t)nsets * G * (size_t)(K * nF) + (size_t)g * 4;
      lp[0] = acc.x; lp[1] = acc.y; lp[2] = acc.z; lp[3] = acc.w;
    }
  }
}

// Stage 2 of the split reduction (fused reconstruction term): complete per_sample from the r_part slots and write the loss
// partials of `chunk` rows per block - no shared memory, so these blocks fit beside resident GEMM CTAs.
__global__ void __launch_bounds__(64) elbo_finish_rows_kernel(const dmvae_elbo_args a, int chunk, int G, float* __restrict__ ws) {
  pdl_wait();
  pdl_launch_dependents();
  const int L = a.L, K = a.K, nF = 2 * L + 1;
  const int g = blockIdx.x, b0 = g * chunk;
  const int nb = min(chunk, a.rows - b0);
  __shared__ float4 wsum[2];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = threadIdx.x; b < nb; b += 64) {
    const float4 v = finish_per_sample(a, b0 + b);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nsets = (a.mode == DMVAE_MODE_VADE) ? 2 : 1;
    float* lp = ws + (size_t)nsets * G * (size_t)(K * nF) + (size_t)g * 4;
    lp[0] = wsum[0].x + wsum[1].x; lp[1] = wsum[0].y + wsum[1].y; lp[2] = wsum[0].z + wsum[1].z; lp[3] = wsum[0].w + wsum[1].w;
  }
}

__global__ void __launch_bounds__(kRedThreads) elbo_reduce_final_kernel(const dmvae_elbo_args a, int G,
                                                                        const float* __restrict__ ws,
                                                                        float* __restrict__ d_means,
                                                                        float* __restrict__ d_log_vars, int accumulate,
                                                                        float* __restrict__ loss_out) {
  pdl_wait();
  pdl_launch_dependents();
  const int L = a.L, K = a.K, nF = 2 * L + 1;
  const int nsets = (a.mode == DMVAE_MODE_VADE) ? 2 : 1;
  const size_t set_stride = (size_t)G * (size_t)(K * nF);
  const float s = a.inv_global_batch, r = a.kl_ratio_dev ? __ldg(a.kl_ratio_dev) : a.kl_ratio;
  // A block owns 32 table elements: warp `sub` sums the chunk partials g = sub, sub + 8, ... for them (coalesced rows of
  // the workspace), the eight warps' sums meet in shared memory and are added in a fixed order (deterministic).
  __shared__ float part[8][6][32];
  const int sub = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 32 + lane;
  const bool active = i < K * L && d_means && d_log_vars;
  {
    const int ii = active ? i : 0;
    const int k = ii / L, l = ii - k * L;
    float U0 = 0.f, U1 = 0.f, Wk = 0.f, V0 = 0.f, V1 = 0.f, Dk = 0.f;
    if (active) {
#pragma unroll 4
      for (int g = sub; g < G; g += 8) {
        const float* p0 = ws + (size_t)g * (K * nF) + (size_t)k * nF;
        U0 += p0[l]; U1 += p0[L + l]; Wk += p0[2 * L];
        if (nsets == 2) {
          const float* p1 = p0 + set_stride;
          V0 += p1[l]; V1 += p1[L + l]; Dk += p1[2 * L];
        }
      }
    }
    part[sub][0][lane] = U0; part[sub][1][lane] = U1; part[sub][2][lane] = Wk;
    part[sub][3][lane] = V0; part[sub][4][lane] = V1; part[sub][5][lane] = Dk;
    __syncthreads();
    if (active && sub == 0) {
      U0 = U1 = Wk = V0 = V1 = Dk = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        U0 += part[w][0][lane]; U1 += part[w][1][lane]; Wk += part[w][2][lane];
        V0 += part[w][3][lane]; V1 += part[w][4][lane]; Dk += part[w][5][lane];
      }
      const float m = a.prior_means[i], plv = a.prior_log_vars[i];
      float dm, dp;
      if (a.mode == DMVAE_MODE_DMVAE_SAMPLED) {
        dm = s * r * U0;
        dp = s * r * U1;
      } else {
        const float iv = expf(-plv);
        dm = -s * r * iv * (U0 - m * Wk);
        dp = 0.5f * s * r * (Wk - iv * (U1 - 2.f * m * U0 + m * m * Wk));
        if (nsets == 2) {
          dm += iv * (V0 - m * Dk);
          dp += 0.5f * iv * (V1 - 2.f * m * V0 + m * m * Dk) - 0.5f * Dk;
        }
      }
      if (accumulate) { d_means[i] += dm; d_log_vars[i] += dp; }
      else { d_means[i] = dm; d_log_vars[i] = dp; }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < 4 && loss_out) {
    const float* lp = ws + (size_t)nsets * set_stride;
    float acc = 0.f;
#pragma unroll 8
    for (int g = 0; g < G; ++g) acc += lp[g * 4 + threadIdx.x];
    loss_out[threadIdx.x] = s * acc;
  }
}

}  // namespace

extern "C" int64_t dmvae_elbo_reduce_workspace(int rows, int L, int K) {
  if (rows <= 0 || L <= 0 || K <= 0) return 0;
  // sized for the smaller of the two chunk sizes in use (scalar kernel: reduce_chunk <= 64; MMA kernel: 64)
  int chunk = reduce_chunk(L, K);
  int64_t G = (rows + chunk - 1) / chunk;
  return 2 * G * (int64_t)K * (2 * L + 1) + 4 * G;
}

extern "C" int dmvae_elbo_reduce(dmvae_ctx* ctx, const dmvae_elbo_args* a, float* d_prior_means, float* d_prior_log_vars,
                                 int accumulate, float* loss_out, float* workspace, void* stream) {
  return dmvae_elbo_reduce_stage(ctx, a, d_prior_means, d_prior_log_vars, accumulate, loss_out, workspace, 0, stream);
}

extern "C" int dmvae_elbo_reduce_stage(dmvae_ctx* ctx, const dmvae_elbo_args* a, float* d_prior_means, float* d_prior_log_vars,
                                       int accumulate, float* loss_out, float* workspace, int stage, void* stream) {
  DMVAE_CHECK_ARG(ctx != nullptr, "elbo_reduce: ctx is NULL");
  DMVAE_CHECK_ARG(stage >= 0 && stage <= 2, "elbo_reduce: stage must be 0 (all), 1 (table partials) or 2 (rows + final sums)");
  int rc = check_elbo_args(a);
  if (rc) return rc;
  DMVAE_CHECK_ARG(workspace != nullptr, "elbo_reduce: workspace is NULL");
  DMVAE_CHECK_ARG((d_prior_means == nullptr) == (d_prior_log_vars == nullptr), "elbo_reduce: pass both prior gradients or neither");
  if (a->rows == 0) return DMVAE_OK;
  const int nsets = (a->mode == DMVAE_MODE_VADE) ? 2 : 1;
  cudaStream_t st = (cudaStream_t)stream;
  const bool mma = elbo_reduce_mma_ok(*a);
  const int chunk = mma ? kRedChunkM : reduce_chunk(a->L, a->K);
  const int G = (a->rows + chunk - 1) / chunk;
  const int do_loss = stage == 0 ? 1 : 0;
  if (stage <= 1) {
    if (mma) {
      rc = launch_elbo_reduce_mma(ctx, *a, G, nsets, workspace, do_loss, st);
      if (rc) return rc;
    } else {
      const size_t smem = sizeof(float) * (size_t)chunk * (size_t)(a->K + 2 * a->L + 1);
      if (smem > 48 * 1024)
        DMVAE_CUDA(cudaFuncSetAttribute(elbo_reduce_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      dmvae_launch(elbo_reduce_partial_kernel, dim3(G, nsets), dim3(kRedThreads), smem, st, true, *a, chunk, G, workspace, do_loss);
      DMVAE_LAUNCH_CHECK(ctx);
    }
    if (stage == 1) return DMVAE_OK;
  } else {
    dmvae_launch(elbo_finish_rows_kernel, dim3(G), dim3(64), 0, st, true, *a, chunk, G, workspace);
    DMVAE_LAUNCH_CHECK(ctx);
  }
  const int fin_blocks = max(1, (a->K * a->L + 31) / 32);
  dmvae_launch(elbo_reduce_final_kernel, dim3(fin_blocks), dim3(kRedThreads), 0, st, true, *a, G, workspace, d_prior_means, d_prior_log_vars,
                                                              accumulate, loss_out);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}
