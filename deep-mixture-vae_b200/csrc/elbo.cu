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

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxK = 128;                 // K values per lane <= 4
constexpr float kEps0 = 1e-20f;

struct ElboParams {
  dmvae_elbo_args a;
  int Ls;          // padded (odd) row stride of the shared prior tables
  int vec_ok;      // 8-wide vector path usable for the streaming part
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
        racc += recon8<INPUT>(x0, d0, g0, s_rec);
        racc += recon8<INPUT>(x1, d1, g1, s_rec);
        Vec8<TD>::store(gr + j, g0);
        Vec8<TD>::store(gr + j + 256, g1);
      }
      if (j < D) {
        float x0[8], d0[8], g0[8];
        Vec8<TX>::load(xr + j, x0);
        Vec8<TD>::load(dr + j, d0);
        racc += recon8<INPUT>(x0, d0, g0, s_rec);
        Vec8<TD>::store(gr + j, g0);
      }
    } else {
      for (int j = lane; j < D; j += 32) {
        float g0;
        racc += recon1<INPUT>(to_f32<TX>(xr[j]), to_f32<TD>(dr[j]), s_rec, g0);
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
  kern<<<blocks, kThreads, smem, st>>>(p);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

template <typename TX, typename TD, int INPUT>
int launch_elbo(dmvae_ctx* ctx, const ElboParams& p, cudaStream_t st) {
  if (p.a.K <= 32) return launch_elbo_k<TX, TD, INPUT, 1>(ctx, p, st);
  if (p.a.K <= 64) return launch_elbo_k<TX, TD, INPUT, 2>(ctx, p, st);
  return launch_elbo_k<TX, TD, INPUT, 4>(ctx, p, st);
}

int check_elbo_args(const dmvae_elbo_args* a) {
  DMVAE_CHECK_ARG(a != nullptr, "elbo: args is NULL");
  DMVAE_CHECK_ARG(a->mode >= 0 && a->mode <= 2, "elbo: unknown mode %d", a->mode);
  DMVAE_CHECK_ARG(a->input_type == DMVAE_INPUT_BINARY || a->input_type == DMVAE_INPUT_REAL,
                  "elbo: input_type %d not implemented (binary | real)", a->input_type);   // base_models.py:84-85
  DMVAE_CHECK_ARG(a->rows >= 0 && a->D > 0 && a->L > 0 && a->K > 0, "elbo: bad sizes rows=%d D=%d L=%d K=%d", a->rows, a->D, a->L, a->K);
  DMVAE_CHECK_ARG(a->K <= kMaxK, "elbo: K=%d exceeds the supported maximum %d", a->K, kMaxK);
  DMVAE_CHECK_ARG(a->X && a->decoded && a->mean && a->log_var && a->prior_means && a->prior_log_vars, "elbo: NULL input");
  DMVAE_CHECK_ARG(a->per_sample && a->qc && a->argmax && a->d_decoded && a->d_mean_kl && a->d_log_var_kl, "elbo: NULL output");
  DMVAE_CHECK_ARG(((uintptr_t)a->per_sample & 15) == 0, "elbo: per_sample must be 16-byte aligned");
  DMVAE_CHECK_ARG(a->ldx >= a->D && a->ld_dec >= a->D && a->ld_ddec >= a->D && a->ddec_cols <= a->ld_ddec, "elbo: leading dimensions too small");
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
  p.Ls = a->L | 1;
  const size_t xs = dmvae_dtype_size(a->x_dtype), ds = dmvae_dtype_size(a->dec_dtype);
  p.vec_ok = (a->D % 8 == 0) && (a->ldx % 8 == 0) && (a->ld_dec % 8 == 0) && (a->ld_ddec % 8 == 0) &&
             (((uintptr_t)a->X) % (8 * xs) == 0) && (((uintptr_t)a->decoded) % (8 * ds) == 0) &&
             (((uintptr_t)a->d_decoded) % (8 * ds) == 0);
  DMVAE_CHECK_ARG(elbo_smem_bytes(a->L, a->K, p.Ls) <= 220 * 1024, "elbo: K*L = %d too large for the shared prior tables", a->K * a->L);
  cudaStream_t st = (cudaStream_t)stream;
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
                                                                          float* __restrict__ ws) {
  extern __shared__ float sm[];
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
  if (set == 0 && threadIdx.x < 32) {
    // loss partials: per_sample[b] = (R, C, Zk, total)
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = threadIdx.x; b < nb; b += 32) {
      float4 v = reinterpret_cast<const float4*>(a.per_sample)[b0 + b];
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

__global__ void __launch_bounds__(kRedThreads) elbo_reduce_final_kernel(const dmvae_elbo_args a, int G,
                                                                        const float* __restrict__ ws,
                                                                        float* __restrict__ d_means,
                                                                        float* __restrict__ d_log_vars, int accumulate,
                                                                        float* __restrict__ loss_out) {
  const int L = a.L, K = a.K, nF = 2 * L + 1;
  const int nsets = (a.mode == DMVAE_MODE_VADE) ? 2 : 1;
  const size_t set_stride = (size_t)G * (size_t)(K * nF);
  const float s = a.inv_global_batch, r = a.kl_ratio_dev ? __ldg(a.kl_ratio_dev) : a.kl_ratio;
  int i = blockIdx.x * kRedThreads + threadIdx.x;
  if (i < K * L && d_means && d_log_vars) {
    int k = i / L, l = i - k * L;
    float U0 = 0.f, U1 = 0.f, Wk = 0.f, V0 = 0.f, V1 = 0.f, Dk = 0.f;
#pragma unroll 8
    for (int g = 0; g < G; ++g) {
      const float* p0 = ws + (size_t)g * (K * nF) + (size_t)k * nF;
      U0 += p0[l]; U1 += p0[L + l]; Wk += p0[2 * L];
      if (nsets == 2) {
        const float* p1 = p0 + set_stride;
        V0 += p1[l]; V1 += p1[L + l]; Dk += p1[2 * L];
      }
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
  int chunk = reduce_chunk(L, K);
  int64_t G = (rows + chunk - 1) / chunk;
  return 2 * G * (int64_t)K * (2 * L + 1) + 4 * G;
}

extern "C" int dmvae_elbo_reduce(dmvae_ctx* ctx, const dmvae_elbo_args* a, float* d_prior_means, float* d_prior_log_vars,
                                 int accumulate, float* loss_out, float* workspace, void* stream) {
  DMVAE_CHECK_ARG(ctx != nullptr, "elbo_reduce: ctx is NULL");
  int rc = check_elbo_args(a);
  if (rc) return rc;
  DMVAE_CHECK_ARG(workspace != nullptr, "elbo_reduce: workspace is NULL");
  DMVAE_CHECK_ARG((d_prior_means == nullptr) == (d_prior_log_vars == nullptr), "elbo_reduce: pass both prior gradients or neither");
  if (a->rows == 0) return DMVAE_OK;
  const int chunk = reduce_chunk(a->L, a->K);
  const int G = (a->rows + chunk - 1) / chunk;
  const int nsets = (a->mode == DMVAE_MODE_VADE) ? 2 : 1;
  const size_t smem = sizeof(float) * (size_t)chunk * (size_t)(a->K + 2 * a->L + 1);
  if (smem > 48 * 1024)
    DMVAE_CUDA(cudaFuncSetAttribute(elbo_reduce_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = (cudaStream_t)stream;
  elbo_reduce_partial_kernel<<<dim3(G, nsets), kRedThreads, smem, st>>>(*a, chunk, G, workspace);
  DMVAE_LAUNCH_CHECK(ctx);
  const int fin_blocks = max(1, (a->K * a->L + kRedThreads - 1) / kRedThreads);
  elbo_reduce_final_kernel<<<fin_blocks, kRedThreads, 0, st>>>(*a, G, workspace, d_prior_means, d_prior_log_vars,
                                                              accumulate, loss_out);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}
