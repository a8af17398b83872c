// Reparameterised sampling: Z = mu + exp(lv/2) eps (priors.py:86-89), concrete / Gumbel-softmax
// sample zeta = softmax((logits + g)/tau) (priors.py:170-181, utils.py:17-19), and the backward of the
// Gaussian reparameterisation.  Noise is either injected (parity runs) or drawn from Philox4x32-10 with
// counter (global_row, column_block, step, stream) and key = seed - see oracle/philox.py for the
// restatement that pins the integer stream.
#include "common.cuh"
#include "philox.cuh"

namespace {

template <typename TZ>
__global__ void __launch_bounds__(256) reparam_fwd_kernel(const dmvae_reparam_args a) {
  pdl_wait();
  pdl_launch_dependents();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int L = a.L, K = a.K;
  const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  const uint32_t step = (uint32_t)(a.step_dev ? *a.step_dev : a.step);
  for (int row = warp; row < a.rows; row += nwarps) {
    const uint32_t grow = (uint32_t)(a.row_offset + (uint64_t)row);
    const float* mean = a.mean + (int64_t)row * a.ld_zh;
    const float* lv = a.log_var + (int64_t)row * a.ld_zh;
    TZ* z = reinterpret_cast<TZ*>(a.Z_out) + (int64_t)row * a.ld_z;
    if (a.fold) {                              // split-weight logits: hi + lo + lo2 partial products (dmvae_split3_bf16)
      float* y = a.fold + (int64_t)row * a.ld_fold;
      for (int k = lane; k < a.fold_K; k += 32) y[k] = (y[2 * a.fold_stride + k] + y[a.fold_stride + k]) + y[k];
      __syncwarp();
    }
    // ---- Gaussian part: one Philox block covers 4 columns ----
    for (int j = lane; j * 4 < L; j += 32) {
      float e[4];
      if (a.eps_in) {
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = (4 * j + i < L) ? a.eps_in[(int64_t)row * L + 4 * j + i] : 0.f;
      } else {
        uint4 x = philox4x32_10(make_uint4(grow, (uint32_t)j, step, 0u), key);
        box_muller(x.x, x.y, e[0], e[1]);
        box_muller(x.z, x.w, e[2], e[3]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int l = 4 * j + i;
        if (l < L) {
          a.eps_out[(int64_t)row * L + l] = e[i];
          z[l] = from_f32<TZ>(mean[l] + expf(0.5f * lv[l]) * e[i]);
        }
      }
    }
    for (int l = L + lane; l < a.z_cols; l += 32) z[l] = from_f32<TZ>(l == L ? 1.f : 0.f);
    // ---- concrete sample ----
    if (a.zeta_out && a.logits) {
      constexpr int KPL = 4;
      float v[KPL];
      float mx = -INFINITY;
      const float inv_tau = 1.f / a.tau;
#pragma unroll
      for (int jk = 0; jk < KPL; ++jk) {
        int k = lane + 32 * jk;
        v[jk] = -INFINITY;
        if (k < K) {
          float g;
          if (a.gumbel_in) g = a.gumbel_in[(int64_t)row * K + k];
          else {
            uint4 x = philox4x32_10(make_uint4(grow, (uint32_t)(k >> 2), step, 1u), key);
            uint32_t xs = (k & 3) == 0 ? x.x : (k & 3) == 1 ? x.y : (k & 3) == 2 ? x.z : x.w;
            g = -logf(1e-20f - logf(u01(xs) + 1e-20f));                  // utils.py:17-19
          }
          v[jk] = (a.logits[(int64_t)row * a.ld_logits + k] + g) * inv_tau;
          mx = fmaxf(mx, v[jk]);
        }
      }
      mx = warp_max(mx);
      float den = 0.f;
#pragma unroll
      for (int jk = 0; jk < KPL; ++jk) {
        v[jk] = (lane + 32 * jk < K) ? expf(v[jk] - mx) : 0.f;
        den += v[jk];
      }
      den = warp_sum(den);
#pragma unroll
      for (int jk = 0; jk < KPL; ++jk) {
        int k = lane + 32 * jk;
        if (k < K) a.zeta_out[(int64_t)row * K + k] = v[jk] / den;
      }
    }
  }
}

template <typename TO>
__global__ void __launch_bounds__(256) reparam_bwd_kernel(int rows, int L, const float* __restrict__ dmk,
                                                           const float* __restrict__ dlk, int64_t ld_kl,
                                                           const float* __restrict__ dZ, int64_t ld_dz,
                                                           const float* __restrict__ dZe, int64_t ld_dze,
                                                           const float* __restrict__ eps, const float* __restrict__ lv,
                                                           int64_t ld_lv, const float* __restrict__ dme, int64_t ld_dme,
                                                           TO* __restrict__ out, int64_t ld_out, int out_cols) {
  pdl_wait();
  pdl_launch_dependents();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < rows; row += nwarps) {
    TO* o = out + (int64_t)row * ld_out;
    for (int j = lane; j < out_cols; j += 32) {
      float v = 0.f;
      if (j < 2 * L) {
        int l = j < L ? j : j - L;
        float dz = dZ[(int64_t)row * ld_dz + l];
        if (dZe) dz += dZe[(int64_t)row * ld_dze + l];
        if (j < L) v = dmk[(int64_t)row * ld_kl + l] + dz + (dme ? dme[(int64_t)row * ld_dme + l] : 0.f);
        else v = dlk[(int64_t)row * ld_kl + l] + 0.5f * dz * eps[(int64_t)row * L + l] * expf(0.5f * lv[(int64_t)row * ld_lv + l]);
      }
      o[j] = from_f32<TO>(v);
    }
  }
}

}  // namespace

extern "C" int dmvae_reparam_fwd(dmvae_ctx* ctx, const dmvae_reparam_args* a, void* stream) {
  DMVAE_CHECK_ARG(ctx && a, "reparam_fwd: NULL argument");
  DMVAE_CHECK_ARG(a->rows >= 0 && a->L > 0 && a->K >= 0 && a->K <= 128, "reparam_fwd: bad sizes rows=%d L=%d K=%d", a->rows, a->L, a->K);
  DMVAE_CHECK_ARG(a->mean && a->log_var && a->Z_out && a->eps_out, "reparam_fwd: mean, log_var, Z_out, eps_out required");  // priors.py:87
  DMVAE_CHECK_ARG(a->z_cols <= a->ld_z && a->ld_z >= a->L, "reparam_fwd: ld_z too small");
  if (a->zeta_out) DMVAE_CHECK_ARG(a->logits && a->tau > 0.f, "reparam_fwd: the concrete sample needs logits and temperature > 0");  // priors.py:171
  if (a->fold) DMVAE_CHECK_ARG(a->fold_K > 0 && a->fold_stride >= a->fold_K && a->ld_fold >= 2 * (int64_t)a->fold_stride + a->fold_K,
                               "reparam_fwd: fold needs K <= stride and 2 stride + K <= ld");
  if (a->rows == 0) return DMVAE_OK;
  int blocks = min(ctx->sm_count * 8, (a->rows + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->z_dtype == DMVAE_F32) dmvae_launch(reparam_fwd_kernel<float>, dim3(blocks), dim3(256), 0, st, true, *a);
  else if (a->z_dtype == DMVAE_BF16) dmvae_launch(reparam_fwd_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, true, *a);
  else {
    dmvae_set_error("reparam_fwd: z_dtype %d unsupported", a->z_dtype);
    return DMVAE_ERR_INVALID;
  }
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

extern "C" int dmvae_reparam_bwd(dmvae_ctx* ctx, int rows, int L, const float* d_mean_kl, const float* d_log_var_kl,
                                 int64_t ld_kl, const float* dZ, int64_t ld_dz, const float* dZ_extra, int64_t ld_dze,
                                 const float* eps, const float* log_var, int64_t ld_lv, const float* d_mean_extra,
                                 int64_t ld_dme, void* out, int out_dtype, int64_t ld_out, int out_cols, void* stream) {
  DMVAE_CHECK_ARG(ctx && d_mean_kl && d_log_var_kl && dZ && eps && log_var && out, "reparam_bwd: NULL pointer");
  DMVAE_CHECK_ARG(rows >= 0 && L > 0 && out_cols >= 2 * L && out_cols <= ld_out, "reparam_bwd: need 2L <= out_cols <= ld_out");
  if (rows == 0) return DMVAE_OK;
  int blocks = min(ctx->sm_count * 8, (rows + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == DMVAE_F32)
    dmvae_launch(reparam_bwd_kernel<float>, dim3(blocks), dim3(256), 0, st, true, rows, L, d_mean_kl, d_log_var_kl, ld_kl, dZ, ld_dz, dZ_extra, ld_dze,
                                                      eps, log_var, ld_lv, d_mean_extra, ld_dme, (float*)out, ld_out, out_cols);
  else if (out_dtype == DMVAE_BF16)
    dmvae_launch(reparam_bwd_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, true, rows, L, d_mean_kl, d_log_var_kl, ld_kl, dZ, ld_dz, dZ_extra,
                                                              ld_dze, eps, log_var, ld_lv, d_mean_extra, ld_dme, (__nv_bfloat16*)out, ld_out, out_cols);
  else {
    dmvae_set_error("reparam_bwd: out_dtype %d unsupported", out_dtype);
    return DMVAE_ERR_INVALID;
  }
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}
