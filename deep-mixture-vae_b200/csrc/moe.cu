// Mixture-of-experts head: gate-weighted mixture of per-expert softmax classifiers / linear regressors
// and its backward (models.py:76-111, :149-163).  The expert logits pred[b, e*O+o] come from one dense
// GEMM over all experts (the reference tiles the input E times, models.py:76-81).
#include "common.cuh"

namespace {

constexpr int kMaxO = 32;
constexpr float kEps0 = 1e-20f;

template <typename TD>
__global__ void __launch_bounds__(128) moe_kernel(const dmvae_moe_args a) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= a.rows) return;
  const int E = a.E, O = a.O;
  const float s = a.inv_global_batch;
  const float* pred = a.pred + (int64_t)row * a.ld_pred;
  const float* gate = a.gate + (int64_t)row * a.ld_gate;
  const float* Y = a.Y + (int64_t)row * a.ldy;
  TD* dpred = reinterpret_cast<TD*>(a.d_pred) + (int64_t)row * a.ld_dpred;
  float* dgate = a.d_gate + (int64_t)row * a.ld_dgate;
  float u[kMaxO];
#pragma unroll
  for (int o = 0; o < kMaxO; ++o) u[o] = 0.f;

  if (a.classification) {
    // pass 1: unnormalised class probabilities u_o = sum_e gate_e softmax_o(pred_e)   (models.py:84-90)
    for (int e = 0; e < E; ++e) {
      const float* pe = pred + e * O;
      float mx = pe[0];
      for (int o = 1; o < O; ++o) mx = fmaxf(mx, pe[o]);
      float den = 0.f;
      for (int o = 0; o < O; ++o) den += expf(pe[o] - mx);
      const float ge = gate[e] / den;
#pragma unroll
      for (int o = 0; o < kMaxO; ++o)
        if (o < O) u[o] += ge * expf(pe[o] - mx);
    }
    float S = 0.f;
    for (int o = 0; o < O; ++o) S += u[o];
    // loss, prediction, d u
    float loss = 0.f, dot = 0.f, best = -1.f;
    int bi = 0;
    float du[kMaxO];
#pragma unroll
    for (int o = 0; o < kMaxO; ++o) {
      du[o] = 0.f;
      if (o < O) {
        float ys = u[o] / S;                                            // models.py:91-93
        a.y_soft[(int64_t)row * O + o] = ys;
        if (ys > best) { best = ys; bi = o; }
        loss -= 1000.f * Y[o] * logf(ys + kEps0);                       // models.py:153-155
        du[o] = -1000.f * s * Y[o] / (ys + kEps0);                      // d loss / d ys
        dot += du[o] * u[o];
      }
    }
#pragma unroll
    for (int o = 0; o < kMaxO; ++o)
      if (o < O) du[o] = du[o] / S - dot / (S * S);
    a.pred_class[row] = bi;
    // error summand: sum_o |Y - onehot(argmax)| / 2                                      models.py:95-103
    float err = 0.f;
    for (int o = 0; o < O; ++o) err += fabsf(Y[o] - (o == bi ? 1.f : 0.f));
    a.per_sample[2 * (int64_t)row] = loss;
    a.per_sample[2 * (int64_t)row + 1] = 0.5f * err;
    // pass 2: gradients
    for (int e = 0; e < E; ++e) {
      const float* pe = pred + e * O;
      float mx = pe[0];
      for (int o = 1; o < O; ++o) mx = fmaxf(mx, pe[o]);
      float den = 0.f;
      for (int o = 0; o < O; ++o) den += expf(pe[o] - mx);
      const float ge = gate[e];
      float dg = 0.f, pdp = 0.f;
      for (int o = 0; o < O; ++o) {
        float p = expf(pe[o] - mx) / den;
        dg += du[o] * p;
        pdp += du[o] * ge * p;
      }
      dgate[e] = dg;
      for (int o = 0; o < O; ++o) {
        float p = expf(pe[o] - mx) / den;
        dpred[e * O + o] = from_f32<TD>(p * (du[o] * ge - pdp));
      }
    }
  } else {
    // regression: Yhat_o = sum_e gate_e pred_eo                                          models.py:105-111
    float err = 0.f, loss = 0.f;
    float dy[kMaxO];
#pragma unroll
    for (int o = 0; o < kMaxO; ++o) {
      dy[o] = 0.f;
      if (o < O) {
        float yh = 0.f;
        for (int e = 0; e < E; ++e) yh += gate[e] * pred[e * O + o];
        a.y_soft[(int64_t)row * O + o] = yh;
        float df = yh - Y[o];
        err += df * df;
        loss += 0.5f * df * df;                                           // models.py:157-159
        dy[o] = s * df;
      }
    }
    if (a.pred_class) a.pred_class[row] = 0;
    a.per_sample[2 * (int64_t)row] = loss;
    a.per_sample[2 * (int64_t)row + 1] = err;          // error = mean_b(err_b) (= mean over B*O times O)
    for (int e = 0; e < E; ++e) {
      float dg = 0.f;
      for (int o = 0; o < O; ++o) {
        dg += dy[o] * pred[e * O + o];
        dpred[e * O + o] = from_f32<TD>(dy[o] * gate[e]);
      }
      dgate[e] = dg;
    }
  }
  for (int j = E * O; j < a.dpred_cols; ++j) dpred[j] = from_f32<TD>(0.f);
}

// Warp-per-row form (E <= 32): lane e owns expert e - its O logits, softmax and gradients stay in registers, the
// mixtures over the experts are warp reductions.  Rows are read / written as contiguous [E*O] segments, 4096 rows fill
// the chip (the thread-per-row kernel above ran 32 blocks and walked each row serially: 165 us at batch 4096).
template <typename TD, int OM>
__global__ void __launch_bounds__(256) moe_warp_kernel(const dmvae_moe_args a) {
  const int E = a.E, O = a.O;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const float s = a.inv_global_batch;
  const bool act = lane < E;
  for (int row = blockIdx.x * 8 + wib; row < a.rows; row += gridDim.x * 8) {
    const float* pred = a.pred + (int64_t)row * a.ld_pred + lane * O;
    const float* Y = a.Y + (int64_t)row * a.ldy;
    TD* dpred = reinterpret_cast<TD*>(a.d_pred) + (int64_t)row * a.ld_dpred;
    const float ge = act ? __ldg(a.gate + (int64_t)row * a.ld_gate + lane) : 0.f;
    float pe[OM], y[OM];
#pragma unroll
    for (int o = 0; o < OM; ++o) {
      pe[o] = (act && o < O) ? __ldg(pred + o) : 0.f;
      y[o] = o < O ? __ldg(Y + o) : 0.f;
    }
    if (a.classification) {
      float mx = -INFINITY;
#pragma unroll
      for (int o = 0; o < OM; ++o) if (o < O) mx = fmaxf(mx, pe[o]);
      float den = 0.f;
#pragma unroll
      for (int o = 0; o < OM; ++o) {
        pe[o] = o < O ? expf(pe[o] - mx) : 0.f;
        den += pe[o];
      }
      const float inv = act ? 1.f / den : 0.f;
      float u[OM], S = 0.f;
#pragma unroll
      for (int o = 0; o < OM; ++o) {
        pe[o] *= inv;                                                     // p_eo = softmax_o(pred_e)
        u[o] = warp_sum(ge * pe[o]);                                      // models.py:84-90
        S += u[o];
      }
      float loss = 0.f, dot = 0.f, best = -1.f, du[OM];
      int bi = 0;
#pragma unroll
      for (int o = 0; o < OM; ++o) {
        du[o] = 0.f;
        if (o < O) {
          const float ys = u[o] / S;                                      // models.py:91-93
          if (lane == 0) a.y_soft[(int64_t)row * O + o] = ys;
          if (ys > best) { best = ys; bi = o; }
          loss -= 1000.f * y[o] * logf(ys + kEps0);                       // models.py:153-155
          du[o] = -1000.f * s * y[o] / (ys + kEps0);
          dot += du[o] * u[o];
        }
      }
      float err = 0.f, dg = 0.f;
#pragma unroll
      for (int o = 0; o < OM; ++o)
        if (o < O) {
          err += fabsf(y[o] - (o == bi ? 1.f : 0.f));                     // models.py:95-103
          du[o] = du[o] / S - dot / (S * S);
          dg += du[o] * pe[o];
        }
      if (lane == 0) {
        a.pred_class[row] = bi;
        a.per_sample[2 * (int64_t)row] = loss;
        a.per_sample[2 * (int64_t)row + 1] = 0.5f * err;
      }
      if (act) {
        a.d_gate[(int64_t)row * a.ld_dgate + lane] = dg;
        const float pdp = ge * dg;
#pragma unroll
        for (int o = 0; o < OM; ++o)
          if (o < O) dpred[lane * O + o] = from_f32<TD>(pe[o] * (du[o] * ge - pdp));
      }
    } else {
      float err = 0.f, loss = 0.f, dy[OM], dg = 0.f;
#pragma unroll
      for (int o = 0; o < OM; ++o) {
        dy[o] = 0.f;
        if (o < O) {
          const float yh = warp_sum(ge * pe[o]);                          // models.py:105-111
          if (lane == 0) a.y_soft[(int64_t)row * O + o] = yh;
          const float df = yh - y[o];
          err += df * df;
          loss += 0.5f * df * df;                                         // models.py:157-159
          dy[o] = s * df;
          dg += dy[o] * pe[o];
        }
      }
      if (lane == 0) {
        if (a.pred_class) a.pred_class[row] = 0;
        a.per_sample[2 * (int64_t)row] = loss;
        a.per_sample[2 * (int64_t)row + 1] = err;
      }
      if (act) {
        a.d_gate[(int64_t)row * a.ld_dgate + lane] = dg;
#pragma unroll
        for (int o = 0; o < OM; ++o)
          if (o < O) dpred[lane * O + o] = from_f32<TD>(dy[o] * ge);
      }
    }
    for (int j = E * O + lane; j < a.dpred_cols; j += 32) dpred[j] = from_f32<TD>(0.f);
  }
}

template <typename TD>
int launch_moe_warp(const dmvae_moe_args& a, int sm_count, cudaStream_t st) {
  const int blocks = max(1, min(sm_count * 8, (a.rows + 7) / 8));
  if (a.O <= 2) moe_warp_kernel<TD, 2><<<blocks, 256, 0, st>>>(a);
  else if (a.O <= 8) moe_warp_kernel<TD, 8><<<blocks, 256, 0, st>>>(a);
  else if (a.O <= 16) moe_warp_kernel<TD, 16><<<blocks, 256, 0, st>>>(a);
  else moe_warp_kernel<TD, 32><<<blocks, 256, 0, st>>>(a);
  return 0;
}

template <typename TD>
__global__ void softmax_bwd_add_kernel(int rows, int K, const float* __restrict__ q, const float* __restrict__ dgate,
                                       int64_t ld_dgate, TD* __restrict__ dlogits, int64_t ld, int accumulate, int cols) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const float* qr = q + (int64_t)row * K;
  const float* dg = dgate + (int64_t)row * ld_dgate;
  float dot = 0.f;
  for (int k = 0; k < K; ++k) dot += qr[k] * dg[k];
  TD* d = dlogits + (int64_t)row * ld;
  for (int k = 0; k < K; ++k) {
    float v = qr[k] * (dg[k] - dot);
    if (accumulate) v += to_f32<TD>(d[k]);
    d[k] = from_f32<TD>(v);
  }
  if (!accumulate)
    for (int k = K; k < cols; ++k) d[k] = from_f32<TD>(0.f);
}

__global__ void softmax_rows_kernel(const float* __restrict__ sc, int64_t ld, int rows, int K, float* __restrict__ q) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const float* s = sc + (int64_t)row * ld;
  float mx = s[0];
  for (int k = 1; k < K; ++k) mx = fmaxf(mx, s[k]);
  float den = 0.f;
  for (int k = 0; k < K; ++k) den += expf(s[k] - mx);
  for (int k = 0; k < K; ++k) q[(int64_t)row * K + k] = expf(s[k] - mx) / den;
}

// out[c] = scale * sum_r src[r, c]; one block, fixed order (deterministic)
__global__ void __launch_bounds__(256) reduce_columns_kernel(const float* __restrict__ src, int64_t ld, int rows, int cols,
                                                             float scale, float* __restrict__ out) {
  __shared__ float sm[256];
  for (int c = 0; c < cols; ++c) {
    float acc = 0.f;
    for (int r = threadIdx.x; r < rows; r += 256) acc += src[(int64_t)r * ld + c];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = scale * sm[0];
    __syncthreads();
  }
}

template <typename TO>
__global__ void stage_features_kernel(const float* __restrict__ src, int64_t ld, int rows, int n, int relu,
                                      TO* __restrict__ out, int64_t ld_out, int out_cols) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < rows; r += nwarps)
    for (int j = lane; j < out_cols; j += 32) {
      float v = j < n ? src[(int64_t)r * ld + j] : (j == n ? 1.f : 0.f);
      if (relu && j < n) v = fmaxf(v, 0.f);
      out[(int64_t)r * ld_out + j] = from_f32<TO>(v);
    }
}

}  // namespace

extern "C" int dmvae_softmax_rows(dmvae_ctx* ctx, const float* scores, int64_t ld, int rows, int K, float* q, void* stream) {
  DMVAE_CHECK_ARG(ctx && scores && q && rows >= 0 && K > 0 && ld >= K, "softmax_rows: bad arguments");
  if (rows == 0) return DMVAE_OK;
  softmax_rows_kernel<<<(rows + 31) / 32, 32, 0, (cudaStream_t)stream>>>(scores, ld, rows, K, q);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

extern "C" int dmvae_reduce_columns(dmvae_ctx* ctx, const float* src, int64_t ld, int rows, int cols, float scale, float* out,
                                    void* stream) {
  DMVAE_CHECK_ARG(ctx && src && out && rows >= 0 && cols > 0 && ld >= cols, "reduce_columns: bad arguments");
  reduce_columns_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(src, ld, rows, cols, scale, out);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

extern "C" int dmvae_stage_features(dmvae_ctx* ctx, const float* src, int64_t ld, int rows, int n, int relu, void* out,
                                    int out_dtype, int64_t ld_out, int out_cols, void* stream) {
  DMVAE_CHECK_ARG(ctx && src && out && rows >= 0 && n > 0 && ld >= n && out_cols > n && out_cols <= ld_out, "stage_features: bad arguments");
  if (rows == 0) return DMVAE_OK;
  int blocks = min(ctx->sm_count * 8, (rows + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == DMVAE_F32) stage_features_kernel<float><<<blocks, 256, 0, st>>>(src, ld, rows, n, relu, (float*)out, ld_out, out_cols);
  else if (out_dtype == DMVAE_BF16) stage_features_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(src, ld, rows, n, relu, (__nv_bfloat16*)out, ld_out, out_cols);
  else { dmvae_set_error("stage_features: out_dtype %d unsupported", out_dtype); return DMVAE_ERR_INVALID; }
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

extern "C" int dmvae_moe_fwd_bwd(dmvae_ctx* ctx, const dmvae_moe_args* a, void* stream) {
  DMVAE_CHECK_ARG(ctx && a, "moe: NULL argument");
  DMVAE_CHECK_ARG(a->rows >= 0 && a->E > 0 && a->O > 0 && a->O <= kMaxO, "moe: bad sizes rows=%d E=%d O=%d (O <= %d)", a->rows, a->E, a->O, kMaxO);
  DMVAE_CHECK_ARG(a->pred && a->gate && a->Y && a->per_sample && a->y_soft && a->d_pred && a->d_gate, "moe: NULL pointer");
  DMVAE_CHECK_ARG(!a->classification || a->pred_class, "moe: pred_class required for classification");
  DMVAE_CHECK_ARG(a->ld_pred >= a->E * a->O && a->ld_dpred >= a->E * a->O && a->dpred_cols <= a->ld_dpred, "moe: leading dimensions too small");
  if (a->rows == 0) return DMVAE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->dpred_dtype != DMVAE_F32 && a->dpred_dtype != DMVAE_BF16) {
    dmvae_set_error("moe: dpred_dtype %d unsupported", a->dpred_dtype);
    return DMVAE_ERR_INVALID;
  }
  if (a->E <= 32) {
    if (a->dpred_dtype == DMVAE_F32) launch_moe_warp<float>(*a, ctx->sm_count, st);
    else launch_moe_warp<__nv_bfloat16>(*a, ctx->sm_count, st);
  } else {
    int blocks = (a->rows + 31) / 32;
    if (a->dpred_dtype == DMVAE_F32) moe_kernel<float><<<blocks, 32, 0, st>>>(*a);
    else moe_kernel<__nv_bfloat16><<<blocks, 32, 0, st>>>(*a);
  }
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

extern "C" int dmvae_softmax_bwd_add(dmvae_ctx* ctx, int rows, int K, const float* q, const float* d_gate, int64_t ld_dgate,
                                     void* d_logits, int dtype, int64_t ld_dlogits, int accumulate, int cols, void* stream) {
  DMVAE_CHECK_ARG(ctx && q && d_gate && d_logits && rows >= 0 && K > 0 && ld_dlogits >= K && cols <= ld_dlogits, "softmax_bwd_add: bad arguments");
  if (rows == 0) return DMVAE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (rows + 31) / 32;
  if (dtype == DMVAE_F32)
    softmax_bwd_add_kernel<float><<<blocks, 32, 0, st>>>(rows, K, q, d_gate, ld_dgate, (float*)d_logits, ld_dlogits, accumulate, cols);
  else if (dtype == DMVAE_BF16)
    softmax_bwd_add_kernel<__nv_bfloat16><<<blocks, 32, 0, st>>>(rows, K, q, d_gate, ld_dgate, (__nv_bfloat16*)d_logits, ld_dlogits, accumulate, cols);
  else {
    dmvae_set_error("softmax_bwd_add: dtype %d unsupported", dtype);
    return DMVAE_ERR_INVALID;
  }
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}
