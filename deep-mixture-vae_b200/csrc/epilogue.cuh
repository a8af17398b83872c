// Epilogue shared by the fp32 SIMT GEMM and the tcgen05 GEMM: bias, ReLU, ReLU-mask (dgrad),
// the ones/zero padding columns of the bias-folding layout, fp32 accumulate / split-K reduction.
#pragma once
#include "common.cuh"

struct EpiParams {
  int out_dtype;       // DMVAE_F32 | DMVAE_BF16
  int act;
  int n_valid, n_block;
  float pad_one;
  const void* mask;    // operand dtype
  int mask_dtype;
  int64_t ld_mask;
  const float* bias;
  int accumulate;      // 0: store, 1: +=, 2: atomic += (split-K)
  // fused reconstruction term (dmvae_recon_fuse): rx == nullptr -> off
  const void* rx;
  int rx_dtype, rx_input, rx_D, r_parts;
  int64_t rx_ld;
  float rx_scale, rx_s;
  float* r_part;
};

static inline EpiParams make_epi_params(const dmvae_gemm_epilogue& e, int operand_dtype) {
  EpiParams p;
  p.out_dtype = e.out_dtype;
  p.act = e.act;
  p.n_valid = e.n_valid;
  p.n_block = e.n_block;
  p.pad_one = e.pad_one;
  p.mask = e.relu_mask;
  p.mask_dtype = operand_dtype;
  p.ld_mask = e.ld_mask;
  p.bias = e.bias;
  p.accumulate = e.split_k > 1 ? 2 : (e.accumulate ? 1 : 0);
  p.rx = nullptr;
  p.rx_dtype = p.rx_input = p.rx_D = p.r_parts = 0;
  p.rx_ld = 0;
  p.rx_scale = 1.f;
  p.rx_s = 0.f;
  p.r_part = nullptr;
  if (e.recon) {
    p.rx = e.recon->X;
    p.rx_dtype = e.recon->x_dtype;
    p.rx_ld = e.recon->ldx;
    p.rx_scale = e.recon->x_scale == 0.f ? 1.f : e.recon->x_scale;
    p.rx_input = e.recon->input_type;
    p.rx_s = e.recon->scale;
    p.rx_D = e.recon->D;
    p.r_part = e.recon->r_part;
    p.r_parts = e.recon->r_parts;
  }
  return p;
}

#ifdef __CUDACC__
// value transform for one output element of column n; maskval is the ReLU-mask source (1 when there is none);
// `first` = this CTA holds the first K-split (adds the bias)
__device__ __forceinline__ float epi_apply(const EpiParams& ep, int n, float acc, float maskval, bool first) {
  float v = acc;
  if (ep.bias && first) v += __ldg(ep.bias + n);
  if (ep.act == DMVAE_ACT_RELU) v = fmaxf(v, 0.f);
  else if (ep.act == DMVAE_ACT_SIGMOID) v = 1.f / (1.f + __expf(-v));
  v = maskval > 0.f ? v : 0.f;
  if (ep.n_valid < ep.n_block) {
    int jb = n % ep.n_block;
    if (jb == ep.n_valid) v = ep.pad_one;
    else if (jb > ep.n_valid) v = 0.f;
  }
  return v;
}

__device__ __forceinline__ float epilogue_value(const EpiParams& ep, int m, int n, float acc, bool first) {
  float mk = 1.f;
  if (ep.mask)
    mk = ep.mask_dtype == DMVAE_BF16
             ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(ep.mask)[(int64_t)m * ep.ld_mask + n])
             : reinterpret_cast<const float*>(ep.mask)[(int64_t)m * ep.ld_mask + n];
  return epi_apply(ep, n, acc, mk, first);
}

__device__ __forceinline__ void epilogue_store_one(const EpiParams& ep, void* C, int64_t ldc, int m, int n, float acc,
                                                   bool first) {
  float v = epilogue_value(ep, m, n, acc, first);
  if (ep.out_dtype == DMVAE_BF16) {
    reinterpret_cast<__nv_bfloat16*>(C)[(int64_t)m * ldc + n] = __float2bfloat16_rn(v);
  } else {
    float* c = reinterpret_cast<float*>(C) + (int64_t)m * ldc + n;
    if (ep.accumulate == 2) atomicAdd(c, v);
    else if (ep.accumulate == 1) *c += v;
    else *c = v;
  }
}
#endif
