// fp32 SIMT GEMM (the 1e-4 "exact" tier) and the dtype dispatcher of the dense-layer entry points.
// C[M,N] = epilogue(op(A) . op(B)); 64x64x16 tiles, 256 threads, 4x4 register tile per thread.
// The bf16 tier lives in gemm_tc.cu (tcgen05 + TMA).
#include "common.cuh"
#include "epilogue.cuh"

int dmvae_gemm_bf16_tc(dmvae_ctx* ctx, int trans_a, int trans_b, const void* A, int64_t lda, const void* B, int64_t ldb,
                       void* C, int64_t ldc, int M, int N, int K, const dmvae_gemm_epilogue* epi, cudaStream_t st);

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PADW = 4;

// 4 consecutive elements along the contiguous direction, zero-filled outside [R, C)
__device__ __forceinline__ float4 load4_guarded(const float* __restrict__ base, int64_t ld, int r, int c, int R, int C,
                                                bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r >= R) return v;
  const float* p = base + (int64_t)r * ld + c;
  if (vec && c + 3 < C) return __ldg(reinterpret_cast<const float4*>(p));
  if (c < C) v.x = __ldg(p);
  if (c + 1 < C) v.y = __ldg(p + 1);
  if (c + 2 < C) v.z = __ldg(p + 2);
  if (c + 3 < C) v.w = __ldg(p + 3);
  return v;
}

template <int TA, int TB>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int64_t lda,
                                                        const float* __restrict__ B, int64_t ldb, void* __restrict__ C,
                                                        int64_t ldc, int M, int N, int K, int k_per_split,
                                                        const EpiParams ep, bool vecA, bool vecB) {
  __shared__ __align__(16) float As[2][BK][BM + PADW];
  __shared__ __align__(16) float Bs[2][BK][BN + PADW];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  const int tx = t & 15, ty = t >> 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra, rb;
  auto gload = [&](int k0) {
    if (TA == 0) ra = load4_guarded(A, lda, m0 + (t >> 2), k0 + (t & 3) * 4, M, k_end, vecA);      // A[m][k]
    else ra = load4_guarded(A, lda, k0 + (t >> 4), m0 + (t & 15) * 4, k_end, M, vecA);             // A[k][m]
    if (TB == 0) rb = load4_guarded(B, ldb, k0 + (t >> 4), n0 + (t & 15) * 4, k_end, N, vecB);     // B[k][n]
    else rb = load4_guarded(B, ldb, n0 + (t >> 2), k0 + (t & 3) * 4, N, k_end, vecB);              // B[n][k]
  };
  auto sstore = [&](int buf) {
    if (TA == 0) {
      int m = t >> 2, kq = (t & 3) * 4;
      As[buf][kq][m] = ra.x; As[buf][kq + 1][m] = ra.y; As[buf][kq + 2][m] = ra.z; As[buf][kq + 3][m] = ra.w;
    } else {
      *reinterpret_cast<float4*>(&As[buf][t >> 4][(t & 15) * 4]) = ra;
    }
    if (TB == 0) {
      *reinterpret_cast<float4*>(&Bs[buf][t >> 4][(t & 15) * 4]) = rb;
    } else {
      int n = t >> 2, kq = (t & 3) * 4;
      Bs[buf][kq][n] = rb.x; Bs[buf][kq + 1][n] = rb.y; Bs[buf][kq + 2][n] = rb.z; Bs[buf][kq + 3][n] = rb.w;
    }
  };

  int buf = 0;
  if (k_begin < k_end) {
    gload(k_begin);
    sstore(0);
  }
  __syncthreads();
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool more = k0 + BK < k_end;
    if (more) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) epilogue_store_one(ep, C, ldc, m, n, acc[i][j], blockIdx.z == 0);
    }
  }
}

}  // namespace

static int check_gemm_common(dmvae_ctx* ctx, const void* A, const void* B, const void* C, int M, int N, int K,
                             const dmvae_gemm_epilogue* epi) {
  DMVAE_CHECK_ARG(ctx && A && B && C && epi, "gemm: NULL argument");
  DMVAE_CHECK_ARG(M >= 0 && N > 0 && K >= 0, "gemm: bad sizes M=%d N=%d K=%d", M, N, K);
  DMVAE_CHECK_ARG(epi->split_k >= 1, "gemm: split_k must be >= 1");
  DMVAE_CHECK_ARG(epi->split_k == 1 || (epi->accumulate && epi->out_dtype == DMVAE_F32 && epi->act == DMVAE_ACT_NONE &&
                                        !epi->relu_mask && epi->n_valid >= epi->n_block),
                  "gemm: split_k > 1 needs accumulate=1, fp32 output and a linear epilogue");
  DMVAE_CHECK_ARG(!epi->accumulate || epi->out_dtype == DMVAE_F32, "gemm: accumulate needs fp32 output");
  DMVAE_CHECK_ARG(epi->n_block > 0, "gemm: n_block must be positive");
  return DMVAE_OK;
}

extern "C" int dmvae_gemm(dmvae_ctx* ctx, int dtype, int trans_a, int trans_b, const void* A, int64_t lda, const void* B,
                          int64_t ldb, void* C, int64_t ldc, int M, int N, int K, const dmvae_gemm_epilogue* epi,
                          void* stream) {
  int rc = check_gemm_common(ctx, A, B, C, M, N, K, epi);
  if (rc) return rc;
  if (M == 0) return DMVAE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DMVAE_BF16) return dmvae_gemm_bf16_tc(ctx, trans_a, trans_b, A, lda, B, ldb, C, ldc, M, N, K, epi, st);
  DMVAE_CHECK_ARG(dtype == DMVAE_F32, "gemm: dtype %d unsupported", dtype);
  DMVAE_CHECK_ARG(epi->out_dtype == DMVAE_F32, "gemm(f32): output must be fp32");
  DMVAE_CHECK_ARG(epi->recon == nullptr, "gemm(f32): the fused reconstruction epilogue exists on the bf16 tensor-core path only");
  EpiParams ep = make_epi_params(*epi, DMVAE_F32);
  const bool vecA = (lda % 4 == 0) && (((uintptr_t)A & 15) == 0);
  const bool vecB = (ldb % 4 == 0) && (((uintptr_t)B & 15) == 0);
  int split = epi->split_k;
  int kps = ((K + split - 1) / split + BK - 1) / BK * BK;
  if (kps <= 0) kps = BK;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, split);
  const float* a = (const float*)A;
  const float* b = (const float*)B;
  if (!trans_a && !trans_b) gemm_f32_kernel<0, 0><<<grid, 256, 0, st>>>(a, lda, b, ldb, C, ldc, M, N, K, kps, ep, vecA, vecB);
  else if (!trans_a && trans_b) gemm_f32_kernel<0, 1><<<grid, 256, 0, st>>>(a, lda, b, ldb, C, ldc, M, N, K, kps, ep, vecA, vecB);
  else if (trans_a && !trans_b) gemm_f32_kernel<1, 0><<<grid, 256, 0, st>>>(a, lda, b, ldb, C, ldc, M, N, K, kps, ep, vecA, vecB);
  else gemm_f32_kernel<1, 1><<<grid, 256, 0, st>>>(a, lda, b, ldb, C, ldc, M, N, K, kps, ep, vecA, vecB);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

extern "C" int dmvae_linear_fwd(dmvae_ctx* ctx, int dtype, const void* X, int64_t ldx, const void* W, int64_t ldw, void* Y,
                                int64_t ldy, int out_dtype, int rows, int n_out_pad, int n_in_pad, int act, int n_valid,
                                int n_block, void* stream) {
  dmvae_gemm_epilogue e;
  memset(&e, 0, sizeof(e));
  e.out_dtype = out_dtype;
  e.act = act;
  e.n_valid = n_valid;
  e.n_block = n_block > 0 ? n_block : n_out_pad;
  e.pad_one = 1.f;
  e.split_k = 1;
  return dmvae_gemm(ctx, dtype, 0, 0, X, ldx, W, ldw, Y, ldy, rows, n_out_pad, n_in_pad, &e, stream);
}

extern "C" int dmvae_linear_dgrad(dmvae_ctx* ctx, int dtype, const void* dY, int64_t lddy, const void* W, int64_t ldw,
                                  const void* act_in, int64_t ld_act, void* dX, int64_t lddx, int out_dtype, int rows,
                                  int n_in_pad, int n_out_pad, int n_valid, int n_block, void* stream) {
  dmvae_gemm_epilogue e;
  memset(&e, 0, sizeof(e));
  e.out_dtype = out_dtype;
  e.act = DMVAE_ACT_NONE;
  e.n_valid = n_valid;
  e.n_block = n_block > 0 ? n_block : n_in_pad;
  e.pad_one = 0.f;
  e.relu_mask = act_in;
  e.ld_mask = ld_act;
  e.split_k = 1;
  // dX[rows, n_in_pad] = dY[rows, n_out_pad] . W[n_in_pad, n_out_pad]^T
  return dmvae_gemm(ctx, dtype, 0, 1, dY, lddy, W, ldw, dX, lddx, rows, n_in_pad, n_out_pad, &e, stream);
}

extern "C" int dmvae_linear_wgrad(dmvae_ctx* ctx, int dtype, const void* X, int64_t ldx, const void* dY, int64_t lddy,
                                  float* dW, int64_t lddw, int rows, int n_in_pad, int n_out_pad, int accumulate,
                                  int split_k, void* stream) {
  dmvae_gemm_epilogue e;
  memset(&e, 0, sizeof(e));
  e.out_dtype = DMVAE_F32;
  e.act = DMVAE_ACT_NONE;
  e.n_valid = n_out_pad;
  e.n_block = n_out_pad;
  e.accumulate = accumulate;
  e.split_k = split_k > 0 ? split_k : 1;
  // dW[n_in_pad, n_out_pad] = X[rows, n_in_pad]^T . dY[rows, n_out_pad]
  return dmvae_gemm(ctx, dtype, 1, 0, X, ldx, dY, lddy, dW, lddw, n_in_pad, n_out_pad, rows, &e, stream);
}
