// Context management, error reporting and the small bandwidth-bound kernels of libdmvae_b200:
// input staging, Adam (TF semantics), data-parallel reduce+Adam over peer memory, evaluation helpers.
#include "common.cuh"

#include <stdlib.h>

static thread_local char g_err[1024] = "";

void dmvae_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool dmvae_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DMVAE_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

void dmvae_prepare_kernel(const void* kern) {
  static thread_local const void* last = nullptr;
  if (kern == last) return;
  last = kern;
  static std::mutex mu;
  static std::unordered_map<const void*, int> seen;
  static int enabled = -1;
  std::lock_guard<std::mutex> g(mu);
  if (enabled < 0) {
    const char* e = getenv("DMVAE_CARVEOUT");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled || seen.count(kern)) return;
  seen[kern] = 1;
  (void)cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  (void)cudaGetLastError();
}

extern "C" int dmvae_abi_version(void) { return DMVAE_B200_ABI_VERSION; }
extern "C" const char* dmvae_last_error(void) { return g_err; }

extern "C" int dmvae_ctx_create(int device, dmvae_ctx** out) {
  DMVAE_CHECK_ARG(out != nullptr, "dmvae_ctx_create: out is NULL");
  int n = 0;
  DMVAE_CUDA(cudaGetDeviceCount(&n));
  DMVAE_CHECK_ARG(device >= 0 && device < n, "dmvae_ctx_create: device %d out of range (%d devices)", device, n);
  DMVAE_CUDA(cudaSetDevice(device));
  dmvae_ctx* c = new dmvae_ctx();
  c->device = device;
  cudaDeviceProp p;
  DMVAE_CUDA(cudaGetDeviceProperties(&p, device));
  c->sm_count = p.multiProcessorCount;
  c->cc_major = p.major;
  c->cc_minor = p.minor;
  // cuTensorMapEncodeTiled through the runtime so that libcuda is not a link-time dependency
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    c->encode_tiled = fn;
  (void)cudaGetLastError();
  *out = c;
  return DMVAE_OK;
}

extern "C" int dmvae_ctx_destroy(dmvae_ctx* ctx) {
  if (ctx)
    for (auto& kv : ctx->scheds) cudaFree(kv.second);
  delete ctx;
  return DMVAE_OK;
}

extern "C" int64_t dmvae_ctx_launch_count(const dmvae_ctx* ctx) { return ctx ? ctx->launches : -1; }
extern "C" int dmvae_ctx_has_tcgen05(const dmvae_ctx* ctx) {
  return ctx && ctx->cc_major == 10 && ctx->encode_tiled != nullptr;
}

// ---------------------------------------------------------------------------------------------
// zero / cast
// ---------------------------------------------------------------------------------------------
__global__ void zero_f32_kernel(float4* __restrict__ p4, float* __restrict__ tail, int64_t n4, int64_t ntail) {
  pdl_wait();
  pdl_launch_dependents();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (blockIdx.x == 0 && threadIdx.x < ntail) tail[threadIdx.x] = 0.f;
}

extern "C" int dmvae_zero_f32(dmvae_ctx* ctx, float* p, int64_t n, void* stream) {
  DMVAE_CHECK_ARG(ctx && p && n >= 0, "dmvae_zero_f32: bad arguments");
  DMVAE_CHECK_ARG(((uintptr_t)p & 15) == 0, "dmvae_zero_f32: pointer must be 16-byte aligned");
  if (n == 0) return DMVAE_OK;
  int64_t n4 = n / 4;
  // 4-warp blocks: they fit beside the tcgen05 GEMM CTAs (one free warp slot per scheduler), see adam_bg_kernel
  int blocks = (int)min((int64_t)ctx->sm_count * 16, (n4 + 127) / 128 + 1);
  dmvae_launch(zero_f32_kernel, dim3(blocks), dim3(128), 0, (cudaStream_t)stream, true, (float4*)p, p + n4 * 4, n4, n - n4 * 4);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    float v[8];
    Vec8<float>::load(src + i, v);
    Vec8<__nv_bfloat16>::store(dst + i, v);
  }
  if (i < n && i + 8 > n)
    for (int64_t j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
}

extern "C" int dmvae_cast_bf16(dmvae_ctx* ctx, const float* src, void* dst, int64_t n, void* stream) {
  DMVAE_CHECK_ARG(ctx && src && dst && n >= 0, "dmvae_cast_bf16: bad arguments");
  DMVAE_CHECK_ARG(((uintptr_t)src & 31) == 0 && ((uintptr_t)dst & 15) == 0, "dmvae_cast_bf16: misaligned pointers");
  if (n == 0) return DMVAE_OK;
  int blocks = (int)min((int64_t)ctx->sm_count * 8, (n / 8 + 255) / 256 + 1);
  dmvae_launch(cast_bf16_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, true, src, (__nv_bfloat16*)dst, n);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// split-weight logits layer: W = hi + lo + lo2 in three column groups of the bf16 operand copy, and the fold of the
// three partial products (see include/dmvae_b200.h)
// ---------------------------------------------------------------------------------------------
__global__ void split3_bf16_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ op, int rows, int64_t ld, int K,
                                   int stride) {
  pdl_wait();
  pdl_launch_dependents();
  const int n = rows * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / K, k = i - r * K;
    const float w = W[(int64_t)r * ld + k];
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const float r1 = w - __bfloat162float(hi);                  // exact (Sterbenz-like: hi is w rounded to 8 bits)
    const __nv_bfloat16 lo = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(lo);                 // exact; fits 8 bits
    __nv_bfloat16* o = op + (int64_t)r * ld + k;
    o[0] = hi;
    o[stride] = lo;
    o[2 * stride] = __float2bfloat16_rn(r2);
  }
}

extern "C" int dmvae_split3_bf16(dmvae_ctx* ctx, const float* W, void* W_bf16, int rows, int64_t ld, int K, int stride,
                                 void* stream) {
  DMVAE_CHECK_ARG(ctx && W && W_bf16 && rows >= 0 && K > 0 && stride >= K && ld >= 2 * (int64_t)stride + K,
                  "split3_bf16: need K <= stride and 2 stride + K <= ld");
  if (rows == 0) return DMVAE_OK;
  const int blocks = min(ctx->sm_count * 4, (rows * K + 255) / 256);
  dmvae_launch(split3_bf16_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, true, W, (__nv_bfloat16*)W_bf16, rows, ld, K, stride);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

__global__ void fold3_kernel(float* __restrict__ Y, int64_t ld, int rows, int K, int stride) {
  pdl_wait();
  pdl_launch_dependents();
  const int n = rows * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / K, k = i - r * K;
    float* y = Y + (int64_t)r * ld + k;
    y[0] = (y[2 * stride] + y[stride]) + y[0];                  // smallest terms first
  }
}

extern "C" int dmvae_fold3(dmvae_ctx* ctx, float* Y, int64_t ld, int rows, int K, int stride, void* stream) {
  DMVAE_CHECK_ARG(ctx && Y && rows >= 0 && K > 0 && stride >= K && ld >= 2 * (int64_t)stride + K,
                  "fold3: need K <= stride and 2 stride + K <= ld");
  if (rows == 0) return DMVAE_OK;
  const int blocks = min(ctx->sm_count * 4, (rows * K + 255) / 256);
  dmvae_launch(fold3_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, true, Y, ld, rows, K, stride);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// input staging: X -> operand matrix with the ones column
// ---------------------------------------------------------------------------------------------
template <typename TX, typename TO>
__global__ void stage_input_kernel(const TX* __restrict__ X, int64_t ldx, TO* __restrict__ A, int64_t lda, int rows,
                                   int D, int vec, float xs) {
  pdl_wait();
  pdl_launch_dependents();
  // one warp per row; 8 elements per lane per iteration when the row pitches allow vector access
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < rows; r += nwarps) {
    const TX* x = X + (int64_t)r * ldx;
    TO* a = A + (int64_t)r * lda;
    int j0 = 0;
    if (vec) {
      const int Dv = D & ~7;
      for (int j = lane * 8; j < Dv; j += 256) {
        float v[8];
        Vec8<TX>::load(x + j, v);
        if (sizeof(TX) == 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] *= xs;
        }
        Vec8<TO>::store(a + j, v);
      }
      j0 = Dv;
    }
    for (int j = j0 + lane; j < (int)lda; j += 32) {
      float v = j < D ? to_f32<TX>(x[j]) * (sizeof(TX) == 1 ? xs : 1.f) : (j == D ? 1.f : 0.f);
      a[j] = from_f32<TO>(v);
    }
  }
}

extern "C" int dmvae_stage_input(dmvae_ctx* ctx, const void* X, int x_dtype, int64_t ldx, void* A0, int out_dtype,
                                 int64_t ld_out, int rows, int D, float x_scale, void* stream) {
  const float xs = x_scale == 0.f ? 1.f : x_scale;
  DMVAE_CHECK_ARG(ctx && X && A0, "dmvae_stage_input: NULL pointer");
  DMVAE_CHECK_ARG(rows >= 0 && D > 0 && ldx >= D && ld_out > D, "dmvae_stage_input: need ldx >= D and ld_out > D");
  if (rows == 0) return DMVAE_OK;
  int blocks = min(ctx->sm_count * 8, (rows + 7) / 8);
  cudaStream_t s = (cudaStream_t)stream;
  const int vec = (ldx % 8 == 0) && (ld_out % 8 == 0) && (((uintptr_t)X) % (8 * dmvae_dtype_size(x_dtype)) == 0) &&
                  (((uintptr_t)A0) % (8 * dmvae_dtype_size(out_dtype)) == 0);
#define STAGE(TX, TO) dmvae_launch(stage_input_kernel<TX, TO>, dim3(blocks), dim3(256), 0, s, true, (const TX*)X, ldx, (TO*)A0, ld_out, rows, D, vec, xs)
  if (x_dtype == DMVAE_F32 && out_dtype == DMVAE_F32) STAGE(float, float);
  else if (x_dtype == DMVAE_F32 && out_dtype == DMVAE_BF16) STAGE(float, __nv_bfloat16);
  else if (x_dtype == DMVAE_U8 && out_dtype == DMVAE_F32) STAGE(uint8_t, float);
  else if (x_dtype == DMVAE_U8 && out_dtype == DMVAE_BF16) STAGE(uint8_t, __nv_bfloat16);
  else if (x_dtype == DMVAE_BF16 && out_dtype == DMVAE_BF16) STAGE(__nv_bfloat16, __nv_bfloat16);
  else if (x_dtype == DMVAE_BF16 && out_dtype == DMVAE_F32) STAGE(__nv_bfloat16, float);
  else {
    dmvae_set_error("dmvae_stage_input: unsupported dtype pair (%d -> %d)", x_dtype, out_dtype);
    return DMVAE_ERR_INVALID;
  }
#undef STAGE
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// shuffled minibatch gather (source may be pinned host memory: zero-copy reads)
// ---------------------------------------------------------------------------------------------
// Background launch shape (see adam_bg_kernel): 4-warp blocks of <= 48 registers, one warp per scheduler, so that a block
// fits BESIDE a resident tcgen05 GEMM CTA.  The reads are bus-latency bound when the source is host memory, so a block
// lives for tens of microseconds: a 256-thread block would not fit next to a GEMM CTA and the step's persistent GEMM
// launches would wait for the gather's blocks to drain (measured: e2e 0.43 instead of 0.31 ms per step).
template <typename V>
__global__ void __launch_bounds__(128, 12) gather_rows_kernel(const uint8_t* __restrict__ src, int64_t src_pitch,
                                                              const int32_t* __restrict__ idx, uint8_t* __restrict__ dst,
                                                              int64_t dst_pitch, int rows, int vecs_per_row) {
  // flat over (row, vector): consecutive threads read consecutive 16-byte (4-byte) pieces of a row; four independent
  // loads in flight per thread (<= 40 registers)
  constexpr int U = 4;
  const int64_t total = (int64_t)rows * vecs_per_row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < total; i += U * stride) {
    V v[U];
    int r[U], c[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int64_t e = i + j * stride;
      r[j] = (int)(e / vecs_per_row);
      c[j] = (int)(e - (int64_t)r[j] * vecs_per_row);
      v[j] = *reinterpret_cast<const V*>(src + (int64_t)idx[r[j]] * src_pitch + c[j] * (int64_t)sizeof(V));
    }
#pragma unroll
    for (int j = 0; j < U; ++j) *reinterpret_cast<V*>(dst + r[j] * dst_pitch + c[j] * (int64_t)sizeof(V)) = v[j];
  }
  for (; i < total; i += stride) {
    const int64_t r = i / vecs_per_row, c = i - r * vecs_per_row;
    *reinterpret_cast<V*>(dst + r * dst_pitch + c * (int64_t)sizeof(V)) =
        *reinterpret_cast<const V*>(src + (int64_t)idx[r] * src_pitch + c * (int64_t)sizeof(V));
  }
}

extern "C" int dmvae_gather_rows(dmvae_ctx* ctx, const void* src, int64_t src_pitch_bytes, const int32_t* idx, void* dst,
                                 int64_t dst_pitch_bytes, int rows, int row_bytes, void* stream) {
  DMVAE_CHECK_ARG(ctx && src && idx && dst && rows >= 0 && row_bytes > 0, "gather_rows: bad arguments");
  DMVAE_CHECK_ARG(src_pitch_bytes >= row_bytes && dst_pitch_bytes >= row_bytes, "gather_rows: pitches smaller than a row");
  DMVAE_CHECK_ARG(row_bytes % 4 == 0 && src_pitch_bytes % 4 == 0 && dst_pitch_bytes % 4 == 0 &&
                      ((uintptr_t)src & 3) == 0 && ((uintptr_t)dst & 3) == 0,
                  "gather_rows: rows and pitches must be multiples of 4 bytes");
  if (rows == 0) return DMVAE_OK;
  const bool v16 = row_bytes % 16 == 0 && src_pitch_bytes % 16 == 0 && dst_pitch_bytes % 16 == 0 &&
                   ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0;
  const int vpr = row_bytes / (v16 ? 16 : 4);
  const int64_t total = (int64_t)rows * vpr;
  // 32 blocks of 128 threads, four 16-byte loads in flight each (256 KB outstanding): enough to keep the bus busy (88 us
  // for 3.2 MB of host rows instead of 76 us with a block on every SM).  A deeper queue of host reads stalls the training
  // step running beside it: with a block on every SM the step's first launches start ~60 us late whenever the gather
  // got going first (scripts/epoch_timeline.py: 362 us per end-to-end step with 148 blocks, 345 us with 32, 351 us with 8).
  // DMVAE_GATHER_BLOCKS overrides the cap (measurements).
  int blocks = (int)min((int64_t)ctx->sm_count, (total + 511) / 512);
  {
    static int cap = -1;
    if (cap < 0) {
      const char* e = getenv("DMVAE_GATHER_BLOCKS");
      cap = e ? atoi(e) : 32;
      if (cap <= 0) cap = ctx->sm_count;
    }
    blocks = min(blocks, cap);
  }
  cudaStream_t st = (cudaStream_t)stream;
  // An SM whose shared-memory carveout was configured for this kernel (0 bytes: maximum L1) must drain before a
  // tcgen05 GEMM CTA (~200 KB of shared memory) can be placed on it - the step's GEMM launches then wait for the whole
  // gather (measured: +75 us per step).  Ask for the maximum shared-memory carveout so both kinds of block co-reside.
  static bool carveout_set = false;
  if (!carveout_set) {
    DMVAE_CUDA(cudaFuncSetAttribute(gather_rows_kernel<uint4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    DMVAE_CUDA(cudaFuncSetAttribute(gather_rows_kernel<uint32_t>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    carveout_set = true;
  }
  if (v16)
    gather_rows_kernel<uint4><<<max(1, blocks), 128, 0, st>>>((const uint8_t*)src, src_pitch_bytes, idx, (uint8_t*)dst,
                                                              dst_pitch_bytes, rows, vpr);
  else
    gather_rows_kernel<uint32_t><<<max(1, blocks), 128, 0, st>>>((const uint8_t*)src, src_pitch_bytes, idx, (uint8_t*)dst,
                                                                 dst_pitch_bytes, rows, vpr);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

// The same gather for {0,1}-valued rows stored ONE BIT per element on the host (little-endian bit order inside a byte,
// as numpy.packbits(bitorder="little")): 8x fewer bytes over the bus, expanded to one uint8 per element while writing the
// device batch.  A thread moves 16 packed bytes (128 elements); consecutive threads read consecutive pieces of a row, so
// a 784-element row is ONE 112-byte request on the bus (a word-per-thread version issued 25 small reads per row and
// disturbed the training step beside it as much as the 8x larger byte gather).
__global__ void __launch_bounds__(128, 12) gather_bits_kernel(const uint8_t* __restrict__ src, int64_t src_pitch,
                                                              const int32_t* __restrict__ idx, uint8_t* __restrict__ dst,
                                                              int64_t dst_pitch, int rows, int D, int vecs_per_row) {
  constexpr int U = 2;
  const int64_t total = (int64_t)rows * vecs_per_row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto expand = [&](const uint4& v, int r, int c) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint8_t* d = dst + (int64_t)r * dst_pitch + 128 * c;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {             // output word q holds bits 4q .. 4q+3 of w[k] as bytes
        const uint32_t n = (w[k] >> (4 * q)) & 0xfu;
        o[q] = (n & 1u) | ((n & 2u) << 7) | ((n & 4u) << 14) | ((n & 8u) << 21);
      }
      const int e0 = 128 * c + 32 * k;
      if (e0 + 16 <= D) *reinterpret_cast<uint4*>(d + 32 * k) = make_uint4(o[0], o[1], o[2], o[3]);
      if (e0 + 32 <= D) *reinterpret_cast<uint4*>(d + 32 * k + 16) = make_uint4(o[4], o[5], o[6], o[7]);
    }
  };
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < total; i += U * stride) {
    uint4 v[U];
    int r[U], c[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int64_t e = i + j * stride;
      r[j] = (int)(e / vecs_per_row);
      c[j] = (int)(e - (int64_t)r[j] * vecs_per_row);
      v[j] = *reinterpret_cast<const uint4*>(src + (int64_t)idx[r[j]] * src_pitch + 16 * c[j]);
    }
#pragma unroll
    for (int j = 0; j < U; ++j) expand(v[j], r[j], c[j]);
  }
  for (; i < total; i += stride) {
    const int r = (int)(i / vecs_per_row), c = (int)(i - (int64_t)r * vecs_per_row);
    expand(*reinterpret_cast<const uint4*>(src + (int64_t)idx[r] * src_pitch + 16 * c), r, c);
  }
}

extern "C" int dmvae_gather_rows_bits(dmvae_ctx* ctx, const void* src_bits, int64_t src_pitch_bytes, const int32_t* idx, void* dst,
                                      int64_t dst_pitch_bytes, int rows, int D, void* stream) {
  DMVAE_CHECK_ARG(ctx && src_bits && idx && dst && rows >= 0 && D > 0, "gather_rows_bits: bad arguments");
  const int vpr = (D + 127) / 128;
  DMVAE_CHECK_ARG(D % 16 == 0 && dst_pitch_bytes >= D && dst_pitch_bytes % 16 == 0 && ((uintptr_t)dst & 15) == 0,
                  "gather_rows_bits: D and the destination pitch must be multiples of 16, destination 16-byte aligned");
  DMVAE_CHECK_ARG(src_pitch_bytes >= 16 * (int64_t)vpr && src_pitch_bytes % 16 == 0 && ((uintptr_t)src_bits & 15) == 0,
                  "gather_rows_bits: packed rows must be padded to whole 16-byte pieces (pitch >= %d bytes, multiple of 16)", 16 * vpr);
  if (rows == 0) return DMVAE_OK;
  static bool carveout_set = false;          // same reason as dmvae_gather_rows: co-residence with the step's GEMM CTAs
  if (!carveout_set) {
    DMVAE_CUDA(cudaFuncSetAttribute(gather_bits_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    carveout_set = true;
  }
  const int64_t total = (int64_t)rows * vpr;
  const int blocks = (int)max((int64_t)1, min((int64_t)32, (total + 255) / 256));
  gather_bits_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>((const uint8_t*)src_bits, src_pitch_bytes, idx, (uint8_t*)dst,
                                                                dst_pitch_bytes, rows, D, vpr);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// Adam, TensorFlow semantics: theta -= lr_t * m / (sqrt(v) + eps)
// 16 B read (p,g,m,v) + 12 B written (p,m,v) per parameter (+2 B bf16 copy, +4 B gradient clear).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_update1(float& p, float g, float& m, float& v, float lr_t, float b1, float b2,
                                             float eps, float gs) {
  const float gi = g * gs;
  m = b1 * m + (1.f - b1) * gi;
  v = b2 * v + (1.f - b2) * gi * gi;
  p -= lr_t * m / (sqrtf(v) + eps);
}

__device__ __forceinline__ void adam_update4(float4& p, const float4 g, float4& m, float4& v, float lr_t, float b1,
                                             float b2, float eps, float gs) {
  adam_update1(p.x, g.x, m.x, v.x, lr_t, b1, b2, eps, gs);
  adam_update1(p.y, g.y, m.y, v.y, lr_t, b1, b2, eps, gs);
  adam_update1(p.z, g.z, m.z, v.z, lr_t, b1, b2, eps, gs);
  adam_update1(p.w, g.w, m.w, v.w, lr_t, b1, b2, eps, gs);
}

__device__ __forceinline__ void adam_item(float* __restrict__ params, float* __restrict__ grads, float* __restrict__ m,
                                          float* __restrict__ v, __nv_bfloat16* __restrict__ pbf, int64_t i, float lr_t,
                                          float b1, float b2, float eps, float gs, int zero_grads, float4 p, float4 g,
                                          float4 mm, float4 vv) {
  adam_update4(p, g, mm, vv, lr_t, b1, b2, eps, gs);
  reinterpret_cast<float4*>(params)[i] = p;
  reinterpret_cast<float4*>(m)[i] = mm;
  reinterpret_cast<float4*>(v)[i] = vv;
  if (zero_grads) reinterpret_cast<float4*>(grads)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pbf) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
    uint2 w = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    reinterpret_cast<uint2*>(pbf)[i] = w;
  }
}

__global__ void __launch_bounds__(256, 6) adam_kernel(float* __restrict__ params, float* __restrict__ grads,
                                                       float* __restrict__ m, float* __restrict__ v,
                                                       __nv_bfloat16* __restrict__ pbf, int64_t n4, float lr_t,
                                                       const float* __restrict__ lr_t_dev, float b1, float b2, float eps,
                                                       float gs, int zero_grads) {
  pdl_wait();
  pdl_launch_dependents();
  if (lr_t_dev) lr_t = __ldg(lr_t_dev);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 p = reinterpret_cast<float4*>(params)[i];
    const float4 g = reinterpret_cast<float4*>(grads)[i];
    const float4 mm = reinterpret_cast<float4*>(m)[i];
    const float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_item(params, grads, m, v, pbf, i, lr_t, b1, b2, eps, gs, zero_grads, p, g, mm, vv);
  }
}

// Background shape of the same update (DMVAE_ADAM_BACKGROUND): 4-warp blocks.  The tcgen05 GEMM CTAs (10 warps x 152
// registers: 3 warps on two of the SM's four register files) leave room for exactly one 40-register warp per scheduler,
// so one such block per SM runs BESIDE a GEMM kernel instead of waiting for its CTAs to exit (a 256-thread block needs
// two warps per scheduler and does not fit); when the GEMM CTAs do exit, up to 16 blocks per SM take over.
__global__ void __launch_bounds__(128, 8) adam_bg_kernel(float* __restrict__ params, float* __restrict__ grads,
                                                          float* __restrict__ m, float* __restrict__ v,
                                                          __nv_bfloat16* __restrict__ pbf, int64_t n4, float lr_t,
                                                          const float* __restrict__ lr_t_dev, float b1, float b2, float eps,
                                                          float gs, int zero_grads) {
  pdl_wait();
  pdl_launch_dependents();
  if (lr_t_dev) lr_t = __ldg(lr_t_dev);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 p = reinterpret_cast<float4*>(params)[i];
    const float4 g = reinterpret_cast<float4*>(grads)[i];
    const float4 mm = reinterpret_cast<float4*>(m)[i];
    const float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_item(params, grads, m, v, pbf, i, lr_t, b1, b2, eps, gs, zero_grads, p, g, mm, vv);
  }
}

extern "C" int dmvae_adam(dmvae_ctx* ctx, float* params, float* grads, float* m, float* v, void* params_bf16, int64_t n,
                          float lr_t, const float* lr_t_dev, float beta1, float beta2, float eps, float grad_scale,
                          int flags, void* stream) {
  DMVAE_CHECK_ARG(ctx && params && grads && m && v, "dmvae_adam: NULL pointer");
  DMVAE_CHECK_ARG(n >= 0 && n % 4 == 0, "dmvae_adam: n (%lld) must be a multiple of 4 (flat padded buffer)", (long long)n);
  DMVAE_CHECK_ARG((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)m | (uintptr_t)v) & 15) == 0 &&
                      ((uintptr_t)params_bf16 & 7) == 0,
                  "dmvae_adam: buffers must be 16-byte aligned");
  DMVAE_CHECK_ARG((flags & ~(DMVAE_ADAM_ZERO_GRADS | DMVAE_ADAM_BACKGROUND)) == 0, "dmvae_adam: unknown flags %d", flags);
  if (n == 0) return DMVAE_OK;
  const int zero_grads = flags & DMVAE_ADAM_ZERO_GRADS;
  int64_t n4 = n / 4;
  if (flags & DMVAE_ADAM_BACKGROUND) {
    int blocks = (int)min((int64_t)ctx->sm_count * 16, (n4 + 127) / 128);
    dmvae_launch(adam_bg_kernel, dim3(blocks), dim3(128), 0, (cudaStream_t)stream, true, params, grads, m, v,
                 (__nv_bfloat16*)params_bf16, n4, lr_t, lr_t_dev, beta1, beta2, eps, grad_scale, zero_grads);
  } else {
    int blocks = (int)min((int64_t)ctx->sm_count * 8, (n4 + 255) / 256);
    dmvae_launch(adam_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, true, params, grads, m, v,
                 (__nv_bfloat16*)params_bf16, n4, lr_t, lr_t_dev, beta1, beta2, eps, grad_scale, zero_grads);
  }
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

// per-step device state for graph replay: {uint64 step; uint32 t; float lr_t}
struct StepState {
  unsigned long long step;
  unsigned int t;
  float lr_t;
};
__global__ void step_tick_kernel(StepState* st, float lr, float b1, float b2) {
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->step += 1ull;
    unsigned int t = st->t + 1u;
    st->t = t;
    st->lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
  }
}
extern "C" int dmvae_step_tick(dmvae_ctx* ctx, void* state_dev, float lr, float beta1, float beta2, void* stream) {
  DMVAE_CHECK_ARG(ctx && state_dev, "dmvae_step_tick: NULL argument");
  dmvae_launch(step_tick_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, true, (StepState*)state_dev, lr, beta1, beta2);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

// Per-step scalars into a ring indexed by the step counter: the end-to-end loop reads the epoch's losses once, after its
// last step, instead of queueing a copy between two step graphs.
__global__ void log_append_kernel(const float* __restrict__ src, int n, float* __restrict__ ring, int cap, const StepState* st,
                                  unsigned long long step) {
  pdl_wait();
  pdl_launch_dependents();
  const unsigned long long s = st ? st->step : step;
  if ((int)threadIdx.x < n) ring[(s % (unsigned long long)cap) * n + threadIdx.x] = src[threadIdx.x];
}
extern "C" int dmvae_log_append(dmvae_ctx* ctx, const float* src, int n, float* ring, int cap, const void* state_dev,
                                uint64_t step, void* stream) {
  DMVAE_CHECK_ARG(ctx && src && ring && n > 0 && n <= 32 && cap > 0, "dmvae_log_append: bad arguments");
  dmvae_launch(log_append_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, true, src, n, ring, cap, (const StepState*)state_dev,
               (unsigned long long)step);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// data parallel: reduce the peers' gradient shards over NVLink, Adam on the owned shard, write the
// updated parameters into every replica (reduce-scatter + Adam + all-gather in one kernel).
// The caller brackets the launch with a cross-rank barrier on each side.
// ---------------------------------------------------------------------------------------------
struct DpPeers {
  float* grads[8];
  float* params[8];          // fp32 master of each replica, or NULL: that replica keeps no master for this shard ...
  float* params_rep[8];      // ... except inside the replicated ranges below (prior tables, logits layer)
  __nv_bfloat16* pbf[8];
  int64_t rep_lo4[4], rep_hi4[4];   // float4 index ranges whose fp32 master every replica needs
  int n_rep;
};

// WORLD > 0: compile-time rank count - the peer loads are unrolled and issued in groups of up to GROUP before the first
// add (a run-time loop waits out one NVLink round trip per peer, ~2 us each); WORLD == 0: generic.
// BG: background shape (4-warp blocks, <= 48 registers, groups of 4 loads) that fits beside a resident GEMM CTA.
template <int WORLD, bool BG>
__global__ void __launch_bounds__(BG ? 128 : 256, BG ? 10 : 3)
dp_reduce_adam_kernel(DpPeers peers, int rank, int world_rt, float* __restrict__ m, float* __restrict__ v, int64_t begin4,
                      int64_t end4, float lr_t, const float* __restrict__ lr_t_dev, float b1, float b2, float eps,
                      int clear_grads) {
  const int world = WORLD > 0 ? WORLD : world_rt;
  constexpr int GROUP = WORLD <= 0 ? 1 : (BG && WORLD > 4 ? 4 : WORLD);
  pdl_wait();                                   // launched like every kernel of the step (programmatic edge)
  if (lr_t_dev) lr_t = __ldg(lr_t_dev);
  int64_t i = begin4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < end4; i += stride) {
    // fixed summation order (rank 0 .. world-1) so that every step is reproducible
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (WORLD > 0) {
#pragma unroll
      for (int r0 = 0; r0 < WORLD; r0 += GROUP) {
        float4 t[GROUP];
#pragma unroll
        for (int r = 0; r < GROUP; ++r) t[r] = reinterpret_cast<const float4*>(peers.grads[r0 + r])[i];
#pragma unroll
        for (int r = 0; r < GROUP; ++r) {
          if (r0 + r == 0) g = t[0];
          else { g.x += t[r].x; g.y += t[r].y; g.z += t[r].z; g.w += t[r].w; }
        }
      }
    } else {
      g = reinterpret_cast<const float4*>(peers.grads[0])[i];
      for (int r = 1; r < world; ++r) {
        float4 t = reinterpret_cast<const float4*>(peers.grads[r])[i];
        g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
      }
    }
    // this rank is the only reader of element i of every replica's gradient: clear it for the next step's
    // split-K accumulation (or leave that to a local pass of each rank after the closing barrier: 7/8 of these
    // stores cross NVLink)
    if (clear_grads)
      for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(peers.grads[r])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 p = reinterpret_cast<float4*>(peers.params[rank])[i];
    const int64_t li = i - begin4;
    float4 mm = reinterpret_cast<float4*>(m)[li];
    float4 vv = reinterpret_cast<float4*>(v)[li];
    adam_update4(p, g, mm, vv, lr_t, b1, b2, eps, 1.f);
    reinterpret_cast<float4*>(m)[li] = mm;
    reinterpret_cast<float4*>(v)[li] = vv;
    __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
    uint2 w = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    bool rep = false;
    for (int q = 0; q < peers.n_rep; ++q) rep = rep || (i >= peers.rep_lo4[q] && i < peers.rep_hi4[q]);
    for (int r = 0; r < world; ++r) {
      float* pp = peers.params[r];                                               // NULL: fp32 master kept by its owner only,
      if (!pp && rep) pp = peers.params_rep[r];                                  // except in the replicated ranges
      if (pp) reinterpret_cast<float4*>(pp)[i] = p;
      if (peers.pbf[r]) reinterpret_cast<uint2*>(peers.pbf[r])[i] = w;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// NVSwitch (NVLS) form of the exchange: the gradient sum is done IN THE SWITCH by one multimem.ld_reduce on the
// multicast address of the symmetric buffer (inbound bytes per rank: its own shard once, instead of that shard from
// every peer), and the updated bf16 operand copy / fp32 parameters go out as ONE multimem.st that the switch
// replicates into every replica (outbound: the shard once instead of N times).  The switch's reduction order is fixed
// by the topology, not rank order 0..N-1: results are reproducible run to run but differ from the peer-pointer kernel
// in the last bits.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3)
dp_reduce_adam_mc_kernel(const float* __restrict__ mc_grads, float* __restrict__ mc_params /* or NULL: master stays local */,
                         __nv_bfloat16* __restrict__ mc_pbf /* or NULL */, float* __restrict__ params_local,
                         float* __restrict__ m, float* __restrict__ v, int64_t begin4, int64_t end4, float lr_t,
                         const float* __restrict__ lr_t_dev, float b1, float b2, float eps) {
  pdl_wait();
  if (lr_t_dev) lr_t = __ldg(lr_t_dev);
  int64_t i = begin4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < end4; i += stride) {
    float4 g;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(g.x), "=f"(g.y), "=f"(g.z), "=f"(g.w)
                 : "l"(reinterpret_cast<const float4*>(mc_grads) + i)
                 : "memory");
    float4 p = reinterpret_cast<float4*>(params_local)[i];
    const int64_t li = i - begin4;
    float4 mm = reinterpret_cast<float4*>(m)[li];
    float4 vv = reinterpret_cast<float4*>(v)[li];
    adam_update4(p, g, mm, vv, lr_t, b1, b2, eps, 1.f);
    reinterpret_cast<float4*>(m)[li] = mm;
    reinterpret_cast<float4*>(v)[li] = vv;
    if (mc_params)
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<float4*>(mc_params) + i),
                   "f"(p.x), "f"(p.y), "f"(p.z), "f"(p.w)
                   : "memory");
    else
      reinterpret_cast<float4*>(params_local)[i] = p;
    if (mc_pbf) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
      asm volatile("multimem.st.relaxed.sys.global.v2.bf16x2 [%0], {%1, %2};" ::"l"(reinterpret_cast<uint2*>(mc_pbf) + i),
                   "r"(*reinterpret_cast<uint32_t*>(&lo)), "r"(*reinterpret_cast<uint32_t*>(&hi))
                   : "memory");
    }
  }
}

extern "C" int dmvae_dp_reduce_adam_mc(dmvae_ctx* ctx, const float* mc_grads, float* mc_params, void* mc_params_bf16,
                                       float* params_local, float* m, float* v, int64_t n, int64_t shard_begin,
                                       int64_t shard_end, float lr_t, const float* lr_t_dev, float beta1, float beta2,
                                       float eps, void* stream) {
  DMVAE_CHECK_ARG(ctx && mc_grads && params_local && m && v, "dmvae_dp_reduce_adam_mc: NULL pointer");
  DMVAE_CHECK_ARG(mc_params || mc_params_bf16, "dmvae_dp_reduce_adam_mc: a multicast fp32 or bf16 parameter pointer is required");
  DMVAE_CHECK_ARG(shard_begin % 4 == 0 && shard_end % 4 == 0 && 0 <= shard_begin && shard_begin <= shard_end && shard_end <= n,
                  "dmvae_dp_reduce_adam_mc: shard [%lld,%lld) must be 4-aligned and inside [0,%lld)", (long long)shard_begin,
                  (long long)shard_end, (long long)n);
  DMVAE_CHECK_ARG((((uintptr_t)mc_grads | (uintptr_t)mc_params | (uintptr_t)params_local | (uintptr_t)m | (uintptr_t)v) & 15) == 0 &&
                      ((uintptr_t)mc_params_bf16 & 7) == 0,
                  "dmvae_dp_reduce_adam_mc: buffers must be 16-byte aligned");
  if (shard_end == shard_begin) return DMVAE_OK;
  int64_t n4 = (shard_end - shard_begin) / 4;
  int blocks = (int)min((int64_t)ctx->sm_count * 8, (n4 + 255) / 256);
  dmvae_launch(dp_reduce_adam_mc_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, true, mc_grads, mc_params,
               (__nv_bfloat16*)mc_params_bf16, params_local, m, v, shard_begin / 4, shard_end / 4, lr_t, lr_t_dev, beta1, beta2,
               eps);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// Cross-GPU barrier over peer-mapped flag pads (data parallel, mode "p2p").  Every rank owns a pad of
// [DMVAE_DP_CHANNELS][8] uint32 in symmetric memory; barrier number e of a channel: write e into slot [channel][rank]
// of every peer's pad (after a system-scope fence), then wait until every slot of the own pad's row has reached e.
// Epochs are kept per channel in ordinary device memory and advanced by the kernel, so the call captures into CUDA
// graphs; all ranks issue the same barrier sequence.  A rank that waits longer than ~10 s traps instead of hanging.
// ---------------------------------------------------------------------------------------------
struct DpPads {
  uint32_t* pad[8];
};

__global__ void __launch_bounds__(32) dp_barrier_kernel(DpPads pads, int rank, int world, uint32_t* __restrict__ epochs,
                                                        int channel) {
  pdl_wait();
  const uint32_t e = epochs[channel] + 1u;
  const int p = threadIdx.x;
  if (p < world && p != rank) {
    __threadfence_system();
    volatile uint32_t* out = pads.pad[p] + channel * 8 + rank;
    *out = e;
    volatile uint32_t* in = pads.pad[rank] + channel * 8 + p;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int32_t)(*in - e) < 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 10000000000ull) __trap();
    }
    __threadfence_system();
  }
  __syncwarp();
  if (p == 0) epochs[channel] = e;
}

extern "C" int dmvae_dp_barrier(dmvae_ctx* ctx, int rank, int world, uint32_t* const* pads_host, uint32_t* epochs,
                                int channel, void* stream) {
  DMVAE_CHECK_ARG(ctx && pads_host && epochs, "dmvae_dp_barrier: NULL pointer");
  DMVAE_CHECK_ARG(world >= 1 && world <= 8 && rank >= 0 && rank < world, "dmvae_dp_barrier: world %d rank %d", world, rank);
  DMVAE_CHECK_ARG(channel >= 0 && channel < DMVAE_DP_CHANNELS, "dmvae_dp_barrier: channel %d outside [0,%d)", channel,
                  DMVAE_DP_CHANNELS);
  DpPads pads;
  memset(&pads, 0, sizeof(pads));
  for (int r = 0; r < world; ++r) {
    DMVAE_CHECK_ARG(pads_host[r] != nullptr, "dmvae_dp_barrier: pad of rank %d is NULL", r);
    pads.pad[r] = pads_host[r];
  }
  dmvae_launch(dp_barrier_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, true, pads, rank, world, epochs, channel);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// peer-mapped buffers through CUDA IPC (data parallel without torch: include/dmvae_b200.h)
// ---------------------------------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(dmvae_ipc_handle), "IPC handle does not fit");

extern "C" int dmvae_dp_alloc(dmvae_ctx* ctx, int64_t bytes, void** local_ptr, dmvae_ipc_handle* handle_out) {
  DMVAE_CHECK_ARG(ctx && local_ptr && handle_out && bytes > 0, "dmvae_dp_alloc: bad arguments");
  DMVAE_CUDA(cudaSetDevice(ctx->device));
  void* p = nullptr;
  DMVAE_CUDA(cudaMalloc(&p, (size_t)bytes));
  DMVAE_CUDA(cudaMemset(p, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    dmvae_set_error("dmvae_dp_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return DMVAE_ERR_CUDA;
  }
  memset(handle_out, 0, sizeof(*handle_out));
  memcpy(handle_out->bytes, &h, sizeof(h));
  *local_ptr = p;
  return DMVAE_OK;
}

extern "C" int dmvae_dp_open(dmvae_ctx* ctx, const dmvae_ipc_handle* peer_handle, void** peer_ptr) {
  DMVAE_CHECK_ARG(ctx && peer_handle && peer_ptr, "dmvae_dp_open: NULL argument");
  DMVAE_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, peer_handle->bytes, sizeof(h));
  DMVAE_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return DMVAE_OK;
}

extern "C" int dmvae_dp_close(dmvae_ctx* ctx, void* peer_ptr) {
  DMVAE_CHECK_ARG(ctx && peer_ptr, "dmvae_dp_close: NULL argument");
  DMVAE_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return DMVAE_OK;
}

extern "C" int dmvae_dp_free(dmvae_ctx* ctx, void* local_ptr) {
  DMVAE_CHECK_ARG(ctx && local_ptr, "dmvae_dp_free: NULL argument");
  DMVAE_CUDA(cudaFree(local_ptr));
  return DMVAE_OK;
}

extern "C" int dmvae_dp_reduce_adam(dmvae_ctx* ctx, int rank, int world, float* const* grads_peers_host,
                                    float* const* params_peers_host, void* const* params_bf16_peers_host,
                                    float* const* params_rep_peers_host, const int64_t* rep_ranges, int n_rep, float* m,
                                    float* v, int64_t n, int64_t shard_begin, int64_t shard_end, float lr_t,
                                    const float* lr_t_dev, float beta1, float beta2, float eps, int flags,
                                    void* stream) {
  DMVAE_CHECK_ARG(ctx && grads_peers_host && params_peers_host && m && v, "dmvae_dp_reduce_adam: NULL pointer");
  DMVAE_CHECK_ARG(world >= 1 && world <= 8 && rank >= 0 && rank < world, "dmvae_dp_reduce_adam: world %d rank %d", world, rank);
  DMVAE_CHECK_ARG(shard_begin % 4 == 0 && shard_end % 4 == 0 && 0 <= shard_begin && shard_begin <= shard_end && shard_end <= n,
                  "dmvae_dp_reduce_adam: shard [%lld,%lld) must be 4-aligned and inside [0,%lld)", (long long)shard_begin,
                  (long long)shard_end, (long long)n);
  DpPeers peers;
  memset(&peers, 0, sizeof(peers));
  for (int r = 0; r < world; ++r) {
    peers.grads[r] = grads_peers_host[r];
    peers.params[r] = params_peers_host[r];
    peers.pbf[r] = params_bf16_peers_host ? (__nv_bfloat16*)params_bf16_peers_host[r] : nullptr;
    peers.params_rep[r] = params_rep_peers_host ? params_rep_peers_host[r] : nullptr;
    DMVAE_CHECK_ARG(peers.grads[r] && (peers.params[r] || (r != rank && peers.pbf[r])),
                    "dmvae_dp_reduce_adam: peer %d: gradient pointer, and fp32 or bf16 parameter pointer, required", r);
  }
  DMVAE_CHECK_ARG(n_rep >= 0 && n_rep <= 4 && (n_rep == 0 || (rep_ranges && params_rep_peers_host)),
                  "dmvae_dp_reduce_adam: at most 4 replicated ranges, with their pointers");
  peers.n_rep = n_rep;
  for (int q = 0; q < n_rep; ++q) {
    DMVAE_CHECK_ARG(rep_ranges[2 * q] % 4 == 0 && rep_ranges[2 * q + 1] % 4 == 0, "dmvae_dp_reduce_adam: replicated ranges must be 4-aligned");
    peers.rep_lo4[q] = rep_ranges[2 * q] / 4;
    peers.rep_hi4[q] = rep_ranges[2 * q + 1] / 4;
  }
  if (shard_end == shard_begin) return DMVAE_OK;
  DMVAE_CHECK_ARG((flags & ~(DMVAE_ADAM_ZERO_GRADS | DMVAE_ADAM_BACKGROUND)) == 0, "dmvae_dp_reduce_adam: unknown flags %d", flags);
  int64_t n4 = (shard_end - shard_begin) / 4;
  // DMVAE_ADAM_BACKGROUND: 4-warp blocks that fit beside the GEMM CTAs (see adam_bg_kernel)
  const int threads = (flags & DMVAE_ADAM_BACKGROUND) ? 128 : 256;
  int blocks = (int)min((int64_t)ctx->sm_count * (2048 / threads), (n4 + threads - 1) / threads);
  const bool bg = (flags & DMVAE_ADAM_BACKGROUND) != 0;
  auto kern = bg ? dp_reduce_adam_kernel<0, true> : dp_reduce_adam_kernel<0, false>;
  if (world == 2) kern = bg ? dp_reduce_adam_kernel<2, true> : dp_reduce_adam_kernel<2, false>;
  else if (world == 4) kern = bg ? dp_reduce_adam_kernel<4, true> : dp_reduce_adam_kernel<4, false>;
  else if (world == 8) kern = bg ? dp_reduce_adam_kernel<8, true> : dp_reduce_adam_kernel<8, false>;
  dmvae_launch(kern, dim3(blocks), dim3(threads), 0, (cudaStream_t)stream, true, peers, rank, world, m, v, shard_begin / 4,
               shard_end / 4, lr_t, lr_t_dev, beta1, beta2, eps, flags & DMVAE_ADAM_ZERO_GRADS);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// evaluation: argmax + contingency matrix  (base_models.py:425-432, utils.py:22-30)
// ---------------------------------------------------------------------------------------------
__global__ void argmax_contingency_kernel(const float* __restrict__ scores, int64_t ld, int rows, int K,
                                          const int32_t* __restrict__ classes, int n_labels,
                                          int32_t* __restrict__ argmax_out, int32_t* __restrict__ counts) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* s = scores + (int64_t)r * ld;
  int best = 0;
  float bv = s[0];
  for (int k = 1; k < K; ++k) {   // np.argmax: first maximum wins
    float x = s[k];
    if (x > bv) { bv = x; best = k; }
  }
  if (argmax_out) argmax_out[r] = best;
  if (counts && classes) {
    int c = classes[r];
    if (c >= 0 && c < n_labels) atomicAdd(&counts[best * n_labels + c], 1);
  }
}

extern "C" int dmvae_argmax_contingency(dmvae_ctx* ctx, const float* scores, int64_t ld, int rows, int K,
                                        const int32_t* classes, int n_labels, int32_t* argmax_out, int32_t* counts,
                                        void* stream) {
  DMVAE_CHECK_ARG(ctx && scores && K > 0 && ld >= K && rows >= 0, "dmvae_argmax_contingency: bad arguments");
  if (rows == 0) return DMVAE_OK;
  argmax_contingency_kernel<<<(rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(scores, ld, rows, K, classes, n_labels,
                                                                                   argmax_out, counts);
  DMVAE_LAUNCH_CHECK(ctx);
  return DMVAE_OK;
}
