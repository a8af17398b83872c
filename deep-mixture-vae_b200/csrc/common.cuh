// Shared host/device helpers for libdmvae_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <map>
#include <unordered_map>
#include <vector>

#include "../../include/dmvae_b200.h"

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct TmapKey {
  uint64_t ptr, d0, d1, ld;
  uint32_t b0, b1, esz, swz;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && ld == o.ld && b0 == o.b0 && b1 == o.b1 && esz == o.esz &&
           swz == o.swz;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = k.ptr * 0x9E3779B97F4A7C15ull;
    h ^= (k.d0 + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.d1 + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.ld + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (((uint64_t)k.b0 << 32 | k.b1) + (h << 6) + (h >> 2));
    h ^= (((uint64_t)k.esz << 32 | k.swz) + (h << 6) + (h >> 2));
    return (size_t)h;
  }
};

struct dmvae_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  int64_t launches = 0;
  void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled, resolved through the runtime
  std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> tmaps;
  // unit -> CTA-pair schedules of the grouped GEMM launches (gemm_chain.cu), keyed by the launch's tiling signature;
  // device arrays [pairs + 1 offsets][unit ids], built at the first (eager) use of a shape and never freed before the context
  std::map<std::vector<int>, int*> scheds;
  std::mutex mu;
};

void dmvae_set_error(const char* fmt, ...);

#define DMVAE_CHECK_ARG(cond, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      dmvae_set_error(__VA_ARGS__);           \
      return DMVAE_ERR_INVALID;               \
    }                                         \
  } while (0)

#define DMVAE_CUDA(call)                                                                        \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      dmvae_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DMVAE_ERR_CUDA;                                                                    \
    }                                                                                           \
  } while (0)

// after a <<<>>> launch
#define DMVAE_LAUNCH_CHECK(ctx)                                                              \
  do {                                                                                       \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) {                                                                \
      dmvae_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DMVAE_ERR_CUDA;                                                                 \
    }                                                                                        \
    (ctx)->launches++;                                                                       \
  } while (0)

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL)
// ---------------------------------------------------------------------------------------------
// A kernel launched through dmvae_launch(..., pdl=true) may be scheduled while its stream predecessor is still
// draining (its CTAs become resident as the predecessor's exit, their prologue - barrier init, TMEM allocation,
// descriptor prefetch, table staging from kernel parameters - overlaps the predecessor's tail and the launch
// latency disappears).  Such a kernel MUST call pdl_wait() before it reads or writes any global memory the
// predecessor may touch; pdl_wait() returns once the predecessor grid has completed and its writes are visible.
// DMVAE_PDL=0 in the environment turns the attribute off (A/B measurements).
bool dmvae_pdl_enabled();
// First launch of a kernel through dmvae_launch: ask for the MAXIMUM shared-memory carveout.  An SM configured for a
// kernel without shared memory (maximum L1) has to drain before a tcgen05 GEMM CTA (~200 KB of shared memory) can be
// placed on it, and the other way round; the step interleaves small streaming kernels with GEMM launches (and runs some
// beside them on a second stream), so every kernel of the library asks for the same carveout and the SMs never
// reconfigure.  DMVAE_CARVEOUT=0 leaves the driver's default (A/B measurements).
void dmvae_prepare_kernel(const void* kern);

template <typename... KArgs, typename... Args>
static inline cudaError_t dmvae_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                       Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && dmvae_pdl_enabled()) ? 1 : 0;
  dmvae_prepare_kernel(reinterpret_cast<const void*>(kern));
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static inline size_t dmvae_dtype_size(int dt) { return dt == DMVAE_F32 ? 4 : dt == DMVAE_BF16 ? 2 : 1; }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<uint8_t>(uint8_t v) { return (float)v; }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive elements <-> 8 floats.  Pointers must be aligned to 8*sizeof(T).
template <typename T>
struct Vec8;
template <>
struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <>
struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <>
struct Vec8<uint8_t> {
  static __device__ __forceinline__ void load(const uint8_t* p, float (&v)[8]) {
    uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = (float)((r.x >> (8 * i)) & 0xffu);
      v[4 + i] = (float)((r.y >> (8 * i)) & 0xffu);
    }
  }
};
#endif  // __CUDACC__
