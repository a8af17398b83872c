// Philox4x32-10 counter-based generator and the transforms the reparameterisation uses (shared by reparam.cu and
// the fused head epilogue of gemm_chain.cu).  oracle/philox.py restates the integer stream.
#pragma once
#include "common.cuh"

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  float rad = sqrtf(-2.f * logf(u01(a)));
  float sn, cs;
  sincospif(2.f * u01(b), &sn, &cs);
  n0 = rad * cs;
  n1 = rad * sn;
}

