// Throughput of the legacy warp-level mma.sync variants on sm_100a (what can a bandwidth-bound kernel afford to do with them?)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void k(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 11, b1 = 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 2)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 3)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(b0));
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = fmaf(c[i][j], 1.0001f, 0.5f);     // FFMA reference: 4 per "op"
      }
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char* name, double macs_per_op, int warps) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 20000;
  k<KIND><<<148, warps * 32>>>(out, 100);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<KIND><<<148, warps * 32>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * warps * iters * 8;
  printf("%-28s warps/SM %2d: %8.3f ms, %7.2f Gops/s per SM-chip (%.1f TMAC/s), ~%.1f clk per op per SMSP at 1.9 GHz\n", name, warps, ms,
         ops / ms * 1e-6, ops * macs_per_op / ms * 1e-9, ms * 1e-3 * 1.9e9 / (iters * 8.0 * warps / 4.0));
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("mma.m16n8k8 tf32", 1024, w);
    run<3>("mma.m16n8k4 tf32", 512, w);
    run<1>("mma.m16n8k16 bf16", 2048, w);
    run<2>("mma.m16n8k16 f16", 2048, w);
    run<4>("4 x FFMA (per lane)", 4 * 32, w);
  }
  return 0;
}
