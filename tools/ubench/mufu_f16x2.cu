// Micro-benchmark: MUFU.EX2 throughput for f32 vs packed f16x2 operands (results per clock per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_f16x2 mufu_f16x2.cu && ./mufu_f16x2
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__global__ void k_f32(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
  for (int i = 0; i < iters; ++i) {
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
    a -= 1.0f; b -= 1.0f; c -= 1.0f; d -= 1.0f;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}
__global__ void k_f16x2(float* out, int iters) {
  unsigned a = 0x3c003800u + threadIdx.x, b = a + 1, c = a + 2, d = a + 3;
  const unsigned one = 0x3c003c00u;
  for (int i = 0; i < iters; ++i) {
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a));
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b));
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(c));
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(d));
    asm volatile("sub.f16x2 %0, %0, %1;" : "+r"(a) : "r"(one));
    asm volatile("sub.f16x2 %0, %0, %1;" : "+r"(b) : "r"(one));
    asm volatile("sub.f16x2 %0, %0, %1;" : "+r"(c) : "r"(one));
    asm volatile("sub.f16x2 %0, %0, %1;" : "+r"(d) : "r"(one));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(a ^ b ^ c ^ d);
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, grid = 148 * 2, block = 1024;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int which = 0; which < 2; ++which) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k_f32<<<grid, block>>>(out, iters); else k_f16x2<<<grid, block>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double instr = (double)grid * block * iters * 4;              // MUFU instructions (thread level)
      const double per_clk_sm = instr / (ms * 1e-3) / 148.0 / (clk_khz * 1e3);
      if (rep == 2)
        printf("%s: %.3f ms, %.2f MUFU thread-instr/clk/SM (at the %d MHz attribute clock) -> %.1f exponentials/clk/SM\n", which ? "ex2.f16x2" : "ex2.f32  ", ms,
               per_clk_sm, clk_khz / 1000, per_clk_sm * (which ? 2 : 1));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
