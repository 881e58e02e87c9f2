// Micro-benchmark of the speculative softmax half-step (fmha_math.cuh: exp_half64<NP>): cycles per 64-column half-step per warp with
// 1, 2 and 4 warps per SM sub-partition, against the static schedule (sum of the SASS stall fields) and the MUFU floor (64 - 8 NP) * 8.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vist3a_b200/csrc -I include -o tools/ubench/exp_half64 tools/ubench/exp_half64.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "fmha_math.cuh"

using namespace v3a;

template <int NP, int V>
__global__ void k(const float* in, float* out, long long* cyc, int iters, uint32_t zero) {
  uint32_t r[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) r[i] = __float_as_uint(in[(threadIdx.x * 64 + i) & 4095]);
  const float c = 0.1275f;
  const uint64_t cc2 = pack2(c, c);
  float m_run = 4.0f, l = 0.f;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[32];
    uint64_t hs[2] = {0ull, 0ull};
    const float nmc = -m_run * c;
    float mh = 0.f;
    if (V == 1) mh = exp_half64<NP>(r, cc2, pack2(nmc, nmc), hs, pk, zero);
    else exp_half64_v2<NP>(r, cc2, pack2(nmc, nmc), hs, pk, zero);
    float s0, s1, s2, s3;
    unpack2(hs[0], s0, s1);
    unpack2(hs[1], s2, s3);
    l += (s0 + s1) + (s2 + s3);
    if (V == 1) { if ((mh - m_run) * c > 8.0f) m_run = mh; }
    else if (!((s0 + s1) + (s2 + s3) <= 256.0f)) m_run += 1.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= pk[i];
    // perturb the inputs so that nothing is loop invariant (one LOP3 per element pair)
#pragma unroll
    for (int i = 0; i < 64; i += 2) r[i] ^= (acc & zero);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + __uint_as_float(acc & 0x3fffffu);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NP, int V>
void run(const float* in, float* out, long long* cyc) {
  for (int warps_per_smsp : {1, 2, 4}) {
    const int iters = 2000, block = 128 * warps_per_smsp;
    k<NP, V><<<148, block>>>(in, out, cyc, iters, 0u);
    k<NP, V><<<148, block>>>(in, out, cyc, iters, 0u);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < 148; ++i) s += (double)h[i];
    const double per = s / 148 / iters;
    printf("v%d NP=%d  %d warp(s)/SMSP: %.0f cycles per half-step per warp, %.0f per half-step of one warp-equivalent (MUFU floor %d)\n", V, NP, warps_per_smsp, per,
           per / warps_per_smsp, (64 - 8 * NP) * 8);
  }
}

int main() {
  float *in, *out;
  long long* cyc;
  cudaMalloc(&in, 4096 * 4);
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = (float)((i * 2654435761u) >> 8 & 0xffff) / 65536.0f * 20.0f - 16.0f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0, 1>(in, out, cyc);
  run<0, 2>(in, out, cyc);
  run<1, 2>(in, out, cyc);
  run<2, 2>(in, out, cyc);
  run<3, 2>(in, out, cyc);
  run<4, 2>(in, out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
