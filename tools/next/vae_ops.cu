// DRAFT (see tools/next/README.md: checked once on a B200, untimed): HBM-bound helper kernels of the planned Wan-VAE device path.
// Per-operation references: oracle/wan_vae_plan.py (rmsnorm_silu, attention's softmax, the time interleave of `upsample`).
// Reference layers: utils/wan_utils.py:150-184 (WanRMS_norm), :428-475 (WanAttentionBlock), :304-306 (temporal interleave).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// y[r, c] = silu?( x[r, c] / max(||x[r, :]||, 1e-12) * sqrt(C) * gamma[c] ), rows = pixels of every frame, C % 8 == 0, C <= 512.
// One warp per row: lane l owns the 16-byte chunks l, l + 32 (C = 96 / 192 / 384 -> 12 / 24 / 48 chunks), one read and one write of the row.
template <bool kSilu>
__global__ void __launch_bounds__(256) rmsnorm_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                           __nv_bfloat16* __restrict__ y, long long rows, int C) {
  const int lane = threadIdx.x & 31;
  const int chunks = C / 8;
  const float root_c = sqrtf((float)C);
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * 8) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + r * C);
    uint4 v[2];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int ch = lane + 32 * k;
      v[k] = ch < chunks ? xr[ch] : make_uint4(0u, 0u, 0u, 0u);
      const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) ss += bf_lo(w[j]) * bf_lo(w[j]) + bf_hi(w[j]) * bf_hi(w[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float scale = root_c / fmaxf(sqrtf(ss), 1e-12f);
    uint4* yr = reinterpret_cast<uint4*>(y + r * C);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int ch = lane + 32 * k;
      if (ch < chunks) {
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + ch * 8), g1 = *reinterpret_cast<const float4*>(gamma + ch * 8 + 4);
        const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        uint32_t o4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = bf_lo(w[j]) * scale * g[2 * j], b = bf_hi(w[j]) * scale * g[2 * j + 1];
          if (kSilu) {
            a = a / (1.f + __expf(-a));
            b = b / (1.f + __expf(-b));
          }
          o4[j] = pack2(a, b);
        }
        yr[ch] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
      }
    }
  }
}

// p[r, :] = softmax(scale * s[r, :]) as bf16; s fp32 [rows, L] (the logits GEMM writes fp32), L % 4 == 0.  One block per row; the row is
// read twice (max + sum in one online pass, then the normalised write): 4096 x 4096 fp32 per frame = 64 MB, L2 resident between passes.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, int L, float scale) {
  const float* sr = s + (long long)blockIdx.x * L;
  __nv_bfloat16* pr = p + (long long)blockIdx.x * L;
  float m = -INFINITY, sum = 0.f;
  for (int i = threadIdx.x * 4; i < L; i += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(sr + i);
    const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) * scale;
    if (mx > m) {
      sum *= __expf(m - mx);   // m = -inf at first: exp(-inf) = 0, sum is 0 anyway
      m = mx;
    }
    sum += __expf(v.x * scale - m) + __expf(v.y * scale - m) + __expf(v.z * scale - m) + __expf(v.w * scale - m);
  }
  __shared__ float sm[8], ssum[8];
  // warp, then block reduction of (m, sum) pairs
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, sum, o);
    const float mn = fmaxf(m, m2);
    sum = (mn == -INFINITY) ? 0.f : sum * __expf(m - mn) + s2 * __expf(m2 - mn);
    m = mn;
  }
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ssum[threadIdx.x >> 5] = sum; }
  __syncthreads();
  float M = -INFINITY;
#pragma unroll
  for (int w = 0; w < 8; ++w) M = fmaxf(M, sm[w]);
  float S = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) S += (sm[w] == -INFINITY) ? 0.f : ssum[w] * __expf(sm[w] - M);
  const float inv = 1.f / S;
  for (int i = threadIdx.x * 4; i < L; i += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(sr + i);
    uint2 o;
    o.x = pack2(__expf(v.x * scale - M) * inv, __expf(v.y * scale - M) * inv);
    o.y = pack2(__expf(v.z * scale - M) * inv, __expf(v.w * scale - M) * inv);
    *reinterpret_cast<uint2*>(pr + i) = o;
  }
}

// temporal up-sampling: y [T, P, 2C] (time_conv output, P = H*W pixels) -> out [2T, P, C]: out[2t + half, p, :] = y[t, p, half*C : half*C + C]
__global__ void __launch_bounds__(256) time_interleave_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ out, long long T,
                                                              long long P, int C) {
  const int cv = C / 8;                                        // 16-byte chunks per output row
  const long long total = 2 * T * P * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    const long long row = i / cv, p = row % P, t2 = row / P;
    const long long t = t2 >> 1, half = t2 & 1;
    reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(y)[((t * P + p) * 2 + half) * cv + c];
  }
}

// [R, C] -> [C, R] (bf16), 32 x 32 tiles through shared memory: V^T for the P V GEMM, and the NCDHW <-> NDHWC ends of the network
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long R, int C) {
  __shared__ __nv_bfloat16 tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.x >> 5; j < 32; j += 8) {
    const long long r = r0 + j;
    const int c = c0 + (threadIdx.x & 31);
    tile[j][threadIdx.x & 31] = (r < R && c < C) ? in[r * C + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int j = threadIdx.x >> 5; j < 32; j += 8) {
    const int c = c0 + j;
    const long long r = r0 + (threadIdx.x & 31);
    if (r < R && c < C) out[(long long)c * R + r] = tile[threadIdx.x & 31][j];
  }
}

}  // namespace

// test entry points of the draft (device pointers, default stream); return cudaGetLastError()
extern "C" int v3a_next_rmsnorm_nhwc(const void* x, const float* gamma, void* y, long long rows, int C, int silu) {
  if (!x || !gamma || !y || rows <= 0 || C <= 0 || C % 8 || C > 512) return -1;
  const unsigned grid = (unsigned)((rows + 7) / 8 < 148 * 16 ? (rows + 7) / 8 : 148 * 16);
  if (silu) rmsnorm_nhwc_kernel<true><<<grid, 256>>>((const __nv_bfloat16*)x, gamma, (__nv_bfloat16*)y, rows, C);
  else rmsnorm_nhwc_kernel<false><<<grid, 256>>>((const __nv_bfloat16*)x, gamma, (__nv_bfloat16*)y, rows, C);
  return (int)cudaGetLastError();
}
extern "C" int v3a_next_softmax_rows(const float* s, void* p, long long rows, int L, float scale) {
  if (!s || !p || rows <= 0 || rows > 0x7fffffff || L <= 0 || L % 4) return -1;
  softmax_rows_kernel<<<(unsigned)rows, 256>>>(s, (__nv_bfloat16*)p, L, scale);
  return (int)cudaGetLastError();
}
extern "C" int v3a_next_time_interleave(const void* y, void* out, long long T, long long P, int C) {
  if (!y || !out || T <= 0 || P <= 0 || C <= 0 || C % 8) return -1;
  time_interleave_kernel<<<148 * 8, 256>>>((const __nv_bfloat16*)y, (__nv_bfloat16*)out, T, P, C);
  return (int)cudaGetLastError();
}
extern "C" int v3a_next_transpose_bf16(const void* in, void* out, long long R, int C) {
  if (!in || !out || R <= 0 || C <= 0 || (R + 31) / 32 > 0x7fffffff || (C + 31) / 32 > 65535) return -1;
  transpose_bf16_kernel<<<dim3((unsigned)((R + 31) / 32), (unsigned)((C + 31) / 32)), 256>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, R, C);
  return (int)cudaGetLastError();
}
