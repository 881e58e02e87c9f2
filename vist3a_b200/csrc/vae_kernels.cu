// HBM-bound kernels of the Wan-2.1 VAE decode on the device (the convolutions themselves are implicit GEMMs: gemm_sm100.cu, conv mode with
// temporal taps).  Activations are NDHWC bf16, one clip: [T, H, W, ld] with ld >= C (the 96-channel layers are stored 128 wide so that a
// convolution tap is a whole number of 64-channel k-blocks; the padding channels of every convolution INPUT are kept at zero here).
// Reference layers (utils/wan_utils.py): WanRMS_norm :150-184, SiLU + conv prologue of WanResidualBlock :366-372, WanAttentionBlock :428-475,
// WanResample up-sampling :226-238 / temporal interleave :304-306, AutoencoderKLWan._decode :1078-1117 (clamp to [-1, 1]).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.cuh"
#include "host_util.cuh"

namespace v3a {

namespace {

__device__ __forceinline__ uint32_t pack2bf(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// y[r, c] = silu?( x[r, c] / max(||x[r, :C]||, 1e-12) * sqrt(C) * gamma[c] ) for c < C, 0 for C <= c < ldy.  LPR lanes per row (16 for rows of
// up to 128 channels: two rows per warp, every lane busy on the 96-channel layers that carry most of the bytes; 32 otherwise), each lane
// owns the 16-byte chunks l, l + LPR, ...; two row groups per warp iteration are in flight (loads of the second issued before the
// reduction of the first): one read and one write of every row.
template <bool kSilu, int LPR>
__global__ void __launch_bounds__(256) vae_rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                                          __nv_bfloat16* __restrict__ y, long long ldy, long long rows, int C) {
  constexpr int RPW = 32 / LPR;                        // rows per warp and pass
  constexpr int NCH = LPR == 16 ? 1 : 2;               // chunks per lane (ld <= 128 with 16 lanes; <= 512 with 32)
  const int lane = threadIdx.x & 31, sub = lane % LPR, rsel = lane / LPR;
  const int chunks = C / 8, ychunks = (int)(ldy / 8);
  const float root_c = sqrtf((float)C);
  const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * 8;
  for (long long r0 = warp0 * (2 * RPW); r0 < rows; r0 += nwarps * (2 * RPW)) {
    uint4 v[2][NCH];
    float ss[2] = {0.f, 0.f};
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const long long r = r0 + g * RPW + rsel;
      const uint4* xr = reinterpret_cast<const uint4*>(x + r * ldx);
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int ch = sub + LPR * k;
        v[g][k] = (r < rows && ch < chunks) ? xr[ch] : make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const uint32_t w[4] = {v[g][k].x, v[g][k].y, v[g][k].z, v[g][k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) ss[g] += bf16_lo(w[j]) * bf16_lo(w[j]) + bf16_hi(w[j]) * bf16_hi(w[j]);
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) ss[g] += __shfl_xor_sync(0xffffffffu, ss[g], o);
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const long long r = r0 + g * RPW + rsel;
      if (r >= rows) continue;
      const float scale = root_c / fmaxf(sqrtf(ss[g]), 1e-12f);
      uint4* yr = reinterpret_cast<uint4*>(y + r * ldy);
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int ch = sub + LPR * k;
        if (ch < chunks) {
          const float4 g0 = *reinterpret_cast<const float4*>(gamma + ch * 8), g1 = *reinterpret_cast<const float4*>(gamma + ch * 8 + 4);
          const uint32_t w[4] = {v[g][k].x, v[g][k].y, v[g][k].z, v[g][k].w};
          const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          uint32_t o4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float a = bf16_lo(w[j]) * scale * gm[2 * j], b = bf16_hi(w[j]) * scale * gm[2 * j + 1];
            if (kSilu) {
              a = __fdividef(a, 1.f + __expf(-a));
              b = __fdividef(b, 1.f + __expf(-b));
            }
            o4[j] = pack2bf(a, b);
          }
          yr[ch] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        } else if (ch < ychunks) {
          yr[ch] = make_uint4(0u, 0u, 0u, 0u);   // padding channels of a convolution input stay zero
        }
      }
    }
  }
}

// p[r, :] = softmax(scale * s[r, :]) as bf16; s fp32 [rows, L] (the logits GEMM writes fp32), L % 4 == 0.  One block per row; the row is
// read twice (online max + sum, then the normalised write): 4096 x 4096 fp32 per frame = 64 MB, L2 resident between the passes.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, int L, int valid, long long ldp, float scale) {
  const float* sr = s + (long long)blockIdx.x * L;
  __nv_bfloat16* pr = p + (long long)blockIdx.x * ldp;
  float m = -INFINITY, sum = 0.f;
  // columns [valid, L) are padding (ragged H*W padded to the GEMM's granularity): they count as -inf and come out as P = 0
  auto masked = [&](int i) {
    float4 v = *reinterpret_cast<const float4*>(sr + i);
    if (i + 3 >= valid) {
      if (i >= valid) v.x = -INFINITY;
      if (i + 1 >= valid) v.y = -INFINITY;
      if (i + 2 >= valid) v.z = -INFINITY;
      v.w = -INFINITY;
    }
    return v;
  };
  for (int i = threadIdx.x * 4; i < L; i += 256 * 4) {
    const float4 v = masked(i);
    const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) * scale;
    if (mx == -INFINITY) continue;
    if (mx > m) {
      sum *= __expf(m - mx);   // m = -inf at first: exp(-inf) = 0, sum is 0 anyway
      m = mx;
    }
    sum += __expf(v.x * scale - m) + __expf(v.y * scale - m) + __expf(v.z * scale - m) + __expf(v.w * scale - m);
  }
  __shared__ float sm[8], ssum[8];
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
    const float4 v = masked(i);
    uint2 o;
    o.x = pack2bf(__expf(v.x * scale - M) * inv, __expf(v.y * scale - M) * inv);
    o.y = pack2bf(__expf(v.z * scale - M) * inv, __expf(v.w * scale - M) * inv);
    *reinterpret_cast<uint2*>(pr + i) = o;
  }
}

// temporal up-sampling: y [T, P, 2C] (time_conv output, P = H*W pixels) -> out [2T, P, C]: out[2t + half, p, :] = y[t, p, half*C : half*C + C]
__global__ void __launch_bounds__(256) time_interleave_kernel(const __nv_bfloat16* __restrict__ y, long long ldy, __nv_bfloat16* __restrict__ out,
                                                              long long ldo, long long T, long long P, int C) {
  const int cv = C / 8;   // 16-byte chunks per output row
  const long long total = 2 * T * P * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    const long long row = i / cv, p = row % P, t2 = row / P;
    const long long t = t2 >> 1, half = t2 & 1;
    *reinterpret_cast<uint4*>(out + row * ldo + c * 8) = *reinterpret_cast<const uint4*>(y + (t * P + p) * ldy + half * C + c * 8);
  }
}

// [R, C] (row stride ld_in) -> [C, R] (row stride ld_out; bf16), 32 x 32 tiles through shared memory: V^T for the P V GEMM of the mid-block attention
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out,
                                                             long long ld_out, long long R, int C) {
  __shared__ __nv_bfloat16 tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.x >> 5; j < 32; j += 8) {
    const long long r = r0 + j;
    const int c = c0 + (threadIdx.x & 31);
    tile[j][threadIdx.x & 31] = (r < R && c < C) ? in[r * ld_in + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int j = threadIdx.x >> 5; j < 32; j += 8) {
    const int c = c0 + j;
    const long long r = r0 + (threadIdx.x & 31);
    if (r < R && c < C) out[(long long)c * ld_out + r] = tile[threadIdx.x & 31][j];
  }
}

// depth-to-space (k = 2) after the parity-decomposed up-sampling convolution: in [n*h*w, 4*C] with column (ph*2 + pw)*C + c
// -> out NHWC [n, 2h, 2w, ldo] (channels [0, C); padding channels [C, ldo) are not written: the consumer is an RMS norm over C)
__global__ void __launch_bounds__(256) depth_to_space2_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n,
                                                                   int h, int w, int C, long long ldo) {
  const int cv = C / 8;
  const long long total = n * (2LL * h) * (2LL * w) * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long pix = i / cv;
    const int X = (int)(pix % (2 * w)), Y = (int)((pix / (2 * w)) % (2 * h));
    const long long img = pix / (4LL * w * h);
    const long long src = ((img * h + (Y >> 1)) * w + (X >> 1)) * (4LL * cv) + (long long)(((Y & 1) * 2 + (X & 1)) * cv + c);
    *reinterpret_cast<uint4*>(out + pix * ldo + c * 8) = reinterpret_cast<const uint4*>(in)[src];
  }
}

// latent [C, T, h, w] (fp32 or bf16, one clip) -> NDHWC bf16 [T, h, w, ld] with channels [C, ld) zero
template <bool kF32>
__global__ void __launch_bounds__(256) latent_to_ndhwc_kernel(const void* __restrict__ z, __nv_bfloat16* __restrict__ out, int C, long long THW, int ld) {
  const long long total = THW * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld);
    const long long p = i / ld;
    float v = 0.f;
    if (c < C) v = kF32 ? reinterpret_cast<const float*>(z)[(long long)c * THW + p] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(z)[(long long)c * THW + p]);
    out[i] = __float2bfloat16_rn(v);
  }
}

// conv_out result [T*H*W, ld] fp32 (channels 0..2) -> frames [3, T, H, W] fp32, clamped to [-1, 1] (AutoencoderKLWan._decode :1115)
__global__ void __launch_bounds__(256) frames_out_kernel(const float* __restrict__ y, int ld, float* __restrict__ out, long long THW) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < THW; p += (long long)gridDim.x * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(y + p * ld);
    out[p] = fminf(fmaxf(v.x, -1.f), 1.f);
    out[THW + p] = fminf(fmaxf(v.y, -1.f), 1.f);
    out[2 * THW + p] = fminf(fmaxf(v.z, -1.f), 1.f);
  }
}

// planar bilinear resize with half-pixel centres (F.interpolate(mode="trilinear", align_corners=False) with the frame count kept is a
// per-frame bilinear resize): in [planes, hi, wi] -> out [planes, ho, wo]; source index = scale * (dst + 0.5) - 0.5 clamped at 0
// (ATen area_pixel_compute_source_index), neighbours clamped at the border
__global__ void __launch_bounds__(256) resize_planes_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int hi, int wi, int ho,
                                                            int wo, float sy, float sx) {
  const long long total = planes * ho * wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % wo), y = (int)((i / wo) % ho);
    const long long pl = i / ((long long)wo * ho);
    const float fy = fmaxf(sy * ((float)y + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)x + 0.5f) - 0.5f, 0.f);
    const int y0 = min((int)fy, hi - 1), x0 = min((int)fx, wi - 1);
    const int y1 = min(y0 + 1, hi - 1), x1 = min(x0 + 1, wi - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float* b = in + pl * hi * wi;
    const float v00 = b[(long long)y0 * wi + x0], v01 = b[(long long)y0 * wi + x1], v10 = b[(long long)y1 * wi + x0], v11 = b[(long long)y1 * wi + x1];
    out[i] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}

inline unsigned blocks_for(long long n, int per_block, int cap) {
  const long long b = (n + per_block - 1) / per_block;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

int vae_rmsnorm_entry(const void* x, long long ldx, const float* gamma, void* y, long long ldy, long long rows, long long C, int silu, cudaStream_t st) {
  V3A_REQUIRE(x && gamma && y && rows > 0 && C > 0 && C % 8 == 0 && C <= 512, VIST3A_ERR_INVALID, "vae_rmsnorm: C must be a multiple of 8, <= 512");
  V3A_REQUIRE(ldx >= C && ldy >= C && ldx % 8 == 0 && ldy % 8 == 0 && ldy <= 512, VIST3A_ERR_INVALID, "vae_rmsnorm: row strides must be multiples of 8 elements, >= C");
  V3A_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma) & 15) == 0, VIST3A_ERR_INVALID, "vae_rmsnorm: pointers must be 16-byte aligned");
  const bool narrow = ldy <= 128 && C <= 128;         // 16 lanes per row
  const unsigned grid = blocks_for(rows, narrow ? 32 : 16, num_sms() * 16);
  const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
  __nv_bfloat16* yb = (__nv_bfloat16*)y;
  if (silu) {
    if (narrow) vae_rmsnorm_kernel<true, 16><<<grid, 256, 0, st>>>(xb, ldx, gamma, yb, ldy, rows, (int)C);
    else vae_rmsnorm_kernel<true, 32><<<grid, 256, 0, st>>>(xb, ldx, gamma, yb, ldy, rows, (int)C);
  } else {
    if (narrow) vae_rmsnorm_kernel<false, 16><<<grid, 256, 0, st>>>(xb, ldx, gamma, yb, ldy, rows, (int)C);
    else vae_rmsnorm_kernel<false, 32><<<grid, 256, 0, st>>>(xb, ldx, gamma, yb, ldy, rows, (int)C);
  }
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

int softmax_rows_entry(const float* s, void* p, long long rows, long long L, long long valid, long long ldp, float scale, cudaStream_t st) {
  V3A_REQUIRE(s && p && rows > 0 && rows <= 0x7fffffff && L > 0 && L % 4 == 0 && L <= 0x7fffffff && ldp >= L && ldp % 4 == 0 && valid > 0 && valid <= L,
              VIST3A_ERR_INVALID, "softmax_rows: L and the output row stride must be positive multiples of 4, stride >= L, 0 < valid <= L");
  softmax_rows_kernel<<<(unsigned)rows, 256, 0, st>>>(s, (__nv_bfloat16*)p, (int)L, (int)valid, ldp, scale);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

int time_interleave_entry(const void* y, long long ldy, void* out, long long ldo, long long T, long long P, long long C, cudaStream_t st) {
  V3A_REQUIRE(y && out && T > 0 && P > 0 && C > 0 && C % 8 == 0 && ldy >= 2 * C && ldo >= C && ldy % 8 == 0 && ldo % 8 == 0, VIST3A_ERR_INVALID,
              "time_interleave: C and the row strides must be multiples of 8, ldy >= 2C, ldo >= C");
  time_interleave_kernel<<<blocks_for(2 * T * P * (C / 8), 256, num_sms() * 8), 256, 0, st>>>((const __nv_bfloat16*)y, ldy, (__nv_bfloat16*)out, ldo, T, P,
                                                                                             (int)C);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

int transpose_bf16_entry(const void* in, long long ld_in, void* out, long long ld_out, long long R, long long C, cudaStream_t st) {
  V3A_REQUIRE(in && out && R > 0 && C > 0 && ld_in >= C && ld_out >= R && (R + 31) / 32 <= 0x7fffffff && (C + 31) / 32 <= 65535, VIST3A_ERR_INVALID, "transpose_bf16: bad shape");
  transpose_bf16_kernel<<<dim3((unsigned)((R + 31) / 32), (unsigned)((C + 31) / 32)), 256, 0, st>>>((const __nv_bfloat16*)in, ld_in, (__nv_bfloat16*)out, ld_out, R, (int)C);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

int depth_to_space2_bf16_entry(const void* in, void* out, long long n, long long h, long long w, long long C, long long ldo, cudaStream_t st) {
  V3A_REQUIRE(in && out && n > 0 && h > 0 && w > 0 && C > 0 && C % 8 == 0 && ldo >= C && ldo % 8 == 0, VIST3A_ERR_INVALID, "depth_to_space2_bf16: bad shape");
  depth_to_space2_bf16_kernel<<<blocks_for(n * 4 * h * w * (C / 8), 256, num_sms() * 16), 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, n, (int)h, (int)w,
                                                                                                        (int)C, ldo);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

int latent_to_ndhwc_entry(const void* z, int dtype, void* out, long long C, long long THW, long long ld, cudaStream_t st) {
  V3A_REQUIRE(z && out && C > 0 && THW > 0 && ld >= C, VIST3A_ERR_INVALID, "latent_to_ndhwc: bad shape");
  V3A_REQUIRE(dtype == VIST3A_DTYPE_F32 || dtype == VIST3A_DTYPE_BF16, VIST3A_ERR_INVALID, "latent_to_ndhwc: dtype");
  const unsigned grid = blocks_for(THW * ld, 256, num_sms() * 8);
  if (dtype == VIST3A_DTYPE_F32) latent_to_ndhwc_kernel<true><<<grid, 256, 0, st>>>(z, (__nv_bfloat16*)out, (int)C, THW, (int)ld);
  else latent_to_ndhwc_kernel<false><<<grid, 256, 0, st>>>(z, (__nv_bfloat16*)out, (int)C, THW, (int)ld);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

int resize_planes_entry(const float* in, float* out, long long planes, long long hi, long long wi, long long ho, long long wo, cudaStream_t st) {
  V3A_REQUIRE(in && out && planes > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0, VIST3A_ERR_INVALID, "resize_planes: bad shape");
  resize_planes_kernel<<<blocks_for(planes * ho * wo, 256, num_sms() * 16), 256, 0, st>>>(in, out, planes, (int)hi, (int)wi, (int)ho, (int)wo, (float)hi / (float)ho,
                                                                                           (float)wi / (float)wo);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

int vae_frames_out_entry(const float* y, long long ld, float* out, long long THW, cudaStream_t st) {
  V3A_REQUIRE(y && out && THW > 0 && ld >= 4 && ld % 4 == 0 && ((uintptr_t)y & 15) == 0, VIST3A_ERR_INVALID, "vae_frames_out: bad arguments");
  frames_out_kernel<<<blocks_for(THW, 256, num_sms() * 16), 256, 0, st>>>(y, (int)ld, out, THW);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

}  // namespace v3a
