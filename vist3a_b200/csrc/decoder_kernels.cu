// HBM-bound kernels of the stitched latent -> 3D-Gaussian decoder: stitching-conv im2col (+ trilinear
// T-upsample), generic NHWC im2col, per-head QK LayerNorm + 2-D RoPE, bilinear resize (+ fused adds),
// depth-to-space, the short-sequence fp32 attention and glue of the camera head, camera construction,
// and the fused per-pixel Gaussian epilogue.  Plain CUDA-core kernels: coalesced / vectorised accesses,
// grids sized from the problem (>= several waves of 148 SMs at the BASELINE shapes).
#include "common.cuh"
#include "host_util.cuh"

namespace v3a {

static inline unsigned grid_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }

// ----------------------------------------------------------------------------------------
// stitching conv im2col
// ----------------------------------------------------------------------------------------
template <bool kF32>
__global__ void __launch_bounds__(256) im2col_stitch_kernel(const void* __restrict__ lat, __nv_bfloat16* __restrict__ A, int B,
                                                            int C, int T, int h, int w) {
  // one thread per (row, c, kt): writes the 9 (ky, kx) taps = 9 consecutive bf16 of A
  const int V = (T - 1) * 4 + 1, oh = h / 2, ow = w / 2;
  const long long total = (long long)B * V * oh * ow * C * 5;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int kt = (int)(i % 5);
  const int c = (int)((i / 5) % C);
  long long row = i / (5LL * C);
  const int ox = (int)(row % ow);
  const int oy = (int)((row / ow) % oh);
  const int v = (int)((row / ((long long)ow * oh)) % V);
  const int b = (int)(row / ((long long)ow * oh * V));
  int f = v + kt - 2;
  f = f < 0 ? 0 : (f > V - 1 ? V - 1 : f);  // replicate padding along T (on the upsampled sequence)
  // align_corners=True linear interpolation T -> V
  float src = (V > 1) ? (float)f * (float)(T - 1) / (float)(V - 1) : 0.f;
  int t0 = (int)floorf(src);
  if (t0 > T - 1) t0 = T - 1;
  const int t1 = t0 + 1 < T ? t0 + 1 : T - 1;
  const float fr = src - (float)t0;
  const long long plane = (long long)h * w;
  const long long base0 = (((long long)b * C + c) * T + t0) * plane;
  const long long base1 = (((long long)b * C + c) * T + t1) * plane;
  __nv_bfloat16* dst = A + row * ((long long)C * 45) + c * 45 + kt * 9;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    int y = 2 * oy + ky - 1;
    y = y < 0 ? 0 : (y > h - 1 ? h - 1 : y);
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      int x = 2 * ox + kx - 1;
      x = x < 0 ? 0 : (x > w - 1 ? w - 1 : x);
      float a, bb;
      if constexpr (kF32) {
        a = reinterpret_cast<const float*>(lat)[base0 + (long long)y * w + x];
        bb = reinterpret_cast<const float*>(lat)[base1 + (long long)y * w + x];
      } else {
        a = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(lat)[base0 + (long long)y * w + x]);
        bb = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(lat)[base1 + (long long)y * w + x]);
      }
      // F.interpolate (linear, align_corners): w0 * a + w1 * b with w0 = 1 - fr
      dst[ky * 3 + kx] = __float2bfloat16_rn((1.0f - fr) * a + fr * bb);
    }
  }
}

int im2col_stitch_entry(const void* lat, int dt, void* A, long long B, long long C, long long T, long long h, long long w,
                        cudaStream_t st) {
  V3A_REQUIRE(lat && A && B > 0 && C > 0 && T > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, VIST3A_ERR_INVALID,
              "im2col_stitch: bad arguments");
  const long long V = (T - 1) * 4 + 1;
  const long long total = B * V * (h / 2) * (w / 2) * C * 5;
  if (dt == VIST3A_DTYPE_F32) im2col_stitch_kernel<true><<<grid_for(total, 256), 256, 0, st>>>(lat, (__nv_bfloat16*)A, (int)B, (int)C, (int)T, (int)h, (int)w);
  else im2col_stitch_kernel<false><<<grid_for(total, 256), 256, 0, st>>>(lat, (__nv_bfloat16*)A, (int)B, (int)C, (int)T, (int)h, (int)w);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// DINOv2 patch embedding as a GEMM: images [n_img, 3, H, W] in [0, 1] -> A [n_img * (H/p) * (W/p), ldA] bf16,
//   A[(n, gy, gx), c*p*p + py*p + px] = (bf16(image[n, c, gy*p + py, gx*p + px]) - mean[c]) / std[c]    (columns >= 3 p^2 zero)
// The stride-p, p x p convolution has no overlap, so its im2col is a permutation; one thread writes two consecutive k.
// ----------------------------------------------------------------------------------------
template <bool kF32>
__global__ void __launch_bounds__(256) patch_embed_im2col_kernel(const void* __restrict__ img, __nv_bfloat16* __restrict__ A, long long ldA, int n_img,
                                                                 int H, int W, int p, float m0, float m1, float m2, float s0, float s1, float s2) {
  const int gh = H / p, gw = W / p;
  const int K = 3 * p * p;
  const long long pairs_per_row = ldA / 2;
  const long long total = (long long)n_img * gh * gw * pairs_per_row;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long row = i / pairs_per_row;
  const int k0 = (int)(i % pairs_per_row) * 2;
  const int gx = (int)(row % gw), gy = (int)((row / gw) % gh);
  const long long n = row / ((long long)gw * gh);
  float v[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int k = k0 + u;
    float x = 0.f;
    if (k < K) {
      const int c = k / (p * p), r = k % (p * p), py = r / p, px = r % p;
      const long long idx = ((n * 3 + c) * H + (gy * p + py)) * (long long)W + (gx * p + px);
      // the reference casts the image to bf16 before normalising (anysplat.py:418: image.to(torch.bfloat16))
      const float raw = kF32 ? __bfloat162float(__float2bfloat16_rn(reinterpret_cast<const float*>(img)[idx]))
                             : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(img)[idx]);
      const float m = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
      x = __fdiv_rn(raw - m, sd);
    }
    v[u] = x;
  }
  *reinterpret_cast<uint32_t*>(A + row * ldA + k0) = pack_bf16(v[0], v[1]);
}

int patch_embed_im2col_entry(const void* img, int dt, void* A, long long ldA, long long n_img, long long H, long long W, int p, const float* mean,
                             const float* std, cudaStream_t st) {
  V3A_REQUIRE(img && A && mean && std && n_img > 0 && H > 0 && W > 0 && p > 0, VIST3A_ERR_INVALID, "patch_embed_im2col: bad arguments");
  V3A_REQUIRE(H % p == 0 && W % p == 0, VIST3A_ERR_INVALID, "patch_embed_im2col: H and W must be multiples of the patch size %d", p);
  V3A_REQUIRE(ldA % 8 == 0 && ldA >= 3ll * p * p, VIST3A_ERR_INVALID, "patch_embed_im2col: ldA must be a multiple of 8 and >= 3 p^2");
  const long long total = n_img * (H / p) * (W / p) * (ldA / 2);
  if (dt == VIST3A_DTYPE_F32)
    patch_embed_im2col_kernel<true><<<grid_for(total, 256), 256, 0, st>>>(img, (__nv_bfloat16*)A, ldA, (int)n_img, (int)H, (int)W, p, mean[0], mean[1], mean[2],
                                                                        std[0], std[1], std[2]);
  else
    patch_embed_im2col_kernel<false><<<grid_for(total, 256), 256, 0, st>>>(img, (__nv_bfloat16*)A, ldA, (int)n_img, (int)H, (int)W, p, mean[0], mean[1], mean[2],
                                                                         std[0], std[1], std[2]);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// RGB views [B,3,V,H,W] in [-1,1] -> zero-padded RGB0 NHWC image [B*V, H, W+8, 4] in [0,1] (3 zero pixels left, 5 right)
// ----------------------------------------------------------------------------------------
template <bool kF32>
__global__ void __launch_bounds__(256) rgb_to_nhwc4pad_kernel(const void* __restrict__ img, float4* __restrict__ out, int B, int V, int H, int W,
                                                              int view_major, float scale, float offset) {
  const int Wp = W + 8;
  const long long total = (long long)B * V * H * Wp;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int xp = (int)(i % Wp);
  const int y = (int)((i / Wp) % H);
  const long long n = i / ((long long)Wp * H);
  const int v = (int)(n % V), b = (int)(n / V);
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  const int x = xp - 3;
  if (x >= 0 && x < W) {
    const long long plane = (long long)H * W;
    // [B,3,V,H,W]: channel stride V * plane;  [B,V,3,H,W] (view_major): channel stride plane
    const long long base = (view_major ? ((long long)b * V + v) * 3 : ((long long)b * 3) * V + v) * plane + (long long)y * W + x;
    const long long cstride = view_major ? plane : (long long)V * plane;
    float c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const long long idx = base + (long long)k * cstride;
      c[k] = kF32 ? reinterpret_cast<const float*>(img)[idx] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(img)[idx]);
    }
    o = make_float4(c[0] * scale + offset, c[1] * scale + offset, c[2] * scale + offset, 0.f);
  }
  out[i] = o;
}

int rgb_to_nhwc4pad_entry(const void* img, int dt, float* out, long long B, long long V, long long H, long long W, int view_major, float scale,
                          float offset, cudaStream_t st) {
  V3A_REQUIRE(img && out && B > 0 && V > 0 && H > 0 && W > 0, VIST3A_ERR_INVALID, "rgb_to_nhwc4pad: bad arguments");
  const long long total = B * V * H * (W + 8);
  if (dt == VIST3A_DTYPE_F32)
    rgb_to_nhwc4pad_kernel<true><<<grid_for(total, 256), 256, 0, st>>>(img, (float4*)out, (int)B, (int)V, (int)H, (int)W, view_major, scale, offset);
  else
    rgb_to_nhwc4pad_kernel<false><<<grid_for(total, 256), 256, 0, st>>>(img, (float4*)out, (int)B, (int)V, (int)H, (int)W, view_major, scale, offset);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// generic NHWC im2col (fp32)
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const float* __restrict__ x, float* __restrict__ A, long long ldA,
                                                          int n_img, int h, int w, int C, int kh, int kw, int stride, int pad,
                                                          int ho, int wo) {
  // one thread per (row, k) with k over the padded row: consecutive threads write consecutive addresses
  const long long total = (long long)n_img * ho * wo * ldA;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % ldA);
  const long long row = i / ldA;
  float v = 0.f;
  if (k < kh * kw * C) {
    const int c = k % C, tap = k / C;
    const int dx = tap % kw, dy = tap / kw;
    const int ox = (int)(row % wo), oy = (int)((row / wo) % ho), n = (int)(row / ((long long)wo * ho));
    const int y = oy * stride + dy - pad, xx = ox * stride + dx - pad;
    if (y >= 0 && y < h && xx >= 0 && xx < w) v = x[(((long long)n * h + y) * w + xx) * C + c];
  }
  A[i] = v;
}

int im2col_nhwc_entry(const float* x, float* A, long long ldA, long long n_img, long long h, long long w, long long C, int kh,
                      int kw, int stride, int pad, cudaStream_t st) {
  V3A_REQUIRE(x && A && n_img > 0 && h > 0 && w > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, VIST3A_ERR_INVALID,
              "im2col_nhwc: bad arguments");
  V3A_REQUIRE(ldA >= (long long)kh * kw * C, VIST3A_ERR_INVALID, "im2col_nhwc: ldA < kh*kw*C");
  const int ho = (int)((h + 2 * pad - kh) / stride + 1), wo = (int)((w + 2 * pad - kw) / stride + 1);
  const long long total = n_img * ho * wo * ldA;
  im2col_nhwc_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, A, ldA, (int)n_img, (int)h, (int)w, (int)C, kh, kw, stride, pad, ho, wo);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// per-head LayerNorm(64) on q, k + 2-D RoPE, in place on a fused bf16 qkv buffer
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qknorm_rope2d_kernel(__nv_bfloat16* __restrict__ qkv, long long ld, long long rows,
                                                            int heads, const float* __restrict__ qw, const float* __restrict__ qb,
                                                            const float* __restrict__ kw, const float* __restrict__ kb, float eps,
                                                            const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                                                            int tokens_per_view, int n_special, int grid_w) {
  // one warp per (row, head); lane l owns elements (2l, 2l+1) of the 64-wide head, for q and for k
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= rows * heads) return;
  const int lane = threadIdx.x & 31;
  const long long row = wid / heads;
  const int head = (int)(wid % heads);
  const int p = (int)(row % tokens_per_view);
  int py = 0, px = 0;
  if (p >= n_special) {
    py = 1 + (p - n_special) / grid_w;
    px = 1 + (p - n_special) % grid_w;
  }
  const int e0 = 2 * lane;                   // element index of .x ; .y = e0 + 1
  const int pos = e0 < 32 ? py : px;         // first half of the head rotates with y, second half with x
  const int j0 = e0 & 15;                    // frequency index of .x (pairs are (j, j+16) inside each 32-wide half)
  const bool upper = (e0 & 16) != 0;         // element is the second member of its pair
  const float c0 = cos_tab[pos * 16 + j0], s0 = sin_tab[pos * 16 + j0];
  const float c1 = cos_tab[pos * 16 + j0 + 1], s1 = sin_tab[pos * 16 + j0 + 1];
  const long long C = (long long)heads * 64;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    __nv_bfloat162* ptr = reinterpret_cast<__nv_bfloat162*>(qkv + row * ld + which * C + head * 64) + lane;
    const float* wgt = which == 0 ? qw : kw;
    const float* bia = which == 0 ? qb : kb;
    const float2 x = __bfloat1622float2(*ptr);
    float s = x.x + x.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / 64.0f);
    const float d0 = x.x - mean, d1 = x.y - mean;
    float ss = d0 * d0 + d1 * d1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * (1.0f / 64.0f) + eps);
    const float n0 = d0 * rstd * wgt[e0] + bia[e0];
    const float n1 = d1 * rstd * wgt[e0 + 1] + bia[e0 + 1];
    // partner element (e +- 16) lives in lane ^ 8
    const float p0 = __shfl_xor_sync(0xffffffffu, n0, 8);
    const float p1 = __shfl_xor_sync(0xffffffffu, n1, 8);
    float o0, o1;
    if (!upper) { o0 = n0 * c0 - p0 * s0; o1 = n1 * c1 - p1 * s1; }   // x*cos + (-x2)*sin
    else        { o0 = n0 * c0 + p0 * s0; o1 = n1 * c1 + p1 * s1; }   // x*cos + ( x1)*sin
    *ptr = __floats2bfloat162_rn(o0, o1);
  }
}

// Same operation with ONE THREAD per (row, q|k, head): the 64 values of a head are 128 contiguous bytes (eight 16-byte loads), the
// LayerNorm statistics and the rotate-half pairs (j, j+16) are thread-local -- no shuffles, 16-byte accesses, and a warp covers the 4 KB of
// one row's q|k heads.  (The warp-per-head kernel above moves 4 bytes per lane and ran at 2.3 TB/s; kept for strides that are not
// multiples of 8 elements.)
__global__ void __launch_bounds__(128) qknorm_rope2d_head_kernel(__nv_bfloat16* __restrict__ qkv, long long ld, long long rows, int heads,
                                                                 const float* __restrict__ qw, const float* __restrict__ qb, const float* __restrict__ kw,
                                                                 const float* __restrict__ kb, float eps, const float* __restrict__ cos_tab,
                                                                 const float* __restrict__ sin_tab, int tokens_per_view, int n_special, int grid_w) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * 2 * heads) return;
  const int head = (int)(t % heads);
  const int which = (int)((t / heads) & 1);
  const long long row = t / (2 * heads);
  const int p = (int)(row % tokens_per_view);
  int py = 0, px = 0;
  if (p >= n_special) {
    py = 1 + (p - n_special) / grid_w;
    px = 1 + (p - n_special) % grid_w;
  }
  uint4* ptr = reinterpret_cast<uint4*>(qkv + row * ld + ((long long)which * heads + head) * 64);
  const float* wgt = which == 0 ? qw : kw;
  const float* bia = which == 0 ? qb : kb;
  // 256-bit accesses: a lane's 128 bytes are 4 whole 32-byte sectors (with 16-byte pieces every warp instruction fetched 32 sectors for
  // half their bytes)
  float x[64];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t u[8];
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                 : "l"(ptr + 2 * k));
#pragma unroll
    for (int e = 0; e < 8; ++e) { x[16 * k + 2 * e] = bf16_lo(u[e]); x[16 * k + 2 * e + 1] = bf16_hi(u[e]); }
  }
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 64; ++e) s += x[e];
  const float mean = s * (1.0f / 64.0f);
  float ss = 0.f;
#pragma unroll
  for (int e = 0; e < 64; ++e) {
    x[e] -= mean;
    ss += x[e] * x[e];
  }
  const float rstd = rsqrtf(ss * (1.0f / 64.0f) + eps);
#pragma unroll
  for (int e = 0; e < 64; ++e) x[e] = x[e] * rstd * __ldg(wgt + e) + __ldg(bia + e);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int pos = h == 0 ? py : px;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float c = __ldg(cos_tab + pos * 16 + j), sn = __ldg(sin_tab + pos * 16 + j);
      const float a = x[h * 32 + j], b = x[h * 32 + j + 16];
      x[h * 32 + j] = a * c - b * sn;
      x[h * 32 + j + 16] = b * c + a * sn;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr + 2 * k), "r"(pack_bf16(x[16 * k], x[16 * k + 1])),
                 "r"(pack_bf16(x[16 * k + 2], x[16 * k + 3])), "r"(pack_bf16(x[16 * k + 4], x[16 * k + 5])), "r"(pack_bf16(x[16 * k + 6], x[16 * k + 7])),
                 "r"(pack_bf16(x[16 * k + 8], x[16 * k + 9])), "r"(pack_bf16(x[16 * k + 10], x[16 * k + 11])),
                 "r"(pack_bf16(x[16 * k + 12], x[16 * k + 13])), "r"(pack_bf16(x[16 * k + 14], x[16 * k + 15]))
                 : "memory");
}

int qknorm_rope2d_entry(void* qkv, long long ld, long long rows, long long heads, const float* qw, const float* qb,
                        const float* kw, const float* kb, float eps, const float* cos_tab, const float* sin_tab,
                        long long max_pos, long long tpv, long long n_special, long long grid_w, cudaStream_t st) {
  V3A_REQUIRE(qkv && qw && qb && kw && kb && cos_tab && sin_tab, VIST3A_ERR_INVALID, "qknorm_rope2d: null pointer");
  V3A_REQUIRE(rows > 0 && heads > 0 && ld >= 3 * heads * 64 && ld % 2 == 0, VIST3A_ERR_INVALID, "qknorm_rope2d: bad sizes");
  V3A_REQUIRE(tpv > 0 && n_special >= 0 && n_special <= tpv && grid_w > 0, VIST3A_ERR_INVALID, "qknorm_rope2d: bad token layout");
  const long long npatch = tpv - n_special;
  const long long need = 1 + (npatch > 0 ? ((npatch - 1) / grid_w + 1 > grid_w ? (npatch - 1) / grid_w + 1 : grid_w) : 0);
  V3A_REQUIRE(max_pos >= need, VIST3A_ERR_INVALID, "qknorm_rope2d: rope table has %lld positions, %lld needed", max_pos, need);
  if (ld % 16 == 0 && ((uintptr_t)qkv & 31) == 0)
    qknorm_rope2d_head_kernel<<<grid_for(rows * 2 * heads, 128), 128, 0, st>>>((__nv_bfloat16*)qkv, ld, rows, (int)heads, qw, qb, kw, kb, eps, cos_tab,
                                                                              sin_tab, (int)tpv, (int)n_special, (int)grid_w);
  else
    qknorm_rope2d_kernel<<<grid_for(rows * heads, 8), 256, 0, st>>>((__nv_bfloat16*)qkv, ld, rows, (int)heads, qw, qb, kw, kb, eps,
                                                                    cos_tab, sin_tab, (int)tpv, (int)n_special, (int)grid_w);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// bilinear resize (align_corners=True), NHWC fp32, 4 channels per thread
// ----------------------------------------------------------------------------------------
// grid: x = chunks of 256 (pixel, 4-channel group) pairs along one output row, y = output row, z = image -- 32-bit index arithmetic only (the first
// version decomposed one flat 64-bit index per thread: three 64-bit divisions made the kernel issue-bound at 2.2 TB/s, ncu: 78 % issue slots)
__global__ void __launch_bounds__(256) bilinear_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int n_img,
                                                            int hi, int wi, int ho, int wo, int C, const float* __restrict__ add,
                                                            const float* __restrict__ pos_x, const float* __restrict__ pos_y,
                                                            float sy, float sx) {
  const int c4 = C >> 2;
  const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (t >= wo * c4) return;
  const int x = t / c4, c = (t - x * c4) * 4;
  const int y = (int)blockIdx.y, n = (int)blockIdx.z;
  // PyTorch area_pixel_compute_source_index(align_corners=True): src = scale * dst
  const float fy = sy * (float)y, fx = sx * (float)x;
  int y0 = (int)fy, x0 = (int)fx;
  y0 = y0 > hi - 1 ? hi - 1 : y0;
  x0 = x0 > wi - 1 ? wi - 1 : x0;
  const int y1 = y0 + (y0 < hi - 1 ? 1 : 0), x1 = x0 + (x0 < wi - 1 ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0;
  const float hy = 1.0f - ly, hx = 1.0f - lx;
  const float* base = in + (long long)n * hi * wi * C + c;
  const float* r0 = base + (long long)y0 * wi * C;
  const float* r1 = base + (long long)y1 * wi * C;
  const float4 v00 = *reinterpret_cast<const float4*>(r0 + x0 * C);
  const float4 v01 = *reinterpret_cast<const float4*>(r0 + x1 * C);
  const float4 v10 = *reinterpret_cast<const float4*>(r1 + x0 * C);
  const float4 v11 = *reinterpret_cast<const float4*>(r1 + x1 * C);
  float4 o;
  o.x = hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
  o.y = hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
  o.z = hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
  o.w = hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
  const long long off = (((long long)n * ho + y) * wo + x) * C + c;
  if (add) {
    const float4 a = *reinterpret_cast<const float4*>(add + off);
    o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
  }
  if (pos_x) {
    const int half = C / 2;
    const float4 pe = c < half ? *reinterpret_cast<const float4*>(pos_x + (long long)x * half + c)
                               : *reinterpret_cast<const float4*>(pos_y + (long long)y * half + (c - half));
    o.x += pe.x; o.y += pe.y; o.z += pe.z; o.w += pe.w;
  }
  *reinterpret_cast<float4*>(out + off) = o;
}

int bilinear_nhwc_entry(const float* in, float* out, long long n_img, long long hi, long long wi, long long ho, long long wo,
                        long long C, const float* add, const float* pos_x, const float* pos_y, cudaStream_t st) {
  V3A_REQUIRE(in && out && n_img > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0 && C > 0 && C % 8 == 0, VIST3A_ERR_INVALID,
              "bilinear_nhwc: bad arguments (C must be a multiple of 8)");
  V3A_REQUIRE((pos_x == nullptr) == (pos_y == nullptr), VIST3A_ERR_INVALID, "bilinear_nhwc: pos_x/pos_y must both be given");
  const float sy = ho > 1 ? (float)(hi - 1) / (float)(ho - 1) : 0.f;
  const float sx = wo > 1 ? (float)(wi - 1) / (float)(wo - 1) : 0.f;
  V3A_REQUIRE(ho <= 65535 && n_img <= 65535 && wo * (C / 4) < (1ll << 31), VIST3A_ERR_INVALID, "bilinear_nhwc: output rows / images exceed the grid limits");
  const dim3 grid((unsigned)((wo * (C / 4) + 255) / 256), (unsigned)ho, (unsigned)n_img);
  bilinear_nhwc_kernel<<<grid, 256, 0, st>>>(in, out, (int)n_img, (int)hi, (int)wi, (int)ho, (int)wo, (int)C, add, pos_x, pos_y, sy, sx);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// depth-to-space after a k = s transposed convolution done as a GEMM
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) depth_to_space_kernel(const float* __restrict__ in, float* __restrict__ out, int n_img, int h,
                                                             int w, int C, int k) {
  const int c4 = C / 4;
  const long long total = (long long)n_img * h * k * w * k * c4;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % c4) * 4;
  long long pix = i / c4;
  const int X = (int)(pix % (w * k)), Y = (int)((pix / (w * k)) % (h * k)), n = (int)(pix / ((long long)w * k * h * k));
  const int x = X / k, dx = X % k, y = Y / k, dy = Y % k;
  const long long src = (((long long)n * h + y) * w + x) * ((long long)k * k * C) + (long long)(dy * k + dx) * C + c;
  *reinterpret_cast<float4*>(out + pix * C + c) = *reinterpret_cast<const float4*>(in + src);
}

int depth_to_space_entry(const float* in, float* out, long long n_img, long long h, long long w, long long C, int k,
                         cudaStream_t st) {
  V3A_REQUIRE(in && out && n_img > 0 && h > 0 && w > 0 && C > 0 && C % 4 == 0 && k > 0, VIST3A_ERR_INVALID, "depth_to_space: bad arguments");
  const long long total = n_img * h * k * w * k * (C / 4);
  depth_to_space_kernel<<<grid_for(total, 256), 256, 0, st>>>(in, out, (int)n_img, (int)h, (int)w, (int)C, k);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// short-sequence fp32 attention (camera head): one warp per (batch, head, query)
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attention_small_kernel(const float* __restrict__ qkv, float* __restrict__ out, int B, int L,
                                                              int H, int D, float scale) {
  const long long wid = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (wid >= (long long)B * H * L) return;
  const int lane = threadIdx.x & 31;
  const int i = (int)(wid % L), h = (int)((wid / L) % H), b = (int)(wid / ((long long)L * H));
  const long long rs = 3LL * H * D;  // row stride of [B, L, 3, H, D]
  const float* q = qkv + ((long long)b * L + i) * rs + (long long)h * D;
  // lane j scores key j
  float sc = -INFINITY;
  if (lane < L) {
    const float* k = qkv + ((long long)b * L + lane) * rs + (long long)(H + h) * D;
    float acc = 0.f;
    for (int d = 0; d < D; d += 4) {
      const float4 a = *reinterpret_cast<const float4*>(q + d);
      const float4 kk = *reinterpret_cast<const float4*>(k + d);
      acc += a.x * kk.x + a.y * kk.y + a.z * kk.z + a.w * kk.w;
    }
    sc = acc * scale;
  }
  float m = sc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float pr = lane < L ? expf(sc - m) : 0.f;
  float s = pr;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  pr /= s;
  // lanes over head-dim columns
  for (int d = lane; d < D; d += 32) {
    float acc = 0.f;
    for (int j = 0; j < L; ++j) {
      const float pj = __shfl_sync(0xffffffffu, pr, j);
      acc += pj * qkv[((long long)b * L + j) * rs + (long long)(2 * H + h) * D + d];
    }
    out[((long long)b * L + i) * ((long long)H * D) + (long long)h * D + d] = acc;
  }
}

int attention_small_entry(const float* qkv, float* out, long long B, long long L, long long H, long long D, float scale,
                          cudaStream_t st) {
  V3A_REQUIRE(qkv && out && B > 0 && L > 0 && L <= 32 && H > 0 && D > 0 && D % 4 == 0, VIST3A_ERR_INVALID,
              "attention_small: need 1 <= L <= 32 and D %% 4 == 0 (got L=%lld D=%lld)", L, D);
  attention_small_kernel<<<grid_for(B * H * L, 4), 128, 0, st>>>(qkv, out, (int)B, (int)L, (int)H, (int)D, scale);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

__global__ void fma_rows_kernel(float* __restrict__ out, long long ldo, const float* __restrict__ a, long long lda,
                                const float* __restrict__ b, long long ldb, const float* __restrict__ c, long long ldc, long long rows,
                                int dim) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * dim) return;
  const long long r = i / dim;
  const int d = (int)(i % dim);
  out[r * ldo + d] = a[r * lda + d] * b[r * ldb + d] + c[r * ldc + d];
}

int fma_rows_entry(float* out, long long ldo, const float* a, long long lda, const float* b, long long ldb, const float* c,
                   long long ldc, long long rows, long long dim, cudaStream_t st) {
  V3A_REQUIRE(out && a && b && c && rows > 0 && dim > 0, VIST3A_ERR_INVALID, "fma_rows: bad arguments");
  fma_rows_kernel<<<grid_for(rows * dim, 256), 256, 0, st>>>(out, ldo, a, lda, b, ldb, c, ldc, rows, (int)dim);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// transposed epilogue of a swapped-operand skinny linear: the tcgen05 GEMM streams the weight matrix as its A operand
// (ct[n, m] = sum_k W[n, k] x[m, k], 16 padded token columns); this kernel finishes y[m, n] = res + gate[n] * act(ct + b[n])
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ float act_apply_f(float v, int act) {
  switch (act) {
    case VIST3A_ACT_GELU_TANH: return gelu_tanh_precise_f(v);
    case VIST3A_ACT_GELU_ERF: return gelu_erf_f(v);
    case VIST3A_ACT_SILU: return v / (1.0f + expf(-v));
    case VIST3A_ACT_RELU: return fmaxf(v, 0.f);
    default: return v;
  }
}
__global__ void __launch_bounds__(256) bias_act_t_kernel(const float* __restrict__ ct, long long ldct, const float* __restrict__ bias,
                                                         int act, const float* __restrict__ gate, const float* __restrict__ res,
                                                         long long ldr, float* __restrict__ y, long long ldy, int M, int N, int splits) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int n = (int)(i % N), m = (int)(i / N);  // consecutive threads -> consecutive n: coalesced y / bias / gate / residual
  // split-K by reshaping (see vist3a_bias_act_t): partial s of output (n, m) sits at row n * splits + s, column m * splits + s
  float v = bias ? bias[n] : 0.f;
  for (int sp = 0; sp < splits; ++sp) v += ct[((long long)n * splits + sp) * ldct + (long long)m * splits + sp];
  v = act_apply_f(v, act);
  if (gate) v *= gate[n];
  if (res) v += res[(long long)m * ldr + n];
  y[(long long)m * ldy + n] = v;
}

int bias_act_t_entry(const float* ct, long long ldct, const float* bias, int act, const float* gate, const float* res, long long ldr,
                     float* y, long long ldy, long long M, long long N, int splits, cudaStream_t st) {
  V3A_REQUIRE(ct && y && M > 0 && N > 0 && ldct >= M * (splits > 0 ? splits : 1) && ldy >= N, VIST3A_ERR_INVALID, "bias_act_t: bad arguments");
  V3A_REQUIRE(!res || ldr >= N, VIST3A_ERR_INVALID, "bias_act_t: residual stride");
  V3A_REQUIRE(splits >= 1 && splits <= 64, VIST3A_ERR_INVALID, "bias_act_t: splits must be in [1, 64]");
  bias_act_t_kernel<<<grid_for(M * N, 256), 256, 0, st>>>(ct, ldct, bias, act, gate, res, ldr, y, ldy, (int)M, (int)N, splits);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// pose encoding -> cameras
// ----------------------------------------------------------------------------------------
__global__ void pose_to_cameras_kernel(const float* __restrict__ pose_raw, float* __restrict__ pose_act, float* __restrict__ extr,
                                       float* __restrict__ intr, float* __restrict__ c2w, float* __restrict__ intr_norm, int S,
                                       float H, float W) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float p[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) p[i] = pose_raw[s * 9 + i];
  p[7] = fmaxf(p[7], 0.f);
  p[8] = fmaxf(p[8], 0.f);
  if (pose_act) {
#pragma unroll
    for (int i = 0; i < 9; ++i) pose_act[s * 9 + i] = p[i];
  }
  const float qi = p[3], qj = p[4], qk = p[5], qr = p[6];
  const float two_s = 2.0f / (qi * qi + qj * qj + qk * qk + qr * qr);
  float R[9];
  R[0] = 1 - two_s * (qj * qj + qk * qk); R[1] = two_s * (qi * qj - qk * qr); R[2] = two_s * (qi * qk + qj * qr);
  R[3] = two_s * (qi * qj + qk * qr); R[4] = 1 - two_s * (qi * qi + qk * qk); R[5] = two_s * (qj * qk - qi * qr);
  R[6] = two_s * (qi * qk - qj * qr); R[7] = two_s * (qj * qk + qi * qr); R[8] = 1 - two_s * (qi * qi + qj * qj);
  if (extr) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      extr[s * 12 + r * 4 + 0] = R[r * 3 + 0]; extr[s * 12 + r * 4 + 1] = R[r * 3 + 1]; extr[s * 12 + r * 4 + 2] = R[r * 3 + 2];
      extr[s * 12 + r * 4 + 3] = p[r];
    }
  }
  const float fy = (H * 0.5f) / (tanf(p[7] * 0.5f) + 1e-3f);
  const float fx = (W * 0.5f) / (tanf(p[8] * 0.5f) + 1e-3f);
  if (intr) {
    float* K = intr + s * 9;
    K[0] = fx; K[1] = 0; K[2] = W * 0.5f; K[3] = 0; K[4] = fy; K[5] = H * 0.5f; K[6] = 0; K[7] = 0; K[8] = 1;
  }
  if (intr_norm) {
    float* K = intr_norm + s * 9;
    K[0] = fx / W; K[1] = 0; K[2] = (W * 0.5f) / W; K[3] = 0; K[4] = fy / H; K[5] = (H * 0.5f) / H; K[6] = 0; K[7] = 0; K[8] = 1;
  }
  if (c2w) {
    // inverse of [R | t; 0 0 0 1] for an orthonormal-up-to-scale R given by a (possibly unnormalised) quaternion: R^-1 = R^T
    // exactly (two_s normalises), t' = -R^T t
    float* M = c2w + s * 16;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      M[r * 4 + 0] = R[0 * 3 + r]; M[r * 4 + 1] = R[1 * 3 + r]; M[r * 4 + 2] = R[2 * 3 + r];
      M[r * 4 + 3] = -(R[0 * 3 + r] * p[0] + R[1 * 3 + r] * p[1] + R[2 * 3 + r] * p[2]);
    }
    M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
  }
}

int pose_to_cameras_entry(const float* pose_raw, float* pose_act, float* extr, float* intr, float* c2w, float* intr_norm,
                          long long S, long long H, long long W, cudaStream_t st) {
  V3A_REQUIRE(pose_raw && S > 0 && H > 0 && W > 0, VIST3A_ERR_INVALID, "pose_to_cameras: bad arguments");
  pose_to_cameras_kernel<<<grid_for(S, 64), 64, 0, st>>>(pose_raw, pose_act, extr, intr, c2w, intr_norm, (int)S, (float)H, (float)W);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// fused per-pixel Gaussian epilogue
// ----------------------------------------------------------------------------------------
constexpr int kGeTile = 64;     // Gaussians per block
constexpr int kGeMaxRaw = 96;   // >= 8 + 3 * d_sh for sh_degree <= 4 (83) rounded up

// Gaussian adapter of one element (gaussian_adapter.py:114-147, common/gaussians.py:8-44): r = raw[0..7] = (density, scales 3, quaternion xyzw 4),
// o[3..5] scales, o[6..9] rotation (xyzw, normalised), o[10] opacity, o[11..19] covariance R S S^T R^T.
__device__ __forceinline__ void gaussian_params(const float* r, float* o) {
  float sc[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float x = r[1 + i];
    const float sp = x > 20.f ? x : log1pf(expf(x));  // F.softplus (threshold 20)
    sc[i] = fminf(0.001f * sp, 0.3f);
    o[3 + i] = sc[i];
  }
  const float qn = sqrtf(r[4] * r[4] + r[5] * r[5] + r[6] * r[6] + r[7] * r[7]) + 1e-8f;
  const float qi = r[4] / qn, qj = r[5] / qn, qk = r[6] / qn, qr = r[7] / qn;
  o[6] = qi; o[7] = qj; o[8] = qk; o[9] = qr;
  o[10] = 1.0f / (1.0f + expf(-r[0]));
  const float two_s = 2.0f / (qi * qi + qj * qj + qk * qk + qr * qr + 1e-8f);
  float R[9];
  R[0] = 1 - two_s * (qj * qj + qk * qk); R[1] = two_s * (qi * qj - qk * qr); R[2] = two_s * (qi * qk + qj * qr);
  R[3] = two_s * (qi * qj + qk * qr); R[4] = 1 - two_s * (qi * qi + qk * qk); R[5] = two_s * (qj * qk - qi * qr);
  R[6] = two_s * (qi * qk - qj * qr); R[7] = two_s * (qj * qk + qi * qr); R[8] = 1 - two_s * (qi * qi + qj * qj);
  const float s2[3] = {sc[0] * sc[0], sc[1] * sc[1], sc[2] * sc[2]};
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b)
      o[11 + a * 3 + b] = R[a * 3 + 0] * s2[0] * R[b * 3 + 0] + R[a * 3 + 1] * s2[1] * R[b * 3 + 1] + R[a * 3 + 2] * s2[2] * R[b * 3 + 2];
}

__global__ void __launch_bounds__(256) gaussian_epilogue_kernel(
    const float* __restrict__ depth_feat, long long ld_df, int cd, const float* __restrict__ depth_w, float depth_b,
    const float* __restrict__ gs_raw, long long ld_raw, const float* __restrict__ extr, const float* __restrict__ intr,
    const float* __restrict__ sh_mask, int d_sh, long long P, int HW, int Wimg, float* __restrict__ depth,
    float* __restrict__ means, float* __restrict__ scales, float* __restrict__ rot, float* __restrict__ opac,
    float* __restrict__ harm, float* __restrict__ cov, float* __restrict__ scene_sum) {
  __shared__ float s_raw[kGeTile][kGeMaxRaw + 1];
  __shared__ float s_out[kGeTile][20];  // means 3, scales 3, rot 4, opac 1, cov 9
  __shared__ float s_norm[8];
  const long long p0 = (long long)blockIdx.x * kGeTile;
  const int n_here = (int)min((long long)kGeTile, P - p0);
  const int nraw = 8 + 3 * d_sh;
  // phase 1: coalesced load of the raw tile
  for (int i = threadIdx.x; i < n_here * nraw; i += blockDim.x) {
    const int g = i / nraw, c = i % nraw;
    s_raw[g][c] = gs_raw[(p0 + g) * ld_raw + c];
  }
  __syncthreads();
  // phase 2: harmonics = raw[8:] * mask, streamed out coalesced
  const int nh = 3 * d_sh;
  for (int i = threadIdx.x; i < n_here * nh; i += blockDim.x) {
    const int g = i / nh, c = i % nh;
    harm[(p0 + g) * nh + c] = s_raw[g][8 + c] * sh_mask[c % d_sh];
  }
  // phase 3: 4 threads per Gaussian compute depth (dot over cd features) then one of them the parameters
  const int g = threadIdx.x >> 2, sub = threadIdx.x & 3;
  float nrm = 0.f;
  const bool valid = g < n_here;
  const long long pidx = p0 + (valid ? g : 0);
  float acc = 0.f;
  if (valid) {
    const float* df = depth_feat + pidx * ld_df;
    for (int c = sub; c < cd; c += 4) acc += df[c] * depth_w[c];
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (valid) {
    if (sub == 0) {
      const float d = expf(acc + depth_b);
      depth[pidx] = d;
      const int view = (int)(pidx / HW), pix = (int)(pidx % HW);
      const float u = (float)(pix % Wimg), v = (float)(pix / Wimg);
      const float* K = intr + view * 9;
      const float* E = extr + view * 12;
      const float xc = (u - K[2]) * d / K[0], yc = (v - K[5]) * d / K[4], zc = d;
      // world = R^T (cam - t)  ==  R^T cam + (-R^T t)
      const float tx = -(E[0] * E[3] + E[4] * E[7] + E[8] * E[11]);
      const float ty = -(E[1] * E[3] + E[5] * E[7] + E[9] * E[11]);
      const float tz = -(E[2] * E[3] + E[6] * E[7] + E[10] * E[11]);
      const float mx = E[0] * xc + E[4] * yc + E[8] * zc + tx;
      const float my = E[1] * xc + E[5] * yc + E[9] * zc + ty;
      const float mz = E[2] * xc + E[6] * yc + E[10] * zc + tz;
      float* o = s_out[g];
      o[0] = mx; o[1] = my; o[2] = mz;
      nrm = sqrtf(mx * mx + my * my + mz * mz);
      gaussian_params(s_raw[g], o);
    }
  }
  // block reduction of |means| for scene_scale
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
  if ((threadIdx.x & 31) == 0) s_norm[threadIdx.x >> 5] = nrm;
  __syncthreads();
  if (threadIdx.x == 0 && scene_sum) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s_norm[i];
    atomicAdd(scene_sum, t);
  }
  // phase 4: coalesced stores of the per-Gaussian records
  for (int i = threadIdx.x; i < n_here * 3; i += blockDim.x) means[p0 * 3 + i] = s_out[i / 3][i % 3];
  for (int i = threadIdx.x; i < n_here * 3; i += blockDim.x) scales[p0 * 3 + i] = s_out[i / 3][3 + i % 3];
  for (int i = threadIdx.x; i < n_here * 4; i += blockDim.x) rot[p0 * 4 + i] = s_out[i / 4][6 + i % 4];
  for (int i = threadIdx.x; i < n_here; i += blockDim.x) opac[p0 + i] = s_out[i][10];
  for (int i = threadIdx.x; i < n_here * 9; i += blockDim.x) cov[p0 * 9 + i] = s_out[i / 9][11 + i % 9];
}

int gaussian_epilogue_entry(const float* depth_feat, long long ld_df, long long cd, const float* depth_w, float depth_b,
                            const float* gs_raw, long long ld_raw, const float* extr, const float* intr, const float* sh_mask,
                            long long d_sh, long long S, long long H, long long W, float* depth, float* means, float* scales,
                            float* rot, float* opac, float* harm, float* cov, float* scene_sum, cudaStream_t st) {
  V3A_REQUIRE(depth_feat && depth_w && gs_raw && extr && intr && sh_mask && depth && means && scales && rot && opac && harm && cov,
              VIST3A_ERR_INVALID, "gaussian_epilogue: null pointer");
  V3A_REQUIRE(S > 0 && H > 0 && W > 0 && cd > 0 && ld_df >= cd && d_sh > 0 && 8 + 3 * d_sh <= kGeMaxRaw && ld_raw >= 8 + 3 * d_sh,
              VIST3A_ERR_INVALID, "gaussian_epilogue: bad sizes (d_sh=%lld)", d_sh);
  const long long P = S * H * W;
  gaussian_epilogue_kernel<<<grid_for(P, kGeTile), 256, 0, st>>>(depth_feat, ld_df, (int)cd, depth_w, depth_b, gs_raw, ld_raw, extr, intr,
                                                                 sh_mask, (int)d_sh, P, (int)(H * W), (int)W, depth, means, scales, rot,
                                                                 opac, harm, cov, scene_sum);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// Gaussian adapter on fused voxels (voxelize=True branch): positions are given, rows hold the raw_gs_dim features
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gaussian_adapter_kernel(const float* __restrict__ pts, const float* __restrict__ feats, long long ld_feats,
                                                               const float* __restrict__ sh_mask, int d_sh, long long P, float* __restrict__ means,
                                                               float* __restrict__ scales, float* __restrict__ rot, float* __restrict__ opac,
                                                               float* __restrict__ harm, float* __restrict__ cov) {
  __shared__ float s_raw[kGeTile][kGeMaxRaw + 1];
  __shared__ float s_out[kGeTile][20];
  const long long p0 = (long long)blockIdx.x * kGeTile;
  const int n_here = (int)min((long long)kGeTile, P - p0);
  const int nraw = 8 + 3 * d_sh;
  for (int i = threadIdx.x; i < n_here * nraw; i += blockDim.x) {
    const int g = i / nraw, c = i % nraw;
    s_raw[g][c] = feats[(p0 + g) * ld_feats + c];
  }
  __syncthreads();
  const int nh = 3 * d_sh;
  for (int i = threadIdx.x; i < n_here * nh; i += blockDim.x) {
    const int g = i / nh, c = i % nh;
    harm[(p0 + g) * nh + c] = s_raw[g][8 + c] * sh_mask[c % d_sh];
  }
  if (threadIdx.x < n_here) gaussian_params(s_raw[threadIdx.x], s_out[threadIdx.x]);
  __syncthreads();
  for (int i = threadIdx.x; i < n_here * 3; i += blockDim.x) means[p0 * 3 + i] = pts[p0 * 3 + i];
  for (int i = threadIdx.x; i < n_here * 3; i += blockDim.x) scales[p0 * 3 + i] = s_out[i / 3][3 + i % 3];
  for (int i = threadIdx.x; i < n_here * 4; i += blockDim.x) rot[p0 * 4 + i] = s_out[i / 4][6 + i % 4];
  for (int i = threadIdx.x; i < n_here; i += blockDim.x) opac[p0 + i] = s_out[i][10];
  for (int i = threadIdx.x; i < n_here * 9; i += blockDim.x) cov[p0 * 9 + i] = s_out[i / 9][11 + i % 9];
}

int gaussian_adapter_entry(const float* pts, const float* feats, long long ld_feats, const float* sh_mask, long long d_sh, long long P, float* means,
                           float* scales, float* rot, float* opac, float* harm, float* cov, cudaStream_t st) {
  V3A_REQUIRE(pts && feats && sh_mask && means && scales && rot && opac && harm && cov, VIST3A_ERR_INVALID, "gaussian_adapter: null pointer");
  V3A_REQUIRE(P > 0 && d_sh > 0 && 8 + 3 * d_sh <= kGeMaxRaw && ld_feats >= 8 + 3 * d_sh, VIST3A_ERR_INVALID, "gaussian_adapter: bad sizes (d_sh=%lld)", d_sh);
  int rc = check_arch();
  if (rc) return rc;
  gaussian_adapter_kernel<<<grid_for(P, kGeTile), 256, 0, st>>>(pts, feats, ld_feats, sh_mask, (int)d_sh, P, means, scales, rot, opac, harm, cov);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

}  // namespace v3a
