// HBM-bound kernels of the DiT path: LayerNorm(+modulation), RMSNorm(+RoPE), AdaLN vectors,
// skinny linear, timestep features, patchify / unpatchify, CFG combine and the scheduler's
// linear update.  All are plain coalesced / vectorised CUDA-core kernels: one warp per row for the
// normalisations (row statistics by shuffle), 16-byte accesses everywhere.
#include <atomic>

#include "common.cuh"
#include "host_util.cuh"

namespace v3a {


struct RowMap3 {
  long long rpg, gstride, goff;
  __device__ __forceinline__ long long operator()(long long row) const {
    return rpg > 0 ? (row / rpg) * gstride + goff + row % rpg : row;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// load / store 8 consecutive elements as fp32
template <bool kF32>
__device__ __forceinline__ void load8(const void* base, long long idx, float (&v)[8]) {
  if constexpr (kF32) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
    v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
  }
}
template <bool kF32>
__device__ __forceinline__ void store8(void* base, long long idx, const float (&v)[8]) {
  if constexpr (kF32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = u;
  }
}

// ----------------------------------------------------------------------------------------
// LayerNorm (+ per-batch scale / shift)
// ----------------------------------------------------------------------------------------
// Row data stays in registers in its STORAGE type (bf16: one uint4 per 8 elements) and is converted on use: 24 instead of 48 registers
// for a 1536-wide row, i.e. 10 instead of 6 resident blocks per SM, and every load of the row is issued before the first use -- the
// kernel is latency-bound (ncu: 1.4 TB/s at 31 % active warps with the fp32 register copy), so bytes in flight are what matters.
template <bool kF32>
struct Row8 {
  uint4 a, b;  // bf16: a only
  __device__ __forceinline__ void load(const void* base, long long idx) {
    if constexpr (kF32) {
      a = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(base) + idx);
      b = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(base) + idx + 4);
    } else {
      a = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    }
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    if constexpr (kF32) {
      v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
      v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
    } else {
      v[0] = bf16_lo(a.x); v[1] = bf16_hi(a.x); v[2] = bf16_lo(a.y); v[3] = bf16_hi(a.y);
      v[4] = bf16_lo(a.z); v[5] = bf16_hi(a.z); v[6] = bf16_lo(a.w); v[7] = bf16_hi(a.w);
    }
  }
};

template <int NCHUNK, bool kInF32, bool kOutF32>
__global__ void __launch_bounds__(128) layernorm_kernel(const void* __restrict__ x, long long ldx, void* __restrict__ out,
                                                        long long ldo, long long rows, int dim, long long rows_per_batch,
                                                        const float* __restrict__ mul, long long mul_bs,
                                                        const float* __restrict__ add, long long add_bs, float eps,
                                                        int mul_plus_one, RowMap3 imap, RowMap3 omap) {
  pdl_launch_dependents();
  pdl_wait();
  const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long xrow = imap(row), orow = omap(row);
  Row8<kInF32> raw[NCHUNK];
#pragma unroll
  for (int i = 0; i < NCHUNK; ++i) {
    const int col = (lane + 32 * i) * 8;
    if (col < dim) raw[i].load(x, xrow * ldx + col);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCHUNK; ++i) {
    const int col = (lane + 32 * i) * 8;
    if (col < dim) {
      float v[8];
      raw[i].get(v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[j];
    }
  }
  const float mean = warp_sum(s) / (float)dim;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NCHUNK; ++i) {
    const int col = (lane + 32 * i) * 8;
    if (col < dim) {
      float v[8];
      raw[i].get(v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; ss += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / (float)dim + eps);
  const long long b = row / rows_per_batch;
  const float* mrow = mul ? mul + b * mul_bs : nullptr;
  const float* arow = add ? add + b * add_bs : nullptr;
#pragma unroll
  for (int i = 0; i < NCHUNK; ++i) {
    const int col = (lane + 32 * i) * 8;
    if (col < dim) {
      float v[8], o[8];
      raw[i].get(v);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd;
      if (mrow) {
        float m[8];
        load8<true>(mrow, col, m);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] *= mul_plus_one ? 1.0f + m[j] : m[j];
      }
      if (arow) {
        float a[8];
        load8<true>(arow, col, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += a[j];
      }
      store8<kOutF32>(out, orow * ldo + col, o);
    }
  }
}

template <int NCHUNK>
static int launch_ln(const void* x, int xdt, long long ldx, void* out, int odt, long long ldo, long long rows, int dim,
                     long long rpb, const float* mul, long long mbs, const float* add, long long abs_, float eps,
                     int mp1, RowMap3 im, RowMap3 om, cudaStream_t st) {
  const unsigned grid = (unsigned)((rows + 3) / 4);
  const bool fi = xdt == VIST3A_DTYPE_F32, fo = odt == VIST3A_DTYPE_F32;
  cudaError_t e;
  if (fi && fo) e = launch_kernel(layernorm_kernel<NCHUNK, true, true>, dim3(grid), dim3(128), 0, st, true, 1, x, ldx, out, ldo, rows, dim, rpb, mul, mbs, add, abs_, eps, mp1, im, om);
  else if (fi) e = launch_kernel(layernorm_kernel<NCHUNK, true, false>, dim3(grid), dim3(128), 0, st, true, 1, x, ldx, out, ldo, rows, dim, rpb, mul, mbs, add, abs_, eps, mp1, im, om);
  else if (fo) e = launch_kernel(layernorm_kernel<NCHUNK, false, true>, dim3(grid), dim3(128), 0, st, true, 1, x, ldx, out, ldo, rows, dim, rpb, mul, mbs, add, abs_, eps, mp1, im, om);
  else e = launch_kernel(layernorm_kernel<NCHUNK, false, false>, dim3(grid), dim3(128), 0, st, true, 1, x, ldx, out, ldo, rows, dim, rpb, mul, mbs, add, abs_, eps, mp1, im, om);
  V3A_CUDA_OK(e);
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

int layernorm_entry(const void* x, int xdt, long long ldx, void* out, int odt, long long ldo, long long rows,
                    long long dim, long long rpb, const float* mul, long long mbs, const float* add, long long abs_,
                    float eps, int mp1, const vist3a_rowmap* in_map, const vist3a_rowmap* out_map, cudaStream_t st) {
  RowMap3 im = {0, 0, 0}, om = {0, 0, 0};
  if (in_map) im = {in_map->rpg, in_map->gstride, in_map->goff};
  if (out_map) om = {out_map->rpg, out_map->gstride, out_map->goff};
  V3A_REQUIRE(x && out, VIST3A_ERR_INVALID, "layernorm: null pointer");
  V3A_REQUIRE(rows > 0 && dim > 0 && dim % 8 == 0 && dim <= 8192, VIST3A_ERR_INVALID,
              "layernorm: dim must be a multiple of 8 and <= 8192 (got %lld)", dim);
  V3A_REQUIRE(ldx % 8 == 0 && ldo % 8 == 0 && ldx >= dim && ldo >= dim, VIST3A_ERR_INVALID, "layernorm: bad row strides");
  V3A_REQUIRE(mbs % 4 == 0 && abs_ % 4 == 0, VIST3A_ERR_INVALID, "layernorm: mul/add batch strides must be multiples of 4");
  if (rpb <= 0) rpb = rows;
  const int nchunk = (int)((dim + 255) / 256);
  if (nchunk <= 4) return launch_ln<4>(x, xdt, ldx, out, odt, ldo, rows, (int)dim, rpb, mul, mbs, add, abs_, eps, mp1, im, om, st);
  if (nchunk <= 6) return launch_ln<6>(x, xdt, ldx, out, odt, ldo, rows, (int)dim, rpb, mul, mbs, add, abs_, eps, mp1, im, om, st);
  if (nchunk <= 8) return launch_ln<8>(x, xdt, ldx, out, odt, ldo, rows, (int)dim, rpb, mul, mbs, add, abs_, eps, mp1, im, om, st);
  if (nchunk <= 16) return launch_ln<16>(x, xdt, ldx, out, odt, ldo, rows, (int)dim, rpb, mul, mbs, add, abs_, eps, mp1, im, om, st);
  if (nchunk <= 20) return launch_ln<20>(x, xdt, ldx, out, odt, ldo, rows, (int)dim, rpb, mul, mbs, add, abs_, eps, mp1, im, om, st);  // Wan-14B (5120)
  return launch_ln<32>(x, xdt, ldx, out, odt, ldo, rows, (int)dim, rpb, mul, mbs, add, abs_, eps, mp1, im, om, st);
}

// ----------------------------------------------------------------------------------------
// RMSNorm across heads (+ interleaved RoPE), in place on bf16
// ----------------------------------------------------------------------------------------
template <int NCHUNK>
__global__ void __launch_bounds__(128) rmsnorm_rope_kernel(__nv_bfloat16* __restrict__ x, long long ldx, long long rows,
                                                           int dim, int head_dim, const float* __restrict__ weight,
                                                           float eps, const float* __restrict__ rcos,
                                                           const float* __restrict__ rsin, long long rope_len, long long seg_stride) {
  pdl_launch_dependents();
  pdl_wait();
  const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  x += (long long)blockIdx.y * seg_stride;   // segment (q | k of a fused qkv buffer) with its own weight vector
  weight += (long long)blockIdx.y * dim;
  const int lane = threadIdx.x & 31;
  Row8<false> raw[NCHUNK];  // bf16 storage type in registers (see layernorm_kernel)
#pragma unroll
  for (int i = 0; i < NCHUNK; ++i) {
    const int col = (lane + 32 * i) * 8;
    if (col < dim) raw[i].load(x, row * ldx + col);
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NCHUNK; ++i) {
    const int col = (lane + 32 * i) * 8;
    if (col < dim) {
      float v[8];
      raw[i].get(v);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += v[j] * v[j];
    }
  }
  const float rinv = rsqrtf(warp_sum(ss) / (float)dim + eps);
  const long long pos = rcos ? row % rope_len : 0;
  const int half = head_dim / 2;
#pragma unroll
  for (int i = 0; i < NCHUNK; ++i) {
    const int col = (lane + 32 * i) * 8;
    if (col < dim) {
      float w[8], o[8], v[8];
      raw[i].get(v);
      load8<true>(weight, col, w);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = v[j] * rinv * w[j];
      if (rcos) {
        const int pj = (col % head_dim) / 2;  // first of 4 rotation pairs
        const float4 cs = *reinterpret_cast<const float4*>(rcos + pos * half + pj);
        const float4 sn = *reinterpret_cast<const float4*>(rsin + pos * half + pj);
        const float c4[4] = {cs.x, cs.y, cs.z, cs.w}, s4[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = o[2 * j], b = o[2 * j + 1];
          o[2 * j] = a * c4[j] - b * s4[j];
          o[2 * j + 1] = a * s4[j] + b * c4[j];
        }
      }
      store8<false>(x, row * ldx + col, o);
    }
  }
}

int rmsnorm_rope_entry(void* x, long long ldx, long long rows, long long dim, long long head_dim, const float* weight,
                       float eps, const float* rcos, const float* rsin, long long rope_len, long long nseg, long long seg_stride,
                       cudaStream_t st) {
  V3A_REQUIRE(nseg >= 1 && nseg <= 8 && (nseg == 1 || (seg_stride >= dim && seg_stride % 8 == 0 && ldx >= (nseg - 1) * seg_stride + dim)),
              VIST3A_ERR_INVALID, "rmsnorm_rope: bad segment layout");
  V3A_REQUIRE(x && weight, VIST3A_ERR_INVALID, "rmsnorm_rope: null pointer");
  V3A_REQUIRE(rows > 0 && dim > 0 && dim % 8 == 0 && dim <= 8192, VIST3A_ERR_INVALID, "rmsnorm_rope: dim %lld unsupported", dim);
  V3A_REQUIRE(head_dim % 8 == 0 && dim % head_dim == 0, VIST3A_ERR_INVALID, "rmsnorm_rope: head_dim must divide dim");
  V3A_REQUIRE(ldx % 8 == 0 && ldx >= dim, VIST3A_ERR_INVALID, "rmsnorm_rope: bad row stride");
  V3A_REQUIRE((rcos == nullptr) == (rsin == nullptr), VIST3A_ERR_INVALID, "rmsnorm_rope: cos/sin must both be given");
  if (rcos) V3A_REQUIRE(rope_len > 0, VIST3A_ERR_INVALID, "rmsnorm_rope: rope_len");
  const dim3 grid((unsigned)((rows + 3) / 4), (unsigned)nseg);
  __nv_bfloat16* xp = reinterpret_cast<__nv_bfloat16*>(x);
  const int nchunk = (int)((dim + 255) / 256);
  cudaError_t e;
  if (nchunk <= 6) e = launch_kernel(rmsnorm_rope_kernel<6>, grid, dim3(128), 0, st, true, 1, xp, ldx, rows, (int)dim, (int)head_dim, weight, eps, rcos, rsin, rope_len, seg_stride);
  else if (nchunk <= 20) e = launch_kernel(rmsnorm_rope_kernel<20>, grid, dim3(128), 0, st, true, 1, xp, ldx, rows, (int)dim, (int)head_dim, weight, eps, rcos, rsin, rope_len, seg_stride);
  else e = launch_kernel(rmsnorm_rope_kernel<32>, grid, dim3(128), 0, st, true, 1, xp, ldx, rows, (int)dim, (int)head_dim, weight, eps, rcos, rsin, rope_len, seg_stride);
  V3A_CUDA_OK(e);
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// per-row RMS reciprocal (read-only pass): out[r] = rsqrt(mean(x[r,:]^2) + eps)
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) row_rinv_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, long long rows, int dim, float eps,
                                                       float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float ss = 0.f;
  for (int col = lane * 8; col < dim; col += 256) {
    float v[8];
    load8<false>(x, row * ldx + col, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) ss = fmaf(v[j], v[j], ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) out[row] = rsqrtf(ss / (float)dim + eps);
}

int row_rinv_entry(const void* x, long long ldx, long long rows, long long dim, float eps, float* out, cudaStream_t st) {
  V3A_REQUIRE(x && out && rows > 0 && dim > 0 && dim % 8 == 0 && ldx % 8 == 0 && ldx >= dim, VIST3A_ERR_INVALID, "row_rinv: bad arguments");
  V3A_CUDA_OK(launch_kernel(row_rinv_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, true, 1,
                            reinterpret_cast<const __nv_bfloat16*>(x), ldx, rows, (int)dim, eps, out));
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// AdaLN modulation vectors
// ----------------------------------------------------------------------------------------
template <bool kModF32>
__global__ void modulation_kernel(const float* __restrict__ table, const void* __restrict__ mod, int broadcast,
                                  float* __restrict__ out, long long batch, int nvec, int dim, unsigned one_plus) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = batch * nvec * dim;
  if (i >= total) return;
  const int d = (int)(i % dim);
  const int j = (int)((i / dim) % nvec);
  const long long b = i / ((long long)dim * nvec);
  const long long mi = broadcast ? b * dim + d : i;
  float m;
  if constexpr (kModF32) m = reinterpret_cast<const float*>(mod)[mi];
  else m = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(mod)[mi]);
  float v = table[(long long)j * dim + d] + m;
  if ((one_plus >> j) & 1u) v += 1.0f;
  out[i] = v;
}

int modulation_entry(const float* table, const void* mod, int mod_dt, int bc, float* out, long long batch,
                     long long nvec, long long dim, unsigned one_plus, cudaStream_t st) {
  V3A_REQUIRE(table && mod && out, VIST3A_ERR_INVALID, "modulation: null pointer");
  V3A_REQUIRE(batch > 0 && nvec > 0 && nvec <= 32 && dim > 0, VIST3A_ERR_INVALID, "modulation: bad sizes");
  const long long total = batch * nvec * dim;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (mod_dt == VIST3A_DTYPE_F32) modulation_kernel<true><<<grid, 256, 0, st>>>(table, mod, bc, out, batch, (int)nvec, (int)dim, one_plus);
  else modulation_kernel<false><<<grid, 256, 0, st>>>(table, mod, bc, out, batch, (int)nvec, (int)dim, one_plus);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// skinny linear: y[m,n] = act(sum_k pre(x[m,k]) W[n,k] + b[n]),  M <= 16.  One warp per output column.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case VIST3A_ACT_GELU_TANH: return gelu_tanh_precise_f(v);
    case VIST3A_ACT_GELU_ERF: return gelu_erf_f(v);
    case VIST3A_ACT_SILU: return v / (1.0f + expf(-v));
    case VIST3A_ACT_RELU: return fmaxf(v, 0.f);
    default: return v;
  }
}

template <int MT, bool kXF32, bool kWF32>
__global__ void __launch_bounds__(256) skinny_linear_kernel(const void* __restrict__ x, long long ldx,
                                                            const void* __restrict__ W, long long ldw,
                                                            const float* __restrict__ bias, void* __restrict__ y,
                                                            int y_f32, long long ldy, int M, int N, int K, int pre_act,
                                                            int act, const float* __restrict__ gate,
                                                            const float* __restrict__ residual, long long ldres) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  float acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = 0.f;
  for (int k = lane * 8; k < K; k += 256) {
    float w[8];
    load8<kWF32>(W, (long long)n * ldw + k, w);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (m < M) {
        float xv[8];
        load8<kXF32>(x, (long long)m * ldx + k, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xx = pre_act ? act_apply(xv[j], pre_act) : xv[j];
          acc[m] = fmaf(xx, w[j], acc[m]);
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = warp_sum(acc[m]);
  if (lane == 0) {
    const float b = bias ? bias[n] : 0.f;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (m < M) {
        float v = act_apply(acc[m] + b, act);
        if (gate) v *= gate[n];
        if (residual) v += residual[(long long)m * ldres + n];
        if (y_f32) reinterpret_cast<float*>(y)[(long long)m * ldy + n] = v;
        else reinterpret_cast<__nv_bfloat16*>(y)[(long long)m * ldy + n] = __float2bfloat16_rn(v);
      }
    }
  }
}

int skinny_linear_entry(const void* x, int xdt, long long ldx, const void* W, int wdt, long long ldw, const float* bias,
                        void* y, int ydt, long long ldy, long long M, long long N, long long K, int pre_act, int act,
                        const float* gate, const float* residual, long long ldres, cudaStream_t st) {
  V3A_REQUIRE(x && W && y, VIST3A_ERR_INVALID, "skinny_linear: null pointer");
  V3A_REQUIRE(M > 0 && M <= 16 && N > 0 && K > 0 && K % 8 == 0, VIST3A_ERR_INVALID,
              "skinny_linear: need 1 <= M <= 16 and K %% 8 == 0 (got M=%lld K=%lld)", M, K);
  V3A_REQUIRE(ldx % 8 == 0 && ldw % 8 == 0 && ldx >= K && ldw >= K, VIST3A_ERR_INVALID, "skinny_linear: bad strides");
  const unsigned grid = (unsigned)((N + 7) / 8);
  const bool xf = xdt == VIST3A_DTYPE_F32, wf = wdt == VIST3A_DTYPE_F32;
  const int yf = ydt == VIST3A_DTYPE_F32;
#define V3A_SKINNY(MT)                                                                                                             \
  do {                                                                                                                             \
    if (xf && wf) skinny_linear_kernel<MT, true, true><<<grid, 256, 0, st>>>(x, ldx, W, ldw, bias, y, yf, ldy, (int)M, (int)N, (int)K, pre_act, act, gate, residual, ldres);   \
    else if (xf) skinny_linear_kernel<MT, true, false><<<grid, 256, 0, st>>>(x, ldx, W, ldw, bias, y, yf, ldy, (int)M, (int)N, (int)K, pre_act, act, gate, residual, ldres);   \
    else if (wf) skinny_linear_kernel<MT, false, true><<<grid, 256, 0, st>>>(x, ldx, W, ldw, bias, y, yf, ldy, (int)M, (int)N, (int)K, pre_act, act, gate, residual, ldres);   \
    else skinny_linear_kernel<MT, false, false><<<grid, 256, 0, st>>>(x, ldx, W, ldw, bias, y, yf, ldy, (int)M, (int)N, (int)K, pre_act, act, gate, residual, ldres);          \
  } while (0)
  if (M <= 2) V3A_SKINNY(2);
  else if (M <= 4) V3A_SKINNY(4);
  else if (M <= 8) V3A_SKINNY(8);
  else V3A_SKINNY(16);
#undef V3A_SKINNY
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// timestep features
// ----------------------------------------------------------------------------------------
__global__ void timestep_features_kernel(const float* __restrict__ t, void* __restrict__ out, int out_f32, int batch,
                                         int dim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * dim) return;
  const int b = i / dim, j = i % dim, half = dim / 2;
  const int fi = j < half ? j : j - half;
  const float freq = expf(-9.210340371976184f * (float)fi / (float)half);
  const float a = t[b] * freq;
  const float v = j < half ? cosf(a) : sinf(a);
  if (out_f32) reinterpret_cast<float*>(out)[i] = v;
  else reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
}

int timestep_features_entry(const float* t, void* out, int odt, long long batch, long long dim, cudaStream_t st) {
  V3A_REQUIRE(t && out && batch > 0 && dim > 0 && dim % 2 == 0, VIST3A_ERR_INVALID, "timestep_features: bad arguments");
  const int total = (int)(batch * dim);
  timestep_features_kernel<<<(total + 255) / 256, 256, 0, st>>>(t, out, odt == VIST3A_DTYPE_F32, (int)batch, (int)dim);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// patchify / unpatchify (Wan patch size (1,2,2))
// ----------------------------------------------------------------------------------------
template <bool kF32>
__global__ void patchify_kernel(const void* __restrict__ x, __nv_bfloat16* __restrict__ A, int B, int C, int T, int H,
                                int W) {
  // one thread per (token, c): 2x2 input pixels -> 4 consecutive bf16 of A
  const int Hp = H / 2, Wp = W / 2;
  const long long total = (long long)B * T * Hp * Wp * C;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  long long tok = i / C;
  const int wp = (int)(tok % Wp); tok /= Wp;
  const int hp = (int)(tok % Hp); tok /= Hp;
  const int t = (int)(tok % T);
  const int b = (int)(tok / T);
  const long long base = ((((long long)b * C + c) * T + t) * H + hp * 2) * W + wp * 2;
  float v[4];
  if constexpr (kF32) {
    const float* p = reinterpret_cast<const float*>(x);
    v[0] = p[base]; v[1] = p[base + 1]; v[2] = p[base + W]; v[3] = p[base + W + 1];
  } else {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(x);
    v[0] = __bfloat162float(p[base]); v[1] = __bfloat162float(p[base + 1]);
    v[2] = __bfloat162float(p[base + W]); v[3] = __bfloat162float(p[base + W + 1]);
  }
  uint2 o;
  o.x = pack_bf16(v[0], v[1]);
  o.y = pack_bf16(v[2], v[3]);
  *reinterpret_cast<uint2*>(A + i * 4) = o;
}

int patchify_entry(const void* x, int xdt, void* A, long long B, long long C, long long T, long long H, long long W,
                   cudaStream_t st) {
  V3A_REQUIRE(x && A && B > 0 && C > 0 && T > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, VIST3A_ERR_INVALID,
              "patchify: bad arguments");
  const long long total = B * T * (H / 2) * (W / 2) * C;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (xdt == VIST3A_DTYPE_F32) patchify_kernel<true><<<grid, 256, 0, st>>>(x, reinterpret_cast<__nv_bfloat16*>(A), (int)B, (int)C, (int)T, (int)H, (int)W);
  else patchify_kernel<false><<<grid, 256, 0, st>>>(x, reinterpret_cast<__nv_bfloat16*>(A), (int)B, (int)C, (int)T, (int)H, (int)W);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

template <bool kInF32, bool kOutF32>
__global__ void unpatchify_kernel(const void* __restrict__ P, long long ldp, void* __restrict__ out, int B, int C, int T,
                                  int H, int W) {
  // one thread per output element (coalesced writes); reads are strided but tiny (L2 resident)
  const long long total = (long long)B * C * T * H * W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int w = (int)(i % W);
  long long r = i / W;
  const int h = (int)(r % H); r /= H;
  const int t = (int)(r % T); r /= T;
  const int c = (int)(r % C);
  const int b = (int)(r / C);
  const int Hp = H / 2, Wp = W / 2;
  const long long tok = (((long long)b * T + t) * Hp + h / 2) * Wp + w / 2;
  const int col = ((h & 1) * 2 + (w & 1)) * C + c;
  float v;
  if constexpr (kInF32) v = reinterpret_cast<const float*>(P)[tok * ldp + col];
  else v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P)[tok * ldp + col]);
  if constexpr (kOutF32) reinterpret_cast<float*>(out)[i] = v;
  else reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
}

int unpatchify_entry(const void* P, int pdt, long long ldp, void* out, int odt, long long B, long long C, long long T,
                     long long H, long long W, cudaStream_t st) {
  V3A_REQUIRE(P && out && B > 0 && C > 0 && T > 0 && H % 2 == 0 && W % 2 == 0 && ldp >= 4 * C, VIST3A_ERR_INVALID,
              "unpatchify: bad arguments");
  const long long total = B * C * T * H * W;
  const unsigned grid = (unsigned)((total + 255) / 256);
  const bool fi = pdt == VIST3A_DTYPE_F32, fo = odt == VIST3A_DTYPE_F32;
  if (fi && fo) unpatchify_kernel<true, true><<<grid, 256, 0, st>>>(P, ldp, out, (int)B, (int)C, (int)T, (int)H, (int)W);
  else if (fi) unpatchify_kernel<true, false><<<grid, 256, 0, st>>>(P, ldp, out, (int)B, (int)C, (int)T, (int)H, (int)W);
  else if (fo) unpatchify_kernel<false, true><<<grid, 256, 0, st>>>(P, ldp, out, (int)B, (int)C, (int)T, (int)H, (int)W);
  else unpatchify_kernel<false, false><<<grid, 256, 0, st>>>(P, ldp, out, (int)B, (int)C, (int)T, (int)H, (int)W);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

// ----------------------------------------------------------------------------------------
// CFG combine and linear multistep update
// ----------------------------------------------------------------------------------------
template <bool kF32>
__global__ void cfg_combine_kernel(const void* __restrict__ cond, const void* __restrict__ uncond, float g,
                                   float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float c, u;
  if constexpr (kF32) { c = reinterpret_cast<const float*>(cond)[i]; u = reinterpret_cast<const float*>(uncond)[i]; }
  else {
    c = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(cond)[i]);
    u = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(uncond)[i]);
  }
  out[i] = u + g * (c - u);
}

int cfg_combine_entry(const void* cond, const void* uncond, int dt, float g, float* out, long long n, cudaStream_t st) {
  V3A_REQUIRE(cond && uncond && out && n > 0, VIST3A_ERR_INVALID, "cfg_combine: bad arguments");
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dt == VIST3A_DTYPE_F32) cfg_combine_kernel<true><<<grid, 256, 0, st>>>(cond, uncond, g, out, n);
  else cfg_combine_kernel<false><<<grid, 256, 0, st>>>(cond, uncond, g, out, n);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

struct AxpbyArgs {
  const float* term[8];
  float coeff[8];
  int n_terms;
};
__global__ void axpby_kernel(float* __restrict__ out, const AxpbyArgs a, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < a.n_terms) v = fmaf(a.coeff[k], a.term[k][i], v);
  out[i] = v;
}

int axpby_entry(float* out, int n_terms, const float* const* terms, const float* coeffs, long long n, cudaStream_t st) {
  V3A_REQUIRE(out && terms && coeffs && n > 0 && n_terms > 0 && n_terms <= 8, VIST3A_ERR_INVALID, "axpby_n: bad arguments");
  AxpbyArgs a;
  a.n_terms = n_terms;
  for (int k = 0; k < 8; ++k) { a.term[k] = k < n_terms ? terms[k] : nullptr; a.coeff[k] = k < n_terms ? coeffs[k] : 0.f; }
  for (int k = 0; k < n_terms; ++k) V3A_REQUIRE(a.term[k], VIST3A_ERR_INVALID, "axpby_n: null term %d", k);
  axpby_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, a, n);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

}  // namespace v3a
