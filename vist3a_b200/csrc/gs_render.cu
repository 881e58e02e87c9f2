// 3D-Gaussian rasteriser (forward), the consumer of the decoder's Gaussians:
//   replaces gsplat 1.4.0 `rasterization(..., covars=, sh_degree=4, render_mode="RGB+D", rasterize_mode="classic", radius_clip=0.1,
//   near_plane=1e-10, packed=False)` as called once per view from DecoderSplattingCUDA.rendering_fn
//   (AS/model/decoder/decoder_splatting_cuda.py:43-125, call at :92-112).
// gsplat is a third-party CUDA package that is absent here (requirements.txt:17 pins gsplat==1.4.0); its published algorithm is
// restated (oracle/gsplat_ref.py carries the same statement, parity unpinned):
//   1. fully_fused_projection: camera-space mean / covariance, perspective projection with the clamped Jacobian (limits
//      (W - cx)/fx + 0.3 tan_fovx ...), eps2d = 0.3 added to the 2-D covariance diagonal, conic = inverse, radius =
//      ceil(3 sqrt(larger eigenvalue)), culling (near / far, det <= 0, radius <= radius_clip, bounding box outside the image)
//   2. view-dependent colour: real spherical harmonics up to degree 4 on the normalised direction mean - camera centre, + 0.5, clamped at 0
//   3. isect_tiles: 16x16 tiles touched by the +-radius box, key = tile id | float bits of the depth, sorted (stable LSD radix sort)
//   4. rasterize_to_pixels: per pixel centre (+0.5), front to back: sigma = 1/2 d^T conic d, alpha = min(0.999, opacity exp(-sigma)), skipped
//      below 1/255, stop when the transmittance would fall to 1e-4; colour / depth accumulated with alpha T, background added with the
//      remaining transmittance, alpha = 1 - T
// Two phases because the number of tile intersections is data dependent: vist3a_gs_project leaves it on the device, the caller reads it
// (as gsplat does) and sizes the intersection workspace of vist3a_gs_rasterize.
#include "common.cuh"
#include "host_util.cuh"
#include "radix_sort.cuh"

namespace v3a {

namespace {

constexpr int kTile = 16;
constexpr int kGeom = 10;          // per Gaussian: x, y, depth, conic a/b/c, opacity, r, g, b
constexpr int kScanBlk = 2048;     // items per block of the intersection-count scan

struct GsCamera {
  float R[9];      // world -> camera rotation, row-major
  float t[3];
  float campos[3];
  float fx, fy, cx, cy;
  int W, H, tiles_x, tiles_y;
  float near_plane, far_plane, radius_clip, eps2d;
};

// Real spherical harmonics up to degree 4 (Sloan, "Efficient Spherical Harmonic Evaluation"; the formulation of gsplat's
// sh_coeffs_to_color_fast).  c = coefficients of one colour channel, (x, y, z) a unit vector.
__device__ __forceinline__ float sh_eval(int deg, const float* __restrict__ c, float x, float y, float z) {
  float r = 0.2820947917738781f * c[0];
  if (deg >= 1) {
    r += 0.48860251190292f * (-y * c[1] + z * c[2] - x * c[3]);
    if (deg >= 2) {
      const float z2 = z * z;
      const float fTmp0B = -1.092548430592079f * z;
      const float fC1 = x * x - y * y, fS1 = 2.f * x * y;
      const float pSH6 = 0.9461746957575601f * z2 - 0.3153915652525201f;
      r += 0.5462742152960395f * fS1 * c[4] + fTmp0B * y * c[5] + pSH6 * c[6] + fTmp0B * x * c[7] + 0.5462742152960395f * fC1 * c[8];
      if (deg >= 3) {
        const float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
        const float fTmp1B = 1.445305721320277f * z;
        const float fC2 = x * fC1 - y * fS1, fS2 = x * fS1 + y * fC1;
        const float pSH12 = z * (1.865881662950577f * z2 - 1.119528997770346f);
        r += -0.5900435899266435f * fS2 * c[9] + fTmp1B * fS1 * c[10] + fTmp0C * y * c[11] + pSH12 * c[12] + fTmp0C * x * c[13] +
             fTmp1B * fC1 * c[14] - 0.5900435899266435f * fC2 * c[15];
        if (deg >= 4) {
          const float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
          const float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
          const float fTmp2B = -1.770130769779931f * z;
          const float fC3 = x * fC2 - y * fS2, fS3 = x * fS2 + y * fC2;
          const float pSH20 = 1.984313483298443f * z * pSH12 - 1.006230589874905f * pSH6;
          r += 0.6258357354491763f * fS3 * c[16] + fTmp2B * fS2 * c[17] + fTmp1C * fS1 * c[18] + fTmp0D * y * c[19] + pSH20 * c[20] +
               fTmp0D * x * c[21] + fTmp1C * fC1 * c[22] + fTmp2B * fC2 * c[23] + 0.6258357354491763f * fC3 * c[24];
        }
      }
    }
  }
  return r;
}

// ---- 1 + 2: projection, culling, colour, tile box -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gs_project_kernel(const float* __restrict__ means, const float* __restrict__ covars, const float* __restrict__ opac,
                                                         const float* __restrict__ harm, int d_sh, int sh_degree, long long N, GsCamera cam,
                                                         float* __restrict__ geom, uint2* __restrict__ bbox, unsigned* __restrict__ touched) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  unsigned ntiles = 0;
  uint2 box = make_uint2(0u, 0u);
  do {
    const float mx = means[3 * i], my = means[3 * i + 1], mz = means[3 * i + 2];
    const float* R = cam.R;
    const float x = R[0] * mx + R[1] * my + R[2] * mz + cam.t[0];
    const float y = R[3] * mx + R[4] * my + R[5] * mz + cam.t[1];
    const float z = R[6] * mx + R[7] * my + R[8] * mz + cam.t[2];
    if (!(z >= cam.near_plane) || z > cam.far_plane) break;
    // covariance to camera space: R S R^T
    float S[9], RS[9], Cc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) S[k] = covars[9 * i + k];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) RS[a * 3 + b] = R[a * 3] * S[b] + R[a * 3 + 1] * S[3 + b] + R[a * 3 + 2] * S[6 + b];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) Cc[a * 3 + b] = RS[a * 3] * R[b * 3] + RS[a * 3 + 1] * R[b * 3 + 1] + RS[a * 3 + 2] * R[b * 3 + 2];
    // perspective projection with the clamped Jacobian
    const float tan_fovx = 0.5f * cam.W / cam.fx, tan_fovy = 0.5f * cam.H / cam.fy;
    const float lim_x_pos = (cam.W - cam.cx) / cam.fx + 0.3f * tan_fovx, lim_x_neg = cam.cx / cam.fx + 0.3f * tan_fovx;
    const float lim_y_pos = (cam.H - cam.cy) / cam.fy + 0.3f * tan_fovy, lim_y_neg = cam.cy / cam.fy + 0.3f * tan_fovy;
    const float rz = 1.f / z, rz2 = rz * rz;
    const float tx = z * fminf(lim_x_pos, fmaxf(-lim_x_neg, x * rz));
    const float ty = z * fminf(lim_y_pos, fmaxf(-lim_y_neg, y * rz));
    // J = [fx rz, 0, -fx tx rz2; 0, fy rz, -fy ty rz2]
    const float j00 = cam.fx * rz, j02 = -cam.fx * tx * rz2, j11 = cam.fy * rz, j12 = -cam.fy * ty * rz2;
    // cov2d = J Cc J^T
    const float u0 = j00 * Cc[0] + j02 * Cc[6], u1 = j00 * Cc[1] + j02 * Cc[7], u2 = j00 * Cc[2] + j02 * Cc[8];
    const float v0 = j11 * Cc[3] + j12 * Cc[6], v1_ = j11 * Cc[4] + j12 * Cc[7], v2 = j11 * Cc[5] + j12 * Cc[8];
    float c00 = u0 * j00 + u2 * j02, c01 = u1 * j11 + u2 * j12, c10 = v0 * j00 + v2 * j02, c11 = v1_ * j11 + v2 * j12;
    const float m2x = cam.fx * x * rz + cam.cx, m2y = cam.fy * y * rz + cam.cy;
    c00 += cam.eps2d;
    c11 += cam.eps2d;
    const float det = c00 * c11 - c01 * c10;
    if (!(det > 0.f)) break;
    const float inv_det = 1.f / det;
    const float ca = c11 * inv_det, cb = -c01 * inv_det, cc = c00 * inv_det;  // conic = inverse(cov2d): (a, b, c)
    const float bmid = 0.5f * (c00 + c11);
    const float ev = bmid + sqrtf(fmaxf(0.01f, bmid * bmid - det));
    const float radius = ceilf(3.f * sqrtf(ev));
    if (radius <= cam.radius_clip) break;
    if (m2x + radius <= 0.f || m2x - radius >= (float)cam.W || m2y + radius <= 0.f || m2y - radius >= (float)cam.H) break;
    // tiles touched by the +-radius box (gsplat isect_tiles)
    const float tile_radius = radius / (float)kTile, tcx = m2x / (float)kTile, tcy = m2y / (float)kTile;
    const unsigned x0 = (unsigned)min(max(0, (int)floorf(tcx - tile_radius)), cam.tiles_x);
    const unsigned x1 = (unsigned)min(max(0, (int)ceilf(tcx + tile_radius)), cam.tiles_x);
    const unsigned y0 = (unsigned)min(max(0, (int)floorf(tcy - tile_radius)), cam.tiles_y);
    const unsigned y1 = (unsigned)min(max(0, (int)ceilf(tcy + tile_radius)), cam.tiles_y);
    ntiles = (x1 - x0) * (y1 - y0);
    if (ntiles == 0) break;
    box = make_uint2(x0 | (x1 << 16), y0 | (y1 << 16));
    // colour
    float dx = mx - cam.campos[0], dy = my - cam.campos[1], dz = mz - cam.campos[2];
    const float inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
    dx *= inorm; dy *= inorm; dz *= inorm;
    float* g = geom + i * kGeom;
    g[0] = m2x; g[1] = m2y; g[2] = z; g[3] = ca; g[4] = cb; g[5] = cc; g[6] = opac[i];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) g[7 + ch] = fmaxf(sh_eval(sh_degree, harm + (i * 3 + ch) * d_sh, dx, dy, dz) + 0.5f, 0.f);
  } while (false);
  bbox[i] = box;
  touched[i] = ntiles;
}

// ---- exclusive scan of the per-Gaussian tile counts (block partials, one-block scan, apply) ------------------------------------------------
__global__ void __launch_bounds__(256) gs_scan_partial_kernel(const unsigned* __restrict__ in, long long N, unsigned* __restrict__ partial) {
  const long long base = (long long)blockIdx.x * kScanBlk;
  unsigned c = 0;
#pragma unroll
  for (int k = 0; k < kScanBlk / 256; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < N) c += in[i];
  }
  __shared__ unsigned ws[8];
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) gs_scan_top_kernel(unsigned* __restrict__ partial, int n, long long* __restrict__ total_out) {
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned long long carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned v = i < n ? partial[i] : 0u;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const unsigned w = warp_tot[lane];
      unsigned wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;
    }
    __syncthreads();
    const unsigned long long excl = carry_s + warp_tot[wid] + incl - v;
    if (i < n) partial[i] = (unsigned)excl;   // offsets stay below 2^32 (checked by the caller through the total)
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = (long long)carry_s;
}

__global__ void __launch_bounds__(256) gs_scan_apply_kernel(const unsigned* __restrict__ in, long long N, const unsigned* __restrict__ partial,
                                                            unsigned* __restrict__ out) {
  // thread t owns 8 consecutive items
  const long long i0 = (long long)blockIdx.x * kScanBlk + (long long)threadIdx.x * 8;
  unsigned v[8], s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v[k] = (i0 + k < N) ? in[i0 + k] : 0u;
    s += v[k];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __shared__ unsigned ws[8];
  if (lane == 31) ws[wid] = incl;
  __syncthreads();
  unsigned woff = 0;
  for (int w = 0; w < wid; ++w) woff += ws[w];
  unsigned run = partial[blockIdx.x] + woff + incl - s;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (i0 + k < N) out[i0 + k] = run;
    run += v[k];
  }
}

// ---- 3: tile intersections ------------------------------------------------------------------------------------------------------------------
__global__ void gs_set_passes_kernel(int* npasses, int v) {
  if (threadIdx.x == 0) *npasses = v;
}

__global__ void __launch_bounds__(256) gs_emit_kernel(const float* __restrict__ geom, const uint2* __restrict__ bbox, const unsigned* __restrict__ touched,
                                                      const unsigned* __restrict__ offsets, long long N, int tiles_x, unsigned long long* __restrict__ keys,
                                                      unsigned* __restrict__ vals) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N || touched[i] == 0) return;
  const uint2 b = bbox[i];
  const unsigned x0 = b.x & 0xffffu, x1 = b.x >> 16, y0 = b.y & 0xffffu, y1 = b.y >> 16;
  const unsigned long long depth_bits = (unsigned long long)__float_as_uint(geom[i * kGeom + 2]);  // depth > 0: the bit pattern orders like the value
  unsigned o = offsets[i];
  for (unsigned ty = y0; ty < y1; ++ty)
    for (unsigned tx = x0; tx < x1; ++tx) {
      keys[o] = ((unsigned long long)(ty * (unsigned)tiles_x + tx) << 32) | depth_bits;
      vals[o] = (unsigned)i;
      ++o;
    }
}

__global__ void __launch_bounds__(256) gs_tile_ranges_kernel(const unsigned long long* __restrict__ keys_a, const unsigned long long* __restrict__ keys_b,
                                                             const int* __restrict__ npasses, long long n, unsigned* __restrict__ tile_start,
                                                             unsigned* __restrict__ tile_end) {
  const unsigned long long* keys = (*npasses & 1) ? keys_b : keys_a;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned tile = (unsigned)(keys[i] >> 32);
  if (i == 0 || (unsigned)(keys[i - 1] >> 32) != tile) tile_start[tile] = (unsigned)i;
  if (i == n - 1 || (unsigned)(keys[i + 1] >> 32) != tile) tile_end[tile] = (unsigned)(i + 1);
}

__global__ void gs_zero_kernel(unsigned* __restrict__ p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0u;
}

// ---- 4: per-tile front-to-back compositing ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile* kTile) gs_rasterize_kernel(const float* __restrict__ geom, const unsigned* __restrict__ vals_a,
                                                                   const unsigned* __restrict__ vals_b, const int* __restrict__ npasses,
                                                                   const unsigned* __restrict__ tile_start, const unsigned* __restrict__ tile_end, int W, int H,
                                                                   int tiles_x, float bg0, float bg1, float bg2, float* __restrict__ rgb,
                                                                   float* __restrict__ depth, float* __restrict__ alpha) {
  const unsigned* vals = (*npasses & 1) ? vals_b : vals_a;
  // batch of 256 Gaussians in shared memory as three float4 per Gaussian: the test of a Gaussian against a pixel needs the first two
  // (position, opacity, conic); colour and depth are read only when it contributes
  __shared__ float4 s_xyod[kTile * kTile];   // x, y, opacity, depth
  __shared__ float4 s_con[kTile * kTile];    // conic a, b, c
  __shared__ float4 s_rgb[kTile * kTile];
  const int tile = blockIdx.y * tiles_x + blockIdx.x;
  const int px_i = blockIdx.x * kTile + (threadIdx.x % kTile), py_i = blockIdx.y * kTile + (threadIdx.x / kTile);
  const float px = (float)px_i + 0.5f, py = (float)py_i + 0.5f;
  const bool inside = px_i < W && py_i < H;
  bool done = !inside;
  const unsigned r0 = tile_start[tile], r1 = tile_end[tile];
  float T = 1.f, acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accd = 0.f;
  for (unsigned base = r0; base < r1; base += kTile * kTile) {
    if (__syncthreads_count(done) == kTile * kTile) break;  // every pixel of the tile is saturated
    const unsigned j = base + threadIdx.x;
    if (j < r1) {
      const float* g = geom + (long long)vals[j] * kGeom;
      s_xyod[threadIdx.x] = make_float4(g[0], g[1], g[6], g[2]);
      s_con[threadIdx.x] = make_float4(g[3], g[4], g[5], 0.f);
      s_rgb[threadIdx.x] = make_float4(g[7], g[8], g[9], 0.f);
    }
    __syncthreads();
    const int cnt = (int)min((unsigned)(kTile * kTile), r1 - base);
    for (int t = 0; t < cnt && !done; ++t) {
      const float4 xo = s_xyod[t], cn = s_con[t];
      const float dx = xo.x - px, dy = xo.y - py;
      const float sigma = 0.5f * (cn.x * dx * dx + cn.z * dy * dy) + cn.y * dx * dy;
      const float a = fminf(0.999f, xo.z * __expf(-sigma));
      if (sigma < 0.f || a < 1.f / 255.f) continue;
      const float next_T = T * (1.f - a);
      if (next_T <= 1e-4f) {
        done = true;
        break;
      }
      const float vis = a * T;
      const float4 c = s_rgb[t];
      acc0 += c.x * vis; acc1 += c.y * vis; acc2 += c.z * vis; accd += xo.w * vis;
      T = next_T;
    }
  }
  if (inside) {
    const long long p = (long long)py_i * W + px_i;
    rgb[p * 3] = acc0 + T * bg0;
    rgb[p * 3 + 1] = acc1 + T * bg1;
    rgb[p * 3 + 2] = acc2 + T * bg2;
    depth[p] = accd;
    alpha[p] = 1.f - T;
  }
}

inline long long align256(long long x) { return (x + 255) & ~255ll; }

struct ProjWs {
  float* geom;
  uint2* bbox;
  unsigned* touched;
  unsigned* offsets;
  unsigned* partial;
  long long bytes;
  int nscan;
};
ProjWs carve_proj(void* ws, long long N) {
  ProjWs w;
  w.nscan = (int)((N + kScanBlk - 1) / kScanBlk);
  char* p = reinterpret_cast<char*>(ws);
  long long off = 0;
  auto take = [&](long long b) { char* q = p ? p + off : nullptr; off += align256(b); return q; };
  w.geom = reinterpret_cast<float*>(take(4ll * kGeom * N));
  w.bbox = reinterpret_cast<uint2*>(take(8 * N));
  w.touched = reinterpret_cast<unsigned*>(take(4 * N));
  w.offsets = reinterpret_cast<unsigned*>(take(4 * N));
  w.partial = reinterpret_cast<unsigned*>(take(4ll * w.nscan));
  w.bytes = off;
  return w;
}

struct IsectWs {
  int* npasses;
  unsigned long long *keys_a, *keys_b;
  unsigned *vals_a, *vals_b;
  unsigned *block_hist, *digit_total, *tile_start, *tile_end;
  long long bytes;
};
IsectWs carve_isect(void* ws, long long n, long long n_tiles) {
  IsectWs w;
  const long long nb = (n + kSortTile - 1) / kSortTile;
  char* p = reinterpret_cast<char*>(ws);
  long long off = 0;
  auto take = [&](long long b) { char* q = p ? p + off : nullptr; off += align256(b); return q; };
  w.npasses = reinterpret_cast<int*>(take(4));
  w.keys_a = reinterpret_cast<unsigned long long*>(take(8 * n));
  w.keys_b = reinterpret_cast<unsigned long long*>(take(8 * n));
  w.vals_a = reinterpret_cast<unsigned*>(take(4 * n));
  w.vals_b = reinterpret_cast<unsigned*>(take(4 * n));
  w.block_hist = reinterpret_cast<unsigned*>(take(4ll * 256 * nb));
  w.digit_total = reinterpret_cast<unsigned*>(take(4 * 256));
  w.tile_start = reinterpret_cast<unsigned*>(take(4 * n_tiles));
  w.tile_end = reinterpret_cast<unsigned*>(take(4 * n_tiles));
  w.bytes = off;
  return w;
}

}  // namespace

long long gs_project_workspace_bytes(long long N) { return N > 0 ? carve_proj(nullptr, N).bytes : 0; }
long long gs_rasterize_workspace_bytes(long long n_isect, long long W, long long H) {
  const long long tiles = ((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile);
  return carve_isect(nullptr, n_isect > 0 ? n_isect : 1, tiles).bytes;
}

int gs_project_entry(const float* means, const float* covars, const float* opac, const float* harm, long long d_sh, int sh_degree, long long N,
                     const float* viewmat, const float* K, long long W, long long H, float near_plane, float far_plane, float radius_clip, float eps2d,
                     void* workspace, long long workspace_bytes, long long* n_isect, cudaStream_t st) {
  V3A_REQUIRE(means && covars && opac && harm && viewmat && K && workspace && n_isect, VIST3A_ERR_INVALID, "gs_project: null pointer");
  V3A_REQUIRE(N > 0 && N < (1ll << 31), VIST3A_ERR_INVALID, "gs_project: N must be in (0, 2^31)");
  V3A_REQUIRE(W > 0 && H > 0 && W <= 16 * 65535 && H <= 16 * 65535, VIST3A_ERR_INVALID, "gs_project: bad image size");
  V3A_REQUIRE(sh_degree >= 0 && sh_degree <= 4 && d_sh >= (sh_degree + 1) * (sh_degree + 1), VIST3A_ERR_INVALID,
              "gs_project: sh_degree must be in [0, 4] and d_sh >= (sh_degree + 1)^2 (got %d, %lld)", sh_degree, d_sh);
  // the sort key holds the float bits of the depth, which order like the value only for depth >= 0
  V3A_REQUIRE(near_plane >= 0.f && far_plane > near_plane, VIST3A_ERR_INVALID, "gs_project: need 0 <= near_plane < far_plane");
  V3A_REQUIRE(((uintptr_t)workspace & 255) == 0, VIST3A_ERR_INVALID, "gs_project: workspace must be 256-byte aligned");
  const ProjWs w = carve_proj(workspace, N);
  V3A_REQUIRE(workspace_bytes >= w.bytes, VIST3A_ERR_INVALID, "gs_project: workspace of %lld bytes needed, %lld given", w.bytes, workspace_bytes);
  int rc = check_arch();
  if (rc) return rc;
  GsCamera cam;
  // viewmat: row-major 4x4 world -> camera;  camera centre = -R^T t
  for (int a = 0; a < 3; ++a) {
    for (int b = 0; b < 3; ++b) cam.R[a * 3 + b] = viewmat[a * 4 + b];
    cam.t[a] = viewmat[a * 4 + 3];
  }
  for (int a = 0; a < 3; ++a) cam.campos[a] = -(cam.R[a] * cam.t[0] + cam.R[3 + a] * cam.t[1] + cam.R[6 + a] * cam.t[2]);
  cam.fx = K[0]; cam.cx = K[2]; cam.fy = K[4]; cam.cy = K[5];
  cam.W = (int)W; cam.H = (int)H;
  cam.tiles_x = (int)((W + kTile - 1) / kTile); cam.tiles_y = (int)((H + kTile - 1) / kTile);
  cam.near_plane = near_plane; cam.far_plane = far_plane; cam.radius_clip = radius_clip; cam.eps2d = eps2d;
  gs_project_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(means, covars, opac, harm, (int)d_sh, sh_degree, N, cam, w.geom, w.bbox, w.touched);
  gs_scan_partial_kernel<<<w.nscan, 256, 0, st>>>(w.touched, N, w.partial);
  gs_scan_top_kernel<<<1, 1024, 0, st>>>(w.partial, w.nscan, n_isect);
  gs_scan_apply_kernel<<<w.nscan, 256, 0, st>>>(w.touched, N, w.partial, w.offsets);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(4);
  return VIST3A_OK;
}

int gs_rasterize_entry(const void* proj_workspace, long long N, long long n_isect, long long W, long long H, const float* background,
                       void* workspace, long long workspace_bytes, float* rgb, float* depth, float* alpha, cudaStream_t st) {
  V3A_REQUIRE(proj_workspace && workspace && rgb && depth && alpha && background, VIST3A_ERR_INVALID, "gs_rasterize: null pointer");
  V3A_REQUIRE(N > 0 && n_isect >= 0 && n_isect < (1ll << 32) - kSortTile && W > 0 && H > 0, VIST3A_ERR_INVALID,
              "gs_rasterize: bad sizes (n_isect %lld must stay below 2^32)", n_isect);
  V3A_REQUIRE(((uintptr_t)workspace & 255) == 0, VIST3A_ERR_INVALID, "gs_rasterize: workspace must be 256-byte aligned");
  const ProjWs pw = carve_proj(const_cast<void*>(proj_workspace), N);
  const int tiles_x = (int)((W + kTile - 1) / kTile), tiles_y = (int)((H + kTile - 1) / kTile);
  const long long n_tiles = (long long)tiles_x * tiles_y;
  const long long n = n_isect > 0 ? n_isect : 1;
  const IsectWs w = carve_isect(workspace, n, n_tiles);
  V3A_REQUIRE(workspace_bytes >= w.bytes, VIST3A_ERR_INVALID, "gs_rasterize: workspace of %lld bytes needed, %lld given", w.bytes, workspace_bytes);
  int rc = check_arch();
  if (rc) return rc;
  long long launches = 0;
  gs_zero_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, st>>>(w.tile_start, (int)n_tiles);  // tiles nothing touches: empty range [0, 0)
  gs_zero_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, st>>>(w.tile_end, (int)n_tiles);
  launches += 2;
  int tile_bits = 0;
  while ((1ll << tile_bits) < n_tiles) ++tile_bits;
  const int passes = (32 + tile_bits + 7) / 8;
  gs_set_passes_kernel<<<1, 32, 0, st>>>(w.npasses, n_isect > 0 ? passes : 0);
  launches += 1;
  if (n_isect > 0) {
    gs_emit_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(pw.geom, pw.bbox, pw.touched, pw.offsets, N, tiles_x, w.keys_a, w.vals_a);
    V3A_CUDA_OK(radix_sort_enqueue(w.keys_a, w.keys_b, w.vals_a, w.vals_b, n_isect, passes, w.npasses, w.block_hist, w.digit_total, st));
    gs_tile_ranges_kernel<<<(unsigned)((n_isect + 255) / 256), 256, 0, st>>>(w.keys_a, w.keys_b, w.npasses, n_isect, w.tile_start, w.tile_end);
    launches += 2 + 3 * passes;
  }
  gs_rasterize_kernel<<<dim3(tiles_x, tiles_y), kTile * kTile, 0, st>>>(pw.geom, w.vals_a, w.vals_b, w.npasses, w.tile_start, w.tile_end, (int)W, (int)H,
                                                                        tiles_x, background[0], background[1], background[2], rgb, depth, alpha);
  launches += 1;
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(launches);
  return VIST3A_OK;
}

}  // namespace v3a
