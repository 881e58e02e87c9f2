// Confidence-quantile branches of the stitched decoder on the device (HBM-bound index work, no tensor cores):
//   render_conf   conf_valid = torch.quantile(depth_conf, conf_threshold); keep the pixels with depth_conf > conf_valid, per batch element
//                 compacted in (view, row, column) order                                  (models/anysplat_stitched.py:381-387, 442-455)
//   opacity_conf  opacity *= sigmoid(depth_conf - quantile)[mask]                         (:463-467)
//   depth_conf    = 1 + exp(second output channel of the depth head)                      (AS/.../heads/head_act.py:102-103, "expp1")
// Pipeline: depth_conf_kernel (one pass over the depth head's 32-wide feature rows) -> order-preserving 32-bit keys -> the shared stable LSD
// radix sort (4 passes of 8 bits) -> the two order statistics around rank q (n - 1), interpolated as torch.quantile / torch.lerp do ->
// flag count per 2048-row block, scan of the block counts, ordered scatter of the kept rows (features, points, damping factor).
#include "common.cuh"
#include "host_util.cuh"
#include "radix_sort.cuh"

namespace v3a {

namespace {

constexpr int kSelTile = 2048;   // rows per block of the compaction (256 threads x 8)

__global__ void __launch_bounds__(256) depth_conf_kernel(const float* __restrict__ feat, long long ld, int C, const float* __restrict__ w, float bias,
                                                         float* __restrict__ conf, long long P) {
  __shared__ float sw[64];
  if (threadIdx.x < C) sw[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const float4* r = reinterpret_cast<const float4*>(feat + p * ld);
    float acc = bias;
    for (int c = 0; c < C / 4; ++c) {
      const float4 v = r[c];
      acc += v.x * sw[4 * c] + v.y * sw[4 * c + 1] + v.z * sw[4 * c + 2] + v.w * sw[4 * c + 3];
    }
    conf[p] = 1.0f + expf(acc);
  }
}

__global__ void __launch_bounds__(256) float_keys_kernel(const float* __restrict__ x, long long n, unsigned long long* __restrict__ keys,
                                                         unsigned* __restrict__ vals) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned b = __float_as_uint(x[i]);
    keys[i] = (unsigned long long)(b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u));   // unsigned order == float order (NaNs sort last)
    vals[i] = (unsigned)i;
  }
}

__global__ void set_int_kernel(int* p, int v) {
  if (threadIdx.x == 0) *p = v;
}

// out = torch.lerp(sorted[lo], sorted[hi], w):  w < 0.5 ? a + w (b - a) : b - (b - a)(1 - w)   (ATen lerp)
__global__ void quantile_pick_kernel(const unsigned long long* __restrict__ keys, long long lo, long long hi, float w, float* __restrict__ out) {
  if (threadIdx.x == 0) {
    auto val = [&](long long i) {
      const unsigned k = (unsigned)keys[i];
      return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
    };
    const float a = val(lo), b = val(hi);
    const float d = b - a;
    *out = w < 0.5f ? a + w * d : b - d * (1.0f - w);
  }
}

__global__ void __launch_bounds__(256) sel_count_kernel(const float* __restrict__ conf, const float* __restrict__ thr, int use_thr, long long n,
                                                        unsigned* __restrict__ block_count) {
  const float t = *thr;
  const long long base = (long long)blockIdx.x * kSelTile;
  unsigned c = 0;
  for (int k = 0; k < kSelTile / 256; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < n && (!use_thr || conf[i] > t)) ++c;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ unsigned s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tot = 0;
    for (int w = 0; w < 8; ++w) tot += s[w];
    block_count[blockIdx.x] = tot;
  }
}

// exclusive scan of the block counts in place (single block; <= 2^22 blocks), total -> *count
__global__ void __launch_bounds__(1024) sel_scan_kernel(unsigned* __restrict__ block_count, int nblocks, long long* __restrict__ count) {
  __shared__ unsigned s[1024];
  unsigned carry = 0;
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned v = i < nblocks ? block_count[i] : 0u;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const unsigned t = threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblocks) block_count[i] = carry + s[threadIdx.x] - v;
    const unsigned tot = s[1023];
    __syncthreads();
    carry += tot;
  }
  if (threadIdx.x == 0) *count = (long long)carry;
}

// ordered scatter: kept row i goes to position block offset + rank inside the block (ranks by ballot + warp prefix, in row order)
__global__ void __launch_bounds__(256) sel_scatter_kernel(const float* __restrict__ conf, const float* __restrict__ thr, int use_thr, long long n,
                                                          const unsigned* __restrict__ block_off, const float* __restrict__ feats, long long ld_feats, int C,
                                                          const float* __restrict__ pts, float* __restrict__ out_feats, float* __restrict__ out_pts,
                                                          float* __restrict__ out_damp) {
  const float t = *thr;
  const long long base = (long long)blockIdx.x * kSelTile;
  __shared__ unsigned warp_cnt[8];
  __shared__ unsigned run;      // kept rows of this block placed so far
  __shared__ unsigned dst_of[256];
  if (threadIdx.x == 0) run = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = 0; k < kSelTile / 256; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    const bool keep = i < n && (!use_thr || conf[i] > t);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    unsigned before = run;
    for (int w = 0; w < wid; ++w) before += warp_cnt[w];
    const unsigned pos = before + __popc(bal & ((1u << lane) - 1u));
    dst_of[threadIdx.x] = keep ? block_off[blockIdx.x] + pos : 0xffffffffu;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned tot = 0;
      for (int w = 0; w < 8; ++w) tot += warp_cnt[w];
      run += tot;
    }
    // rows are moved by the whole block: thread t copies element t, t + 256, ... of the 256 x C row block (coalesced reads of the kept rows)
    for (int e = threadIdx.x; e < 256 * C; e += 256) {
      const int rr = e / C, cc = e - rr * C;
      const unsigned d = dst_of[rr];
      if (d != 0xffffffffu) out_feats[(long long)d * C + cc] = feats[(base + k * 256 + rr) * ld_feats + cc];
    }
    if (keep) {
      const unsigned d = dst_of[threadIdx.x];
      out_pts[3ll * d] = pts[3 * i];
      out_pts[3ll * d + 1] = pts[3 * i + 1];
      out_pts[3ll * d + 2] = pts[3 * i + 2];
      if (out_damp) out_damp[d] = 1.0f / (1.0f + expf(-(conf[i] - t)));
    }
    __syncthreads();
  }
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

int depth_conf_entry(const float* feat, long long ld, long long C, const float* w, float bias, float* conf, long long P, cudaStream_t st) {
  V3A_REQUIRE(feat && w && conf && P > 0 && C > 0 && C <= 64 && C % 4 == 0 && ld >= C && ld % 4 == 0 && ((uintptr_t)feat & 15) == 0, VIST3A_ERR_INVALID,
              "depth_conf: feature rows of <= 64 channels (multiple of 4), 16-byte aligned");
  const long long blocks = (P + 255) / 256;
  depth_conf_kernel<<<(unsigned)(blocks < num_sms() * 16 ? blocks : num_sms() * 16), 256, 0, st>>>(feat, ld, (int)C, w, bias, conf, P);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

long long quantile_workspace_bytes(long long n) {
  const long long nblocks = (n + kSortTile - 1) / kSortTile;
  return (long long)(2 * align256(8 * (size_t)n) + 2 * align256(4 * (size_t)n) + align256(4 * 256 * (size_t)nblocks) + align256(4 * 256) + 256);
}

// out (device float) = torch.quantile(x, q) of n floats with linear interpolation: rank = q (n - 1) in fp32, as ATen computes it
int quantile_entry(const float* x, long long n, float q, float* out, void* ws, long long ws_bytes, cudaStream_t st) {
  V3A_REQUIRE(x && out && ws && n > 0 && n < (1ll << 31), VIST3A_ERR_INVALID, "quantile: bad arguments");
  V3A_REQUIRE(q >= 0.f && q <= 1.f, VIST3A_ERR_INVALID, "quantile: q must be in [0, 1]");
  V3A_REQUIRE(ws_bytes >= quantile_workspace_bytes(n) && ((uintptr_t)ws & 255) == 0, VIST3A_ERR_INVALID, "quantile: workspace too small or not 256-byte aligned");
  char* p = reinterpret_cast<char*>(ws);
  auto take = [&](size_t bytes) { char* r = p; p += align256(bytes); return r; };
  const long long nblocks = (n + kSortTile - 1) / kSortTile;
  unsigned long long* keys_a = (unsigned long long*)take(8 * (size_t)n);
  unsigned long long* keys_b = (unsigned long long*)take(8 * (size_t)n);
  unsigned* vals_a = (unsigned*)take(4 * (size_t)n);
  unsigned* vals_b = (unsigned*)take(4 * (size_t)n);
  unsigned* block_hist = (unsigned*)take(4 * 256 * (size_t)nblocks);
  unsigned* digit_total = (unsigned*)take(4 * 256);
  int* npasses = (int*)take(4);
  const long long blocks = (n + 255) / 256;
  float_keys_kernel<<<(unsigned)(blocks < num_sms() * 16 ? blocks : num_sms() * 16), 256, 0, st>>>(x, n, keys_a, vals_a);
  set_int_kernel<<<1, 32, 0, st>>>(npasses, 4);
  V3A_CUDA_OK(radix_sort_enqueue(keys_a, keys_b, vals_a, vals_b, n, 4, npasses, block_hist, digit_total, st));   // 4 passes: result back in keys_a
  // ATen: ranks = q * (n - 1) (fp32), lower = floor, upper = ceil, weight = ranks - lower
  const float rank = q * (float)(n - 1);
  const float lo_f = floorf(rank);
  long long lo = (long long)lo_f, hi = (long long)ceilf(rank);
  if (hi > n - 1) hi = n - 1;
  quantile_pick_kernel<<<1, 32, 0, st>>>(keys_a, lo, hi, rank - lo_f, out);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(3 + 12);
  return VIST3A_OK;
}

long long compact_rows_workspace_bytes(long long n) { return (long long)align256(4 * (size_t)((n + kSelTile - 1) / kSelTile)) + 256; }

// rows i with conf[i] > *thr (all rows when use_thr == 0), in order: out_feats [count, C], out_pts [count, 3], out_damp [count] = sigmoid(conf - *thr)
// (optional); *count (device int64) receives the number of kept rows
int compact_rows_entry(const float* conf, const float* thr, int use_thr, long long n, const float* feats, long long ld_feats, long long C, const float* pts,
                       float* out_feats, float* out_pts, float* out_damp, long long* count, void* ws, long long ws_bytes, cudaStream_t st) {
  V3A_REQUIRE(conf && thr && feats && pts && out_feats && out_pts && count && ws && n > 0 && n < (1ll << 31) && C > 0 && C <= 4096 && ld_feats >= C,
              VIST3A_ERR_INVALID, "compact_rows: bad arguments");
  V3A_REQUIRE(ws_bytes >= compact_rows_workspace_bytes(n), VIST3A_ERR_INVALID, "compact_rows: workspace too small");
  const int nblocks = (int)((n + kSelTile - 1) / kSelTile);
  unsigned* block_count = reinterpret_cast<unsigned*>(ws);
  sel_count_kernel<<<nblocks, 256, 0, st>>>(conf, thr, use_thr, n, block_count);
  sel_scan_kernel<<<1, 1024, 0, st>>>(block_count, nblocks, count);
  sel_scatter_kernel<<<nblocks, 256, 0, st>>>(conf, thr, use_thr, n, block_count, feats, ld_feats, (int)C, pts, out_feats, out_pts, out_damp);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(3);
  return VIST3A_OK;
}

}  // namespace v3a
