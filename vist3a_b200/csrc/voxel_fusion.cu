// Voxelised Gaussian fusion on the device (HBM-bound integer / index work, no tensor cores):
//   replaces EncoderAnySplat.voxelizaton_with_fusion  (AS/model/encoder/anysplat.py:298-335, called from
//   models/anysplat_stitched.py:419-440 when cfg.voxelize is set -- the released AnySplat configs set it, voxel_size 0.002)
//
//   voxel = round_half_even(p / voxel_size) as int32 per axis                       (anysplat.py:304)
//   unique voxels in lexicographic (x, y, z) order, inverse index, counts           (torch.unique(dim=0), :305-307)
//   per voxel: w_i = exp(conf_i - max conf) / (sum_j exp(conf_j - max conf) + 1e-6) (:313-319)
//              voxel_pts = sum_i w_i p_i ;  voxel_feats = sum_i w_i f_i             (:322-333)
//
// Pipeline (every stage streams its arrays once, coalesced; nothing returns to the host):
//   1. voxel_range_kernel      min / max voxel coordinate per axis                              (reads pts)
//   2. voxel_key_kernel        order-preserving 64-bit key: the three offset coordinates packed with just the bits their
//                              ranges need (x most significant), value = point index            (reads pts)
//   3. LSD radix sort, 8-bit digits, only ceil(bits / 8) passes execute (later launches exit at once): per pass a
//      per-block digit histogram, one exclusive scan over the digit-major (digit, block) table, and a STABLE scatter
//      (warp-private digit counters + match.any ranking), so members of a voxel stay in point order
//   4. segment heads -> voxel ids (block partials + scan), inverse / counts / segment starts
//   5. voxel_reduce_kernel     one warp per voxel: softmax weights over the members' confidences, then the 3 + C weighted
//                              sums in member (= point) order; lanes own output channels, member rows are read coalesced
#include <limits.h>

#include "common.cuh"
#include "host_util.cuh"
#include "radix_sort.cuh"

namespace v3a {

namespace {

constexpr int kScanTile = 2048;                         // items per block of the segment-head scan (256 threads x 8)

struct VoxelHeader {   // lives at the start of the workspace
  int mn[3], mx[3];
  int bits[3];         // bits per axis
  int total_bits;
  int npasses;         // radix passes that execute
  int n_voxels;
  int error;           // 1: coordinate ranges need more than 64 key bits
};

__device__ __forceinline__ int voxel_coord(float p, float voxel_size) {
  // (p / voxel_size).round().int(): IEEE division, round half to even, then float -> int32 (saturating, NaN -> 0)
  return __float2int_rz(rintf(__fdiv_rn(p, voxel_size)));
}

__global__ void voxel_init_kernel(VoxelHeader* h) {
  if (threadIdx.x < 3) {
    h->mn[threadIdx.x] = INT_MAX;
    h->mx[threadIdx.x] = INT_MIN;
  }
  if (threadIdx.x == 0) {
    h->n_voxels = 0;
    h->error = 0;
  }
}

__global__ void __launch_bounds__(256) voxel_range_kernel(const float* __restrict__ pts, long long N, float voxel_size, VoxelHeader* h) {
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
  // 4 points = 12 floats = three 16-byte loads per thread and iteration
  const float4* p4 = reinterpret_cast<const float4*>(pts);
  const long long groups = N / 4;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
    const float4 a = p4[3 * g], b = p4[3 * g + 1], c = p4[3 * g + 2];
    const float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const int q = voxel_coord(v[k], voxel_size);
      mn[k % 3] = min(mn[k % 3], q);
      mx[k % 3] = max(mx[k % 3], q);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(N - 4 * groups)) {  // up to 3 tail points
    const long long i = 4 * groups + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int q = voxel_coord(pts[3 * i + k], voxel_size);
      mn[k] = min(mn[k], q);
      mx[k] = max(mx[k], q);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    mn[k] = __reduce_min_sync(0xffffffffu, mn[k]);
    mx[k] = __reduce_max_sync(0xffffffffu, mx[k]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (mn[k] != INT_MAX) atomicMin(&h->mn[k], mn[k]);
      if (mx[k] != INT_MIN) atomicMax(&h->mx[k], mx[k]);
    }
  }
}

__global__ void voxel_bits_kernel(VoxelHeader* h) {
  if (threadIdx.x == 0) {
    int total = 0;
    for (int k = 0; k < 3; ++k) {
      const unsigned long long range = (unsigned long long)((long long)h->mx[k] - (long long)h->mn[k]);
      const int b = range == 0 ? 0 : 64 - __clzll(range);
      h->bits[k] = b;
      total += b;
    }
    h->total_bits = total;
    h->error = total > 64 ? 1 : 0;
    h->npasses = total > 64 ? 0 : (total + 7) / 8;
  }
}

__global__ void __launch_bounds__(256) voxel_key_kernel(const float* __restrict__ pts, long long N, float voxel_size, const VoxelHeader* __restrict__ h,
                                                        unsigned long long* __restrict__ keys, unsigned* __restrict__ vals) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int by = h->bits[1], bz = h->bits[2];
  const unsigned long long x = (unsigned long long)((long long)voxel_coord(pts[3 * i + 0], voxel_size) - h->mn[0]);
  const unsigned long long y = (unsigned long long)((long long)voxel_coord(pts[3 * i + 1], voxel_size) - h->mn[1]);
  const unsigned long long z = (unsigned long long)((long long)voxel_coord(pts[3 * i + 2], voxel_size) - h->mn[2]);
  // shifts of 64 are undefined: by + bz <= 64 always, and x == 0 whenever bits[0] == 0
  const int sxy = by + bz;
  unsigned long long key = z;
  if (bz < 64) key |= y << bz;
  if (sxy < 64) key |= x << sxy;
  keys[i] = key;
  vals[i] = (unsigned)i;
}

// ------------------------------------------------------------------------------------------------
// segment heads -> voxel ids
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ const unsigned long long* sorted_keys(const VoxelHeader* h, const unsigned long long* a, const unsigned long long* b) {
  return (h->npasses & 1) ? b : a;
}
__device__ __forceinline__ const unsigned* sorted_vals(const VoxelHeader* h, const unsigned* a, const unsigned* b) { return (h->npasses & 1) ? b : a; }

__global__ void __launch_bounds__(256) seg_count_kernel(const unsigned long long* __restrict__ keys_a, const unsigned long long* __restrict__ keys_b, long long N,
                                                        const VoxelHeader* __restrict__ h, unsigned* __restrict__ partial) {
  const unsigned long long* keys = sorted_keys(h, keys_a, keys_b);
  const long long base = (long long)blockIdx.x * kScanTile;
  unsigned c = 0;
#pragma unroll
  for (int k = 0; k < kScanTile / 256; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < N) c += (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
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

// one block: exclusive scan of the block partials; the total is the number of voxels
__global__ void __launch_bounds__(1024) seg_scan_partials_kernel(unsigned* __restrict__ partial, int n, VoxelHeader* h, long long* __restrict__ n_voxels_out,
                                                                 unsigned* __restrict__ seg_start, long long N) {
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned carry_s;
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
    const unsigned excl = carry_s + warp_tot[wid] + incl - v;
    if (i < n) partial[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const unsigned nv = h->error ? 0u : carry_s;
    h->n_voxels = (int)nv;
    if (n_voxels_out) *n_voxels_out = h->error ? -1ll : (long long)nv;
    seg_start[nv] = (unsigned)N;  // sentinel: end of the last segment
  }
}

__global__ void __launch_bounds__(256) seg_assign_kernel(const unsigned long long* __restrict__ keys_a, const unsigned long long* __restrict__ keys_b,
                                                         const unsigned* __restrict__ vals_a, const unsigned* __restrict__ vals_b, long long N,
                                                         const VoxelHeader* __restrict__ h, const unsigned* __restrict__ partial, unsigned* __restrict__ seg_start,
                                                         unsigned* __restrict__ seg_of, int* __restrict__ inverse) {
  const unsigned long long* keys = sorted_keys(h, keys_a, keys_b);
  const unsigned* vals = sorted_vals(h, vals_a, vals_b);
  // thread t owns 8 consecutive sorted positions
  const long long i0 = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * 8;
  unsigned head[8];
  unsigned s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const long long i = i0 + k;
    head[k] = (i < N && (i == 0 || keys[i] != keys[i - 1])) ? 1u : 0u;
    s += head[k];
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
  unsigned run = partial[blockIdx.x] + woff + incl - s;  // heads before position i0
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const long long i = i0 + k;
    if (i < N) {
      run += head[k];
      const unsigned seg = run - 1u;
      if (head[k]) seg_start[seg] = (unsigned)i;
      seg_of[i] = seg;
      if (inverse) inverse[vals[i]] = (int)seg;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// per-voxel softmax-weighted fusion.  One warp owns 32 consecutive voxels = one contiguous range of sorted members:
//   phase 1  lane-per-voxel softmax statistics (max, denominator) in member order; voxels with one member need none
//   phase 2  the members of the range, 32 at a time: lane-per-member weights, then four members in flight at once -- their rows
//            (3 + C values, lanes own channels: coalesced 128-byte reads) are fetched before any is accumulated, so a warp keeps
//            twelve independent loads outstanding instead of one dependent chain per voxel; the accumulator is flushed
//            (coalesced row store) whenever the member's voxel changes.  Sums run in member (= point) order.
// ------------------------------------------------------------------------------------------------
template <int NCH>  // channel rounds: 3 + C <= 32 * NCH
__global__ void __launch_bounds__(256) voxel_reduce_kernel(const float* __restrict__ pts, const float* __restrict__ feats, long long ld_feats, int C,
                                                           const float* __restrict__ conf, long long conf_stride, const unsigned* __restrict__ vals_a,
                                                           const unsigned* __restrict__ vals_b, const VoxelHeader* __restrict__ h,
                                                           const unsigned* __restrict__ seg_start, const unsigned* __restrict__ seg_of,
                                                           float* __restrict__ voxel_pts, float* __restrict__ voxel_feats, int* __restrict__ counts) {
  constexpr unsigned kFull = 0xffffffffu;
  const unsigned* vals = sorted_vals(h, vals_a, vals_b);
  const long long nv = h->n_voxels;
  const int lane = threadIdx.x & 31;
  const long long nbatch = (nv + 31) / 32;
  for (long long bt = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); bt < nbatch; bt += (long long)gridDim.x * 8) {
    const long long v0 = bt * 32;
    const int nvalid = (int)min(32ll, nv - v0);
    const bool vok = lane < nvalid;
    const unsigned s0 = seg_start[v0 + (vok ? lane : nvalid - 1)];
    const unsigned s1 = seg_start[v0 + (vok ? lane : nvalid - 1) + 1];
    const unsigned n = vok ? s1 - s0 : 0u;
    if (counts && vok) counts[v0 + lane] = (int)n;
    // ---- phase 1
    float m = 0.f, den = 1.0f;  // one member: exp(0) = 1
    const unsigned nmax = __reduce_max_sync(kFull, n);
    if (nmax > 1 && nmax <= 64) {
      if (n > 1) {
        m = -INFINITY;
        for (unsigned j = s0; j < s1; ++j) m = fmaxf(m, conf[(long long)vals[j] * conf_stride]);
        den = 0.f;
        for (unsigned j = s0; j < s1; ++j) den += expf(conf[(long long)vals[j] * conf_stride] - m);
      }
    } else if (nmax > 64) {  // long segments: the warp shares each voxel's members
      for (int k = 0; k < nvalid; ++k) {
        const unsigned a = __shfl_sync(kFull, s0, k), b = __shfl_sync(kFull, s1, k);
        if (b - a <= 1) continue;
        float mk = -INFINITY;
        for (unsigned j = a + lane; j < b; j += 32) mk = fmaxf(mk, conf[(long long)vals[j] * conf_stride]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mk = fmaxf(mk, __shfl_xor_sync(kFull, mk, o));
        float sk = 0.f;
        for (unsigned j = a + lane; j < b; j += 32) sk += expf(conf[(long long)vals[j] * conf_stride] - mk);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sk += __shfl_xor_sync(kFull, sk, o);
        if (lane == k) { m = mk; den = sk; }
      }
    }
    den += 1e-6f;
    // ---- phase 2
    const unsigned r0 = __shfl_sync(kFull, s0, 0);
    const unsigned r1 = __shfl_sync(kFull, s1, nvalid - 1);
    float acc[NCH];
#pragma unroll
    for (int r = 0; r < NCH; ++r) acc[r] = 0.f;
    int cur = 0;  // voxel (relative to v0) the accumulator belongs to
    auto flush = [&](int rel) {
      const long long v = v0 + rel;
#pragma unroll
      for (int r = 0; r < NCH; ++r) {
        const int c = lane + 32 * r;
        if (c < 3) voxel_pts[v * 3 + c] = acc[r];
        else if (c < 3 + C) voxel_feats[v * C + (c - 3)] = acc[r];
        acc[r] = 0.f;
      }
    };
    for (unsigned base = r0; base < r1; base += 32) {
      const unsigned j = base + lane;
      const bool ok = j < r1;
      const long long idx = ok ? (long long)vals[j] : 0ll;
      const int rel = ok ? (int)((long long)seg_of[j] - v0) : 0;
      const unsigned nn = __shfl_sync(kFull, n, rel);
      const float mm = __shfl_sync(kFull, m, rel), dd = __shfl_sync(kFull, den, rel);
      const float e = (ok && nn > 1) ? expf(conf[idx * conf_stride] - mm) : 1.0f;
      const float w = __fdiv_rn(e, dd);
      const int cnt = (int)min(32u, r1 - base);
      for (int k = 0; k < cnt; k += 4) {
        float x[4][NCH];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (k + u < cnt) {
            const long long iu = __shfl_sync(kFull, idx, k + u);
#pragma unroll
            for (int r = 0; r < NCH; ++r) {
              const int c = lane + 32 * r;
              x[u][r] = c < 3 ? pts[iu * 3 + c] : (c < 3 + C ? feats[iu * ld_feats + (c - 3)] : 0.f);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (k + u < cnt) {
            const int ru = __shfl_sync(kFull, rel, k + u);
            const float wu = __shfl_sync(kFull, w, k + u);
            if (ru != cur) {  // warp-uniform
              flush(cur);
              cur = ru;
            }
#pragma unroll
            for (int r = 0; r < NCH; ++r) acc[r] = __fadd_rn(acc[r], __fmul_rn(x[u][r], wu));
          }
        }
      }
    }
    if (r1 > r0) flush(cur);
  }
}

inline long long align256(long long x) { return (x + 255) & ~255ll; }

struct VoxelWorkspace {
  VoxelHeader* header;
  unsigned long long *keys_a, *keys_b;
  unsigned *vals_a, *vals_b;
  unsigned* block_hist;
  unsigned* digit_total;
  unsigned* partial;
  unsigned* seg_start;
  unsigned* seg_of;
  long long bytes;
  int nblocks, nscan;
};

VoxelWorkspace carve(void* ws, long long N) {
  VoxelWorkspace w;
  w.nblocks = (int)((N + kSortTile - 1) / kSortTile);
  w.nscan = (int)((N + kScanTile - 1) / kScanTile);
  char* p = reinterpret_cast<char*>(ws);
  long long off = 0;
  auto take = [&](long long bytes) { char* q = p ? p + off : nullptr; off += align256(bytes); return q; };
  w.header = reinterpret_cast<VoxelHeader*>(take(sizeof(VoxelHeader)));
  w.keys_a = reinterpret_cast<unsigned long long*>(take(8 * N));
  w.keys_b = reinterpret_cast<unsigned long long*>(take(8 * N));
  w.vals_a = reinterpret_cast<unsigned*>(take(4 * N));
  w.vals_b = reinterpret_cast<unsigned*>(take(4 * N));
  w.block_hist = reinterpret_cast<unsigned*>(take(4ll * 256 * w.nblocks));
  w.digit_total = reinterpret_cast<unsigned*>(take(4 * 256));
  w.partial = reinterpret_cast<unsigned*>(take(4ll * w.nscan));
  w.seg_start = reinterpret_cast<unsigned*>(take(4 * (N + 1)));
  w.seg_of = reinterpret_cast<unsigned*>(take(4 * N));
  w.bytes = off;
  return w;
}

}  // namespace

long long voxel_fusion_workspace_bytes(long long N) { return N > 0 ? carve(nullptr, N).bytes : 0; }

int voxel_fusion_entry(const float* pts, const float* feats, long long ld_feats, long long C, const float* conf, long long conf_stride, long long N,
                       float voxel_size, float* voxel_pts, float* voxel_feats, int* inverse, int* counts, long long* n_voxels, void* workspace,
                       long long workspace_bytes, cudaStream_t st) {
  V3A_REQUIRE(pts && feats && conf && voxel_pts && voxel_feats && n_voxels && workspace, VIST3A_ERR_INVALID, "voxel_fusion: null pointer");
  V3A_REQUIRE(N > 0 && N < (1ll << 31) - kSortTile, VIST3A_ERR_INVALID, "voxel_fusion: N must be in (0, 2^31) (got %lld)", N);
  V3A_REQUIRE(C > 0 && C + 3 <= 128 && ld_feats >= C && conf_stride >= 1, VIST3A_ERR_INVALID, "voxel_fusion: feature dim must be in [1, 125] (got %lld)", C);
  V3A_REQUIRE(voxel_size > 0.f, VIST3A_ERR_INVALID, "voxel_fusion: voxel_size must be positive");
  V3A_REQUIRE(((uintptr_t)workspace & 255) == 0, VIST3A_ERR_INVALID, "voxel_fusion: workspace must be 256-byte aligned");
  const VoxelWorkspace w = carve(workspace, N);
  V3A_REQUIRE(workspace_bytes >= w.bytes, VIST3A_ERR_INVALID, "voxel_fusion: workspace of %lld bytes needed, %lld given", w.bytes, workspace_bytes);
  int rc = check_arch();
  if (rc) return rc;
  const int sms = num_sms();
  long long launches = 0;
  V3A_REQUIRE(((uintptr_t)pts & 15) == 0, VIST3A_ERR_INVALID, "voxel_fusion: pts must be 16-byte aligned");
  voxel_init_kernel<<<1, 32, 0, st>>>(w.header);
  voxel_range_kernel<<<(unsigned)min((long long)sms * 8, (N / 4 + 255) / 256 + 1), 256, 0, st>>>(pts, N, voxel_size, w.header);
  voxel_bits_kernel<<<1, 32, 0, st>>>(w.header);
  voxel_key_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(pts, N, voxel_size, w.header, w.keys_a, w.vals_a);
  launches += 4;
  // passes >= ceil(bits / 8) return immediately (device-side count: no host round trip)
  V3A_CUDA_OK(radix_sort_enqueue(w.keys_a, w.keys_b, w.vals_a, w.vals_b, N, 8, &w.header->npasses, w.block_hist, w.digit_total, st));
  launches += 24;
  seg_count_kernel<<<w.nscan, 256, 0, st>>>(w.keys_a, w.keys_b, N, w.header, w.partial);
  seg_scan_partials_kernel<<<1, 1024, 0, st>>>(w.partial, w.nscan, w.header, n_voxels, w.seg_start, N);
  seg_assign_kernel<<<w.nscan, 256, 0, st>>>(w.keys_a, w.keys_b, w.vals_a, w.vals_b, N, w.header, w.partial, w.seg_start, w.seg_of, inverse);
  launches += 3;
  const unsigned rgrid = (unsigned)min((N + 255) / 256, (long long)sms * 32);  // a warp per 32 voxels, 8 warps per block
  const int nch = (int)((3 + C + 31) / 32);
#define V3A_REDUCE(NCH_)                                                                                                                   \
  voxel_reduce_kernel<NCH_><<<rgrid, 256, 0, st>>>(pts, feats, ld_feats, (int)C, conf, conf_stride, w.vals_a, w.vals_b, w.header, w.seg_start, \
                                                   w.seg_of, voxel_pts, voxel_feats, counts)
  if (nch == 1) V3A_REDUCE(1);
  else if (nch == 2) V3A_REDUCE(2);
  else if (nch == 3) V3A_REDUCE(3);
  else V3A_REDUCE(4);
#undef V3A_REDUCE
  launches += 1;
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(launches);
  return VIST3A_OK;
}

}  // namespace v3a
