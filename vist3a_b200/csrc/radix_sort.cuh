// Stable LSD radix sort of (64-bit key, 32-bit value) pairs, 8-bit digits, shared by the voxelised fusion and the Gaussian rasteriser.
// One pass = radix_hist_kernel (per-block digit histogram, digit-major table) -> radix_scan_kernel (one block per digit: exclusive scan
// over the sort blocks + digit totals) -> radix_scatter_kernel (warp-private digit counters + match.any ranking: stable; the tile is
// reordered in shared memory so that digit runs leave as contiguous stores).  Buffers
// ping-pong between (keys_a, vals_a) and (keys_b, vals_b) with the pass index; `*npasses` (device memory) says how many passes run --
// launches for later passes return at once, so the host can enqueue the maximum without knowing the key width.  After the sort the data
// sit in the a-buffers if *npasses is even, in the b-buffers otherwise.
#pragma once
#include "common.cuh"
#include "host_util.cuh"

namespace v3a {
namespace {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;                          // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;    // 4096 keys per block

// ------------------------------------------------------------------------------------------------
// radix sort pass p (digit = bits [8p, 8p+8)); in/out buffers alternate with the pass index
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const unsigned long long* __restrict__ keys_a, const unsigned long long* __restrict__ keys_b,
                                                                  long long N, int pass, const int* __restrict__ npasses, unsigned* __restrict__ block_hist,
                                                                  int nblocks) {
  if (pass >= *npasses) return;
  const unsigned long long* keys = (pass & 1) ? keys_b : keys_a;
  __shared__ unsigned cnt[256];
  cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kSortTile;
  const int shift = pass * 8;
#pragma unroll 4
  for (int k = 0; k < kSortItems; ++k) {
    const long long i = base + k * kSortThreads + threadIdx.x;
    if (i < N) atomicAdd(&cnt[(unsigned)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  block_hist[(long long)threadIdx.x * nblocks + blockIdx.x] = cnt[threadIdx.x];
}

// Block d scans row d of the digit-major table in place (exclusive, over the sort blocks) and records the digit total; the scatter
// kernel adds the exclusive scan of the 256 totals (recomputed per block in shared memory: 256 values).
__global__ void __launch_bounds__(256) radix_scan_kernel(unsigned* __restrict__ data, int nblocks, unsigned* __restrict__ digit_total, int pass,
                                                         const int* __restrict__ npasses) {
  if (pass >= *npasses) return;
  __shared__ unsigned warp_tot[8];
  __shared__ unsigned carry_s;
  unsigned* row = data + (long long)blockIdx.x * nblocks;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  constexpr int PER = 4;
  for (int base = 0; base < nblocks; base += 256 * PER) {
    unsigned v[PER];
    unsigned s = 0;
    const int i0 = base + (int)threadIdx.x * PER;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      v[k] = (i0 + k < nblocks) ? row[i0 + k] : 0u;
      s += v[k];
    }
    unsigned incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    unsigned woff = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) woff += (w < wid) ? warp_tot[w] : 0u;
    unsigned run = carry_s + woff + (incl - s);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      if (i0 + k < nblocks) row[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 255) carry_s = run;  // carry + chunk total
    __syncthreads();
  }
  if (threadIdx.x == 0) digit_total[blockIdx.x] = carry_s;
}

constexpr int kScatterSmem = kSortTile * 12 + (kSortThreads / 32) * 256 * 4 + 3 * 256 * 4 + 64;

// Stable scatter of one 4096-key tile.  Ranking: warp w owns a contiguous slice of the tile and walks it in rounds of 32 keys; match.any
// groups the lanes of a round by digit, the group's first lane bumps the warp-private digit counter, so a key's rank is (keys of that digit
// earlier in the slice) -- no atomics, order preserved.  The tile is then laid out in digit order in shared memory and written from there:
// neighbouring threads hold neighbouring keys of a digit run, i.e. the global stores of a run are contiguous (16 keys = 128 bytes on
// average) instead of 4096 scattered 8- and 4-byte stores.
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const unsigned long long* __restrict__ keys_a, unsigned long long* __restrict__ keys_b_,
                                                                     const unsigned* __restrict__ vals_a, unsigned* __restrict__ vals_b_, long long N, int pass,
                                                                     const int* __restrict__ npasses, const unsigned* __restrict__ block_off, int nblocks,
                                                                     const unsigned* __restrict__ digit_total) {
  if (pass >= *npasses) return;
  // pass parity selects the direction: even a -> b, odd b -> a
  const unsigned long long* kin = (pass & 1) ? keys_b_ : keys_a;
  unsigned long long* kout = (pass & 1) ? const_cast<unsigned long long*>(keys_a) : keys_b_;
  const unsigned* vin = (pass & 1) ? vals_b_ : vals_a;
  unsigned* vout = (pass & 1) ? const_cast<unsigned*>(vals_a) : vals_b_;
  constexpr int NW = kSortThreads / 32;
  extern __shared__ __align__(16) unsigned char radix_smem[];
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(radix_smem);          // [4096] tile in digit order
  unsigned* s_vals = reinterpret_cast<unsigned*>(radix_smem + kSortTile * 8);              // [4096]
  unsigned(*wcnt)[256] = reinterpret_cast<unsigned(*)[256]>(radix_smem + kSortTile * 12);   // per-warp digit counters, then tile-local bases
  unsigned* dbase = reinterpret_cast<unsigned*>(radix_smem + kSortTile * 12 + NW * 1024);   // exclusive scan of the digit totals
  unsigned* gdel = dbase + 256;                                                             // global position - tile-local position, per digit
  unsigned* lscan = gdel + 256;                                                             // scratch of the block scans
  unsigned* dwarp = lscan + 256;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  auto block_excl_scan = [&](unsigned v) {  // exclusive scan over the 256 threads (digits)
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    __syncthreads();   // dwarp free again
    if (lane == 31) dwarp[wid] = incl;
    __syncthreads();
    unsigned woff = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) woff += (w < wid) ? dwarp[w] : 0u;
    return woff + incl - v;
  };
  dbase[threadIdx.x] = block_excl_scan(digit_total[threadIdx.x]);
  for (int d = lane; d < 256; d += 32) wcnt[wid][d] = 0;
  __syncwarp();
  // warp w owns the contiguous slice [w * 512, (w + 1) * 512) of the block tile; 16 rounds of 32 keys in order
  const long long tile0 = (long long)blockIdx.x * kSortTile;
  const long long wbase = tile0 + (long long)wid * (kSortTile / NW);
  const int shift = pass * 8;
  unsigned long long key[kSortItems];
  unsigned rank[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const long long i = wbase + r * 32 + lane;
    const bool ok = i < N;
    key[r] = ok ? kin[i] : ~0ull;
    const unsigned d = (unsigned)(key[r] >> shift) & 255u;
    const unsigned act = __ballot_sync(0xffffffffu, ok);
    unsigned peers = __match_any_sync(0xffffffffu, ok ? d : 256u + (unsigned)lane) & act;
    const unsigned before = __popc(peers & ((1u << lane) - 1u));
    unsigned basev = 0;
    const int leader = __ffs(peers) - 1;
    if (ok && lane == leader) {
      basev = wcnt[wid][d];
      wcnt[wid][d] = basev + __popc(peers);
    }
    basev = __shfl_sync(0xffffffffu, basev, leader < 0 ? 0 : leader);
    rank[r] = basev + before;
    __syncwarp();
  }
  __syncthreads();
  // per digit (thread d): tile-local base of the digit, then of every warp inside it; and the shift to the global position
  {
    const int d = threadIdx.x;  // kSortThreads == 256 digits
    unsigned c[NW], tot = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      c[w] = wcnt[w][d];
      tot += c[w];
    }
    const unsigned lbase = block_excl_scan(tot);
    gdel[d] = dbase[d] + block_off[(long long)d * nblocks + blockIdx.x] - lbase;
    unsigned run = lbase;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      wcnt[w][d] = run;
      run += c[w];
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const long long i = wbase + r * 32 + lane;
    if (i < N) {
      const unsigned d = (unsigned)(key[r] >> shift) & 255u;
      const unsigned lp = wcnt[wid][d] + rank[r];
      s_keys[lp] = key[r];
      s_vals[lp] = vin[i];
    }
  }
  __syncthreads();
  const int count = (int)min((long long)kSortTile, N - tile0);
#pragma unroll 4
  for (int k = 0; k < kSortItems; ++k) {
    const int j = k * kSortThreads + threadIdx.x;
    if (j < count) {
      const unsigned long long kk = s_keys[j];
      const unsigned pos = (unsigned)j + gdel[(unsigned)(kk >> shift) & 255u];
      kout[pos] = kk;
      vout[pos] = s_vals[j];
    }
  }
}

// enqueue `max_passes` passes; block_hist holds 256 * nblocks counters, digit_total 256
inline cudaError_t radix_sort_enqueue(unsigned long long* keys_a, unsigned long long* keys_b, unsigned* vals_a, unsigned* vals_b, long long n, int max_passes,
                               const int* npasses, unsigned* block_hist, unsigned* digit_total, cudaStream_t st) {
  const int nblocks = (int)((n + kSortTile - 1) / kSortTile);
  static std::atomic<unsigned long long> attr_done{0};  // one copy of the kernel (and of this record) per translation unit; one bit per device
  if (ensure_dynamic_smem(radix_scatter_kernel, kScatterSmem, attr_done) != cudaSuccess) return cudaGetLastError();
  for (int pass = 0; pass < max_passes; ++pass) {
    radix_hist_kernel<<<nblocks, kSortThreads, 0, st>>>(keys_a, keys_b, n, pass, npasses, block_hist, nblocks);
    radix_scan_kernel<<<256, 256, 0, st>>>(block_hist, nblocks, digit_total, pass, npasses);
    radix_scatter_kernel<<<nblocks, kSortThreads, kScatterSmem, st>>>(keys_a, keys_b, vals_a, vals_b, n, pass, npasses, block_hist, nblocks, digit_total);
  }
  return cudaSuccess;
}

}  // namespace
}  // namespace v3a
