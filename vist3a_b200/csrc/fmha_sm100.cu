// tcgen05 fused attention forward (non-causal, no mask):  O = softmax(Q K^T * scale) V
//
// One CTA per (batch, head, 256-row query block) = TWO independent 128-row query tiles; each tile has its own MMA
// issuing warp and its own softmax warpgroup, so the tiles never wait for each other.
//   warp 0        TMA producer: Q tiles once, then K and V tiles of BKV keys through mbarrier rings
//   warp 1 / 3    MMA issuer of query tile 0 / 1:  S = Q K^T (SS) into TMEM,  O += P V (TS: P from TMEM, V MN-major smem).
//                 The whole warp runs the warp-uniform control flow, one elected lane issues: descriptors stay in
//                 uniform registers (no per-MMA R2UR / elect waterfall).
//   warp 2        TMEM allocator
//   warps 4..     softmax: SPLIT (1 or 2) warpgroups per query tile; a thread owns (a column half of) query row t (TMEM lane t):
//                 tcgen05.ld S, online max (halves exchange their partial max through smem + a 64-thread named barrier) with lazy
//                 rescale of O (only when the running max grows by more than 2^8), exponentials, bf16 P -> tcgen05.st
// MUFU.EX2 is the co-bottleneck of attention on this part (16 results/clk/SM: a 128x128 score tile costs as many SM
// cycles in exponentials as its two d=128 MMAs, twice as many at d=64).  Hence:
//   * the score tile of the next step is always produced while the warpgroup is still exponentiating the current one
//       d=128: 64 keys per step, S double-buffered, P(j) overwrites the buffer S(j) came from; the in-order tensor pipe
//              orders QK(j+2) behind P(j) V, so no "S free" handshake exists at all
//       d=64:  128 keys per step, S and P in separate columns; QK(j+1) is issued as soon as the softmax warps have pulled
//              S(j) into registers;
//   * kPoly of every 8 column pairs take their exp2 on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial,
//     rel. error 1e-4 << bf16 rounding of P) instead of the MUFU.
// TMEM map, 256 columns per query tile:  d=128: S0|P0 [0,64)  S1|P1 [64,128)  O [128,256)
//                                        d=64:  S [0,128)  P [128,192)  O [192,256)
#include "common.cuh"
#include "fmha_math.cuh"
#include "host_util.cuh"

namespace v3a {

struct FmhaParams {
  void* O;
  long long o_bs, o_rs, o_hs;
  int len_q, len_kv;
  float scale_log2;  // scale * log2(e)
  const float* row_scale;  // optional per-(batch, query row) positive factor on the logits
  int single_issuer;       // one MMA-issuing warp for both query tiles (flags bit 2)
  int direct_store;        // per-thread output stores instead of shared memory + bulk tensor store (flags bit 6, A/B)
  long long* trace;        // debug: 32 clock64 stamps / phase sums per CTA (v3a_debug_fmha_trace), normally null
  uint32_t zero;           // 0 (a value ptxas cannot fold: scheduling aid of the speculative softmax)
  int skip_softmax;        // debug (flags bit 12): the softmax warps only hand the barriers on (garbage output): tensor-side ceiling
  // Tail rows.  When len_q is a few rows more than a multiple of the CTA's 256 query rows (the decoder's frames: 1029 = 4 * 256 + 5 tokens),
  // a whole CTA would run every key step for them.  The tensor-core CTAs then cover len_q - tail_rows rows (len_q above is that number)
  // and the grid gets extra z-slices whose CTAs compute the tail rows of one (batch, head) each on the CUDA cores (fmha_tail_rows).
  int tail_rows, batch, heads;
  const __nv_bfloat16 *Qg, *Kg, *Vg;
  long long q_bs, q_rs, q_hs, k_bs, k_rs, k_hs, v_bs, v_rs, v_hs;
};

// debug hook (tools/fmha_trace.py), off unless armed: process-wide by design, read once per launch
static std::atomic<long long*> g_fmha_trace{nullptr};
extern "C" void v3a_debug_fmha_trace(void* buf) { g_fmha_trace.store(reinterpret_cast<long long*>(buf)); }
#define FMHA_TRACE_ADD(slot, val)                                                                                     \
  do {                                                                                                                \
    if (p.trace) p.trace[((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 32 + (slot)] = (val); \
  } while (0)
#define FMHA_TRACE(slot)                                                                                              \
  do {                                                                                                                \
    if (p.trace) p.trace[((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 32 + (slot)] = clock64(); \
  } while (0)

// The last tail_rows (<= 8) query rows of one (batch, head), head_dim 64, on the CUDA cores in fp32.  The CTA has one thread block's worth of
// latency hiding (the kernel's shared-memory request allows one CTA per SM), so nothing here may wait on global memory more than twice: the
// whole K of the head, then the whole V, come in by TMA through the kernel's own tensor maps (128-key boxes, 128-byte swizzle: a thread that
// walks "its" key row reads conflict-free), everything else runs from shared memory.
//   smem (1024-aligned): K | V [nbox][128 keys][128 B]  (later: per-warp partial outputs)  |  scores / P [tail_rows][Lp] fp32  |  q [8][64] fp32
//                        |  per-warp row max [8][32], row sum [8][32]  |  mbarrier
constexpr int kTailMax = 8;
static size_t fmha_tail_smem_bytes(int tail_rows, long long len_kv) {
  const long long nbox = (len_kv + 127) / 128, Lp = (len_kv + 3) & ~3ll;
  return (size_t)(1024 + nbox * 16384 + (tail_rows * Lp + kTailMax * 64 + 2 * kTailMax * 32) * 4 + 16);
}
// (T = tail_rows as a template parameter: every row loop unrolls without branches, so that the loads of a group are issued ahead of its FMAs --
//  with one CTA of 12-20 warps per SM nothing else hides their latency)
template <int THREADS, int T>
__device__ __noinline__ void fmha_tail_rows(const FmhaParams& p, const CUtensorMap* tmK, const CUtensorMap* tmV, uint8_t* smem_raw, int bh) {
  constexpr int D = 64, NW = THREADS / 32;
  const int b = bh / p.heads, h = bh - b * p.heads;
  const int L = p.len_kv, tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbox = (L + 127) >> 7, Lp = (L + 3) & ~3;
  const int row0 = p.len_q;   // first tail row
  const uint32_t kv = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* kvg = smem_raw + (kv - smem_u32(smem_raw));
  float* sc = reinterpret_cast<float*>(kvg + (size_t)nbox * 16384);   // [T][Lp]
  float* qs = sc + (size_t)T * Lp;                                      // [kTailMax][D], pre-multiplied by scale * log2(e) [* row scale]
  float* wmax = qs + kTailMax * D;                                      // [kTailMax][32]
  float* wsum = wmax + kTailMax * 32;                                   // [kTailMax][32]
  const uint32_t bar = kv + (uint32_t)nbox * 16384u + (uint32_t)((T * Lp + kTailMax * D + 2 * kTailMax * 32) * 4);
  if (tid == 0) {
    FMHA_TRACE(0);
    mbar_init(bar, 1);
    fence_barrier_init();
    mbar_expect_tx(bar, (uint32_t)nbox * 16384u);
    for (int x = 0; x < nbox; ++x) tma_load_4d(kv + (uint32_t)x * 16384u, tmK, bar, 0, h, x * 128, b);   // keys >= len_kv: zero-filled
  }
  for (int e = tid; e < kTailMax * D; e += THREADS) {
    const int r = e / D, d = e - r * D;
    float v = 0.0f;
    if (r < T) {
      const float c = p.row_scale ? p.scale_log2 * p.row_scale[(long long)b * (p.len_q + T) + row0 + r] : p.scale_log2;
      v = c * __bfloat162float(p.Qg[(long long)b * p.q_bs + (long long)(row0 + r) * p.q_rs + (long long)h * p.q_hs + d]);
    }
    qs[e] = v;
  }
  __syncthreads();   // q staged, barrier initialised
  mbar_wait(bar, 0);
  if (tid == 0) FMHA_TRACE(1);
  // ---- scores (log2 domain): a thread owns up to KPT keys at a time, so that every q chunk read from shared memory serves all of them ----
  constexpr int KPT = 3;
  for (int kb = 0; kb < L; kb += KPT * THREADS) {
    float acc[KPT][T];
    int key[KPT];
#pragma unroll
    for (int u = 0; u < KPT; ++u) {
      key[u] = kb + u * THREADS + tid;
#pragma unroll
      for (int r = 0; r < T; ++r) acc[u][r] = 0.0f;
    }
#pragma unroll
    for (int ch = 0; ch < D / 8; ++ch) {
      float kf[KPT][8];
#pragma unroll
      for (int u = 0; u < KPT; ++u) {
        const int k = min(key[u], nbox * 128 - 1);
        const uint4 w = *reinterpret_cast<const uint4*>(kvg + (size_t)k * 128 + ((ch ^ (k & 7)) << 4));
        const uint32_t kw[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          kf[u][2 * q] = __uint_as_float(kw[q] << 16);
          kf[u][2 * q + 1] = __uint_as_float(kw[q] & 0xffff0000u);
        }
      }
#pragma unroll
      for (int r = 0; r < T; ++r) {
        const float4 qa = *reinterpret_cast<const float4*>(qs + r * D + ch * 8), qb = *reinterpret_cast<const float4*>(qs + r * D + ch * 8 + 4);
#pragma unroll
        for (int u = 0; u < KPT; ++u) {
          acc[u][r] = fmaf(kf[u][0], qa.x, fmaf(kf[u][1], qa.y, fmaf(kf[u][2], qa.z, fmaf(kf[u][3], qa.w, acc[u][r]))));
          acc[u][r] = fmaf(kf[u][4], qb.x, fmaf(kf[u][5], qb.y, fmaf(kf[u][6], qb.z, fmaf(kf[u][7], qb.w, acc[u][r]))));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < KPT; ++u) {
      if (key[u] < L) {
#pragma unroll
        for (int r = 0; r < T; ++r) sc[r * Lp + key[u]] = acc[u][r];
      }
    }
  }
  __syncthreads();   // scores complete; every read of K is done
  fence_proxy_async_smem();
  if (tid == 0) {
    FMHA_TRACE(2);
    mbar_expect_tx(bar, (uint32_t)nbox * 16384u);
    for (int x = 0; x < nbox; ++x) tma_load_4d(kv + (uint32_t)x * 16384u, tmV, bar, 0, h, x * 128, b);   // V over K, under the softmax
  }
  // ---- softmax: every warp takes a share of every row; per-warp maxima / sums are combined through shared memory ----
  {
    float mx[T];
#pragma unroll
    for (int r = 0; r < T; ++r) mx[r] = -INFINITY;
    for (int k = tid; k < L; k += THREADS) {
#pragma unroll
      for (int r = 0; r < T; ++r) mx[r] = fmaxf(mx[r], sc[r * Lp + k]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
      for (int r = 0; r < T; ++r) mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], o));
    }
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < T; ++r) wmax[r * 32 + warp] = mx[r];
    }
    __syncthreads();
    float sum[T];
#pragma unroll
    for (int r = 0; r < T; ++r) {
      mx[r] = wmax[r * 32 + (lane < NW ? lane : 0)];
      sum[r] = 0.0f;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
      for (int r = 0; r < T; ++r) mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], o));
    }
    for (int k = tid; k < L; k += THREADS) {
#pragma unroll
      for (int r = 0; r < T; ++r) {
        const float e = ex2_approx(sc[r * Lp + k] - mx[r]);
        sc[r * Lp + k] = e;
        sum[r] += e;
      }
    }
    if (tid < Lp - L) {   // padding of the rows (read by the 4-key groups below)
#pragma unroll
      for (int r = 0; r < T; ++r) sc[r * Lp + L + tid] = 0.0f;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
      for (int r = 0; r < T; ++r) sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < T; ++r) wsum[r * 32 + warp] = sum[r];
    }
  }
  __syncthreads();   // P and the row sums are complete
  if (tid == 0) FMHA_TRACE(3);
  mbar_wait(bar, 1);
  if (tid == 0) FMHA_TRACE(4);
  // ---- P V: warp w owns every NW-th group of four keys, a thread two head-dim columns ----
  float a0[T], a1[T];
#pragma unroll
  for (int r = 0; r < T; ++r) a0[r] = a1[r] = 0.0f;
  for (int kg = warp; kg * 4 < L; kg += NW) {
    float v0[4], v1[4];
    float4 pr[T];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = kg * 4 + u;   // (< nbox * 128; keys >= len_kv hold zeros and meet P = 0)
      const uint32_t vv = *reinterpret_cast<const uint32_t*>(kvg + (size_t)k * 128 + (((lane >> 2) ^ (k & 7)) << 4) + ((lane & 3) << 2));
      v0[u] = __uint_as_float(vv << 16);
      v1[u] = __uint_as_float(vv & 0xffff0000u);
    }
#pragma unroll
    for (int r = 0; r < T; ++r) pr[r] = *reinterpret_cast<const float4*>(sc + r * Lp + kg * 4);
#pragma unroll
    for (int r = 0; r < T; ++r) {
      a0[r] = fmaf(pr[r].x, v0[0], fmaf(pr[r].y, v0[1], fmaf(pr[r].z, v0[2], fmaf(pr[r].w, v0[3], a0[r]))));
      a1[r] = fmaf(pr[r].x, v1[0], fmaf(pr[r].y, v1[1], fmaf(pr[r].z, v1[2], fmaf(pr[r].w, v1[3], a1[r]))));
    }
  }
  __syncthreads();   // V is no longer read: its place takes the per-warp partial outputs [NW][kTailMax][D]
  if (tid == 0) FMHA_TRACE(5);
  float* part = reinterpret_cast<float*>(kvg);
#pragma unroll
  for (int r = 0; r < T; ++r) {
    part[((size_t)warp * kTailMax + r) * D + 2 * lane] = a0[r];
    part[((size_t)warp * kTailMax + r) * D + 2 * lane + 1] = a1[r];
  }
  __syncthreads();
  for (int e = tid; e < T * D; e += THREADS) {
    const int r = e / D, d = e - r * D;
    float acc = 0.0f, sum = 0.0f;
    for (int w = 0; w < NW; ++w) {
      acc += part[((size_t)w * kTailMax + r) * D + d];
      sum += wsum[r * 32 + w];
    }
    static_cast<__nv_bfloat16*>(p.O)[(long long)b * p.o_bs + (long long)(row0 + r) * p.o_rs + (long long)h * p.o_hs + d] = __float2bfloat16(acc / sum);
  }
  if (tid == 0) FMHA_TRACE(6);
}

// QT_ = 1 (round 2): ONE 128-row query tile per CTA with two softmax threads per row.  Tensor memory then has room for double-buffered
// 128-key score tiles at d=128 (S0|P0 [0,128)  S1|P1 [128,256)  O [256,384)), so Q K^T of step j+1 runs under the softmax of step j without
// the 64-key steps whose SS-mode MMAs are shared-memory bound, and 384 threads leave 168+ registers to the softmax threads.
template <int D, int BKV_, int POLY_, int SPLIT_, int FAST_ = 0, int QT_ = 2>
struct FmhaCfg {
  static constexpr int BQ = 128, QT = QT_;           // query tiles per CTA
  static constexpr int BKV = BKV_;                   // keys per step
  static constexpr bool ALIAS = (D == 128);          // P(j) overwrites S(j)
  static constexpr int NSB = (ALIAS && (BKV == 64 || QT == 1)) ? 2 : 1;  // score buffers per query tile
  static constexpr int SLABS = D / 64;               // 64-element (128-byte) column slabs
  static constexpr int Q_SLAB_BYTES = BQ * 128;
  static constexpr int KV_SLAB_BYTES = BKV * 128;
  static constexpr int Q_TILE_BYTES = SLABS * Q_SLAB_BYTES;
  static constexpr int KV_TILE_BYTES = SLABS * KV_SLAB_BYTES;   // 16 KB either way
  static constexpr int KV_STAGES = KV_TILE_BYTES > 16384 ? 2 : 4;
  static constexpr int PT = NSB + 1 + 4 + 4;         // barriers per query tile: s_full[NSB], s_free, p_full[2], pv_done[2], p_half[2], pvh_done[2]
  static constexpr int NBARS = 1 + 4 * KV_STAGES + 2 * PT;
  static constexpr int SPLIT = SPLIT_;               // threads per query row (softmax warpgroups per tile)
  static constexpr int THREADS = 128 + 128 * QT * SPLIT;
  static constexpr int HC = BKV / SPLIT;             // score columns per softmax thread and step
  static constexpr int XCH_BYTES = 4 * 2 * 2 * 128 * 4;    // row-max / row-sum / slow-path flag exchange between the two threads of a row: [slot][tile][half][row]
  static constexpr int SMEM_BYTES = Q_TILE_BYTES * QT + KV_TILE_BYTES * 2 * KV_STAGES + 1024 + 8 * NBARS + 16 + XCH_BYTES;
  static_assert(SPLIT == 1 || SPLIT == 2, "SPLIT");
  static constexpr uint32_t TILE_COLS = QT == 1 ? 512 : 256;
  static constexpr uint32_t TM_S = 0;
  static constexpr uint32_t S_STRIDE = BKV;          // between the NSB score buffers (ALIAS only)
  static constexpr uint32_t TM_P = ALIAS ? 0 : 128;
  static constexpr uint32_t P_STRIDE = (NSB == 2) ? BKV : 0;
  static constexpr uint32_t TM_O = ALIAS ? NSB * BKV : 192;
  static_assert(QT == 1 || QT == 2, "QT");
  static_assert(QT == 2 || (SPLIT_ == 2 && BKV_ == 128), "one tile per CTA: two threads per row, 128-key steps");
  static_assert(D == 128 || BKV == 128, "d=64 runs 128-key steps");
  static constexpr int POLY = POLY_;                     // of every 8 column pairs, this many use the FMA-pipe exp2
  static constexpr bool FAST = FAST_ != 0;               // speculative (stale-maximum) softmax in 64-column half-steps, one thread per row
  // aliased 128-key steps (S single-buffered: softmax -> P V -> next Q K^T is a serial chain per tile): the first 64 keys of P(j) are handed
  // to the tensor pipe while the other 64 are still being exponentiated
  static constexpr bool HANDOFF = (FAST_ == 1 || FAST_ == 3) && ALIAS && BKV_ == 128;
  // FAST_ == 1: one thread per query row, two 64-column half-steps per 128 keys in sequence; FAST_ == 2: two threads per row, one 64-column half
  // each (128-key steps only), which agree on the rare slow path through shared memory and a 64-thread named barrier AFTER the exponentials
  static_assert(FAST_ == 0 || ((FAST_ == 1 || FAST_ == 3) && SPLIT_ == 1) || (FAST_ == 2 && SPLIT_ == 2 && BKV_ == 128), "speculative softmax variants");
  static_assert(TM_O + D <= TILE_COLS, "TMEM budget");
};


// registers per softmax thread of the two-threads-per-row speculative variant (96 = the launch-time allocation, no setmaxnreg).  setmaxnreg
// redistributes only what the CTA was given at launch (640 x 96): raising 16 warps to 112 needs 8192 registers, shrinking the 4 utility warps
// to 40 frees 7168 -- that dead-locks; 104 works but the utility warps then spill in their per-step loops: 630 vs 693 TFLOP/s at 13 377 keys.
#ifndef V3A_FAST2_REGS
#define V3A_FAST2_REGS 96
#endif
constexpr int kFast2Regs = V3A_FAST2_REGS;


template <int D, int BKV_, int POLY_, int SPLIT_, int FAST_ = 0, int QT_ = 2>
__global__ void __launch_bounds__(128 + 128 * QT_ * SPLIT_, 1)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const FmhaParams p) {
  using Cfg = FmhaCfg<D, BKV_, POLY_, SPLIT_, FAST_, QT_>;
  constexpr int SPLIT = Cfg::SPLIT;
  constexpr int HC = Cfg::HC;
  constexpr int ST = Cfg::KV_STAGES;
  constexpr int NSB = Cfg::NSB;
  constexpr int BKV = Cfg::BKV;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto smem_q = [&](int i) { return smem_base + Cfg::Q_TILE_BYTES * i; };
  auto smem_k = [&](int s) { return smem_base + Cfg::Q_TILE_BYTES * Cfg::QT + Cfg::KV_TILE_BYTES * s; };
  auto smem_v = [&](int s) { return smem_base + Cfg::Q_TILE_BYTES * Cfg::QT + Cfg::KV_TILE_BYTES * (ST + s); };
  const uint32_t bar_base = smem_base + Cfg::Q_TILE_BYTES * Cfg::QT + Cfg::KV_TILE_BYTES * 2 * ST;
  const uint32_t q_full = bar_base;
  auto k_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bar_base + 8u * (1 + ST + s); };
  auto v_full = [&](int s) { return bar_base + 8u * (1 + 2 * ST + s); };
  auto v_empty = [&](int s) { return bar_base + 8u * (1 + 3 * ST + s); };
  // Per query tile.  p_full / pv_done alternate between two barriers (step parity) so that a waiter is never two
  // completions behind the barrier it polls (mbarrier parity waits only distinguish the current and the previous phase).
  auto s_full = [&](int i, int b) { return bar_base + 8u * (1 + 4 * ST + i * Cfg::PT + b); };
  auto s_free = [&](int i) { return bar_base + 8u * (1 + 4 * ST + i * Cfg::PT + NSB); };
  auto p_full = [&](int i, int b) { return bar_base + 8u * (1 + 4 * ST + i * Cfg::PT + NSB + 1 + b); };
  auto pv_done = [&](int i, int b) { return bar_base + 8u * (1 + 4 * ST + i * Cfg::PT + NSB + 3 + b); };
  auto p_half = [&](int i, int b) { return bar_base + 8u * (1 + 4 * ST + i * Cfg::PT + NSB + 5 + b); };
  auto pvh_done = [&](int i, int b) { return bar_base + 8u * (1 + 4 * ST + i * Cfg::PT + NSB + 7 + b); };   // first half of P(j) V(j) has completed
  const uint32_t tmem_slot = bar_base + 8u * Cfg::NBARS;
  const uint32_t xch_base = tmem_slot + 16u;  // float [2 parities][2 tiles][2 halves][128 rows]

  if ((int)blockIdx.z >= p.batch) {
    // (extra z-slices of the grid: the tail rows of one (batch, head) per CTA; they are scheduled last and fill the grid's last, partial wave)
    const int bh = (int)(((blockIdx.z - p.batch) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
    pdl_launch_dependents();
    pdl_wait();
    if constexpr (D == 64) {
      if (bh < p.batch * p.heads) {
        switch (p.tail_rows) {
          case 1: fmha_tail_rows<Cfg::THREADS, 1>(p, &tmK, &tmV, smem_raw, bh); break;
          case 2: fmha_tail_rows<Cfg::THREADS, 2>(p, &tmK, &tmV, smem_raw, bh); break;
          case 3: fmha_tail_rows<Cfg::THREADS, 3>(p, &tmK, &tmV, smem_raw, bh); break;
          case 4: fmha_tail_rows<Cfg::THREADS, 4>(p, &tmK, &tmV, smem_raw, bh); break;
          case 5: fmha_tail_rows<Cfg::THREADS, 5>(p, &tmK, &tmV, smem_raw, bh); break;
          case 6: fmha_tail_rows<Cfg::THREADS, 6>(p, &tmK, &tmV, smem_raw, bh); break;
          case 7: fmha_tail_rows<Cfg::THREADS, 7>(p, &tmK, &tmV, smem_raw, bh); break;
          default: fmha_tail_rows<Cfg::THREADS, 8>(p, &tmK, &tmV, smem_raw, bh); break;
        }
      }
    }
    return;
  }
  const uint32_t warp = warp_id_sync();
  const uint32_t lane = lane_id();
  const int q0 = blockIdx.x * (Cfg::BQ * Cfg::QT);
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_kv = (p.len_kv + BKV - 1) / BKV;
  const int nq = (Cfg::QT == 2 && q0 + Cfg::BQ < p.len_q) ? 2 : 1;  // query tiles of this CTA that hold at least one row
  if (threadIdx.x == 0) {
    FMHA_TRACE(0);
    if (p.trace) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.trace[((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 32 + 12] = smid;
    }
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), (uint32_t)nq);  // one commit per MMA warp
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), (uint32_t)nq);
    }
    for (int i = 0; i < 2; ++i) {
      for (int b = 0; b < NSB; ++b) mbar_init(s_full(i, b), 1);
      mbar_init(s_free(i), 4 * SPLIT);  // one arrival per softmax warp of the tile
      for (int b = 0; b < 2; ++b) {
        mbar_init(p_full(i, b), 4 * SPLIT);
        mbar_init(pv_done(i, b), 1);
        mbar_init(p_half(i, b), 4 * SPLIT);
        mbar_init(pvh_done(i, b), 1);
      }
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (threadIdx.x == 0) FMHA_TRACE(1);
  pdl_launch_dependents();
  pdl_wait();  // PDL: q/k/v of the previous kernel are read (and O written) only after this point
  if (threadIdx.x == 0) FMHA_TRACE(2);

  if (warp < 4) {
    if constexpr (Cfg::THREADS == 384) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if constexpr (FAST_ == 2 && Cfg::THREADS == 640 && kFast2Regs != 96) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    // (SPLIT == 2: 20 warps at the launch-time 96 registers; see the note at the softmax branch)
    if (warp == 0) {
      // ------------------------------ TMA producer (warp-uniform control flow, one elected lane issues) ---------
      if (elect_one()) {
        mbar_expect_tx(q_full, (uint32_t)(nq * Cfg::Q_TILE_BYTES));
        for (int i = 0; i < nq; ++i) {
#pragma unroll
          for (int sl = 0; sl < Cfg::SLABS; ++sl)
            tma_load_4d(smem_q(i) + sl * Cfg::Q_SLAB_BYTES, &tmQ, q_full, sl * 64, head, q0 + i * Cfg::BQ, batch);
        }
      }
      __syncwarp();
      if (lane == 0) FMHA_TRACE(3);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(k_empty(s), ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(k_full(s), Cfg::KV_TILE_BYTES);
#pragma unroll
          for (int sl = 0; sl < Cfg::SLABS; ++sl)
            tma_load_4d(smem_k(s) + sl * Cfg::KV_SLAB_BYTES, &tmK, k_full(s), sl * 64, head, j * BKV, batch);
        }
        __syncwarp();
        mbar_wait(v_empty(s), ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(v_full(s), Cfg::KV_TILE_BYTES);
#pragma unroll
          for (int sl = 0; sl < Cfg::SLABS; ++sl)
            tma_load_4d(smem_v(s) + sl * Cfg::KV_SLAB_BYTES, &tmV, v_full(s), sl * 64, head, j * BKV, batch);
        }
        __syncwarp();
        if (++s == ST) { s = 0; ph ^= 1u; }
      }
    } else if (warp != 2) {
      // ------------------------------ MMA issue ------------------------------
      constexpr uint32_t idesc_qk = make_idesc(kFmtBF16, 128, BKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(kFmtBF16, 128, D, 0, 1);  // B (V) is MN-major
      auto issue_qk = [&](int i, int j, int s, uint32_t ph) {  // (s, ph) = ring slot / phase of key tile j
        mbar_wait(k_full(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + (uint32_t)i * Cfg::TILE_COLS + Cfg::TM_S + (uint32_t)(j % NSB) * Cfg::S_STRIDE;
          const uint64_t qdesc = make_smem_desc_sw128(smem_q(i), 1024, 0);
          const uint64_t bdesc = make_smem_desc_sw128(smem_k(s), 1024, 0);
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            // +32 B along K inside a swizzle atom = +2 in the (addr >> 4) field; next 64-wide slab = + SLAB_BYTES
            const uint64_t ao = (uint64_t)(((kk >> 2) * Cfg::Q_SLAB_BYTES + (kk & 3) * 32) >> 4);
            const uint64_t bo = (uint64_t)(((kk >> 2) * Cfg::KV_SLAB_BYTES + (kk & 3) * 32) >> 4);
            umma_f16_ss<1>(d_tmem, qdesc + ao, bdesc + bo, idesc_qk, kk ? 1u : 0u);
          }
          umma_commit(s_full(i, j % NSB));
          umma_commit(k_empty(s));
        }
        __syncwarp();
      };
      // part: -1 = the whole P(j) V(j); 0 / 1 = its first / second 64 keys (HANDOFF: the halves of P arrive separately)
      auto issue_pv = [&](int i, int j, int s, uint32_t ph, int part = -1) {
        if (part <= 0) mbar_wait(v_full(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t t_base = tmem_base + (uint32_t)i * Cfg::TILE_COLS;
          const uint32_t p_tmem = t_base + Cfg::TM_P + (uint32_t)(j & 1) * Cfg::P_STRIDE;
          // V tile: kv rows at a 128 B pitch (K dimension), 64-wide head-dim slabs LBO apart (MN dimension)
          const uint64_t bdesc = make_smem_desc_sw128(smem_v(s), 1024, Cfg::KV_SLAB_BYTES);
          constexpr int KS = BKV / 16;
          const int k_lo = part == 1 ? KS / 2 : 0, k_hi = part == 0 ? KS / 2 : KS;
#pragma unroll
          for (int kk = 0; kk < KS; ++kk) {
            if (kk >= k_lo && kk < k_hi)
              umma_f16_ts(t_base + Cfg::TM_O, p_tmem + (uint32_t)(kk * 8), bdesc + (uint64_t)((kk * 2048) >> 4), idesc_pv, (j | kk) ? 1u : 0u);
          }
          if (part != 0) {
            umma_commit(pv_done(i, j & 1));
            umma_commit(v_empty(s));
          } else {
            umma_commit(pvh_done(i, j & 1));
          }
        }
        __syncwarp();
      };
      auto adv = [&](int& s_, uint32_t& ph_) { if (++s_ == ST) { s_ = 0; ph_ ^= 1u; } };
      const bool trm = p.trace != nullptr && warp == 1;
      long long mm_wait = 0, mm_issue = 0, mt = 0;
      if (p.single_issuer) {
        // ONE warp serves both query tiles, polling their barriers and issuing whichever block (P V of step j, then Q K^T of step
        // j + NSB) is ready as a whole: the in-order tensor pipe then finishes a tile's block in its own 512 cycles instead of
        // interleaving it MMA by MMA with the other tile's (two issuing warps), which delayed BOTH score tiles by a full 1024.
        if (warp == 1) {
          mbar_wait(q_full, 0);
          if (lane == 0) FMHA_TRACE(4);
          int qs[2] = {0, 0}, vs[2] = {0, 0}, jq[2] = {0, 0}, jp[2] = {0, 0};
          uint32_t qph[2] = {0, 0}, vph[2] = {0, 0};
          for (int j = 0; j < NSB && j < n_kv; ++j) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              if (i < nq) {
                issue_qk(i, j, qs[i], qph[i]);
                adv(qs[i], qph[i]);
                jq[i] = j + 1;
              }
            }
            if (j == 0 && lane == 0) FMHA_TRACE(5);
          }
          int remaining = nq * n_kv;
          if (trm) mt = clock64();
          while (remaining > 0) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              if (i >= nq) continue;
              if (!Cfg::ALIAS && jq[i] < n_kv) {  // Q K^T of the next step as soon as the softmax warps hold S in registers
                if (__any_sync(0xffffffffu, mbar_try_wait(s_free(i), (uint32_t)(jq[i] - 1) & 1u))) {
                  tc_fence_after();
                  if (trm) { const long long t2 = clock64(); mm_wait += t2 - mt; mt = t2; }
                  issue_qk(i, jq[i], qs[i], qph[i]);
                  adv(qs[i], qph[i]);
                  ++jq[i];
                  if (trm) { const long long t2 = clock64(); mm_issue += t2 - mt; mt = t2; }
                }
              }
              if (jp[i] < n_kv) {
                const int j = jp[i];
                if (__any_sync(0xffffffffu, mbar_try_wait(p_full(i, j & 1), (uint32_t)(j >> 1) & 1u))) {
                  tc_fence_after();
                  if (trm) { const long long t2 = clock64(); mm_wait += t2 - mt; mt = t2; }
                  issue_pv(i, j, vs[i], vph[i]);
                  adv(vs[i], vph[i]);
                  if (Cfg::ALIAS && j + NSB < n_kv) {  // overwrites S(j)|P(j): ordered behind P(j) V by the in-order tensor pipe
                    issue_qk(i, j + NSB, qs[i], qph[i]);
                    adv(qs[i], qph[i]);
                  }
                  ++jp[i];
                  --remaining;
                  if (trm) { const long long t2 = clock64(); mm_issue += t2 - mt; mt = t2; }
                }
              }
            }
          }
          if (trm && lane == 0) { FMHA_TRACE_ADD(21, mm_wait); FMHA_TRACE_ADD(22, mm_issue); }
        }
      } else if ((int)(warp >> 1) < nq) {
        // one issuing warp per query tile (warp 1 -> tile 0, warp 3 -> tile 1)
        const int i = (int)(warp >> 1);
        mbar_wait(q_full, 0);
        if (warp == 1 && lane == 0) FMHA_TRACE(4);
        int qs = 0, vs = 0;  // ring positions of the key tile the next QK / PV reads
        uint32_t qph = 0, vph = 0;
        for (int j = 0; j < NSB && j < n_kv; ++j) {
          issue_qk(i, j, qs, qph);
          adv(qs, qph);
          if (j == 0 && warp == 1 && lane == 0) FMHA_TRACE(5);
        }
        for (int j = 0; j < n_kv; ++j) {
          const bool more = j + NSB < n_kv;
          if (trm) mt = clock64();
          if (!Cfg::ALIAS && more) {
            mbar_wait(s_free(i), (uint32_t)j & 1u);  // the softmax warps hold S(j) in registers
            tc_fence_after();
            issue_qk(i, j + 1, qs, qph);
            adv(qs, qph);
          }
          if constexpr (Cfg::HANDOFF) {
            mbar_wait(p_half(i, j & 1), (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            if (trm) { const long long t2 = clock64(); mm_wait += t2 - mt; mt = t2; }
            issue_pv(i, j, vs, vph, 0);
            if (trm) { const long long t2 = clock64(); mm_issue += t2 - mt; mt = t2; }
          }
          mbar_wait(p_full(i, j & 1), (uint32_t)(j >> 1) & 1u);
          tc_fence_after();
          if (trm) { const long long t2 = clock64(); mm_wait += t2 - mt; mt = t2; }
          issue_pv(i, j, vs, vph, Cfg::HANDOFF ? 1 : -1);
          adv(vs, vph);
          if (Cfg::ALIAS && more) {  // overwrites S(j)|P(j): ordered behind P(j) V by the in-order tensor pipe
            issue_qk(i, j + NSB, qs, qph);
            adv(qs, qph);
          }
          if (trm) mm_issue += clock64() - mt;
        }
        if (trm && lane == 0) { FMHA_TRACE_ADD(21, mm_wait); FMHA_TRACE_ADD(22, mm_issue); }
      }
    }
  } else {
    if constexpr (Cfg::THREADS == 384) asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    if constexpr (FAST_ == 2 && Cfg::THREADS == 640 && kFast2Regs != 96) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kFast2Regs));
    // SPLIT == 2 keeps the launch-time allocation (96 registers x 640 threads): setmaxnreg.inc above 96 deadlocked on B200 --
    // the per-warp register allocation is coarser than the PTX granularity of 8, and 16 warps x 4096 registers is the whole file.
    // ------------------------------ softmax / correction / epilogue ------------------------------
    // SPLIT threads share a query row: thread (tile i, half h, row) owns score columns [h*HC, (h+1)*HC) of every step, half of the
    // O columns for the (rare) rescale and the output store; the two halves exchange their partial row max through shared memory
    // and a 64-thread named barrier per step, so both take identical rescale decisions.
    const int sw = (int)warp - 4;
    const int i = sw / (4 * SPLIT);          // query tile
    const int h = (sw >> 2) % SPLIT;         // column half
    if (i < nq) {
      const uint32_t wq = warp & 3u;         // TMEM lane quadrant this warp may access
      const int rit = (int)(wq * 32u + lane);  // row in tile
      const uint32_t tile_base = tmem_base + ((wq * 32u) << 16) + (uint32_t)i * Cfg::TILE_COLS;
      const int row = q0 + i * Cfg::BQ + rit;
      constexpr int OC = D / SPLIT;          // O columns per thread
      const uint32_t o_addr = tile_base + Cfg::TM_O + (uint32_t)(h * OC);
      auto xch = [&](int par, int half) { return xch_base + 4u * (uint32_t)(((par * 2 + i) * 2 + half) * 128 + rit); };
      const uint32_t pair_bar = 1u + (uint32_t)i * 4u + wq;  // named barrier shared by the two warps that own these 32 rows
      float m_run = -INFINITY;  // running (possibly stale) row max of raw scores
      float l_run = 0.0f;       // running sum of exp2((s - m_run) * scale_log2) over this thread's columns
      const float c = (p.row_scale && row < p.len_q) ? p.scale_log2 * p.row_scale[(long long)batch * (p.len_q + p.tail_rows) + row] : p.scale_log2;
      const uint64_t cc2 = pack2(c, c);
      const bool tr = p.trace != nullptr && warp == 4;
      long long ph_wait = 0, ph_ld = 0, ph_max = 0, ph_exp = 0, ph_st = 0, tt = 0;
      if constexpr (FAST_ == 2) {
        // ---- speculative softmax, two threads per row: thread (row, h) owns score columns [64 h, 64 h + 64) of every 128-key step.  The
        //      exponentials run against the stale maximum with no exchange in front of them; afterwards the two threads of a row exchange ONE
        //      flag (row sum of the half above 2^8 or not finite) and only then store P, so that a flagged step can still reload its scores from
        //      tensor memory, agree on the row maximum, rescale O / the row sums once and redo the exponentials exactly. ----
        float nmc = 0.0f;
        uint64_t mc2 = 0ull;
        auto max64 = [&](const uint32_t* r) {
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int k = 0; k < 64; k += 8) {
#pragma unroll
            for (int u = 0; u < 4; ++u) mx[u] = fmaxf(fmaxf(mx[u], __uint_as_float(r[k + 2 * u])), __uint_as_float(r[k + 2 * u + 1]));
          }
          return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        };
        auto exchange = [&](int slot, float mine) {   // value of the other thread of this row (both call it at the same point)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch(slot, h)), "f"(mine) : "memory");
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          float other;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(xch(slot, h ^ 1)) : "memory");
          return other;
        };
        for (int j = 0; j < n_kv; ++j) {
          const int sb = j % NSB;
          if (tr) tt = clock64();
          mbar_wait(s_full(i, sb), (uint32_t)(j / NSB) & 1u);
          tc_fence_after();
          if (tr) { const long long t2 = clock64(); ph_wait += t2 - tt; tt = t2; }
          if (j == 0 && warp == 4 && lane == 0) FMHA_TRACE(6);
          const uint32_t s_addr = tile_base + Cfg::TM_S + (uint32_t)sb * Cfg::S_STRIDE + (uint32_t)(h * 64);
          const uint32_t p_addr = tile_base + Cfg::TM_P + (uint32_t)(j & 1) * Cfg::P_STRIDE + (uint32_t)(h * 32);
          const int valid = p.len_kv - j * BKV - h * 64;   // columns of this thread that hold existing keys
          uint32_t pk[32];
          uint64_t hsum[2] = {0ull, 0ull};
          {
            uint32_t r[64];
            tmem_ld_x32(s_addr, r);
            tmem_ld_x32(s_addr + 32u, r + 32);
            tmem_ld_wait();
            if (tr) { const long long t2 = clock64(); ph_ld += t2 - tt; tt = t2; }
            if (valid < 64) {
#pragma unroll
              for (int k = 0; k < 64; ++k)
                if (k >= valid) r[k] = 0xff800000u;  // -inf
            }
            if (j == 0) {   // first scores of the row: the reference maximum is the maximum over both halves
              const float mine = max64(r);
              m_run = fmaxf(mine, exchange(2, mine));
              nmc = -m_run * c;
              mc2 = pack2(nmc, nmc);
            }
            exp_half64<Cfg::POLY, false>(r, cc2, mc2, hsum, pk, p.zero);
          }
          float hs;
          {
            float s0, s1, s2, s3;
            unpack2(hsum[0], s0, s1);
            unpack2(hsum[1], s2, s3);
            hs = (s0 + s1) + (s2 + s3);
          }
          if (tr) { const long long t2 = clock64(); ph_exp += t2 - tt; tt = t2; }
          const float flag = (hs <= 256.0f) ? 0.0f : 1.0f;
          const float flag_other = exchange(j & 1, flag);
          if (__any_sync(0xffffffffu, (flag + flag_other) != 0.0f)) {
            // (rare) exact path for the rows of this warp pair: the scores are still in tensor memory (P is stored below, and Q K^T of the next
            // step waits for s_free / for P V of this step)
            uint32_t r[64];
            tmem_ld_x32(s_addr, r);
            tmem_ld_x32(s_addr + 32u, r + 32);
            tmem_ld_wait();
            if (valid < 64) {
#pragma unroll
              for (int k = 0; k < 64; ++k)
                if (k >= valid) r[k] = 0xff800000u;  // -inf
            }
            const float mine = max64(r);
            const float m_new = fmaxf(m_run, fmaxf(mine, exchange(3, mine)));
            const bool need = (m_new - m_run) * c > 8.0f;
            if (__any_sync(0xffffffffu, need)) {
              const float f = need ? ex2_approx((m_run - m_new) * c) : 1.0f;
              if (need) m_run = m_new;
              l_run *= f;
              if (j >= 1) {   // O holds P(0..j-1) V: P(j-1) V must have finished before the rows are rescaled (each thread its column half)
                mbar_wait(pv_done(i, (j - 1) & 1), (uint32_t)((j - 1) >> 1) & 1u);
                tc_fence_after();
#pragma unroll 1
                for (int cb = 0; cb < OC / 16; ++cb) {
                  uint32_t o[16];
                  tmem_ld_x16(o_addr + (uint32_t)(cb * 16), o);
                  tmem_ld_wait();
#pragma unroll
                  for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * f);
                  tmem_st_x16(o_addr + (uint32_t)(cb * 16), o);
                }
                tmem_st_wait();
              }
              nmc = -m_run * c;
              mc2 = pack2(nmc, nmc);
            }
            // (always redone here, also when no row needed the rescale: the fast path's P need not stay live across this branch)
            hsum[0] = hsum[1] = 0ull;
            exp_half64_exact(r, cc2, mc2, hsum, pk);
            float s0, s1, s2, s3;
            unpack2(hsum[0], s0, s1);
            unpack2(hsum[1], s2, s3);
            hs = (s0 + s1) + (s2 + s3);
          }
          l_run += hs;
          if (!Cfg::ALIAS) {
            // every thread of the row is done with S(j): Q K^T of the next step may overwrite it.  The single P buffer is read by P(j-1) V.
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free(i));
            if (j >= 1) {
              mbar_wait(pv_done(i, (j - 1) & 1), (uint32_t)((j - 1) >> 1) & 1u);
              tc_fence_after();
            }
          }
          tmem_st_x32(p_addr, pk);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full(i, j & 1));
          if (tr) { const long long t2 = clock64(); ph_st += t2 - tt; tt = t2; }
          if (j == 0 && warp == 4 && lane == 0) FMHA_TRACE(7);
        }
      } else if constexpr (Cfg::FAST) {
        // ---- speculative softmax: 64-column half-steps against the stale running maximum (fmha_math.cuh: exp_half64) ----
        constexpr int NHALF = BKV / 64;
        float nmc = 0.0f;
        uint64_t mc2 = 0ull;
        for (int j = 0; j < n_kv; ++j) {
          const int sb = j % NSB;
          if (tr) tt = clock64();
          mbar_wait(s_full(i, sb), (uint32_t)(j / NSB) & 1u);
          tc_fence_after();
          if (tr) { const long long t2 = clock64(); ph_wait += t2 - tt; tt = t2; }
          if (j == 0 && warp == 4 && lane == 0) FMHA_TRACE(6);
          const uint32_t p_addr = tile_base + Cfg::TM_P + (uint32_t)(j & 1) * Cfg::P_STRIDE;
          if (p.skip_softmax) {   // debug: no score loads, no exponentials -- the barriers alone
            if (!Cfg::ALIAS) { __syncwarp(); if (lane == 0) mbar_arrive(s_free(i)); }
            if (!Cfg::ALIAS && j >= 1) mbar_wait(pv_done(i, (j - 1) & 1), (uint32_t)((j - 1) >> 1) & 1u);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (Cfg::HANDOFF) mbar_arrive(p_half(i, j & 1));
              mbar_arrive(p_full(i, j & 1));
            }
            l_run = 1.0f;
            continue;
          }
          // d=64 (128-key steps, S not aliased): both halves of the score row are pulled into registers at once -- one TMEM round trip per step
          // instead of two, and Q K^T of the next step may overwrite S right away
          constexpr bool kPreload = !Cfg::ALIAS && NHALF == 2 && FAST_ == 3;   // A/B variant (flags bit 15): measured slower than two loads per step
          uint32_t rr[kPreload ? NHALF : 1][64];
          if constexpr (kPreload) {
#pragma unroll
            for (int hh = 0; hh < NHALF; ++hh) {
              const uint32_t s_addr = tile_base + Cfg::TM_S + (uint32_t)sb * Cfg::S_STRIDE + (uint32_t)(hh * 64);
              tmem_ld_x32(s_addr, rr[hh]);
              tmem_ld_x32(s_addr + 32u, rr[hh] + 32);
            }
            tmem_ld_wait();
            if (tr) { const long long t2 = clock64(); ph_ld += t2 - tt; tt = t2; }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free(i));
          }
#pragma unroll
          for (int hh = 0; hh < NHALF; ++hh) {
            uint32_t (&r)[64] = rr[kPreload ? hh : 0];
            if constexpr (!kPreload) {
              const uint32_t s_addr = tile_base + Cfg::TM_S + (uint32_t)sb * Cfg::S_STRIDE + (uint32_t)(hh * 64);
              tmem_ld_x32(s_addr, r);
              tmem_ld_x32(s_addr + 32u, r + 32);
              tmem_ld_wait();
              if (tr) { const long long t2 = clock64(); ph_ld += t2 - tt; tt = t2; }
              if (!Cfg::ALIAS && hh == NHALF - 1) {   // the whole score tile is in registers: Q K^T of the next step may overwrite it
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(s_free(i));
              }
            }
            const int valid = p.len_kv - j * BKV - hh * 64;   // columns of this half that hold existing keys
            if (valid < 64) {
#pragma unroll
              for (int k = 0; k < 64; ++k)
                if (k >= valid) r[k] = 0xff800000u;  // -inf
            }
            if (j == 0 && hh == 0) {   // first scores of the row: the reference maximum is their maximum
              float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
              for (int k = 0; k < 64; k += 8) {
#pragma unroll
                for (int u = 0; u < 4; ++u) mx[u] = fmaxf(fmaxf(mx[u], __uint_as_float(r[k + 2 * u])), __uint_as_float(r[k + 2 * u + 1]));
              }
              m_run = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
              nmc = -m_run * c;
              mc2 = pack2(nmc, nmc);
            }
            uint32_t pk[32];
            uint64_t hsum[2] = {0ull, 0ull};
            const float m_half = exp_half64<Cfg::POLY>(r, cc2, mc2, hsum, pk, p.zero);
            float hs;
            {
              float s0, s1, s2, s3;
              unpack2(hsum[0], s0, s1);
              unpack2(hsum[1], s2, s3);
              hs = (s0 + s1) + (s2 + s3);
            }
            {
              const bool need = (m_half - m_run) * c > 8.0f;
              if (__any_sync(0xffffffffu, need)) {
                // (rare) the stale maximum is too small for some row of this warp: rescale what has been accumulated, redo the half-step
                const float f = need ? ex2_approx((m_run - m_half) * c) : 1.0f;
                if (need) m_run = m_half;
                l_run *= f;
                if (j >= 1 || (Cfg::HANDOFF && hh > 0)) {   // O holds P(0..j-1) V: P(j-1) V must have finished before the rows are rescaled
                  if (j >= 1) mbar_wait(pv_done(i, (j - 1) & 1), (uint32_t)((j - 1) >> 1) & 1u);
                  // (the first half of P(j) has been handed over: its product is being added to O -- wait for it; the second half is not
                  //  issued before this warp arrives on p_full, so O is quiescent afterwards)
                  if (Cfg::HANDOFF && hh > 0) mbar_wait(pvh_done(i, j & 1), (uint32_t)(j >> 1) & 1u);
                  tc_fence_after();
#pragma unroll 1
                  for (int cb = 0; cb < D / 16; ++cb) {
                    uint32_t o[16];
                    tmem_ld_x16(o_addr + (uint32_t)(cb * 16), o);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * f);
                    tmem_st_x16(o_addr + (uint32_t)(cb * 16), o);
                  }
                }
                if (hh > 0 && !Cfg::HANDOFF) {   // the first half of P(j) is in tensor memory, taken against the old maximum
                  tmem_st_wait();
#pragma unroll 1
                  for (int cb = 0; cb < 2; ++cb) {
                    uint32_t q[16];
                    tmem_ld_x16(p_addr + (uint32_t)(cb * 16), q);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 16; ++k) q[k] = pack_bf16(bf16_lo(q[k]) * f, bf16_hi(q[k]) * f);
                    tmem_st_x16(p_addr + (uint32_t)(cb * 16), q);
                  }
                }
                tmem_st_wait();
                nmc = -m_run * c;
                mc2 = pack2(nmc, nmc);
                hsum[0] = hsum[1] = 0ull;
                exp_half64_exact(r, cc2, mc2, hsum, pk);
                float s0, s1, s2, s3;
                unpack2(hsum[0], s0, s1);
                unpack2(hsum[1], s2, s3);
                hs = (s0 + s1) + (s2 + s3);
              }
            }
            l_run += hs;
            if (tr) { const long long t2 = clock64(); ph_exp += t2 - tt; tt = t2; }
            // d=64: the single P buffer is read by P(j-1) V.  (d=128: the commit behind s_full of this step covers the last reader of
            // the buffer S(j) | P(j) lives in.)
            if (!Cfg::ALIAS && hh == 0 && j >= 1) {
              mbar_wait(pv_done(i, (j - 1) & 1), (uint32_t)((j - 1) >> 1) & 1u);
              tc_fence_after();
            }
            tmem_st_x32(p_addr + (uint32_t)(hh * 32), pk);
            if (Cfg::HANDOFF && hh == 0) {
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(p_half(i, j & 1));
            }
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full(i, j & 1));
          if (tr) { const long long t2 = clock64(); ph_st += t2 - tt; tt = t2; }
          if (j == 0 && warp == 4 && lane == 0) FMHA_TRACE(7);
        }
      } else
      for (int j = 0; j < n_kv; ++j) {
        const int sb = j % NSB;
        if (tr) tt = clock64();
        mbar_wait(s_full(i, sb), (uint32_t)(j / NSB) & 1u);
        tc_fence_after();
        if (tr) { const long long t2 = clock64(); ph_wait += t2 - tt; tt = t2; }
        if (j == 0 && warp == 4 && lane == 0) FMHA_TRACE(6);
        uint32_t r[HC];
        const uint32_t s_addr = tile_base + Cfg::TM_S + (uint32_t)sb * Cfg::S_STRIDE + (uint32_t)(h * HC);
#pragma unroll
        for (int cb = 0; cb < HC / 32; ++cb) tmem_ld_x32(s_addr + (uint32_t)(cb * 32), r + cb * 32);
        tmem_ld_wait();
        if (tr) { const long long t2 = clock64(); ph_ld += t2 - tt; tt = t2; }
        if (!Cfg::ALIAS) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free(i));
        }
        const int valid = p.len_kv - j * BKV - h * HC;  // columns of this thread that hold existing keys
        if (valid < HC) {
#pragma unroll
          for (int k = 0; k < HC; ++k)
            if (k >= valid) r[k] = 0xff800000u;  // -inf
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int k = 0; k < HC; k += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            mx[u] = fmaxf(fmaxf(mx[u], __uint_as_float(r[k + 2 * u])), __uint_as_float(r[k + 2 * u + 1]));
        }
        float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        if constexpr (SPLIT == 2) {
          // (alias layout: this barrier also orders the partner's S loads before our P stores into the columns it read)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch(j & 1, h)), "f"(m_tile) : "memory");
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          float other;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(xch(j & 1, h ^ 1)) : "memory");
          m_tile = fmaxf(m_tile, other);
        }
        const float m_new = fmaxf(m_run, m_tile);
        // d=64: the single P buffer is read by P(j-1) V.  (d=128: the commit behind s_full of this step covers the last reader
        // of the buffer S(j) | P(j) lives in.)
        if (!Cfg::ALIAS && j >= 1) mbar_wait(pv_done(i, (j - 1) & 1), (uint32_t)((j - 1) >> 1) & 1u);
        if (j == 0) {
          m_run = m_new;
        } else {
          const bool need = (m_new - m_run) * c > 8.0f;
          if (__any_sync(0xffffffffu, need)) {
            // O is accumulated by P(j-1) V: it must have finished before the rows are rescaled
            if (Cfg::ALIAS) mbar_wait(pv_done(i, (j - 1) & 1), (uint32_t)((j - 1) >> 1) & 1u);
            tc_fence_after();
            const float f = need ? ex2_approx((m_run - m_new) * c) : 1.0f;
            if (need) m_run = m_new;
            l_run *= f;
#pragma unroll 1
            for (int cb = 0; cb < OC / 16; ++cb) {  // 16 columns at a time: the score row stays live in registers
              uint32_t o[16];
              tmem_ld_x16(o_addr + (uint32_t)(cb * 16), o);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * f);
              tmem_st_x16(o_addr + (uint32_t)(cb * 16), o);
            }
            tmem_st_wait();
          }
        }
        tc_fence_after();
        if (tr) { const long long t2 = clock64(); ph_max += t2 - tt; tt = t2; }
        const float nmc = -m_run * c;
        const uint64_t mc2 = pack2(nmc, nmc);
        uint64_t sum2[2] = {0ull, 0ull};  // packed (even, odd) column partial sums
        const uint32_t p_addr = tile_base + Cfg::TM_P + (uint32_t)(j & 1) * Cfg::P_STRIDE + (uint32_t)(h * (HC / 2));
#pragma unroll
        for (int cb = 0; cb < HC / 32; ++cb) {
          uint32_t pk[16];
          if (cb * 32 >= valid) {  // (warp-uniform) a 32-column chunk past the last key: P = 0 without 32 exponentials per thread
#pragma unroll
            for (int k = 0; k < 16; ++k) pk[k] = 0u;
            tmem_st_x16(p_addr + (uint32_t)(cb * 16), pk);
            continue;
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float y0, y1, e0, e1;
            unpack2(fma2(pack2(__uint_as_float(r[cb * 32 + 2 * k]), __uint_as_float(r[cb * 32 + 2 * k + 1])), cc2, mc2), y0, y1);
            if ((k & 7) < Cfg::POLY) {
              exp2_poly2(y0, y1, e0, e1);
            } else {
              e0 = ex2_approx(y0);
              e1 = ex2_approx(y1);
            }
            sum2[k & 1] = add2(sum2[k & 1], pack2(e0, e1));
            pk[k] = pack_bf16(e0, e1);
          }
          tmem_st_x16(p_addr + (uint32_t)(cb * 16), pk);
        }
        {
          float s0, s1, s2, s3;
          unpack2(sum2[0], s0, s1);
          unpack2(sum2[1], s2, s3);
          l_run += (s0 + s1) + (s2 + s3);
        }
        if (tr) { const long long t2 = clock64(); ph_exp += t2 - tt; tt = t2; }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(i, j & 1));
        if (tr) { const long long t2 = clock64(); ph_st += t2 - tt; tt = t2; }
        if (j == 0 && warp == 4 && lane == 0) FMHA_TRACE(7);
      }
      if (warp == 4 && lane == 0) {
        FMHA_TRACE(8);
        FMHA_TRACE_ADD(16, ph_wait); FMHA_TRACE_ADD(17, ph_ld); FMHA_TRACE_ADD(18, ph_max); FMHA_TRACE_ADD(19, ph_exp); FMHA_TRACE_ADD(20, ph_st);
      }
      // ---- epilogue: O / l -> bf16 -> global ----
      if constexpr (SPLIT == 2) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch(n_kv & 1, h)), "f"(l_run) : "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        float other;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(xch(n_kv & 1, h ^ 1)) : "memory");
        l_run += other;
      }
      mbar_wait(pv_done(i, (n_kv - 1) & 1), (uint32_t)((n_kv - 1) >> 1) & 1u);  // the commit covers every earlier MMA too
      tc_fence_after();
      if (warp == 4 && lane == 0) FMHA_TRACE(9);
      const float inv_l = 1.0f / l_run;
      if (p.direct_store) {  // A/B (flags bit 6): per-thread row stores
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.O) + (long long)batch * p.o_bs + (long long)row * p.o_rs + (long long)head * p.o_hs + h * OC;
#pragma unroll
        for (int cb = 0; cb < OC / 32; ++cb) {
          uint32_t o[32];
          tmem_ld_x32(o_addr + (uint32_t)(cb * 32), o);
          tmem_ld_wait();
          if (row < p.len_q) {
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
              uint4 w;
              w.x = pack_bf16(__uint_as_float(o[k]) * inv_l, __uint_as_float(o[k + 1]) * inv_l);
              w.y = pack_bf16(__uint_as_float(o[k + 2]) * inv_l, __uint_as_float(o[k + 3]) * inv_l);
              w.z = pack_bf16(__uint_as_float(o[k + 4]) * inv_l, __uint_as_float(o[k + 5]) * inv_l);
              w.w = pack_bf16(__uint_as_float(o[k + 6]) * inv_l, __uint_as_float(o[k + 7]) * inv_l);
              *reinterpret_cast<uint4*>(orow + cb * 32 + k) = w;
            }
          }
        }
      } else {
      // O / l -> bf16 -> this tile's Q buffer (free: every MMA has completed) in the 128B-swizzled box layout, then ONE bulk tensor store
      // per 64-column slab.  (Per-thread row stores are uncoalesced -- 32 rows per instruction -- and cost ~7000 cycles per CTA.)
      // Rows past len_q are clipped by the TMA unit.
#pragma unroll
      for (int cb = 0; cb < OC / 32; ++cb) {
        uint32_t o[32];
        tmem_ld_x32(o_addr + (uint32_t)(cb * 32), o);
        tmem_ld_wait();
        const int col0 = h * OC + cb * 32;  // first of 32 consecutive output columns
        const uint32_t srow = smem_q(i) + (uint32_t)((col0 >> 6) * Cfg::Q_SLAB_BYTES + rit * 128);
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          const uint32_t chunk = (uint32_t)(((col0 & 63) + k) >> 3);
          const uint32_t dst = srow + ((chunk ^ ((uint32_t)rit & 7u)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                       "r"(pack_bf16(__uint_as_float(o[k]) * inv_l, __uint_as_float(o[k + 1]) * inv_l)),
                       "r"(pack_bf16(__uint_as_float(o[k + 2]) * inv_l, __uint_as_float(o[k + 3]) * inv_l)),
                       "r"(pack_bf16(__uint_as_float(o[k + 4]) * inv_l, __uint_as_float(o[k + 5]) * inv_l)),
                       "r"(pack_bf16(__uint_as_float(o[k + 6]) * inv_l, __uint_as_float(o[k + 7]) * inv_l))
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, %1;" ::"r"(9u + (uint32_t)i), "r"(128u * SPLIT) : "memory");  // all threads of this tile
      if (h == 0 && rit == 0) {
#pragma unroll
        for (int sl = 0; sl < Cfg::SLABS; ++sl)
          tma_store_4d(&tmO, smem_q(i) + sl * Cfg::Q_SLAB_BYTES, sl * 64, head, q0 + i * Cfg::BQ, batch);
        tma_store_commit();
        tma_store_wait_read<0>();  // the buffer must stay intact until the TMA unit has read it (the CTA exits next)
      }
      }
      if (warp == 4 && lane == 0) FMHA_TRACE(10);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) FMHA_TRACE(11);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}


static int make_qkv_map(CUtensorMap* tm, const void* ptr, long long B, long long H, long long L, long long D,
                        long long bs, long long rs, long long hs, uint32_t box_rows) {
  // dims (fastest first): head_dim, heads, rows, batch
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)H, (uint64_t)L, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)hs * 2, (uint64_t)rs * 2, (uint64_t)bs * 2};
  uint32_t box[4] = {64, 1, box_rows, 1};
  return encode_tensor_map(tm, ptr, 2, false, 4, dims, strides, box, true);
}

template <int D, int BKV_, int POLY_, int SPLIT_, int FAST_ = 0, int QT_ = 2>
static int launch_fmha(const vist3a_fmha_args& a, cudaStream_t stream) {
  using Cfg = FmhaCfg<D, BKV_, POLY_, SPLIT_, FAST_, QT_>;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  // a few rows more than a multiple of the CTA's query rows: those rows can go to CUDA-core tail CTAs (FmhaParams).  OPT-IN (flags bit 21):
  // measured on B200 at the decoder's frame attention (13 x 16 heads, 1029 x 1029, d 64; tools/fmha_tail_trace.py): a tail CTA takes 21.2 k
  // cycles (K by TMA 3.0 k, scores 5.5 k, softmax 3.4 k, P V 8.0 k: shared-memory wavefronts of the broadcast reads, 12 warps per SM) against
  // 25.3 k of the one-tile tensor-core CTA it replaces -- 377 vs 377 TFLOP/s for the launch, so the default stays without it
  const long long rows_per_cta = Cfg::BQ * Cfg::QT;
  long long tail = a.len_q % rows_per_cta;
  if (D != 64 || Cfg::BKV != 128 || tail > kTailMax || a.len_q < rows_per_cta || !(a.flags & (1u << 21)) || (a.flags & 4096u) ||
      a.len_kv <= 256 || fmha_tail_smem_bytes((int)tail, a.len_kv) > (size_t)Cfg::SMEM_BYTES)
    tail = 0;
  const long long len_q_main = a.len_q - tail;
  if ((rc = make_qkv_map(&tmQ, a.Q, a.batch, a.heads, len_q_main, D, a.q_bs, a.q_rs, a.q_hs, Cfg::BQ))) return rc;
  if ((rc = make_qkv_map(&tmK, a.K, a.batch, a.heads, a.len_kv, D, a.k_bs, a.k_rs, a.k_hs, Cfg::BKV))) return rc;
  if ((rc = make_qkv_map(&tmV, a.V, a.batch, a.heads, a.len_kv, D, a.v_bs, a.v_rs, a.v_hs, Cfg::BKV))) return rc;
  if ((rc = make_qkv_map(&tmO, a.O, a.batch, a.heads, len_q_main, D, a.o_bs, a.o_rs, a.o_hs, Cfg::BQ))) return rc;
  FmhaParams p;
  p.O = a.O; p.o_bs = a.o_bs; p.o_rs = a.o_rs; p.o_hs = a.o_hs;
  p.len_q = (int)len_q_main; p.len_kv = (int)a.len_kv;
  p.tail_rows = (int)tail; p.batch = (int)a.batch; p.heads = (int)a.heads;
  p.Qg = static_cast<const __nv_bfloat16*>(a.Q); p.Kg = static_cast<const __nv_bfloat16*>(a.K); p.Vg = static_cast<const __nv_bfloat16*>(a.V);
  p.q_bs = a.q_bs; p.q_rs = a.q_rs; p.q_hs = a.q_hs;
  p.k_bs = a.k_bs; p.k_rs = a.k_rs; p.k_hs = a.k_hs;
  p.v_bs = a.v_bs; p.v_rs = a.v_rs; p.v_hs = a.v_hs;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.row_scale = a.q_row_scale;
  p.trace = g_fmha_trace.load(std::memory_order_relaxed);
  p.zero = 0u;
  p.skip_softmax = (a.flags & 4096u) ? 1 : 0;
  p.single_issuer = (a.flags & 4u) ? 1 : 0;
  p.direct_store = (a.flags & 64u) ? 1 : 0;
  auto kern = fmha_fwd_kernel<D, BKV_, POLY_, SPLIT_, FAST_, QT_>;
  static std::atomic<unsigned long long> attr_done{0};  // per template instantiation, one bit per device
  V3A_CUDA_OK(ensure_dynamic_smem(kern, Cfg::SMEM_BYTES, attr_done));
  dim3 grid((unsigned)((len_q_main + rows_per_cta - 1) / rows_per_cta), (unsigned)a.heads, (unsigned)a.batch);
  if (tail) grid.z += (unsigned)((a.batch * a.heads + (long long)grid.x * grid.y - 1) / ((long long)grid.x * grid.y));
  V3A_REQUIRE(grid.z <= 65535u, VIST3A_ERR_INVALID, "fmha: batch exceeds grid limits");
  V3A_CUDA_OK(launch_kernel(kern, grid, dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, /*pdl=*/true, 1, tmQ, tmK, tmV, tmO, p));
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

int fmha_pair_entry(const vist3a_fmha_args& a, int variant, cudaStream_t stream, long long* ws_query);   // fmha_pair_sm100.cu

// ws_query != nullptr: nothing is launched; *ws_query = bytes of workspace the selected kernel would use for this problem (0: none)
int fmha_entry(const vist3a_fmha_args* args, cudaStream_t stream, long long* ws_query) {
  if (ws_query) *ws_query = 0;
  V3A_REQUIRE(args != nullptr, VIST3A_ERR_INVALID, "fmha: null args");
  const vist3a_fmha_args& a = *args;
  V3A_REQUIRE(a.Q && a.K && a.V && a.O, VIST3A_ERR_INVALID, "fmha: null Q/K/V/O");
  V3A_REQUIRE(a.batch > 0 && a.heads > 0 && a.len_q > 0 && a.len_kv > 0, VIST3A_ERR_INVALID,
              "fmha: batch/heads/len_q/len_kv must be positive");
  V3A_REQUIRE(a.head_dim == 64 || a.head_dim == 128, VIST3A_ERR_UNSUPPORTED, "fmha: head_dim %lld not in {64,128}",
              (long long)a.head_dim);
  V3A_REQUIRE(a.batch <= 65535 && a.heads <= 65535, VIST3A_ERR_INVALID, "fmha: batch/heads exceed grid limits");
  const long long st[] = {a.q_bs, a.q_rs, a.q_hs, a.k_bs, a.k_rs, a.k_hs, a.v_bs, a.v_rs, a.v_hs, a.o_bs, a.o_rs, a.o_hs};
  for (long long s : st) V3A_REQUIRE(s % 8 == 0 && s >= 0, VIST3A_ERR_INVALID, "fmha: strides must be multiples of 8 elements");
  V3A_REQUIRE((((uintptr_t)a.Q | (uintptr_t)a.K | (uintptr_t)a.V | (uintptr_t)a.O) & 15) == 0, VIST3A_ERR_INVALID,
              "fmha: pointers must be 16-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  // Default: two threads per query row (4 softmax warpgroups, 640 threads), 64-key double-buffered steps at d=128, 128-key steps at
  // d=64 -- the fastest of the variants measured on B200 (tools/fmha_variants.py).  flags bit0 selects one thread per row (2 softmax
  // warpgroups, setmaxnreg-enlarged register file), bit1 (d=128) the 128-key aliased steps, bit2 a single MMA-issuing warp for both query tiles (measured slower: one thread cannot keep
  // the pipe fed); kept for A/B measurements.
  // head_dim 128: the CTA-pair kernel (fmha_pair_sm100.cu); flags bit 8 selects it, bits 9-11 its variant (A/B measurements)
  if (a.head_dim == 128 && (a.flags & 256u)) return fmha_pair_entry(a, (int)((a.flags >> 9) & 7u), stream, ws_query);
  const bool one = (a.flags & 1u) != 0;
  // Default (flags == 0), from same-process A/B runs on B200 (tools/fmha_variants.py, tools/fmha_pair_check.py; profiles/README.md): the speculative
  // softmax (stale running maximum, 64-column half-steps, one thread per row, 2-3 of 8 exponentials on the FMA pipe) everywhere; at head_dim 128
  // and >= 1024 keys on CTA pairs (+5.5 % over the former default at 4096 keys, device time in a CUDA graph), below that on one CTA (+1.7 % at 512
  // keys; +5.5 % on the decoder's 1029-key frame attention, +2.7 % on its 13 377-key global attention).  flags bit 13
  // selects the former default (two threads per row, exact running maximum) for A/B.
  // (flags bits 17-20 steer the work decomposition of the pair kernel, bit 21 switches the tail-row CTAs on; they do not select a kernel)
  if ((a.flags & ~(63u << 17)) == 0u) {
    if (ws_query && (a.head_dim == 64 || a.len_kv < 512)) return VIST3A_OK;
    // (head_dim 64, long key sequences -- the decoder's global attention: two threads per row, +2 % over one thread per row at 13 377 keys; -7 % at 1029)
    if (a.head_dim == 64) return a.len_kv >= 4096 ? launch_fmha<64, 128, 0, 2, 2>(a, stream) : launch_fmha<64, 128, 2, 1, 1>(a, stream);
    // (>= 512 keys: CTA pairs with 2 of every 8 exponentials on the FMA pipe.  On the persistent kernel 1 or 2 of 8 are 2.3 % ahead of 3 of 8 --
    //  the round-2 choice for >= 1024 keys -- at 4096 keys: 1151-1156 vs 1123-1132 TFLOP/s, and 817 vs 780-790 at the 512 text keys;
    //  tools/runs/gpu_r4s.sh, profiles/r2_fmha_variants_persistent.jsonl)
    if (a.len_kv >= 512) return fmha_pair_entry(a, 4, stream, ws_query);
    return launch_fmha<128, 64, 2, 1, 1>(a, stream);
  }
  if (ws_query) return VIST3A_OK;   // the single-CTA variants below use no workspace
  if (a.flags & 65536u) {   // ONE query tile per CTA, speculative softmax with two threads per row; bits 3-5 = FMA-pipe exponentials per 8 pairs
    const unsigned np = (a.flags >> 3) & 7u;
    if (a.head_dim == 64) {
      switch (np) {
        case 0: return launch_fmha<64, 128, 0, 2, 2, 1>(a, stream);
        case 1: return launch_fmha<64, 128, 1, 2, 2, 1>(a, stream);
        case 2: return launch_fmha<64, 128, 2, 2, 2, 1>(a, stream);
        default: return launch_fmha<64, 128, 3, 2, 2, 1>(a, stream);
      }
    }
    switch (np) {
      case 0: return launch_fmha<128, 128, 0, 2, 2, 1>(a, stream);
      case 1: return launch_fmha<128, 128, 1, 2, 2, 1>(a, stream);
      case 2: return launch_fmha<128, 128, 2, 2, 2, 1>(a, stream);
      default: return launch_fmha<128, 128, 3, 2, 2, 1>(a, stream);
    }
  }
  if (a.flags & 16384u) {   // speculative softmax, two threads per row (128-key steps); bits 3-5 = FMA-pipe exponentials per 8 column pairs
    const unsigned np = (a.flags >> 3) & 7u;
    if (a.head_dim == 64) {
      switch (np) {
        case 0: return launch_fmha<64, 128, 0, 2, 2>(a, stream);
        case 1: return launch_fmha<64, 128, 1, 2, 2>(a, stream);
        case 2: return launch_fmha<64, 128, 2, 2, 2>(a, stream);
        default: return launch_fmha<64, 128, 3, 2, 2>(a, stream);
      }
    }
    switch (np) {
      case 0: return launch_fmha<128, 128, 0, 2, 2>(a, stream);
      case 2: return launch_fmha<128, 128, 2, 2, 2>(a, stream);
      default: return launch_fmha<128, 128, 3, 2, 2>(a, stream);
    }
  }
  if (a.flags & 128u) {   // speculative softmax, one thread per row; bits 3-5 = FMA-pipe exponentials per 8 column pairs
    const unsigned np = (a.flags >> 3) & 7u;
    if (a.head_dim == 64) {
      if (a.flags & 32768u) return launch_fmha<64, 128, 2, 1, 3>(a, stream);   // both score halves of a step loaded at once
      switch (np) {
        case 0: return launch_fmha<64, 128, 0, 1, 1>(a, stream);
        case 1: return launch_fmha<64, 128, 1, 1, 1>(a, stream);
        case 2: return launch_fmha<64, 128, 2, 1, 1>(a, stream);
        case 3: return launch_fmha<64, 128, 3, 1, 1>(a, stream);
        default: return launch_fmha<64, 128, 4, 1, 1>(a, stream);
      }
    }
    if (a.flags & 2u) {   // aliased 128-key steps, P handed over in two halves
      switch (np) {
        case 0: return launch_fmha<128, 128, 0, 1, 1>(a, stream);
        case 1: return launch_fmha<128, 128, 1, 1, 1>(a, stream);
        case 2: return launch_fmha<128, 128, 2, 1, 1>(a, stream);
        case 3: return launch_fmha<128, 128, 3, 1, 1>(a, stream);
        default: return launch_fmha<128, 128, 4, 1, 1>(a, stream);
      }
    }
    switch (np) {
      case 0: return launch_fmha<128, 64, 0, 1, 1>(a, stream);
      case 1: return launch_fmha<128, 64, 1, 1, 1>(a, stream);
      case 2: return launch_fmha<128, 64, 2, 1, 1>(a, stream);
      case 3: return launch_fmha<128, 64, 3, 1, 1>(a, stream);
      default: return launch_fmha<128, 64, 4, 1, 1>(a, stream);
    }
  }
  // Share of the exponentials moved from the MUFU to the FMA pipe (Cody-Waite + polynomial), per 8 column pairs.  Measured on B200 with
  // two threads per query row (4 softmax warpgroups): the softmax is bound by issue slots and dependent-latency chains rather than by
  // the MUFU, so the default is 0 (all MUFU.EX2): +4 % at d=128 / 4096 keys, +9 % at d=64 / 13377 keys over the former 2 / 3 of 8; short
  // key sequences (cross-attention, 512 keys) keep 2 of 8.  flags >> 3 = 1 + share selects a variant explicitly (A/B).
  const unsigned poly_sel = (a.flags >> 3) & 7u;
  if (a.head_dim == 64) {
    if (one) return launch_fmha<64, 128, 3, 1>(a, stream);
    if (poly_sel == 2u) return launch_fmha<64, 128, 1, 2>(a, stream);
    if (poly_sel == 3u) return launch_fmha<64, 128, 2, 2>(a, stream);
    if (poly_sel == 4u) return launch_fmha<64, 128, 3, 2>(a, stream);
    return launch_fmha<64, 128, 0, 2>(a, stream);
  }
  if (a.flags & 2u) return one ? launch_fmha<128, 128, 2, 1>(a, stream) : launch_fmha<128, 128, 2, 2>(a, stream);
  if (one) return launch_fmha<128, 64, 2, 1>(a, stream);
  if (poly_sel == 1u) return launch_fmha<128, 64, 0, 2>(a, stream);
  if (poly_sel == 2u) return launch_fmha<128, 64, 1, 2>(a, stream);
  if (poly_sel == 3u) return launch_fmha<128, 64, 2, 2>(a, stream);
  if (poly_sel == 4u) return launch_fmha<128, 64, 3, 2>(a, stream);
  return a.len_kv >= 2048 ? launch_fmha<128, 64, 0, 2>(a, stream) : launch_fmha<128, 64, 2, 2>(a, stream);
}

}  // namespace v3a
