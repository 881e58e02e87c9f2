// tcgen05 GEMM with fused epilogues:  C[M,N] = epilogue(A[M,K] · W[N,K]^T)
//
// Persistent, warp-specialised kernel (one CTA — or one CTA pair — per SM):
//   warp 0       TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, mbarrier ring)
//   warp 1       MMA issuer     (one elected thread, tcgen05.mma, accumulators in TMEM)
//   warp 2       TMEM allocator
//   warps 4..11  epilogue       (tcgen05.ld -> registers -> bias/act/gate/residual -> global)
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
// kCta == 2 uses cta_group::2: the pair computes a 256 x BN tile, each CTA stages its own 128 rows
// of A and half of the W tile, the leader CTA issues the MMAs for both.
//
// Convolution mode (implicit GEMM): the A tile of 128 rows is an 8 x 16 patch of output pixels of one
// NHWC image; for k-block (tap dy,dx ; channel block cb) the producer issues ONE 4-D TMA load at
// coordinates (cb*BK, x0+dx-pad, y0+dy-pad, n): the box lands in smem exactly as the 128B-swizzled
// K-major tile the MMA wants, and taps that fall outside the image are zero-filled by the TMA unit,
// so there is no im2col buffer and no padding branch anywhere.
#include "common.cuh"
#include "host_util.cuh"

namespace v3a {

struct RowMap {
  long long rpg, gstride, goff;
  __device__ __forceinline__ long long operator()(long long row) const {
    return rpg > 0 ? (row / rpg) * gstride + goff + row % rpg : row;
  }
};

struct GemmEpilogue {
  void* C;
  const float* bias;
  const float* gate;
  const void* residual;
  const void* residual2;
  long long ldc, ldr;
  long long rows_per_batch, gate_bstride;
  RowMap cmap, rmap;
  int out_fp32;
  int act, post_act;
  int round_linear;
  int round_gate;
  long long* trace;  // debug (v3a_debug_gemm_trace): 8 cycle counters per CTA, normally null
};

// debug hook (tools/gemm_trace.py), off unless armed: process-wide by design, read once per launch
static std::atomic<long long*> g_gemm_trace{nullptr};
extern "C" void v3a_debug_gemm_trace(void* buf) { g_gemm_trace.store(reinterpret_cast<long long*>(buf)); }

constexpr int kConvTW = 16, kConvTH = 8;  // pixel patch of one 128-row A tile

struct GemmShape {
  int M, N, K;
  int tiles_m, tiles_n;  // tiles_m counts 128*kCta-row tiles
  int mc;                // clusters of two CTA pairs on neighbouring column tiles, the shared A rows multicast between them (kCta == 2 only)
  // conv mode
  int conv, kh, kw, pad_y, pad_x, n_img, h, w, c_in, h_out, w_out, tiles_x, tiles_y, cblocks;
  int kt;                // causal temporal taps (>= 1): the image coordinate of tap dt is img + dt - (kt - 1); negative = zero fill
};

template <int BN, int kCta, bool kTF32>
struct GemmCfg {
  static constexpr int BM = 128;                       // rows per CTA
  static constexpr int ES = kTF32 ? 4 : 2;             // operand element size
  static constexpr int BK = 128 / ES;                  // one 128-byte swizzle row of K per stage
  static constexpr int UK = 32 / ES;                   // K per tcgen05.mma
  static constexpr int BN_LOCAL = BN / kCta;           // W rows staged by this CTA
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BN_LOCAL * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = 8 * 32 * 128;   // per epilogue warp: 32 rows x 128 bytes
  static constexpr int RING_BUDGET = 227 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/ - STAGING_BYTES;
  static constexpr int STAGES = RING_BUDGET / STAGE_BYTES > 8 ? 8 : RING_BUDGET / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;  // two accumulator stages
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + STAGING_BYTES;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  // BN = 176 (pairs only, A/B flag): 9 column tiles of a 1536-wide output fill 4 waves of 74 CTA pairs to 97 % (6 tiles of 256: 3 waves at 86 %)
  // BN = 96 / 192: output widths that are multiples of 96 but not of 128 (the Wan VAE's 96 / 192 / 384-channel layers, the 84-wide Gaussian
  // head): a 128 / 256-wide tile would spend a quarter of its MMA columns on padding
  static_assert(BN == 64 || BN == 96 || BN == 128 || BN == 192 || BN == 256 || (BN == 176 && kCta == 2), "BN");
  static_assert(BN % 16 == 0 && BN_LOCAL % 8 == 0, "MMA N granularity (16) / swizzle atom rows (8)");
  static_assert(TMEM_COLS <= 512, "TMEM");
};

template <int ACT>
__device__ __forceinline__ void act32(float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if constexpr (ACT == VIST3A_ACT_GELU_TANH) v[j] = gelu_tanh_f(v[j]);
    else if constexpr (ACT == VIST3A_ACT_GELU_ERF) v[j] = gelu_erf_fast_f(v[j]);
    else if constexpr (ACT == VIST3A_ACT_SILU) v[j] = silu_f(v[j]);
    else if constexpr (ACT == VIST3A_ACT_RELU) v[j] = fmaxf(v[j], 0.0f);
  }
}
// one warp-uniform dispatch per 32-column chunk (a per-element switch compiles to a jump table per element)
__device__ __forceinline__ void apply_act32(float (&v)[32], int act) {
  switch (act) {
    case VIST3A_ACT_GELU_TANH: act32<VIST3A_ACT_GELU_TANH>(v); break;
    case VIST3A_ACT_GELU_ERF: act32<VIST3A_ACT_GELU_ERF>(v); break;
    case VIST3A_ACT_SILU: act32<VIST3A_ACT_SILU>(v); break;
    case VIST3A_ACT_RELU: act32<VIST3A_ACT_RELU>(v); break;
    default: break;
  }
}

template <bool kF32>
__device__ __forceinline__ void add_row32(float (&v)[32], const void* base, long long off, int ncols) {
  if constexpr (kF32) {
    const float* p = reinterpret_cast<const float*>(base) + off;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j < ncols) {
        const float4 x = *reinterpret_cast<const float4*>(p + j);
        v[j] += x.x; v[j + 1] += x.y; v[j + 2] += x.z; v[j + 3] += x.w;
      }
    }
  } else {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base) + off;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (j < ncols) {
        const uint4 x = *reinterpret_cast<const uint4*>(p + j);
        v[j] += bf16_lo(x.x); v[j + 1] += bf16_hi(x.x);
        v[j + 2] += bf16_lo(x.y); v[j + 3] += bf16_hi(x.y);
        v[j + 4] += bf16_lo(x.z); v[j + 5] += bf16_hi(x.z);
        v[j + 6] += bf16_lo(x.w); v[j + 7] += bf16_hi(x.w);
      }
    }
  }
}

constexpr int kGemmThreads = 384;  // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue

// kDirect: fp32 output written (and an fp32 residual read) by 256-bit per-thread accesses in the accumulator's row layout, no staging buffer.
// A separate instantiation, so that its registers do not weigh on the staged path (a spill there goes to L2: the kernel leaves no L1).
template <int BN, int kCta, bool kTF32, bool kDirect = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmA64,
                    const GemmShape shape, const GemmEpilogue ep) {
  using Cfg = GemmCfg<BN, kCta, kTF32>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms need 1024-byte alignment
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  auto smem_a = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
  auto smem_b = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };

  const uint32_t warp = warp_id_sync();
  const uint32_t lane = lane_id();
  // Multicast mode (shape.mc): the cluster holds TWO pairs working on column tiles 2 tnp and 2 tnp + 1 of the same 256 rows.  Every CTA
  // fetches one 64-row half of its 128 A rows and multicasts it to its twin in the other pair (same rank inside the pair), so the A
  // operand crosses L2 -> SM once per cluster instead of once per pair (-25 % operand traffic per flop).
  const uint32_t crank = (kCta == 2) ? cluster_ctarank() : 0u;
  const uint32_t rank = crank & 1u;       // rank inside the CTA pair
  const uint32_t pr = crank >> 1;         // pair inside the cluster (0 unless shape.mc)
  const int csize = (kCta == 2 && shape.mc) ? 4 : kCta;
  const bool leader = (rank == 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), shape.mc ? 2 : 1);  // multicast: a slot is free once BOTH pairs of the cluster have consumed it
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8 * kCta);  // one arrival per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kCta>(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  if constexpr (kCta == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // PDL: the set-up above overlapped the tail of the previous kernel; its results are needed from here on
  pdl_launch_dependents();
  pdl_wait();

  const int num_kb = shape.conv ? shape.kt * shape.kh * shape.kw * shape.cblocks : (shape.K + Cfg::BK - 1) / Cfg::BK;
  const int num_tiles = shape.mc ? shape.tiles_m * (shape.tiles_n / 2) : shape.tiles_m * shape.tiles_n;
  const int worker = blockIdx.x / csize;
  const int num_workers = gridDim.x / csize;
  auto tile_n_of = [&](int tile) { return shape.mc ? 2 * (tile / shape.tiles_m) + (int)pr : tile / shape.tiles_m; };

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    // The whole warp runs the (warp-uniform) control flow and the barrier waits; one elected lane issues, so the
    // coordinates / descriptors stay in uniform registers (no per-instruction R2UR + elect waterfall).
    int s = 0;
    uint32_t ph = 0;
    long long tr_wait = 0;
    const long long tr_t0 = ep.trace ? clock64() : 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers) {
      const int tm = tile % shape.tiles_m, tn = tile_n_of(tile);
      const int mt = tm * kCta + (int)rank;  // this CTA's 128-row tile
      const int m0 = mt * Cfg::BM;
      const int n0 = tn * BN + (int)rank * Cfg::BN_LOCAL;
      int img = 0, y0 = 0, x0 = 0;
      if (shape.conv) {
        const int per_img = shape.tiles_x * shape.tiles_y;
        img = mt / per_img - (shape.kt - 1);
        const int t2 = mt % per_img;
        y0 = (t2 / shape.tiles_x) * kConvTH - shape.pad_y;
        x0 = (t2 % shape.tiles_x) * kConvTW - shape.pad_x;
      }
      int cb = 0, dy = 0, dx = 0, dt = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (ep.trace) {
          const long long t1 = clock64();
          mbar_wait(empty_bar(s), ph ^ 1u);
          tr_wait += clock64() - t1;
        } else {
          mbar_wait(empty_bar(s), ph ^ 1u);
        }
        if (elect_one()) {
          if constexpr (kCta == 1) {
            mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
            if (shape.conv) tma_load_4d(smem_a(s), &tmA, full_bar(s), cb * Cfg::BK, x0 + dx, y0 + dy, img + dt);
            else tma_load_2d(smem_a(s), &tmA, full_bar(s), kb * Cfg::BK, m0);
            tma_load_2d(smem_b(s), &tmB, full_bar(s), kb * Cfg::BK, n0);
          } else {
            if (leader) mbar_expect_tx(full_bar(s), 2 * Cfg::STAGE_BYTES);
            if (shape.conv) tma_load_4d_2sm(smem_a(s), &tmA, full_bar(s), cb * Cfg::BK, x0 + dx, y0 + dy, img + dt);
            else if (shape.mc)
              tma_load_2d_2sm_mc(smem_a(s) + pr * (Cfg::A_BYTES / 2), &tmA64, full_bar(s), kb * Cfg::BK, m0 + (int)pr * (Cfg::BM / 2),
                                 (uint16_t)((1u << rank) | (1u << (rank + 2))));
            else tma_load_2d_2sm(smem_a(s), &tmA, full_bar(s), kb * Cfg::BK, m0);
            tma_load_2d_2sm(smem_b(s), &tmB, full_bar(s), kb * Cfg::BK, n0);
          }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
        if (shape.conv && ++cb == shape.cblocks) {
          cb = 0;
          if (++dx == shape.kw) {
            dx = 0;
            if (++dy == shape.kh) { dy = 0; ++dt; }
          }
        }
      }
    }
    if (ep.trace && lane == 0) {
      ep.trace[blockIdx.x * 8 + 0] = clock64() - tr_t0;  // producer: total
      ep.trace[blockIdx.x * 8 + 1] = tr_wait;            // producer: waiting for a free stage
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (same scheme: uniform control flow, elected issue) -------------
    if (leader) {
      constexpr uint32_t idesc = make_idesc(kTF32 ? kFmtTF32 : kFmtBF16, 128 * kCta, BN, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      long long tr_full = 0, tr_tempty = 0;
      const long long tr_t0 = ep.trace ? clock64() : 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        if (ep.trace) {
          const long long t1 = clock64();
          mbar_wait(tempty_bar(as), aph ^ 1u);
          tr_tempty += clock64() - t1;
        } else {
          mbar_wait(tempty_bar(as), aph ^ 1u);
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          if (ep.trace) {
            const long long t1 = clock64();
            mbar_wait(full_bar(s), ph);
            tr_full += clock64() - t1;
          } else {
            mbar_wait(full_bar(s), ph);
          }
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = make_smem_desc_sw128(smem_a(s), 1024, 0);
            const uint64_t bdesc = make_smem_desc_sw128(smem_b(s), 1024, 0);
#pragma unroll
            for (int k = 0; k < Cfg::BK / Cfg::UK; ++k) {
              // advance 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
              if constexpr (kTF32) umma_tf32_ss<kCta>(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) ? 1u : 0u);
              else umma_f16_ss<kCta>(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) ? 1u : 0u);
            }
            if constexpr (kCta == 1) umma_commit(empty_bar(s)); else umma_commit_2sm_mc(empty_bar(s), shape.mc ? 15 : 3);
            if (kb == num_kb - 1) {
              if constexpr (kCta == 1) umma_commit(tfull_bar(as)); else umma_commit_2sm_mc(tfull_bar(as), (uint16_t)(3u << (2 * pr)));
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
      if (ep.trace && lane == 0) {
        ep.trace[blockIdx.x * 8 + 2] = clock64() - tr_t0;  // MMA warp: total
        ep.trace[blockIdx.x * 8 + 3] = tr_full;            // MMA warp: waiting for operands (TMA)
        ep.trace[blockIdx.x * 8 + 4] = tr_tempty;          // MMA warp: waiting for a free accumulator stage (epilogue)
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    // Two warps per TMEM lane quadrant: warp w reads lanes [32 (w%4), +32) and every other 32-column chunk.
    const uint32_t q = warp & 3u;
    const int half = (int)((warp - 4u) >> 2);
    const uint32_t stage_base = bar_base + 256u + (warp - 4u) * 4096u;  // this warp's staging buffer
    constexpr int NCHUNK = (BN + 31) / 32;  // BN = 176: the last chunk holds 16 columns (the TMEM read runs into the other stage, unused)
    int as = 0;
    uint32_t aph = 0;
    long long tr_tfull = 0;
    const long long tr_t0 = ep.trace ? clock64() : 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers) {
      const int tm = tile % shape.tiles_m, tn = tile_n_of(tile);
      const int mt = tm * kCta + (int)rank;
      const int r_in_tile = (int)(q * 32u + lane);
      long long row;
      bool row_ok;
      if (shape.conv) {
        const int per_img = shape.tiles_x * shape.tiles_y;
        const int img = mt / per_img, t2 = mt % per_img;
        const int y = (t2 / shape.tiles_x) * kConvTH + (r_in_tile >> 4);
        const int x = (t2 % shape.tiles_x) * kConvTW + (r_in_tile & 15);
        row_ok = img < shape.n_img && y < shape.h_out && x < shape.w_out;
        row = ((long long)img * shape.h_out + y) * shape.w_out + x;
      } else {
        row = (long long)mt * Cfg::BM + r_in_tile;
        row_ok = row < shape.M;
      }
      const int n_tile = tn * BN;
      const long long b = row_ok ? row / ep.rows_per_batch : 0;
      const float* gate_row = ep.gate ? ep.gate + b * ep.gate_bstride : nullptr;
      const long long crow = row_ok ? ep.cmap(row) : 0;
      const long long rrow = (row_ok && ep.residual) ? ep.rmap(row) : 0;
      if (ep.trace) {
        const long long t1 = clock64();
        mbar_wait(tfull_bar(as), aph);
        tr_tfull += clock64() - t1;
      } else {
        mbar_wait(tfull_bar(as), aph);
      }
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((q * 32u) << 16) + (uint32_t)(as * BN);
      // Chunk k of this warp.  bf16 output: the two warps of a lane quadrant take alternating PAIRS of adjacent 32-column chunks
      // (0,1,4,5 | 2,3,6,7), so a warp stages 64 columns = one full 128-byte line per row before it writes; fp32 output (and BN = 64):
      // one 32-column chunk at a time.
      constexpr bool kPairs = NCHUNK >= 4;
      auto chunk_of = [&](int k) { return kPairs ? (k >> 1) * 4 + half * 2 + (k & 1) : 2 * k + half; };
      auto chunk_ok = [&](int k) { const int c = chunk_of(k); return c < NCHUNK && n_tile + c * 32 < shape.N; };
      uint32_t r[32];
      if (chunk_ok(0)) tmem_ld_x32(taddr + (uint32_t)(chunk_of(0) * 32), r);
      bool released = false;
      int blk_n0 = 0, blk_bytes = 0;  // first column / valid bytes per row of the block being staged
      if constexpr (kDirect) {
        // ---- fp32 output without staging: every thread owns 32 consecutive columns of its row = 4 sectors of 32 bytes, which LDG.256 /
        //      STG.256 move whole.  The staging that coalesces 16-byte pieces into lines costs 512 KB of shared-memory traffic per
        //      128 x 256 tile with a residual, on top of what the main loop moves through the same 128 B/clk port.
        (void)blk_n0; (void)blk_bytes;
#pragma unroll 1
        for (int k = 0; chunk_ok(k); ++k) {
          const int n0 = n_tile + chunk_of(k) * 32;
          float4 bvv[8];
          if (ep.bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) bvv[j] = __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + j);
          }
          const char* rp = (ep.residual && row_ok) ? reinterpret_cast<const char*>(ep.residual) + (rrow * ep.ldr + n0) * (ep.out_fp32 ? 4 : 2) : nullptr;
          uint32_t rs[32];   // fp32: 32 values; bf16: 16 packed pairs
          if (rp) {   // issued before the wait for the accumulator
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (ep.out_fp32 || j < 16)
                asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(rs[j]), "=r"(rs[j + 1]), "=r"(rs[j + 2]), "=r"(rs[j + 3]), "=r"(rs[j + 4]), "=r"(rs[j + 5]), "=r"(rs[j + 6]), "=r"(rs[j + 7])
                             : "l"(rp + 4 * j));
            }
          }
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (chunk_ok(k + 1)) {
            tmem_ld_x32(taddr + (uint32_t)(chunk_of(k + 1) * 32), r);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (kCta == 1) mbar_arrive(tempty_bar(as)); else mbar_arrive_cluster(tempty_bar(as), 2 * pr);
            }
            released = true;
          }
          if (row_ok) {
            if (ep.bias) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 bv = bvv[j >> 2];
                v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
              }
            }
            if (ep.act != VIST3A_ACT_NONE) apply_act32(v, ep.act);
            if (ep.round_linear) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);
            }
            if (gate_row) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gate_row + n0 + j));
                v[j] *= g.x; v[j + 1] *= g.y; v[j + 2] *= g.z; v[j + 3] *= g.w;
              }
              if (ep.round_gate) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);
              }
            }
            if (rp) {
              if (ep.out_fp32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(rs[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) { v[2 * j] += bf16_lo(rs[j]); v[2 * j + 1] += bf16_hi(rs[j]); }
              }
            }
            if (ep.post_act == VIST3A_ACT_RELU) act32<VIST3A_ACT_RELU>(v);
            if (ep.out_fp32) {
              float* cp = reinterpret_cast<float*>(ep.C) + crow * ep.ldc + n0;
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(cp + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]),
                             "f"(v[j + 4]), "f"(v[j + 5]), "f"(v[j + 6]), "f"(v[j + 7])
                             : "memory");
            } else {
              __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(ep.C) + crow * ep.ldc + n0;
#pragma unroll
              for (int j = 0; j < 32; j += 16)
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(cp + j), "r"(pack_bf16(v[j], v[j + 1])), "r"(pack_bf16(v[j + 2], v[j + 3])),
                             "r"(pack_bf16(v[j + 4], v[j + 5])), "r"(pack_bf16(v[j + 6], v[j + 7])), "r"(pack_bf16(v[j + 8], v[j + 9])),
                             "r"(pack_bf16(v[j + 10], v[j + 11])), "r"(pack_bf16(v[j + 12], v[j + 13])), "r"(pack_bf16(v[j + 14], v[j + 15]))
                             : "memory");
            }
          }
        }
      } else
#pragma unroll 1
      for (int k = 0; chunk_ok(k); ++k) {
        const int n0 = n_tile + chunk_of(k) * 32;
        // bias of this chunk: issued before the wait for the accumulator so that its (L2) latency overlaps the TMEM read
        // (ncu: the FADDs consuming a just-issued bias load were the hottest stall of the epilogue)
        float4 bvv[8];
        if (ep.bias) {
#pragma unroll
          for (int j = 0; j < 8; ++j) bvv[j] = (n0 + 4 * j < shape.N) ? __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        // prefetch the next chunk of this warp while this one is processed; after the last read the accumulator stage goes back to
        // the MMA warp at once (the arithmetic and the stores below no longer need it)
        const bool more = chunk_ok(k + 1);
        if (more) {
          tmem_ld_x32(taddr + (uint32_t)(chunk_of(k + 1) * 32), r);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kCta == 1) mbar_arrive(tempty_bar(as)); else mbar_arrive_cluster(tempty_bar(as), 2 * pr);
          }
          released = true;
        }
        const int ncols = min(min(32, shape.N - n0), n_tile + BN - n0);
        const int es = ep.out_fp32 ? 4 : 2;
        const bool first_of_block = ep.out_fp32 || !kPairs || (k & 1) == 0;
        if (first_of_block) { blk_n0 = n0; blk_bytes = 0; }
        const uint32_t srow = stage_base + lane * 128u;
        const uint32_t byte0 = (uint32_t)blk_bytes;  // offset of this chunk inside the staged row (0 or 64)
        if (ep.residual && first_of_block) {
          // the residual rows of this block come in the same way the results go out: row-contiguous 16-byte pieces (full lines)
          // into the staging buffer; every thread then picks its own row's values from shared memory
          int blk_cols = ncols;
          if (!ep.out_fp32 && kPairs && chunk_ok(k + 1)) blk_cols += min(min(32, shape.N - n0 - 32), n_tile + BN - n0 - 32);
          const int rbytes = blk_cols * es;
          const uint32_t piece = lane & 7u;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (int)(lane >> 3);
            const long long rrow_r = __shfl_sync(0xffffffffu, rrow, rr);
            const int ok_r = __shfl_sync(0xffffffffu, (int)row_ok, rr);
            if (ok_r && (int)(piece << 4) < rbytes) {
              const uint4 val = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(ep.residual) + (rrow_r * ep.ldr + n0) * es + (piece << 4));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_base + (uint32_t)rr * 128u + ((piece ^ ((uint32_t)rr & 7u)) << 4)),
                           "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w)
                           : "memory");
            }
          }
          __syncwarp();
        }
        if (row_ok) {
          if (ep.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = bvv[j >> 2];
              v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
            }
          }
          if (ep.act != VIST3A_ACT_NONE) apply_act32(v, ep.act);
          if (ep.round_linear) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);
          }
          if (gate_row) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j < ncols) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gate_row + n0 + j));
                v[j] *= g.x; v[j + 1] *= g.y; v[j + 2] *= g.z; v[j + 3] *= g.w;
              }
            }
            if (ep.round_gate) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);
            }
          }
          // stage the finished values in this warp's 32 x 128-byte buffer (16-byte pieces XOR-swizzled by the row: conflict-free
          // both for these per-row writes and for the row-contiguous reads below)
          if (ep.out_fp32) {
            if (ep.residual) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (j < ncols) {
                  float4 x;
                  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                               : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                               : "r"(srow + (((uint32_t)(j >> 2)) ^ (lane & 7u)) * 16u)
                               : "memory");
                  v[j] += x.x; v[j + 1] += x.y; v[j + 2] += x.z; v[j + 3] += x.w;
                }
              }
            }
            if (ep.residual2) add_row32<true>(v, ep.residual2, crow * ep.ldc + n0, ncols);
            if (ep.post_act == VIST3A_ACT_RELU) act32<VIST3A_ACT_RELU>(v);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j < ncols) {
                const uint32_t piece = (uint32_t)(j >> 2);
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((piece ^ (lane & 7u)) << 4)), "f"(v[j]), "f"(v[j + 1]),
                             "f"(v[j + 2]), "f"(v[j + 3])
                             : "memory");
              }
            }
          } else {
            if (ep.residual) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                if (j < ncols) {
                  uint4 x;
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                               : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
                               : "r"(srow + ((((byte0 >> 4) + (uint32_t)(j >> 3)) ^ (lane & 7u)) << 4))
                               : "memory");
                  v[j] += bf16_lo(x.x); v[j + 1] += bf16_hi(x.x);
                  v[j + 2] += bf16_lo(x.y); v[j + 3] += bf16_hi(x.y);
                  v[j + 4] += bf16_lo(x.z); v[j + 5] += bf16_hi(x.z);
                  v[j + 6] += bf16_lo(x.w); v[j + 7] += bf16_hi(x.w);
                }
              }
            }
            if (ep.residual2) add_row32<false>(v, ep.residual2, crow * ep.ldc + n0, ncols);
            if (ep.post_act == VIST3A_ACT_RELU) act32<VIST3A_ACT_RELU>(v);
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (j < ncols) {
                const uint32_t piece = (byte0 >> 4) + (uint32_t)(j >> 3);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((piece ^ (lane & 7u)) << 4)), "r"(pack_bf16(v[j], v[j + 1])),
                             "r"(pack_bf16(v[j + 2], v[j + 3])), "r"(pack_bf16(v[j + 4], v[j + 5])), "r"(pack_bf16(v[j + 6], v[j + 7]))
                             : "memory");
              }
            }
          }
        }
        blk_bytes += ncols * es;
        // flush: rows of the block are written as contiguous runs -- lane l moves piece l % 8 of row l / 8 + 4 it (up to 4 full 128-byte
        // lines per instruction instead of 32 scattered 16-byte pieces)
        const bool flush = ep.out_fp32 || !kPairs || (k & 1) == 1 || !more;
        if (flush) {
          __syncwarp();
          const uint32_t piece = lane & 7u;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (int)(lane >> 3);
            const long long crow_r = __shfl_sync(0xffffffffu, crow, rr);
            const int ok_r = __shfl_sync(0xffffffffu, (int)row_ok, rr);
            if (ok_r && (int)(piece << 4) < blk_bytes) {
              uint4 val;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                           : "r"(stage_base + (uint32_t)rr * 128u + ((piece ^ ((uint32_t)rr & 7u)) << 4))
                           : "memory");
              char* dst = reinterpret_cast<char*>(ep.C) + (crow_r * ep.ldc + blk_n0) * es + (piece << 4);
              *reinterpret_cast<uint4*>(dst) = val;
            }
          }
          __syncwarp();
        }
      }
      if (!released) {  // no chunk of this tile belongs to this warp: still hand the accumulator stage back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kCta == 1) mbar_arrive(tempty_bar(as)); else mbar_arrive_cluster(tempty_bar(as), 2 * pr);
        }
      }
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
    if (ep.trace && warp == 4 && lane == 0) {
      ep.trace[blockIdx.x * 8 + 5] = clock64() - tr_t0;  // epilogue warp 4: total
      ep.trace[blockIdx.x * 8 + 6] = tr_tfull;           // epilogue warp 4: waiting for an accumulator
    }
  }

  // ------------------------------ teardown ------------------------------
  tc_fence_before();
  if constexpr (kCta == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kCta>(tmem_base, Cfg::TMEM_COLS);
  }
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------

template <int BN, int kCta, bool kTF32>
static int launch_gemm(const vist3a_gemm_args& a, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, kCta, kTF32>;
  CUtensorMap tmA, tmB, tmA64;
  GemmShape shape = {};
  shape.M = (int)a.M; shape.N = (int)a.N; shape.K = (int)a.K;
  if (a.conv.enabled) {
    const vist3a_conv& c = a.conv;
    shape.conv = 1; shape.kh = c.kh; shape.kw = c.kw; shape.pad_y = c.pad_y; shape.pad_x = c.pad_x;
    shape.n_img = c.n_img; shape.h = c.h; shape.w = c.w; shape.c_in = c.c_in; shape.kt = c.kt > 1 ? c.kt : 1;
    shape.h_out = c.h + 2 * c.pad_y - c.kh + 1;
    shape.w_out = c.w + 2 * c.pad_x - c.kw + 1;
    shape.tiles_x = (shape.w_out + kConvTW - 1) / kConvTW;
    shape.tiles_y = (shape.h_out + kConvTH - 1) / kConvTH;
    shape.cblocks = c.c_in / Cfg::BK;
    const long long mtiles = (long long)c.n_img * shape.tiles_x * shape.tiles_y;
    shape.tiles_m = (int)((mtiles + kCta - 1) / kCta);
    uint64_t dims[4] = {(uint64_t)c.c_in, (uint64_t)c.w, (uint64_t)c.h, (uint64_t)c.n_img};
    const uint64_t pix = c.pix_stride ? (uint64_t)c.pix_stride : (uint64_t)c.c_in;
    const uint64_t row = c.row_stride ? (uint64_t)c.row_stride : (uint64_t)c.w * c.c_in;
    const uint64_t im = c.img_stride ? (uint64_t)c.img_stride : (uint64_t)c.h * c.w * c.c_in;
    uint64_t strides[3] = {pix * Cfg::ES, row * Cfg::ES, im * Cfg::ES};
    uint32_t box[4] = {(uint32_t)Cfg::BK, (uint32_t)kConvTW, (uint32_t)kConvTH, 1};
    int rc = encode_tensor_map(&tmA, a.A, Cfg::ES, kTF32, 4, dims, strides, box, true);
    if (rc) return rc;
  } else {
    shape.tiles_m = (int)((a.M + Cfg::BM * kCta - 1) / (Cfg::BM * kCta));
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
    uint64_t strides[1] = {(uint64_t)a.lda * Cfg::ES};
    uint32_t box[2] = {(uint32_t)Cfg::BK, (uint32_t)Cfg::BM};
    int rc = encode_tensor_map(&tmA, a.A, Cfg::ES, kTF32, 2, dims, strides, box, true);
    if (rc) return rc;
    uint32_t box64[2] = {(uint32_t)Cfg::BK, (uint32_t)Cfg::BM / 2};
    rc = encode_tensor_map(&tmA64, a.A, Cfg::ES, kTF32, 2, dims, strides, box64, true);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    uint64_t strides[1] = {(uint64_t)a.ldw * Cfg::ES};
    uint32_t box[2] = {(uint32_t)Cfg::BK, (uint32_t)Cfg::BN_LOCAL};
    int rc = encode_tensor_map(&tmB, a.W, Cfg::ES, kTF32, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  shape.tiles_n = (int)((a.N + BN - 1) / BN);
  GemmEpilogue ep;
  ep.C = a.C; ep.bias = a.bias; ep.gate = a.gate; ep.residual = a.residual; ep.residual2 = a.residual2;
  ep.ldc = a.ldc; ep.ldr = a.ldr;
  ep.rows_per_batch = a.rows_per_batch > 0 ? a.rows_per_batch : a.M;
  ep.gate_bstride = a.gate_bstride;
  ep.cmap = {a.cmap.rpg, a.cmap.gstride, a.cmap.goff};
  ep.rmap = {a.rmap.rpg, a.rmap.gstride, a.rmap.goff};
  ep.out_fp32 = a.out_dtype == VIST3A_DTYPE_F32;
  ep.act = a.act; ep.post_act = a.post_act; ep.round_linear = a.round_linear; ep.round_gate = a.round_gate;
  ep.trace = g_gemm_trace.load(std::memory_order_relaxed);

  // fp32 output through 256-bit per-thread accesses when the layout allows it (whole 32-column chunks, 32-byte aligned rows)
  const long long al = ep.out_fp32 ? 8 : 16;   // elements per 32 bytes
  const bool direct_ok = BN % 32 == 0 /* (176-wide tiles end in a 16-column chunk) */ && a.N % 32 == 0 && a.ldc % al == 0 && ((uintptr_t)a.C & 31) == 0 &&
                         (!a.residual || (a.ldr % al == 0 && ((uintptr_t)a.residual & 31) == 0)) && !a.residual2;
  // direct by default -- same-process A/B (tools/gemm_epilogue_bench.py): fp32 + gate + residual -12 % on the K <= 1536 projections, bf16 + bf16
  // residual -11..-13 %, plain bf16 -2..-3 % on QKV / FFN1; VIST3A_GEMM_FLAG_STAGED forces the staging-buffer epilogue
  const bool direct = direct_ok && !(a.flags & VIST3A_GEMM_FLAG_STAGED);
  auto kern = direct ? gemm_tcgen05_kernel<BN, kCta, kTF32, true> : gemm_tcgen05_kernel<BN, kCta, kTF32, false>;
  static std::atomic<unsigned long long> attr_done[2];  // per template instantiation (staged / direct), one bit per device
  V3A_CUDA_OK(ensure_dynamic_smem(kern, Cfg::SMEM_BYTES, attr_done[direct ? 1 : 0]));
  if (a.conv.enabled) tmA64 = tmA;
  shape.mc = (kCta == 2 && BN == 256 && !a.conv.enabled && (a.flags & VIST3A_GEMM_FLAG_MULTICAST) && shape.tiles_n % 2 == 0) ? 1 : 0;
  const int csize = shape.mc ? 4 : kCta;
  const int tiles = shape.mc ? shape.tiles_m * (shape.tiles_n / 2) : shape.tiles_m * shape.tiles_n;
  int workers = num_sms() / csize;
  if (shape.mc) {
    // clusters of 4 must sit inside one GPC: fewer than num_sms / 4 may be resident at once, and a persistent kernel must not launch more
    static int max_clusters = 0;
    if (max_clusters == 0) {
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3((unsigned)(workers * 4));
      qc.blockDim = dim3(kGemmThreads);
      qc.dynamicSmemBytes = Cfg::SMEM_BYTES;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 4; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      qc.attrs = qa; qc.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &qc) != cudaSuccess || n <= 0) n = workers;
      max_clusters = n;
    }
    if (workers > max_clusters) workers = max_clusters;
  }
  if (workers > tiles) workers = tiles;
  V3A_CUDA_OK(launch_kernel(kern, dim3((unsigned)(workers * csize)), dim3(kGemmThreads), Cfg::SMEM_BYTES, stream, /*pdl=*/true, csize,
                            tmA, tmB, tmA64, shape, ep));
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

template <bool kTF32>
static int dispatch_gemm(const vist3a_gemm_args& a, cudaStream_t stream) {
  const bool two = (a.flags & VIST3A_GEMM_FLAG_2CTA) && !(a.flags & VIST3A_GEMM_FLAG_1CTA) && a.M > 128;
  if (a.N > 128) {
    if (two && !a.conv.enabled && (a.flags & VIST3A_GEMM_FLAG_BN176)) {
      // Wave quantisation experiment (A/B flag, off by default): cost ~ waves x tile width over the CTA pairs; 176-wide tiles fill 4 waves
      // of 74 pairs to 97 % on a 1536-wide output (256-wide: 3 waves at 86 %).  Measured on B200 (tools/kernel_bench.py, 8192 x 1536):
      // 944 vs 993 TFLOP/s at K = 1536 and -1.3 % at K = 8960 -- a 176-wide tile pulls 77 B/clk/SM of operands through L2 instead of
      // 62.5 and the MMA warp already waits for operands a third of its time (tools/gemm_trace.py), which costs more than the fuller
      // last wave returns.
      const long long pairs = num_sms() / 2, tm = (a.M + 255) / 256;
      auto cost = [&](long long bn) { return ((tm * ((a.N + bn - 1) / bn) + pairs - 1) / pairs) * bn; };
      if (cost(176) * 103 < cost(256) * 100) return launch_gemm<176, 2, kTF32>(a, stream);
    }
    // 192-wide tiles when they cover N with strictly fewer padded columns than 256-wide ones (N = 192, 384, 1152, ...)
    if ((a.N + 191) / 192 * 192 < (a.N + 255) / 256 * 256) return two ? launch_gemm<192, 2, kTF32>(a, stream) : launch_gemm<192, 1, kTF32>(a, stream);
    return two ? launch_gemm<256, 2, kTF32>(a, stream) : launch_gemm<256, 1, kTF32>(a, stream);
  }
  if (a.N > 96) return two ? launch_gemm<128, 2, kTF32>(a, stream) : launch_gemm<128, 1, kTF32>(a, stream);
  if (a.N > 64) return two ? launch_gemm<96, 2, kTF32>(a, stream) : launch_gemm<96, 1, kTF32>(a, stream);
  return launch_gemm<64, 1, kTF32>(a, stream);
}

int gemm_entry(const vist3a_gemm_args* args, cudaStream_t stream) {
  V3A_REQUIRE(args != nullptr, VIST3A_ERR_INVALID, "gemm: null args");
  const vist3a_gemm_args& a = *args;
  V3A_REQUIRE(a.A && a.W && a.C, VIST3A_ERR_INVALID, "gemm: null A/W/C");
  V3A_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, VIST3A_ERR_INVALID, "gemm: M,N,K must be positive (got %lld,%lld,%lld)",
              (long long)a.M, (long long)a.N, (long long)a.K);
  V3A_REQUIRE(a.in_dtype == VIST3A_DTYPE_BF16 || a.in_dtype == VIST3A_DTYPE_F32, VIST3A_ERR_INVALID, "gemm: in_dtype");
  V3A_REQUIRE(a.out_dtype == VIST3A_DTYPE_BF16 || a.out_dtype == VIST3A_DTYPE_F32, VIST3A_ERR_INVALID,
              "gemm: out_dtype");
  const int es = a.in_dtype == VIST3A_DTYPE_F32 ? 4 : 2;
  const int kalign = 16 / es;
  V3A_REQUIRE(a.ldw >= a.K && a.ldw % kalign == 0, VIST3A_ERR_INVALID,
              "gemm: ldw must be >= K and a multiple of %d elements (16-byte TMA stride)", kalign);
  if (a.conv.enabled) {
    const vist3a_conv& c = a.conv;
    const int bk = 128 / es;
    V3A_REQUIRE(c.kh > 0 && c.kw > 0 && c.pad_y >= 0 && c.pad_x >= 0 && c.n_img > 0 && c.h > 0 && c.w > 0 && c.c_in > 0, VIST3A_ERR_INVALID,
                "gemm(conv): bad geometry");
    V3A_REQUIRE(c.pix_stride >= 0 && c.row_stride >= 0 && c.img_stride >= 0 && (c.pix_stride * es) % 16 == 0 && (c.row_stride * es) % 16 == 0 &&
                    (c.img_stride * es) % 16 == 0, VIST3A_ERR_INVALID, "gemm(conv): pixel / row / image strides must be multiples of 16 bytes");
    V3A_REQUIRE(c.c_in % bk == 0, VIST3A_ERR_UNSUPPORTED, "gemm(conv): c_in (%d) must be a multiple of %d", c.c_in, bk);
    V3A_REQUIRE(c.kt >= 0 && c.kt <= 8, VIST3A_ERR_INVALID, "gemm(conv): kt must be in [0, 8]");
    V3A_REQUIRE(a.K == (int64_t)(c.kt > 1 ? c.kt : 1) * c.kh * c.kw * c.c_in, VIST3A_ERR_INVALID, "gemm(conv): K must equal kt*kh*kw*c_in");
    const long long ho = c.h + 2 * c.pad_y - c.kh + 1, wo = c.w + 2 * c.pad_x - c.kw + 1;
    V3A_REQUIRE(ho > 0 && wo > 0 && a.M == (int64_t)c.n_img * ho * wo, VIST3A_ERR_INVALID, "gemm(conv): M must equal n_img*h_out*w_out");
  } else {
    V3A_REQUIRE(a.lda >= a.K && a.lda % kalign == 0, VIST3A_ERR_INVALID,
                "gemm: lda must be >= K and a multiple of %d elements (16-byte TMA stride)", kalign);
  }
  V3A_REQUIRE(((uintptr_t)a.A & 15) == 0 && ((uintptr_t)a.W & 15) == 0 && ((uintptr_t)a.C & 15) == 0,
              VIST3A_ERR_INVALID, "gemm: A/W/C must be 16-byte aligned");
  const int nalign = a.out_dtype == VIST3A_DTYPE_F32 ? 4 : 8;
  V3A_REQUIRE(a.N % nalign == 0 && a.ldc % nalign == 0 && a.ldc >= a.N, VIST3A_ERR_INVALID,
              "gemm: N and ldc must be multiples of %d and ldc >= N", nalign);
  if (a.residual)
    V3A_REQUIRE(a.ldr % nalign == 0 && a.ldr >= a.N && ((uintptr_t)a.residual & 15) == 0, VIST3A_ERR_INVALID,
                "gemm: residual stride/alignment");
  if (a.residual2) V3A_REQUIRE(((uintptr_t)a.residual2 & 15) == 0, VIST3A_ERR_INVALID, "gemm: residual2 alignment");
  if (a.bias) V3A_REQUIRE(((uintptr_t)a.bias & 15) == 0, VIST3A_ERR_INVALID, "gemm: bias alignment");
  if (a.gate) V3A_REQUIRE(((uintptr_t)a.gate & 15) == 0 && a.gate_bstride % 4 == 0, VIST3A_ERR_INVALID, "gemm: gate alignment");
  V3A_REQUIRE(a.cmap.rpg >= 0 && a.rmap.rpg >= 0, VIST3A_ERR_INVALID, "gemm: row maps");
  V3A_REQUIRE(a.post_act == VIST3A_ACT_NONE || a.post_act == VIST3A_ACT_RELU, VIST3A_ERR_UNSUPPORTED, "gemm: post_act must be none or relu");
  V3A_REQUIRE(a.M < (1ll << 31) && a.N < (1ll << 31) && a.K < (1ll << 31), VIST3A_ERR_INVALID, "gemm: dims exceed int32");
  int rc = check_arch();
  if (rc) return rc;
  return a.in_dtype == VIST3A_DTYPE_F32 ? dispatch_gemm<true>(a, stream) : dispatch_gemm<false>(a, stream);
}

}  // namespace v3a
