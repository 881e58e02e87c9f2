// tcgen05 fused attention forward (non-causal, no mask):  O = softmax(Q K^T * scale) V
//
// One CTA per (batch, head, 256-row query block) = TWO 128-row query tiles processed in ping-pong, so the
// tensor pipe works on one tile while the softmax of the other runs (MUFU.EX2 is the co-bottleneck of
// attention on this part: 128x128 exponentials cost as many SM cycles as the 2 MMAs of a d=128 tile).
//   warp 0        TMA producer: Q tiles once, then K and V tiles of 128 keys through mbarrier rings
//   warp 1        MMA issuer:   S_i = Q_i K^T (SS) into TMEM,  O_i += P_i V (TS: P from TMEM, V MN-major smem)
//   warp 2        TMEM allocator
//   warps 4..7    softmax of query tile 0   } thread t owns query row t (TMEM lane t): tcgen05.ld S, online
//   warps 8..11   softmax of query tile 1   } max with lazy rescale of O (only when the running max grows by
//                                             more than 2^8), FFMA2 + MUFU.EX2, bf16 P -> tcgen05.st
// TMEM map (512 columns), d = 128:  S0|P0 [0,128)  S1|P1 [128,256)  O0 [256,384)  O1 [384,512)
//     P_i overwrites the first 64 columns of S_i; the in-order tensor pipe makes QK(j+1) wait for PV(j).
//                         d = 64:   S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512)
//     P has its own columns, so QK_i(j+1) is issued as soon as the softmax warps have pulled S_i(j) into
//     registers and the score tile of step j+1 is ready before the exponentials of step j are done.
// Control warps shrink to 56 registers (setmaxnreg) so that each softmax thread can hold its 128-wide row.
#include "common.cuh"
#include "host_util.cuh"

namespace v3a {

struct FmhaParams {
  void* O;
  long long o_bs, o_rs, o_hs;
  int len_q, len_kv;
  float scale_log2;  // scale * log2(e)
};

template <int D>
struct FmhaCfg {
  static constexpr int BQ = 128, BKV = 128, QT = 2;  // two query tiles per CTA
  static constexpr int SLABS = D / 64;               // 64-element (128-byte) column slabs
  static constexpr int SLAB_BYTES = 128 * 128;       // 128 rows x 128 B
  static constexpr int TILE_BYTES = SLABS * SLAB_BYTES;
  static constexpr bool ALIAS = (D == 128);
  static constexpr int KV_STAGES = (D == 128) ? 2 : 4;
  static constexpr int NBARS = 1 + 4 * KV_STAGES + 8;
  static constexpr int SMEM_BYTES = TILE_BYTES * (QT + 2 * KV_STAGES) + 1024 + 8 * NBARS + 16;
  static constexpr uint32_t TM_S = 0;
  static constexpr uint32_t TM_P = ALIAS ? 0 : 256;
  static constexpr uint32_t P_STRIDE = ALIAS ? 128 : 64;
  static constexpr uint32_t TM_O = ALIAS ? 256 : 384;
};

constexpr int kFmhaThreads = 384;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// (y0, y1) = (x0, x1) * (c, c) + (b, b) as one packed FFMA2
__device__ __forceinline__ void fma2(float x0, float x1, uint64_t cc, uint64_t bb, float& y0, float& y1) {
  uint64_t xx, yy;
  asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x0), "f"(x1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(yy) : "l"(xx), "l"(cc), "l"(bb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(y0), "=f"(y1) : "l"(yy));
}
__device__ __forceinline__ uint64_t splat2(float v) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
  return r;
}

template <int D>
__global__ void __launch_bounds__(kFmhaThreads, 1)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const FmhaParams p) {
  using Cfg = FmhaCfg<D>;
  constexpr int ST = Cfg::KV_STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto smem_q = [&](int i) { return smem_base + Cfg::TILE_BYTES * i; };
  auto smem_k = [&](int s) { return smem_base + Cfg::TILE_BYTES * (Cfg::QT + s); };
  auto smem_v = [&](int s) { return smem_base + Cfg::TILE_BYTES * (Cfg::QT + ST + s); };
  const uint32_t bar_base = smem_base + Cfg::TILE_BYTES * (Cfg::QT + 2 * ST);
  const uint32_t q_full = bar_base;
  auto k_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bar_base + 8u * (1 + ST + s); };
  auto v_full = [&](int s) { return bar_base + 8u * (1 + 2 * ST + s); };
  auto v_empty = [&](int s) { return bar_base + 8u * (1 + 3 * ST + s); };
  auto s_full = [&](int i) { return bar_base + 8u * (1 + 4 * ST + i); };
  auto s_free = [&](int i) { return bar_base + 8u * (3 + 4 * ST + i); };
  auto p_full = [&](int i) { return bar_base + 8u * (5 + 4 * ST + i); };
  auto pv_done = [&](int i) { return bar_base + 8u * (7 + 4 * ST + i); };
  const uint32_t tmem_slot = bar_base + 8u * Cfg::NBARS;

  const uint32_t warp = warp_id_sync();
  const uint32_t lane = lane_id();
  const int q0 = blockIdx.x * (Cfg::BQ * Cfg::QT);
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_kv = (p.len_kv + Cfg::BKV - 1) / Cfg::BKV;
  const int nq = (q0 + Cfg::BQ < p.len_q) ? 2 : 1;  // query tiles of this CTA that hold at least one row

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full(i), 1);
      mbar_init(s_free(i), 4);  // one arrival per softmax warp of the tile
      mbar_init(p_full(i), 4);
      mbar_init(pv_done(i), 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0 && lane == 0) {
      // ------------------------------ TMA producer ------------------------------
      mbar_expect_tx(q_full, (uint32_t)(nq * Cfg::TILE_BYTES));
      for (int i = 0; i < nq; ++i) {
#pragma unroll
        for (int sl = 0; sl < Cfg::SLABS; ++sl)
          tma_load_4d(smem_q(i) + sl * Cfg::SLAB_BYTES, &tmQ, q_full, sl * 64, head, q0 + i * Cfg::BQ, batch);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(k_empty(s), ph ^ 1u);
        mbar_expect_tx(k_full(s), Cfg::TILE_BYTES);
#pragma unroll
        for (int sl = 0; sl < Cfg::SLABS; ++sl)
          tma_load_4d(smem_k(s) + sl * Cfg::SLAB_BYTES, &tmK, k_full(s), sl * 64, head, j * Cfg::BKV, batch);
        mbar_wait(v_empty(s), ph ^ 1u);
        mbar_expect_tx(v_full(s), Cfg::TILE_BYTES);
#pragma unroll
        for (int sl = 0; sl < Cfg::SLABS; ++sl)
          tma_load_4d(smem_v(s) + sl * Cfg::SLAB_BYTES, &tmV, v_full(s), sl * 64, head, j * Cfg::BKV, batch);
        if (++s == ST) { s = 0; ph ^= 1u; }
      }
    } else if (warp == 1 && lane == 0) {
      // ------------------------------ MMA issuer ------------------------------
      constexpr uint32_t idesc_qk = make_idesc(kFmtBF16, 128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(kFmtBF16, 128, D, 0, 1);  // B (V) is MN-major
      auto issue_qk = [&](int i, int j) {
        const int s = j % ST;
        mbar_wait(k_full(s), (uint32_t)(j / ST) & 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + Cfg::TM_S + (uint32_t)(i * 128);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (uint32_t)((kk >> 2) * Cfg::SLAB_BYTES + (kk & 3) * 32);
          const uint64_t adesc = make_smem_desc_sw128(smem_q(i) + off, 1024, 0);
          const uint64_t bdesc = make_smem_desc_sw128(smem_k(s) + off, 1024, 0);
          umma_f16_ss<1>(d_tmem, adesc, bdesc, idesc_qk, kk ? 1u : 0u);
        }
        umma_commit(s_full(i));
        if (i == nq - 1) umma_commit(k_empty(s));
      };
      auto issue_pv = [&](int i, int j) {
        const int s = j % ST;
        mbar_wait(v_full(s), (uint32_t)(j / ST) & 1u);
        tc_fence_after();
        const uint32_t o_tmem = tmem_base + Cfg::TM_O + (uint32_t)(i * D);
        const uint32_t p_tmem = tmem_base + Cfg::TM_P + (uint32_t)(i * Cfg::P_STRIDE);
#pragma unroll
        for (int kk = 0; kk < Cfg::BKV / 16; ++kk) {
          // V tile: kv rows at a 128 B pitch (K dimension), 64-wide head-dim slabs LBO apart (MN dimension)
          const uint64_t bdesc = make_smem_desc_sw128(smem_v(s) + kk * 2048, 1024, Cfg::SLAB_BYTES);
          umma_f16_ts(o_tmem, p_tmem + (uint32_t)(kk * 8), bdesc, idesc_pv, (j | kk) ? 1u : 0u);
        }
        umma_commit(pv_done(i));
        if (i == nq - 1) umma_commit(v_empty(s));
      };
      mbar_wait(q_full, 0);
      for (int i = 0; i < nq; ++i) issue_qk(i, 0);
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        const bool more = j + 1 < n_kv;
        for (int i = 0; i < nq; ++i) {
          if (!Cfg::ALIAS && more) {
            mbar_wait(s_free(i), ph);  // softmax warps hold S_i(j) in registers
            tc_fence_after();
            issue_qk(i, j + 1);
          }
          mbar_wait(p_full(i), ph);
          tc_fence_after();
          issue_pv(i, j);
          if (Cfg::ALIAS && more) issue_qk(i, j + 1);  // overwrites S_i|P_i: ordered behind P_i V by the in-order pipe
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ------------------------------ softmax / correction / epilogue ------------------------------
    const int i = (int)(warp >> 2) - 1;  // query tile of this warpgroup
    if (i < nq) {
      const uint32_t wq = warp & 3u;     // TMEM lane quadrant this warp may access
      const uint32_t lane_sel = (wq * 32u) << 16;
      const int row = q0 + i * Cfg::BQ + (int)(wq * 32u + lane);
      const uint32_t s_addr = tmem_base + lane_sel + Cfg::TM_S + (uint32_t)(i * 128);
      const uint32_t p_addr = tmem_base + lane_sel + Cfg::TM_P + (uint32_t)(i * Cfg::P_STRIDE);
      const uint32_t o_addr = tmem_base + lane_sel + Cfg::TM_O + (uint32_t)(i * D);
      float m_run = -INFINITY;  // running (possibly stale) row max of raw scores
      float l_run = 0.0f;       // running row sum of exp2((s - m_run) * scale_log2)
      const float c = p.scale_log2;
      const uint64_t cc2 = splat2(c);
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        mbar_wait(s_full(i), ph);
        tc_fence_after();
        uint32_t r[128];
        tmem_ld_x32(s_addr, r);
        tmem_ld_x32(s_addr + 32, r + 32);
        tmem_ld_x32(s_addr + 64, r + 64);
        tmem_ld_x32(s_addr + 96, r + 96);
        tmem_ld_wait();
        if (!Cfg::ALIAS) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free(i));
        }
        const int valid = p.len_kv - j * Cfg::BKV;  // keys of this tile that exist
        if (valid < Cfg::BKV) {
#pragma unroll
          for (int k = 0; k < 128; ++k)
            if (k >= valid) r[k] = 0xff800000u;  // -inf
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int k = 0; k < 128; k += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            mx[u] = fmaxf(fmaxf(mx[u], __uint_as_float(r[k + 2 * u])), __uint_as_float(r[k + 2 * u + 1]));
        }
        const float m_new = fmaxf(m_run, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])));
        if (j == 0) {
          m_run = m_new;
        } else {
          if (!Cfg::ALIAS) {
            // P_i is read by P_i(j-1) V and O_i is accumulated by it: both must be finished before we touch them.
            // (d = 128: the commit behind s_full(i) of this step already covers that MMA.)
            mbar_wait(pv_done(i), ph ^ 1u);
            tc_fence_after();
          }
          const bool need = (m_new - m_run) * c > 8.0f;
          if (__any_sync(0xffffffffu, need)) {
            const float f = need ? ex2_approx((m_run - m_new) * c) : 1.0f;
            if (need) m_run = m_new;
            l_run *= f;
#pragma unroll
            for (int cb = 0; cb < D / 32; ++cb) {
              uint32_t o[32];
              tmem_ld_x32(o_addr + (uint32_t)(cb * 32), o);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * f);
              tmem_st_x32(o_addr + (uint32_t)(cb * 32), o);
            }
            tmem_st_wait();
          }
        }
        const uint64_t mc2 = splat2(-m_run * c);
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          uint32_t pk[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float y0, y1;
            fma2(__uint_as_float(r[cb * 32 + 2 * k]), __uint_as_float(r[cb * 32 + 2 * k + 1]), cc2, mc2, y0, y1);
            const float e0 = ex2_approx(y0), e1 = ex2_approx(y1);
            sum[(2 * k) & 3] += e0;
            sum[(2 * k + 1) & 3] += e1;
            pk[k] = pack_bf16(e0, e1);
          }
          tmem_st_x16(p_addr + (uint32_t)(cb * 16), pk);
        }
        l_run += (sum[0] + sum[1]) + (sum[2] + sum[3]);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(i));
      }
      // ---- epilogue: O / l -> bf16 -> global ----
      mbar_wait(pv_done(i), (uint32_t)(n_kv - 1) & 1u);
      tc_fence_after();
      const float inv_l = 1.0f / l_run;
      __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.O) + (long long)batch * p.o_bs + (long long)row * p.o_rs +
                            (long long)head * p.o_hs;
#pragma unroll
      for (int cb = 0; cb < D / 32; ++cb) {
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
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}


static int make_qkv_map(CUtensorMap* tm, const void* ptr, long long B, long long H, long long L, long long D,
                        long long bs, long long rs, long long hs) {
  // dims (fastest first): head_dim, heads, rows, batch
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)H, (uint64_t)L, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)hs * 2, (uint64_t)rs * 2, (uint64_t)bs * 2};
  uint32_t box[4] = {64, 1, 128, 1};
  return encode_tensor_map(tm, ptr, 2, false, 4, dims, strides, box, true);
}

template <int D>
static int launch_fmha(const vist3a_fmha_args& a, cudaStream_t stream) {
  using Cfg = FmhaCfg<D>;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_qkv_map(&tmQ, a.Q, a.batch, a.heads, a.len_q, D, a.q_bs, a.q_rs, a.q_hs))) return rc;
  if ((rc = make_qkv_map(&tmK, a.K, a.batch, a.heads, a.len_kv, D, a.k_bs, a.k_rs, a.k_hs))) return rc;
  if ((rc = make_qkv_map(&tmV, a.V, a.batch, a.heads, a.len_kv, D, a.v_bs, a.v_rs, a.v_hs))) return rc;
  FmhaParams p;
  p.O = a.O; p.o_bs = a.o_bs; p.o_rs = a.o_rs; p.o_hs = a.o_hs;
  p.len_q = (int)a.len_q; p.len_kv = (int)a.len_kv;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  auto kern = fmha_fwd_kernel<D>;
  static bool attr_set = false;
  if (!attr_set) {
    V3A_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const long long rows_per_cta = Cfg::BQ * Cfg::QT;
  dim3 grid((unsigned)((a.len_q + rows_per_cta - 1) / rows_per_cta), (unsigned)a.heads, (unsigned)a.batch);
  kern<<<grid, kFmhaThreads, Cfg::SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  V3A_CUDA_OK(cudaGetLastError());
  launch_counter().fetch_add(1);
  return VIST3A_OK;
}

int fmha_entry(const vist3a_fmha_args* args, cudaStream_t stream) {
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
  return a.head_dim == 128 ? launch_fmha<128>(a, stream) : launch_fmha<64>(a, stream);
}

}  // namespace v3a
