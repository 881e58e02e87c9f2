// tcgen05 fused attention forward (non-causal, no mask):  O = softmax(Q K^T * scale) V
//
// One CTA per (batch, head, 128-row query tile).  Warp roles:
//   warp 0      TMA producer: Q once, then K/V tiles of 128 keys through 2-deep mbarrier rings
//   warp 1      MMA issuer:   S = Q K^T (SS, both K-major) into a double-buffered TMEM score tile,
//                             O += P V (TS: P from TMEM, V MN-major from smem) into a TMEM accumulator
//   warp 2      TMEM allocator
//   warps 4..7  softmax: thread t owns query row t (TMEM lane t): tcgen05.ld S, online max with lazy
//               rescale of O (only when the running max grows by more than 2^8), exp2, bf16 P -> tcgen05.st
// QK^T of tile j+1 is issued before P V of tile j so the tensor pipe runs under the softmax of tile j.
// TMEM map (512 columns): S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O [384, 384+D).
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
  static constexpr int BQ = 128, BKV = 128;
  static constexpr int SLABS = D / 64;              // 64-element (128-byte) column slabs
  static constexpr int SLAB_BYTES = 128 * 128;      // 128 rows x 128 B
  static constexpr int TILE_BYTES = SLABS * SLAB_BYTES;
  static constexpr int KV_STAGES = 2;
  static constexpr int SMEM_BYTES = TILE_BYTES * (1 + 2 * KV_STAGES) + 1024 + 256;
  static constexpr uint32_t TM_S = 0, TM_P = 256, TM_O = 384;
};

template <int D>
__global__ void __launch_bounds__(256, 1)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const FmhaParams p) {
  using Cfg = FmhaCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_q = smem_base;
  auto smem_k = [&](int s) { return smem_base + Cfg::TILE_BYTES * (1 + s); };
  auto smem_v = [&](int s) { return smem_base + Cfg::TILE_BYTES * (1 + Cfg::KV_STAGES + s); };
  const uint32_t bar_base = smem_base + Cfg::TILE_BYTES * (1 + 2 * Cfg::KV_STAGES);
  const uint32_t q_full = bar_base;
  auto k_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bar_base + 8u * (3 + s); };
  auto v_full = [&](int s) { return bar_base + 8u * (5 + s); };
  auto v_empty = [&](int s) { return bar_base + 8u * (7 + s); };
  auto s_full = [&](int s) { return bar_base + 8u * (9 + s); };
  auto p_full = [&](int s) { return bar_base + 8u * (11 + s); };
  auto pv_done = [&](int s) { return bar_base + 8u * (13 + s); };
  const uint32_t tmem_slot = bar_base + 8u * 15;

  const uint32_t warp = warp_id_sync();
  const uint32_t lane = lane_id();
  const int q0 = blockIdx.x * Cfg::BQ;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_kv = (p.len_kv + Cfg::BKV - 1) / Cfg::BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
      mbar_init(s_full(s), 1);
      mbar_init(p_full(s), 4);  // one arrival per softmax warp
      mbar_init(pv_done(s), 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      mbar_expect_tx(q_full, Cfg::TILE_BYTES);
#pragma unroll
      for (int sl = 0; sl < Cfg::SLABS; ++sl)
        tma_load_4d(smem_q + sl * Cfg::SLAB_BYTES, &tmQ, q_full, sl * 64, head, q0, batch);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1u;
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
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc(kFmtBF16, 128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(kFmtBF16, 128, D, 0, 1);  // B (V) is MN-major
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        mbar_wait(k_full(s), (uint32_t)(j >> 1) & 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + Cfg::TM_S + (uint32_t)(s * 128);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (uint32_t)((kk >> 2) * Cfg::SLAB_BYTES + (kk & 3) * 32);
          const uint64_t adesc = make_smem_desc_sw128(smem_q + off, 1024, 0);
          const uint64_t bdesc = make_smem_desc_sw128(smem_k(s) + off, 1024, 0);
          umma_f16_ss<1>(d_tmem, adesc, bdesc, idesc_qk, kk ? 1u : 0u);
        }
        umma_commit(k_empty(s));
        umma_commit(s_full(s));
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1u;
        if (j + 1 < n_kv) issue_qk(j + 1);
        mbar_wait(p_full(s), ph);
        mbar_wait(v_full(s), ph);
        tc_fence_after();
        const uint32_t o_tmem = tmem_base + Cfg::TM_O;
        const uint32_t p_tmem = tmem_base + Cfg::TM_P + (uint32_t)(s * 64);
#pragma unroll
        for (int kk = 0; kk < Cfg::BKV / 16; ++kk) {
          // V tile: kv rows at a 128 B pitch (K dimension), 64-wide head-dim slabs LBO apart (MN dimension)
          const uint64_t bdesc = make_smem_desc_sw128(smem_v(s) + kk * 2048, 1024, Cfg::SLAB_BYTES);
          umma_f16_ts(o_tmem, p_tmem + (uint32_t)(kk * 8), bdesc, idesc_pv, (j | kk) ? 1u : 0u);
        }
        umma_commit(v_empty(s));
        umma_commit(pv_done(s));
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ softmax / correction / epilogue ------------------------------
    const uint32_t q = warp & 3u;
    const uint32_t lane_sel = (q * 32u) << 16;
    const int row = q0 + (int)(q * 32u + lane);
    float m_run = -INFINITY;  // running (possibly stale) row max of raw scores
    float l_run = 0.0f;       // running row sum of exp2((s - m_run) * scale_log2)
    const float c = p.scale_log2;
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      mbar_wait(s_full(s), (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      uint32_t r[128];
      {
        const uint32_t sa = tmem_base + lane_sel + Cfg::TM_S + (uint32_t)(s * 128);
        tmem_ld_x32(sa, r);
        tmem_ld_x32(sa + 32, r + 32);
        tmem_ld_x32(sa + 64, r + 64);
        tmem_ld_x32(sa + 96, r + 96);
        tmem_ld_wait();
      }
      const int valid = p.len_kv - j * Cfg::BKV;  // keys of this tile that exist
      if (valid < Cfg::BKV) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= valid) r[i] = 0xff800000u;  // -inf
      }
      float m_tile = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; ++i) m_tile = fmaxf(m_tile, __uint_as_float(r[i]));
      const float m_new = fmaxf(m_run, m_tile);
      if (j == 0) {
        m_run = m_new;
      } else {
        const bool need = (m_new - m_run) * c > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // O is being accumulated by P V of tile j-1: wait for it, then rescale this row
          mbar_wait(pv_done((j - 1) & 1), (uint32_t)((j - 1) >> 1) & 1u);
          tc_fence_after();
          const float f = need ? exp2f((m_run - m_new) * c) : 1.0f;
          if (need) m_run = m_new;
          l_run *= f;
#pragma unroll
          for (int cc = 0; cc < D / 32; ++cc) {
            uint32_t o[32];
            const uint32_t oa = tmem_base + lane_sel + Cfg::TM_O + (uint32_t)(cc * 32);
            tmem_ld_x32(oa, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_x32(oa, o);
          }
          tmem_st_wait();
        }
      }
      if (j >= 2) mbar_wait(pv_done(s), (uint32_t)((j - 2) >> 1) & 1u);  // P buffer s is free again
      const float mc = m_run * c;
      float sum = 0.0f;
      const uint32_t pa = tmem_base + lane_sel + Cfg::TM_P + (uint32_t)(s * 64);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float e0 = exp2f(fmaf(__uint_as_float(r[cc * 32 + 2 * i]), c, -mc));
          const float e1 = exp2f(fmaf(__uint_as_float(r[cc * 32 + 2 * i + 1]), c, -mc));
          sum += e0 + e1;
          pk[i] = pack_bf16(e0, e1);
        }
        tmem_st_x16(pa + (uint32_t)(cc * 16), pk);
      }
      l_run += sum;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(s));
    }
    // ---- epilogue: O / l -> bf16 -> global ----
    mbar_wait(pv_done((n_kv - 1) & 1), (uint32_t)((n_kv - 1) >> 1) & 1u);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.O) + (long long)batch * p.o_bs + (long long)row * p.o_rs +
                          (long long)head * p.o_hs;
#pragma unroll
    for (int cc = 0; cc < D / 32; ++cc) {
      uint32_t o[32];
      tmem_ld_x32(tmem_base + lane_sel + Cfg::TM_O + (uint32_t)(cc * 32), o);
      tmem_ld_wait();
      if (row < p.len_q) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
          w.y = pack_bf16(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
          w.z = pack_bf16(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
          w.w = pack_bf16(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + cc * 32 + i) = w;
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
  dim3 grid((unsigned)((a.len_q + Cfg::BQ - 1) / Cfg::BQ), (unsigned)a.heads, (unsigned)a.batch);
  kern<<<grid, 256, Cfg::SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
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
