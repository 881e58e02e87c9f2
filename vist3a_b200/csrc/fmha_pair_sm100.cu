// tcgen05 fused attention forward on CTA PAIRS (cta_group::2), head_dim 128:  O = softmax(Q K^T * scale) V, non-causal, no mask.
//
// Why pairs.  The single-CTA kernel (fmha_sm100.cu) issues S = Q K^T as SS-mode MMAs of 128 x 64 x 16: every MMA re-reads its 4 KB Q slice
// and a 2 KB K slice from shared memory = 192 B/clk against the 128 B/clk an SM's shared memory delivers, so the QK half of the tensor work
// runs at 2/3 speed (traced floor: 1280 cycles per 1024 cycles of tensor work).  With cta_group::2 one MMA covers 256 query rows x 128 keys:
// each CTA supplies its OWN 128 query rows (4 KB per MMA) and only HALF of the key tile (64 keys, 2 KB) for 64 cycles of tensor time
// = 96 B/clk, and P V (A = P from tensor memory, B = half of V's head-dim columns: 2 KB per MMA) needs 32 B/clk.  TMA traffic halves too
// (each CTA loads half of every K and V tile: 32 KB per 128-key step).
//
// Work decomposition (FmhaPairParams, plan_pair): PERSISTENT clusters of two CTAs, one per SM pair, each walking its own list of items
// (built in shared memory at start-up).  An item is a unit = (batch, head, 256 QT query rows) over all keys, or -- for the units of the grid's
// last, partial wave -- a range of its key steps: those units are laid end to end and cut into one equal range per cluster; partial pieces
// go to a workspace and fmha_pair_combine_kernel merges them.  Within a tile of the pair CTA r owns query rows [128 r, 128 r + 128).
// Barrier parities follow a step counter that runs across the items of a cluster.  Between items: the output leaves through its own 32 KB
// staging buffer (OSTAGE; the two tiles take turns), so the producer reloads the Q buffers as soon as the item's last Q K^T has completed
// and the MMA warp issues the next item's first Q K^T right behind the last P V.
//   warp 0        TMA producer (both CTAs): per item its own Q tiles; per 128-key step the CTA's half of K (keys [64 r, 64 r + 64), all of d) and its half
//                 of V (all 128 keys, head-dim columns [64 r, 64 r + 64)); "full" barriers live in the leader CTA (2-SM TMA completion)
//   warp 1        MMA issuer, leader CTA only: S(j) = Q K(j)^T into TMEM score buffer j % 2, O += P(j) V(j) (P read from TMEM), for both CTAs;
//                 completion is multicast to both CTAs' barriers (tcgen05.commit ... multicast::cluster)
//   warp 2        TMEM allocator (cta_group::2, 512 columns)
//   warp 3        output store (OSTAGE): waits for a tile's softmax threads to stage their rows, issues the bulk tensor store (to O, or to
//                 the workspace for a partial piece) and releases the staging buffer to the other tile
//   warps 4..     softmax, SPLIT threads per query row (thread = TMEM lane x column slice): tcgen05.ld S, row max (slices exchange through
//                 shared memory + a named barrier), lazy rescale of O, exp2, bf16 P -> tcgen05.st over the S buffer it came from, arrive on
//                 the LEADER's p_full barrier (remote arrive from the peer CTA)
// Two shapes of the same kernel (template QT = query tiles per CTA):
//   QT = 1   one 128-row tile per CTA, S double-buffered.  TMEM: S0 | P0 [0,128)  S1 | P1 [128,256)  O [256,384).  Q K(j+2)^T is queued behind
//            P(j) V(j) in the in-order tensor pipe (it overwrites the buffer P(j) lives in).  All softmax warps of the CTA work on the same score
//            tile in lockstep, so the MUFU idles while they load / reduce / store (measured: 36 % tensor pipe) -- kept for A/B.
//   QT = 2   two 128-row tiles per CTA (512 query rows per cluster), each with its own softmax warps, S single-buffered per tile
//            (TMEM per tile: S | P [0,128)  O [128,256)): while tile A's warps exponentiate, the tensor pipe runs tile B's P V and next Q K^T
//            and tile B's warps are in their load / max / store phases -- the two tiles keep MUFU and tensor pipe busy alternately, and every
//            K / V tile is used for 512 query rows.
#include "common.cuh"
#include "fmha_math.cuh"
#include "host_util.cuh"
#include <algorithm>
#include <vector>

namespace v3a {

struct FmhaPairParams {
  int len_q, len_kv;
  float scale_log2;        // scale * log2(e)
  const float* row_scale;  // optional per-(batch, query row) positive factor on the logits
  long long* trace;        // debug (v3a_debug_fmha_pair_trace): clock64 stamps of CTA (0,0,0), normally null
  uint32_t zero;           // 0 (a value ptxas cannot fold: scheduling aid of the speculative softmax)
  // Work decomposition (1-D grid of G persistent clusters, cluster c = blockIdx.x / 2).  A unit = (batch, head, block of 256 QT query rows),
  // q-block fastest.  Units [0, n_full) -- whole waves of the grid -- go round-robin: cluster c takes c, c + G, ... over all keys.  The
  // tail_units behind them, which would leave most SMs idle for a whole pass over the keys, are laid end to end as tail_units * n_kv key
  // steps and cut into G equal ranges: cluster c takes steps [c T / G, (c + 1) T / G), i.e. the end of one unit and / or the beginning of the
  // next.  A partial piece writes its normalised O (bf16) and per row (reference maximum * c, row sum) to workspace slot 2 c (the cluster's
  // first piece) or 2 c + 1 (its last) and fmha_pair_combine_kernel merges the pieces of every unit.
  int q_blocks, heads;
  int n_full, tail_units;
  float* ws_ml;            // [slot][256 QT rows][2] fp32 (behind the partial O tiles the workspace tensor map tmW addresses)
};

// debug hook (tools/fmha_pair_trace.py), off unless armed: [step][tile][8] stamps of the leader CTA of cluster 0:
//   0 MMA warp sees P ready   1 MMA warp has issued P V + next Q K^T   2 softmax sees S ready   3 S in registers   4 row max exchanged
//   5 exponentials done       6 P stored + arrived
static std::atomic<long long*> g_pair_trace{nullptr};
extern "C" void v3a_debug_fmha_pair_trace(void* buf) { g_pair_trace.store(reinterpret_cast<long long*>(buf)); }
#define PAIR_TRACE(j, i, slot)                                                                         \
  do {                                                                                                 \
    if (p.trace && blockIdx.x == 0 && item == 0 && (j) < 64) p.trace[((j) * 2 + (i)) * 8 + (slot)] = clock64(); \
  } while (0)

// second part of the debug buffer, [item number in the cluster][cluster][8] at offset 1024 (512 rows): %globaltimer (ns) in the leader CTA of the cluster that works on the item:
//   0 producer turns to the item   1 Q requested (the previous item's output has left the Q buffers)   2 tile 0 sees its first S
//   3 tile 0 has handed over its last P   4 last P V complete   5 tile 0's output stored   6 kernel entry of the cluster   7 SM id | key steps << 16
#define PAIR_STAMP(slot)                                                                                    \
  do {                                                                                                      \
    if (p.trace && rank == 0 && item * (int)(gridDim.x >> 1) + (int)(blockIdx.x >> 1) < 512)             \
      p.trace[1024 + (long long)(item * (int)(gridDim.x >> 1) + (int)(blockIdx.x >> 1)) * 8 + (slot)] = (long long)globaltimer_ns(); \
  } while (0)

// ---- the key split of the tail: index arithmetic shared by the attention kernel (its item list), the merge kernel and the host-side
//      self-check (v3a_debug_fmha_pair_plan_check, tests/test_abi_cpu.py) ------------------------------------------------------------------
// pieces of the tail range [lo, hi) of cluster c, in order: f(unit in the tail, first key step, key steps, workspace slot or -1 for a whole unit)
template <class F>
__host__ __device__ inline void pair_tail_pieces(int c, int lo, int hi, int n_kv_all, F&& f) {
  for (int pos = lo; pos < hi;) {
    const int u = pos / n_kv_all, end = hi < (u + 1) * n_kv_all ? hi : (u + 1) * n_kv_all;
    f(u, pos - u * n_kv_all, end - pos, end - pos == n_kv_all ? -1 : 2 * c + (pos == lo ? 0 : 1));
    pos = end;
  }
}
// merge side: the next piece of the unit with tail steps [u_lo, u_hi), walking the clusters from cc on (lo[] = the range boundaries); returns its
// workspace slot, or -1 when the unit has no further piece.  A cluster's piece is its first item (slot 2 cc) when its range starts inside
// the unit, else its last (slot 2 cc + 1).
__host__ __device__ inline int pair_next_piece_slot(const int* lo, int clusters, int u_lo, int u_hi, int& cc) {
  while (cc < clusters && lo[cc] < u_hi) {
    const int l = lo[cc], h = lo[cc + 1];
    const int at = cc++;
    if (h > (l > u_lo ? l : u_lo)) return 2 * at + (l >= u_lo ? 0 : 1);
  }
  return -1;
}
// the cut units: a boundary on a unit's edge cuts nothing; of several boundaries inside one unit the first one stands for it
__host__ inline int pair_cut_list(const int* lo, int clusters, int n_kv_all, int* cut) {
  int n = 0;
  for (int cc = 1; cc < clusters; ++cc) {
    const int u_lo = lo[cc] / n_kv_all * n_kv_all;
    if (lo[cc] != u_lo && lo[cc - 1] <= u_lo) cut[n++] = cc;
  }
  return n;
}

template <int QT_, int SPLIT_, int POLY_, int FAST_ = 0>
struct FmhaPairCfg {
  static constexpr int D = 128, BQ = 128, BKV = 128;
  static constexpr int QT = QT_;                       // 128-row query tiles per CTA
  static constexpr int NSB = QT == 1 ? 2 : 1;          // score buffers per tile
  static constexpr int SPLIT = SPLIT_;                 // softmax threads per query row
  static constexpr int POLY = POLY_;                   // of every 8 column pairs, this many take exp2 on the FMA pipe
  static constexpr int HC = BKV / SPLIT;               // score columns per softmax thread and step
  static constexpr int OC = D / SPLIT;                 // O columns per softmax thread (rescale, epilogue)
  static constexpr int THREADS = 128 + QT * 128 * SPLIT;
  static constexpr int Q_SLAB_BYTES = BQ * 128;        // 64 head-dim columns (128 B) x 128 rows
  static constexpr int Q_TILE_BYTES = 2 * Q_SLAB_BYTES;
  static constexpr int K_SLAB_BYTES = (BKV / 2) * 128; // this CTA's 64 keys x 64 head-dim columns
  static constexpr int K_HALF_BYTES = 2 * K_SLAB_BYTES;  // 16 KB
  static constexpr int V_HALF_BYTES = BKV * 128;       // 128 keys x this CTA's 64 head-dim columns = 16 KB
  static constexpr int ST = 4;                         // ring stages of K and of V
  static constexpr int NH = SPLIT == 1 ? 2 : 1;        // P(j) is handed to the MMA warp in NH key halves (one thread per row: after 64 keys each)
  // OSTAGE: the output leaves through its own 32 KB staging buffer (shared by the CTA's two tiles, which finish half a step apart) instead of
  // the Q buffers, so that the next item's Q is loaded and its first Q K^T issued while the epilogue of the current item runs
  static constexpr bool OSTAGE = QT_ == 2 && SPLIT_ == 1;
  static constexpr int NBARS = 1 + 4 * ST + 8 * QT + 2 * QT + 1 + 2;   // q_full | k_full, k_empty, v_full, v_empty | per tile: s_full[2], p_full[2][2], pv_done[2] | pvh_done[2] | q_free | ostage_done[2]
  static constexpr bool FAST = FAST_ != 0;             // speculative (stale-maximum) softmax in 64-column half-steps (fmha_math.cuh), one thread per row
  static_assert(!FAST || (SPLIT_ == 1 && QT_ == 2), "the speculative softmax runs one thread per row on two tiles per CTA");
  static constexpr int XCH_BYTES = SPLIT == 1 ? 0 : 2 * QT * SPLIT * 128 * 4;  // [step parity][tile][slice][row] fp32 (slices of a row exchange max / sum)
  static constexpr int OSTAGE_BYTES = OSTAGE ? Q_TILE_BYTES : 0;
  static constexpr int MAX_ITEMS = 60;                 // items of one cluster (its list lives in shared memory)
  static constexpr int SMEM_BYTES = QT * Q_TILE_BYTES + ST * (K_HALF_BYTES + V_HALF_BYTES) + OSTAGE_BYTES + 1024 + 8 * NBARS + 16 + XCH_BYTES + 16 * (MAX_ITEMS + 1);
  static constexpr uint32_t TILE_COLS = 256, TM_S = 0, S_STRIDE = 128, TM_O = NSB * 128;
  static_assert(SPLIT == 1 || SPLIT == 2 || SPLIT == 4, "SPLIT");
  static_assert(QT == 1 || (QT == 2 && SPLIT <= 2), "two tiles per CTA run one or two threads per row (384 / 640 threads)");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <int QT_, int SPLIT_, int POLY_, int FAST_ = 0>
__global__ void __launch_bounds__(128 + QT_ * 128 * SPLIT_, 1)
fmha_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmW, const FmhaPairParams p) {
  using Cfg = FmhaPairCfg<QT_, SPLIT_, POLY_, FAST_>;
  constexpr int QT = Cfg::QT, NSB = Cfg::NSB, NH = Cfg::NH, SPLIT = Cfg::SPLIT, HC = Cfg::HC, OC = Cfg::OC, ST = Cfg::ST, BKV = Cfg::BKV, D = Cfg::D;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto smem_q = [&](int i) { return smem_base + Cfg::Q_TILE_BYTES * i; };
  auto smem_k = [&](int s) { return smem_base + QT * Cfg::Q_TILE_BYTES + Cfg::K_HALF_BYTES * s; };
  auto smem_v = [&](int s) { return smem_base + QT * Cfg::Q_TILE_BYTES + Cfg::K_HALF_BYTES * ST + Cfg::V_HALF_BYTES * s; };
  const uint32_t smem_ostage = smem_base + QT * Cfg::Q_TILE_BYTES + ST * (Cfg::K_HALF_BYTES + Cfg::V_HALF_BYTES);
  const uint32_t bar_base = smem_ostage + Cfg::OSTAGE_BYTES;
  const uint32_t q_full = bar_base;
  auto k_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bar_base + 8u * (1 + ST + s); };
  auto v_full = [&](int s) { return bar_base + 8u * (1 + 2 * ST + s); };
  auto v_empty = [&](int s) { return bar_base + 8u * (1 + 3 * ST + s); };
  // per tile i; every barrier exists twice (b = step parity): a waiter is never two completions behind the barrier it polls
  auto s_full = [&](int i, int b) { return bar_base + 8u * (1 + 4 * ST + 8 * i + b); };
  auto p_full = [&](int i, int b, int hh) { return bar_base + 8u * (1 + 4 * ST + 8 * i + 2 + 2 * b + hh); };
  auto pv_done = [&](int i, int b) { return bar_base + 8u * (1 + 4 * ST + 8 * i + 6 + b); };
  auto pvh_done = [&](int i, int b) { return bar_base + 8u * (1 + 4 * ST + 8 * QT + 2 * i + b); };   // first key half of P_i(j) V(j) has completed
  const uint32_t q_free = bar_base + 8u * (1 + 4 * ST + 10 * QT);   // this CTA's tiles have stored the item's output (the Q buffers double as staging)
  auto ostage_done = [&](int i) { return bar_base + 8u * (2 + 4 * ST + 10 * QT + i); };   // tile i's output has been read out of the staging buffer
  const uint32_t tmem_slot = bar_base + 8u * Cfg::NBARS;
  const uint32_t xch_base = tmem_slot + 16u;
  const uint32_t items_base = xch_base + Cfg::XCH_BYTES;   // [0]: number of items, [1 + n]: {unit, first key step, key steps, workspace slot or -1}

  const uint32_t warp = warp_id_sync();
  const uint32_t lane = lane_id();
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const bool leader = rank == 0;
  // Persistent clusters: every role walks the cluster's item list (FmhaPairParams; built below by one thread before the set-up sync).
  // A unit owns 256 * QT consecutive query rows: tile i of the pair = rows [256 i, 256 i + 256), this CTA's half = [128 rank, +128).
  const int n_kv_all = (p.len_kv + BKV - 1) / BKV;
  struct Item {
    int qb, head, batch;
    int kv0, n_kv;   // first 128-key step and number of steps
    int part;        // >= 0: partial piece, its slot in the workspace
  };
  auto get_item = [&](int item) {
    Item w;
    int unit;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(unit), "=r"(w.kv0), "=r"(w.n_kv), "=r"(w.part) : "r"(items_base + 16u * (uint32_t)(item + 1)));
    w.qb = unit % p.q_blocks;
    w.head = (unit / p.q_blocks) % p.heads;
    w.batch = unit / (p.q_blocks * p.heads);
    return w;
  };
  auto q0_of = [&](const Item& w, int i) { return w.qb * (2 * QT * Cfg::BQ) + (2 * i + (int)rank) * Cfg::BQ; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_free, QT);
    for (int i = 0; i < 2; ++i) mbar_init(ostage_done(i), 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    for (int i = 0; i < QT; ++i) {
      for (int b = 0; b < 2; ++b) {
        mbar_init(s_full(i, b), 1);
        for (int hh = 0; hh < 2; ++hh) mbar_init(p_full(i, b, hh), 2 * 4 * SPLIT);   // one arrival per softmax warp of the tile in BOTH CTAs (leader's copy)
        mbar_init(pv_done(i, b), 1);
        mbar_init(pvh_done(i, b), 1);
      }
    }
    fence_barrier_init();
    // this cluster's items
    const int c = (int)(blockIdx.x >> 1), G = (int)(gridDim.x >> 1);
    int n = 0;
    auto put = [&](int unit, int kv0, int n_kv, int part) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(items_base + 16u * (uint32_t)(n + 1)), "r"(unit), "r"(kv0), "r"(n_kv), "r"(part) : "memory");
      ++n;
    };
    for (int u = c; u < p.n_full && n < Cfg::MAX_ITEMS; u += G) put(u, 0, n_kv_all, -1);
    if (p.tail_units > 0) {
      const long long T = (long long)p.tail_units * n_kv_all;
      const int lo = (int)(c * T / G), hi = (int)((c + 1) * T / G);
      pair_tail_pieces(c, lo, hi, n_kv_all, [&](int u, int kv0, int n_kv, int slot) {
        if (n < Cfg::MAX_ITEMS) put(p.n_full + u, kv0, n_kv, slot);
      });
    }
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(items_base), "r"(n) : "memory");
  }
  if (warp == 2) tmem_alloc<2>(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  int n_local;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(n_local) : "r"(items_base));
  pdl_launch_dependents();
  pdl_wait();  // PDL: q / k / v written by the previous kernel are read (and O written) only after this point
  if (p.trace && rank == 0 && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const long long now = (long long)globaltimer_ns();
    for (int item = 0; item < n_local; ++item) {
      const int row = item * (int)(gridDim.x >> 1) + (int)(blockIdx.x >> 1);
      if (row < 512) {
        p.trace[1024 + (long long)row * 8 + 6] = now;
        p.trace[1024 + (long long)row * 8 + 7] = smid | ((long long)get_item(item).n_kv << 16);
      }
    }
  }

  if (warp == 0) {
    // ------------------------------ TMA producer (warp-uniform control flow, one elected lane issues) ------------------------------
    int s = 0, n_it = 0, s_last = 0;
    uint32_t ph = 0, ph_last = 0;   // (s_last, ph_last): ring slot and phase of the previous item's last key tile
    for (int item = 0; item < n_local; ++item, ++n_it) {
      const Item w = get_item(item);
      if (lane == 0) PAIR_STAMP(0);
      if (n_it > 0) {
        if constexpr (Cfg::OSTAGE) {
          // the Q buffers are free once the previous item's last Q K^T has completed: the event that releases the ring slot of its key tile.
          // (Not s_full: this warp runs up to ST steps = two phases of that barrier ahead of the MMAs, which a parity wait cannot tell apart.)
          mbar_wait(k_empty(s_last), ph_last);
        } else {
          mbar_wait(q_free, (uint32_t)(n_it - 1) & 1u);   // the previous item's output has been read out of the Q buffers
        }
      }

      if (elect_one()) {
        if (leader) mbar_expect_tx(q_full, 2u * QT * Cfg::Q_TILE_BYTES);
#pragma unroll
        for (int i = 0; i < QT; ++i) {
#pragma unroll
          for (int sl = 0; sl < 2; ++sl) tma_load_4d_2sm(smem_q(i) + sl * Cfg::Q_SLAB_BYTES, &tmQ, q_full, sl * 64, w.head, q0_of(w, i), w.batch);
        }
      }
      __syncwarp();
      if (lane == 0) PAIR_STAMP(1);
      for (int j = 0; j < w.n_kv; ++j) {
        mbar_wait(k_empty(s), ph ^ 1u);
        if (elect_one()) {
          if (leader) mbar_expect_tx(k_full(s), 2u * Cfg::K_HALF_BYTES);
#pragma unroll
          for (int sl = 0; sl < 2; ++sl)
            tma_load_4d_2sm(smem_k(s) + sl * Cfg::K_SLAB_BYTES, &tmK, k_full(s), sl * 64, w.head, (w.kv0 + j) * BKV + (int)rank * (BKV / 2), w.batch);
        }
        __syncwarp();
        mbar_wait(v_empty(s), ph ^ 1u);
        if (elect_one()) {
          if (leader) mbar_expect_tx(v_full(s), 2u * Cfg::V_HALF_BYTES);
          tma_load_4d_2sm(smem_v(s), &tmV, v_full(s), (int)rank * 64, w.head, (w.kv0 + j) * BKV, w.batch);
        }
        __syncwarp();
        s_last = s;
        ph_last = ph;
        if (++s == ST) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issue (leader CTA; the whole warp runs the control flow, one elected lane issues) ----------
    if (leader) {
      constexpr uint32_t idesc_qk = make_idesc(kFmtBF16, 256, BKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(kFmtBF16, 256, D, 0, 1);   // B (V) is MN-major
      int qs = 0, vs = 0;
      uint32_t qph = 0, vph = 0;
      // S_i(j) = Q_i K(j)^T for tile i; all tiles use key tile j back to back: the first waits for it, the last releases its ring slot
      // g = number of the step counted over all items of this cluster: score buffer and barrier parities follow it
      auto issue_qk = [&](int i, int g) {
        if (i == 0) mbar_wait(k_full(qs), qph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + (uint32_t)i * Cfg::TILE_COLS + Cfg::TM_S + (uint32_t)(g % NSB) * Cfg::S_STRIDE;
          const uint64_t qdesc = make_smem_desc_sw128(smem_q(i), 1024, 0);
          const uint64_t kdesc = make_smem_desc_sw128(smem_k(qs), 1024, 0);
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            // +32 B along the head dim inside a swizzle atom = +2 in the (addr >> 4) field; next 64-column slab = + SLAB_BYTES
            const uint64_t ao = (uint64_t)(((kk >> 2) * Cfg::Q_SLAB_BYTES + (kk & 3) * 32) >> 4);
            const uint64_t bo = (uint64_t)(((kk >> 2) * Cfg::K_SLAB_BYTES + (kk & 3) * 32) >> 4);
            umma_f16_ss<2>(d_tmem, qdesc + ao, kdesc + bo, idesc_qk, kk ? 1u : 0u);
          }
          umma_commit_2sm_mc(s_full(i, g & 1), 3);
          if (i == QT - 1) umma_commit_2sm_mc(k_empty(qs), 3);
        }
        __syncwarp();
        if (i == QT - 1) { if (++qs == ST) { qs = 0; qph ^= 1u; } }
      };
      auto issue_pv = [&](int i, int j, int g, int hh) {   // j: step within the item; hh: key half of P(j) (NH == 1: the whole tile)
        if (i == 0 && hh == 0) mbar_wait(v_full(vs), vph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t t_base = tmem_base + (uint32_t)i * Cfg::TILE_COLS;
          const uint32_t p_tmem = t_base + Cfg::TM_S + (uint32_t)(g % NSB) * Cfg::S_STRIDE;
          // this CTA's V half: key rows at a 128 B pitch (K dimension of the MMA), 64 head-dim columns = one swizzle row (MN dimension)
          const uint64_t vdesc = make_smem_desc_sw128(smem_v(vs), 1024, Cfg::V_HALF_BYTES);
          constexpr int KH = BKV / 16 / NH;
#pragma unroll
          for (int k2 = 0; k2 < KH; ++k2) {
            const int kk = hh * KH + k2;
            umma_f16_ts_2sm(t_base + Cfg::TM_O, p_tmem + (uint32_t)(kk * 8), vdesc + (uint64_t)((kk * 2048) >> 4), idesc_pv, (j | kk) ? 1u : 0u);
          }
          if (hh == NH - 1) {
            umma_commit_2sm_mc(pv_done(i, g & 1), 3);
            if (i == QT - 1) umma_commit_2sm_mc(v_empty(vs), 3);
          } else if (Cfg::FAST) {
            umma_commit_2sm_mc(pvh_done(i, g & 1), 3);
          }
        }
        __syncwarp();
        if (i == QT - 1 && hh == NH - 1) { if (++vs == ST) { vs = 0; vph ^= 1u; } }
      };
      int g0 = 0, n_it = 0;
      for (int item = 0; item < n_local; ++item, ++n_it) {
        const Item w = get_item(item);
        const bool has_next = item + 1 < n_local;
        if (!Cfg::OSTAGE || n_it == 0) {
          // Q of this item is in shared memory (without the staging buffer: both CTAs have stored the previous item's output from there)
          mbar_wait(q_full, (uint32_t)n_it & 1u);
          for (int j = 0; j < NSB && j < w.n_kv; ++j) {
#pragma unroll
            for (int i = 0; i < QT; ++i) issue_qk(i, g0 + j);
          }
        }
        for (int j = 0; j < w.n_kv; ++j) {
          const int g = g0 + j;
#pragma unroll
          for (int i = 0; i < QT; ++i) {
#pragma unroll
            for (int hh = 0; hh < NH; ++hh) {
              mbar_wait(p_full(i, g & 1, hh), (uint32_t)(g >> 1) & 1u);   // (this half of) P_i(j) of both CTAs is in tensor memory
              tc_fence_after();
              if (lane == 0 && hh == 0) PAIR_TRACE(j, i, 0);
              issue_pv(i, j, g, hh);
            }
            if (j + NSB < w.n_kv) {
              issue_qk(i, g + NSB);   // overwrites S_i(j) | P_i(j): ordered behind P_i(j) V(j) by the in-order tensor pipe
            } else if (Cfg::OSTAGE && has_next) {
              // first Q K^T of the NEXT item, behind this item's last P V.  Its P V (which starts O afresh) waits for P like any other, and the
              // softmax warps hand that P over only after they have read this item's O out of tensor memory
              if (i == 0) mbar_wait(q_full, (uint32_t)(n_it + 1) & 1u);
              issue_qk(i, g + 1);
            }
            if (lane == 0) PAIR_TRACE(j, i, 1);
          }
        }
        g0 += w.n_kv;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ softmax / correction / epilogue (both CTAs) ------------------------------
    const int i = (int)(warp - 4u) / (4 * SPLIT);    // query tile of this warp
    const int h = ((int)(warp - 4u) >> 2) % SPLIT;   // column slice of this warp
    const uint32_t wq = warp & 3u;                   // TMEM lane quadrant this warp may access
    const int rit = (int)(wq * 32u + lane);          // row in this CTA's tile
    const uint32_t lane_base = tmem_base + ((wq * 32u) << 16) + (uint32_t)i * Cfg::TILE_COLS;
    const uint32_t o_addr = lane_base + Cfg::TM_O + (uint32_t)(h * OC);
    auto xch = [&](int par, int slice) { return xch_base + 4u * (uint32_t)(((par * QT + i) * SPLIT + slice) * 128 + rit); };
    const uint32_t quad_bar = 1u + (uint32_t)i * 4u + wq;   // named barrier of the SPLIT warps that own these 32 rows
    int g = 0;                                         // step counted over all items of this cluster (barrier parities)
    for (int item = 0; item < n_local; ++item) {
    // (only what the key loop needs stays live across it; the epilogue decodes the item again)
    int n_kv, kvalid;                                // steps of this item; keys from its first step to the end of the sequence
    float c;
    {
      const Item w = get_item(item);
      n_kv = w.n_kv;
      kvalid = p.len_kv - w.kv0 * BKV;
      const int row = q0_of(w, i) + rit;
      c = (p.row_scale && row < p.len_q) ? p.scale_log2 * p.row_scale[(long long)w.batch * p.len_q + row] : p.scale_log2;
    }
    float m_run = -INFINITY;                         // running (possibly stale) row max of raw scores
    float l_run = 0.0f;                              // running sum of exp2((s - m_run) * c) over this thread's columns
    const uint64_t cc2 = pack2(c, c);
    if constexpr (Cfg::FAST) {
      // ---- speculative softmax: 64-column half-steps against the stale running maximum; each half of P(j) is handed to the tensor pipe as
      //      soon as it is stored (fmha_sm100.cu runs the same scheme on one CTA) ----
      float nmc = 0.0f;
      uint64_t mc2 = 0ull;
      for (int j = 0; j < n_kv; ++j, ++g) {
        const int b = g & 1;
        mbar_wait(s_full(i, b), (uint32_t)(g >> 1) & 1u);
        tc_fence_after();
        const bool tr = wq == 0 && lane == 0;
        if (tr) PAIR_TRACE(j, i, 2);
        if (tr && i == 0 && j == 0) PAIR_STAMP(2);
        const uint32_t p_addr = lane_base + Cfg::TM_S;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[64];
          tmem_ld_x32(p_addr + (uint32_t)(hh * 64), r);
          tmem_ld_x32(p_addr + (uint32_t)(hh * 64) + 32u, r + 32);
          tmem_ld_wait();
          if (tr && hh == 0) PAIR_TRACE(j, i, 3);
          const int valid = kvalid - j * BKV - hh * 64;
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
          const bool need = (m_half - m_run) * c > 8.0f;
          if (__any_sync(0xffffffffu, need)) {
            // (rare) the stale maximum is too small for some row of this warp: rescale what has been accumulated, redo the half-step
            const float f = need ? ex2_approx((m_run - m_half) * c) : 1.0f;
            if (need) m_run = m_half;
            l_run *= f;
            if (j >= 1 || hh > 0) {
              // O holds P(0..j-1) V [+ the first half of P(j) V, handed over already]: those products must have completed; the second half
              // of P(j) V is not issued before this warp arrives on p_full, so O is quiescent afterwards
              if (j >= 1) mbar_wait(pv_done(i, (g - 1) & 1), (uint32_t)((g - 1) >> 1) & 1u);
              if (hh > 0) mbar_wait(pvh_done(i, b), (uint32_t)(g >> 1) & 1u);
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
              tmem_st_wait();
            }
            nmc = -m_run * c;
            mc2 = pack2(nmc, nmc);
            hsum[0] = hsum[1] = 0ull;
            exp_half64_exact(r, cc2, mc2, hsum, pk);
          }
          {
            float s0, s1, s2, s3;
            unpack2(hsum[0], s0, s1);
            unpack2(hsum[1], s2, s3);
            l_run += (s0 + s1) + (s2 + s3);
          }
          if (tr && hh == 1) PAIR_TRACE(j, i, 5);
          tmem_st_x32(p_addr + (uint32_t)(hh * 32), pk);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (tr && hh == 1) PAIR_TRACE(j, i, 6);
          if (lane == 0) {
            if (leader) mbar_arrive(p_full(i, b, hh)); else mbar_arrive_remote(p_full(i, b, hh), 0);
          }
        }
      }
    } else
    for (int j = 0; j < n_kv; ++j, ++g) {
      const int b = g & 1, sb = g % NSB;
      mbar_wait(s_full(i, b), (uint32_t)(g >> 1) & 1u);
      tc_fence_after();
      const bool tr = h == 0 && wq == 0 && lane == 0;
      if (tr) PAIR_TRACE(j, i, 2);
      uint32_t r[HC];
      const uint32_t s_addr = lane_base + Cfg::TM_S + (uint32_t)sb * Cfg::S_STRIDE + (uint32_t)(h * HC);
#pragma unroll
      for (int cb = 0; cb < HC / 32; ++cb) tmem_ld_x32(s_addr + (uint32_t)(cb * 32), r + cb * 32);
      tmem_ld_wait();
      if (tr) PAIR_TRACE(j, i, 3);
      const int valid = kvalid - j * BKV - h * HC;   // columns of this thread that hold existing keys
      if (valid < HC) {
#pragma unroll
        for (int k = 0; k < HC; ++k)
          if (k >= valid) r[k] = 0xff800000u;  // -inf
      }
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int k = 0; k < HC; k += 8) {
#pragma unroll
        for (int u = 0; u < 4; ++u) mx[u] = fmaxf(fmaxf(mx[u], __uint_as_float(r[k + 2 * u])), __uint_as_float(r[k + 2 * u + 1]));
      }
      float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      // the slices of a row exchange their partial max; the barrier also orders every slice's S loads before anybody's P stores into
      // the columns they were read from (P aliases the first 64 columns of the score buffer)
      if constexpr (SPLIT > 1) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch(b, h)), "f"(m_tile) : "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(quad_bar), "n"(32 * SPLIT) : "memory");
#pragma unroll
        for (int o = 1; o < SPLIT; ++o) {
          float other;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(xch(b, (h + o) % SPLIT)) : "memory");
          m_tile = fmaxf(m_tile, other);
        }
      }
      const float m_new = fmaxf(m_run, m_tile);
      if (tr) PAIR_TRACE(j, i, 4);
      if (j == 0) {
        m_run = m_new;
      } else {
        const bool need = (m_new - m_run) * c > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // O is accumulated by P(j-1) V(j-1): it must have finished before the rows are rescaled
          mbar_wait(pv_done(i, (g - 1) & 1), (uint32_t)((g - 1) >> 1) & 1u);
          tc_fence_after();
          const float f = need ? ex2_approx((m_run - m_new) * c) : 1.0f;
          if (need) m_run = m_new;
          l_run *= f;
#pragma unroll 1
          for (int cb = 0; cb < OC / 16; ++cb) {   // 16 columns at a time: the score row stays live in registers
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
      const float nmc = -m_run * c;
      const uint64_t mc2 = pack2(nmc, nmc);
      uint64_t sum2[2] = {0ull, 0ull};   // packed (even, odd) column partial sums
      const uint32_t p_addr = lane_base + Cfg::TM_S + (uint32_t)sb * Cfg::S_STRIDE + (uint32_t)(h * (HC / 2));
#pragma unroll
      for (int cb = 0; cb < HC / 32; ++cb) {
        uint32_t pk[16];
        if (cb * 32 >= valid) {   // (warp-uniform) a 32-column chunk past the last key: P = 0 without its exponentials
#pragma unroll
          for (int k = 0; k < 16; ++k) pk[k] = 0u;
        } else {
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
        }
        tmem_st_x16(p_addr + (uint32_t)(cb * 16), pk);
        if constexpr (SPLIT == 1) {
          if (cb == 1) {   // the first 64 keys of P(j) are complete: the tensor pipe starts on them while the other 64 are exponentiated
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (leader) mbar_arrive(p_full(i, b, 0)); else mbar_arrive_remote(p_full(i, b, 0), 0);
            }
          }
        }
      }
      {
        float s0, s1, s2, s3;
        unpack2(sum2[0], s0, s1);
        unpack2(sum2[1], s2, s3);
        l_run += (s0 + s1) + (s2 + s3);
      }
      if (tr) PAIR_TRACE(j, i, 5);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (tr) PAIR_TRACE(j, i, 6);
      if (lane == 0) {
        // (CTA-scope release: the data handed over is in tensor memory, ordered by the tcgen05 fences; a cluster-scope release costs a
        //  MEMBAR.GPU + L1 invalidate per step -- measured 9 % of all warp samples)
        if (leader) mbar_arrive(p_full(i, b, NH - 1)); else mbar_arrive_remote(p_full(i, b, NH - 1), 0);
      }
    }
    // ---- epilogue: O / l -> bf16 -> shared memory (this CTA's Q buffer: every MMA has completed) -> one bulk tensor store per slab ----
    if (i == 0 && h == 0 && rit == 0) PAIR_STAMP(3);
    if constexpr (SPLIT > 1) {
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch(g & 1, h)), "f"(l_run) : "memory");
      asm volatile("bar.sync %0, %1;" ::"r"(quad_bar), "n"(32 * SPLIT) : "memory");
#pragma unroll
      for (int o = 1; o < SPLIT; ++o) {
        float other;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(xch(g & 1, (h + o) % SPLIT)) : "memory");
        l_run += other;
      }
    }
    mbar_wait(pv_done(i, (g - 1) & 1), (uint32_t)((g - 1) >> 1) & 1u);   // the commit covers every earlier MMA too
    tc_fence_after();
    if (i == 0 && h == 0 && rit == 0) PAIR_STAMP(4);
    const float inv_l = 1.0f / l_run;
    const uint32_t stage = Cfg::OSTAGE ? smem_ostage : smem_q(i);
    if constexpr (Cfg::OSTAGE) {
      // the two tiles take turns on the staging buffer: tile 1 after tile 0 of the same item, tile 0 after tile 1 of the previous item
      const int n_done = item;
      if (i == 1) mbar_wait(ostage_done(0), (uint32_t)n_done & 1u);
      else if (n_done > 0) mbar_wait(ostage_done(1), (uint32_t)(n_done - 1) & 1u);
    }
#pragma unroll
    for (int cb = 0; cb < OC / 32; ++cb) {
      uint32_t o[32];
      tmem_ld_x32(o_addr + (uint32_t)(cb * 32), o);
      tmem_ld_wait();
      const int col0 = h * OC + cb * 32;   // first of 32 consecutive output columns
      const uint32_t srow = stage + (uint32_t)((col0 >> 6) * Cfg::Q_SLAB_BYTES + rit * 128);
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
    if constexpr (Cfg::OSTAGE) {
      // warp 3 stores the staged tile (below): the softmax threads only signal it and turn to the next item
      asm volatile("bar.arrive %0, %1;" ::"r"(9u + (uint32_t)i), "n"(128 * SPLIT + 32) : "memory");
    } else {
      asm volatile("bar.sync %0, %1;" ::"r"(9u + (uint32_t)i), "n"(128 * SPLIT) : "memory");   // all softmax threads of this tile
    }
    int item_e = item;
    asm volatile("" : "+r"(item_e));   // (decoded again rather than kept in registers across the key loop)
    const Item w = get_item(item_e);
    const int part = w.part, head = w.head, batch = w.batch, q0 = q0_of(w, i);
    if (part >= 0 && h == 0) {   // chunk cluster: what the merge needs to weigh this partial result (row sums are complete in every slice)
      const int wrow = part * (2 * QT * Cfg::BQ) + (2 * i + (int)rank) * Cfg::BQ + rit;
      *reinterpret_cast<float2*>(p.ws_ml + 2ll * wrow) = make_float2(m_run * c, l_run);
    }
    if (!Cfg::OSTAGE && h == 0 && rit == 0) {
      if (part >= 0) {
#pragma unroll
        for (int sl = 0; sl < 2; ++sl)
          tma_store_4d(&tmW, stage + sl * Cfg::Q_SLAB_BYTES, sl * 64, 0, part * (2 * QT * Cfg::BQ) + (2 * i + (int)rank) * Cfg::BQ, 0);
      } else {
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) tma_store_4d(&tmO, stage + sl * Cfg::Q_SLAB_BYTES, sl * 64, head, q0, batch);   // rows >= len_q are clipped
      }
      tma_store_commit();
      tma_store_wait_read<0>();   // the buffer must stay intact until the TMA unit has read it
      if (i == 0) PAIR_STAMP(5);
      mbar_arrive(q_free);        // the producer may load the next item's Q over it
    }
    }   // items
  } else if (warp == 3 && Cfg::OSTAGE) {
    // ------------------------------ output store (both CTAs): staging buffer -> global memory, tile 0 and tile 1 of every item in turn ------
    for (int item = 0; item < n_local; ++item) {
      const Item w = get_item(item);
#pragma unroll 1
      for (int i = 0; i < QT; ++i) {
        asm volatile("bar.sync %0, %1;" ::"r"(9u + (uint32_t)i), "n"(128 * SPLIT + 32) : "memory");   // tile i's softmax threads have staged their rows
        if (lane == 0) {
          if (w.part >= 0) {
#pragma unroll
            for (int sl = 0; sl < 2; ++sl)
              tma_store_4d(&tmW, smem_ostage + sl * Cfg::Q_SLAB_BYTES, sl * 64, 0, w.part * (2 * QT * Cfg::BQ) + (2 * i + (int)rank) * Cfg::BQ, 0);
          } else {
#pragma unroll
            for (int sl = 0; sl < 2; ++sl)
              tma_store_4d(&tmO, smem_ostage + sl * Cfg::Q_SLAB_BYTES, sl * 64, w.head, q0_of(w, i), w.batch);   // rows >= len_q are clipped
          }
          tma_store_commit();
          tma_store_wait_read<0>();   // the buffer must stay intact until the TMA unit has read it
          if (i == 0) PAIR_STAMP(5);
          mbar_arrive(ostage_done(i));   // the other tile may stage its output
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // the peer's tensor memory / shared memory is in use until both CTAs are done
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

static int make_map4(CUtensorMap* tm, const void* ptr, long long B, long long H, long long L, long long D, long long bs, long long rs,
                     long long hs, uint32_t box_rows) {
  // dims (fastest first): head_dim, heads, rows, batch; box = 64 head-dim columns (128 B: one swizzle row) x box_rows rows
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)H, (uint64_t)L, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)hs * 2, (uint64_t)rs * 2, (uint64_t)bs * 2};
  uint32_t box[4] = {64, 1, box_rows, 1};
  return encode_tensor_map(tm, ptr, 2, false, 4, dims, strides, box, true);
}

constexpr int kPairMaxClusters = 128;
// ---- merge of the pieces of the units that were cut: one warp per (range boundary, query row) ---------------------------------------------
// O[row] = sum_k w_k O_k[row] / sum_k w_k,  w_k = l_k 2^(m_k - max_k m_k)   (m_k = the piece's reference maximum in the log2 domain, l_k its row sum)
struct FmhaPairCombineParams {
  const __nv_bfloat16* ws_o;   // [slot][rows_per_unit][128] normalised partial O
  const float* ws_ml;          // [slot][rows_per_unit][2]
  __nv_bfloat16* O;
  long long o_bs, o_rs, o_hs;
  int len_q, q_blocks, heads, rows_per_unit;
  int n_full, tail_units, n_kv, clusters, n_cut;
  int lo[kPairMaxClusters + 1];   // lo[c] = c T / clusters: first tail step of cluster c (host-computed: the kernel would spend its time in 64-bit divisions)
  int cut[kPairMaxClusters];      // the cut units: for each, the first range boundary c (between clusters c - 1 and c) inside it
};
// One warp per (cut unit, 4 query rows); a lane owns 4 head-dim columns.  The kernel is pure latency (15 MB out of L2 for the whole merge):
// the loads of up to 4 pieces x 4 rows are issued together, and the grid is one wave.
constexpr int kCombineRows = 4, kCombineBatch = 4;
__global__ void __launch_bounds__(256) fmha_pair_combine_kernel(const FmhaPairCombineParams p) {
  pdl_launch_dependents();
  const int w = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = (int)(threadIdx.x & 31u);
  const int groups = p.rows_per_unit / kCombineRows;
  const int ci = w / groups, r0 = (w - ci * groups) * kCombineRows;
  if (ci >= p.n_cut) return;
  const int c = p.cut[ci];
  const int u = p.lo[c] / p.n_kv, u_lo = u * p.n_kv, u_hi = u_lo + p.n_kv;
  const int unit = p.n_full + u;
  const int qb = unit % p.q_blocks, head = (unit / p.q_blocks) % p.heads, batch = unit / (p.q_blocks * p.heads);
  // pieces: clusters c - 1, c, ... while their range starts inside the unit; a cluster's piece is its first item (slot 2 cc) when its range
  // starts inside this unit, else its last (slot 2 cc + 1)
  float mx[kCombineRows], wsum[kCombineRows], acc[kCombineRows][4];
#pragma unroll
  for (int rr = 0; rr < kCombineRows; ++rr) {
    mx[rr] = -INFINITY;
    wsum[rr] = 0.f;
    acc[rr][0] = acc[rr][1] = acc[rr][2] = acc[rr][3] = 0.f;
  }
  pdl_wait();
  int cc = c - 1;
  bool more = true;
  while (more) {
    int slot[kCombineBatch];
#pragma unroll
    for (int k = 0; k < kCombineBatch; ++k) {
      slot[k] = more ? pair_next_piece_slot(p.lo, p.clusters, u_lo, u_hi, cc) : -1;
      more = slot[k] >= 0;
    }
    float2 ml[kCombineBatch][kCombineRows];
    uint2 v[kCombineBatch][kCombineRows];
#pragma unroll
    for (int k = 0; k < kCombineBatch; ++k) {
#pragma unroll
      for (int rr = 0; rr < kCombineRows; ++rr) {
        const long long wrow = (long long)max(slot[k], 0) * p.rows_per_unit + r0 + rr;
        ml[k][rr] = *reinterpret_cast<const float2*>(p.ws_ml + 2 * wrow);
        v[k][rr] = *reinterpret_cast<const uint2*>(p.ws_o + wrow * 128 + lane * 4);
      }
    }
#pragma unroll
    for (int k = 0; k < kCombineBatch; ++k) {
      if (slot[k] >= 0) {
#pragma unroll
        for (int rr = 0; rr < kCombineRows; ++rr) {
          const float m_new = fmaxf(mx[rr], ml[k][rr].x);
          const float f = ex2_approx(mx[rr] - m_new), wk = ml[k][rr].y * ex2_approx(ml[k][rr].x - m_new);
          mx[rr] = m_new;
          wsum[rr] = wsum[rr] * f + wk;
          acc[rr][0] = fmaf(wk, __uint_as_float(v[k][rr].x << 16), acc[rr][0] * f);
          acc[rr][1] = fmaf(wk, __uint_as_float(v[k][rr].x & 0xffff0000u), acc[rr][1] * f);
          acc[rr][2] = fmaf(wk, __uint_as_float(v[k][rr].y << 16), acc[rr][2] * f);
          acc[rr][3] = fmaf(wk, __uint_as_float(v[k][rr].y & 0xffff0000u), acc[rr][3] * f);
        }
      }
    }
  }
#pragma unroll
  for (int rr = 0; rr < kCombineRows; ++rr) {
    const int row = qb * p.rows_per_unit + r0 + rr;
    if (row < p.len_q) {
      const float inv = 1.0f / wsum[rr];
      uint2 o;
      o.x = pack_bf16(acc[rr][0] * inv, acc[rr][1] * inv);
      o.y = pack_bf16(acc[rr][2] * inv, acc[rr][3] * inv);
      *reinterpret_cast<uint2*>(p.O + (long long)batch * p.o_bs + (long long)row * p.o_rs + (long long)head * p.o_hs + lane * 4) = o;
    }
  }
}

// The decomposition (FmhaPairParams).  `slots` clusters run at a time.  The key split pays when the balanced share of the tail plus the cost of
// two partial pieces and the merge (kSplitCost steps) is shorter than the whole pass over the keys the tail units would otherwise take.
struct PairPlan {
  int clusters, n_full, tail_units;
};
static PairPlan plan_pair(long long units, int n_kv, int slots, bool allow_split, bool persistent, int max_items) {
  constexpr int kMinSteps = 4, kSplitCost = 4, kMinKv = 8;
  PairPlan whole{(int)std::min<long long>(slots, units), (int)units, 0};
  if (!persistent || slots <= 0 || (units + whole.clusters - 1) / whole.clusters > max_items) return PairPlan{(int)units, (int)units, 0};
  if (!allow_split || n_kv < kMinKv) return whole;
  // small problems: as many clusters as there are kMinSteps-step pieces, at most one per SM pair
  const int G = (int)std::min<long long>(slots, std::max<long long>(units, units * n_kv / kMinSteps));
  const int n_full = (int)(units / G) * G, tail = (int)(units - n_full);
  if (tail == 0) return whole;
  const long long share = ((long long)tail * n_kv + G - 1) / G;
  if (share + kSplitCost >= n_kv || n_full / G + share / n_kv + 3 > max_items) return whole;
  return PairPlan{G, n_full, tail};
}
static constexpr long long kPairSlotBytes(int rows_per_unit) { return (long long)rows_per_unit * (128 * 2 + 2 * 4); }

template <int QT_, int SPLIT_, int POLY_, int FAST_ = 0>
static int pair_slots(int* slots) {
  // clusters of this kernel the device runs at a time (cached per device)
  using Cfg = FmhaPairCfg<QT_, SPLIT_, POLY_, FAST_>;
  static std::atomic<int> cache[64];
  int dev = 0;
  V3A_CUDA_OK(cudaGetDevice(&dev));
  int v = dev < 64 ? cache[dev].load(std::memory_order_acquire) : 0;
  if (v <= 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * 1024);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fmha_pair_kernel<QT_, SPLIT_, POLY_, FAST_>, &cfg) != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      int sms = 0;
      V3A_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      n = sms / 2;
    }
    v = n;
    if (dev < 64) cache[dev].store(v, std::memory_order_release);
  }
  *slots = v;
  return VIST3A_OK;
}

// flags bit 17: no key split; bit 20: one cluster per unit instead of persistent clusters (the round-2 decomposition; A/B measurements)
template <int QT_, int SPLIT_, int POLY_, int FAST_ = 0>
static int launch_fmha_pair(const vist3a_fmha_args& a, cudaStream_t stream, long long* ws_query) {
  using Cfg = FmhaPairCfg<QT_, SPLIT_, POLY_, FAST_>;
  auto kern = fmha_pair_kernel<QT_, SPLIT_, POLY_, FAST_>;
  static std::atomic<unsigned long long> attr_done{0};
  V3A_CUDA_OK(ensure_dynamic_smem(kern, Cfg::SMEM_BYTES, attr_done));
  const int rows_per_unit = 2 * Cfg::QT * Cfg::BQ;
  const long long q_blocks = (a.len_q + rows_per_unit - 1) / rows_per_unit;
  const long long units = q_blocks * a.heads * a.batch;
  V3A_REQUIRE(units < (1ll << 29), VIST3A_ERR_INVALID, "fmha: too many query blocks");
  const int n_kv = (int)((a.len_kv + Cfg::BKV - 1) / Cfg::BKV);
  int slots = 0, rc = pair_slots<QT_, SPLIT_, POLY_, FAST_>(&slots);
  if (rc) return rc;
  slots = std::min(slots, kPairMaxClusters);
  const bool persistent = !(a.flags & (1u << 20));
  PairPlan plan = plan_pair(units, n_kv, slots, !(a.flags & (1u << 17)), persistent, Cfg::MAX_ITEMS);
  const long long ws_bytes = plan.tail_units ? 2ll * plan.clusters * kPairSlotBytes(rows_per_unit) : 0;
  if (ws_query) {
    *ws_query = ws_bytes;
    return VIST3A_OK;
  }
  if (plan.tail_units && (a.workspace == nullptr || a.workspace_bytes < ws_bytes || ((uintptr_t)a.workspace & 127) != 0))
    plan = plan_pair(units, n_kv, slots, false, persistent, Cfg::MAX_ITEMS);   // no (usable) workspace: every unit over all keys
  CUtensorMap tmQ, tmK, tmV, tmO, tmW;
  if ((rc = make_map4(&tmQ, a.Q, a.batch, a.heads, a.len_q, 128, a.q_bs, a.q_rs, a.q_hs, Cfg::BQ))) return rc;
  if ((rc = make_map4(&tmK, a.K, a.batch, a.heads, a.len_kv, 128, a.k_bs, a.k_rs, a.k_hs, Cfg::BKV / 2))) return rc;
  if ((rc = make_map4(&tmV, a.V, a.batch, a.heads, a.len_kv, 128, a.v_bs, a.v_rs, a.v_hs, Cfg::BKV))) return rc;
  if ((rc = make_map4(&tmO, a.O, a.batch, a.heads, a.len_q, 128, a.o_bs, a.o_rs, a.o_hs, Cfg::BQ))) return rc;
  FmhaPairParams p;
  p.len_q = (int)a.len_q;
  p.len_kv = (int)a.len_kv;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.row_scale = a.q_row_scale;
  p.trace = g_pair_trace.load(std::memory_order_relaxed);
  p.zero = 0u;
  p.q_blocks = (int)q_blocks;
  p.heads = (int)a.heads;
  p.n_full = plan.n_full;
  p.tail_units = plan.tail_units;
  p.ws_ml = nullptr;
  if (plan.tail_units) {
    const long long wrows = 2ll * plan.clusters * rows_per_unit;
    if ((rc = make_map4(&tmW, a.workspace, 1, 1, wrows, 128, wrows * 128, 128, 128, Cfg::BQ))) return rc;
    p.ws_ml = reinterpret_cast<float*>(static_cast<char*>(a.workspace) + wrows * 128 * 2);
  } else {
    tmW = tmO;
  }
  dim3 grid((unsigned)(2 * plan.clusters));
  V3A_CUDA_OK(launch_kernel(kern, grid, dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, /*pdl=*/true, 2, tmQ, tmK, tmV, tmO, tmW, p));
  launch_counter().fetch_add(1);
  if (plan.tail_units) {
    FmhaPairCombineParams c;
    c.ws_o = static_cast<const __nv_bfloat16*>(a.workspace);
    c.ws_ml = p.ws_ml;
    c.O = static_cast<__nv_bfloat16*>(a.O);
    c.o_bs = a.o_bs;
    c.o_rs = a.o_rs;
    c.o_hs = a.o_hs;
    c.len_q = (int)a.len_q;
    c.q_blocks = (int)q_blocks;
    c.heads = (int)a.heads;
    c.rows_per_unit = rows_per_unit;
    c.n_full = plan.n_full;
    c.tail_units = plan.tail_units;
    c.n_kv = n_kv;
    c.clusters = plan.clusters;
    for (int cc = 0; cc <= plan.clusters; ++cc) c.lo[cc] = (int)((long long)cc * plan.tail_units * n_kv / plan.clusters);
    c.n_cut = pair_cut_list(c.lo, plan.clusters, n_kv, c.cut);
    const long long warps = (long long)c.n_cut * (rows_per_unit / kCombineRows);
    if (warps > 0 && !(a.flags & (1u << 22))) {   // (flags bit 22, timing only: no merge -- the cut units keep their stale output)
      V3A_CUDA_OK(launch_kernel(fmha_pair_combine_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, stream, /*pdl=*/true, 1, c));
      launch_counter().fetch_add(1);
    }
  }
  return VIST3A_OK;
}

// Host-side self-check of the decomposition (no CUDA call; tests/test_abi_cpu.py sweeps it): builds every cluster's item list and the merge's
// piece lists with the functions the kernels use and verifies that (1) every key step of every unit is covered exactly once, (2) no cluster
// exceeds max_items, (3) every workspace slot is written at most once, (4) the merge visits, for every cut unit, exactly the slots its partial
// pieces were written to and each uncut unit is written whole.  Returns 0, or the number of the violated property; out = {clusters, n_full,
// tail_units, cut units, partial pieces}.
extern "C" int v3a_debug_fmha_pair_plan_check(long long units, int n_kv, int slots, int allow_split, int persistent, int* out) {
  constexpr int kMaxItems = FmhaPairCfg<2, 1, 3, 1>::MAX_ITEMS;
  if (units <= 0 || units > (1 << 20) || n_kv <= 0 || n_kv > (1 << 14) || slots <= 0) return -1;
  slots = std::min(slots, kPairMaxClusters);
  const PairPlan plan = plan_pair(units, n_kv, slots, allow_split != 0, persistent != 0, kMaxItems);
  const int G = plan.clusters;
  std::vector<int> covered((size_t)units * n_kv, 0), slot_unit(2 * (size_t)G, -1), lo(G + 1, 0);
  std::vector<char> whole(units, 0);
  int partial = 0;
  for (int c = 0; c <= G; ++c) lo[c] = (int)((long long)c * plan.tail_units * n_kv / G);
  for (int c = 0; c < G; ++c) {
    int items = 0;
    for (int u = c; u < plan.n_full; u += G) {
      ++items;
      whole[u] = 1;
      for (int j = 0; j < n_kv; ++j) ++covered[(size_t)u * n_kv + j];
    }
    int bad = 0;
    if (plan.tail_units) {
      pair_tail_pieces(c, lo[c], lo[c + 1], n_kv, [&](int u, int kv0, int n, int slot) {
        ++items;
        const int unit = plan.n_full + u;
        for (int j = 0; j < n; ++j) ++covered[(size_t)unit * n_kv + kv0 + j];
        if (slot < 0) {
          whole[unit] = 1;
        } else {
          ++partial;
          if (slot >= 2 * G || slot_unit[slot] != -1) bad = 3;
          else slot_unit[slot] = unit;
        }
      });
    }
    if (bad) return bad;
    if (items > kMaxItems) return 2;
  }
  for (int v : covered)
    if (v != 1) return 1;
  int n_cut = 0;
  if (plan.tail_units) {
    if (G > kPairMaxClusters) return -2;
    int cut[kPairMaxClusters];
    n_cut = pair_cut_list(lo.data(), G, n_kv, cut);
    std::vector<char> merged(units, 0);
    int visited = 0;
    for (int i = 0; i < n_cut; ++i) {
      const int u = lo[cut[i]] / n_kv, unit = plan.n_full + u;
      if (whole[unit] || merged[unit]) return 4;
      merged[unit] = 1;
      int cc = cut[i] - 1;
      for (int slot; (slot = pair_next_piece_slot(lo.data(), G, u * n_kv, (u + 1) * n_kv, cc)) >= 0; ++visited)
        if (slot_unit[slot] != unit) return 4;
    }
    if (visited != partial) return 4;
    for (long long u = 0; u < units; ++u)
      if (!whole[u] && !merged[u]) return 4;
  } else if (partial) {
    return 4;
  }
  if (out) {
    out[0] = G; out[1] = plan.n_full; out[2] = plan.tail_units; out[3] = n_cut; out[4] = partial;
  }
  return 0;
}

// head_dim 128 on CTA pairs; `variant` (A/B measurements): 0 = default: two query tiles per CTA, one thread per query row, P handed over in
// two key halves, all exponentials on the MUFU; 1 / 2 = the same with 1 / 2 of every 8 column pairs on the FMA pipe; 3 = speculative
// softmax (stale maximum, 64-column half-steps), all MUFU; 4 = the same with 2 of 8 on the FMA pipe; 5 = default with 3 of 8; 6 = speculative, 1 of 8; 7 = speculative, 3 of 8
// ws_query != nullptr: no launch, *ws_query = workspace bytes the key split of this problem would use (0: no split)
int fmha_pair_entry(const vist3a_fmha_args& a, int variant, cudaStream_t stream, long long* ws_query) {
  switch (variant) {
    case 1: return launch_fmha_pair<2, 1, 1>(a, stream, ws_query);
    case 2: return launch_fmha_pair<2, 1, 2>(a, stream, ws_query);
    case 3: return launch_fmha_pair<2, 1, 0, 1>(a, stream, ws_query);
    case 4: return launch_fmha_pair<2, 1, 2, 1>(a, stream, ws_query);
    case 5: return launch_fmha_pair<2, 1, 3>(a, stream, ws_query);
    case 6: return launch_fmha_pair<2, 1, 1, 1>(a, stream, ws_query);
    case 7: return launch_fmha_pair<2, 1, 3, 1>(a, stream, ws_query);
    default: return launch_fmha_pair<2, 1, 0>(a, stream, ws_query);
  }
}

}  // namespace v3a
