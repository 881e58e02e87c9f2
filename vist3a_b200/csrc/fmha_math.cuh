// Packed-fp32 / exp2 helpers shared by the attention kernels (fmha_sm100.cu, fmha_pair_sm100.cu).
#pragma once
#include "common.cuh"

namespace v3a {

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {  // packed 2 x fp32 FADD2
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {  // packed 2 x fp32 FFMA2
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// exp2 of two values on the FMA / ALU pipes: y = n + f (n integer, |f| <= 0.5), 2^f by a degree-3 minimax polynomial
// (max rel. error 1.0e-4), 2^n by adding n to the exponent field.  y is clamped to >= -126.
__device__ __forceinline__ void exp2_poly2(float y0, float y1, float& e0, float& e1) {
  const uint64_t one = pack2(1.0f, 1.0f);
  const uint64_t magic = pack2(12582912.0f, 12582912.0f);      // 1.5 * 2^23: rounds to nearest integer
  const uint64_t nmagic = pack2(-12582912.0f, -12582912.0f);
  const uint64_t neg1 = pack2(-1.0f, -1.0f);
  const uint64_t yy = pack2(fmaxf(y0, -126.0f), fmaxf(y1, -126.0f));
  const uint64_t t = fma2(yy, one, magic);
  const uint64_t n = fma2(t, one, nmagic);
  const uint64_t f = fma2(n, neg1, yy);
  uint64_t q = fma2(f, pack2(0.055008664727211f, 0.055008664727211f), pack2(0.24221056699752808f, 0.24221056699752808f));
  q = fma2(q, f, pack2(0.6932829022407532f, 0.6932829022407532f));
  q = fma2(q, f, one);
  float t0, t1, q0, q1;
  unpack2(t, t0, t1);
  unpack2(q, q0, q1);
  e0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// ---- speculative softmax half-step (one thread per query row, 64 score columns) ------------------------------------------------------------
// The exponentials are taken against the STALE running reference maximum m_run while the maximum of the 64 new scores is reduced in the issue
// slots the MUFU leaves free; the caller checks afterwards whether that maximum exceeds m_run by more than the lazy-rescale threshold (rare)
// and only then rescales O / the row sum and recomputes the half-step from the intact scores.  This takes the load -> max -> compare chain
// (240 cycles per step in the traced kernels) off the critical path of every step.
//   r[64]  raw scores;  cc2 = packed (c, c), c = scale * log2(e) [* row scale];  mc2 = packed (-m_run c, -m_run c)
//   pk[32] packed bf16 pairs of exp2((s - m_run) c);  hsum[2] packed partial sums of this half-step (added to);  returns max(r)
// NP of every 8 column pairs take exp2 on the FMA pipe (Cody-Waite + cubic); they are spread over the group so that ptxas finds FMA-pipe
// work for the issue slots between MUFU instructions (a warp's MUFU.EX2 issues once per 8 cycles: 4 lanes per clock per SM sub-partition).
template <int NP>
__device__ __forceinline__ constexpr bool exp_pair_is_poly(int k) {
  return NP >= 4 ? (k & 1) : NP == 3 ? (k == 1 || k == 4 || k == 7) : NP == 2 ? (k == 2 || k == 6) : NP == 1 ? (k == 4) : false;
}
// exp2 of two values on the FMA pipe with the lower clamp bound passed in (see exp_half64 for why it is a register)
template <bool kClampHigh = false>
__device__ __forceinline__ void exp2_poly2_b(float y0, float y1, float bound, float& e0, float& e1) {
  // kClampHigh: callers that detect an exponent above the lazy-rescale threshold from the row SUM (no running maximum) need the polynomial path
  // to return something huge for y > 127 instead of a wrapped exponent field (2^127 sends the sum over any threshold; the MUFU path gives +inf)
  if (kClampHigh) { y0 = fminf(y0, 127.0f); y1 = fminf(y1, 127.0f); }
  const uint64_t one = pack2(1.0f, 1.0f);
  const uint64_t magic = pack2(12582912.0f, 12582912.0f);      // 1.5 * 2^23: rounds to nearest integer
  const uint64_t nmagic = pack2(-12582912.0f, -12582912.0f);
  const uint64_t neg1 = pack2(-1.0f, -1.0f);
  const uint64_t yy = pack2(fmaxf(y0, bound), fmaxf(y1, bound));
  const uint64_t t = fma2(yy, one, magic);
  const uint64_t n = fma2(t, one, nmagic);
  const uint64_t f = fma2(n, neg1, yy);
  uint64_t q = fma2(f, pack2(0.055008664727211f, 0.055008664727211f), pack2(0.24221056699752808f, 0.24221056699752808f));
  q = fma2(q, f, pack2(0.6932829022407532f, 0.6932829022407532f));
  q = fma2(q, f, one);
  float t0, t1, q0, q1;
  unpack2(t, t0, t1);
  unpack2(q, q0, q1);
  e0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
// `zero` is a run-time zero (a kernel parameter ptxas cannot fold).  ptxas schedules the long polynomial chains of ALL groups of the unrolled
// half-step first and leaves a MUFU-only tail (measured: 570 static cycles per half-step against 384 of MUFU time); to stagger them, the clamp
// bound of group g's polynomial pairs is made to depend on a MUFU result from the middle of group g - 1 (bound | (e & zero): one LOP3), so
// that every group's FMA-pipe work lands in the issue slots between the MUFU instructions of its own neighbourhood.
// TRACK = false drops the maximum (returns -inf): the caller then tests the half-step's row sum instead (a sum above 2^8 or a non-finite one
// means that some exponent was above the lazy-rescale threshold, or several were close to it)
template <int NP, bool TRACK = true>
__device__ __forceinline__ float exp_half64(const uint32_t* r, uint64_t cc2, uint64_t mc2, uint64_t* hsum, uint32_t* pk, uint32_t zero) {
  float mx0 = -INFINITY, mx1 = -INFINITY;
  uint32_t dep = 0u;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float bound = __uint_as_float(0xc2fc0000u | (dep & zero));   // -126.0f
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float s0 = __uint_as_float(r[g * 16 + 2 * k]), s1 = __uint_as_float(r[g * 16 + 2 * k + 1]);
      if (TRACK) {
        if (k & 1) mx1 = fmaxf(fmaxf(mx1, s0), s1);
        else mx0 = fmaxf(fmaxf(mx0, s0), s1);
      }
      float y0, y1, e0, e1;
      unpack2(fma2(pack2(s0, s1), cc2, mc2), y0, y1);
      if (exp_pair_is_poly<NP>(k)) {
        exp2_poly2_b<!TRACK>(y0, y1, bound, e0, e1);
      } else {
        e0 = ex2_approx(y0);
        e1 = ex2_approx(y1);
        if (k == 3 || (NP >= 3 && k == 2)) dep = __float_as_uint(e0);
      }
      hsum[k & 1] = add2(hsum[k & 1], pack2(e0, e1));
      pk[g * 8 + k] = pack_bf16(e0, e1);
    }
  }
  return fmaxf(mx0, mx1);
}
// Second version of the half-step: (1) no running maximum in the fast path -- the caller triggers its slow path when the half-step's row sum
// exceeds 2^8 (some exponent was above the lazy-rescale threshold, or several came close) or is not finite; (2) the consumers of a MUFU result
// (row sum, bf16 pack) are held back by one group of 8 column pairs.  A MUFU.EX2 result returns ~40 cycles after issue, ptxas places its
// consumer one pair (16 cycles) behind, and a warp that runs alone on its SM sub-partition then stalls on every pair: 1285 cycles per half-step
// measured against 512 of MUFU time (tools/ubench/exp_half64.cu).  The consumers of group g are therefore made to depend (e | (x & zero), one
// LOP3 per pair) on the first MUFU result of group g + 1, which puts them into the issue slots between that group's later MUFU instructions.
template <int NP>
__device__ __forceinline__ void exp_half64_v2(const uint32_t* r, uint64_t cc2, uint64_t mc2, uint64_t* hsum, uint32_t* pk, uint32_t zero) {
  float e[2][16];
  uint32_t dep_poly = 0u;
#pragma unroll
  for (int g = 0; g <= 4; ++g) {
    uint32_t dep = 0u;
    if (g < 4) {
      const float bound = __uint_as_float(0xc2fc0000u | (dep_poly & zero));   // -126.0f
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float y0, y1;
        unpack2(fma2(pack2(__uint_as_float(r[g * 16 + 2 * k]), __uint_as_float(r[g * 16 + 2 * k + 1])), cc2, mc2), y0, y1);
        float& e0 = e[g & 1][2 * k];
        float& e1 = e[g & 1][2 * k + 1];
        if (exp_pair_is_poly<NP>(k)) {
          exp2_poly2_b<true>(y0, y1, bound, e0, e1);
        } else {
          e0 = ex2_approx(y0);
          e1 = ex2_approx(y1);
        }
      }
      // first MUFU result of this group (pair 0 is never a polynomial pair)
      dep = __float_as_uint(e[g & 1][0]) & zero;
      dep_poly = __float_as_uint(e[g & 1][NP >= 4 ? 4 : 6]);   // pair 2 / 3: a MUFU pair for the given NP
    }
    if (g >= 1) {
      const int h = (g - 1) & 1;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float e0 = __uint_as_float(__float_as_uint(e[h][2 * k]) | dep), e1 = e[h][2 * k + 1];
        hsum[k & 1] = add2(hsum[k & 1], pack2(e0, e1));
        pk[(g - 1) * 8 + k] = pack_bf16(e0, e1);
      }
    }
  }
}
// the same half-step on the MUFU only, against a reference maximum that is known to cover the scores (slow path / first step)
__device__ __forceinline__ void exp_half64_exact(const uint32_t* r, uint64_t cc2, uint64_t mc2, uint64_t* hsum, uint32_t* pk) {
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    float y0, y1;
    unpack2(fma2(pack2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1])), cc2, mc2), y0, y1);
    const float e0 = ex2_approx(y0), e1 = ex2_approx(y1);
    hsum[k & 1] = add2(hsum[k & 1], pack2(e0, e1));
    pk[k] = pack_bf16(e0, e1);
  }
}

}  // namespace v3a
