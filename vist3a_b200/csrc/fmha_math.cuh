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

}  // namespace v3a
