// Exact-erf GELU (models/convnext.py:31 nn.GELU()) and its derivative for bf16 activations.
//
// The elementwise GELU passes and the fused GEMM epilogues are bound by the fp32 FMA pipe, not by HBM
// (profiles/r01_ops_bench_v5.txt: 19 FMA-pipe instructions per element at 2 cycles each == the measured time),
// so the formula is arranged for the fewest FMA-pipe instructions:
//
//   h(v)    = Phi(-|v|) = 0.5 erfc(|v| / sqrt 2) = t (a1/2 + t (a2/2 + ... )) 2^(-xs^2),   t = 1 / (1 + p |v| / sqrt 2),
//             xs = |v| sqrt(log2(e) / 2)                      (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7 on erf)
//   GELU(v) = max(v, 0) - |v| h             (forward: h from the 2^polynomial form below, not from the A&S expression)
//   GELU'(v) = Phi(v) + v phi(v) = 0.5 + copysign(0.5 - h, v) + v 2^(-xs^2) / sqrt(2 pi)
//
// 10 FMA-pipe instructions + MUFU.RCP + MUFU.EX2 (the .ftz approx forms: no denormal fix-up code) per GELU, 14 per
// GELU'.  Error vs double precision: 3.3e-7 / 3.0e-7 absolute (checked over [-12, 12]), far below the 2^-9
// relative step of the bf16 results.
//
// Forward GELU, round 2: Phi(-|v|) = 2^P7(|v|), P7 = the degree-7 weighted least-squares fit of log2 Phi(-x) on [0, 6.5]
// (input clamped there: Phi(-6.5) = 4e-11).  7 FMAs + ONE MUFU.EX2 instead of 6 FMAs + 4 multiplies + MUFU.RCP + MUFU.EX2:
// the XU pipe (16 lanes / clk / SM) was the busiest pipe of the GELU epilogues.  Error of GELU(v) against double precision:
// 9.6e-8 absolute over [-12, 12] (the A&S form: 2.1e-7; profiles/fit_gelu_poly.py prints the coefficients and both numbers).  The derivative keeps the A&S form (it needs exp(-v^2/2) anyway).
//
// b200at_gelu2 / b200at_gelu_grad2 evaluate two elements with the packed fp32x2 forms (sm_100 FFMA2 / FMUL2 / FADD2):
// the same operations in the same order on each lane, so the results are bit-identical to the scalar functions, at half
// the FMA-pipe instructions -- the MUFU pair per element is then what bounds a GELU epilogue.
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ float2 b200at_ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 b200at_fadd2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 b200at_fmul2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 b200at_dup2(float x) { return make_float2(x, x); }

__device__ __forceinline__ float b200at_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float b200at_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// h = Phi(-|v|); *e = exp(-v*v/2)
__device__ __forceinline__ float b200at_phi_tail(float ax, float* e) {
  const float xs = ax * 0.8493218f;
  const float t = b200at_rcp(fmaf(ax, 0.23164189f, 1.0f));
  float q = fmaf(0.5307027f, t, -0.72657603f);
  q = fmaf(q, t, 0.7107069f);
  q = fmaf(q, t, -0.14224836f);
  q = fmaf(q, t, 0.1274148f);
  *e = b200at_ex2(xs * -xs);
  return (q * t) * (*e);
}
// coefficients of P7 in t = -|v| (the sign flips folded in)
#define B200AT_GELU_P0 (-9.999995828e-01f)
#define B200AT_GELU_P1 (1.151118398e+00f)
#define B200AT_GELU_P2 (-4.591108263e-01f)
#define B200AT_GELU_P3 (5.278837308e-02f)
#define B200AT_GELU_P4 (7.496536244e-03f)
#define B200AT_GELU_P5 (4.811376275e-04f)
#define B200AT_GELU_P6 (-3.767985982e-05f)
#define B200AT_GELU_P7 (-6.787956409e-06f)
__device__ __forceinline__ float b200at_gelu(float v) {
  const float nax = -fabsf(v);                        // NaN propagates; +-inf gives NaN (inf * 0) or -inf * 4e-11
  const float t = fmaxf(nax, -6.5f);
  float p = fmaf(B200AT_GELU_P7, t, B200AT_GELU_P6);
  p = fmaf(p, t, B200AT_GELU_P5);
  p = fmaf(p, t, B200AT_GELU_P4);
  p = fmaf(p, t, B200AT_GELU_P3);
  p = fmaf(p, t, B200AT_GELU_P2);
  p = fmaf(p, t, B200AT_GELU_P1);
  p = fmaf(p, t, B200AT_GELU_P0);
  return fmaf(nax, b200at_ex2(p), fmaxf(v, 0.0f));
}
// d/dv [ v * Phi(v) ] = Phi(v) + v * phi(v)
__device__ __forceinline__ float b200at_gelu_grad(float v) {
  float e;
  const float h = b200at_phi_tail(fabsf(v), &e);
  const float cdf = 0.5f + copysignf(0.5f - h, v);
  return fmaf(v * 0.3989422804014327f, e, cdf);
}

// ---- two elements at a time (bit-identical to the scalar forms above).  nax = -|v|: the sign flips are folded into the
// constants, (-|v|)(-k) == |v| k exactly.
__device__ __forceinline__ float2 b200at_phi_tail2(float2 nax, float2* e) {
  const float2 xs = b200at_fmul2(nax, b200at_dup2(-0.8493218f));
  const float2 d = b200at_ffma2(nax, b200at_dup2(-0.23164189f), b200at_dup2(1.0f));
  const float2 t = make_float2(b200at_rcp(d.x), b200at_rcp(d.y));
  float2 q = b200at_ffma2(b200at_dup2(0.5307027f), t, b200at_dup2(-0.72657603f));
  q = b200at_ffma2(q, t, b200at_dup2(0.7107069f));
  q = b200at_ffma2(q, t, b200at_dup2(-0.14224836f));
  q = b200at_ffma2(q, t, b200at_dup2(0.1274148f));
  const float2 x2 = b200at_fmul2(xs, xs);
  *e = make_float2(b200at_ex2(-x2.x), b200at_ex2(-x2.y));
  return b200at_fmul2(b200at_fmul2(q, t), *e);
}
__device__ __forceinline__ float2 b200at_gelu2(float2 v) {
  const float2 nax = make_float2(-fabsf(v.x), -fabsf(v.y));
  const float2 t = make_float2(fmaxf(nax.x, -6.5f), fmaxf(nax.y, -6.5f));
  float2 p = b200at_ffma2(b200at_dup2(B200AT_GELU_P7), t, b200at_dup2(B200AT_GELU_P6));
  p = b200at_ffma2(p, t, b200at_dup2(B200AT_GELU_P5));
  p = b200at_ffma2(p, t, b200at_dup2(B200AT_GELU_P4));
  p = b200at_ffma2(p, t, b200at_dup2(B200AT_GELU_P3));
  p = b200at_ffma2(p, t, b200at_dup2(B200AT_GELU_P2));
  p = b200at_ffma2(p, t, b200at_dup2(B200AT_GELU_P1));
  p = b200at_ffma2(p, t, b200at_dup2(B200AT_GELU_P0));
  const float2 h = make_float2(b200at_ex2(p.x), b200at_ex2(p.y));
  return b200at_ffma2(nax, h, make_float2(fmaxf(v.x, 0.0f), fmaxf(v.y, 0.0f)));
}
__device__ __forceinline__ float2 b200at_gelu_grad2(float2 v) {
  float2 e;
  const float2 nax = make_float2(-fabsf(v.x), -fabsf(v.y));
  const float2 h = b200at_phi_tail2(nax, &e);
  const float2 r = b200at_ffma2(h, b200at_dup2(-1.0f), b200at_dup2(0.5f));        // 0.5 - h, one rounding
  const float2 cdf = b200at_fadd2(b200at_dup2(0.5f), make_float2(copysignf(r.x, v.x), copysignf(r.y, v.y)));
  return b200at_ffma2(b200at_fmul2(v, b200at_dup2(0.3989422804014327f)), e, cdf);
}
