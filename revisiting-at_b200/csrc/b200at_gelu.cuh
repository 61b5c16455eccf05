// Exact-erf GELU (models/convnext.py:31 nn.GELU()) and its derivative for bf16 activations.
//
// The elementwise GELU passes and the fused GEMM epilogues are bound by the fp32 FMA pipe, not by HBM
// (profiles/r01_ops_bench_v5.txt: 19 FMA-pipe instructions per element at 2 cycles each == the measured time),
// so the formula is arranged for the fewest FMA-pipe instructions:
//
//   h(v)    = Phi(-|v|) = 0.5 erfc(|v| / sqrt 2) = t (a1/2 + t (a2/2 + ... )) 2^(-xs^2),   t = 1 / (1 + p |v| / sqrt 2),
//             xs = |v| sqrt(log2(e) / 2)                      (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7 on erf)
//   GELU(v) = max(v, 0) - |v| h
//   GELU'(v) = Phi(v) + v phi(v) = 0.5 + copysign(0.5 - h, v) + v 2^(-xs^2) / sqrt(2 pi)
//
// 10 FMA-pipe instructions + MUFU.RCP + MUFU.EX2 (the .ftz approx forms: no denormal fix-up code) per GELU, 14 per
// GELU'.  Error vs double precision: 3.3e-7 / 3.0e-7 absolute (checked over [-12, 12]), far below the 2^-9
// relative step of the bf16 results.
#pragma once
#include <cuda_runtime.h>

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
__device__ __forceinline__ float b200at_gelu(float v) {
  float e;
  const float ax = fabsf(v);                          // NaN propagates; +-inf gives NaN (inf * 0), finite bf16 is exact
  const float h = b200at_phi_tail(ax, &e);
  return fmaf(-ax, h, fmaxf(v, 0.0f));
}
// d/dv [ v * Phi(v) ] = Phi(v) + v * phi(v)
__device__ __forceinline__ float b200at_gelu_grad(float v) {
  float e;
  const float h = b200at_phi_tail(fabsf(v), &e);
  const float cdf = 0.5f + copysignf(0.5f - h, v);
  return fmaf(v * 0.3989422804014327f, e, cdf);
}
