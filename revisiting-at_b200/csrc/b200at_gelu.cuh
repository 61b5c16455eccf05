// Exact-erf GELU (models/convnext.py:31 nn.GELU()) and its derivative for bf16 activations.
// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the 2^-9 relative step of the bf16 result):
// one MUFU.RCP + one MUFU.EX2 + ~10 FMAs instead of erff()'s branchy ~30-instruction path.  The elementwise
// GELU passes and the fused GEMM epilogues were ALU-bound on erff, not HBM-bound (profiles/r01_gemm_bench_v2.txt).
#pragma once
#include <cuda_runtime.h>

// returns erf(v / sqrt(2)); *e = exp(-v*v/2)
__device__ __forceinline__ float b200at_erf_half(float v, float* e) {
  const float x = fabsf(v) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  *e = __expf(-x * x);
  const float y = 1.0f - p * t * (*e);
  return copysignf(y, v);
}
__device__ __forceinline__ float b200at_gelu(float v) {
  float e;
  return 0.5f * v * (1.0f + b200at_erf_half(v, &e));
}
// d/dv [ v * Phi(v) ] = Phi(v) + v * phi(v)
__device__ __forceinline__ float b200at_gelu_grad(float v) {
  float e;
  const float cdf = 0.5f * (1.0f + b200at_erf_half(v, &e));
  return fmaf(v * 0.3989422804014327f, e, cdf);
}
