// K11 (first layer): the CvSt stem's first stage as ONE kernel per direction, for the attack's evaluations.
//
//   forward     y = GELU(LN_C0(conv3x3_s2_p1((x - mean) / std) + bias))      x fp32 NCHW in [0,1] -> y NHWC bf16
//   input-grad  dx = d/dx of the above, given dy (NHWC bf16)                  -> dx fp32 NCHW
//
// Reference: utils_architecture.py:198-217 (ConvBlock1) / :174-195 (ConvBlock3) first conv + channels-first
// LayerNorm (:57-81) + GELU, behind ImageNormalizer (:86-98).  With 3 input channels this layer is memory /
// FMA-issue bound, not a tensor-core GEMM (27 MACs per output value): one thread owns one output pixel and all
// C0 channels of it, so the LayerNorm statistics need no shuffles, the 3x3x3 input window lives in registers,
// the weights are broadcast from shared memory as 16-byte loads feeding packed fp32x2 FMAs, and the NHWC
// row of the pixel leaves through a padded shared-memory tile as full 128-byte lines.
// The input gradient recomputes the pre-LN activation from x instead of reading a saved copy (the layer's
// output is the largest activation of the network: 154 MB per forward at batch 128), pushes dy through
// GELU' and the LayerNorm backward in registers, contracts with the weights to the 27 window gradients of
// the pixel, and a second phase gathers the (up to 4) overlapping windows of every input pixel of the
// CTA's 32x32 input tile from shared memory -- no atomics, deterministic.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "b200at_gelu.cuh"
#include "../../include/b200at_model.h"

namespace {

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

struct StemParams {
  const float* x;       // [B][3][H][W] fp32
  const float* wk;      // [27][C0] fp32, k = c*9 + kh*3 + kw
  const float* bias;    // [C0]
  const float* ln_w;    // [C0]
  const float* ln_b;    // [C0]
  float mean[3], inv_std[3];
  int B, H, W, Ho, Wo;
  float eps;
};

// the pixel's normalised 3x3x3 window (zero outside the image: padding is applied after normalisation)
__device__ __forceinline__ void load_window(const StemParams& p, int n, int ho, int wo, float* v) {
  const float* xn = p.x + (int64_t)n * 3 * p.H * p.W;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = 2 * ho - 1 + kh;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int w = 2 * wo - 1 + kw;
        float t = 0.f;
        if (h >= 0 && h < p.H && w >= 0 && w < p.W) t = (__ldg(xn + ((int64_t)c * p.H + h) * p.W + w) - p.mean[c]) * p.inv_std[c];
        v[c * 9 + kh * 3 + kw] = t;
      }
    }
  }
}

// acc[j] = bias + sum_k v[k] * wk[k][2j..2j+1]   (weights broadcast from shared memory, 16 bytes per load)
template <int C0>
__device__ __forceinline__ void conv_pixel(const float* __restrict__ wsm, const float* __restrict__ bsm, const float* v,
                                           float2* acc) {
#pragma unroll
  for (int j = 0; j < C0 / 2; ++j) acc[j] = make_float2(bsm[2 * j], bsm[2 * j + 1]);
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    const float2 vv = make_float2(v[k], v[k]);
    const float4* wr = reinterpret_cast<const float4*>(wsm + k * C0);
#pragma unroll
    for (int j = 0; j < C0 / 4; ++j) {
      const float4 w4 = wr[j];
      acc[2 * j] = ffma2(vv, make_float2(w4.x, w4.y), acc[2 * j]);
      acc[2 * j + 1] = ffma2(vv, make_float2(w4.z, w4.w), acc[2 * j + 1]);
    }
  }
}

template <int C0>
__device__ __forceinline__ void ln_stats(const float2* acc, float eps, float& mu, float& rs) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < C0 / 2; ++j) s += acc[j].x + acc[j].y;
  mu = s * (1.0f / C0);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < C0 / 2; ++j) {
    const float a = acc[j].x - mu, b = acc[j].y - mu;
    q += a * a + b * b;
  }
  rs = rsqrtf(q * (1.0f / C0) + eps);
}

constexpr int kFwdThreads = 128;

template <int C0>
__global__ void __launch_bounds__(kFwdThreads) stem0_fwd_kernel(const StemParams p, bf16* __restrict__ y) {
  constexpr int kPitch = C0 * 2 + 16;                       // bytes per staged pixel row (pad: conflict-free 16 B stores)
  __shared__ __align__(16) float wsm[27 * C0];
  __shared__ __align__(16) float bsm[C0], lw[C0], lb[C0];
  __shared__ __align__(16) uint8_t stage[kFwdThreads * kPitch];
  for (int i = threadIdx.x; i < 27 * C0; i += kFwdThreads) wsm[i] = p.wk[i];
  for (int i = threadIdx.x; i < C0; i += kFwdThreads) { bsm[i] = p.bias[i]; lw[i] = p.ln_w[i]; lb[i] = p.ln_b[i]; }
  __syncthreads();
  const int64_t total = (int64_t)p.B * p.Ho * p.Wo;
  const int64_t pix0 = (int64_t)blockIdx.x * kFwdThreads;
  const int64_t pix = pix0 + threadIdx.x;
  if (pix < total) {
    const int wo = (int)(pix % p.Wo);
    const int ho = (int)((pix / p.Wo) % p.Ho);
    const int n = (int)(pix / ((int64_t)p.Wo * p.Ho));
    float v[27];
    load_window(p, n, ho, wo, v);
    float2 acc[C0 / 2];
    conv_pixel<C0>(wsm, bsm, v, acc);
    float mu, rs;
    ln_stats<C0>(acc, p.eps, mu, rs);
    uint8_t* row = stage + threadIdx.x * kPitch;
#pragma unroll
    for (int j = 0; j < C0 / 2; j += 4) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = 2 * (j + i);
        const float a = b200at_gelu((acc[j + i].x - mu) * rs * lw[c] + lb[c]);
        const float b = b200at_gelu((acc[j + i].y - mu) * rs * lw[c + 1] + lb[c + 1]);
        bf162 t = __floats2bfloat162_rn(a, b);
        w[i] = *reinterpret_cast<uint32_t*>(&t);
      }
      *reinterpret_cast<uint4*>(row + j * 4) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  __syncthreads();
  // the CTA's pixels are consecutive NHWC rows: one contiguous run of (live pixels) * C0 * 2 bytes
  const int64_t live = (total - pix0) < kFwdThreads ? (total - pix0) : kFwdThreads;
  constexpr int kPieces = C0 * 2 / 16;                      // 16-byte pieces per pixel
  uint4* dst = reinterpret_cast<uint4*>(y + pix0 * C0);
  for (int q = threadIdx.x; q < (int)live * kPieces; q += kFwdThreads) {
    const int r = q / kPieces, piece = q % kPieces;
    dst[q] = *reinterpret_cast<const uint4*>(stage + r * kPitch + piece * 16);
  }
}

// ---------------------------------------------------------------------------------------------- input gradient
constexpr int kBwdTileOut = 16;                 // output pixels per tile side (input tile: 32 x 32)
constexpr int kBwdSide = kBwdTileOut + 1;       // + one halo row/column of output pixels
constexpr int kBwdPix = kBwdSide * kBwdSide;    // 289
constexpr int kBwdThreads = 320;

template <int C0>
__global__ void __launch_bounds__(kBwdThreads) stem0_bwd_kernel(const StemParams p, const bf16* __restrict__ dy,
                                                                float* __restrict__ dx, int tiles_w, int tiles_h) {
  __shared__ __align__(16) float wsm[27 * C0];
  __shared__ __align__(16) float bsm[C0], lw[C0], lb[C0];
  __shared__ float vbuf[kBwdPix][27];           // stride 27 words: conflict-free for the per-pixel rows
  for (int i = threadIdx.x; i < 27 * C0; i += kBwdThreads) wsm[i] = p.wk[i];
  for (int i = threadIdx.x; i < C0; i += kBwdThreads) { bsm[i] = p.bias[i]; lw[i] = p.ln_w[i]; lb[i] = p.ln_b[i]; }
  __syncthreads();
  int bid = blockIdx.x;
  const int tw = bid % tiles_w; bid /= tiles_w;
  const int th = bid % tiles_h; bid /= tiles_h;
  const int n = bid;
  const int ho0 = th * kBwdTileOut, wo0 = tw * kBwdTileOut;
  if (threadIdx.x < kBwdPix) {
    const int lr = threadIdx.x / kBwdSide, lc = threadIdx.x % kBwdSide;
    const int ho = ho0 + lr, wo = wo0 + lc;
    float* vrow = vbuf[threadIdx.x];
    if (ho < p.Ho && wo < p.Wo) {
      float v[27];
      load_window(p, n, ho, wo, v);
      float2 acc[C0 / 2];
      conv_pixel<C0>(wsm, bsm, v, acc);
      float mu, rs;
      ln_stats<C0>(acc, p.eps, mu, rs);
      // g = dy * GELU'(pre) * ln_w ; dconv = rs * (g - mean(g) - xhat * mean(g * xhat))   (acc <- xhat, then dconv)
      const uint4* dyr = reinterpret_cast<const uint4*>(dy + (((int64_t)n * p.Ho + ho) * p.Wo + wo) * C0);
      float2 g[C0 / 2];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < C0 / 2; j += 4) {
        const uint4 u = __ldg(dyr + j / 4);
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 2 * (j + i);
          const float2 d = __bfloat1622float2(*reinterpret_cast<const bf162*>(&w4[i]));
          const float xa = (acc[j + i].x - mu) * rs, xb = (acc[j + i].y - mu) * rs;
          const float ga = d.x * b200at_gelu_grad(xa * lw[c] + lb[c]) * lw[c];
          const float gb = d.y * b200at_gelu_grad(xb * lw[c + 1] + lb[c + 1]) * lw[c + 1];
          acc[j + i] = make_float2(xa, xb);
          g[j + i] = make_float2(ga, gb);
          s1 += ga + gb;
          s2 += ga * xa + gb * xb;
        }
      }
      const float m1 = s1 * (1.0f / C0), m2 = s2 * (1.0f / C0);
#pragma unroll
      for (int j = 0; j < C0 / 2; ++j) {
        g[j].x = rs * (g[j].x - m1 - acc[j].x * m2);
        g[j].y = rs * (g[j].y - m1 - acc[j].y * m2);
      }
      // window gradient: vrow[k] = sum_co dconv[co] * wk[k][co].  Rolled over k (3 at a time): fully unrolled,
      // ptxas hoisted all 27*C0 weight loads ahead of the FMAs and spilled kilobytes per thread.
#pragma unroll 1
      for (int k = 0; k < 27; k += 3) {
        float2 a[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const float4* wr = reinterpret_cast<const float4*>(wsm + (k + u) * C0);
          a[u] = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < C0 / 4; ++j) {
            const float4 w4 = wr[j];
            a[u] = ffma2(g[2 * j], make_float2(w4.x, w4.y), a[u]);
            a[u] = ffma2(g[2 * j + 1], make_float2(w4.z, w4.w), a[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) vrow[k + u] = a[u].x + a[u].y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 27; ++k) vrow[k] = 0.f;
    }
  }
  __syncthreads();
  // gather: input pixel (h, w) of the 32x32 tile receives window entry (kh, kw) of output pixel
  // ((h + 1 - kh) / 2, (w + 1 - kw) / 2) whenever those are integers
  const int h0 = 2 * ho0, w0 = 2 * wo0;
  float* dxn = dx + (int64_t)n * 3 * p.H * p.W;
  for (int q = threadIdx.x; q < 3 * 32 * 32; q += kBwdThreads) {
    const int lw_ = q & 31, lh = (q >> 5) & 31, c = q >> 10;
    const int h = h0 + lh, w = w0 + lw_;
    if (h >= p.H || w >= p.W) continue;
    float s = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int th2 = lh + 1 - kh;                 // = 2 * (local output row)
      if (th2 < 0 || (th2 & 1)) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tw2 = lw_ + 1 - kw;
        if (tw2 < 0 || (tw2 & 1)) continue;
        s += vbuf[(th2 >> 1) * kBwdSide + (tw2 >> 1)][c * 9 + kh * 3 + kw];
      }
    }
    dxn[((int64_t)c * p.H + h) * p.W + w] = s * p.inv_std[c];
  }
}

bool fill(StemParams& p, const float* x, const float* mean3, const float* std3, const float* wk, const float* bias,
          const float* ln_w, const float* ln_b, int64_t B, int64_t H, int64_t W, float eps) {
  p.x = x; p.wk = wk; p.bias = bias; p.ln_w = ln_w; p.ln_b = ln_b;
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = mean3 ? mean3[c] : 0.f;
    const float sd = std3 ? std3[c] : 1.f;
    if (!(sd > 0.f)) return false;
    p.inv_std[c] = 1.0f / sd;
  }
  p.B = (int)B; p.H = (int)H; p.W = (int)W;
  p.Ho = (int)((H - 1) / 2 + 1); p.Wo = (int)((W - 1) / 2 + 1);
  p.eps = eps;
  return true;
}


// ImageNormalizer + cast + layout change of the TRAINING forward's first stem stage (utils_architecture.py:86-98 in
// front of the library convolution): y[b][h][w][c] = bf16((x[b][c][h][w] - mean[c]) / std[c]).  One pass (12 B read,
// 6 B written per pixel) instead of torch's sub, div, cast and channels_last copy (4 passes over the batch).
// Same fp32 arithmetic as the eager expression (IEEE subtraction and division, then one rounding to bf16).
struct NormParams { float mean[3], std[3]; };
__global__ void __launch_bounds__(256) normalize_nhwc_kernel(const float* __restrict__ x, bf16* __restrict__ y,
                                                             const NormParams np, int64_t hw, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t b = i / hw, pix = i - b * hw;
    const float* src = x + b * 3 * hw + pix;
    bf16* dst = y + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      dst[c] = __float2bfloat16_rn(__fdiv_rn(__fsub_rn(__ldcs(src + c * hw), np.mean[c]), np.std[c]));
  }
}

}  // namespace

extern "C" {

int b200at_stem0_fwd(const float* x, const float* mean3, const float* std3, const float* wk, const float* bias,
                     const float* ln_w, const float* ln_b, void* y, int64_t B, int64_t H, int64_t W, int64_t C0,
                     float eps, void* stream) {
  if (B <= 0) return 0;
  StemParams p;
  if (H < 1 || W < 1 || !fill(p, x, mean3, std3, wk, bias, ln_w, ln_b, B, H, W, eps)) return (int)cudaErrorInvalidValue;
  const int64_t total = (int64_t)p.B * p.Ho * p.Wo;
  const int64_t grid = (total + kFwdThreads - 1) / kFwdThreads;
  if (grid > 0x7fffffff) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  switch (C0) {
    case 48: stem0_fwd_kernel<48><<<(unsigned)grid, kFwdThreads, 0, s>>>(p, (bf16*)y); break;
    case 64: stem0_fwd_kernel<64><<<(unsigned)grid, kFwdThreads, 0, s>>>(p, (bf16*)y); break;
    case 96: stem0_fwd_kernel<96><<<(unsigned)grid, kFwdThreads, 0, s>>>(p, (bf16*)y); break;
    default: return (int)cudaErrorInvalidValue;
  }
  return (int)cudaGetLastError();
}

int b200at_stem0_bwd_input(const void* dy, const float* x, const float* mean3, const float* std3, const float* wk,
                           const float* bias, const float* ln_w, const float* ln_b, float* dx, int64_t B, int64_t H,
                           int64_t W, int64_t C0, float eps, void* stream) {
  if (B <= 0) return 0;
  StemParams p;
  if (H < 1 || W < 1 || !fill(p, x, mean3, std3, wk, bias, ln_w, ln_b, B, H, W, eps)) return (int)cudaErrorInvalidValue;
  // tiles must cover every INPUT pixel: input row h belongs to the tile of output row h / 2
  const int tiles_h = (int)((H + 2 * kBwdTileOut - 1) / (2 * kBwdTileOut));
  const int tiles_w = (int)((W + 2 * kBwdTileOut - 1) / (2 * kBwdTileOut));
  const int64_t grid = B * tiles_h * tiles_w;
  if (grid > 0x7fffffff) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  switch (C0) {
    case 48: stem0_bwd_kernel<48><<<(unsigned)grid, kBwdThreads, 0, s>>>(p, (const bf16*)dy, dx, tiles_w, tiles_h); break;
    case 64: stem0_bwd_kernel<64><<<(unsigned)grid, kBwdThreads, 0, s>>>(p, (const bf16*)dy, dx, tiles_w, tiles_h); break;
    case 96: stem0_bwd_kernel<96><<<(unsigned)grid, kBwdThreads, 0, s>>>(p, (const bf16*)dy, dx, tiles_w, tiles_h); break;
    default: return (int)cudaErrorInvalidValue;
  }
  return (int)cudaGetLastError();
}

int b200at_normalize_nhwc_bf16(const float* x, const float* mean3, const float* std3, void* y, int64_t B, int64_t H,
                               int64_t W, void* stream) {
  if (B <= 0) return 0;
  NormParams np;
  for (int c = 0; c < 3; ++c) {
    np.mean[c] = mean3 ? mean3[c] : 0.f;
    np.std[c] = std3 ? std3[c] : 1.f;
    if (!(np.std[c] > 0.f)) return (int)cudaErrorInvalidValue;
  }
  const int64_t hw = H * W, total = B * hw;
  int64_t grid = (total + 255) / 256;
  if (grid > 148 * 32) grid = 148 * 32;
  normalize_nhwc_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y, np, hw, total);
  return (int)cudaGetLastError();
}

}  // extern "C"
