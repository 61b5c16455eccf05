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
#include <stdlib.h>

#include "b200at_gelu.cuh"
#include "b200at_launch.cuh"
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

// SAVE (the training forward): also the convolution output WITHOUT its bias (bf16, what the unfused path hands to the
// LayerNorm kernels as `x` with `pre_bias`) and the LayerNorm statistics, so that the backward can stay b200at_ln_bwd_bias +
// the convolution's weight gradient.
template <int C0, bool SAVE = false>
__global__ void __launch_bounds__(kFwdThreads) stem0_fwd_kernel(const StemParams p, bf16* __restrict__ y,
                                                                bf16* __restrict__ y_pre = nullptr,
                                                                float* __restrict__ mean_out = nullptr,
                                                                float* __restrict__ rstd_out = nullptr) {
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
  float2 acc[C0 / 2];
  uint8_t* row = stage + threadIdx.x * kPitch;
  if (pix < total) {
    const int wo = (int)(pix % p.Wo);
    const int ho = (int)((pix / p.Wo) % p.Ho);
    const int n = (int)(pix / ((int64_t)p.Wo * p.Ho));
    float v[27];
    load_window(p, n, ho, wo, v);
    conv_pixel<C0>(wsm, bsm, v, acc);
    float mu, rs;
    ln_stats<C0>(acc, p.eps, mu, rs);
    if (SAVE) { mean_out[pix] = mu; rstd_out[pix] = rs; }
#pragma unroll
    for (int j = 0; j < C0 / 2; j += 4) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = 2 * (j + i);
        const float2 ab = b200at_gelu2(make_float2((acc[j + i].x - mu) * rs * lw[c] + lb[c],
                                                   (acc[j + i].y - mu) * rs * lw[c + 1] + lb[c + 1]));
        bf162 t = __floats2bfloat162_rn(ab.x, ab.y);
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
  if (SAVE) {
    __syncthreads();
    if (pix < total) {
#pragma unroll
      for (int j = 0; j < C0 / 2; j += 4) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 2 * (j + i);
          bf162 t = __floats2bfloat162_rn(acc[j + i].x - bsm[c], acc[j + i].y - bsm[c + 1]);
          w[i] = *reinterpret_cast<uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(row + j * 4) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    __syncthreads();
    uint4* dst2 = reinterpret_cast<uint4*>(y_pre + pix0 * C0);
    for (int q = threadIdx.x; q < (int)live * kPieces; q += kFwdThreads) {
      const int r = q / kPieces, piece = q % kPieces;
      dst2[q] = *reinterpret_cast<const uint4*>(stage + r * kPitch + piece * 16);
    }
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
          const float2 gp = b200at_gelu_grad2(make_float2(xa * lw[c] + lb[c], xb * lw[c + 1] + lb[c + 1]));
          const float ga = d.x * gp.x * lw[c];
          const float gb = d.y * gp.y * lw[c + 1];
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

// ================================================================================================ tensor-core form
// Second design of the two kernels above.  With one thread per output pixel the 27 x C0 contraction is 27 * C0 / 2 packed
// FMAs per pixel and direction, each pair fed by a broadcast LDS.128 of weights: FMA pipe and shared-memory wavefronts
// tie (profiles/r01_ncu_dwconv_stem0_v11_summary.txt), 143 registers allow one 10-warp CTA per SM, and the input
// gradient ran at 0.06 of the HBM rate (755 us for 308 MB).  Here the contraction is mma.sync.m16n8k16 on bf16 pairs:
//   forward      U[16 pixels][C0] = X[16][27 -> 32] . Wk[32][C0]          2 k-steps x C0/8 n-tiles
//   input grad   V[16 pixels][27 -> 32] = dU[16][C0] . Wk^T[C0][32]       C0/16 k-steps x 4 n-tiles
// and the accumulator fragment of U (pixel g / g + 8; channels 8t + 2q, 8t + 2q + 1) is exactly the layout the LayerNorm
// statistics (12 values per lane + two quad shuffles), the GELU / GELU' tail, the dy loads (4-byte, 16 contiguous bytes
// per quad) and -- repacked in place as bf16 pairs -- the A operand of the second product want.  To keep the fp32
// quality of the scalar kernels (x_adv moves by 4/255: a bf16 x would quantise the perturbation to ~12 % of eps) both
// operands enter as hi + lo bf16 splits and every product is three MMAs (hi.hi + lo.hi + hi.lo, error 2^-16 relative).
// What is left per pixel is the elementwise tail (~22 instructions per channel in the backward), spread evenly over the
// lanes; weights sit in shared memory in fragment order (one conflict-free LDS.64 per MMA operand).
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint2 b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
// (a, b) -> bf16x2 of the values and bf16x2 of what the rounding dropped
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const bf162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const bf162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Window slots.  The contraction index is k = ((c * 3 + kh) * 4 + kw'), kw' = 0..3, padded to 48: kw' = 3 is a dummy
// with zero weight.  The input tile sits in shared memory as bf16 hi / lo planes whose column 0 is the image column
// 2 wo0 - 1, so the window of local output column wo starts at the EVEN column 2 wo and every A register -- the pair
// (k, k + 1) = (kw' 0, 1) or (kw' 2, 3) of one (c, kh) -- is one aligned 32-bit shared-memory load (the first version
// gathered the 27 taps from global memory with bounds checks: 25 instructions per tap, more than the FMAs it replaced).
constexpr int kStemSlots = 36;
__host__ __device__ constexpr int stem_tap_of_slot(int k) {      // wk row (c * 9 + kh * 3 + kw) or -1
  return (k < kStemSlots && (k & 3) < 3) ? (k >> 2) * 3 + (k & 3) : -1;
}

// shared-memory weight tables in fragment order: Bf[s][t][lane] (forward: k-step s of 3, channel tile t) and
// Bt[s][t][lane] (input gradient: channel k-step s, slot tile t of 5), each a uint2 of the lane's two B registers
template <int C0>
struct StemTc {
  static constexpr int NT = C0 / 8;          // channel tiles of the forward
  static constexpr int KS = C0 / 16;         // channel k-steps of the input gradient
  static constexpr int kBf = 3 * NT * 32, kBt = KS * 5 * 32;
};

template <int C0>
__device__ __forceinline__ void build_tables(const float* __restrict__ wk, uint2* bf_hi, uint2* bf_lo, uint2* bt_hi,
                                             uint2* bt_lo, int nthreads) {
  typedef StemTc<C0> T;
  for (int i = threadIdx.x; i < T::kBf; i += nthreads) {
    const int lane = i & 31, t = (i >> 5) % T::NT, s = (i >> 5) / T::NT;
    const int g = lane >> 2, q = lane & 3, co = 8 * t + g, k0 = 16 * s + 2 * q;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int tap = stem_tap_of_slot(k0 + (j & 1) + 8 * (j >> 1));
      v[j] = tap >= 0 ? wk[tap * C0 + co] : 0.f;
    }
    split2(v[0], v[1], bf_hi[i].x, bf_lo[i].x);
    split2(v[2], v[3], bf_hi[i].y, bf_lo[i].y);
  }
  if (bt_hi) {
    for (int i = threadIdx.x; i < T::kBt; i += nthreads) {
      const int lane = i & 31, t = (i >> 5) % 5, s = (i >> 5) / 5;
      const int g = lane >> 2, q = lane & 3, tap = stem_tap_of_slot(8 * t + g), c0 = 16 * s + 2 * q;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = tap >= 0 ? wk[tap * C0 + c0 + (j & 1) + 8 * (j >> 1)] : 0.f;
      split2(v[0], v[1], bt_hi[i].x, bt_lo[i].x);
      split2(v[2], v[3], bt_hi[i].y, bt_lo[i].y);
    }
  }
}

// Normalised input tile as bf16 hi / lo planes: xt[hl][c][row][col], row 0 = image row h_base, column 0 = image column
// w_base (zero outside the image: the padding applies after normalisation), `pitch` bf16 per row (even)
__device__ __forceinline__ void fill_tile(const StemParams& p, const float* __restrict__ xn, int h_base, int w_base, int rows,
                                          int pitch, uint16_t* xt, int nthreads) {
  const int plane = rows * pitch;
  for (int i = threadIdx.x; i < 3 * plane; i += nthreads) {
    const int c = i / plane, r = (i - c * plane) / pitch, j = i - c * plane - r * pitch;
    const int h = h_base + r, w = w_base + j;
    float v = 0.f;
    if (h >= 0 && h < p.H && w >= 0 && w < p.W) {
      const float m = c == 0 ? p.mean[0] : (c == 1 ? p.mean[1] : p.mean[2]);
      const float is = c == 0 ? p.inv_std[0] : (c == 1 ? p.inv_std[1] : p.inv_std[2]);
      v = (__ldg(xn + ((int64_t)c * p.H + h) * p.W + w) - m) * is;
    }
    const bf16 hi = __float2bfloat16_rn(v);
    const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    xt[i] = *reinterpret_cast<const uint16_t*>(&hi);
    xt[3 * plane + i] = *reinterpret_cast<const uint16_t*>(&lo);
  }
}

// The lane's 6 A-register slots: pair index kp = 8 s + q (+ 4), s = 0..2; (c, kh) = kp / 2, columns 2 (kp % 2), +1.
// off = bf16 offset of the pair inside the tile relative to the window origin, or -1 (padding: zero register).
struct StemPairs { int off[6]; };
__device__ __forceinline__ StemPairs make_pairs(int q, int rows, int pitch) {
  StemPairs t;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int kp = 8 * (i >> 1) + q + 4 * (i & 1);
    const int ck = kp >> 1, c = ck / 3, kh = ck - 3 * c;
    t.off[i] = kp < kStemSlots / 2 ? (c * rows + kh) * pitch + 2 * (kp & 1) : -1;
  }
  return t;
}

// U = X . Wk + bias for one 16-pixel tile: rows g (window origin oa, bf16 offset in the tile) and g + 8 (origin ob)
template <int C0>
__device__ __forceinline__ void conv_tile(const uint16_t* xt, int lo_plane, const StemPairs& pr, int oa, int ob,
                                          const uint2* bf_hi, const uint2* bf_lo, const float* bsm, int lane,
                                          float (&acc)[C0 / 8][4]) {
  typedef StemTc<C0> T;
  const int q = lane & 3;
  uint32_t ah[3][4], al[3][4];
#pragma unroll
  for (int s = 0; s < 3; ++s) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int off = pr.off[2 * s + h];
      if (off >= 0) {
        ah[s][2 * h] = *reinterpret_cast<const uint32_t*>(xt + oa + off);
        ah[s][2 * h + 1] = *reinterpret_cast<const uint32_t*>(xt + ob + off);
        al[s][2 * h] = *reinterpret_cast<const uint32_t*>(xt + lo_plane + oa + off);
        al[s][2 * h + 1] = *reinterpret_cast<const uint32_t*>(xt + lo_plane + ob + off);
      } else {
        ah[s][2 * h] = ah[s][2 * h + 1] = al[s][2 * h] = al[s][2 * h + 1] = 0u;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < T::NT; ++t) {
    const float b0 = bsm[8 * t + 2 * q], b1 = bsm[8 * t + 2 * q + 1];
    acc[t][0] = b0; acc[t][1] = b1; acc[t][2] = b0; acc[t][3] = b1;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const uint2 wh = bf_hi[(s * T::NT + t) * 32 + lane], wl = bf_lo[(s * T::NT + t) * 32 + lane];
      mma16816(acc[t], ah[s], wh);
      mma16816(acc[t], al[s], wh);
      mma16816(acc[t], ah[s], wl);
    }
  }
}

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// LayerNorm statistics of the two pixels of the lane's quad (rows g and g + 8)
template <int C0>
__device__ __forceinline__ void ln_stats_frag(const float (&acc)[C0 / 8][4], float eps, float (&mu)[2], float (&rs)[2]) {
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int t = 0; t < C0 / 8; ++t) { sa += acc[t][0] + acc[t][1]; sb += acc[t][2] + acc[t][3]; }
  mu[0] = quad_sum(sa) * (1.0f / C0); mu[1] = quad_sum(sb) * (1.0f / C0);
  float qa = 0.f, qb = 0.f;
#pragma unroll
  for (int t = 0; t < C0 / 8; ++t) {
    const float a0 = acc[t][0] - mu[0], a1 = acc[t][1] - mu[0], b0 = acc[t][2] - mu[1], b1 = acc[t][3] - mu[1];
    qa += a0 * a0 + a1 * a1; qb += b0 * b0 + b1 * b1;
  }
  rs[0] = rsqrtf(quad_sum(qa) * (1.0f / C0) + eps); rs[1] = rsqrtf(quad_sum(qb) * (1.0f / C0) + eps);
}

// forward: a CTA owns a 16 x 16 tile of output pixels of one image (16 M tiles = the 16 output rows, two per warp)
constexpr int kTcFwdThreads = 256;
constexpr int kTcFwdTile = 16;
constexpr int kTcFwdRows = 2 * kTcFwdTile + 1;      // input rows of the tile
constexpr int kTcFwdPitch = 2 * kTcFwdTile + 2;     // input columns + the dummy kw' = 3 column (even)

template <int C0>
__global__ void __launch_bounds__(kTcFwdThreads) stem0_fwd_tc_kernel(const StemParams p, bf16* __restrict__ y, int tiles_w,
                                                                    int tiles_h) {
  typedef StemTc<C0> T;
  constexpr int kPitch = C0 * 2 + 16;                     // bytes per staged pixel row
  extern __shared__ __align__(16) uint8_t stem_smem[];
  uint2* bf_hi = reinterpret_cast<uint2*>(stem_smem);
  uint2* bf_lo = bf_hi + T::kBf;
  float* bsm = reinterpret_cast<float*>(bf_lo + T::kBf);
  float* lw = bsm + C0;
  float* lb = lw + C0;
  uint16_t* xt = reinterpret_cast<uint16_t*>(lb + C0);
  constexpr int kPlane3 = 3 * kTcFwdRows * kTcFwdPitch;   // bf16 per hi (or lo) half
  uint8_t* stage = reinterpret_cast<uint8_t*>(xt + 2 * kPlane3 + 8);
  stage += (16 - (reinterpret_cast<uintptr_t>(stage) & 15)) & 15;
  int bid = blockIdx.x;
  const int tw = bid % tiles_w; bid /= tiles_w;
  const int th = bid % tiles_h; bid /= tiles_h;
  const int n = bid;
  const int ho0 = th * kTcFwdTile, wo0 = tw * kTcFwdTile;
  build_tables<C0>(p.wk, bf_hi, bf_lo, nullptr, nullptr, kTcFwdThreads);
  for (int i = threadIdx.x; i < C0; i += kTcFwdThreads) { bsm[i] = p.bias[i]; lw[i] = p.ln_w[i]; lb[i] = p.ln_b[i]; }
  fill_tile(p, p.x + (int64_t)n * 3 * p.H * p.W, 2 * ho0 - 1, 2 * wo0 - 1, kTcFwdRows, kTcFwdPitch, xt, kTcFwdThreads);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const StemPairs pr = make_pairs(q, kTcFwdRows, kTcFwdPitch);
#pragma unroll 1
  for (int m = warp; m < kTcFwdTile; m += kTcFwdThreads / 32) {     // M tile m = output row ho0 + m, columns wo0 .. wo0 + 15
    const int oa = (2 * m) * kTcFwdPitch + 2 * g, ob = oa + 16;
    float acc[T::NT][4];
    conv_tile<C0>(xt, kPlane3, pr, oa, ob, bf_hi, bf_lo, bsm, lane, acc);
    float mu[2], rs[2];
    ln_stats_frag<C0>(acc, p.eps, mu, rs);
    uint8_t* rowa = stage + (16 * m + g) * kPitch;
    uint8_t* rowb = rowa + 8 * kPitch;
#pragma unroll
    for (int t = 0; t < T::NT; ++t) {
      const int c = 8 * t + 2 * q;
      const float w0 = lw[c], w1 = lw[c + 1], b0 = lb[c], b1 = lb[c + 1];
      const float2 ga2 = b200at_gelu2(make_float2((acc[t][0] - mu[0]) * rs[0] * w0 + b0, (acc[t][1] - mu[0]) * rs[0] * w1 + b1));
      const float2 gb2 = b200at_gelu2(make_float2((acc[t][2] - mu[1]) * rs[1] * w0 + b0, (acc[t][3] - mu[1]) * rs[1] * w1 + b1));
      const bf162 oa2 = __floats2bfloat162_rn(ga2.x, ga2.y);
      const bf162 ob2 = __floats2bfloat162_rn(gb2.x, gb2.y);
      *reinterpret_cast<bf162*>(rowa + c * 2) = oa2;
      *reinterpret_cast<bf162*>(rowb + c * 2) = ob2;
    }
  }
  __syncthreads();
  // an output row of the tile is a contiguous run of (live columns) * C0 * 2 bytes in y
  constexpr int kPieces = C0 * 2 / 16;
  const int live_w = min(kTcFwdTile, p.Wo - wo0), live_h = min(kTcFwdTile, p.Ho - ho0);
  for (int i = threadIdx.x; i < live_h * kTcFwdTile * kPieces; i += kTcFwdThreads) {
    const int piece = i % kPieces, px = (i / kPieces) % kTcFwdTile, r = i / (kPieces * kTcFwdTile);
    if (px < live_w)
      reinterpret_cast<uint4*>(y + (((int64_t)n * p.Ho + ho0 + r) * p.Wo + wo0 + px) * C0)[piece] =
          *reinterpret_cast<const uint4*>(stage + (16 * r + px) * kPitch + piece * 16);
  }
}

constexpr int kTcBwdThreads = 320;     // 10 warps; the 17 x 17 pixels of a tile are 19 tiles of 16 (two per warp)
constexpr int kTcBwdRows = 2 * kBwdSide + 1;        // 35 input rows
constexpr int kTcBwdPitch = 2 * kBwdSide + 2;       // 36
constexpr int kTcVPitch = 37;                       // vbuf words per pixel (36 slots; odd: fewer bank conflicts)

template <int C0>
__global__ void __launch_bounds__(kTcBwdThreads, (C0 <= 48 ? 2 : 1)) stem0_bwd_tc_kernel(const StemParams p, const bf16* __restrict__ dy,
                                                                    float* __restrict__ dx, int tiles_w, int tiles_h) {
  typedef StemTc<C0> T;
  extern __shared__ __align__(16) uint8_t stem_smem[];      // dynamic (opt-in above 48 KB)
  uint2* bf_hi = reinterpret_cast<uint2*>(stem_smem);
  uint2* bf_lo = bf_hi + T::kBf;
  uint2* bt_hi = bf_lo + T::kBf;
  uint2* bt_lo = bt_hi + T::kBt;
  float* bsm = reinterpret_cast<float*>(bt_lo + T::kBt);
  float* lw = bsm + C0;
  float* lb = lw + C0;
  float* vbuf = lb + C0;                                   // [kBwdPix][kTcVPitch]
  uint16_t* xt = reinterpret_cast<uint16_t*>(vbuf + kBwdPix * kTcVPitch);
  constexpr int kPlane3 = 3 * kTcBwdRows * kTcBwdPitch;
  int bid = blockIdx.x;
  const int tw = bid % tiles_w; bid /= tiles_w;
  const int th = bid % tiles_h; bid /= tiles_h;
  const int n = bid;
  const int ho0 = th * kBwdTileOut, wo0 = tw * kBwdTileOut;
  build_tables<C0>(p.wk, bf_hi, bf_lo, bt_hi, bt_lo, kTcBwdThreads);
  for (int i = threadIdx.x; i < C0; i += kTcBwdThreads) { bsm[i] = p.bias[i]; lw[i] = p.ln_w[i]; lb[i] = p.ln_b[i]; }
  fill_tile(p, p.x + (int64_t)n * 3 * p.H * p.W, 2 * ho0 - 1, 2 * wo0 - 1, kTcBwdRows, kTcBwdPitch, xt, kTcBwdThreads);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const StemPairs pr = make_pairs(q, kTcBwdRows, kTcBwdPitch);
#pragma unroll 1
  for (int m = warp; m < (kBwdPix + 15) / 16; m += kTcBwdThreads / 32) {
    int idx[2], ho[2], wo[2], org[2]; bool live[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      idx[r] = 16 * m + g + 8 * r;
      const int ic = idx[r] < kBwdPix ? idx[r] : 0;
      const int lr = ic / kBwdSide, lc = ic - lr * kBwdSide;
      ho[r] = ho0 + lr; wo[r] = wo0 + lc;
      org[r] = (2 * lr) * kTcBwdPitch + 2 * lc;
      live[r] = idx[r] < kBwdPix && ho[r] < p.Ho && wo[r] < p.Wo;
    }
    float acc[T::NT][4];
    conv_tile<C0>(xt, kPlane3, pr, org[0], org[1], bf_hi, bf_lo, bsm, lane, acc);
    float mu[2], rs[2];
    ln_stats_frag<C0>(acc, p.eps, mu, rs);
    // g = dy * GELU'(pre) * ln_w ; dconv = rs * (g - mean(g) - xhat * mean(g * xhat))   (acc <- xhat; gg <- g)
    const bf16* dya = dy + (((int64_t)n * p.Ho + (live[0] ? ho[0] : 0)) * p.Wo + (live[0] ? wo[0] : 0)) * C0 + 2 * q;
    const bf16* dyb = dy + (((int64_t)n * p.Ho + (live[1] ? ho[1] : 0)) * p.Wo + (live[1] ? wo[1] : 0)) * C0 + 2 * q;
    float gg[T::NT][4];
    float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
    for (int t = 0; t < T::NT; ++t) {
      const int c = 8 * t + 2 * q;
      const float w0 = lw[c], w1 = lw[c + 1], b0 = lb[c], b1 = lb[c + 1];
      const uint32_t ua = live[0] ? __ldg(reinterpret_cast<const uint32_t*>(dya + 8 * t)) : 0u;
      const uint32_t ub = live[1] ? __ldg(reinterpret_cast<const uint32_t*>(dyb + 8 * t)) : 0u;
      const float2 da = __bfloat1622float2(*reinterpret_cast<const bf162*>(&ua));
      const float2 db = __bfloat1622float2(*reinterpret_cast<const bf162*>(&ub));
      const float x0 = (acc[t][0] - mu[0]) * rs[0], x1 = (acc[t][1] - mu[0]) * rs[0];
      const float x2 = (acc[t][2] - mu[1]) * rs[1], x3 = (acc[t][3] - mu[1]) * rs[1];
      const float2 gpa = b200at_gelu_grad2(make_float2(x0 * w0 + b0, x1 * w1 + b1));
      const float2 gpb = b200at_gelu_grad2(make_float2(x2 * w0 + b0, x3 * w1 + b1));
      gg[t][0] = da.x * gpa.x * w0;
      gg[t][1] = da.y * gpa.y * w1;
      gg[t][2] = db.x * gpb.x * w0;
      gg[t][3] = db.y * gpb.y * w1;
      acc[t][0] = x0; acc[t][1] = x1; acc[t][2] = x2; acc[t][3] = x3;
      s1a += gg[t][0] + gg[t][1]; s2a += gg[t][0] * x0 + gg[t][1] * x1;
      s1b += gg[t][2] + gg[t][3]; s2b += gg[t][2] * x2 + gg[t][3] * x3;
    }
    const float m1a = quad_sum(s1a) * (1.0f / C0), m2a = quad_sum(s2a) * (1.0f / C0);
    const float m1b = quad_sum(s1b) * (1.0f / C0), m2b = quad_sum(s2b) * (1.0f / C0);
    // dU as the A operand of the second product: channel tiles (2s, 2s + 1) of the fragment are k-step s
    float v[5][4];
#pragma unroll
    for (int t = 0; t < 5; ++t) { v[t][0] = 0.f; v[t][1] = 0.f; v[t][2] = 0.f; v[t][3] = 0.f; }
#pragma unroll
    for (int s = 0; s < T::KS; ++s) {
      uint32_t ah[4], al[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int t = 2 * s + h;
        const float d0 = rs[0] * (gg[t][0] - m1a - acc[t][0] * m2a), d1 = rs[0] * (gg[t][1] - m1a - acc[t][1] * m2a);
        const float d2 = rs[1] * (gg[t][2] - m1b - acc[t][2] * m2b), d3 = rs[1] * (gg[t][3] - m1b - acc[t][3] * m2b);
        split2(d0, d1, ah[2 * h], al[2 * h]);            // row g
        split2(d2, d3, ah[2 * h + 1], al[2 * h + 1]);    // row g + 8
      }
#pragma unroll
      for (int t = 0; t < 5; ++t) {
        const uint2 wh = bt_hi[(s * 5 + t) * 32 + lane], wl = bt_lo[(s * 5 + t) * 32 + lane];
        mma16816(v[t], ah, wh);
        mma16816(v[t], al, wh);
        mma16816(v[t], ah, wl);
      }
    }
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      const int k = 8 * t + 2 * q;                          // slots k, k + 1 (< 36 for t < 4; t == 4: q < 2)
      if (k < kStemSlots) {
        if (idx[0] < kBwdPix) { vbuf[idx[0] * kTcVPitch + k] = v[t][0]; vbuf[idx[0] * kTcVPitch + k + 1] = v[t][1]; }
        if (idx[1] < kBwdPix) { vbuf[idx[1] * kTcVPitch + k] = v[t][2]; vbuf[idx[1] * kTcVPitch + k + 1] = v[t][3]; }
      }
    }
  }
  __syncthreads();
  // gather: input pixel (h, w) of the 32x32 tile receives window entry (kh, kw) of output pixel
  // ((h + 1 - kh) / 2, (w + 1 - kw) / 2) whenever those are integers
  const int h0 = 2 * ho0, w0 = 2 * wo0;
  float* dxn = dx + (int64_t)n * 3 * p.H * p.W;
  for (int i = threadIdx.x; i < 3 * 32 * 32; i += kTcBwdThreads) {
    const int lw_ = i & 31, lh = (i >> 5) & 31, c = i >> 10;
    const int h = h0 + lh, w = w0 + lw_;
    if (h >= p.H || w >= p.W) continue;
    float sum = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int th2 = lh + 1 - kh;
      if (th2 < 0 || (th2 & 1)) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tw2 = lw_ + 1 - kw;
        if (tw2 < 0 || (tw2 & 1)) continue;
        sum += vbuf[((th2 >> 1) * kBwdSide + (tw2 >> 1)) * kTcVPitch + (c * 3 + kh) * 4 + kw];
      }
    }
    const float is = c == 0 ? p.inv_std[0] : (c == 1 ? p.inv_std[1] : p.inv_std[2]);
    dxn[((int64_t)c * p.H + h) * p.W + w] = sum * is;
  }
}

template <int C0>
constexpr int stem_fwd_tc_smem() {
  return 2 * StemTc<C0>::kBf * 8 + 3 * C0 * 4 + 2 * 3 * kTcFwdRows * kTcFwdPitch * 2 + 16 + 16 + 256 * (C0 * 2 + 16);
}
template <int C0>
constexpr int stem_bwd_tc_smem() {
  return (2 * StemTc<C0>::kBf + 2 * StemTc<C0>::kBt) * 8 + 3 * C0 * 4 + kBwdPix * kTcVPitch * 4 +
         2 * 3 * kTcBwdRows * kTcBwdPitch * 2;
}

// B200AT_STEM0_TC: unset = tensor-core input gradient (472 vs 745 us at batch 128) and FMA forward (210 vs 225 us: with the
// contraction gone both forwards are bound by the same LayerNorm + GELU tail, and the MMA form pays more integer
// overhead per tile -- profiles/r02_stem0_tc_ncu.txt); "0" = FMA kernels only; "all" = tensor-core kernels for both.
inline int stem_tc_mode() {
  static const int mode = [] {
    const char* e = getenv("B200AT_STEM0_TC");
    if (e == nullptr) return 1;
    if (e[0] == '0') return 0;
    return (e[0] == 'a') ? 2 : 1;
  }();
  return mode;
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
  if (stem_tc_mode() == 2) {
    const int tiles_h = (p.Ho + kTcFwdTile - 1) / kTcFwdTile, tiles_w = (p.Wo + kTcFwdTile - 1) / kTcFwdTile;
    const int64_t gtc = (int64_t)p.B * tiles_h * tiles_w;
    if (gtc > 0x7fffffff) return (int)cudaErrorInvalidValue;
#define B200AT_STEM_FWD(CC)                                                                                 \
    {                                                                                                       \
      static b200at::SmemConfig conf;                                                                       \
      cudaError_t e = b200at::ensure_dynamic_smem(stem0_fwd_tc_kernel<CC>, stem_fwd_tc_smem<CC>(), conf);   \
      if (e != cudaSuccess) return (int)e;                                                                  \
      stem0_fwd_tc_kernel<CC><<<(unsigned)gtc, kTcFwdThreads, stem_fwd_tc_smem<CC>(), s>>>(p, (bf16*)y, tiles_w, tiles_h); \
    }
    switch (C0) {
      case 48: B200AT_STEM_FWD(48) break;
      case 64: B200AT_STEM_FWD(64) break;
      case 96: B200AT_STEM_FWD(96) break;
      default: return (int)cudaErrorInvalidValue;
    }
#undef B200AT_STEM_FWD
    return (int)cudaGetLastError();
  }
  switch (C0) {
    case 48: stem0_fwd_kernel<48><<<(unsigned)grid, kFwdThreads, 0, s>>>(p, (bf16*)y); break;
    case 64: stem0_fwd_kernel<64><<<(unsigned)grid, kFwdThreads, 0, s>>>(p, (bf16*)y); break;
    case 96: stem0_fwd_kernel<96><<<(unsigned)grid, kFwdThreads, 0, s>>>(p, (bf16*)y); break;
    default: return (int)cudaErrorInvalidValue;
  }
  return (int)cudaGetLastError();
}

int b200at_stem0_fwd_save(const float* x, const float* mean3, const float* std3, const float* wk, const float* bias,
                          const float* ln_w, const float* ln_b, void* y, void* y_pre, float* mean, float* rstd, int64_t B,
                          int64_t H, int64_t W, int64_t C0, float eps, void* stream) {
  if (B <= 0) return 0;
  StemParams p;
  if (H < 1 || W < 1 || !fill(p, x, mean3, std3, wk, bias, ln_w, ln_b, B, H, W, eps)) return (int)cudaErrorInvalidValue;
  if (y_pre == nullptr || mean == nullptr || rstd == nullptr) return (int)cudaErrorInvalidValue;
  const int64_t total = (int64_t)p.B * p.Ho * p.Wo;
  const int64_t grid = (total + kFwdThreads - 1) / kFwdThreads;
  if (grid > 0x7fffffff) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  switch (C0) {
    case 48: stem0_fwd_kernel<48, true><<<(unsigned)grid, kFwdThreads, 0, s>>>(p, (bf16*)y, (bf16*)y_pre, mean, rstd); break;
    case 64: stem0_fwd_kernel<64, true><<<(unsigned)grid, kFwdThreads, 0, s>>>(p, (bf16*)y, (bf16*)y_pre, mean, rstd); break;
    case 96: stem0_fwd_kernel<96, true><<<(unsigned)grid, kFwdThreads, 0, s>>>(p, (bf16*)y, (bf16*)y_pre, mean, rstd); break;
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
  if (stem_tc_mode() >= 1) {
#define B200AT_STEM_BWD(CC)                                                                                           \
    {                                                                                                                 \
      constexpr int smem = stem_bwd_tc_smem<CC>();                                                                    \
      static b200at::SmemConfig conf;                                                                                 \
      cudaError_t e = b200at::ensure_dynamic_smem(stem0_bwd_tc_kernel<CC>, smem, conf);                               \
      if (e != cudaSuccess) return (int)e;                                                                            \
      stem0_bwd_tc_kernel<CC><<<(unsigned)grid, kTcBwdThreads, smem, s>>>(p, (const bf16*)dy, dx, tiles_w, tiles_h); \
    }
    switch (C0) {
      case 48: B200AT_STEM_BWD(48) break;
      case 64: B200AT_STEM_BWD(64) break;
      case 96: B200AT_STEM_BWD(96) break;
      default: return (int)cudaErrorInvalidValue;
    }
#undef B200AT_STEM_BWD
    return (int)cudaGetLastError();
  }
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
