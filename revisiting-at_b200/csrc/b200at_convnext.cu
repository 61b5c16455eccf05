// sm_100a kernels for the memory-bound ConvNeXt-CvSt layer ops (include/b200at_model.h):
//   K7  depthwise 7x7 conv, NHWC bf16: forward, input-gradient (same kernel, flipped taps), weight-gradient
//   K8  per-pixel LayerNorm over C (channels-last; also the stems' "channels_first" LN, which on NHWC
//       memory is the same op): forward (+optional fused GELU), input-gradient, gamma/beta gradients
//   K9  bias+GELU(erf) forward/backward on the 4C hidden, layer-scale + residual
// Reference math: /root/reference/models/convnext.py:37-50, utils_architecture.py:57-81.
// Activations are NHWC bf16 with fp32 statistics / accumulation.  Every pass is 8- or 16-byte vectorised
// and coalesced along C; the depthwise conv stages its halo tile through shared memory and is register
// tiled (a thread slides one output column through 7 input columns) because at bf16 it is FMA-issue
// bound rather than HBM bound on B200.  Entry points never allocate or synchronise.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "b200at_gelu.cuh"
#include "b200at_launch.cuh"
#include "b200at_tma.cuh"
#include "../../include/b200at_model.h"

namespace {

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

__device__ __forceinline__ float gelu_f(float v) { return b200at_gelu(v); }
__device__ __forceinline__ float gelu_grad_f(float v) { return b200at_gelu_grad(v); }
// packed fp32x2 forms (sm_100 FFMA2): one issue slot for the two channels a thread carries
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return b200at_ffma2(a, b, c); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return b200at_fadd2(a, b); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return b200at_fmul2(a, b); }
// bf16x2 word -> two fp32 (exact): low half << 16, high half masked -- two ALU-pipe instructions, none on the FMA pipe
__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2_to_bf2(float2 v) {
  const bf162 t = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ void unpack4(const uint2& u, float* f) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const bf162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const bf162*>(&u.y));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ uint2 pack4(const float* f) {
  uint2 u;
  bf162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  return u;
}

// ------------------------------------------------------------------------------------------------ K8
// A group of G lanes (4/8/16/32) owns one pixel row of C channels; each lane holds VPL 8-byte vectors
// (4 bf16) in registers, so a warp works on 32/G rows at once (small C: more rows in flight per warp).
constexpr int kLnThreads = 256;
constexpr int kLnMaxVec = 12 * 32;        // float4 groups of the widest row the plain LayerNorm kernels take (C <= 1536)

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Row of pixel (b, h, w) in the "2x2 patch" layout [B][H/2][W/2][2][2][C]: the 4 pixels of a stride-2 2x2 window are
// adjacent, so the downsample convolution (models/convnext.py:79-82) is a plain GEMM over rows of 4C channels.
// pW == 0: identity.
__device__ __forceinline__ int64_t patch2_row(int64_t row, int pH, int pW) {
  if (pW == 0) return row;
  const int w = (int)(row % pW);
  const int64_t t = row / pW;
  const int h = (int)(t % pH);
  const int64_t b = t / pH;
  return (((b * (pH >> 1) + (h >> 1)) * (pW >> 1) + (w >> 1)) << 2) + ((h & 1) << 1) + (w & 1);
}

template <int G, int VPL, bool GELU>
__global__ void __launch_bounds__(kLnThreads) ln_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ b, bf16* __restrict__ y,
                                                            float* __restrict__ mean, float* __restrict__ rstd,
                                                            int64_t M, int C, float eps, int pH, int pW,
                                                            const float* __restrict__ pre_bias) {
  const int gl = threadIdx.x % G;                       // lane within the row group
  const int nv = C >> 2;
  constexpr int kRows = kLnThreads / G;                 // rows per CTA iteration
  // weight, bias and the optional pre-bias live in shared memory, not in 12 VPL registers per thread: at C = 384 / 768 the
  // register form ran 24 / 16 warps per SM (80 / 128 registers), too few loads in flight for the small maps of stages 2-3
  __shared__ float4 sw4[kLnMaxVec], sb4[kLnMaxVec], spb4[kLnMaxVec];
  const bool has_pb = pre_bias != nullptr;
  for (int c = threadIdx.x; c < C; c += kLnThreads) {
    reinterpret_cast<float*>(sw4)[c] = w[c];
    reinterpret_cast<float*>(sb4)[c] = b[c];
    if (has_pb) reinterpret_cast<float*>(spb4)[c] = pre_bias[c];
  }
  __syncthreads();
  const float inv_c = 1.0f / (float)C;
  const int64_t rows_pad = (M + kRows - 1) / kRows * kRows;   // keep whole warps in the loop (shuffles)
  for (int64_t row = (int64_t)blockIdx.x * kRows + threadIdx.x / G; row < rows_pad; row += (int64_t)gridDim.x * kRows) {
    const bool live = row < M;
    const uint2* xr = reinterpret_cast<const uint2*>(x + row * C);
    float f[VPL][4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int v = gl + G * j;
      if (live && v < nv) {
        unpack4(__ldcs(xr + v), f[j]);
        if (has_pb) {
          const float4 pb = spb4[v];
          f[j][0] += pb.x; f[j][1] += pb.y; f[j][2] += pb.z; f[j][3] += pb.w;
        }
        s += (f[j][0] + f[j][1]) + (f[j][2] + f[j][3]);
      } else {
        f[j][0] = f[j][1] = f[j][2] = f[j][3] = 0.f;
      }
    }
    const float mu = group_sum<G>(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int v = gl + G * j;
      if (v < nv) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float d = f[j][k] - mu; q += d * d; }
      }
    }
    const float rs = rsqrtf(group_sum<G>(q) * inv_c + eps);
    if (!live) continue;
    uint2* yr = reinterpret_cast<uint2*>(y + patch2_row(row, pH, pW) * C);
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int v = gl + G * j;
      if (v < nv) {
        float o[4];
        const float4 w4 = sw4[v], b4 = sb4[v];
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          o[k] = (f[j][k] - mu) * rs * wv[k] + bv[k];
          if (GELU) o[k] = gelu_f(o[k]);
        }
        yr[v] = pack4(o);
      }
    }
    if (gl == 0) { mean[row] = mu; rstd[row] = rs; }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w  (dy first multiplied by gelu'(pre) if GELU)
template <int G, int VPL, bool GELU, bool PGRAD>
__global__ void __launch_bounds__(kLnThreads) ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                            const float* __restrict__ w, const float* __restrict__ b,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, bf16* __restrict__ dx,
                                                            float* __restrict__ dw, float* __restrict__ db,
                                                            int64_t M, int C, int pH, int pW,
                                                            const float* __restrict__ pre_bias) {
  extern __shared__ float red[];  // PGRAD: [2][C] accumulated with shared-memory atomics
  const int gl = threadIdx.x % G;
  const int nv = C >> 2;
  constexpr int kRows = kLnThreads / G;
  float aw[VPL][4], ab[VPL][4];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { aw[j][k] = 0.f; ab[j][k] = 0.f; }
  }
  __shared__ float4 sw4[kLnMaxVec], sb4[kLnMaxVec], spb4[kLnMaxVec];   // see ln_fwd_kernel
  const bool has_pb = pre_bias != nullptr;
  for (int c = threadIdx.x; c < C; c += kLnThreads) {
    reinterpret_cast<float*>(sw4)[c] = w[c];
    if (GELU) reinterpret_cast<float*>(sb4)[c] = b[c];
    if (has_pb) reinterpret_cast<float*>(spb4)[c] = pre_bias[c];
  }
  if (!PGRAD) __syncthreads();
  if (PGRAD) {
    for (int c = threadIdx.x; c < 2 * C; c += kLnThreads) red[c] = 0.f;
    __syncthreads();
  }
  const float inv_c = 1.0f / (float)C;
  const int64_t rows_pad = (M + kRows - 1) / kRows * kRows;
  for (int64_t row = (int64_t)blockIdx.x * kRows + threadIdx.x / G; row < rows_pad; row += (int64_t)gridDim.x * kRows) {
    const bool live = row < M;
    const uint2* xr = reinterpret_cast<const uint2*>(x + row * C);
    const uint2* gr = reinterpret_cast<const uint2*>(dy + patch2_row(live ? row : 0, pH, pW) * C);
    const float mu = live ? mean[row] : 0.f, rs = live ? rstd[row] : 0.f;
    float xh[VPL][4], g[VPL][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int v = gl + G * j;
      if (live && v < nv) {
        float xv[4], dv[4];
        unpack4(__ldcs(xr + v), xv);
        unpack4(__ldcs(gr + v), dv);
        const float4 w4 = sw4[v];
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
        float bv[4] = {0.f, 0.f, 0.f, 0.f}, pv[4] = {0.f, 0.f, 0.f, 0.f};
        if (GELU) { const float4 b4 = sb4[v]; bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w; }
        if (has_pb) { const float4 p4 = spb4[v]; pv[0] = p4.x; pv[1] = p4.y; pv[2] = p4.z; pv[3] = p4.w; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          xh[j][k] = ((xv[k] + pv[k]) - mu) * rs;
          float d = dv[k];
          if (GELU) d *= gelu_grad_f(xh[j][k] * wv[k] + bv[k]);
          if (PGRAD) { aw[j][k] += d * xh[j][k]; ab[j][k] += d; }
          g[j][k] = d * wv[k];
          s1 += g[j][k];
          s2 += g[j][k] * xh[j][k];
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { xh[j][k] = 0.f; g[j][k] = 0.f; }
      }
    }
    const float m1 = group_sum<G>(s1) * inv_c, m2 = group_sum<G>(s2) * inv_c;
    if (!live) continue;
    uint2* dr = reinterpret_cast<uint2*>(dx + row * C);
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int v = gl + G * j;
      if (v < nv) {
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = rs * (g[j][k] - m1 - xh[j][k] * m2);
        dr[v] = pack4(o);
      }
    }
  }
  if (PGRAD) {
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int v = gl + G * j;
      if (v < nv) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          atomicAdd(&red[v * 4 + k], aw[j][k]);
          atomicAdd(&red[C + v * 4 + k], ab[j][k]);
        }
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kLnThreads) {
      atomicAdd(dw + c, red[c]);
      atomicAdd(db + c, red[C + c]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ K8, pipelined form
// The kernels above keep one row per lane group in flight (3 x 8-byte loads per lane): 3.7-4.0 TB/s of the 6.5 TB/s copy
// rate (profiles/r01_ncu_dwconv_stem0_v11_summary.txt) -- not enough bytes in flight per SM.  Rows of [M][C] are
// contiguous, so a tile of rows is ONE cp.async.bulk (global -> shared, mbarrier completion): persistent CTAs run a
// kRingStages-deep ring of input tiles (>= 100 KB in flight per SM), the same lane-group arithmetic reads its rows from
// shared memory, results leave through a double-buffered output tile and one cp.async.bulk store per tile.  The
// per-row statistics ride in the ring as two more (small) bulk copies in the backward.
constexpr int kRingThreads = 512;
constexpr int kRingStages = 3;

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   b200at::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(b200at::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(b200at::smem_u32(src)),
               "r"(bytes)
               : "memory");
}

// C == 4 * G * VPL exactly (no lane predicates); arithmetic on packed fp32 pairs (add / mul / fma .f32x2): the scalar form
// of this loop issued 25 instructions per element and was ISSUE-bound at 0.52 of the HBM rate
// (profiles/r02_ln_ring_ncu_v1.txt), this one ~6.
template <int G, int VPL, bool GELU>
__global__ void __launch_bounds__(kRingThreads) ln_fwd_ring_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                                  const float* __restrict__ b, bf16* __restrict__ y,
                                                                  float* __restrict__ mean, float* __restrict__ rstd,
                                                                  int64_t M, float eps, int tile_rows,
                                                                  const float* __restrict__ pre_bias) {
  constexpr int C = 4 * G * VPL;
  constexpr int kRowBytes = 2 * C;
  extern __shared__ __align__(128) uint8_t ring_smem[];
  const int tile_bytes = tile_rows * kRowBytes;            // multiple of 16
  const int tile_pad = (tile_bytes + 127) & ~127;
  uint8_t* in0 = ring_smem;
  uint8_t* out0 = ring_smem + kRingStages * tile_pad;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring_smem + (kRingStages + 2) * tile_pad);
  const int tid = threadIdx.x, gl = tid % G;
  constexpr int kRows = kRingThreads / G;
  float2 wr[VPL][2], br[VPL][2], pbr[VPL][2];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int c = (gl + G * j) * 4;
    wr[j][0] = make_float2(w[c], w[c + 1]); wr[j][1] = make_float2(w[c + 2], w[c + 3]);
    br[j][0] = make_float2(b[c], b[c + 1]); br[j][1] = make_float2(b[c + 2], b[c + 3]);
    pbr[j][0] = pre_bias ? make_float2(pre_bias[c], pre_bias[c + 1]) : make_float2(0.f, 0.f);
    pbr[j][1] = pre_bias ? make_float2(pre_bias[c + 2], pre_bias[c + 3]) : make_float2(0.f, 0.f);
  }
  constexpr float inv_c = 1.0f / (float)C;
  const int64_t tiles = (M + tile_rows - 1) / tile_rows;
  auto rows_of = [&](int64_t t) { const int64_t left = M - t * tile_rows; return (int)(left < tile_rows ? left : tile_rows); };
  if (tid == 0) {
    for (int s = 0; s < kRingStages; ++s) b200at::mbar_init(&full[s], 1);
    b200at::mbar_fence_init();
    for (int s = 0; s < kRingStages; ++s) {
      const int64_t t = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (t < tiles) {
        const uint32_t bytes = (uint32_t)rows_of(t) * kRowBytes;
        b200at::mbar_expect_tx(&full[s], bytes);
        bulk_load(in0 + s * tile_pad, x + t * tile_rows * C, bytes, &full[s]);
      }
    }
  }
  __syncthreads();
  int k = 0;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++k) {
    const int s = k % kRingStages, ob = k & 1;
    const int rows = rows_of(t);
    b200at::mbar_wait(&full[s], (uint32_t)((k / kRingStages) & 1));
    const uint8_t* in = in0 + s * tile_pad;
    uint8_t* out = out0 + ob * tile_pad;
    float* mean_t = mean + t * tile_rows;
    float* rstd_t = rstd + t * tile_rows;
    const int rows_pad = (rows + kRows - 1) / kRows * kRows;     // whole warps stay in the loop (shuffles)
    for (int r = tid / G; r < rows_pad; r += kRows) {
      const bool live = r < rows;
      const uint2* xr = reinterpret_cast<const uint2*>(in + (live ? r : 0) * kRowBytes) + gl;
      float2 f[VPL][2];
      float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const uint2 u = xr[G * j];
        f[j][0] = fadd2(bf2_to_f2(u.x), pbr[j][0]); f[j][1] = fadd2(bf2_to_f2(u.y), pbr[j][1]);
        sum2 = fadd2(sum2, fadd2(f[j][0], f[j][1]));
      }
      const float mu = group_sum<G>(sum2.x + sum2.y) * inv_c;
      const float2 nmu = make_float2(-mu, -mu);
      float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        f[j][0] = fadd2(f[j][0], nmu); f[j][1] = fadd2(f[j][1], nmu);
        q2 = ffma2(f[j][0], f[j][0], q2);
        q2 = ffma2(f[j][1], f[j][1], q2);
      }
      const float rs = rsqrtf(group_sum<G>(q2.x + q2.y) * inv_c + eps);
      if (!live) continue;
      const float2 rs2 = make_float2(rs, rs);
      uint2* yr = reinterpret_cast<uint2*>(out + r * kRowBytes) + gl;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        float2 o0 = ffma2(f[j][0], fmul2(rs2, wr[j][0]), br[j][0]);
        float2 o1 = ffma2(f[j][1], fmul2(rs2, wr[j][1]), br[j][1]);
        if (GELU) {
          o0 = b200at_gelu2(o0); o1 = b200at_gelu2(o1);
        }
        yr[G * j] = make_uint2(f2_to_bf2(o0), f2_to_bf2(o1));
      }
      if (gl == 0) { mean_t[r] = mu; rstd_t[r] = rs; }
    }
    b200at::fence_proxy_async();
    if (tid == 0) b200at::tma_store_wait_read();      // the store of tile k-1 has drained the OTHER output buffer
    __syncthreads();
    if (tid == 0) {
      bulk_store(y + t * tile_rows * C, out, (uint32_t)rows * kRowBytes);
      b200at::tma_store_commit();
      const int64_t tn = t + (int64_t)kRingStages * gridDim.x;
      if (tn < tiles) {
        const uint32_t bytes = (uint32_t)rows_of(tn) * kRowBytes;
        b200at::mbar_expect_tx(&full[s], bytes);
        bulk_load(in0 + s * tile_pad, x + tn * tile_rows * C, bytes, &full[s]);
      }
    }
  }
  if (tid == 0) b200at::tma_store_wait_all();
}

template <int G, int VPL, bool GELU, bool PGRAD>
__global__ void __launch_bounds__(kRingThreads) ln_bwd_ring_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                                  const float* __restrict__ w, const float* __restrict__ b,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd, bf16* __restrict__ dx,
                                                                  float* __restrict__ dw, float* __restrict__ db,
                                                                  int64_t M, int tile_rows,
                                                                  const float* __restrict__ pre_bias) {
  constexpr int C = 4 * G * VPL;
  constexpr int kRowBytes = 2 * C;
  extern __shared__ __align__(128) uint8_t ring_smem[];
  const int tile_bytes = tile_rows * kRowBytes;
  const int stat_bytes = tile_rows * 4;                    // tile_rows % 4 == 0
  const int stage_pad = ((2 * tile_bytes + 2 * stat_bytes) + 127) & ~127;   // [dy tile][x tile][mean][rstd]
  const int out_pad = (tile_bytes + 127) & ~127;
  uint8_t* in0 = ring_smem;
  uint8_t* out0 = ring_smem + kRingStages * stage_pad;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring_smem + kRingStages * stage_pad + 2 * out_pad);
  float* red = reinterpret_cast<float*>(full + kRingStages);   // PGRAD: [2][C]
  const int tid = threadIdx.x, gl = tid % G;
  constexpr int kRows = kRingThreads / G;
  float2 wr[VPL][2], br[VPL][2], aw[VPL][2], ab[VPL][2], pbr[VPL][2];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int c = (gl + G * j) * 4;
    pbr[j][0] = pre_bias ? make_float2(pre_bias[c], pre_bias[c + 1]) : make_float2(0.f, 0.f);
    pbr[j][1] = pre_bias ? make_float2(pre_bias[c + 2], pre_bias[c + 3]) : make_float2(0.f, 0.f);
    wr[j][0] = make_float2(w[c], w[c + 1]); wr[j][1] = make_float2(w[c + 2], w[c + 3]);
    br[j][0] = GELU ? make_float2(b[c], b[c + 1]) : make_float2(0.f, 0.f);
    br[j][1] = GELU ? make_float2(b[c + 2], b[c + 3]) : make_float2(0.f, 0.f);
    aw[j][0] = aw[j][1] = ab[j][0] = ab[j][1] = make_float2(0.f, 0.f);
  }
  if (PGRAD)
    for (int c = tid; c < 2 * C; c += kRingThreads) red[c] = 0.f;
  constexpr float inv_c = 1.0f / (float)C;
  const int64_t tiles = (M + tile_rows - 1) / tile_rows;
  auto rows_of = [&](int64_t t) { const int64_t left = M - t * tile_rows; return (int)(left < tile_rows ? left : tile_rows); };
  auto issue = [&](int s, int64_t t) {                      // thread 0 only
    const int rows = rows_of(t);
    const uint32_t bytes = (uint32_t)rows * kRowBytes, sb = (uint32_t)(rows & ~3) * 4;   // statistics: whole 16-byte pieces
    uint8_t* st = in0 + s * stage_pad;
    float* mu_s = reinterpret_cast<float*>(st + 2 * tile_bytes);
    float* rs_s = reinterpret_cast<float*>(st + 2 * tile_bytes + stat_bytes);
    for (int r = rows & ~3; r < rows; ++r) {                // the last tile's 1-3 odd rows: plain copies (every reader
      mu_s[r] = mean[t * tile_rows + r];                    // passes a __syncthreads before it uses this stage)
      rs_s[r] = rstd[t * tile_rows + r];
    }
    b200at::mbar_expect_tx(&full[s], 2 * bytes + 2 * sb);
    bulk_load(st, dy + t * tile_rows * C, bytes, &full[s]);
    bulk_load(st + tile_bytes, x + t * tile_rows * C, bytes, &full[s]);
    if (sb) {
      bulk_load(mu_s, mean + t * tile_rows, sb, &full[s]);
      bulk_load(rs_s, rstd + t * tile_rows, sb, &full[s]);
    }
  };
  if (tid == 0) {
    for (int s = 0; s < kRingStages; ++s) b200at::mbar_init(&full[s], 1);
    b200at::mbar_fence_init();
    for (int s = 0; s < kRingStages; ++s) {
      const int64_t t = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (t < tiles) issue(s, t);
    }
  }
  __syncthreads();
  int k = 0;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++k) {
    const int s = k % kRingStages, ob = k & 1;
    const int rows = rows_of(t);
    b200at::mbar_wait(&full[s], (uint32_t)((k / kRingStages) & 1));
    const uint8_t* st = in0 + s * stage_pad;
    const float* mu_t = reinterpret_cast<const float*>(st + 2 * tile_bytes);
    const float* rs_t = reinterpret_cast<const float*>(st + 2 * tile_bytes + stat_bytes);
    uint8_t* out = out0 + ob * out_pad;
    const int rows_pad = (rows + kRows - 1) / kRows * kRows;
    for (int r = tid / G; r < rows_pad; r += kRows) {
      const bool live = r < rows;
      const int rr = live ? r : 0;
      const uint2* gr = reinterpret_cast<const uint2*>(st + rr * kRowBytes) + gl;
      const uint2* xr = reinterpret_cast<const uint2*>(st + tile_bytes + rr * kRowBytes) + gl;
      const float mu = mu_t[rr], rs = live ? rs_t[rr] : 0.f;
      const float2 rs2 = make_float2(rs, rs), nmr = make_float2(-mu * rs, -mu * rs);
      float2 xh[VPL][2], g[VPL][2];
      float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const uint2 ux = xr[G * j], ud = gr[G * j];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          xh[j][h] = ffma2(fadd2(bf2_to_f2(h ? ux.y : ux.x), pbr[j][h]), rs2, nmr);   // (x + pre_bias - mean) * rstd
          float2 d = bf2_to_f2(h ? ud.y : ud.x);
          if (!live) d = make_float2(0.f, 0.f);
          if (GELU) {
            const float2 pre = ffma2(xh[j][h], wr[j][h], br[j][h]);
            d = fmul2(d, b200at_gelu_grad2(pre));
          }
          if (PGRAD) { aw[j][h] = ffma2(d, xh[j][h], aw[j][h]); ab[j][h] = fadd2(ab[j][h], d); }
          g[j][h] = fmul2(d, wr[j][h]);
          s1 = fadd2(s1, g[j][h]);
          s2 = ffma2(g[j][h], xh[j][h], s2);
        }
      }
      const float m1 = group_sum<G>(s1.x + s1.y) * inv_c, m2 = group_sum<G>(s2.x + s2.y) * inv_c;
      if (!live) continue;
      const float2 nm2 = make_float2(-m2, -m2), nm1 = make_float2(-m1, -m1);
      uint2* dr = reinterpret_cast<uint2*>(out + r * kRowBytes) + gl;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        // rstd * (g - mean(g) - xhat * mean(g * xhat))
        const float2 o0 = fmul2(rs2, fadd2(ffma2(xh[j][0], nm2, g[j][0]), nm1));
        const float2 o1 = fmul2(rs2, fadd2(ffma2(xh[j][1], nm2, g[j][1]), nm1));
        dr[G * j] = make_uint2(f2_to_bf2(o0), f2_to_bf2(o1));
      }
    }
    b200at::fence_proxy_async();
    if (tid == 0) b200at::tma_store_wait_read();
    __syncthreads();
    if (tid == 0) {
      bulk_store(dx + t * tile_rows * C, out, (uint32_t)rows * kRowBytes);
      b200at::tma_store_commit();
      const int64_t tn = t + (int64_t)kRingStages * gridDim.x;
      if (tn < tiles) issue(s, tn);
    }
  }
  if (tid == 0) b200at::tma_store_wait_all();
  if (PGRAD) {
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int c = (gl + G * j) * 4;
      atomicAdd(&red[c], aw[j][0].x); atomicAdd(&red[c + 1], aw[j][0].y);
      atomicAdd(&red[c + 2], aw[j][1].x); atomicAdd(&red[c + 3], aw[j][1].y);
      atomicAdd(&red[C + c], ab[j][0].x); atomicAdd(&red[C + c + 1], ab[j][0].y);
      atomicAdd(&red[C + c + 2], ab[j][1].x); atomicAdd(&red[C + c + 3], ab[j][1].y);
    }
    __syncthreads();
    for (int c = tid; c < C; c += kRingThreads) {
      atomicAdd(dw + c, red[c]);
      atomicAdd(db + c, red[C + c]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ K9
// h = gelu(z + bias) on [M][N] bf16, 16 bytes per thread per load, two loads in flight.  The launcher makes the
// thread count a multiple of n8 = N/8, so a thread keeps its 8 columns (and their bias) for the whole loop.
__global__ void __launch_bounds__(256) bias_gelu_fwd_kernel(const bf16* __restrict__ z, const float* __restrict__ bias,
                                                            bf16* __restrict__ h, int64_t total8, int n8) {
  const int64_t q0 = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * 256;
  const int c0 = (int)(q0 % n8) * 8;
  const float4 b0 = *reinterpret_cast<const float4*>(bias + c0), b1 = *reinterpret_cast<const float4*>(bias + c0 + 4);
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  for (int64_t q = q0; q < total8; q += 2 * stride) {
    const bool two = q + stride < total8;
    const uint4 u0 = __ldcs(reinterpret_cast<const uint4*>(z) + q);
    const uint4 u1 = two ? __ldcs(reinterpret_cast<const uint4*>(z) + q + stride) : u0;
    float f[8];
    unpack4(make_uint2(u0.x, u0.y), f); unpack4(make_uint2(u0.z, u0.w), f + 4);
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const float2 r = b200at_gelu2(fadd2(make_float2(f[k], f[k + 1]), make_float2(bb[k], bb[k + 1])));
      f[k] = r.x; f[k + 1] = r.y;
    }
    uint2 lo = pack4(f), hi = pack4(f + 4);
    reinterpret_cast<uint4*>(h)[q] = make_uint4(lo.x, lo.y, hi.x, hi.y);
    if (two) {
      unpack4(make_uint2(u1.x, u1.y), f); unpack4(make_uint2(u1.z, u1.w), f + 4);
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
      const float2 r = b200at_gelu2(fadd2(make_float2(f[k], f[k + 1]), make_float2(bb[k], bb[k + 1])));
      f[k] = r.x; f[k + 1] = r.y;
    }
      lo = pack4(f); hi = pack4(f + 4);
      reinterpret_cast<uint4*>(h)[q + stride] = make_uint4(lo.x, lo.y, hi.x, hi.y);
    }
  }
}

// dz = dh * gelu'(z + bias);  DBIAS: also dbias[c] += sum_rows dz[.,c] (the pwconv1 bias gradient).  The launcher
// makes the total thread count a multiple of n8, so a thread keeps its 8 columns for the whole grid-stride loop
// and carries their partial sums in registers; one shared-memory + one global atomic pass at the end.
template <bool DBIAS>
__global__ void __launch_bounds__(256) bias_gelu_bwd_kernel(const bf16* __restrict__ dh, const bf16* __restrict__ z,
                                                            const float* __restrict__ bias, bf16* __restrict__ dz,
                                                            float* __restrict__ dbias, int64_t total8, int n8) {
  extern __shared__ float colred[];   // DBIAS: [8 * n8]
  const int64_t q0 = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int c0 = (int)(q0 % n8) * 8;
  const float4 b0 = *reinterpret_cast<const float4*>(bias + c0), b1 = *reinterpret_cast<const float4*>(bias + c0 + 4);
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (DBIAS) {
    for (int c = threadIdx.x; c < 8 * n8; c += 256) colred[c] = 0.f;
    __syncthreads();
  }
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t q = q0; q < total8; q += 2 * stride) {
    const bool two = q + stride < total8;
    uint4 u[2], d[2];
    u[0] = __ldcs(reinterpret_cast<const uint4*>(z) + q);
    d[0] = __ldcs(reinterpret_cast<const uint4*>(dh) + q);
    u[1] = two ? __ldcs(reinterpret_cast<const uint4*>(z) + q + stride) : u[0];
    d[1] = two ? __ldcs(reinterpret_cast<const uint4*>(dh) + q + stride) : d[0];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      if (w == 1 && !two) break;
      float f[8], g[8];
      unpack4(make_uint2(u[w].x, u[w].y), f); unpack4(make_uint2(u[w].z, u[w].w), f + 4);
      unpack4(make_uint2(d[w].x, d[w].y), g); unpack4(make_uint2(d[w].z, d[w].w), g + 4);
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const float2 r = fmul2(make_float2(g[k], g[k + 1]),
                               b200at_gelu_grad2(fadd2(make_float2(f[k], f[k + 1]), make_float2(bb[k], bb[k + 1]))));
        g[k] = r.x; g[k + 1] = r.y;
      }
      const uint2 lo = pack4(g), hi = pack4(g + 4);
      reinterpret_cast<uint4*>(dz)[q + w * stride] = make_uint4(lo.x, lo.y, hi.x, hi.y);
      if (DBIAS) {   // sum what the weight-gradient GEMM will see: the bf16-rounded dz
        float r[8];
        unpack4(lo, r); unpack4(hi, r + 4);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += r[k];
      }
    }
  }
  if (DBIAS) {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&colred[c0 + k], acc[k]);
    __syncthreads();
    for (int c = threadIdx.x; c < 8 * n8; c += 256) atomicAdd(dbias + c, colred[c]);
  }
}

// colsum[c] += sum_rows a[.,c]  on [M][N] bf16 (bias gradients of the second pwconv: column sums of dout)
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ a, float* __restrict__ out, int64_t total8,
                                                     int n8) {
  extern __shared__ float colred[];
  const int64_t q0 = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int c0 = (int)(q0 % n8) * 8;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int c = threadIdx.x; c < 8 * n8; c += 256) colred[c] = 0.f;
  __syncthreads();
  for (int64_t q = q0; q < total8; q += (int64_t)gridDim.x * 256) {
    const uint4 u = __ldcs(reinterpret_cast<const uint4*>(a) + q);
    float f[8];
    unpack4(make_uint2(u.x, u.y), f); unpack4(make_uint2(u.z, u.w), f + 4);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) atomicAdd(&colred[c0 + k], acc[k]);
  __syncthreads();
  for (int c = threadIdx.x; c < 8 * n8; c += 256) atomicAdd(out + c, colred[c]);
}

// out = res + gamma * (z + bias)   (layer scale + residual; models/convnext.py:45-49)
__global__ void __launch_bounds__(256) scale_residual_kernel(const bf16* __restrict__ z, const float* __restrict__ bias,
                                                             const float* __restrict__ gamma,
                                                             const bf16* __restrict__ res, bf16* __restrict__ out,
                                                             int64_t total8, int n8) {
  for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < total8; q += (int64_t)gridDim.x * 256) {
    const int c0 = (int)(q % n8) * 8;
    const uint4 u = __ldcs(reinterpret_cast<const uint4*>(z) + q);
    const uint4 r = __ldcs(reinterpret_cast<const uint4*>(res) + q);
    float f[8], g[8];
    unpack4(make_uint2(u.x, u.y), f); unpack4(make_uint2(u.z, u.w), f + 4);
    unpack4(make_uint2(r.x, r.y), g); unpack4(make_uint2(r.z, r.w), g + 4);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] += gamma[c0 + k] * (f[k] + bias[c0 + k]);
    const uint2 lo = pack4(g), hi = pack4(g + 4);
    reinterpret_cast<uint4*>(out)[q] = make_uint4(lo.x, lo.y, hi.x, hi.y);
  }
}

// dz = dout * gamma  [, dres_out = dres_in + ... handled by the caller]
__global__ void __launch_bounds__(256) scale_bwd_kernel(const bf16* __restrict__ dout, const float* __restrict__ gamma,
                                                        bf16* __restrict__ dz, int64_t total8, int n8) {
  for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < total8; q += (int64_t)gridDim.x * 256) {
    const int c0 = (int)(q % n8) * 8;
    const uint4 d = __ldcs(reinterpret_cast<const uint4*>(dout) + q);
    float g[8];
    unpack4(make_uint2(d.x, d.y), g); unpack4(make_uint2(d.z, d.w), g + 4);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] *= gamma[c0 + k];
    const uint2 lo = pack4(g), hi = pack4(g + 4);
    reinterpret_cast<uint4*>(dz)[q] = make_uint4(lo.x, lo.y, hi.x, hi.y);
  }
}

// c = a + b  (bf16, residual-gradient join)
__global__ void __launch_bounds__(256) add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                                  bf16* __restrict__ c, int64_t total8) {
  for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < total8; q += (int64_t)gridDim.x * 256) {
    const uint4 u = __ldcs(reinterpret_cast<const uint4*>(a) + q);
    const uint4 d = __ldcs(reinterpret_cast<const uint4*>(b) + q);
    float f[8], g[8];
    unpack4(make_uint2(u.x, u.y), f); unpack4(make_uint2(u.z, u.w), f + 4);
    unpack4(make_uint2(d.x, d.y), g); unpack4(make_uint2(d.z, d.w), g + 4);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] += f[k];
    const uint2 lo = pack4(g), hi = pack4(g + 4);
    reinterpret_cast<uint4*>(c)[q] = make_uint4(lo.x, lo.y, hi.x, hi.y);
  }
}

// ------------------------------------------------------------------------------------------------ K7
// Depthwise 7x7, stride 1, pad 3, NHWC bf16 -- persistent, TMA-fed, double buffered, register tiled.
// A CTA works through a contiguous run of tiles; a tile is (NB images) x (TH x TW output pixels) x 32 channels.
// Its (TH+6) x (TW+6) halo box is ONE cp.async.bulk.tensor.4d request by one thread: the tensor map's zero fill
// supplies the padding (negative / past-the-edge coordinates), so there is no staging code, no address math and
// no load latency in the compute warps -- the box of tile i+1 lands in the other buffer while tile i is computed
// (the st.shared staging loop this replaces held 47 % of the warp-stall samples: profiles/r01_dwconv_ncu_v5.txt).
// Thread (cp, t): channel pair cp (16 per CTA) and TWO adjacent output columns.  The thread slides through the 8
// input columns its two outputs touch, keeping the previous column in registers, so every shared-memory load
// (one bf16x2 -> fp32x2) feeds up to 14 packed fp32x2 FMAs: the kernel is bound by the FMA pipe
// (2 cycles per FFMA2), not by instruction issue (the one-column form issued 3 instructions per FFMA2).
constexpr int kDwCh = 32;            // channels per tile (16 bf16x2 lanes)

template <int TH, int TW, int NB>
struct DwTile {
  static constexpr int IH = TH + 6, IW = TW + 6;
  static constexpr int kPairs = (TW / 2) * NB;                           // column pairs = threads / 16
  static constexpr int kThreads = kPairs * 16;
  static constexpr int kTileWords = NB * IH * IW * (kDwCh / 2);          // bf16x2 words per input buffer
  static constexpr int kTileBytes = kTileWords * 4;
  static constexpr int kTilePad = (kTileBytes + 127) / 128 * 128;        // TMA destination: 128-byte aligned
  static constexpr int kWBytes = 49 * (kDwCh / 2) * 8;                   // fp32x2 taps of the channel group
  static constexpr int kSmem = 2 * kTilePad + kWBytes + 64 + 128;        // + barriers + alignment slack
  // resident CTAs per SM the register budget is cut for: the half-height tile keeps 7 instead of 14 accumulator rows
  // per column, which buys a third CTA (21 instead of 14 warps per SM: the full-height form is latency-bound at 2.9
  // active warps per scheduler, profiles/r01_ncu_dwconv_stem0_v11_summary.txt)
  static constexpr int kMinCtas = (TH == 7 && TW >= 14) ? 3 : 2;
};

struct DwParams {
  const float* wt;      // [49][C] taps
  const float* bias;    // [C] or null
  const bf16* add;      // [B][H][W][C] or null
  bf16* y;
  int B, H, W, C;
  int tiles_w, tiles_h, groups_b;   // spatial tiles per image, image groups
  int total_tiles;
};

template <int TH, int TW, int NB, bool BIAS, bool ADD>
__global__ void __launch_bounds__(DwTile<TH, TW, NB>::kThreads, DwTile<TH, TW, NB>::kMinCtas)
    dwconv7_kernel(const __grid_constant__ CUtensorMap map_x, const DwParams p) {
  typedef DwTile<TH, TW, NB> T;
  extern __shared__ __align__(128) uint8_t dw_smem_raw[];
  // 128-byte aligned TMA destinations; offset arithmetic on the __shared__ array keeps the loads LDS (not generic LD)
  uint8_t* smem = dw_smem_raw + ((128u - (b200at::smem_u32(dw_smem_raw) & 127u)) & 127u);
  float2* wk = reinterpret_cast<float2*>(smem + 2 * T::kTilePad);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * T::kTilePad + T::kWBytes);

  const int tid = threadIdx.x;
  const int cp = tid & 15, t = tid >> 4;
  const int img = t / (TW / 2), col = 2 * (t % (TW / 2));
  // this CTA's contiguous run of tiles (channel group slowest, so the taps are reloaded at most a few times)
  const int q = p.total_tiles / (int)gridDim.x, rem = p.total_tiles % (int)gridDim.x;
  const int b = (int)blockIdx.x;
  const int first = b * q + (b < rem ? b : rem);
  const int count = q + (b < rem ? 1 : 0);
  const int spatial = p.tiles_w * p.tiles_h * p.groups_b;

  auto decode = [&](int tile, int& cg, int& n0, int& h0, int& w0) {
    cg = tile / spatial;
    int s = tile - cg * spatial;
    const int tw = s % p.tiles_w; s /= p.tiles_w;
    const int th = s % p.tiles_h; s /= p.tiles_h;
    n0 = s * NB; h0 = th * TH; w0 = tw * TW;
  };
  if (tid == 0) {
    b200at::tma_prefetch_desc(&map_x);
    b200at::mbar_init(&bars[0], 1);
    b200at::mbar_init(&bars[1], 1);
    b200at::mbar_fence_init();
    if (count > 0) {
      int cg, n0, h0, w0;
      decode(first, cg, n0, h0, w0);
      b200at::mbar_expect_tx(&bars[0], T::kTileBytes);
      b200at::tma_load_4d(&map_x, &bars[0], smem, cg * kDwCh, w0 - 3, h0 - 3, n0);
    }
  }
  __syncthreads();

  int cur_cg = -1;
  for (int it = 0; it < count; ++it) {
    const int cur = it & 1;
    int cg, n0, h0, w0;
    decode(first + it, cg, n0, h0, w0);
    const int c0 = cg * kDwCh;
    if (cg != cur_cg) {   // new channel group (CTA-uniform, rare): reload the taps once the previous tile is done
      cur_cg = cg;
      __syncthreads();
      for (int k = tid; k < 49 * (kDwCh / 2); k += T::kThreads) {
        const int tap = k >> 4, c2 = k & 15;
        wk[k] = *reinterpret_cast<const float2*>(p.wt + (int64_t)tap * p.C + c0 + c2 * 2);
      }
    }
    b200at::mbar_wait(&bars[cur], (uint32_t)((it >> 1) & 1));   // this tile's box has landed
    __syncthreads();   // taps visible; everybody is done with the previous tile => the other buffer is free
    if (tid == 0 && it + 1 < count) {
      int cg1, n1, h1, w1;
      decode(first + it + 1, cg1, n1, h1, w1);
      b200at::mbar_expect_tx(&bars[cur ^ 1], T::kTileBytes);
      b200at::tma_load_4d(&map_x, &bars[cur ^ 1], smem + (cur ^ 1) * T::kTilePad, cg1 * kDwCh, w1 - 3, h1 - 3, n1);
    }
    const bf162* tile = reinterpret_cast<const bf162*>(smem + cur * T::kTilePad) +
                        (img * (T::IH * T::IW) + col) * (kDwCh / 2) + cp;
    float2 acc0[TH], acc1[TH], vp[T::IH];
    const float2 b2 = BIAS ? make_float2(p.bias[c0 + cp * 2], p.bias[c0 + cp * 2 + 1]) : make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < TH; ++r) { acc0[r] = b2; acc1[r] = b2; }
#pragma unroll
    for (int r = 0; r < T::IH; ++r) vp[r] = __bfloat1622float2(tile[(r * T::IW) * (kDwCh / 2)]);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      float2 wj[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) wj[i] = wk[(i * 7 + j) * (kDwCh / 2) + cp];
#pragma unroll
      for (int r = 0; r < T::IH; ++r) {
        // input column (col + j) feeds output column col, input column (col + j + 1) output column col + 1
        const float2 vn = __bfloat1622float2(tile[(r * T::IW + j + 1) * (kDwCh / 2)]);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          const int o = r - i;  // output row fed by input row r through tap row i
          if (o >= 0 && o < TH) {
            acc0[o] = ffma2(vp[r], wj[i], acc0[o]);
            acc1[o] = ffma2(vn, wj[i], acc1[o]);
          }
        }
        vp[r] = vn;
      }
    }
    const int n = n0 + img, w = w0 + col;
    if (n < p.B && w < p.W) {
      const bool second = w + 1 < p.W;
      const int64_t off0 = (((int64_t)n * p.H + h0) * p.W + w) * p.C + c0 + cp * 2;
      const int64_t rstride = (int64_t)p.W * p.C;
      if (ADD) {
        // residual-gradient join fused into the input-gradient pass.  All loads first (read-only path), then the
        // stores: interleaved, every load had to wait behind the previous store (y and add may alias as far as the
        // compiler knows), which serialised 2*TH memory round trips per tile (dgrad 200 us vs forward 131 us).
        uint32_t a0[TH], a1[TH];
#pragma unroll
        for (int r = 0; r < TH; ++r) {
          a0[r] = 0u; a1[r] = 0u;
          if (h0 + r < p.H) {
            a0[r] = __ldg(reinterpret_cast<const uint32_t*>(p.add + off0 + r * rstride));
            if (second) a1[r] = __ldg(reinterpret_cast<const uint32_t*>(p.add + off0 + r * rstride + p.C));
          }
        }
#pragma unroll
        for (int r = 0; r < TH; ++r) {
          const float2 f0 = __bfloat1622float2(*reinterpret_cast<const bf162*>(&a0[r]));
          const float2 f1 = __bfloat1622float2(*reinterpret_cast<const bf162*>(&a1[r]));
          acc0[r].x += f0.x; acc0[r].y += f0.y;
          acc1[r].x += f1.x; acc1[r].y += f1.y;
        }
      }
#pragma unroll
      for (int r = 0; r < TH; ++r) {
        if (h0 + r < p.H) {
          const int64_t off = off0 + r * rstride;
          *reinterpret_cast<bf162*>(p.y + off) = __floats2bfloat162_rn(acc0[r].x, acc0[r].y);
          if (second) *reinterpret_cast<bf162*>(p.y + off + p.C) = __floats2bfloat162_rn(acc1[r].x, acc1[r].y);
        }
      }
    }
  }
}

// tile staging of the weight-gradient kernel below (plain loads; that kernel is not TMA-fed yet)
constexpr int kDwTile = 14;
constexpr int kDwIn = kDwTile + 6;   // 20

// weight gradient: dw[tap][c] += sum_{pixels} dy[p][c] * x[p + tap][c];  db[c] += sum dy
//
// Round-1 form: one CTA per (image, 14 x 14 tile, 32 channels), the 16 column-threads' partial sums of a tap column reduced
// through shared memory for every tile -- 15 barriers per tile on 8 warps, 28 % of the FFMA2 rate (210 us at 56 x 56 x 96).
// This form is persistent: a CTA owns a channel group and walks every `splits`-th (image, tile); thread (channel pair,
// output column) keeps all 49 tap sums of its column in registers ACROSS tiles (98 registers; 7 independent accumulators per
// loaded value keep the FMA pipe busy from two warps per scheduler), the x and dy tiles arrive by cp.async one tile
// ahead (zero fill = padding and ragged edges), and the cross-column reduction happens once per CTA.  One barrier per tile.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

constexpr int kWgXBytes = kDwIn * kDwIn * kDwCh * 2;          // 25600: [20][20][32 ch] bf16
constexpr int kWgGBytes = kDwTile * kDwTile * kDwCh * 2;      // 12544: [14][14][32 ch] bf16
constexpr int kWgBuf = kWgXBytes + kWgGBytes;

__global__ void __launch_bounds__(256, 1) dwconv7_wgrad_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                               float* __restrict__ dw, float* __restrict__ db, int H,
                                                               int W, int C, int tiles_w, int tiles_h, int spatial,
                                                               int splits) {
  extern __shared__ __align__(16) uint8_t wg_smem[];            // [2][kWgBuf], then red7 [7][16][16] float2
  float2 (*red7)[16][kDwCh / 2] = reinterpret_cast<float2 (*)[16][kDwCh / 2]>(wg_smem + 2 * kWgBuf);
  const uint32_t smem_a = b200at::smem_u32(wg_smem);
  const int cp = threadIdx.x & 15, t = threadIdx.x >> 4;
  const int cgroups = C / kDwCh;
  const int cg = blockIdx.x % cgroups, split = blockIdx.x / cgroups;
  const int c0 = cg * kDwCh;
  auto request = [&](int sp, int buf) {                          // both tiles of item `sp` into buffer `buf`
    int bid = sp;
    const int tw = bid % tiles_w; bid /= tiles_w;
    const int th = bid % tiles_h; bid /= tiles_h;
    const int64_t img = (int64_t)bid * H * W * C;
    const int h0 = th * kDwTile, w0 = tw * kDwTile;
    const uint32_t xa = smem_a + (uint32_t)(buf * kWgBuf), ga = xa + kWgXBytes;
    for (int q = threadIdx.x; q < kDwIn * kDwIn * 4; q += 256) {
      const int pix = q >> 2, piece = q & 3;
      const int r = pix / kDwIn, c = pix - r * kDwIn;
      const int hh = h0 + r - 3, ww = w0 + c - 3;
      const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
      cp_async16(xa + (uint32_t)(q * 16), ok ? (const void*)(x + img + ((int64_t)hh * W + ww) * C + c0 + piece * 8) : (const void*)x, ok);
    }
    for (int q = threadIdx.x; q < kDwTile * kDwTile * 4; q += 256) {
      const int pix = q >> 2, piece = q & 3;
      const int r = pix / kDwTile, c = pix - r * kDwTile;
      const int hh = h0 + r, ww = w0 + c;
      const bool ok = hh < H && ww < W;
      cp_async16(ga + (uint32_t)(q * 16), ok ? (const void*)(dy + img + ((int64_t)hh * W + ww) * C + c0 + piece * 8) : (const void*)dy, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float2 acc[7][7];                                              // [tap column j][tap row i]
#pragma unroll
  for (int j = 0; j < 7; ++j) {
#pragma unroll
    for (int i = 0; i < 7; ++i) acc[j][i] = make_float2(0.f, 0.f);
  }
  float2 gsum = make_float2(0.f, 0.f);
  if (split < spatial) request(split, 0);
  int k = 0;
  for (int sp = split; sp < spatial; sp += splits, ++k) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                             // tile k has landed; tile k - 1 has been consumed
    if (sp + splits < spatial) request(sp + splits, (k + 1) & 1);
    if (t < kDwTile) {
      const bf162* xt = reinterpret_cast<const bf162*>(wg_smem + (k & 1) * kWgBuf);
      const bf162* gt = reinterpret_cast<const bf162*>(wg_smem + (k & 1) * kWgBuf + kWgXBytes);
      float2 g[kDwTile];
#pragma unroll
      for (int r = 0; r < kDwTile; ++r) {
        g[r] = __bfloat1622float2(gt[(r * kDwTile + t) * (kDwCh / 2) + cp]);
        gsum.x += g[r].x; gsum.y += g[r].y;
      }
#pragma unroll
      for (int j = 0; j < 7; ++j) {
#pragma unroll
        for (int r = 0; r < kDwIn; ++r) {
          const float2 v = __bfloat1622float2(xt[(r * kDwIn + t + j) * (kDwCh / 2) + cp]);
#pragma unroll
          for (int i = 0; i < 7; ++i) {
            const int o = r - i;
            if (o >= 0 && o < kDwTile) acc[j][i] = ffma2(v, g[o], acc[j][i]);
          }
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // once per CTA: the 16 column-threads' sums of a tap column through shared memory, 7 x 16 threads add them up
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 7; ++i) red7[i][t][cp] = acc[j][i];
    __syncthreads();
    if (threadIdx.x < 7 * 16) {
      const int i = threadIdx.x >> 4, c2 = threadIdx.x & 15;
      float2 s = make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 16; ++q) { s.x += red7[i][q][c2].x; s.y += red7[i][q][c2].y; }
      atomicAdd(dw + (i * 7 + j) * C + c0 + c2 * 2, s.x);
      atomicAdd(dw + (i * 7 + j) * C + c0 + c2 * 2 + 1, s.y);
    }
  }
  __syncthreads();
  red7[0][t][cp] = gsum;
  __syncthreads();
  if (t == 0) {
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 16; ++q) { s.x += red7[0][q][cp].x; s.y += red7[0][q][cp].y; }
    atomicAdd(db + c0 + cp * 2, s.x);
    atomicAdd(db + c0 + cp * 2 + 1, s.y);
  }
}

// The round-1 form (one CTA per tile, cross-column reduction per tile): kept for the 7 x 7 maps of the last stage, where a
// 14 x 14 tile is three quarters padding and the persistent form's 8 warps per SM lose to 32 (61 vs 88 us at 7 x 7 x 768).
__global__ void __launch_bounds__(256) dwconv7_wgrad_tile_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                            float* __restrict__ dw, float* __restrict__ db, int H,
                                                            int W, int C, int tiles_w, int tiles_h) {
  __shared__ __align__(16) bf162 tile[kDwIn][kDwIn][kDwCh / 2];
  __shared__ float2 red[16][kDwCh / 2];
  __shared__ float2 red7[7][16][kDwCh / 2];
  const int cp = threadIdx.x & 15, t = threadIdx.x >> 4;
  const int cgroups = C / kDwCh;
  int bid = blockIdx.x;
  const int cg = bid % cgroups; bid /= cgroups;
  const int tw = bid % tiles_w; bid /= tiles_w;
  const int th = bid % tiles_h; bid /= tiles_h;
  const int n = bid;
  const int h0 = th * kDwTile, w0 = tw * kDwTile, c0 = cg * kDwCh;
  const bf16* xin = x + (int64_t)n * H * W * C;
  const bf16* gin = dy + (int64_t)n * H * W * C;
  // stage the (20 x 20) halo tile: 4 consecutive lanes fetch the 64 contiguous bytes of one pixel (16 B each)
  for (int q = threadIdx.x; q < kDwIn * kDwIn * 4; q += 256) {
    const int pix = q >> 2, piece = q & 3;
    const int r = pix / kDwIn, c = pix % kDwIn;
    const int hh = h0 + r - 3, ww = w0 + c - 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (hh >= 0 && hh < H && ww >= 0 && ww < W)
      v = __ldg(reinterpret_cast<const uint4*>(xin + ((int64_t)hh * W + ww) * C + c0 + piece * 8));
    *reinterpret_cast<uint4*>(&tile[r][c][piece * 4]) = v;
  }
  __syncthreads();
  // this thread's dy column (zero outside the image)
  float2 g[kDwTile];
  float2 gsum = make_float2(0.f, 0.f);
  const bool active = t < kDwTile && w0 + t < W;
#pragma unroll
  for (int r = 0; r < kDwTile; ++r) {
    g[r] = make_float2(0.f, 0.f);
    if (active && h0 + r < H)
      g[r] = __bfloat1622float2(*reinterpret_cast<const bf162*>(gin + ((int64_t)(h0 + r) * W + w0 + t) * C + c0 + cp * 2));
    gsum.x += g[r].x; gsum.y += g[r].y;
  }
  // taps are produced one tap-column j at a time (7 taps per thread); the 16 column-threads' partial sums of all 7
  // go to shared memory at once and 7 x 16 threads add them up: 2 barriers per tap column instead of 2 per tap
  for (int j = 0; j < 7; ++j) {
    float2 a[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) a[i] = make_float2(0.f, 0.f);
    if (active) {
#pragma unroll
      for (int r = 0; r < kDwIn; ++r) {
        const float2 v = __bfloat1622float2(tile[r][t + j][cp]);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          const int o = r - i;
          if (o >= 0 && o < kDwTile) a[i] = ffma2(v, g[o], a[i]);
        }
      }
    }
    __syncthreads();                       // previous column's sums have been read
#pragma unroll
    for (int i = 0; i < 7; ++i) red7[i][t][cp] = a[i];
    __syncthreads();
    if (threadIdx.x < 7 * 16) {
      const int i = threadIdx.x >> 4, c2 = threadIdx.x & 15;
      float2 s = make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 16; ++q) { s.x += red7[i][q][c2].x; s.y += red7[i][q][c2].y; }
      atomicAdd(dw + (i * 7 + j) * C + c0 + c2 * 2, s.x);
      atomicAdd(dw + (i * 7 + j) * C + c0 + c2 * 2 + 1, s.y);
    }
  }
  __syncthreads();
  red[t][cp] = gsum;
  __syncthreads();
  if (t == 0) {
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 16; ++q) { s.x += red[q][cp].x; s.y += red[q][cp].y; }
    atomicAdd(db + c0 + cp * 2, s.x);
    atomicAdd(db + c0 + cp * 2 + 1, s.y);
  }
}

// ------------------------------------------------------------------------------------------------ K9b
// Per-optimiser-step weight preparation of one block's MLP (ops._prepared) in ONE launch instead of seven torch ops:
//   w1b = bf16(W1) [4C][C],  w1t = w1b^T [C][4C],  w2g = bf16(gamma[:,None] * W2) [C][4C],  w2gt = w2g^T [4C][C],
//   b2g = gamma * b2.  32 x 32 tiles through a padded shared-memory tile (coalesced reads and transposed writes).
__global__ void __launch_bounds__(256) prepare_mlp_weights_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                                                  const float* __restrict__ b2, const float* __restrict__ gamma,
                                                                  bf16* __restrict__ w1b, bf16* __restrict__ w1t,
                                                                  bf16* __restrict__ w2g, bf16* __restrict__ w2gt,
                                                                  float* __restrict__ b2g, int C) {
  __shared__ float tile[32][33];
  const int H = 4 * C;
  const int tiles_r1 = H / 32, tiles_c1 = C / 32;          // W1 is [H][C], W2 is [C][H]
  const int n1 = tiles_r1 * tiles_c1;
  int t = blockIdx.x;
  const bool second = t >= n1;
  if (second) t -= n1;
  const int rows = second ? C : H, cols = second ? H : C;
  const int tr = t / (cols / 32), tc = t % (cols / 32);
  const float* src = second ? w2 : w1;
  bf16* dst = second ? w2g : w1b;
  bf16* dst_t = second ? w2gt : w1t;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = tr * 32 + ty + 8 * i, c = tc * 32 + tx;
    float v = src[(int64_t)r * cols + c];
    if (second) v *= gamma[r];
    const bf16 h = __float2bfloat16_rn(v);
    dst[(int64_t)r * cols + c] = h;
    tile[ty + 8 * i][tx] = __bfloat162float(h);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = tc * 32 + ty + 8 * i, c = tr * 32 + tx;  // transposed matrix is [cols][rows]
    dst_t[(int64_t)r * rows + c] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
  }
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += 256) b2g[c] = gamma[c] * b2[c];
}

// Tail of one block's parameter gradients (ops._ConvNeXtBlock.backward) in ONE launch instead of eight torch ops.  The
// weight-gradient GEMM yields dW2g, the gradient w.r.t. the layer-scale-folded matrix gamma[:,None] * W2; `col` is the column
// sum of the block's upstream gradient:
//   dW2 = gamma[:,None] * dW2g,   db2 = col * gamma,   dgamma = rowsum(dW2g * W2) + col * b2.      One warp per row.
__global__ void __launch_bounds__(256) finish_mlp_grads_kernel(const float* __restrict__ dw2g, const float* __restrict__ w2,
                                                               const float* __restrict__ col, const float* __restrict__ b2,
                                                               const float* __restrict__ gamma, float* __restrict__ dw2,
                                                               float* __restrict__ db2, float* __restrict__ dgamma, int C) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= C) return;
  const int H = 4 * C;
  const float gm = gamma[row];
  const float4* a = reinterpret_cast<const float4*>(dw2g + (int64_t)row * H);
  const float4* w = reinterpret_cast<const float4*>(w2 + (int64_t)row * H);
  float4* o = reinterpret_cast<float4*>(dw2 + (int64_t)row * H);
  float acc = 0.f;
  for (int k = lane; k < H / 4; k += 32) {
    const float4 g4 = a[k], w4 = w[k];
    acc += (g4.x * w4.x + g4.y * w4.y) + (g4.z * w4.z + g4.w * w4.w);
    o[k] = make_float4(gm * g4.x, gm * g4.y, gm * g4.z, gm * g4.w);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    const float cl = col[row];
    dgamma[row] = acc + cl * b2[row];
    db2[row] = cl * gm;
  }
}

inline int flat_grid(int64_t total8) {
  int64_t g = (total8 + 255) / 256;
  const int64_t cap = 148 * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

// grid whose thread count (256 per CTA) is a multiple of n8, so a grid-stride thread keeps its column group
inline int column_grid(int64_t total8, int n8, int ctas_per_sm) {
  int a = n8, b = 256;
  while (b) { const int t = a % b; a = b; b = t; }     // a = gcd(n8, 256)
  const int unit = n8 / a;                             // grid must be a multiple of this
  int64_t want = (total8 + 255) / 256;
  const int64_t cap = 148 * (int64_t)ctas_per_sm;
  if (want > cap) want = cap;
  int64_t g = (want + unit - 1) / unit * unit;
  return (int)(g < unit ? unit : g);
}

// (G, VPL) with the best lane utilisation for nv = C/4 vectors per row; ties go to the wider group
inline void ln_shape(int nv, int& G, int& VPL) {
  double best = -1.0;
  G = 32; VPL = (nv + 31) / 32;
  for (int g : {32, 16, 8, 4}) {
    const int v = (nv + g - 1) / g;
    if (v > 12) continue;
    const double u = (double)nv / (double)(g * v);
    if (u > best + 1e-9) { best = u; G = g; VPL = v; }
  }
}
inline int env_int_ln(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e != nullptr && e[0] != 0) ? atoi(e) : dflt;
}
inline int ln_grid(int64_t M, int G) {
  const int rows = kLnThreads / G;
  int64_t g = (M + rows - 1) / rows;
  static const int per_sm = env_int_ln("B200AT_LN_CTAS", 4);
  const int64_t cap = 148 * (int64_t)per_sm;   // 4: one resident wave of the <= 64-register forms (weights in shared memory)
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

#define B200AT_LN_FOR_VPL(G, X) X(G, 1) X(G, 2) X(G, 3) X(G, 4) X(G, 5) X(G, 6) X(G, 7) X(G, 8) X(G, 9) X(G, 10) X(G, 11) X(G, 12)
#define B200AT_LN_FOR_ALL(X) B200AT_LN_FOR_VPL(32, X) B200AT_LN_FOR_VPL(16, X) B200AT_LN_FOR_VPL(8, X) B200AT_LN_FOR_VPL(4, X)

// (G, VPL) pairs the pipelined kernels are built for: every ConvNeXt-T/S/B/L and ViT-S width (C = 48 ... 1536)
#define B200AT_RING_FOR_VPL(G, X) X(G, 1) X(G, 2) X(G, 3) X(G, 4) X(G, 6)
#define B200AT_RING_FOR_ALL(X) B200AT_RING_FOR_VPL(32, X) B200AT_RING_FOR_VPL(16, X) B200AT_RING_FOR_VPL(8, X) B200AT_RING_FOR_VPL(4, X)
inline bool ring_shape_ok(int G, int VPL, int C) { return (VPL == 1 || VPL == 2 || VPL == 3 || VPL == 4 || VPL == 6) && C == 4 * G * VPL; }
inline bool use_ring(int64_t M, int C) {
  static const bool on = [] { const char* e = getenv("B200AT_LN_RING"); return e == nullptr || e[0] != '0'; }();
  return on && C % 8 == 0 && M * C >= (int64_t)1 << 24;   // >= 32 MB of rows (stages 0-1, stems): below that the
                                                           // launch is latency-bound and the plain kernels are as fast
}
// rows per tile: whole passes of the CTA (kRingThreads / G rows each), ~12 KB, a multiple of 4 rows
inline int ring_tile_rows(int G, int C) {
  const int pass = kRingThreads / G;
  int passes = 12288 / (pass * C * 2);
  if (passes < 1) passes = 1;
  return pass * passes;
}
inline int ring_grid(int64_t M, int tile_rows, size_t smem) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);     // 4 x 512 threads = the SM's 2048
  const int64_t tiles = (M + tile_rows - 1) / tile_rows, cap = (int64_t)sms * per_sm;
  return (int)(tiles < cap ? tiles : cap);
}

template <bool GELU>
int launch_ln_fwd(const bf16* x, const float* w, const float* b, bf16* y, float* mean, float* rstd, int64_t M, int C,
                  float eps, cudaStream_t s, int pH = 0, int pW = 0, const float* pre_bias = nullptr) {
  int G, VPL;
  ln_shape(C / 4, G, VPL);
  if (pW == 0 && use_ring(M, C) && ring_shape_ok(G, VPL, C)) {
    const int tr = ring_tile_rows(G, C);
    const size_t smem = (size_t)(kRingStages + 2) * (((size_t)tr * C * 2 + 127) & ~(size_t)127) + 64;
    if (smem <= 200 * 1024) {
      const int grid = ring_grid(M, tr, smem);
#define B200AT_CASE(GG, V)                                                                                               \
      if (G == GG && VPL == V) {                                                                                         \
        static b200at::SmemConfig conf;                                                                            \
        cudaError_t e = b200at::ensure_dynamic_smem(ln_fwd_ring_kernel<GG, V, GELU>, (int)smem, conf);                   \
        if (e != cudaSuccess) return (int)e;                                                                             \
        ln_fwd_ring_kernel<GG, V, GELU><<<grid, kRingThreads, smem, s>>>(x, w, b, y, mean, rstd, M, eps, tr, pre_bias);        \
        return (int)cudaGetLastError();                                                                                  \
      }
      B200AT_RING_FOR_ALL(B200AT_CASE)
#undef B200AT_CASE
    }
  }
  const int g = ln_grid(M, G);
#define B200AT_CASE(GG, V) \
  if (G == GG && VPL == V) { ln_fwd_kernel<GG, V, GELU><<<g, kLnThreads, 0, s>>>(x, w, b, y, mean, rstd, M, C, eps, pH, pW, pre_bias); return (int)cudaGetLastError(); }
  B200AT_LN_FOR_ALL(B200AT_CASE)
#undef B200AT_CASE
  return (int)cudaErrorInvalidValue;
}

template <bool GELU, bool PGRAD>
int launch_ln_bwd(const bf16* dy, const bf16* x, const float* w, const float* b, const float* mean, const float* rstd,
                  bf16* dx, float* dw, float* db, int64_t M, int C, cudaStream_t s, int pH = 0, int pW = 0,
                  const float* pre_bias = nullptr) {
  int G, VPL;
  ln_shape(C / 4, G, VPL);
  if (pW == 0 && use_ring(M, C) && ring_shape_ok(G, VPL, C) && (reinterpret_cast<uintptr_t>(mean) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(rstd) & 15) == 0) {
    const int tr = ring_tile_rows(G, C);
    const size_t tb = (size_t)tr * C * 2;
    const size_t smem = (size_t)kRingStages * ((2 * tb + 8 * (size_t)tr + 127) & ~(size_t)127) + 2 * ((tb + 127) & ~(size_t)127) +
                        64 + (PGRAD ? sizeof(float) * 2 * C : 0);
    if (smem <= 200 * 1024) {
      const int grid = ring_grid(M, tr, smem);
#define B200AT_CASE(GG, V)                                                                                               \
      if (G == GG && VPL == V) {                                                                                         \
        static b200at::SmemConfig conf;                                                                            \
        cudaError_t e = b200at::ensure_dynamic_smem(ln_bwd_ring_kernel<GG, V, GELU, PGRAD>, (int)smem, conf);            \
        if (e != cudaSuccess) return (int)e;                                                                             \
        ln_bwd_ring_kernel<GG, V, GELU, PGRAD><<<grid, kRingThreads, smem, s>>>(dy, x, w, b, mean, rstd, dx, dw, db, M, tr, pre_bias); \
        return (int)cudaGetLastError();                                                                                  \
      }
      B200AT_RING_FOR_ALL(B200AT_CASE)
#undef B200AT_CASE
    }
  }
  const int g = ln_grid(M, G);
  const size_t sm = PGRAD ? sizeof(float) * 2 * C : 0;
#define B200AT_CASE(GG, V) \
  if (G == GG && VPL == V) { ln_bwd_kernel<GG, V, GELU, PGRAD><<<g, kLnThreads, sm, s>>>(dy, x, w, b, mean, rstd, dx, dw, db, M, C, pH, pW, pre_bias); return (int)cudaGetLastError(); }
  B200AT_LN_FOR_ALL(B200AT_CASE)
#undef B200AT_CASE
  return (int)cudaErrorInvalidValue;
}

template <int TH, int TW, int NB>
int launch_dwconv(const void* x, const float* wt, const float* bias, const void* add, void* y, int64_t B, int64_t H,
                  int64_t W, int64_t C, cudaStream_t s) {
  typedef DwTile<TH, TW, NB> T;
  DwParams p;
  p.wt = wt; p.bias = bias; p.add = (const bf16*)add; p.y = (bf16*)y;
  p.B = (int)B; p.H = (int)H; p.W = (int)W; p.C = (int)C;
  p.tiles_w = (int)((W + TW - 1) / TW); p.tiles_h = (int)((H + TH - 1) / TH); p.groups_b = (int)((B + NB - 1) / NB);
  const int64_t total = (int64_t)p.tiles_w * p.tiles_h * p.groups_b * (C / kDwCh);
  if (total > 0x7fffffff) return (int)cudaErrorInvalidValue;
  p.total_tiles = (int)total;
  CUtensorMap map;
  if (!b200at::make_map_nhwc_bf16(&map, x, B, H, W, C, kDwCh, T::IW, T::IH, NB)) return (int)cudaErrorUnknown;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = (228 * 1024) / (T::kSmem + 1024);      // resident CTAs per SM by shared memory (launch bounds: 2)
  per_sm = per_sm < 1 ? 1 : (per_sm > T::kMinCtas ? T::kMinCtas : per_sm);
  const int64_t cap = (int64_t)sms * per_sm;
  const int grid = (int)(total < cap ? total : cap);
#define B200AT_DW(BI, AD)                                                                                     \
  do {                                                                                                        \
    static b200at::SmemConfig configured;                                                               \
    cudaError_t e = b200at::ensure_dynamic_smem(dwconv7_kernel<TH, TW, NB, BI, AD>, T::kSmem, configured);    \
    if (e != cudaSuccess) return (int)e;                                                                      \
    dwconv7_kernel<TH, TW, NB, BI, AD><<<grid, T::kThreads, T::kSmem, s>>>(map, p);                            \
  } while (0)
  if (bias) { if (add) B200AT_DW(true, true); else B200AT_DW(true, false); }
  else { if (add) B200AT_DW(false, true); else B200AT_DW(false, false); }
#undef B200AT_DW
  return (int)cudaGetLastError();
}

}  // namespace

// b200at_dwconv_mma.cu: -1 = shape not covered, else the launch's cudaError_t
int b200at_dwconv7_mma_launch(const void* x, const float* wt, const float* bias, const void* add, void* y, int64_t B,
                              int64_t H, int64_t W, int64_t C, void* stream);

extern "C" {

int b200at_ln_fwd(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd, int64_t M,
                  int64_t C, float eps, int fuse_gelu, void* stream) {
  if (M <= 0) return 0;
  if (C % 4 || C > 1536) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  return fuse_gelu ? launch_ln_fwd<true>((const bf16*)x, w, b, (bf16*)y, mean, rstd, M, (int)C, eps, s)
                   : launch_ln_fwd<false>((const bf16*)x, w, b, (bf16*)y, mean, rstd, M, (int)C, eps, s);
}

int b200at_ln_bwd(const void* dy, const void* x, const float* w, const float* b, const float* mean, const float* rstd,
                  void* dx, float* dw, float* db, int64_t M, int64_t C, int fuse_gelu, void* stream) {
  if (M <= 0) return 0;
  if (C % 4 || C > 1536) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  const bool pg = dw != nullptr;
  if (pg != (db != nullptr)) return (int)cudaErrorInvalidValue;
#define B200AT_GO(G, P) launch_ln_bwd<G, P>((const bf16*)dy, (const bf16*)x, w, b, mean, rstd, (bf16*)dx, dw, db, M, (int)C, s)
  if (fuse_gelu) return pg ? B200AT_GO(true, true) : B200AT_GO(true, false);
  return pg ? B200AT_GO(false, true) : B200AT_GO(false, false);
#undef B200AT_GO
}

int b200at_ln_fwd_bias(const void* x, const float* pre_bias, const float* w, const float* b, void* y, float* mean,
                       float* rstd, int64_t M, int64_t C, float eps, int fuse_gelu, void* stream) {
  if (M <= 0) return 0;
  if (C % 4 || C > 1536) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  return fuse_gelu ? launch_ln_fwd<true>((const bf16*)x, w, b, (bf16*)y, mean, rstd, M, (int)C, eps, s, 0, 0, pre_bias)
                   : launch_ln_fwd<false>((const bf16*)x, w, b, (bf16*)y, mean, rstd, M, (int)C, eps, s, 0, 0, pre_bias);
}

int b200at_ln_bwd_bias(const void* dy, const void* x, const float* pre_bias, const float* w, const float* b,
                       const float* mean, const float* rstd, void* dx, float* dw, float* db, int64_t M, int64_t C,
                       int fuse_gelu, void* stream) {
  if (M <= 0) return 0;
  if (C % 4 || C > 1536) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  const bool pg = dw != nullptr;
  if (pg != (db != nullptr)) return (int)cudaErrorInvalidValue;
#define B200AT_GO(G, P) launch_ln_bwd<G, P>((const bf16*)dy, (const bf16*)x, w, b, mean, rstd, (bf16*)dx, dw, db, M, (int)C, s, 0, 0, pre_bias)
  if (fuse_gelu) return pg ? B200AT_GO(true, true) : B200AT_GO(true, false);
  return pg ? B200AT_GO(false, true) : B200AT_GO(false, false);
#undef B200AT_GO
}

int b200at_ln_fwd_patch2(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd, int64_t B,
                         int64_t H, int64_t W, int64_t C, float eps, void* stream) {
  if (B <= 0) return 0;
  if (C % 4 || C > 1536 || H % 2 || W % 2 || H <= 0 || W <= 0) return (int)cudaErrorInvalidValue;
  return launch_ln_fwd<false>((const bf16*)x, w, b, (bf16*)y, mean, rstd, B * H * W, (int)C, eps, (cudaStream_t)stream,
                              (int)H, (int)W);
}

int b200at_ln_bwd_patch2(const void* dy, const void* x, const float* w, const float* b, const float* mean,
                         const float* rstd, void* dx, float* dw, float* db, int64_t B, int64_t H, int64_t W, int64_t C,
                         void* stream) {
  if (B <= 0) return 0;
  if (C % 4 || C > 1536 || H % 2 || W % 2 || H <= 0 || W <= 0) return (int)cudaErrorInvalidValue;
  if ((dw != nullptr) != (db != nullptr)) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t M = B * H * W;
  return dw ? launch_ln_bwd<false, true>((const bf16*)dy, (const bf16*)x, w, b, mean, rstd, (bf16*)dx, dw, db, M, (int)C,
                                         s, (int)H, (int)W)
            : launch_ln_bwd<false, false>((const bf16*)dy, (const bf16*)x, w, b, mean, rstd, (bf16*)dx, dw, db, M, (int)C,
                                          s, (int)H, (int)W);
}

int b200at_bias_gelu_fwd(const void* z, const float* bias, void* h, int64_t M, int64_t N, void* stream) {
  if (M <= 0) return 0;
  if (N % 8) return (int)cudaErrorInvalidValue;
  const int64_t total8 = M * N / 8;
  bias_gelu_fwd_kernel<<<column_grid(total8, (int)(N / 8), 16), 256, 0, (cudaStream_t)stream>>>((const bf16*)z, bias, (bf16*)h, total8, (int)(N / 8));
  return (int)cudaGetLastError();
}

int b200at_bias_gelu_bwd(const void* dh, const void* z, const float* bias, void* dz, float* dbias, int64_t M, int64_t N,
                         void* stream) {
  if (M <= 0) return 0;
  if (N % 8 || N > 8192) return (int)cudaErrorInvalidValue;
  const int64_t total8 = M * N / 8;
  const int n8 = (int)(N / 8);
  cudaStream_t s = (cudaStream_t)stream;
  if (dbias)
    bias_gelu_bwd_kernel<true><<<column_grid(total8, n8, 16), 256, sizeof(float) * N, s>>>((const bf16*)dh, (const bf16*)z, bias, (bf16*)dz, dbias, total8, n8);
  else
    bias_gelu_bwd_kernel<false><<<column_grid(total8, n8, 16), 256, 0, s>>>((const bf16*)dh, (const bf16*)z, bias, (bf16*)dz, nullptr, total8, n8);
  return (int)cudaGetLastError();
}

int b200at_colsum_bf16(const void* a, float* out, int64_t M, int64_t N, void* stream) {
  if (M <= 0) return 0;
  if (N % 8 || N > 8192) return (int)cudaErrorInvalidValue;
  const int64_t total8 = M * N / 8;
  const int n8 = (int)(N / 8);
  colsum_kernel<<<column_grid(total8, n8, 4), 256, sizeof(float) * N, (cudaStream_t)stream>>>((const bf16*)a, out, total8, n8);
  return (int)cudaGetLastError();
}

int b200at_scale_residual_fwd(const void* z, const float* bias, const float* gamma, const void* res, void* out,
                              int64_t M, int64_t N, void* stream) {
  if (M <= 0) return 0;
  if (N % 8) return (int)cudaErrorInvalidValue;
  const int64_t total8 = M * N / 8;
  scale_residual_kernel<<<flat_grid(total8), 256, 0, (cudaStream_t)stream>>>((const bf16*)z, bias, gamma, (const bf16*)res, (bf16*)out, total8, (int)(N / 8));
  return (int)cudaGetLastError();
}

int b200at_scale_bwd(const void* dout, const float* gamma, void* dz, int64_t M, int64_t N, void* stream) {
  if (M <= 0) return 0;
  if (N % 8) return (int)cudaErrorInvalidValue;
  const int64_t total8 = M * N / 8;
  scale_bwd_kernel<<<flat_grid(total8), 256, 0, (cudaStream_t)stream>>>((const bf16*)dout, gamma, (bf16*)dz, total8, (int)(N / 8));
  return (int)cudaGetLastError();
}

int b200at_add_bf16(const void* a, const void* b, void* c, int64_t total, void* stream) {
  if (total <= 0) return 0;
  if (total % 8) return (int)cudaErrorInvalidValue;
  add_kernel<<<flat_grid(total / 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b, (bf16*)c, total / 8);
  return (int)cudaGetLastError();
}

int b200at_prepare_mlp_weights(const float* w1, const float* w2, const float* b2, const float* gamma, void* w1b, void* w1t,
                               void* w2g, void* w2gt, float* b2g, int64_t C, void* stream) {
  if (C <= 0 || C % 32) return (int)cudaErrorInvalidValue;
  const int tiles = 2 * (int)((4 * C / 32) * (C / 32));
  prepare_mlp_weights_kernel<<<tiles, 256, 0, (cudaStream_t)stream>>>(w1, w2, b2, gamma, (bf16*)w1b, (bf16*)w1t, (bf16*)w2g,
                                                                      (bf16*)w2gt, b2g, (int)C);
  return (int)cudaGetLastError();
}

int b200at_finish_mlp_grads(const float* dw2g, const float* w2, const float* col, const float* b2, const float* gamma,
                            float* dw2, float* db2, float* dgamma, int64_t C, void* stream) {
  if (C <= 0 || C % 4) return (int)cudaErrorInvalidValue;
  finish_mlp_grads_kernel<<<(unsigned)((C + 7) / 8), 256, 0, (cudaStream_t)stream>>>(dw2g, w2, col, b2, gamma, dw2, db2, dgamma, (int)C);
  return (int)cudaGetLastError();
}

int b200at_dwconv7_fwd(const void* x, const float* wt, const float* bias, const void* add, void* y, int64_t B,
                       int64_t H, int64_t W, int64_t C, void* stream) {
  if (B <= 0) return 0;
  // tensor-core (banded Toeplitz) kernel for every shape it covers: b200at_dwconv_mma.cu; B200AT_DW_MMA=0 keeps the
  // fp32-FMA kernel below (A/B measurements, and shapes wider than 80 columns)
  static const bool use_mma = [] { const char* e = getenv("B200AT_DW_MMA"); return e == nullptr || e[0] != '0'; }();
  if (use_mma) {
    const int r = b200at_dwconv7_mma_launch(x, wt, bias, add, y, B, H, W, C, stream);
    if (r >= 0) return r;
  }
  if (C % kDwCh || (reinterpret_cast<uintptr_t>(x) & 15)) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  // tile shapes: wide maps 14 x 28; 14-wide maps two images side by side; the 7 x 7 maps of the last stage three
  // images of 7 x 8 (one masked column)
  static const bool half_height = [] { const char* e = getenv("B200AT_DW_TH7"); return e == nullptr || e[0] != '0'; }();
  if (W > 14) return half_height ? launch_dwconv<7, 28, 1>(x, wt, bias, add, y, B, H, W, C, s)
                                 : launch_dwconv<14, 28, 1>(x, wt, bias, add, y, B, H, W, C, s);
  // 14-wide maps: the half-height tile (3 CTAs / SM) there too: 41.8 / 40.6 vs 43.8 / 42.7 us at 14 x 14 x 384, batch 128
  // (profiles/r02_ops_bench_dwconv_half14.txt); B200AT_DW_TH7_14=0 keeps the full-height tile
  static const bool half14 = [] { const char* e = getenv("B200AT_DW_TH7_14"); return e == nullptr || e[0] != '0'; }();
  if ((W > 8 || H > 7) && half14 && H > 7) return launch_dwconv<7, 14, 2>(x, wt, bias, add, y, B, H, W, C, s);
  if (W > 8 || H > 7) return launch_dwconv<14, 14, 2>(x, wt, bias, add, y, B, H, W, C, s);
  return launch_dwconv<7, 8, 3>(x, wt, bias, add, y, B, H, W, C, s);
}

int b200at_dwconv7_wgrad(const void* x, const void* dy, float* dw, float* db, int64_t B, int64_t H, int64_t W,
                         int64_t C, void* stream) {
  if (B <= 0) return 0;
  if (C % kDwCh) return (int)cudaErrorInvalidValue;
  const int tw = (int)((W + kDwTile - 1) / kDwTile), th = (int)((H + kDwTile - 1) / kDwTile);
  const int64_t spatial = B * th * tw;
  if (spatial > 0x7fffffff) return (int)cudaErrorInvalidValue;
  const int cgroups = (int)(C / kDwCh);
  if (H <= 8 && W <= 8) {                                 // small maps: one CTA per (image, channel group)
    const int64_t grid = spatial * cgroups;
    if (grid > 0x7fffffff) return (int)cudaErrorInvalidValue;
    dwconv7_wgrad_tile_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)dy, dw, db, (int)H, (int)W, (int)C, tw, th);
    return (int)cudaGetLastError();
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t splits = sms / cgroups;                         // one persistent CTA per SM (98 accumulator registers per thread)
  if (splits < 1) splits = 1;
  if (splits > spatial) splits = spatial;
  const size_t smem = 2 * (size_t)kWgBuf + sizeof(float2) * 7 * 16 * (kDwCh / 2);
  static b200at::SmemConfig configured;
  cudaError_t e = b200at::ensure_dynamic_smem(dwconv7_wgrad_kernel, (int)smem, configured);
  if (e != cudaSuccess) return (int)e;
  dwconv7_wgrad_kernel<<<(unsigned)(splits * cgroups), 256, smem, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)dy, dw, db, (int)H, (int)W, (int)C, tw, th, (int)spatial, (int)splits);
  return (int)cudaGetLastError();
}

}  // extern "C"
